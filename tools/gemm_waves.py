"""Back-to-back timing of one GEMM shape family around the wave boundaries (tiles = 148, 296, 320, 444, 592 ...):
separates main-loop efficiency from wave quantisation. Diagnostic only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scaledreamer_b200 import nn_ops as O

dev = torch.device("cuda:0")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2880
for N in (160, 320):
    for mt in (148, 296, 320, 444, 592, 640):
        M = 128 * mt // (N // 160)
        a = torch.randn(M, K, device=dev, dtype=torch.float16)
        b = torch.randn(N, K, device=dev, dtype=torch.float16)
        bias = torch.randn(N, device=dev, dtype=torch.float16)
        out = torch.empty(M, N, device=dev, dtype=torch.float16)
        for _ in range(3):
            O.gemm(a, b, bias=bias, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            O.gemm(a, b, bias=bias, out=out)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        tiles = (M // 128) * (N // 160)
        print(f"K{K} M{M} N{N}: tiles {tiles} ({tiles/148:.2f}/SM) {us:.1f} us  {2*M*N*K/us/1e6:.0f} TFLOP/s  per-SM-tile {us/max(1,-(-tiles//148)):.1f} us")
