"""Main-loop vs epilogue time of ONE tile per CTA (per-CTA %globaltimer stamps: set-up done / accumulator ready / exit) for
each tile width, on launches of exactly 148 tiles (one per SM) and 296 tiles (two per SM). Diagnostic only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scaledreamer_b200 import lib as L, nn_ops as O

lib = L.load()
dev = torch.device("cuda:0")
buf = torch.zeros(296 * 4, dtype=torch.int64, device=dev)
CASES = []
for K in (320, 1280, 2880):
    for N in (128, 160, 256):
        for mt in (148, 296):
            CASES.append((128 * mt, N, K))
for (M, N, K) in CASES:
    a = torch.randn(M, K, device=dev, dtype=torch.float16) * 0.1
    b = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.1
    bias = torch.randn(N, device=dev, dtype=torch.float16)
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    for _ in range(3):
        O.gemm(a, b, bias=bias, out=out)
    torch.cuda.synchronize()
    buf.zero_()
    lib.sdb_gemm_debug_timeline(buf.data_ptr())
    O.gemm(a, b, bias=bias, out=out)
    torch.cuda.synchronize()
    lib.sdb_gemm_debug_timeline(None)
    t = buf.view(296, 4).cpu()
    t = t[t[:, 0] > 0].double()
    t0 = t[:, 0].min()
    nkb = K // 64
    main = (t[:, 2] - t[:, 1]) / 1e3
    epi = (t[:, 3] - t[:, 2]) / 1e3
    print(f"M{M} N{N} K{K}: ctas {t.shape[0]} tiles {M//128}; setup {float((t[:,1]-t[:,0]).mean())/1e3:.2f} us; "
          f"main loop {float(main.mean()):.2f} us ({float(main.mean())/nkb*1e3:.0f} ns per k-block; tensor-pipe ideal {N*1.02:.0f}); "
          f"epilogue {float(epi.mean()):.2f} us (min {float(epi.min()):.2f}); span {float(t[:,3].max()-t0)/1e3:.1f} us", flush=True)
