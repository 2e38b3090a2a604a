"""One GEMM shape for an ncu capture: python tools/ncu_gemm_one.py M N K [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scaledreamer_b200 import nn_ops as O
M, N, K = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev, dtype=torch.float16) * 0.1
b = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.1
bias = torch.randn(N, device=dev, dtype=torch.float16)
out = torch.empty(M, N, device=dev, dtype=torch.float16)
for _ in range(reps):
    O.gemm(a, b, bias=bias, out=out)
torch.cuda.synchronize()
