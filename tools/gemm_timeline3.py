"""Per-CTA %globaltimer stamps of stream-K launches: when do CTAs produce their first accumulator and when do they exit?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scaledreamer_b200 import lib as L, nn_ops as O

lib = L.load()
dev = torch.device("cuda:0")
buf = torch.zeros(296 * 4, dtype=torch.int64, device=dev)
for (M, N, K) in [(20480, 320, 1280), (4096, 512, 4608), (1280, 1280, 5120), (5120, 640, 2560)]:
    a = torch.randn(M, K, device=dev, dtype=torch.float16) * 0.1
    b = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.1
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    for _ in range(3):
        O.gemm(a, b, out=out)
    torch.cuda.synchronize()
    buf.zero_()
    lib.sdb_gemm_debug_timeline(buf.data_ptr())
    O.gemm(a, b, out=out)
    torch.cuda.synchronize()
    lib.sdb_gemm_debug_timeline(None)
    t = buf.view(296, 4).cpu()
    t = t[t[:, 0] > 0].double()
    t0 = t[:, 0].min()
    q = lambda x: [round(float(v), 1) for v in torch.quantile(x, torch.tensor([0.0, 0.25, 0.5, 0.75, 1.0], dtype=torch.float64))]
    print(f"M{M} N{N} K{K}: ctas {t.shape[0]}; entry {q((t[:,0]-t0)/1e3)}; first accum at {q((t[:,2]-t0)/1e3)}; exit at {q((t[:,3]-t0)/1e3)} us", flush=True)
