"""Implicit-conv A operand vs a plain matrix of the same M x N x K: per-CTA stamps (entry / first accumulator / exit)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scaledreamer_b200 import lib as L, nn_ops as O

lib = L.load()
dev = torch.device("cuda:0")
buf = torch.zeros(296 * 4, dtype=torch.int64, device=dev)

def stamp(run, label):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    buf.zero_()
    lib.sdb_gemm_debug_timeline(buf.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.sdb_gemm_debug_timeline(None)
    t = buf.view(296, 4).cpu()
    t = t[t[:, 0] > 0].double()
    t0 = t[:, 0].min()
    q = lambda x: [round(float(v), 1) for v in torch.quantile(x, torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64))]
    print(f"{label}: back-to-back {e0.elapsed_time(e1)/10*1e3:.1f} us; ctas {t.shape[0]}; first accum at {q((t[:,2]-t0)/1e3)}; exit at {q((t[:,3]-t0)/1e3)} us", flush=True)

for (n, h, w, cin, cout) in [(1, 64, 64, 512, 512), (5, 64, 64, 320, 320), (5, 32, 32, 640, 640), (5, 16, 16, 1280, 1280), (5, 8, 8, 1280, 1280), (1, 512, 512, 128, 128)]:
    x = torch.randn(n, h, w, cin, device=dev, dtype=torch.float16) * 0.1
    wt = torch.randn(cout, 3, 3, cin, device=dev, dtype=torch.float16) * 0.02
    M, K = n * h * w, 9 * cin
    a = torch.randn(M, K, device=dev, dtype=torch.float16) * 0.1
    b = wt.reshape(cout, K)
    out = torch.empty(M, cout, device=dev, dtype=torch.float16)
    stamp(lambda: O.conv3x3(x, wt), f"conv {n}x{h}x{w}x{cin}->{cout} (M{M} N{cout} K{K})")
    stamp(lambda: O.gemm(a, b, out=out), f"  mm same shape")
    stamp(lambda: torch.matmul(a, b.t()), f"  cublas") if False else None
