"""Per-CTA %globaltimer stamps of one GEMM launch (entry / set-up done / first accumulator / exit): where does a small
GEMM spend its time? Diagnostic only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scaledreamer_b200 import lib as L, nn_ops as O

lib = L.load()
dev = torch.device("cuda:0")
buf = torch.zeros(296 * 4, dtype=torch.int64, device=dev)
for (M, N, K, res) in [(20480, 320, 320, True), (1280, 1280, 1280, True), (5120, 640, 640, True), (20480, 2560, 320, False),
                       (20480, 320, 2880, False)]:
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    b = torch.randn(N, K, device=dev, dtype=torch.float16)
    bias = torch.randn(N, device=dev, dtype=torch.float16)
    r = torch.randn(M, N, device=dev, dtype=torch.float16) if res else None
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    for _ in range(3):
        O.gemm(a, b, bias=bias, residual=r, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        O.gemm(a, b, bias=bias, residual=r, out=out)
    e1.record(); torch.cuda.synchronize()
    buf.zero_()
    lib.sdb_gemm_debug_timeline(buf.data_ptr())
    O.gemm(a, b, bias=bias, residual=r, out=out)
    torch.cuda.synchronize()
    lib.sdb_gemm_debug_timeline(None)
    t = buf.view(296, 4).cpu()
    t = t[t[:, 0] > 0].double()
    t0 = t[:, 0].min()
    print(f"M{M} N{N} K{K} res={res}: back-to-back {e0.elapsed_time(e1)/20*1e3:.1f} us/launch, ctas {t.shape[0]}; "
          f"entry spread {float(t[:,0].max()-t0)/1e3:.1f} us; setup {float((t[:,1]-t[:,0]).mean())/1e3:.2f} us; "
          f"to first accum {float((t[:,2]-t[:,1]).mean())/1e3:.2f} us; epilogue+rest {float((t[:,3]-t[:,2]).mean())/1e3:.2f} us; "
          f"kernel span {float(t[:,3].max()-t0)/1e3:.1f} us")
