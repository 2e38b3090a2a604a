"""Times the C5 Triplane-Transformer generator (12 blocks, 768 channels, 16 heads, 3072 tokens, 4 prompts) forward +
backward: the native path (tcgen05 tf32 GEMMs + fp32 kernels) against the same network in torch (fp32 and tf32 cuBLAS +
SDPA). CUDA events, 3 warm-ups, 5 timed repeats. Also prints the per-entry-point device time of one native step."""
import sys
import time

import torch

sys.path.insert(0, ".")
from scaledreamer_b200 import lib as L  # noqa: E402
from scaledreamer_b200.amortized import TriplaneTransformer  # noqa: E402

cfg = {"inner_dim": 768, "condition_dim": 1024, "triplane_low_res": 32, "triplane_high_res": 64, "triplane_dim": 32,
       "num_layers": 12, "num_heads": 16, "flash_attention": False, "local_text": True}
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda")
torch.manual_seed(0)
gen = TriplaneTransformer(**cfg).to(dev)
emb = torch.randn(N, 77, 1024, device=dev)
d = torch.randn(N, 3, 64, 64, 32, device=dev).permute(0, 1, 4, 2, 3)


def run(fn, reps=5, warm=3):
    for _ in range(warm):
        for p in gen.parameters():
            p.grad = None
        fn(emb).backward(d)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        for p in gen.parameters():
            p.grad = None
        fn(emb).backward(d)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# flops of one forward: per block q/o (x2 attn) + k,v of self + mlp + scores
L_, C, H = 3072, 768, 16
lin = lambda m, n, k: 2.0 * m * n * k
fwd = 12 * (N * (lin(L_, C, C) * 4 + lin(L_, C, C) * 2 + lin(77, C, 1024) * 2 + lin(L_, 4 * C, C) * 2)
            + N * H * (lin(L_, L_, 48) * 2 + lin(L_, 77, 48) * 2))
print(f"prompts {N}: forward {fwd / 1e12:.2f} TFLOP, forward+backward ~{3 * fwd / 1e12:.2f} TFLOP (recomputed scores not counted)")
ms = run(gen.forward)
print(f"native  tcgen05 tf32            : {ms:8.2f} ms / step   {3 * fwd / ms / 1e9:7.1f} TFLOP/s useful")
torch.backends.cuda.matmul.allow_tf32 = True
ms_t = run(gen.forward_torch)
print(f"torch   cuBLAS tf32 + SDPA      : {ms_t:8.2f} ms / step   {3 * fwd / ms_t / 1e9:7.1f} TFLOP/s useful")
torch.backends.cuda.matmul.allow_tf32 = False
ms_f = run(gen.forward_torch, reps=2, warm=1)
print(f"torch   cuBLAS fp32 + SDPA (the reference's `precision: 32` default): {ms_f:8.2f} ms / step")
for p in gen.parameters():
    p.grad = None
L.call_timer_begin()
gen(emb).backward(d)
rec = L.call_timer_end()
tot = sum(v["ms"] for v in rec.values())
print(f"one native step by entry point ({tot:.2f} ms on the device, {sum(v['launches'] for v in rec.values())} launches):")
for k, v in sorted(rec.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:36s} {v['calls']:5d} calls {v['ms']:8.2f} ms")
print(f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")

# ---- per-shape GEMM times of one native step (CUDA events around every sdb_gemm_tf32 call)
from collections import defaultdict  # noqa: E402

from scaledreamer_b200 import transformer_ops as TO  # noqa: E402

_orig = TO.gemm
_recs = []


def _timed(A, B, M, N, K, out, batch=1, zdiv=1, **kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _orig(A, B, M, N, K, out, batch=batch, zdiv=zdiv, **kw)
    b.record()
    _recs.append(((M, N, K, batch), a, b))


TO.gemm = _timed
import scaledreamer_b200.triplane_native as TN  # noqa: E402

TN.T.gemm = _timed
for p in gen.parameters():
    p.grad = None
gen(emb).backward(d)
torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for shp, a, b in _recs:
    agg[shp][0] += 1
    agg[shp][1] += a.elapsed_time(b)
print("GEMM shapes of one step (M, N, K, batch): calls, total ms, TFLOP/s, output GB/s")
for shp, (n, ms_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M_, N_, K_, Bz = shp
    fl = 2.0 * M_ * N_ * K_ * Bz * n
    print(f"  {str(shp):28s} {n:4d} {ms_:8.2f} ms {fl / ms_ / 1e9:8.1f} TF/s {4.0 * M_ * N_ * Bz * n / ms_ / 1e6:8.0f} GB/s")
