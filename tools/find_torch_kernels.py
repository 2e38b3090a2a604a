"""Which torch (at::) kernels run inside one C2 training step, and from where: torch.profiler with stacks, one step after
warm-up. Prints the aten ops that launched device kernels with their Python call sites."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SDB_SYNTHETIC_WEIGHTS", "1")
os.environ.setdefault("SDB_NO_TRIAL_DIRS", "1")
import torch
from torch.profiler import ProfilerActivity, profile
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
job = bench.Job(wl, 0, 1, torch.device("cuda:0"), tempfile.mkdtemp())
for _ in range(4):
    job.step(job.to_device(job.host_batch()))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    job.step(job.to_device(job.host_batch()))
    torch.cuda.synchronize()
tot = 0.0
out = []
for k in prof.key_averages(group_by_stack_n=12):
    t = getattr(k, "self_device_time_total", None)
    if t is None:
        t = getattr(k, "self_cuda_time_total", 0.0)
    if not k.key.startswith("aten::") or t <= 0:
        continue
    st = [s for s in (k.stack or []) if "scaledreamer_b200" in s or "bench.py" in s]
    out.append((t, k.count, k.key, st[0] if st else (k.stack[0] if k.stack else "?")))
    tot += t
print(f"aten ops with device time of their own: {tot / 1e3:.3f} ms per step")
for t, c, n, s in sorted(out, key=lambda r: -r[0])[:32]:
    print(f"{t:8.1f} us  x{c:3d}  {n:26s} {s[-120:]}")
