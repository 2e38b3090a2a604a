"""Kernel-level table of ONE C4 (or C5) training step under torch.profiler -- diagnosis only, never a bench number."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SDB_SYNTHETIC_WEIGHTS", "1"); os.environ.setdefault("SDB_NO_TRIAL_DIRS", "1")
import torch
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
job = bench.Job(wl, 0, 1, dev, tempfile.mkdtemp())
for _ in range(3):
    job.step(job.to_device(job.host_batch()))
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    job.step(job.to_device(job.host_batch()))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
