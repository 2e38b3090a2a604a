"""C2 with lambda_orient forced on: device time of the orientation entry points per step (call timer), both kernels."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SDB_SYNTHETIC_WEIGHTS", "1")
os.environ.setdefault("SDB_NO_TRIAL_DIRS", "1")
import torch
import bench
from scaledreamer_b200 import lib as L

bench.WORKLOADS["C2"]["cli"] = bench.WORKLOADS["C2"]["cli"] + ["system.loss.lambda_orient=100.0"]
torch.cuda.set_device(0)
job = bench.Job("C2", 0, 1, torch.device("cuda:0"), tempfile.mkdtemp())
for _ in range(4):
    job.step(job.to_device(job.host_batch()))
torch.cuda.synchronize()
L.call_timer_begin()
n = 5
for _ in range(n):
    job.step(job.to_device(job.host_batch()))
rec = L.call_timer_end()
for k in ("sdb_render_orient_forward", "sdb_render_orient_backward", "sdb_render_nerf_backward_tape_zv", "sdb_render_nerf_backward_tape"):
    if k in rec:
        print(f"{k:36s} {rec[k]['ms'] / n:7.2f} ms per step")
