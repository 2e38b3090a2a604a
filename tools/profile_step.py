"""Runs W warm-up training steps of a BASELINE workload, then brackets ONE step with cudaProfilerStart/Stop (use with
`ncu --profile-from-start off ...`). Never a benchmark: numbers under a profiler are not bench values.
    python tools/profile_step.py [W] [C2|C3|C4|C5] [orient]      (`orient`: C2/C3 with lambda_orient forced to 100)"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SDB_SYNTHETIC_WEIGHTS", "1")
os.environ.setdefault("SDB_NO_TRIAL_DIRS", "1")
import torch
import bench

W = int(sys.argv[1]) if len(sys.argv) > 1 else 3
wl = sys.argv[2] if len(sys.argv) > 2 else "C2"
if len(sys.argv) > 3 and sys.argv[3] == "orient":
    bench.WORKLOADS[wl]["cli"] = bench.WORKLOADS[wl]["cli"] + ["system.loss.lambda_orient=100.0"]
torch.cuda.set_device(0)
job = bench.Job(wl, 0, 1, torch.device("cuda:0"), tempfile.mkdtemp())
for _ in range(W):
    job.step(job.to_device(job.host_batch()))
torch.cuda.synchronize()
torch.cuda.profiler.start()
job.step(job.to_device(job.host_batch()))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
