"""Per-step device times of a bench workload (CUDA events around every step): shows whether a mean hides outliers.
usage: python tools/step_times.py C4 [steps]"""
import argparse
import os
import sys
import tempfile

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
os.environ.setdefault("SDB_SYNTHETIC_WEIGHTS", "1")
os.environ.setdefault("SDB_NO_TRIAL_DIRS", "1")
dev = torch.device("cuda:0")
job = bench.Job(name, 0, 1, dev, tempfile.mkdtemp())
for _ in range(4):
    job.step(job.to_device(job.host_batch()))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
mem0 = torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"]
ev[0].record()
for i in range(steps):
    job.step(job.to_device(job.host_batch()))
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
st = torch.cuda.memory_stats()
print(name, "per-step ms:", " ".join(f"{v:.1f}" for v in ms))
print("mean", round(sum(ms) / len(ms), 2), "min", round(min(ms), 2), "max", round(max(ms), 2), "alloc retries / device allocs during the run:",
      st["num_alloc_retries"] - mem0[0], st["num_device_alloc"] - mem0[1], "reserved GiB", round(st["reserved_bytes.all.peak"] / 2**30, 1))
