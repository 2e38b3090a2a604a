"""profiles/<tag>_ncu_full_summary.csv from the raw pages of the `ncu --set full` captures (gpurun_out/<tag>_full_*_raw.csv,
written by tools/capture_evidence.sh): one row per captured launch with the metrics DESIGN.md argues from. Also copies the
raw pages next to it. Usage: python tools/summarize_full.py r2"""
import csv
import glob
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "dram__bytes_read.sum",
           "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__grid_size", "launch__block_size", "smsp__inst_executed_op_global_red.sum",
           "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^(void )?((dense::)?(<unnamed>|\(anonymous namespace\))::)*", "", name)
    return re.sub(r"^(void )?unnamed>::", "", name)[:80]  # ncu prints `dense::<unnamed>::f<...>` as `unnamed>::f<...>`


out = [["capture", "kernel"] + METRICS]
for path in sorted(glob.glob(os.path.join(SRC, f"{TAG}_full_*_raw.csv"))):
    cap = os.path.basename(path)[len(TAG) + 6:-8]
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        line = [cap, short(r[ki])]
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                line.append(f"{r[i]} {units[i]}".strip())
            else:
                line.append("")
        out.append(line)
    shutil.copy(path, os.path.join(DST, os.path.basename(path)))
with open(os.path.join(DST, f"{TAG}_ncu_full_summary.csv"), "w", newline="") as f:
    csv.writer(f).writerows(out)
print(len(out) - 1, "captured launches summarised")
