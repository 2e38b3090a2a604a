"""Turns the raw ncu CSVs of tools/capture_evidence.sh (gpurun_out/<tag>_*) into the summaries committed under
profiles/: per-kernel totals of the bench launch list, per-kernel duration + DRAM traffic of one step, and
profiles/<tag>_traffic.json (average DRAM bytes per launch of each kernel family, read by bench.py for
roofline.traffic). Usage: python tools/summarize_profiles.py r1f"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1f"
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

FAMILIES = [("gemm", "gemm_f16_kernel"), ("flash_attn", "flash_attn_f16_kernel"), ("render_fwd", "render_nerf_fwd2_kernel"),
            ("render_field_bwd", "render_field_bwd"), ("render_composite_bwd", "render_composite_bwd_kernel"),
            ("groupnorm", "gn_"), ("render_orient_fwd", "render_orient_fwd_kernel"),
            ("render_orient_bwd", "render_orient_bwd_kernel"), ("hyper_field_fwd", "hyper_field_fwd_kernel"),
            ("hyper_field_bwd", "hyper_field_bwd"), ("volsdf", "volsdf_"), ("gemm_tf32", "gemm_tf32_kernel"),
            ("softmax_f32", "softmax_f32_")]


# CUDA sources each group of kernel families is built from: a capture counts for a group only while these files are
# unchanged (editing the transformer kernels does not invalidate the render or tensor numbers)
SOURCE_GROUPS = {
    "tensor": ["gemm_sm100.cu", "flash_attn_sm100.cu", "ptx_sm100.cuh", "dense.h", "common.cuh"],
    "render": ["render_fwd2.cu", "render_bwd2.cu", "render_bwd_tc.cu", "field_bwd_tc.cuh", "render_tape.cuh",
               "render_types.cuh", "field.cuh", "common.cuh"],
    "hyper_field": ["hyper_field.cu", "field_bwd_tc.cuh", "render_tape.cuh", "render_types.cuh", "field.cuh", "common.cuh"],
}


def lib_sources_sha(group: str = None) -> str:
    """Hash of the CUDA sources a group of kernels is built from (all sources when group is None): bench.py only reports
    `roofline.traffic` from a capture whose hash equals the one of the sources it runs (a stale capture reads as null
    instead of a wrong number)."""
    import glob
    import hashlib

    base = os.path.join(ROOT, "scaledreamer_b200", "csrc")
    if group is None:
        files = sorted(glob.glob(os.path.join(base, "*.cu")) + glob.glob(os.path.join(base, "*.cuh")) +
                       glob.glob(os.path.join(base, "*.h")))
    else:
        files = [os.path.join(base, f) for f in SOURCE_GROUPS[group]]
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def rows_of(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(lines[start:]))


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^(void )?((dense::)?(<unnamed>|\(anonymous namespace\))::)*", "", name)
    return re.sub(r"^(void )?unnamed>::", "", name)[:80]  # ncu prints `dense::<unnamed>::f<...>` as `unnamed>::f<...>`


def launches(path):
    """id -> {name, ms, rd, wr}"""
    out = collections.OrderedDict()
    for r in rows_of(path):
        d = out.setdefault(r["ID"], {"name": short(r["Kernel Name"]), "ms": 0.0, "rd": 0.0, "wr": 0.0})
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["ms"] = to_ms(r["Metric Value"], r["Metric Unit"])
        elif m == "dram__bytes_read.sum":
            d["rd"] = to_bytes(r["Metric Value"], r["Metric Unit"])
        elif m == "dram__bytes_write.sum":
            d["wr"] = to_bytes(r["Metric Value"], r["Metric Unit"])
    return list(out.values())


def per_kernel(ls, path, traffic):
    agg = collections.OrderedDict()
    for d in ls:
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["ms"]
        a[2] += d["rd"]
        a[3] += d["wr"]
    total = sum(a[1] for a in agg.values())
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share", "avg_us"] + (["dram_read_MB", "dram_write_MB"] if traffic else []))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], f"{a[1]:.4f}", f"{a[1] / total:.4f}", f"{1e3 * a[1] / a[0]:.2f}"]
                       + ([f"{a[2] / 1e6:.2f}", f"{a[3] / 1e6:.2f}"] if traffic else []))
        w.writerow(["TOTAL", sum(a[0] for a in agg.values()), f"{total:.4f}", "1.0", ""])
    return agg, total


def main():
    p = os.path.join(SRC, f"{TAG}_launches_bench.csv")
    if os.path.exists(p):
        ls = launches(p)
        agg, total = per_kernel(ls, os.path.join(DST, f"{TAG}_launches_bench_per_kernel.csv"), False)
        print("bench launch list:", len(ls), "launches,", f"{total:.1f} ms under ncu")
    p = os.path.join(SRC, f"{TAG}_step_traffic.csv")
    if os.path.exists(p):
        ls = launches(p)
        agg, total = per_kernel(ls, os.path.join(DST, f"{TAG}_step_kernels_traffic.csv"), True)
        fam = {}
        for key, pat in FAMILIES:
            sel = [d for d in ls if pat in d["name"]]
            if sel:
                fam[key] = {"launches_per_step": len(sel), "ms_per_step_under_ncu": round(sum(d["ms"] for d in sel), 4),
                            "share_of_step": round(sum(d["ms"] for d in sel) / total, 4),
                            "dram_bytes_per_launch": round(sum(d["rd"] + d["wr"] for d in sel) / len(sel), 1)}
        for extra, label in (("c4", "C4"), ("orient", "C2 with lambda_orient = 100, render kernels only"),
                             ("generator", "Triplane-Transformer generator, 2 of 12 blocks, 4 prompts, forward + backward")):
            pe = os.path.join(SRC, f"{TAG}_{extra}_step_traffic.csv")
            if not os.path.exists(pe):
                continue
            le = launches(pe)
            _, tot_e = per_kernel(le, os.path.join(DST, f"{TAG}_{extra}_step_kernels_traffic.csv"), True)
            print(f"one {label} step: {len(le)} launches, {tot_e:.1f} ms under ncu")
            for key, pat in FAMILIES:
                sel = [d for d in le if pat in d["name"]]
                if sel and key not in fam:
                    fam[key] = {"launches_per_step": len(sel), "ms_per_step_under_ncu": round(sum(d["ms"] for d in sel), 4),
                                "share_of_step": round(sum(d["ms"] for d in sel) / tot_e, 4),
                                "dram_bytes_per_launch": round(sum(d["rd"] + d["wr"] for d in sel) / len(sel), 1),
                                "workload": label}
        fam["_source"] = f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one step of tools/profile_step.py ({TAG})"
        fam["_lib_sources_sha"] = lib_sources_sha()
        fam["_group_sources_sha"] = {g: lib_sources_sha(g) for g in SOURCE_GROUPS}
        json.dump(fam, open(os.path.join(DST, f"{TAG}_traffic.json"), "w"), indent=1)
        print("one step:", len(ls), "launches,", f"{total:.1f} ms under ncu")
        print(json.dumps(fam, indent=1))


if __name__ == "__main__":
    main()
