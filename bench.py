"""ASD-step benchmark (BASELINE.json metric): steps/s of the full Asynchronous-Score-Distillation training step --
random camera -> 256x256 fused NeRF render -> 512x512 VAE encode -> SD-2.1-shape UNet on the 5-way Perp-Neg batch at
t and t+dt -> score gradient -> VAE-encoder / render backward -> AdamW -- on N B200s (one process per GPU, weak
scaling, one NCCL all-reduce of generator gradients per step).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference math on the host cores (oracle port)

One JSON line on stdout (rank 0). Weights are seeded synthetic (no checkpoints on the box); data is synthetic
random cameras. See DESIGN.md "Measurement" for how every field is produced.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# no checkpoints on the box: the benchmark runs on seeded synthetic weights and says so in its JSON line
os.environ.setdefault("SDB_SYNTHETIC_WEIGHTS", "1")

METRIC = "ASD steps/sec (256^2 render->UNet)"
WORKLOAD = "C2: single-prompt ASD-SD, hash-grid iNGP NeRF, 256x256x1 view, Perp-Neg UNet batch 5 @64x64 latents, VAE @512x512"
CFG_YAML = os.path.join(ROOT, "tests", "configs", "asd_sd_nerf.yaml")
H = W = 256


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel families the roofline objects
    describe, from the newest committed ncu capture (profiles/*_traffic.json, written by tools/summarize_profiles.py from
    one profiled step of the same workload). Empty when no capture is committed: traffic is then null."""
    import glob

    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "*_traffic.json")))
    if not files:
        return {}
    t = json.load(open(files[-1]))
    out = {}
    fam = [t[k] for k in ("gemm", "flash_attn") if k in t]
    if fam:
        n = sum(f["launches_per_step"] for f in fam)
        out["tensor"] = sum(f["dram_bytes_per_launch"] * f["launches_per_step"] for f in fam) / n
    if "render_fwd" in t:
        out["render_fwd"] = t["render_fwd"]["dram_bytes_per_launch"]
    if "render_field_bwd" in t:
        out["render_bwd"] = t["render_field_bwd"]["dram_bytes_per_launch"] + t.get("render_composite_bwd", {}).get(
            "dram_bytes_per_launch", 0.0)
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.gpu), "-lms", "200"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU baseline
def cpu_baseline_step_seconds():
    """The reference's math on the host cores (oracle port, fp32 PyTorch CPU), on a BOUNDED sample of the C2 step,
    scaled linearly to the full step: render 32x32 rays of the 256x256 view (x64), VAE encoder fwd+bwd on a
    128x128 crop of the 512x512 input (x16), UNet forward of 1 of the 5 batch entries at 16x16 of the 64x64 latents
    (x80)."""
    import torch
    import torch.nn.functional as F

    from oracle import ldm_oracle as lo, render_oracle as ro
    from scaledreamer_b200 import nets
    from tests.helpers import scene

    torch.set_num_threads(os.cpu_count() or 1)
    t_parts = {}
    sc = scene(H=32, W=32, B=1, seed=0, table_scale=1e-4)
    P = {k: v.clone().requires_grad_(True) for k, v in sc["P"].items()}
    t0 = time.perf_counter()
    out = ro.render(sc["rays_o"], sc["rays_d"], sc["jitter"], None, sc["binary"].numpy(), float(sc["occs"].mean()), P,
                    sc["fcfg"], sc["mcfg"], 1024)
    out["comp_rgb"].square().sum().backward()
    t_parts["render_32x32"] = time.perf_counter() - t0

    vae_specs = [(n, s) for n, s in _specs("vae", 1, 128, 128)]
    sd_v = nets.random_state_dict(vae_specs, 1)
    x = (torch.rand(1, 3, 128, 128) * 2 - 1).requires_grad_(True)
    t0 = time.perf_counter()
    h = lo.vae_encoder_forward(sd_v, x)
    h.square().sum().backward()
    t_parts["vae_128_fwd_bwd"] = time.perf_counter() - t0
    del sd_v

    sd_u = nets.random_state_dict(_specs("unet", 1, 16, 16), 0)
    with torch.no_grad():
        xin, ctx = torch.randn(1, 4, 16, 16), torch.randn(1, 77, 1024)
        t0 = time.perf_counter()
        lo.unet_forward(sd_u, xin, torch.tensor([500.0]), ctx)
        t_parts["unet_1x16x16"] = time.perf_counter() - t0
    step_s = t_parts["render_32x32"] * 64 + t_parts["vae_128_fwd_bwd"] * 16 + t_parts["unet_1x16x16"] * 80
    sample = ("oracle port: render fwd+bwd 32x32 rays x64 + VAE enc fwd+bwd 128x128 x16 + UNet fwd 1x16x16 latents x80; "
              + ", ".join(f"{k}={v:.2f}s" for k, v in t_parts.items()))
    return step_s, sample, torch.get_num_threads()


def _specs(kind, B, Hh, Ww):
    import ctypes as C

    from scaledreamer_b200 import lib as L

    lib = L.load()
    h = C.c_void_p()
    if kind == "vae":
        c = L.VaeCfgC(3, 128, 4, (C.c_int * 4)(1, 2, 4, 4), 2, 4)
        L.check(lib.sdb_vae_encoder_create(C.byref(c), B, Hh, Ww, C.byref(h)), "create")
    else:
        c = L.UNetCfgC(4, 4, 320, 4, (C.c_int * 4)(1, 2, 4, 4), 2, 3, 64, 1024, 77, 0, 1)
        L.check(lib.sdb_unet_create(C.byref(c), B, Hh, Ww, C.byref(h)), "create")
    name, ndim, shape = C.c_char_p(), C.c_int(), (C.c_int * 4)()
    out = []
    for i in range(lib.sdb_net_num_params(h)):
        L.check(lib.sdb_net_param(h, i, C.byref(name), C.byref(ndim), shape), "param")
        out.append((name.value.decode(), tuple(shape[: ndim.value])))
    lib.sdb_net_destroy(h)
    return out


def run_reference(args, rank: int) -> None:
    if rank != 0:
        return
    times = []
    sample, cores = "", 1
    for i in range(args.warmup + args.steps):
        s, sample, cores = cpu_baseline_step_seconds()
        if i >= args.warmup:
            times.append(s)
    step_s = sum(times) / len(times)
    v = 1.0 / step_s
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference math (oracle port) on host cores, bounded sample scaled to one step"},
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------- CUDA arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="C2", choices=["C2", "C3"],
                    help="C2 = the BASELINE metric's configuration (default); C3 = ASD-MVDream, 4 views of 256x256 per step "
                         "(informational: same loop, not the headline line)")
    args = ap.parse_args()
    global WORKLOAD, CFG_YAML
    extra_cli = []
    if args.workload == "C3":
        WORKLOAD = ("C3: single-prompt ASD-MVDream, hash-grid iNGP NeRF, 256x256x4 views, multi-view UNet batch 12 @32x32 "
                    "latents (cond/uncond/t+dt x 4 views), VAE @256x256")
        CFG_YAML = os.path.join(ROOT, "tests", "configs", "asd_mv_nerf.yaml")
        extra_cli = ["data.batch_size=[4,4]"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import scaledreamer_b200 as sd
    from scaledreamer_b200 import lib as L
    from scaledreamer_b200.systems import FusedAdamW, Trainer

    import random

    torch.manual_seed(1234 + rank)
    random.seed(1234 + rank)
    cfg = sd.load_config(CFG_YAML, cli_args=["system.prompt_processor.prompt=a DSLR photo of a hamburger",
                                             f"data.width=[{W},{W}]", f"data.height=[{H},{H}]"] + extra_cli)
    dm = sd.find(cfg.data_type)(cfg.data)
    system = sd.find(cfg.system_type)(cfg.system)
    dm.setup("fit")
    ds = dm.train_dataset
    system.train()
    system.on_fit_start()
    opt = system.configure_optimizers()
    if isinstance(opt, FusedAdamW):
        opt.grad_scale = 1.0 / world
    params = [p for g in opt.param_groups for p in g["params"]]
    trainer = Trainer(max_steps=0, distributed=world > 1)
    step_no = [0]

    def step(batch_dev):
        ds.update_step(0, step_no[0])
        system.true_global_step = step_no[0]
        system.do_update_step(0, step_no[0])
        out = system.training_step(batch_dev, step_no[0])
        out["loss"].backward()
        if world > 1:
            trainer._allreduce_grads(params)
        opt.step()
        opt.zero_grad(set_to_none=False)
        system.do_update_step_end(0, step_no[0])
        step_no[0] += 1
        return out["loss"]

    def host_batch():
        # host tensors; ds.to_device packs them into a pinned staging slot and uploads them with one asynchronous copy
        return ds.collate({})

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (builds the networks, the occupancy grid, every TMA descriptor)
    for _ in range(args.warmup):
        step(ds.to_device(host_batch(), dev))
    sync_all()

    # ---- timed region 1: inputs resident in HBM
    resident = [ds.to_device(host_batch(), dev) for _ in range(args.steps)]
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in resident:
        step(b)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - n0

    # ---- timed region 2: end to end through the plugin API with host buffers (pinned H2D in, loss D2H out)
    h2d = d2h = 0
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    # The loss of step i is copied to pinned host memory asynchronously and read on the host one step later (after
    # step i+1 has been enqueued), the way a logging trainer consumes it: every step still pays its H2D inputs and a
    # D2H result, but the host never drains the GPU queue between steps.
    loss_pinned = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_events = [torch.cuda.Event(), torch.cuda.Event()]
    losses_host = []
    for i in range(args.steps):
        hb = host_batch()
        h2d = sum(v.numel() * v.element_size() for v in hb.values() if torch.is_tensor(v))
        loss = step(ds.to_device(hb, dev))
        loss_pinned[i & 1].copy_(loss.detach(), non_blocking=True)
        loss_events[i & 1].record()
        if i > 0:
            loss_events[(i - 1) & 1].synchronize()
            losses_host.append(float(loss_pinned[(i - 1) & 1]))
        d2h = 4
    loss_events[(args.steps - 1) & 1].synchronize()
    losses_host.append(float(loss_pinned[(args.steps - 1) & 1]))
    assert len(losses_host) == args.steps and all(v == v for v in losses_host)
    e3.record()
    sync_all()
    ms_e2e = e2.elapsed_time(e3)
    clk = clocks.stop() if rank == 0 else {}

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- profiling pass (same workload, outside the timed regions): per-phase and per-kernel-class device times
    import ctypes as C

    lib = L.load()
    prof = {}
    if rank == 0:
        ev = lambda: torch.cuda.Event(enable_timing=True)
        acc = {"render_fwd": 0.0, "guidance_fwd": 0.0, "backward": 0.0, "optimizer": 0.0}
        n_prof = 3
        lib.sdb_gemm_profile_begin()
        if os.environ.get("SDB_GEMM_CSV"):
            lib.sdb_gemm_profile_dump(os.environ["SDB_GEMM_CSV"].encode())
        samples_kept = 0
        fwd_each = []
        for _ in range(n_prof):
            b = ds.to_device(host_batch(), dev)
            a0, a1, a2, a3, a4 = ev(), ev(), ev(), ev(), ev()
            a0.record()
            out = system(b)
            a1.record()
            g = system.guidance(out["comp_rgb"], system.prompt_utils, **b, rgb_as_latents=False)
            lossp = g["loss_asd"] + 30.0 * (out["opacity"] ** 2 + 0.01).sqrt().mean()
            a2.record()
            lossp.backward()
            a3.record()
            opt.step()
            opt.zero_grad(set_to_none=False)
            a4.record()
            torch.cuda.synchronize()
            fwd_each.append(a0.elapsed_time(a1))
            for k, (x, y) in zip(acc, ((a0, a1), (a1, a2), (a2, a3), (a3, a4))):
                acc[k] += x.elapsed_time(y) / n_prof
        gm, gf, gl = C.c_double(), C.c_double(), C.c_int()
        L.check(lib.sdb_gemm_profile_end(C.byref(gm), C.byref(gf), C.byref(gl)), "profile_end")
        # render kernels alone (sample count for the algorithmic bytes)
        from scaledreamer_b200 import render_ops as R

        rr = system.renderer
        b = ds.to_device(host_batch(), dev)
        P = {k: v.detach() for k, v in rr._params().items()}
        march = R.MarchSpec(render_step_size=rr.render_step_size, prune=True, grid_res=32)
        ro_, rd_ = b["rays_o"].reshape(-1, 3).contiguous(), b["rays_d"].reshape(-1, 3).contiguous()
        n_rays = ro_.shape[0]  # views x H x W
        jit = torch.rand(n_rays, device=dev)
        tape = R.RenderTape.acquire(march, rr._spec().radius, n_rays, dev)
        o = R.render_forward_v2_raw(rr._spec(), march, P, rr._occ_grid(dev), ro_, rd_, jit, None, H * W, tape)
        samples_kept = int(tape.counter[0].item())
        tape.check_overflow()
        grads = {k: torch.zeros_like(v) for k, v in P.items()}
        g_rgb = torch.randn_like(o["comp_rgb"])
        t_f, t_b = [], []
        for _ in range(3):  # best of three: the first call after a large allocation is not representative
            f0, f1, f2 = ev(), ev(), ev()
            f0.record()
            o = R.render_forward_v2_raw(rr._spec(), march, P, rr._occ_grid(dev), ro_, rd_, jit, None, H * W, tape)
            f1.record()
            R.render_backward_tape_raw(rr._spec(), march, P, grads, rd_, None, H * W, o, tape, g_rgb)
            f2.record()
            torch.cuda.synchronize()
            t_f.append(f0.elapsed_time(f1))
            t_b.append(f1.elapsed_time(f2))
        tape.release()
        torch.cuda.synchronize()
        prof = dict(phase_ms=acc, gemm_ms_per_step=gm.value / n_prof, gemm_tflop_per_step=gf.value / n_prof / 1e12,
                    gemm_launches_per_step=gl.value // n_prof, render_fwd_kernel_ms=min(t_f),
                    render_bwd_kernel_ms=min(t_b), render_samples_kept=samples_kept,
                    render_tapes_allocated=R.RenderTape.n_allocated, render_fwd_phase_ms_each=fwd_each)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    pk = peaks()
    tr = ncu_traffic()
    gemm_tfs = prof["gemm_tflop_per_step"] / (prof["gemm_ms_per_step"] / 1e3)
    # algorithmic bytes of the render kernels: samples x 16 levels x 8 corners x 2 features x 4 B (SURVEY.md 8d, E = 1)
    n_views = 4 if args.workload == "C3" else 1
    rbytes_f = prof["render_samples_kept"] * 1024 + n_views * H * W * (24 + 28)
    rbytes_b = 2 * prof["render_samples_kept"] * 1024 + n_views * H * W * (24 + 28)
    roof_gemm = {"kernel": "gemm_f16_kernel + flash_attn_f16_kernel (tcgen05 GEMM / implicit conv / fused attention, all launches of one step)", "bound": "tensor",
                 "achieved": gemm_tfs, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tfs / pk["tf_sustained"],
                 "traffic": tr.get("tensor"), "ms_per_step": prof["gemm_ms_per_step"], "peak_source": pk["src"] + " sustained bf16"}
    rb = rbytes_b / 1e9 / (prof["render_bwd_kernel_ms"] / 1e3)
    roof_rbwd = {"kernel": "render_composite_bwd_kernel + render_field_bwd_kernel (tape backward)", "bound": "hbm", "achieved": rb, "peak": pk["hbm"], "unit": "GB/s",
                 "frac": rb / pk["hbm"], "traffic": tr.get("render_bwd"), "ms_per_step": prof["render_bwd_kernel_ms"],
                 "peak_source": pk["src"] + " copy bandwidth"}
    rf = rbytes_f / 1e9 / (prof["render_fwd_kernel_ms"] / 1e3)
    roof_rfwd = {"kernel": "render_bg_kernel + render_nerf_fwd2_kernel (march + encode + MLPs + composite + tape)", "bound": "hbm", "achieved": rf, "peak": pk["hbm"], "unit": "GB/s",
                 "frac": rf / pk["hbm"], "traffic": tr.get("render_fwd"), "ms_per_step": prof["render_fwd_kernel_ms"],
                 "peak_source": pk["src"] + " copy bandwidth"}
    roofs = sorted([roof_gemm, roof_rbwd, roof_rfwd], key=lambda r: -r["ms_per_step"])
    line = {
        "metric": METRIC, "value": world * args.steps / (ms / 1e3), "unit": "steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (render: f32)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "weights": "seeded synthetic (865.9 M UNet, 34.2 M VAE encoder)",
                   "parallelism": f"dp{world}: one prompt+camera per GPU, one flat all-reduce of generator grads per step",
                   "l2": "no explicit flush: each step streams 1.73 GB of UNet weights + >1 GB activations (>> 126 MB L2)"},
        "e2e": {"value": world * args.steps / (ms_e2e / 1e3), "unit": "steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "gpu_launches_per_step": int(launches // args.steps),
        "clocks": clk, "roofline": roofs[0], "roofline_other_kernels": roofs[1:], "profile": prof,
    }
    line["views_per_s"] = line["value"] * n_views
    if world == 1 and not args.no_cpu_baseline:
        step_s, sample, cores = cpu_baseline_step_seconds()
        line["cpu_baseline"] = {"value": 1.0 / step_s, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
