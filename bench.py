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
os.environ.setdefault("SDB_NO_TRIAL_DIRS", "1")

METRIC = "ASD steps/sec (256^2 render->UNet)"
H = W = 256
CFG_DIR = os.path.join(ROOT, "tests", "configs")
# BASELINE.json configs[1..4]. C2 is the configuration the metric is quoted on (default at every N, so that the driver's
# 1 -> 8 GPU efficiency compares like with like); C4 / C5 are the prompt-sharded configurations north_star names for
# 8 GPUs (`--workload C4`; with --gpus N > 1 the default C2 line also carries a measured "c4" block).
WORKLOADS = {
    "C2": dict(yaml="asd_sd_nerf.yaml", views=1, multiprompt=False, cli=[],
               desc="C2: single-prompt ASD-SD, hash-grid iNGP NeRF, 256x256x1 view, Perp-Neg UNet batch 5 @64x64 latents, "
                    "VAE @512x512"),
    "C3": dict(yaml="asd_mv_nerf.yaml", views=4, multiprompt=False, cli=["data.batch_size=[4,4]"],
               desc="C3: single-prompt ASD-MVDream, hash-grid iNGP NeRF, 256x256x4 views, multi-view UNet batch 12 @32x32 "
                    "latents (cond/uncond/t+dt x 4 views), VAE @256x256"),
    "C4": dict(yaml="asd_sd_hyper_iNGP.yaml", views=1, multiprompt=True, cli=["data.batch_size=1"],
               desc="C4: multi-prompt Hyper-iNGP (MG15-size library, rank-strided), ASD-SD, VolSDF importance renderer "
                    "256x256, one prompt per GPU, Perp-Neg UNet batch 5 @64x64 latents, VAE @512x512"),
    "C5": dict(yaml="asd_mv_triplane_transformer.yaml", views=4, multiprompt=True, cli=[],
               desc="C5: multi-prompt Triplane-Transformer, ASD-MVDream, VolSDF importance renderer 256x256x4 views, one "
                    "prompt per GPU, multi-view UNet batch 12 @32x32 latents, accumulate_grad_batches 2"),
}


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel families the roofline objects
    describe, from the newest committed ncu capture (profiles/*_traffic.json, written by tools/summarize_profiles.py from
    one profiled step of the same workload). Empty when no capture is committed: traffic is then null."""
    import glob

    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "*_traffic.json")),
                   key=os.path.getmtime)
    if not files:
        return {}
    t = json.load(open(files[-1]))
    # only a capture of THESE kernels counts: the file carries hashes of the CUDA sources each kernel group was built from
    # (tools/summarize_profiles.py); a group whose sources changed since reads as null rather than as a stale number
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        from summarize_profiles import SOURCE_GROUPS, lib_sources_sha
        shas = t.get("_group_sources_sha", {})
        ok = {g: shas.get(g) == lib_sources_sha(g) for g in SOURCE_GROUPS}
    except Exception:
        return {}
    out = {}
    fam = [t[k] for k in ("gemm", "flash_attn") if k in t]
    if fam and ok["tensor"]:
        n = sum(f["launches_per_step"] for f in fam)
        out["tensor"] = sum(f["dram_bytes_per_launch"] * f["launches_per_step"] for f in fam) / n
    if "render_fwd" in t and ok["render"]:
        out["render_fwd"] = t["render_fwd"]["dram_bytes_per_launch"]
    if "render_field_bwd" in t and ok["render"]:
        out["render_bwd"] = t["render_field_bwd"]["dram_bytes_per_launch"] + t.get("render_composite_bwd", {}).get(
            "dram_bytes_per_launch", 0.0)
    for k in ("hyper_field_fwd", "hyper_field_bwd"):
        if k in t and ok["hyper_field"]:
            out[k] = t[k]["dram_bytes_per_launch"]
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled every 200 ms during the timed region. NVML is read in-process from a thread
    (two driver queries per sample); the `nvidia-smi -lms` subprocess this replaces takes a driver-wide lock for tens of
    milliseconds per sample, which showed up as 65 ms host stalls in the end-to-end leg and as 10-15 ms per step on the
    host-paced C4 workload (step times without a sampler: 66.1-70.4 ms, tools/step_times.py). nvidia-smi stays as the
    fallback when NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index
        self.thread = None
        self.samples = []  # (sm_mhz, reasons bitmask)

    def _nvml_handle(self):
        import pynvml
        import torch

        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu)

    def start(self):
        import threading

        try:
            nv, h = self._nvml_handle()
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.stop_flag = threading.Event()

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(reasons_fn(h))))
                    except Exception:
                        pass
                    self.stop_flag.wait(0.2)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.gpu), "-lms", "200"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            sm = sorted(v for v, _ in self.samples)
            mask = 0
            for _, m in self.samples:
                mask |= m
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max,
                    "reasons": sorted(k for k, b in self.BITS.items() if mask & b), "samples": len(sm), "source": "nvml"}
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
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------- CPU baseline
# Bounded sample of one C2 step on the host cores (the reference has no CPU path of its own and cannot travel to the GPU
# box: SURVEY.md 8c/8d; the oracle port is pinned to the reference's vendored LDM at 1e-4, tests/test_oracle_ldm.py).
# The legs and the factors that scale them to one full step:
#   render   fwd+bwd  32 x 32 rays of the 256 x 256 view          x 64
#   VAE enc  fwd+bwd  256 x 256 crop of the 512 x 512 input       x 4
#   UNet     fwd      ONE of the five Perp-Neg batch entries at the FULL 64 x 64 latent (attention is quadratic in the
#                     token count, so the latent is not cropped)   x 5
CPU_SCALE = {"render_32x32_fwd_bwd": 64.0, "vae_256_fwd_bwd": 4.0, "unet_1x64x64_fwd": 5.0}


def cpu_sample_seconds():
    """One bounded sample (see CPU_SCALE) on all host cores with the oracle only: nothing of the product is imported
    or loaded (the parameter shapes come from oracle/ldm_param_specs.json). -> (seconds per leg, threads)."""
    import torch

    from oracle import ldm_oracle as lo, render_oracle as ro

    torch.set_num_threads(os.cpu_count() or 1)
    t_parts = {}
    fcfg = ro.FieldCfg()
    mcfg = ro.MarchCfg(render_step_size=1.732 * 2 * fcfg.radius / 512, prune=True)
    P = {k: v.clone().requires_grad_(True) for k, v in ro.make_field_params(fcfg, seed=0, table_scale=1e-4).items()}
    c2w = ro.look_at_c2w(torch.tensor([15.0]), torch.tensor([30.0]), torch.tensor([1.25]))
    rays_o, rays_d = ro.get_rays(c2w, torch.deg2rad(torch.tensor([55.0])), 32, 32)
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    occs, binary, _ = ro.occ_grid_from_density({k: v.detach() for k, v in P.items()}, fcfg, mcfg, seed=2)
    jitter = torch.rand(32 * 32, generator=torch.Generator().manual_seed(1))
    t0 = time.perf_counter()
    out = ro.render(rays_o, rays_d, jitter, None, binary.numpy(), float(occs.mean()), P, fcfg, mcfg, 1024)
    out["comp_rgb"].square().sum().backward()
    t_parts["render_32x32_fwd_bwd"] = time.perf_counter() - t0

    sd_v = lo.seeded_state_dict("vae", 1)
    x = (torch.rand(1, 3, 256, 256) * 2 - 1).requires_grad_(True)
    t0 = time.perf_counter()
    h = lo.vae_encoder_forward(sd_v, x)
    h.square().sum().backward()
    t_parts["vae_256_fwd_bwd"] = time.perf_counter() - t0
    del sd_v, h

    sd_u = lo.seeded_state_dict("unet_sd", 0)
    with torch.no_grad():
        xin, ctx = torch.randn(1, 4, 64, 64), torch.randn(1, 77, 1024)
        t0 = time.perf_counter()
        lo.unet_forward(sd_u, xin, torch.tensor([500.0]), ctx)
        t_parts["unet_1x64x64_fwd"] = time.perf_counter() - t0
    return t_parts, torch.get_num_threads()


def cpu_baseline_object(parts_list, cores):
    """cpu_baseline JSON object from one or more bounded samples: value = extrapolated full steps/s."""
    mean = {k: sum(p[k] for p in parts_list) / len(parts_list) for k in parts_list[0]}
    step_s = sum(mean[k] * CPU_SCALE[k] for k in mean)
    sample = ("oracle port, fp32 torch CPU: " + "; ".join(f"{k} {mean[k]:.2f} s x {CPU_SCALE[k]:g}" for k in mean)
              + f" -> {step_s:.1f} s per full C2 step (extrapolated)")
    return {"value": 1.0 / step_s, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample,
            "extrapolated": True, "sample_seconds": sum(mean.values()), "full_step_seconds": step_s}


def run_reference(args, rank: int) -> None:
    """`--impl reference`: the reference's math on the host cores. Each of the W + K steps really runs one bounded sample
    (about 5 s); `ms_per_step` is that measured time (steps x ms_per_step is the wall clock of the loop), `value` is the
    full-step rate extrapolated with CPU_SCALE and says so (`extrapolated: true`)."""
    if rank != 0:
        return
    parts, cores = [], 1
    t_loop = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        p, cores = cpu_sample_seconds()
        if i >= args.warmup:
            parts.append(p)
            t_loop.append(time.perf_counter() - t0)
    cb = cpu_baseline_object(parts, cores)
    v = cb["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(t_loop) / len(t_loop), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "extrapolated": True,
        "config": {"workload": WORKLOADS["C2"]["desc"],
                   "note": "reference math (oracle port pinned to the vendored LDM; the Python reference cannot travel to "
                           "the GPU box) on the host cores; every step runs one bounded sample, ms_per_step is its "
                           "measured time, value is the full-step rate extrapolated by the stated factors"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------- CUDA arm
class Job:
    """One workload built through the plugin API exactly as launch.py builds it (yaml -> data module + system), plus the
    step closure of Trainer.fit: hooks, training_step, backward, [all-reduce on a side stream], optimizer."""

    def __init__(self, name: str, rank: int, world: int, dev, tmpdir: str):
        import random

        import torch

        import scaledreamer_b200 as sd
        from scaledreamer_b200.systems import FusedAdamW, FusedAdan, Trainer

        self.name, self.rank, self.world, self.dev = name, rank, world, dev
        wl = WORKLOADS[name]
        self.views = wl["views"]
        torch.manual_seed(1234 + rank)
        random.seed(1234 + rank)
        cli = [f"data.width={W}", f"data.height={H}"] if wl["multiprompt"] else [f"data.width=[{W},{W}]",
                                                                                  f"data.height=[{H},{H}]"]
        if wl["multiprompt"]:
            # MG15-sized synthetic library; rank r keeps library[r::world] (custom/amortized/data/multiprompt.py:180-186)
            os.makedirs(os.path.join(tmpdir, "load"), exist_ok=True)
            lib_path = os.path.join(tmpdir, "load", "bench_lib.json")
            if not os.path.exists(lib_path):
                prompts = [f"a synthetic benchmark object number {i}" for i in range(max(16, 2 * world))]
                with open(lib_path + f".{rank}", "w") as f:
                    json.dump({"train": prompts, "val": prompts[:1], "test": prompts[:1]}, f)
                os.replace(lib_path + f".{rank}", lib_path)
            cli += ["system.prompt_processor.prompt_library=bench_lib",
                    f"data.prompt_library_dir={os.path.join(tmpdir, 'load')}",
                    f"system.prompt_processor.prompt_library_dir={os.path.join(tmpdir, 'load')}"]
        else:
            cli += ["system.prompt_processor.prompt=a DSLR photo of a hamburger"]
        cfg = sd.load_config(os.path.join(CFG_DIR, wl["yaml"]), cli_args=cli + wl["cli"])
        self.cfg = cfg
        self.accumulate = int(cfg.trainer.get("accumulate_grad_batches", 1) or 1)
        self.dm = sd.find(cfg.data_type)(cfg.data)
        self.system = sd.find(cfg.system_type)(cfg.system)
        self.dm.setup("fit")
        self.ds = self.dm.train_dataset
        self.system.train()
        self.system.on_fit_start()
        self.trainer = Trainer(max_steps=0, distributed=world > 1, accumulate_grad_batches=self.accumulate)
        self.trainer.sync_initial_state(self.system)  # what DDP does at wrap time: rank 0's generator everywhere
        self.opt = self.system.configure_optimizers()
        if isinstance(self.opt, (FusedAdamW, FusedAdan)):
            self.opt.grad_scale = 1.0 / (world * self.accumulate)
        self.params = [p for g in self.opt.param_groups for p in g["params"]]
        self.n_params = sum(p.numel() for p in self.params)
        self.step_no = 0
        self.marks = None  # when a list: (start, backward done, all-reduce done, end) events per step

    def host_batch(self):
        return self.ds.collate({})

    def to_device(self, hb):
        return self.ds.to_device(hb, self.dev)

    def step(self, batch_dev):
        import torch

        ev = (lambda: torch.cuda.Event(enable_timing=True)) if self.marks is not None else None
        m = [ev()] if ev else None
        if m:
            m[0].record()
        i = self.step_no
        self.ds.update_step(0, i)
        self.system.true_global_step = i // self.accumulate
        self.system.do_update_step(0, i // self.accumulate)
        out = self.system.training_step(batch_dev, i)
        out["loss"].backward()
        if m:
            m.append(ev())
            m[1].record()
        if (i + 1) % self.accumulate == 0:
            if self.world > 1:
                self.trainer._allreduce_grads(self.params)
            if m:
                m.append(ev())
                m[2].record()
            self.opt.step()
            self.opt.zero_grad(set_to_none=False)
        elif m:
            m.append(m[1])
        self.system.do_update_step_end(0, i // self.accumulate)
        if m:
            m.append(ev())
            m[3].record()
            self.marks.append(m)
        self.step_no += 1
        return out["loss"]


def timed_regions(job: Job, args, dist, sample_clocks: bool):
    """Region 1: K steps with the camera batches already resident in HBM. Region 2: the same K steps end to end (pinned
    host batch -> H2D every step, loss -> D2H every step). Both bracketed by barrier + synchronize; device-timed."""
    import torch

    from scaledreamer_b200 import lib as L

    world = job.world

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # the sampler starts BEFORE the warm-up: the first NVML queries of a process take tens of milliseconds with a driver
    # lock held (one run in three showed 97 ms of stalled launches at the start of the timed region otherwise); its
    # samples are discarded when the timed region begins
    clocks = ClockSampler(job.dev.index or 0)
    if sample_clocks:
        clocks.start()
    for w in range(args.warmup):
        # the last warm-up step also records the per-step timeline events, so that nothing (timing-enabled events, the
        # all-reduce side stream) is created for the first time inside the timed region
        job.marks = [] if w == args.warmup - 1 else None
        job.step(job.to_device(job.host_batch()))
    job.marks = None
    sync_all()
    resident = [job.to_device(job.host_batch()) for _ in range(args.steps)]
    sync_all()
    # The end-to-end leg keeps the host at most one step ahead of the GPU, so a host pause lands in the number: one
    # generation-2 garbage collection (~90 ms with torch + the UNet executor's objects alive) showed up as a 26 ms
    # average in one run out of five. Collect now and keep the collector off inside the timed regions.
    import gc
    gc.collect()
    gc.disable()
    clocks.samples.clear()
    n0 = L.launch_count()
    job.marks = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in resident:
        job.step(b)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - n0
    marks, job.marks = job.marks, None
    # per-step timeline of THIS rank: compute (start -> backward done), all-reduce (incl. waiting for the slowest rank),
    # optimizer + hooks
    tl = torch.tensor([[m[0].elapsed_time(m[1]), m[1].elapsed_time(m[2]), m[2].elapsed_time(m[3])] for m in marks],
                      device=job.dev, dtype=torch.float64)

    h2d = d2h = 0
    # The loss of step i is copied to pinned host memory asynchronously and read on the host one step later (after
    # step i+1 has been enqueued), the way a logging trainer consumes it: every step still pays its H2D inputs and a
    # D2H result, but the host never drains the GPU queue between steps.
    loss_pinned = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_events = [torch.cuda.Event(), torch.cuda.Event()]
    # one untimed step through exactly this path: the first pinned allocation / first D2H read of a process cost 97 ms of
    # host time on a fresh box (page-ins), which landed in step 0 of the timed leg
    loss = job.step(job.to_device(job.host_batch()))
    loss_pinned[0].copy_(loss.detach(), non_blocking=True)
    loss_events[0].record()
    loss_events[0].synchronize()
    float(loss_pinned[0])
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    losses_host = []
    host_s = host_max = 0.0
    for i in range(args.steps):
        t_host = time.perf_counter()
        hb = job.host_batch()
        h2d = sum(v.numel() * v.element_size() for v in hb.values() if torch.is_tensor(v))
        loss = job.step(job.to_device(hb))
        dt_host = time.perf_counter() - t_host
        host_s += dt_host
        host_max = max(host_max, dt_host)
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
    gc.enable()
    clk = clocks.stop() if sample_clocks else {}

    t = torch.tensor([ms, ms_e2e], device=job.dev, dtype=torch.float64)
    timeline = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allr = [torch.empty_like(tl) for _ in range(world)]
        dist.all_gather(allr, tl)
        st = torch.stack(allr)  # [rank, step, 3]
        comp = st[:, :, 0]
        timeline = {
            "compute_ms_mean_over_ranks": float(comp.mean()),
            "compute_ms_max_over_ranks_mean_over_steps": float(comp.max(0).values.mean()),
            "rank_skew_ms": float((comp.max(0).values - comp.mean(0)).mean()),
            # a rank that is slow on EVERY step is slower hardware (power / thermals), step-to-step changes are data
            "compute_ms_mean_per_rank": [round(float(v), 2) for v in comp.mean(1)],
            "compute_ms_std_over_steps_per_rank": [round(float(v), 2) for v in comp.std(1)],
            "allreduce_ms_fastest_rank_mean": float(st[:, :, 1].min(0).values.mean()),  # the last rank to arrive: pure NCCL
            "allreduce_ms_mean": float(st[:, :, 1].mean()),
            "optimizer_ms_mean": float(st[:, :, 2].mean()),
            "flat_gradient_mb": job.n_params * 4 / 1e6,
            "how": "CUDA events on each rank's stream inside the timed region; all-reduce runs on a side stream the "
                   "optimizer waits on; its time on a rank includes waiting for the slowest rank's backward",
        }
    else:
        timeline = {"compute_ms_mean_over_ranks": float(tl[:, 0].mean()), "optimizer_ms_mean": float(tl[:, 2].mean()),
                    # every step of the timed region (camera-dependent: 18-25 ms on C2); a single step far outside that
                    # range is a stall of the box, not of the step
                    "compute_ms_per_step": [round(float(v), 2) for v in tl[:, 0]],
                    "compute_ms_median": float(tl[:, 0].median())}
    return dict(ms=float(t[0]), ms_e2e=float(t[1]), launches=int(launches), h2d=h2d, d2h=d2h, clocks=clk,
                timeline=timeline, host_enqueue_ms=1e3 * host_s / args.steps, host_enqueue_ms_max=1e3 * host_max)


def profile_pass(job: Job, n_prof: int = 3):
    """Outside the timed regions (same workload): per-phase CUDA-event times, every tcgen05 launch bracketed by events on
    its stream (sdb_gemm_profile_*), and every other C-ABI call bracketed the same way (lib.call_timer_*)."""
    import ctypes as C

    import torch

    from scaledreamer_b200 import lib as L

    lib = L.load()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    acc = {"forward_render_and_guidance": 0.0, "backward": 0.0, "optimizer": 0.0}
    lib.sdb_gemm_profile_begin()
    if os.environ.get("SDB_GEMM_CSV"):
        lib.sdb_gemm_profile_dump(os.environ["SDB_GEMM_CSV"].encode())
    L.call_timer_begin()
    system, opt = job.system, job.opt
    for i in range(n_prof):
        b = job.to_device(job.host_batch())
        a0, a2, a3, a4 = ev(), ev(), ev(), ev()
        a0.record()
        lossp = system.training_step(b, job.step_no + i)["loss"]  # the yaml's own loss terms (eikonal, sparsity, ...)
        a2.record()
        lossp.backward()
        a3.record()
        opt.step()
        opt.zero_grad(set_to_none=False)
        a4.record()
        torch.cuda.synchronize()
        for k, (x, y) in zip(acc, ((a0, a2), (a2, a3), (a3, a4))):
            acc[k] += x.elapsed_time(y) / n_prof
    calls = L.call_timer_end()
    gm, gf, gl = C.c_double(), C.c_double(), C.c_int()
    L.check(L.load().sdb_gemm_profile_end(C.byref(gm), C.byref(gf), C.byref(gl)), "profile_end")
    per_call = {k: {"ms_per_step": v["ms"] / n_prof, "calls_per_step": v["calls"] / n_prof,
                    "launches_per_step": v["launches"] / n_prof} for k, v in calls.items()}
    return dict(phase_ms=acc, gemm_ms_per_step=gm.value / n_prof, gemm_tflop_per_step=gf.value / n_prof / 1e12,
                gemm_launches_per_step=gl.value // n_prof, abi_calls=per_call)


def nerf_kernels_alone(job: Job):
    """C2 / C3: the fused render kernels alone on one camera batch (sample count for the algorithmic bytes)."""
    import torch

    from scaledreamer_b200 import render_ops as R

    ev = lambda: torch.cuda.Event(enable_timing=True)
    rr, dev = job.system.renderer, job.dev
    b = job.to_device(job.host_batch())
    P = {k: v.detach() for k, v in rr._params().items()}
    march = R.MarchSpec(render_step_size=rr.render_step_size, prune=True, grid_res=32)
    ro_, rd_ = b["rays_o"].reshape(-1, 3).contiguous(), b["rays_d"].reshape(-1, 3).contiguous()
    n_rays = ro_.shape[0]
    jit = torch.rand(n_rays, device=dev)
    tape = R.RenderTape.acquire(march, rr._spec().radius, n_rays, dev)
    o = R.render_forward_v2_raw(rr._spec(), march, P, rr._occ_grid(dev), ro_, rd_, jit, None, H * W, tape)
    kept = int(tape.counter[0].item())
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
    return dict(render_fwd_kernel_ms=min(t_f), render_bwd_kernel_ms=min(t_b), render_samples_kept=kept,
                render_tapes_allocated=R.RenderTape.n_allocated)


def rooflines(job: Job, prof: dict, pk: dict):
    """roofline objects, most expensive family first. Algorithmic units are SURVEY.md 8(d)'s."""
    tr = ncu_traffic()
    n_rays = job.views * H * W
    gemm_tfs = prof["gemm_tflop_per_step"] / (prof["gemm_ms_per_step"] / 1e3)
    roofs = [{"kernel": "gemm_f16_kernel + flash_attn_f16_kernel (tcgen05 GEMM / implicit conv / fused attention, all "
                        "launches of one step)", "bound": "tensor", "achieved": gemm_tfs, "peak": pk["tf_sustained"],
              "unit": "TFLOP/s", "frac": gemm_tfs / pk["tf_sustained"], "traffic": tr.get("tensor"),
              "ms_per_step": prof["gemm_ms_per_step"], "peak_source": pk["src"] + " sustained bf16"}]

    def hbm(kernel, gbytes, ms, traffic=None):
        a = gbytes / (ms / 1e3)
        return {"kernel": kernel, "bound": "hbm", "achieved": a, "peak": pk["hbm"], "unit": "GB/s", "frac": a / pk["hbm"],
                "traffic": traffic, "ms_per_step": ms, "peak_source": pk["src"] + " copy bandwidth"}

    if "render_samples_kept" in prof:  # NeRF path: samples x 16 levels x 8 corners x 2 features x 4 B, E = 1
        ns = prof["render_samples_kept"]
        roofs.append(hbm("render_composite_bwd_kernel + render_field_bwd_tc_kernel (tape backward)",
                         (2 * ns * 1024 + n_rays * 52) / 1e9, prof["render_bwd_kernel_ms"], tr.get("render_bwd")))
        roofs.append(hbm("render_bg_kernel + render_nerf_fwd2_kernel (march + encode + MLPs + composite + tape)",
                         (ns * 1024 + n_rays * 52) / 1e9, prof["render_fwd_kernel_ms"], tr.get("render_fwd")))
    else:  # VolSDF importance path: Nr x (nc proposal + S x 4) field evaluations per step
        rcfg = job.system.renderer.cfg
        nc, S = rcfg.num_samples_per_ray_importance, rcfg.num_samples_per_ray_importance + rcfg.num_samples_per_ray + 1
        per_point = 1024 if job.name == "C4" else 1536
        calls = prof["abi_calls"]
        fwd_names = ("sdb_hyper_field_forward",) if job.name == "C4" else ("sdb_triplane_sample_forward",)
        bwd_names = ("sdb_hyper_field_backward",) if job.name == "C4" else ("sdb_triplane_sample_backward",)
        ms_f = sum(calls[k]["ms_per_step"] for k in fwd_names if k in calls)
        ms_b = sum(calls[k]["ms_per_step"] for k in bwd_names if k in calls)
        # the environment map of C4 goes through the same entry point: n_rays more points per direction
        extra = n_rays if job.name == "C4" else 0
        if ms_f > 0:
            roofs.append(hbm(" + ".join(fwd_names) + " (field lookups of the VolSDF renderer: proposal, centres, three "
                             "finite-difference offsets)", ((nc + 4 * S) * n_rays + extra) * per_point / 1e9, ms_f,
                             tr.get("hyper_field_fwd") if job.name == "C4" else None))
        if ms_b > 0:
            roofs.append(hbm(" + ".join(bwd_names) + " (field backward: MLP contractions + scatter)",
                             (4 * S * n_rays + extra) * per_point / 1e9, ms_b,
                             tr.get("hyper_field_bwd") if job.name == "C4" else None))
        if job.name == "C5" and "sdb_gemm_tf32" in calls:
            # the trained generator's contractions: 2 * M * N * K of every sdb_gemm_tf32 of one step (forward 3.87 TFLOP at
            # 4 prompts (0.97 at the one prompt x 4 views of a C5 step), backward twice that plus the recomputed score products). tf32 runs at half the bf16 rate: the
            # peak is half the measured bf16 figure (no tf32 number in MEASURED_PEAKS.json).
            gen = job.system.geometry.space_generator
            L_, Cg = gen.pos_embed.shape[1], gen.pos_embed.shape[2]
            Hh, nl, nb = gen.layers[0].self_attn.heads, len(gen.layers), max(1, int(job.cfg.data["batch_size"]) // int(job.cfg.data.get("n_view", 1)))
            lin = lambda m, n, k: 2.0 * m * n * k
            fwd = nl * nb * (6 * lin(L_, Cg, Cg) + 2 * lin(77, Cg, 1024) + 2 * lin(L_, 4 * Cg, Cg)
                             + Hh * 2 * (lin(L_, L_, Cg // Hh) + lin(L_, 77, Cg // Hh)))
            score = nl * nb * Hh * (lin(L_, L_, Cg // Hh) + lin(L_, 77, Cg // Hh))
            tfl = (3 * fwd + score) / 1e12  # backward = 2 x forward + the recomputed score product S
            ms_g = calls["sdb_gemm_tf32"]["ms_per_step"]
            roofs.append({"kernel": "gemm_tf32_kernel (Triplane-Transformer generator, forward + backward)", "bound": "tensor",
                          "achieved": tfl / (ms_g / 1e3), "peak": pk["tf_sustained"] / 2, "unit": "TFLOP/s",
                          "frac": tfl / (ms_g / 1e3) / (pk["tf_sustained"] / 2), "traffic": None, "ms_per_step": ms_g,
                          "peak_source": pk["src"] + " sustained bf16 / 2 (tf32)",
                          "note": "60 of the step's products are 3072 x 3072 x 48 score matrices written in fp32: those "
                                  "are bound by their 2.4 GB outputs (4.3 TB/s), not by the tensor pipe"})
    return sorted(roofs, key=lambda r: -r["ms_per_step"])


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS),
                    help="C2 = the BASELINE metric's configuration (default); C3 / C4 / C5 = the other BASELINE configs at "
                         "256x256 (same loop; C4 is also measured into the C2 line's `c4` block when --gpus > 1)")
    ap.add_argument("--no-c4", action="store_true", help="skip the extra C4 measurement of multi-GPU runs")
    args = ap.parse_args()
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
    tmpdir = os.path.join(tempfile.gettempdir(), f"sdb_bench_{os.environ.get('MASTER_PORT', 'single')}_{os.getppid()}")
    os.makedirs(tmpdir, exist_ok=True)

    job = Job(args.workload, rank, world, dev, tmpdir)
    res = timed_regions(job, args, dist, sample_clocks=rank == 0)
    prof, roofs = {}, None
    if rank == 0:
        prof = profile_pass(job)
        if args.workload in ("C2", "C3"):
            prof.update(nerf_kernels_alone(job))
        roofs = rooflines(job, prof, peaks())
    c4 = None
    if world > 1 and args.workload == "C2" and not args.no_c4:
        # the prompt-sharded configuration north_star names for 8 GPUs, measured in the same run (every rank takes part)
        del job
        torch.cuda.empty_cache()
        job4 = Job("C4", rank, world, dev, tmpdir)
        a4 = argparse.Namespace(steps=max(4, args.steps // 2), warmup=3)
        r4 = timed_regions(job4, a4, dist, sample_clocks=False)
        c4 = {"workload": WORKLOADS["C4"]["desc"], "value": world * a4.steps / (r4["ms"] / 1e3), "unit": "steps/s",
              "steps": a4.steps, "ms_per_step": r4["ms"] / a4.steps,
              "e2e": {"value": world * a4.steps / (r4["ms_e2e"] / 1e3), "ms_per_step": r4["ms_e2e"] / a4.steps},
              "timeline": r4["timeline"], "gpu_launches_per_step": r4["launches"] // a4.steps}
        job = job4

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    ms, ms_e2e, steps = res["ms"], res["ms_e2e"], args.steps
    wl = WORKLOADS[args.workload]
    line = {
        "metric": METRIC, "value": world * steps / (ms / 1e3), "unit": "steps/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (render: f32)",
        "data": "synthetic",
        "config": {"workload": wl["desc"], "weights": "seeded synthetic (865.9 M UNet, 34.2 M VAE encoder)",
                   "parallelism": f"dp{world}: one prompt+camera batch per GPU, one flat all-reduce of generator grads per "
                                  "optimizer step",
                   "l2": "no explicit flush: each step streams 1.73 GB of UNet weights + >1 GB activations (>> 126 MB L2)"},
        "e2e": {"value": world * steps / (ms_e2e / 1e3), "unit": "steps/s", "h2d_bytes_per_step": res["h2d"],
                "d2h_bytes_per_step": res["d2h"], "ms_per_step": ms_e2e / steps,
                "host_enqueue_ms_per_step": res["host_enqueue_ms"], "host_enqueue_ms_max": res["host_enqueue_ms_max"],
                "gc": "collected before, disabled inside the timed regions"},
        "gpu_launches": res["launches"], "gpu_launches_per_step": res["launches"] // steps,
        "clocks": res["clocks"], "roofline": roofs[0], "roofline_other_kernels": roofs[1:], "profile": prof,
        "timeline": res["timeline"],
    }
    line["views_per_s"] = line["value"] * wl["views"]
    if c4 is not None:
        line["c4"] = c4
    if world == 1 and not args.no_cpu_baseline and args.workload == "C2":
        parts, cores = cpu_sample_seconds()
        line["cpu_baseline"] = cpu_baseline_object([parts], cores)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
