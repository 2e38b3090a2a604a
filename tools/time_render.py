"""Times the render kernels alone on the bench scene (C2, 256x256) -- a diagnostic, never a bench number."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scaledreamer_b200 as sd
from scaledreamer_b200 import render_ops as R

CFG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "configs", "asd_sd_nerf.yaml")
torch.manual_seed(1234); random.seed(1234)
HW = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = sd.load_config(CFG, cli_args=["system.prompt_processor.prompt=a DSLR photo of a hamburger", f"data.width=[{HW},{HW}]", f"data.height=[{HW},{HW}]"])
dev = torch.device("cuda:0")
dm = sd.find(cfg.data_type)(cfg.data); dm.setup("fit"); ds = dm.train_dataset
system = sd.find(cfg.system_type)(cfg.system); system.train()
rr = system.renderer
ds.update_step(0, 0); system.do_update_step(0, 0)   # occupancy refresh at step 0
march = R.MarchSpec(render_step_size=rr.render_step_size, prune=True, grid_res=32)
P = {k: v.detach() for k, v in rr._params().items()}
spec = rr._spec()
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(6):
    b = ds.to_device(ds.collate({}), dev)
    ro_, rd_ = b["rays_o"].reshape(-1, 3).contiguous(), b["rays_d"].reshape(-1, 3).contiguous()
    jit = torch.rand(HW * HW, device=dev)
    bgo = torch.rand(1, 3, device=dev) if it % 2 else None
    tape = R.RenderTape.acquire(march, spec.radius, HW * HW, dev)
    grads = {k: torch.zeros_like(v) for k, v in P.items()}
    g_rgb = torch.randn(HW * HW, 3, device=dev)
    torch.cuda.synchronize()
    e = [ev() for _ in range(6)]
    e[0].record()
    o1 = R.render_forward_raw(spec, march, P, rr._occ_grid(dev), ro_, rd_, jit, bgo, HW * HW, 0)
    e[1].record()
    o2n = R.render_forward_v2_raw(spec, march, P, rr._occ_grid(dev), ro_, rd_, jit, bgo, HW * HW, None)
    e[2].record()
    o2 = R.render_forward_v2_raw(spec, march, P, rr._occ_grid(dev), ro_, rd_, jit, bgo, HW * HW, tape)
    e[3].record()
    R.render_backward_tape_raw(spec, march, P, grads, rd_, bgo, HW * HW, o2, tape, g_rgb)
    e[4].record()
    R.render_backward_raw(spec, march, P, grads, rr._occ_grid(dev), ro_, rd_, jit, bgo, HW * HW, o1, g_rgb)
    e[5].record()
    torch.cuda.synchronize()
    n = int(tape.counter[0].item()); tape.release()
    d = (o1["comp_rgb"] - o2["comp_rgb"]).abs().max().item()
    print(f"it{it} bg_override={bgo is not None} kept={n} fwd_v1={e[0].elapsed_time(e[1]):.2f} fwd_v2_notape={e[1].elapsed_time(e[2]):.2f} "
          f"fwd_v2_tape={e[2].elapsed_time(e[3]):.2f} bwd_tape={e[3].elapsed_time(e[4]):.2f} bwd_v1={e[4].elapsed_time(e[5]):.2f} maxdiff={d:.2e}", flush=True)
