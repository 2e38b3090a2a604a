import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scaledreamer_b200 as sd
CFG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "configs", "asd_sd_nerf.yaml")
res = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = sd.load_config(CFG, cli_args=["system.prompt_processor.prompt=a hamburger", f"data.width=[{res},{res}]", f"data.height=[{res},{res}]"])
dev = torch.device("cuda:0")
dm = sd.find(cfg.data_type)(cfg.data); dm.setup("fit"); ds = dm.train_dataset
system = sd.find(cfg.system_type)(cfg.system); system.train(); system.on_fit_start()
opt = system.configure_optimizers()
def st(name, t):
    t = t.float(); print(f"  {name:10s} shape {tuple(t.shape)} finite {bool(torch.isfinite(t).all())} absmean {float(t.abs().mean()):.4e} max {float(t.abs().max()):.4e}")
for step in range(3):
    ds.update_step(0, step); system.true_global_step = step; system.do_update_step(0, step)
    b = ds.to_device(ds.collate({}), dev)
    out = system.training_step(b, step)
    out["loss"].backward()
    torch.cuda.synchronize()
    g = system.guidance
    print("step", step, "loss", float(out["loss"]), "loss_asd", float(system.logged["train/loss_asd"]), "t", g._last["t"].tolist(), "t_plus", g.buf["t_plus"].tolist())
    for k in ("img", "h", "latents", "unet_x", "unet_t", "eps", "neg_w", "grad", "d_h", "d_img", "d_rgb"): st(k, g.buf[k])
    for n, p in system.named_parameters():
        if p.grad is not None: st("g:" + n[-28:], p.grad)
    opt.step(); opt.zero_grad(set_to_none=False)
