"""Runs W warm-up ASD steps, then brackets ONE step with cudaProfilerStart/Stop (use with
`ncu --profile-from-start off ...`). Never a benchmark: numbers under a profiler are not bench values."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scaledreamer_b200 as sd

CFG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "configs", "asd_sd_nerf.yaml")
torch.manual_seed(1234); random.seed(1234)
cfg = sd.load_config(CFG, cli_args=["system.prompt_processor.prompt=a DSLR photo of a hamburger", "data.width=[256,256]", "data.height=[256,256]"])
dev = torch.device("cuda:0")
dm = sd.find(cfg.data_type)(cfg.data); dm.setup("fit"); ds = dm.train_dataset
system = sd.find(cfg.system_type)(cfg.system); system.train(); system.on_fit_start()
opt = system.configure_optimizers()
def step(i):
    ds.update_step(0, i); system.true_global_step = i; system.do_update_step(0, i)
    b = ds.to_device(ds.collate({}), dev)
    out = system.training_step(b, i); out["loss"].backward(); opt.step(); opt.zero_grad(set_to_none=False)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(W): step(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(W)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
