import torch, sys
sys.path.insert(0, '/root/repo')
from oracle import render_oracle as ro
from tests.helpers import *
from scaledreamer_b200 import render_ops as R, lib as L
dev = torch.device('cuda:0')
# raygen
sc = scene(H=20, W=28, B=3, seed=29)
o = torch.empty(3, 20, 28, 3, device=dev); d = torch.empty_like(o)
L.check(L.load().sdb_raygen(L.ptr(sc["c2w"].to(dev)), L.ptr(sc["fovy"].to(dev)), 3, 20, 28, L.ptr(o), L.ptr(d), L.stream_ptr()), "raygen")
dd = (d.cpu().reshape(-1,3) - sc["rays_d"]).abs()
print("raygen max diff per comp", dd.max(0).values, "per image", dd.reshape(3,-1,3).amax((1,2)))
print("c2w", sc["c2w"][0]); print(d[0,0,0].cpu(), sc["rays_d"][0])
# hashgrid
for cfg in (ro.GridCfg(), ro.GridCfg(4, 2, 19, 4, 4.0)):
    n = ro.grid_meta(cfg)["n_entries"]
    g = torch.Generator().manual_seed(0)
    table = torch.randn(n, 2, generator=g).requires_grad_(True)
    x = torch.rand(4099, 3, generator=g)
    x[:3] = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0, 0.5, 1.0]])
    enc = ro.hashgrid_encode(x, table, cfg)
    got = R.hashgrid_forward(x.to(dev), table.detach().to(dev), vars(cfg)).cpu()
    print("hashgrid fwd rel", rel_l2(got, enc.detach()), "maxabs", (got-enc.detach()).abs().max().item())
    per_level = ((got-enc.detach())**2).reshape(4099, -1, 2).sum((0,2)).sqrt()
    print(" per-level err", per_level)
    go = torch.randn(enc.shape, generator=g)
    enc.backward(go)
    gt = R.hashgrid_backward(x.to(dev), go.to(dev), n, vars(cfg)).cpu()
    print("hashgrid bwd rel", rel_l2(gt, table.grad))
# render bwd
sc = scene(H=24, W=24, B=2, seed=17)
P = {k: v.clone().requires_grad_(True) for k, v in sc["P"].items()}
ref = ro.render(sc["rays_o"], sc["rays_d"], sc["jitter"], None, sc["binary"].numpy(), float(sc["occs"].mean()), P, sc["fcfg"], sc["mcfg"], 576)
g = torch.Generator().manual_seed(1)
g_rgb, g_op, g_dp = torch.randn(1152, 3, generator=g), torch.randn(1152, generator=g), torch.randn(1152, generator=g)
(ref["comp_rgb"] * g_rgb).sum().add((ref["opacity"] * g_op).sum()).add((ref["depth"] * g_dp).sum()).backward()
spec = field_spec_from_oracle(sc["fcfg"]); march = march_spec_from_oracle(sc["mcfg"])
occ = R.OccGrid(32, dev); occ.set_binaries(sc["binary"], sc["occs"])
Pd = {k: v.to(dev).contiguous() for k, v in sc["P"].items()}
out = R.render_forward_raw(spec, march, Pd, occ, sc["rays_o"].to(dev).contiguous(), sc["rays_d"].to(dev).contiguous(), sc["jitter"].to(dev), None, 576, 0)
for k in ("comp_rgb","opacity","depth"): print("fwd", k, rel_l2(out[k].cpu(), ref[k].detach()))
grads = {k: torch.zeros_like(v) for k, v in Pd.items()}
R.render_backward_raw(spec, march, Pd, grads, occ, sc["rays_o"].to(dev).contiguous(), sc["rays_d"].to(dev).contiguous(), sc["jitter"].to(dev), None, 576, out, g_rgb.to(dev), g_op.to(dev), g_dp.to(dev))
torch.cuda.synchronize()
for k in R.PARAM_KEYS: print("bwd", k, rel_l2(grads[k].cpu(), P[k].grad))
m = ro.grid_meta(sc["fcfg"].grid)
gt, gr = grads["table"].cpu(), P["table"].grad
for l in range(16):
    a, b = m["offset"][l], m["offset"][l]+m["size"][l]
    print(" level", l, rel_l2(gt[a:b], gr[a:b]), float(gr[a:b].norm()))
