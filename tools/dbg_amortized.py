import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import amortized_oracle as ao, render_oracle as ro
from tests.helpers import rel_l2
import scaledreamer_b200 as sd
dev = torch.device("cuda:0")
torch.manual_seed(0)
geo = sd.find("Hyper-iNGP")({"radius": 2.0, "sdf_bias": "sphere", "sdf_bias_params": 0.5,
    "hypernet_config": {"c_dim": 1024, "out_dims": {"sdf_weights": [64, 1], "feature_weights": [64, 3]}, "spectral_norm": False, "n_neurons": 64, "n_hidden_layers": 1}}).to(dev)
mat = sd.find("no-material")({"n_output_dims": 3, "color_activation": "sigmoid", "requires_normal": True}).to(dev)
bgm = sd.find("multiprompt-neural-hashgrid-environment-map-background")({"color_activation": "sigmoid", "random_aug": False,
    "pos_encoding_config": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16, "per_level_scale": 1.0}}).to(dev)
ren = sd.find("generative-space-volsdf-volume-renderer")({"radius": 2.0, "use_volsdf": True, "trainable_variance": False, "learned_variance_init": 0.340119,
    "estimator": "importance", "num_samples_per_ray": 64, "num_samples_per_ray_importance": 128, "near_plane": 0.1, "far_plane": 4.0}, geometry=geo, material=mat, background=bgm).to(dev)
ren.train(); geo.update_step(0, 0)
with torch.no_grad():
    geo.encoding.encoding.params.mul_(500.0)
B, H, W = 2, 6, 5
g = torch.Generator().manual_seed(3)
o = (torch.tensor([0.0, -1.6, 0.3]) + 0.05 * torch.randn(B, 1, 1, 3, generator=g)).expand(B, H, W, 3).contiguous()
d = torch.nn.functional.normalize(torch.tensor([0.0, 1.0, -0.15]) + 0.2 * torch.randn(B, H, W, 3, generator=g), dim=-1)
emb = torch.randn(B, 1024, generator=g)
uc, uf = torch.rand(B * H * W, generator=g), torch.rand(B * H * W, generator=g)
gimg = torch.randn(B * H * W, 3, generator=g)
hcfg, vcfg = ao.HyperCfg(), ao.VolSDFCfg()
hp = {k: v.detach().cpu() for k, v in geo.hypernet.state_dict().items()}
bg_ref = torch.rand(B*H*W, 3, generator=g)
for name in ("fg", "opacity", "depth", "eik"):
    geo.zero_grad()
    out = ren(o.to(dev), d.to(dev), None, bg_color=bg_ref.view(B,H,W,3).to(dev), text_embed=emb.to(dev), u_coarse=uc.to(dev), u_fine=uf.to(dev))
    table = geo.encoding.table.detach().cpu().view(-1, 2).clone().requires_grad_(True)
    cache = ao.hypernet_forward(hp, emb, {"sdf_weights": [32, 64, 1], "feature_weights": [32, 64, 3]})
    ref = ao.render(o.view(-1, 3), d.view(-1, 3), H * W, table, cache, bg_ref, hcfg, vcfg, uc, uf)
    if name == "fg":
        l1, l2 = (out["comp_rgb_fg"].view(-1,3) * gimg.to(dev)).sum(), (ref["comp_rgb_fg"] * gimg).sum()
    elif name == "opacity":
        l1, l2 = out["opacity"].sum(), ref["opacity"].sum()
    elif name == "depth":
        l1, l2 = out["depth"].sum(), ref["depth"].sum()
    else:
        l1, l2 = ((torch.linalg.norm(out["sdf_grad"], ord=2, dim=-1) - 1.0) ** 2).mean(), ao.eikonal_loss(ref["sdf_grad"])
    l1.backward(); l2.backward()
    gt = geo.encoding.table.grad.cpu().view(-1, 2)
    print(name, 'loss', float(l1), float(l2), 'table grad rel', rel_l2(gt, table.grad), 'norms', float(gt.norm()), float(table.grad.norm()),
          'max t diff', float((out["t_points"].view(-1).cpu() - ref["t_points"].reshape(-1)).abs().max()))
