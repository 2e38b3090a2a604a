"""Parity at BASELINE sizes (256 x 256 rays per view), one test per render path:

  C2  fused NeRF renderer (csrc/render_fwd2.cu), every one of the 65 536 rays against oracle/render_oracle.py::render
  C4  VolSDF importance renderer over the Hyper-iNGP field (csrc/volsdf.cu + hyper_field.cu): the device renders the
      full 256 x 256 view; the oracle (a few minutes of CPU for all rays) checks every 4th ray -- rays are independent
      given their random draws, so a strided subset is an unbiased check of the full-size launch.

The oracle side runs in ray chunks under no_grad to bound its memory. Forward outputs only (gradients are covered at
small sizes in test_render_gpu.py / test_amortized_gpu.py; the size-independent properties checked here on ALL rays are
comp_rgb = fg + bg (1 - opacity), 0 <= opacity <= 1, and depth inside [near, far] x opacity).
"""
import pytest
import torch

from oracle import amortized_oracle as ao, render_oracle as ro
from tests.helpers import field_spec_from_oracle, march_spec_from_oracle, rel_l2, scene

pytestmark = pytest.mark.gpu


def test_c2_fused_nerf_render_256x256_matches_oracle(cuda_device):
    from scaledreamer_b200 import render_ops as R

    H = W = 256
    sc = scene(H=H, W=W, B=1, seed=41)
    spec, march = field_spec_from_oracle(sc["fcfg"]), march_spec_from_oracle(sc["mcfg"])
    occ = R.OccGrid(32, cuda_device)
    occ.set_binaries(sc["binary"], sc["occs"])
    P = {k: v.to(cuda_device) for k, v in sc["P"].items()}
    tape = R.RenderTape.acquire(march, spec.radius, H * W, cuda_device)
    out = R.render_forward_v2_raw(spec, march, P, occ, sc["rays_o"].to(cuda_device), sc["rays_d"].to(cuda_device),
                                  sc["jitter"].to(cuda_device), None, H * W, tape)
    torch.cuda.synchronize()
    kept = int(tape.counter[0].item())
    tape.check_overflow()
    tape.release()
    ref = {k: [] for k in ("comp_rgb", "comp_rgb_fg", "comp_rgb_bg", "opacity", "depth", "z_variance")}
    n_ref = 0
    with torch.no_grad():
        for s0 in range(0, H * W, 4096):
            sl = slice(s0, s0 + 4096)
            r = ro.render(sc["rays_o"][sl], sc["rays_d"][sl], sc["jitter"][sl], None, sc["binary"].numpy(),
                          float(sc["occs"].mean()), sc["P"], sc["fcfg"], sc["mcfg"], H * W)
            n_ref += int(r["weights"].numel())
            for k in ref:
                ref[k].append(r[k])
    ref = {k: torch.cat(v) for k, v in ref.items()}
    print(f"C2 256x256: kept samples device {kept} / oracle {n_ref}")
    assert abs(kept - n_ref) <= max(64, n_ref // 5000)  # alpha-threshold flips at fp32 rounding
    for k in ("comp_rgb", "comp_rgb_fg", "comp_rgb_bg", "opacity", "depth"):
        err = rel_l2(out[k].cpu(), ref[k])
        print(k, err)
        assert err < 1e-3, (k, err)
    assert (out["comp_rgb"].cpu() - ref["comp_rgb"]).abs().max() < 2e-2
    zerr = rel_l2(out["z_variance"].cpu(), ref["z_variance"])
    assert zerr < 5e-3, zerr
    # size-independent properties on every ray
    op = out["opacity"]
    assert float(op.min()) >= 0.0 and float(op.max()) <= 1.0 + 1e-5
    blend = out["comp_rgb_fg"] + out["comp_rgb_bg"] * (1.0 - op[:, None])
    assert (blend - out["comp_rgb"]).abs().max() < 1e-6
    assert float(op.mean()) > 0.05  # the blob is in view: the comparison is not about an empty image


def test_c4_volsdf_hyper_render_256x256_matches_oracle_on_strided_rays(cuda_device):
    import scaledreamer_b200 as sd

    dev = cuda_device
    torch.manual_seed(0)
    geo = sd.find("Hyper-iNGP")({"radius": 2.0, "sdf_bias": "sphere", "sdf_bias_params": 0.5,
                                 "hypernet_config": {"c_dim": 1024, "out_dims": {"sdf_weights": [64, 1], "feature_weights": [64, 3]},
                                                     "spectral_norm": False, "n_neurons": 64, "n_hidden_layers": 1}}).to(dev)
    mat = sd.find("no-material")({"n_output_dims": 3, "color_activation": "sigmoid", "requires_normal": True}).to(dev)
    bgm = sd.find("multiprompt-neural-hashgrid-environment-map-background")(
        {"color_activation": "sigmoid", "random_aug": False,
         "pos_encoding_config": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19,
                                 "base_resolution": 16, "per_level_scale": 1.0}}).to(dev)
    ren = sd.find("generative-space-volsdf-volume-renderer")(
        {"radius": 2.0, "use_volsdf": True, "trainable_variance": False, "learned_variance_init": 0.340119,
         "estimator": "importance", "num_samples_per_ray": 64, "num_samples_per_ray_importance": 128, "near_plane": 0.1,
         "far_plane": 4.0}, geometry=geo, material=mat, background=bgm).to(dev)
    ren.train()
    geo.update_step(0, 0)
    with torch.no_grad():
        geo.encoding.encoding.params.mul_(500.0)
    H = W = 256
    c2w = ro.look_at_c2w(torch.tensor([20.0]), torch.tensor([35.0]), torch.tensor([1.6]))
    o, d = ro.get_rays(c2w, torch.deg2rad(torch.tensor([60.0])), H, W)
    g = torch.Generator().manual_seed(7)
    emb = torch.randn(1, 1024, generator=g)
    uc, uf = torch.rand(H * W, generator=g), torch.rand(H * W, generator=g)
    with torch.no_grad():
        out = ren(o.view(1, H, W, 3).to(dev), d.view(1, H, W, 3).to(dev), None, text_embed=emb.to(dev),
                  u_coarse=uc.to(dev), u_fine=uf.to(dev))
    torch.cuda.synchronize()

    idx = torch.arange(0, H * W, 4)
    hcfg, vcfg = ao.HyperCfg(), ao.VolSDFCfg()
    table = geo.encoding.table.detach().cpu().view(-1, 2)
    with torch.no_grad():
        cache = ao.hypernet_forward({k: v.detach().cpu() for k, v in geo.hypernet.state_dict().items()}, emb,
                                    {"sdf_weights": [32, 64, 1], "feature_weights": [32, 64, 3]})
        bg_cache = ao.hypernet_forward({k: v.detach().cpu() for k, v in bgm.hypernet.state_dict().items()}, emb,
                                       {"bg_weights": [32, 64, 3]})
        oo, dd = o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx]
        ref = {k: [] for k in ("comp_rgb", "comp_rgb_fg", "comp_rgb_bg", "opacity", "depth")}
        for s0 in range(0, idx.numel(), 2048):
            sl = slice(s0, s0 + 2048)
            n = oo[sl].shape[0]
            bg_ref = ao.hyper_background(dd[sl].view(1, n, 3), bgm.encoding.table.detach().cpu().view(-1, 2),
                                         bg_cache["bg_weights"], ro.GridCfg(16, 2, 19, 16, 1.0)).view(-1, 3)
            r = ao.render(oo[sl], dd[sl], n, table, cache, bg_ref, hcfg, vcfg, uc[idx][sl], uf[idx][sl])
            for k in ref:
                ref[k].append(r[k].reshape(n, -1))
    ref = {k: torch.cat(v) for k, v in ref.items()}
    for k in ref:
        got = out[k].reshape(H * W, -1).cpu()[idx]
        err = rel_l2(got, ref[k])
        print("C4 256x256", k, err)
        assert err < 1e-3, (k, err)
    op = out["opacity"].reshape(-1)
    assert float(op.min()) >= 0.0 and float(op.max()) <= 1.0 + 1e-5 and float(op.max()) > 0.5
    blend = out["comp_rgb_fg"] + out["comp_rgb_bg"] * (1.0 - out["opacity"])
    assert (blend - out["comp_rgb"]).abs().max() < 1e-6
    dp = out["depth"].reshape(-1)
    assert float((dp - 4.0 * op).max()) <= 1e-4 and float((dp - 0.1 * op).min()) >= -1e-4
