"""Host-side pieces against vectors produced by the reference's own definitions (tests/golden/make_host_golden.py:
threestudio/utils/misc.py C, utils/ops.py ray / projection helpers, prompt_processors/base.py Perp-Neg defaults)."""
import os

import torch

from oracle import render_oracle as ro

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_golden.pt"))


def test_scheduled_scalar_matches_reference_C():
    from scaledreamer_b200.core import C

    assert len(GOLD["C"]) > 100
    for c in GOLD["C"]:
        got = C(c["value"], c["epoch"], c["global_step"], c["interpolation"])
        assert abs(float(got) - c["out"]) <= 1e-12 * max(1.0, abs(c["out"])), c


def test_projection_and_mvp_match_reference():
    from scaledreamer_b200 import data as D

    r = GOLD["rays"]
    proj = D.get_projection_matrix(r["fovy"], r["W"] / r["H"], 0.1, 1000.0)
    torch.testing.assert_close(proj, r["proj_mtx"], atol=0, rtol=0)
    torch.testing.assert_close(D.get_mvp_matrix(r["c2w"], proj), r["mvp_mtx"], atol=1e-6, rtol=1e-6)


def test_oracle_rays_match_reference_get_rays():
    r = GOLD["rays"]
    rays_o, rays_d = ro.get_rays(r["c2w"], r["fovy"], r["H"], r["W"])
    torch.testing.assert_close(rays_o.reshape(r["rays_o"].shape), r["rays_o"], atol=0, rtol=0)
    torch.testing.assert_close(rays_d.reshape(r["rays_d"].shape), r["rays_d"], atol=2e-7, rtol=1e-6)


def test_perp_neg_defaults_match_reference_config():
    import scaledreamer_b200 as sd

    cfg = sd.find("stable-diffusion-prompt-processor").Config
    p = GOLD["prompt"]
    for k in ("perp_neg_f_sb", "perp_neg_f_fsb", "perp_neg_f_fs", "perp_neg_f_sf"):
        assert tuple(float(v) for v in getattr(cfg, k)) == tuple(float(v) for v in p[k]), k
    assert abs(p["decay_check"] - (1.0 * torch.exp(torch.tensor(-0.5 * 0.25)).item() - 0.606)) < 1e-6


def _replay_unet_out(g):
    B = g["t"].shape[0]
    gen = torch.Generator().manual_seed(g["unet_out_seed"])
    unet_out = torch.randn(5 * B, 4, 64, 64, generator=gen)
    torch.testing.assert_close(unet_out.flatten()[::97], g["unet_out_sample"], atol=0, rtol=0)
    return unet_out


def test_oracle_guidance_arithmetic_matches_reference_call():
    """oracle/ldm_oracle.py t_plus / alphas_cumprod / q-sample / asd_grad against the reference's own
    SDTimestepShiftedScoreDistillationGuidance.__call__ + get_t_plus + get_eps (stable_diffusion_asd_guidance.py:211-428)
    run with a recorded UNet output: the schedule, the timestep shift, the UNet input batch, the Perp-Neg CFG
    combination, w(t), the loss and its gradient."""
    from oracle import ldm_oracle as lo

    g = GOLD["guidance"]
    B = g["t"].shape[0]
    ac = lo.alphas_cumprod()
    torch.testing.assert_close(ac, g["alphas_cumprod"], atol=0, rtol=0)
    tp = lo.t_plus(g["t"], g["u"], g["plus_ratio"], g["min_step"])
    assert torch.equal(tp, g["t_plus"]) and (tp >= g["t"]).all() and (tp <= 999).all()
    assert torch.equal(g["unet_t"].long(), torch.cat([g["t"]] * 4 + [tp]))
    z = g["latents_bhwc"].permute(0, 3, 1, 2).clone().requires_grad_(True)
    qs = lambda tt: ac[tt].sqrt().view(-1, 1, 1, 1) * z + (1 - ac[tt]).sqrt().view(-1, 1, 1, 1) * g["noise"]
    x_in = torch.cat([qs(g["t"])] * 4 + [qs(tp)], 0)
    torch.testing.assert_close(x_in.detach().flatten()[::97], g["unet_x_sample"], atol=1e-6, rtol=1e-6)
    # Perp-Neg weights as get_eps scales them: prompt weights * -1 * guidance_perp_neg
    p = GOLD["prompt"]
    sel = [3, 8]
    torch.testing.assert_close(p["elevation"][sel], g["elevation"])
    neg_w = p["neg_weights"][sel] * -1 * g["guidance_perp_neg"]
    grad, loss, gnorm = lo.asd_grad(_replay_unet_out(g), z, g["t"], ac, B, g["guidance_scale"], neg_w)
    loss.backward()
    torch.testing.assert_close(loss.detach(), g["loss_asd"], atol=0, rtol=1e-5)
    torch.testing.assert_close(gnorm, g["grad_norm"], atol=0, rtol=1e-5)
    torch.testing.assert_close(z.grad.permute(0, 2, 3, 1), g["grad_bhwc"], atol=1e-6, rtol=1e-5)
    # context batch order of the UNet call: [vd, uncond, neg (2B, sample-major), vd]
    ctx = g["ctx"]
    assert ctx.shape[0] == 5 * B and torch.equal(ctx[:B], ctx[4 * B:])


def test_oracle_mvdream_guidance_matches_reference_call():
    """Same for the multi-view guidance (mvdream_asd_guidance.py:167-304): ONE timestep for the four views, plain CFG
    (no Perp-Neg), camera conditioning through extern/mvdream/camera_utils.py normalize_camera."""
    from oracle import ldm_oracle as lo
    from scaledreamer_b200.guidance import normalize_camera

    g = GOLD["mv_guidance"]
    B = 4
    assert g["t"].shape == (1,) and g["u"].shape == (1,) and g["num_frames"] == 4
    ac = lo.alphas_cumprod()
    t = g["t"].repeat(B)
    tp = lo.t_plus(g["t"], g["u"], g["plus_ratio"], g["min_step"]).repeat(B)
    assert torch.equal(g["unet_t"].long(), torch.cat([t, t, tp]))
    cam = normalize_camera(g["c2w"].clone())
    torch.testing.assert_close(cam.repeat(3, 1), g["camera"], atol=1e-6, rtol=1e-6)
    z = g["latents_bhwc"].permute(0, 3, 1, 2).clone().requires_grad_(True)
    qs = lambda tt: ac[tt].sqrt().view(-1, 1, 1, 1) * z + (1 - ac[tt]).sqrt().view(-1, 1, 1, 1) * g["noise"]
    x_in = torch.cat([qs(t), qs(t), qs(tp)], 0)
    torch.testing.assert_close(x_in.detach().flatten()[::53], g["unet_x_sample"], atol=1e-6, rtol=1e-6)
    gen = torch.Generator().manual_seed(g["unet_out_seed"])
    unet_out = torch.randn(3 * B, 4, 32, 32, generator=gen)
    torch.testing.assert_close(unet_out.flatten()[::53], g["unet_out_sample"], atol=0, rtol=0)
    grad, loss, gnorm = lo.asd_grad(unet_out, z, t, ac, B, g["guidance_scale"], None)
    loss.backward()
    torch.testing.assert_close(loss.detach(), g["loss_asd"], atol=0, rtol=1e-5)
    torch.testing.assert_close(gnorm, g["grad_norm"], atol=0, rtol=1e-5)
    torch.testing.assert_close(z.grad.permute(0, 2, 3, 1), g["grad_bhwc"], atol=1e-6, rtol=1e-5)
    # context: the global (not view-dependent) pair, each repeated over the four views: [cond x4, uncond x4, cond x4]
    p = GOLD["prompt"]
    torch.testing.assert_close(g["ctx"][:4], p["text_embeddings"].expand(4, -1, -1))
    torch.testing.assert_close(g["ctx"][4:8], p["uncond_text_embeddings"].expand(4, -1, -1))
    assert torch.equal(g["ctx"][:4], g["ctx"][8:])


def test_config_resolvers_match_reference():
    """The `${name:args}` resolvers of the yaml configs and C_max (threestudio/utils/config.py:10-49)."""
    from scaledreamer_b200 import core

    r = GOLD["resolvers"]
    table = core._resolver_table(1)
    assert set(r["names"]) <= set(table), set(r["names"]) - set(table)
    for name, args, ref in r["calls"]:
        got = table[name](*args)
        assert got == ref and type(got) is type(ref), (name, args, got, ref)
    for v, ref in r["c_max"]:
        assert core._cmax(v) == ref, (v, core._cmax(v), ref)
