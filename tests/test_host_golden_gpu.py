"""Device-side glue against vectors produced by the reference's own definitions (tests/golden/make_host_golden.py):
the ray generator (threestudio/utils/ops.py:183-269) and the view-dependent / Perp-Neg text-embedding selection
(threestudio/models/prompt_processors/base.py:53-167), both through the C ABI."""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_golden.pt"))


def test_raygen_matches_reference_get_rays(cuda_device):
    from scaledreamer_b200.data import rays_on_device

    r = GOLD["rays"]
    rays_o, rays_d, _, _ = rays_on_device(r["c2w"], r["fovy"], r["H"], r["W"], cuda_device)
    torch.testing.assert_close(rays_o.cpu(), r["rays_o"], atol=0, rtol=0)
    torch.testing.assert_close(rays_d.cpu(), r["rays_d"], atol=5e-7, rtol=1e-5)
    torch.testing.assert_close(rays_d.norm(dim=-1).cpu(), torch.ones(r["rays_d"].shape[:-1]), atol=1e-6, rtol=0)


def _output(dev):
    from scaledreamer_b200.prompts import PromptProcessorOutput

    p = GOLD["prompt"]
    cfg = types.SimpleNamespace(use_perp_neg=True, front_threshold=p["front_threshold"],
                                back_threshold=p["back_threshold"], overhead_threshold=p["overhead_threshold"],
                                **{k: p[k] for k in ("perp_neg_f_sb", "perp_neg_f_fsb", "perp_neg_f_fs", "perp_neg_f_sf")})
    h = lambda k: p[k].to(dev, torch.float16).contiguous()
    return p, PromptProcessorOutput(h("text_embeddings"), h("uncond_text_embeddings"), h("text_embeddings_vd"),
                                    h("uncond_text_embeddings_vd"), cfg, "p", ["a", "b", "c", "d"])


def test_view_dependent_embeddings_match_reference(cuda_device):
    """Direction index from (elevation, azimuth) with the 30 / 30 / 60 degree thresholds: side, front, back, overhead,
    incl. azimuths outside [-180, 180] and samples next to every threshold. The tables are fp16-exact, so selection is
    exact."""
    p, out = _output(cuda_device)
    el, az = p["elevation"].to(cuda_device), p["azimuth"].to(cuda_device)
    vd = out.get_text_embeddings(el, az, None, True)
    assert torch.equal(vd.float().cpu(), p["vd"])
    gl = out.get_text_embeddings(el, az, None, False)
    assert torch.equal(gl.float().cpu(), p["global"])


def test_perp_neg_embeddings_and_weights_match_reference(cuda_device):
    """Front-side / side-back interpolation of the positive embedding, the two negative embeddings per sample and the
    shifted-exponential-decay weights; the overhead branch returns the unconditional embedding twice with zero weights."""
    p, out = _output(cuda_device)
    el, az = p["elevation"].to(cuda_device), p["azimuth"].to(cuda_device)
    ctx, w = out.get_text_embeddings_perp_neg(el, az, None, True)
    B = el.shape[0]
    assert ctx.shape == p["perp_neg"].shape and w.shape == (B, 2)
    # positive block: an fp32 blend of two fp16 rows rounded to fp16 on store
    torch.testing.assert_close(ctx[:B].float().cpu(), p["perp_neg"][:B], atol=2e-3, rtol=1e-3)
    assert torch.equal(ctx[B:].float().cpu(), p["perp_neg"][B:])  # unconditional and negative rows are copies
    torch.testing.assert_close(w.cpu(), p["neg_weights"], atol=1e-6, rtol=1e-5)
    overhead = p["elevation"] > p["overhead_threshold"]
    assert overhead.any() and (w.cpu()[overhead] == 0).all()


def test_asd_glue_kernels_match_reference_call(cuda_device):
    """sdb_asd_t_plus / sdb_asd_prologue / sdb_asd_epilogue against the reference's own guidance __call__ (run with a
    recorded UNet output, tests/golden/make_host_golden.py): timestep shift, q-sample batch assembly, Perp-Neg CFG,
    w(t), loss, grad norm and the gradient that reaches the latents. The posterior is made a point mass (identity
    quant_conv for the mean, log-variance -30, zero posterior noise, scaling factor 1) so that z equals the golden
    latents."""
    from scaledreamer_b200 import lib as L

    lib = L.load()
    g = GOLD["guidance"]
    dev = cuda_device
    B, HW, R = g["t"].shape[0], 64 * 64, 4
    t = g["t"].to(dev, torch.int32)
    u = g["u"].to(dev, torch.float32)
    tp = torch.empty(B, dtype=torch.int32, device=dev)
    L.check(lib.sdb_asd_t_plus(L.ptr(t), L.ptr(u), B, g["plus_ratio"], g["min_step"], 1000, L.ptr(tp), L.stream_ptr()), "t_plus")
    assert torch.equal(tp.cpu().long(), g["t_plus"])

    h = torch.cat([g["latents_bhwc"].reshape(B, HW, 4), torch.zeros(B, HW, 4)], -1).to(dev).contiguous()
    qw = torch.zeros(8, 8)
    qw[:4, :4] = torch.eye(4)
    qb = torch.cat([torch.zeros(4), torch.full((4,), -30.0)])
    qw, qb = qw.to(dev), qb.to(dev)
    eps_post = torch.zeros(B, HW, 4, device=dev)
    noise = g["noise"].permute(0, 2, 3, 1).reshape(B, HW, 4).to(dev).contiguous()
    ac = g["alphas_cumprod"].to(dev)
    latents = torch.empty(B, HW, 4, device=dev)
    unet_x = torch.empty((R + 1) * B, HW, 4, device=dev, dtype=torch.float16)
    unet_t = torch.empty((R + 1) * B, device=dev)
    L.check(lib.sdb_asd_prologue(L.ptr(h), L.ptr(qw), L.ptr(qb), L.ptr(eps_post), L.ptr(noise), L.ptr(t), L.ptr(tp),
                                 L.ptr(ac), 1.0, B, HW, R, L.ptr(latents), L.ptr(unet_x), L.ptr(unet_t), L.stream_ptr()),
            "prologue")
    torch.testing.assert_close(latents.cpu(), g["latents_bhwc"].reshape(B, HW, 4), atol=1e-6, rtol=0)
    assert torch.equal(unet_t.cpu(), g["unet_t"].float())
    x_nchw = unet_x.view((R + 1) * B, 64, 64, 4).permute(0, 3, 1, 2).float().cpu()
    torch.testing.assert_close(x_nchw.flatten()[::97], g["unet_x_sample"], atol=4e-3, rtol=2e-3)  # fp16 UNet input

    gen = torch.Generator().manual_seed(g["unet_out_seed"])
    unet_out = torch.randn(5 * B, 4, 64, 64, generator=gen)
    torch.testing.assert_close(unet_out.flatten()[::97], g["unet_out_sample"], atol=0, rtol=0)
    eps = unet_out.permute(0, 2, 3, 1).reshape(5 * B, HW, 4).to(dev).contiguous()
    neg_w = (GOLD["prompt"]["neg_weights"][[3, 8]] * -1 * g["guidance_perp_neg"]).to(dev).contiguous()
    grad, d_h = torch.empty(B, HW, 4, device=dev), torch.empty(B, HW, 8, device=dev)
    loss, gnorm = torch.empty(1, device=dev), torch.empty(1, device=dev)
    L.check(lib.sdb_asd_epilogue(L.ptr(eps), L.ptr(h), L.ptr(qw), L.ptr(qb), L.ptr(eps_post), L.ptr(t), L.ptr(ac),
                                 L.ptr(neg_w), g["guidance_scale"], 0, 0.0, 1.0, 1.0, B, HW, R, L.ptr(grad), L.ptr(d_h),
                                 L.ptr(loss), L.ptr(gnorm), L.stream_ptr()), "epilogue")
    torch.testing.assert_close(loss.cpu()[0], g["loss_asd"], atol=0, rtol=2e-5)
    torch.testing.assert_close(gnorm.cpu()[0], g["grad_norm"], atol=0, rtol=2e-5)
    # d loss / d latents = grad / B reaches h through the identity rows of quant_conv
    ref = g["grad_bhwc"].reshape(B, HW, 4)
    torch.testing.assert_close(d_h[..., :4].cpu(), ref, atol=2e-6, rtol=2e-5)
    torch.testing.assert_close(grad.cpu() / B, ref, atol=2e-6, rtol=2e-5)


def test_asd_glue_kernels_match_reference_mvdream_call(cuda_device):
    """The same kernels in the multi-view arrangement (mvdream_asd_guidance.py:167-304): one shared timestep, UNet batch
    [x_t, x_t, x_{t+}] (num_repeats 2), plain CFG (neg_weights NULL), 32x32 latents."""
    from scaledreamer_b200 import lib as L

    lib = L.load()
    g = GOLD["mv_guidance"]
    dev = cuda_device
    B, HW, R = 4, 32 * 32, 2
    t = g["t"].repeat(B).to(dev, torch.int32)
    u = g["u"].repeat(B).to(dev, torch.float32)
    tp = torch.empty(B, dtype=torch.int32, device=dev)
    L.check(lib.sdb_asd_t_plus(L.ptr(t), L.ptr(u), B, g["plus_ratio"], g["min_step"], 1000, L.ptr(tp), L.stream_ptr()), "t_plus")
    assert torch.equal(torch.cat([t, t, tp]).cpu().long(), g["unet_t"].long())
    h = torch.cat([g["latents_bhwc"].reshape(B, HW, 4), torch.zeros(B, HW, 4)], -1).to(dev).contiguous()
    qw = torch.zeros(8, 8)
    qw[:4, :4] = torch.eye(4)
    qw, qb = qw.to(dev), torch.cat([torch.zeros(4), torch.full((4,), -30.0)]).to(dev)
    eps_post = torch.zeros(B, HW, 4, device=dev)
    noise = g["noise"].permute(0, 2, 3, 1).reshape(B, HW, 4).to(dev).contiguous()
    ac = g["alphas_cumprod"].to(dev)
    latents = torch.empty(B, HW, 4, device=dev)
    unet_x = torch.empty((R + 1) * B, HW, 4, device=dev, dtype=torch.float16)
    unet_t = torch.empty((R + 1) * B, device=dev)
    L.check(lib.sdb_asd_prologue(L.ptr(h), L.ptr(qw), L.ptr(qb), L.ptr(eps_post), L.ptr(noise), L.ptr(t), L.ptr(tp),
                                 L.ptr(ac), 1.0, B, HW, R, L.ptr(latents), L.ptr(unet_x), L.ptr(unet_t), L.stream_ptr()),
            "prologue")
    assert torch.equal(unet_t.cpu(), g["unet_t"].float())
    x_nchw = unet_x.view((R + 1) * B, 32, 32, 4).permute(0, 3, 1, 2).float().cpu()
    torch.testing.assert_close(x_nchw.flatten()[::53], g["unet_x_sample"], atol=4e-3, rtol=2e-3)
    gen = torch.Generator().manual_seed(g["unet_out_seed"])
    unet_out = torch.randn(3 * B, 4, 32, 32, generator=gen)
    eps = unet_out.permute(0, 2, 3, 1).reshape(3 * B, HW, 4).to(dev).contiguous()
    grad, d_h = torch.empty(B, HW, 4, device=dev), torch.empty(B, HW, 8, device=dev)
    loss, gnorm = torch.empty(1, device=dev), torch.empty(1, device=dev)
    L.check(lib.sdb_asd_epilogue(L.ptr(eps), L.ptr(h), L.ptr(qw), L.ptr(qb), L.ptr(eps_post), L.ptr(t), L.ptr(ac),
                                 None, g["guidance_scale"], 0, 0.0, 1.0, 1.0, B, HW, R, L.ptr(grad), L.ptr(d_h),
                                 L.ptr(loss), L.ptr(gnorm), L.stream_ptr()), "epilogue")
    torch.testing.assert_close(loss.cpu()[0], g["loss_asd"], atol=0, rtol=2e-5)
    torch.testing.assert_close(gnorm.cpu()[0], g["grad_norm"], atol=0, rtol=2e-5)
    torch.testing.assert_close(d_h[..., :4].cpu(), g["grad_bhwc"].reshape(B, HW, 4), atol=2e-6, rtol=2e-5)
    # camera conditioning of the plugin (normalize_camera, flattened 4x4 per view, repeated for the three UNet blocks)
    from scaledreamer_b200.guidance import normalize_camera

    cam = normalize_camera(g["c2w"].to(dev))
    torch.testing.assert_close(cam.repeat(3, 1).cpu(), g["camera"], atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("case,plugin", [("c2", "random-camera-datamodule"), ("generic", "random-camera-datamodule"),
                                          ("mv", "mvdream-random-multiview-camera-datamodule"),
                                          ("mv_zoom", "mvdream-random-multiview-camera-datamodule")])
def test_device_rays_of_training_batches_match_reference_datasets(cuda_device, case, plugin):
    """collate (host scalars, bit-identical to the reference under the same seeds: tests/test_data_cpu.py) + to_device
    (sdb_raygen) against the rays the reference's own data sets put into the same batches (tests/golden/data_golden.pt)."""
    import random

    import scaledreamer_b200 as sd

    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_golden.pt"))[case]
    dm = sd.find(plugin)(dict(g["config"]))
    dm.setup("fit")
    ds = dm.train_dataset
    for rec in g["batches"]:
        ds.update_step(0, rec["step"])
        random.seed(rec["seed"])
        torch.manual_seed(rec["seed"])
        b = ds.to_device(ds.collate({}), cuda_device)
        ref = rec["batch"]
        assert b["rays_o"].shape == (ds.batch_size, ref["height"], ref["width"], 3)
        torch.testing.assert_close(b["rays_o"].reshape(-1, 3)[::37].cpu(), ref["rays_o_sample"], atol=0, rtol=0)
        torch.testing.assert_close(b["rays_d"].reshape(-1, 3)[::37].cpu(), ref["rays_d_sample"], atol=1e-6, rtol=1e-5)


def test_multi_prompt_embeddings_match_reference(cuda_device):
    """sdb_asd_text_embeddings_multi through the plugin's MultiPromptProcessorOutput against the reference's own
    MultiPromptProcessorOutput (custom/amortized/models/prompt_processors/base.py:409-568): every sample of the batch
    has its own prompt (index into the stacked table), view-dependent selection, local / global embeddings, Perp-Neg
    interpolation and weights (default and explicit guidance_scale_neg)."""
    from scaledreamer_b200.amortized import MultiPromptProcessorOutput

    m, p = GOLD["multi_prompt"], GOLD["prompt"]
    dev = cuda_device
    h = lambda t: t.to(dev, torch.float16).contiguous()
    cfg = types.SimpleNamespace(use_perp_neg=True, use_local_text_embeddings=False, front_threshold=p["front_threshold"],
                                back_threshold=p["back_threshold"], overhead_threshold=p["overhead_threshold"],
                                **{k: p[k] for k in ("perp_neg_f_sb", "perp_neg_f_fsb", "perp_neg_f_fs", "perp_neg_f_sf")})
    proc = types.SimpleNamespace(cfg=cfg, vd_table=h(m["vd_table"]), local_table=h(m["local_table"][:, None]),
                                 uncond=h(m["uncond"][None]), uncond_vd=h(m["uncond_vd"]),
                                 global_table=m["global_table"].to(dev))
    out = MultiPromptProcessorOutput(proc, m["prompt_idx"].to(dev, torch.int32), ["p"] * 6)
    el, az = m["elevation"].to(dev), m["azimuth"].to(dev)
    B = el.shape[0]
    assert torch.equal(out.get_text_embeddings(el, az, None, True).float().cpu(), m["vd"])
    assert torch.equal(out.get_text_embeddings(el, az, None, False).float().cpu(), m["local"])
    torch.testing.assert_close(out.get_global_text_embeddings().cpu(), m["global"], atol=0, rtol=0)
    ctx, w = out.get_text_embeddings_perp_neg(el, az, None, True)
    torch.testing.assert_close(ctx[:B].float().cpu(), m["perp_neg"][:B], atol=2e-3, rtol=1e-3)
    assert torch.equal(ctx[B:].float().cpu(), m["perp_neg"][B:])
    torch.testing.assert_close(w.cpu(), m["neg_weights"], atol=1e-6, rtol=1e-5)
    _, w2 = out.get_text_embeddings_perp_neg(el, az, None, True, 0.5)
    torch.testing.assert_close(w2.cpu(), m["perp_neg_scaled"], atol=1e-6, rtol=1e-5)
