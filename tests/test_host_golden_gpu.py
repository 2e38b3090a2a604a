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
