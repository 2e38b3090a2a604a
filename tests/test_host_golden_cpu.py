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
