"""CPU tests of the render oracle: analytic invariants and an independent scalar restatement."""
import math

import numpy as np
import torch

from oracle import render_oracle as ro
from tests.helpers import scene


def test_grid_geometry_matches_published_numbers():
    # levels 0-4 dense (res 16,24,34,49,71), 5-15 hashed, 6 299 960 entries (50.4 MB fp32 x2)
    m = ro.grid_meta(ro.GridCfg())
    assert m["res"][:6] == [16, 24, 34, 49, 71, 102]
    assert m["hashed"] == [False] * 5 + [True] * 11
    assert m["n_entries"] == 6299960
    assert m["res"][-1] == 4096
    mb = ro.grid_meta(ro.GridCfg(4, 2, 19, 4, 4.0))
    assert mb["res"] == [4, 16, 64, 256] and mb["n_entries"] == 64 + 4096 + 262144 + 524288


def _encode_scalar(x, table, cfg):
    """Independent pure-python restatement of the tcnn grid lookup for one point."""
    m = ro.grid_meta(cfg)
    out = []
    for l in range(cfg.n_levels):
        s = np.float32(m["scale"][l])
        pos = [np.float32(np.float64(np.float32(v)) * np.float64(s) + 0.5) for v in x]
        g = [int(math.floor(p)) for p in pos]
        w = [np.float32(p - np.float32(gi)) for p, gi in zip(pos, g)]
        acc = np.zeros(2, np.float64)
        for c in range(8):
            b = [(c >> d) & 1 for d in range(3)]
            cc = [g[d] + b[d] for d in range(3)]
            if m["hashed"][l]:
                idx = ((cc[0] * 1) & 0xFFFFFFFF) ^ ((cc[1] * 2654435761) & 0xFFFFFFFF) ^ ((cc[2] * 805459861) & 0xFFFFFFFF)
            else:
                idx = cc[0] + cc[1] * m["res"][l] + cc[2] * m["res"][l] ** 2
            idx %= m["size"][l]
            wt = 1.0
            for d in range(3):
                wt *= float(w[d]) if b[d] else 1.0 - float(w[d])
            acc += wt * table[m["offset"][l] + idx].double().numpy()
        out += list(acc)
    return np.array(out)


def test_hashgrid_vectorised_matches_scalar():
    cfg = ro.GridCfg()
    n = ro.grid_meta(cfg)["n_entries"]
    table = torch.randn(n, 2, generator=torch.Generator().manual_seed(0))
    x = torch.rand(7, 3, generator=torch.Generator().manual_seed(1))
    x[0] = torch.tensor([0.0, 0.5, 1.0])  # box faces
    enc = ro.hashgrid_encode(x, table, cfg)
    for i in range(x.shape[0]):
        ref = _encode_scalar(x[i].tolist(), table, cfg)
        np.testing.assert_allclose(enc[i].numpy(), ref, rtol=2e-4, atol=2e-5)


def test_render_invariants():
    sc = scene(H=16, W=16, B=2, seed=3)
    out = ro.render(sc["rays_o"], sc["rays_d"], sc["jitter"], None, sc["binary"].numpy(), float(sc["occs"].mean()),
                    sc["P"], sc["fcfg"], sc["mcfg"], sc["H"] * sc["W"])
    # comp = fg + bg (1 - opacity)
    torch.testing.assert_close(out["comp_rgb"], out["comp_rgb_fg"] + out["comp_rgb_bg"] * (1 - out["opacity"][:, None]))
    # sum_i w_i + T_final = 1 per ray
    n_rays = sc["rays_o"].shape[0]
    sd = out["density"] * (out["t_ends"] - out["t_starts"])
    tot = torch.zeros(n_rays).index_add(0, out["ray_indices"], sd)
    torch.testing.assert_close(out["opacity"] + torch.exp(-tot), torch.ones(n_rays), atol=1e-4, rtol=0)
    assert out["opacity"].max() > 0.5, "scene must contain an opaque blob"
    assert (out["opacity"] < 1e-3).any(), "and empty rays"
    # samples sorted along each ray, all inside the box
    ts, ri = out["t_starts"], out["ray_indices"]
    same = ri[1:] == ri[:-1]
    assert (ts[1:][same] > ts[:-1][same]).all()
    tm = 0.5 * (out["t_starts"] + out["t_ends"])
    pos = sc["rays_o"][ri] + sc["rays_d"][ri] * tm[:, None]
    assert pos.abs().max() <= sc["fcfg"].radius + 1e-4


def test_no_prune_keeps_every_lattice_sample():
    sc = scene(H=8, W=8, seed=4, prune=False, n_samples=64)
    out = ro.render(sc["rays_o"], sc["rays_d"], None, None, sc["binary"].numpy(), None, sc["P"], sc["fcfg"],
                    sc["mcfg"], 64)
    # every ray that crosses the box carries (t1 - t0) / step samples, at most n_samples + 1
    counts = torch.bincount(out["ray_indices"], minlength=64)
    assert counts.max() <= 65 and counts.max() >= 20


def test_get_rays_unit_and_centered():
    c2w = ro.look_at_c2w(torch.tensor([15.0]), torch.tensor([30.0]), torch.tensor([1.2]))
    o, d = ro.get_rays(c2w, torch.deg2rad(torch.tensor([50.0])), 8, 8)
    torch.testing.assert_close(d.norm(dim=-1), torch.ones(1, 8, 8))
    centre = torch.nn.functional.normalize(d[0, 3:5, 3:5].mean((0, 1)), dim=0)
    torch.testing.assert_close(centre, torch.nn.functional.normalize(-o[0, 0, 0], dim=0), atol=1e-5, rtol=0)


def test_frequency_field_matches_reference_golden():
    """C1 field: the oracle's ProgressiveBandFrequency / VanillaMLP / density bias + activation restatement against
    vectors produced by the reference's own class definitions (tests/golden/make_field_golden.py)."""
    import os

    import torch

    from oracle import render_oracle as ro

    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "field_golden.pt"))
    assert set(gold) == {"f4", "f6_xyz_masked", "f12"}
    for name, c in gold.items():
        mask = ro.freq_mask(c["n_frequencies"], c["n_masking_step"], c["global_step"])
        torch.testing.assert_close(mask, c["mask"], atol=1e-6, rtol=0)
        bias = c["density_bias"]
        fcfg = ro.FieldCfg(radius=c["radius"], encoding="frequency", n_frequencies=c["n_frequencies"],
                           include_xyz=c["include_xyz"], n_hidden_layers=c["n_hidden_layers"],
                           density_bias=bias if isinstance(bias, str) else "const",
                           density_bias_const=0.0 if isinstance(bias, str) else bias,
                           density_activation=c["density_activation"])
        P = {"freq_mask": mask}
        for head, ws in (("d", c["density_weights"]), ("f", c["feature_weights"])):
            P["w1" + head], P["w2" + head] = ws[0], ws[-1]
            if len(ws) == 3:
                P["wm" + head] = ws[1]
        x01 = (c["points"] + c["radius"]) / (2 * c["radius"])
        enc = ro.freq_encode(x01, c["n_frequencies"], mask, c["include_xyz"])
        torch.testing.assert_close(enc, c["enc"], atol=1e-6, rtol=1e-6)
        fcfg.fd_eps = c["fd_eps"]
        out = ro.field_forward(c["points"], P, fcfg, output_normal=True)
        torch.testing.assert_close(out["density"], c["density"], atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(out["features"], c["features"], atol=1e-5, rtol=1e-5)
        # finite-difference normals of the reference's own ImplicitVolume.forward (implicit_volume.py:167-177): offsets
        # clamped to the box, -(sigma(x + eps e_k) - sigma(x)) / eps, normalised. fp32 differences of nearly equal
        # densities: compared by direction
        cos = (out["normal"] * c["normal"]).sum(-1)
        assert cos.min() > 0.999 and cos.median() > 0.99999, (name, float(cos.min()))
