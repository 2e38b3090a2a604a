"""GPU parity of the packed (unfused) renderer stages and the C1 vanilla-MLP implicit volume against the oracle and the
reference-generated golden vectors, through the C ABI / the threestudio plugin classes.

Tolerances: index work (candidate counts, ray indices) is exact; float outputs 1e-3 relative (north_star); the
visibility test is a hard threshold, so a sample whose alpha sits within rounding of it may flip (see
test_render_gpu.py) -- the per-sample comparisons below run with pruning off, the image comparisons keep the looser
max-abs bound."""
import os

import pytest
import torch

from oracle import render_oracle as ro
from tests.helpers import march_spec_from_oracle, rel_l2, scene

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _gold():
    return torch.load(os.path.join(HERE, "golden", "field_golden.pt"))


def _geometry(c, dev):
    """implicit-volume plugin configured like golden case c, carrying its weights."""
    import scaledreamer_b200 as sd

    geo = sd.find("implicit-volume")({
        "radius": c["radius"], "density_bias": c["density_bias"], "density_activation": c["density_activation"],
        "pos_encoding_config": {"otype": "ProgressiveBandFrequency", "n_frequencies": c["n_frequencies"],
                                "n_masking_step": c["n_masking_step"], "include_xyz": c["include_xyz"]},
        "mlp_network_config": {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none",
                               "n_neurons": 64, "n_hidden_layers": c["n_hidden_layers"]}}).to(dev)
    for net, ws in ((geo.density_network, c["density_weights"]), (geo.feature_network, c["feature_weights"])):
        for p, w in zip(net.weights(), ws):
            p.data.copy_(w)
    geo.do_update_step(0, c["global_step"])
    return geo


@pytest.mark.parametrize("case", ["f4", "f6_xyz_masked", "f12"])
def test_frequency_field_matches_reference_golden(cuda_device, case):
    """sdb_freq_encode + sdb_mlp3_forward + density bias / activation == the reference's own classes."""
    c = _gold()[case]
    geo = _geometry(c, cuda_device)
    assert not geo.fusable and geo.encoding.n_output_dims == c["enc"].shape[1]
    assert [tuple(k for k in geo.state_dict() if "network" in k)] == [(
        *[f"density_network.layers.{2 * i}.weight" for i in range(c["n_hidden_layers"] + 1)],
        *[f"feature_network.layers.{2 * i}.weight" for i in range(c["n_hidden_layers"] + 1)])]
    torch.testing.assert_close(geo.encoding.mask.cpu(), c["mask"], atol=1e-6, rtol=0)
    pts = c["points"].to(cuda_device)
    x01 = (pts + c["radius"]) / (2 * c["radius"])
    enc = geo.encoding(x01)
    assert enc.shape[1] % 8 == 0 and (enc[:, c["enc"].shape[1]:] == 0).all()
    torch.testing.assert_close(enc[:, : c["enc"].shape[1]].cpu(), c["enc"], atol=2e-6, rtol=1e-5)
    out = geo(pts)
    assert rel_l2(out["density"][:, 0].cpu(), c["density"]) < 1e-5
    assert rel_l2(out["features"].cpu(), c["features"]) < 1e-5
    assert rel_l2(geo.forward_density(pts)[:, 0].cpu(), c["density"]) < 1e-5
    # finite-difference normals against the reference's own ImplicitVolume.forward(output_normal=True)
    nrm = geo(pts, output_normal=True)["normal"].detach().cpu()
    cos = (nrm * c["normal"]).sum(-1)
    assert cos.min() > 0.99 and cos.median() > 0.9999, float(cos.min())


def _c1_scene(H, W, B=1, seed=3, prune=True, n_hidden=2, n_samples=256):
    fcfg = ro.FieldCfg(encoding="frequency", n_frequencies=6, n_hidden_layers=n_hidden)
    return scene(H=H, W=W, B=B, seed=seed, prune=prune, n_samples=n_samples, fcfg=fcfg)


def _occ(sc, dev):
    from scaledreamer_b200 import render_ops as R

    occ = R.OccGrid(sc["mcfg"].grid_res, dev)
    occ.set_binaries(sc["binary"], sc["occs"])
    return occ


@pytest.mark.parametrize("jitter", [True, False])
def test_march_packed_matches_oracle(cuda_device, jitter):
    """Candidate lattice samples: identical counts / ray indices, t within an ulp, sorted by (ray, t); rays that miss
    the box contribute nothing."""
    from scaledreamer_b200 import render_ops as R

    sc = _c1_scene(24, 40, B=2)
    jit = sc["jitter"] if jitter else None
    ri, ts, te, tm = ro.march_candidates(sc["rays_o"], sc["rays_d"], jit, sc["binary"].numpy(), sc["fcfg"], sc["mcfg"])
    smp = R.march_packed(march_spec_from_oracle(sc["mcfg"]), sc["fcfg"].radius, _occ(sc, cuda_device),
                         sc["rays_o"].to(cuda_device), sc["rays_d"].to(cuda_device),
                         jit.to(cuda_device) if jitter else None)
    assert smp["ray_indices"].numel() == ri.numel() > 0
    assert torch.equal(smp["ray_indices"].cpu().long(), ri)
    counts = torch.bincount(ri, minlength=sc["rays_o"].shape[0])
    assert torch.equal(smp["offsets"].cpu()[1:] - smp["offsets"].cpu()[:-1], counts)
    assert (counts == 0).any()
    torch.testing.assert_close(smp["t_starts"].cpu(), ts, atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(smp["t_ends"].cpu(), te, atol=1e-6, rtol=1e-6)
    pos = sc["rays_o"][ri] + sc["rays_d"][ri] * tm[:, None]
    torch.testing.assert_close(smp["positions"].cpu(), pos, atol=2e-6, rtol=1e-5)


def test_march_packed_empty_grid_and_no_rays(cuda_device):
    from scaledreamer_b200 import render_ops as R

    sc = _c1_scene(8, 8)
    occ = R.OccGrid(32, cuda_device)  # nothing occupied
    m = march_spec_from_oracle(sc["mcfg"])
    smp = R.march_packed(m, 1.0, occ, sc["rays_o"].to(cuda_device), sc["rays_d"].to(cuda_device), None)
    assert smp["ray_indices"].numel() == 0 and int(smp["offsets"][-1]) == 0
    out = R.composite_packed(torch.zeros(0, device=cuda_device), torch.zeros(0, 3, device=cuda_device), smp)
    assert (out["opacity"] == 0).all() and (out["comp_rgb_fg"] == 0).all() and (out["z_variance"] == 0).all()
    smp0 = R.march_packed(m, 1.0, occ, torch.zeros(0, 3, device=cuda_device), torch.zeros(0, 3, device=cuda_device), None)
    assert smp0["offsets"].numel() == 1


def test_packed_composite_forward_backward(cuda_device):
    """Ragged rays (0 .. 150 samples, i.e. several 32-sample chunks) against the oracle's packed arithmetic in fp64."""
    from scaledreamer_b200 import render_ops as R

    g = torch.Generator().manual_seed(0)
    counts = torch.tensor([0, 1, 31, 32, 33, 150, 0, 64, 7])
    n_rays, n = counts.numel(), int(counts.sum())
    ray_idx = torch.repeat_interleave(torch.arange(n_rays), counts)
    dt = 0.01
    ts = torch.cat([torch.arange(int(c)) * dt + 0.3 for c in counts]).float()
    te = ts + dt
    sigma = (torch.rand(n, generator=g) * 40).double().requires_grad_(True)
    rgb = torch.rand(n, 3, generator=g).double().requires_grad_(True)
    sd = sigma * (te - ts).double()
    T = torch.exp(-ro._excl_cumsum_packed(sd, ray_idx, n_rays))
    w = T * (1 - torch.exp(-sd))
    tm = 0.5 * (ts + te).double()
    acc = lambda v: torch.zeros(n_rays, v.shape[-1], dtype=torch.float64).index_add(0, ray_idx, v)
    op, dep, fg = acc(w[:, None]), acc((w * tm)[:, None]), acc(w[:, None] * rgb)
    wn = w / op.clamp(min=1e-5)[ray_idx, 0]
    zm = acc((wn * tm)[:, None])
    zv = acc((wn * (tm - zm[ray_idx, 0]) ** 2)[:, None]) * (op > 0.5)
    go, gd, gf = torch.randn(n_rays, 1, generator=g), torch.randn(n_rays, 1, generator=g), torch.randn(n_rays, 3, generator=g)
    ((op * go).sum() + (dep * gd).sum() + (fg * gf).sum()).backward(retain_graph=True)

    dev = cuda_device
    off = torch.zeros(n_rays + 1, dtype=torch.int64)
    off[1:] = counts.cumsum(0)
    smp = {"t_starts": ts.to(dev), "t_ends": te.to(dev), "offsets": off.to(dev)}
    sg = sigma.detach().float().to(dev).requires_grad_(True)
    cg = rgb.detach().float().to(dev).requires_grad_(True)
    out = R.composite_packed(sg, cg, smp)
    assert rel_l2(out["opacity"].cpu(), op[:, 0]) < 1e-5
    assert rel_l2(out["depth"].cpu(), dep[:, 0]) < 1e-5
    assert rel_l2(out["comp_rgb_fg"].cpu(), fg) < 1e-5
    assert rel_l2(out["weights"].cpu(), w.detach()) < 1e-5
    assert rel_l2(out["z_variance"].cpu(), zv[:, 0].detach()) < 1e-4
    ((out["opacity"] * go[:, 0].to(dev)).sum() + (out["depth"] * gd[:, 0].to(dev)).sum()
     + (out["comp_rgb_fg"] * gf.to(dev)).sum()).backward()
    assert rel_l2(sg.grad.cpu(), sigma.grad) < 1e-4
    assert rel_l2(cg.grad.cpu(), rgb.grad) < 1e-5
    # each gradient input on its own (NULL for the others)
    (d_only,) = torch.autograd.grad(R.composite_packed(sg, cg, smp)["depth"].sum(), sg)
    (ref_d,) = torch.autograd.grad(acc((T * (1 - torch.exp(-sd)) * tm)[:, None]).sum(), sigma)
    assert rel_l2(d_only.cpu(), ref_d) < 1e-4


def test_visibility_pruning_matches_oracle(cuda_device):
    from scaledreamer_b200 import render_ops as R

    sc = _c1_scene(32, 32)
    ri, ts, te, tm = ro.march_candidates(sc["rays_o"], sc["rays_d"], sc["jitter"], sc["binary"].numpy(), sc["fcfg"],
                                         sc["mcfg"])
    pos = sc["rays_o"][ri] + sc["rays_d"][ri] * tm[:, None]
    sig = ro.field_density(pos, sc["P"], sc["fcfg"])[0]
    n_rays = sc["rays_o"].shape[0]
    sd = sig * (te - ts)
    T = torch.exp(-ro._excl_cumsum_packed(sd, ri, n_rays))
    alpha = 1 - torch.exp(-sd)
    thre = min(sc["mcfg"].alpha_thre, float(sc["occs"].mean()))
    keep = (alpha >= thre) & (T >= sc["mcfg"].early_stop_eps)
    margin = ((alpha - thre).abs() < 1e-5) | ((T - sc["mcfg"].early_stop_eps).abs() < 1e-6)
    occ = _occ(sc, cuda_device)
    m = march_spec_from_oracle(sc["mcfg"])
    smp = R.march_packed(m, 1.0, occ, sc["rays_o"].to(cuda_device), sc["rays_d"].to(cuda_device),
                         sc["jitter"].to(cuda_device))
    kept = R.prune_packed(smp, sig.to(cuda_device), m, occ)
    assert 0 < keep.sum() < keep.numel()
    assert abs(kept["ray_indices"].numel() - int(keep.sum())) <= int(margin.sum())
    if int(margin.sum()) == 0:
        assert torch.equal(kept["ray_indices"].cpu().long(), ri[keep])
        torch.testing.assert_close(kept["t_starts"].cpu(), ts[keep], atol=1e-6, rtol=1e-6)
    cnt = kept["offsets"].cpu()[1:] - kept["offsets"].cpu()[:-1]
    assert torch.equal(cnt, torch.bincount(kept["ray_indices"].cpu().long(), minlength=n_rays))


def _renderer(sc, dev, n_hidden, n_samples, prune, requires_normal=False):
    """nerf-volume-renderer over the C1 geometry, carrying the oracle scene's parameters and occupancy grid."""
    import scaledreamer_b200 as sd

    f = sc["fcfg"]
    geo = sd.find("implicit-volume")({
        "radius": f.radius, "pos_encoding_config": {"otype": "ProgressiveBandFrequency", "n_frequencies": f.n_frequencies},
        "mlp_network_config": {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none",
                               "n_neurons": 64, "n_hidden_layers": n_hidden}}).to(dev)
    mat = sd.find("no-material")({"requires_normal": requires_normal}).to(dev)
    bg = sd.find("neural-environment-map-background")({"random_aug": False, "dir_encoding_config": {
        "otype": "HashGrid", "n_features_per_level": 2, "log2_hashmap_size": 19, "n_levels": 4, "base_resolution": 4,
        "per_level_scale": 4.0}}).to(dev)
    P = sc["P"]
    for net, keys in ((geo.density_network, ("w1d", "wmd", "w2d")), (geo.feature_network, ("w1f", "wmf", "w2f"))):
        for p, w in zip(net.weights(), [P[k] for k in keys if k in P]):
            p.data.copy_(w)
    bg.encoding.encoding.params.data.copy_(P["bg_table"].reshape(-1))
    for p, k in zip(bg.network.weights(), ("bg_w1", "bg_w2", "bg_w3")):
        p.data.copy_(P[k])
    rend = sd.find("nerf-volume-renderer")({"radius": f.radius, "num_samples_per_ray": n_samples,
                                            "grid_prune": prune}, geometry=geo, material=mat, background=bg).to(dev)
    if prune:
        rend._occ_grid(dev).set_binaries(sc["binary"], sc["occs"])
    return rend, geo, bg


@pytest.mark.parametrize("n_hidden,prune", [(2, True), (1, True), (2, False)])
def test_c1_renderer_matches_oracle(cuda_device, n_hidden, prune, monkeypatch):
    """The plugin-level C1 render (frequency encoding + VanillaMLP, packed stages) against oracle.render on the same
    rays / jitter / parameters: images, packed extras and the gradients of an image loss to every trainable tensor."""
    H = W = 24
    n_samples = 256 if prune else 64
    sc = _c1_scene(H, W, seed=5 + n_hidden, prune=prune, n_hidden=n_hidden, n_samples=n_samples)
    P = {k: v.clone().requires_grad_(True) for k, v in sc["P"].items()}
    occ_mean = float(sc["occs"].mean()) if prune else None
    ref = ro.render(sc["rays_o"], sc["rays_d"], sc["jitter"], None, sc["binary"].numpy(), occ_mean, P, sc["fcfg"],
                    sc["mcfg"], H * W)
    g = torch.Generator().manual_seed(1)
    go, gop, gdp = torch.randn(H * W, 3, generator=g), torch.randn(H * W, generator=g), torch.randn(H * W, generator=g)
    ((ref["comp_rgb"] * go).sum() + (ref["opacity"] * gop).sum() + (ref["depth"] * gdp).sum()).backward()

    dev = cuda_device
    rend, geo, bg = _renderer(sc, dev, n_hidden, n_samples, prune)
    rend.train()
    monkeypatch.setattr(torch, "rand", lambda n, device=None, **kw: sc["jitter"].to(device))  # the renderer's jitter draw
    out = rend(sc["rays_o"].view(1, H, W, 3).to(dev), sc["rays_d"].view(1, H, W, 3).to(dev), None)
    monkeypatch.undo()
    for k in ("comp_rgb", "comp_rgb_fg", "comp_rgb_bg", "opacity", "depth"):
        a, b = out[k].reshape(ref[k].shape).cpu(), ref[k].detach()
        assert rel_l2(a, b) < 1e-3, (k, rel_l2(a, b))
        assert (a - b).abs().max() < 2e-2, k
    assert rel_l2(out["z_variance"].reshape(-1).cpu(), ref["z_variance"]) < 5e-3
    if not prune:  # no threshold flips: the packed extras line up sample for sample
        assert torch.equal(out["ray_indices"].cpu(), ref["ray_indices"])
        assert rel_l2(out["weights"][:, 0].cpu(), ref["weights"].detach()) < 1e-4
        assert rel_l2(out["density"][:, 0].cpu(), ref["density"].detach()) < 1e-4
        assert out["t_points"].shape == out["t_intervals"].shape == out["weights"].shape
        assert out["points"].shape == out["t_dirs"].shape == (out["weights"].shape[0], 3)
    loss = ((out["comp_rgb"].reshape(-1, 3) * go.to(dev)).sum() + (out["opacity"].reshape(-1) * gop.to(dev)).sum()
            + (out["depth"].reshape(-1) * gdp.to(dev)).sum())
    loss.backward()
    pairs = [(geo.density_network.weights(), ("w1d", "wmd", "w2d")), (geo.feature_network.weights(), ("w1f", "wmf", "w2f")),
             (bg.network.weights(), ("bg_w1", "bg_w2", "bg_w3"))]
    for ws, keys in pairs:
        for w, k in zip(ws, [k for k in keys if k in P]):
            assert rel_l2(w.grad.cpu(), P[k].grad) < 2e-2, (k, rel_l2(w.grad.cpu(), P[k].grad))
    assert rel_l2(bg.encoding.encoding.params.grad.view(-1, 2).cpu(), P["bg_table"].grad) < 2e-2


def test_c1_normals_and_occupancy_refresh(cuda_device):
    """material.requires_normal: finite-difference normals of the packed field (implicit_volume.py:167-177) against the
    oracle, and the occupancy refresh from caller-evaluated densities."""
    sc = _c1_scene(8, 8, seed=9)
    dev = cuda_device
    rend, geo, _ = _renderer(sc, dev, 2, 256, True, requires_normal=True)
    pts = (torch.rand(2000, 3, generator=torch.Generator().manual_seed(3)) * 2 - 1) * 0.7
    ref = ro.field_forward(pts, sc["P"], sc["fcfg"], output_normal=True)
    out = geo(pts.to(dev), output_normal=True)
    cos = (out["normal"].cpu() * ref["normal"]).sum(-1)
    assert cos.median() > 0.999 and (cos > 0.99).float().mean() > 0.97
    assert out["normal"].requires_grad
    rend.train()
    o = rend(sc["rays_o"].view(1, 8, 8, 3).to(dev), sc["rays_d"].view(1, 8, 8, 3).to(dev), None)
    assert o["normal"].shape == o["points"].shape and o["shading_normal"].shape == o["points"].shape
    # occupancy refresh (warm-up branch: every cell) == oracle occ_grid_from_density with the same cell jitter
    rend.occ = None
    torch.manual_seed(0)
    real_rand = torch.rand
    try:
        torch.rand = lambda *a, **k: sc["cell_rand"].to(dev) if a[:2] == (32 ** 3, 3) else real_rand(*a, **k)
        rend.update_step(0, 0)
    finally:
        torch.rand = real_rand
    occ = rend.occ
    assert rel_l2(occ.occs.cpu(), sc["occs"].reshape(-1)) < 1e-4
    agree = (occ.binaries().cpu() == sc["binary"]).float().mean()
    assert agree > 0.9995
    assert abs(float(occ.mean) - float(sc["occs"].mean())) < 1e-6


def test_c1_training_step_end_to_end(cuda_device):
    """BASELINE configs[0]: single-prompt ASD-SD, vanilla-MLP implicit volume, 64x64x1 view -- from the reference-schema
    yaml through data module, packed renderer, SD guidance, losses (incl. the orientation loss through the FD normals),
    backward and AdamW."""
    import scaledreamer_b200 as sd
    from scaledreamer_b200.systems import Trainer

    torch.manual_seed(4)
    cfg = sd.load_config(os.path.join(HERE, "configs", "asd_sd_vanilla_mlp.yaml"),
                         cli_args=["system.prompt_processor.prompt=a DSLR photo of a hamburger", "trainer.max_steps=3",
                                   "trainer.log_every_n_steps=1"])
    dm = sd.find(cfg.data_type)(cfg.data)
    system = sd.find(cfg.system_type)(cfg.system)
    assert not system.geometry.fusable
    keys = set(system.state_dict())
    assert {"geometry.density_network.layers.0.weight", "geometry.density_network.layers.4.weight",
            "geometry.feature_network.layers.4.weight", "background.network.layers.4.weight"} <= keys
    assert not any(k.startswith("geometry.encoding") for k in keys)  # the frequency encoding is parameter-free
    before = {k: v.detach().clone() for k, v in system.state_dict().items() if v.numel() > 0}
    tr = Trainer(**cfg.trainer)
    tr.fit(system, dm)
    torch.cuda.synchronize()
    assert tr.global_step == 3
    last = tr.history[-1]
    assert all(k in last for k in ("train/loss_asd", "train/loss_orient", "train/loss_sparsity"))
    assert last["train/loss_asd"] > 0 and last["train/loss_asd"] == last["train/loss_asd"]
    after = system.state_dict()
    for k in ("geometry.density_network.layers.0.weight", "geometry.density_network.layers.2.weight",
              "geometry.feature_network.layers.4.weight"):
        assert (after[k] != before[k]).any(), k
