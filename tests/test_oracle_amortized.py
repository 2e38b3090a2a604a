"""CPU checks of the amortized-path oracle (oracle/amortized_oracle.py): analytic invariants of the restated nerfacc
pieces (parity is unpinned for nerfacc, so the oracle is held to what can be proved) and its agreement
with torch autograd / closed forms."""
import math

import torch

from oracle import amortized_oracle as ao, render_oracle as ro


def _setup(B=2, seed=0, table_scale=0.05):
    hcfg, vcfg = ao.HyperCfg(), ao.VolSDFCfg()
    g = torch.Generator().manual_seed(seed)
    n = ro.grid_meta(hcfg.grid)["n_entries"]
    table = (torch.rand(n, 2, generator=g) * 2 - 1) * table_scale
    out_dims = {"sdf_weights": [32, 64, 1], "feature_weights": [32, 64, 3]}
    hp = ao.make_hypernet(hcfg.c_dim, hcfg.n_neurons, 32 * 64 * 2 + 64 + 192, seed + 1)
    emb = torch.randn(B, hcfg.c_dim, generator=g)
    cache = ao.hypernet_forward(hp, emb, out_dims)
    return hcfg, vcfg, table, cache, g


def test_hypernet_split_shapes():
    _, _, _, cache, _ = _setup()
    assert [tuple(m.shape) for m in cache["sdf_weights"]] == [(2, 32, 64), (2, 64, 1)]
    assert [tuple(m.shape) for m in cache["feature_weights"]] == [(2, 32, 64), (2, 64, 3)]


def test_volsdf_density_closed_form():
    s = torch.linspace(-0.3, 0.3, 101)
    a = 30.0
    ref = torch.where(s > 0, 0.5 * a * torch.exp(-s * a), a * (1 - 0.5 * torch.exp(s * a)))
    ref[50] = 0.5 * a  # sign(0) = 0
    torch.testing.assert_close(ao.volsdf_density(s, a), ref, atol=1e-5, rtol=1e-5)
    assert float(ao.volsdf_density(torch.tensor([1.0]), 1e3)) < 1e-10  # inv_std clamps at 80


def test_importance_sampling_uniform_and_monotone():
    Nr, n = 7, 128
    unit = torch.tensor([[0.0, 1.0]]).expand(Nr, 2)
    u = torch.rand(Nr, generator=torch.Generator().manual_seed(3))
    s = ao.importance_sampling(unit, unit, n, u)
    torch.testing.assert_close(s, (torch.arange(n + 1)[None] + u[:, None]) / (n + 1))
    cdf = torch.sort(torch.rand(Nr, 33), -1)[0]
    cdf[:, 0], cdf[:, -1] = 0.0, 1.0
    edges = torch.linspace(0, 1, 33)[None].expand(Nr, 33).contiguous()
    f = ao.importance_sampling(edges, cdf, 64, u)
    assert (f[:, 1:] >= f[:, :-1]).all() and (f >= 0).all() and (f <= 1).all()
    # a point mass in one bin pulls every fine edge into that bin
    cdf2 = torch.zeros(1, 33)
    cdf2[:, 17:] = 1.0
    f2 = ao.importance_sampling(edges[:1], cdf2, 16, torch.tensor([0.5]))
    assert (f2 >= edges[0, 16] - 1e-6).all() and (f2 <= edges[0, 17] + 1e-6).all()


def test_render_invariants_and_fd_normals():
    hcfg, vcfg, table, cache, g = _setup(B=2)
    HW = 12
    o = torch.tensor([[0.0, -1.6, 0.3]]).repeat(2 * HW, 1)
    d = torch.nn.functional.normalize(torch.tensor([[0.0, 1.0, -0.15]]) + 0.15 * torch.randn(2 * HW, 3, generator=g), dim=-1)
    bg = torch.rand(2 * HW, 3, generator=g)
    out = ao.render(o, d, HW, table, cache, bg, hcfg, vcfg, torch.rand(2 * HW, generator=g), torch.rand(2 * HW, generator=g))
    t = out["t"]
    assert t.shape == (2 * HW, vcfg.n_coarse + vcfg.n_fine + 2)
    assert (t[:, 1:] >= t[:, :-1]).all() and t.min() >= vcfg.near and t.max() <= vcfg.far
    w = out["weights"]
    alpha = out["t_intervals"].abs() * ao.volsdf_density(out["sdf"].view_as(w), vcfg.inv_std)
    T_final = torch.prod(1 - alpha, -1)
    torch.testing.assert_close(w.sum(-1) + T_final, torch.ones(2 * HW), atol=1e-5, rtol=0)
    torch.testing.assert_close(out["comp_rgb"], out["comp_rgb_fg"] + bg * (1 - out["opacity"][:, None]))
    assert out["opacity"].max() > 0.9  # the sphere-biased SDF is hit by the central rays
    # finite-difference sdf_grad agrees with autograd of the SDF (sphere bias dominates: |grad| ~ 1)
    # (at tcnn's 1e-4 table initialisation; a rough table makes the eps = 0.01 difference a poor derivative)
    p = (torch.rand(1, 50, 3, generator=g) * 2 - 1).requires_grad_(True)
    c1 = {k: [m[:1] for m in v] for k, v in cache.items()}
    table = table * (1e-4 / 0.05)
    sdf = ao.hyper_sdf(p, table, c1, hcfg)
    (ga,) = torch.autograd.grad(sdf.sum(), p)
    fd = ao.hyper_field(p.detach(), table, c1, hcfg, output_normal=True)["sdf_grad"].view(1, 50, 3)
    assert torch.nn.functional.cosine_similarity(fd, ga, dim=-1).min() > 0.95
    assert abs(float(ao.eikonal_loss(fd.view(-1, 3)))) < 0.5


def test_sample_from_planes_axes_and_borders():
    """Plane 0 reads (x, y), plane 1 (x, z), plane 2 (z, y); texel centres reproduce the texel; outside is zero."""
    C, H, W = 4, 8, 8
    planes = torch.zeros(1, 3, C, H, W)
    planes[0, 0, 0] = torch.arange(W)[None, :].float().expand(H, W)   # varies along grid-x (W)
    planes[0, 1, 0] = torch.arange(H)[:, None].float().expand(H, W)   # varies along grid-y (H)
    planes[0, 2, 0] = torch.arange(W)[None, :].float().expand(H, W)
    centre = lambda i, n: (2 * i + 1) / n - 1
    p = torch.tensor([[[centre(2, W), centre(5, H), centre(6, W)]]])
    enc = ao.sample_from_planes(planes, p)
    assert enc.shape == (1, 1, 12)
    assert abs(float(enc[0, 0, 0]) - 2.0) < 1e-5      # plane 0: u = x -> column 2
    assert abs(float(enc[0, 0, 4]) - 6.0) < 1e-5      # plane 1: v = z -> row 6
    assert abs(float(enc[0, 0, 8]) - 6.0) < 1e-5      # plane 2: u = z -> column 6
    assert float(ao.sample_from_planes(planes + 1.0, torch.tensor([[[3.0, 3.0, 3.0]]])).abs().max()) == 0.0


def test_adan_matches_closed_form_first_step():
    g = torch.tensor([0.5, -2.0])
    p = torch.tensor([1.0, 1.0])
    st = {}
    ao.adan_step(p, g, st, 1, 0.1, (0.98, 0.92, 0.99), 1e-15)
    # step 1: diff = 0, m = (1-b1) g, n = (1-b3) g^2 -> update = lr * sign(g) (bias corrections cancel)
    torch.testing.assert_close(p, torch.tensor([0.9, 1.1]), atol=1e-6, rtol=0)
