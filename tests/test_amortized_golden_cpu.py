"""Amortized-path oracle restatements and host modules against vectors produced by the reference's own definitions
(tests/golden/make_amortized_golden.py): Adan, VolSDF density, BCE, LinearHyperNetwork, triplane lookup."""
import os

import torch

from oracle import amortized_oracle as ao

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "amortized_golden.pt"))


def test_adan_oracle_matches_reference_optimizer():
    """Six steps of threestudio/systems/optimizers.py Adan (single-tensor path), proximal and decoupled weight decay."""
    for tag in ("prox", "no_prox", "plain"):
        c = GOLD[f"adan_{tag}"]
        p, state = c["p0"].clone(), {}
        for i in range(6):
            ao.adan_step(p, c["grads"][i], state, i + 1, c["lr"], c["betas"], c["eps"], c["weight_decay"], c["no_prox"])
            torch.testing.assert_close(p, c["params"][i], atol=1e-6, rtol=1e-5)


def test_volsdf_density_and_bce_match_reference():
    c = GOLD["volsdf_density"]
    for inv, ref in zip(c["inv_std"], c["out"]):
        torch.testing.assert_close(ao.volsdf_density(c["sdf"], float(inv)), ref, atol=1e-6, rtol=1e-6)
    from scaledreamer_b200.systems import binary_cross_entropy

    torch.testing.assert_close(binary_cross_entropy(GOLD["bce"]["x"], GOLD["bce"]["x"]), GOLD["bce"]["out"])


def test_hypernetwork_matches_reference_module():
    """Same state-dict keys, same split of the flat output into per-head [in, out] matrices -- for the oracle function
    and for the plugin's LinearHyperNetwork."""
    from scaledreamer_b200.amortized import LinearHyperNetwork

    c = GOLD["hypernet"]
    net = LinearHyperNetwork(c["n_input_dims"], c["config"])
    assert set(net.state_dict()) == set(c["state_dict"]) and net.n_output_dims == c["n_output_dims"]
    net.load_state_dict(c["state_dict"])
    with torch.no_grad():
        mine = net(c["c"])
    out_dims = {k: [c["n_input_dims"]] + v for k, v in c["config"]["out_dims"].items()}
    orc = ao.hypernet_forward(c["state_dict"], c["c"], out_dims)
    for name, mats in c["out"].items():
        assert len(mine[name]) == len(orc[name]) == len(mats) == 2
        for a, b, ref in zip(mine[name], orc[name], mats):
            torch.testing.assert_close(a, ref, atol=1e-6, rtol=1e-5)
            torch.testing.assert_close(b, ref, atol=1e-6, rtol=1e-5)


def test_triplane_lookup_oracle_matches_reference():
    c = GOLD["triplane"]
    torch.testing.assert_close(ao.sample_from_planes(c["planes"], c["points"]), c["out"], atol=1e-6, rtol=1e-5)
    # plane 0 samples (x, y), plane 1 (x, z), plane 2 (z, y)
    p, pts = c["proj"].view(2, 3, 5, 2), c["points"][:, :5]
    torch.testing.assert_close(p[:, 0], pts[..., [0, 1]])
    torch.testing.assert_close(p[:, 1], pts[..., [0, 2]])
    torch.testing.assert_close(p[:, 2], pts[..., [2, 1]])


def test_triplane_field_oracle_matches_reference_forward():
    """ao.triplane_field against the reference's own TriplaneTransformerSDF.forward(output_normal=True): contraction to
    [-1, 1], plane lookup, VanillaMLP heads, sphere sdf bias, finite-difference sdf_grad with clamped offsets, normals."""
    c = GOLD["triplane_geometry"]
    out = ao.triplane_field(c["points"], c["space_cache"], c["sdf_weights"], c["feature_weights"], c["radius"],
                            c["sdf_bias_radius"], c["fd_eps"], output_normal=True)
    ref = c["out"]
    torch.testing.assert_close(out["sdf"], ref["sdf"], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(out["features"], ref["features"], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(out["sdf_grad"], ref["sdf_grad"], atol=2e-4, rtol=1e-3)  # differences divided by 0.01
    cos = (out["normal"] * ref["normal"]).sum(-1)
    assert cos.min() > 0.9999
    assert torch.equal(ref["normal"], ref["shading_normal"])


def test_hyper_field_oracle_matches_reference_forward():
    """ao.hyper_field against the reference's own Hypernet_Sdf.forward(output_normal=True) (per-prompt bmm chain, sphere
    bias, clamped finite differences, [B*N, .] layout). The encoding inside that golden is the oracle's own hash grid
    (tiny-cuda-nn is not installable), so this pins everything around the encoding, not the encoding."""
    from oracle import render_oracle as ro

    c = GOLD["hyper_geometry"]
    cfg = ao.HyperCfg(grid=ro.GridCfg(**c["grid"]), radius=c["radius"], sdf_bias_radius=c["sdf_bias_radius"], fd_eps=c["fd_eps"])
    out = ao.hyper_field(c["points"], c["table"], c["cache"], cfg, output_normal=True)
    ref = c["out"]
    torch.testing.assert_close(out["sdf"], ref["sdf"], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(out["features"], ref["features"], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(out["sdf_grad"], ref["sdf_grad"], atol=2e-4, rtol=1e-3)
    assert ((out["normal"] * ref["normal"]).sum(-1)).min() > 0.9999
