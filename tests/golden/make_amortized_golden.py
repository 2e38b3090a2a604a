"""Golden vectors for the amortized (multi-prompt) pieces, produced by the REFERENCE's own definitions.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_amortized_golden.py
    threestudio/systems/optimizers.py                        Adan (the whole file is plain torch: executed as is)
    threestudio/models/renderers/neus_volume_renderer.py     volsdf_density
    threestudio/utils/ops.py                                 binary_cross_entropy, get_activation
    custom/amortized/models/geometry/hyper_iNGP.py           LinearHyperNetwork
    custom/amortized/models/geometry/utils.py                planes, project_onto_planes, sample_from_planes,
                                                             contract_to_unisphere_custom
    custom/amortized/models/geometry/triplane_transformer.py TriplaneTransformerSDF.forward / forward_sdf /
                                                             interpolate_encodings / get_shifted_sdf (+ networks.py VanillaMLP)
    custom/amortized/models/geometry/hyper_iNGP.py           Hypernet_Sdf.forward / forward_sdf / hypernet_forward /
                                                             get_shifted_sdf (encoding = the oracle's hash grid: tcnn is absent)
(definitions taken out by name with `ast` where the module itself cannot be imported).
Output: tests/golden/amortized_golden.pt (about 100 kB).
"""
import ast
import math
import os
import runpy

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "amortized_golden.pt")


class _Any:
    def __getitem__(self, item):
        return self


def pieces(path, names, ns):
    """exec the named top-level functions / classes / assignments of a reference file into ns, unchanged."""
    s = open(path).read()
    found = set()
    for node in ast.parse(s).body:
        name = None
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            name = node.name
        elif isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
        if name in names:
            exec(compile(ast.get_source_segment(s, node), path, "exec"), ns)
            found.add(name)
    assert found == set(names), set(names) - found


def dict_update(d, **kw):
    d.update(kw)
    return d


def main():
    gold = {}
    g = torch.Generator().manual_seed(5)
    ns = {"torch": torch, "nn": nn, "F": F, "math": math, "Float": _Any(), "Tensor": torch.Tensor, "Callable": object,
          "ListConfig": list, "trunc_exp": None}

    # ---- Adan, six steps, both weight-decay forms (optimizers.py:23-315; C5 uses betas (0.98, 0.92, 0.99))
    adan = runpy.run_path(f"{ROOT}/threestudio/systems/optimizers.py")["Adan"]
    for tag, kw in (("prox", dict(weight_decay=0.02, no_prox=False)), ("no_prox", dict(weight_decay=0.02, no_prox=True)),
                    ("plain", dict(weight_decay=0.0, no_prox=False))):
        p = nn.Parameter(torch.randn(257, generator=g))
        p0 = p.detach().clone()
        opt = adan([p], lr=1e-2, betas=(0.98, 0.92, 0.99), eps=1e-8, max_grad_norm=0.0, foreach=False, **kw)
        grads, traj = [], []
        for _ in range(6):
            gr = torch.randn(257, generator=g)
            p.grad = gr.clone()
            opt.step()
            grads.append(gr)
            traj.append(p.detach().clone())
        gold[f"adan_{tag}"] = {"p0": p0, "grads": torch.stack(grads), "params": torch.stack(traj), "lr": 1e-2,
                               "betas": (0.98, 0.92, 0.99), "eps": 1e-8, **kw}

    # ---- VolSDF density (neus_volume_renderer.py:19-23) and the numerically plain BCE (ops.py:365-369)
    pieces(f"{ROOT}/threestudio/models/renderers/neus_volume_renderer.py", ["volsdf_density"], ns)
    sdf = torch.cat([torch.linspace(-0.5, 0.5, 41), torch.tensor([0.0, 1e-6, -1e-6, 2.0, -2.0])])
    inv = torch.tensor([0.5, 10.0, 29.96, 80.0, 200.0, -1.0])
    gold["volsdf_density"] = {"sdf": sdf, "inv_std": inv,
                              "out": torch.stack([ns["volsdf_density"](sdf, i) for i in inv])}
    pieces(f"{ROOT}/threestudio/utils/ops.py", ["binary_cross_entropy", "get_activation"], ns)
    x = torch.rand(64, generator=g).clamp(1e-3, 1 - 1e-3)
    gold["bce"] = {"x": x, "out": ns["binary_cross_entropy"](x, x)}

    # ---- LinearHyperNetwork (hyper_iNGP.py:18-111): Linear(no bias) -> LayerNorm -> SiLU -> Linear, split per head
    pieces(f"{ROOT}/custom/amortized/models/geometry/hyper_iNGP.py", ["LinearHyperNetwork"], ns)
    torch.manual_seed(9)
    cfg = {"c_dim": 24, "out_dims": {"sdf_weights": [16, 1], "feature_weights": [16, 3]}, "spectral_norm": False,
           "n_neurons": 12, "n_hidden_layers": 1}
    net = ns["LinearHyperNetwork"](8, cfg)
    with torch.no_grad():
        net.layers[1].weight.add_(0.1 * torch.randn(12, generator=g))  # LayerNorm / bias away from their trivial init
        net.layers[1].bias.add_(0.1 * torch.randn(12, generator=g))
        net.layers[3].bias.add_(0.05 * torch.randn(net.n_output_dims, generator=g))
    c = torch.randn(3, 24, generator=g)
    with torch.no_grad():
        out = net(c)
    gold["hypernet"] = {"n_input_dims": 8, "config": {**cfg, "out_dims": {"sdf_weights": [16, 1], "feature_weights": [16, 3]}},
                        "state_dict": {k: v.clone() for k, v in net.state_dict().items()}, "c": c,
                        "out": {k: [t.clone() for t in v] for k, v in out.items()}, "n_output_dims": net.n_output_dims}

    # ---- triplane lookup (geometry/utils.py:28-97): plane axes, projection by inverse axes, grid_sample, plane-major concat
    pieces(f"{ROOT}/custom/amortized/models/geometry/utils.py", ["planes", "project_onto_planes", "sample_from_planes"], ns)
    feats = torch.randn(2, 3, 8, 16, 12, generator=g)  # H != W on purpose
    pts = torch.rand(2, 301, 3, generator=g) * 2.4 - 1.2
    pts[0, :4] = torch.tensor([[1.0, -1.0, 0.0], [-1.0, 1.0, 1.0], [0.3, -0.7, 0.99], [0.0, 0.0, 0.0]])
    gold["triplane"] = {"planes": feats, "points": pts, "out": ns["sample_from_planes"](feats, pts),
                        "proj": ns["project_onto_planes"](ns["planes"], pts[:, :5])}
    # ---- TriplaneTransformerSDF.forward / forward_sdf / interpolate_encodings / get_shifted_sdf
    # (custom/amortized/models/geometry/triplane_transformer.py:103-239) bound to a stand-in that carries a recorded space
    # cache and the reference's own VanillaMLP heads: sdf + sphere bias, features, finite-difference sdf_grad / normals
    import types

    tns = dict(ns, Updateable=object, Sized=(list, tuple), Dict=_Any(), Any=object, Tuple=_Any(), Optional=_Any(),
               Union=_Any(), threestudio=types.SimpleNamespace(debug=lambda *a, **k: None))
    pieces(f"{ROOT}/threestudio/models/networks.py", ["VanillaMLP"], tns)
    pieces(f"{ROOT}/custom/amortized/models/geometry/utils.py", ["contract_to_unisphere_custom"], tns)
    pieces(f"{ROOT}/threestudio/utils/ops.py", ["scale_tensor"], dict_update(tns, Num=_Any(), ValidScale=object))
    src = open(f"{ROOT}/custom/amortized/models/geometry/triplane_transformer.py").read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "TriplaneTransformerSDF")
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("forward", "forward_sdf", "interpolate_encodings", "get_shifted_sdf"):
            body = "\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno])
            exec(compile("\n".join(line[4:] for line in body.splitlines()), "triplane_transformer.py", "exec"), tns)
    torch.manual_seed(13)
    mlp_cfg = {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none", "n_neurons": 64, "n_hidden_layers": 2}
    C3 = 8
    sdf_net, feat_net = tns["VanillaMLP"](3 * C3, 1, mlp_cfg), tns["VanillaMLP"](3 * C3, 3, mlp_cfg)
    radius = 1.0
    geo = types.SimpleNamespace(
        cfg=types.SimpleNamespace(normal_type="finite_difference", n_feature_dims=3, radius=radius, sdf_bias="sphere",
                                  sdf_bias_params=0.8),
        bbox=torch.tensor([[-radius] * 3, [radius] * 3]), unbounded=False, finite_difference_normal_eps=0.01,
        sdf_network=sdf_net, feature_network=feat_net)
    for name in ("forward_sdf", "interpolate_encodings", "get_shifted_sdf"):
        setattr(geo, name, types.MethodType(tns[name], geo))
    cache = torch.randn(2, 3, C3, 16, 16, generator=g) * 0.5
    pts3 = (torch.rand(2, 200, 3, generator=g) * 2 - 1) * 0.98
    with torch.no_grad():
        out = tns["forward"](geo, pts3.clone(), cache, output_normal=True)
    gold["triplane_geometry"] = {
        "space_cache": cache, "points": pts3, "radius": radius, "sdf_bias_radius": 0.8, "fd_eps": 0.01,
        "sdf_weights": [m.weight.detach().clone() for m in sdf_net.layers if isinstance(m, nn.Linear)],
        "feature_weights": [m.weight.detach().clone() for m in feat_net.layers if isinstance(m, nn.Linear)],
        "out": {k: v.clone() for k, v in out.items()}}
    # ---- Hypernet_Sdf.forward / forward_sdf / hypernet_forward / get_shifted_sdf (hyper_iNGP.py:206-349) bound to a stand-in.
    # tiny-cuda-nn is not installable here, so `self.encoding` is the ORACLE's hash-grid restatement: this pins the per-prompt
    # bmm chain, the sphere bias, the clamped finite-difference sdf_grad and the output layout -- not the encoding itself.
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import render_oracle as ro

    hns = dict(tns, Callable=_Any())
    pieces(f"{ROOT}/threestudio/models/geometry/base.py", ["contract_to_unisphere"], hns)
    src = open(f"{ROOT}/custom/amortized/models/geometry/hyper_iNGP.py").read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "Hypernet_Sdf")
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("forward", "forward_sdf", "hypernet_forward", "get_shifted_sdf"):
            body = "\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno])
            exec(compile("\n".join(line[4:] for line in body.splitlines()), "hyper_iNGP.py", "exec"), hns)
    grid = ro.GridCfg(n_levels=16, n_features_per_level=2, log2_hashmap_size=10, base_resolution=4, per_level_scale=1.3)
    table = (torch.rand(ro.grid_meta(grid)["n_entries"], 2, generator=g) * 2 - 1) * 0.3
    enc_mod = types.SimpleNamespace(n_output_dims=32)
    encoding = lambda x: ro.hashgrid_encode(x, table, grid)
    hradius = 2.0
    hgeo = types.SimpleNamespace(
        cfg=types.SimpleNamespace(normal_type="finite_difference", n_feature_dims=3, n_input_dims=3, radius=hradius,
                                  sdf_bias="sphere", sdf_bias_params=0.5),
        bbox=torch.tensor([[-hradius] * 3, [hradius] * 3]), unbounded=False, finite_difference_normal_eps=0.01)
    hgeo.encoding = type("Enc", (), {"n_output_dims": 32, "__call__": lambda self, x: encoding(x)})()
    for name in ("forward_sdf", "hypernet_forward", "get_shifted_sdf"):
        setattr(hgeo, name, types.MethodType(hns[name], hgeo))
    Bh = 2
    hcache = {"sdf_weights": [torch.randn(Bh, 32, 64, generator=g) * 0.3, torch.randn(Bh, 64, 1, generator=g) * 0.3],
              "feature_weights": [torch.randn(Bh, 32, 64, generator=g) * 0.3, torch.randn(Bh, 64, 3, generator=g) * 0.3]}
    hpts = (torch.rand(Bh, 150, 3, generator=g) * 2 - 1) * 1.98
    with torch.no_grad():
        hout = hns["forward"](hgeo, hpts.clone(), hcache, output_normal=True)
    gold["hyper_geometry"] = {"grid": vars(grid), "table": table, "cache": hcache, "points": hpts, "radius": hradius,
                              "sdf_bias_radius": 0.5, "fd_eps": 0.01, "out": {k: v.clone() for k, v in hout.items()}}
    torch.save(gold, OUT)
    print("wrote", OUT, list(gold))


if __name__ == "__main__":
    main()
