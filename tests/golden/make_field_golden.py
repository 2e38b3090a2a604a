"""Generates golden vectors for the C1 field (frequency encoding + VanillaMLP implicit volume) from the REFERENCE's
own class definitions.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_field_golden.py
`threestudio.models.networks` cannot be imported here (it needs tinycudann / pytorch_lightning at module level), so
the class and function definitions the C1 path uses are taken out of the reference files by name with `ast`, compiled
unchanged, and executed against a namespace that supplies only their imports (torch, nn, F, math) and inert stand-ins
for `Updateable` / `threestudio.debug`:
    threestudio/models/networks.py      ProgressiveBandFrequency, CompositeEncoding, VanillaMLP
    threestudio/utils/ops.py            get_activation
    threestudio/models/geometry/implicit_volume.py   ImplicitVolume.get_activated_density (bound to a cfg holder)
Outputs: tests/golden/field_golden.pt (inputs, seeded weights and the reference outputs; a few tens of kB).
"""
import ast
import math
import os
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF = "/root/reference/threestudio"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "field_golden.pt")


def definitions(path, names):
    """Source of the named top-level classes / functions (or Class.method) of a reference file, unchanged."""
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            out[node.name] = ast.get_source_segment(src, node)
        if isinstance(node, ast.ClassDef):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and f"{node.name}.{sub.name}" in names:
                    out[f"{node.name}.{sub.name}"] = ast.get_source_segment(src, sub)
    missing = set(names) - set(out)
    assert not missing, missing
    return out


class _Any:  # stands for the jaxtyping annotations (Float[Tensor, "..."]) evaluated at def time
    def __getitem__(self, item):
        return self


def namespace():
    ts = types.SimpleNamespace(debug=lambda *a, **k: None)
    ns = {"torch": torch, "nn": nn, "F": F, "math": math, "threestudio": ts, "Updateable": object,
          "trunc_exp": None, "Callable": object, "Float": _Any(), "Tensor": torch.Tensor, "Tuple": _Any(),
          "Union": _Any()}
    ops = definitions(f"{REF}/utils/ops.py", ["get_activation"])
    exec(compile(ops["get_activation"], "ops.py", "exec"), ns)
    net = definitions(f"{REF}/models/networks.py", ["ProgressiveBandFrequency", "CompositeEncoding", "VanillaMLP"])
    for name in ("ProgressiveBandFrequency", "CompositeEncoding", "VanillaMLP"):
        exec(compile(net[name], "networks.py", "exec"), ns)
    ns.update({"Dict": _Any(), "Num": _Any(), "ValidScale": object})
    ops2 = definitions(f"{REF}/utils/ops.py", ["scale_tensor"])
    exec(compile(ops2["scale_tensor"], "ops.py", "exec"), ns)
    base = definitions(f"{REF}/models/geometry/base.py", ["contract_to_unisphere"])
    exec(compile(base["contract_to_unisphere"], "base.py", "exec"), ns)
    geo = definitions(f"{REF}/models/geometry/implicit_volume.py",
                      ["ImplicitVolume.get_activated_density", "ImplicitVolume.forward", "ImplicitVolume.forward_density"])
    for name in ("get_activated_density", "forward", "forward_density"):
        code = geo["ImplicitVolume." + name]
        code = "\n".join(line[4:] if line.startswith("    ") else line for line in code.splitlines())
        exec(compile(code, "implicit_volume.py", "exec"), ns)
    return ns


def main():
    ns = namespace()
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(11)
    cases = {}
    for name, n_freq, include_xyz, n_mask, step, n_hidden, bias, act in [
        ("f4", 4, False, 0, None, 2, "blob_magic3d", "softplus"),
        ("f6_xyz_masked", 6, True, 1000, 400, 1, "blob_dreamfusion", "exp"),
        ("f12", 12, False, 0, None, 2, 0.5, "softplus"),
    ]:
        cfg = {"otype": "ProgressiveBandFrequency", "n_frequencies": n_freq, "n_masking_step": n_mask}
        enc = ns["ProgressiveBandFrequency"](3, cfg)
        enc.update_step(0, step)
        comp = ns["CompositeEncoding"](enc, include_xyz=include_xyz, xyz_scale=2.0, xyz_offset=-1.0)
        mlp_cfg = {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none", "n_neurons": 64,
                   "n_hidden_layers": n_hidden}
        dnet = ns["VanillaMLP"](comp.n_output_dims, 1, mlp_cfg)
        fnet = ns["VanillaMLP"](comp.n_output_dims, 3, mlp_cfg)
        radius = 1.0
        points = (torch.rand(257, 3, generator=g) * 2 - 1) * radius
        x01 = (points + radius) / (2 * radius)
        holder = types.SimpleNamespace(cfg=types.SimpleNamespace(
            density_bias=bias, density_blob_scale=10.0, density_blob_std=0.5, density_activation=act))
        with torch.no_grad():
            e = comp(x01)
            raw, density = ns["get_activated_density"](holder, points, dnet(e))
            feat = fnet(e)
            # the whole ImplicitVolume.forward (implicit_volume.py:109-196) with finite-difference normals, bound to a
            # stand-in that carries the reference encoding / networks
            vol = types.SimpleNamespace(
                cfg=types.SimpleNamespace(n_input_dims=3, n_feature_dims=3, normal_type="finite_difference",
                                          finite_difference_normal_eps=0.01, radius=radius, density_bias=bias,
                                          density_blob_scale=10.0, density_blob_std=0.5, density_activation=act),
                bbox=torch.tensor([[-radius] * 3, [radius] * 3]), unbounded=False, encoding=comp, density_network=dnet,
                feature_network=fnet)
            vol.get_activated_density = types.MethodType(ns["get_activated_density"], vol)
            vol.forward_density = types.MethodType(ns["forward_density"], vol)
            full = ns["forward"](vol, points.clone(), output_normal=True)
            assert torch.equal(full["density"][:, 0], density[:, 0]) and torch.equal(full["features"], feat)
        cases[name] = {
            "n_frequencies": n_freq, "include_xyz": include_xyz, "n_masking_step": n_mask, "global_step": step,
            "n_hidden_layers": n_hidden, "density_bias": bias, "density_activation": act, "radius": radius,
            "points": points, "mask": enc.mask.clone(), "enc": e, "density": density[:, 0], "features": feat,
            "normal": full["normal"], "fd_eps": 0.01,
            "density_weights": [m.weight.detach().clone() for m in dnet.layers if isinstance(m, nn.Linear)],
            "feature_weights": [m.weight.detach().clone() for m in fnet.layers if isinstance(m, nn.Linear)],
        }
    torch.save(cases, OUT)
    print("wrote", OUT, {k: tuple(v["enc"].shape) for k, v in cases.items()})


if __name__ == "__main__":
    main()
