"""State-dict keys (names, shapes, dtypes) of the trained modules as the REFERENCE's own `configure` methods lay them
out, so that checkpoints written by either side load on the other (SURVEY.md §8b "sub-module attribute names are part
of the contract", §8f rank 4).

Run in the build container only (it reads /root/reference):
    python tests/golden/make_state_keys_golden.py
The reference modules cannot be imported (tinycudann / nerfacc / pytorch_lightning are absent), so — as in
make_field_golden.py — the definitions are taken out of the reference files by name with `ast`, compiled unchanged and
executed against a namespace that supplies their imports:
    threestudio/models/networks.py            TCNNEncoding, CompositeEncoding, get_encoding, VanillaMLP, get_mlp
    threestudio/models/geometry/base.py       BaseImplicitGeometry.configure (bbox buffer)
    threestudio/models/geometry/implicit_volume.py            ImplicitVolume.configure
    threestudio/models/background/neural_environment_map_background.py   NeuralEnvironmentMapBackground.configure
    custom/amortized/models/geometry/hyper_iNGP.py            LinearHyperNetwork, Hypernet_Sdf.configure
    custom/amortized/models/background/multiprompt_neural_environment_hashgrid_map_background.py   ….configure
Two third-party classes are absent and are stood in for, by name, with what their published source registers:
    tcnn.Encoding            one flat fp32 `params` Parameter (size = the grid's parameter count, or 0)
    nerfacc.OccGridEstimator (v0.5.2 estimators/occ_grid.py) persistent buffers resolution / aabbs / occs / binaries
Output: tests/golden/state_keys_golden.json  {config: {key: [shape, dtype]}}.
"""
import json
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_field_golden import _Any, definitions  # noqa: E402

ROOT = "/root/reference"
REF = f"{ROOT}/threestudio"
AMO = f"{ROOT}/custom/amortized"
OUT = os.path.join(HERE, "state_keys_golden.json")


class Cfg(dict):
    """OmegaConf's DictConfig as far as the configure methods use it: attribute access and .get()."""

    __getattr__ = dict.__getitem__


def grid_param_count(c) -> int:
    """tcnn GridEncoding parameter count (published level geometry)."""
    import math

    total = 0
    for lvl in range(c["n_levels"]):
        scale = 2.0 ** (lvl * math.log2(c["per_level_scale"])) * c["base_resolution"] - 1.0
        res = math.ceil(scale) + 1
        n = min(-(-res ** 3 // 8) * 8, 2 ** c["log2_hashmap_size"])
        total += n * c["n_features_per_level"]
    return total


class FakeTcnnEncoding(nn.Module):
    def __init__(self, n_input_dims, config, dtype=torch.float32):
        super().__init__()
        n = grid_param_count(config) if config["otype"] in ("HashGrid", "Grid") else 0
        self.params = nn.Parameter(torch.zeros(n, dtype=dtype))
        self.n_output_dims = (config["n_levels"] * config["n_features_per_level"] if n else config.get("degree", 3) ** 2)


class FakeOccGridEstimator(nn.Module):
    def __init__(self, roi_aabb, resolution=32, levels=1):
        super().__init__()
        self.register_buffer("resolution", torch.tensor([resolution] * 3, dtype=torch.int32))
        self.register_buffer("aabbs", roi_aabb.reshape(1, 6).float())
        self.register_buffer("occs", torch.zeros(levels * resolution ** 3))
        self.register_buffer("binaries", torch.zeros([levels] + [resolution] * 3, dtype=torch.bool))
        self.register_buffer("grid_coords", torch.zeros(resolution ** 3, 3, dtype=torch.long), persistent=False)
        self.register_buffer("grid_indices", torch.arange(resolution ** 3), persistent=False)


def run(ns, path, names):
    """Top-level definitions are executed as they are; methods (`Class.method`) are kept as source for `configured`."""
    for name, code in definitions(path, names).items():
        if "." in name:
            ns.setdefault("_methods", {})[name] = code
        else:
            exec(compile(code, path, "exec"), ns)


def namespace():
    import torch.nn.functional as F

    tcnn = types.SimpleNamespace(Encoding=FakeTcnnEncoding)
    cuda = types.SimpleNamespace(device=lambda *_: __import__("contextlib").nullcontext())
    ns = {"torch": torch, "nn": nn, "F": F, "tcnn": tcnn, "get_rank": lambda: 0, "Updateable": object,
          "config_to_primitive": dict, "ListConfig": (), "Float": _Any(), "Tensor": torch.Tensor, "Optional": _Any(),
          "IsosurfaceHelper": object, "threestudio": types.SimpleNamespace(debug=lambda *a, **k: None),
          "get_activation": lambda name: None}
    real_device = torch.cuda.device
    torch.cuda.device = cuda.device  # TCNNEncoding.__init__ enters torch.cuda.device(rank); harmless on the CPU
    ns["_restore"] = lambda: setattr(torch.cuda, "device", real_device)
    run(ns, f"{REF}/models/networks.py", ["TCNNEncoding", "CompositeEncoding", "get_encoding", "VanillaMLP", "get_mlp"])
    run(ns, f"{REF}/models/geometry/base.py", ["BaseImplicitGeometry.configure"])
    run(ns, f"{REF}/models/geometry/implicit_volume.py", ["ImplicitVolume.configure"])
    run(ns, f"{REF}/models/background/neural_environment_map_background.py", ["NeuralEnvironmentMapBackground.configure"])
    run(ns, f"{AMO}/models/geometry/hyper_iNGP.py", ["LinearHyperNetwork", "Hypernet_Sdf.configure"])
    run(ns, f"{AMO}/models/background/multiprompt_neural_environment_hashgrid_map_background.py",
        ["MultipromptNeuralHashgridEnvironmentMapBackground.configure"])
    return ns


def configured(ns, method, cfg, base=None):
    """An nn.Module carrying `cfg`, configured by the reference method `Class.configure`, compiled unchanged inside a
    class body (so its zero-argument `super().configure()` resolves) whose parent runs the `base` method."""
    env = dict(ns)
    parent = "nn.Module"
    if base is not None:
        exec(compile("class Parent(nn.Module):\n    " + ns["_methods"][base] + "\n", base, "exec"), env)
        parent = "Parent"
    exec(compile(f"class Holder({parent}):\n    " + ns["_methods"][method] + "\n", method, "exec"), env)
    m = env["Holder"]()
    m.cfg = cfg
    m.configure()
    return m


GRID = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16,
        "per_level_scale": 1.447269237440378}
MLP = {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none", "n_neurons": 64, "n_hidden_layers": 1}


def c2_system(ns):
    """configs/single-prompt_benchmark/asd_sd_nerf.yaml:38-75; BaseLift3DSystem.configure (systems/base.py:292-303)
    assigns geometry / material / background / renderer as children of the system."""
    sys_ = nn.Module()
    sys_.geometry = configured(ns, "ImplicitVolume.configure", Cfg(
        radius=1.0, n_input_dims=3, n_feature_dims=3, normal_type="finite_difference",
        pos_encoding_config=Cfg(GRID), mlp_network_config=Cfg(MLP)), base="BaseImplicitGeometry.configure")
    sys_.material = nn.Module()  # NoMaterial: no parameters (materials/no_material.py:15-54)
    bg_grid = dict(GRID, n_levels=4, base_resolution=4, per_level_scale=4.0)  # asd_sd_nerf.yaml:66-73
    sys_.background = configured(ns, "NeuralEnvironmentMapBackground.configure", Cfg(
        n_output_dims=3, dir_encoding_config=Cfg(bg_grid), mlp_network_config=Cfg(dict(MLP, n_neurons=16, n_hidden_layers=2))))
    ren = nn.Module()  # renderers/base.py:37-58 (bbox buffer) + nerf_volume_renderer.py:60-65 (self.estimator)
    ren.register_buffer("bbox", torch.tensor([[-1.0] * 3, [1.0] * 3]))
    ren.estimator = FakeOccGridEstimator(ren.bbox.view(-1), resolution=32, levels=1)
    sys_.renderer = ren
    return sys_


def c4_system(ns):
    """configs/multi-prompt_benchmark/asd_sd_hyper_iNGP_50k.yaml geometry / background sections (class defaults for
    the keys the yaml leaves out: hyper_iNGP.py:117-163, multiprompt_neural_environment_hashgrid_map_background.py:20-47)
    + the renderer's LearnedVariance (generative_space_volsdf_volume_renderer.py:23-34,75)."""
    hyper = {"c_dim": 1024, "n_neurons": 64, "n_hidden_layers": 1, "spectral_norm": False, "output_activation": "none"}
    sys_ = nn.Module()
    sys_.geometry = configured(ns, "Hypernet_Sdf.configure", Cfg(
        radius=2.0, n_input_dims=3, backbone="linear_hypernetwork", normal_type="finite_difference",
        isosurface_deformable_grid=False, pos_encoding_config=Cfg(GRID),
        hypernet_config=Cfg(dict(hyper, out_dims={"sdf_weights": [64, 1], "feature_weights": [64, 3]}))),
        base="BaseImplicitGeometry.configure")
    bg_grid = dict(GRID, per_level_scale=1.0)  # asd_sd_hyper_iNGP_50k.yaml background.pos_encoding_config
    sys_.background = configured(ns, "MultipromptNeuralHashgridEnvironmentMapBackground.configure", Cfg(
        pos_encoding_config=Cfg(bg_grid), hypernet_config=Cfg(dict(hyper, out_dims={"bg_weights": [64, 3]}))))
    ren = nn.Module()
    ren.register_buffer("bbox", torch.tensor([[-2.0] * 3, [2.0] * 3]))
    run(ns, f"{AMO}/models/renderers/generative_space_volsdf_volume_renderer.py", ["LearnedVariance"])
    ren.variance = ns["LearnedVariance"](0.340119, requires_grad=False)
    sys_.renderer = ren
    return sys_


def describe(module):
    return {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in module.state_dict().items()}


def main():
    ns = namespace()
    try:
        out = {"C2": describe(c2_system(ns)), "C4": describe(c4_system(ns))}
    finally:
        ns["_restore"]()
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, len(v), "keys")
        for name, d in v.items():
            print("   ", name, d)


if __name__ == "__main__":
    main()
