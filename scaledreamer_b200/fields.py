"""Field and renderer plugins of the single-prompt NeRF path (threestudio names, Config keys, attribute names and
state-dict keys), all evaluated by the fused sm_100a render kernels:

  "implicit-volume"                     threestudio/models/geometry/implicit_volume.py:19
  "no-material"                         threestudio/models/materials/no_material.py:15
  "neural-environment-map-background"   threestudio/models/background/neural_environment_map_background.py:15
  "nerf-volume-renderer"                threestudio/models/renderers/nerf_volume_renderer.py:20

Sub-module attribute names are part of the contract (optimizer groups and checkpoints address
`geometry.encoding`, `geometry.density_network`, `geometry.feature_network`, `background.encoding`,
`background.network`; configs/single-prompt_benchmark/asd_sd_nerf.yaml:115-125).
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import lib as L
from . import render_ops as R
from .core import BaseModule, Updateable, register

DEFAULT_GRID = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19,
                "base_resolution": 16, "per_level_scale": 1.447269237440378}
DEFAULT_MLP = {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none", "n_neurons": 64,
               "n_hidden_layers": 1}


class _GridParams(nn.Module):
    """Stands where tcnn.Encoding stands: one flat fp32 `params` vector in tcnn's level-major layout,
    U(-1e-4, 1e-4) initialised (tcnn default)."""

    def __init__(self, n_params: int, seed: int = 1337):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.params = nn.Parameter((torch.rand(n_params, generator=g) * 2 - 1) * 1e-4)


class _TCNNEncoding(nn.Module):
    """Stands where networks.py:55-64 (TCNNEncoding) stands: `.encoding` is the tcnn.Encoding stand-in, so the
    parameter's state-dict path below a geometry is `encoding.encoding.encoding.params`, as in the reference's
    checkpoints. `.params` is kept as a shortcut to the same Parameter."""

    def __init__(self, n_params: int):
        super().__init__()
        self.encoding = _GridParams(n_params)

    @property
    def params(self) -> nn.Parameter:
        return self.encoding.params


class HashGridEncoding(nn.Module):
    """networks.py:194-211 (get_encoding): CompositeEncoding(`.encoding` = TCNNEncoding(`.encoding` = tcnn.Encoding
    with the flat `params`)), include_xyz false. n_output_dims = n_levels * n_features_per_level."""

    def __init__(self, n_input_dims: int, config: dict):
        super().__init__()
        if config.get("otype") not in ("HashGrid", "Grid"):
            raise NotImplementedError(f"encoding otype {config.get('otype')} is not implemented by the sm_100a kernels "
                                      "(HashGrid only)")
        if n_input_dims != 3:
            raise NotImplementedError("HashGrid encodings take 3-D inputs")
        self.grid_cfg = {k: config[k] for k in ("n_levels", "n_features_per_level", "log2_hashmap_size",
                                                "base_resolution", "per_level_scale")}
        self.n_input_dims = n_input_dims
        self.n_output_dims = int(config["n_levels"]) * int(config["n_features_per_level"])
        self.n_entries = L.grid_num_entries(self.grid_cfg)
        self.encoding = _TCNNEncoding(self.n_entries * int(config["n_features_per_level"]))
        self._register_load_state_dict_pre_hook(self._accept_flat_key)

    @staticmethod
    def _accept_flat_key(state_dict, prefix, *unused) -> None:
        """Checkpoints written by earlier builds of this package held the table one level higher
        (`<prefix>encoding.params`); move it to the reference's path before the regular load."""
        old, new = prefix + "encoding.params", prefix + "encoding.encoding.params"
        if old in state_dict and new not in state_dict:
            state_dict[new] = state_dict.pop(old)

    @property
    def table(self) -> torch.Tensor:
        return self.encoding.encoding.params

    def forward(self, x01: torch.Tensor) -> torch.Tensor:
        return R.hashgrid_forward(x01.reshape(-1, 3), self.table.detach().view(-1, 2), self.grid_cfg)


class SphericalHarmonicsEncoding(nn.Module):
    """tiny-cuda-nn "SphericalHarmonics" (un-vendored; real SH basis on the direction 2 x01 - 1), degrees 1..3. It has no
    parameters and runs once per RAY, so it stays a few element-wise torch ops."""

    def __init__(self, degree: int):
        super().__init__()
        if not 1 <= degree <= 3:
            raise NotImplementedError("SphericalHarmonics encodings of degree 1..3 are implemented")
        self.degree, self.n_input_dims, self.n_output_dims = degree, 3, degree * degree

    def forward(self, x01: torch.Tensor) -> torch.Tensor:
        d = x01 * 2.0 - 1.0
        x, y, z = d[..., 0], d[..., 1], d[..., 2]
        out = [torch.full_like(x, 0.28209479177387814)]
        if self.degree > 1:
            out += [-0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x]
        if self.degree > 2:
            out += [1.0925484305920792 * x * y, -1.0925484305920792 * y * z,
                    0.94617469575755997 * z * z - 0.31539156525251999, -1.0925484305920792 * x * z,
                    0.54627421529603959 * (x * x - y * y)]
        return torch.stack(out, dim=-1)


class _HashGridFn(torch.autograd.Function):
    """tcnn.Encoding forward / backward (sdb_hashgrid_forward / sdb_hashgrid_backward) with a gradient to the table."""

    @staticmethod
    def forward(ctx, x01, table, grid_cfg):
        ctx.save_for_backward(x01.detach())
        ctx.grid_cfg, ctx.n_entries = grid_cfg, table.shape[0]
        return R.hashgrid_forward(x01.detach(), table.detach(), grid_cfg)

    @staticmethod
    def backward(ctx, g_out):
        (x01,) = ctx.saved_tensors
        return None, R.hashgrid_backward(x01, g_out, ctx.n_entries, ctx.grid_cfg), None


class VanillaMLP(nn.Module):
    """networks.py:214-251: bias-free Linear + ReLU stack; keys layers.{0,2,...}.weight."""

    def __init__(self, dim_in: int, dim_out: int, config: dict):
        super().__init__()
        if config.get("otype", "VanillaMLP") != "VanillaMLP":
            raise NotImplementedError(f"network otype {config.get('otype')} is not implemented (VanillaMLP only)")
        if config.get("output_activation", "none") not in (None, "none"):
            raise NotImplementedError("VanillaMLP output_activation must be none")
        self.n_neurons, self.n_hidden_layers = int(config["n_neurons"]), int(config["n_hidden_layers"])
        layers = [nn.Linear(dim_in, self.n_neurons, bias=False), nn.ReLU(inplace=True)]
        for _ in range(self.n_hidden_layers - 1):
            layers += [nn.Linear(self.n_neurons, self.n_neurons, bias=False), nn.ReLU(inplace=True)]
        layers += [nn.Linear(self.n_neurons, dim_out, bias=False)]
        self.layers = nn.Sequential(*layers)

    def weights(self):
        return [m.weight for m in self.layers if isinstance(m, nn.Linear)]


class _TinyMLP(torch.autograd.Function):
    """y [n, k] = W3 relu(W2 relu(W1 x)) for the bias-free 64-wide VanillaMLP heads (networks.py:214-251), native
    fp32 kernels (csrc/tiny_mlp.cu). Nothing but x is kept for the backward: the hidden layers are recomputed."""

    @staticmethod
    def forward(ctx, x, w1, w2, w3):
        lib = L.load()
        x = x.contiguous().float()
        w1, w2, w3 = (w.detach().contiguous().float() for w in (w1, w2, w3))
        n, d_in, k = x.shape[0], x.shape[1], w3.shape[0]
        y = torch.empty(n, k, device=x.device)
        L.check(lib.sdb_mlp3_forward(L.ptr(x.detach()), n, d_in, L.ptr(w1), L.ptr(w2), L.ptr(w3), k, L.ptr(y),
                                     L.stream_ptr()), "sdb_mlp3_forward")
        ctx.save_for_backward(x.detach(), w1, w2, w3)
        return y

    @staticmethod
    def backward(ctx, d_y):
        x, w1, w2, w3 = ctx.saved_tensors
        n, d_in, k = x.shape[0], x.shape[1], w3.shape[0]
        d_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        g1, g2, g3 = torch.zeros_like(w1), torch.zeros_like(w2), torch.zeros_like(w3)
        L.check(L.load().sdb_mlp3_backward(L.ptr(x), n, d_in, L.ptr(w1), L.ptr(w2), L.ptr(w3), k,
                                           L.ptr(d_y.contiguous().float()), L.ptr(d_x) if d_x is not None else None,
                                           0, L.ptr(g1), L.ptr(g2), L.ptr(g3), L.stream_ptr()), "sdb_mlp3_backward")
        return d_x, g1, g2, g3


def tiny_mlp(mlp: "VanillaMLP", x: torch.Tensor) -> torch.Tensor:
    """Evaluate a VanillaMLP (64 neurons, 1 or 2 hidden layers, 1 or 3 outputs) on x [..., d] with the native kernels.
    x may carry zero padding columns beyond the MLP's input width (the first layer is padded to match). A single
    hidden layer runs as the two-layer kernel with an identity second layer: relu(I relu(h)) == relu(h) exactly."""
    ws = list(mlp.weights())
    d = x.shape[-1]
    if mlp.n_neurons != 64 or len(ws) not in (2, 3) or ws[-1].shape[0] not in (1, 3) or d % 8 or not 8 <= d <= 96 \
            or ws[0].shape[1] > d:
        raise NotImplementedError("native tiny MLP: d_in -> 64 [-> 64] -> {1,3}, d_in <= 96 (rows padded to 8)")
    w1 = ws[0] if ws[0].shape[1] == d else torch.nn.functional.pad(ws[0], (0, d - ws[0].shape[1]))
    w2 = ws[1] if len(ws) == 3 else torch.eye(64, device=x.device)
    return _TinyMLP.apply(x.reshape(-1, d), w1, w2, ws[-1]).view(*x.shape[:-1], ws[-1].shape[0])


class ProgressiveBandFrequency(nn.Module, Updateable):
    """networks.py:16-52 wrapped as get_encoding does (CompositeEncoding, :170-206): sin / cos of 2^f x01 with the
    coarse-to-fine mask, optional leading xyz*2-1 block. Parameter-free; rows come back zero-padded to a multiple of 8
    columns (`padded_dims`) for the MLP kernels."""

    def __init__(self, n_input_dims: int, config: dict):
        super().__init__()
        if n_input_dims != 3:
            raise NotImplementedError("ProgressiveBandFrequency takes 3-D inputs")
        self.N_freqs = int(config["n_frequencies"])
        self.n_masking_step = int(config.get("n_masking_step", 0))
        self.include_xyz = bool(config.get("include_xyz", False))
        self.n_input_dims = n_input_dims
        self.n_output_dims = 3 * int(self.include_xyz) + 6 * self.N_freqs
        self.padded_dims = -(-self.n_output_dims // 8) * 8
        self.register_buffer("mask", torch.ones(self.N_freqs), persistent=False)

    def forward(self, x01: torch.Tensor) -> torch.Tensor:
        return R.freq_encode(x01, self.N_freqs, self.mask, self.include_xyz)

    def update_step(self, epoch, global_step, on_load_weights: bool = False):
        if self.n_masking_step <= 0 or global_step is None:
            self.mask.fill_(1.0)
        else:
            ramp = (global_step / self.n_masking_step * self.N_freqs - torch.arange(0, self.N_freqs)).clamp(0, 1)
            self.mask.copy_((1.0 - torch.cos(math.pi * ramp)) / 2.0)


@register("implicit-volume")
class ImplicitVolume(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        radius: float = 1.0
        isosurface: bool = True
        isosurface_method: str = "mt"
        isosurface_resolution: int = 128
        isosurface_threshold: Union[float, str] = 25.0
        isosurface_chunk: int = 0
        isosurface_coarse_to_fine: bool = True
        isosurface_deformable_grid: bool = False
        isosurface_remove_outliers: bool = True
        isosurface_outlier_n_faces_threshold: Union[int, float] = 0.01
        n_input_dims: int = 3
        n_feature_dims: int = 3
        density_activation: Optional[str] = "softplus"
        density_bias: Union[float, str] = "blob_magic3d"
        density_blob_scale: float = 10.0
        density_blob_std: float = 0.5
        pos_encoding_config: dict = field(default_factory=lambda: dict(DEFAULT_GRID))
        mlp_network_config: dict = field(default_factory=lambda: dict(DEFAULT_MLP))
        normal_type: Optional[str] = "finite_difference"
        finite_difference_normal_eps: float = 0.01
        anneal_density_blob_std_config: Optional[dict] = None

    cfg: Config

    def configure(self) -> None:
        r = self.cfg.radius
        self.register_buffer("bbox", torch.as_tensor([[-r, -r, -r], [r, r, r]], dtype=torch.float32))
        self.unbounded = False
        if self.cfg.n_feature_dims != 3:
            raise NotImplementedError("the renderers evaluate a 3-channel feature network")
        if self.cfg.normal_type not in (None, "finite_difference"):
            raise NotImplementedError(f"normal_type {self.cfg.normal_type} is not implemented (finite_difference only)")
        mlp, enc = self.cfg.mlp_network_config, self.cfg.pos_encoding_config
        if int(mlp["n_neurons"]) != 64 or int(mlp["n_hidden_layers"]) not in (1, 2):
            raise NotImplementedError("the sm_100a MLP kernels are built for n_neurons 64 and 1 or 2 hidden layers")
        if enc.get("otype") == "ProgressiveBandFrequency":
            self.encoding = ProgressiveBandFrequency(self.cfg.n_input_dims, enc)
        else:
            self.encoding = HashGridEncoding(self.cfg.n_input_dims, enc)
        n_enc = self.encoding.n_output_dims
        if n_enc > 96:
            raise NotImplementedError("encodings wider than 96 are not covered by the MLP kernels")
        # The fused ray-march + field + composite kernels cover the iNGP layout of the BASELINE configs C2/C3
        # (32-wide hash grid, 32-64-{1,3} MLPs); anything else (C1: frequency encoding / deeper MLP) is rendered by
        # the packed stages with this module's own forward().
        self.fusable = isinstance(self.encoding, HashGridEncoding) and n_enc == 32 and int(mlp["n_hidden_layers"]) == 1
        self.density_network = VanillaMLP(n_enc, 1, mlp)
        self.feature_network = VanillaMLP(n_enc, self.cfg.n_feature_dims, mlp)

    def field_params(self) -> Dict[str, torch.Tensor]:
        if not self.fusable:
            raise NotImplementedError("this geometry is not in the fused renderer's layout (32-wide HashGrid, one "
                                      "hidden layer); it renders through the packed stages")
        w1d, w2d = self.density_network.weights()
        w1f, w2f = self.feature_network.weights()
        return {"table": self.encoding.table, "w1d": w1d, "w2d": w2d, "w1f": w1f, "w2f": w2f}

    def field_spec_kwargs(self) -> dict:
        return dict(grid=self.encoding.grid_cfg, radius=self.cfg.radius, density_bias=self.cfg.density_bias,
                    density_blob_scale=self.cfg.density_blob_scale, density_blob_std=self.cfg.density_blob_std,
                    density_activation=self.cfg.density_activation or "softplus",
                    fd_eps=self.cfg.finite_difference_normal_eps)

    def _spec(self) -> R.FieldSpec:
        return R.FieldSpec(bg_grid=self.encoding.grid_cfg, **self.field_spec_kwargs())

    # ---- packed path: differentiable evaluation on arbitrary points (implicit_volume.py:80-107, 109-207)
    def _encode(self, points: torch.Tensor) -> torch.Tensor:
        x01 = ((points - self.bbox[0]) / (self.bbox[1] - self.bbox[0])).detach()  # contract_to_unisphere, bounded
        if isinstance(self.encoding, HashGridEncoding):
            return _HashGridFn.apply(x01, self.encoding.table.view(-1, 2), self.encoding.grid_cfg)
        return self.encoding(x01)

    def _activated_density(self, points: torch.Tensor, raw: torch.Tensor) -> torch.Tensor:
        bias = self.cfg.density_bias
        if bias == "blob_dreamfusion":
            raw = raw + self.cfg.density_blob_scale * torch.exp(-0.5 * (points ** 2).sum(-1) / self.cfg.density_blob_std ** 2)
        elif bias == "blob_magic3d":
            raw = raw + self.cfg.density_blob_scale * (1 - torch.sqrt((points ** 2).sum(-1)) / self.cfg.density_blob_std)
        elif isinstance(bias, (int, float)):
            raw = raw + float(bias)
        else:
            raise ValueError(f"Unknown density bias {bias}")
        act = self.cfg.density_activation or "softplus"
        if act == "softplus":
            return torch.nn.functional.softplus(raw)
        if act == "exp":
            return torch.exp(raw)
        raise NotImplementedError(f"density_activation {act} is not implemented (softplus, exp)")

    def _density(self, points: torch.Tensor, enc: Optional[torch.Tensor] = None) -> torch.Tensor:
        enc = self._encode(points) if enc is None else enc
        return self._activated_density(points, tiny_mlp(self.density_network, enc)[..., 0])

    def _forward_packed(self, points: torch.Tensor, output_normal: bool) -> Dict[str, torch.Tensor]:
        pts = points.reshape(-1, 3)
        enc = self._encode(pts)
        density = self._density(pts, enc)
        out = {"density": density[:, None], "features": tiny_mlp(self.feature_network, enc)}
        if output_normal:
            eps, r = self.cfg.finite_difference_normal_eps, self.cfg.radius
            offs = (pts[:, None, :] + eps * torch.eye(3, device=pts.device)).clamp(-r, r)
            d_off = self._density(offs.reshape(-1, 3)).view(-1, 3)
            normal = torch.nn.functional.normalize(-(d_off - density[:, None]) / eps, dim=-1)
            out.update(normal=normal, shading_normal=normal)
        return out

    def forward(self, points: torch.Tensor, output_normal: bool = False) -> Dict[str, torch.Tensor]:
        """implicit_volume.py:109-196 on arbitrary points. In the fused layout this entry is gradient-free (training
        gradients flow through the fused renderer; it serves occupancy refresh, export and inspection); otherwise it
        is the differentiable field the packed renderer calls."""
        shp = points.shape[:-1]
        if not self.fusable:
            out = self._forward_packed(points, output_normal)
            return {k: v.view(*shp, v.shape[-1]) for k, v in out.items()}
        P = {k: v.detach() for k, v in self.field_params().items()}
        d, f, n = R.field_forward(self._spec(), P, points.reshape(-1, 3), want_features=True, want_normal=output_normal)
        out = {"density": d.view(*shp, 1), "features": f.view(*shp, 3)}
        if output_normal:
            out["normal"] = n.view(*shp, 3)
            out["shading_normal"] = out["normal"]
        return out

    def forward_density(self, points: torch.Tensor) -> torch.Tensor:
        shp = points.shape[:-1]
        if not self.fusable:
            return self._density(points.reshape(-1, 3)).view(*shp, 1)
        P = {k: v.detach() for k, v in self.field_params().items()}
        d, _, _ = R.field_forward(self._spec(), P, points.reshape(-1, 3), want_features=False)
        return d.view(*shp, 1)


@register("no-material")
class NoMaterial(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        n_output_dims: int = 3
        color_activation: str = "sigmoid"
        input_feature_dims: Optional[int] = None
        mlp_network_config: Optional[dict] = None
        requires_normal: bool = False

    cfg: Config
    requires_tangent = False

    def configure(self) -> None:
        if self.cfg.input_feature_dims is not None and self.cfg.mlp_network_config is not None:
            raise NotImplementedError("no-material with an MLP head is not implemented by the fused renderer")
        if self.cfg.color_activation not in R.COLOR_ACTS:
            raise NotImplementedError(f"color_activation {self.cfg.color_activation} is not implemented")
        self.requires_normal = self.cfg.requires_normal

    def forward(self, features: torch.Tensor, **kwargs) -> torch.Tensor:
        x = features.view(-1, self.cfg.n_output_dims)
        s = torch.sigmoid(x)
        if self.cfg.color_activation == "sigmoid-mipnerf":
            s = s * 1.002 - 0.001
        return s.view(*features.shape[:-1], self.cfg.n_output_dims)


@register("neural-environment-map-background")
class NeuralEnvironmentMapBackground(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        n_output_dims: int = 3
        color_activation: str = "sigmoid"
        dir_encoding_config: dict = field(default_factory=lambda: {"otype": "SphericalHarmonics", "degree": 3})
        mlp_network_config: dict = field(default_factory=lambda: {"otype": "VanillaMLP", "activation": "ReLU",
                                                                  "n_neurons": 16, "n_hidden_layers": 2})
        random_aug: bool = False
        random_aug_prob: float = 0.5
        eval_color: Optional[Tuple[float, float, float]] = None

    cfg: Config

    def configure(self) -> None:
        enc_cfg, mlp = self.cfg.dir_encoding_config, self.cfg.mlp_network_config
        if enc_cfg.get("otype") == "SphericalHarmonics":  # the class default; used by the Triplane configs
            self.encoding = SphericalHarmonicsEncoding(int(enc_cfg.get("degree", 3)))
        else:
            self.encoding = HashGridEncoding(3, enc_cfg)
        self.fusable = (isinstance(self.encoding, HashGridEncoding) and self.encoding.n_output_dims == 8
                        and int(mlp["n_neurons"]) == 16 and int(mlp["n_hidden_layers"]) == 2)
        self.network = VanillaMLP(self.encoding.n_output_dims, self.cfg.n_output_dims, mlp)

    def field_params(self) -> Dict[str, torch.Tensor]:
        if not self.fusable:
            raise NotImplementedError("the fused NeRF renderer is built for a 4-level x 2 hash grid and an 8-16-16-3 MLP "
                                      "environment map")
        w1, w2, w3 = self.network.weights()
        return {"bg_table": self.encoding.table, "bg_w1": w1, "bg_w2": w2, "bg_w3": w3}

    def forward(self, dirs: torch.Tensor) -> torch.Tensor:
        """Stand-alone differentiable evaluation (neural_environment_map_background.py:46-67) for renderers that call
        the background as a module (the VolSDF renderer of the amortized path). The NeRF renderer fuses it instead."""
        shp = dirs.shape[:-1]
        x01 = ((dirs + 1.0) / 2.0).reshape(-1, 3)
        if isinstance(self.encoding, SphericalHarmonicsEncoding):
            enc = self.encoding(x01)
        else:
            enc = _HashGridFn.apply(x01, self.encoding.table.view(-1, 2), self.encoding.grid_cfg)
        color = torch.sigmoid(self.network.layers(enc))
        if self.cfg.color_activation == "sigmoid-mipnerf":
            color = color * 1.002 - 0.001
        color = color.view(*shp, self.cfg.n_output_dims)
        if self.training and self.cfg.random_aug and random.random() < self.cfg.random_aug_prob:
            color = color * 0 + torch.rand(dirs.shape[0], 1, 1, self.cfg.n_output_dims, device=dirs.device).expand(
                *shp, -1)
        return color

    def sample_override(self, batch_size: int, device) -> Optional[torch.Tensor]:
        """Random solid colour with probability random_aug_prob while training (…background.py:56-66): host coin,
        device colours. The environment map receives no gradient on those steps (color * 0 + rand)."""
        if self.training and self.cfg.random_aug and random.random() < self.cfg.random_aug_prob:
            return torch.rand(batch_size, 3, device=device)
        return None


@register("nerf-volume-renderer")
class NeRFVolumeRenderer(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        radius: float = 1.0
        num_samples_per_ray: int = 512
        eval_chunk_size: int = 160000
        randomized: bool = True
        near_plane: float = 0.0
        far_plane: float = 1e10
        return_comp_normal: bool = False
        return_normal_perturb: bool = False
        estimator: str = "occgrid"
        grid_prune: bool = True
        prune_alpha_threshold: bool = True
        proposal_network_config: Optional[dict] = None
        prop_optimizer_config: Optional[dict] = None
        prop_scheduler_config: Optional[dict] = None
        num_samples_per_ray_proposal: int = 64
        num_samples_per_ray_importance: int = 64

    cfg: Config
    OCC_RES = 32

    def configure(self, geometry, material, background) -> None:
        @dataclass
        class SubModules:  # kept out of nn.Module registration (renderers/base.py:28-35)
            geometry: Any
            material: Any
            background: Any

        self.sub_modules = SubModules(geometry, material, background)
        r = self.cfg.radius
        self.register_buffer("bbox", torch.as_tensor([[-r, -r, -r], [r, r, r]], dtype=torch.float32))
        if self.cfg.estimator != "occgrid":
            raise NotImplementedError("Unknown estimator, should be in ['occgrid'] for the fused sm_100a renderer "
                                      f"(got {self.cfg.estimator})")
        if self.cfg.return_comp_normal or self.cfg.return_normal_perturb:
            raise NotImplementedError("comp_normal outputs are not implemented by the fused renderer")
        self.render_step_size = 1.732 * 2 * self.cfg.radius / self.cfg.num_samples_per_ray
        self.randomized = self.cfg.randomized
        self.occ: Optional[R.OccGrid] = None
        self.packed_capacity = 0  # set > 0 by a system that consumes per-sample outputs
        # set by the system while loss.lambda_orient > 0: the fused path then returns `orient` [B,H,W,1], the per-ray
        # sum_i w_i relu(n_i . d)^2 (gradients through the finite-difference normals), instead of per-sample tensors
        self.orient_loss = False

    geometry = property(lambda self: self.sub_modules.geometry)
    material = property(lambda self: self.sub_modules.material)
    background = property(lambda self: self.sub_modules.background)

    def _occ_grid(self, device) -> R.OccGrid:
        if self.occ is None:
            self.occ = R.OccGrid(self.OCC_RES, device, all_occupied=not self.cfg.grid_prune)
        elif self.occ.occs.device.type != torch.device(device).type:  # module moved after the grid was created
            self.occ.occs, self.occ.bits, self.occ.mean = (t.to(device) for t in (self.occ.occs, self.occ.bits,
                                                                                   self.occ.mean))
        return self.occ

    # The occupancy state travels in checkpoints under the names of nerfacc.OccGridEstimator's persistent buffers
    # (nerfacc 0.5.2 estimators/occ_grid.py; `self.estimator` at nerf_volume_renderer.py:60-65): `estimator.resolution`
    # int32 [3], `estimator.aabbs` [1,6], `estimator.occs` float [res^3] (cell = (x*res + y)*res + z) and
    # `estimator.binaries` bool [1,res,res,res], so a Lightning checkpoint of the reference resumes here with its grid.
    def _save_to_state_dict(self, destination, prefix, keep_vars) -> None:
        super()._save_to_state_dict(destination, prefix, keep_vars)
        occ, res, dev = self._occ_grid(self.bbox.device), self.OCC_RES, self.bbox.device
        destination[prefix + "estimator.resolution"] = torch.full((3,), res, dtype=torch.int32, device=dev)
        destination[prefix + "estimator.aabbs"] = self.bbox.detach().reshape(1, 6).clone()
        destination[prefix + "estimator.occs"] = occ.occs.detach().clone()
        destination[prefix + "estimator.binaries"] = occ.binaries().reshape(1, res, res, res)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs) -> None:
        est = {k[len(prefix) + len("estimator."):]: state_dict.pop(k)
               for k in [k for k in state_dict if k.startswith(prefix + "estimator.")]}
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        if "occs" in est and "binaries" in est:
            n = self.OCC_RES ** 3
            if est["occs"].numel() != n or est["binaries"].numel() != n:
                error_msgs.append(f"{prefix}estimator: occupancy grid of {est['occs'].numel()} cells, expected {n} "
                                  f"(resolution {self.OCC_RES}, one level)")
                return
            self._occ_grid(self.bbox.device).set_binaries(est["binaries"].reshape(-1).bool(), est["occs"].float())
        elif strict and est.keys() & {"occs", "binaries"}:
            missing_keys.append(prefix + ("estimator.binaries" if "occs" in est else "estimator.occs"))

    def _spec(self) -> R.FieldSpec:
        return R.FieldSpec(bg_grid=self.background.encoding.grid_cfg,
                           color_activation=self.material.cfg.color_activation,
                           bg_color_activation=self.background.cfg.color_activation,
                           **self.geometry.field_spec_kwargs())

    def _params(self) -> Dict[str, torch.Tensor]:
        P = dict(self.geometry.field_params())
        P.update(self.background.field_params())
        P["table"] = P["table"].view(-1, 2)
        P["bg_table"] = P["bg_table"].view(-1, 2)
        return P

    def forward(self, rays_o, rays_d, light_positions=None, bg_color=None, **kwargs) -> Dict[str, torch.Tensor]:
        B, H, W = rays_o.shape[:3]
        dev = rays_o.device
        n_rays = B * H * W
        march = R.MarchSpec(render_step_size=self.render_step_size, near_plane=self.cfg.near_plane,
                            far_plane=self.cfg.far_plane, prune=self.cfg.grid_prune and self.cfg.prune_alpha_threshold,
                            alpha_thre=0.01, grid_res=self.OCC_RES,
                            output_normal=bool(self.material.requires_normal and self.packed_capacity > 0))
        jitter = torch.rand(n_rays, device=dev) if self.randomized else None
        if not getattr(self.geometry, "fusable", False):
            return self._forward_packed(rays_o, rays_d, light_positions, bg_color, march, jitter)
        if bg_color is not None:
            bg_override = bg_color.to(dev, torch.float32).reshape(-1, 3).expand(B, 3).contiguous()
        else:
            bg_override = self.background.sample_override(B, dev)
        want_orient = bool(self.orient_loss and self.training and self.material.requires_normal)
        out = R.render_nerf(self._spec(), march, self._occ_grid(dev), self._params(), rays_o.reshape(-1, 3),
                            rays_d.reshape(-1, 3), jitter, bg_override, H * W, self.packed_capacity, want_orient)
        res = {"comp_rgb": out["comp_rgb"].view(B, H, W, 3), "comp_rgb_fg": out["comp_rgb_fg"].view(B, H, W, 3),
               "comp_rgb_bg": out["comp_rgb_bg"].view(B, H, W, 3), "opacity": out["opacity"].view(B, H, W, 1),
               "depth": out["depth"].view(B, H, W, 1), "z_variance": out["z_variance"].view(B, H, W, 1)}
        if "orient" in out:
            res["orient"] = out["orient"].view(B, H, W, 1)
        pk = out.get("packed")
        if self.training and pk is not None:
            res.update({"weights": pk["weights"], "t_starts": pk["t_starts"], "t_ends": pk["t_ends"],
                        "ray_indices": pk["ray_indices"], "density": pk["density"], "n_samples": pk["counter"]})
            if pk.get("normal") is not None:
                res["normal"] = pk["normal"]
        return res

    def _forward_packed(self, rays_o, rays_d, light_positions, bg_color, march: "R.MarchSpec", jitter):
        """The reference's own stage order (nerf_volume_renderer.py:139-386) over packed samples, for geometries outside
        the fused layout: march -> sigma_fn pruning -> geometry / material on the kept samples -> compositing."""
        B, H, W = rays_o.shape[:3]
        dev = rays_o.device
        n_rays = B * H * W
        ro, rd = rays_o.reshape(-1, 3).contiguous().float(), rays_d.reshape(-1, 3).contiguous().float()
        occ = self._occ_grid(dev)
        smp = R.march_packed(march, self.cfg.radius, occ, ro, rd, jitter)
        if march.prune and smp["positions"].shape[0] > 0:
            with torch.no_grad():
                sigma = self.geometry.forward_density(smp["positions"])[..., 0]
            smp = R.prune_packed(smp, sigma, march, occ)
        ray_idx = smp["ray_indices"].long()
        positions, t_dirs = smp["positions"], rd[ray_idx]
        geo = self.geometry(positions, output_normal=bool(self.material.requires_normal))
        lp = light_positions.reshape(-1, 1, 1, 3).expand(-1, H, W, -1).reshape(-1, 3)[ray_idx] \
            if light_positions is not None else None
        rgb_fg = self.material(viewdirs=t_dirs, positions=positions, light_positions=lp, **geo)
        if bg_color is not None:
            comp_rgb_bg = bg_color.to(dev, torch.float32).reshape(-1, 1, 1, 3).expand(B, H, W, 3)
        else:
            comp_rgb_bg = self.background(dirs=rays_d)
        comp_rgb_bg = comp_rgb_bg.reshape(n_rays, 3)
        acc = R.composite_packed(geo["density"][..., 0], rgb_fg, smp)
        opacity = acc["opacity"][:, None]
        comp_rgb = acc["comp_rgb_fg"] + comp_rgb_bg * (1.0 - opacity)
        res = {"comp_rgb": comp_rgb.view(B, H, W, 3), "comp_rgb_fg": acc["comp_rgb_fg"].view(B, H, W, 3),
               "comp_rgb_bg": comp_rgb_bg.view(B, H, W, 3), "opacity": opacity.view(B, H, W, 1),
               "depth": acc["depth"].view(B, H, W, 1), "z_variance": acc["z_variance"].view(B, H, W, 1)}
        if self.training:
            res.update({"weights": acc["weights"][:, None], "t_points": (0.5 * (smp["t_starts"] + smp["t_ends"]))[:, None],
                        "t_intervals": (smp["t_ends"] - smp["t_starts"])[:, None], "t_dirs": t_dirs,
                        "ray_indices": ray_idx, "points": positions, **geo})
        return res

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False) -> None:
        """nerfacc OccGridEstimator.update_every_n_steps(step, occ_eval_fn = sigma * step_size, occ_thre=0.01,
        ema_decay=0.95, warmup_steps=256, n=16) (nerf_volume_renderer.py:430-444)."""
        if not (self.cfg.grid_prune and self.training and not on_load_weights) or global_step % 16 != 0:
            return
        dev = self.geometry.bbox.device
        occ = self._occ_grid(dev)
        n_cells = self.OCC_RES ** 3
        if global_step < 256:
            idx = torch.arange(n_cells, device=dev, dtype=torch.int32)
        else:
            n = n_cells // 4
            uniform = torch.randint(n_cells, (n,), device=dev)
            occupied = torch.nonzero(occ.binaries().flatten())[:, 0]
            if occupied.numel() > n:
                occupied = occupied[torch.randint(occupied.numel(), (n,), device=dev)]
            idx = torch.cat([uniform, occupied]).to(torch.int32)
        rand = torch.rand(idx.numel(), 3, device=dev)
        if not getattr(self.geometry, "fusable", False):
            res, r = self.OCC_RES, self.cfg.radius
            i = idx.long()
            cell = torch.stack([i // (res * res), (i // res) % res, i % res], -1).float()
            with torch.no_grad():
                sigma = self.geometry.forward_density(-r + (cell + rand) * (2.0 * r / res))
            R.occgrid_update_values(occ, idx, sigma * self.render_step_size)
            return
        P = {k: v.detach() for k, v in self._params().items()}
        occ.update(self._spec(), P, idx, rand, self.render_step_size)

    def update_step_end(self, epoch: int, global_step: int) -> None:
        pass

    def train(self, mode=True):
        self.randomized = mode and self.cfg.randomized
        return super().train(mode=mode)

    def eval(self):
        self.randomized = False
        return super().eval()
