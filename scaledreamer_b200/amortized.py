"""Plugins of the amortized (multi-prompt) generator path, threestudio names / Config keys / attribute names:

  "Hyper-iNGP"                                   custom/amortized/models/geometry/hyper_iNGP.py:114
  "multiprompt-neural-hashgrid-environment-map-background"   custom/amortized/models/background/…background.py:17
  "generative-space-volsdf-volume-renderer"      custom/amortized/models/renderers/generative_space_volsdf_volume_renderer.py:37
  "Triplane-transformer-sdf"                     custom/amortized/models/geometry/triplane_transformer.py:20
  "multiprompt-camera-datamodule"                custom/amortized/data/multiprompt.py:166
  "multiprompt-multiview-camera-datamodule"      custom/amortized/data/multiview_multiprompt.py:78
  "stable-diffusion-multi-prompt-processor"      custom/amortized/models/prompt_processors/stable_diffusion_multi_prompt_processor.py
  "multiprompt-radience-field-generator-system"  custom/amortized/systems/multiprompt_radience_field_generator.py:18

Every field evaluation (hash-grid encode + per-prompt 32-64-{1,3} MLPs, forward and backward) runs in the sm_100a
kernels of csrc/hyper_field.cu through `hyper_field()`; importance resampling and VolSDF compositing run in
csrc/volsdf.cu. The hypernetworks themselves (Linear -> LayerNorm -> SiLU -> Linear on a [B, 1024] embedding,
0.3 MFLOP per prompt) are plain torch modules so that their state-dict keys match the reference checkpoints.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import math
import os
import random
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib as L
from . import core
from .core import BaseModule, BaseObject, find, get_rank, parse_structured, register
from .data import (RandomCameraDataModuleConfig, RandomCameraDataset, RandomCameraIterableDataset,
                   RandomMultiviewCameraIterableDataset)
from .fields import DEFAULT_GRID, HashGridEncoding, _TinyMLP, tiny_mlp
from .prompts import DIRECTIONS, PromptProcessorOutput, hash_prompt
from .systems import BaseSystem, binary_cross_entropy


# ------------------------------------------------------------------------------------------------ field op
class _HyperField(torch.autograd.Function):
    """(out_a [B,N], out_b [B,N,3]) = heads of relu(enc(x01) W1[b]) W2[b]; gradients to the table and the weights."""

    @staticmethod
    def forward(ctx, grid_c, pts01, table, w1a, w2a, w1b, w2b):
        lib = L.load()
        B, N = pts01.shape[0], pts01.shape[1]
        dev = pts01.device
        has_a, has_b = w1a is not None, w1b is not None
        cont = lambda t: None if t is None else t.detach().contiguous().float()
        pts01 = pts01.detach().contiguous().float()
        w1a_c, w2a_c, w1b_c, w2b_c = cont(w1a), cont(w2a), cont(w1b), cont(w2b)
        out_a = torch.empty(B, N, device=dev) if has_a else torch.empty(0, device=dev)
        out_b = torch.empty(B, N, 3, device=dev) if has_b else torch.empty(0, device=dev)
        need = any(ctx.needs_input_grad[2:])
        tape = torch.empty(lib.sdb_hyper_field_tape_floats(B, N), device=dev) if need else None
        tbl = table.detach()
        L.check(lib.sdb_hyper_field_forward(C.byref(grid_c), L.ptr(tbl), L.ptr(pts01), B, N, L.ptr(w1a_c), L.ptr(w2a_c),
                                            L.ptr(w1b_c), L.ptr(w2b_c), L.ptr(out_a) if has_a else None,
                                            L.ptr(out_b) if has_b else None, L.ptr(tape), L.stream_ptr()),
                "sdb_hyper_field_forward")
        ctx.grid_c, ctx.has = grid_c, (has_a, has_b)
        ctx.table_shape = table.shape
        ctx.saved = (pts01, w1a_c, w2a_c, w1b_c, w2b_c, tape)
        return out_a, out_b

    @staticmethod
    def backward(ctx, d_a, d_b):
        lib = L.load()
        pts01, w1a, w2a, w1b, w2b, tape = ctx.saved
        has_a, has_b = ctx.has
        B, N = pts01.shape[0], pts01.shape[1]
        dev = pts01.device
        d_a = d_a.contiguous().float() if (has_a and d_a is not None) else None
        d_b = d_b.contiguous().float() if (has_b and d_b is not None) else None
        g_table = torch.zeros(ctx.table_shape, device=dev)
        z = lambda t: torch.zeros_like(t) if t is not None else None
        g1a, g2a, g1b, g2b = z(w1a), z(w2a), z(w1b), z(w2b)
        L.check(lib.sdb_hyper_field_backward(C.byref(ctx.grid_c), L.ptr(pts01), B, N, L.ptr(w1a), L.ptr(w2a), L.ptr(w1b),
                                             L.ptr(w2b), L.ptr(tape), L.ptr(d_a), L.ptr(d_b), L.ptr(g_table), L.ptr(g1a),
                                             L.ptr(g2a), L.ptr(g1b), L.ptr(g2b), L.stream_ptr()),
                "sdb_hyper_field_backward")
        ctx.saved = None
        return None, None, g_table, g1a, g2a, g1b, g2b


def hyper_field(grid_cfg: dict, pts01: torch.Tensor, table: torch.Tensor, head_a=None, head_b=None):
    """pts01 [B,N,3] in [0,1]; head = (W1 [B,32,64], W2 [B,64,k]) or None -> (out_a [B,N] | None, out_b [B,N,3] | None)."""
    w1a, w2a = head_a if head_a is not None else (None, None)
    w1b, w2b = head_b if head_b is not None else (None, None)
    a, b = _HyperField.apply(L.grid_cfg_c(grid_cfg), pts01, table, w1a, w2a, w1b, w2b)
    return (a if head_a is not None else None), (b if head_b is not None else None)


# ------------------------------------------------------------------------------------------------ hypernetwork
class _HyperNet(torch.autograd.Function):
    """out = W1 silu(LayerNorm(W0 x)) + b1 (sdb_hypernet_forward / sdb_hypernet_backward): one launch forward, two
    backward, fixed summation order. x (the text embedding) receives no gradient, as in the reference."""

    @staticmethod
    def forward(ctx, x, w0, ln_w, ln_b, w1, b1, eps):
        lib = L.load()
        B, c_dim = x.shape
        n_out = w1.shape[0]
        w0c, w1c = w0.detach().contiguous(), w1.detach().contiguous()
        hidden = torch.empty(B, 64, device=x.device)
        out = torch.empty(B, n_out, device=x.device)
        L.check(lib.sdb_hypernet_forward(L.ptr(x), B, c_dim, L.ptr(w0c), L.ptr(ln_w.detach()), L.ptr(ln_b.detach()),
                                         float(eps), L.ptr(w1c), L.ptr(b1.detach()) if b1 is not None else None, n_out,
                                         L.ptr(hidden), L.ptr(out), L.stream_ptr()), "sdb_hypernet_forward")
        ctx.save_for_backward(x, ln_w.detach(), ln_b.detach(), w1c, hidden)
        ctx.eps, ctx.has_bias, ctx.w0_shape = float(eps), b1 is not None, tuple(w0.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = L.load()
        x, ln_w, ln_b, w1, hidden = ctx.saved_tensors
        B, c_dim = x.shape
        n_out = w1.shape[0]
        dev = x.device
        g_w0, g_w1 = torch.empty(ctx.w0_shape, device=dev), torch.empty_like(w1)
        g_lw, g_lb = torch.empty(64, device=dev), torch.empty(64, device=dev)
        g_b1 = torch.empty(n_out, device=dev) if ctx.has_bias else None
        scratch = torch.empty(lib.sdb_hypernet_scratch_floats(B, n_out), device=dev)
        L.check(lib.sdb_hypernet_backward(L.ptr(x), B, c_dim, L.ptr(ln_w), L.ptr(ln_b), ctx.eps, L.ptr(w1), n_out,
                                          L.ptr(hidden), L.ptr(d_out.contiguous().float()), L.ptr(g_w0), L.ptr(g_lw),
                                          L.ptr(g_lb), L.ptr(g_w1), L.ptr(g_b1), L.ptr(scratch), L.stream_ptr()),
                "sdb_hypernet_backward")
        return None, g_w0, g_lw, g_lb, g_w1, g_b1, None


class LinearHyperNetwork(nn.Module):
    """hyper_iNGP.py:18-111: text embedding -> flat weight vector, split into [in, out] matrices per head."""

    def __init__(self, n_input_dims: int, config: dict):
        super().__init__()
        self.c_dim = int(config["c_dim"])
        self.out_dims: Dict[str, List[int]] = {}
        for key, val in dict(config.get("out_dims", {"sdf_weights": [64, 1], "feature_weights": [64, 3]})).items():
            self.out_dims[key] = [n_input_dims] + (list(val) if isinstance(val, (list, tuple)) else [val])
        if config.get("spectral_norm", False):
            raise NotImplementedError("spectral_norm hypernetworks are not implemented")
        if config.get("output_activation", None) not in (None, "none"):
            raise NotImplementedError("hypernetwork output_activation is not implemented")
        self.n_output_dims = sum(i * o for ch in self.out_dims.values() for i, o in zip(ch[:-1], ch[1:]))
        n, hl = int(config["n_neurons"]), int(config["n_hidden_layers"])
        layers: List[nn.Module] = [self._linear(self.c_dim, n, False), nn.LayerNorm(n), nn.SiLU(inplace=True)]
        for _ in range(hl - 1):
            layers += [self._linear(n, n, True), nn.LayerNorm(n), nn.SiLU(inplace=True)]
        layers += [self._linear(n, self.n_output_dims, True)]
        self.layers = nn.Sequential(*layers)

    @staticmethod
    def _linear(i, o, bias):
        layer = nn.Linear(i, o, bias=bias)
        if bias:
            nn.init.zeros_(layer.bias)
        nn.init.xavier_normal_(layer.weight, gain=1.0)
        return layer

    def _native_layout(self, x: torch.Tensor) -> bool:
        """Linear(no bias) -> LayerNorm(64) -> SiLU -> Linear: the layout of every BASELINE config (csrc/hypernet.cu)."""
        return (x.is_cuda and x.dim() == 2 and len(self.layers) == 4 and self.layers[0].out_features == 64
                and 0 < x.shape[0] <= 64 and x.shape[1] * 4 <= 48 * 1024)

    def forward(self, x: torch.Tensor) -> Dict[str, List[torch.Tensor]]:
        if self._native_layout(x):
            lin0, ln, _, lin1 = self.layers
            out = _HyperNet.apply(x.float().contiguous(), lin0.weight, ln.weight, ln.bias, lin1.weight, lin1.bias,
                                  float(ln.eps))
        else:  # deeper hypernetworks (n_hidden_layers > 1) and CPU inspection: torch modules
            out = self.layers(x.float())
        res, start = {}, 0
        for name, ch in self.out_dims.items():
            mats = []
            for i, o in zip(ch[:-1], ch[1:]):
                mats.append(out[:, start:start + i * o].reshape(*x.shape[:-1], i, o))
                start += i * o
            res[name] = mats
        return res


def _check_head(mats, what):
    if len(mats) != 2 or mats[0].shape[-2:] != (32, 64) or mats[1].shape[-2] != 64:
        raise NotImplementedError(f"{what}: the sm_100a field kernels evaluate 32-64-k heads (got "
                                  f"{[tuple(m.shape) for m in mats]})")
    return mats[0], mats[1]


# ------------------------------------------------------------------------------------------------ geometry
@register("Hyper-iNGP")
class HypernetSdf(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        radius: float = 1.0
        isosurface: bool = True
        isosurface_method: str = "mt"
        isosurface_resolution: int = 128
        isosurface_threshold: Union[float, str] = 0.0
        isosurface_chunk: int = 0
        isosurface_coarse_to_fine: bool = True
        isosurface_deformable_grid: bool = False
        isosurface_remove_outliers: bool = False
        isosurface_outlier_n_faces_threshold: Union[int, float] = 0.01
        n_input_dims: int = 3
        n_feature_dims: int = 3
        hypernet_config: dict = field(default_factory=lambda: {
            "c_dim": 768, "out_dims": {"sdf_weights": [64, 1], "feature_weights": [64, 3]}, "spectral_norm": False,
            "n_neurons": 64, "n_hidden_layers": 1, "output_activation": None})
        pos_encoding_config: dict = field(default_factory=lambda: dict(DEFAULT_GRID))
        backbone: str = "linear_hypernetwork"
        normal_type: Optional[str] = "finite_difference"
        finite_difference_normal_eps: Union[float, str] = 0.01
        shape_init: Optional[str] = None
        shape_init_params: Optional[Any] = None
        shape_init_mesh_up: str = "+z"
        shape_init_mesh_front: str = "+x"
        force_shape_init: bool = False
        sdf_bias: Union[float, str] = 0.0
        sdf_bias_params: Optional[Any] = None

    cfg: Config

    def configure(self) -> None:
        r = self.cfg.radius
        self.register_buffer("bbox", torch.as_tensor([[-r, -r, -r], [r, r, r]], dtype=torch.float32))
        self.unbounded = False
        if self.cfg.backbone != "linear_hypernetwork":
            raise NotImplementedError(f"backbone {self.cfg.backbone}")
        if self.cfg.normal_type != "finite_difference":
            raise NotImplementedError(f"normal_type == {self.cfg.normal_type} is not implemented yet.")
        if self.cfg.n_feature_dims != 3:
            raise NotImplementedError("Hyper-iNGP: n_feature_dims must be 3")
        if self.cfg.isosurface_deformable_grid:
            raise NotImplementedError("isosurface_deformable_grid is outside the ASD hot path")
        self.encoding = HashGridEncoding(self.cfg.n_input_dims, self.cfg.pos_encoding_config)
        if self.encoding.n_output_dims != 32:
            raise NotImplementedError("Hyper-iNGP: the field kernels need a 16-level x 2-feature encoding")
        self.hypernet = LinearHyperNetwork(self.encoding.n_output_dims, self.cfg.hypernet_config)
        self.finite_difference_normal_eps: Optional[float] = None

    def initialize_shape(self) -> None:
        if self.cfg.shape_init is None and not self.cfg.force_shape_init:
            return
        if self.cfg.weights is not None and not self.cfg.force_shape_init:
            return
        raise NotImplementedError  # as in the reference (hyper_iNGP.py:204)

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False):
        if not isinstance(self.cfg.finite_difference_normal_eps, float):
            raise NotImplementedError("progressive finite_difference_normal_eps is not implemented yet.")
        self.finite_difference_normal_eps = self.cfg.finite_difference_normal_eps

    def generate_space_cache(self, styles=None, text_embed: Optional[torch.Tensor] = None):
        return self.hypernet(text_embed)  # noises are not used by the hypernetwork

    def _sdf_bias(self, points: torch.Tensor) -> Union[float, torch.Tensor]:
        if self.cfg.sdf_bias == "sphere":
            assert isinstance(self.cfg.sdf_bias_params, float)
            return points.norm(dim=-1) - self.cfg.sdf_bias_params
        if self.cfg.sdf_bias == "ellipsoid":
            size = torch.as_tensor(self.cfg.sdf_bias_params).to(points)
            return ((points / size) ** 2).sum(-1).sqrt() - 1.0
        if isinstance(self.cfg.sdf_bias, (int, float)):
            return float(self.cfg.sdf_bias)
        raise ValueError(f"Unknown sdf bias {self.cfg.sdf_bias}")

    def _contract(self, points: torch.Tensor) -> torch.Tensor:
        return (points - self.bbox[0]) / (self.bbox[1] - self.bbox[0])

    def forward_sdf(self, points: torch.Tensor, space_cache: Dict) -> torch.Tensor:
        """points [B, ..., 3] -> sdf [B, ..., 1] (hyper_iNGP.py:324-349)."""
        B = points.shape[0]
        flat = points.reshape(B, -1, 3)
        a, _ = hyper_field(self.encoding.grid_cfg, self._contract(flat), self.encoding.table.view(-1, 2),
                           head_a=_check_head(space_cache["sdf_weights"], "sdf_weights"))
        return (a + self._sdf_bias(flat)).view(*points.shape[:-1], 1)

    def forward(self, points: torch.Tensor, space_cache: Dict, output_normal: bool = False) -> Dict[str, torch.Tensor]:
        B, N, _ = points.shape
        a, b = hyper_field(self.encoding.grid_cfg, self._contract(points), self.encoding.table.view(-1, 2),
                           head_a=_check_head(space_cache["sdf_weights"], "sdf_weights"),
                           head_b=_check_head(space_cache["feature_weights"], "feature_weights"))
        sdf = a + self._sdf_bias(points)
        out = {"sdf": sdf.reshape(B * N, 1), "features": b.reshape(B * N, 3)}
        if output_normal:
            assert self.finite_difference_normal_eps is not None
            eps = self.finite_difference_normal_eps
            offs = (points[..., None, :] + eps * torch.eye(3, device=points.device)).clamp(-self.cfg.radius,
                                                                                           self.cfg.radius)
            sdf_off = self.forward_sdf(offs.reshape(B, N * 3, 3), space_cache).view(B, N, 3)
            sdf_grad = (sdf_off - sdf[..., None]) / eps
            normal = F.normalize(sdf_grad, dim=-1)
            out.update(normal=normal.reshape(B * N, 3), shading_normal=normal.reshape(B * N, 3),
                       sdf_grad=sdf_grad.reshape(B * N, 3))
        return out

    def train(self, mode=True):
        return super().train(mode)


# ------------------------------------------------------------------------------------------------ triplane
class _TriplaneSample(torch.autograd.Function):
    """enc [B, N, 3C] = bilinear samples of planes_cl [B, 3, H, W, C] (channels-last) at points [B, N, 3] in [-1, 1]."""

    @staticmethod
    def forward(ctx, planes_cl, points):
        lib = L.load()
        B, _, H, W, Cc = planes_cl.shape
        N = points.shape[1]
        planes_cl, points = planes_cl.contiguous().float(), points.detach().contiguous().float()
        enc = torch.empty(B, N, 3 * Cc, device=points.device)
        L.check(lib.sdb_triplane_sample_forward(L.ptr(planes_cl.detach()), L.ptr(points), B, N, H, W, Cc, L.ptr(enc),
                                                L.stream_ptr()), "sdb_triplane_sample_forward")
        ctx.save_for_backward(points)
        ctx.shape = (B, H, W, Cc, N)
        return enc

    @staticmethod
    def backward(ctx, d_enc):
        (points,) = ctx.saved_tensors
        B, H, W, Cc, N = ctx.shape
        d_planes = torch.zeros(B, 3, H, W, Cc, device=points.device)
        L.check(L.load().sdb_triplane_sample_backward(L.ptr(d_enc.contiguous().float()), L.ptr(points), B, N, H, W, Cc,
                                                      L.ptr(d_planes), L.stream_ptr()), "sdb_triplane_sample_backward")
        return d_planes, None


class _Attention(nn.Module):
    """diffusers.models.attention_processor.Attention as the reference instantiates it (bias-free q/k/v, biased
    to_out[0], no dropout; triplane_transformer_modules.py:45-53) on torch SDPA."""

    def __init__(self, query_dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        B, Lq, _ = x.shape
        sp = lambda t: t.view(B, t.shape[1], self.heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(sp(self.to_q(x)), sp(self.to_k(ctx)), sp(self.to_v(ctx)))
        return self.to_out[0](o.transpose(1, 2).reshape(B, Lq, -1))


class _BlockCross(nn.Module):  # ConditionModulationBlock (local text: cross-attention to the 77 token embeddings)
    def __init__(self, inner_dim, cond_dim, num_heads, eps, mlp_ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(inner_dim, eps)
        self.cross_attn = _Attention(inner_dim, num_heads, inner_dim // num_heads, cond_dim)
        self.norm2 = nn.LayerNorm(inner_dim, eps)
        self.self_attn = _Attention(inner_dim, num_heads, inner_dim // num_heads, inner_dim)
        self.norm3 = nn.LayerNorm(inner_dim, eps)
        hid = int(inner_dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(inner_dim, hid), nn.GELU(), nn.Dropout(0.0), nn.Linear(hid, inner_dim),
                                 nn.Dropout(0.0))

    def forward(self, x, cond):
        x = x + self.cross_attn(self.norm1(x), cond)
        x = x + self.self_attn(self.norm2(x))
        return x + self.mlp(self.norm3(x))


class _BlockConcat(nn.Module):  # ConditionModulationBlockwoCrossAttn (global text: condition token prepended)
    def __init__(self, inner_dim, cond_dim, num_heads, eps, mlp_ratio):
        super().__init__()
        self.norm2 = nn.LayerNorm(inner_dim, eps)
        self.self_attn = _Attention(inner_dim, num_heads, inner_dim // num_heads, inner_dim)
        self.norm3 = nn.LayerNorm(inner_dim, eps)
        hid = int(inner_dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.GELU(), nn.Linear(inner_dim, hid), nn.GELU(), nn.Linear(hid, inner_dim),
                                 nn.Dropout(0.0))

    def forward(self, x, cond):
        x = torch.cat([cond, x], dim=1)
        x = x + self.self_attn(self.norm2(x))
        x = x + self.mlp(self.norm3(x))
        return x[:, 1:, :]


class TriplaneTransformer(nn.Module):
    """custom/amortized/extern/triplane_transformer_modules.py:115-187 (state-dict compatible). A TRAINED dense network
    (SURVEY.md §8f rank 1). With `local_text: true` (the C5 yaml) forward and backward run on this library's kernels
    (triplane_native.py: tcgen05 kind::tf32 GEMMs + fp32 LayerNorm / softmax / GELU); `forward_torch` is the same network
    in plain torch, kept as the comparison for the tests and for the `local_text: false` variant no benchmark selects."""

    def __init__(self, inner_dim: int, condition_dim: int, triplane_low_res: int, triplane_high_res: int,
                 triplane_dim: int, num_layers: int, num_heads: int, local_text: bool, mlp_ratio: float = 4.0,
                 eps: float = 1e-6, flash_attention: bool = False):
        super().__init__()
        self.triplane_low_res, self.triplane_high_res, self.triplane_dim = triplane_low_res, triplane_high_res, triplane_dim
        self.pos_embed = nn.Parameter(torch.randn(1, 3 * triplane_low_res ** 2, inner_dim) * (1.0 / inner_dim) ** 0.5)
        self.needs_local_text = local_text
        blk = _BlockCross if local_text else _BlockConcat
        self.layers = nn.ModuleList([blk(inner_dim, condition_dim, num_heads, eps, mlp_ratio) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(inner_dim, eps=eps)
        self.deconv = nn.ConvTranspose2d(inner_dim, triplane_dim, kernel_size=2, stride=2, padding=0, bias=False)
        if not local_text:
            self.proj = nn.Linear(condition_dim, inner_dim)

    def forward(self, text_embed: torch.Tensor) -> torch.Tensor:
        if self.needs_local_text:
            from .triplane_native import generate_planes

            return generate_planes(self, text_embed)  # raises off-CUDA: no CPU path
        return self.forward_torch(text_embed)

    def forward_torch(self, text_embed: torch.Tensor) -> torch.Tensor:
        N, H = text_embed.shape[0], self.triplane_low_res
        if not self.needs_local_text:
            text_embed = self.proj(text_embed).unsqueeze(1)
        x = self.pos_embed.repeat(N, 1, 1)
        for layer in self.layers:
            x = layer(x, text_embed)
        x = self.norm(x).view(N, 3, H, H, -1)
        x = torch.einsum("nihwd->indhw", x).contiguous().view(3 * N, -1, H, H)
        x = self.deconv(x)
        x = x.view(3, N, *x.shape[-3:])
        return torch.einsum("indhw->nidhw", x).contiguous()  # [N, 3, C, 2H, 2H]


@register("Triplane-transformer-sdf")
class TriplaneTransformerSDF(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        radius: float = 1.0
        isosurface: bool = True
        isosurface_method: str = "mt"
        isosurface_resolution: int = 128
        isosurface_threshold: Union[float, str] = 0.0
        isosurface_chunk: int = 0
        isosurface_coarse_to_fine: bool = True
        isosurface_deformable_grid: bool = False
        isosurface_remove_outliers: bool = False
        isosurface_outlier_n_faces_threshold: Union[int, float] = 0.01
        n_feature_dims: int = 3
        space_generator_config: dict = field(default_factory=lambda: {
            "inner_dim": 768, "condition_dim": 1024, "triplane_low_res": 32, "triplane_high_res": 64, "triplane_dim": 32,
            "num_layers": 12, "num_heads": 16, "flash_attention": False, "local_text": False})
        mlp_network_config: dict = field(default_factory=lambda: {
            "otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none", "n_neurons": 64,
            "n_hidden_layers": 2})
        backbone: str = "triplane_transformer"
        normal_type: Optional[str] = "finite_difference"
        finite_difference_normal_eps: Union[float, str] = 0.01
        sdf_bias: Union[float, str] = 0.0
        sdf_bias_params: Optional[Any] = None

    cfg: Config

    def configure(self) -> None:
        from .fields import VanillaMLP

        r = self.cfg.radius
        self.register_buffer("bbox", torch.as_tensor([[-r, -r, -r], [r, r, r]], dtype=torch.float32))
        self.unbounded = False
        if self.cfg.backbone != "triplane_transformer":
            raise ValueError(f"Unknown backbone {self.cfg.backbone}")
        if self.cfg.normal_type != "finite_difference":
            raise NotImplementedError(f"normal_type == {self.cfg.normal_type} is not implemented yet.")
        if self.cfg.isosurface_deformable_grid:
            raise NotImplementedError("isosurface_deformable_grid is outside the ASD hot path")
        self.space_generator = TriplaneTransformer(**self.cfg.space_generator_config)
        input_dim = int(self.cfg.space_generator_config["triplane_dim"]) * 3
        self.sdf_network = VanillaMLP(input_dim, 1, self.cfg.mlp_network_config)
        self.feature_network = VanillaMLP(input_dim, self.cfg.n_feature_dims, self.cfg.mlp_network_config)
        self.finite_difference_normal_eps: Optional[float] = None

    def initialize_shape(self) -> None:
        pass

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False):
        if not isinstance(self.cfg.finite_difference_normal_eps, float):
            raise NotImplementedError("progressive finite_difference_normal_eps is not implemented yet.")
        self.finite_difference_normal_eps = self.cfg.finite_difference_normal_eps

    def generate_space_cache(self, styles=None, text_embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.space_generator(text_embed=text_embed)

    _sdf_bias = HypernetSdf._sdf_bias

    def interpolate_encodings(self, points: torch.Tensor, space_cache: torch.Tensor) -> torch.Tensor:
        """points [B, N, 3] in [-1, 1], space_cache [B, 3, C, H, W] -> [B, N, 3C] (utils.py:80-97; box_warp = 2)."""
        planes_cl = space_cache.permute(0, 1, 3, 4, 2).contiguous()
        return _TriplaneSample.apply(planes_cl, points)

    def _contract(self, points: torch.Tensor) -> torch.Tensor:  # scale_tensor(x, bbox, (-1, 1))
        return (points - self.bbox[0]) / (self.bbox[1] - self.bbox[0]) * 2.0 - 1.0

    def forward_sdf(self, points: torch.Tensor, space_cache: torch.Tensor) -> torch.Tensor:
        B = points.shape[0]
        flat = points.reshape(B, -1, 3)
        enc = self.interpolate_encodings(self._contract(flat), space_cache)
        sdf = tiny_mlp(self.sdf_network, enc)[..., 0] + self._sdf_bias(flat)
        return sdf.view(*points.shape[:-1], 1)

    def forward(self, points: torch.Tensor, space_cache: torch.Tensor, output_normal: bool = False):
        B, N, _ = points.shape
        enc = self.interpolate_encodings(self._contract(points), space_cache)
        sdf = tiny_mlp(self.sdf_network, enc)[..., 0] + self._sdf_bias(points)
        out = {"sdf": sdf.reshape(B * N, 1), "features": tiny_mlp(self.feature_network, enc).reshape(B * N, -1)}
        if output_normal:
            assert self.finite_difference_normal_eps is not None
            eps = self.finite_difference_normal_eps
            offs = (points[..., None, :] + eps * torch.eye(3, device=points.device)).clamp(-self.cfg.radius,
                                                                                           self.cfg.radius)
            sdf_off = self.forward_sdf(offs.reshape(B, N * 3, 3), space_cache).view(B, N, 3)
            sdf_grad = (sdf_off - sdf[..., None]) / eps
            normal = F.normalize(sdf_grad, dim=-1)
            out.update(normal=normal.reshape(B * N, 3), shading_normal=normal.reshape(B * N, 3),
                       sdf_grad=sdf_grad.reshape(B * N, 3))
        return out


# ------------------------------------------------------------------------------------------------ background
@register("multiprompt-neural-hashgrid-environment-map-background")
class MultipromptNeuralHashgridEnvironmentMapBackground(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        n_output_dims: int = 3
        color_activation: str = "sigmoid"
        pos_encoding_config: dict = field(default_factory=lambda: {
            "otype": "HashGrid", "n_levels": 8, "n_features_per_level": 2, "log2_hashmap_size": 19,
            "base_resolution": 4, "per_level_scale": 1.8114473285278132})
        hypernet_config: dict = field(default_factory=lambda: {
            "c_dim": 1024, "out_dims": {"bg_weights": [64, 3]}, "spectral_norm": False, "n_neurons": 64,
            "n_hidden_layers": 1, "output_activation": None})
        random_aug: bool = False
        random_aug_prob: float = 0.5
        eval_color: Optional[Tuple[float, float, float]] = None

    cfg: Config

    def configure(self) -> None:
        self.encoding = HashGridEncoding(3, self.cfg.pos_encoding_config)
        if self.encoding.n_output_dims != 32:
            raise NotImplementedError("the field kernels need a 16-level x 2-feature encoding for the environment map")
        if self.cfg.color_activation != "sigmoid" or self.cfg.n_output_dims != 3:
            raise NotImplementedError("environment map: sigmoid RGB only")
        self.hypernet = LinearHyperNetwork(self.encoding.n_output_dims, self.cfg.hypernet_config)
        self.enabling_hypernet = True

    def forward(self, dirs: torch.Tensor, text_embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, H, W, _ = dirs.shape
        if not self.training and self.cfg.eval_color is not None:
            return torch.ones(B, H, W, 3, device=dirs.device) * torch.as_tensor(self.cfg.eval_color).to(dirs)
        bg_cache = self.hypernet(text_embed)
        _, col = hyper_field(self.encoding.grid_cfg, ((dirs + 1.0) / 2.0).reshape(B, H * W, 3),
                             self.encoding.table.view(-1, 2), head_b=_check_head(bg_cache["bg_weights"], "bg_weights"))
        color = torch.sigmoid(col).view(B, H, W, 3)
        if self.training and self.cfg.random_aug and random.random() < self.cfg.random_aug_prob:
            color = color * 0 + torch.rand(B, 1, 1, 3, device=dirs.device).expand(B, H, W, 3)
        return color


# ------------------------------------------------------------------------------------------------ renderer
def volsdf_density(sdf: torch.Tensor, inv_std: torch.Tensor) -> torch.Tensor:
    """neus_volume_renderer.py:19-23"""
    inv_std = inv_std.clamp(0.0, 80.0)
    return inv_std * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() * inv_std))


class LearnedVariance(nn.Module):
    def __init__(self, init_val, requires_grad=True):
        super().__init__()
        self.register_parameter("_inv_std", nn.Parameter(torch.tensor(float(init_val)), requires_grad=requires_grad))

    @property
    def inv_std(self):
        return torch.exp(self._inv_std * 10.0)

    def forward(self, x):
        return torch.ones_like(x) * self.inv_std.clamp(1.0e-6, 1.0e6)


class _VolSDFComposite(torch.autograd.Function):
    """Dense [Nr, S] VolSDF compositing (get_alpha + render_weight_from_alpha + the five accumulate_along_rays of
    generative_space_volsdf_volume_renderer.py:356-397) in one launch each way (csrc/volsdf.cu)."""

    @staticmethod
    def forward(ctx, sdf, feat, normal, t_mid, delta, inv_std: float, color_act: int = 0):
        lib = L.load()
        Nr, S = sdf.shape
        dev = sdf.device
        sdf, feat, normal = sdf.contiguous().float(), feat.contiguous().float(), normal.detach().contiguous().float()
        t_mid, delta = t_mid.contiguous().float(), delta.contiguous().float()
        weights = torch.empty(Nr, S, device=dev)
        opacity, depth, zvar = (torch.empty(Nr, device=dev) for _ in range(3))
        fg, cn = torch.empty(Nr, 3, device=dev), torch.empty(Nr, 3, device=dev)
        L.check(lib.sdb_volsdf_composite_forward(L.ptr(sdf), L.ptr(feat), L.ptr(normal), L.ptr(t_mid), L.ptr(delta), Nr,
                                                 S, float(inv_std), int(color_act), L.ptr(weights), L.ptr(opacity), L.ptr(depth),
                                                 L.ptr(fg), L.ptr(zvar), L.ptr(cn), L.stream_ptr()),
                "sdb_volsdf_composite_forward")
        ctx.save_for_backward(sdf, feat, t_mid, delta, weights, opacity, depth, fg)
        ctx.inv_std, ctx.color_act = float(inv_std), int(color_act)
        ctx.mark_non_differentiable(weights, zvar, cn)
        return fg, opacity, depth, zvar, weights, cn

    @staticmethod
    def backward(ctx, g_fg, g_op, g_depth, g_zvar, g_w, g_cn):
        lib = L.load()
        sdf, feat, t_mid, delta, weights, opacity, depth, fg = ctx.saved_tensors
        Nr, S = sdf.shape
        c = lambda g, like: (g.contiguous().float() if g is not None else torch.zeros_like(like))
        g_fg, g_op, g_depth = c(g_fg, fg), c(g_op, opacity), c(g_depth, depth)
        d_sdf, d_feat = torch.empty_like(sdf), torch.empty_like(feat)
        L.check(lib.sdb_volsdf_composite_backward(L.ptr(sdf), L.ptr(feat), L.ptr(t_mid), L.ptr(delta), L.ptr(weights),
                                                  L.ptr(opacity), L.ptr(depth), L.ptr(fg), L.ptr(g_fg), L.ptr(g_op),
                                                  L.ptr(g_depth), Nr, S, ctx.inv_std, ctx.color_act, L.ptr(d_sdf),
                                                  L.ptr(d_feat), L.stream_ptr()), "sdb_volsdf_composite_backward")
        return d_sdf, d_feat, None, None, None, None, None


@register("generative-space-volsdf-volume-renderer")
class GenerativeSpaceVolSDFVolumeRenderer(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        radius: float = 1.0
        num_samples_per_ray: int = 512
        randomized: bool = True
        eval_chunk_size: int = 320000
        learned_variance_init: float = 0.3
        cos_anneal_end_steps: int = 0
        use_volsdf: bool = False
        near_plane: float = 0.0
        far_plane: float = 1e10
        trainable_variance: bool = True
        estimator: str = "occgrid"
        grid_prune: bool = True
        prune_alpha_threshold: bool = True
        num_samples_per_ray_importance: int = 64
        train_chunk_size: int = 0

    cfg: Config

    def configure(self, geometry, material, background) -> None:
        @dataclass
        class SubModules:  # kept out of nn.Module registration (renderers/base.py:28-35)
            geometry: Any
            material: Any
            background: Any

        self.sub_modules = SubModules(geometry, material, background)
        r = self.cfg.radius
        self.register_buffer("bbox", torch.as_tensor([[-r, -r, -r], [r, r, r]], dtype=torch.float32))
        self.variance = LearnedVariance(self.cfg.learned_variance_init, requires_grad=self.cfg.trainable_variance)
        if self.cfg.estimator == "occgrid":
            raise NotImplementedError("Occgrid estimator not supported for generative-space-volsdf-volume-renderer")
        if self.cfg.estimator != "importance":
            raise NotImplementedError(f"Estimator {self.cfg.estimator} not implemented")
        if not self.cfg.use_volsdf:
            raise ValueError("Currently only VolSDF supports importance sampling.")
        if self.cfg.trainable_variance:
            raise NotImplementedError("trainable_variance=true: the compositing kernels take inv_std as a constant "
                                      "(the multi-prompt configs freeze it)")
        self.randomized = self.cfg.randomized
        # Training memory: the per-sample tensors of one step (encodings, MLP activations, their gradients) grow with
        # rays x 193 intervals x 4 field evaluations. `train_chunk_size` (the reference's key, rays per chunk; its
        # chunk_batch_original at generative_space_volsdf_volume_renderer.py:241-250) bounds them: rays are rendered in
        # chunks whose intermediates are dropped after the forward and recomputed chunk by chunk in the backward. With
        # 0 (the yaml default) the chunk is chosen so that a chunk's intermediates stay under CHUNK_BUDGET_BYTES; a
        # batch that fits is rendered in one piece (BASELINE C4 does, C5 at 256 x 256 x 4 views needs ~80 GB and does not).
        self.last_chunk_rays = 0

    geometry = property(lambda self: self.sub_modules.geometry)
    material = property(lambda self: self.sub_modules.material)
    background = property(lambda self: self.sub_modules.background)

    def sample_intervals(self, rays_o, rays_d, rays_per_prompt, space_cache, u_coarse=None, u_fine=None):
        """ImportanceEstimator.sampling (estimators.py:23-101) with one VolSDF proposal: -> sorted t [Nr, nc+nf+2]."""
        lib = L.load()
        Nr, dev = rays_o.shape[0], rays_o.device
        B = Nr // rays_per_prompt
        nc, nf = self.cfg.num_samples_per_ray_importance, self.cfg.num_samples_per_ray
        near, far = float(self.cfg.near_plane), float(self.cfg.far_plane)
        with torch.no_grad():
            if u_coarse is None:
                u_coarse = torch.rand(Nr, device=dev) if self.randomized else torch.full((Nr,), 0.5, device=dev)
            if u_fine is None:
                u_fine = torch.rand(Nr, device=dev) if self.randomized else torch.full((Nr,), 0.5, device=dev)
            u_coarse, u_fine = u_coarse.contiguous().float(), u_fine.contiguous().float()
            pts = torch.empty(Nr, nc, 3, device=dev)
            L.check(lib.sdb_volsdf_coarse_points(L.ptr(rays_o), L.ptr(rays_d), L.ptr(u_coarse), Nr, nc, near, far,
                                                 L.ptr(pts), L.stream_ptr()), "sdb_volsdf_coarse_points")
            sdf = self.geometry.forward_sdf(pts.view(B, rays_per_prompt * nc, 3), space_cache).reshape(Nr, nc)
            t_all = torch.empty(Nr, nc + nf + 2, device=dev)
            L.check(lib.sdb_volsdf_resample(L.ptr(sdf.contiguous()), L.ptr(u_coarse), L.ptr(u_fine), Nr, nc, nf, near, far,
                                            float(self.variance.inv_std.clamp(1.0e-6, 1.0e6)), L.ptr(t_all),
                                            L.stream_ptr()), "sdb_volsdf_resample")
        return t_all

    def forward(self, rays_o, rays_d, light_positions=None, bg_color=None, noise=None, space_cache=None,
                text_embed=None, u_coarse=None, u_fine=None, **kwargs) -> Dict[str, torch.Tensor]:
        B, H, W = rays_o.shape[:3]
        Bc = text_embed.shape[0] if text_embed is not None else B
        if space_cache is None:
            space_cache = self.geometry.generate_space_cache(styles=noise, text_embed=text_embed)
        if Bc != B and not self.training:
            # inference over an orbit of ONE generated space: one view at a time (the reference's chunk_batch_original
            # with chunk_size 1, generative_space_volsdf_volume_renderer.py:131-157), so an evaluation batch of 120
            # 512x512 views never holds more than one view's samples
            assert Bc == 1, "batch_size of space_cache must be 1 or equal to batch_size of rays_o"
            per_view = lambda t, i, n: t[i * n:(i + 1) * n] if torch.is_tensor(t) else t
            outs = [self.forward(rays_o[i:i + 1], rays_d[i:i + 1],
                                 None if light_positions is None else light_positions[i:i + 1],
                                 bg_color=per_view(bg_color, i, 1) if torch.is_tensor(bg_color) and bg_color.dim() > 1
                                 and bg_color.shape[0] == B else bg_color,
                                 noise=noise, space_cache=space_cache, text_embed=text_embed,
                                 u_coarse=per_view(u_coarse, i, H * W), u_fine=per_view(u_fine, i, H * W))
                    for i in range(B)]
            return {k: torch.cat([o[k] for o in outs], dim=0) for k in outs[0]}
        if Bc != B:
            if not self.training:
                assert Bc == 1, "batch_size of space_cache must be 1 or equal to batch_size of rays_o"
            assert B % Bc == 0
            if torch.is_tensor(space_cache):
                space_cache = space_cache.repeat_interleave(B // Bc, dim=0)
            else:
                space_cache = {k: [m.repeat_interleave(B // Bc, dim=0) for m in v] for k, v in space_cache.items()}
            if text_embed is not None:
                text_embed = text_embed.repeat_interleave(B // Bc, dim=0)
        Nr, HW = B * H * W, H * W
        o, d = rays_o.reshape(Nr, 3).contiguous(), rays_d.reshape(Nr, 3).contiguous()
        dev = o.device
        if u_coarse is None:
            u_coarse = torch.rand(Nr, device=dev) if self.randomized else torch.full((Nr,), 0.5, device=dev)
        if u_fine is None:
            u_fine = torch.rand(Nr, device=dev) if self.randomized else torch.full((Nr,), 0.5, device=dev)
        S = self.cfg.num_samples_per_ray_importance + self.cfg.num_samples_per_ray + 1
        chunk = self._chunk_rays(B, HW, S) if (self.training and torch.is_grad_enabled()) else 0
        self.last_chunk_rays = chunk
        extras = None
        if chunk <= 0 or chunk >= HW:
            fg, opacity, depth, z_var, comp_normal, sdf_grad, extras = self._render_rays(o, d, u_coarse, u_fine,
                                                                                       space_cache, B, True)
        else:
            # rays [B, HW] -> chunks [B, c]: every chunk keeps the batch grouping the geometry's space cache needs
            from torch.utils.checkpoint import checkpoint

            o3, d3 = o.view(B, HW, 3), d.view(B, HW, 3)
            uc2, uf2 = u_coarse.view(B, HW), u_fine.view(B, HW)
            parts = []
            for s0 in range(0, HW, chunk):
                s1 = min(HW, s0 + chunk)
                args = (o3[:, s0:s1].reshape(-1, 3).contiguous(), d3[:, s0:s1].reshape(-1, 3).contiguous(),
                        uc2[:, s0:s1].reshape(-1).contiguous(), uf2[:, s0:s1].reshape(-1).contiguous())
                fn = lambda a, b_, c_, e_: self._render_rays(a, b_, c_, e_, space_cache, B, False)[:6]
                parts.append(checkpoint(fn, *args, use_reentrant=False, preserve_rng_state=False))
            cat = lambda i: torch.cat([p[i].view(B, -1, *p[i].shape[1:]) for p in parts], dim=1)
            fg, opacity, depth, z_var, comp_normal = (cat(i).reshape(Nr, *parts[0][i].shape[1:]) for i in range(5))
            sdf_grad = torch.cat([p[5].view(B, -1, S, 3) for p in parts], dim=1).reshape(-1, 3)
        color_bg_module = self.background
        if getattr(color_bg_module, "enabling_hypernet", False):
            comp_rgb_bg = color_bg_module(dirs=rays_d, text_embed=text_embed)
        else:
            comp_rgb_bg = color_bg_module(dirs=rays_d)
        if bg_color is None:
            bg_color = comp_rgb_bg
        bg_flat = bg_color.reshape(Nr, -1) if bg_color.shape[:-1] == (B, H, W) else bg_color
        comp_rgb = fg + bg_flat * (1.0 - opacity[:, None])
        out = {"comp_rgb": comp_rgb.view(B, H, W, 3), "comp_rgb_fg": fg.view(B, H, W, 3),
               "comp_rgb_bg": comp_rgb_bg.view(B, H, W, 3), "opacity": opacity.view(B, H, W, 1),
               "depth": depth.view(B, H, W, 1), "z_variance": z_var.view(B, H, W, 1),
               "comp_normal": comp_normal.view(B, H, W, 3)}
        if self.training:
            if extras is not None:
                out.update(extras)
            else:  # chunked: the per-sample tensors are not kept (only what the eikonal loss reads)
                out["sdf_grad"] = sdf_grad
            out["inv_std"] = self.variance.inv_std
        return out

    CHUNK_BUDGET_BYTES = 24 << 30

    def _chunk_rays(self, B: int, HW: int, S: int) -> int:
        """Rays per batch element and chunk (0: render everything at once)."""
        if self.cfg.train_chunk_size > 0:
            return max(1, int(self.cfg.train_chunk_size) // B)
        # bytes of intermediates per field evaluation: encodings kept for the MLP backward (fp32 x width), their
        # gradient, hidden-layer recompute scratch and the outputs; the hash-grid field keeps a 128-byte tape instead
        width = 96 if type(self.geometry).__name__ == "TriplaneTransformerSDF" else 32
        per_point = 3 * 4 * width + 96
        total = B * HW * S * 4 * per_point
        if total <= self.CHUNK_BUDGET_BYTES:
            return 0
        c = max(256, int(self.CHUNK_BUDGET_BYTES // (B * S * 4 * per_point)))
        return 1 << (c.bit_length() - 1)  # power of two: equal chunks for the usual image sizes

    def _render_rays(self, o, d, u_coarse, u_fine, space_cache, B: int, want_extras: bool):
        """Importance sampling -> geometry (centres + finite-difference offsets) -> VolSDF compositing for rays grouped
        by batch element ([B * c, 3]); generative_space_volsdf_volume_renderer.py:209-424."""
        Nr = o.shape[0]
        c = Nr // B
        t = self.sample_intervals(o, d, c, space_cache, u_coarse, u_fine)
        t_mid, delta = 0.5 * (t[:, :-1] + t[:, 1:]), t[:, 1:] - t[:, :-1]
        S = t_mid.shape[1]
        positions = o[:, None, :] + d[:, None, :] * t_mid[..., None]
        geo = self.geometry(positions.view(B, c * S, 3), space_cache=space_cache, output_normal=True)
        inv_std = float(self.variance.inv_std.clamp(1.0e-6, 1.0e6))
        color_act = {"sigmoid": 0, "sigmoid-mipnerf": 1}.get(self.material.cfg.color_activation)
        if color_act is None or type(self.material).__name__ != "NoMaterial":
            raise NotImplementedError("the VolSDF compositing kernel fuses no-material with a sigmoid / sigmoid-mipnerf "
                                      "colour activation")
        fg, opacity, depth, z_var, weights, comp_normal = _VolSDFComposite.apply(
            geo["sdf"].view(Nr, S), geo["features"].view(Nr, S, 3), geo["normal"].view(Nr, S, 3), t_mid, delta, inv_std,
            color_act)
        extras = None
        if want_extras and self.training:
            ray_indices = torch.arange(Nr, device=o.device).unsqueeze(-1).expand(-1, S).reshape(-1)
            extras = {"weights": weights.reshape(-1, 1), "t_points": t_mid.reshape(-1, 1),
                      "t_intervals": delta.reshape(-1, 1), "t_dirs": d[ray_indices], "ray_indices": ray_indices,
                      "points": positions.reshape(-1, 3), **geo}
        return fg, opacity, depth, z_var, comp_normal, geo["sdf_grad"], extras

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False) -> None:
        pass

    def train(self, mode=True):
        self.randomized = mode and self.cfg.randomized
        return super().train(mode=mode)

    def eval(self):
        self.randomized = False
        return super().eval()


# ------------------------------------------------------------------------------------------------ data
@dataclass
class MultipromptRandomCameraDataModuleConfig(RandomCameraDataModuleConfig):
    dim_gaussian: int = 512
    prompt_library: str = "magic3d_prompt_library"
    prompt_library_dir: str = "load"
    prompt_library_format: str = "json"
    eval_prompt: Optional[str] = None
    target_prompt: Optional[str] = None
    eval_fix_camera: Optional[int] = None


def load_prompt_library(cfg, rank: int, world: int) -> Dict[str, List[str]]:
    """Each process only keeps every world-th prompt (multiprompt.py:176-186)."""
    path = os.path.join(cfg.prompt_library_dir, cfg.prompt_library) + "." + cfg.prompt_library_format
    with open(path, "r") as f:
        lib = json.load(f)
    return {k: v[rank::world] for k, v in lib.items()}


def _world() -> Tuple[int, int]:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return get_rank(), int(os.environ.get("WORLD_SIZE", "1"))


class MultipromptRandomCameraIterableDataset(RandomCameraIterableDataset):
    config_cls = MultipromptRandomCameraDataModuleConfig

    def __init__(self, cfg: Any, prompt_library: Dict) -> None:
        super().__init__(cfg)
        assert "train" in prompt_library, "prompt library must contain train split"
        self.prompt_library = prompt_library["train"]

    def collate(self, batch=None) -> Dict[str, Any]:
        out = super().collate(batch)
        out["noise"] = torch.randn(self.batch_size, self.cfg.dim_gaussian)
        if len(self.prompt_library) < self.batch_size:
            out["prompt"] = random.choices(self.prompt_library, k=self.batch_size)
        else:
            out["prompt"] = random.sample(self.prompt_library, k=self.batch_size)
        return out


class MultipromptRandomCameraDataset4Test:
    """Evaluation set of the multi-prompt data modules (multiprompt.py:85-122): one batch per prompt of the split holding
    the WHOLE evaluation orbit (all n_val_views / n_test_views cameras) and one noise row; the renderer walks the views
    one at a time against the single generated space."""

    def __init__(self, cfg: Any, split: str, prompt_library: Dict) -> None:
        self.dataset = RandomCameraDataset(cfg, split)
        self.cfg, self.n_views = cfg, self.dataset.n_views
        start_point = torch.randn(cfg.dim_gaussian)
        end_point = torch.randn(cfg.dim_gaussian)
        self.noises = torch.stack([start_point + (end_point - start_point) * i / self.n_views
                                   for i in range(self.n_views)])
        # the reference indexes prompt_library["val"] as the fall-back; a train-only library simply has nothing to evaluate
        self.prompt_library = prompt_library[split] if split in prompt_library else prompt_library.get("val", [])
        self._views: Optional[Dict[str, Any]] = None

    def __len__(self) -> int:
        return len(self.prompt_library)

    def __iter__(self):
        for prompt in self.prompt_library:
            yield {"prompt": [prompt]}

    def collate(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        if self._views is None:
            self._views = self.dataset.collate([self.dataset[i] for i in range(self.n_views)])
        out = dict(self._views)
        out["noise"] = self.noises[0][None, :]
        out.update(batch)
        return out

    def to_device(self, batch: Dict[str, Any], device) -> Dict[str, Any]:
        return self.dataset.to_device(batch, device)


class MultipromptRandomCameraDataset4FixPrompt:
    """`data.eval_prompt=...` (multiprompt.py:125-164): one view per batch, always the same prompt and a zero noise row;
    with `target_prompt` the text embedding is interpolated along the orbit (`ratio` = linspace(0, 1, n_views)), with
    `eval_fix_camera` every batch uses that one camera. Batches are collated with batch size 1 (multiprompt.py:230-234)."""

    def __init__(self, cfg: Any, split: str) -> None:
        self.dataset = RandomCameraDataset(cfg, split)
        self.cfg, self.n_views = cfg, self.dataset.n_views
        self.noise = torch.zeros(cfg.dim_gaussian)
        self.eval_prompt, self.target_prompt = cfg.eval_prompt, cfg.target_prompt
        self.ratios = torch.linspace(0, 1, self.n_views)
        self.fix_camera = cfg.eval_fix_camera

    def __len__(self) -> int:
        return self.n_views

    def __getitem__(self, idx: int) -> Dict[str, Any]:
        item = self.dataset[self.fix_camera] if self.fix_camera else self.dataset[idx]
        item.update(noise=self.noise, prompt=self.eval_prompt, index=idx)
        if self.target_prompt is not None:
            item.update(prompt_target=self.target_prompt, ratio=self.ratios[idx])
        item["name"] = "_to_".join([self.eval_prompt, self.target_prompt]) if self.target_prompt is not None \
            else self.eval_prompt
        return item

    def __iter__(self):
        for idx in range(self.n_views):
            yield self.collate([self[idx]])

    def collate(self, items: List[Dict[str, Any]]) -> Dict[str, Any]:
        strings = {k: [it[k] for it in items] for k in ("prompt", "prompt_target", "name") if k in items[0]}
        out = self.dataset.collate([{k: v for k, v in it.items() if k not in strings} for it in items])
        out.update(strings)
        return out

    def to_device(self, batch: Dict[str, Any], device) -> Dict[str, Any]:
        return self.dataset.to_device(batch, device)


class _MultipromptEvalLoaders:
    """setup / val_dataloader / test_dataloader shared by the two multi-prompt data modules (multiprompt.py:189-238,
    multiview_multiprompt.py:100-111)."""

    train_cls: Any = None

    def setup(self, stage=None) -> None:
        if stage in (None, "fit"):
            self.train_dataset = self.train_cls(self.cfg, self.prompt_library)
        if stage in (None, "fit", "validate"):
            self.val_dataset = MultipromptRandomCameraDataset4Test(self.cfg, "val", self.prompt_library)
        if stage in (None, "test", "predict"):
            if self.cfg.eval_prompt is not None:
                self.test_dataset = MultipromptRandomCameraDataset4FixPrompt(self.cfg, "test")
            else:
                self.test_dataset = MultipromptRandomCameraDataset4Test(self.cfg, "test", self.prompt_library)

    @staticmethod
    def _loader(ds):
        if isinstance(ds, MultipromptRandomCameraDataset4FixPrompt):
            yield from ds
        else:
            for item in ds:
                yield ds.collate(item)

    def val_dataloader(self):
        if getattr(self, "val_dataset", None) is None:
            self.setup("validate")
        return self._loader(self.val_dataset)

    def test_dataloader(self):
        if getattr(self, "test_dataset", None) is None:
            self.setup("test")
        return self._loader(self.test_dataset)

    def train_dataloader(self):
        if self.train_dataset is None:
            self.setup("fit")
        while True:
            yield self.train_dataset.collate({})


@register("multiprompt-camera-datamodule")
class MultipromptCameraDataModule(_MultipromptEvalLoaders):
    train_cls = MultipromptRandomCameraIterableDataset

    def __init__(self, cfg=None) -> None:
        self.cfg = parse_structured(MultipromptRandomCameraDataModuleConfig, cfg)
        rank, world = _world()
        self.prompt_library = load_prompt_library(self.cfg, rank, world)
        self.train_dataset = self.val_dataset = self.test_dataset = None


@dataclass
class MultiviewMultipromptRandomCameraDataModuleConfig(MultipromptRandomCameraDataModuleConfig):
    relative_radius: bool = True
    n_view: int = 1
    zoom_range: Tuple[float, float] = (1.0, 1.0)


class MultiviewMultipromptRandomCameraIterableDataset(RandomMultiviewCameraIterableDataset):
    config_cls = MultiviewMultipromptRandomCameraDataModuleConfig

    def __init__(self, cfg: Any, prompt_library: Dict) -> None:
        super().__init__(cfg)
        assert "train" in prompt_library, "prompt library must contain train split"
        self.prompt_library = prompt_library["train"]
        self.n_view = self.cfg.n_view

    def collate(self, batch=None) -> Dict[str, Any]:
        n_prompts = self.batch_size // self.n_view
        out = super().collate(batch)
        out["noise"] = torch.randn(n_prompts, self.cfg.dim_gaussian)
        if len(self.prompt_library) < n_prompts:
            out["prompt"] = random.choices(self.prompt_library, k=n_prompts)
        else:
            out["prompt"] = random.sample(self.prompt_library, k=n_prompts)
        return out


@register("multiprompt-multiview-camera-datamodule")
class MultiviewMultipromptCameraDataModule(_MultipromptEvalLoaders):
    train_cls = MultiviewMultipromptRandomCameraIterableDataset

    def __init__(self, cfg=None) -> None:
        self.cfg = parse_structured(MultiviewMultipromptRandomCameraDataModuleConfig, cfg)
        rank, world = _world()
        self.prompt_library = load_prompt_library(self.cfg, rank, world)
        self.train_dataset = self.val_dataset = self.test_dataset = None


# ------------------------------------------------------------------------------------------------ prompts
class MultiPromptProcessorOutput:
    """custom/amortized/models/prompt_processors/base.py:410-568. The processor keeps one stacked device table
    [P, 4, 77, 1024] for its (rank-local) prompt library; a batch is a vector of prompt indices, and direction
    selection / Perp-Neg interpolation for the whole batch is ONE launch of sdb_asd_text_embeddings_multi (the
    reference loops over the batch in Python with .item() syncs). Block order as in the reference:
    [pos (B), uncond (B), neg (2B, sample-major)]."""

    def __init__(self, proc: "StableDiffusionMultiPromptProcessor", prompt_idx: torch.Tensor, prompts: List[str]):
        self.proc, self.prompt_idx, self.prompts = proc, prompt_idx, prompts
        self.use_perp_neg = bool(proc.cfg.use_perp_neg)

    def get_global_text_embeddings(self) -> torch.Tensor:
        """[B, 1024] pooled embeddings, or the [B, 77, 1024] token embeddings when use_local_text_embeddings (:426-432)."""
        if self.proc.cfg.use_local_text_embeddings:
            return self.proc.local_table[self.prompt_idx.long(), 0].float()
        return self.proc.global_table[self.prompt_idx.long()]

    def prompt_cfg_c(self, view_dependent: bool, perp_neg: bool) -> L.PromptCfgC:
        cfg = self.proc.cfg
        pc = L.PromptCfgC()
        pc.view_dependent, pc.perp_neg = int(view_dependent), int(perp_neg)
        pc.front_threshold, pc.back_threshold = float(cfg.front_threshold), float(cfg.back_threshold)
        pc.overhead_threshold = float(cfg.overhead_threshold)
        for name in ("f_sb", "f_fsb", "f_fs", "f_sf"):
            arr = getattr(pc, name)
            for i, v in enumerate(getattr(cfg, "perp_neg_" + name)):
                arr[i] = float(v)
        pc.neg_scale = 1.0
        return pc

    def _idx_for(self, n: int) -> torch.Tensor:
        reps = n // self.prompt_idx.shape[0]  # several views of one prompt are contiguous
        return self.prompt_idx.repeat_interleave(reps).contiguous() if reps > 1 else self.prompt_idx

    def fill_context(self, pc: L.PromptCfgC, elevation, azimuth, ctx, neg_w) -> None:
        p = self.proc
        table = p.vd_table if pc.view_dependent else p.local_table
        unc = p.uncond_vd if pc.view_dependent else p.uncond
        idx = self._idx_for(elevation.shape[0])
        L.check(L.load().sdb_asd_text_embeddings_multi(C.byref(pc), L.ptr(table), L.ptr(unc), L.ptr(idx), table.shape[0],
                                                       L.ptr(elevation), L.ptr(azimuth), elevation.shape[0],
                                                       table.shape[-2], table.shape[-1], L.ptr(ctx), L.ptr(neg_w),
                                                       L.stream_ptr()),
                "sdb_asd_text_embeddings_multi")

    def _run(self, elevation, azimuth, view_dependent, perp_neg):
        B = elevation.shape[0]
        dev = self.proc.vd_table.device
        pc = self.prompt_cfg_c(view_dependent, perp_neg)
        ctx = torch.empty((5 if perp_neg else 3) * B, *self.proc.vd_table.shape[-2:], device=dev, dtype=torch.float16)
        neg_w = torch.zeros(B, 2, device=dev)
        self.fill_context(pc, elevation.to(dev, torch.float32).contiguous(), azimuth.to(dev, torch.float32).contiguous(),
                          ctx, neg_w)
        return ctx, neg_w

    def get_text_embeddings(self, elevation, azimuth, camera_distances, view_dependent_prompting: bool = True):
        ctx, _ = self._run(elevation, azimuth, view_dependent_prompting, False)
        return ctx[: 2 * elevation.shape[0]]

    def get_text_embeddings_perp_neg(self, elevation, azimuth, camera_distances, view_dependent_prompting: bool = True,
                                     guidance_scale_neg: Optional[float] = None):
        assert view_dependent_prompting, "Perp-Neg only works with view-dependent prompting"
        ctx, neg_w = self._run(elevation, azimuth, True, True)
        scale = -1.0 if guidance_scale_neg is None else float(guidance_scale_neg)
        return ctx[: 4 * elevation.shape[0]], neg_w * (-scale)  # kernel: -(a e^{-b r} + c); reference: (...) * scale


@register("stable-diffusion-multi-prompt-processor")
class StableDiffusionMultiPromptProcessor(BaseObject):
    @dataclass
    class Config(BaseObject.Config):
        prompt_library: str = "magic3d_prompt_library"
        prompt_library_dir: str = "load"
        prompt_library_format: str = "json"
        eval_prompt: Optional[str] = None
        eval_prompt_target: Optional[str] = None
        negative_prompt: str = ""
        pretrained_model_name_or_path: str = "runwayml/stable-diffusion-v1-5"
        overhead_threshold: float = 60.0
        front_threshold: float = 45.0
        back_threshold: float = 45.0
        view_dependent_prompt_front: bool = False
        use_cache: bool = True
        spawn: bool = False
        use_perp_neg: bool = False
        perp_neg_f_sb: Tuple[float, float, float] = (1, 0.5, -0.606)
        perp_neg_f_fsb: Tuple[float, float, float] = (1, 0.5, +0.967)
        perp_neg_f_fs: Tuple[float, float, float] = (4, 0.5, -2.426)
        perp_neg_f_sf: Tuple[float, float, float] = (4, 0.5, -2.426)
        use_prompt_debiasing: bool = False
        pretrained_model_name_or_path_prompt_debiasing: str = "bert-base-uncased"
        prompt_debiasing_mask_ids: Optional[List[int]] = None
        use_local_text_embeddings: bool = False

    cfg: Config
    embed_dim, n_tokens = 1024, 77

    def configure(self) -> None:
        if self.cfg.use_prompt_debiasing:
            raise NotImplementedError("Prompt debiasing is not implemented yet")
        if self.cfg.eval_prompt is None:
            rank, world = _world()
            lib = load_prompt_library(self.cfg, rank, world)
            self.prompt_library = [p for split in lib for p in lib[split]]
        else:
            self.prompt_library = [self.cfg.eval_prompt] + ([self.cfg.eval_prompt_target]
                                                            if self.cfg.eval_prompt_target else [])
        self.prompt_library = list(dict.fromkeys(self.prompt_library))
        self.negative_prompt = self.cfg.negative_prompt
        if self.cfg.view_dependent_prompt_front:
            fmt = {"side": "side view of {}", "front": "front view of {}", "back": "backside view of {}",
                   "overhead": "overhead view of {}"}
        else:
            fmt = {"side": "{}, side view", "front": "{}, front view", "back": "{}, back view",
                   "overhead": "{}, overhead view"}
        self.prompt2idx = {p: i for i, p in enumerate(self.prompt_library)}
        dev = self.device
        h16 = lambda ts: torch.stack(ts, 0).to(dev, torch.float16).contiguous()
        self.uncond = h16([self._embedding(self.negative_prompt, "local")])
        self.uncond_vd = h16([self._embedding(self.negative_prompt, "local") for _ in DIRECTIONS])
        self.local_table = h16([self._embedding(p, "local")[None] for p in self.prompt_library])          # [P,1,77,1024]
        self.vd_table = h16([torch.stack([self._embedding(fmt[d].format(p), "local") for d in DIRECTIONS], 0)
                             for p in self.prompt_library])                                             # [P,4,77,1024]
        self.global_table = torch.stack([self._embedding(p, "global") for p in self.prompt_library], 0).to(dev)

    def _embedding(self, prompt: str, kind: str) -> torch.Tensor:
        """Reference cache files `<md5(model-prompt-kind)>.pt` when present, else deterministic synthetic N(0,1)
        (no CLIP weights on the box): 'local' = [77,1024] token embeddings, 'global' = [1024] pooled embedding."""
        key = hashlib.md5(f"{self.cfg.pretrained_model_name_or_path}-{prompt}-{kind}".encode()).hexdigest()
        path = os.path.join(".threestudio_cache/text_embeddings", f"{key}.pt")
        shape = (self.n_tokens, self.embed_dim) if kind == "local" else (self.embed_dim,)
        if self.cfg.use_cache and os.path.exists(path):
            return torch.load(path, map_location="cpu").reshape(shape).float()
        core.synthetic_or_raise("text embeddings", path)
        g = torch.Generator().manual_seed(int(key[:8], 16) % (2 ** 31))
        return torch.randn(*shape, generator=g)

    def __call__(self, prompt: Union[str, List[str]]) -> MultiPromptProcessorOutput:
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        for p in prompts:
            if p not in self.prompt2idx:
                raise ValueError(f"Prompt [{p}] is not in the prompt library.")
        idx = torch.tensor([self.prompt2idx[p] for p in prompts], dtype=torch.int32, device=self.device)
        return MultiPromptProcessorOutput(self, idx, prompts)


# ------------------------------------------------------------------------------------------------ system
@register("multiprompt-radience-field-generator-system")
class MultipromptRadienceFieldGeneratorSystem(BaseSystem):
    @dataclass
    class Config(BaseSystem.Config):
        geometry_type: str = ""
        geometry: dict = field(default_factory=dict)
        geometry_convert_from: Optional[str] = None
        geometry_convert_inherit_texture: bool = False
        geometry_convert_override: dict = field(default_factory=dict)
        material_type: str = ""
        material: dict = field(default_factory=dict)
        background_type: str = ""
        background: dict = field(default_factory=dict)
        renderer_type: str = ""
        renderer: dict = field(default_factory=dict)
        guidance_type: str = ""
        guidance: dict = field(default_factory=dict)
        prompt_processor_type: str = ""
        prompt_processor: dict = field(default_factory=dict)
        exporter_type: str = "mesh-exporter"
        exporter: dict = field(default_factory=dict)
        stage: str = "coarse"
        visualize_samples: bool = False
        validation_via_video: bool = False
        rgb_as_latents: bool = False
        initialize_shape: bool = True
        train_guidance: bool = False

    cfg: Config

    def configure(self) -> None:
        if self.cfg.stage != "coarse":
            raise NotImplementedError(f"stage '{self.cfg.stage}' is not implemented (coarse only)")
        if self.cfg.rgb_as_latents or self.cfg.train_guidance:
            raise NotImplementedError("rgb_as_latents / train_guidance are outside the ASD hot path")
        dev = core.get_device()
        self.geometry = find(self.cfg.geometry_type)(self.cfg.geometry).to(dev)
        self.material = find(self.cfg.material_type)(self.cfg.material).to(dev)
        self.background = find(self.cfg.background_type)(self.cfg.background).to(dev)
        self.renderer = find(self.cfg.renderer_type)(self.cfg.renderer, geometry=self.geometry, material=self.material,
                                                     background=self.background).to(dev)

    def on_fit_start(self) -> None:
        self.guidance = find(self.cfg.guidance_type)(self.cfg.guidance)
        if self.cfg.initialize_shape and hasattr(self.geometry, "initialize_shape"):
            self.geometry.initialize_shape()
        self.prompt_processor = find(self.cfg.prompt_processor_type)(self.cfg.prompt_processor)

    def on_test_start(self) -> None:
        """multiprompt_radience_field_generator.py:83-91: evaluation without a preceding fit only needs the prompt processor
        (the guidance networks are never built for --validate / --test)."""
        if not hasattr(self, "prompt_processor"):
            self.prompt_processor = find(self.cfg.prompt_processor_type)(self.cfg.prompt_processor)

    on_validation_start = on_predict_start = on_test_start

    def forward(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        self.prompt_utils = self.prompt_processor(prompt=batch["prompt"])
        if "prompt_target" in batch:
            target = self.prompt_processor(prompt=batch["prompt_target"])
            ratio = batch["ratio"]
            batch["text_embed"] = ratio * self.prompt_utils.get_global_text_embeddings() \
                + (1 - ratio) * target.get_global_text_embeddings()
        else:
            batch["text_embed"] = self.prompt_utils.get_global_text_embeddings()
        return {**self.renderer(**batch)}

    def training_step(self, batch, batch_idx):
        out = self(batch)
        guidance_out = self.guidance(out["comp_rgb"], self.prompt_utils, **batch, rgb_as_latents=False)
        loss = 0.0
        lam = self.cfg.loss
        for name, value in guidance_out.items():
            self.log(f"train/{name}", value)
            if name.startswith("loss_"):
                loss = loss + value * self.C(lam[name.replace("loss_", "lambda_")])
        if self.C(lam.get("lambda_orient", 0.0)) > 0:
            dot = (out["normal"] * out["t_dirs"]).sum(-1, keepdim=True)
            loss_orient = (out["weights"].detach() * dot.clamp_min(0.0) ** 2).sum() / (out["opacity"] > 0).sum()
            self.log("train/loss_orient", loss_orient)
            loss = loss + loss_orient * self.C(lam["lambda_orient"])
        if self.C(lam.get("lambda_sparsity", 0.0)) > 0:
            loss_sparsity = (out["opacity"] ** 2 + 0.01).sqrt().mean()
            self.log("train/loss_sparsity", loss_sparsity)
            loss = loss + loss_sparsity * self.C(lam["lambda_sparsity"])
        if self.C(lam.get("lambda_opaque", 0.0)) > 0:
            op = out["opacity"].clamp(1.0e-3, 1.0 - 1.0e-3)
            loss_opaque = binary_cross_entropy(op, op)
            self.log("train/loss_opaque", loss_opaque)
            loss = loss + loss_opaque * self.C(lam["lambda_opaque"])
        if self.C(lam.get("lambda_z_variance", 0.0)) > 0:
            raise NotImplementedError("lambda_z_variance > 0: z_variance is a non-differentiable output of the "
                                      "compositing kernel")
        if "lambda_eikonal" in lam and self.C(lam["lambda_eikonal"]) > 0:
            if "sdf_grad" not in out:
                raise ValueError("sdf is required for eikonal loss, no sdf is found in the output.")
            loss_eikonal = ((torch.linalg.norm(out["sdf_grad"], ord=2, dim=-1) - 1.0) ** 2).mean()
            self.log("train/loss_eikonal", loss_eikonal)
            loss = loss + loss_eikonal * self.C(lam["lambda_eikonal"])
            self.log("train/inv_std", out["inv_std"])
        return {"loss": loss}

    def _eval_images(self, batch, name_key: str) -> Dict[str, Any]:
        """validation_step / test_step (multiprompt_radience_field_generator.py:218-300, 318-385) up to the image grid:
        every view of the batch as rgb | normal | opacity | per-view min-max normalised depth, filed under the prompt
        (`it{step}-val/{name}/{index}.png` in the reference). Writing files / videos stays with the caller."""
        out = self(batch)
        label = batch[name_key][0] if name_key in batch else batch["prompt"][0]
        res = {"name": label.replace(",", "").replace(".", "").replace(" ", "_"), "index": batch["index"],
               "comp_rgb": out["comp_rgb"], "opacity": out["opacity"]}
        if "comp_normal" in out:
            res["comp_normal"] = out["comp_normal"]
        if "depth" in out:
            d = out["depth"][..., 0]
            lo, hi = d.amin(dim=(1, 2), keepdim=True), d.amax(dim=(1, 2), keepdim=True)
            res["depth"] = (d - lo) / (hi - lo)
        return res

    def validation_step(self, batch, batch_idx):
        if self.cfg.visualize_samples:
            raise NotImplementedError
        return self._eval_images(batch, "prompt")

    def test_step(self, batch, batch_idx):
        return self._eval_images(batch, "name")
