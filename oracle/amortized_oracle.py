"""CPU oracle of the amortized (multi-prompt) generator path -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU
legs may import it; the product path never does).

Restates, in plain fp32 PyTorch on the CPU, the reference's
  LinearHyperNetwork / Hypernet_Sdf      custom/amortized/models/geometry/hyper_iNGP.py:18-111, 206-349
  multi-prompt environment map           custom/amortized/models/background/multiprompt_neural_environment_hashgrid_map_background.py:83-116
  ImportanceEstimator.sampling           threestudio/models/estimators.py:23-101
  volsdf_density / get_alpha             threestudio/models/renderers/neus_volume_renderer.py:19-23, 93-96
  GenerativeSpaceVolSDFVolumeRenderer    custom/amortized/models/renderers/generative_space_volsdf_volume_renderer.py:89-446
  eikonal / sparsity / opaque losses     custom/amortized/systems/multiprompt_radience_field_generator.py:127-216

PINNED against the reference's own code (tests/golden/make_amortized_golden.py -> tests/test_amortized_golden_cpu.py):
Adan, volsdf_density, LinearHyperNetwork, sample_from_planes, and the complete Hypernet_Sdf.forward /
TriplaneTransformerSDF.forward (heads, sphere bias, clamped finite-difference sdf_grad, normals).
PARITY UNPINNED for the nerfacc pieces (nerfacc v0.5.2 is an un-vendored dependency, not installable here):
`importance_sampling` is restated as inverse-CDF sampling at u_j = (j + b) / (n + 1), j = 0..n, with one offset b per
ray (U[0,1) when stratified, 0.5 otherwise) and linear interpolation inside a CDF bin; `render_weight_from_alpha` as
T_i = prod_{j<i} (1 - alpha_j); `render_transmittance_from_density` as T_i = exp(-sum_{j<i} sigma_j dt_j). The random
offsets are explicit inputs so that oracle and kernels consume identical draws. The hash grid follows
oracle/render_oracle.py (tiny-cuda-nn semantics, also unpinned).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from oracle import render_oracle as ro


@dataclass
class HyperCfg:
    grid: ro.GridCfg = field(default_factory=ro.GridCfg)
    radius: float = 2.0
    sdf_bias_radius: float = 0.5     # sdf_bias: sphere, sdf_bias_params: 0.5
    fd_eps: float = 0.01
    c_dim: int = 1024
    n_neurons: int = 64


@dataclass
class VolSDFCfg:
    near: float = 0.1
    far: float = 4.0
    n_coarse: int = 128              # num_samples_per_ray_importance (proposal intervals)
    n_fine: int = 64                 # num_samples_per_ray
    inv_std: float = math.exp(0.340119 * 10.0)


def make_hypernet(c_dim: int, n_neurons: int, n_out: int, seed: int) -> Dict[str, torch.Tensor]:
    """LinearHyperNetwork parameters (hyper_iNGP.py:58-77, 101-108): xavier-normal weights, zero biases."""
    g = torch.Generator().manual_seed(seed)
    xavier = lambda o, i: torch.randn(o, i, generator=g) * math.sqrt(2.0 / (i + o))
    return {"layers.0.weight": xavier(n_neurons, c_dim), "layers.1.weight": torch.ones(n_neurons),
            "layers.1.bias": torch.zeros(n_neurons), "layers.3.weight": xavier(n_out, n_neurons),
            "layers.3.bias": torch.zeros(n_out)}


def hypernet_forward(P: Dict[str, torch.Tensor], text_embed: torch.Tensor, out_dims: Dict[str, list]):
    """-> {name: [W_in_hidden [B,32,64], W_hidden_out [B,64,k]]} (hyper_iNGP.py:79-99)."""
    h = text_embed @ P["layers.0.weight"].t()
    h = F.layer_norm(h, (h.shape[-1],), P["layers.1.weight"], P["layers.1.bias"], 1e-5)
    h = F.silu(h)
    out = h @ P["layers.3.weight"].t() + P["layers.3.bias"]
    res, start = {}, 0
    for name, ch in out_dims.items():
        mats = []
        for i, o in zip(ch[:-1], ch[1:]):
            mats.append(out[:, start:start + i * o].reshape(-1, i, o))
            start += i * o
        res[name] = mats
    return res


def hyper_mlp(enc: torch.Tensor, mats) -> torch.Tensor:
    """hypernet_forward (hyper_iNGP.py:238-259): bmm chain with ReLU between, no bias."""
    x = enc
    for i, w in enumerate(mats):
        x = torch.bmm(x, w)
        if i < len(mats) - 1:
            x = torch.relu(x)
    return x


def hyper_sdf(points: torch.Tensor, table: torch.Tensor, cache, cfg: HyperCfg) -> torch.Tensor:
    """forward_sdf (hyper_iNGP.py:324-349). points [B, N, 3] -> [B, N]."""
    B, N, _ = points.shape
    x01 = (points + cfg.radius) / (2 * cfg.radius)
    enc = ro.hashgrid_encode(x01.reshape(-1, 3), table, cfg.grid).view(B, N, -1)
    sdf = hyper_mlp(enc, cache["sdf_weights"])[..., 0]
    return sdf + (points.norm(dim=-1) - cfg.sdf_bias_radius)


def hyper_field(points: torch.Tensor, table: torch.Tensor, cache, cfg: HyperCfg, output_normal: bool = False):
    """Hypernet_Sdf.forward (hyper_iNGP.py:261-322)."""
    B, N, _ = points.shape
    x01 = (points + cfg.radius) / (2 * cfg.radius)
    enc = ro.hashgrid_encode(x01.reshape(-1, 3), table, cfg.grid).view(B, N, -1)
    sdf = hyper_mlp(enc, cache["sdf_weights"])[..., 0] + (points.norm(dim=-1) - cfg.sdf_bias_radius)
    out = {"sdf": sdf.reshape(B * N, 1), "features": hyper_mlp(enc, cache["feature_weights"]).reshape(B * N, 3)}
    if output_normal:
        eps = cfg.fd_eps
        offs = (points[..., None, :] + eps * torch.eye(3)).clamp(-cfg.radius, cfg.radius)   # [B,N,3,3]
        sdf_off = hyper_sdf(offs.reshape(B, N * 3, 3), table, cache, cfg).view(B, N, 3)
        sdf_grad = (sdf_off - sdf[..., None]) / eps
        normal = F.normalize(sdf_grad, dim=-1)
        out.update(normal=normal.reshape(B * N, 3), shading_normal=normal.reshape(B * N, 3),
                   sdf_grad=sdf_grad.reshape(B * N, 3))
    return out


def hyper_background(dirs: torch.Tensor, table: torch.Tensor, mats, grid: ro.GridCfg) -> torch.Tensor:
    """dirs [B, HW, 3] (unit) -> sigmoid colour [B, HW, 3]."""
    B, N, _ = dirs.shape
    enc = ro.hashgrid_encode(((dirs + 1.0) / 2.0).reshape(-1, 3), table, grid).view(B, N, -1)
    return torch.sigmoid(hyper_mlp(enc, mats))


def volsdf_density(sdf: torch.Tensor, inv_std: float) -> torch.Tensor:
    a = min(max(inv_std, 0.0), 80.0)
    return a * (0.5 + 0.5 * torch.sign(sdf) * torch.expm1(-sdf.abs() * a))


def importance_sampling(edges: torch.Tensor, cdfs: torch.Tensor, n: int, offset: torch.Tensor) -> torch.Tensor:
    """Inverse-CDF resampling to n intervals (n + 1 sorted edges). edges/cdfs [Nr, m]; offset [Nr] in [0,1)."""
    u = (torch.arange(n + 1, dtype=edges.dtype)[None, :] + offset[:, None]) / (n + 1)
    idx = torch.searchsorted(cdfs.contiguous(), u.contiguous(), right=True).clamp(1, cdfs.shape[1] - 1)
    c_lo, c_hi = torch.gather(cdfs, 1, idx - 1), torch.gather(cdfs, 1, idx)
    e_lo, e_hi = torch.gather(edges, 1, idx - 1), torch.gather(edges, 1, idx)
    w = ((u - c_lo) / (c_hi - c_lo).clamp_min(1e-10)).clamp(0.0, 1.0)
    return e_lo + (e_hi - e_lo) * w


def sample_intervals(rays_o, rays_d, n_rays_per_prompt, table, cache, hcfg: HyperCfg, vcfg: VolSDFCfg,
                     u_coarse: torch.Tensor, u_fine: torch.Tensor) -> torch.Tensor:
    """ImportanceEstimator.sampling with one VolSDF proposal (estimators.py:60-101) -> sorted t [Nr, n_c + n_f + 2]."""
    Nr = rays_o.shape[0]
    B = Nr // n_rays_per_prompt
    with torch.no_grad():
        unit = torch.tensor([[0.0, 1.0]]).expand(Nr, 2)
        s_c = importance_sampling(unit, unit, vcfg.n_coarse, u_coarse)
        t_c = vcfg.near + s_c * (vcfg.far - vcfg.near)
        mid = 0.5 * (t_c[:, :-1] + t_c[:, 1:])
        pts = rays_o[:, None, :] + rays_d[:, None, :] * mid[..., None]
        sdf = hyper_sdf(pts.reshape(B, -1, 3), table, cache, hcfg).reshape(Nr, -1)
        sigma = volsdf_density(sdf, vcfg.inv_std)
        sd = sigma * (t_c[:, 1:] - t_c[:, :-1])
        trans = torch.exp(-(torch.cumsum(sd, -1) - sd))
        cdfs = 1.0 - torch.cat([trans, torch.zeros_like(trans[:, :1])], -1)
        s_f = importance_sampling(s_c, cdfs, vcfg.n_fine, u_fine)
        t_f = vcfg.near + s_f * (vcfg.far - vcfg.near)
        t_all, _ = torch.sort(torch.cat([t_c, t_f], -1), -1)
    return t_all


def composite(sdf, rgb, normal, t_mid, delta, inv_std):
    """get_alpha (VolSDF) + render_weight_from_alpha + accumulate_along_rays on dense [Nr, S] samples."""
    alpha = delta.abs() * volsdf_density(sdf, inv_std)
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha[:, :-1]], -1), -1)
    w = T * alpha
    opacity = w.sum(-1)
    depth = (w * t_mid).sum(-1)
    fg = (w[..., None] * rgb).sum(-2)
    z_var = (w * (t_mid - depth[:, None]) ** 2).sum(-1)
    cn = F.normalize((w[..., None] * normal).sum(-2), dim=-1)
    cn = torch.lerp(torch.zeros_like(cn), (cn.detach() + 1.0) / 2.0, opacity[:, None])
    return dict(weights=w, opacity=opacity, depth=depth, comp_rgb_fg=fg, z_variance=z_var, comp_normal=cn)


def render(rays_o, rays_d, n_rays_per_prompt, table, cache, bg_rgb, hcfg: HyperCfg, vcfg: VolSDFCfg, u_coarse, u_fine):
    """GenerativeSpaceVolSDFVolumeRenderer._forward. rays [Nr,3]; bg_rgb [Nr,3] -> dict of per-ray / per-sample outputs."""
    Nr = rays_o.shape[0]
    B = Nr // n_rays_per_prompt
    t = sample_intervals(rays_o, rays_d, n_rays_per_prompt, table, cache, hcfg, vcfg, u_coarse, u_fine)
    t_mid, delta = 0.5 * (t[:, :-1] + t[:, 1:]), t[:, 1:] - t[:, :-1]
    S = t_mid.shape[1]
    pts = rays_o[:, None, :] + rays_d[:, None, :] * t_mid[..., None]
    geo = hyper_field(pts.reshape(B, -1, 3), table, cache, hcfg, output_normal=True)
    rgb = torch.sigmoid(geo["features"]).view(Nr, S, 3)
    out = composite(geo["sdf"].view(Nr, S), rgb, geo["normal"].view(Nr, S, 3), t_mid, delta, vcfg.inv_std)
    out["comp_rgb_bg"] = bg_rgb
    out["comp_rgb"] = out["comp_rgb_fg"] + bg_rgb * (1.0 - out["opacity"][:, None])
    out.update(t=t, t_points=t_mid, t_intervals=delta, sdf=geo["sdf"], sdf_grad=geo["sdf_grad"], normal=geo["normal"])
    return out


def eikonal_loss(sdf_grad: torch.Tensor) -> torch.Tensor:
    return ((torch.linalg.norm(sdf_grad, ord=2, dim=-1) - 1.0) ** 2).mean()


# ------------------------------------------------------------------------------------------------ triplane / Adan
_PLANE_AXES = torch.tensor([[[1, 0, 0], [0, 1, 0], [0, 0, 1]], [[1, 0, 0], [0, 0, 1], [0, 1, 0]],
                            [[0, 0, 1], [0, 1, 0], [1, 0, 0]]], dtype=torch.float32)


def sample_from_planes(plane_features: torch.Tensor, coordinates: torch.Tensor, box_warp: float = 2.0) -> torch.Tensor:
    """custom/amortized/models/geometry/utils.py:67-97 restated with the same torch primitives (inverse plane matrices,
    bmm, F.grid_sample bilinear / zeros / align_corners=False). plane_features [N,3,C,H,W], coordinates [N,M,3]."""
    N, n_planes, C, H, W = plane_features.shape
    M = coordinates.shape[1]
    pf = plane_features.reshape(N * n_planes, C, H, W)
    coords = (2.0 / box_warp) * coordinates
    c = coords.unsqueeze(1).expand(-1, n_planes, -1, -1).reshape(N * n_planes, M, 3)
    inv = torch.linalg.inv(_PLANE_AXES).unsqueeze(0).expand(N, -1, -1, -1).reshape(N * n_planes, 3, 3)
    proj = torch.bmm(c, inv)[..., :2].unsqueeze(1)
    out = F.grid_sample(pf, proj.float(), mode="bilinear", padding_mode="zeros", align_corners=False)
    out = out.permute(0, 3, 2, 1).reshape(N, n_planes, M, C)
    return out.permute(0, 2, 1, 3).reshape(N, M, n_planes * C).contiguous()


def vanilla_mlp(x: torch.Tensor, weights) -> torch.Tensor:
    """threestudio/models/networks.py:214-251: bias-free Linear + ReLU stack."""
    for i, w in enumerate(weights):
        x = x @ w.t()
        if i < len(weights) - 1:
            x = torch.relu(x)
    return x


def triplane_field(points, planes, w_sdf, w_feat, radius, sdf_bias_radius, fd_eps, output_normal=False):
    """TriplaneTransformerSDF.forward (triplane_transformer.py:153-239). points [B,N,3] world, planes [B,3,C,H,W]."""
    B, N, _ = points.shape
    contract = lambda p: (p + radius) / (2 * radius) * 2.0 - 1.0
    sdf_of = lambda p: vanilla_mlp(sample_from_planes(planes, contract(p)), w_sdf)[..., 0] + (p.norm(dim=-1) - sdf_bias_radius)
    enc = sample_from_planes(planes, contract(points))
    sdf = vanilla_mlp(enc, w_sdf)[..., 0] + (points.norm(dim=-1) - sdf_bias_radius)
    out = {"sdf": sdf.reshape(B * N, 1), "features": vanilla_mlp(enc, w_feat).reshape(B * N, -1)}
    if output_normal:
        offs = (points[..., None, :] + fd_eps * torch.eye(3)).clamp(-radius, radius)
        sdf_grad = (sdf_of(offs.reshape(B, N * 3, 3)).view(B, N, 3) - sdf[..., None]) / fd_eps
        out.update(sdf_grad=sdf_grad.reshape(B * N, 3), normal=F.normalize(sdf_grad, dim=-1).reshape(B * N, 3))
    return out


def adan_step(p, g, state, step, lr, betas, eps, weight_decay=0.0, no_prox=False):
    """threestudio/systems/optimizers.py:141-250 (_single_tensor_adan with the step-1 neg_pre_grad initialisation of
    :171-172), max_grad_norm = 0. state: dict of exp_avg / exp_avg_sq / exp_avg_diff / neg_pre_grad (created on step 1)."""
    b1, b2, b3 = betas
    if step == 1:
        for k in ("exp_avg", "exp_avg_sq", "exp_avg_diff"):
            state[k] = torch.zeros_like(p)
        state["neg_pre_grad"] = g.clone().mul_(-1.0)
    bc1, bc2, bc3 = 1 - b1 ** step, 1 - b2 ** step, 1 - b3 ** step
    npg = state["neg_pre_grad"]
    npg.add_(g)
    state["exp_avg"].mul_(b1).add_(g, alpha=1 - b1)
    state["exp_avg_diff"].mul_(b2).add_(npg, alpha=1 - b2)
    npg.mul_(b2).add_(g)
    state["exp_avg_sq"].mul_(b3).addcmul_(npg, npg, value=1 - b3)
    denom = (state["exp_avg_sq"].sqrt() / math.sqrt(bc3)).add_(eps)
    if no_prox:
        p.mul_(1 - lr * weight_decay)
    p.addcdiv_(state["exp_avg"], denom, value=-lr / bc1)
    p.addcdiv_(state["exp_avg_diff"], denom, value=-lr * b2 / bc2)
    if not no_prox:
        p.div_(1 + lr * weight_decay)
    npg.zero_().add_(g, alpha=-1.0)
    return p
