"""CPU oracle for the NeRF render half of the ASD step.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product path (scaledreamer_b200/) never does.

PARITY STATUS: **unpinned** for the hash grid and the occupancy-grid sampling / compositing.  The reference has no
tests or golden vectors for this path and that arithmetic lives in two un-vendored CUDA extensions that
are absent from /root/reference and from this image: tiny-cuda-nn @ master (README.md:65) and nerfacc v0.5.2
(README.md:66).  This file restates their published algorithms and drives them with the reference's own formulas,
each cited below (paths relative to /root/reference).  **Pinned** where the reference's own code can run here
(tests/golden/make_field_golden.py, make_host_golden.py: definitions taken out of the reference files with `ast` and
executed unchanged): the frequency encoding + VanillaMLP field incl. density bias / activation and finite-difference
normals (ImplicitVolume.forward), and get_rays / look-at cameras.  Everything is plain differentiable float32 torch on the CPU, so
torch.autograd of these functions is the gradient oracle as well.

One documented choice where nerfacc's exact behaviour cannot be checked here: the lattice of candidate
samples.  nerfacc marches with a constant step `render_step_size` from a per-ray jittered near plane and
emits [t, t+step] intervals whose midpoint lies in an occupied cell.  We anchor the lattice at
near_plane + u*step (u = the per-ray stratified draw, an explicit input) for the whole ray:
t_start_k = near_j + k*step; a candidate exists iff its midpoint lies inside the ray/AABB span and in an
occupied cell.  The CUDA kernel implements exactly this definition.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch

PRIME_Y = 2654435761
PRIME_Z = 805459861


# --------------------------------------------------------------------------------------------- hash grid
@dataclass
class GridCfg:
    """pos_encoding_config of configs/single-prompt_benchmark/asd_sd_nerf.yaml:47-53."""
    n_levels: int = 16
    n_features_per_level: int = 2
    log2_hashmap_size: int = 19
    base_resolution: int = 16
    per_level_scale: float = 1.447269237440378


def grid_meta(cfg: GridCfg):
    """tiny-cuda-nn GridEncoding level geometry (restated): scale_l = 2^(l*log2(s))*base - 1,
    res_l = ceil(scale_l)+1, entries_l = min(round_up(res_l^3, 8), 2^log2_hashmap_size);
    a level is hashed iff res_l^3 exceeds its entries."""
    scales, ress, sizes, offsets, hashed = [], [], [], [], []
    off = 0
    # tcnn reads per_level_scale from JSON as a float32; the level scale is evaluated in float64 here and in the
    # library (csrc/capi_render.cu::resolve_grid) and rounded once to float32 (tcnn itself calls exp2f on device).
    log2_pls = math.log2(float(np.float32(cfg.per_level_scale)))
    for l in range(cfg.n_levels):
        scale = float(np.float32(2.0 ** (l * log2_pls) * cfg.base_resolution - 1.0))
        res = int(math.ceil(scale)) + 1
        dense = res ** 3
        n = min((dense + 7) // 8 * 8, 1 << cfg.log2_hashmap_size)
        scales.append(scale)
        ress.append(res)
        sizes.append(n)
        offsets.append(off)
        hashed.append(dense > n)
        off += n
    return dict(scale=scales, res=ress, size=sizes, offset=offsets, hashed=hashed, n_entries=off)


def hashgrid_encode(x01: torch.Tensor, table: torch.Tensor, cfg: GridCfg) -> torch.Tensor:
    """tcnn.Encoding(HashGrid, Linear interpolation) forward: x01 [N,3] in [0,1] -> [N, L*F] level-major.
    pos = x*scale + 0.5; corner index = dense x + y*res + z*res^2, or the coherent prime hash
    x ^ y*2654435761 ^ z*805459861 (uint32) on hashed levels; index % entries."""
    meta = grid_meta(cfg)
    x01 = x01.to(torch.float32)
    outs = []
    for l in range(cfg.n_levels):
        s = torch.tensor(meta["scale"][l], dtype=torch.float32)
        res, size, off = meta["res"][l], meta["size"][l], meta["offset"][l]
        # tcnn evaluates pos = fma(scale, x, 0.5f): one rounding. Emulated exactly through float64.
        pos = (x01.double() * s.double() + 0.5).float()
        g = torch.floor(pos)
        w = pos - g
        gi = g.to(torch.int64)
        acc = torch.zeros(x01.shape[0], cfg.n_features_per_level, dtype=torch.float32)
        for c in range(8):
            bx, by, bz = c & 1, (c >> 1) & 1, (c >> 2) & 1
            cx, cy, cz = gi[:, 0] + bx, gi[:, 1] + by, gi[:, 2] + bz
            if meta["hashed"][l]:
                idx = ((cx & 0xFFFFFFFF) ^ ((cy * PRIME_Y) & 0xFFFFFFFF) ^ ((cz * PRIME_Z) & 0xFFFFFFFF)) % size
            else:
                idx = (cx + cy * res + cz * res * res) % size
            wc = (w[:, 0] if bx else 1 - w[:, 0]) * (w[:, 1] if by else 1 - w[:, 1]) * (w[:, 2] if bz else 1 - w[:, 2])
            acc = acc + wc[:, None] * table[off + idx]
        outs.append(acc)
    return torch.cat(outs, dim=-1)


# --------------------------------------------------------------------------------------------- field
@dataclass
class FieldCfg:
    """geometry / material / background keys of asd_sd_nerf.yaml:28-73."""
    grid: GridCfg = field(default_factory=GridCfg)
    radius: float = 1.0
    density_bias: str = "blob_magic3d"
    density_bias_const: float = 0.0
    density_blob_scale: float = 10.0
    density_blob_std: float = 0.5
    density_activation: str = "softplus"
    fd_eps: float = 0.01
    color_activation: str = "sigmoid"
    bg_grid: GridCfg = field(default_factory=lambda: GridCfg(4, 2, 19, 4, 4.0))
    bg_color_activation: str = "sigmoid"
    # C1 (vanilla-MLP implicit volume): pos_encoding_config.otype ProgressiveBandFrequency, deeper VanillaMLP
    encoding: str = "hashgrid"  # "hashgrid" | "frequency"
    n_frequencies: int = 0
    include_xyz: bool = False
    n_hidden_layers: int = 1


def make_field_params(cfg: FieldCfg, seed: int = 0, table_scale: float = 1e-4) -> Dict[str, torch.Tensor]:
    """Random parameters with the reference's shapes: tcnn table U(-1e-4,1e-4) (scaled up in tests so the
    features matter), nn.Linear(bias=False) default init (networks.py:248-250)."""
    g = torch.Generator().manual_seed(seed)

    def lin(o, i):
        b = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * b

    n = grid_meta(cfg.grid)["n_entries"]
    nb = grid_meta(cfg.bg_grid)["n_entries"]
    if cfg.encoding == "frequency":
        d = 6 * cfg.n_frequencies + (3 if cfg.include_xyz else 0)
        P = {"w1d": lin(64, d), "w2d": lin(1, 64), "w1f": lin(64, d), "w2f": lin(3, 64)}
    else:
        P = {"table": (torch.rand(n, 2, generator=g) * 2 - 1) * table_scale,
             "w1d": lin(64, 32), "w2d": lin(1, 64), "w1f": lin(64, 32), "w2f": lin(3, 64)}
    if cfg.n_hidden_layers == 2:
        P["wmd"], P["wmf"] = lin(64, 64), lin(64, 64)
    P.update({"bg_table": (torch.rand(nb, 2, generator=g) * 2 - 1) * table_scale,
              "bg_w1": lin(16, 8), "bg_w2": lin(16, 16), "bg_w3": lin(3, 16)})
    return P


def freq_encode(x01: torch.Tensor, n_frequencies: int, mask: Optional[torch.Tensor] = None,
                include_xyz: bool = False) -> torch.Tensor:
    """ProgressiveBandFrequency.forward (threestudio/models/networks.py:16-52) as get_encoding wraps it in
    CompositeEncoding (:170-206): for every band 2^f, sin then cos of 2^f * x (3 channels each), times mask[f];
    with include_xyz the block x*2-1 comes first."""
    mask = torch.ones(n_frequencies) if mask is None else mask
    out = [x01 * 2.0 - 1.0] if include_xyz else []
    for f in range(n_frequencies):
        for fn in (torch.sin, torch.cos):
            out.append(fn((2.0 ** f) * x01) * mask[f])
    return torch.cat(out, -1)


def freq_mask(n_frequencies: int, n_masking_step: int, global_step: Optional[int]) -> torch.Tensor:
    """ProgressiveBandFrequency.update_step (networks.py:36-52)."""
    if n_masking_step <= 0 or global_step is None:
        return torch.ones(n_frequencies)
    ramp = (global_step / n_masking_step * n_frequencies - torch.arange(0, n_frequencies)).clamp(0, 1)
    return (1.0 - torch.cos(math.pi * ramp)) / 2.0


def _encode(points: torch.Tensor, P, cfg: FieldCfg) -> torch.Tensor:
    r = cfg.radius
    x01 = (points + r) / (2 * r)
    if cfg.encoding == "frequency":
        return freq_encode(x01, cfg.n_frequencies, P.get("freq_mask"), cfg.include_xyz)
    return hashgrid_encode(x01, P["table"], cfg.grid)


def _mlp(enc: torch.Tensor, P, head: str) -> torch.Tensor:
    """VanillaMLP networks.py:214-251 (bias-free Linear + ReLU; 1 or 2 hidden layers)."""
    h = torch.relu(enc @ P["w1" + head].T)
    if ("wm" + head) in P:
        h = torch.relu(h @ P["wm" + head].T)
    return h @ P["w2" + head].T


def _color_act(name: str, x: torch.Tensor) -> torch.Tensor:
    s = torch.sigmoid(x)
    if name == "sigmoid-mipnerf":  # threestudio/utils/ops.py:106-107
        return s * (1 + 2 * 0.001) - 0.001
    return s


def field_density(points: torch.Tensor, P: Dict[str, torch.Tensor], cfg: FieldCfg):
    """ImplicitVolume.forward_density (implicit_volume.py:198-207) + get_activated_density (:80-107) +
    contract_to_unisphere bounded branch (geometry/base.py:20-32). Returns (density [N], enc [N,32])."""
    enc = _encode(points, P, cfg)
    raw = _mlp(enc, P, "d")[:, 0]
    if cfg.density_bias == "blob_magic3d":
        raw = raw + cfg.density_blob_scale * (1 - torch.sqrt((points ** 2).sum(-1)) / cfg.density_blob_std)
    elif cfg.density_bias == "blob_dreamfusion":
        raw = raw + cfg.density_blob_scale * torch.exp(-0.5 * (points ** 2).sum(-1) / cfg.density_blob_std ** 2)
    else:
        raw = raw + cfg.density_bias_const
    if cfg.density_activation == "softplus":
        sigma = torch.nn.functional.softplus(raw)
    else:
        sigma = torch.exp(raw)
    return sigma, enc


def field_forward(points: torch.Tensor, P, cfg: FieldCfg, output_normal: bool = False):
    """ImplicitVolume.forward (implicit_volume.py:109-196): density, raw features, FD normals (:167-177)."""
    sigma, enc = field_density(points, P, cfg)
    feat = _mlp(enc, P, "f")
    out = {"density": sigma, "features": feat}
    if output_normal:
        eps = cfg.fd_eps
        offs = torch.eye(3) * eps
        sig_off = []
        for k in range(3):
            p = (points + offs[k]).clamp(-cfg.radius, cfg.radius)
            sig_off.append(field_density(p, P, cfg)[0])
        n = -(torch.stack(sig_off, -1) - sigma[:, None]) / eps
        out["normal"] = torch.nn.functional.normalize(n, dim=-1)
    return out


def background(dirs: torch.Tensor, P, cfg: FieldCfg) -> torch.Tensor:
    """NeuralEnvironmentMapBackground.forward (neural_environment_map_background.py:46-55), dirs [N,3]."""
    enc = hashgrid_encode((dirs + 1.0) / 2.0, P["bg_table"], cfg.bg_grid)
    h = torch.relu(enc @ P["bg_w1"].T)
    h = torch.relu(h @ P["bg_w2"].T)
    return _color_act(cfg.bg_color_activation, h @ P["bg_w3"].T)


# --------------------------------------------------------------------------------------------- marching
@dataclass
class MarchCfg:
    """renderer keys (asd_sd_nerf.yaml:75-79) resolved as nerf_volume_renderer.py:60-68,139-180."""
    render_step_size: float = 1.732 * 2 * 1.0 / 512
    near_plane: float = 0.0
    far_plane: float = 1e10
    prune: bool = True
    alpha_thre: float = 0.01
    early_stop_eps: float = 1e-4
    grid_res: int = 32


def _f32(x):
    return np.float32(x)


def ray_box(o: np.ndarray, d: np.ndarray, r: float, near: float, far: float):
    """Slab test against [-r,r]^3, float32 like the kernel."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = _f32(1.0) / d
        a = (_f32(-r) - o) * inv
        b = (_f32(r) - o) * inv
    tmin = np.minimum(a, b).max(-1)
    tmax = np.maximum(a, b).min(-1)
    t0 = np.maximum(tmin, _f32(near))
    t1 = np.minimum(tmax, _f32(far))
    return t0, t1, t1 > t0


def _fma32(a, b, c):
    """float32 fused multiply-add emulated through float64 (exact product, one rounding that matters)."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def march_candidates(rays_o, rays_d, jitter, occ: np.ndarray, fcfg: FieldCfg, mcfg: MarchCfg):
    """Candidate lattice samples in occupied cells (nerfacc OccGridEstimator.sampling -> traverse_grids with
    cone_angle=0, restated). occ: bool [res,res,res] indexed [x,y,z]. Returns packed (ray_idx, t_start, t_end,
    t_mid) sorted by (ray, t)."""
    o = rays_o.numpy().astype(np.float32)
    d = rays_d.numpy().astype(np.float32)
    u = np.zeros(len(o), np.float32) if jitter is None else jitter.numpy().astype(np.float32)
    step = _f32(mcfg.render_step_size)
    r = fcfg.radius
    t0, t1, hit = ray_box(o, d, r, mcfg.near_plane, mcfg.far_plane)
    near_j = _fma32(u, step, np.full_like(u, mcfg.near_plane))
    c0 = _fma32(np.full_like(u, 0.5), step, near_j)
    kmax = int(math.ceil((2 * math.sqrt(3.0) * r + 1e-3) / float(step))) + 4
    if np.isfinite(t1[hit]).any():
        kmax = max(kmax, int(np.nanmax(np.where(hit, (t1 - c0) / step, 0))) + 4)
    ks = np.arange(kmax, dtype=np.float32)
    res = mcfg.grid_res
    ray_idx, ts_l, te_l, tm_l = [], [], [], []
    chunk = 4096
    for s in range(0, len(o), chunk):
        e = min(s + chunk, len(o))
        tm = (ks[None, :].astype(np.float64) * np.float64(step) + c0[s:e, None].astype(np.float64)).astype(np.float32)
        inside = hit[s:e, None] & (tm >= t0[s:e, None]) & (tm < t1[s:e, None])
        p = (d[s:e, None, :].astype(np.float64) * tm[..., None].astype(np.float64)
             + o[s:e, None, :].astype(np.float64)).astype(np.float32)
        kf = _f32(res * 0.5 / r)
        cell = np.floor((p + _f32(r)) * kf).astype(np.int64).clip(0, res - 1)
        occd = occ[cell[..., 0], cell[..., 1], cell[..., 2]]
        keep = inside & occd
        ri, ki = np.nonzero(keep)
        nearj = (c0[s:e] - _f32(0.5) * step).astype(np.float32)
        ts = (ki.astype(np.float64) * np.float64(step) + nearj[ri].astype(np.float64)).astype(np.float32)
        ray_idx.append(ri + s)
        ts_l.append(ts)
        te_l.append((ts + step).astype(np.float32))
        tm_l.append(tm[ri, ki])
    cat = lambda xs, dt: torch.from_numpy(np.concatenate(xs).astype(dt)) if xs else torch.zeros(0)
    return cat(ray_idx, np.int64), cat(ts_l, np.float32), cat(te_l, np.float32), cat(tm_l, np.float32)


def _excl_cumsum_packed(x: torch.Tensor, ray_idx: torch.Tensor, n_rays: int) -> torch.Tensor:
    """nerfacc exclusive_sum over packed per-ray segments (samples sorted by ray)."""
    cs = torch.cumsum(x, 0)
    seg_total = torch.zeros(n_rays, dtype=x.dtype).index_add(0, ray_idx, x)
    seg_start = torch.cumsum(seg_total, 0) - seg_total
    return cs - x - seg_start[ray_idx]


def render(rays_o, rays_d, jitter, bg_override, occ, occ_mean: Optional[float], P, fcfg: FieldCfg, mcfg: MarchCfg,
           rays_per_image: int, output_normal: bool = False):
    """NeRFVolumeRenderer.forward, training branch, occgrid estimator (nerf_volume_renderer.py:118-386)."""
    n_rays = rays_o.shape[0]
    ray_idx, ts, te, tm = march_candidates(rays_o, rays_d, jitter, occ, fcfg, mcfg)
    pos = rays_o[ray_idx] + rays_d[ray_idx] * tm[:, None]
    if mcfg.prune and len(ray_idx) > 0:
        # sigma_fn pass + nerfacc render_visibility_from_density (:153-180)
        with torch.no_grad():
            sig = field_density(pos, P, fcfg)[0]
            sd = sig * (te - ts)
            T = torch.exp(-_excl_cumsum_packed(sd, ray_idx, n_rays))
            alpha = 1 - torch.exp(-sd)
            thre = mcfg.alpha_thre if occ_mean is None else min(mcfg.alpha_thre, occ_mean)
            keep = (alpha >= thre) & (T >= mcfg.early_stop_eps)
        ray_idx, ts, te, tm, pos = ray_idx[keep], ts[keep], te[keep], tm[keep], pos[keep]
    geo = field_forward(pos, P, fcfg, output_normal=output_normal)
    rgb = _color_act(fcfg.color_activation, geo["features"])  # NoMaterial no_material.py:41-54
    sigma = geo["density"]
    sd = sigma * (te - ts)
    T = torch.exp(-_excl_cumsum_packed(sd, ray_idx, n_rays))  # nerfacc render_weight_from_density (:313-319)
    w = T * (1 - torch.exp(-sd))
    acc = lambda v: torch.zeros(n_rays, v.shape[-1] if v.ndim > 1 else 1).index_add(
        0, ray_idx, v if v.ndim > 1 else v[:, None])
    opacity = acc(w)
    depth = acc(w * tm)
    fg = acc(w[:, None] * rgb)
    wn = w / opacity.clamp(min=1e-5)[ray_idx, 0]
    z_mean = acc(wn * tm)
    z_var = acc(wn * (tm - z_mean[ray_idx, 0]) ** 2) * (opacity > 0.5).float()
    bg = background(rays_d, P, fcfg)
    if bg_override is not None:  # random_aug branch (neural_environment_map_background.py:56-66)
        img = torch.arange(n_rays) // rays_per_image
        bg = bg * 0 + bg_override[img]
    comp = fg + bg * (1.0 - opacity)
    out = {"comp_rgb": comp, "comp_rgb_fg": fg, "comp_rgb_bg": bg, "opacity": opacity[:, 0], "depth": depth[:, 0],
           "z_variance": z_var[:, 0], "ray_indices": ray_idx, "t_starts": ts, "t_ends": te, "weights": w,
           "density": sigma, "rgb": rgb}
    if output_normal:
        out["normal"] = geo["normal"]
    return out


def occ_grid_from_density(P, fcfg: FieldCfg, mcfg: MarchCfg, occ_thre: float = 0.01, seed: int = 0):
    """One warm-up occupancy refresh (nerfacc update_every_n_steps, warm-up branch: every cell, jittered
    point, occ = sigma*step, thre = min(mean, occ_thre)). Returns (occs float [res^3], binary bool [res]^3,
    cell_rand [res^3,3])."""
    res = mcfg.grid_res
    g = torch.Generator().manual_seed(seed)
    idx = torch.arange(res ** 3)
    cz, cy, cx = idx % res, (idx // res) % res, idx // (res * res)
    rand = torch.rand(res ** 3, 3, generator=g)
    cell = 2 * fcfg.radius / res
    pts = -fcfg.radius + (torch.stack([cx, cy, cz], -1).float() + rand) * cell
    with torch.no_grad():
        sig = field_density(pts, P, fcfg)[0]
    occs = torch.maximum(torch.zeros(res ** 3), sig * mcfg.render_step_size)
    thre = min(float(occs.mean()), occ_thre)
    binary = (occs > thre).reshape(res, res, res)
    return occs, binary, rand


def get_rays(c2w: torch.Tensor, fovy: torch.Tensor, H: int, W: int):
    """get_ray_directions + get_rays (threestudio/utils/ops.py:183-269) as used by uncond.py:302-315."""
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32) + 0.5, torch.arange(H, dtype=torch.float32) + 0.5,
                          indexing="xy")
    focal = 0.5 * H / torch.tan(0.5 * fovy)
    dirs = torch.stack([(i - W / 2)[None] / focal[:, None, None], -(j - H / 2)[None] / focal[:, None, None],
                        -torch.ones(len(fovy), H, W)], -1)
    rays_d = (dirs[:, :, :, None, :] * c2w[:, None, None, :3, :3]).sum(-1)
    rays_o = c2w[:, None, None, :3, 3].expand(rays_d.shape)
    return rays_o.contiguous(), torch.nn.functional.normalize(rays_d, dim=-1)


def look_at_c2w(elev_deg, azim_deg, dist):
    """Camera pose from spherical coordinates, uncond.py:192-199,292-301 with zero perturbations."""
    e, a = torch.deg2rad(elev_deg), torch.deg2rad(azim_deg)
    pos = torch.stack([dist * torch.cos(e) * torch.cos(a), dist * torch.cos(e) * torch.sin(a), dist * torch.sin(e)], -1)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(pos)
    lookat = torch.nn.functional.normalize(-pos, dim=-1)
    right = torch.nn.functional.normalize(torch.cross(lookat, up, dim=-1), dim=-1)
    up2 = torch.nn.functional.normalize(torch.cross(right, lookat, dim=-1), dim=-1)
    c2w = torch.zeros(len(pos), 4, 4)
    c2w[:, :3, 0], c2w[:, :3, 1], c2w[:, :3, 2], c2w[:, :3, 3] = right, up2, -lookat, pos
    c2w[:, 3, 3] = 1.0
    return c2w
