"""Host-side wrappers of the fused render kernels: tensor plumbing + torch.autograd glue.

The arithmetic lives in csrc/render_fwd.cu / render_bwd.cu behind the C ABI (include/sdb200.h); this file only
allocates outputs, fills the parameter structs and registers the backward.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import lib as L

BIAS_TYPES = {"const": 0, "blob_magic3d": 1, "blob_dreamfusion": 2}
DENSITY_ACTS = {"softplus": 0, "exp": 1, "trunc_exp": 2}
COLOR_ACTS = {"sigmoid": 0, "sigmoid-mipnerf": 1}


@dataclass
class FieldSpec:
    """Static description of the iNGP field (geometry + material + background configs, resolved)."""
    grid: dict
    bg_grid: dict
    radius: float = 1.0
    density_bias: object = "blob_magic3d"
    density_blob_scale: float = 10.0
    density_blob_std: float = 0.5
    density_activation: str = "softplus"
    fd_eps: float = 0.01
    color_activation: str = "sigmoid"
    bg_color_activation: str = "sigmoid"


def _enum(table: dict, key, what: str) -> int:
    if key not in table:
        raise NotImplementedError(f"{what} '{key}' is not implemented by the sm_100a render kernels "
                                  f"(supported: {sorted(table)})")
    return table[key]


def make_field_c(spec: FieldSpec, params: Dict[str, torch.Tensor]) -> L.FieldC:
    """params keys: table, w1d, w2d, w1f, w2f, bg_table, bg_w1, bg_w2, bg_w3 (fp32 CUDA, contiguous)."""
    f = L.FieldC()
    f.grid = L.grid_cfg_c(spec.grid)
    f.bg_grid = L.grid_cfg_c(spec.bg_grid)
    for name, key in (("table", "table"), ("w1_density", "w1d"), ("w2_density", "w2d"), ("w1_feature", "w1f"),
                      ("w2_feature", "w2f"), ("bg_table", "bg_table"), ("bg_w1", "bg_w1"), ("bg_w2", "bg_w2"),
                      ("bg_w3", "bg_w3")):
        t = params.get(key)
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError(f"{key} must be float32")
        setattr(f, name, L.ptr(t))
    f.radius = float(spec.radius)
    if isinstance(spec.density_bias, str):
        f.density_bias_type = _enum(BIAS_TYPES, spec.density_bias, "density_bias")
        f.density_bias_const = 0.0
    else:
        f.density_bias_type = 0
        f.density_bias_const = float(spec.density_bias)
    f.density_blob_scale = float(spec.density_blob_scale)
    f.density_blob_std = float(spec.density_blob_std)
    f.density_activation = _enum(DENSITY_ACTS, spec.density_activation, "density_activation")
    f.fd_normal_eps = float(spec.fd_eps)
    f.color_activation = _enum(COLOR_ACTS, spec.color_activation, "color_activation")
    f.bg_color_activation = _enum(COLOR_ACTS, spec.bg_color_activation, "background color_activation")
    return f


@dataclass
class MarchSpec:
    render_step_size: float
    near_plane: float = 0.0
    far_plane: float = 1e10
    prune: bool = True
    alpha_thre: float = 0.01
    early_stop_eps: float = 1e-4
    grid_res: int = 32
    output_normal: bool = False

    def to_c(self) -> L.MarchCfgC:
        return L.MarchCfgC(float(self.render_step_size), float(self.near_plane), float(min(self.far_plane, 3e38)),
                           int(self.prune), float(self.alpha_thre), float(self.early_stop_eps), int(self.grid_res),
                           int(self.output_normal))


def hashgrid_forward(x01: torch.Tensor, table: torch.Tensor, grid_cfg) -> torch.Tensor:
    lib = L.load()
    c = L.grid_cfg_c(grid_cfg)
    x01 = x01.contiguous().float()
    out = torch.empty(x01.shape[0], c.n_levels * c.n_features_per_level, device=x01.device, dtype=torch.float32)
    L.check(lib.sdb_hashgrid_forward(C.byref(c), L.ptr(table), L.ptr(x01), x01.shape[0], L.ptr(out), L.stream_ptr()),
            "sdb_hashgrid_forward")
    return out


def hashgrid_backward(x01: torch.Tensor, g_out: torch.Tensor, n_entries: int, grid_cfg) -> torch.Tensor:
    lib = L.load()
    c = L.grid_cfg_c(grid_cfg)
    x01 = x01.contiguous().float()
    g_out = g_out.contiguous().float()
    g_table = torch.zeros(n_entries, c.n_features_per_level, device=x01.device, dtype=torch.float32)
    L.check(lib.sdb_hashgrid_backward(C.byref(c), L.ptr(x01), L.ptr(g_out), x01.shape[0], L.ptr(g_table),
                                      L.stream_ptr()), "sdb_hashgrid_backward")
    return g_table


def field_forward(spec: FieldSpec, params, points: torch.Tensor, want_features=True, want_normal=False):
    lib = L.load()
    f = make_field_c(spec, {**params, "bg_table": params.get("bg_table"), "bg_w1": params.get("bg_w1"),
                            "bg_w2": params.get("bg_w2"), "bg_w3": params.get("bg_w3")})
    pts = points.reshape(-1, 3).contiguous().float()
    n = pts.shape[0]
    density = torch.empty(n, device=pts.device)
    features = torch.empty(n, 3, device=pts.device) if want_features else None
    normal = torch.empty(n, 3, device=pts.device) if want_normal else None
    L.check(lib.sdb_field_forward(C.byref(f), L.ptr(pts), n, L.ptr(density), L.ptr(features), L.ptr(normal),
                                  L.stream_ptr()), "sdb_field_forward")
    return density, features, normal


class OccGrid:
    """Device-resident occupancy state of nerfacc.OccGridEstimator(resolution=32, levels=1)
    (nerf_volume_renderer.py:60-65): float occs, packed binaries, running mean."""

    def __init__(self, res: int, device, all_occupied: bool = False):
        self.res = res
        n = res ** 3
        self.occs = torch.zeros(n, device=device)
        self.bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device=device)
        self.mean = torch.zeros(1, device=device)
        if all_occupied:  # grid_prune: false
            self.occs.fill_(1.0)
            self.bits.fill_(-1)
            self.mean.fill_(1.0)

    def binaries(self) -> torch.Tensor:
        """bool [res,res,res] view (x,y,z) of the packed bit-field (tests / debugging)."""
        b = self.bits.view(-1, 1) >> torch.arange(32, device=self.bits.device, dtype=torch.int32)
        return (b & 1).bool().reshape(-1)[: self.res ** 3].reshape(self.res, self.res, self.res)

    def set_binaries(self, binary: torch.Tensor, occs: Optional[torch.Tensor] = None) -> None:
        flat = binary.reshape(-1).to(self.bits.device).to(torch.int64)
        pad = (-flat.numel()) % 32
        if pad:
            flat = torch.cat([flat, flat.new_zeros(pad)])
        words = (flat.view(-1, 32) << torch.arange(32, device=flat.device)).sum(-1)
        words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
        self.bits.copy_(words.to(torch.int32))
        if occs is not None:
            self.occs.copy_(occs.reshape(-1).to(self.occs.device))
            self.mean.copy_(self.occs.mean().reshape(1))

    def update(self, spec: FieldSpec, params, cell_idx: torch.Tensor, cell_rand: torch.Tensor, step_size: float,
               ema_decay: float = 0.95, occ_thre: float = 0.01) -> None:
        lib = L.load()
        f = make_field_c(spec, params)
        cell_idx = cell_idx.to(torch.int32).contiguous()
        cell_rand = cell_rand.float().contiguous()
        L.check(lib.sdb_occgrid_update(C.byref(f), L.ptr(cell_idx), L.ptr(cell_rand), cell_idx.numel(), self.res,
                                       float(step_size), float(ema_decay), float(occ_thre), L.ptr(self.occs),
                                       L.ptr(self.bits), L.ptr(self.mean), L.stream_ptr()), "sdb_occgrid_update")


PARAM_KEYS = ("table", "w1d", "w2d", "w1f", "w2f", "bg_table", "bg_w1", "bg_w2", "bg_w3")


def render_forward_raw(spec: FieldSpec, march: MarchSpec, params, occ: OccGrid, rays_o, rays_d, jitter, bg_override,
                       rays_per_image: int, packed_capacity: int = 0):
    """Runs sdb_render_nerf_forward. rays_* [Nr,3]. Returns dict of per-ray tensors (+ packed samples)."""
    lib = L.load()
    f = make_field_c(spec, params)
    m = march.to_c()
    dev = rays_o.device
    n = rays_o.shape[0]
    out = {k: torch.empty(n, 3, device=dev) for k in ("comp_rgb", "comp_rgb_fg", "comp_rgb_bg")}
    out.update({k: torch.empty(n, device=dev) for k in ("opacity", "depth", "z_variance")})
    work = torch.zeros(1, dtype=torch.int32, device=dev)
    pk = L.PackedSamplesC()
    packed = None
    if packed_capacity > 0:
        cap = packed_capacity
        packed = {"counter": torch.zeros(1, dtype=torch.int32, device=dev),
                  "ray_indices": torch.empty(cap, dtype=torch.int32, device=dev),
                  "t_starts": torch.empty(cap, device=dev), "t_ends": torch.empty(cap, device=dev),
                  "weights": torch.empty(cap, device=dev), "density": torch.empty(cap, device=dev),
                  "rgb": torch.empty(cap, 3, device=dev),
                  "normal": torch.empty(cap, 3, device=dev) if march.output_normal else None}
        pk.counter = L.ptr(packed["counter"])
        pk.capacity = cap
        for k in ("ray_indices", "t_starts", "t_ends", "weights", "density", "rgb", "normal"):
            setattr(pk, k, L.ptr(packed[k]))
    L.check(lib.sdb_render_nerf_forward(
        C.byref(f), C.byref(m), L.ptr(occ.bits), L.ptr(occ.mean), L.ptr(rays_o), L.ptr(rays_d), L.ptr(jitter),
        L.ptr(bg_override), n, int(rays_per_image), L.ptr(out["comp_rgb"]), L.ptr(out["comp_rgb_fg"]),
        L.ptr(out["comp_rgb_bg"]), L.ptr(out["opacity"]), L.ptr(out["depth"]), L.ptr(out["z_variance"]),
        C.byref(pk) if packed is not None else None, L.ptr(work), L.stream_ptr()), "sdb_render_nerf_forward")
    out["packed"] = packed
    return out


def render_backward_raw(spec: FieldSpec, march: MarchSpec, params, grads, occ: OccGrid, rays_o, rays_d, jitter,
                        bg_override, rays_per_image: int, saved, g_comp_rgb, g_opacity=None, g_depth=None) -> None:
    """Runs sdb_render_nerf_backward; accumulates into `grads` (dict keyed like params)."""
    lib = L.load()
    f = make_field_c(spec, params)
    m = march.to_c()
    g = L.FieldGradsC()
    for name, key in (("table", "table"), ("w1_density", "w1d"), ("w2_density", "w2d"), ("w1_feature", "w1f"),
                      ("w2_feature", "w2f"), ("bg_table", "bg_table"), ("bg_w1", "bg_w1"), ("bg_w2", "bg_w2"),
                      ("bg_w3", "bg_w3")):
        setattr(g, name, L.ptr(grads[key]))
    n = rays_o.shape[0]
    work = torch.zeros(1, dtype=torch.int32, device=rays_o.device)
    L.check(lib.sdb_render_nerf_backward(
        C.byref(f), C.byref(g), C.byref(m), L.ptr(occ.bits), L.ptr(occ.mean), L.ptr(rays_o), L.ptr(rays_d),
        L.ptr(jitter), L.ptr(bg_override), n, int(rays_per_image), L.ptr(saved["comp_rgb_fg"]),
        L.ptr(saved["comp_rgb_bg"]), L.ptr(saved["opacity"]), L.ptr(saved["depth"]), L.ptr(g_comp_rgb),
        L.ptr(g_opacity), L.ptr(g_depth), L.ptr(work), L.stream_ptr()), "sdb_render_nerf_backward")


class RenderTape:
    """Device buffers of the v2 renderer's sample tape (include/sdb200.h sdb_render_tape), sized for the worst case
    (every lattice sample of every ray kept) so the forward can never overflow it. Tapes are pooled: a forward takes
    a free one, its backward hands it back (two forwards before a backward simply use two tapes)."""

    _pool: Dict[tuple, list] = {}
    n_allocated = 0  # tapes ever created (diagnostics: should stay at the number of concurrently live forwards)

    def __init__(self, capacity: int, max_chunks: int, n_rays: int, device):
        RenderTape.n_allocated += 1
        self.capacity, self.max_chunks, self.n_rays = capacity, max_chunks, n_rays
        self.counter = torch.zeros(2, dtype=torch.int32, device=device)
        self.enc = torch.empty(capacity * 32, device=device)
        self.pos = torch.empty(3 * capacity, device=device)
        self.sample = torch.empty(8 * capacity, device=device)
        self.ray_chunks = torch.empty(n_rays * max_chunks, dtype=torch.int32, device=device)
        self.ray_nchunks = torch.zeros(n_rays, dtype=torch.int32, device=device)
        self.key = None
        self._og = None

    @property
    def og(self) -> torch.Tensor:
        """[4][capacity] partial derivatives of the orientation term (allocated on first use: 16 B per slot)."""
        if self._og is None:
            self._og = torch.empty(4 * self.capacity, device=self.enc.device)
        return self._og

    def to_c(self) -> L.RenderTapeC:
        return L.RenderTapeC(self.capacity, self.max_chunks, L.ptr(self.counter), L.ptr(self.enc), L.ptr(self.pos),
                             L.ptr(self.sample), L.ptr(self.ray_chunks), L.ptr(self.ray_nchunks))

    @classmethod
    def acquire(cls, march: "MarchSpec", radius: float, n_rays: int, device) -> "RenderTape":
        lib = L.load()
        m = march.to_c()
        cap, mc = C.c_longlong(), C.c_int()
        L.check(lib.sdb_render_tape_geometry(C.byref(m), float(radius), int(n_rays), C.byref(cap), C.byref(mc)),
                "sdb_render_tape_geometry")
        if cap.value > (1 << 27):
            raise RuntimeError(f"render tape of {cap.value} samples exceeds 2^27 slots; split the ray batch")
        key = (int(cap.value), int(mc.value), int(n_rays), str(device))
        free = cls._pool.setdefault(key, [])
        tape = free.pop() if free else cls(int(cap.value), int(mc.value), int(n_rays), device)
        tape.key = key
        return tape

    def release(self) -> None:
        if self.key is not None:
            self._pool.setdefault(self.key, []).append(self)
            self.key = None

    def check_overflow(self) -> None:
        if int(self.counter[1].item()) != 0:
            raise RuntimeError("render tape overflow: kept samples exceeded the worst-case capacity")


def render_forward_v2_raw(spec: FieldSpec, march: MarchSpec, params, occ: OccGrid, rays_o, rays_d, jitter, bg_override,
                          rays_per_image: int, tape: Optional[RenderTape]):
    """Runs sdb_render_nerf_forward_v2 (tape may be None: no gradient wanted)."""
    lib = L.load()
    f = make_field_c(spec, params)
    m = march.to_c()
    dev = rays_o.device
    n = rays_o.shape[0]
    out = {k: torch.empty(n, 3, device=dev) for k in ("comp_rgb", "comp_rgb_fg", "comp_rgb_bg")}
    out.update({k: torch.empty(n, device=dev) for k in ("opacity", "depth", "z_variance")})
    work = torch.zeros(1, dtype=torch.int32, device=dev)
    tc = tape.to_c() if tape is not None else None
    L.check(lib.sdb_render_nerf_forward_v2(
        C.byref(f), C.byref(m), L.ptr(occ.bits), L.ptr(occ.mean), L.ptr(rays_o), L.ptr(rays_d), L.ptr(jitter),
        L.ptr(bg_override), n, int(rays_per_image), L.ptr(out["comp_rgb"]), L.ptr(out["comp_rgb_fg"]),
        L.ptr(out["comp_rgb_bg"]), L.ptr(out["opacity"]), L.ptr(out["depth"]), L.ptr(out["z_variance"]),
        C.byref(tc) if tc is not None else None, L.ptr(work), L.stream_ptr()), "sdb_render_nerf_forward_v2")
    out["packed"] = None
    return out


def _grads_c(grads) -> L.FieldGradsC:
    g = L.FieldGradsC()
    for name, key in (("table", "table"), ("w1_density", "w1d"), ("w2_density", "w2d"), ("w1_feature", "w1f"),
                      ("w2_feature", "w2f"), ("bg_table", "bg_table"), ("bg_w1", "bg_w1"), ("bg_w2", "bg_w2"),
                      ("bg_w3", "bg_w3")):
        setattr(g, name, L.ptr(grads[key]))
    return g


def render_orient_forward_raw(spec: FieldSpec, params, rays_d, tape: RenderTape) -> torch.Tensor:
    """sdb_render_orient_forward on the tape of the forward that just ran: per-ray sum_i w_i relu(n_i . d)^2 (the
    numerator of loss_orient, scaledreamer.py:70-80); the partial derivatives stay in tape.og for the backward."""
    f = make_field_c(spec, params)
    orient = torch.empty(rays_d.shape[0], device=rays_d.device)
    tc = tape.to_c()
    L.check(L.load().sdb_render_orient_forward(C.byref(f), L.ptr(rays_d), rays_d.shape[0], C.byref(tc), L.ptr(orient),
                                               L.ptr(tape.og), L.stream_ptr()), "sdb_render_orient_forward")
    return orient


def render_orient_backward_raw(spec: FieldSpec, params, grads, tape: RenderTape, g_orient: torch.Tensor) -> None:
    f = make_field_c(spec, params)
    g = _grads_c(grads)
    tc = tape.to_c()
    L.check(L.load().sdb_render_orient_backward(C.byref(f), C.byref(g), g_orient.numel(), C.byref(tc), L.ptr(tape.og),
                                                L.ptr(g_orient), L.stream_ptr()), "sdb_render_orient_backward")


def render_backward_tape_raw(spec: FieldSpec, march: MarchSpec, params, grads, rays_d, bg_override,
                             rays_per_image: int, saved, tape: RenderTape, g_comp_rgb, g_opacity=None,
                             g_depth=None, g_z_variance=None) -> None:
    """Runs sdb_render_nerf_backward_tape; accumulates into `grads` (dict keyed like params)."""
    lib = L.load()
    f = make_field_c(spec, params)
    m = march.to_c()
    g = L.FieldGradsC()
    for name, key in (("table", "table"), ("w1_density", "w1d"), ("w2_density", "w2d"), ("w1_feature", "w1f"),
                      ("w2_feature", "w2f"), ("bg_table", "bg_table"), ("bg_w1", "bg_w1"), ("bg_w2", "bg_w2"),
                      ("bg_w3", "bg_w3")):
        setattr(g, name, L.ptr(grads[key]))
    tc = tape.to_c()
    zv = saved.get("z_variance") if g_z_variance is not None else None
    L.check(lib.sdb_render_nerf_backward_tape_zv(
        C.byref(f), C.byref(g), C.byref(m), L.ptr(rays_d), L.ptr(bg_override), rays_d.shape[0], int(rays_per_image),
        L.ptr(saved["comp_rgb_fg"]), L.ptr(saved["comp_rgb_bg"]), L.ptr(saved["opacity"]), L.ptr(saved["depth"]),
        L.ptr(zv), L.ptr(g_comp_rgb), L.ptr(g_opacity), L.ptr(g_depth), L.ptr(g_z_variance), C.byref(tc),
        L.stream_ptr()), "sdb_render_nerf_backward_tape_zv")


class _RenderNeRF(torch.autograd.Function):
    """comp_rgb, opacity, depth = render(params...) with the fused backward. Gradient flows to the nine field
    parameter tensors only (rays are data). Default path: v2 forward + sample tape + tape backward. When per-sample
    extras are requested (packed_capacity > 0) the v1 kernels run instead (forward with packed outputs, backward by
    re-marching)."""

    @staticmethod
    def forward(ctx, spec, march, occ, rays_o, rays_d, jitter, bg_override, rays_per_image, packed_capacity,
                holder, *param_tensors):
        params = dict(zip(PARAM_KEYS, param_tensors))
        ctx.spec, ctx.march, ctx.occ, ctx.rpi = spec, march, occ, rays_per_image
        ctx.has_jitter, ctx.has_bg = jitter is not None, bg_override is not None
        ctx.tape = None
        ctx.set_materialize_grads(False)
        want_orient = bool(holder.pop("want_orient", False))
        ctx.has_orient = False
        need_grad = any(ctx.needs_input_grad[10:])
        if packed_capacity > 0:
            out = render_forward_raw(spec, march, params, occ, rays_o, rays_d, jitter, bg_override, rays_per_image,
                                     packed_capacity)
            ctx.bits_snapshot = occ.bits.clone()  # the grid may be refreshed before backward runs
            ctx.mean_snapshot = occ.mean.clone()
        else:
            if need_grad:
                ctx.tape = RenderTape.acquire(march, spec.radius, rays_o.shape[0], rays_o.device)
            out = render_forward_v2_raw(spec, march, params, occ, rays_o, rays_d, jitter, bg_override,
                                        rays_per_image, ctx.tape)
        # orientation term (scaledreamer.py:70-80) on the tape, before anything can overwrite its raw densities
        if want_orient and ctx.tape is not None:
            orient = render_orient_forward_raw(spec, params, rays_d, ctx.tape)
            ctx.has_orient = True
        else:
            orient = torch.zeros(rays_o.shape[0], device=rays_o.device)
        ctx.save_for_backward(rays_o, rays_d, jitter if jitter is not None else rays_o.new_zeros(0),
                              bg_override if bg_override is not None else rays_o.new_zeros(0),
                              out["comp_rgb_fg"], out["comp_rgb_bg"], out["opacity"], out["depth"], out["z_variance"],
                              *param_tensors)
        holder.update(out)  # non-differentiable extras (fg / bg / packed) for the caller
        holder["tape"] = ctx.tape
        non_diff = [out["comp_rgb_fg"], out["comp_rgb_bg"]]
        if ctx.tape is None:  # the re-marching (v1) backward has no z-variance / orientation gradient
            non_diff += [out["z_variance"], orient]
        elif not ctx.has_orient:
            non_diff.append(orient)
        ctx.mark_non_differentiable(*non_diff)
        return out["comp_rgb"], out["opacity"], out["depth"], out["z_variance"], orient

    @staticmethod
    def backward(ctx, g_rgb, g_op, g_depth, g_zv=None, g_orient=None):
        rays_o, rays_d, jitter, bg_override, fg, bg, op, depth, zvar, *param_tensors = ctx.saved_tensors
        params = dict(zip(PARAM_KEYS, param_tensors))
        grads = {k: torch.zeros_like(v) for k, v in params.items()}
        cg = lambda g: g.contiguous().float() if g is not None else None
        g_rgb = g_rgb.contiguous() if g_rgb is not None else torch.zeros_like(fg)
        g_op, g_depth, g_zv, g_orient = cg(g_op), cg(g_depth), cg(g_zv), cg(g_orient)
        saved = {"comp_rgb_fg": fg, "comp_rgb_bg": bg, "opacity": op, "depth": depth, "z_variance": zvar}
        if ctx.tape is not None:
            if ctx.has_orient and g_orient is not None:
                render_orient_backward_raw(ctx.spec, params, grads, ctx.tape, g_orient)
            render_backward_tape_raw(ctx.spec, ctx.march, params, grads, rays_d, bg_override if ctx.has_bg else None,
                                     ctx.rpi, saved, ctx.tape, g_rgb, g_op, g_depth, g_zv)
            ctx.tape.release()
            ctx.tape = None
        else:
            occ = ctx.occ
            snap = OccGrid.__new__(OccGrid)
            snap.res, snap.bits, snap.mean, snap.occs = occ.res, ctx.bits_snapshot, ctx.mean_snapshot, occ.occs
            render_backward_raw(ctx.spec, ctx.march, params, grads, snap, rays_o, rays_d,
                                jitter if ctx.has_jitter else None, bg_override if ctx.has_bg else None, ctx.rpi,
                                saved, g_rgb, g_op, g_depth)
        return (None,) * 10 + tuple(grads[k] for k in PARAM_KEYS)


def render_nerf(spec, march, occ, params, rays_o, rays_d, jitter, bg_override, rays_per_image, packed_capacity=0,
                want_orient=False):
    """Differentiable fused render. Returns dict with comp_rgb / opacity / depth / z_variance (grad) and the extras;
    with want_orient (tape path, gradients on) also `orient` [Nr] = sum_i w_i relu(n_i . d)^2 per ray, differentiable
    with respect to the density network and the table through the finite-difference normals."""
    holder: dict = {"want_orient": want_orient and packed_capacity == 0}
    rgb, op, depth, zvar, orient = _RenderNeRF.apply(spec, march, occ, rays_o.contiguous(), rays_d.contiguous(), jitter,
                                                     bg_override, rays_per_image, packed_capacity, holder,
                                                     *[params[k] for k in PARAM_KEYS])
    out = dict(holder)
    out["comp_rgb"], out["opacity"], out["depth"], out["z_variance"] = rgb, op, depth, zvar
    if want_orient and holder.get("tape") is not None:
        out["orient"] = orient
    return out


# ------------------------------------------------------------------------------------------------ packed (unfused) stages
def freq_encode(x01: torch.Tensor, n_frequencies: int, mask: torch.Tensor, include_xyz: bool = False,
                pad_to: int = 8) -> torch.Tensor:
    """ProgressiveBandFrequency under CompositeEncoding (networks.py:16-52, 170-190) on x01 [N,3]; the rows are
    zero-padded to a multiple of `pad_to` columns for the MLP kernels."""
    x01 = x01.reshape(-1, 3).contiguous().float()
    used = (3 if include_xyz else 0) + 6 * n_frequencies
    stride = -(-used // pad_to) * pad_to
    out = torch.empty(x01.shape[0], stride, device=x01.device)
    mask = mask.to(x01.device, torch.float32).contiguous()
    L.check(L.load().sdb_freq_encode(L.ptr(x01), x01.shape[0], int(n_frequencies), L.ptr(mask), int(include_xyz),
                                     stride, L.ptr(out), L.stream_ptr()), "sdb_freq_encode")
    return out


def _offsets(counts: torch.Tensor) -> torch.Tensor:
    off = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts, 0, out=off[1:])
    return off


def march_packed(march: MarchSpec, radius: float, occ: OccGrid, rays_o, rays_d, jitter) -> Dict[str, torch.Tensor]:
    """Candidate samples of every ray, packed and sorted by (ray, t) (nerfacc OccGridEstimator.sampling without
    sigma_fn). One host read of the total (the reference's packed tensors are sized the same way)."""
    lib = L.load()
    m = march.to_c()
    rays_o, rays_d = rays_o.reshape(-1, 3).contiguous().float(), rays_d.reshape(-1, 3).contiguous().float()
    n_rays, dev = rays_o.shape[0], rays_o.device
    counts = torch.zeros(n_rays, dtype=torch.int32, device=dev)
    L.check(lib.sdb_march_count(C.byref(m), float(radius), L.ptr(occ.bits), L.ptr(rays_o), L.ptr(rays_d),
                                L.ptr(jitter), n_rays, L.ptr(counts), L.stream_ptr()), "sdb_march_count")
    offsets = _offsets(counts)
    n = int(offsets[-1].item())
    out = {"offsets": offsets, "ray_indices": torch.empty(n, dtype=torch.int32, device=dev),
           "t_starts": torch.empty(n, device=dev), "t_ends": torch.empty(n, device=dev),
           "positions": torch.empty(n, 3, device=dev)}
    if n:
        L.check(lib.sdb_march_fill(C.byref(m), float(radius), L.ptr(occ.bits), L.ptr(rays_o), L.ptr(rays_d),
                                   L.ptr(jitter), n_rays, L.ptr(offsets), L.ptr(out["ray_indices"]),
                                   L.ptr(out["t_starts"]), L.ptr(out["t_ends"]), L.ptr(out["positions"]),
                                   L.stream_ptr()), "sdb_march_fill")
    return out


def prune_packed(samples: Dict[str, torch.Tensor], sigma: torch.Tensor, march: MarchSpec,
                 occ: Optional[OccGrid]) -> Dict[str, torch.Tensor]:
    """sigma_fn visibility pruning (nerf_volume_renderer.py:153-180): drops samples with alpha < min(alpha_thre,
    mean occupancy) or transmittance < early_stop_eps; returns the compacted packed samples."""
    n_rays = samples["offsets"].numel() - 1
    n = sigma.numel()
    dev = sigma.device
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    kept = torch.zeros(n_rays, dtype=torch.int32, device=dev)
    if n:
        L.check(L.load().sdb_packed_visibility(
            L.ptr(sigma.contiguous().float()), L.ptr(samples["t_starts"]), L.ptr(samples["t_ends"]),
            L.ptr(samples["offsets"]), n_rays, float(march.alpha_thre), L.ptr(occ.mean) if occ is not None else None,
            float(march.early_stop_eps), L.ptr(keep), L.ptr(kept), L.stream_ptr()), "sdb_packed_visibility")
    idx = torch.nonzero(keep)[:, 0]
    out = {k: samples[k][idx] for k in ("ray_indices", "t_starts", "t_ends", "positions")}
    out["offsets"] = _offsets(kept)
    return out


class _PackedComposite(torch.autograd.Function):
    """(opacity [R], depth [R], comp_rgb_fg [R,3]) = front-to-back compositing of packed samples; z_variance and the
    per-sample weights come back through `holder` (not differentiable, as the fused renderer's)."""

    @staticmethod
    def forward(ctx, sigma, rgb, t_starts, t_ends, offsets, holder):
        lib = L.load()
        n_rays, dev = offsets.numel() - 1, offsets.device
        sigma, rgb = sigma.contiguous().float(), rgb.contiguous().float()
        n = sigma.numel()
        weights, trans = torch.empty(n, device=dev), torch.empty(n, device=dev)
        op, depth = torch.empty(n_rays, device=dev), torch.empty(n_rays, device=dev)
        fg, zvar = torch.empty(n_rays, 3, device=dev), torch.empty(n_rays, device=dev)
        L.check(lib.sdb_packed_composite_forward(
            L.ptr(sigma.detach()), L.ptr(rgb.detach()), L.ptr(t_starts), L.ptr(t_ends), L.ptr(offsets), n_rays,
            L.ptr(weights), L.ptr(trans), L.ptr(op), L.ptr(depth), L.ptr(fg), L.ptr(zvar), L.stream_ptr()),
            "sdb_packed_composite_forward")
        ctx.save_for_backward(rgb.detach(), t_starts, t_ends, offsets, weights, trans)
        holder["weights"], holder["z_variance"] = weights, zvar
        return op, depth, fg

    @staticmethod
    def backward(ctx, g_op, g_depth, g_fg):
        rgb, t_starts, t_ends, offsets, weights, trans = ctx.saved_tensors
        n_rays = offsets.numel() - 1
        d_sigma, d_rgb = torch.zeros(weights.numel(), device=rgb.device), torch.zeros_like(rgb)
        cg = lambda g: g.contiguous().float() if g is not None else None
        g_op, g_depth, g_fg = cg(g_op), cg(g_depth), cg(g_fg)
        L.check(L.load().sdb_packed_composite_backward(
            L.ptr(rgb), L.ptr(t_starts), L.ptr(t_ends), L.ptr(offsets), n_rays, L.ptr(weights), L.ptr(trans),
            L.ptr(g_op), L.ptr(g_depth), L.ptr(g_fg), L.ptr(d_sigma), L.ptr(d_rgb), L.stream_ptr()),
            "sdb_packed_composite_backward")
        return d_sigma, d_rgb, None, None, None, None


def composite_packed(sigma, rgb, samples: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    holder: dict = {}
    op, depth, fg = _PackedComposite.apply(sigma.reshape(-1), rgb.reshape(-1, 3), samples["t_starts"],
                                           samples["t_ends"], samples["offsets"], holder)
    return {"opacity": op, "depth": depth, "comp_rgb_fg": fg, "weights": holder["weights"],
            "z_variance": holder["z_variance"]}


def occgrid_update_values(occ: OccGrid, cell_idx: torch.Tensor, values: torch.Tensor, ema_decay: float = 0.95,
                          occ_thre: float = 0.01) -> None:
    cell_idx, values = cell_idx.to(torch.int32).contiguous(), values.reshape(-1).float().contiguous()
    L.check(L.load().sdb_occgrid_update_values(L.ptr(cell_idx), L.ptr(values), cell_idx.numel(), occ.res,
                                               float(ema_decay), float(occ_thre), L.ptr(occ.occs), L.ptr(occ.bits),
                                               L.ptr(occ.mean), L.stream_ptr()), "sdb_occgrid_update_values")
