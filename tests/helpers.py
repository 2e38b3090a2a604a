"""Shared scene builders for the parity tests (seeded synthetic inputs)."""
import math

import torch

from oracle import render_oracle as ro


def scene(H=32, W=32, B=1, seed=0, table_scale=0.5, prune=True, n_samples=512, fcfg=None):
    """Random field + random cameras + one warm-up occupancy refresh, all on CPU (oracle side)."""
    fcfg = ro.FieldCfg() if fcfg is None else fcfg
    mcfg = ro.MarchCfg(render_step_size=1.732 * 2 * fcfg.radius / n_samples, prune=prune)
    P = ro.make_field_params(fcfg, seed=seed, table_scale=table_scale)
    g = torch.Generator().manual_seed(seed + 1)
    elev = torch.rand(B, generator=g) * 55 - 10
    azim = (torch.rand(B, generator=g) + torch.arange(B)) / B * 360 - 180
    dist = torch.rand(B, generator=g) * 0.5 + 1.0
    fovy = torch.deg2rad(torch.rand(B, generator=g) * 30 + 40)
    c2w = ro.look_at_c2w(elev, azim, dist)
    rays_o, rays_d = ro.get_rays(c2w, fovy, H, W)
    jitter = torch.rand(B * H * W, generator=g)
    occs, binary, cell_rand = ro.occ_grid_from_density(P, fcfg, mcfg, seed=seed + 2)
    if not prune:
        binary = torch.ones_like(binary)
    return dict(fcfg=fcfg, mcfg=mcfg, P=P, c2w=c2w, fovy=fovy, rays_o=rays_o.reshape(-1, 3),
                rays_d=rays_d.reshape(-1, 3), jitter=jitter, occs=occs, binary=binary, cell_rand=cell_rand,
                H=H, W=W, B=B, elev=elev, azim=azim, dist=dist)


def field_spec_from_oracle(fcfg):
    from scaledreamer_b200.render_ops import FieldSpec

    return FieldSpec(grid=vars(fcfg.grid), bg_grid=vars(fcfg.bg_grid), radius=fcfg.radius,
                     density_bias=fcfg.density_bias, density_blob_scale=fcfg.density_blob_scale,
                     density_blob_std=fcfg.density_blob_std, density_activation=fcfg.density_activation,
                     fd_eps=fcfg.fd_eps, color_activation=fcfg.color_activation,
                     bg_color_activation=fcfg.bg_color_activation)


def march_spec_from_oracle(mcfg, output_normal=False):
    from scaledreamer_b200.render_ops import MarchSpec

    return MarchSpec(render_step_size=mcfg.render_step_size, near_plane=mcfg.near_plane, far_plane=mcfg.far_plane,
                     prune=mcfg.prune, alpha_thre=mcfg.alpha_thre, early_stop_eps=mcfg.early_stop_eps,
                     grid_res=mcfg.grid_res, output_normal=output_normal)


def rel_l2(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
