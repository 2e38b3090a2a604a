"""GPU probe: times the fused render forward/backward at 256x256 (C2 shape) with CUDA events."""
import math, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scaledreamer_b200 import render_ops as R, lib as L


def make_scene(dev, H=256, W=256, B=1, seed=0, table_scale=1e-4, n_samples=512):
    g = torch.Generator().manual_seed(seed)
    grid = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, per_level_scale=1.447269237440378)
    bg_grid = dict(n_levels=4, n_features_per_level=2, log2_hashmap_size=19, base_resolution=4, per_level_scale=4.0)
    spec = R.FieldSpec(grid=grid, bg_grid=bg_grid)
    n, nb = L.grid_num_entries(grid), L.grid_num_entries(bg_grid)
    lin = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) / math.sqrt(i)
    P = {"table": (torch.rand(n, 2, generator=g) * 2 - 1) * table_scale, "w1d": lin(64, 32), "w2d": lin(1, 64),
         "w1f": lin(64, 32), "w2f": lin(3, 64), "bg_table": (torch.rand(nb, 2, generator=g) * 2 - 1) * table_scale,
         "bg_w1": lin(16, 8), "bg_w2": lin(16, 16), "bg_w3": lin(3, 16)}
    P = {k: v.to(dev).contiguous() for k, v in P.items()}
    elev = torch.rand(B, generator=g) * 55 - 10
    azim = torch.rand(B, generator=g) * 360 - 180
    dist = torch.rand(B, generator=g) * 0.5 + 1.0
    fovy = torch.deg2rad(torch.rand(B, generator=g) * 30 + 40)
    e, a = torch.deg2rad(elev), torch.deg2rad(azim)
    pos = torch.stack([dist * torch.cos(e) * torch.cos(a), dist * torch.cos(e) * torch.sin(a), dist * torch.sin(e)], -1)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(pos)
    nz = torch.nn.functional.normalize
    look = nz(-pos, dim=-1); right = nz(torch.cross(look, up, dim=-1), dim=-1); up2 = nz(torch.cross(right, look, dim=-1), dim=-1)
    c2w = torch.zeros(B, 4, 4); c2w[:, :3, 0], c2w[:, :3, 1], c2w[:, :3, 2], c2w[:, :3, 3] = right, up2, -look, pos; c2w[:, 3, 3] = 1
    o = torch.empty(B, H, W, 3, device=dev); d = torch.empty_like(o)
    L.check(L.load().sdb_raygen(L.ptr(c2w.to(dev)), L.ptr(fovy.to(dev)), B, H, W, L.ptr(o), L.ptr(d), L.stream_ptr()), "raygen")
    march = R.MarchSpec(render_step_size=1.732 * 2 * 1.0 / n_samples)
    return spec, march, P, o.reshape(-1, 3), d.reshape(-1, 3), torch.rand(B * H * W, generator=g).to(dev)


def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[n // 2]


def main():
    dev = torch.device("cuda:0")
    res = {}
    for name, prune_grid in (("occgrid", True), ("allcells", False)):
        spec, march, P, o, d, jit = make_scene(dev)
        occ = R.OccGrid(32, dev, all_occupied=not prune_grid)
        if prune_grid:
            idx = torch.arange(32 ** 3, device=dev)
            for _ in range(2):
                occ.update(spec, P, idx, torch.rand(32 ** 3, 3, device=dev), march.render_step_size)
        cap = 1 << 25
        out = R.render_forward_raw(spec, march, P, occ, o, d, jit, None, 65536, cap)
        ns = int(out["packed"]["counter"].item())
        del out
        grads = {k: torch.zeros_like(v) for k, v in P.items()}
        f = lambda: R.render_forward_raw(spec, march, P, occ, o, d, jit, None, 65536, 0)
        out = f()
        g_rgb = torch.randn_like(out["comp_rgb"])
        b = lambda: R.render_backward_raw(spec, march, P, grads, occ, o, d, jit, None, 65536, out, g_rgb)
        tf, tb = timeit(f), timeit(b)
        res[name] = dict(kept_samples=ns, occ_cells=int(occ.binaries().sum()), fwd_ms=tf, bwd_ms=tb,
                         opacity_mean=float(out["opacity"].mean()))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
