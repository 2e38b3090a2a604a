"""Pins the dense-path oracle (oracle/ldm_oracle.py) to golden vectors produced by the REFERENCE's own vendored LDM
modules (tests/golden/make_ldm_golden.py): UNetModel, MultiViewUNetModel and the VAE Encoder, forward and (VAE)
data gradient, fp32 CPU."""
import os

import pytest
import torch

from oracle import ldm_oracle as lo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ldm_golden.pt")


def _specs_from_lib(kind, B, H, W):
    """Parameter names/shapes come from the library's own enumeration (host-only call, no GPU needed)."""
    import ctypes as C

    from scaledreamer_b200 import lib as L, nets

    lib = L.load()
    h = C.c_void_p()
    if kind == "vae":
        c = L.VaeCfgC(3, 128, 4, (C.c_int * 4)(1, 2, 4, 4), 2, 4)
        L.check(lib.sdb_vae_encoder_create(C.byref(c), B, H, W, C.byref(h)), "create")
    else:
        cfg = nets.MVDREAM_UNET if kind == "unet_mv" else nets.SD21_UNET
        c = L.UNetCfgC(4, 4, 320, 4, (C.c_int * 4)(1, 2, 4, 4), 2, 3, 64, 1024, 77, cfg["camera_dim"], cfg["num_frames"])
        L.check(lib.sdb_unet_create(C.byref(c), B, H, W, C.byref(h)), "create")
    name, ndim, shape = C.c_char_p(), C.c_int(), (C.c_int * 4)()
    specs = []
    for i in range(lib.sdb_net_num_params(h)):
        L.check(lib.sdb_net_param(h, i, C.byref(name), C.byref(ndim), shape), "param")
        specs.append((name.value.decode(), tuple(shape[: ndim.value])))
    lib.sdb_net_destroy(h)
    return specs


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("which", ["unet_sd", "unet_mv"])
def test_unet_oracle_matches_reference_modules(which):
    from scaledreamer_b200 import nets

    c = torch.load(GOLD)[which]
    B, _, H, W = c["x"].shape
    sd = {k: v.half().float() for k, v in nets.random_state_dict(_specs_from_lib(which, B, H, W), c["seed"]).items()}
    cfg = dict(lo.SD, camera_dim=16, num_frames=4) if which == "unet_mv" else lo.SD
    cam = c["camera"].float() if "camera" in c else None
    with torch.no_grad():
        y = lo.unet_forward(sd, c["x"], c["t"], c["ctx"].float(), cam, cfg)
    assert rel(y, c["y"]) < 1e-4


def test_vae_oracle_matches_reference_modules():
    from scaledreamer_b200 import nets

    c = torch.load(GOLD)["vae"]
    B, _, H, W = c["x"].shape
    sd = {k: v.half().float() for k, v in nets.random_state_dict(_specs_from_lib("vae", B, H, W), c["seed"]).items()}
    x = c["x"].clone().requires_grad_(True)
    h = lo.vae_encoder_forward(sd, x)
    (h * c["d_h"]).sum().backward()
    assert rel(h.detach(), c["h"]) < 1e-4
    assert rel(x.grad, c["d_x"]) < 1e-4


def test_schedule_and_t_plus_invariants():
    ac = lo.alphas_cumprod()
    assert ac.shape == (1000,) and abs(float(ac[0]) - 0.99915) < 1e-5 and abs(float(ac[-1]) - 0.004660) < 1e-5
    t = torch.tensor([20, 500, 980, 999])
    tp = lo.t_plus(t, torch.tensor([0.999, 0.5, 0.999, 0.3]), 0.1, 20)
    assert (tp >= t).all() and (tp <= 999).all() and tp[0] == 20 and tp[1] == 524
    x, y = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    p = lo.perpendicular_component(x, y)
    assert torch.allclose((p * y).sum(dim=[1, 2, 3]), torch.zeros(2), atol=1e-4)
