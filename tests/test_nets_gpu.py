"""GPU parity of the native UNet / VAE-encoder executors against golden vectors produced by the REFERENCE's own
vendored LDM modules (tests/golden/make_ldm_golden.py, fp32 CPU, same fp16-rounded weights and inputs).

Tolerance: north_star asks eps-pred within 1e-3 relative of the reference path. The reference SD path itself runs
fp16 weights/activations (stable_diffusion_asd_guidance.py:57-59); our path stores activations in fp16 with fp32
accumulation, so against the fp32 golden the expected error is the accumulated fp16 rounding of ~100 layers.
The bound below (relative L2) is what is asserted; the measured value is printed.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ldm_golden.pt")


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


@pytest.mark.parametrize("which", ["unet_sd", "unet_mv"])
def test_unet_matches_reference_ldm(cuda_device, gold, which):
    from scaledreamer_b200 import nets

    c = gold[which]
    B, _, H, W = c["x"].shape
    cfg = nets.MVDREAM_UNET if which == "unet_mv" else nets.SD21_UNET
    net = nets.UNet(cfg, B, H, W, cuda_device)
    assert net.num_parameters() == (867_572_164 if which == "unet_mv" else 865_910_724)
    net.load_state_dict(nets.random_state_dict(net.specs, c["seed"]))
    x = c["x"].permute(0, 2, 3, 1).contiguous().half().to(cuda_device)
    cam = c["camera"].to(cuda_device) if "camera" in c else None
    y = net.forward(x, c["t"].to(cuda_device), c["ctx"].half().to(cuda_device), cam)
    torch.cuda.synchronize()
    err = rel(y.permute(0, 3, 1, 2), c["y"])
    print(f"{which}: rel_l2 = {err:.3e}, launches = {net.launches()}")
    assert torch.isfinite(y).all()
    assert err < 3e-3
    # replay is bitwise reproducible (no atomics on the forward path)
    y2 = net.forward(x, c["t"].to(cuda_device), c["ctx"].half().to(cuda_device), cam)
    assert torch.equal(y2, y)


def test_vae_encoder_forward_backward_match_reference_ldm(cuda_device, gold):
    from scaledreamer_b200 import nets

    c = gold["vae"]
    B, _, H, W = c["x"].shape
    net = nets.VaeEncoder(nets.SD_VAE, B, H, W, cuda_device)
    assert net.num_parameters() == 34_163_592
    net.load_state_dict(nets.random_state_dict(net.specs, c["seed"]))
    x = c["x"].permute(0, 2, 3, 1).contiguous().to(cuda_device)
    h = net.forward(x)
    e_f = rel(h.permute(0, 3, 1, 2), c["h"])
    d_x = net.backward(c["d_h"].permute(0, 2, 3, 1).contiguous().to(cuda_device))
    torch.cuda.synchronize()
    e_b = rel(d_x.permute(0, 3, 1, 2), c["d_x"])
    print(f"vae: forward rel_l2 = {e_f:.3e}, backward rel_l2 = {e_b:.3e}, launches = {net.launches()}/{net.launches(True)}")
    assert e_f < 3e-3
    assert e_b < 1e-2
