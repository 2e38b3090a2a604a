"""Generates the UNet / VAE-encoder golden vectors from the REFERENCE's own vendored LDM modules.

Run in the build container only (it imports /root/reference, which does not exist on the GPU box):
    python tests/golden/make_ldm_golden.py
Weights are the seeded synthetic parameters of scaledreamer_b200.nets.random_state_dict (no pretrained
checkpoints exist here); the reference modules are instantiated unchanged
(extern/mvdream/ldm/modules/diffusionmodules/openaimodel.py UNetModel / MultiViewUNetModel,
ldm/modules/diffusionmodules/model.py Encoder), loaded with those weights and run in fp32 on the CPU.
Outputs: tests/golden/ldm_golden.pt (inputs + outputs, a few hundred kB).
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

# the vendored package imports omegaconf at module level only (model_zoo.py:4, openaimodel.py:882)
om = types.ModuleType("omegaconf")
om.OmegaConf = type("OmegaConf", (), {})
om.listconfig = types.ModuleType("omegaconf.listconfig")
om.listconfig.ListConfig = list
sys.modules["omegaconf"] = om
sys.modules["omegaconf.listconfig"] = om.listconfig

from extern.mvdream.ldm.modules.diffusionmodules.model import Encoder  # noqa: E402
from extern.mvdream.ldm.modules.diffusionmodules.openaimodel import MultiViewUNetModel, UNetModel  # noqa: E402

from scaledreamer_b200.nets import random_state_dict  # noqa: E402

torch.manual_seed(0)
torch.set_grad_enabled(False)


def specs_of(module, prefix=""):
    """(name, my-layout shape) for every parameter: 3x3 conv [Co,Ci,3,3] -> [Co,3,3,Ci]; 1x1 conv -> [Co,Ci]."""
    out = []
    for k, v in module.state_dict().items():
        s = tuple(v.shape)
        if len(s) == 4 and s[2] == 3:
            s = (s[0], 3, 3, s[1])
        elif len(s) == 4:
            s = (s[0], s[1])
        out.append((prefix + k, s))
    return out


def load_seeded(module, seed, prefix=""):
    sd = random_state_dict(specs_of(module, prefix), seed)
    ref = module.state_dict()
    # fp16-rounded weights: the product path (and the reference's half_precision_weights) stores fp16
    module.load_state_dict({k: sd[prefix + k].reshape(ref[k].shape).half().float() for k in ref})


def unet_case(multiview, B, HW, seed):
    kw = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
              num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
              use_linear_in_transformer=True, transformer_depth=1, context_dim=1024, use_checkpoint=False, legacy=False)
    net = (MultiViewUNetModel(camera_dim=16, **kw) if multiview else UNetModel(**kw)).eval()
    load_seeded(net, seed)
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.randn(B, 4, HW, HW, generator=g).half().float()
    t = torch.tensor([20.0, 250.0, 500.0, 750.0, 980.0, 333.0, 111.0, 999.0])[:B]
    ctx = torch.randn(B, 77, 1024, generator=g).half().float()
    case = dict(x=x, t=t, ctx=ctx, seed=seed)
    if multiview:
        cam = torch.randn(B, 16, generator=g).half().float()
        case["camera"] = cam
        y = net(x, t, context=ctx, camera=cam, num_frames=4)
    else:
        y = net(x, t, context=ctx)
    case["y"] = y
    print("unet", "mv" if multiview else "sd", tuple(y.shape), float(y.abs().mean()), float(y.std()))
    return case


def vae_case(B, HW, seed):
    enc = Encoder(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0,
                  in_channels=3, resolution=256, z_channels=4, double_z=True).eval()
    load_seeded(enc, seed, prefix="encoder.")
    g = torch.Generator().manual_seed(200 + seed)
    x = (torch.rand(B, 3, HW, HW, generator=g) * 2 - 1).requires_grad_(True)
    with torch.enable_grad():
        h = enc(x)
        d_h = torch.randn(h.shape, generator=g)
        (h * d_h).sum().backward()
    print("vae", tuple(h.shape), float(h.abs().mean()), float(x.grad.abs().mean()))
    return dict(x=x.detach(), h=h.detach(), d_h=d_h, d_x=x.grad.detach(), seed=seed)


if __name__ == "__main__":
    out = {"unet_sd": unet_case(False, 2, 16, 1), "unet_mv": unet_case(True, 4, 8, 2), "vae": vae_case(1, 64, 3)}
    for v in out.values():  # inputs are fp16-exact: store them compactly
        for kk in ('x', 'ctx', 'camera'):
            if kk in v and kk != 'x':
                v[kk] = v[kk].half()
    torch.save(out, os.path.join(ROOT, "tests", "golden", "ldm_golden.pt"))
    print("saved", os.path.getsize(os.path.join(ROOT, "tests", "golden", "ldm_golden.pt")), "bytes")
