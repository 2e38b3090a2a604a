"""CPU oracle for the dense half of the ASD step.  TEST INFRASTRUCTURE ONLY (see oracle/render_oracle.py header).

Plain fp32 PyTorch restatement of the frozen networks and of the guidance arithmetic, driven by a flat
{name: tensor} state dict with the reference's key names and layouts:
  unet_forward        UNetModel / MultiViewUNetModel.forward  extern/mvdream/ldm/modules/diffusionmodules/openaimodel.py:777-808, 1175-1213
                      (ResBlock :255-275, Down/Upsample :91-160, SpatialTransformer(3D) attention.py:393-412,
                       BasicTransformerBlock(3D) :271-275/:348-354, CrossAttention :163-194, GEGLU :49-57)
  vae_encoder_forward Encoder.forward  ldm/modules/diffusionmodules/model.py:518-543 (ResnetBlock :129-149,
                      Downsample :80-87, AttnBlock :179-203)
  asd_*               stable_diffusion_asd_guidance.py:211-316,333-428 / mvdream_asd_guidance.py:141-304

PARITY STATUS: **pinned**.  Networks: tests/test_oracle_ldm.py checks this file against tests/golden/ldm_golden.pt,
which tests/golden/make_ldm_golden.py produced by running the REFERENCE's own vendored modules in this container.
Guidance arithmetic (schedule, t+dt, q-sample batch, CFG / Perp-Neg, w(t), loss and gradient):
tests/test_host_golden_cpu.py checks it against tests/golden/host_golden.pt, produced by the reference's own SD and
MVDream guidance `__call__` / `get_t_plus` / `get_eps` methods (taken out of the guidance files with `ast`, executed
unchanged with a recorded UNet output; tests/golden/make_host_golden.py).  Not executable here: the diffusers VAE / UNet
classes themselves (their vendored LDM twins are) and DDPMScheduler.add_noise (== q_sample on the same schedule).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = dict(in_channels=4, out_channels=4, model_channels=320, num_levels=4, channel_mult=(1, 2, 4, 4), num_res_blocks=2,
          attn_levels=3, head_dim=64, context_dim=1024, camera_dim=0, num_frames=1)


def _gn(x, sd, name, eps):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps)


def _conv(x, sd, name, **kw):
    w = sd[name + ".weight"]
    if w.dim() == 2:  # 1x1 convolution stored as [out, in]
        w = w.reshape(w.shape[0], w.shape[1], 1, 1)
    return F.conv2d(x, w, sd.get(name + ".bias"), **kw)


def _lin(x, sd, name):
    w = sd[name + ".weight"]
    return F.linear(x, w.reshape(w.shape[0], -1), sd.get(name + ".bias"))


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _resblock(x, emb, sd, name):
    h = _conv(F.silu(_gn(x, sd, name + ".in_layers.0", 1e-5)), sd, name + ".in_layers.2", padding=1)
    h = h + _lin(F.silu(emb), sd, name + ".emb_layers.1")[:, :, None, None]
    h = _conv(F.silu(_gn(h, sd, name + ".out_layers.0", 1e-5)), sd, name + ".out_layers.3", padding=1)
    if name + ".skip_connection.weight" in sd:
        x = _conv(x, sd, name + ".skip_connection")
    return x + h


def _attn(q, k, v, heads):
    b, lq, c = q.shape
    sp = lambda t: t.view(b, t.shape[1], heads, c // heads).transpose(1, 2)
    s = torch.einsum("bhid,bhjd->bhij", sp(q), sp(k)) * (c // heads) ** -0.5
    o = torch.einsum("bhij,bhjd->bhid", s.softmax(-1), sp(v))
    return o.transpose(1, 2).reshape(b, lq, c)


def _transformer(x, ctx, sd, name, head_dim, num_frames):
    b, c, h, w = x.shape
    heads = c // head_dim
    t = _gn(x, sd, name + ".norm", 1e-6).permute(0, 2, 3, 1).reshape(b, h * w, c)
    t = _lin(t, sd, name + ".proj_in")
    blk = name + ".transformer_blocks.0"
    ln = lambda z, n: F.layer_norm(z, (c,), sd[blk + n + ".weight"], sd[blk + n + ".bias"])
    n1 = ln(t, ".norm1").reshape(b // num_frames, num_frames * h * w, c)  # self-attention across the frames
    a = _attn(_lin(n1, sd, blk + ".attn1.to_q"), _lin(n1, sd, blk + ".attn1.to_k"), _lin(n1, sd, blk + ".attn1.to_v"), heads)
    t = _lin(a, sd, blk + ".attn1.to_out.0").reshape(b, h * w, c) + t
    n2 = ln(t, ".norm2")
    a = _attn(_lin(n2, sd, blk + ".attn2.to_q"), _lin(ctx, sd, blk + ".attn2.to_k"), _lin(ctx, sd, blk + ".attn2.to_v"), heads)
    t = _lin(a, sd, blk + ".attn2.to_out.0") + t
    g = _lin(ln(t, ".norm3"), sd, blk + ".ff.net.0.proj")
    a, gate = g.chunk(2, dim=-1)
    t = _lin(a * F.gelu(gate), sd, blk + ".ff.net.2") + t
    t = _lin(t, sd, name + ".proj_out")
    return t.reshape(b, h, w, c).permute(0, 3, 1, 2) + x


def unet_forward(sd: Dict[str, torch.Tensor], x, t, ctx, camera=None, cfg=SD):
    """x [B,4,H,W] NCHW, t [B], ctx [B,77,1024] -> eps [B,4,H,W]."""
    mc, hd, nf = cfg["model_channels"], cfg["head_dim"], cfg["num_frames"]
    emb = _lin(F.silu(_lin(timestep_embedding(t, mc), sd, "time_embed.0")), sd, "time_embed.2")
    if cfg["camera_dim"]:
        emb = emb + _lin(F.silu(_lin(camera, sd, "camera_embed.0")), sd, "camera_embed.2")
    h = _conv(x, sd, "input_blocks.0.0", padding=1)
    hs = [h]
    ib = 1
    for level in range(cfg["num_levels"]):
        for _ in range(cfg["num_res_blocks"]):
            h = _resblock(h, emb, sd, f"input_blocks.{ib}.0")
            if level < cfg["attn_levels"]:
                h = _transformer(h, ctx, sd, f"input_blocks.{ib}.1", hd, nf)
            hs.append(h)
            ib += 1
        if level != cfg["num_levels"] - 1:
            h = _conv(h, sd, f"input_blocks.{ib}.0.op", stride=2, padding=1)
            hs.append(h)
            ib += 1
    h = _resblock(h, emb, sd, "middle_block.0")
    h = _transformer(h, ctx, sd, "middle_block.1", hd, nf)
    h = _resblock(h, emb, sd, "middle_block.2")
    ob = 0
    for level in reversed(range(cfg["num_levels"])):
        for i in range(cfg["num_res_blocks"] + 1):
            h = _resblock(torch.cat([h, hs.pop()], dim=1), emb, sd, f"output_blocks.{ob}.0")
            sub = 1
            if level < cfg["attn_levels"]:
                h = _transformer(h, ctx, sd, f"output_blocks.{ob}.1", hd, nf)
                sub = 2
            if level and i == cfg["num_res_blocks"]:
                h = _conv(F.interpolate(h, scale_factor=2, mode="nearest"), sd, f"output_blocks.{ob}.{sub}.conv", padding=1)
            ob += 1
    return _conv(F.silu(_gn(h, sd, "out.0", 1e-5)), sd, "out.2", padding=1)


def _vae_resblock(x, sd, name):
    h = _conv(F.silu(_gn(x, sd, name + ".norm1", 1e-6)), sd, name + ".conv1", padding=1)
    h = _conv(F.silu(_gn(h, sd, name + ".norm2", 1e-6)), sd, name + ".conv2", padding=1)
    if name + ".nin_shortcut.weight" in sd:
        w = sd[name + ".nin_shortcut.weight"]
        x = F.conv2d(x, w.reshape(w.shape[0], -1, 1, 1), sd[name + ".nin_shortcut.bias"])
    return x + h


def vae_encoder_forward(sd: Dict[str, torch.Tensor], x, ch_mult=(1, 2, 4, 4), num_res_blocks=2):
    """x [B,3,H,W] in [-1,1] -> conv_out output [B,8,H/8,W/8] (before quant_conv)."""
    c1 = lambda z, n: F.conv2d(z, sd[n + ".weight"].reshape(sd[n + ".weight"].shape[0], -1, 1, 1), sd[n + ".bias"])
    h = _conv(x, sd, "encoder.conv_in", padding=1)
    for level in range(len(ch_mult)):
        for r in range(num_res_blocks):
            h = _vae_resblock(h, sd, f"encoder.down.{level}.block.{r}")
        if level != len(ch_mult) - 1:
            h = _conv(F.pad(h, (0, 1, 0, 1)), sd, f"encoder.down.{level}.downsample.conv", stride=2)
    h = _vae_resblock(h, sd, "encoder.mid.block_1")
    n = "encoder.mid.attn_1"
    b, c, hh, ww = h.shape
    z = _gn(h, sd, n + ".norm", 1e-6)
    q, k, v = (c1(z, n + s).reshape(b, c, hh * ww).permute(0, 2, 1) for s in (".q", ".k", ".v"))
    a = torch.bmm((torch.bmm(q, k.transpose(1, 2)) * c ** -0.5).softmax(-1), v)
    h = h + c1(a.permute(0, 2, 1).reshape(b, c, hh, ww), n + ".proj_out")
    h = _vae_resblock(h, sd, "encoder.mid.block_2")
    return _conv(F.silu(_gn(h, sd, "encoder.norm_out", 1e-6)), sd, "encoder.conv_out", padding=1)


# ------------------------------------------------------------------------------------------------ guidance
def alphas_cumprod():
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0).float()


def t_plus(t, u, plus_ratio, min_step, T=1000):
    """second get_t_plus definition (stable_diffusion_asd_guidance.py:294-316)."""
    tp = (plus_ratio * (t - min_step)).float()
    tp = torch.minimum(torch.maximum(tp, torch.zeros_like(tp)), (T - t - 1).float())
    if u is not None:
        tp = tp * u
    return torch.clamp(t + tp.to(torch.long), 1, T - 1)


def sample_latents(h, quant_w, quant_b, eps_post, sf=0.18215):
    """h [B,8,H,W] -> z [B,4,H,W] = (mean + std*eps) * sf after quant_conv (autoencoder.py:81-85, distributions.py)."""
    m = F.conv2d(h, quant_w.reshape(8, 8, 1, 1), quant_b)
    mean, logvar = m.chunk(2, dim=1)
    return (mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * eps_post) * sf


def perpendicular_component(x, y):
    eps = torch.ones_like(x[:, 0, 0, 0]) * 1e-6
    return x - ((x * y).sum(dim=[1, 2, 3]) / torch.maximum((y * y).sum(dim=[1, 2, 3]), eps)).view(-1, 1, 1, 1) * y


def asd_grad(eps, latents, t, ac, B, guidance_scale, neg_w=None, weighting="sds"):
    """eps: UNet output in batch order [vd, uncond, (neg 2B), second]; returns (grad, loss, grad_norm)."""
    e_c, e_u = eps[0:B], eps[B:2 * B]
    pos = e_c - e_u
    if neg_w is not None:
        e_n, e_s = eps[2 * B:4 * B], eps[4 * B:5 * B]
        acc = 0
        for i in range(2):
            acc = acc + neg_w[:, i].view(-1, 1, 1, 1) * perpendicular_component(e_n[i::2] - e_u, pos)
        pred = (pos + acc) * guidance_scale + e_u
    else:
        e_s = eps[2 * B:3 * B]
        pred = pos * guidance_scale + e_u
    a = ac[t].view(-1, 1, 1, 1)
    w = {"sds": 1 - a, "uniform": torch.ones_like(a), "fantasia3d": a ** 0.5 * (1 - a)}[weighting]
    grad = torch.nan_to_num(w * (pred - e_s))
    target = (latents - grad).detach()
    loss = 0.5 * F.mse_loss(latents, target, reduction="sum") / B
    return grad, loss, grad.norm()


def seeded_state_dict(kind: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic parameters for `unet_forward` ("unet_sd") / `vae_encoder_forward` ("vae") from the committed
    name -> shape table oracle/ldm_param_specs.json (the reference modules' state-dict names; 3x3 kernels listed as
    [Co,3,3,Ci] and returned as [Co,Ci,3,3]). Lets the CPU baseline run without touching the product library."""
    import json
    import math
    import os
    import zlib

    specs = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ldm_param_specs.json")))[kind]
    sd = {}
    for name, shape in specs:
        rs = (shape[0], shape[3], shape[1], shape[2]) if len(shape) == 4 else tuple(shape)
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
        if len(rs) >= 2:
            t = torch.randn(rs, generator=g) / math.sqrt(math.prod(rs[1:]))
        elif name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(rs, generator=g)
        else:
            t = 0.05 * torch.randn(rs, generator=g)
        sd[name] = t
    return sd

