"""Host-side handles of the native UNet / VAE-encoder executors (include/sdb200_nn.h, csrc/models.cu).

The arithmetic and the layer graph live in the C++ library; this module owns the two torch-allocated arenas,
feeds parameters from a reference-layout state dict (the vendored LDM / diffusers key names and NCHW conv
weights) and exposes forward / backward on torch tensors.
"""
from __future__ import annotations

import ctypes as C
import math
import zlib
from typing import Dict, List, Optional, Tuple

import torch

from . import lib as L

SD21_UNET = dict(in_channels=4, out_channels=4, model_channels=320, num_levels=4, channel_mult=(1, 2, 4, 4),
                 num_res_blocks=2, attn_levels=3, head_dim=64, context_dim=1024, context_len=77, camera_dim=0,
                 num_frames=1)  # extern/mvdream/configs/sd-v2-base.yaml:10-27 minus camera_dim
MVDREAM_UNET = dict(SD21_UNET, camera_dim=16, num_frames=4)
SD_VAE = dict(in_channels=3, ch=128, num_levels=4, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4)


class _Net:
    def __init__(self, handle, device):
        self._h = handle
        self.device = device
        lib = L.load()
        wb, kb = C.c_longlong(), C.c_longlong()
        L.check(lib.sdb_net_sizes(handle, C.byref(wb), C.byref(kb)), "sdb_net_sizes")
        self.weight_bytes, self.work_bytes = wb.value, kb.value
        self._weights = torch.empty(wb.value, dtype=torch.uint8, device=device)
        self._work = torch.zeros(kb.value, dtype=torch.uint8, device=device)
        L.check(lib.sdb_net_bind(handle, L.ptr(self._weights), L.ptr(self._work)), "sdb_net_bind")
        self.specs: List[Tuple[str, Tuple[int, ...]]] = []
        name, ndim, shape = C.c_char_p(), C.c_int(), (C.c_int * 4)()
        for i in range(lib.sdb_net_num_params(handle)):
            L.check(lib.sdb_net_param(handle, i, C.byref(name), C.byref(ndim), shape), "sdb_net_param")
            self.specs.append((name.value.decode(), tuple(shape[: ndim.value])))
        self._finalized = False

    def __del__(self):
        try:
            if self._h:
                L.load().sdb_net_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def num_parameters(self) -> int:
        return sum(math.prod(s) for _, s in self.specs)

    def launches(self, backward=False) -> int:
        return int(L.load().sdb_net_num_launches(self._h, int(backward)))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """sd: reference-layout tensors (conv weights [Cout,Cin,kh,kw]); every parameter must be present."""
        lib = L.load()
        for name, shape in self.specs:
            if name not in sd:
                raise KeyError(f"state dict lacks '{name}'")
            t = sd[name]
            if len(shape) == 4:
                if tuple(t.shape) != (shape[0], shape[3], shape[1], shape[2]):
                    raise RuntimeError(f"{name}: expected {(shape[0], shape[3], shape[1], shape[2])}, got {tuple(t.shape)}")
                t = t.permute(0, 2, 3, 1)
            elif t.numel() != math.prod(shape):
                raise RuntimeError(f"{name}: expected {shape}, got {tuple(t.shape)}")
            t = t.to(device=self.device, dtype=torch.float16).contiguous()
            L.check(lib.sdb_net_load_param(self._h, name.encode(), L.ptr(t), t.numel(), L.stream_ptr()),
                    f"sdb_net_load_param({name})")
            del t
        L.check(lib.sdb_net_finalize(self._h, L.stream_ptr()), "sdb_net_finalize")
        torch.cuda.synchronize()
        self._finalized = True

    def _ready(self):
        if not self._finalized:
            raise RuntimeError("load_state_dict() must be called before running the network")


def reference_shape(shape: Tuple[int, ...]) -> Tuple[int, ...]:
    return (shape[0], shape[3], shape[1], shape[2]) if len(shape) == 4 else shape


def random_state_dict(specs, seed: int = 0, device="cpu") -> Dict[str, torch.Tensor]:
    """Seeded synthetic parameters in the reference layout (there are no pretrained weights on the box).
    Every tensor is drawn from its own generator keyed by (seed, name), so the result does not depend on
    enumeration order; the reference's zero-initialised layers are random like the rest."""
    sd = {}
    for name, shape in specs:
        rs = reference_shape(tuple(shape))
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
        if len(rs) >= 2:
            fan_in = math.prod(rs[1:])
            t = torch.randn(rs, generator=g) / math.sqrt(fan_in)
        elif name.endswith(".weight"):  # norm scale
            t = 1.0 + 0.1 * torch.randn(rs, generator=g)
        else:
            t = 0.05 * torch.randn(rs, generator=g)
        sd[name] = t.to(device)
    return sd


class UNet(_Net):
    """UNetModel / MultiViewUNetModel forward (openaimodel.py:777-808, 1175-1213), frozen, no grad."""

    def __init__(self, cfg: dict, batch: int, height: int, width: int, device):
        lib = L.load()
        c = L.UNetCfgC()
        for k in ("in_channels", "out_channels", "model_channels", "num_levels", "num_res_blocks", "attn_levels",
                  "head_dim", "context_dim", "context_len", "camera_dim", "num_frames"):
            setattr(c, k, int(cfg[k]))
        for i, m in enumerate(cfg["channel_mult"]):
            c.channel_mult[i] = int(m)
        h = C.c_void_p()
        L.check(lib.sdb_unet_create(C.byref(c), batch, height, width, C.byref(h)), "sdb_unet_create")
        self.cfg, self.batch, self.height, self.width = dict(cfg), batch, height, width
        super().__init__(h, device)

    def forward(self, x: torch.Tensor, t: torch.Tensor, ctx: torch.Tensor, camera: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x fp16 [B,H,W,Cin] (channels-last); t fp32 [B]; ctx fp16 [B,77,1024]; -> fp32 [B,H,W,Cout]."""
        self._ready()
        B, H, W = self.batch, self.height, self.width
        assert x.shape == (B, H, W, self.cfg["in_channels"]) and x.dtype == torch.float16
        assert ctx.shape == (B, self.cfg["context_len"], self.cfg["context_dim"]) and ctx.dtype == torch.float16
        t = t.to(torch.float32)
        if out is None:
            out = torch.empty(B, H, W, self.cfg["out_channels"], device=x.device, dtype=torch.float32)
        cam = None
        if self.cfg["camera_dim"]:
            cam = camera.to(torch.float16).contiguous()
        L.check(L.load().sdb_unet_forward(self._h, L.ptr(x), L.ptr(t), L.ptr(ctx), L.ptr(cam), L.ptr(out),
                                          L.stream_ptr()), "sdb_unet_forward")
        return out


class VaeEncoder(_Net):
    """AutoencoderKL encoder (model.py:452-543) up to conv_out, with the data-gradient backward."""

    def __init__(self, cfg: dict, batch: int, height: int, width: int, device):
        lib = L.load()
        c = L.VaeCfgC()
        for k in ("in_channels", "ch", "num_levels", "num_res_blocks", "z_channels"):
            setattr(c, k, int(cfg[k]))
        for i, m in enumerate(cfg["ch_mult"]):
            c.ch_mult[i] = int(m)
        h = C.c_void_p()
        L.check(lib.sdb_vae_encoder_create(C.byref(c), batch, height, width, C.byref(h)), "sdb_vae_encoder_create")
        self.cfg, self.batch, self.height, self.width = dict(cfg), batch, height, width
        super().__init__(h, device)

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x fp32 [B,H,W,3] in [-1,1] -> fp32 [B,H/8,W/8,8] (conv_out output, before quant_conv)."""
        self._ready()
        B, H, W = self.batch, self.height, self.width
        assert x.shape == (B, H, W, self.cfg["in_channels"]) and x.dtype == torch.float32 and x.is_contiguous()
        if out is None:
            out = torch.empty(B, H // 8, W // 8, 2 * self.cfg["z_channels"], device=x.device, dtype=torch.float32)
        L.check(L.load().sdb_vae_encoder_forward(self._h, L.ptr(x), L.ptr(out), L.stream_ptr()),
                "sdb_vae_encoder_forward")
        return out

    def backward(self, d_h: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._ready()
        B, H, W = self.batch, self.height, self.width
        assert d_h.shape == (B, H // 8, W // 8, 2 * self.cfg["z_channels"]) and d_h.dtype == torch.float32
        d_h = d_h.contiguous()
        if out is None:
            out = torch.empty(B, H, W, self.cfg["in_channels"], device=d_h.device, dtype=torch.float32)
        L.check(L.load().sdb_vae_encoder_backward(self._h, L.ptr(d_h), L.ptr(out), L.stream_ptr()),
                "sdb_vae_encoder_backward")
        return out
