"""ASD guidance plugins (threestudio names and Config keys) over the native VAE / UNet executors.

Registered names:
  "stable-diffusion-asynchronous-score-distillation-guidance"  threestudio/models/guidance/stable_diffusion_asd_guidance.py:24
  "mvdream-asynchronous-score-distillation-guidance"           threestudio/models/guidance/mvdream_asd_guidance.py:26
`__call__` keeps the reference signature and returns {"loss_asd", "grad_norm", "min_step", "max_step"}; loss_asd is
a 0-d tensor whose backward runs the hand-written VAE-encoder data-gradient and resize backward into `rgb`.
"""
from __future__ import annotations

import ctypes as C_
import os
from dataclasses import dataclass, field
from typing import Any, Optional

import torch

from . import checkpoints, core, lib as L, nets
from .core import C, BaseObject, register

NUM_TRAIN_TIMESTEPS = 1000
SCALING_FACTOR = 0.18215
WEIGHTING = {"sds": 0, "uniform": 1, "fantasia3d": 2}


def alphas_cumprod(device) -> torch.Tensor:
    """DDPM "scaled_linear" schedule of stable-diffusion-2-1-base == make_beta_schedule("linear", 1000, 0.00085, 0.012)
    (extern/mvdream/ldm/modules/diffusionmodules/util.py:37-40, ldm/interface.py:56-73): float64 then float32."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, NUM_TRAIN_TIMESTEPS, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0).to(torch.float32).to(device)


def normalize_camera(c2w: torch.Tensor) -> torch.Tensor:
    """extern/mvdream/camera_utils.py:45-57: translation rescaled to unit norm, flattened to 16."""
    cam = c2w.reshape(-1, 4, 4).clone()
    tr = cam[:, :3, 3]
    cam[:, :3, 3] = tr / (tr.norm(dim=1, keepdim=True) + 1e-8)
    return cam.reshape(-1, 16)


class _ASDLoss(torch.autograd.Function):
    """loss_asd(rgb): forward = resize -> VAE encoder -> posterior sample + q-sample -> UNet -> score gradient;
    backward = VAE-encoder data gradient -> resize backward (the only differentiable input is rgb)."""

    @staticmethod
    def forward(ctx, rgb, g, prompt_utils, elevation, azimuth, c2w, rng):
        out = g._forward_device(rgb.detach(), prompt_utils, elevation, azimuth, c2w, rng)
        ctx.g = g
        ctx.shape = rgb.shape
        ctx.mark_non_differentiable(out["grad_norm"])
        return out["loss"], out["grad_norm"]

    @staticmethod
    def backward(ctx, g_loss, _g_norm):
        d_rgb = ctx.g._backward_device(ctx.shape)
        return d_rgb * g_loss, None, None, None, None, None, None


class _ASDGuidanceBase(BaseObject):
    """Shared machinery; subclasses fix resolution, UNet flavour and batch layout."""

    image_size = 512
    unet_cfg = nets.SD21_UNET
    per_sample_t = True

    def _init_common(self) -> None:
        self.num_train_timesteps = NUM_TRAIN_TIMESTEPS
        self.alphas = alphas_cumprod(self.device)
        self.grad_clip_val: Optional[float] = None
        self.set_min_max_steps(C(self.cfg.min_step_percent, 0, 0), C(self.cfg.max_step_percent, 0, 0))
        self._nets_for = None
        self.loss_scale = float(os.environ.get("SDB_VAE_LOSS_SCALE", "1"))
        self.weights_seed = int(os.environ.get("SDB_WEIGHTS_SEED", "0"))

    def set_min_max_steps(self, min_step_percent=0.02, max_step_percent=0.98):
        self.min_step = int(self.num_train_timesteps * min_step_percent)
        self.max_step = int(self.num_train_timesteps * max_step_percent)

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False):
        if self.cfg.grad_clip is not None:
            self.grad_clip_val = C(self.cfg.grad_clip, epoch, global_step)
        self.set_min_max_steps(C(self.cfg.min_step_percent, epoch, global_step),
                               C(self.cfg.max_step_percent, epoch, global_step))

    # ---- network set-up (lazy: the batch size is only known at the first call) ----
    def _load_weights(self, vae: nets.VaeEncoder, unet: nets.UNet):
        """Pretrained weights: a diffusers pipeline directory (unet/, vae/; keys renamed by checkpoints.py) or a
        single LDM-layout checkpoint file load by name. Anything else raises FileNotFoundError unless
        SDB_SYNTHETIC_WEIGHTS=1 asks for seeded synthetic parameters (none exist on this box)."""
        path = getattr(self.cfg, "ckpt_path", None) or getattr(self.cfg, "pretrained_model_name_or_path", "")
        if checkpoints.is_diffusers_dir(path):  # the layout StableDiffusionPipeline.from_pretrained reads (:68-114)
            usd, vsd = checkpoints.load_diffusers_pipeline(path)
            unet.load_state_dict(usd)
            vae.load_state_dict({k: v for k, v in vsd.items() if k.startswith("encoder.")})
            self.quant_w = vsd["quant_conv.weight"].reshape(8, 8).float().to(self.device).contiguous()
            self.quant_b = vsd["quant_conv.bias"].float().to(self.device).contiguous()
            return
        if path and os.path.isfile(path):
            sd = torch.load(path, map_location="cpu")
            sd = sd.get("state_dict", sd)
            unet.load_state_dict({k[len("model.diffusion_model."):]: v for k, v in sd.items()
                                  if k.startswith("model.diffusion_model.")})
            vae.load_state_dict({k[len("first_stage_model."):]: v for k, v in sd.items()
                                 if k.startswith("first_stage_model.encoder.")})
            self.quant_w = sd["first_stage_model.quant_conv.weight"].reshape(8, 8).float().to(self.device).contiguous()
            self.quant_b = sd["first_stage_model.quant_conv.bias"].float().to(self.device).contiguous()
            return
        core.synthetic_or_raise("UNet / VAE weights", path)
        unet.load_state_dict(nets.random_state_dict(unet.specs, self.weights_seed))
        vae.load_state_dict(nets.random_state_dict(vae.specs, self.weights_seed + 1))
        q = nets.random_state_dict([("quant_conv.weight", (8, 8)), ("quant_conv.bias", (8,))], self.weights_seed + 2)
        self.quant_w = q["quant_conv.weight"].float().to(self.device).contiguous()
        self.quant_b = q["quant_conv.bias"].float().to(self.device).contiguous()

    def _ensure_nets(self, B: int, src_hw):
        key = (B, tuple(src_hw))
        if self._nets_for == key:
            return
        S, R = self.image_size, self.num_repeats
        dev = self.device
        self.vae = nets.VaeEncoder(nets.SD_VAE, B, S, S, dev)
        self.unet = nets.UNet(self.unet_cfg, (R + 1) * B, S // 8, S // 8, dev)
        self._load_weights(self.vae, self.unet)
        hw = (S // 8) ** 2
        f32 = dict(device=dev, dtype=torch.float32)
        self.buf = dict(
            img=torch.empty(B, S, S, 3, **f32), h=torch.empty(B, S // 8, S // 8, 8, **f32),
            latents=torch.empty(B, hw, 4, **f32), unet_x=torch.empty((R + 1) * B, S // 8, S // 8, 4, device=dev,
                                                                      dtype=torch.float16),
            unet_t=torch.empty((R + 1) * B, **f32), eps=torch.empty((R + 1) * B, S // 8, S // 8, 4, **f32),
            ctx=torch.empty((R + 1) * B, 77, 1024, device=dev, dtype=torch.float16),
            neg_w=torch.empty(B, 2, **f32), grad=torch.empty(B, hw, 4, **f32), d_h=torch.empty(B, hw, 8, **f32),
            loss=torch.zeros(1, **f32), grad_norm=torch.zeros(1, **f32), t_plus=torch.empty(B, device=dev,
                                                                                          dtype=torch.int32),
            d_img=torch.empty(B, S, S, 3, **f32), d_rgb=torch.empty(B, src_hw[0], src_hw[1], 3, **f32))
        self._nets_for = key

    # ---- device pipeline ----
    def _prompt_cfg(self, prompt_utils) -> L.PromptCfgC:
        pc = prompt_utils.prompt_cfg_c(view_dependent=bool(self.cfg.view_dependent_prompting),
                                       perp_neg=self.use_perp_neg)
        pc.neg_scale = -1.0 * float(getattr(self.cfg, "guidance_perp_neg", 0.0))
        return pc

    def _camera_cond(self, c2w):
        return None

    def _forward_device(self, rgb, prompt_utils, elevation, azimuth, c2w, rng):
        lib = L.load()
        B, H, W, _ = rgb.shape
        self._ensure_nets(B, (H, W))
        S, R, b, st = self.image_size, self.num_repeats, self.buf, L.stream_ptr()
        hw = (S // 8) ** 2
        dev = self.device
        rgb = rgb.contiguous().float()
        # (1) resize to the VAE resolution, x*2-1
        L.check(lib.sdb_resize_bilinear_forward(L.ptr(rgb), B, H, W, 3, L.ptr(b["img"]), S, S, 2.0, -1.0, st), "resize")
        # (2) frozen VAE encoder
        self.vae.forward(b["img"], out=b["h"])
        # (3) random draws (explicit inputs for parity tests; device RNG otherwise)
        rng = rng or {}
        noise = rng.get("noise")
        if noise is None:
            noise = torch.randn(B, hw, 4, device=dev)
        eps_post = rng.get("eps_post")
        if eps_post is None:
            eps_post = torch.randn(B, hw, 4, device=dev)
        t = rng.get("t")
        if t is None:
            n_t = B if self.per_sample_t else 1
            t = torch.randint(self.min_step, self.max_step + 1, [n_t], device=dev, dtype=torch.int32)
            if not self.per_sample_t:
                t = t.repeat(B)
        t = t.to(torch.int32).contiguous()
        u = rng.get("u")
        if u is None and self.cfg.plus_random:
            u = torch.rand(B if self.per_sample_t else 1, device=dev)
            if not self.per_sample_t:
                u = u.repeat(B)
        L.check(lib.sdb_asd_t_plus(L.ptr(t), L.ptr(u) if self.cfg.plus_random else None, B, float(self.cfg.plus_ratio),
                                   int(self.min_step), self.num_train_timesteps, L.ptr(b["t_plus"]), st), "t_plus")
        # (4) posterior sample, q-sample at t and t+dt, UNet input batch
        L.check(lib.sdb_asd_prologue(L.ptr(b["h"]), L.ptr(self.quant_w), L.ptr(self.quant_b), L.ptr(eps_post),
                                     L.ptr(noise), L.ptr(t), L.ptr(b["t_plus"]), L.ptr(self.alphas), SCALING_FACTOR, B,
                                     hw, R, L.ptr(b["latents"]), L.ptr(b["unet_x"]), L.ptr(b["unet_t"]), st), "prologue")
        # (5) text context (view-dependent / Perp-Neg) built on device
        pc = self._prompt_cfg(prompt_utils)
        elevation = elevation.to(dev, torch.float32).contiguous()
        azimuth = azimuth.to(dev, torch.float32).contiguous()
        prompt_utils.fill_context(pc, elevation, azimuth, b["ctx"], b["neg_w"])
        # (6) frozen UNet on the concatenated batch
        cam = self._camera_cond(c2w)
        self.unet.forward(b["unet_x"], b["unet_t"], b["ctx"], cam, out=b["eps"])
        # (7) CFG / Perp-Neg / w(t) -> grad, loss, dLoss/d(VAE moments)
        clip = float(self.grad_clip_val) if self.grad_clip_val is not None else 0.0
        L.check(lib.sdb_asd_epilogue(L.ptr(b["eps"]), L.ptr(b["h"]), L.ptr(self.quant_w), L.ptr(self.quant_b),
                                     L.ptr(eps_post), L.ptr(t), L.ptr(self.alphas),
                                     L.ptr(b["neg_w"]) if self.use_perp_neg else None, float(self.cfg.guidance_scale),
                                     WEIGHTING[self.cfg.weighting_strategy], clip, SCALING_FACTOR, self.loss_scale, B, hw,
                                     R, L.ptr(b["grad"]), L.ptr(b["d_h"]), L.ptr(b["loss"]), L.ptr(b["grad_norm"]), st),
                "epilogue")
        self._last = dict(t=t, noise=noise, eps_post=eps_post)
        return {"loss": b["loss"][0].clone(), "grad_norm": b["grad_norm"][0].clone()}

    def _backward_device(self, shape):
        lib = L.load()
        B, H, W, _ = shape
        S, b, st = self.image_size, self.buf, L.stream_ptr()
        self.vae.backward(b["d_h"].view(B, S // 8, S // 8, 8), out=b["d_img"])
        L.check(lib.sdb_resize_bilinear_backward(L.ptr(b["d_img"]), B, H, W, 3, L.ptr(b["d_rgb"]), S, S,
                                                 2.0 / self.loss_scale, st), "resize_backward")
        return b["d_rgb"]

    def _call(self, rgb, prompt_utils, elevation, azimuth, camera_distances, c2w=None, rgb_as_latents=False, rng=None):
        if rgb_as_latents:
            raise NotImplementedError("rgb_as_latents=True is not implemented by the sm_100a guidance path")
        if not (self.min_step is not None and self.max_step is not None):
            raise RuntimeError("min/max step not set")
        loss, grad_norm = _ASDLoss.apply(rgb, self, prompt_utils, elevation, azimuth, c2w, rng)
        return {"loss_asd": loss, "grad_norm": grad_norm, "min_step": self.min_step, "max_step": self.max_step}


@register("stable-diffusion-asynchronous-score-distillation-guidance")
class SDTimestepShiftedScoreDistillationGuidance(_ASDGuidanceBase):
    @dataclass
    class Config(BaseObject.Config):
        pretrained_model_name_or_path: str = "stabilityai/stable-diffusion-2-1-base"
        enable_memory_efficient_attention: bool = False
        enable_sequential_cpu_offload: bool = False
        enable_attention_slicing: bool = False
        enable_channels_last_format: bool = True
        guidance_scale: float = 7.5
        grad_clip: Optional[Any] = None
        half_precision_weights: bool = True
        min_step_percent: Any = 0.02
        max_step_percent: Any = 0.98
        weighting_strategy: str = "sds"
        plus_ratio: float = 0.1
        plus_random: bool = False
        view_dependent_prompting: bool = True
        guidance_perp_neg: float = 0.0

    cfg: Config
    image_size = 512
    unet_cfg = nets.SD21_UNET
    per_sample_t = True

    def configure(self) -> None:
        if self.cfg.weighting_strategy not in WEIGHTING:
            raise ValueError(f"Unknown weighting strategy: {self.cfg.weighting_strategy}")
        self.use_perp_neg = self.cfg.guidance_perp_neg != 0
        self.num_repeats = 4 if self.use_perp_neg else 2
        self._init_common()

    def __call__(self, rgb, prompt_utils, elevation, azimuth, camera_distances, rgb_as_latents=False,
                 guidance_eval=False, **kwargs):
        if self.use_perp_neg and not prompt_utils.use_perp_neg:
            raise AssertionError("guidance_perp_neg != 0 needs a prompt processor with use_perp_neg: true")
        return self._call(rgb, prompt_utils, elevation, azimuth, camera_distances, None, rgb_as_latents,
                          kwargs.get("_rng"))


@register("mvdream-asynchronous-score-distillation-guidance")
class MVDreamTimestepShiftedScoreDistillationGuidance(_ASDGuidanceBase):
    @dataclass
    class Config(BaseObject.Config):
        model_name: str = "sd-v2.1-base-4view"
        ckpt_path: Optional[str] = None
        grad_clip: Optional[Any] = None
        half_precision_weights: bool = True
        guidance_scale: float = 7.5
        n_view: int = 4
        min_step_percent: Any = 0.02
        max_step_percent: Any = 0.98
        weighting_strategy: str = "sds"
        plus_ratio: float = 0.1
        plus_random: bool = False
        camera_condition_type: str = "rotation"
        view_dependent_prompting: bool = False

    cfg: Config
    image_size = 256
    per_sample_t = False  # one timestep shared by the whole view batch (mvdream_asd_guidance.py:214-221)

    def configure(self) -> None:
        if self.cfg.weighting_strategy not in WEIGHTING:
            raise ValueError(f"Unknown weighting strategy: {self.cfg.weighting_strategy}")
        if self.cfg.camera_condition_type != "rotation":
            raise NotImplementedError(f"Unknown camera_condition_type={self.cfg.camera_condition_type}")
        self.use_perp_neg = False
        self.num_repeats = 2
        self.unet_cfg = dict(nets.MVDREAM_UNET, num_frames=int(self.cfg.n_view))
        self._init_common()

    def _camera_cond(self, c2w):
        if c2w is None:
            raise ValueError("the multi-view guidance needs c2w")
        cam = normalize_camera(c2w.to(self.device, torch.float32))
        return cam.repeat(self.num_repeats + 1, 1)

    def __call__(self, rgb, prompt_utils, elevation, azimuth, camera_distances, c2w, rgb_as_latents=False, fovy=None,
                 input_is_latent=False, **kwargs):
        return self._call(rgb, prompt_utils, elevation, azimuth, camera_distances, c2w, rgb_as_latents,
                          kwargs.get("_rng"))
