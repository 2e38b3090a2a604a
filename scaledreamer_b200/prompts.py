"""Prompt processor plugin ("stable-diffusion-prompt-processor"): view-dependent prompt set + device-resident
embedding tables. The CLIP text encoder itself is out of scope (offline, cached on disk in the reference:
threestudio/models/prompt_processors/base.py:19-23, 348-420); embeddings come from
  (a) the reference's own cache files `.threestudio_cache/text_embeddings/<md5>.pt` when they exist, else
  (b) deterministic synthetic N(0,1) embeddings keyed by the same md5 (no weights / tokenizer on this box).
The per-step part -- direction selection and Perp-Neg interpolation (base.py:53-167) -- runs on device through
sdb_asd_text_embeddings with no host sync.
"""
from __future__ import annotations

import ctypes as C_
import hashlib
import json
import os
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from . import core, lib as L
from .core import BaseObject, register

DIRECTIONS = ("side", "front", "back", "overhead")
CACHE_DIR = ".threestudio_cache/text_embeddings"


def hash_prompt(model: str, prompt: str) -> str:
    identifier = f"{model}-{prompt}"
    return hashlib.md5(identifier.encode()).hexdigest()


def shift_azimuth_deg(azimuth: torch.Tensor) -> torch.Tensor:
    return (azimuth + 180) % 360 - 180


class PromptProcessorOutput:
    """What `prompt_processor()` returns (base.py:37-51): the four view-dependent embeddings, their unconditional
    twins, the global pair, and the Perp-Neg coefficients."""

    def __init__(self, text_embeddings, uncond_text_embeddings, text_embeddings_vd, uncond_text_embeddings_vd, cfg,
                 prompt: str, prompts_vd: List[str]):
        self.text_embeddings = text_embeddings              # [1,77,1024] fp16 device
        self.uncond_text_embeddings = uncond_text_embeddings
        self.text_embeddings_vd = text_embeddings_vd        # [4,77,1024]
        self.uncond_text_embeddings_vd = uncond_text_embeddings_vd
        self.cfg = cfg
        self.use_perp_neg = bool(cfg.use_perp_neg)
        self.perp_neg_f_sb, self.perp_neg_f_fsb = tuple(cfg.perp_neg_f_sb), tuple(cfg.perp_neg_f_fsb)
        self.perp_neg_f_fs, self.perp_neg_f_sf = tuple(cfg.perp_neg_f_fs), tuple(cfg.perp_neg_f_sf)
        self.prompt, self.prompts_vd = prompt, prompts_vd

    def prompt_cfg_c(self, view_dependent: bool, perp_neg: bool) -> L.PromptCfgC:
        pc = L.PromptCfgC()
        pc.view_dependent, pc.perp_neg = int(view_dependent), int(perp_neg)
        pc.front_threshold, pc.back_threshold = float(self.cfg.front_threshold), float(self.cfg.back_threshold)
        pc.overhead_threshold = float(self.cfg.overhead_threshold)
        for name in ("f_sb", "f_fsb", "f_fs", "f_sf"):
            arr = getattr(pc, name)
            for i, v in enumerate(getattr(self, "perp_neg_" + name)):
                arr[i] = float(v)
        pc.neg_scale = 1.0
        return pc

    def tables(self, view_dependent: bool):
        if view_dependent:
            return self.text_embeddings_vd, self.uncond_text_embeddings_vd
        return self.text_embeddings, self.uncond_text_embeddings

    def fill_context(self, pc: L.PromptCfgC, elevation, azimuth, ctx, neg_w) -> None:
        """Writes the UNet context batch of the ASD step ([vd, uncond, (neg x2,) vd] blocks) and the Perp-Neg weights
        into the guidance's persistent buffers: one launch, no host sync."""
        emb, unc = self.tables(bool(pc.view_dependent))
        L.check(L.load().sdb_asd_text_embeddings(C_.byref(pc), L.ptr(emb), L.ptr(unc), L.ptr(elevation), L.ptr(azimuth),
                                                 elevation.shape[0], emb.shape[-2], emb.shape[-1], L.ptr(ctx), L.ptr(neg_w),
                                                 L.stream_ptr()), "sdb_asd_text_embeddings")

    def _run(self, elevation, azimuth, view_dependent, perp_neg):
        B = elevation.shape[0]
        dev = self.text_embeddings_vd.device
        pc = self.prompt_cfg_c(view_dependent, perp_neg)
        n_rows = 5 * B if perp_neg else 3 * B
        ctx = torch.empty(n_rows, self.text_embeddings_vd.shape[-2], self.text_embeddings_vd.shape[-1], device=dev,
                          dtype=torch.float16)
        neg_w = torch.zeros(B, 2, device=dev)
        el = elevation.to(dev, torch.float32).contiguous()
        az = azimuth.to(dev, torch.float32).contiguous()
        emb, unc = self.tables(view_dependent)
        L.check(L.load().sdb_asd_text_embeddings(C_.byref(pc), L.ptr(emb), L.ptr(unc), L.ptr(el), L.ptr(az), B,
                                                 emb.shape[-2], emb.shape[-1], L.ptr(ctx), L.ptr(neg_w), L.stream_ptr()),
                "sdb_asd_text_embeddings")
        return ctx, neg_w

    def get_text_embeddings(self, elevation, azimuth, camera_distances, view_dependent_prompting: bool = True):
        """-> [2B,77,1024]: (cond, uncond) (base.py:53-80)."""
        B = elevation.shape[0]
        ctx, _ = self._run(elevation, azimuth, view_dependent_prompting, False)
        return ctx[: 2 * B]

    def get_text_embeddings_perp_neg(self, elevation, azimuth, camera_distances, view_dependent_prompting: bool = True):
        """-> ([4B,77,1024] = pos, uncond, neg(2B)), [B,2] weights (base.py:82-167)."""
        assert view_dependent_prompting, "Perp-Neg only works with view-dependent prompting"
        B = elevation.shape[0]
        ctx, neg_w = self._run(elevation, azimuth, True, True)
        return ctx[: 4 * B], neg_w


@register("stable-diffusion-prompt-processor")
class StableDiffusionPromptProcessor(BaseObject):
    @dataclass
    class Config(BaseObject.Config):
        prompt: str = "a hamburger"
        prompt_front: Optional[str] = None
        prompt_side: Optional[str] = None
        prompt_back: Optional[str] = None
        prompt_overhead: Optional[str] = None
        negative_prompt: str = ""
        pretrained_model_name_or_path: str = "runwayml/stable-diffusion-v1-5"
        overhead_threshold: float = 60.0
        front_threshold: float = 45.0
        back_threshold: float = 45.0
        view_dependent_prompt_front: bool = False
        use_cache: bool = True
        spawn: bool = True
        use_perp_neg: bool = False
        perp_neg_f_sb: Tuple[float, float, float] = (1, 0.5, -0.606)
        perp_neg_f_fsb: Tuple[float, float, float] = (1, 0.5, +0.967)
        perp_neg_f_fs: Tuple[float, float, float] = (4, 0.5, -2.426)
        perp_neg_f_sf: Tuple[float, float, float] = (4, 0.5, -2.426)
        use_prompt_debiasing: bool = False
        pretrained_model_name_or_path_prompt_debiasing: str = "bert-base-uncased"
        prompt_debiasing_mask_ids: Optional[List[int]] = None

    cfg: Config
    embed_dim = 1024
    n_tokens = 77

    def configure(self) -> None:
        if self.cfg.use_prompt_debiasing:
            raise NotImplementedError("prompt debiasing needs a BERT model and is outside the ASD hot path")
        if os.path.exists("load/prompt_library.json"):
            self.prompt_library = json.load(open("load/prompt_library.json"))
        else:
            self.prompt_library = {}
        self.prompt = self.preprocess_prompt(self.cfg.prompt)
        self.negative_prompt = self.cfg.negative_prompt
        if self.cfg.view_dependent_prompt_front:
            fmt = {"side": "side view of {}", "front": "front view of {}", "back": "backside view of {}",
                   "overhead": "overhead view of {}"}
        else:
            fmt = {"side": "{}, side view", "front": "{}, front view", "back": "{}, back view",
                   "overhead": "{}, overhead view"}
        self.prompts_vd = [getattr(self.cfg, f"prompt_{d}") or fmt[d].format(self.prompt) for d in DIRECTIONS]
        self.negative_prompts_vd = [self.negative_prompt for _ in DIRECTIONS]
        self.load_text_embeddings()

    def preprocess_prompt(self, prompt: str) -> str:
        if prompt.startswith("lib:"):  # keyword lookup in load/prompt_library.json (base.py:422-441)
            keywords = prompt[4:].lower().split("_")
            candidate = None
            for p in self.prompt_library.get("dreamfusion", []):
                if all(k in p.lower() for k in keywords):
                    if candidate is not None:
                        raise ValueError(f"Multiple prompts matched with keywords {keywords} in library")
                    candidate = p
            if candidate is None:
                raise ValueError(f"Cannot find prompt with keywords {keywords} in library")
            return candidate
        return prompt

    def _embedding(self, prompt: str) -> torch.Tensor:
        key = hash_prompt(self.cfg.pretrained_model_name_or_path, prompt)
        path = os.path.join(CACHE_DIR, f"{key}.pt")
        if self.cfg.use_cache and os.path.exists(path):
            return torch.load(path, map_location="cpu").reshape(self.n_tokens, self.embed_dim).float()
        core.synthetic_or_raise("text embeddings", path)
        g = torch.Generator().manual_seed(int(key[:8], 16) % (2 ** 31))
        return torch.randn(self.n_tokens, self.embed_dim, generator=g)

    def load_text_embeddings(self) -> None:
        to = lambda ts: torch.stack(ts, 0).to(self.device, torch.float16).contiguous()
        self.text_embeddings = to([self._embedding(self.prompt)])
        self.uncond_text_embeddings = to([self._embedding(self.negative_prompt)])
        self.text_embeddings_vd = to([self._embedding(p) for p in self.prompts_vd])
        self.uncond_text_embeddings_vd = to([self._embedding(p) for p in self.negative_prompts_vd])

    def __call__(self) -> PromptProcessorOutput:
        return PromptProcessorOutput(self.text_embeddings, self.uncond_text_embeddings, self.text_embeddings_vd,
                                     self.uncond_text_embeddings_vd, self.cfg, self.prompt, self.prompts_vd)
