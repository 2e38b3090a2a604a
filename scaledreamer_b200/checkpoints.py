"""Checkpoint key layouts for the frozen guidance networks.

The native executors (csrc/models.cu) enumerate their parameters under the vendored-LDM names
(`extern/mvdream/ldm/modules/diffusionmodules/openaimodel.py`, `model.py`: `input_blocks.N.M...`, `encoder.down.i.block.j...`),
which is what MVDream's `sd-v2.1-base-4view.pt` uses. The SD guidance of the reference loads a **diffusers** pipeline
directory instead (`StableDiffusionPipeline.from_pretrained(pretrained_model_name_or_path)`,
threestudio/models/guidance/stable_diffusion_asd_guidance.py:68-114): `unet/diffusion_pytorch_model.safetensors` with
`UNet2DConditionModel` names and `vae/diffusion_pytorch_model.safetensors` with `AutoencoderKL` names. This module renames
those to the LDM layout (the inverse of diffusers' `convert_ldm_unet_checkpoint` / `convert_ldm_vae_checkpoint`), so
the same plugin configs load the same files. Tensors are not touched except for the VAE attention projections, which
are `Linear [C, C]` in diffusers and `Conv2d [C, C, 1, 1]` in LDM (the executors take either: same element order).
"""
from __future__ import annotations

import os
import re
from typing import Dict, Optional, Sequence, Tuple

import torch

_RESNET = {"norm1": "in_layers.0", "conv1": "in_layers.2", "time_emb_proj": "emb_layers.1", "norm2": "out_layers.0",
           "conv2": "out_layers.3", "conv_shortcut": "skip_connection"}


def _resnet(rest: str) -> str:
    head, _, tail = rest.partition(".")
    if head not in _RESNET:
        raise KeyError(f"unknown diffusers ResnetBlock2D parameter '{rest}'")
    return f"{_RESNET[head]}.{tail}"


def diffusers_unet_to_ldm(sd: Dict[str, torch.Tensor], layers_per_block: int = 2,
                          attention_levels: Sequence[bool] = (True, True, True, False)) -> Dict[str, torch.Tensor]:
    """UNet2DConditionModel state dict -> UNetModel names (openaimodel.py:422-808 module order).

    input_blocks: 0 = conv_in; level i holds `layers_per_block` (ResBlock[, SpatialTransformer]) entries followed by one
    Downsample entry (except the last level). output_blocks mirror it with `layers_per_block + 1` entries per level,
    the last of which carries the Upsample (index 2 after an attention, 1 otherwise)."""
    n_levels, lpb = len(attention_levels), layers_per_block
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        m = re.match(r"(down_blocks|up_blocks)\.(\d+)\.(resnets|attentions|downsamplers|upsamplers)\.(\d+)\.(.*)", k)
        if k.startswith("time_embedding.linear_1."):
            nk = "time_embed.0." + k.rsplit(".", 1)[1]
        elif k.startswith("time_embedding.linear_2."):
            nk = "time_embed.2." + k.rsplit(".", 1)[1]
        elif k.startswith("conv_in."):
            nk = "input_blocks.0.0." + k.rsplit(".", 1)[1]
        elif k.startswith("conv_norm_out."):
            nk = "out.0." + k.rsplit(".", 1)[1]
        elif k.startswith("conv_out."):
            nk = "out.2." + k.rsplit(".", 1)[1]
        elif k.startswith("mid_block.resnets."):
            j, rest = k[len("mid_block.resnets."):].split(".", 1)
            nk = f"middle_block.{2 * int(j)}.{_resnet(rest)}"
        elif k.startswith("mid_block.attentions.0."):
            nk = "middle_block.1." + k[len("mid_block.attentions.0."):]
        elif m:
            side, i, kind, j, rest = m.group(1), int(m.group(2)), m.group(3), int(m.group(4)), m.group(5)
            if side == "down_blocks":
                if kind == "downsamplers":
                    nk = f"input_blocks.{(i + 1) * (lpb + 1)}.0.op.{rest[len('conv.'):]}"
                else:
                    blk = 1 + i * (lpb + 1) + j
                    nk = f"input_blocks.{blk}.0.{_resnet(rest)}" if kind == "resnets" else f"input_blocks.{blk}.1.{rest}"
            else:
                if kind == "upsamplers":
                    # up level i of diffusers is LDM level n_levels-1-i; attention there decides the slot of Upsample
                    slot = 2 if attention_levels[n_levels - 1 - i] else 1
                    nk = f"output_blocks.{i * (lpb + 1) + lpb}.{slot}.conv.{rest[len('conv.'):]}"
                else:
                    blk = i * (lpb + 1) + j
                    nk = f"output_blocks.{blk}.0.{_resnet(rest)}" if kind == "resnets" else f"output_blocks.{blk}.1.{rest}"
        else:
            raise KeyError(f"unknown diffusers UNet2DConditionModel parameter '{k}'")
        if nk in out:
            raise KeyError(f"two diffusers parameters map to '{nk}'")
        out[nk] = v
    return out


_VAE_ATTN = {"group_norm": "norm", "query": "q", "key": "k", "value": "v", "proj_attn": "proj_out",  # diffusers < 0.20
             "to_q": "q", "to_k": "k", "to_v": "v", "to_out.0": "proj_out"}                          # diffusers >= 0.20


def diffusers_vae_to_ldm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """AutoencoderKL state dict -> first_stage_model names (ldm model.py:452-543, autoencoder.py:32); decoder and
    post_quant_conv entries (not on the ASD path) are dropped."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.startswith(("decoder.", "post_quant_conv.")):
            continue
        if k.startswith("quant_conv.") or k.startswith(("encoder.conv_in.", "encoder.conv_out.")):
            nk = k
        elif k.startswith("encoder.conv_norm_out."):
            nk = "encoder.norm_out." + k.rsplit(".", 1)[1]
        elif k.startswith("encoder.mid_block.resnets."):
            j, rest = k[len("encoder.mid_block.resnets."):].split(".", 1)
            nk = f"encoder.mid.block_{int(j) + 1}.{rest}"
        elif k.startswith("encoder.mid_block.attentions.0."):
            rest = k[len("encoder.mid_block.attentions.0."):]
            name, leaf = rest.rsplit(".", 1)
            if name not in _VAE_ATTN:
                raise KeyError(f"unknown diffusers VAE attention parameter '{k}'")
            nk = f"encoder.mid.attn_1.{_VAE_ATTN[name]}.{leaf}"
            if leaf == "weight" and v.ndim == 2 and _VAE_ATTN[name] != "norm":
                v = v[:, :, None, None]
        else:
            m = re.match(r"encoder\.down_blocks\.(\d+)\.(resnets|downsamplers)\.(\d+)\.(.*)", k)
            if not m:
                raise KeyError(f"unknown diffusers AutoencoderKL parameter '{k}'")
            i, kind, j, rest = int(m.group(1)), m.group(2), int(m.group(3)), m.group(4)
            if kind == "downsamplers":
                nk = f"encoder.down.{i}.downsample.{rest}"
            else:
                nk = f"encoder.down.{i}.block.{j}.{rest.replace('conv_shortcut', 'nin_shortcut')}"
        out[nk] = v
    return out


def _read(path_stem: str) -> Optional[Dict[str, torch.Tensor]]:
    for ext in (".safetensors", ".fp16.safetensors", ".bin"):
        p = path_stem + ext
        if os.path.isfile(p):
            if p.endswith(".safetensors"):
                from safetensors.torch import load_file

                return load_file(p)
            return torch.load(p, map_location="cpu")
    return None


def is_diffusers_dir(path: str) -> bool:
    return bool(path) and os.path.isdir(os.path.join(path, "unet")) and os.path.isdir(os.path.join(path, "vae"))


def load_diffusers_pipeline(path: str) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """(unet, first_stage) state dicts in LDM names from a diffusers pipeline directory (unet/, vae/ sub-folders)."""
    unet = _read(os.path.join(path, "unet", "diffusion_pytorch_model"))
    vae = _read(os.path.join(path, "vae", "diffusion_pytorch_model"))
    if unet is None or vae is None:
        raise FileNotFoundError(f"{path}: unet/ or vae/ diffusion_pytorch_model.{{safetensors,bin}} not found")
    return diffusers_unet_to_ldm(unet), diffusers_vae_to_ldm(vae)
