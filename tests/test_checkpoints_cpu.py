"""diffusers -> LDM key renaming (scaledreamer_b200/checkpoints.py) against an independent restatement of the
UNet2DConditionModel / AutoencoderKL module trees of stable-diffusion-2-1-base (the checkpoint the reference's SD
guidance loads, stable_diffusion_asd_guidance.py:68-114). No weights exist on the box: the test builds the diffusers
key -> shape table from the architecture, renames it, and requires the result to be exactly the parameter list (names and
shapes) the native executors enumerate. Tensor identity is checked by tagging every tensor with its own index."""
import ctypes as C

import torch

from scaledreamer_b200 import checkpoints as ck, lib as L, nets


def _specs(kind):
    lib = L.load()
    h = C.c_void_p()
    if kind == "unet":
        c = L.UNetCfgC()
        cfg = nets.SD21_UNET
        for k in ("in_channels", "out_channels", "model_channels", "num_levels", "num_res_blocks", "attn_levels",
                  "head_dim", "context_dim", "context_len", "camera_dim", "num_frames"):
            setattr(c, k, int(cfg[k]))
        for i, m in enumerate(cfg["channel_mult"]):
            c.channel_mult[i] = int(m)
        L.check(lib.sdb_unet_create(C.byref(c), 1, 32, 32, C.byref(h)), "create")
    else:
        c = L.VaeCfgC()
        cfg = nets.SD_VAE
        for k in ("in_channels", "ch", "num_levels", "num_res_blocks", "z_channels"):
            setattr(c, k, int(cfg[k]))
        for i, m in enumerate(cfg["ch_mult"]):
            c.ch_mult[i] = int(m)
        L.check(lib.sdb_vae_encoder_create(C.byref(c), 1, 64, 64, C.byref(h)), "create")
    name, ndim, shape = C.c_char_p(), C.c_int(), (C.c_int * 4)()
    out = {}
    for i in range(lib.sdb_net_num_params(h)):
        L.check(lib.sdb_net_param(h, i, C.byref(name), C.byref(ndim), shape), "param")
        out[name.value.decode()] = nets.reference_shape(tuple(shape[: ndim.value]))
    lib.sdb_net_destroy(h)
    return out


def _wb(d, prefix, w_shape):
    d[prefix + ".weight"] = tuple(w_shape)
    d[prefix + ".bias"] = (w_shape[0],)


def _resnet(d, p, cin, cout, temb=1280):
    _wb(d, p + ".norm1", (cin,))
    _wb(d, p + ".conv1", (cout, cin, 3, 3))
    if temb:
        _wb(d, p + ".time_emb_proj", (cout, temb))
    _wb(d, p + ".norm2", (cout,))
    _wb(d, p + ".conv2", (cout, cout, 3, 3))
    if cin != cout:
        _wb(d, p + ".conv_shortcut", (cout, cin, 1, 1))


def _transformer(d, p, c, ctx=1024):
    _wb(d, p + ".norm", (c,))
    _wb(d, p + ".proj_in", (c, c))  # use_linear_projection=True in SD 2.x
    t = p + ".transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        _wb(d, f"{t}.{n}", (c,))
    for a, kv in (("attn1", c), ("attn2", ctx)):
        d[f"{t}.{a}.to_q.weight"] = (c, c)
        d[f"{t}.{a}.to_k.weight"] = (c, kv)
        d[f"{t}.{a}.to_v.weight"] = (c, kv)
        _wb(d, f"{t}.{a}.to_out.0", (c, c))
    _wb(d, f"{t}.ff.net.0.proj", (8 * c, c))
    _wb(d, f"{t}.ff.net.2", (c, 4 * c))
    _wb(d, p + ".proj_out", (c, c))


def diffusers_unet_table():
    d = {}
    ch = (320, 640, 1280, 1280)
    _wb(d, "conv_in", (320, 4, 3, 3))
    _wb(d, "time_embedding.linear_1", (1280, 320))
    _wb(d, "time_embedding.linear_2", (1280, 1280))
    cin = 320
    for i, c in enumerate(ch):
        for j in range(2):
            _resnet(d, f"down_blocks.{i}.resnets.{j}", cin, c)
            if i < 3:
                _transformer(d, f"down_blocks.{i}.attentions.{j}", c)
            cin = c
        if i < 3:
            _wb(d, f"down_blocks.{i}.downsamplers.0.conv", (c, c, 3, 3))
    _resnet(d, "mid_block.resnets.0", 1280, 1280)
    _transformer(d, "mid_block.attentions.0", 1280)
    _resnet(d, "mid_block.resnets.1", 1280, 1280)
    skips = [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]  # outputs of input_blocks 0..11
    cin = 1280
    for i, c in enumerate(reversed(ch)):
        level = 3 - i
        for j in range(3):
            _resnet(d, f"up_blocks.{i}.resnets.{j}", cin + skips.pop(), c)
            if level < 3:
                _transformer(d, f"up_blocks.{i}.attentions.{j}", c)
            cin = c
        if level > 0:
            _wb(d, f"up_blocks.{i}.upsamplers.0.conv", (c, c, 3, 3))
    _wb(d, "conv_norm_out", (320,))
    _wb(d, "conv_out", (4, 320, 3, 3))
    return d


def diffusers_vae_table(new_attention_names):
    d = {}
    ch = (128, 256, 512, 512)
    _wb(d, "encoder.conv_in", (128, 3, 3, 3))
    cin = 128
    for i, c in enumerate(ch):
        for j in range(2):
            _resnet(d, f"encoder.down_blocks.{i}.resnets.{j}", cin, c, temb=0)
            cin = c
        if i < 3:
            _wb(d, f"encoder.down_blocks.{i}.downsamplers.0.conv", (c, c, 3, 3))
    _resnet(d, "encoder.mid_block.resnets.0", 512, 512, temb=0)
    a = "encoder.mid_block.attentions.0"
    names = ("group_norm", "to_q", "to_k", "to_v", "to_out.0") if new_attention_names else \
        ("group_norm", "query", "key", "value", "proj_attn")
    _wb(d, f"{a}.{names[0]}", (512,))
    for n in names[1:]:
        _wb(d, f"{a}.{n}", (512, 512))
    _resnet(d, "encoder.mid_block.resnets.1", 512, 512, temb=0)
    _wb(d, "encoder.conv_norm_out", (512,))
    _wb(d, "encoder.conv_out", (8, 512, 3, 3))
    _wb(d, "quant_conv", (8, 8, 1, 1))
    _wb(d, "post_quant_conv", (4, 4, 1, 1))
    _wb(d, "decoder.conv_in", (512, 4, 3, 3))  # decoder entries must be ignored
    return d


def _tagged(table):
    """Tiny stand-in tensors: shape is carried separately, the value identifies the source key."""
    keys = sorted(table)
    return {k: torch.full((1,), float(i)) for i, k in enumerate(keys)}, keys


def test_unet_diffusers_names_map_onto_the_executor_parameters():
    table = diffusers_unet_table()
    assert len(table) == 686
    specs = _specs("unet")
    sd, keys = _tagged(table)
    mapped = ck.diffusers_unet_to_ldm(sd)
    assert set(mapped) == set(specs), (sorted(set(mapped) - set(specs))[:5], sorted(set(specs) - set(mapped))[:5])
    for nk, t in mapped.items():
        src = keys[int(t.item())]
        want, got = specs[nk], table[src]
        assert tuple(want) == tuple(got) or (len(want) == 2 and tuple(got) == (*want, 1, 1)), (src, nk, got, want)
    # spot checks of the module order (openaimodel.py:422-808)
    inv = {keys[int(t.item())]: nk for nk, t in mapped.items()}
    assert inv["down_blocks.0.downsamplers.0.conv.weight"] == "input_blocks.3.0.op.weight"
    assert inv["down_blocks.3.resnets.1.conv2.bias"] == "input_blocks.11.0.out_layers.3.bias"
    assert inv["up_blocks.0.upsamplers.0.conv.weight"] == "output_blocks.2.1.conv.weight"   # no attention at 8x8
    assert inv["up_blocks.1.upsamplers.0.conv.weight"] == "output_blocks.5.2.conv.weight"
    assert inv["up_blocks.3.attentions.2.transformer_blocks.0.attn2.to_k.weight"] == \
        "output_blocks.11.1.transformer_blocks.0.attn2.to_k.weight"
    assert inv["mid_block.resnets.1.time_emb_proj.weight"] == "middle_block.2.emb_layers.1.weight"


def test_vae_diffusers_names_map_onto_the_executor_parameters():
    specs = _specs("vae")
    for new_names in (False, True):
        table = diffusers_vae_table(new_names)
        sd = {k: torch.zeros(s) for k, s in table.items()}
        mapped = ck.diffusers_vae_to_ldm(sd)
        enc = {k: v for k, v in mapped.items() if k.startswith("encoder.")}
        assert set(enc) == set(specs), (sorted(set(enc) - set(specs))[:5], sorted(set(specs) - set(enc))[:5])
        for k, v in enc.items():
            want = specs[k]
            assert v.numel() == torch.Size(want).numel() and (v.ndim != 4 or len(want) != 4 or tuple(v.shape) == want), k
        assert set(mapped) - set(enc) == {"quant_conv.weight", "quant_conv.bias"}
        assert mapped["encoder.mid.attn_1.q.weight"].shape == (512, 512, 1, 1)
        assert "encoder.down.1.block.0.nin_shortcut.weight" in mapped and "encoder.down.2.downsample.conv.bias" in mapped


def test_unknown_keys_are_refused():
    import pytest

    with pytest.raises(KeyError):
        ck.diffusers_unet_to_ldm({"down_blocks.0.resnets.0.conv3.weight": torch.zeros(1)})
    with pytest.raises(KeyError):
        ck.diffusers_vae_to_ldm({"encoder.up_blocks.0.x": torch.zeros(1)})


def test_pipeline_directory_round_trip(tmp_path):
    from safetensors.torch import save_file

    (tmp_path / "unet").mkdir()
    (tmp_path / "vae").mkdir()
    assert not ck.is_diffusers_dir(str(tmp_path / "unet"))
    save_file({k: torch.zeros(1) for k in diffusers_unet_table()}, str(tmp_path / "unet" / "diffusion_pytorch_model.safetensors"))
    torch.save({k: torch.zeros(s) for k, s in diffusers_vae_table(False).items()},
               str(tmp_path / "vae" / "diffusion_pytorch_model.bin"))
    assert ck.is_diffusers_dir(str(tmp_path))
    unet, vae = ck.load_diffusers_pipeline(str(tmp_path))
    assert set(unet) == set(_specs("unet")) and "quant_conv.weight" in vae
