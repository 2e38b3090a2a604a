"""How far is the REFERENCE's own half-precision UNet from its fp32 output?  (evidence for the eps-pred tolerance)

north_star states "eps-pred within 1e-3 relative of the reference diffusers path". That path is fp16
(stable_diffusion_asd_guidance.py:38,57-59: half_precision_weights -> torch_dtype=float16 for VAE and UNet), so the only
fp32-exact statement available is relative to an fp32 run of the same network. This script runs the reference's vendored
LDM UNet (the architecture twin of UNet2DConditionModel, see make_ldm_golden.py) TWICE on the golden inputs: in fp32
(= ldm_golden.pt's `y`) and with `.half()` parameters and activations the way diffusers runs it (GroupNorm statistics in
fp32 via GroupNorm32, attention logits in fp32 via _ATTN_PRECISION, everything else fp16), and records
rel_l2(y_fp16, y_fp32) next to the golden. tests/test_nets_gpu.py then asserts err(this repo) <= max(1e-3, that).

Run in the build container only:  python tests/golden/make_ldm_fp16_evidence.py   -> tests/golden/ldm_fp16_evidence.json
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ldm_golden as G  # noqa: E402  (installs the omegaconf stub, imports the reference modules)

torch.set_grad_enabled(False)


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm())


def _patch_attention_einsum():
    """CrossAttention.forward (attention.py:163-194) keeps the fp32 logits / softmax of _ATTN_PRECISION == "fp32" and then
    contracts them with the half-precision values, which only type-checks under autocast. Outside autocast the
    probabilities are cast to the values' dtype first -- what torch SDPA / diffusers' AttnProcessor2_0 do (fp32 softmax,
    fp16 P V)."""
    import extern.mvdream.ldm.modules.attention as A

    plain = torch.einsum

    def einsum(eq, *ops):
        if len({o.dtype for o in ops}) > 1:
            ops = tuple(o.to(ops[-1].dtype) for o in ops)
        return plain(eq, *ops)

    A.einsum = einsum


def unet_fp16(multiview, case):
    kw = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
              num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
              use_linear_in_transformer=True, transformer_depth=1, context_dim=1024, use_checkpoint=False, legacy=False)
    net = (G.MultiViewUNetModel(camera_dim=16, **kw) if multiview else G.UNetModel(**kw)).eval()
    G.load_seeded(net, case["seed"])
    y32 = net(case["x"].float(), case["t"], context=case["ctx"].float(),
              **({"camera": case["camera"].float(), "num_frames": 4} if multiview else {}))
    assert rel(y32, case["y"]) < 1e-6, "the fp32 run must reproduce the committed golden"
    _patch_attention_einsum()
    net = net.half()
    net.dtype = torch.float16
    for m in net.modules():  # GroupNorm32 normalises in fp32 (util.py:229-231): its affine parameters stay fp32
        if isinstance(m, torch.nn.GroupNorm):
            m.float()
    cast = lambda mod, args: tuple(a.half() if torch.is_tensor(a) and a.is_floating_point() else a for a in args)
    net.time_embed.register_forward_pre_hook(cast)  # diffusers: t_emb.to(dtype=sample.dtype)
    if multiview:
        net.camera_embed.register_forward_pre_hook(cast)
    y16 = net(case["x"].half(), case["t"], context=case["ctx"].half(),
              **({"camera": case["camera"].half(), "num_frames": 4} if multiview else {}))
    return rel(y16.float(), y32)


if __name__ == "__main__":
    gold = torch.load(os.path.join(HERE, "ldm_golden.pt"))
    out = {}
    for name, mv in (("unet_sd", False), ("unet_mv", True)):
        out[name] = {"reference_fp16_vs_fp32_rel_l2": unet_fp16(mv, gold[name]),
                     "how": "reference LDM module .half() on CPU (torch %s), inputs of ldm_golden.pt" % torch.__version__}
        print(name, out[name])
    json.dump(out, open(os.path.join(HERE, "ldm_fp16_evidence.json"), "w"), indent=1)
