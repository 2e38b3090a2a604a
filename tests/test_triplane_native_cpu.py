"""Host side of the native Triplane-Transformer (no GPU): the flat parameter list the autograd node works on covers every
parameter of the state-dict-compatible module exactly once, in the documented order, and the product path refuses to
run without CUDA instead of falling back to torch."""
import os

import pytest
import torch

from scaledreamer_b200.amortized import TriplaneTransformer
from scaledreamer_b200.triplane_native import PER_LAYER, flat_parameters


@pytest.mark.parametrize("layers,dim,heads", [(1, 64, 4), (3, 128, 2)])
def test_flat_parameter_list_covers_the_module(layers, dim, heads):
    gen = TriplaneTransformer(inner_dim=dim, condition_dim=1024, triplane_low_res=8, triplane_high_res=16, triplane_dim=32,
                              num_layers=layers, num_heads=heads, local_text=True)
    ps = flat_parameters(gen)
    own = list(gen.parameters())
    assert len(ps) == len(own) == 1 + PER_LAYER * layers + 3
    assert {id(p) for p in ps} == {id(p) for p in own}
    names = {id(p): n for n, p in gen.named_parameters()}
    order = [names[id(p)] for p in ps]
    assert order[0] == "pos_embed" and order[-3:] == ["norm.weight", "norm.bias", "deconv.weight"]
    blk = [n.split(".", 2)[2] for n in order[1:1 + PER_LAYER]]
    assert blk == ["norm1.weight", "norm1.bias", "cross_attn.to_q.weight", "cross_attn.to_k.weight", "cross_attn.to_v.weight",
                   "cross_attn.to_out.0.weight", "cross_attn.to_out.0.bias", "norm2.weight", "norm2.bias",
                   "self_attn.to_q.weight", "self_attn.to_k.weight", "self_attn.to_v.weight", "self_attn.to_out.0.weight",
                   "self_attn.to_out.0.bias", "norm3.weight", "norm3.bias", "mlp.0.weight", "mlp.0.bias", "mlp.3.weight",
                   "mlp.3.bias"]


def test_no_cpu_path():
    gen = TriplaneTransformer(inner_dim=64, condition_dim=1024, triplane_low_res=8, triplane_high_res=16, triplane_dim=32,
                              num_layers=1, num_heads=4, local_text=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        gen(torch.randn(1, 77, 1024))
    # the plain-torch restatement (tests' comparison, and the local_text = False variant) does run anywhere
    assert gen.forward_torch(torch.randn(1, 77, 1024)).shape == (1, 3, 32, 16, 16)


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "triplane_generator_golden.pt")


@pytest.mark.parametrize("tag", ["local_text", "global_text"])
def test_forward_torch_matches_the_reference_module(tag):
    """`TriplaneTransformer.forward_torch` (the comparison the GPU parity tests use, and the path of the local_text = False
    variant) against planes computed by the REFERENCE's own triplane_transformer_modules.py, executed unchanged by
    tests/golden/make_triplane_generator_golden.py (diffusers' Attention restated there: it is not installed). The strict
    state-dict load also pins every parameter name and shape to the reference's."""
    g = torch.load(GOLD)[tag]
    gen = TriplaneTransformer(**g["cfg"])
    gen.load_state_dict(g["state_dict"], strict=True)
    with torch.no_grad():
        out = gen.forward_torch(g["text_embed"])
    assert out.shape == g["planes"].shape
    err = float((out - g["planes"]).norm() / g["planes"].norm())
    assert err < 1e-5, err
