"""Golden vectors for the Triplane-Transformer generator, produced by the REFERENCE's own module.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_triplane_generator_golden.py
The whole file custom/amortized/extern/triplane_transformer_modules.py is executed unchanged (ConditionModulationBlock,
ConditionModulationBlockwoCrossAttn, TriplaneTransformer: block wiring, LayerNorm eps, GELU MLPs, pos_embed.repeat, the
einsum / ConvTranspose2d plane assembly). Its one external dependency, `diffusers.models.attention_processor.Attention`,
is not installed here (diffusers 0.19 per the reference's requirements.txt): the stand-in below restates the documented
behaviour of that class for the arguments the reference passes (query_dim, heads, dim_head, cross_attention_dim,
dropout = 0, bias = False): `to_q / to_k / to_v` = Linear without bias, `to_out = [Linear(inner, query_dim) with bias,
Dropout]`, softmax(q k^T * dim_head ** -0.5) v per head, no residual, no rescaling.
Output: tests/golden/triplane_generator_golden.pt (two small configurations, about 0.5 MB).
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF = "/root/reference/custom/amortized/extern/triplane_transformer_modules.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "triplane_generator_golden.pt")


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False):
        super().__init__()
        inner = heads * dim_head
        cross = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cross, inner, bias=bias)
        self.to_v = nn.Linear(cross, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        B, L, _ = hidden_states.shape
        split = lambda t: t.view(B, t.shape[1], self.heads, -1).transpose(1, 2)
        q, k, v = split(self.to_q(hidden_states)), split(self.to_k(ctx)), split(self.to_v(ctx))
        p = torch.softmax(q @ k.transpose(-1, -2) * self.scale, dim=-1)
        return self.to_out[1](self.to_out[0]((p @ v).transpose(1, 2).reshape(B, L, -1)))


def load_reference():
    stubs = {}
    for name in ("diffusers", "diffusers.models", "diffusers.models.attention_processor", "threestudio", "threestudio.utils",
                 "threestudio.utils.typing"):
        stubs[name] = types.ModuleType(name)
    stubs["diffusers.models.attention_processor"].Attention = Attention
    exec("from typing import *", stubs["threestudio.utils.typing"].__dict__)
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        ns = {"__name__": "ref_triplane_transformer_modules"}
        exec(compile(open(REF).read(), REF, "exec"), ns)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ns["TriplaneTransformer"]


def main():
    RefGen = load_reference()
    gold = {}
    for tag, local_text, cond_shape in (("local_text", True, (2, 7, 48)), ("global_text", False, (2, 48))):
        cfg = dict(inner_dim=64, condition_dim=48, triplane_low_res=4, triplane_high_res=8, triplane_dim=8, num_layers=2,
                   num_heads=4, local_text=local_text)
        torch.manual_seed(11 if local_text else 12)
        gen = RefGen(**cfg).double()
        with torch.no_grad():  # non-trivial norms and biases
            for n, p in gen.named_parameters():
                if "norm" in n or n.endswith("bias"):
                    p.add_(0.1 * torch.randn_like(p))
        emb = torch.randn(*cond_shape, dtype=torch.float64)
        with torch.no_grad():
            planes = gen(emb)
        gold[tag] = {"cfg": cfg, "state_dict": {k: v.float() for k, v in gen.state_dict().items()}, "text_embed": emb.float(),
                     "planes": planes.float()}
        print(tag, tuple(planes.shape), "parameters", sum(p.numel() for p in gen.parameters()))
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
