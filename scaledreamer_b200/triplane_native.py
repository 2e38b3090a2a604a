"""The Triplane-Transformer generator (custom/amortized/extern/triplane_transformer_modules.py:33-71, 115-187) forward AND
backward on this library's kernels: every contraction is `sdb_gemm_tf32` (tcgen05 kind::tf32, fp32 in / out, the
arithmetic class of the reference's `precision: 32` cuBLAS path), LayerNorm / softmax / GELU / transposes / bias sums are
the fp32 kernels of csrc/transformer_ops.cu. One autograd node covers the whole network, so autograd itself launches
nothing (no gradient-accumulation adds, no permute copies) between the text embeddings and the planes.

Layout: activations are [N prompts, L = 3 * low_res^2 tokens, C] fp32; attention heads are strided views of them
(row stride C, head stride d, prompt stride L*C) passed to the GEMM as 4-D tensor maps, scores are [N*heads][Lq][Lk]
fp32 matrices that live only inside one attention call. The backward recomputes the scores and never transposes them:
dV = P^T dO and dK = dS^T Q hand P / dS to the GEMM as they are ([query][key]) with `a_mn_major`, i.e. the tensor core
reads its A operand MN-major; delta = sum_k P dP comes from the same pass that turns dP into dS.

Rounding: tcgen05 kind::tf32 ignores the low 13 mantissa bits of its fp32 operands (truncation; measured here as 5x the
error of torch's tf32 path after two blocks, because the bias compounds through chained GEMMs). Every GEMM operand is
therefore rounded to the NEAREST tf32 where it is produced -- LayerNorm / GELU / softmax outputs, GEMM epilogues of
q, k, v, O and their gradients, transposes, and one rounded copy of each weight matrix per step -- so the tensor core
sees exactly-representable inputs; the residual stream, pre-activations, scores and all reductions stay fp32.

Supported variant: `local_text: true` (ConditionModulationBlock: cross-attention to the 77 token embeddings), the one
the C5 yaml selects (configs/multi-prompt_benchmark/asd_mv_triplane_transformer_10k.yaml:52)."""
from __future__ import annotations

from typing import List

import torch

from . import transformer_ops as T
from .transformer_ops import Operand, mat

PER_LAYER = 20  # tensors per block in the flat parameter list (see flat_parameters)


def flat_parameters(gen) -> List[torch.Tensor]:
    """pos_embed, then per block [norm1 w b, cross q k v o ob, norm2 w b, self q k v o ob, norm3 w b, mlp w1 b1 w2 b2],
    then the final norm w b and the deconvolution weight."""
    ps = [gen.pos_embed]
    for blk in gen.layers:
        ps += [blk.norm1.weight, blk.norm1.bias]
        for a in (blk.cross_attn, blk.self_attn):
            ps += [a.to_q.weight, a.to_k.weight, a.to_v.weight, a.to_out[0].weight, a.to_out[0].bias]
            if a is blk.cross_attn:
                ps += [blk.norm2.weight, blk.norm2.bias]
        ps += [blk.norm3.weight, blk.norm3.bias, blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[3].weight, blk.mlp[3].bias]
    ps += [gen.norm.weight, gen.norm.bias, gen.deconv.weight]
    return ps


def _ceil4(n: int) -> int:
    return (n + 3) // 4 * 4


def _empty(*shape, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(*shape, device=like.device, dtype=torch.float32)


# ---- linear layers on [M, K] matrices ----------------------------------------------------------------------------
def _linear(x, W, b=None, residual=None, act=T.ACT_NONE, round_out=False):
    """x [M, K] (tf32-rounded) @ W [N, K]^T (tf32-rounded) + b (+ residual)."""
    M, K = x.shape
    N = W.shape[0]
    out = _empty(M, N, like=x)
    T.gemm(mat(x), mat(W), M, N, K, mat(out), bias=b, residual=mat(residual) if residual is not None else None, act=act,
           round_out=round_out)
    return out


def _dgrad(dy, W, into=None, round_out=False):
    """dy [M, N] (tf32-rounded) @ W [N, K] -> [M, K]; `into` (same shape) is accumulated into when given."""
    M, N = dy.shape
    K = W.shape[1]
    WT = T.transpose(W, N, K)[0]  # [K, N4], rounded on the way
    out = into if into is not None else _empty(M, K, like=dy)
    T.gemm(mat(dy), Operand(WT, WT.shape[-1]), M, K, N, mat(out), residual=mat(out) if into is not None else None,
           round_out=round_out)
    return out


def _tr(x, rows, cols):
    """[rows, cols] -> its tf32-rounded transpose [cols, rows4] (rows padded to a multiple of 4; the padding is never read)."""
    return T.transpose(x, rows, cols)[0]


def _wgrad(dyT, xT, M):
    """dy^T [N, M4] , x^T [K, M4] -> dW [N, K] = dy^T x. A 768 x 768 result is 36 tiles for 148 SMs: the contraction is
    cut into `s` slices run as one batched GEMM (partials [s][N][K]) and summed in slice order (deterministic)."""
    N, K = dyT.shape[0], xT.shape[0]
    tiles = ((N + 127) // 128) * ((K + 127) // 128)
    s = 1
    while tiles * s < 148 and s < 8 and M % (8 * s) == 0:
        s *= 2
    if s == 1:
        out = _empty(N, K, like=dyT)
        T.gemm(Operand(dyT, dyT.shape[-1]), Operand(xT, xT.shape[-1]), N, K, M, mat(out))
        return out
    part = _empty(s, N, K, like=dyT)
    T.gemm(Operand(dyT, dyT.shape[-1], M // s), Operand(xT, xT.shape[-1], M // s), N, K, M // s, Operand(part, K, N * K),
           batch=s)
    return T.colsum(part, s, N * K).view(N, K)


# ---- attention ---------------------------------------------------------------------------------------------------
class _Heads:
    """Strided per-(prompt, head) operands of a [B, L, heads * d] activation and of [B*heads][rows][ld] score matrices."""

    def __init__(self, B, heads, d):
        self.B, self.H, self.d, self.C = B, heads, d, heads * d

    def act(self, t, L):  # rows = tokens, K = d
        return Operand(t, self.C, L * self.C, self.d)

    def act_T(self, tT, ld):  # tT [B, C, ld]: rows = d (of one head), K = tokens
        return Operand(tT, ld, self.C * ld, self.d * ld)

    def scores(self, S, rows, ld):
        return Operand(S, ld, self.H * rows * ld, rows * ld)


def _attn_forward(xn, ctx, Wq, Wk, Wv, Wo, bo, x_res, heads):
    """x_res + to_out(softmax(q k^T / sqrt(d)) v); xn [B, L, C], ctx [B, Lk, Cc], the weights: all tf32-rounded.
    -> (y, saved)"""
    B, L, C = xn.shape
    Lk, Cc = ctx.shape[1], ctx.shape[2]
    d = C // heads
    hv = _Heads(B, heads, d)
    q = _linear(xn.view(B * L, C), Wq, round_out=True)
    k = _linear(ctx.view(B * Lk, Cc), Wk, round_out=True)
    v = _linear(ctx.view(B * Lk, Cc), Wv, round_out=True)
    Lkp = _ceil4(Lk)
    vT = _empty(B, C, Lkp, like=xn)  # V^T per prompt straight from the operand-swapped projection: no transpose kernel
    T.gemm(mat(Wv), Operand(ctx, Cc, Lk * Cc), C, Lk, Cc, Operand(vT, Lkp, C * Lkp), batch=B, round_out=True)
    S = _empty(B * heads, L, Lkp, like=xn)
    T.gemm(hv.act(q, L), hv.act(k, Lk), L, Lk, d, hv.scores(S, L, Lkp), batch=B * heads, zdiv=heads, alpha=d ** -0.5)
    lse = T.softmax_forward_(S, B * heads * L, Lk, Lkp, round_out=True)
    O = _empty(B, L, C, like=xn)
    T.gemm(hv.scores(S, L, Lkp), hv.act_T(vT, Lkp), L, d, Lk, hv.act(O, L), batch=B * heads, zdiv=heads, round_out=True)
    del S, vT
    y = _linear(O.view(B * L, C), Wo, bo, residual=x_res.view(B * L, C)).view(B, L, C)
    return y, (q, k, v, O, lse)


def _attn_backward(dy, saved, xn, ctx, Wq, Wk, Wv, Wo, heads, self_attn):
    """dy [B, L, C] = d loss / d (block output). -> (d xn, [dWq, dWk, dWv, dWo, dbo])"""
    q, k, v, O, lse = saved
    B, L, C = xn.shape
    Lk, Cc = ctx.shape[1], ctx.shape[2]
    d = C // heads
    hv = _Heads(B, heads, d)
    M, Mk = B * L, B * Lk
    scale = d ** -0.5
    dy2 = dy.view(M, C)
    dWo = _wgrad(_tr(dy2, M, C), _tr(O.view(M, C), M, C), M)
    dbo = T.colsum(dy2, M, C)
    dO = _dgrad(T.round_tf32(dy2), Wo, round_out=True).view(B, L, C)
    Lkp, Lp = _ceil4(Lk), _ceil4(L)
    # S -> P, dP -> dS (row form: delta = sum_k P dP from the very dP the rows are corrected with, see
    # softmax_f32_bwd_rows_kernel), dq = scale * dS k
    S = _empty(B * heads, L, Lkp, like=xn)
    T.gemm(hv.act(q, L), hv.act(k, Lk), L, Lk, d, hv.scores(S, L, Lkp), batch=B * heads, zdiv=heads, alpha=scale)
    dP = _empty(B * heads, L, Lkp, like=xn)
    T.gemm(hv.act(dO, L), hv.act(v, Lk), L, Lk, d, hv.scores(dP, L, Lkp), batch=B * heads, zdiv=heads)
    T.softmax_backward_rows_(S, dP, B * heads * L, Lk, Lkp, lse, round_out=True)
    kT = T.transpose(k, Lk, C, batch=B, ld_out=Lkp)  # [B, C, Lkp]
    dq = _empty(B, L, C, like=xn)
    T.gemm(hv.scores(dP, L, Lkp), hv.act_T(kT, Lkp), L, d, Lk, hv.act(dq, L), batch=B * heads, zdiv=heads, alpha=scale,
           round_out=True)
    # dv = P^T dO and dk = scale * dS^T q read P / dS [query][key] as they are: the GEMM takes A transposed (MN-major
    # tiles), so neither the 2.4 GB score matrices nor their gradients are transposed or recomputed key-major
    dOT = T.transpose(dO, L, C, batch=B, ld_out=Lp)
    dv = _empty(B, Lk, C, like=xn)
    T.gemm(hv.scores(S, L, Lkp), hv.act_T(dOT, Lp), Lk, d, L, hv.act(dv, Lk), batch=B * heads, zdiv=heads, round_out=True,
           a_mn_major=True)
    qT = T.transpose(q, L, C, batch=B, ld_out=Lp)
    dk = _empty(B, Lk, C, like=xn)
    T.gemm(hv.scores(dP, L, Lkp), hv.act_T(qT, Lp), Lk, d, L, hv.act(dk, Lk), batch=B * heads, zdiv=heads, alpha=scale,
           round_out=True, a_mn_major=True)
    del S, dP, kT, dOT, qT
    # projections
    xnT = _tr(xn.view(M, C), M, C)
    ctxT = xnT if self_attn else _tr(ctx.view(Mk, Cc), Mk, Cc)
    dWq = _wgrad(_tr(dq.view(M, C), M, C), xnT, M)
    dWk = _wgrad(_tr(dk.view(Mk, C), Mk, C), ctxT, Mk)
    dWv = _wgrad(_tr(dv.view(Mk, C), Mk, C), ctxT, Mk)
    dxn = _dgrad(dq.view(M, C), Wq)
    if self_attn:
        _dgrad(dk.view(Mk, C), Wk, into=dxn)
        _dgrad(dv.view(Mk, C), Wv, into=dxn)
    return dxn.view(B, L, C), [dWq, dWk, dWv, dWo, dbo]


# ---- the whole generator -------------------------------------------------------------------------------------------
class TriplaneTransformerFn(torch.autograd.Function):
    """planes [N, 3, D, 2H, 2H] (a view of channels-last storage [N, 3, 2H, 2H, D]) = generator(text_embed [N, 77, Cc])."""

    @staticmethod
    def forward(ctx, text_embed, heads: int, eps: float, low_res: int, *params):
        if not text_embed.is_cuda:
            raise RuntimeError("the Triplane-Transformer runs on this library's CUDA kernels only (no CPU path)")
        cond = T.round_tf32(text_embed.detach().float().contiguous())
        N = cond.shape[0]
        ps = [p.detach() for p in params]
        # one nearest-tf32 copy of every weight matrix per step (biases, norms and pos_embed are not GEMM operands)
        pr = [T.round_tf32(p) if (p.dim() == 2 and i > 0) else p for i, p in enumerate(ps)]
        n_layers = (len(ps) - 4) // PER_LAYER
        pos = ps[0]
        L, C = pos.shape[1], pos.shape[2]
        x = T.broadcast(pos[0], N)  # [N, L, C]
        saved = []
        for li in range(n_layers):
            (n1w, n1b, cq, ck, cv, co, cob, n2w, n2b, sq, sk, sv, so, sob, n3w, n3b, w1, b1, w2, b2) = \
                pr[1 + li * PER_LAYER: 1 + (li + 1) * PER_LAYER]
            x0 = x
            xn1, m1, r1 = T.layernorm_forward(x0, n1w, n1b, eps, round_out=True)
            x1, sa1 = _attn_forward(xn1, cond, cq, ck, cv, co, cob, x0, heads)
            xn2, m2, r2 = T.layernorm_forward(x1, n2w, n2b, eps, round_out=True)
            x2, sa2 = _attn_forward(xn2, xn2, sq, sk, sv, so, sob, x1, heads)
            xn3, m3, r3 = T.layernorm_forward(x2, n3w, n3b, eps, round_out=True)
            h = _linear(xn3.view(N * L, C), w1, b1)
            g = T.gelu_forward(h, round_out=True)
            x = _linear(g, w2, b2, residual=x2.view(N * L, C)).view(N, L, C)
            saved.append((x0, m1, r1, xn1, sa1, x1, m2, r2, xn2, sa2, x2, m3, r3, xn3, h, g))
        nw, nb, wd = ps[-3], ps[-2], ps[-1]
        xf, mf, rf = T.layernorm_forward(x, nw, nb, eps, round_out=True)
        D = wd.shape[1]
        wd_flat = wd.reshape(C, D * 4)
        wdT = _tr(wd_flat, C, D * 4)  # [4D, C]
        t = _linear(xf.view(N * L, C), wdT)  # C is a multiple of 32 (LayerNorm kernel), so wdT is exactly [4D, C]
        planes_cl = T.deconv_shuffle(t, N * 3, low_res, low_res, D, inverse=False).view(N, 3, 2 * low_res, 2 * low_res, D)
        del pr
        ctx.stuff = (cond, ps, saved, x, mf, rf, xf, heads, eps, low_res, n_layers)
        return planes_cl.permute(0, 1, 4, 2, 3)

    @staticmethod
    def backward(ctx, d_planes):
        cond, ps, saved, x_last, mf, rf, xf, heads, eps, low_res, n_layers = ctx.stuff
        N, L, C = x_last.shape
        M = N * L
        nw, nb, wd = ps[-3], ps[-2], ps[-1]
        D = wd.shape[1]
        d_cl = d_planes.permute(0, 1, 3, 4, 2)
        if not d_cl.is_contiguous():
            d_cl = d_cl.contiguous()
        d_cl = d_cl.float()
        dt = T.deconv_shuffle(d_cl, N * 3, low_res, low_res, D, inverse=True)  # [M, 4D]
        d_wd = _wgrad(_tr(xf.view(M, C), M, C), _tr(dt, M, 4 * D), M).view_as(wd)
        # dt @ wd_flat^T: wd_flat is already [N = C, K = 4D]
        dxf = _linear(T.round_tf32(dt), T.round_tf32(wd.reshape(C, 4 * D)))
        dx, d_nw, d_nb = T.layernorm_backward(x_last, nw, mf, rf, dxf.view(N, L, C), None)
        grads = [None] * len(ps)
        grads[-3], grads[-2], grads[-1] = d_nw, d_nb, d_wd
        for li in reversed(range(n_layers)):
            (n1w, n1b, cq, ck, cv, co, cob, n2w, n2b, sq, sk, sv, so, sob, n3w, n3b, w1, b1, w2, b2) = \
                ps[1 + li * PER_LAYER: 1 + (li + 1) * PER_LAYER]
            (x0, m1, r1, xn1, sa1, x1, m2, r2, xn2, sa2, x2, m3, r3, xn3, h, g) = saved[li]
            base = 1 + li * PER_LAYER
            # x_out = x2 + W2 gelu(W1 LN3(x2) + b1) + b2
            dy = dx.view(M, C)
            dyT = _tr(dy, M, C)
            Hd = w1.shape[0]
            grads[base + 18] = _wgrad(dyT, _tr(g, M, Hd), M)
            grads[base + 19] = T.colsum(dy, M, C)
            dh = T.gelu_backward_(h, _dgrad(T.round_tf32(dy), w2), round_out=True)
            grads[base + 16] = _wgrad(_tr(dh, M, Hd), _tr(xn3.view(M, C), M, C), M)
            grads[base + 17] = T.colsum(dh, M, Hd)
            dxn3 = _dgrad(dh, w1)
            del dh
            dx, grads[base + 14], grads[base + 15] = T.layernorm_backward(x2, n3w, m3, r3, dxn3.view(N, L, C), dx)
            # x2 = x1 + self_attn(LN2(x1))
            dxn2, ga = _attn_backward(dx, sa2, xn2, xn2, sq, sk, sv, so, heads, True)
            grads[base + 9: base + 14] = ga
            dx, grads[base + 7], grads[base + 8] = T.layernorm_backward(x1, n2w, m2, r2, dxn2, dx)
            # x1 = x0 + cross_attn(LN1(x0), cond)
            dxn1, ga = _attn_backward(dx, sa1, xn1, cond, cq, ck, cv, co, heads, False)
            grads[base + 2: base + 7] = ga
            dx, grads[base + 0], grads[base + 1] = T.layernorm_backward(x0, n1w, m1, r1, dxn1, dx)
        grads[0] = T.colsum(dx.view(N, L * C), N, L * C).view(1, L, C)
        return (None, None, None, None, *grads)


def generate_planes(gen, text_embed: torch.Tensor) -> torch.Tensor:
    """The generator's forward on the native kernels; `gen` is the state-dict-compatible module holding the parameters."""
    heads = gen.layers[0].self_attn.heads
    eps = float(gen.norm.eps)
    return TriplaneTransformerFn.apply(text_embed, heads, eps, int(gen.triplane_low_res), *flat_parameters(gen))
