"""Host-side wrappers of the kernel-level dense C ABI (include/sdb200_nn.h). Tensor plumbing only.

Layout convention: activations are fp16 channels-last, `[N, H, W, C]` == token matrix `[N*H*W, C]`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L

ACT = {None: 0, "none": 0, "silu": 1, "gelu": 2}


def _h(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if t.dtype != torch.float16:
        raise RuntimeError(f"expected float16 tensor, got {t.dtype}")
    return L.ptr(t)


def gemm(a: torch.Tensor, b: torch.Tensor, bias=None, rowbias=None, rows_per_group=1, residual=None, alpha=1.0,
         act=None, out_fp32=False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = act(alpha * a @ b.T + bias + rowbias[row // rows_per_group]) + residual.
    a [M,K] or [Z,M,K]; b [N,K] or [Z,N,K] (fp16, row-major, K % 8 == 0)."""
    lib = L.load()
    batched = a.dim() == 3
    Z = a.shape[0] if batched else 1
    M, K = a.shape[-2], a.shape[-1]
    N = b.shape[-2]
    if out is None:
        shape = (Z, M, N) if batched else (M, N)
        out = torch.empty(shape, device=a.device, dtype=torch.float32 if out_fp32 else torch.float16)
    args = L.GemmArgsC()
    args.A, args.lda = _h(a), a.stride(-2)
    args.B, args.ldb = _h(b), b.stride(-2)
    args.out, args.ldc = L.ptr(out), out.stride(-2)
    args.M, args.N, args.K = M, N, K
    args.bias = _h(bias)
    args.rowbias = L.ptr(rowbias) if rowbias is not None else None
    args.rows_per_group = rows_per_group
    args.residual = _h(residual)
    args.ldr = residual.stride(-2) if residual is not None else 0
    args.alpha = alpha
    args.act = ACT[act]
    args.out_fp32 = int(out_fp32)
    args.batch = Z
    args.a_zs = a.stride(0) if batched else 0
    args.b_zs = b.stride(0) if b.dim() == 3 else 0
    args.out_zs = out.stride(0) if batched else 0
    L.check(lib.sdb_gemm_f16(C.byref(args), L.stream_ptr()), "sdb_gemm_f16")
    return out


def conv3x3(x: torch.Tensor, w: torch.Tensor, bias=None, rowbias=None, residual=None, act=None) -> torch.Tensor:
    """x [N,H,W,Cin] fp16, w [Cout,3,3,Cin] fp16 -> [N,H,W,Cout]."""
    lib = L.load()
    n, h, wd, cin = x.shape
    cout = w.shape[0]
    out = torch.empty(n, h, wd, cout, device=x.device, dtype=torch.float16)
    L.check(lib.sdb_conv3x3_f16(_h(x), n, h, wd, cin, _h(w), cout, _h(bias),
                                L.ptr(rowbias) if rowbias is not None else None, _h(residual), ACT[act], L.ptr(out),
                                L.stream_ptr()), "sdb_conv3x3_f16")
    return out


def conv3x3_small(x: torch.Tensor, w: torch.Tensor, bias=None, out_fp32=False) -> torch.Tensor:
    lib = L.load()
    n, h, wd, cin = x.shape
    cout = w.shape[0]
    out = torch.empty(n, h, wd, cout, device=x.device, dtype=torch.float32 if out_fp32 else torch.float16)
    L.check(lib.sdb_conv3x3_small(L.ptr(x), int(x.dtype == torch.float32), _h(w), _h(bias), L.ptr(out), int(out_fp32),
                                  n, h, wd, cin, cout, L.stream_ptr()), "sdb_conv3x3_small")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """q [B,Lq,heads*64], k/v [B,Lk,heads*64] -> [B,Lq,heads*64]."""
    lib = L.load()
    B, Lq, inner = q.shape
    Lk = k.shape[1]
    lds = (Lk + 7) // 8 * 8
    scores = torch.empty(B * heads * Lq * lds, device=q.device, dtype=torch.float16)
    out = torch.empty(B, Lq, inner, device=q.device, dtype=torch.float16)
    L.check(lib.sdb_attention_f16(_h(q), q.stride(1), _h(k), k.stride(1), _h(v), v.stride(1), B, heads, Lq, Lk,
                                  L.ptr(scores), L.ptr(out), out.stride(1), L.stream_ptr()), "sdb_attention_f16")
    return out


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """Same contract as attention(), one fused kernel (no score matrix). q/k/v may be column slices of a fused
    [B, L, 3*heads*64] projection (row stride = stride(1))."""
    lib = L.load()
    B, Lq, inner = q.shape
    Lk = k.shape[1]
    out = torch.empty(B, Lq, inner, device=q.device, dtype=torch.float16)
    for t in (q, k, v):
        if t.dtype != torch.float16 or t.stride(2) != 1 or t.stride(0) != t.shape[1] * t.stride(1):
            raise RuntimeError("flash_attention: fp16 tensors with unit channel stride and packed batches expected")
    L.check(lib.sdb_flash_attention_f16(q.data_ptr(), q.stride(1), k.data_ptr(), k.stride(1), v.data_ptr(), v.stride(1),
                                        B, heads, Lq, Lk, L.ptr(out), out.stride(1), L.stream_ptr()),
            "sdb_flash_attention_f16")
    return out


def groupnorm(x: torch.Tensor, gamma, beta, groups=32, eps=1e-5, silu=False):
    """x [N, ..., C] channels-last. Returns (y, stats)."""
    lib = L.load()
    n, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (n * c)
    y = torch.empty_like(x)
    stats = torch.empty(lib.sdb_groupnorm_workspace_floats(n, hw, c, groups), device=x.device, dtype=torch.float32)
    L.check(lib.sdb_groupnorm_f16(_h(x), _h(gamma), _h(beta), L.ptr(y), L.ptr(stats), n, hw, c, groups, eps, int(silu),
                                  L.stream_ptr()), "sdb_groupnorm_f16")
    return y, stats


def groupnorm_backward(x, gamma, beta, stats, dy, groups=32, eps=1e-5, silu=False):
    lib = L.load()
    n, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (n * c)
    dx = torch.empty_like(x)
    scratch = torch.empty(lib.sdb_groupnorm_workspace_floats(n, hw, c, groups), device=x.device, dtype=torch.float32)
    L.check(lib.sdb_groupnorm_backward_f16(_h(x), _h(gamma), _h(beta), L.ptr(stats), _h(dy), L.ptr(dx), L.ptr(scratch),
                                           n, hw, c, groups, eps, int(silu), L.stream_ptr()),
            "sdb_groupnorm_backward_f16")
    return dx


def layernorm(x, gamma, beta, eps=1e-5):
    lib = L.load()
    c = x.shape[-1]
    y = torch.empty_like(x)
    L.check(lib.sdb_layernorm_f16(_h(x), _h(gamma), _h(beta), L.ptr(y), x.numel() // c, c, eps, L.stream_ptr()),
            "sdb_layernorm_f16")
    return y


def geglu(xg):
    lib = L.load()
    inner = xg.shape[-1] // 2
    rows = xg.numel() // (2 * inner)
    y = torch.empty(*xg.shape[:-1], inner, device=xg.device, dtype=torch.float16)
    L.check(lib.sdb_geglu_f16(_h(xg), L.ptr(y), rows, inner, L.stream_ptr()), "sdb_geglu_f16")
    return y


def upsample2x(x):
    lib = L.load()
    n, h, w, c = x.shape
    y = torch.empty(n, 2 * h, 2 * w, c, device=x.device, dtype=torch.float16)
    L.check(lib.sdb_upsample2x_f16(_h(x), L.ptr(y), n, h, w, c, L.stream_ptr()), "sdb_upsample2x_f16")
    return y


def im2col3x3s2(x, pad_lo):
    lib = L.load()
    n, h, w, c = x.shape
    col = torch.empty(n, h // 2, w // 2, 9 * c, device=x.device, dtype=torch.float16)
    L.check(lib.sdb_im2col3x3s2_f16(_h(x), L.ptr(col), n, h, w, c, pad_lo, L.stream_ptr()), "sdb_im2col3x3s2_f16")
    return col


def col2im3x3s2(col, h, w, pad_lo):
    lib = L.load()
    n, c = col.shape[0], col.shape[-1] // 9
    dx = torch.empty(n, h, w, c, device=col.device, dtype=torch.float16)
    L.check(lib.sdb_col2im3x3s2_f16(_h(col), L.ptr(dx), n, h, w, c, pad_lo, L.stream_ptr()), "sdb_col2im3x3s2_f16")
    return dx
