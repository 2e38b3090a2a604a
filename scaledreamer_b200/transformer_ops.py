"""Tensor-level wrappers of the fp32 / tf32 kernels the native Triplane-Transformer is built from
(include/sdb200_nn.h, "fp32 / tf32 kernels of the trained Triplane-Transformer generator"). Every function launches
this library's kernels on the current stream and nothing else; there is no torch fallback."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import lib as L

ACT_NONE, ACT_GELU = 0, 2


class Operand:
    """One GEMM operand: a device pointer with a row stride and two batch strides (floats). `t` keeps the storage alive."""

    __slots__ = ("t", "offset", "ld", "zs_hi", "zs_lo")

    def __init__(self, t: torch.Tensor, ld: int, zs_hi: int = 0, zs_lo: int = 0, offset: int = 0):
        if t.dtype != torch.float32 or not t.is_cuda:
            raise RuntimeError("tf32 GEMM operands are fp32 CUDA tensors")
        self.t, self.ld, self.zs_hi, self.zs_lo, self.offset = t, int(ld), int(zs_hi), int(zs_lo), int(offset)

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + 4 * self.offset


def mat(t: torch.Tensor) -> Operand:
    """A contiguous [rows, cols] (or [..., cols] flattened) matrix shared by every batch index."""
    return Operand(t, t.shape[-1])


def gemm(A: Operand, B: Operand, M: int, N: int, K: int, out: Operand, batch: int = 1, zdiv: int = 1,
         bias: Optional[torch.Tensor] = None, residual: Optional[Operand] = None, alpha: float = 1.0,
         act: int = ACT_NONE, round_out: bool = False, a_mn_major: bool = False) -> None:
    """out[z] = act(alpha * A[z] B[z]^T + bias) + residual[z]; z = hi * zdiv + lo < batch. round_out: the result is
    written rounded to the nearest tf32 (for tensors that only feed further GEMMs; the tensor core truncates otherwise).
    a_mn_major: A is handed over TRANSPOSED, [K rows][M] with A.ld floats between K rows (no transpose pass)."""
    a = L.GemmTf32ArgsC()
    a.A, a.lda, a.a_zs_hi, a.a_zs_lo = A.ptr, A.ld, A.zs_hi, A.zs_lo
    a.B, a.ldb, a.b_zs_hi, a.b_zs_lo = B.ptr, B.ld, B.zs_hi, B.zs_lo
    a.M, a.N, a.K, a.batch, a.zdiv = M, N, K, batch, zdiv
    a.out, a.ldc, a.out_zs_hi, a.out_zs_lo = out.ptr, out.ld, out.zs_hi, out.zs_lo
    a.bias = bias.data_ptr() if bias is not None else None
    if residual is not None:
        a.residual, a.ldr, a.res_zs_hi, a.res_zs_lo = residual.ptr, residual.ld, residual.zs_hi, residual.zs_lo
    a.alpha, a.act, a.round_out = alpha, act, 1 if round_out else 0
    a.a_mn_major = 1 if a_mn_major else 0
    L.check(L.load().sdb_gemm_tf32(a, L.stream_ptr()), "sdb_gemm_tf32")


def round_tf32(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Nearest-tf32 copy of a contiguous fp32 tensor (out=src rounds in place)."""
    out = torch.empty_like(src) if out is None else out
    L.check(L.load().sdb_round_tf32_f32(src.data_ptr(), out.data_ptr(), src.numel(), L.stream_ptr()), "sdb_round_tf32_f32")
    return out


def transpose(src: torch.Tensor, rows: int, cols: int, batch: int = 1, ld_out: Optional[int] = None,
              round_out: bool = True) -> torch.Tensor:
    """[batch, rows, cols] contiguous -> [batch, cols, ld_out >= rows] (ld_out: rows rounded up to a multiple of 4).
    Transposes exist to feed GEMMs, so the values are rounded to tf32 on the way unless round_out is False."""
    ld_out = ld_out or (rows + 3) // 4 * 4
    out = torch.empty(batch, cols, ld_out, device=src.device, dtype=torch.float32)
    L.check(L.load().sdb_transpose_f32(src.data_ptr(), cols, rows * cols, out.data_ptr(), ld_out, cols * ld_out, rows,
                                       cols, batch, 1 if round_out else 0, L.stream_ptr()), "sdb_transpose_f32")
    return out


def layernorm_forward(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, round_out: bool = False
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    L.check(L.load().sdb_layernorm_f32_forward(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                                               mean.data_ptr(), rstd.data_ptr(), rows, C, eps, 1 if round_out else 0,
                                               L.stream_ptr()),
            "sdb_layernorm_f32_forward")
    return y, mean, rstd


def layernorm_backward(x, gamma, mean, rstd, dy, dskip: Optional[torch.Tensor]
                       ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (dx [+ dskip], d_gamma, d_beta)"""
    lib = L.load()
    C = x.shape[-1]
    rows = x.numel() // C
    dx = torch.empty_like(x)
    ws = torch.empty(int(lib.sdb_layernorm_f32_backward_ws_floats(rows, C)), device=x.device, dtype=torch.float32)
    dgb = torch.empty(2, C, device=x.device, dtype=torch.float32)
    L.check(lib.sdb_layernorm_f32_backward(x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                           dy.data_ptr(), dskip.data_ptr() if dskip is not None else None,
                                           dx.data_ptr(), ws.data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr(), rows, C,
                                           L.stream_ptr()), "sdb_layernorm_f32_backward")
    return dx, dgb[0], dgb[1]


def softmax_forward_(x: torch.Tensor, rows: int, cols: int, ld: int, round_out: bool = False) -> torch.Tensor:
    """In place over x viewed as [rows, ld]; -> lse [rows]."""
    lse = torch.empty(rows, device=x.device, dtype=torch.float32)
    L.check(L.load().sdb_softmax_f32_forward(x.data_ptr(), rows, cols, ld, lse.data_ptr(), 1 if round_out else 0,
                                             L.stream_ptr()),
            "sdb_softmax_f32_forward")
    return lse


def softmax_backward_rows_(X: torch.Tensor, Y: torch.Tensor, rows: int, cols: int, ld: int, lse: torch.Tensor,
                           round_out: bool = False, write_p: bool = True) -> torch.Tensor:
    """X (scores) <- P (unless write_p is False), Y (dP) <- dS, -> delta [rows] = sum_k P dP."""
    delta = torch.empty(rows, device=X.device, dtype=torch.float32)
    L.check(L.load().sdb_softmax_f32_backward_rows(X.data_ptr(), Y.data_ptr(), rows, cols, ld, lse.data_ptr(),
                                                   delta.data_ptr(), 1 if round_out else 0, 1 if write_p else 0,
                                                   L.stream_ptr()), "sdb_softmax_f32_backward_rows")
    return delta


def softmax_backward_stats_(X: torch.Tensor, Y: torch.Tensor, batch: int, rows: int, cols: int, ld: int,
                            lse: torch.Tensor, delta: torch.Tensor, by_col: bool, round_out: bool = False) -> None:
    L.check(L.load().sdb_softmax_f32_backward_stats(X.data_ptr(), Y.data_ptr(), batch, rows, cols, ld, lse.data_ptr(),
                                                    delta.data_ptr(), 1 if by_col else 0, 1 if round_out else 0,
                                                    L.stream_ptr()), "sdb_softmax_f32_backward_stats")


def attn_delta(dO: torch.Tensor, O: torch.Tensor, B: int, Lq: int, heads: int, d: int) -> torch.Tensor:
    delta = torch.empty(B * heads * Lq, device=O.device, dtype=torch.float32)
    L.check(L.load().sdb_attn_delta_f32(dO.data_ptr(), O.data_ptr(), delta.data_ptr(), B, Lq, heads, d, L.stream_ptr()),
            "sdb_attn_delta_f32")
    return delta


def gelu_forward(h: torch.Tensor, round_out: bool = False) -> torch.Tensor:
    g = torch.empty_like(h)
    L.check(L.load().sdb_gelu_f32_forward(h.data_ptr(), g.data_ptr(), h.numel(), 1 if round_out else 0, L.stream_ptr()),
            "sdb_gelu_f32_forward")
    return g


def gelu_backward_(h: torch.Tensor, dg: torch.Tensor, round_out: bool = False) -> torch.Tensor:
    L.check(L.load().sdb_gelu_f32_backward(h.data_ptr(), dg.data_ptr(), h.numel(), 1 if round_out else 0, L.stream_ptr()),
            "sdb_gelu_f32_backward")
    return dg


def colsum(x: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    lib = L.load()
    ws = torch.empty(int(lib.sdb_colsum_f32_ws_floats(rows, cols)), device=x.device, dtype=torch.float32)
    out = torch.empty(cols, device=x.device, dtype=torch.float32)
    L.check(lib.sdb_colsum_f32(x.data_ptr(), rows, cols, cols, ws.data_ptr(), out.data_ptr(), L.stream_ptr()),
            "sdb_colsum_f32")
    return out


def broadcast(src: torch.Tensor, copies: int) -> torch.Tensor:
    out = torch.empty(copies, *src.shape, device=src.device, dtype=torch.float32)
    L.check(L.load().sdb_broadcast_f32(src.data_ptr(), src.numel(), out.data_ptr(), copies, L.stream_ptr()),
            "sdb_broadcast_f32")
    return out


def deconv_shuffle(src: torch.Tensor, planes: int, H: int, W: int, D: int, inverse: bool) -> torch.Tensor:
    """inverse False: t [planes*H*W, 4D] -> p [planes, 2H, 2W, D]; True: the other way (gradient)."""
    out = (torch.empty(planes * H * W, 4 * D, device=src.device, dtype=torch.float32) if inverse
           else torch.empty(planes, 2 * H, 2 * W, D, device=src.device, dtype=torch.float32))
    L.check(L.load().sdb_deconv_shuffle_f32(src.data_ptr(), out.data_ptr(), planes, H, W, D, 1 if inverse else 0,
                                            L.stream_ptr()), "sdb_deconv_shuffle_f32")
    return out
