// extern "C" entry points of the fp32 / tf32 Triplane-Transformer kernels (include/sdb200_nn.h): validation + dispatch.
#include "../../include/sdb200_nn.h"
#include "dense.h"

using namespace dense;

extern "C" {

int sdb_gemm_tf32(const sdb_gemm_tf32_args* a, void* stream) {
  SDB_CHECK_ARG(a && a->A && a->B && a->out, "gemm_tf32: NULL argument");
  SDB_CHECK_ARG((a->ldc & 3) == 0 || a->N < 4, "gemm_tf32: ldc=%lld should be a multiple of 4", a->ldc);
  Tf32Operand A{a->A, a->lda, a->a_zs_hi, a->a_zs_lo, a->a_mn_major}, B{a->B, a->ldb, a->b_zs_hi, a->b_zs_lo, 0};
  Tf32Epilogue ep;
  ep.bias = a->bias;
  ep.residual = a->residual;
  ep.ldr = a->ldr, ep.res_zs_hi = a->res_zs_hi, ep.res_zs_lo = a->res_zs_lo;
  ep.alpha = a->alpha;
  ep.act = a->act;
  ep.round_out = a->round_out;
  return gemm_tf32(A, B, a->M, a->N, a->K, a->out, a->ldc, a->batch > 0 ? a->batch : 1, a->zdiv > 0 ? a->zdiv : 1,
                   a->out_zs_hi, a->out_zs_lo, ep, (cudaStream_t)stream);
}

int sdb_round_tf32_f32(const float* in, float* out, long long n, void* stream) {
  SDB_CHECK_ARG(in && out && n > 0, "round_tf32_f32: bad arguments");
  return round_tf32_f32(in, out, n, (cudaStream_t)stream);
}

int sdb_transpose_f32(const float* in, long long ld_in, long long zs_in, float* out, long long ld_out, long long zs_out,
                      int rows, int cols, int batch, int round_out, void* stream) {
  SDB_CHECK_ARG(in && out, "transpose_f32: NULL argument");
  return transpose_f32(in, ld_in, zs_in, out, ld_out, zs_out, rows, cols, batch, round_out, (cudaStream_t)stream);
}

int sdb_layernorm_f32_forward(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                              int rows, int C, float eps, int round_out, void* stream) {
  SDB_CHECK_ARG(x && gamma && beta && y && mean && rstd && rows > 0, "layernorm_f32_forward: bad arguments");
  return layernorm_f32_forward(x, gamma, beta, y, mean, rstd, rows, C, eps, round_out, (cudaStream_t)stream);
}

long long sdb_layernorm_f32_backward_ws_floats(int rows, int C) { return layernorm_f32_backward_ws_floats(rows, C); }

int sdb_layernorm_f32_backward(const float* x, const float* gamma, const float* mean, const float* rstd, const float* dy,
                               const float* dskip, float* dx, float* ws, float* d_gamma, float* d_beta, int rows, int C,
                               void* stream) {
  SDB_CHECK_ARG(x && gamma && mean && rstd && dy && dx && ws && d_gamma && d_beta && rows > 0,
                "layernorm_f32_backward: bad arguments");
  return layernorm_f32_backward(x, gamma, mean, rstd, dy, dskip, dx, ws, d_gamma, d_beta, rows, C, (cudaStream_t)stream);
}

int sdb_softmax_f32_forward(float* x, long long rows, int cols, long long ld, float* lse, int round_out, void* stream) {
  SDB_CHECK_ARG(x && lse && rows > 0, "softmax_f32_forward: bad arguments");
  return softmax_f32_forward(x, rows, cols, ld, lse, round_out, (cudaStream_t)stream);
}

int sdb_softmax_f32_backward_rows(float* X, float* Y, long long rows, int cols, long long ld, const float* lse,
                                  float* delta, int round_out, int write_p, void* stream) {
  SDB_CHECK_ARG(X && Y && lse && delta && rows > 0 && cols > 0, "softmax_f32_backward_rows: bad arguments");
  return softmax_f32_backward_rows(X, Y, rows, cols, ld, lse, delta, round_out, write_p, (cudaStream_t)stream);
}

int sdb_softmax_f32_backward_stats(float* X, float* Y, int batch, int rows, int cols, long long ld, const float* lse,
                                   const float* delta, int by_col, int round_out, void* stream) {
  SDB_CHECK_ARG(X && Y && lse && delta && batch > 0 && rows > 0 && cols > 0, "softmax_f32_backward_stats: bad arguments");
  return softmax_f32_backward_stats(X, Y, batch, rows, cols, ld, lse, delta, by_col, round_out, (cudaStream_t)stream);
}

int sdb_attn_delta_f32(const float* dO, const float* O, float* delta, int B, int L, int heads, int head_dim, void* stream) {
  SDB_CHECK_ARG(dO && O && delta && B > 0 && L > 0 && heads > 0, "attn_delta_f32: bad arguments");
  return attn_delta_f32(dO, O, delta, B, L, heads, head_dim, (cudaStream_t)stream);
}

int sdb_gelu_f32_forward(const float* h, float* g, long long n, int round_out, void* stream) {
  SDB_CHECK_ARG(h && g && n > 0, "gelu_f32_forward: bad arguments");
  return gelu_f32_forward(h, g, n, round_out, (cudaStream_t)stream);
}

int sdb_gelu_f32_backward(const float* h, float* dg_inout, long long n, int round_out, void* stream) {
  SDB_CHECK_ARG(h && dg_inout && n > 0, "gelu_f32_backward: bad arguments");
  return gelu_f32_backward(h, dg_inout, n, round_out, (cudaStream_t)stream);
}

long long sdb_colsum_f32_ws_floats(long long rows, int cols) { return colsum_f32_ws_floats(rows, cols); }

int sdb_colsum_f32(const float* x, long long rows, int cols, long long ld, float* ws, float* out, void* stream) {
  SDB_CHECK_ARG(x && ws && out && rows > 0 && cols > 0, "colsum_f32: bad arguments");
  return colsum_f32(x, rows, cols, ld, ws, out, (cudaStream_t)stream);
}

int sdb_broadcast_f32(const float* src, long long n, float* out, int copies, void* stream) {
  SDB_CHECK_ARG(src && out && n > 0 && copies > 0, "broadcast_f32: bad arguments");
  return broadcast_f32(src, n, out, copies, (cudaStream_t)stream);
}

int sdb_deconv_shuffle_f32(const float* in, float* out, int planes, int H, int W, int D, int inverse, void* stream) {
  SDB_CHECK_ARG(in && out && planes > 0 && H > 0 && W > 0 && D > 0, "deconv_shuffle_f32: bad arguments");
  return deconv_shuffle_f32(in, out, planes, H, W, D, inverse, (cudaStream_t)stream);
}

}  // extern "C"
