// extern "C" entry points of include/sdb200_nn.h (kernel-level): validation + dispatch into dense::.
#include "../../include/sdb200_nn.h"
#include "dense.h"

using namespace dense;

extern "C" {

int sdb_gemm_f16(const sdb_gemm_args* a, void* stream) {
  SDB_CHECK_ARG(a && a->A && a->B && a->out, "gemm: NULL argument");
  Epilogue ep;
  ep.out = a->out;
  ep.out_fp32 = a->out_fp32;
  ep.ldc = a->ldc;
  ep.bias = reinterpret_cast<const __half*>(a->bias);
  ep.rowbias = a->rowbias;
  ep.rows_per_group = a->rows_per_group;
  ep.residual = reinterpret_cast<const __half*>(a->residual);
  ep.ldr = a->ldr;
  ep.alpha = a->alpha;
  ep.act = a->act;
  GemmPlan plan;
  int rc = plan_gemm(&plan, reinterpret_cast<const __half*>(a->A), a->lda, reinterpret_cast<const __half*>(a->B),
                     a->ldb, a->M, a->N, a->K, ep, a->batch > 0 ? a->batch : 1, a->a_zs, a->b_zs, a->out_zs);
  if (rc) return rc;
  return run_gemm(plan, (cudaStream_t)stream);
}

int sdb_conv3x3_f16(const void* x, int n, int h, int w, int cin, const void* weight, int cout, const void* bias,
                    const float* rowbias, const void* residual, int act, void* out, void* stream) {
  SDB_CHECK_ARG(x && weight && out && n > 0 && h > 0 && w > 0, "conv3x3: bad arguments");
  Epilogue ep;
  ep.out = out;
  ep.ldc = cout;
  ep.bias = reinterpret_cast<const __half*>(bias);
  ep.rowbias = rowbias;
  ep.rows_per_group = h * w;
  ep.residual = reinterpret_cast<const __half*>(residual);
  ep.ldr = cout;
  ep.act = act;
  GemmPlan plan;
  int rc = plan_conv3x3(&plan, reinterpret_cast<const __half*>(x), n, h, w, cin, reinterpret_cast<const __half*>(weight),
                        cout, ep);
  if (rc) return rc;
  return run_gemm(plan, (cudaStream_t)stream);
}

int sdb_conv3x3_small(const void* x, int x_fp32, const void* weight, const void* bias, void* out, int out_fp32,
                      int n, int h, int w, int cin, int cout, void* stream) {
  SDB_CHECK_ARG(x && weight && out, "conv3x3_small: NULL argument");
  return conv3x3_small(x, x_fp32, reinterpret_cast<const __half*>(weight), reinterpret_cast<const __half*>(bias), out,
                       out_fp32, n, h, w, cin, cout, (cudaStream_t)stream);
}

int sdb_attention_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                      int batch, int heads, int lq, int lk, void* scores, void* out, long long ldo, void* stream) {
  SDB_CHECK_ARG(q && k && v && scores && out && batch > 0 && heads > 0 && lq > 0 && lk > 0, "attention: bad arguments");
  const long long lds = (lk + 7) / 8 * 8;
  GemmPlan ps, pa;
  int rc = plan_attn_scores(&ps, reinterpret_cast<const __half*>(q), ldq, reinterpret_cast<const __half*>(k), ldk, batch,
                            heads, 64, lq, lk, reinterpret_cast<__half*>(scores), lds, 0.125f);
  if (rc) return rc;
  rc = plan_attn_apply(&pa, reinterpret_cast<const __half*>(scores), lds, reinterpret_cast<const __half*>(v), ldv, batch,
                       heads, 64, lq, lk, reinterpret_cast<__half*>(out), ldo);
  if (rc) return rc;
  rc = run_gemm(ps, (cudaStream_t)stream);
  if (rc) return rc;
  rc = softmax_rows(reinterpret_cast<__half*>(scores), (long long)batch * heads * lq, lk, lds, (cudaStream_t)stream);
  if (rc) return rc;
  return run_gemm(pa, (cudaStream_t)stream);
}

int sdb_flash_attention_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                            int batch, int heads, int lq, int lk, void* out, long long ldo, void* stream) {
  SDB_CHECK_ARG(q && k && v && out && batch > 0 && heads > 0 && lq > 0 && lk > 0, "flash_attention: bad arguments");
  FlashPlan fp;
  int rc = plan_flash_attn(&fp, reinterpret_cast<const __half*>(q), ldq, reinterpret_cast<const __half*>(k), ldk,
                           reinterpret_cast<const __half*>(v), ldv, batch, heads, 64, lq, lk,
                           reinterpret_cast<__half*>(out), ldo, 0.125f);
  if (rc) return rc;
  return run_flash_attn(fp, (cudaStream_t)stream);
}

long long sdb_groupnorm_workspace_floats(int n, int hw, int c, int groups) {
  return groupnorm_workspace_floats(n, hw, c, groups);
}

int sdb_groupnorm_f16(const void* x, const void* gamma, const void* beta, void* y, float* stats, int n, int hw,
                      int c, int groups, float eps, int silu, void* stream) {
  SDB_CHECK_ARG(x && gamma && beta && y && stats, "groupnorm: NULL argument");
  return groupnorm_forward(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(gamma),
                           reinterpret_cast<const __half*>(beta), reinterpret_cast<__half*>(y), stats, n, hw, c, groups,
                           eps, silu, (cudaStream_t)stream);
}

int sdb_groupnorm_backward_f16(const void* x, const void* gamma, const void* beta, const float* stats,
                               const void* dy, void* dx, float* scratch, int n, int hw, int c, int groups,
                               float eps, int silu, void* stream) {
  SDB_CHECK_ARG(x && gamma && beta && stats && dy && dx && scratch, "groupnorm_backward: NULL argument");
  return groupnorm_backward(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(gamma),
                            reinterpret_cast<const __half*>(beta), stats, reinterpret_cast<const __half*>(dy),
                            reinterpret_cast<__half*>(dx), scratch, n, hw, c, groups, eps, silu, (cudaStream_t)stream);
}

int sdb_layernorm_f16(const void* x, const void* gamma, const void* beta, void* y, int rows, int c, float eps,
                      void* stream) {
  SDB_CHECK_ARG(x && gamma && beta && y, "layernorm: NULL argument");
  return layernorm_forward(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(gamma),
                           reinterpret_cast<const __half*>(beta), reinterpret_cast<__half*>(y), rows, c, eps,
                           (cudaStream_t)stream);
}

int sdb_geglu_f16(const void* xg, void* y, long long rows, int inner, void* stream) {
  SDB_CHECK_ARG(xg && y, "geglu: NULL argument");
  return geglu(reinterpret_cast<const __half*>(xg), reinterpret_cast<__half*>(y), rows, inner, (cudaStream_t)stream);
}

int sdb_upsample2x_f16(const void* x, void* y, int n, int h, int w, int c, void* stream) {
  SDB_CHECK_ARG(x && y, "upsample2x: NULL argument");
  return upsample_nearest2x(reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), n, h, w, c,
                            (cudaStream_t)stream);
}

int sdb_im2col3x3s2_f16(const void* x, void* col, int n, int h, int w, int c, int pad_lo, void* stream) {
  SDB_CHECK_ARG(x && col, "im2col: NULL argument");
  return im2col_3x3_s2(reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(col), n, h, w, c, pad_lo,
                       (cudaStream_t)stream);
}

int sdb_col2im3x3s2_f16(const void* col, void* dx, int n, int h, int w, int c, int pad_lo, void* stream) {
  SDB_CHECK_ARG(col && dx, "col2im: NULL argument");
  return col2im_3x3_s2(reinterpret_cast<const __half*>(col), reinterpret_cast<__half*>(dx), n, h, w, c, pad_lo,
                       (cudaStream_t)stream);
}

int sdb_gemm_profile_begin(void) {
  profile_begin();
  return SDB_OK;
}

int sdb_gemm_debug_timeline(unsigned long long* device_buf) {
  debug_timeline(device_buf);
  return SDB_OK;
}

int sdb_gemm_profile_dump(const char* csv_path) {
  profile_dump_to(csv_path);
  return SDB_OK;
}

int sdb_gemm_profile_end(double* total_ms, double* total_flops, int* launches) {
  SDB_CHECK_ARG(total_ms && total_flops && launches, "gemm_profile_end: NULL argument");
  return profile_end(total_ms, total_flops, launches);
}

}  // extern "C"
