/*
 * sdb200_nn.h — C ABI of the dense (tensor-core) half of libsdb200.so: the frozen VAE encoder and UNet of the
 * ASD guidance step and the kernels they are built from.
 *
 * The reference reaches this arithmetic through diffusers / the vendored LDM (cuDNN, cuBLAS, SDPA); each entry
 * point cites the reference call site it replaces (paths relative to the reference tree). Activations are fp16,
 * channels-last: an image batch is [N, H, W, C] and is at the same time the token matrix [N*H*W, C].
 * All pointers are DEVICE pointers owned by the caller; `stream` is a cudaStream_t; return 0 or a negative code
 * (sdb_last_error()). Nothing synchronises.
 */
#ifndef SDB200_NN_H
#define SDB200_NN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_ACT_NONE 0
#define SDB_ACT_SILU 1
#define SDB_ACT_GELU 2

/* ---- tcgen05 GEMM: out[M,N] = act(alpha * A[M,K] B[N,K]^T + bias[N] + rowbias[m / rows_per_group, N]) + residual
 * A, B, bias, residual fp16; rowbias fp32; out fp16 or fp32 (out_fp32). Leading dimensions in elements, multiples
 * of 8. batch > 1 repeats over z with element strides a_zs / b_zs / out_zs (0 = shared operand).
 * Replaces nn.Linear / torch.bmm calls, e.g. extern/mvdream/ldm/modules/attention.py:49-76,163-194. */
typedef struct {
  const void* A;
  long long lda;
  const void* B;
  long long ldb;
  void* out;
  long long ldc;
  int M, N, K;
  const void* bias;
  const float* rowbias;
  int rows_per_group;
  const void* residual;
  long long ldr;
  float alpha;
  int act;
  int out_fp32;
  int batch;
  long long a_zs, b_zs, out_zs;
} sdb_gemm_args;
int sdb_gemm_f16(const sdb_gemm_args* args, void* stream);

/* Measurement hooks: between begin and end every tcgen05 GEMM launch of the library is bracketed by CUDA events on
 * its own stream; end synchronises the device and returns the summed durations, algorithmic FLOPs (2*M*N*K per
 * launch) and launch count. */
int sdb_gemm_profile_begin(void);
int sdb_gemm_profile_end(double* total_ms, double* total_flops, int* launches);
/* HOST path: the next sdb_gemm_profile_end() also writes one CSV row per launch (M,N,K,batch,splits,bn,mode,ms). */
int sdb_gemm_profile_dump(const char* csv_path);
/* Diagnostics: non-NULL device buffer of >= 296*4 uint64 -> every GEMM CTA stamps %globaltimer at entry / after
 * set-up / first accumulator ready / exit. NULL switches it off. */
int sdb_gemm_debug_timeline(unsigned long long* device_buf);

/* 3x3 stride-1 pad-1 convolution as implicit GEMM: x [N,H,W,Cin], w [Cout, 3,3,Cin] (= [Cout, 9*Cin]),
 * out [N,H,W,Cout]; epilogue as sdb_gemm_f16 with rows_per_group = H*W (the per-image timestep-embedding add
 * of ResBlock, openaimodel.py:262-272). Cin % 64 == 0. Replaces conv_nd(...) of openaimodel.py:203,230 and
 * torch.nn.Conv2d of model.py:101-123. */
int sdb_conv3x3_f16(const void* x, int n, int h, int w, int cin, const void* weight, int cout, const void* bias,
                    const float* rowbias, const void* residual, int act, void* out, void* stream);

/* Direct (CUDA-core) 3x3 conv for the 3/4/8-channel ends of the networks; x fp32 or fp16, out fp32 or fp16. */
int sdb_conv3x3_small(const void* x, int x_fp32, const void* weight, const void* bias, void* out, int out_fp32,
                      int n, int h, int w, int cin, int cout, void* stream);

/* Multi-head attention with head_dim 64: q [B,Lq,heads*64] (row stride ldq), k/v [B,Lk,heads*64];
 * out[b,l,h*64+d] = softmax(q_h k_h^T / 8) v_h. scores: fp16 scratch of B*heads*Lq*round_up(Lk,8) elements.
 * Replaces CrossAttention.forward (attention.py:163-194) / diffusers AttnProcessor2_0. */
int sdb_attention_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                      int batch, int heads, int lq, int lk, void* scores, void* out, long long ldo, void* stream);
/* Same result without the scores buffer: one fused tcgen05 kernel (QK^T into tensor memory, online softmax,
 * PV from a shared-memory P tile). head_dim 64. Used for every self-attention of the UNet. */
int sdb_flash_attention_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                            int batch, int heads, int lq, int lk, void* out, long long ldo, void* stream);

/* GroupNorm over [N, HW, C] with optional fused SiLU (GroupNorm32 + SiLU, openaimodel.py:200-203; Normalize +
 * nonlinearity, model.py:129-141). stats / scratch: fp32 buffers of sdb_groupnorm_workspace_floats() elements; the
 * first [N, groups, 2] hold (sum, sum of squares), the rest per-block partials of the order-fixed reduction. */
long long sdb_groupnorm_workspace_floats(int n, int hw, int c, int groups);
int sdb_groupnorm_f16(const void* x, const void* gamma, const void* beta, void* y, float* stats, int n, int hw,
                      int c, int groups, float eps, int silu, void* stream);
/* dx of the above given dy; scratch: [N, groups, 2] fp32. */
int sdb_groupnorm_backward_f16(const void* x, const void* gamma, const void* beta, const float* stats,
                               const void* dy, void* dx, float* scratch, int n, int hw, int c, int groups,
                               float eps, int silu, void* stream);
int sdb_layernorm_f16(const void* x, const void* gamma, const void* beta, void* y, int rows, int c, float eps,
                      void* stream);
/* y[rows, inner] = xg[:, :inner] * gelu(xg[:, inner:]) (GEGLU, attention.py:49-57) */
int sdb_geglu_f16(const void* xg, void* y, long long rows, int inner, void* stream);
int sdb_upsample2x_f16(const void* x, void* y, int n, int h, int w, int c, void* stream);
int sdb_im2col3x3s2_f16(const void* x, void* col, int n, int h, int w, int c, int pad_lo, void* stream);
int sdb_col2im3x3s2_f16(const void* col, void* dx, int n, int h, int w, int c, int pad_lo, void* stream);

/* ---- network executors -------------------------------------------------------------------------------------
 * A net is created for a fixed (batch, height, width). Protocol:
 *   create -> sdb_net_sizes -> caller allocates the two arenas -> sdb_net_bind -> for every parameter reported by
 *   sdb_net_param: sdb_net_load_param(name, fp16 device tensor in the reported layout) -> sdb_net_finalize ->
 *   forward / backward any number of times. Parameter names are the reference's state-dict keys
 *   (extern/mvdream/ldm: "input_blocks.1.0.in_layers.2.weight", "encoder.down.0.block.0.conv1.weight", ...);
 *   layouts: 3x3 conv [Cout,3,3,Cin] (the reference's [Cout,Cin,3,3] permuted 0,2,3,1), 1x1 conv / linear
 *   [out,in], vectors [n]. */
typedef struct sdb_net sdb_net;

/* ---- fp32 / tf32 kernels of the trained Triplane-Transformer generator ---------------------------------------
 * custom/amortized/extern/triplane_transformer_modules.py:33-71 (ConditionModulationBlock: LayerNorm ->
 * diffusers Attention -> residual, x3) and :115-187 (TriplaneTransformer: pos_embed, 12 blocks, LayerNorm,
 * ConvTranspose2d). The reference runs them through torch.nn.functional / cuBLAS at `precision: 32`
 * (configs/multi-prompt_benchmark/asd_mv_triplane_transformer_10k.yaml:127); here every contraction is
 * sdb_gemm_tf32 (tcgen05 kind::tf32, fp32 accumulate) and the rest are the fp32 kernels below. */
typedef struct {
  const float* A;              /* [M, K] rows `lda` floats apart (multiple of 4); with a_mn_major: stored [K, M] */
  long long lda, a_zs_hi, a_zs_lo; /* batch strides in floats; 0 = shared along that batch coordinate */
  const float* B;              /* [N, K] */
  long long ldb, b_zs_hi, b_zs_lo;
  int M, N, K;
  int batch, zdiv;             /* batch index z = hi * zdiv + lo */
  float* out;                  /* out[hi*out_zs_hi + lo*out_zs_lo + m*ldc + n] */
  long long ldc, out_zs_hi, out_zs_lo;
  const float* bias;           /* [N] or NULL */
  const float* residual;       /* added after the activation, indexed like `out` with its own strides; or NULL */
  long long ldr, res_zs_hi, res_zs_lo;
  float alpha;
  int act;                     /* SDB_ACT_NONE | SDB_ACT_GELU */
  int round_out;               /* 1: results are rounded to the nearest tf32 (outputs that only feed further GEMMs) */
  int a_mn_major;              /* 1: A is given transposed, [K rows][M] with lda floats between K rows: out = A^T-free A B^T */
} sdb_gemm_tf32_args;
/* out = act(alpha * A B^T + bias) + residual, fp32 in / out, tf32 products. K, M, N need no padding.
 * The tensor core IGNORES the low 13 mantissa bits of its fp32 operands (truncation, a bias that compounds through
 * chained GEMMs): producers therefore write GEMM-only tensors already rounded to the nearest tf32 -- the `round_out`
 * flags below and sdb_round_tf32_f32 -- which is the rounding cuBLAS' tf32 path applies to its operands. */
int sdb_gemm_tf32(const sdb_gemm_tf32_args* a, void* stream);
/* out = nearest tf32 of in (in place allowed) */
int sdb_round_tf32_f32(const float* in, float* out, long long n, void* stream);
/* out[z][c][r] = in[z][r][c] */
int sdb_transpose_f32(const float* in, long long ld_in, long long zs_in, float* out, long long ld_out, long long zs_out,
                      int rows, int cols, int batch, int round_out, void* stream);
/* torch.nn.LayerNorm over the last dimension (C multiple of 32, <= 1024); mean / rstd [rows] are kept for the backward */
int sdb_layernorm_f32_forward(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                              int rows, int C, float eps, int round_out, void* stream);
long long sdb_layernorm_f32_backward_ws_floats(int rows, int C);
/* dx = dLayerNorm(dy) + dskip (dskip NULL: no skip branch); d_gamma / d_beta [C] are written, fixed summation order */
int sdb_layernorm_f32_backward(const float* x, const float* gamma, const float* mean, const float* rstd, const float* dy,
                               const float* dskip, float* dx, float* ws, float* d_gamma, float* d_beta, int rows, int C,
                               void* stream);
/* in-place row softmax of the first `cols` (<= 4096) entries of each row; lse[row] = log sum exp */
int sdb_softmax_f32_forward(float* x, long long rows, int cols, long long ld, float* lse, int round_out, void* stream);
/* Row form of the softmax backward: X holds scores [rows][cols] and Y d loss / d P. X <- P = exp(X - lse[row]) (only
 * when write_p is set), delta[row] = sum_k P dP (WRITTEN), Y <- dS = P (dP - delta) */
int sdb_softmax_f32_backward_rows(float* X, float* Y, long long rows, int cols, long long ld, const float* lse,
                                  float* delta, int round_out, int write_p, void* stream);
/* X <- P = exp(X - lse), Y <- P * (Y - delta): X holds scores [batch][rows][cols] and Y d loss / d P; the statistics
 * are indexed by row (by_col 0) or by column (by_col 1: X holds the TRANSPOSED scores) */
int sdb_softmax_f32_backward_stats(float* X, float* Y, int batch, int rows, int cols, long long ld, const float* lse,
                                   const float* delta, int by_col, int round_out, void* stream);
/* delta[(b*heads + h)*L + q] = sum_j dO[b,q,h*d+j] * O[b,q,h*d+j] */
int sdb_attn_delta_f32(const float* dO, const float* O, float* delta, int B, int L, int heads, int head_dim, void* stream);
int sdb_gelu_f32_forward(const float* h, float* g, long long n, int round_out, void* stream);
int sdb_gelu_f32_backward(const float* h, float* dg_inout, long long n, int round_out, void* stream);
long long sdb_colsum_f32_ws_floats(long long rows, int cols);
/* out[c] = sum_r x[r, c] (bias gradients; position-embedding gradient), fixed summation order */
int sdb_colsum_f32(const float* x, long long rows, int cols, long long ld, float* ws, float* out, void* stream);
/* out[k][i] = src[i], k < copies (pos_embed.repeat(N, 1, 1), triplane_transformer_modules.py:172) */
int sdb_broadcast_f32(const float* src, long long n, float* out, int copies, void* stream);
/* ConvTranspose2d(k 2, s 2) written as a GEMM leaves t[(plane,h,w)][d*4 + a*2 + b]; this moves it to the channels-last
 * plane p[plane][2h+a][2w+b][d] the triplane sampler reads (inverse 1: plane gradient -> GEMM layout) */
int sdb_deconv_shuffle_f32(const float* in, float* out, int planes, int H, int W, int D, int inverse, void* stream);

typedef struct {
  int in_channels, out_channels, model_channels;
  int num_levels;
  int channel_mult[4];
  int num_res_blocks;
  int attn_levels;   /* levels [0, attn_levels) carry a SpatialTransformer */
  int head_dim;
  int context_dim, context_len;
  int camera_dim;    /* 0 = single-view UNetModel; 16 = MultiViewUNetModel camera embedding */
  int num_frames;    /* 1, or 4 for MVDream (self-attention across the views of one object) */
} sdb_unet_cfg;

typedef struct {
  int in_channels, ch, num_levels;
  int ch_mult[4];
  int num_res_blocks;
  int z_channels;
} sdb_vae_cfg;

/* UNetModel.forward / MultiViewUNetModel.forward (openaimodel.py:777-808, 1175-1213); diffusers
 * UNet2DConditionModel as called by stable_diffusion_asd_guidance.py:318-331. */
int sdb_unet_create(const sdb_unet_cfg* cfg, int batch, int height, int width, sdb_net** out);
/* x fp16 [B,H,W,in_ch]; t fp32 [B]; ctx fp16 [B,context_len,context_dim]; camera fp16 [B,camera_dim] or NULL;
 * out fp32 [B,H,W,out_ch] */
int sdb_unet_forward(sdb_net* net, const void* x, const float* t, const void* ctx, const void* camera, float* out,
                     void* stream);

/* AutoencoderKL.encode up to (not including) quant_conv (ldm/models/autoencoder.py:81-85, model.py:518-543);
 * diffusers vae.encode as called by stable_diffusion_asd_guidance.py:170-178. */
int sdb_vae_encoder_create(const sdb_vae_cfg* cfg, int batch, int height, int width, sdb_net** out);
/* x fp32 [B,H,W,3]; h fp32 [B,H/8,W/8,2*z_channels] */
int sdb_vae_encoder_forward(sdb_net* net, const float* x, float* h, void* stream);
/* data gradient of the last forward: d_h fp32 -> d_x fp32 [B,H,W,3] (what autograd computes through the frozen
 * encoder in the reference's loss.backward()) */
int sdb_vae_encoder_backward(sdb_net* net, const float* d_h, float* d_x, void* stream);

void sdb_net_destroy(sdb_net* net);
int sdb_net_sizes(sdb_net* net, long long* weight_bytes, long long* work_bytes);
int sdb_net_bind(sdb_net* net, void* weights, void* work);
int sdb_net_num_params(sdb_net* net);
int sdb_net_param(sdb_net* net, int index, const char** name, int* ndim, int* shape4);
int sdb_net_load_param(sdb_net* net, const char* name, const void* src, long long numel, void* stream);
int sdb_net_finalize(sdb_net* net, void* stream);
int sdb_net_num_launches(sdb_net* net, int backward);

#ifdef __cplusplus
}
#endif
#endif /* SDB200_NN_H */
