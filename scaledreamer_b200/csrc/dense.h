// Internal C++ interface of the dense (tensor-core) half of libsdb200: the tcgen05 GEMM / implicit-conv kernel
// plans and the element-wise / normalisation kernels the UNet and VAE executors are built from.
// Activations are fp16, channels-last (NHWC == [tokens, C]); accumulation and statistics are fp32.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace dense {

enum OperandMode : int {
  kMatrix = 0,  // 3D map {K, rows, Z}
  kConv3x3 = 1, // 4D map {C, W, H, N}: implicit-GEMM 3x3 stride-1 pad-1 over NHWC, K ordered (kh, kw, cin)
  kHeads = 2,   // 4D map {d, heads, L, B}: per-head view of a [B, L, heads*d] tensor, z = b*heads + h
};

// kActGeglu: the weight rows are interleaved (interleave_geglu_rows) so every 32-column chunk holds 16 value columns
// followed by their 16 gate columns; the epilogue writes value * gelu(gate) into N/2 output columns.
enum Act : int { kActNone = 0, kActSilu = 1, kActGelu = 2, kActGeglu = 3 };

struct OperandGeom {
  int mode;
  int batched;     // kMatrix: z indexes dim 2 (else coordinate 0)
  int heads;       // kHeads
  int mn_major;    // B only: tile is [K rows][64 MN] (e.g. V of attention); requires BN == 64
  int W, H;        // kConv3x3 image size
  int bw, bh, bn;  // kConv3x3: the 128 pixels of a tile are bw x bh x bn (columns, rows, images)
  int split_dim;   // kConv3x3: the tile is fetched as two 64-pixel boxes, halved along w (1), h (2) or n (3)
  int cin_blocks;  // kConv3x3: Cin / 64
};

struct GemmParams {
  int M, N, K;
  int num_k_blocks;
  OperandGeom a, b;
  // epilogue: v = alpha*acc + bias[n] + rowbias[(m / rows_per_group)*N + n]; v = act(v); v += residual[m*ldr + n]
  void* out;
  int out_fp32;
  long long ldc;
  int out_zdiv;                 // output offset(z) = (z / out_zdiv)*out_zs_hi + (z % out_zdiv)*out_zs_lo
  long long out_zs_hi, out_zs_lo;
  const __half* bias;
  const float* rowbias;
  int rows_per_group;
  long long rowbias_ld;
  const __half* residual;
  long long ldr;
  float alpha;
  int act;
  int splits;          // > 1: split-K over blockIdx.z, fp32 partial planes in ws, epilogue by splitk_finalize_kernel
  float* ws;           // [splits][M][N] fp32
  int tiles_m, tiles_n, tiles_z;  // tile grid walked by the persistent CTAs (filled at launch)
  unsigned long long* dbg;  // optional [gridDim.x][4] globaltimer stamps: entry, setup done, first accumulator, exit
  int row_softmax;     // BN == 80 only: the epilogue applies softmax over the (single-tile) row of N <= 80 scores
  int tma_store;       // fp16 results leave through shared memory + cp.async.bulk.tensor stores (plan->tc)
};

struct GemmPlan {
  CUtensorMap ta, tb;
  CUtensorMap tc;  // output [M, N] fp16 as 16-column x 32-row boxes (32-byte swizzle), valid when p.tma_store
  GemmParams p;
  dim3 grid;
  int bn;  // 64, 128 or 160
};

// dtype: 0 = fp16, 1 = fp32 (the tf32 GEMM). dims/box innermost first; strides in BYTES for dims 1..rank-1.
int make_tmap(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle_bytes = 128, int dtype = 0);

struct Epilogue {
  void* out = nullptr;
  int out_fp32 = 0;
  long long ldc = 0;
  const __half* bias = nullptr;
  const float* rowbias = nullptr;
  int rows_per_group = 1;
  long long rowbias_ld = 0;  // 0 = N
  const __half* residual = nullptr;
  long long ldr = 0;
  float alpha = 1.f;
  int act = kActNone;
  float* splitk_ws = nullptr;  // gemm_splits(M,N,K) * M * N floats: lets the planner split K when M*N is small
};

// out[M,N] = A[M,K] * B[N,K]^T. lda/ldb in elements (multiples of 8). Optional batching over z with element strides.
int plan_gemm(GemmPlan* plan, const __half* A, long long lda, const __half* B, long long ldb, int M, int N, int K,
              const Epilogue& ep, int batch = 1, long long a_zs = 0, long long b_zs = 0, long long out_zs = 0);
// out[(n,h,w), Cout] = conv3x3(x[N,H,W,Cin], w[Cout, 9*Cin]) stride 1 pad 1.
int plan_conv3x3(GemmPlan* plan, const __half* x, int N, int H, int W, int Cin, const __half* w, int Cout,
                 const Epilogue& ep);
// S[b,h,Lq,Lk] = alpha * Q_h K_h^T for Q [B,Lq,heads*64] (row stride ldq), K [B,Lk,heads*64] (ldk). S row stride lds.
// fuse_softmax (Lk <= 80 only): S holds softmax(alpha Q K^T) and columns [Lk, lds) are zeroed.
int plan_attn_scores(GemmPlan* plan, const __half* Q, long long ldq, const __half* K, long long ldk, int B, int heads,
                     int head_dim, int Lq, int Lk, __half* S, long long lds, float alpha, int fuse_softmax = 0);
// O[b,Lq,h*64:(h+1)*64] = P[b,h,Lq,Lk] V_h for V [B,Lk,heads*64] (ldv); O row stride ldo.
int plan_attn_apply(GemmPlan* plan, const __half* P, long long ldp, const __half* V, long long ldv, int B, int heads,
                    int head_dim, int Lq, int Lk, __half* O, long long ldo, float alpha = 1.f);
// K splits plan_gemm / plan_conv3x3 use for this shape when the epilogue carries a workspace (1 = no split).
int gemm_splits(int M, int N, int K);
int run_gemm(const GemmPlan& plan, cudaStream_t stream);

// Fused attention (flash_attn_sm100.cu): O[b,Lq,h*64:(h+1)*64] = softmax(alpha Q_h K_h^T) V_h, head_dim 64, no score
// matrix in HBM. Q [B,Lq,heads*64] (row stride ldq), K/V [B,Lk,heads*64] (ldk/ldv), O row stride ldo.
struct FlashPlan {
  CUtensorMap tq, tk, tv;
  int Lq, Lk, heads;
  float scale_log2;
  __half* out;
  long long ldo;
  dim3 grid;
  double flops;
};
int plan_flash_attn(FlashPlan* plan, const __half* Q, long long ldq, const __half* K, long long ldk, const __half* V,
                    long long ldv, int B, int heads, int head_dim, int Lq, int Lk, __half* O, long long ldo, float alpha);
int run_flash_attn(const FlashPlan& plan, cudaStream_t stream);
// per-launch CUDA-event timing of every tcgen05 GEMM launched between begin and end (bench.py roofline)
int profile_mark_begin(double flops, int M, int N, int K, int z, cudaStream_t stream);  // -1 when not profiling
void profile_mark_end(int idx, cudaStream_t stream);
void debug_timeline(unsigned long long* device_buf);  // non-null: every GEMM launch stamps its CTAs into device_buf
void profile_begin();
void profile_dump_to(const char* path);  // the next profile_end() also writes one CSV row per launch
int profile_end(double* ms, double* flops, int* launches);

// ---- fp32-in / fp32-out GEMM on tcgen05 kind::tf32 (gemm_tf32_sm100.cu): the trained Triplane-Transformer -----------
// One operand of out[z][M, N] = alpha * A[z][M, K] * B[z][N, K]^T: row-major fp32, `ld` floats between rows (multiple of
// 4), batch index z = hi * zdiv + lo with element strides zs_hi / zs_lo (0 = the operand is shared along that index).
struct Tf32Operand {
  const float* ptr;
  long long ld, zs_hi, zs_lo;
  // A only: the matrix is stored TRANSPOSED, [K rows][M] with `ld` floats between K rows (M contiguous): the product
  // reads A^T without a transpose pass (MN-major shared-memory tiles). dV = P^T dO and dK = dS^T Q of the attention backward.
  int mn_major = 0;
};
struct Tf32Epilogue {
  const float* bias = nullptr;      // [N]
  const float* residual = nullptr;  // added after the activation; indexed like the output with its own strides
  long long ldr = 0, res_zs_hi = 0, res_zs_lo = 0;
  float alpha = 1.f;
  int act = kActNone;  // kActNone | kActGelu
  int round_out = 0;   // write results rounded to tf32 (for outputs that only feed further tf32 GEMMs)
};
// out[(z / zdiv) * out_zs_hi + (z % zdiv) * out_zs_lo + m * ldc + n], z < batch.
int gemm_tf32(const Tf32Operand& A, const Tf32Operand& B, int M, int N, int K, float* out, long long ldc, int batch,
              int zdiv, long long out_zs_hi, long long out_zs_lo, const Tf32Epilogue& ep, cudaStream_t stream);

// fp32 companions of the tf32 GEMM (transformer_ops.cu); every reduction has a fixed order.
// out[z][c][r] = in[z][r][c]; round_out: values rounded to tf32 on the way
int transpose_f32(const float* in, long long ld_in, long long zs_in, float* out, long long ld_out, long long zs_out, int R,
                  int C, int Z, int round_out, cudaStream_t s);
int round_tf32_f32(const float* in, float* out, long long n, cudaStream_t s);  // out = nearest tf32 of in (in == out allowed)
int layernorm_f32_forward(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                          int rows, int C, float eps, int round_out, cudaStream_t s);
long long layernorm_f32_backward_ws_floats(int rows, int C);
// dx = LayerNorm'(dy) + dskip (dskip may be null); dgamma / dbeta are WRITTEN
int layernorm_f32_backward(const float* x, const float* gamma, const float* mean, const float* rstd, const float* dy,
                           const float* dskip, float* dx, float* ws, float* dgamma, float* dbeta, int rows, int C,
                           cudaStream_t s);
// the softmax kernels write P / dS rounded to tf32 when round_out is set (they only feed GEMMs)
int softmax_f32_forward(float* x, long long rows, int cols, long long ld, float* lse, int round_out, cudaStream_t s);
// write_p = 0: X (the scores) is only read -- the caller does not need P afterwards
int softmax_f32_backward_rows(float* X, float* Y, long long rows, int cols, long long ld, const float* lse, float* delta,
                              int round_out, int write_p, cudaStream_t s);
int softmax_f32_backward_stats(float* X, float* Y, int Z, int R, int cols, long long ld, const float* lse,
                               const float* delta, int by_col, int round_out, cudaStream_t s);
int attn_delta_f32(const float* dO, const float* O, float* delta, int B, int L, int heads, int d, cudaStream_t s);
int gelu_f32_forward(const float* h, float* g, long long n, int round_out, cudaStream_t s);
int gelu_f32_backward(const float* h, float* dg, long long n, int round_out, cudaStream_t s);  // dg *= gelu'(h)
long long colsum_f32_ws_floats(long long rows, int cols);
int colsum_f32(const float* x, long long rows, int cols, long long ld, float* ws, float* out, cudaStream_t s);
int broadcast_f32(const float* src, long long n, float* out, int copies, cudaStream_t s);
int deconv_shuffle_f32(const float* in, float* out, int planes, int H, int W, int D, int inverse, cudaStream_t s);

// ---- normalisation / element-wise kernels (dense_ops.cu) ------------------------------------------------
// Size (floats) of the `stats` / `scratch2` buffers below: [N,groups,2] results followed by per-block partials
// (the reduction is two-stage and order-fixed, so statistics are bitwise reproducible).
long long groupnorm_workspace_floats(int N, int HW, int C, int groups);
// GroupNorm(32 groups) over NHWC fp16 with optional fused SiLU. stats: fp32 (sum, sumsq) per (n, group).
int groupnorm_forward(const __half* x, const __half* gamma, const __half* beta, __half* y, float* stats, int N, int HW,
                      int C, int groups, float eps, int silu, cudaStream_t s);
// dx for y = silu?(GN(x)): needs x, gamma, beta and the forward stats; scratch2: [N,groups,2] fp32.
int groupnorm_backward(const __half* x, const __half* gamma, const __half* beta, const float* stats, const __half* dy,
                       __half* dx, float* scratch2, int N, int HW, int C, int groups, float eps, int silu,
                       cudaStream_t s);
int layernorm_forward(const __half* x, const __half* gamma, const __half* beta, __half* y, int rows, int C, float eps,
                      cudaStream_t s);
// in-place softmax over the first `cols` entries of each row (row stride ld); entries [cols, ld) are zeroed.
int softmax_rows(__half* x, long long rows, int cols, long long ld, cudaStream_t s);
int geglu(const __half* xg, __half* y, long long rows, int inner, cudaStream_t s);  // xg [rows, 2*inner]
int silu_f32_to_f16(const float* x, __half* y, long long n, cudaStream_t s);
int upsample_nearest2x(const __half* x, __half* y, int N, int H, int W, int C, cudaStream_t s);
int concat_channels(const __half* a, int Ca, const __half* b, int Cb, __half* y, long long rows, cudaStream_t s);
int add_f16(const __half* a, const __half* b, __half* y, long long n, cudaStream_t s);
int transpose_f16(const __half* x, __half* y, int rows, int cols, cudaStream_t s);  // y[cols, rows]
// im2col for 3x3 stride-2 convs: pad_lo = 1 (UNet Downsample, padding=1) or 0 (VAE Downsample, pad (0,1,0,1)).
int im2col_3x3_s2(const __half* x, __half* col, int N, int H, int W, int C, int pad_lo, cudaStream_t s);
int col2im_3x3_s2(const __half* col, __half* dx, int N, int H, int W, int C, int pad_lo, cudaStream_t s);
// Direct (CUDA-core) 3x3 stride-1 pad-1 conv for tiny channel counts. w: [Cout, 3, 3, Cin] fp16, bias [Cout].
int conv3x3_small(const void* x, int x_fp32, const __half* w, const __half* bias, void* y, int y_fp32, int N, int H,
                  int W, int Cin, int Cout, cudaStream_t s);
int timestep_embedding(const float* t, __half* out, int n, int dim, float max_period, cudaStream_t s);
// y[rows, N] (fp32) = act?(x[rows,K]) * w[N,K]^T + bias: tiny-M linear on CUDA cores (time / camera embeddings).
int linear_small(const __half* x, const __half* w, const __half* bias, void* y, int y_fp32, int rows, int N, int K,
                 int silu_in, cudaStream_t s);

// dst[chunk*32 + j] = src[chunk*16 + j] (j < 16) | src[half_rows + chunk*16 + j - 16]: rows of `cols` halfs.
int interleave_geglu_rows(const __half* src, __half* dst, int half_rows, int cols, cudaStream_t s);
int rotate_w3x3(const __half* w, __half* wr, int Cout, int Cin, cudaStream_t s);  // [Cout,3,3,Cin] -> [Cin,3,3,Cout] flipped
int softmax_rows_backward(const __half* P, __half* dP, long long rows, int cols, long long ld, float scale,
                          cudaStream_t s);
int add_silu_f32_to_f16(const float* a, const float* b, __half* y, long long n, cudaStream_t s);  // silu(a + b?)

}  // namespace dense
