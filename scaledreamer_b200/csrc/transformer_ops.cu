// fp32 element-wise / normalisation / softmax kernels of the native Triplane-Transformer generator
// (custom/amortized/extern/triplane_transformer_modules.py:33-71 blocks, :115-187 generator; diffusers Attention with
// bias-free to_q / to_k / to_v and a biased to_out[0]). The contractions run in gemm_tf32_sm100.cu; everything here is
// HBM-bound and written for coalesced 16-byte accesses. Statistics and parameter gradients are reduced in a fixed order
// (two-stage, no atomics), so a step is bitwise reproducible.
#include <cstdio>

#include "common.cuh"
#include "dense.h"

namespace dense {
namespace {

// ---------------------------------------------------------------------------------------------- transpose
// out[z][c][r] = in[z][r][c]; 32 x 32 tiles through shared memory (+1 padding), both sides coalesced.
__global__ void __launch_bounds__(256)
transpose_f32_kernel(const float* __restrict__ in, long long ld_in, long long zs_in, float* __restrict__ out,
                     long long ld_out, long long zs_out, int R, int C, int round_out) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int z = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = in + (long long)z * zs_in;
  float* dst = out + (long long)z * zs_out;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = r0 + ty + 8 * j, c = c0 + tx;
    if (r < R && c < C) {
      const float v = src[(long long)r * ld_in + c];
      tile[ty + 8 * j][tx] = round_out ? round_tf32(v) : v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + ty + 8 * j, r = r0 + tx;
    if (r < R && c < C) dst[(long long)c * ld_out + r] = tile[tx][ty + 8 * j];
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm
constexpr int kLnMaxPerLane = 32;  // channels per lane: C <= 1024 (instantiated for 24 = the generator's 768, and 32)

// Warp per row: the row is read once into registers (C / 32 values per lane), mean / variance / apply from there.
template <int PER>
__global__ void __launch_bounds__(256)
layernorm_f32_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
                         int C, float eps, int round_out) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int per = C / 32;
  const float* xr = x + (long long)row * C;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i)
    if (i < per) {
      v[i] = xr[i * 32 + lane];
      s += v[i];
    }
  const float mean = warp_sum(s) / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i)
    if (i < per) {
      const float d = v[i] - mean;
      ss = fmaf(d, d, ss);
    }
  const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
  float* yr = y + (long long)row * C;
#pragma unroll
  for (int i = 0; i < PER; ++i)
    if (i < per) {
      const int c = i * 32 + lane;
      const float o = fmaf((v[i] - mean) * rstd, __ldg(gamma + c), __ldg(beta + c));
      yr[c] = round_out ? round_tf32(o) : o;
    }
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// dx = rstd * (dy*g - mean_c(dy*g) - xhat * mean_c(dy*g*xhat)) (+ dskip), warp per row like the forward.
// (A first version also accumulated d gamma / d beta per lane over a slab of rows: 175 registers, one CTA per SM, 277 us
// for 151 MB. Split in two, the row kernel streams at the forward's rate and the parameter sums are a column reduction.)
template <int PER>
__global__ void __launch_bounds__(256)
layernorm_f32_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean_in,
                            const float* __restrict__ rstd_in, const float* __restrict__ dy, const float* __restrict__ dskip,
                            float* __restrict__ dx, int rows, int C) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int per = C / 32;
  const float mean = mean_in[row], rstd = rstd_in[row];
  const float* xr = x + (long long)row * C;
  const float* dr = dy + (long long)row * C;
  const float* sk = dskip ? dskip + (long long)row * C : nullptr;
  float xh[PER], dz[PER], skv[PER];
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i)
    if (i < per) {
      const int c = i * 32 + lane;
      xh[i] = (xr[c] - mean) * rstd;
      dz[i] = dr[c] * __ldg(gamma + c);
      skv[i] = sk ? sk[c] : 0.f;
      s0 += dz[i];
      s1 = fmaf(dz[i], xh[i], s1);
    }
  s0 = warp_sum(s0) / (float)C;
  s1 = warp_sum(s1) / (float)C;
  float* o = dx + (long long)row * C;
#pragma unroll
  for (int i = 0; i < PER; ++i)
    if (i < per) o[i * 32 + lane] = fmaf(rstd, dz[i] - s0 - xh[i] * s1, skv[i]);
}

// partial[blk][0][c] = sum_r dy[r][c] * xhat[r][c], partial[blk][1][c] = sum_r dy[r][c] over the block's rows; 32 columns
// x 8 row lanes per CTA (every warp reads 128 contiguous bytes per operand), ln_param_finalize sums the blocks in order.
__global__ void __launch_bounds__(256)
ln_param_partial_kernel(const float* __restrict__ x, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                        const float* __restrict__ dy, float* __restrict__ partial, int rows, int C, int rows_per_block) {
  pdl_prologue();
  __shared__ float sm[2][8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float g0 = 0.f, g1 = 0.f, b0 = 0.f, b1 = 0.f;
  int r = r0 + ty;
  for (; r + 8 < r1; r += 16) {
    const float d0 = dy[(long long)r * C + c], d1 = dy[(long long)(r + 8) * C + c];
    const float x0 = (x[(long long)r * C + c] - mean_in[r]) * rstd_in[r];
    const float x1 = (x[(long long)(r + 8) * C + c] - mean_in[r + 8]) * rstd_in[r + 8];
    g0 = fmaf(d0, x0, g0), g1 = fmaf(d1, x1, g1);
    b0 += d0, b1 += d1;
  }
  if (r < r1) {
    const float d0 = dy[(long long)r * C + c];
    g0 = fmaf(d0, (x[(long long)r * C + c] - mean_in[r]) * rstd_in[r], g0);
    b0 += d0;
  }
  sm[0][ty][tx] = g0 + g1;
  sm[1][ty][tx] = b0 + b1;
  __syncthreads();
  if (ty < 2) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += sm[ty][w][tx];
    partial[((long long)blockIdx.y * 2 + ty) * C + c] = a;
  }
}

// out[k][c] = sum over blocks (in order) of partial[blk][k][c]
__global__ void ln_param_finalize_kernel(const float* __restrict__ partial, int nblk, int n, float* __restrict__ dgamma,
                                         float* __restrict__ dbeta, int C) {
  pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int b = 0;
  for (; b + 3 < nblk; b += 4) {
    a0 += partial[(long long)b * n + t];
    a1 += partial[(long long)(b + 1) * n + t];
    a2 += partial[(long long)(b + 2) * n + t];
    a3 += partial[(long long)(b + 3) * n + t];
  }
  for (; b < nblk; ++b) a0 += partial[(long long)b * n + t];
  const float a = (a0 + a1) + (a2 + a3);
  if (t < C) dgamma[t] = a;
  else dbeta[t - C] = a;
}

// ---------------------------------------------------------------------------------------------- softmax
// In-place softmax over the first `cols` entries of each row (row stride ld) with T threads per row (T = 32: one warp,
// T = 256: the block); lse[row] = max + log(sum exp(x - max)). Up to 16 * T columns stay in registers.
template <int T>
__global__ void __launch_bounds__(256)
softmax_f32_fwd_kernel(float* __restrict__ x, long long rows, int cols, long long ld, float* __restrict__ lse,
                       int round_out) {
  pdl_prologue();
  constexpr int kRowsPerBlock = 256 / T;
  __shared__ float red[8];
  const int t = threadIdx.x % T;
  const long long row = (long long)blockIdx.x * kRowsPerBlock + threadIdx.x / T;
  const bool on = row < rows;
  float* xr = x + (on ? row : 0) * ld;
  float v[16];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = i * T + t;
    v[i] = (on && c < cols) ? xr[c] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
  if (T == 256) {
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = __expf(v[i] - mx);  // exp(-inf) = 0 for the padding
    s += v[i];
  }
  s = warp_sum(s);
  if (T == 256) {
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
  }
  if (!on) return;
  const float inv = 1.f / s;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = i * T + t;
    if (c < cols) xr[c] = round_out ? round_tf32(v[i] * inv) : v[i] * inv;
  }
  if (t == 0) lse[row] = mx + __logf(s);
}

// Row form of the softmax backward: X holds the scores S of one row, Y holds dP. X <- P = exp(S - lse[row]),
// delta[row] = sum_k P dP, Y <- dS = P (dP - delta). delta is reduced from the SAME dP the row is corrected with, so
// sum_k dS = 0 holds to fp32 rounding (with delta taken from dO . O instead, the tf32 rounding of the two products
// differs and leaves a per-row offset that the key / query projections' gradients amplify tenfold).
template <int T>
__global__ void __launch_bounds__(256)
softmax_f32_bwd_rows_kernel(float* __restrict__ X, float* __restrict__ Y, long long rows, int cols, long long ld,
                            const float* __restrict__ lse, float* __restrict__ delta, int round_out, int write_p) {
  pdl_prologue();
  constexpr int kRowsPerBlock = 256 / T;
  __shared__ float red[8];
  const int t = threadIdx.x % T;
  const long long row = (long long)blockIdx.x * kRowsPerBlock + threadIdx.x / T;
  const bool on = row < rows;
  float* xr = X + (on ? row : 0) * ld;
  float* yr = Y + (on ? row : 0) * ld;
  const float l = on ? lse[row] : 0.f;
  float pv[16], dv[16];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = i * T + t;
    const bool in = on && c < cols;
    pv[i] = in ? __expf(xr[c] - l) : 0.f;
    dv[i] = in ? yr[c] : 0.f;
    s = fmaf(pv[i], dv[i], s);
  }
  s = warp_sum(s);
  if (T == 256) {
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
  }
  if (!on) return;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = i * T + t;
    if (c < cols) {
      const float ds = pv[i] * (dv[i] - s);
      if (write_p) xr[c] = round_out ? round_tf32(pv[i]) : pv[i];
      yr[c] = round_out ? round_tf32(ds) : ds;
    }
  }
  if (t == 0) delta[row] = s;
}

// X <- P = exp(X - lse[i]);  Y <- dS = P * (Y - delta[i]) over [Z][R][cols] (row stride ld), i = z*R + r (by_col = 0:
// X holds scores S) or i = z*cols + c (by_col = 1: X holds S^T, the statistics belong to the columns).
__global__ void __launch_bounds__(256)
softmax_f32_bwd_stats_kernel(float* __restrict__ X, float* __restrict__ Y, int R, int cols, long long ld,
                             const float* __restrict__ lse, const float* __restrict__ delta, int by_col, long long total4,
                             int round_out) {
  pdl_prologue();
  const int c4n = (cols + 3) >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / c4n;  // global row z*R + r
    const int c = (int)(i - row * c4n) * 4;
    float* xp = X + row * ld + c;
    float* yp = Y + row * ld + c;
    const long long z = row / R;
    if (c + 4 <= cols) {
      float4 xv = *reinterpret_cast<float4*>(xp), yv = *reinterpret_cast<float4*>(yp);
      float l[4], d[4];
      if (by_col) {
        const float4 lv = *reinterpret_cast<const float4*>(lse + z * cols + c);
        const float4 dv = *reinterpret_cast<const float4*>(delta + z * cols + c);
        l[0] = lv.x, l[1] = lv.y, l[2] = lv.z, l[3] = lv.w;
        d[0] = dv.x, d[1] = dv.y, d[2] = dv.z, d[3] = dv.w;
      } else {
        l[0] = l[1] = l[2] = l[3] = lse[row];
        d[0] = d[1] = d[2] = d[3] = delta[row];
      }
      xv.x = __expf(xv.x - l[0]), xv.y = __expf(xv.y - l[1]), xv.z = __expf(xv.z - l[2]), xv.w = __expf(xv.w - l[3]);
      yv.x = xv.x * (yv.x - d[0]), yv.y = xv.y * (yv.y - d[1]), yv.z = xv.z * (yv.z - d[2]), yv.w = xv.w * (yv.w - d[3]);
      if (round_out) {
        xv.x = round_tf32(xv.x), xv.y = round_tf32(xv.y), xv.z = round_tf32(xv.z), xv.w = round_tf32(xv.w);
        yv.x = round_tf32(yv.x), yv.y = round_tf32(yv.y), yv.z = round_tf32(yv.z), yv.w = round_tf32(yv.w);
      }
      *reinterpret_cast<float4*>(xp) = xv;
      *reinterpret_cast<float4*>(yp) = yv;
    } else {
      for (int j = 0; j < 4 && c + j < cols; ++j) {
        const long long si = by_col ? z * cols + c + j : row;
        const float pv = __expf(xp[j] - lse[si]);
        const float ds = pv * (yp[j] - delta[si]);
        xp[j] = round_out ? round_tf32(pv) : pv;
        yp[j] = round_out ? round_tf32(ds) : ds;
      }
    }
  }
}

// delta[(b*heads + h)*L + q] = sum_j dO[b, q, h*d + j] * O[b, q, h*d + j]  (= sum_k dP[q,k] P[q,k])
__global__ void attn_delta_f32_kernel(const float* __restrict__ dO, const float* __restrict__ O, float* __restrict__ delta,
                                      int B, int L, int heads, int d) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (b, q, h) with h fastest: coalesced-ish reads
  if (i >= (long long)B * L * heads) return;
  const int h = (int)(i % heads);
  const long long bq = i / heads;
  const int q = (int)(bq % L), b = (int)(bq / L);
  const float* a = dO + bq * (long long)heads * d + (long long)h * d;
  const float* o = O + bq * (long long)heads * d + (long long)h * d;
  float s = 0.f;
  for (int j = 0; j < d; j += 4) {
    const float4 u = *reinterpret_cast<const float4*>(a + j), w = *reinterpret_cast<const float4*>(o + j);
    s = fmaf(u.x, w.x, fmaf(u.y, w.y, fmaf(u.z, w.z, fmaf(u.w, w.w, s))));
  }
  delta[((long long)b * heads + h) * L + q] = s;
}

// ---------------------------------------------------------------------------------------------- GELU (erf form)
__global__ void gelu_f32_fwd_kernel(const float* __restrict__ h, float* __restrict__ g, long long n4, int round_out) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(h)[i];
    v.x = 0.5f * v.x * (1.f + erff(v.x * 0.70710678118654752f));
    v.y = 0.5f * v.y * (1.f + erff(v.y * 0.70710678118654752f));
    v.z = 0.5f * v.z * (1.f + erff(v.z * 0.70710678118654752f));
    v.w = 0.5f * v.w * (1.f + erff(v.w * 0.70710678118654752f));
    if (round_out) v.x = round_tf32(v.x), v.y = round_tf32(v.y), v.z = round_tf32(v.z), v.w = round_tf32(v.w);
    reinterpret_cast<float4*>(g)[i] = v;
  }
}
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__global__ void gelu_f32_bwd_kernel(const float* __restrict__ h, float* __restrict__ dg, long long n4, int round_out) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(h)[i];
    float4 d = reinterpret_cast<float4*>(dg)[i];
    d.x *= gelu_grad(v.x), d.y *= gelu_grad(v.y), d.z *= gelu_grad(v.z), d.w *= gelu_grad(v.w);
    if (round_out) d.x = round_tf32(d.x), d.y = round_tf32(d.y), d.z = round_tf32(d.z), d.w = round_tf32(d.w);
    reinterpret_cast<float4*>(dg)[i] = d;
  }
}

// ---------------------------------------------------------------------------------------------- column sums
// partial[blk][c] = sum of rows [blk*rpb, (blk+1)*rpb) of x[:, c]; finalize sums the blocks in order. 32 columns x 8
// row lanes per CTA: every warp reads 128 contiguous bytes.
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ x, long long rows, int cols, long long ld, int rows_per_block,
                  float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float sm[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f;
  if (c < cols) {
    long long r = r0 + ty;
    for (; r + 8 < r1; r += 16) {
      a0 += x[r * ld + c];
      a1 += x[(r + 8) * ld + c];
    }
    if (r < r1) a0 += x[r * ld + c];
  }
  sm[ty][tx] = a0 + a1;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += sm[w][tx];
    partial[(long long)blockIdx.y * cols + c] = a;
  }
}
// few rows (the slices of a split weight-gradient product, the prompts of the position-embedding gradient): one thread
// per four columns sums the rows in order
__global__ void colsum_small_kernel(const float* __restrict__ x, int rows, long long cols4, long long ld4,
                                    float* __restrict__ out) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cols4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<const float4*>(x)[i];
    for (int r = 1; r < rows; ++r) {
      const float4 v = reinterpret_cast<const float4*>(x)[(long long)r * ld4 + i];
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
  }
}
__global__ void colsum_finalize_kernel(const float* __restrict__ partial, int nblk, int cols, float* __restrict__ out) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float a = 0.f;
  for (int b = 0; b < nblk; ++b) a += partial[(long long)b * cols + c];
  out[c] = a;
}

__global__ void round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  pdl_prologue();
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(in)[i];
    v.x = round_tf32(v.x), v.y = round_tf32(v.y), v.z = round_tf32(v.z), v.w = round_tf32(v.w);
    reinterpret_cast<float4*>(out)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[n4 * 4 + threadIdx.x] = round_tf32(in[n4 * 4 + threadIdx.x]);
}

// out[k][i] = src[i] for k < copies (the position embedding repeated over the prompts of the batch)
__global__ void broadcast_f32_kernel(const float* __restrict__ src, long long n4, float* __restrict__ out, int copies) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    for (int k = 0; k < copies; ++k) reinterpret_cast<float4*>(out)[(long long)k * n4 + i] = v;
  }
}

// ConvTranspose2d(kernel 2, stride 2) as a GEMM leaves t[(plane, h, w)][d*4 + a*2 + b]; the channels-last plane the
// triplane sampler reads is p[plane][2h + a][2w + b][d]. inverse = 1 maps a plane gradient back to the GEMM layout.
__global__ void deconv_shuffle_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int H, int W,
                                          int D, int inverse) {
  pdl_prologue();
  const long long total = (long long)planes * H * W * D * 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes the plane layout [plane][2H][2W][D]
    const int d = (int)(i % D);
    long long r = i / D;
    const int xw = (int)(r % (2 * W));
    r /= 2 * W;
    const int yh = (int)(r % (2 * H));
    const int pl = (int)(r / (2 * H));
    const long long tok = ((long long)pl * H + (yh >> 1)) * W + (xw >> 1);
    const long long j = tok * (D * 4) + d * 4 + (yh & 1) * 2 + (xw & 1);
    if (inverse) out[j] = in[i];
    else out[i] = in[j];
  }
}

int ew_blocks(long long n) {
  const long long b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > kNumSMs * 16 ? kNumSMs * 16 : b));
}

}  // namespace

int transpose_f32(const float* in, long long ld_in, long long zs_in, float* out, long long ld_out, long long zs_out, int R,
                  int C, int Z, int round_out, cudaStream_t s) {
  if (R <= 0 || C <= 0 || Z <= 0 || Z > 65535) {
    sdb_set_error("transpose_f32: bad shape R=%d C=%d Z=%d", R, C, Z);
    return SDB_ERR_ARG;
  }
  sdb_launch(transpose_f32_kernel, dim3((C + 31) / 32, (R + 31) / 32, Z), dim3(256), 0, s, in, ld_in, zs_in, out, ld_out,
             zs_out, R, C, round_out);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("transpose_f32");
  return SDB_OK;
}

int layernorm_f32_forward(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                          int rows, int C, float eps, int round_out, cudaStream_t s) {
  if (C % 32 || C > 32 * kLnMaxPerLane) {
    sdb_set_error("layernorm_f32: C=%d must be a multiple of 32, at most %d", C, 32 * kLnMaxPerLane);
    return SDB_ERR_UNSUPPORTED;
  }
  if (C <= 768)
    sdb_launch(layernorm_f32_fwd_kernel<24>, dim3((rows + 7) / 8), dim3(256), 0, s, x, gamma, beta, y, mean, rstd, rows, C, eps, round_out);
  else
    sdb_launch(layernorm_f32_fwd_kernel<32>, dim3((rows + 7) / 8), dim3(256), 0, s, x, gamma, beta, y, mean, rstd, rows, C, eps, round_out);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("layernorm_f32_fwd");
  return SDB_OK;
}

static int ln_bwd_blocks(int rows, int* rpb) {
  int r = (rows + 63) / 64;
  if (r < 16) r = 16;
  *rpb = r;
  return (rows + r - 1) / r;
}
long long layernorm_f32_backward_ws_floats(int rows, int C) {
  int rpb;
  return (long long)ln_bwd_blocks(rows, &rpb) * 2 * C;
}
int layernorm_f32_backward(const float* x, const float* gamma, const float* mean, const float* rstd, const float* dy,
                           const float* dskip, float* dx, float* ws, float* dgamma, float* dbeta, int rows, int C,
                           cudaStream_t s) {
  if (C % 32 || C > 32 * kLnMaxPerLane) {
    sdb_set_error("layernorm_f32: C=%d must be a multiple of 32, at most %d", C, 32 * kLnMaxPerLane);
    return SDB_ERR_UNSUPPORTED;
  }
  if (C <= 768)
    sdb_launch(layernorm_f32_bwd_dx_kernel<24>, dim3((rows + 7) / 8), dim3(256), 0, s, x, gamma, mean, rstd, dy, dskip, dx,
               rows, C);
  else
    sdb_launch(layernorm_f32_bwd_dx_kernel<32>, dim3((rows + 7) / 8), dim3(256), 0, s, x, gamma, mean, rstd, dy, dskip, dx,
               rows, C);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("layernorm_f32_bwd_dx");
  int rpb;
  const int nblk = ln_bwd_blocks(rows, &rpb);
  sdb_launch(ln_param_partial_kernel, dim3(C / 32, nblk), dim3(256), 0, s, x, mean, rstd, dy, ws, rows, C, rpb);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("ln_param_partial");
  sdb_launch(ln_param_finalize_kernel, dim3((2 * C + 255) / 256), dim3(256), 0, s, (const float*)ws, nblk, 2 * C, dgamma,
             dbeta, C);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("ln_param_finalize");
  return SDB_OK;
}

int softmax_f32_forward(float* x, long long rows, int cols, long long ld, float* lse, int round_out, cudaStream_t s) {
  if (cols <= 0 || cols > 4096) {
    sdb_set_error("softmax_f32: cols=%d must be in [1, 4096]", cols);
    return SDB_ERR_UNSUPPORTED;
  }
  if (cols <= 512) {
    sdb_launch(softmax_f32_fwd_kernel<32>, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, s, x, rows, cols, ld, lse, round_out);
  } else {
    sdb_launch(softmax_f32_fwd_kernel<256>, dim3((unsigned)rows), dim3(256), 0, s, x, rows, cols, ld, lse, round_out);
  }
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("softmax_f32_fwd");
  return SDB_OK;
}

int softmax_f32_backward_rows(float* X, float* Y, long long rows, int cols, long long ld, const float* lse, float* delta,
                              int round_out, int write_p, cudaStream_t s) {
  if (cols <= 0 || cols > 4096) {
    sdb_set_error("softmax_f32: cols=%d must be in [1, 4096]", cols);
    return SDB_ERR_UNSUPPORTED;
  }
  if (cols <= 512)
    sdb_launch(softmax_f32_bwd_rows_kernel<32>, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, s, X, Y, rows, cols, ld, lse,
               delta, round_out, write_p);
  else
    sdb_launch(softmax_f32_bwd_rows_kernel<256>, dim3((unsigned)rows), dim3(256), 0, s, X, Y, rows, cols, ld, lse, delta,
               round_out, write_p);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("softmax_f32_bwd_rows");
  return SDB_OK;
}

int softmax_f32_backward_stats(float* X, float* Y, int Z, int R, int cols, long long ld, const float* lse,
                               const float* delta, int by_col, int round_out, cudaStream_t s) {
  if ((ld & 3) || (by_col && (cols & 3))) {
    sdb_set_error("softmax_f32_backward_stats: ld=%lld (and cols=%d when the statistics run along columns) must be multiples of 4",
                  ld, cols);
    return SDB_ERR_ARG;
  }
  const long long total4 = (long long)Z * R * ((cols + 3) >> 2);
  sdb_launch(softmax_f32_bwd_stats_kernel, dim3(ew_blocks(total4)), dim3(256), 0, s, X, Y, R, cols, ld, lse, delta, by_col,
             total4, round_out);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("softmax_f32_bwd_stats");
  return SDB_OK;
}

int attn_delta_f32(const float* dO, const float* O, float* delta, int B, int L, int heads, int d, cudaStream_t s) {
  if (d % 4) {
    sdb_set_error("attn_delta_f32: head_dim=%d must be a multiple of 4", d);
    return SDB_ERR_ARG;
  }
  const long long n = (long long)B * L * heads;
  sdb_launch(attn_delta_f32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, dO, O, delta, B, L, heads, d);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("attn_delta_f32");
  return SDB_OK;
}

int gelu_f32_forward(const float* h, float* g, long long n, int round_out, cudaStream_t s) {
  if (n % 4) {
    sdb_set_error("gelu_f32: n=%lld must be a multiple of 4", n);
    return SDB_ERR_ARG;
  }
  sdb_launch(gelu_f32_fwd_kernel, dim3(ew_blocks(n / 4)), dim3(256), 0, s, h, g, n / 4, round_out);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("gelu_f32_fwd");
  return SDB_OK;
}
int gelu_f32_backward(const float* h, float* dg, long long n, int round_out, cudaStream_t s) {
  if (n % 4) {
    sdb_set_error("gelu_f32: n=%lld must be a multiple of 4", n);
    return SDB_ERR_ARG;
  }
  sdb_launch(gelu_f32_bwd_kernel, dim3(ew_blocks(n / 4)), dim3(256), 0, s, h, dg, n / 4, round_out);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("gelu_f32_bwd");
  return SDB_OK;
}

static int colsum_blocks(long long rows, int* rpb) {
  long long r = (rows + 63) / 64;
  if (r < 16) r = 16;
  *rpb = (int)r;
  return (int)((rows + r - 1) / r);
}
long long colsum_f32_ws_floats(long long rows, int cols) {
  int rpb;
  return (long long)colsum_blocks(rows, &rpb) * cols;
}
int colsum_f32(const float* x, long long rows, int cols, long long ld, float* ws, float* out, cudaStream_t s) {
  if (rows <= 16 && (cols & 3) == 0 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    sdb_launch(colsum_small_kernel, dim3(ew_blocks(cols / 4)), dim3(256), 0, s, x, (int)rows, (long long)(cols / 4), ld / 4,
               out);
    SDB_COUNT_LAUNCH();
    SDB_CHECK_LAUNCH("colsum_small");
    return SDB_OK;
  }
  int rpb;
  const int nblk = colsum_blocks(rows, &rpb);
  sdb_launch(colsum_f32_kernel, dim3((cols + 31) / 32, nblk), dim3(256), 0, s, x, rows, cols, ld, rpb, ws);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("colsum_f32");
  sdb_launch(colsum_finalize_kernel, dim3((cols + 255) / 256), dim3(256), 0, s, (const float*)ws, nblk, cols, out);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("colsum_finalize");
  return SDB_OK;
}

int round_tf32_f32(const float* in, float* out, long long n, cudaStream_t s) {
  sdb_launch(round_tf32_kernel, dim3(ew_blocks((n + 3) / 4)), dim3(256), 0, s, in, out, n);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("round_tf32");
  return SDB_OK;
}

int broadcast_f32(const float* src, long long n, float* out, int copies, cudaStream_t s) {
  if (n % 4) {
    sdb_set_error("broadcast_f32: n=%lld must be a multiple of 4", n);
    return SDB_ERR_ARG;
  }
  sdb_launch(broadcast_f32_kernel, dim3(ew_blocks(n / 4)), dim3(256), 0, s, src, n / 4, out, copies);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("broadcast_f32");
  return SDB_OK;
}

int deconv_shuffle_f32(const float* in, float* out, int planes, int H, int W, int D, int inverse, cudaStream_t s) {
  const long long total = (long long)planes * H * W * D * 4;
  sdb_launch(deconv_shuffle_f32_kernel, dim3(ew_blocks(total)), dim3(256), 0, s, in, out, planes, H, W, D, inverse);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("deconv_shuffle_f32");
  return SDB_OK;
}

}  // namespace dense
