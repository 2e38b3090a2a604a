// LinearHyperNetwork of the amortized generators (custom/amortized/models/geometry/hyper_iNGP.py:18-111; the same class
// emits the environment-map weights, custom/.../multiprompt_neural_environment_hashgrid_map_background.py:82-99):
//     out[b] = W1 silu(LayerNorm(W0 x[b])) + b1,   W0 [64, c_dim] without bias, W1 [n_out, 64]
// for a handful of prompts per step (B <= 64). 0.3 MFLOP per prompt: the point of a native version is one launch
// forward and two backward instead of six eager launches and four cuBLAS GEMVs, and fixed summation orders.
#include "../../include/sdb200.h"
#include "common.cuh"

namespace {

constexpr int kHn = 64;       // n_neurons
constexpr int kHnMaxB = 64;   // prompts per call

__device__ __forceinline__ float hn_silu(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float hn_silu_grad(float x) {
  const float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}

// LayerNorm statistics of one 64-vector held as 2 values per lane of a warp
__device__ __forceinline__ void ln_stats(float v0, float v1, float eps, float* mean, float* rstd) {
  const float m = warp_sum(v0 + v1) * (1.f / kHn);
  const float d0 = v0 - m, d1 = v1 - m;
  const float var = warp_sum(d0 * d0 + d1 * d1) * (1.f / kHn);  // biased variance, as torch.nn.LayerNorm
  *mean = m;
  *rstd = rsqrtf(var + eps);
}

__global__ void __launch_bounds__(256)
hypernet_fwd_kernel(const float* __restrict__ x, int c_dim, const float* __restrict__ w0, const float* __restrict__ ln_g,
                    const float* __restrict__ ln_b, float eps, const float* __restrict__ w1,
                    const float* __restrict__ b1, int n_out, float* __restrict__ h_out, float* __restrict__ out) {
  extern __shared__ float xs[];  // [c_dim]
  __shared__ float hs[kHn], as[kHn];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < c_dim; i += blockDim.x) xs[i] = x[(size_t)b * c_dim + i];
  __syncthreads();
  for (int j = warp; j < kHn; j += 8) {
    const float* wr = w0 + (size_t)j * c_dim;
    float s = 0.f;
    for (int i = lane; i < c_dim; i += 32) s = fmaf(__ldg(wr + i), xs[i], s);
    s = warp_sum(s);
    if (lane == 0) {
      hs[j] = s;
      h_out[b * kHn + j] = s;
    }
  }
  __syncthreads();
  if (warp == 0) {
    float mean, rstd;
    const float v0 = hs[lane], v1 = hs[lane + 32];
    ln_stats(v0, v1, eps, &mean, &rstd);
    as[lane] = hn_silu(fmaf((v0 - mean) * rstd, ln_g[lane], ln_b[lane]));
    as[lane + 32] = hn_silu(fmaf((v1 - mean) * rstd, ln_g[lane + 32], ln_b[lane + 32]));
  }
  __syncthreads();
  for (int o = tid; o < n_out; o += blockDim.x) {
    const float4* wr = reinterpret_cast<const float4*>(w1 + (size_t)o * kHn);
    float s = b1 ? b1[o] : 0.f;
#pragma unroll
    for (int q = 0; q < kHn / 4; ++q) {
      const float4 w = __ldg(wr + q);
      s = fmaf(w.x, as[4 * q], fmaf(w.y, as[4 * q + 1], fmaf(w.z, as[4 * q + 2], fmaf(w.w, as[4 * q + 3], s))));
    }
    out[(size_t)b * n_out + o] = s;
  }
}

// One block per 64 output rows of W1: dW1, db1 and this block's share of d a (fixed-order partials in `scratch`).
__global__ void __launch_bounds__(256)
hypernet_bwd_out_kernel(int B, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps,
                        const float* __restrict__ w1, int n_out, const float* __restrict__ h,
                        const float* __restrict__ d_out, float* __restrict__ g_w1, float* __restrict__ g_b1,
                        float* __restrict__ scratch) {
  __shared__ float as[kHnMaxB][kHn];
  __shared__ float ds[kHnMaxB][kHn];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int o0 = blockIdx.x * kHn;
  for (int b = warp; b < B; b += 8) {
    float mean, rstd;
    const float v0 = h[b * kHn + lane], v1 = h[b * kHn + lane + 32];
    ln_stats(v0, v1, eps, &mean, &rstd);
    as[b][lane] = hn_silu(fmaf((v0 - mean) * rstd, ln_g[lane], ln_b[lane]));
    as[b][lane + 32] = hn_silu(fmaf((v1 - mean) * rstd, ln_g[lane + 32], ln_b[lane + 32]));
  }
  for (int i = tid; i < B * kHn; i += blockDim.x) {
    const int b = i / kHn, oo = i % kHn;
    ds[b][oo] = o0 + oo < n_out ? d_out[(size_t)b * n_out + o0 + oo] : 0.f;
  }
  __syncthreads();
  const int k = tid & 63, q = tid >> 6;
  for (int oo = q; oo < kHn; oo += 4) {
    if (o0 + oo >= n_out) break;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(ds[b][oo], as[b][k], s);
    g_w1[(size_t)(o0 + oo) * kHn + k] = s;
  }
  if (tid < kHn && o0 + tid < n_out && g_b1) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += ds[b][tid];
    g_b1[o0 + tid] = s;
  }
  for (int b = q; b < B; b += 4) {
    float s = 0.f;
    for (int oo = 0; oo < kHn; ++oo)
      if (o0 + oo < n_out) s = fmaf(__ldg(w1 + (size_t)(o0 + oo) * kHn + k), ds[b][oo], s);
    scratch[((size_t)blockIdx.x * B + b) * kHn + k] = s;
  }
}

// One block per hidden unit j: d a (summed over the partials in block order) -> SiLU / LayerNorm backward -> dW0[j, :];
// block 0 also writes the LayerNorm parameter gradients.
__global__ void __launch_bounds__(256)
hypernet_bwd_in_kernel(const float* __restrict__ x, int B, int c_dim, const float* __restrict__ ln_g,
                       const float* __restrict__ ln_b, float eps, const float* __restrict__ h,
                       const float* __restrict__ scratch, int nblk, float* __restrict__ g_w0, float* __restrict__ g_ln_g,
                       float* __restrict__ g_ln_b) {
  __shared__ float da[kHnMaxB][kHn];
  __shared__ float dhj[kHnMaxB];
  __shared__ float dg[8][kHn], db[8][kHn];
  const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < B * kHn; i += blockDim.x) {
    const int b = i / kHn, k = i % kHn;
    float s = 0.f;
    for (int blk = 0; blk < nblk; ++blk) s += scratch[((size_t)blk * B + b) * kHn + k];
    da[b][k] = s;
  }
  for (int i = tid; i < 8 * kHn; i += blockDim.x) dg[i / kHn][i % kHn] = db[i / kHn][i % kHn] = 0.f;
  __syncthreads();
  for (int b = warp; b < B; b += 8) {  // warp per prompt: lanes hold hidden units lane, lane + 32
    float mean, rstd;
    const float v0 = h[b * kHn + lane], v1 = h[b * kHn + lane + 32];
    ln_stats(v0, v1, eps, &mean, &rstd);
    const float z0 = (v0 - mean) * rstd, z1 = (v1 - mean) * rstd;
    const float g0 = ln_g[lane], g1 = ln_g[lane + 32];
    const float dy0 = da[b][lane] * hn_silu_grad(fmaf(z0, g0, ln_b[lane]));
    const float dy1 = da[b][lane + 32] * hn_silu_grad(fmaf(z1, g1, ln_b[lane + 32]));
    const float dz0 = dy0 * g0, dz1 = dy1 * g1;
    const float m1 = warp_sum(dz0 + dz1) * (1.f / kHn);
    const float m2 = warp_sum(dz0 * z0 + dz1 * z1) * (1.f / kHn);
    const float dh0 = rstd * (dz0 - m1 - z0 * m2), dh1 = rstd * (dz1 - m1 - z1 * m2);
    if (j == lane) dhj[b] = dh0;
    if (j == lane + 32) dhj[b] = dh1;
    dg[warp][lane] += dy0 * z0;
    dg[warp][lane + 32] += dy1 * z1;
    db[warp][lane] += dy0;
    db[warp][lane + 32] += dy1;
  }
  __syncthreads();
  if (j == 0 && tid < kHn) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s0 += dg[w][tid], s1 += db[w][tid];
    g_ln_g[tid] = s0;
    g_ln_b[tid] = s1;
  }
  for (int i = tid; i < c_dim; i += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dhj[b], __ldg(x + (size_t)b * c_dim + i), s);
    g_w0[(size_t)j * c_dim + i] = s;
  }
}

}  // namespace

extern "C" {

long long sdb_hypernet_scratch_floats(int n_prompts, int n_out) {
  return (long long)((n_out + kHn - 1) / kHn) * n_prompts * kHn;
}

int sdb_hypernet_forward(const float* x, int n_prompts, int c_dim, const float* w0, const float* ln_weight,
                         const float* ln_bias, float ln_eps, const float* w1, const float* b1, int n_out,
                         float* hidden, float* out, void* stream) {
  SDB_CHECK_ARG(x && w0 && ln_weight && ln_bias && w1 && hidden && out, "hypernet_forward: NULL buffer");
  SDB_CHECK_ARG(n_prompts >= 0 && n_prompts <= kHnMaxB && c_dim > 0 && n_out > 0,
                "hypernet_forward: 0 <= n_prompts <= 64, c_dim > 0, n_out > 0");
  SDB_CHECK_ARG((size_t)c_dim * 4 <= 48 * 1024, "hypernet_forward: c_dim must be <= 12288");
  if (n_prompts == 0) return SDB_OK;
  hypernet_fwd_kernel<<<n_prompts, 256, (size_t)c_dim * 4, (cudaStream_t)stream>>>(x, c_dim, w0, ln_weight, ln_bias, ln_eps,
                                                                                  w1, b1, n_out, hidden, out);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("hypernet_fwd");
  return SDB_OK;
}

int sdb_hypernet_backward(const float* x, int n_prompts, int c_dim, const float* ln_weight, const float* ln_bias,
                          float ln_eps, const float* w1, int n_out, const float* hidden, const float* d_out,
                          float* g_w0, float* g_ln_weight, float* g_ln_bias, float* g_w1, float* g_b1, float* scratch,
                          void* stream) {
  SDB_CHECK_ARG(x && ln_weight && ln_bias && w1 && hidden && d_out && g_w0 && g_ln_weight && g_ln_bias && g_w1 && scratch,
                "hypernet_backward: NULL buffer");
  SDB_CHECK_ARG(n_prompts > 0 && n_prompts <= kHnMaxB && c_dim > 0 && n_out > 0,
                "hypernet_backward: 0 < n_prompts <= 64, c_dim > 0, n_out > 0");
  const int nblk = (n_out + kHn - 1) / kHn;
  hypernet_bwd_out_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(n_prompts, ln_weight, ln_bias, ln_eps, w1, n_out, hidden,
                                                                 d_out, g_w1, g_b1, scratch);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("hypernet_bwd_out");
  hypernet_bwd_in_kernel<<<kHn, 256, 0, (cudaStream_t)stream>>>(x, n_prompts, c_dim, ln_weight, ln_bias, ln_eps, hidden,
                                                               scratch, nblk, g_w0, g_ln_weight, g_ln_bias);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("hypernet_bwd_in");
  return SDB_OK;
}

}  // extern "C"
