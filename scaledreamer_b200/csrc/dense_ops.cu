// Normalisation and element-wise kernels of the dense path (HBM-bound; fp16 storage, fp32 math).
// Reference semantics: GroupNorm32 / Normalize (extern/mvdream/ldm/modules/diffusionmodules/util.py:229-231,
// model.py:46-47, attention.py:88-89), nn.LayerNorm + GEGLU (attention.py:49-57,271-273), softmax
// (attention.py:186), nearest Upsample (openaimodel.py:109-118), timestep_embedding (util.py:165-186).
#include <cstdlib>

#include "dense.h"

namespace dense {
namespace {

__device__ __forceinline__ float silu(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_grad(float x) {
  const float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// ---------------------------------------------------------------------------------------------- GroupNorm
// Thread mapping shared by the statistics and apply kernels: a block is PL pixel lanes x C8N channel octets; thread
// (pl, c8) owns channels [8 c8, 8 c8 + 8) of pixels pl, pl + stride, ... of image blockIdx.y. The channel octet never
// changes, so every per-channel constant (mean, rstd, gamma, beta) is computed once and lives in registers, and each
// access is one 16-byte load / store, fully coalesced across the block (NHWC).
struct GnChan {  // per-thread constants of the 4 channel pairs of an octet
  float mean[4], rstd[4];
};
__device__ __forceinline__ GnChan gn_chan(const float* __restrict__ stats, int n, int groups, int cpg, int c8, float cnt,
                                          float eps) {
  GnChan k;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int g = (c8 * 8 + 2 * e) / cpg;
    const float s = stats[(n * groups + g) * 2], ss = stats[(n * groups + g) * 2 + 1];
    k.mean[e] = s / cnt;
    k.rstd[e] = rsqrtf(fmaxf(ss / cnt - k.mean[e] * k.mean[e], 0.f) + eps);
  }
  return k;
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __half22float2(h[e]);
    f[2 * e] = t.x, f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
  return v;
}

// mode 0: per-pair (sum x, sum x^2); mode 1 (backward): (sum dxhat, sum dxhat*xhat). Each block writes one partial per
// group; gn_finalize_kernel sums them in a fixed order (bitwise reproducible statistics).
template <int MODE, int UN>
__global__ void __launch_bounds__(512)
gn_reduce_kernel(const __half* __restrict__ x, const __half* __restrict__ dy, const __half* __restrict__ gamma,
                 const __half* __restrict__ beta, const float* __restrict__ stats, float* __restrict__ out, int HW,
                 int C, int groups, int C8N, int PL, int pix_per_block, float eps, int act_silu) {
  pdl_prologue();
  extern __shared__ float sm[];  // [PL][C/2][2]
  const int n = blockIdx.y;
  const int c8 = threadIdx.x % C8N, pl = threadIdx.x / C8N;
  const int cpg = C / groups;
  const int p_begin = blockIdx.x * pix_per_block, p_end = min(HW, p_begin + pix_per_block);
  float g[8], b[8];
  GnChan k;
  if (MODE == 1) {
    k = gn_chan(stats, n, groups, cpg, c8, (float)HW * (float)cpg, eps);
    unpack8(reinterpret_cast<const uint4*>(gamma)[c8], g);
    unpack8(reinterpret_cast<const uint4*>(beta)[c8], b);
  }
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  const uint4* xp = reinterpret_cast<const uint4*>(x) + (size_t)n * HW * C8N + c8;
  const uint4* dp = reinterpret_cast<const uint4*>(dy) + (size_t)n * HW * C8N + c8;
  auto accumulate = [&](const uint4& xv, const uint4& dv) {
    float xf[8];
    unpack8(xv, xf);
    if (MODE == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s0[e] += xf[2 * e] + xf[2 * e + 1];
        s1[e] = fmaf(xf[2 * e], xf[2 * e], fmaf(xf[2 * e + 1], xf[2 * e + 1], s1[e]));
      }
    } else {
      float df[8];
      unpack8(dv, df);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = j >> 1;
        const float xh = (xf[j] - k.mean[e]) * k.rstd[e];
        float dz = df[j];
        if (act_silu) dz *= silu_grad(fmaf(xh, g[j], b[j]));
        const float dxh = dz * g[j];
        s0[e] += dxh;
        s1[e] = fmaf(dxh, xh, s1[e]);
      }
    }
  };
  int p = p_begin + pl;
  for (; p + (UN - 1) * PL < p_end; p += UN * PL) {  // UN independent 16-byte loads in flight per operand
    uint4 xv[UN], dv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      xv[u] = xp[(size_t)(p + u * PL) * C8N];
      if (MODE == 1) dv[u] = dp[(size_t)(p + u * PL) * C8N];
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) accumulate(xv[u], dv[u]);
  }
  for (; p < p_end; p += PL) {
    uint4 dv = make_uint4(0, 0, 0, 0);
    if (MODE == 1) dv = dp[(size_t)p * C8N];
    accumulate(xp[(size_t)p * C8N], dv);
  }
  const int n_pairs = C / 2;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    sm[(pl * n_pairs + c8 * 4 + e) * 2] = s0[e];
    sm[(pl * n_pairs + c8 * 4 + e) * 2 + 1] = s1[e];
  }
  __syncthreads();
  // thread (gi, k) sums statistic k of group gi over pixel lanes and the group's pairs, in a fixed order
  const int ppg = cpg / 2;
  for (int t = threadIdx.x; t < groups * 2; t += blockDim.x) {
    const int gi = t >> 1, kk = t & 1;
    float acc = 0.f;
    for (int yy = 0; yy < PL; ++yy)
      for (int q = gi * ppg; q < (gi + 1) * ppg; ++q) acc += sm[(yy * n_pairs + q) * 2 + kk];
    out[(((size_t)n * gridDim.x + blockIdx.x) * groups + gi) * 2 + kk] = acc;
  }
}

// One warp per (n, group, k): lanes stride over the per-block partials, then a butterfly sum -- a fixed order, so the
// statistics stay bitwise reproducible run to run.
__global__ void gn_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int nblk, int groups,
                                   int total) {
  pdl_prologue();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // (n, g, k)
  const int lane = threadIdx.x & 31;
  if (i >= total) return;
  const int k = i & 1, g = (i >> 1) % groups, n = (i >> 1) / groups;
  float a4[4] = {0.f, 0.f, 0.f, 0.f};  // four independent chains: the loads are latency-bound, the order stays fixed
  int b = lane;
  for (; b + 96 < nblk; b += 128) {
#pragma unroll
    for (int u = 0; u < 4; ++u) a4[u] += partial[(((size_t)n * nblk + b + 32 * u) * groups + g) * 2 + k];
  }
  for (; b < nblk; b += 32) a4[0] += partial[(((size_t)n * nblk + b) * groups + g) * 2 + k];
  float acc = warp_sum((a4[0] + a4[1]) + (a4[2] + a4[3]));
  if (lane == 0) out[i] = acc;
}

// y = act(GN(x)) (MODE 0) or dx of it (MODE 1), same thread mapping as gn_reduce_kernel.
template <int MODE, int UN>
__global__ void __launch_bounds__(512)
gn_apply_kernel(const __half* __restrict__ x, const __half* __restrict__ dy, const __half* __restrict__ gamma,
                const __half* __restrict__ beta, const float* __restrict__ stats, const float* __restrict__ red,
                __half* __restrict__ y, int HW, int C, int groups, int C8N, int PL, float eps, int act_silu) {
  pdl_prologue();
  const int n = blockIdx.y;
  const int c8 = threadIdx.x % C8N, pl = threadIdx.x / C8N;
  const int cpg = C / groups;
  const float cnt = (float)HW * (float)cpg;
  const GnChan k = gn_chan(stats, n, groups, cpg, c8, cnt, eps);
  float g[8], b[8], r0[4], r1[4];
  unpack8(reinterpret_cast<const uint4*>(gamma)[c8], g);
  unpack8(reinterpret_cast<const uint4*>(beta)[c8], b);
  if (MODE == 0) {  // fold the normalisation into one fma per element: y = x * g' + b'
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sc = k.rstd[j >> 1] * g[j];
      b[j] = fmaf(-k.mean[j >> 1], sc, b[j]);
      g[j] = sc;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gi = (c8 * 8 + 2 * e) / cpg;
      r0[e] = red[(n * groups + gi) * 2] / cnt;
      r1[e] = red[(n * groups + gi) * 2 + 1] / cnt;
    }
  }
  const size_t img = (size_t)n * HW * C8N + c8;
  const uint4* xp = reinterpret_cast<const uint4*>(x) + img;
  const uint4* dp = reinterpret_cast<const uint4*>(dy) + img;
  uint4* yp = reinterpret_cast<uint4*>(y) + img;
  auto apply = [&](const uint4& xv, const uint4& dv) {
    float xf[8], o[8];
    unpack8(xv, xf);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = fmaf(xf[j], g[j], b[j]);
        if (act_silu) o[j] = silu(o[j]);
      }
    } else {
      float df[8];
      unpack8(dv, df);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = j >> 1;
        const float xh = (xf[j] - k.mean[e]) * k.rstd[e];
        float dz = df[j];
        if (act_silu) dz *= silu_grad(fmaf(xh, g[j], b[j]));
        o[j] = k.rstd[e] * (dz * g[j] - r0[e] - xh * r1[e]);
      }
    }
    return pack8(o);
  };
  const int step = gridDim.x * PL;
  int p = blockIdx.x * PL + pl;
  for (; p + (UN - 1) * step < HW; p += UN * step) {
    uint4 xv[UN], dv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      xv[u] = xp[(size_t)(p + u * step) * C8N];
      if (MODE == 1) dv[u] = dp[(size_t)(p + u * step) * C8N];
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) yp[(size_t)(p + u * step) * C8N] = apply(xv[u], dv[u]);
  }
  for (; p < HW; p += step) {
    uint4 dv = make_uint4(0, 0, 0, 0);
    if (MODE == 1) dv = dp[(size_t)p * C8N];
    yp[(size_t)p * C8N] = apply(xp[(size_t)p * C8N], dv);
  }
}

// Single-launch GroupNorm forward for many small (image, group) slices (the UNet: N x 32 >= 128 slices of <= 160 k
// elements that sit in L2): one CTA per slice does the statistics pass and the apply pass back to back, in a fixed
// summation order. Replaces reduce + finalize + apply (three launches of ~5-17 us each on ~13 MB tensors).
__global__ void __launch_bounds__(256)
gn_fused_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma, const __half* __restrict__ beta,
                __half* __restrict__ y, float* __restrict__ stats, int HW, int C, int groups, float eps, int act_silu) {
  pdl_prologue();
  __shared__ float red[2][8];
  __shared__ float s_mean, s_rstd;
  const int n = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = C / groups, PP = cpg / 2;  // channel pairs of the group
  const int PL = 256 / PP;                   // pixel lanes
  const int pp = threadIdx.x % PP, pl = threadIdx.x / PP;
  const bool active = pl < PL;
  const size_t base2 = ((size_t)n * HW * C + (size_t)g * cpg) / 2 + pp;  // half2 index of (pixel 0, this pair)
  const size_t stride2 = (size_t)C / 2;
  const __half2* x2 = reinterpret_cast<const __half2*>(x);
  float s = 0.f, ss = 0.f;
  if (active)
    for (int p = pl; p < HW; p += PL) {
      const float2 v = __half22float2(x2[base2 + (size_t)p * stride2]);
      s += v.x + v.y;
      ss = fmaf(v.x, v.x, fmaf(v.y, v.y, ss));
    }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[0][warp] = s, red[1][warp] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[0][w], b += red[1][w];
    stats[(n * groups + g) * 2] = a;
    stats[(n * groups + g) * 2 + 1] = b;
    const float cnt = (float)HW * (float)cpg;
    const float mean = a / cnt;
    s_mean = mean;
    s_rstd = rsqrtf(fmaxf(b / cnt - mean * mean, 0.f) + eps);
  }
  __syncthreads();
  if (!active) return;
  const float mean = s_mean, rstd = s_rstd;
  const float2 gm = __half22float2(reinterpret_cast<const __half2*>(gamma)[(g * cpg) / 2 + pp]);
  const float2 bt = __half22float2(reinterpret_cast<const __half2*>(beta)[(g * cpg) / 2 + pp]);
  __half2* y2 = reinterpret_cast<__half2*>(y);
  for (int p = pl; p < HW; p += PL) {
    const size_t idx = base2 + (size_t)p * stride2;
    const float2 v = __half22float2(x2[idx]);
    float o0 = (v.x - mean) * rstd * gm.x + bt.x, o1 = (v.y - mean) * rstd * gm.y + bt.y;
    if (act_silu) {
      o0 = silu(o0);
      o1 = silu(o1);
    }
    y2[idx] = __floats2half2_rn(o0, o1);
  }
}

// Block shape (C8N channel octets x PL pixel lanes, 128..512 threads) and the pixel split of the statistics pass
// (~4 waves of blocks over N x nblk).
void gn_geometry(int HW, int C, int N, int* C8N, int* PL, int* ppb, int* nblk) {
  static const int thr = getenv("SDB_GN_THREADS") ? atoi(getenv("SDB_GN_THREADS")) : 256;
  static const int waves = getenv("SDB_GN_WAVES") ? atoi(getenv("SDB_GN_WAVES")) : 4;
  const int c8n = C / 8;
  int pl = thr / c8n;
  if (pl < 1) pl = 1;
  if (pl > HW) pl = HW;
  int splits = (waves * kNumSMs + N - 1) / N;
  int pix = (HW + splits - 1) / splits;
  if (pix < pl * 8) pix = pl * 8;
  if (pix > HW) pix = HW;
  *C8N = c8n;
  *PL = pl;
  *ppb = pix;
  *nblk = (HW + pix - 1) / pix;
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// Warp per row. Rows of up to 32 * 8 * LN_ITEMS channels (multiple of 8) are read ONCE with 16-byte loads and stay in
// registers between the mean / variance / apply passes; wider rows fall back to three strided passes.
constexpr int LN_ITEMS = 5;  // 1280 channels
__global__ void layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                                 const __half* __restrict__ beta, __half* __restrict__ y, int rows, int C, float eps) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  if ((C & 7) == 0 && C <= 256 * LN_ITEMS) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * C);
    const int n8 = C >> 3;
    float v[LN_ITEMS][8];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < LN_ITEMS; ++it) {
      const int i = lane + 32 * it;
      uint4 raw = make_uint4(0, 0, 0, 0);
      if (i < n8) raw = xr[i];
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = __half22float2(h2[e]);
        v[it][2 * e] = t.x, v[it][2 * e + 1] = t.y;
        s += t.x + t.y;
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < LN_ITEMS; ++it)
      if (lane + 32 * it < n8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += (v[it][e] - mean) * (v[it][e] - mean);
      }
    const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
    uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * C);
#pragma unroll
    for (int it = 0; it < LN_ITEMS; ++it) {
      const int i = lane + 32 * it;
      if (i >= n8) continue;
      const uint4 gr = reinterpret_cast<const uint4*>(gamma)[i], br = reinterpret_cast<const uint4*>(beta)[i];
      const __half2* g2 = reinterpret_cast<const __half2*>(&gr);
      const __half2* b2 = reinterpret_cast<const __half2*>(&br);
      uint4 o;
      __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 g = __half22float2(g2[e]), b = __half22float2(b2[e]);
        o2[e] = __floats2half2_rn((v[it][2 * e] - mean) * rstd * g.x + b.x, (v[it][2 * e + 1] - mean) * rstd * g.y + b.y);
      }
      yr[i] = o;
    }
    return;
  }
  const __half2* xr = reinterpret_cast<const __half2*>(x + (size_t)row * C);
  const int n2 = C / 2;
  float s = 0.f;
  for (int i = lane; i < n2; i += 32) {
    const float2 v = __half22float2(xr[i]);
    s += v.x + v.y;
  }
  s = warp_sum(s);
  const float mean = s / (float)C;
  float ss = 0.f;
  for (int i = lane; i < n2; i += 32) {
    const float2 v = __half22float2(xr[i]);
    ss += (v.x - mean) * (v.x - mean) + (v.y - mean) * (v.y - mean);
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / (float)C + eps);
  __half2* yr = reinterpret_cast<__half2*>(y + (size_t)row * C);
  for (int i = lane; i < n2; i += 32) {
    const float2 v = __half22float2(xr[i]);
    const float2 g = __half22float2(reinterpret_cast<const __half2*>(gamma)[i]);
    const float2 b = __half22float2(reinterpret_cast<const __half2*>(beta)[i]);
    yr[i] = __floats2half2_rn((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y);
  }
}

// ---------------------------------------------------------------------------------------------- softmax
// One block of 32*WARPS threads per row; each thread owns ITEMS groups of 8 consecutive columns (16-byte accesses),
// the row stays in registers between the max / sum / normalise passes. cols <= 256 * WARPS * ITEMS.
template <int WARPS, int ITEMS>
__global__ void __launch_bounds__(32 * WARPS) softmax_kernel(__half* __restrict__ x, int cols, long long ld) {
  pdl_prologue();
  __shared__ float red[WARPS];
  __half* row = x + (size_t)blockIdx.x * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float v[ITEMS][8];
  float mx = -INFINITY;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int c0 = (tid + it * 32 * WARPS) * 8;
    uint4 raw = make_uint4(0, 0, 0, 0);
    if (c0 < ld) raw = *reinterpret_cast<const uint4*>(row + c0);
    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h2[e]);
      v[it][2 * e] = c0 + 2 * e < cols ? f.x : -INFINITY;
      v[it][2 * e + 1] = c0 + 2 * e + 1 < cols ? f.y : -INFINITY;
      mx = fmaxf(mx, fmaxf(v[it][2 * e], v[it][2 * e + 1]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
  if (WARPS > 1) {
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
  }
  float sum = 0.f;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[it][e] = __expf(v[it][e] - mx);  // exp(-inf) = 0 for the masked tail
      sum += v[it][e];
    }
  sum = warp_sum(sum);
  if (WARPS > 1) {
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = red[0];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) sum += red[w];
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int c0 = (tid + it * 32 * WARPS) * 8;
    if (c0 >= ld) continue;
    uint4 ov;
    __half2* h2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(v[it][2 * e] * inv, v[it][2 * e + 1] * inv);
    *reinterpret_cast<uint4*>(row + c0) = ov;
  }
}

__global__ void geglu_kernel(const __half* __restrict__ xg, __half* __restrict__ y, long long rows, int inner) {
  pdl_prologue();
  const int i2 = inner / 2;
  const long long total = rows * i2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / i2;
    const int c = (int)(i - r * i2);
    const __half2* base = reinterpret_cast<const __half2*>(xg + r * 2 * inner);
    const float2 a = __half22float2(base[c]);
    const float2 g = __half22float2(base[i2 + c]);
    reinterpret_cast<__half2*>(y + r * inner)[c] = __floats2half2_rn(a.x * gelu_erf(g.x), a.y * gelu_erf(g.y));
  }
}

__global__ void upsample2x_kernel(const __half* __restrict__ x, __half* __restrict__ y, int N, int H, int W, int C8) {
  pdl_prologue();
  const long long total = (long long)N * 2 * H * 2 * W * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int ox = (int)(r % (2 * W));
    r /= 2 * W;
    const int oy = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    reinterpret_cast<uint4*>(y)[i] =
        reinterpret_cast<const uint4*>(x)[(((long long)n * H + oy / 2) * W + ox / 2) * C8 + c];
  }
}

__global__ void concat_kernel(const __half* __restrict__ a, int Ca8, const __half* __restrict__ b, int Cb8,
                              __half* __restrict__ y, long long rows) {
  pdl_prologue();
  const int C8 = Ca8 + Cb8;
  const long long total = rows * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C8;
    const int c = (int)(i - r * C8);
    reinterpret_cast<uint4*>(y)[i] = c < Ca8 ? reinterpret_cast<const uint4*>(a)[r * Ca8 + c]
                                             : reinterpret_cast<const uint4*>(b)[r * Cb8 + (c - Ca8)];
  }
}

__global__ void add_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ y,
                           long long n2) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<__half2*>(y)[i] =
        __hadd2(reinterpret_cast<const __half2*>(a)[i], reinterpret_cast<const __half2*>(b)[i]);
}

__global__ void transpose_kernel(const __half* __restrict__ x, __half* __restrict__ y, int rows, int cols) {
  pdl_prologue();
  __shared__ __half tile[32][34];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = by + j, c = bx + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = x[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = bx + j, r = by + threadIdx.x;
    if (r < rows && c < cols) y[(size_t)c * rows + r] = tile[threadIdx.x][j];
  }
}

// col[(n,oy,ox), (kh,kw,c)] = x[n, 2*oy + kh - pad_lo, 2*ox + kw - pad_lo, c] (zero outside)
__global__ void im2col_s2_kernel(const __half* __restrict__ x, __half* __restrict__ col, int N, int H, int W, int C8,
                                 int pad_lo) {
  pdl_prologue();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * 9 * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int tap = (int)(r % 9);
    r /= 9;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const int iy = 2 * oy + tap / 3 - pad_lo, ix = 2 * ox + tap % 3 - pad_lo;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = reinterpret_cast<const uint4*>(x)[(((long long)n * H + iy) * W + ix) * C8 + c];
    reinterpret_cast<uint4*>(col)[i] = v;
  }
}

// dx[n,iy,ix,c] = sum over (oy,ox,tap) with 2*oy + kh - pad_lo == iy, 2*ox + kw - pad_lo == ix of col[...]
__global__ void col2im_s2_kernel(const __half* __restrict__ col, __half* __restrict__ dx, int N, int H, int W, int C8,
                                 int pad_lo) {
  pdl_prologue();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * H * W * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int ix = (int)(r % W);
    r /= W;
    const int iy = (int)(r % H);
    const int n = (int)(r / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int kh = 0; kh < 3; ++kh) {
      const int ty = iy + pad_lo - kh;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= Ho) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int tx = ix + pad_lo - kw;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= Wo) continue;
        const uint4 v =
            reinterpret_cast<const uint4*>(col)[((((long long)n * Ho + oy) * Wo + ox) * 9 + kh * 3 + kw) * C8 + c];
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          acc[2 * e] += f.x;
          acc[2 * e + 1] += f.y;
        }
      }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
    reinterpret_cast<uint4*>(dx)[i] = o;
  }
}

// Direct 3x3 conv for tiny Cin (3 / 4 / 8: the RGB / latent / moment ends of the networks). Thread = PX consecutive
// pixels of a row x 8 output channels: the (3 x (PX+2) x CIN) input window stays in registers and every pair of
// LDS.128 weight reads feeds 8 * PX FMAs (the first version, one pixel per thread, was bound by the shared-memory
// pipe at one LDS.128 per four FMAs). Weights sit in shared memory as fp32, transposed to [k][CT].
template <typename TIn, int CIN, int PX>
__global__ void __launch_bounds__(128)
conv_small_cin_kernel(const TIn* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ bias,
                      void* __restrict__ y, int y_fp32, int N, int H, int W, int Cout, int CT) {
  pdl_prologue();
  extern __shared__ float swf[];  // [9*CIN][CT]: the CT output channels of slice blockIdx.y
  constexpr int K = 9 * CIN;
  const int co0 = blockIdx.y * CT;
  for (int i = threadIdx.x; i < CT * K; i += blockDim.x) {
    const int co = i / K, k = i - co * K;
    swf[k * CT + co] = __half2float(w[(size_t)(co0 + co) * K + k]);
  }
  __syncthreads();
  const int co8n = CT / 8, wq = (W + PX - 1) / PX;
  const long long total = (long long)N * H * wq * co8n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % co8n);
    long long r = i / co8n;
    const int px0 = (int)(r % wq) * PX;
    r /= wq;
    const int py = (int)(r % H);
    const int n = (int)(r / H);
    float in[3][PX + 2][CIN];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = py + ky - 1;
#pragma unroll
      for (int q = 0; q < PX + 2; ++q) {
        const int ix = px0 + q - 1;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
        for (int c = 0; c < CIN; ++c)
          in[ky][q][c] = ok ? (float)x[(((long long)n * H + iy) * W + ix) * CIN + c] : 0.f;
      }
    }
    float acc[PX][8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float bv = bias ? __half2float(bias[co0 + cg * 8 + o]) : 0.f;
#pragma unroll
      for (int p = 0; p < PX; ++p) acc[p][o] = bv;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const int k = (ky * 3 + kx) * CIN + c;
          const float4 w0 = *reinterpret_cast<const float4*>(swf + k * CT + cg * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(swf + k * CT + cg * 8 + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[p][o] = fmaf(in[ky][kx + p][c], wv[o], acc[p][o]);
        }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (px0 + p >= W) break;
      const long long ob = (((long long)n * H + py) * W + px0 + p) * Cout + co0 + cg * 8;
      if (y_fp32) {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + ob)[0] = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + ob)[1] = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
      } else {
        uint4 ov;
        __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(acc[p][2 * e], acc[p][2 * e + 1]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(y) + ob) = ov;
      }
    }
  }
}

// Direct 3x3 conv for tiny Cout (3 / 4 / 8: the last convolution of the UNet, conv_out of the VAE, and the data
// gradients of their first convolutions) on LARGE images. Thread = PX consecutive pixels of a row x all COUT outputs, looping over
// taps and 8-channel chunks: one 16-byte load per pixel and 2 * COUT LDS.128 (broadcast) feed 32 * COUT FMAs. The first
// warp-per-pixel kernel below stays for small images, where a thread per PX pixels would leave the GPU empty
// (64 x 64: 1024 threads).
template <int COUT, int PX>
__global__ void __launch_bounds__(128)
conv_small_cout_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const __half* __restrict__ bias,
                       void* __restrict__ y, int y_fp32, int N, int H, int W, int Cin) {
  pdl_prologue();
  extern __shared__ float swo[];  // [9][Cin][COUT] fp32
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < COUT * K; i += blockDim.x) {
    const int o = i / K, k = i - o * K;
    swo[k * COUT + o] = __half2float(w[(size_t)o * K + k]);
  }
  __syncthreads();
  const int wq = (W + PX - 1) / PX, c8n = Cin / 8;
  const long long total = (long long)N * H * wq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int px0 = (int)(i % wq) * PX;
    const int py = (int)((i / wq) % H);
    const int n = (int)(i / ((long long)wq * H));
    float acc[PX][COUT];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[p][o] = 0.f;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = py + tap / 3 - 1, dx = tap % 3 - 1;
      if (iy < 0 || iy >= H) continue;
      const uint4* row = reinterpret_cast<const uint4*>(x + ((long long)n * H + iy) * W * Cin);
      const float* wt = swo + (size_t)tap * Cin * COUT;
#pragma unroll 1
      for (int c8 = 0; c8 < c8n; ++c8) {
        uint4 xv[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const int ix = px0 + p + dx;
          xv[p] = (ix >= 0 && ix < W) ? row[(long long)ix * c8n + c8] : make_uint4(0, 0, 0, 0);
        }
        float wv[8 * COUT];  // the chunk's weights, [e][o] contiguous: 2 * COUT LDS.128
#pragma unroll
        for (int q = 0; q < 2 * COUT; ++q) {
          const float4 t = reinterpret_cast<const float4*>(wt + (size_t)c8 * 8 * COUT)[q];
          wv[4 * q] = t.x, wv[4 * q + 1] = t.y, wv[4 * q + 2] = t.z, wv[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const __half2* h2 = reinterpret_cast<const __half2*>(&xv[p]);
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const float2 xf = __half22float2(h2[e2]);
#pragma unroll
            for (int o = 0; o < COUT; ++o)
              acc[p][o] = fmaf(xf.x, wv[(2 * e2) * COUT + o], fmaf(xf.y, wv[(2 * e2 + 1) * COUT + o], acc[p][o]));
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (px0 + p >= W) break;
      const long long pix = ((long long)n * H + py) * W + px0 + p;
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const float v = acc[p][o] + (bias ? __half2float(bias[o]) : 0.f);
        if (y_fp32)
          reinterpret_cast<float*>(y)[pix * COUT + o] = v;
        else
          reinterpret_cast<__half*>(y)[pix * COUT + o] = __float2half_rn(v);
      }
    }
  }
}

// Small images: one warp per pixel, lanes split Cin, warp-shuffle reduction of the COUT sums.
template <int COUT>
__global__ void conv_small_cout_warp_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                            const __half* __restrict__ bias, void* __restrict__ y, int y_fp32, int N,
                                            int H, int W, int Cin) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)N * H * W;
  const int K = 9 * Cin;
  for (long long pix = wid; pix < total; pix += nw) {
    const int px = (int)(pix % W);
    const int py = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = py + tap / 3 - 1, ix = px + tap % 3 - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const __half2* xr = reinterpret_cast<const __half2*>(x + (((long long)n * H + iy) * W + ix) * Cin);
      for (int c2 = lane; c2 < Cin / 2; c2 += 32) {
        const float2 xv = __half22float2(xr[c2]);
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
          const float2 wv = __half22float2(reinterpret_cast<const __half2*>(w + (size_t)o * K + tap * Cin)[c2]);
          acc[o] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, acc[o]));
        }
      }
    }
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = warp_sum(acc[o]);
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const float v = acc[o] + (bias ? __half2float(bias[o]) : 0.f);
        if (y_fp32)
          reinterpret_cast<float*>(y)[pix * COUT + o] = v;
        else
          reinterpret_cast<__half*>(y)[pix * COUT + o] = __float2half_rn(v);
      }
    }
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, __half* __restrict__ out, int n, int dim,
                                          float max_period) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= n * half) return;
  const int b = i / half, j = i % half;
  const float freq = expf(-logf(max_period) * (float)j / (float)half);
  const float a = t[b] * freq;
  out[(size_t)b * dim + j] = __float2half_rn(cosf(a));
  out[(size_t)b * dim + half + j] = __float2half_rn(sinf(a));
}

// One warp per output column; rows <= 16.
__global__ void linear_small_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                    const __half* __restrict__ bias, void* __restrict__ y, int y_fp32, int rows, int N,
                                    int K, int silu_in) {
  pdl_prologue();
  extern __shared__ float sx[];  // [rows][K]
  for (int i = threadIdx.x; i < rows * K; i += blockDim.x) {
    float v = __half2float(x[i]);
    sx[i] = silu_in ? silu(v) : v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (col >= N) return;
  float acc[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) acc[r] = 0.f;
  const __half2* wr = reinterpret_cast<const __half2*>(w + (size_t)col * K);
  for (int k2 = lane; k2 < K / 2; k2 += 32) {
    const float2 wv = __half22float2(wr[k2]);
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (r < rows) acc[r] = fmaf(wv.x, sx[r * K + 2 * k2], fmaf(wv.y, sx[r * K + 2 * k2 + 1], acc[r]));
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    if (r >= rows) break;
    const float v = warp_sum(acc[r]) + (bias ? __half2float(bias[col]) : 0.f);
    if (lane == 0) {
      if (y_fp32)
        reinterpret_cast<float*>(y)[(size_t)r * N + col] = v;
      else
        reinterpret_cast<__half*>(y)[(size_t)r * N + col] = __float2half_rn(v);
    }
  }
}

__global__ void silu_f32_to_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, long long n) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(silu(x[i]));
}

inline int ew_grid(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

inline bool gn_un8() {
  static const bool v = getenv("SDB_GN_UNROLL") && atoi(getenv("SDB_GN_UNROLL")) == 8;
  return v;
}

// Blocks per image of the apply pass: ~8 blocks per SM over the whole batch, at least 4 pixels per thread.
inline int gn_apply_blocks(int HW, int pl, int N) {
  static const int per_sm = getenv("SDB_GN_APPLY_BLOCKS") ? atoi(getenv("SDB_GN_APPLY_BLOCKS")) : 8;
  const int want = (per_sm * kNumSMs + N - 1) / N, most = (HW + 4 * pl - 1) / (4 * pl);
  return want < most ? want : (most > 0 ? most : 1);
}

}  // namespace

long long groupnorm_workspace_floats(int N, int HW, int C, int groups) {
  int c8n, pl, ppb, nblk;
  gn_geometry(HW, C, N, &c8n, &pl, &ppb, &nblk);
  return (long long)N * groups * 2 * (1 + nblk);
}

int groupnorm_forward(const __half* x, const __half* gamma, const __half* beta, __half* y, float* stats, int N, int HW,
                      int C, int groups, float eps, int act_silu, cudaStream_t s) {
  if (C % 8 || C % groups || (C / groups) % 2 || C > 4096) {
    sdb_set_error("groupnorm: C=%d must be a multiple of 8 (<= 4096) with an even group size", C);
    return SDB_ERR_UNSUPPORTED;
  }
  const int cpg = C / groups;
  // One launch instead of three pays off only while the tensor is small enough for launch latency to dominate; its
  // per-group slices are strided (cpg * 2 contiguous bytes per pixel), so above a few MB the coalesced passes win
  // (C2 step: 30.0 ms with everything fused, 29.5 ms with nothing fused, 28.0 ms with the 8 MB switch).
  static const long long fused_max = getenv("SDB_GN_FUSED_MAX") ? atoll(getenv("SDB_GN_FUSED_MAX")) : (8ll << 20);
  if (N * groups >= 128 && cpg / 2 <= 256 && (long long)N * HW * C * 2 <= fused_max) {
    sdb_launch(gn_fused_kernel, N * groups, 256, 0, s, x, gamma, beta, y, stats, HW, C, groups, eps, act_silu);
    SDB_COUNT_LAUNCH();
    SDB_CHECK_LAUNCH("gn_fused");
    return SDB_OK;
  }
  int c8n, pl, ppb, nblk;
  gn_geometry(HW, C, N, &c8n, &pl, &ppb, &nblk);
  float* partial = stats + (size_t)N * groups * 2;
  const int threads = c8n * pl;
  const size_t smem = sizeof(float) * (size_t)pl * C;
  sdb_launch((gn_un8() ? gn_reduce_kernel<0, 8> : gn_reduce_kernel<0, 4>), dim3(nblk, N), threads, smem, s, x, nullptr, nullptr, nullptr, nullptr, partial, HW, C, groups,
                                                           c8n, pl, ppb, eps, 0);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("gn_stats");
  sdb_launch(gn_finalize_kernel, (N * groups * 2 + 3) / 4, 128, 0, s, partial, stats, nblk, groups, N * groups * 2);
  SDB_COUNT_LAUNCH();
  sdb_launch((gn_un8() ? gn_apply_kernel<0, 8> : gn_apply_kernel<0, 4>), dim3(gn_apply_blocks(HW, pl, N), N), threads, 0, s, x, nullptr, gamma, beta, stats, nullptr, y,
                                                                             HW, C, groups, c8n, pl, eps, act_silu);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("gn_apply");
  return SDB_OK;
}

int groupnorm_backward(const __half* x, const __half* gamma, const __half* beta, const float* stats, const __half* dy,
                       __half* dx, float* scratch2, int N, int HW, int C, int groups, float eps, int act_silu,
                       cudaStream_t s) {
  int c8n, pl, ppb, nblk;
  gn_geometry(HW, C, N, &c8n, &pl, &ppb, &nblk);
  float* partial = scratch2 + (size_t)N * groups * 2;
  const int threads = c8n * pl;
  const size_t smem = sizeof(float) * (size_t)pl * C;
  sdb_launch((gn_un8() ? gn_reduce_kernel<1, 8> : gn_reduce_kernel<1, 4>), dim3(nblk, N), threads, smem, s, x, dy, gamma, beta, stats, partial, HW, C, groups, c8n, pl,
                                                           ppb, eps, act_silu);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("gn_bwd_reduce");
  sdb_launch(gn_finalize_kernel, (N * groups * 2 + 3) / 4, 128, 0, s, partial, scratch2, nblk, groups, N * groups * 2);
  SDB_COUNT_LAUNCH();
  sdb_launch((gn_un8() ? gn_apply_kernel<1, 8> : gn_apply_kernel<1, 4>), dim3(gn_apply_blocks(HW, pl, N), N), threads, 0, s, x, dy, gamma, beta, stats, scratch2, dx, HW,
                                                                             C, groups, c8n, pl, eps, act_silu);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("gn_bwd_apply");
  return SDB_OK;
}

int layernorm_forward(const __half* x, const __half* gamma, const __half* beta, __half* y, int rows, int C, float eps,
                      cudaStream_t s) {
  sdb_launch(layernorm_kernel, (rows + 3) / 4, 128, 0, s, x, gamma, beta, y, rows, C, eps);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("layernorm");
  return SDB_OK;
}

int softmax_rows(__half* x, long long rows, int cols, long long ld, cudaStream_t s) {
  if (cols > 4096 || ld > 4096 || (ld & 7) || (reinterpret_cast<uintptr_t>(x) & 15)) {
    sdb_set_error("softmax: at most 4096 columns, row stride a multiple of 8, 16-byte aligned base (got %d, ld %lld)",
                  cols, ld);
    return SDB_ERR_UNSUPPORTED;
  }
  const unsigned g = (unsigned)rows;
  if (ld <= 256) sdb_launch(softmax_kernel<1, 1>, g, 32, 0, s, x, cols, ld);
  else if (ld <= 1024) sdb_launch(softmax_kernel<4, 1>, g, 128, 0, s, x, cols, ld);
  else if (ld <= 2048) sdb_launch(softmax_kernel<4, 2>, g, 128, 0, s, x, cols, ld);
  else sdb_launch(softmax_kernel<4, 4>, g, 128, 0, s, x, cols, ld);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("softmax");
  return SDB_OK;
}

int geglu(const __half* xg, __half* y, long long rows, int inner, cudaStream_t s) {
  sdb_launch(geglu_kernel, ew_grid(rows * inner / 2), 256, 0, s, xg, y, rows, inner);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("geglu");
  return SDB_OK;
}

int silu_f32_to_f16(const float* x, __half* y, long long n, cudaStream_t s) {
  sdb_launch(silu_f32_to_f16_kernel, ew_grid(n), 256, 0, s, x, y, n);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("silu");
  return SDB_OK;
}

int upsample_nearest2x(const __half* x, __half* y, int N, int H, int W, int C, cudaStream_t s) {
  sdb_launch(upsample2x_kernel, ew_grid((long long)N * 4 * H * W * C / 8), 256, 0, s, x, y, N, H, W, C / 8);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("upsample2x");
  return SDB_OK;
}

int concat_channels(const __half* a, int Ca, const __half* b, int Cb, __half* y, long long rows, cudaStream_t s) {
  sdb_launch(concat_kernel, ew_grid(rows * (Ca + Cb) / 8), 256, 0, s, a, Ca / 8, b, Cb / 8, y, rows);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("concat");
  return SDB_OK;
}

int add_f16(const __half* a, const __half* b, __half* y, long long n, cudaStream_t s) {
  sdb_launch(add_kernel, ew_grid(n / 2), 256, 0, s, a, b, y, n / 2);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("add");
  return SDB_OK;
}

int transpose_f16(const __half* x, __half* y, int rows, int cols, cudaStream_t s) {
  sdb_launch(transpose_kernel, dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, s, x, y, rows, cols);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("transpose");
  return SDB_OK;
}

int im2col_3x3_s2(const __half* x, __half* col, int N, int H, int W, int C, int pad_lo, cudaStream_t s) {
  sdb_launch(im2col_s2_kernel, ew_grid((long long)N * (H / 2) * (W / 2) * 9 * C / 8), 256, 0, s, x, col, N, H, W, C / 8,
                                                                                        pad_lo);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("im2col_s2");
  return SDB_OK;
}

int col2im_3x3_s2(const __half* col, __half* dx, int N, int H, int W, int C, int pad_lo, cudaStream_t s) {
  sdb_launch(col2im_s2_kernel, ew_grid((long long)N * H * W * C / 8), 256, 0, s, col, dx, N, H, W, C / 8, pad_lo);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("col2im_s2");
  return SDB_OK;
}

int conv3x3_small(const void* x, int x_fp32, const __half* w, const __half* bias, void* y, int y_fp32, int N, int H,
                  int W, int Cin, int Cout, cudaStream_t s) {
  if ((Cin == 3 || Cin == 4 || Cin == 8) && Cout % 8 == 0) {
    int CT = 128;  // output-channel slice per block: the largest multiple of 8 <= 128 dividing Cout
    while (Cout % CT) CT -= 8;
    const size_t smem = sizeof(float) * CT * 9 * Cin;  // <= 36 KB
    const int px = Cin == 8 ? 2 : 4;  // pixels per thread (register budget: 3 x (px + 2) x Cin inputs)
    const long long total = (long long)N * H * ((W + px - 1) / px) * (CT / 8);
    const dim3 grid(ew_grid(total, 128), Cout / CT);
#define SDB_SMALL_CIN(T, CI, PX) \
  sdb_launch(conv_small_cin_kernel<T, CI, PX>, grid, 128, smem, s, reinterpret_cast<const T*>(x), w, bias, y, y_fp32, N, H, W, Cout, CT)
    if (x_fp32) {
      if (Cin == 3) SDB_SMALL_CIN(float, 3, 4);
      else if (Cin == 4) SDB_SMALL_CIN(float, 4, 4);
      else SDB_SMALL_CIN(float, 8, 2);
    } else {
      if (Cin == 3) SDB_SMALL_CIN(__half, 3, 4);
      else if (Cin == 4) SDB_SMALL_CIN(__half, 4, 4);
      else SDB_SMALL_CIN(__half, 8, 2);
    }
#undef SDB_SMALL_CIN
  } else if (Cout <= 8 && !x_fp32 && Cin % 8 == 0 && (size_t)Cout * 9 * Cin * 4 <= 200 * 1024 &&
             (long long)N * H * W >= 131072) {
    // large images: thread per PX pixels (2 until there are >= 512 k pixels, so that >= 64 k threads exist)
    const int px = (long long)N * H * W >= 524288 ? 4 : 2;
    const long long total = (long long)N * H * ((W + px - 1) / px);
    const int grid = ew_grid(total, 128);
    const size_t smem = sizeof(float) * Cout * 9 * Cin;
    const __half* xh = reinterpret_cast<const __half*>(x);
#define SDB_SMALL_COUT(CO, PX)                                                                                     \
  do {                                                                                                             \
    cudaFuncSetAttribute(conv_small_cout_kernel<CO, PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    sdb_launch(conv_small_cout_kernel<CO, PX>, grid, 128, smem, s, xh, w, bias, y, y_fp32, N, H, W, Cin);                  \
  } while (0)
    switch (Cout * 10 + px) {
      case 32: SDB_SMALL_COUT(3, 2); break;
      case 34: SDB_SMALL_COUT(3, 4); break;
      case 42: SDB_SMALL_COUT(4, 2); break;
      case 44: SDB_SMALL_COUT(4, 4); break;
      case 82: SDB_SMALL_COUT(8, 2); break;
      case 84: SDB_SMALL_COUT(8, 4); break;
      default:
        sdb_set_error("conv3x3_small: Cout=%d not instantiated", Cout);
        return SDB_ERR_UNSUPPORTED;
    }
#undef SDB_SMALL_COUT
  } else if (Cout <= 8 && !x_fp32 && Cin % 2 == 0) {
    const long long total = (long long)N * H * W;
    const int grid = ew_grid(total * 32);
    const __half* xh = reinterpret_cast<const __half*>(x);
    switch (Cout) {
      case 3: sdb_launch(conv_small_cout_warp_kernel<3>, grid, 256, 0, s, xh, w, bias, y, y_fp32, N, H, W, Cin); break;
      case 4: sdb_launch(conv_small_cout_warp_kernel<4>, grid, 256, 0, s, xh, w, bias, y, y_fp32, N, H, W, Cin); break;
      case 8: sdb_launch(conv_small_cout_warp_kernel<8>, grid, 256, 0, s, xh, w, bias, y, y_fp32, N, H, W, Cin); break;
      default:
        sdb_set_error("conv3x3_small: Cout=%d not instantiated", Cout);
        return SDB_ERR_UNSUPPORTED;
    }
  } else {
    sdb_set_error("conv3x3_small: unsupported Cin=%d Cout=%d", Cin, Cout);
    return SDB_ERR_UNSUPPORTED;
  }
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("conv3x3_small");
  return SDB_OK;
}

int timestep_embedding(const float* t, __half* out, int n, int dim, float max_period, cudaStream_t s) {
  sdb_launch(timestep_embedding_kernel, (n * dim / 2 + 127) / 128, 128, 0, s, t, out, n, dim, max_period);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("timestep_embedding");
  return SDB_OK;
}

int linear_small(const __half* x, const __half* w, const __half* bias, void* y, int y_fp32, int rows, int N, int K,
                 int silu_in, cudaStream_t s) {
  if (rows > 16 || (K & 1)) {
    sdb_set_error("linear_small: rows=%d must be <= 16 and K even", rows);
    return SDB_ERR_UNSUPPORTED;
  }
  const size_t smem = sizeof(float) * rows * K;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(linear_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr = true;
  }
  sdb_launch(linear_small_kernel, (N + 7) / 8, 256, smem, s, x, w, bias, y, y_fp32, rows, N, K, silu_in);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("linear_small");
  return SDB_OK;
}

}  // namespace dense

// ---- additions used by the network executors -----------------------------------------------------------------
namespace dense {
namespace {

// wr[ci, kh, kw, co] = w[co, 2-kh, 2-kw, ci]: weights of the data-gradient convolution.
__global__ void rotate_w3x3_kernel(const __half* __restrict__ w, __half* __restrict__ wr, int Cout, int Cin) {
  pdl_prologue();
  const long long total = (long long)Cout * 9 * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    long long r = i / Cout;
    const int tap = (int)(r % 9);
    const int ci = (int)(r / 9);
    wr[i] = w[((long long)co * 9 + (8 - tap)) * Cin + ci];
  }
}

// dS = P * (dP - sum_j dP_j P_j) * scale, in place on dP. One 128-thread block per row, cols <= 4096.
__global__ void __launch_bounds__(128) softmax_bwd_kernel(const __half* __restrict__ P, __half* __restrict__ dP,
                                                          int cols, long long ld, float scale) {
  pdl_prologue();
  __shared__ float red[4];
  const __half* pr = P + (size_t)blockIdx.x * ld;
  __half* dr = dP + (size_t)blockIdx.x * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float p[32], d[32];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int c = tid + j * 128;
    p[j] = c < cols ? __half2float(pr[c]) : 0.f;
    d[j] = c < cols ? __half2float(dr[c]) : 0.f;
    dot = fmaf(p[j], d[j], dot);
  }
  dot = warp_sum(dot);
  if (lane == 0) red[warp] = dot;
  __syncthreads();
  dot = red[0] + red[1] + red[2] + red[3];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int c = tid + j * 128;
    if (c < ld && c < 4096) dr[c] = __float2half_rn(c < cols ? p[j] * (d[j] - dot) * scale : 0.f);
  }
}

__global__ void add_silu_kernel(const float* __restrict__ a, const float* __restrict__ b, __half* __restrict__ y,
                                long long n) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = a[i] + (b ? b[i] : 0.f);
    y[i] = __float2half_rn(v / (1.f + __expf(-v)));
  }
}

}  // namespace

namespace {
__global__ void interleave_geglu_rows_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int half_rows,
                                             int cols) {
  pdl_prologue();
  const long long total = 2LL * half_rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const long long r = i / cols;
    const int chunk = (int)(r / 32), j = (int)(r % 32);
    const long long sr = j < 16 ? (long long)chunk * 16 + j : (long long)half_rows + chunk * 16 + (j - 16);
    dst[i] = src[sr * cols + c];
  }
}
}  // namespace

int interleave_geglu_rows(const __half* src, __half* dst, int half_rows, int cols, cudaStream_t s) {
  if (half_rows % 16) {
    sdb_set_error("geglu interleave: inner width %d must be a multiple of 16", half_rows);
    return SDB_ERR_UNSUPPORTED;
  }
  sdb_launch(interleave_geglu_rows_kernel, ew_grid(2LL * half_rows * cols), 256, 0, s, src, dst, half_rows, cols);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("interleave_geglu_rows");
  return SDB_OK;
}

int rotate_w3x3(const __half* w, __half* wr, int Cout, int Cin, cudaStream_t s) {
  sdb_launch(rotate_w3x3_kernel, ew_grid((long long)Cout * 9 * Cin), 256, 0, s, w, wr, Cout, Cin);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("rotate_w3x3");
  return SDB_OK;
}

int softmax_rows_backward(const __half* P, __half* dP, long long rows, int cols, long long ld, float scale,
                          cudaStream_t s) {
  if (cols > 4096 || ld > 4096) {
    sdb_set_error("softmax_backward: at most 4096 columns");
    return SDB_ERR_UNSUPPORTED;
  }
  sdb_launch(softmax_bwd_kernel, (unsigned)rows, 128, 0, s, P, dP, cols, ld, scale);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("softmax_bwd");
  return SDB_OK;
}

int add_silu_f32_to_f16(const float* a, const float* b, __half* y, long long n, cudaStream_t s) {
  sdb_launch(add_silu_kernel, ew_grid(n), 256, 0, s, a, b, y, n);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("add_silu");
  return SDB_OK;
}

}  // namespace dense
