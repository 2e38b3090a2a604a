// VolSDF importance renderer pieces of the amortized path (sm_100a), dense [n_rays, S] sample layout, warp per ray:
//   volsdf_coarse_points   proposal samples: stratified uniform lattice between the near / far planes
//   volsdf_resample        proposal SDF -> VolSDF density -> transmittance CDF -> inverse-CDF fine edges -> merged,
//                          sorted interval edges (warp prefix scans + binary searches in shared memory)
//   volsdf_composite_fwd   alpha = |dt| sigma(sdf), T = prod (1 - alpha) (multiplicative warp scan), the five
//                          accumulations + composited normal in one pass
//   volsdf_composite_bwd   analytic backward with the running-prefix form of the suffix sum
// Replaces ImportanceEstimator.sampling (threestudio/models/estimators.py:23-101: nerfacc importance_sampling x2,
// render_transmittance_from_density, torch.sort), volsdf_density / get_alpha (renderers/neus_volume_renderer.py:19-23,
// 93-96) and nerfacc.render_weight_from_alpha + accumulate_along_rays x5 + comp_normal
// (custom/amortized/models/renderers/generative_space_volsdf_volume_renderer.py:356-424).
#include "../../include/sdb200.h"
#include "common.cuh"

namespace {

__device__ __forceinline__ float volsdf_sigma(float sdf, float a) {
  const float sgn = sdf > 0.f ? 1.f : (sdf < 0.f ? -1.f : 0.f);
  return a * (0.5f + 0.5f * sgn * expm1f(-fabsf(sdf) * a));
}

// edge j of the proposal lattice in [0,1]: (j + u) / (n + 1), j = 0..n
__device__ __forceinline__ float lattice(int j, float u, int n) { return ((float)j + u) / (float)(n + 1); }

__global__ void volsdf_coarse_points_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                            const float* __restrict__ u_coarse, int n_rays, int nc, float near_p,
                                            float far_p, float* __restrict__ pts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_rays * nc) return;
  const int ray = (int)(i / nc), j = (int)(i % nc);
  const float u = u_coarse[ray];
  const float t0 = near_p + lattice(j, u, nc) * (far_p - near_p);
  const float t1 = near_p + lattice(j + 1, u, nc) * (far_p - near_p);
  const float tm = 0.5f * (t0 + t1);
#pragma unroll
  for (int c = 0; c < 3; ++c) pts[i * 3 + c] = rays_o[ray * 3 + c] + rays_d[ray * 3 + c] * tm;
}

// number of elements of the sorted array a[0..n) that are < x (strict) or <= x
__device__ __forceinline__ int count_less(const float* a, int n, float x, bool or_equal) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const bool left = or_equal ? a[mid] <= x : a[mid] < x;
    if (left) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int kRsWarps = 4;

__global__ void __launch_bounds__(kRsWarps * 32)
volsdf_resample_kernel(const float* __restrict__ sdf, const float* __restrict__ u_coarse,
                       const float* __restrict__ u_fine, int n_rays, int nc, int nf, float near_p, float far_p,
                       float inv_std, float* __restrict__ t_all) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 2 * (nc + 1) + (nf + 1);
  float* sc = sm + warp * per_warp;   // proposal edges in [0,1]   [nc + 1]
  float* cdf = sc + nc + 1;           // [nc + 1]
  float* sf = cdf + nc + 1;           // fine edges in [0,1]       [nf + 1]
  const int ray = blockIdx.x * kRsWarps + warp;
  if (ray >= n_rays) return;
  const float a = fminf(fmaxf(inv_std, 0.f), 80.f);
  const float uc = u_coarse[ray], uf = u_fine[ray], span = far_p - near_p;
  for (int j = lane; j <= nc; j += 32) sc[j] = lattice(j, uc, nc);
  __syncwarp();
  // transmittance before each proposal interval: exp(-exclusive prefix of sigma dt)
  float carry = 0.f;
  for (int j0 = 0; j0 < nc; j0 += 32) {
    const int j = j0 + lane;
    float sd = 0.f;
    if (j < nc) {
      const float t0 = near_p + sc[j] * span, t1 = near_p + sc[j + 1] * span;
      sd = volsdf_sigma(sdf[(size_t)ray * nc + j], a) * (t1 - t0);
    }
    const float incl = warp_scan_incl(sd, lane);
    if (j < nc) cdf[j] = 1.f - expf(-(carry + incl - sd));
    carry += __shfl_sync(kFullMask, incl, 31);
  }
  if (lane == 0) cdf[nc] = 1.f;
  __syncwarp();
  // inverse CDF at (j + b) / (nf + 1)
  for (int j = lane; j <= nf; j += 32) {
    const float u = lattice(j, uf, nf);
    int idx = count_less(cdf, nc + 1, u, true);  // searchsorted(right=True)
    idx = min(max(idx, 1), nc);
    const float c_lo = cdf[idx - 1], c_hi = cdf[idx];
    float w = (u - c_lo) / fmaxf(c_hi - c_lo, 1e-10f);
    w = fminf(fmaxf(w, 0.f), 1.f);
    sf[j] = sc[idx - 1] + (sc[idx] - sc[idx - 1]) * w;
  }
  __syncwarp();
  // stable merge by rank (proposal edges first on ties)
  float* out = t_all + (size_t)ray * (nc + nf + 2);
  for (int j = lane; j <= nc; j += 32) out[j + count_less(sf, nf + 1, sc[j], false)] = near_p + sc[j] * span;
  for (int j = lane; j <= nf; j += 32) out[j + count_less(sc, nc + 1, sf[j], true)] = near_p + sf[j] * span;
}

__device__ __forceinline__ float warp_scan_prod_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}

constexpr int kCpWarps = 8;

__global__ void __launch_bounds__(kCpWarps * 32)
volsdf_composite_fwd_kernel(const float* __restrict__ sdf, const float* __restrict__ feat,
                            const float* __restrict__ normal, const float* __restrict__ t_mid,
                            const float* __restrict__ delta, int n_rays, int S, float inv_std, int color_act,
                            float* __restrict__ weights, float* __restrict__ opacity, float* __restrict__ depth,
                            float* __restrict__ fg, float* __restrict__ zvar, float* __restrict__ cn) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kCpWarps + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const float a = fminf(fmaxf(inv_std, 0.f), 80.f);
  const size_t base = (size_t)ray * S;
  float T_in = 1.f;
  float a_w = 0.f, a_wt = 0.f, a_wtt = 0.f, a_c[3] = {0.f, 0.f, 0.f}, a_n[3] = {0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < S; j0 += 32) {
    const int j = j0 + lane;
    float alpha = 0.f, t = 0.f;
    if (j < S) {
      alpha = fabsf(delta[base + j]) * volsdf_sigma(sdf[base + j], a);
      t = t_mid[base + j];
    }
    const float incl = warp_scan_prod_incl(1.f - alpha, lane);
    float excl = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) excl = 1.f;
    const float w = T_in * excl * alpha;
    T_in *= __shfl_sync(kFullMask, incl, 31);
    if (j < S) {
      weights[base + j] = w;
      a_w += w;
      a_wt = fmaf(w, t, a_wt);
      a_wtt = fmaf(w * t, t, a_wtt);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float col = 1.f / (1.f + expf(-feat[(base + j) * 3 + c]));
        if (color_act == 1) col = col * 1.002f - 0.001f;
        a_c[c] = fmaf(w, col, a_c[c]);
        a_n[c] = fmaf(w, normal[(base + j) * 3 + c], a_n[c]);
      }
    }
  }
  a_w = warp_sum(a_w);
  a_wt = warp_sum(a_wt);
  a_wtt = warp_sum(a_wtt);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    a_c[c] = warp_sum(a_c[c]);
    a_n[c] = warp_sum(a_n[c]);
  }
  if (lane == 0) {
    opacity[ray] = a_w;
    depth[ray] = a_wt;
    zvar[ray] = a_wtt - a_wt * a_wt * (2.f - a_w);  // sum w (t - depth)^2
    const float inv = 1.f / fmaxf(sqrtf(a_n[0] * a_n[0] + a_n[1] * a_n[1] + a_n[2] * a_n[2]), 1e-12f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      fg[ray * 3 + c] = a_c[c];
      cn[ray * 3 + c] = a_w * (a_n[c] * inv + 1.f) * 0.5f;  // lerp(0, (n + 1) / 2, opacity)
    }
  }
}

__global__ void __launch_bounds__(kCpWarps * 32)
volsdf_composite_bwd_kernel(const float* __restrict__ sdf, const float* __restrict__ feat,
                            const float* __restrict__ t_mid, const float* __restrict__ delta,
                            const float* __restrict__ weights, const float* __restrict__ opacity,
                            const float* __restrict__ depth, const float* __restrict__ fg,
                            const float* __restrict__ g_fg, const float* __restrict__ g_op,
                            const float* __restrict__ g_depth, int n_rays, int S, float inv_std, int color_act,
                            float* __restrict__ d_sdf, float* __restrict__ d_feat) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kCpWarps + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const float a = fminf(fmaxf(inv_std, 0.f), 80.f);
  const size_t base = (size_t)ray * S;
  const float gC0 = g_fg[ray * 3], gC1 = g_fg[ray * 3 + 1], gC2 = g_fg[ray * 3 + 2];
  const float gO = g_op[ray], gD = g_depth[ray];
  // R = sum_k g_k w_k from the saved per-ray outputs
  const float R = fmaf(gC0, fg[ray * 3], fmaf(gC1, fg[ray * 3 + 1], fmaf(gC2, fg[ray * 3 + 2],
                  fmaf(gO, opacity[ray], gD * depth[ray]))));
  float P = 0.f, T_in = 1.f;
  (void)weights;
  for (int j0 = 0; j0 < S; j0 += 32) {
    const int j = j0 + lane;
    float gw = 0.f, g = 0.f, s = 0.f, dl = 0.f, alpha = 0.f, c[3] = {0.f, 0.f, 0.f}, sg[3] = {0.f, 0.f, 0.f};
    const float cs = color_act == 1 ? 1.002f : 1.f, co = color_act == 1 ? -0.001f : 0.f;
    if (j < S) {
      s = sdf[base + j];
      dl = fabsf(delta[base + j]);
      alpha = dl * volsdf_sigma(s, a);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        sg[k] = 1.f / (1.f + expf(-feat[(base + j) * 3 + k]));
        c[k] = fmaf(sg[k], cs, co);
      }
      g = fmaf(gC0, c[0], fmaf(gC1, c[1], fmaf(gC2, c[2], fmaf(gD, t_mid[base + j], gO))));
    }
    // transmittance re-derived with the forward's multiplicative scan
    const float inclT = warp_scan_prod_incl(1.f - alpha, lane);
    float exclT = __shfl_up_sync(kFullMask, inclT, 1);
    if (lane == 0) exclT = 1.f;
    const float T = T_in * exclT;
    T_in *= __shfl_sync(kFullMask, inclT, 31);
    const float w = T * alpha;
    gw = g * w;
    const float incl = warp_scan_incl(gw, lane);
    if (j < S) {
      // dL/dalpha_i = g_i T_i - (sum_{k>i} g_k w_k) / (1 - alpha_i)
      const float one_m = fmaxf(1.f - alpha, 1e-10f);
      const float suffix = R - (P + incl);
      const float dalpha = g * T - suffix / one_m;
      const float dsigma_ds = s == 0.f ? 0.f : -0.5f * a * a * expf(-fabsf(s) * a);
      d_sdf[base + j] = dalpha * dl * dsigma_ds;
      d_feat[(base + j) * 3 + 0] = w * gC0 * cs * sg[0] * (1.f - sg[0]);
      d_feat[(base + j) * 3 + 1] = w * gC1 * cs * sg[1] * (1.f - sg[1]);
      d_feat[(base + j) * 3 + 2] = w * gC2 * cs * sg[2] * (1.f - sg[2]);
    }
    P += __shfl_sync(kFullMask, incl, 31);
  }
}

}  // namespace

extern "C" {

int sdb_volsdf_coarse_points(const float* rays_o, const float* rays_d, const float* u_coarse, int n_rays,
                             int n_coarse, float near_plane, float far_plane, float* points, void* stream) {
  SDB_CHECK_ARG(rays_o && rays_d && u_coarse && points && n_rays >= 0 && n_coarse > 0 && far_plane > near_plane,
                "volsdf_coarse_points: bad arguments");
  if (n_rays == 0) return SDB_OK;
  const long long total = (long long)n_rays * n_coarse;
  volsdf_coarse_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, u_coarse, n_rays, n_coarse, near_plane, far_plane, points);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("volsdf_coarse_points");
  return SDB_OK;
}

int sdb_volsdf_resample(const float* sdf, const float* u_coarse, const float* u_fine, int n_rays, int n_coarse,
                        int n_fine, float near_plane, float far_plane, float inv_std, float* t_all, void* stream) {
  SDB_CHECK_ARG(sdf && u_coarse && u_fine && t_all && n_rays >= 0 && n_coarse > 0 && n_fine > 0 &&
                    far_plane > near_plane, "volsdf_resample: bad arguments");
  SDB_CHECK_ARG(n_coarse <= 2048 && n_fine <= 2048, "volsdf_resample: at most 2048 proposal / fine intervals");
  if (n_rays == 0) return SDB_OK;
  const size_t smem = sizeof(float) * kRsWarps * (2 * (n_coarse + 1) + (n_fine + 1));
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(volsdf_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  volsdf_resample_kernel<<<(n_rays + kRsWarps - 1) / kRsWarps, kRsWarps * 32, smem, (cudaStream_t)stream>>>(
      sdf, u_coarse, u_fine, n_rays, n_coarse, n_fine, near_plane, far_plane, inv_std, t_all);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("volsdf_resample");
  return SDB_OK;
}

int sdb_volsdf_composite_forward(const float* sdf, const float* features, const float* normal, const float* t_mid,
                                 const float* delta, int n_rays, int n_samples, float inv_std, int color_activation,
                                 float* weights, float* opacity, float* depth, float* comp_rgb_fg, float* z_variance,
                                 float* comp_normal, void* stream) {
  SDB_CHECK_ARG(sdf && features && normal && t_mid && delta && weights && opacity && depth && comp_rgb_fg &&
                    z_variance && comp_normal && n_rays >= 0 && n_samples > 0, "volsdf_composite_forward: bad arguments");
  if (n_rays == 0) return SDB_OK;
  volsdf_composite_fwd_kernel<<<(n_rays + kCpWarps - 1) / kCpWarps, kCpWarps * 32, 0, (cudaStream_t)stream>>>(
      sdf, features, normal, t_mid, delta, n_rays, n_samples, inv_std, color_activation, weights, opacity, depth, comp_rgb_fg,
      z_variance, comp_normal);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("volsdf_composite_forward");
  return SDB_OK;
}

int sdb_volsdf_composite_backward(const float* sdf, const float* features, const float* t_mid, const float* delta,
                                  const float* weights, const float* opacity, const float* depth,
                                  const float* comp_rgb_fg, const float* g_comp_rgb_fg, const float* g_opacity,
                                  const float* g_depth, int n_rays, int n_samples, float inv_std, int color_activation,
                                  float* d_sdf, float* d_features, void* stream) {
  SDB_CHECK_ARG(sdf && features && t_mid && delta && weights && opacity && depth && comp_rgb_fg && g_comp_rgb_fg &&
                    g_opacity && g_depth && d_sdf && d_features && n_rays >= 0 && n_samples > 0,
                "volsdf_composite_backward: bad arguments");
  if (n_rays == 0) return SDB_OK;
  volsdf_composite_bwd_kernel<<<(n_rays + kCpWarps - 1) / kCpWarps, kCpWarps * 32, 0, (cudaStream_t)stream>>>(
      sdf, features, t_mid, delta, weights, opacity, depth, comp_rgb_fg, g_comp_rgb_fg, g_opacity, g_depth, n_rays,
      n_samples, inv_std, color_activation, d_sdf, d_features);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("volsdf_composite_backward");
  return SDB_OK;
}

}  // extern "C"
