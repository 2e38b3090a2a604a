// Unfused ("packed") NeRF renderer stages for geometries the fused kernels do not cover -- C1's vanilla-MLP
// implicit volume with a frequency encoding, or any plugin geometry evaluated through its own
// forward(points). The reference runs the same stages as separate nerfacc / torch ops
// (threestudio/models/renderers/nerf_volume_renderer.py:139-180 sampling + sigma_fn pruning, :313-373 weights and
// per-ray accumulations); here each stage is one kernel over packed samples sorted by (ray, t):
//   march_count / march_fill      candidate lattice samples in occupied cells (nerfacc traverse_grids, cone_angle 0)
//   packed_visibility             keep = alpha >= min(alpha_thre, mean occ) && T >= early_stop_eps
//   packed_composite fwd / bwd    w = T (1 - exp(-sigma dt)), opacity / depth / colour / z-variance per ray
//   freq_encode                   ProgressiveBandFrequency (threestudio/models/networks.py:16-52)
//   occgrid_update_values         occupancy EMA from densities evaluated by the caller
// One warp per ray; a ray's samples are contiguous, so all sample traffic is coalesced.
#include "../../include/sdb200.h"
#include "field.cuh"

int launch_occ_update(const FieldMeta&, const FieldPtrs&, const int*, const float*, int, int, float, float, float,
                      float*, uint32_t*, float*, cudaStream_t);

namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ void stage_occ(uint32_t* sOcc, const uint32_t* __restrict__ bits, int res) {
  const int words = (res * res * res + 31) / 32;
  for (int i = threadIdx.x; i < words; i += blockDim.x) sOcc[i] = bits[i];
  __syncthreads();
}

// FILL = false: counts[ray] = number of candidates. FILL = true: writes the candidates at offsets[ray]....
template <bool FILL>
__global__ void __launch_bounds__(kWarps * 32)
march_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ jitter,
             int n_rays, const __grid_constant__ MarchMeta m, float radius, const uint32_t* __restrict__ occ_bits,
             int* __restrict__ counts, const long long* __restrict__ offsets, int* __restrict__ ray_idx,
             float* __restrict__ t_start, float* __restrict__ t_end, float* __restrict__ positions) {
  __shared__ uint32_t sOcc[1024];
  stage_occ(sOcc, occ_bits, m.grid_res);
  const int lane = threadIdx.x & 31;
  for (int ray = blockIdx.x * kWarps + (threadIdx.x >> 5); ray < n_rays; ray += gridDim.x * kWarps) {
    const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
    const float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
    const float jit = jitter ? jitter[ray] : 0.f;
    Marcher mar;
    mar.init(ox, oy, oz, dx, dy, dz, jit, m, radius);
    const float near_j = fmaf(jit, m.step, m.near_plane);
    long long slot = FILL ? offsets[ray] : 0;
    int total = 0;
    while (true) {
      int k;
      const int n = mar.next(sOcc, m.grid_res, radius, lane, &k);
      if (n == 0) break;
      if (FILL && lane < n) {
        const long long s = slot + lane;
        const float ts = fmaf((float)k, m.step, near_j), tm = mar.tmid(k);
        ray_idx[s] = ray;
        t_start[s] = ts;
        t_end[s] = ts + m.step;
        positions[3 * s + 0] = fmaf(dx, tm, ox);
        positions[3 * s + 1] = fmaf(dy, tm, oy);
        positions[3 * s + 2] = fmaf(dz, tm, oz);
      }
      slot += n;
      total += n;
    }
    if (!FILL && lane == 0) counts[ray] = total;
  }
}

// Exclusive prefix of sigma*dt along one ray, 32 samples at a time.
struct RayScan {
  float carry;
  __device__ __forceinline__ float excl(float sd, int lane) {
    const float inc = warp_scan_incl(sd, lane);
    const float e = carry + (inc - sd);
    carry += __shfl_sync(kFullMask, inc, 31);
    return e;
  }
};

__global__ void __launch_bounds__(kWarps * 32)
visibility_kernel(const float* __restrict__ sigma, const float* __restrict__ t_start, const float* __restrict__ t_end,
                  const long long* __restrict__ offsets, int n_rays, float alpha_thre,
                  const float* __restrict__ occ_mean, float early_stop_eps, unsigned char* __restrict__ keep,
                  int* __restrict__ kept_counts) {
  const int lane = threadIdx.x & 31;
  const float thre = occ_mean ? fminf(alpha_thre, *occ_mean) : alpha_thre;
  for (int ray = blockIdx.x * kWarps + (threadIdx.x >> 5); ray < n_rays; ray += gridDim.x * kWarps) {
    const long long b = offsets[ray], e = offsets[ray + 1];
    RayScan scan{0.f};
    int kept = 0;
    for (long long s0 = b; s0 < e; s0 += 32) {
      const long long s = s0 + lane;
      const bool in = s < e;
      const float sd = in ? sigma[s] * (t_end[s] - t_start[s]) : 0.f;
      const float T = expf(-scan.excl(sd, lane));
      const float alpha = 1.f - expf(-sd);
      const bool k = in && alpha >= thre && T >= early_stop_eps;
      if (in) keep[s] = k ? 1 : 0;
      kept += __popc(__ballot_sync(kFullMask, k));
    }
    if (lane == 0) kept_counts[ray] = kept;
  }
}

__global__ void __launch_bounds__(kWarps * 32)
composite_fwd_kernel(const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ t_start,
                     const float* __restrict__ t_end, const long long* __restrict__ offsets, int n_rays,
                     float* __restrict__ weights, float* __restrict__ trans, float* __restrict__ opacity,
                     float* __restrict__ depth, float* __restrict__ fg, float* __restrict__ z_variance) {
  const int lane = threadIdx.x & 31;
  for (int ray = blockIdx.x * kWarps + (threadIdx.x >> 5); ray < n_rays; ray += gridDim.x * kWarps) {
    const long long b = offsets[ray], e = offsets[ray + 1];
    RayScan scan{0.f};
    float op = 0.f, dep = 0.f, r = 0.f, g = 0.f, bl = 0.f;
    for (long long s0 = b; s0 < e; s0 += 32) {
      const long long s = s0 + lane;
      const bool in = s < e;
      float ts = 0.f, te = 0.f, sg = 0.f;
      if (in) ts = t_start[s], te = t_end[s], sg = sigma[s];
      const float sd = sg * (te - ts);
      const float T = expf(-scan.excl(sd, lane));
      const float w = T * (1.f - expf(-sd));
      if (in) {
        weights[s] = w;
        trans[s] = T;
        const float tm = 0.5f * (ts + te);
        op += w;
        dep = fmaf(w, tm, dep);
        r = fmaf(w, rgb[3 * s], r);
        g = fmaf(w, rgb[3 * s + 1], g);
        bl = fmaf(w, rgb[3 * s + 2], bl);
      }
    }
    op = warp_sum(op), dep = warp_sum(dep), r = warp_sum(r), g = warp_sum(g), bl = warp_sum(bl);
    // z-variance (:356-373): weights normalised by the clamped opacity, masked where opacity <= 0.5
    const float inv = 1.f / fmaxf(op, 1e-5f);
    const float zmean = dep * inv;
    float zv = 0.f;
    for (long long s = b + lane; s < e; s += 32) {
      const float d = 0.5f * (t_start[s] + t_end[s]) - zmean;
      zv = fmaf(weights[s] * inv, d * d, zv);
    }
    zv = warp_sum(zv);
    if (lane == 0) {
      opacity[ray] = op;
      depth[ray] = dep;
      fg[3 * ray] = r, fg[3 * ray + 1] = g, fg[3 * ray + 2] = bl;
      z_variance[ray] = op > 0.5f ? zv : 0.f;
    }
  }
}

// With G_i = d loss / d w_i = g_op + g_depth t_i + g_fg . rgb_i and x_i = sigma_i dt_i:
//   d w_i / d x_i = T_i - w_i,  d w_j / d x_i = -w_j for j > i
//   => d loss / d x_i = G_i (T_i - w_i) - sum_{j > i} G_j w_j.
__global__ void __launch_bounds__(kWarps * 32)
composite_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ t_start, const float* __restrict__ t_end,
                     const long long* __restrict__ offsets, int n_rays, const float* __restrict__ weights,
                     const float* __restrict__ trans, const float* __restrict__ g_opacity,
                     const float* __restrict__ g_depth, const float* __restrict__ g_fg, float* __restrict__ d_sigma,
                     float* __restrict__ d_rgb) {
  const int lane = threadIdx.x & 31;
  for (int ray = blockIdx.x * kWarps + (threadIdx.x >> 5); ray < n_rays; ray += gridDim.x * kWarps) {
    const long long b = offsets[ray], e = offsets[ray + 1];
    const float go = g_opacity ? g_opacity[ray] : 0.f, gd = g_depth ? g_depth[ray] : 0.f;
    float gr = 0.f, gg = 0.f, gb = 0.f;
    if (g_fg) gr = g_fg[3 * ray], gg = g_fg[3 * ray + 1], gb = g_fg[3 * ray + 2];
    float total = 0.f;
    for (long long s = b + lane; s < e; s += 32) {
      const float tm = 0.5f * (t_start[s] + t_end[s]);
      const float G = go + gd * tm + gr * rgb[3 * s] + gg * rgb[3 * s + 1] + gb * rgb[3 * s + 2];
      total = fmaf(G, weights[s], total);
    }
    total = warp_sum(total);
    float carry = 0.f;  // sum of G_j w_j over samples before this chunk
    for (long long s0 = b; s0 < e; s0 += 32) {
      const long long s = s0 + lane;
      const bool in = s < e;
      float G = 0.f, w = 0.f, T = 0.f, dt = 0.f;
      if (in) {
        const float ts = t_start[s], te = t_end[s];
        dt = te - ts;
        w = weights[s];
        T = trans[s];
        G = go + gd * 0.5f * (ts + te) + gr * rgb[3 * s] + gg * rgb[3 * s + 1] + gb * rgb[3 * s + 2];
      }
      const float gw = G * w;
      const float inc = warp_scan_incl(gw, lane);
      const float suffix = total - (carry + inc);
      carry += __shfl_sync(kFullMask, inc, 31);
      if (in) {
        d_sigma[s] = dt * (G * (T - w) - suffix);
        d_rgb[3 * s] = w * gr, d_rgb[3 * s + 1] = w * gg, d_rgb[3 * s + 2] = w * gb;
      }
    }
  }
}

// out[i, (2 f + fn) * 3 + c] = fn(2^f x[i, c]) * mask[f], fn in (sin, cos); optional leading xyz*2-1 block
// (CompositeEncoding include_xyz, networks.py:170-190); columns beyond the encoding are zero (row padding).
__global__ void __launch_bounds__(256)
freq_encode_kernel(const float* __restrict__ x, long long n, int n_freq, const float* __restrict__ mask,
                   int include_xyz, int stride, float* __restrict__ out) {
  const int lead = include_xyz ? 3 : 0, used = lead + 6 * n_freq;
  const long long total = n * stride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / stride;
    const int col = (int)(i - row * stride);
    float v = 0.f;
    if (col < lead) {
      v = fmaf(x[3 * row + col], 2.f, -1.f);
    } else if (col < used) {
      const int q = col - lead, c = q % 3, ff = q / 3, f = ff >> 1;
      const float a = exp2f((float)f) * x[3 * row + c];
      v = ((ff & 1) ? cosf(a) : sinf(a)) * mask[f];
    }
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256)
occ_values_kernel(const int* __restrict__ cell_idx, const float* __restrict__ values, int n, float decay,
                  float* __restrict__ occs) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int idx = cell_idx[i];
    occs[idx] = fmaxf(occs[idx] * decay, values[i]);
  }
}

int ray_grid(int n_rays) {
  const int want = (n_rays + kWarps - 1) / kWarps;
  return want < kNumSMs * 8 ? (want > 0 ? want : 1) : kNumSMs * 8;
}

int march_meta(const sdb_march_cfg* c, MarchMeta* m) {
  SDB_CHECK_ARG(c && c->render_step_size > 0.f, "march: render_step_size must be > 0");
  SDB_CHECK_ARG(c->grid_resolution >= 1 && c->grid_resolution <= 32, "march: grid_resolution must be in [1,32]");
  m->step = c->render_step_size;
  m->near_plane = c->near_plane;
  m->far_plane = c->far_plane;
  m->prune = c->prune;
  m->alpha_thre = c->alpha_thre;
  m->early_stop_eps = c->early_stop_eps;
  m->grid_res = c->grid_resolution;
  m->output_normal = 0;
  return SDB_OK;
}

}  // namespace

extern "C" {

int sdb_march_count(const sdb_march_cfg* march, float radius, const uint32_t* occ_bits, const float* rays_o,
                    const float* rays_d, const float* jitter, int n_rays, int* counts, void* stream) {
  MarchMeta m;
  int rc = march_meta(march, &m);
  if (rc) return rc;
  SDB_CHECK_ARG(n_rays >= 0 && radius > 0.f, "march_count: bad arguments");
  if (n_rays == 0) return SDB_OK;
  SDB_CHECK_ARG(occ_bits && rays_o && rays_d && counts, "march_count: NULL pointer");
  march_kernel<false><<<ray_grid(n_rays), kWarps * 32, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, jitter, n_rays, m, radius, occ_bits, counts, nullptr, nullptr, nullptr, nullptr, nullptr);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("march_count");
  return SDB_OK;
}

int sdb_march_fill(const sdb_march_cfg* march, float radius, const uint32_t* occ_bits, const float* rays_o,
                   const float* rays_d, const float* jitter, int n_rays, const long long* offsets, int* ray_indices,
                   float* t_starts, float* t_ends, float* positions, void* stream) {
  MarchMeta m;
  int rc = march_meta(march, &m);
  if (rc) return rc;
  SDB_CHECK_ARG(n_rays >= 0 && radius > 0.f, "march_fill: bad arguments");
  if (n_rays == 0) return SDB_OK;
  SDB_CHECK_ARG(occ_bits && rays_o && rays_d && offsets && ray_indices && t_starts && t_ends && positions,
                "march_fill: NULL pointer");
  march_kernel<true><<<ray_grid(n_rays), kWarps * 32, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, jitter, n_rays, m, radius, occ_bits, nullptr, offsets, ray_indices, t_starts, t_ends, positions);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("march_fill");
  return SDB_OK;
}

int sdb_packed_visibility(const float* sigma, const float* t_starts, const float* t_ends, const long long* offsets,
                          int n_rays, float alpha_thre, const float* occ_mean, float early_stop_eps,
                          unsigned char* keep, int* kept_counts, void* stream) {
  SDB_CHECK_ARG(n_rays >= 0, "packed_visibility: bad arguments");
  if (n_rays == 0) return SDB_OK;
  SDB_CHECK_ARG(offsets && kept_counts, "packed_visibility: NULL pointer");
  visibility_kernel<<<ray_grid(n_rays), kWarps * 32, 0, (cudaStream_t)stream>>>(
      sigma, t_starts, t_ends, offsets, n_rays, alpha_thre, occ_mean, early_stop_eps, keep, kept_counts);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("packed_visibility");
  return SDB_OK;
}

int sdb_packed_composite_forward(const float* sigma, const float* rgb, const float* t_starts, const float* t_ends,
                                 const long long* offsets, int n_rays, float* weights, float* trans, float* opacity,
                                 float* depth, float* comp_rgb_fg, float* z_variance, void* stream) {
  SDB_CHECK_ARG(n_rays >= 0, "packed_composite_forward: bad arguments");
  if (n_rays == 0) return SDB_OK;
  SDB_CHECK_ARG(offsets && opacity && depth && comp_rgb_fg && z_variance, "packed_composite_forward: NULL pointer");
  composite_fwd_kernel<<<ray_grid(n_rays), kWarps * 32, 0, (cudaStream_t)stream>>>(
      sigma, rgb, t_starts, t_ends, offsets, n_rays, weights, trans, opacity, depth, comp_rgb_fg, z_variance);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("packed_composite_forward");
  return SDB_OK;
}

int sdb_packed_composite_backward(const float* rgb, const float* t_starts, const float* t_ends,
                                  const long long* offsets, int n_rays, const float* weights, const float* trans,
                                  const float* g_opacity, const float* g_depth, const float* g_comp_rgb_fg,
                                  float* d_sigma, float* d_rgb, void* stream) {
  SDB_CHECK_ARG(n_rays >= 0, "packed_composite_backward: bad arguments");
  if (n_rays == 0) return SDB_OK;
  SDB_CHECK_ARG(offsets, "packed_composite_backward: NULL pointer");
  composite_bwd_kernel<<<ray_grid(n_rays), kWarps * 32, 0, (cudaStream_t)stream>>>(
      rgb, t_starts, t_ends, offsets, n_rays, weights, trans, g_opacity, g_depth, g_comp_rgb_fg, d_sigma, d_rgb);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("packed_composite_backward");
  return SDB_OK;
}

int sdb_freq_encode(const float* x01, long long n, int n_frequencies, const float* mask, int include_xyz,
                    int out_stride, float* out, void* stream) {
  SDB_CHECK_ARG(n >= 0 && n_frequencies >= 1 && n_frequencies <= 32, "freq_encode: bad arguments");
  SDB_CHECK_ARG(out_stride >= (include_xyz ? 3 : 0) + 6 * n_frequencies, "freq_encode: out_stride too small");
  if (n == 0) return SDB_OK;
  SDB_CHECK_ARG(x01 && mask && out, "freq_encode: NULL pointer");
  const long long total = n * out_stride, want = (total + 255) / 256;
  const int grid = (int)(want < (long long)kNumSMs * 8 ? want : (long long)kNumSMs * 8);
  freq_encode_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x01, n, n_frequencies, mask, include_xyz, out_stride, out);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("freq_encode");
  return SDB_OK;
}

int sdb_occgrid_update_values(const int* cell_idx, const float* values, int n_cells, int resolution, float ema_decay,
                              float occ_thre, float* occs, uint32_t* occ_bits, float* occ_mean, void* stream) {
  SDB_CHECK_ARG(resolution >= 1 && resolution <= 32, "occgrid: resolution must be in [1,32]");
  SDB_CHECK_ARG(occs && occ_bits && occ_mean && n_cells >= 0, "occgrid: bad arguments");
  SDB_CHECK_ARG(n_cells == 0 || (cell_idx && values), "occgrid: cell list is NULL");
  if (n_cells > 0) {
    const int grid = (n_cells + 255) / 256 < kNumSMs * 4 ? (n_cells + 255) / 256 : kNumSMs * 4;
    occ_values_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cell_idx, values, n_cells, ema_decay, occs);
    SDB_COUNT_LAUNCH();
    SDB_CHECK_LAUNCH("occ_values");
  }
  FieldMeta fm{};
  FieldPtrs fp{};
  return launch_occ_update(fm, fp, nullptr, nullptr, 0, resolution, 0.f, ema_decay, occ_thre, occs, occ_bits, occ_mean,
                           (cudaStream_t)stream);
}

}  // extern "C"
