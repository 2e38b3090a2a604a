// Fused NeRF volume-render forward for sm_100a: ray march over the occupancy lattice + hash-grid encode +
// density / colour MLPs + visibility pruning + front-to-back compositing + background blend, ONE launch.
//
// Replaces, on the reference's hot path (threestudio/models/renderers/nerf_volume_renderer.py:118-428):
//   nerfacc.OccGridEstimator.sampling + sigma_fn pass (:153-180), geometry(...) (:282-284),
//   material (:285-291), background (:292), render_weight_from_density (:313-319),
//   accumulate_along_rays x5 (:324-348), comp_rgb blend (:364).
//
// Work decomposition: one warp owns a bundle of 32 rays (strided across the image so every bundle mixes
// heavy and empty rays). Backgrounds are evaluated thread-per-ray; the march is warp-per-ray with the 32
// lanes holding 32 consecutive *occupied* lattice samples, so the transmittance scan is a warp shuffle scan.
#include "field.cuh"

namespace {

constexpr int kFwdWarps = 8;

struct FwdSmem {
  float w1d[kHidden * kEncDim];
  float w2d[kHidden];
  float w1f[kHidden * kEncDim];
  float w2f[3 * kHidden];
  float b1[kBgHidden * kBgEncDim];
  float b2[kBgHidden * kBgHidden];
  float b3[3 * kBgHidden];
  uint32_t occ[1024];
};

__device__ __forceinline__ void load_weights(FwdSmem& s, const FieldPtrs& p, const uint32_t* occ_bits, int occ_words) {
  for (int i = threadIdx.x; i < kHidden * kEncDim; i += blockDim.x) {
    s.w1d[i] = p.w1d[i];
    s.w1f[i] = p.w1f[i];
  }
  for (int i = threadIdx.x; i < kHidden; i += blockDim.x) s.w2d[i] = p.w2d[i];
  for (int i = threadIdx.x; i < 3 * kHidden; i += blockDim.x) s.w2f[i] = p.w2f[i];
  for (int i = threadIdx.x; i < kBgHidden * kBgEncDim; i += blockDim.x) s.b1[i] = p.bg_w1[i];
  for (int i = threadIdx.x; i < kBgHidden * kBgHidden; i += blockDim.x) s.b2[i] = p.bg_w2[i];
  for (int i = threadIdx.x; i < 3 * kBgHidden; i += blockDim.x) s.b3[i] = p.bg_w3[i];
  for (int i = threadIdx.x; i < occ_words; i += blockDim.x) s.occ[i] = occ_bits[i];
}

__global__ void __launch_bounds__(kFwdWarps * 32, 2)
render_nerf_fwd_kernel(const __grid_constant__ FieldMeta f, const FieldPtrs p, const __grid_constant__ MarchMeta m,
                       const RayIO io, const PackedOut pk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& s = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int occ_words = (m.grid_res * m.grid_res * m.grid_res + 31) / 32;
  load_weights(s, p, io.occ_bits, occ_words);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int n_items = (io.n_rays + 31) / 32;
  const float2* table = reinterpret_cast<const float2*>(p.table);
  const float2* bg_table = reinterpret_cast<const float2*>(p.bg_table);
  float thre = 0.f, eps_T = 0.f;
  if (m.prune) {
    thre = m.alpha_thre;
    if (io.occ_mean) thre = fminf(thre, __ldg(io.occ_mean));
    eps_T = m.early_stop_eps;
  }

  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(io.work_counter, 1);
    item = __shfl_sync(kFullMask, item, 0);
    if (item >= n_items) break;

    // ---- phase 1: per-lane ray setup + background (thread per ray) ----
    const int my_ray = lane * n_items + item;
    const bool my_valid = my_ray < io.n_rays;
    float mo[3] = {0.f, 0.f, 0.f}, md[3] = {0.f, 0.f, 1.f}, mjit = 0.f, mbg[3] = {0.f, 0.f, 0.f};
    if (my_valid) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mo[c] = __ldg(io.rays_o + 3 * my_ray + c);
        md[c] = __ldg(io.rays_d + 3 * my_ray + c);
      }
      if (io.jitter) mjit = __ldg(io.jitter + my_ray);
      BgActs a;
      bg_forward(f, bg_table, s.b1, s.b2, s.b3, md[0], md[1], md[2], a, mbg);
      if (io.bg_override) {
        const int img = my_ray / io.rays_per_image;
#pragma unroll
        for (int c = 0; c < 3; ++c) mbg[c] = __ldg(io.bg_override + 3 * img + c);
      }
    }
    float r_op = 0.f, r_depth = 0.f, r_tt = 0.f, r_fg[3] = {0.f, 0.f, 0.f};

    // ---- phase 2: warp-per-ray march over the bundle ----
    for (int r = 0; r < 32; ++r) {
      const int ray = r * n_items + item;
      if (ray >= io.n_rays) break;
      Marcher mc;
      mc.init(__shfl_sync(kFullMask, mo[0], r), __shfl_sync(kFullMask, mo[1], r), __shfl_sync(kFullMask, mo[2], r),
              __shfl_sync(kFullMask, md[0], r), __shfl_sync(kFullMask, md[1], r), __shfl_sync(kFullMask, md[2], r),
              __shfl_sync(kFullMask, mjit, r), m, f.radius);
      float S_all = 0.f, S_kept = 0.f;
      float a_w = 0.f, a_wt = 0.f, a_wtt = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f;
      while (true) {
        int my_k;
        const int filled = mc.next(s.occ, m.grid_res, f.radius, lane, &my_k);
        if (filled == 0) break;
        const bool valid = lane < filled;
        float enc[kEncDim];
        float sigma = 0.f, raw = 0.f, tm = 0.f, ts = 0.f, te = 0.f, px = 0.f, py = 0.f, pz = 0.f, sd = 0.f;
        if (valid) {
          tm = mc.tmid(my_k);
          ts = fmaf((float)my_k, mc.step, mc.c0 - 0.5f * mc.step);
          te = ts + mc.step;
          px = fmaf(mc.dx, tm, mc.ox);
          py = fmaf(mc.dy, tm, mc.oy);
          pz = fmaf(mc.dz, tm, mc.oz);
          sigma = field_density(f, table, s.w1d, s.w2d, px, py, pz, enc, &raw);
          sd = sigma * (te - ts);
        }
        const float incl = warp_scan_incl(sd, lane);
        const float T_all = expf(-(S_all + (incl - sd)));
        const float alpha = 1.f - expf(-sd);
        const bool vis = valid && (!m.prune || (alpha >= thre && T_all >= eps_T));
        const float sdk = vis ? sd : 0.f;
        const float inclk = warp_scan_incl(sdk, lane);
        const float T = expf(-(S_kept + (inclk - sdk)));
        const float w = vis ? T * alpha : 0.f;
        S_all += __shfl_sync(kFullMask, incl, 31);
        S_kept += __shfl_sync(kFullMask, inclk, 31);
        float rgb[3] = {0.f, 0.f, 0.f};
        if (vis) {
          float feat[3];
          mlp32_64_3(s.w1f, s.w2f, enc, feat);
#pragma unroll
          for (int c = 0; c < 3; ++c) rgb[c] = color_activation(f.color_act, feat[c]);
          a_w += w;
          a_wt = fmaf(w, tm, a_wt);
          a_wtt = fmaf(w * tm, tm, a_wtt);
          a_r = fmaf(w, rgb[0], a_r);
          a_g = fmaf(w, rgb[1], a_g);
          a_b = fmaf(w, rgb[2], a_b);
        }
        if (pk.counter) {
          const uint32_t vm = __ballot_sync(kFullMask, vis);
          int slot0 = 0;
          if (lane == 0 && vm) slot0 = atomicAdd(pk.counter, __popc(vm));
          slot0 = __shfl_sync(kFullMask, slot0, 0);
          const int slot = slot0 + __popc(vm & ((1u << lane) - 1u));
          if (vis && slot < pk.capacity) {
            pk.ray_idx[slot] = ray;
            pk.t_start[slot] = ts;
            pk.t_end[slot] = te;
            pk.weight[slot] = w;
            pk.density[slot] = sigma;
#pragma unroll
            for (int c = 0; c < 3; ++c) pk.rgb[3 * slot + c] = rgb[c];
            if (m.output_normal && pk.normal) {
              float n3[3];
              field_fd_normal(f, table, s.w1d, s.w2d, px, py, pz, sigma, n3);
#pragma unroll
              for (int c = 0; c < 3; ++c) pk.normal[3 * slot + c] = n3[c];
            }
          }
        }
        if (m.prune && expf(-S_all) < eps_T) break;  // every later candidate fails the T test
      }
      a_w = warp_sum(a_w);
      a_wt = warp_sum(a_wt);
      a_wtt = warp_sum(a_wtt);
      a_r = warp_sum(a_r);
      a_g = warp_sum(a_g);
      a_b = warp_sum(a_b);
      if (lane == r) {
        r_op = a_w;
        r_depth = a_wt;
        r_tt = a_wtt;
        r_fg[0] = a_r;
        r_fg[1] = a_g;
        r_fg[2] = a_b;
      }
    }

    // ---- phase 3: per-lane outputs ----
    if (my_valid) {
      const float one_m = 1.f - r_op;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        io.comp_rgb_fg[3 * my_ray + c] = r_fg[c];
        io.comp_rgb_bg[3 * my_ray + c] = mbg[c];
        io.comp_rgb[3 * my_ray + c] = fmaf(mbg[c], one_m, r_fg[c]);
      }
      io.opacity[my_ray] = r_op;
      io.depth[my_ray] = r_depth;
      // z-variance (nerf_volume_renderer.py:335-349) in one pass: sum w~ (t - zbar)^2 with w~ = w / clamp(op, 1e-5)
      const float cl = fmaxf(r_op, 1e-5f);
      const float zbar = r_depth / cl;
      const float zv = r_tt / cl - 2.f * zbar * (r_depth / cl) + zbar * zbar * (r_op / cl);
      io.z_variance[my_ray] = r_op > 0.5f ? fmaxf(zv, 0.f) : 0.f;
    }
  }
}

// Stand-alone field evaluation, thread per point (geometry.forward / forward_density outside the renderer and
// the occupancy-grid refresh). out_features / out_normal may be null.
__global__ void __launch_bounds__(256)
field_eval_kernel(const __grid_constant__ FieldMeta f, const FieldPtrs p, const float* __restrict__ points, int n,
                  float* __restrict__ out_density, float* __restrict__ out_features, float* __restrict__ out_normal) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* w1d = reinterpret_cast<float*>(smem_raw);
  float* w2d = w1d + kHidden * kEncDim;
  float* w1f = w2d + kHidden;
  float* w2f = w1f + kHidden * kEncDim;
  for (int i = threadIdx.x; i < kHidden * kEncDim; i += blockDim.x) {
    w1d[i] = p.w1d[i];
    w1f[i] = p.w1f[i];
  }
  for (int i = threadIdx.x; i < kHidden; i += blockDim.x) w2d[i] = p.w2d[i];
  for (int i = threadIdx.x; i < 3 * kHidden; i += blockDim.x) w2f[i] = p.w2f[i];
  __syncthreads();
  const float2* table = reinterpret_cast<const float2*>(p.table);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float px = points[3 * i], py = points[3 * i + 1], pz = points[3 * i + 2];
    float enc[kEncDim], raw;
    const float sigma = field_density(f, table, w1d, w2d, px, py, pz, enc, &raw);
    out_density[i] = sigma;
    if (out_features) {
      float feat[3];
      mlp32_64_3(w1f, w2f, enc, feat);
#pragma unroll
      for (int c = 0; c < 3; ++c) out_features[3 * i + c] = feat[c];
    }
    if (out_normal) {
      float n3[3];
      field_fd_normal(f, table, w1d, w2d, px, py, pz, sigma, n3);
#pragma unroll
      for (int c = 0; c < 3; ++c) out_normal[3 * i + c] = n3[c];
    }
  }
}

// Stand-alone hash-grid encode (KATs against the oracle), thread per point.
__global__ void __launch_bounds__(256)
hashgrid_fwd_kernel(const __grid_constant__ GridMeta gm, const float* __restrict__ table,
                    const float* __restrict__ x01, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float enc[kEncDim];
#pragma unroll
  for (int k = 0; k < kEncDim; ++k) enc[k] = 0.f;
  const float2* t = reinterpret_cast<const float2*>(table);
  if (gm.n_levels == 16)
    grid_encode<16>(t, gm, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], enc);
  else
    grid_encode<4>(t, gm, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], enc);
  for (int k = 0; k < 2 * gm.n_levels; ++k) out[(size_t)i * 2 * gm.n_levels + k] = enc[k];
}

__global__ void __launch_bounds__(256)
hashgrid_bwd_kernel(const __grid_constant__ GridMeta gm, const float* __restrict__ x01,
                    const float* __restrict__ g_out, int n, float* __restrict__ g_table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g[kEncDim];
#pragma unroll
  for (int k = 0; k < kEncDim; ++k) g[k] = 0.f;
  for (int k = 0; k < 2 * gm.n_levels; ++k) g[k] = g_out[(size_t)i * 2 * gm.n_levels + k];
  float2* t = reinterpret_cast<float2*>(g_table);
  if (gm.n_levels == 16)
    grid_scatter<16>(t, gm, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], g);
  else
    grid_scatter<4>(t, gm, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], g);
}

// Occupancy refresh (nerfacc OccGridEstimator.update_every_n_steps semantics):
//   occs[idx] = max(occs[idx] * decay, sigma(x_idx) * step) for the sampled cells, then
//   binaries = occs > min(mean(occs), occ_thre), packed into a bit-field; mean kept on device.
__global__ void __launch_bounds__(256)
occ_update_kernel(const __grid_constant__ FieldMeta f, const FieldPtrs p, const int* __restrict__ cell_idx,
                  const float* __restrict__ cell_rand, int n_cells, int res, float step, float decay,
                  float* __restrict__ occs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* w1d = reinterpret_cast<float*>(smem_raw);
  float* w2d = w1d + kHidden * kEncDim;
  for (int i = threadIdx.x; i < kHidden * kEncDim; i += blockDim.x) w1d[i] = p.w1d[i];
  for (int i = threadIdx.x; i < kHidden; i += blockDim.x) w2d[i] = p.w2d[i];
  __syncthreads();
  const float2* table = reinterpret_cast<const float2*>(p.table);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += gridDim.x * blockDim.x) {
    const int idx = cell_idx[i];
    const int cz = idx % res, cy = (idx / res) % res, cx = idx / (res * res);
    const float cell = 2.f * f.radius / (float)res;
    const float px = -f.radius + ((float)cx + cell_rand[3 * i + 0]) * cell;
    const float py = -f.radius + ((float)cy + cell_rand[3 * i + 1]) * cell;
    const float pz = -f.radius + ((float)cz + cell_rand[3 * i + 2]) * cell;
    float enc[kEncDim], raw;
    const float sigma = field_density(f, table, w1d, w2d, px, py, pz, enc, &raw);
    occs[idx] = fmaxf(occs[idx] * decay, sigma * step);
  }
}

__global__ void __launch_bounds__(1024)
occ_binarize_kernel(const float* __restrict__ occs, int n, float occ_thre, uint32_t* __restrict__ bits,
                    float* __restrict__ mean_out) {
  __shared__ float red[32];
  __shared__ float s_thr;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += occs[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      const float mean = v / (float)n;
      *mean_out = mean;
      s_thr = fminf(mean, occ_thre);
    }
  }
  __syncthreads();
  const float thr = s_thr;
  for (int w = threadIdx.x; w < (n + 31) / 32; w += blockDim.x) {
    uint32_t word = 0u;
    for (int b = 0; b < 32; ++b) {
      const int i = w * 32 + b;
      if (i < n && occs[i] > thr) word |= (1u << b);
    }
    bits[w] = word;
  }
}

}  // namespace

// ---------------------------------------------------------------- host wrappers (called from capi.cu)
int launch_render_fwd(const FieldMeta& f, const FieldPtrs& p, const MarchMeta& m, const RayIO& io,
                      const PackedOut& pk, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(render_nerf_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem));
    attr_set = true;
  }
  cudaMemsetAsync(io.work_counter, 0, sizeof(int), stream);
  if (pk.counter) cudaMemsetAsync(pk.counter, 0, sizeof(int), stream);
  const int n_items = (io.n_rays + 31) / 32;
  const int max_ctas = kNumSMs * 2;
  const int grid = min(max_ctas, (n_items + kFwdWarps - 1) / kFwdWarps);
  render_nerf_fwd_kernel<<<grid, kFwdWarps * 32, sizeof(FwdSmem), stream>>>(f, p, m, io, pk);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("render_nerf_fwd");
  return SDB_OK;
}

int launch_field_eval(const FieldMeta& f, const FieldPtrs& p, const float* points, int n, float* density,
                      float* features, float* normal, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (2 * kHidden * kEncDim + 4 * kHidden);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(field_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const int grid = max(1, min(kNumSMs * 8, (n + 255) / 256));
  field_eval_kernel<<<grid, 256, smem, stream>>>(f, p, points, n, density, features, normal);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("field_eval");
  return SDB_OK;
}

int launch_hashgrid_fwd(const GridMeta& gm, const float* table, const float* x01, int n, float* out,
                        cudaStream_t stream) {
  hashgrid_fwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gm, table, x01, n, out);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("hashgrid_fwd");
  return SDB_OK;
}

int launch_hashgrid_bwd(const GridMeta& gm, const float* x01, const float* g_out, int n, float* g_table,
                        cudaStream_t stream) {
  hashgrid_bwd_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gm, x01, g_out, n, g_table);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("hashgrid_bwd");
  return SDB_OK;
}

int launch_occ_update(const FieldMeta& f, const FieldPtrs& p, const int* cell_idx, const float* cell_rand,
                      int n_cells, int res, float step, float decay, float occ_thre, float* occs, uint32_t* bits,
                      float* mean_out, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (kHidden * kEncDim + kHidden);
  if (n_cells > 0) {
    const int grid = max(1, min(kNumSMs * 4, (n_cells + 255) / 256));
    occ_update_kernel<<<grid, 256, smem, stream>>>(f, p, cell_idx, cell_rand, n_cells, res, step, decay, occs);
    SDB_COUNT_LAUNCH();
    SDB_CHECK_LAUNCH("occ_update");
  }
  occ_binarize_kernel<<<1, 1024, 0, stream>>>(occs, res * res * res, occ_thre, bits, mean_out);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("occ_binarize");
  return SDB_OK;
}
