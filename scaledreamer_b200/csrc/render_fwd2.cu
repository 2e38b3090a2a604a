// Fused NeRF volume-render forward, v2 (sm_100a): warp-per-ray lattice march, rolled 16-level hash-grid encode staged
// through a per-warp shared-memory tile, density / colour MLPs as 32-sample x 64-unit register tiles, warp-shuffle
// transmittance scans, front-to-back compositing, background blend -- ONE launch -- and, when a tape is supplied,
// a record of every kept sample for the tape-based backward (render_bwd2.cu).
//
// Replaces the same reference lines as render_fwd.cu (threestudio/models/renderers/nerf_volume_renderer.py:118-428:
// nerfacc sampling + sigma_fn pass :153-180, geometry :282-284, material :285-291, background :292,
// render_weight_from_density :313-319, accumulate_along_rays :324-348, blend :364).
#include "render_tape.cuh"

namespace {

constexpr int kF2Warps = 4;

struct Fwd2Smem {
  float wpd[kWpSize];  // density W1, permuted (stage_w1_perm)
  float wpf[kWpSize];  // feature W1, permuted
  float w2d[kHidden];
  float w2f[3 * kHidden];
  float b1[kBgHidden * kBgEncDim];
  float b2[kBgHidden * kBgHidden];
  float b3[3 * kBgHidden];
  uint32_t occ[1024];
  float et[kF2Warps][kEncDim * 32];  // per-warp encoding tile, feature-major: et[k][sample]
};

// Encodes this lane's point through all levels straight into column `lane` of the warp's tile.
__device__ __forceinline__ void encode_to_tile(const float2* __restrict__ table, const GridMeta& gm, float x, float y,
                                               float z, bool valid, float* __restrict__ et_lane) {
#pragma unroll 4
  for (int l = 0; l < kMaxLevels; ++l) {
    float ax = 0.f, ay = 0.f;
    if (valid) {
      const uint32_t res = gm.res[l], size = gm.size[l], hashed = gm.hashed[l];
      const float2* tl = table + gm.offset[l];
      const LevelCell c = level_cell(gm.scale[l], x, y, z);
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        v[k] = __ldg(tl + grid_index(hashed, res, size, c.ix + (k & 1), c.iy + ((k >> 1) & 1), c.iz + ((k >> 2) & 1)));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float w = corner_weight(c, k);
        ax = fmaf(w, v[k].x, ax);
        ay = fmaf(w, v[k].y, ay);
      }
    }
    et_lane[(2 * l) * 32] = ax;
    et_lane[(2 * l + 1) * 32] = ay;
  }
}

__global__ void __launch_bounds__(kF2Warps * 32, 4)
render_nerf_fwd2_kernel(const __grid_constant__ FieldMeta f, const FieldPtrs p, const __grid_constant__ MarchMeta m,
                        const RayIO io, const RenderTape tape, const int has_tape) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Fwd2Smem& s = *reinterpret_cast<Fwd2Smem*>(smem_raw);
  const int occ_words = (m.grid_res * m.grid_res * m.grid_res + 31) / 32;
  stage_w1_perm(s.wpd, p.w1d, threadIdx.x, blockDim.x);
  stage_w1_perm(s.wpf, p.w1f, threadIdx.x, blockDim.x);
  for (int i = threadIdx.x; i < kHidden; i += blockDim.x) s.w2d[i] = p.w2d[i];
  for (int i = threadIdx.x; i < 3 * kHidden; i += blockDim.x) s.w2f[i] = p.w2f[i];
  for (int i = threadIdx.x; i < kBgHidden * kBgEncDim; i += blockDim.x) s.b1[i] = p.bg_w1[i];
  for (int i = threadIdx.x; i < kBgHidden * kBgHidden; i += blockDim.x) s.b2[i] = p.bg_w2[i];
  for (int i = threadIdx.x; i < 3 * kBgHidden; i += blockDim.x) s.b3[i] = p.bg_w3[i];
  for (int i = threadIdx.x; i < occ_words; i += blockDim.x) s.occ[i] = io.occ_bits[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int li = lane >> 3, lj = lane & 7;
  float* et = s.et[warp];
  const int n_items = (io.n_rays + 31) / 32;
  const float2* table = reinterpret_cast<const float2*>(p.table);
  const float2* bg_table = reinterpret_cast<const float2*>(p.bg_table);
  const float inv2r = 0.5f / f.radius;
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
      if (io.bg_override) {
        const int img = my_ray / io.rays_per_image;
#pragma unroll
        for (int c = 0; c < 3; ++c) mbg[c] = __ldg(io.bg_override + 3 * img + c);
      } else {
        BgActs a;
        bg_forward(f, bg_table, s.b1, s.b2, s.b3, md[0], md[1], md[2], a, mbg);
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
      int nchunk = 0;
      while (true) {
        int my_k;
        const int filled = mc.next(s.occ, m.grid_res, f.radius, lane, &my_k);
        if (filled == 0) break;
        const bool valid = lane < filled;
        float tm = 0.f, ts = 0.f, te = 0.f, px = 0.f, py = 0.f, pz = 0.f;
        if (valid) {
          tm = mc.tmid(my_k);
          ts = fmaf((float)my_k, mc.step, mc.c0 - 0.5f * mc.step);
          te = ts + mc.step;
          px = fmaf(mc.dx, tm, mc.ox);
          py = fmaf(mc.dy, tm, mc.oy);
          pz = fmaf(mc.dz, tm, mc.oz);
        }
        const float x01 = (px + f.radius) * inv2r, y01 = (py + f.radius) * inv2r, z01 = (pz + f.radius) * inv2r;
        __syncwarp();  // the previous chunk's readers are done with the tile
        encode_to_tile(table, f.grid, x01, y01, z01, valid, et + lane);
        __syncwarp();

        // density MLP: hidden tile, then second layer + reduce-scatter so lane l holds sample l
        float acc[8][8];
        hidden_tile(et, 32, s.wpd, li, lj, acc);
        float part[8];
        {
          float w2[8];
#pragma unroll
          for (int b = 0; b < 8; ++b) w2[b] = s.w2d[hidden_of(lj, b)];
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            float sum = 0.f;
#pragma unroll
            for (int b = 0; b < 8; ++b) sum = fmaf(w2[b], fmaxf(acc[a][b], 0.f), sum);
            part[a] = sum;
          }
        }
        const float raw_mlp = reduce_scatter8(part, lj);
        float raw = 0.f, sigma = 0.f, sd = 0.f;
        if (valid) {
          raw = raw_mlp + density_bias(f, px, py, pz);
          sigma = density_activation(f.density_act, raw);
          sd = sigma * (te - ts);
        }
        const float incl = warp_scan_incl(sd, lane);
        const float T_all = expf(-(S_all + (incl - sd)));
        const float e_sd = expf(-sd);
        const float alpha = 1.f - e_sd;
        const bool vis = valid && (!m.prune || (alpha >= thre && T_all >= eps_T));
        const float sdk = vis ? sd : 0.f;
        const float inclk = warp_scan_incl(sdk, lane);
        const float T = expf(-(S_kept + (inclk - sdk)));
        const float w = vis ? T * alpha : 0.f;
        S_all += __shfl_sync(kFullMask, incl, 31);
        S_kept += __shfl_sync(kFullMask, inclk, 31);
        const uint32_t vm = __ballot_sync(kFullMask, vis);
        if (vm) {
          // colour MLP on the whole tile (pruned lanes ride along)
          hidden_tile(et, 32, s.wpf, li, lj, acc);
          float o[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float w2[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) w2[b] = s.w2f[c * kHidden + hidden_of(lj, b)];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
              float sum = 0.f;
#pragma unroll
              for (int b = 0; b < 8; ++b) sum = fmaf(w2[b], fmaxf(acc[a][b], 0.f), sum);
              part[a] = sum;
            }
            o[c] = reduce_scatter8(part, lj);
          }
          if (vis) {
            const float c0 = color_activation(f.color_act, o[0]), c1 = color_activation(f.color_act, o[1]),
                        c2 = color_activation(f.color_act, o[2]);
            a_w += w;
            a_wt = fmaf(w, tm, a_wt);
            a_wtt = fmaf(w * tm, tm, a_wtt);
            a_r = fmaf(w, c0, a_r);
            a_g = fmaf(w, c1, a_g);
            a_b = fmaf(w, c2, a_b);
          }
          if (has_tape) {
            const int n = __popc(vm);
            int slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(tape.counter, n);
            slot0 = __shfl_sync(kFullMask, slot0, 0);
            if (slot0 + n > tape.capacity || nchunk >= tape.max_chunks) {
              if (lane == 0) atomicExch(tape.counter + 1, 1);  // overflow: the host must enlarge the tape
            } else {
              if (vis) {
                const int slot = slot0 + __popc(vm & ((1u << lane) - 1u));
                const size_t cap = (size_t)tape.capacity;
                tape.pos[slot] = x01;
                tape.pos[cap + slot] = y01;
                tape.pos[2 * cap + slot] = z01;
                float* sp = tape.sample + slot;
                sp[0] = raw;
                sp[cap] = o[0];
                sp[2 * cap] = o[1];
                sp[3 * cap] = o[2];
                sp[4 * cap] = w;
                sp[5 * cap] = T * e_sd;
                sp[6 * cap] = tm;
                sp[7 * cap] = te - ts;
                float* ep = tape.enc + (size_t)(slot >> 5) * (kEncDim * 32) + (slot & 31);
#pragma unroll 8
                for (int k = 0; k < kEncDim; ++k) ep[k * 32] = et[k * 32 + lane];
              }
              if (lane == 0)
                tape.ray_chunks[(size_t)ray * tape.max_chunks + nchunk] = ((uint32_t)slot0 << 5) | (uint32_t)(n - 1);
              ++nchunk;
            }
          }
        }
        if (m.prune && expf(-S_all) < eps_T) break;  // every later candidate fails the T test
      }
      if (has_tape && lane == 0) tape.ray_nchunks[ray] = nchunk;
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

}  // namespace

int launch_render_fwd2(const FieldMeta& f, const FieldPtrs& p, const MarchMeta& m, const RayIO& io,
                       const RenderTape* tape, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(render_nerf_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(Fwd2Smem));
    if (e != cudaSuccess) {
      sdb_set_error("render_nerf_fwd2: smem attribute: %s", cudaGetErrorString(e));
      return SDB_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaMemsetAsync(io.work_counter, 0, sizeof(int), stream);
  RenderTape t;
  memset(&t, 0, sizeof(t));
  if (tape) {
    t = *tape;
    cudaMemsetAsync(t.counter, 0, 2 * sizeof(int), stream);
  }
  const int n_items = (io.n_rays + 31) / 32;
  const int grid = min(kNumSMs * 4, (n_items + kF2Warps - 1) / kF2Warps);
  render_nerf_fwd2_kernel<<<grid, kF2Warps * 32, sizeof(Fwd2Smem), stream>>>(f, p, m, io, t, tape ? 1 : 0);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("render_nerf_fwd2");
  return SDB_OK;
}
