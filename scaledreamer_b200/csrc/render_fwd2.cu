// Fused NeRF volume-render forward, v2 (sm_100a): warp-per-ray lattice march, rolled 16-level hash-grid encode staged
// through a per-warp shared-memory tile, density / colour MLPs as 32-sample x 64-unit register tiles, warp-shuffle
// transmittance scans, front-to-back compositing and background blend in one march launch (the environment map is a
// thread-per-ray pre-pass) and, when a tape is supplied,
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
  uint32_t occ[1024];
  float et[kF2Warps][kEncDim * 32];  // per-warp encoding tile, feature-major: et[k][sample]
};

// Environment map (or the random-colour override) for every ray, thread per ray -> comp_rgb_bg.
__global__ void __launch_bounds__(128)
render_bg_kernel(const __grid_constant__ FieldMeta f, const FieldPtrs p, const RayIO io) {
  __shared__ float b1[kBgHidden * kBgEncDim], b2[kBgHidden * kBgHidden], b3[3 * kBgHidden];
  for (int i = threadIdx.x; i < kBgHidden * kBgEncDim; i += blockDim.x) b1[i] = p.bg_w1[i];
  for (int i = threadIdx.x; i < kBgHidden * kBgHidden; i += blockDim.x) b2[i] = p.bg_w2[i];
  for (int i = threadIdx.x; i < 3 * kBgHidden; i += blockDim.x) b3[i] = p.bg_w3[i];
  __syncthreads();
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= io.n_rays) return;
  float bg[3];
  if (io.bg_override) {
    const int img = ray / io.rays_per_image;
#pragma unroll
    for (int c = 0; c < 3; ++c) bg[c] = __ldg(io.bg_override + 3 * img + c);
  } else {
    BgActs a;
    bg_forward(f, reinterpret_cast<const float2*>(p.bg_table), b1, b2, b3, __ldg(io.rays_d + 3 * ray),
               __ldg(io.rays_d + 3 * ray + 1), __ldg(io.rays_d + 3 * ray + 2), a, bg);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) io.comp_rgb_bg[3 * ray + c] = bg[c];
}

constexpr int kRaysPerItem = 4;  // work-stealing granularity: consecutive rays of one image row

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
  for (int i = threadIdx.x; i < occ_words; i += blockDim.x) s.occ[i] = io.occ_bits[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int li = lane >> 3, lj = lane & 7;
  float* et = s.et[warp];
  const int n_items = (io.n_rays + kRaysPerItem - 1) / kRaysPerItem;
  const float2* table = reinterpret_cast<const float2*>(p.table);
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

    for (int r = 0; r < kRaysPerItem; ++r) {
      const int ray = item * kRaysPerItem + r;
      if (ray >= io.n_rays) break;
      Marcher mc;
      mc.init(__ldg(io.rays_o + 3 * ray), __ldg(io.rays_o + 3 * ray + 1), __ldg(io.rays_o + 3 * ray + 2),
              __ldg(io.rays_d + 3 * ray), __ldg(io.rays_d + 3 * ray + 1), __ldg(io.rays_d + 3 * ray + 2),
              io.jitter ? __ldg(io.jitter + ray) : 0.f, m, f.radius);
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
      if (lane == 0) {
        const float fg[3] = {a_r, a_g, a_b};
        const float one_m = 1.f - a_w;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          io.comp_rgb_fg[3 * ray + c] = fg[c];
          io.comp_rgb[3 * ray + c] = fmaf(io.comp_rgb_bg[3 * ray + c], one_m, fg[c]);  // bg: render_bg_kernel
        }
        io.opacity[ray] = a_w;
        io.depth[ray] = a_wt;
        // z-variance (nerf_volume_renderer.py:335-349) in one pass: sum w~ (t - zbar)^2 with w~ = w / clamp(op, 1e-5)
        const float cl = fmaxf(a_w, 1e-5f);
        const float zbar = a_wt / cl;
        const float zv = a_wtt / cl - 2.f * zbar * (a_wt / cl) + zbar * zbar * (a_w / cl);
        io.z_variance[ray] = a_w > 0.5f ? fmaxf(zv, 0.f) : 0.f;
      }
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
  render_bg_kernel<<<(io.n_rays + 127) / 128, 128, 0, stream>>>(f, p, io);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("render_bg");
  const int n_items = (io.n_rays + kRaysPerItem - 1) / kRaysPerItem;
  const int grid = min(kNumSMs * 4, (n_items + kF2Warps - 1) / kF2Warps);
  render_nerf_fwd2_kernel<<<grid, kF2Warps * 32, sizeof(Fwd2Smem), stream>>>(f, p, m, io, t, tape ? 1 : 0);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("render_nerf_fwd2");
  return SDB_OK;
}
