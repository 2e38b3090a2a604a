// Fused NeRF volume-render backward for sm_100a (recompute-based, ONE launch).
//
// Replaces the autograd chain the reference runs after nerf_volume_renderer.py:118-428:
//   backward of accumulate_along_rays / render_weight_from_density (nerfacc autograd),
//   of the colour + density MLPs (cuBLAS) and of the tcnn hash grid (atomic scatter),
//   plus the background network.
//
// Each ray is re-marched front to back exactly as in the forward. With the per-ray forward outputs saved,
// the suffix sum that the compositing gradient needs collapses to a running prefix:
//   dL/dsigma_i = delta_i * ( g_i (T_i - w_i) - (R - P_i) ),   g_i = dL/dw_i,
//   R = <gC, C_fg> + gO * opacity + gD * depth,  P_i = sum_{j<=i} g_j w_j.
// MLP weight gradients: per-warp smem tiles hold hidden activations / hidden grads of the 32 samples of a
// chunk; lane l owns input column l of dW1 (128 register accumulators) so the tile reads are conflict free.
#include "field.cuh"

namespace {

constexpr int kBwdWarps = 8;
constexpr int kCat = 2 * kHidden;     // concatenated hidden units: [0,64) density, [64,128) feature
constexpr int kTileStride = kCat + 4;  // 132: keeps float4 alignment, spreads rows over 16-byte bank groups
constexpr int kEncStride = 33;

struct BwdSmem {
  float w1[kCat * kEncDim];  // rows 0..63 = density W1, rows 64..127 = feature W1
  float w2d[kHidden];
  float w2f[3 * kHidden];
  float b1[kBgHidden * kBgEncDim];
  float b2[kBgHidden * kBgHidden];
  float b3[3 * kBgHidden];
  uint32_t occ[1024];
  float gW1[kCat * kEncDim];
  float gW2d[kHidden];
  float gW2f[3 * kHidden];
  float gB1[kBgHidden * kBgEncDim];
  float gB2[kBgHidden * kBgHidden];
  float gB3[3 * kBgHidden];
  float tile_h[kBwdWarps][32 * kTileStride];
  float tile_enc[kBwdWarps][kEncDim * kEncStride];
};

__global__ void __launch_bounds__(kBwdWarps * 32, 1)
render_nerf_bwd_kernel(const __grid_constant__ FieldMeta f, const FieldPtrs p, const FieldGrads g,
                       const __grid_constant__ MarchMeta m, const RayIO io) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& s = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int occ_words = (m.grid_res * m.grid_res * m.grid_res + 31) / 32;
  for (int i = threadIdx.x; i < kHidden * kEncDim; i += blockDim.x) {
    s.w1[i] = p.w1d[i];
    s.w1[kHidden * kEncDim + i] = p.w1f[i];
  }
  for (int i = threadIdx.x; i < kHidden; i += blockDim.x) s.w2d[i] = p.w2d[i];
  for (int i = threadIdx.x; i < 3 * kHidden; i += blockDim.x) s.w2f[i] = p.w2f[i];
  for (int i = threadIdx.x; i < kBgHidden * kBgEncDim; i += blockDim.x) s.b1[i] = p.bg_w1[i];
  for (int i = threadIdx.x; i < kBgHidden * kBgHidden; i += blockDim.x) s.b2[i] = p.bg_w2[i];
  for (int i = threadIdx.x; i < 3 * kBgHidden; i += blockDim.x) s.b3[i] = p.bg_w3[i];
  for (int i = threadIdx.x; i < occ_words; i += blockDim.x) s.occ[i] = io.occ_bits[i];
  for (int i = threadIdx.x; i < kCat * kEncDim; i += blockDim.x) s.gW1[i] = 0.f;
  for (int i = threadIdx.x; i < kHidden; i += blockDim.x) s.gW2d[i] = 0.f;
  for (int i = threadIdx.x; i < 3 * kHidden; i += blockDim.x) s.gW2f[i] = 0.f;
  for (int i = threadIdx.x; i < kBgHidden * kBgEncDim; i += blockDim.x) s.gB1[i] = 0.f;
  for (int i = threadIdx.x; i < kBgHidden * kBgHidden; i += blockDim.x) s.gB2[i] = 0.f;
  for (int i = threadIdx.x; i < 3 * kBgHidden; i += blockDim.x) s.gB3[i] = 0.f;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float* th = s.tile_h[warp];
  float* tenc = s.tile_enc[warp];
  const int n_items = (io.n_rays + 31) / 32;
  const float2* table = reinterpret_cast<const float2*>(p.table);
  const float2* bg_table = reinterpret_cast<const float2*>(p.bg_table);
  float2* g_table = reinterpret_cast<float2*>(g.table);
  float2* g_bg_table = reinterpret_cast<float2*>(g.bg_table);
  float thre = 0.f, eps_T = 0.f;
  if (m.prune) {
    thre = m.alpha_thre;
    if (io.occ_mean) thre = fminf(thre, __ldg(io.occ_mean));
    eps_T = m.early_stop_eps;
  }

  float acc[kCat];  // dW1[:, lane]
#pragma unroll
  for (int j = 0; j < kCat; ++j) acc[j] = 0.f;
  float g2d0 = 0.f, g2d1 = 0.f, g2f[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};

  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(io.work_counter, 1);
    item = __shfl_sync(kFullMask, item, 0);
    if (item >= n_items) break;

    // ---- phase 1: per-lane ray setup, output grads, background backward (thread per ray) ----
    const int my_ray = lane * n_items + item;
    const bool my_valid = my_ray < io.n_rays;
    float mo[3] = {0.f, 0.f, 0.f}, md[3] = {0.f, 0.f, 1.f}, mjit = 0.f;
    float gC[3] = {0.f, 0.f, 0.f}, gO = 0.f, gD = 0.f, Rtot = 0.f;
    {
      float dbg[3] = {0.f, 0.f, 0.f};
      BgActs a;
      float bgc[3] = {0.f, 0.f, 0.f};
      if (my_valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          mo[c] = __ldg(io.rays_o + 3 * my_ray + c);
          md[c] = __ldg(io.rays_d + 3 * my_ray + c);
          gC[c] = __ldg(io.g_comp_rgb + 3 * my_ray + c);
        }
        if (io.jitter) mjit = __ldg(io.jitter + my_ray);
        const float op = __ldg(io.opacity + my_ray);
        const float dp = __ldg(io.depth + my_ray);
        gO = io.g_opacity ? __ldg(io.g_opacity + my_ray) : 0.f;
        gD = io.g_depth ? __ldg(io.g_depth + my_ray) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float b = __ldg(io.comp_rgb_bg + 3 * my_ray + c);
          gO = fmaf(-gC[c], b, gO);  // comp_rgb = fg + bg (1 - op)
          Rtot = fmaf(gC[c], __ldg(io.comp_rgb_fg + 3 * my_ray + c), Rtot);
          dbg[c] = gC[c] * (1.f - op);
        }
        Rtot = fmaf(gO, op, Rtot);
        Rtot = fmaf(gD, dp, Rtot);
      }
      if (!io.bg_override) {  // random-colour augmentation detaches the environment map (color * 0 + rand)
        float dpre[3] = {0.f, 0.f, 0.f}, dh2[kBgHidden], dh1[kBgHidden], denc[kBgEncDim];
        if (my_valid) {
          bg_forward(f, bg_table, s.b1, s.b2, s.b3, md[0], md[1], md[2], a, bgc);
#pragma unroll
          for (int c = 0; c < 3; ++c) dpre[c] = dbg[c] * color_activation_grad(f.bg_color_act, a.pre[c]);
        } else {
#pragma unroll
          for (int i = 0; i < kBgEncDim; ++i) a.enc[i] = 0.f;
#pragma unroll
          for (int i = 0; i < kBgHidden; ++i) a.h1[i] = a.h2[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < kBgHidden; ++i) {
          float v = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) v = fmaf(s.b3[c * kBgHidden + i], dpre[c], v);
          dh2[i] = a.h2[i] > 0.f ? v : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kBgHidden; ++i) {
          float v = 0.f;
#pragma unroll
          for (int j = 0; j < kBgHidden; ++j) v = fmaf(s.b2[j * kBgHidden + i], dh2[j], v);
          dh1[i] = a.h1[i] > 0.f ? v : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kBgEncDim; ++i) {
          float v = 0.f;
#pragma unroll
          for (int j = 0; j < kBgHidden; ++j) v = fmaf(s.b1[j * kBgEncDim + i], dh1[j], v);
          denc[i] = v;
        }
        if (my_valid)
          grid_scatter<4>(g_bg_table, f.bg_grid, (md[0] + 1.f) * 0.5f, (md[1] + 1.f) * 0.5f, (md[2] + 1.f) * 0.5f,
                          denc);
        // weight grads: reduce over the 32 rays of the bundle, one shared-memory atomic per weight
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < kBgHidden; ++i) {
            const float v = warp_sum(a.h2[i] * dpre[c]);
            if (lane == 0) atomicAdd(&s.gB3[c * kBgHidden + i], v);
          }
#pragma unroll
        for (int j = 0; j < kBgHidden; ++j)
#pragma unroll
          for (int i = 0; i < kBgHidden; ++i) {
            const float v = warp_sum(a.h1[i] * dh2[j]);
            if (lane == 0) atomicAdd(&s.gB2[j * kBgHidden + i], v);
          }
#pragma unroll
        for (int j = 0; j < kBgHidden; ++j)
#pragma unroll
          for (int i = 0; i < kBgEncDim; ++i) {
            const float v = warp_sum(a.enc[i] * dh1[j]);
            if (lane == 0) atomicAdd(&s.gB1[j * kBgEncDim + i], v);
          }
      }
    }

    // ---- phase 2: warp-per-ray re-march with gradient propagation ----
    for (int r = 0; r < 32; ++r) {
      const int ray = r * n_items + item;
      if (ray >= io.n_rays) break;
      Marcher mc;
      mc.init(__shfl_sync(kFullMask, mo[0], r), __shfl_sync(kFullMask, mo[1], r), __shfl_sync(kFullMask, mo[2], r),
              __shfl_sync(kFullMask, md[0], r), __shfl_sync(kFullMask, md[1], r), __shfl_sync(kFullMask, md[2], r),
              __shfl_sync(kFullMask, mjit, r), m, f.radius);
      const float rgC0 = __shfl_sync(kFullMask, gC[0], r), rgC1 = __shfl_sync(kFullMask, gC[1], r),
                  rgC2 = __shfl_sync(kFullMask, gC[2], r);
      const float rgO = __shfl_sync(kFullMask, gO, r), rgD = __shfl_sync(kFullMask, gD, r);
      const float rR = __shfl_sync(kFullMask, Rtot, r);
      float S_all = 0.f, S_kept = 0.f, P = 0.f;
      while (true) {
        int my_k;
        const int filled = mc.next(s.occ, m.grid_res, f.radius, lane, &my_k);
        if (filled == 0) break;
        const bool valid = lane < filled;
        float sigma = 0.f, raw = 0.f, tm = 0.f, sd = 0.f, delta = 0.f;
        float x01 = 0.f, y01 = 0.f, z01 = 0.f;
        float enc[kEncDim];
        // A1: encode + density hidden layer (hidden activations parked in the warp tile)
        if (valid) {
          tm = mc.tmid(my_k);
          const float ts = fmaf((float)my_k, mc.step, mc.c0 - 0.5f * mc.step);
          delta = (ts + mc.step) - ts;
          const float px = fmaf(mc.dx, tm, mc.ox), py = fmaf(mc.dy, tm, mc.oy), pz = fmaf(mc.dz, tm, mc.oz);
          const float inv2r = 0.5f / f.radius;
          x01 = (px + f.radius) * inv2r;
          y01 = (py + f.radius) * inv2r;
          z01 = (pz + f.radius) * inv2r;
          grid_encode<kMaxLevels>(table, f.grid, x01, y01, z01, enc);
          float out = 0.f;
#pragma unroll 2
          for (int j4 = 0; j4 < kHidden; j4 += 4) {
            float h4[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float4* w = reinterpret_cast<const float4*>(s.w1 + (j4 + jj) * kEncDim);
              float h0 = 0.f, h1 = 0.f;
#pragma unroll
              for (int q = 0; q < kEncDim / 4; q += 2) {
                const float4 a = w[q], b = w[q + 1];
                h0 = fmaf(a.x, enc[4 * q + 0], h0);
                h0 = fmaf(a.y, enc[4 * q + 1], h0);
                h0 = fmaf(a.z, enc[4 * q + 2], h0);
                h0 = fmaf(a.w, enc[4 * q + 3], h0);
                h1 = fmaf(b.x, enc[4 * q + 4], h1);
                h1 = fmaf(b.y, enc[4 * q + 5], h1);
                h1 = fmaf(b.z, enc[4 * q + 6], h1);
                h1 = fmaf(b.w, enc[4 * q + 7], h1);
              }
              h4[jj] = fmaxf(h0 + h1, 0.f);
              out = fmaf(s.w2d[j4 + jj], h4[jj], out);
            }
            *reinterpret_cast<float4*>(th + lane * kTileStride + j4) = make_float4(h4[0], h4[1], h4[2], h4[3]);
          }
          raw = out + density_bias(f, px, py, pz);
          sigma = density_activation(f.density_act, raw);
          sd = sigma * delta;
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

        float draw = 0.f, dfeat[3] = {0.f, 0.f, 0.f};
        float gw_w = 0.f, g_w = 0.f;
        if (vis) {
          // feature hidden layer + colour
          float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 2
          for (int j4 = 0; j4 < kHidden; j4 += 4) {
            float h4[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float4* wv = reinterpret_cast<const float4*>(s.w1 + (kHidden + j4 + jj) * kEncDim);
              float h0 = 0.f, h1 = 0.f;
#pragma unroll
              for (int q = 0; q < kEncDim / 4; q += 2) {
                const float4 a = wv[q], b = wv[q + 1];
                h0 = fmaf(a.x, enc[4 * q + 0], h0);
                h0 = fmaf(a.y, enc[4 * q + 1], h0);
                h0 = fmaf(a.z, enc[4 * q + 2], h0);
                h0 = fmaf(a.w, enc[4 * q + 3], h0);
                h1 = fmaf(b.x, enc[4 * q + 4], h1);
                h1 = fmaf(b.y, enc[4 * q + 5], h1);
                h1 = fmaf(b.z, enc[4 * q + 6], h1);
                h1 = fmaf(b.w, enc[4 * q + 7], h1);
              }
              h4[jj] = fmaxf(h0 + h1, 0.f);
              o0 = fmaf(s.w2f[j4 + jj], h4[jj], o0);
              o1 = fmaf(s.w2f[kHidden + j4 + jj], h4[jj], o1);
              o2 = fmaf(s.w2f[2 * kHidden + j4 + jj], h4[jj], o2);
            }
            *reinterpret_cast<float4*>(th + lane * kTileStride + kHidden + j4) =
                make_float4(h4[0], h4[1], h4[2], h4[3]);
          }
          const float c0 = color_activation(f.color_act, o0), c1 = color_activation(f.color_act, o1),
                      c2 = color_activation(f.color_act, o2);
          g_w = fmaf(rgC0, c0, fmaf(rgC1, c1, fmaf(rgC2, c2, fmaf(rgD, tm, rgO))));
          gw_w = g_w * w;
          dfeat[0] = w * rgC0 * color_activation_grad(f.color_act, o0);
          dfeat[1] = w * rgC1 * color_activation_grad(f.color_act, o1);
          dfeat[2] = w * rgC2 * color_activation_grad(f.color_act, o2);
          // park enc transposed for the dW1 pass
#pragma unroll
          for (int i = 0; i < kEncDim; ++i) tenc[i * kEncStride + lane] = enc[i];
        }
        const float inclP = warp_scan_incl(gw_w, lane);
        if (vis) {
          const float dsigma = delta * (g_w * (T * e_sd) - (rR - (P + inclP)));
          draw = dsigma * density_activation_grad(f.density_act, raw);
        }
        P += __shfl_sync(kFullMask, inclP, 31);
        __syncwarp();

        if (vm) {
          // B1: second-layer weight grads, lane l owns hidden units {l, l+32} of each net
          for (uint32_t mm = vm; mm; mm &= mm - 1u) {
            const int sidx = __ffs(mm) - 1;
            const float dr = __shfl_sync(kFullMask, draw, sidx);
            const float d0 = __shfl_sync(kFullMask, dfeat[0], sidx), d1 = __shfl_sync(kFullMask, dfeat[1], sidx),
                        d2 = __shfl_sync(kFullMask, dfeat[2], sidx);
            const float* row = th + sidx * kTileStride;
            const float ha = row[lane], hb = row[lane + 32], hc = row[kHidden + lane], hd = row[kHidden + lane + 32];
            g2d0 = fmaf(ha, dr, g2d0);
            g2d1 = fmaf(hb, dr, g2d1);
            g2f[0][0] = fmaf(hc, d0, g2f[0][0]);
            g2f[0][1] = fmaf(hd, d0, g2f[0][1]);
            g2f[1][0] = fmaf(hc, d1, g2f[1][0]);
            g2f[1][1] = fmaf(hd, d1, g2f[1][1]);
            g2f[2][0] = fmaf(hc, d2, g2f[2][0]);
            g2f[2][1] = fmaf(hd, d2, g2f[2][1]);
          }
          __syncwarp();
          // A2: hidden grads (overwrite the tile row in place) and d_enc
          if (vis) {
            float denc[kEncDim];
#pragma unroll
            for (int i = 0; i < kEncDim; ++i) denc[i] = 0.f;
#pragma unroll 1
            for (int j4 = 0; j4 < kCat; j4 += 4) {
              float4 hv = *reinterpret_cast<float4*>(th + lane * kTileStride + j4);
              float hh[4] = {hv.x, hv.y, hv.z, hv.w};
              float dh[4];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = j4 + jj;
                float v;
                if (j4 < kHidden) {
                  v = s.w2d[j] * draw;
                } else {
                  const int jf = j - kHidden;
                  v = fmaf(s.w2f[jf], dfeat[0], fmaf(s.w2f[kHidden + jf], dfeat[1], s.w2f[2 * kHidden + jf] * dfeat[2]));
                }
                dh[jj] = hh[jj] > 0.f ? v : 0.f;
                const float4* wv = reinterpret_cast<const float4*>(s.w1 + j * kEncDim);
#pragma unroll
                for (int q = 0; q < kEncDim / 4; ++q) {
                  const float4 a = wv[q];
                  denc[4 * q + 0] = fmaf(a.x, dh[jj], denc[4 * q + 0]);
                  denc[4 * q + 1] = fmaf(a.y, dh[jj], denc[4 * q + 1]);
                  denc[4 * q + 2] = fmaf(a.z, dh[jj], denc[4 * q + 2]);
                  denc[4 * q + 3] = fmaf(a.w, dh[jj], denc[4 * q + 3]);
                }
              }
              *reinterpret_cast<float4*>(th + lane * kTileStride + j4) = make_float4(dh[0], dh[1], dh[2], dh[3]);
            }
            grid_scatter<kMaxLevels>(g_table, f.grid, x01, y01, z01, denc);
          }
          __syncwarp();
          // B2: first-layer weight grads, lane l owns input column l
          for (uint32_t mm = vm; mm; mm &= mm - 1u) {
            const int sidx = __ffs(mm) - 1;
            const float e = tenc[lane * kEncStride + sidx];
            const float4* row = reinterpret_cast<const float4*>(th + sidx * kTileStride);
#pragma unroll
            for (int q = 0; q < kCat / 4; ++q) {
              const float4 d = row[q];
              acc[4 * q + 0] = fmaf(d.x, e, acc[4 * q + 0]);
              acc[4 * q + 1] = fmaf(d.y, e, acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(d.z, e, acc[4 * q + 2]);
              acc[4 * q + 3] = fmaf(d.w, e, acc[4 * q + 3]);
            }
          }
          __syncwarp();
        }
        if (m.prune && expf(-S_all) < eps_T) break;
      }
    }
  }

  // ---- flush: registers -> CTA shared accumulators -> global ----
#pragma unroll
  for (int j = 0; j < kCat; ++j) atomicAdd(&s.gW1[j * kEncDim + lane], acc[j]);
  atomicAdd(&s.gW2d[lane], g2d0);
  atomicAdd(&s.gW2d[lane + 32], g2d1);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    atomicAdd(&s.gW2f[c * kHidden + lane], g2f[c][0]);
    atomicAdd(&s.gW2f[c * kHidden + lane + 32], g2f[c][1]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHidden * kEncDim; i += blockDim.x) {
    atomicAdd(g.w1d + i, s.gW1[i]);
    atomicAdd(g.w1f + i, s.gW1[kHidden * kEncDim + i]);
  }
  for (int i = threadIdx.x; i < kHidden; i += blockDim.x) atomicAdd(g.w2d + i, s.gW2d[i]);
  for (int i = threadIdx.x; i < 3 * kHidden; i += blockDim.x) atomicAdd(g.w2f + i, s.gW2f[i]);
  if (!io.bg_override) {
    for (int i = threadIdx.x; i < kBgHidden * kBgEncDim; i += blockDim.x) atomicAdd(g.bg_w1 + i, s.gB1[i]);
    for (int i = threadIdx.x; i < kBgHidden * kBgHidden; i += blockDim.x) atomicAdd(g.bg_w2 + i, s.gB2[i]);
    for (int i = threadIdx.x; i < 3 * kBgHidden; i += blockDim.x) atomicAdd(g.bg_w3 + i, s.gB3[i]);
  }
}

}  // namespace

int launch_render_bwd(const FieldMeta& f, const FieldPtrs& p, const FieldGrads& g, const MarchMeta& m,
                      const RayIO& io, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(render_nerf_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(BwdSmem));
    if (e != cudaSuccess) {
      sdb_set_error("render_nerf_bwd: smem attribute: %s", cudaGetErrorString(e));
      return SDB_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaMemsetAsync(io.work_counter, 0, sizeof(int), stream);
  const int n_items = (io.n_rays + 31) / 32;
  const int grid = min(kNumSMs, (n_items + kBwdWarps - 1) / kBwdWarps);
  render_nerf_bwd_kernel<<<grid, kBwdWarps * 32, sizeof(BwdSmem), stream>>>(f, p, g, m, io);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("render_nerf_bwd");
  return SDB_OK;
}
