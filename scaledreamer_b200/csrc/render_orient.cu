// Orientation loss of the fused NeRF renderer on the sample tape (sm_100a).
//
// Reference: threestudio/systems/scaledreamer.py:70-80
//     loss_orient = (weights.detach() * dot(normal, t_dirs).clamp_min(0)^2).sum() / (opacity > 0).sum()
// with the finite-difference normals of threestudio/models/geometry/implicit_volume.py:137-177
//     n = normalize(-(sigma(clamp(x + eps e_k, -r, r)) - sigma(x)) / eps),  k = x, y, z
// evaluated on the samples the renderer kept (nerf_volume_renderer.py:282-284, 375-386). The reference materialises
// per-sample `normal`, `weights`, `t_dirs` tensors; here nothing per-sample leaves the device:
//
//   render_orient_fwd_kernel   warp per ray over the ray's taped chunks: three offset densities per sample (hash-grid
//                              encode into the warp's tile + the density MLP as a 32-sample x 64-unit register tile),
//                              the normal, orient[ray] = sum_i w_i relu(n_i . d)^2, and the four partial derivatives
//                              d term / d raw density (centre, x, y, z offsets) into og [4][capacity].
//   render_orient_scale_kernel warp per ray: og *= g_orient[ray] (the upstream gradient of the per-ray sums).
//   render_orient_bwd_kernel   sample-parallel, 128-sample tiles, four passes (centre + three offsets): re-encode the
//                              point, density-MLP backward as register-tiled fp32 contractions (dH, dW2, dE = dH W1,
//                              dW1 += dH^T E) and the trilinear red.v2 scatter into the table gradient.
// The weights are detached in the reference, so no gradient reaches w / the compositing: only the density network and
// the hash table receive one.
#include "render_tape.cuh"

namespace {

constexpr int kOrWarps = 4;

struct OrientFwdSmem {
  float wpd[kWpSize];
  float w2d[kHidden];
  float et[kOrWarps][kEncDim * 32];
};

__global__ void __launch_bounds__(kOrWarps * 32, 4)
render_orient_fwd_kernel(const __grid_constant__ FieldMeta f, const FieldPtrs p, const float* __restrict__ rays_d,
                         const int n_rays, const RenderTape tape, float* __restrict__ orient, float* __restrict__ og) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  OrientFwdSmem& s = *reinterpret_cast<OrientFwdSmem*>(smem_raw);
  stage_w1_perm(s.wpd, p.w1d, threadIdx.x, blockDim.x);
  for (int i = threadIdx.x; i < kHidden; i += blockDim.x) s.w2d[i] = p.w2d[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int li = lane >> 3, lj = lane & 7;
  float* et = s.et[warp];
  const float2* table = reinterpret_cast<const float2*>(p.table);
  const float r = f.radius, inv2r = 0.5f / f.radius, eps = f.fd_eps, inv_eps = 1.f / f.fd_eps;
  const size_t cap = (size_t)tape.capacity;
  const int warps_total = gridDim.x * kOrWarps;
  float w2[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) w2[b] = s.w2d[hidden_of(lj, b)];

  for (int ray = blockIdx.x * kOrWarps + warp; ray < n_rays; ray += warps_total) {
    const int nch = __ldg(tape.ray_nchunks + ray);
    float acc_ray = 0.f;
    if (nch > 0) {
      const float dx = __ldg(rays_d + 3 * ray), dy = __ldg(rays_d + 3 * ray + 1), dz = __ldg(rays_d + 3 * ray + 2);
      const uint32_t* chunks = tape.ray_chunks + (size_t)ray * tape.max_chunks;
      for (int c0 = 0; c0 < nch; c0 += 32) {
        const uint32_t mine = (c0 + lane < nch) ? __ldg(chunks + c0 + lane) : 0u;
        const int nc = min(32, nch - c0);
        for (int c = 0; c < nc; ++c) {
          const uint32_t ch = __shfl_sync(kFullMask, mine, c);
          const int slot0 = (int)(ch >> 5), cnt = (int)(ch & 31u) + 1;
          const bool act = lane < cnt;
          const int slot = slot0 + lane;
          float px = 0.f, py = 0.f, pz = 0.f, raw0 = 0.f, w = 0.f;
          if (act) {
            px = fmaf(tape.pos[slot], 2.f * r, -r);
            py = fmaf(tape.pos[cap + slot], 2.f * r, -r);
            pz = fmaf(tape.pos[2 * cap + slot], 2.f * r, -r);
            raw0 = tape.sample[slot];
            w = tape.sample[4 * cap + slot];
          }
          float sig[3], dact[3];
#pragma unroll 1
          for (int k = 0; k < 3; ++k) {
            float qx = px, qy = py, qz = pz;
            if (k == 0) qx = fminf(fmaxf(px + eps, -r), r);
            if (k == 1) qy = fminf(fmaxf(py + eps, -r), r);
            if (k == 2) qz = fminf(fmaxf(pz + eps, -r), r);
            __syncwarp();  // the previous tile's readers are done
            encode_to_tile<32>(table, f.grid, (qx + r) * inv2r, (qy + r) * inv2r, (qz + r) * inv2r, act, et + lane);
            __syncwarp();
            float acc[8][8], part[8];
            hidden_tile(et, 32, s.wpd, li, lj, acc);
#pragma unroll
            for (int a = 0; a < 8; ++a) {
              float sum = 0.f;
#pragma unroll
              for (int b = 0; b < 8; ++b) sum = fmaf(w2[b], fmaxf(acc[a][b], 0.f), sum);
              part[a] = sum;
            }
            const float raw_k = reduce_scatter8(part, lj) + density_bias(f, qx, qy, qz);
            sig[k] = density_activation(f.density_act, raw_k);
            dact[k] = density_activation_grad(f.density_act, raw_k);
          }
          if (act) {
            const float sigma0 = density_activation(f.density_act, raw0);
            const float gx = -(sig[0] - sigma0) * inv_eps, gy = -(sig[1] - sigma0) * inv_eps,
                        gz = -(sig[2] - sigma0) * inv_eps;
            const float norm = sqrtf(gx * gx + gy * gy + gz * gz);
            const bool tiny = norm < 1e-12f;  // F.normalize: x / max(|x|, 1e-12)
            const float inv = 1.f / fmaxf(norm, 1e-12f);
            const float nx = gx * inv, ny = gy * inv, nz = gz * inv;
            const float cs = nx * dx + ny * dy + nz * dz;
            const float rc = fmaxf(cs, 0.f);
            acc_ray = fmaf(w * rc, rc, acc_ray);
            // d term / d g = 2 w relu(c) (d - c n) / |g|   (no projection while the norm sits under the clamp)
            const float coef = 2.f * w * rc * inv;
            const float pc = tiny ? 0.f : cs;
            const float tgx = coef * (dx - pc * nx), tgy = coef * (dy - pc * ny), tgz = coef * (dz - pc * nz);
            og[slot] = (tgx + tgy + tgz) * inv_eps * density_activation_grad(f.density_act, raw0);
            og[cap + slot] = -tgx * inv_eps * dact[0];
            og[2 * cap + slot] = -tgy * inv_eps * dact[1];
            og[3 * cap + slot] = -tgz * inv_eps * dact[2];
          }
        }
      }
    }
    acc_ray = warp_sum(acc_ray);
    if (lane == 0) orient[ray] = acc_ray;
  }
}

// og[k][slot] *= g_orient[ray of slot]
__global__ void __launch_bounds__(256)
render_orient_scale_kernel(const int n_rays, const RenderTape tape, float* __restrict__ og,
                           const float* __restrict__ g_orient) {
  const int lane = threadIdx.x & 31;
  const size_t cap = (size_t)tape.capacity;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ray < n_rays; ray += warps_total) {
    const int nch = __ldg(tape.ray_nchunks + ray);
    if (nch == 0) continue;
    const float g = __ldg(g_orient + ray);
    const uint32_t* chunks = tape.ray_chunks + (size_t)ray * tape.max_chunks;
    for (int c = 0; c < nch; ++c) {
      const uint32_t ch = __ldg(chunks + c);
      const int slot0 = (int)(ch >> 5), cnt = (int)(ch & 31u) + 1;
      if (lane < cnt) {
#pragma unroll
        for (int k = 0; k < 4; ++k) og[k * cap + slot0 + lane] *= g;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ density-net backward
constexpr int kObThreads = 128;
constexpr int kObTile = 128;
constexpr int kObEtStride = 36;
constexpr int kObDhStride = kObTile + 4;

struct OrientBwdSmem {
  float wp[kWpSize];                       // W1 (density) permuted for the hidden recompute
  float w1[kHidden * kEncDim];             // W1 row-major [hidden][enc] for dE = dH W1
  float w2d[kHidden];
  float et[4][kEncDim * kObEtStride];      // encodings of the current pass, feature-major per 32-sample sub-tile
  float dht[kHidden * kObDhStride];        // dH^T: [hidden][sample]
  float pos[3][kObTile];                   // x01 of the points of the current pass
  float dout[kObTile];                     // d raw of the current pass
  float g2d[kHidden];
};

__global__ void __launch_bounds__(kObThreads, 2)
render_orient_bwd_kernel(const __grid_constant__ FieldMeta f, const FieldPtrs p, const FieldGrads g, const RenderTape tape,
                         const float* __restrict__ og) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  OrientBwdSmem& s = *reinterpret_cast<OrientBwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int li = lane >> 3, lj = lane & 7;
  stage_w1_perm(s.wp, p.w1d, tid, kObThreads);
  for (int i = tid; i < kHidden * kEncDim; i += kObThreads) s.w1[i] = p.w1d[i];
  for (int i = tid; i < kHidden; i += kObThreads) s.w2d[i] = p.w2d[i], s.g2d[i] = 0.f;
  __syncthreads();

  const int n = min(__ldg(tape.counter), tape.capacity);
  const int n_tiles = (n + kObTile - 1) / kObTile;
  const size_t cap = (size_t)tape.capacity;
  const float2* table = reinterpret_cast<const float2*>(p.table);
  float2* g_table = reinterpret_cast<float2*>(g.table);
  const float r = f.radius, inv2r = 0.5f / f.radius, eps = f.fd_eps;

  float accw[4][4];  // dW1[hidden hg + 16 a][enc eg + 8 b]
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) accw[a][b] = 0.f;
  float g2d[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) g2d[b] = 0.f;
  const int hg = tid >> 3, eg = tid & 7;
  float w2[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) w2[b] = s.w2d[hidden_of(lj, b)];

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int base = tile * kObTile;
    const int slot = base + tid;
    const bool valid = slot < n;
    float cx01 = 0.f, cy01 = 0.f, cz01 = 0.f;
    if (valid) {
      cx01 = tape.pos[slot];
      cy01 = tape.pos[cap + slot];
      cz01 = tape.pos[2 * cap + slot];
    }
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      // ---- stage this pass: point, upstream gradient, encoding
      float x01 = cx01, y01 = cy01, z01 = cz01;
      if (k > 0) {
        float q = fmaf(k == 1 ? cx01 : (k == 2 ? cy01 : cz01), 2.f * r, -r);
        q = (fminf(fmaxf(q + eps, -r), r) + r) * inv2r;
        if (k == 1) x01 = q;
        if (k == 2) y01 = q;
        if (k == 3) z01 = q;
      }
      const float dr = valid ? __ldg(og + k * cap + slot) : 0.f;
      s.pos[0][tid] = x01;
      s.pos[1][tid] = y01;
      s.pos[2][tid] = z01;
      s.dout[tid] = dr;
      const bool any = __ballot_sync(kFullMask, dr != 0.f) != 0u;  // a warp whose 32 samples carry no gradient skips the gather
      encode_to_tile<kObEtStride>(table, f.grid, x01, y01, z01, valid && any, s.et[warp] + lane);
      __syncthreads();

      // (a) hidden recompute, (b) dH and dW2 partials
      float dE[8][4];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) dE[a][c] = 0.f;
      {
        float acc[8][8];
        hidden_tile(s.et[warp], kObEtStride, s.wp, li, lj, acc);
        const int s0 = warp * 32 + 8 * li;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const float d = s.dout[s0 + a];
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            const float h = fmaxf(acc[a][b], 0.f);
            g2d[b] = fmaf(h, d, g2d[b]);
            acc[a][b] = acc[a][b] > 0.f ? w2[b] * d : 0.f;
          }
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          float* row = s.dht + hidden_of(lj, b) * kObDhStride + s0;
          *reinterpret_cast<float4*>(row) = make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
          *reinterpret_cast<float4*>(row + 4) = make_float4(acc[4][b], acc[5][b], acc[6][b], acc[7][b]);
        }
      }
      __syncwarp();
      // (c) dE = dH W1 for this warp's samples
      {
        const float* dcol = s.dht + warp * 32 + 8 * li;
        const float* wrow = s.w1 + 4 * lj;
#pragma unroll 2
        for (int h = 0; h < kHidden; ++h) {
          const float4 d0 = *reinterpret_cast<const float4*>(dcol + h * kObDhStride);
          const float4 d1 = *reinterpret_cast<const float4*>(dcol + h * kObDhStride + 4);
          const float4 wv = *reinterpret_cast<const float4*>(wrow + h * kEncDim);
          const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            dE[a][0] = fmaf(d[a], wv.x, dE[a][0]);
            dE[a][1] = fmaf(d[a], wv.y, dE[a][1]);
            dE[a][2] = fmaf(d[a], wv.z, dE[a][2]);
            dE[a][3] = fmaf(d[a], wv.w, dE[a][3]);
          }
        }
      }
      __syncthreads();
      // (d) dW1 += dH^T E over the 128 samples of the tile
#pragma unroll 1
      for (int sub = 0; sub < 4; ++sub) {
#pragma unroll 1
        for (int s4 = 0; s4 < 8; ++s4) {
          float4 dv[4], ev[4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
            dv[a] = *reinterpret_cast<const float4*>(s.dht + (hg + 16 * a) * kObDhStride + sub * 32 + s4 * 4);
#pragma unroll
          for (int b = 0; b < 4; ++b)
            ev[b] = *reinterpret_cast<const float4*>(&s.et[sub][(eg + 8 * b) * kObEtStride + s4 * 4]);
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              float v = accw[a][b];
              v = fmaf(dv[a].x, ev[b].x, v);
              v = fmaf(dv[a].y, ev[b].y, v);
              v = fmaf(dv[a].z, ev[b].z, v);
              v = fmaf(dv[a].w, ev[b].w, v);
              accw[a][b] = v;
            }
        }
      }
      // ---- scatter: lane pairs, x-neighbour corners in one instruction (scatter_encoding_grads, render_tape.cuh)
      {
        const int sl0 = warp * 32 + 8 * li;
        scatter_encoding_grads(f.grid, g_table, dE, lj, &s.pos[0][sl0], &s.pos[1][sl0], &s.pos[2][sl0], n - (base + sl0));
      }
      __syncthreads();  // pos / dout / et / dht are rewritten by the next pass
    }
  }

#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) atomicAdd(g.w1d + (hg + 16 * a) * kEncDim + eg + 8 * b, accw[a][b]);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    float v = g2d[b];
    v += __shfl_xor_sync(kFullMask, v, 8);
    v += __shfl_xor_sync(kFullMask, v, 16);
    if (li == 0) atomicAdd(&s.g2d[hidden_of(lj, b)], v);
  }
  __syncthreads();
  for (int i = tid; i < kHidden; i += kObThreads) atomicAdd(g.w2d + i, s.g2d[i]);
}

}  // namespace

int launch_render_orient_fwd(const FieldMeta& f, const FieldPtrs& p, const float* rays_d, int n_rays,
                             const RenderTape& tape, float* orient, float* og, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(render_orient_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(OrientFwdSmem));
    if (e != cudaSuccess) {
      sdb_set_error("render_orient_fwd: smem attribute: %s", cudaGetErrorString(e));
      return SDB_ERR_CUDA;
    }
    attr_set = true;
  }
  const int grid = max(1, min(kNumSMs * 4, (n_rays + kOrWarps - 1) / kOrWarps));
  render_orient_fwd_kernel<<<grid, kOrWarps * 32, sizeof(OrientFwdSmem), stream>>>(f, p, rays_d, n_rays, tape, orient, og);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("render_orient_fwd");
  return SDB_OK;
}

int launch_render_orient_bwd(const FieldMeta& f, const FieldPtrs& p, const FieldGrads& g, int n_rays,
                             const RenderTape& tape, float* og, const float* g_orient, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(render_orient_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(OrientBwdSmem));
    if (e != cudaSuccess) {
      sdb_set_error("render_orient_bwd: smem attribute: %s", cudaGetErrorString(e));
      return SDB_ERR_CUDA;
    }
    attr_set = true;
  }
  const int grid_s = max(1, min(kNumSMs * 8, (n_rays + 7) / 8));
  render_orient_scale_kernel<<<grid_s, 256, 0, stream>>>(n_rays, tape, og, g_orient);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("render_orient_scale");
  const int max_tiles = (tape.capacity + kObTile - 1) / kObTile;
  const int grid_b = max(1, min(kNumSMs * 2, max_tiles));
  render_orient_bwd_kernel<<<grid_b, kObThreads, sizeof(OrientBwdSmem), stream>>>(f, p, g, tape, og);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("render_orient_bwd");
  return SDB_OK;
}
