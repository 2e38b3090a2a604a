// Sample tape + warp-cooperative tiny-MLP tiles shared by the v2 render kernels (render_fwd2.cu, render_bwd2.cu).
//
// v1 (render_fwd.cu / render_bwd.cu) evaluates the 32-64-{1,3} MLPs thread-per-sample with the weights broadcast
// from shared memory: one LDS.128 per four FFMA and a fully unrolled 16-level encode -- ncu showed it stalled on
// instruction fetch and the LSU at 12-15 % occupancy. v2 keeps the warp-per-ray march but
//   * evaluates the hidden layers as 32-sample x 64-unit register tiles (lane (i,j) = 8 samples x 8 units, operands
//     read as LDS.128 with 16 FFMA per load),
//   * records every kept sample on a TAPE (encoding, position, pre-activations, weight, transmittance) so the
//     backward never re-marches or re-gathers: a light per-ray pass turns the image gradient into per-sample
//     d(raw density) / d(features), and a sample-parallel pass does the MLP backward + hash-grid scatter.
// Reference maths: see field.cuh / render_fwd.cu headers (nerf_volume_renderer.py:118-428, networks.py:214-251).
#pragma once
#include "field.cuh"

// Kept samples of one forward launch. All buffers are caller-owned device memory.
struct RenderTape {
  int capacity;          // sample slots (multiple of 128)
  int max_chunks;        // stride of ray_chunks (>= ceil(max candidates per ray / 32) + 1)
  int* counter;          // [2]: kept samples written, overflow flag
  float* enc;            // [capacity/32][32 (feature k)][32 (slot)]  tile-transposed encodings
  float* pos;            // [3][capacity]  x01, y01, z01
  float* sample;         // [8][capacity]  raw, o0, o1, o2, w, T*exp(-sigma*delta), t_mid, delta
                         //                (the backward overwrites rows 0..3 with d raw, d o0..2)
  uint32_t* ray_chunks;  // [n_rays][max_chunks]  (slot0 << 5) | (count - 1), front to back
  int* ray_nchunks;      // [n_rays]
};

static constexpr int kWpSize = kEncDim * kHidden;  // 2048 floats per net

// Hidden unit owned by lane column j, register b.
__device__ __forceinline__ int hidden_of(int j, int b) { return j + 8 * b; }

// Stages W1 [64][32] (nn.Linear layout) as Wp[k][half][j][c] = W1[hidden_of(j, 4*half + c)][k] so that lane column j
// reads its eight units of input feature k as two conflict-free LDS.128.
__device__ __forceinline__ void stage_w1_perm(float* __restrict__ Wp, const float* __restrict__ W1, int tid, int nthr) {
  for (int idx = tid; idx < kWpSize; idx += nthr) {
    const int k = idx >> 6, rem = idx & 63, half = rem >> 5, j = (rem & 31) >> 2, c = rem & 3;
    Wp[idx] = W1[hidden_of(j, 4 * half + c) * kEncDim + k];
  }
}

// acc[a][b] = sum_k ET[k][8i + a] * W1[hidden_of(j, b)][k]   (ET: [32][es] floats, feature-major)
__device__ __forceinline__ void hidden_tile(const float* __restrict__ ET, int es, const float* __restrict__ Wp, int i,
                                            int j, float (&acc)[8][8]) {
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
#pragma unroll 1
  for (int k = 0; k < kEncDim; ++k) {
    const float4 e0 = *reinterpret_cast<const float4*>(ET + k * es + 8 * i);
    const float4 e1 = *reinterpret_cast<const float4*>(ET + k * es + 8 * i + 4);
    const float4 w0 = *reinterpret_cast<const float4*>(Wp + (k * 16 + j) * 4);
    const float4 w1 = *reinterpret_cast<const float4*>(Wp + (k * 16 + 8 + j) * 4);
    const float e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(e[a], w[b], acc[a][b]);
  }
}

// v[a] holds this lane's partial sum for sample 8i + a; returns the total of sample 8i + j (= this lane's own
// sample) over the eight lanes of the i-group: a reduce-scatter in 7 shuffles.
__device__ __forceinline__ float reduce_scatter8(const float (&v)[8], int j) {
  float u[4], t[2];
  const bool h4 = (j & 4) != 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float send = h4 ? v[q] : v[q + 4];
    const float keep = h4 ? v[q + 4] : v[q];
    u[q] = keep + __shfl_xor_sync(kFullMask, send, 4);
  }
  const bool h2 = (j & 2) != 0;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float send = h2 ? u[q] : u[q + 2];
    const float keep = h2 ? u[q + 2] : u[q];
    t[q] = keep + __shfl_xor_sync(kFullMask, send, 2);
  }
  const bool h1 = (j & 1) != 0;
  const float send = h1 ? t[0] : t[1];
  const float keep = h1 ? t[1] : t[0];
  return keep + __shfl_xor_sync(kFullMask, send, 1);
}

// One level of the hash grid for point (x,y,z) in [0,1]^3: cell corner + trilinear fractions.
struct LevelCell {
  uint32_t ix, iy, iz;
  float wx, wy, wz;
};
__device__ __forceinline__ LevelCell level_cell(float s, float x, float y, float z) {
  LevelCell c;
  const float px = fmaf(x, s, 0.5f), py = fmaf(y, s, 0.5f), pz = fmaf(z, s, 0.5f);
  const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
  c.ix = (uint32_t)(int)fx;
  c.iy = (uint32_t)(int)fy;
  c.iz = (uint32_t)(int)fz;
  c.wx = px - fx;
  c.wy = py - fy;
  c.wz = pz - fz;
  return c;
}
__device__ __forceinline__ float corner_weight(const LevelCell& c, int k) {
  return ((k & 1) ? c.wx : 1.f - c.wx) * (((k >> 1) & 1) ? c.wy : 1.f - c.wy) * (((k >> 2) & 1) ? c.wz : 1.f - c.wz);
}

// Hash-table gradient scatter for the field backward kernels. Lane (li, lj) of a warp holds dE[a][c] = d loss / d encoding
// feature 4 lj + c (levels 2 lj, 2 lj + 1) of its eight samples a; px/py/pz point at those samples' positions (shared
// memory), the first n_valid of them exist. Lane PAIRS (lj, lj ^ 1) walk levels 4 (lj >> 1) .. + 3 together: one lane
// sends the four corners with x = cx, the other those with x = cx + 1 (the partner's half of dE arrives by shuffle).
// The two entries are neighbours in the table (dense levels: idx, idx + 1; hashed levels: x enters the hash
// un-multiplied), and two lanes of ONE red.v2 on one 32-byte sector cost one sector operation: 327 G lane-ops/s
// against 193 G/s for lanes on their own sectors (tools/red_probe.cu mode 3, profiles/r2t_red_probe.log).
// Consecutive samples that stay in one cell (coarse levels, samples a step apart on one ray) are summed per corner in
// registers and sent once when the cell changes. Must be called by all 32 lanes (it shuffles). level_mask: bit l = send
// level l.
__device__ __forceinline__ void scatter_encoding_grads(const GridMeta& gm, float2* __restrict__ g_table,
                                                       const float (&dE)[8][4], int lj, const float* px, const float* py,
                                                       const float* pz, int n_valid, uint32_t level_mask = 0xFFFFu) {
  const int xp = lj & 1;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int lvl = 4 * (lj >> 1) + q;
    const bool mine = (q >> 1) == xp;  // this lane computed dE for level lvl (else the partner did)
    float gxs[8], gys[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const float ox = __shfl_xor_sync(kFullMask, dE[a][2 * (q & 1)], 1);
      const float oy = __shfl_xor_sync(kFullMask, dE[a][2 * (q & 1) + 1], 1);
      gxs[a] = mine ? dE[a][2 * (q & 1)] : ox;
      gys[a] = mine ? dE[a][2 * (q & 1) + 1] : oy;
    }
    if (!((level_mask >> lvl) & 1u)) continue;
    const float sc = gm.scale[lvl];
    const uint32_t res = gm.res[lvl], size = gm.size[lvl], hashed = gm.hashed[lvl];
    float2* tl = g_table + gm.offset[lvl];
    uint32_t cx = 0u, cy = 0u, cz = 0u;
    float ax[4], ay[4];
    bool open = false;
    auto flush = [&]() {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t idx = grid_index(hashed, res, size, cx + xp, cy + (j & 1), cz + (j >> 1));
        atomicAdd(tl + idx, make_float2(ax[j], ay[j]));  // red.global.add.v2.f32
      }
    };
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      if (a >= n_valid || (gxs[a] == 0.f && gys[a] == 0.f)) continue;
      const LevelCell c = level_cell(sc, px[a], py[a], pz[a]);
      if (open && (c.ix != cx || c.iy != cy || c.iz != cz)) {
        flush();
        open = false;
      }
      if (!open) {
        cx = c.ix, cy = c.iy, cz = c.iz;
#pragma unroll
        for (int j = 0; j < 4; ++j) ax[j] = ay[j] = 0.f;
        open = true;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w = corner_weight(c, xp | (j << 1));
        ax[j] = fmaf(w, gxs[a], ax[j]);
        ay[j] = fmaf(w, gys[a], ay[j]);
      }
    }
    if (open) flush();
  }
}

// Encodes this lane's point through all levels straight into column `lane` of the warp's tile.
// `kStride` = floats per feature row of the tile (32 for the forward's per-warp tile).
template <int kStride = 32>
__device__ __forceinline__ void encode_to_tile(const float2* __restrict__ table, const GridMeta& gm, float x, float y,
                                               float z, bool valid, float* __restrict__ et_lane) {
#pragma unroll 2
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
    et_lane[(2 * l) * kStride] = ax;
    et_lane[(2 * l + 1) * kStride] = ay;
  }
}

int launch_render_fwd2(const FieldMeta& f, const FieldPtrs& p, const MarchMeta& m, const RayIO& io,
                       const RenderTape* tape, cudaStream_t stream);
int launch_render_bwd2(const FieldMeta& f, const FieldPtrs& p, const FieldGrads& g, const MarchMeta& m, const RayIO& io,
                       const RenderTape& tape, cudaStream_t stream);
// Field backward with the MLP contractions on the tensor cores (render_bwd_tc.cu); same contract as the fp32 kernel.
int launch_render_field_bwd_tc(const FieldMeta& f, const FieldPtrs& p, const FieldGrads& g, const RenderTape& tape,
                               int scatter_on, cudaStream_t stream);
// Orientation term on the taped samples (render_orient.cu): orient[ray] = sum_i w_i relu(n_i . d)^2 with finite-difference
// normals; og [4][capacity] receives d term / d raw density at the sample and its three offset points.
int launch_render_orient_fwd(const FieldMeta& f, const FieldPtrs& p, const float* rays_d, int n_rays,
                             const RenderTape& tape, float* orient, float* og, cudaStream_t stream);
int launch_render_orient_bwd(const FieldMeta& f, const FieldPtrs& p, const FieldGrads& g, int n_rays,
                             const RenderTape& tape, float* og, const float* g_orient, cudaStream_t stream);
