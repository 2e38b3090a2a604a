// Device-side building blocks of the iNGP field: hash-grid encode / scatter, tiny MLPs,
// density bias + activation, background environment map.
//
// Follows (restates, does not copy) the maths of the reference call sites:
//   threestudio/models/geometry/implicit_volume.py:80-107,109-196  (density bias, activation, FD normals)
//   threestudio/models/networks.py:55-64,214-251                   (tcnn HashGrid wrapper, bias-free ReLU MLP)
//   threestudio/models/background/neural_environment_map_background.py:46-67
// Hash-grid semantics are those of tiny-cuda-nn's GridEncoding (un-vendored dependency).
#pragma once
#include "render_types.cuh"

__device__ __forceinline__ uint32_t grid_index(uint32_t hashed, uint32_t res, uint32_t size, uint32_t cx,
                                               uint32_t cy, uint32_t cz) {
  uint32_t idx;
  if (hashed) {
    idx = (cx ^ (cy * 2654435761u) ^ (cz * 805459861u)) & (size - 1u);  // hashed levels are 2^k sized
  } else {
    idx = cx + cy * res + cz * res * res;
    if (idx >= size) idx %= size;
  }
  return idx;
}

// Encodes one point x in [0,1]^3 through L levels (2 features each) into enc[2L].
template <int L>
__device__ __forceinline__ void grid_encode(const float2* __restrict__ table, const GridMeta& gm, float x,
                                            float y, float z, float* enc) {
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const float s = gm.scale[l];
    const uint32_t res = gm.res[l], size = gm.size[l], hashed = gm.hashed[l];
    const float2* tl = table + gm.offset[l];
    const float px = fmaf(x, s, 0.5f), py = fmaf(y, s, 0.5f), pz = fmaf(z, s, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const uint32_t ix = (uint32_t)(int)fx, iy = (uint32_t)(int)fy, iz = (uint32_t)(int)fz;
    const float wx = px - fx, wy = py - fy, wz = pz - fz;
    float2 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t idx = grid_index(hashed, res, size, ix + (c & 1), iy + ((c >> 1) & 1), iz + ((c >> 2) & 1));
      v[c] = __ldg(tl + idx);
    }
    float ax = 0.f, ay = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float w = ((c & 1) ? wx : 1.f - wx) * (((c >> 1) & 1) ? wy : 1.f - wy) * (((c >> 2) & 1) ? wz : 1.f - wz);
      ax = fmaf(w, v[c].x, ax);
      ay = fmaf(w, v[c].y, ay);
    }
    enc[2 * l] = ax;
    enc[2 * l + 1] = ay;
  }
}

// Scatter-adds d_enc[2L] into the table gradient with the same trilinear weights.
template <int L>
__device__ __forceinline__ void grid_scatter(float2* __restrict__ gtable, const GridMeta& gm, float x, float y,
                                             float z, const float* d_enc) {
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const float gx = d_enc[2 * l], gy = d_enc[2 * l + 1];
    if (gx == 0.f && gy == 0.f) continue;
    const float s = gm.scale[l];
    const uint32_t res = gm.res[l], size = gm.size[l], hashed = gm.hashed[l];
    float2* tl = gtable + gm.offset[l];
    const float px = fmaf(x, s, 0.5f), py = fmaf(y, s, 0.5f), pz = fmaf(z, s, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const uint32_t ix = (uint32_t)(int)fx, iy = (uint32_t)(int)fy, iz = (uint32_t)(int)fz;
    const float wx = px - fx, wy = py - fy, wz = pz - fz;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t idx = grid_index(hashed, res, size, ix + (c & 1), iy + ((c >> 1) & 1), iz + ((c >> 2) & 1));
      const float w = ((c & 1) ? wx : 1.f - wx) * (((c >> 1) & 1) ? wy : 1.f - wy) * (((c >> 2) & 1) ? wz : 1.f - wz);
      atomicAdd(tl + idx, make_float2(w * gx, w * gy));  // red.global.add.v2.f32 (sm_90+)
    }
  }
}

// hidden = relu(W1 enc) ; out = W2 hidden, one output. W1 rows are read as broadcast float4s.
__device__ __forceinline__ float mlp32_64_1(const float* __restrict__ sW1, const float* __restrict__ sW2,
                                            const float* enc) {
  float out = 0.f;
#pragma unroll 4
  for (int j = 0; j < kHidden; ++j) {
    const float4* w = reinterpret_cast<const float4*>(sW1 + j * kEncDim);
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
    const float h = fmaxf(h0 + h1, 0.f);
    out = fmaf(sW2[j], h, out);
  }
  return out;
}

__device__ __forceinline__ void mlp32_64_3(const float* __restrict__ sW1, const float* __restrict__ sW2,
                                           const float* enc, float* out3) {
  float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 4
  for (int j = 0; j < kHidden; ++j) {
    const float4* w = reinterpret_cast<const float4*>(sW1 + j * kEncDim);
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
    const float h = fmaxf(h0 + h1, 0.f);
    o0 = fmaf(sW2[j], h, o0);
    o1 = fmaf(sW2[kHidden + j], h, o1);
    o2 = fmaf(sW2[2 * kHidden + j], h, o2);
  }
  out3[0] = o0;
  out3[1] = o1;
  out3[2] = o2;
}

__device__ __forceinline__ float density_bias(const FieldMeta& f, float x, float y, float z) {
  if (f.bias_type == 1) return f.blob_scale * (1.f - sqrtf(x * x + y * y + z * z) / f.blob_std);
  if (f.bias_type == 2) return f.blob_scale * expf(-0.5f * (x * x + y * y + z * z) / (f.blob_std * f.blob_std));
  return f.bias_const;
}

__device__ __forceinline__ float density_activation(int act, float raw) {
  if (act == 0) return raw > 20.f ? raw : log1pf(expf(raw));  // F.softplus(beta=1, threshold=20)
  return expf(raw);                                            // exp / trunc_exp forward
}

// d activation / d raw
__device__ __forceinline__ float density_activation_grad(int act, float raw) {
  if (act == 0) return raw > 20.f ? 1.f : 1.f / (1.f + expf(-raw));
  if (act == 2) return expf(fminf(raw, 15.f));  // trunc_exp backward clamps at 15
  return expf(raw);
}

__device__ __forceinline__ float color_activation(int act, float x) {
  const float s = 1.f / (1.f + expf(-x));
  return act == 1 ? s * 1.002f - 0.001f : s;
}
// derivative expressed through the plain sigmoid value s
__device__ __forceinline__ float color_activation_grad(int act, float x) {
  const float s = 1.f / (1.f + expf(-x));
  return (act == 1 ? 1.002f : 1.f) * s * (1.f - s);
}

// Density at world-space point p (|p_i| <= radius). Returns sigma, writes raw pre-activation.
__device__ __forceinline__ float field_density(const FieldMeta& f, const float2* __restrict__ table,
                                               const float* sW1d, const float* sW2d, float px, float py, float pz,
                                               float* enc, float* raw_out) {
  const float inv2r = 0.5f / f.radius;
  grid_encode<kMaxLevels>(table, f.grid, (px + f.radius) * inv2r, (py + f.radius) * inv2r, (pz + f.radius) * inv2r,
                          enc);
  const float raw = mlp32_64_1(sW1d, sW2d, enc) + density_bias(f, px, py, pz);
  *raw_out = raw;
  return density_activation(f.density_act, raw);
}

// Finite-difference normal: n = normalize(-(sigma(x + eps e_k) - sigma(x)) / eps), offsets clamped to the box.
__device__ __forceinline__ void field_fd_normal(const FieldMeta& f, const float2* __restrict__ table,
                                                const float* sW1d, const float* sW2d, float px, float py, float pz,
                                                float sigma, float* n3) {
  float enc[kEncDim], raw;
  const float r = f.radius, e = f.fd_eps;
  const float sx = field_density(f, table, sW1d, sW2d, fminf(fmaxf(px + e, -r), r), py, pz, enc, &raw);
  const float sy = field_density(f, table, sW1d, sW2d, px, fminf(fmaxf(py + e, -r), r), pz, enc, &raw);
  const float sz = field_density(f, table, sW1d, sW2d, px, py, fminf(fmaxf(pz + e, -r), r), enc, &raw);
  const float nx = -(sx - sigma) / e, ny = -(sy - sigma) / e, nz = -(sz - sigma) / e;
  const float inv = 1.f / fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-12f);
  n3[0] = nx * inv;
  n3[1] = ny * inv;
  n3[2] = nz * inv;
}

// Background colour for direction d (thread-per-ray). Keeps the activations needed by the backward.
struct BgActs {
  float enc[kBgEncDim];
  float h1[kBgHidden];
  float h2[kBgHidden];
  float pre[3];
};

__device__ __forceinline__ void bg_forward(const FieldMeta& f, const float2* __restrict__ bg_table,
                                           const float* sB1, const float* sB2, const float* sB3, float dx, float dy,
                                           float dz, BgActs& a, float* rgb) {
  grid_encode<4>(bg_table, f.bg_grid, (dx + 1.f) * 0.5f, (dy + 1.f) * 0.5f, (dz + 1.f) * 0.5f, a.enc);
#pragma unroll
  for (int j = 0; j < kBgHidden; ++j) {
    float h = 0.f;
#pragma unroll
    for (int i = 0; i < kBgEncDim; ++i) h = fmaf(sB1[j * kBgEncDim + i], a.enc[i], h);
    a.h1[j] = fmaxf(h, 0.f);
  }
#pragma unroll
  for (int j = 0; j < kBgHidden; ++j) {
    float h = 0.f;
#pragma unroll
    for (int i = 0; i < kBgHidden; ++i) h = fmaf(sB2[j * kBgHidden + i], a.h1[i], h);
    a.h2[j] = fmaxf(h, 0.f);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float o = 0.f;
#pragma unroll
    for (int i = 0; i < kBgHidden; ++i) o = fmaf(sB3[c * kBgHidden + i], a.h2[i], o);
    a.pre[c] = o;
    rgb[c] = color_activation(f.bg_color_act, o);
  }
}

// Ray / axis-aligned-box slab test against [-r, r]^3. Returns false on a miss.
__device__ __forceinline__ bool ray_box(float ox, float oy, float oz, float dx, float dy, float dz, float r,
                                        float near_plane, float far_plane, float* t0, float* t1) {
  const float ix = 1.f / dx, iy = 1.f / dy, iz = 1.f / dz;
  float a = (-r - ox) * ix, b = (r - ox) * ix;
  float tmin = fminf(a, b), tmax = fmaxf(a, b);
  a = (-r - oy) * iy;
  b = (r - oy) * iy;
  tmin = fmaxf(tmin, fminf(a, b));
  tmax = fminf(tmax, fmaxf(a, b));
  a = (-r - oz) * iz;
  b = (r - oz) * iz;
  tmin = fmaxf(tmin, fminf(a, b));
  tmax = fminf(tmax, fmaxf(a, b));
  *t0 = fmaxf(tmin, near_plane);
  *t1 = fminf(tmax, far_plane);
  return *t1 > *t0;
}

// Occupancy bit of the cell containing p.
__device__ __forceinline__ bool occ_lookup(const uint32_t* sOcc, int res, float r, float px, float py, float pz) {
  const float k = (float)res * 0.5f / r;
  int cx = (int)floorf((px + r) * k), cy = (int)floorf((py + r) * k), cz = (int)floorf((pz + r) * k);
  cx = min(max(cx, 0), res - 1);
  cy = min(max(cy, 0), res - 1);
  cz = min(max(cz, 0), res - 1);
  const int bit = (cx * res + cy) * res + cz;
  return (sOcc[bit >> 5] >> (bit & 31)) & 1u;
}

// Candidate collector: walks the fixed-step lattice of one ray and hands out, 32 at a time, the
// lattice indices whose midpoint lies inside the box and in an occupied cell.
struct Marcher {
  int k_next, k_end;
  uint32_t pending;
  int base;
  float c0, step, t0, t1;
  float ox, oy, oz, dx, dy, dz;

  __device__ __forceinline__ float tmid(int k) const { return fmaf((float)k, step, c0); }

  __device__ __forceinline__ void init(float ox_, float oy_, float oz_, float dx_, float dy_, float dz_, float jit,
                                       const MarchMeta& m, float radius) {
    ox = ox_; oy = oy_; oz = oz_; dx = dx_; dy = dy_; dz = dz_;
    step = m.step;
    pending = 0u;
    base = 0;
    const bool hit = ray_box(ox, oy, oz, dx, dy, dz, radius, m.near_plane, m.far_plane, &t0, &t1);
    const float near_j = fmaf(jit, step, m.near_plane);
    c0 = fmaf(0.5f, step, near_j);
    if (!hit) {
      k_next = 0;
      k_end = 0;
    } else {
      k_next = max(0, (int)floorf((t0 - c0) / step) - 1);
      k_end = max(0, (int)ceilf((t1 - c0) / step) + 1);
    }
  }

  // Returns the number of candidates handed out (<=32); lane i < n receives its lattice index in *my_k.
  __device__ __forceinline__ int next(const uint32_t* sOcc, int grid_res, float radius, int lane, int* my_k) {
    int filled = 0;
    *my_k = -1;
    while (filled < 32) {
      if (pending == 0u) {
        if (k_next >= k_end) break;
        base = k_next;
        k_next += 32;
        const int k = base + lane;
        bool occ = false;
        if (k < k_end) {
          const float tm = tmid(k);
          if (tm >= t0 && tm < t1)
            occ = occ_lookup(sOcc, grid_res, radius, fmaf(dx, tm, ox), fmaf(dy, tm, oy), fmaf(dz, tm, oz));
        }
        pending = __ballot_sync(kFullMask, occ);
        continue;
      }
      const int n = __popc(pending);
      const int take = min(n, 32 - filled);
      const int idx = lane - filled;
      if (idx >= 0 && idx < take) *my_k = base + (int)__fns(pending, 0, idx + 1);
      if (take == n) {
        pending = 0u;
      } else {
        const uint32_t bit = __fns(pending, 0, take + 1);
        pending &= ~((1u << bit) - 1u);
      }
      filled += take;
    }
    return filled;
  }
};
