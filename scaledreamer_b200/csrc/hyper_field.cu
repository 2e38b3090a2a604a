// Prompt-conditioned hash-grid field (amortized generators): out = relu(enc(x) W1[b]) W2[b] with the per-prompt weight
// matrices produced by a hypernetwork, for up to two heads sharing one encoding:
//   head a: 32 -> 64 -> 1   (signed distance of Hyper-iNGP; reference hyper_iNGP.py:261-349 `hypernet_forward` = torch.bmm)
//   head b: 32 -> 64 -> 3   (features of Hyper-iNGP; colour of the multi-prompt environment map,
//                            multiprompt_neural_environment_hashgrid_map_background.py:83-101)
// Points are given already contracted to [0,1]^3, laid out [B, N, 3]; 128-point tiles never straddle prompts.
// Forward: rolled 16-level encode into a per-warp tile + register-tiled hidden layers (render_tape.cuh), optional
// tape of the encodings. Backward: the tape-based field backward of render_bwd2.cu generalised to per-prompt
// weights: hidden recompute, dE = dH W1^T, dW1 += E^T dH, dW2 += H^T d out, trilinear scatter into the table gradient.
#include <cstdlib>

#include "field_bwd_tc.cuh"

namespace {

constexpr int kHfThreads = 128;
constexpr int kHfTile = 128;

struct HfArgs {
  const float* table;
  const float* pts;   // [B, N, 3] in [0,1]
  int B, N, n_pad;    // n_pad = ceil(N / 128) * 128: tape slots per prompt
  const float* w1a;   // [B, 32, 64] or null
  const float* w2a;   // [B, 64, 1]
  const float* w1b;   // [B, 32, 64] or null
  const float* w2b;   // [B, 64, 3]
  float* out_a;       // [B, N]
  float* out_b;       // [B, N, 3]
  float* tape;        // [B * n_pad / 32][32][32] or null
  // backward
  const float* d_a;   // [B, N] or null
  const float* d_b;   // [B, N, 3] or null
  float* g_table;
  float* g_w1a;
  float* g_w2a;
  float* g_w1b;
  float* g_w2b;
};

// W1 [32 in][64 out] (enc @ W1 layout) -> Wp[k][half][j][c] (see stage_w1_perm)
__device__ __forceinline__ void stage_w1_inout(float* __restrict__ Wp, const float* __restrict__ W1, int tid, int nthr) {
  for (int idx = tid; idx < kWpSize; idx += nthr) {
    const int k = idx >> 6, rem = idx & 63, half = rem >> 5, j = (rem & 31) >> 2, c = rem & 3;
    Wp[idx] = W1[k * kHidden + hidden_of(j, 4 * half + c)];
  }
}

struct HfFwdSmem {
  float wpa[kWpSize];
  float wpb[kWpSize];
  float w2a[kHidden];
  float w2b[3 * kHidden];  // [c][h]
  float et[4][kEncDim * 32];
};

__device__ __forceinline__ void encode_tile(const float2* __restrict__ table, const GridMeta& gm, float x, float y,
                                            float z, bool valid, float* __restrict__ et_lane, int stride) {
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
    et_lane[(2 * l) * stride] = ax;
    et_lane[(2 * l + 1) * stride] = ay;
  }
}

__global__ void __launch_bounds__(kHfThreads, 4)
hyper_field_fwd_kernel(const __grid_constant__ GridMeta gm, const HfArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HfFwdSmem& s = *reinterpret_cast<HfFwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int li = lane >> 3, lj = lane & 7;
  float* et = s.et[warp];
  const float2* table = reinterpret_cast<const float2*>(a.table);
  const int tiles_per_b = a.n_pad / kHfTile;
  const int total = a.B * tiles_per_b;
  int cur_b = -1;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int b = tile / tiles_per_b, t = tile - b * tiles_per_b;
    if (b != cur_b) {
      __syncthreads();
      if (a.w1a) {
        stage_w1_inout(s.wpa, a.w1a + (size_t)b * kWpSize, tid, kHfThreads);
        for (int i = tid; i < kHidden; i += kHfThreads) s.w2a[i] = a.w2a[(size_t)b * kHidden + i];
      }
      if (a.w1b) {
        stage_w1_inout(s.wpb, a.w1b + (size_t)b * kWpSize, tid, kHfThreads);
        for (int i = tid; i < 3 * kHidden; i += kHfThreads)
          s.w2b[(i % 3) * kHidden + i / 3] = a.w2b[(size_t)b * 3 * kHidden + i];
      }
      __syncthreads();
      cur_b = b;
    }
    const int i = t * kHfTile + warp * 32 + lane;
    const bool valid = i < a.N;
    float x = 0.f, y = 0.f, z = 0.f;
    if (valid) {
      const float* p = a.pts + ((size_t)b * a.N + i) * 3;
      x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    }
    __syncwarp();
    encode_tile(table, gm, x, y, z, valid, et + lane, 32);
    __syncwarp();
    float acc[8][8], part[8];
    if (a.w1a) {
      hidden_tile(et, 32, s.wpa, li, lj, acc);
      float w2[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) w2[q] = s.w2a[hidden_of(lj, q)];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) sum = fmaf(w2[q], fmaxf(acc[r][q], 0.f), sum);
        part[r] = sum;
      }
      const float o = reduce_scatter8(part, lj);
      if (valid) a.out_a[(size_t)b * a.N + i] = o;
    }
    if (a.w1b) {
      hidden_tile(et, 32, s.wpb, li, lj, acc);
      float o[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float w2[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) w2[q] = s.w2b[c * kHidden + hidden_of(lj, q)];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float sum = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) sum = fmaf(w2[q], fmaxf(acc[r][q], 0.f), sum);
          part[r] = sum;
        }
        o[c] = reduce_scatter8(part, lj);
      }
      if (valid) {
        float* ob = a.out_b + ((size_t)b * a.N + i) * 3;
        ob[0] = o[0], ob[1] = o[1], ob[2] = o[2];
      }
    }
    if (a.tape) {  // padded slots of the last tile of a prompt are written too (zeros): the backward reads whole tiles
      const size_t slot = (size_t)b * a.n_pad + (size_t)t * kHfTile + warp * 32 + lane;
      float* ep = a.tape + (slot >> 5) * (kEncDim * 32) + (slot & 31);
#pragma unroll 8
      for (int k = 0; k < kEncDim; ++k) ep[k * 32] = et[k * 32 + lane];
    }
  }
}

// ------------------------------------------------------------------------------------------------- backward
constexpr int kEtStride = 36;
constexpr int kDhStride = kHfTile + 4;

struct HfBwdSmem {
  float wp[2][kWpSize];            // permuted W1 (head a, head b) for the hidden recompute
  float w1t[2][kHidden * kEncDim]; // W1^T: [hidden][enc] for dE = dH W1^T
  float w2a[kHidden];
  float w2b[3 * kHidden];          // [c][h]
  float et[4][kEncDim * kEtStride];
  float dht[kHidden * kDhStride];
  float pos[2][3][kHfTile];
  float dout[2][4][kHfTile];
  float g2a[kHidden];
  float g2b[3 * kHidden];
};

__global__ void __launch_bounds__(kHfThreads, 2)
hyper_field_bwd_kernel(const __grid_constant__ GridMeta gm, const HfArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HfBwdSmem& s = *reinterpret_cast<HfBwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int li = lane >> 3, lj = lane & 7;
  const int hg = tid >> 3, eg = tid & 7;
  const int tiles_per_b = a.n_pad / kHfTile;
  const int total = a.B * tiles_per_b;
  float2* g_table = reinterpret_cast<float2*>(a.g_table);
  const bool has[2] = {a.w1a != nullptr && a.d_a != nullptr, a.w1b != nullptr && a.d_b != nullptr};

  float accw[2][4][4];
  float g2a[8], g2b[3][8];
  auto zero_acc = [&]() {
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) accw[q][r][c] = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) g2a[q] = g2b[0][q] = g2b[1][q] = g2b[2][q] = 0.f;
  };
  zero_acc();

  // weight gradients of prompt b: registers -> global (dW1 is [enc][hidden], dW2a [hidden], dW2b [hidden][3])
  auto flush = [&](int b) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int idx = (eg + 8 * c) * kHidden + hg + 16 * r;
        if (has[0]) atomicAdd(a.g_w1a + (size_t)b * kWpSize + idx, accw[0][r][c]);
        if (has[1]) atomicAdd(a.g_w1b + (size_t)b * kWpSize + idx, accw[1][r][c]);
      }
    for (int i = tid; i < kHidden; i += kHfThreads) s.g2a[i] = 0.f;
    for (int i = tid; i < 3 * kHidden; i += kHfThreads) s.g2b[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v = g2a[q];
      v += __shfl_xor_sync(kFullMask, v, 8);
      v += __shfl_xor_sync(kFullMask, v, 16);
      if (li == 0) atomicAdd(&s.g2a[hidden_of(lj, q)], v);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float u = g2b[c][q];
        u += __shfl_xor_sync(kFullMask, u, 8);
        u += __shfl_xor_sync(kFullMask, u, 16);
        if (li == 0) atomicAdd(&s.g2b[c * kHidden + hidden_of(lj, q)], u);
      }
    }
    __syncthreads();
    if (has[0])
      for (int i = tid; i < kHidden; i += kHfThreads) atomicAdd(a.g_w2a + (size_t)b * kHidden + i, s.g2a[i]);
    if (has[1])
      for (int i = tid; i < 3 * kHidden; i += kHfThreads)
        atomicAdd(a.g_w2b + (size_t)b * 3 * kHidden + (i % kHidden) * 3 + i / kHidden, s.g2b[i]);
    zero_acc();
  };

  auto issue_tile = [&](int tile, int buf) {
    const int b = tile / tiles_per_b, t = tile - b * tiles_per_b;
    const float* src = a.tape + ((size_t)b * a.n_pad + (size_t)t * kHfTile) * kEncDim;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int q = tid + kHfThreads * r;
      const int sub = q >> 8, k = (q & 255) >> 3, s4 = q & 7;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.et[sub][k * kEtStride + s4 * 4]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)q * 4) : "memory");
    }
    const int i = t * kHfTile + tid;
    const int valid = i < a.N ? 4 : 0;
    const size_t pi = (size_t)b * a.N + (i < a.N ? i : 0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.pos[buf][c][tid]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(a.pts + pi * 3 + c), "r"(valid)
                   : "memory");
    }
    {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.dout[buf][0][tid]);
      const float* sp = a.d_a ? a.d_a + pi : a.pts;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(sp), "r"(a.d_a ? valid : 0)
                   : "memory");
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.dout[buf][1 + c][tid]);
      const float* sp = a.d_b ? a.d_b + pi * 3 + c : a.pts;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(sp), "r"(a.d_b ? valid : 0)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0, cur_b = -1;
  if ((int)blockIdx.x < total) issue_tile(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, buf ^= 1) {
    const int b = tile / tiles_per_b, t = tile - b * tiles_per_b;
    if (b != cur_b) {
      if (cur_b >= 0) flush(cur_b);
      __syncthreads();
#pragma unroll
      for (int net = 0; net < 2; ++net) {
        const float* w1 = net == 0 ? a.w1a : a.w1b;
        if (!w1) continue;
        w1 += (size_t)b * kWpSize;
        stage_w1_inout(s.wp[net], w1, tid, kHfThreads);
        for (int i = tid; i < kWpSize; i += kHfThreads) {
          const int h = i >> 5, e = i & 31;
          s.w1t[net][i] = w1[e * kHidden + h];
        }
      }
      if (a.w1a)
        for (int i = tid; i < kHidden; i += kHfThreads) s.w2a[i] = a.w2a[(size_t)b * kHidden + i];
      if (a.w1b)
        for (int i = tid; i < 3 * kHidden; i += kHfThreads)
          s.w2b[(i % 3) * kHidden + i / 3] = a.w2b[(size_t)b * 3 * kHidden + i];
      cur_b = b;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    float dE[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) dE[r][c] = 0.f;

#pragma unroll
    for (int net = 0; net < 2; ++net) {
      if (!has[net]) continue;  // uniform across the CTA
      {
        float acc[8][8];
        hidden_tile(s.et[warp], kEtStride, s.wp[net], li, lj, acc);
        const int s0 = warp * 32 + 8 * li;
        if (net == 0) {
          float w2[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) w2[q] = s.w2a[hidden_of(lj, q)];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float dr = s.dout[buf][0][s0 + r];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float h = fmaxf(acc[r][q], 0.f);
              g2a[q] = fmaf(h, dr, g2a[q]);
              acc[r][q] = acc[r][q] > 0.f ? w2[q] * dr : 0.f;
            }
          }
        } else {
          float w20[8], w21[8], w22[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            w20[q] = s.w2b[hidden_of(lj, q)];
            w21[q] = s.w2b[kHidden + hidden_of(lj, q)];
            w22[q] = s.w2b[2 * kHidden + hidden_of(lj, q)];
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float d0 = s.dout[buf][1][s0 + r], d1 = s.dout[buf][2][s0 + r], d2 = s.dout[buf][3][s0 + r];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float h = fmaxf(acc[r][q], 0.f);
              g2b[0][q] = fmaf(h, d0, g2b[0][q]);
              g2b[1][q] = fmaf(h, d1, g2b[1][q]);
              g2b[2][q] = fmaf(h, d2, g2b[2][q]);
              acc[r][q] = acc[r][q] > 0.f ? fmaf(w20[q], d0, fmaf(w21[q], d1, w22[q] * d2)) : 0.f;
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float* row = s.dht + hidden_of(lj, q) * kDhStride + s0;
          *reinterpret_cast<float4*>(row) = make_float4(acc[0][q], acc[1][q], acc[2][q], acc[3][q]);
          *reinterpret_cast<float4*>(row + 4) = make_float4(acc[4][q], acc[5][q], acc[6][q], acc[7][q]);
        }
      }
      __syncwarp();
      {
        const float* dcol = s.dht + warp * 32 + 8 * li;
        const float* wrow = s.w1t[net] + 4 * lj;
#pragma unroll 2
        for (int h = 0; h < kHidden; ++h) {
          const float4 d0 = *reinterpret_cast<const float4*>(dcol + h * kDhStride);
          const float4 d1 = *reinterpret_cast<const float4*>(dcol + h * kDhStride + 4);
          const float4 wv = *reinterpret_cast<const float4*>(wrow + h * kEncDim);
          const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            dE[r][0] = fmaf(d[r], wv.x, dE[r][0]);
            dE[r][1] = fmaf(d[r], wv.y, dE[r][1]);
            dE[r][2] = fmaf(d[r], wv.z, dE[r][2]);
            dE[r][3] = fmaf(d[r], wv.w, dE[r][3]);
          }
        }
      }
      __syncthreads();
#pragma unroll 1
      for (int sub = 0; sub < 4; ++sub) {
#pragma unroll 1
        for (int s4 = 0; s4 < 8; ++s4) {
          float4 dv[4], ev[4];
#pragma unroll
          for (int r = 0; r < 4; ++r)
            dv[r] = *reinterpret_cast<const float4*>(s.dht + (hg + 16 * r) * kDhStride + sub * 32 + s4 * 4);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            ev[c] = *reinterpret_cast<const float4*>(&s.et[sub][(eg + 8 * c) * kEtStride + s4 * 4]);
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float v = accw[net][r][c];
              v = fmaf(dv[r].x, ev[c].x, v);
              v = fmaf(dv[r].y, ev[c].y, v);
              v = fmaf(dv[r].z, ev[c].z, v);
              v = fmaf(dv[r].w, ev[c].w, v);
              accw[net][r][c] = v;
            }
        }
      }
      __syncthreads();
    }

    if (tile + (int)gridDim.x < total) issue_tile(tile + gridDim.x, buf ^ 1);

    // scatter: lane pairs, x-neighbour corners in one instruction (scatter_encoding_grads, render_tape.cuh)
    {
      const int sl0 = warp * 32 + 8 * li;
      scatter_encoding_grads(gm, g_table, dE, lj, &s.pos[buf][0][sl0], &s.pos[buf][1][sl0], &s.pos[buf][2][sl0],
                             a.N - (t * kHfTile + sl0));
    }
  }
  if (cur_b >= 0) flush(cur_b);
}

// The same backward with the contractions on the tensor cores (field_bwd_tc.cuh: mma.sync tf32, 3xTF32 for the hidden
// recompute) and the lane-pair scatter reading dE from shared memory; the default. Per-prompt weights are re-staged
// (transposed to [hidden][feature]) and the register accumulators flushed whenever a CTA crosses a prompt boundary.
__global__ void __launch_bounds__(fbtc::kTcThreads, 2)
hyper_field_bwd_tc_kernel(const __grid_constant__ GridMeta gm, const HfArgs a) {
  using namespace fbtc;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int hb = warp * 16;
  const int tiles_per_b = a.n_pad / kTcTile;
  const int total = a.B * tiles_per_b;
  float2* g_table = reinterpret_cast<float2*>(a.g_table);
  const bool has0 = a.w1a != nullptr && a.d_a != nullptr, has1 = a.w1b != nullptr && a.d_b != nullptr;

  float accw1[2][2][2][4];  // [net][feature block of 16][hidden block of 8][c]: dW1^T[e][h]
  float accw2[2][2][4];     // [net][hidden block of 8][c]: dW2^T[o][h], rows o = gq
  auto zero_acc = [&]() {
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          accw1[q][0][b][c] = accw1[q][1][b][c] = 0.f;
          accw2[q][b][c] = 0.f;
        }
  };
  zero_acc();

  // weight gradients of prompt b: registers -> global (dW1 is [feature][hidden], dW2a [hidden], dW2b [hidden][3])
  auto flush = [&](int b) {
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      if (!(net == 0 ? has0 : has1)) continue;
      float* gw1 = (net == 0 ? a.g_w1a : a.g_w1b) + (size_t)b * kWpSize;
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int e = 16 * mb + gq + 8 * (c >> 1), h = hb + 8 * nb + 2 * tq + (c & 1);
            atomicAdd(gw1 + e * kHidden + h, accw1[net][mb][nb][c]);
          }
      if (gq < (net == 0 ? 1 : 3)) {
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int h = hb + 8 * nb + 2 * tq + c;
            if (net == 0) atomicAdd(a.g_w2a + (size_t)b * kHidden + h, accw2[0][nb][c]);
            else atomicAdd(a.g_w2b + (size_t)b * 3 * kHidden + h * 3 + gq, accw2[1][nb][c]);
          }
      }
    }
    zero_acc();
  };

  auto issue_tile = [&](int tile, int buf) {
    const int b = tile / tiles_per_b, t = tile - b * tiles_per_b;
    const float* src = a.tape + ((size_t)b * a.n_pad + (size_t)t * kTcTile) * kEncDim;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int q = tid + kTcThreads * r;
      const int sub = q >> 8, k = (q & 255) >> 3, s4 = q & 7;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.et[sub][k * kTcEt + s4 * 4]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)q * 4) : "memory");
    }
    const int i = t * kTcTile + tid;
    const int valid = i < a.N ? 4 : 0;
    const size_t pi = (size_t)b * a.N + (i < a.N ? i : 0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.pos[buf][c][tid]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(a.pts + pi * 3 + c), "r"(valid)
                   : "memory");
    }
    {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.dout[buf][0][tid]);
      const float* sp = a.d_a ? a.d_a + pi : a.pts;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(sp), "r"(a.d_a ? valid : 0)
                   : "memory");
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.dout[buf][1 + c][tid]);
      const float* sp = a.d_b ? a.d_b + pi * 3 + c : a.pts;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(sp), "r"(a.d_b ? valid : 0)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0, cur_b = -1;
  if ((int)blockIdx.x < total) issue_tile(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, buf ^= 1) {
    const int b = tile / tiles_per_b, t = tile - b * tiles_per_b;
    if (b != cur_b) {
      if (cur_b >= 0) flush(cur_b);
      // (every warp is past the last barrier of the previous tile's contractions: nobody reads the old weights any more)
      for (int i = tid; i < kWpSize; i += kTcThreads) {
        const int e = i >> 6, h = i & 63;  // w1 [feature][hidden]
        if (a.w1a) s.w1[0][h * kTcW1 + e] = a.w1a[(size_t)b * kWpSize + i];
        if (a.w1b) s.w1[1][h * kTcW1 + e] = a.w1b[(size_t)b * kWpSize + i];
      }
      if (a.w1a)
        for (int i = tid; i < kHidden; i += kTcThreads) s.w2d[i] = a.w2a[(size_t)b * kHidden + i];
      if (a.w1b)
        for (int i = tid; i < 3 * kHidden; i += kTcThreads) s.w2f[(i % 3) * kHidden + i / 3] = a.w2b[(size_t)b * 3 * kHidden + i];
      cur_b = b;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    contract_tile(s, buf, warp, lane, has0, has1, accw1, accw2);
    if (tile + (int)gridDim.x < total) issue_tile(tile + gridDim.x, buf ^ 1);
    scatter_tile(gm, g_table, s, buf, warp, lane, a.N - t * kTcTile, 0xffff, nullptr, 0, 0, 1);
  }
  if (cur_b >= 0) flush(cur_b);
}

}  // namespace

int launch_hyper_field_fwd(const GridMeta& gm, const float* table, const float* pts, int B, int N, const float* w1a,
                           const float* w2a, const float* w1b, const float* w2b, float* out_a, float* out_b,
                           float* tape, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(hyper_field_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HfFwdSmem));
    attr_set = true;
  }
  HfArgs a;
  memset(&a, 0, sizeof(a));
  a.table = table, a.pts = pts, a.B = B, a.N = N, a.n_pad = (N + kHfTile - 1) / kHfTile * kHfTile;
  a.w1a = w1a, a.w2a = w2a, a.w1b = w1b, a.w2b = w2b, a.out_a = out_a, a.out_b = out_b, a.tape = tape;
  const int total = B * (a.n_pad / kHfTile);
  const int grid = max(1, min(kNumSMs * 4, total));
  hyper_field_fwd_kernel<<<grid, kHfThreads, sizeof(HfFwdSmem), stream>>>(gm, a);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("hyper_field_fwd");
  return SDB_OK;
}

int launch_hyper_field_bwd(const GridMeta& gm, const float* pts, int B, int N, const float* w1a, const float* w2a,
                           const float* w1b, const float* w2b, const float* tape, const float* d_a, const float* d_b,
                           float* g_table, float* g_w1a, float* g_w2a, float* g_w1b, float* g_w2b,
                           cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(hyper_field_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(HfBwdSmem));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(hyper_field_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(fbtc::TcSmem));
    if (e != cudaSuccess) {
      sdb_set_error("hyper_field_bwd: smem attribute: %s", cudaGetErrorString(e));
      return SDB_ERR_CUDA;
    }
    attr_set = true;
  }
  HfArgs a;
  memset(&a, 0, sizeof(a));
  a.pts = pts, a.B = B, a.N = N, a.n_pad = (N + kHfTile - 1) / kHfTile * kHfTile;
  a.w1a = w1a, a.w2a = w2a, a.w1b = w1b, a.w2b = w2b, a.tape = const_cast<float*>(tape);
  a.d_a = d_a, a.d_b = d_b, a.g_table = g_table, a.g_w1a = g_w1a, a.g_w2a = g_w2a, a.g_w1b = g_w1b, a.g_w2b = g_w2b;
  const int total = B * (a.n_pad / kHfTile);
  const int grid = max(1, min(kNumSMs * 2, total));
  // tensor-core contractions by default; SDB_HF_TC=0 selects the fp32 CUDA-core kernel (the cross-check)
  static const int use_tc = (getenv("SDB_HF_TC") && atoi(getenv("SDB_HF_TC")) == 0) ? 0 : 1;
  if (use_tc) hyper_field_bwd_tc_kernel<<<grid, fbtc::kTcThreads, sizeof(fbtc::TcSmem), stream>>>(gm, a);
  else hyper_field_bwd_kernel<<<grid, kHfThreads, sizeof(HfBwdSmem), stream>>>(gm, a);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("hyper_field_bwd");
  return SDB_OK;
}
