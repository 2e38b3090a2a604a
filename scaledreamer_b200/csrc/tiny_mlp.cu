// Bias-free ReLU MLP  D -> 64 -> 64 -> k  (k = 1 or 3) on rows of precomputed encodings, forward and backward, fp32 on the
// CUDA cores with the register-tile scheme of render_tape.cuh (lane = 8 rows x 8 units, operands as LDS.128).
// This is the VanillaMLP (threestudio/models/networks.py:214-251, n_neurons 64, n_hidden_layers 2) behind the triplane
// lookup of "Triplane-transformer-sdf" (custom/amortized/models/geometry/triplane_transformer.py:63-78, 176-190:
// sdf_network / feature_network on the 96-wide encoding); the reference runs it as three cuBLAS GEMMs + autograd.
//
// Forward: one warp per 32 rows; the encoding tile is transposed into shared memory, layer 1 -> ReLU -> the hidden
// tile is written back over it (unit-major) -> layer 2 -> ReLU -> layer 3 by a reduce-scatter.
// Backward (128-row CTA tiles): hidden recompute with ReLU masks, dH2 = mask2 * W3^T dY, dH1 = mask1 * dH2 W2,
// dX = dH1 W1, and the three weight gradients as CTA-level register tiles accumulated across tiles.
#include <cstdlib>

#include "../../include/sdb200.h"
#include "render_tape.cuh"

namespace {

constexpr int kMlpThreads = 128;
constexpr int kTile = 128;
constexpr int kTs = 36;         // floats per row of a per-warp [K][32] tile
constexpr int kCs = kTile + 4;  // floats per row of a CTA-wide [64][128] tile

// W [64 out][K in] (nn.Linear) -> Wp[k][half][j][c] = W[hidden_of(j, 4 half + c)][k]
__device__ __forceinline__ void stage_perm_out_in(float* __restrict__ Wp, const float* __restrict__ W, int K, int tid) {
  for (int idx = tid; idx < K * kHidden; idx += kMlpThreads) {
    const int k = idx >> 6, rem = idx & 63, half = rem >> 5, j = (rem & 31) >> 2, c = rem & 3;
    Wp[idx] = W[hidden_of(j, 4 * half + c) * K + k];
  }
}
// W [64 out][64 in] -> Wq[k = out][half][j][c] = W[k][hidden_of(j, 4 half + c)]  (contraction over the OUTPUT index)
__device__ __forceinline__ void stage_perm_in_out(float* __restrict__ Wq, const float* __restrict__ W, int tid) {
  for (int idx = tid; idx < kHidden * kHidden; idx += kMlpThreads) {
    const int k = idx >> 6, rem = idx & 63, half = rem >> 5, j = (rem & 31) >> 2, c = rem & 3;
    Wq[idx] = W[k * kHidden + hidden_of(j, 4 * half + c)];
  }
}

// acc[a][b] = sum_{k<K} T[k][8i + a] * Wp[k][unit hidden_of(j, b)]
__device__ __forceinline__ void tile_mm(const float* __restrict__ T, int ts, const float* __restrict__ Wp, int K, int i,
                                        int j, float (&acc)[8][8]) {
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
#pragma unroll 1
  for (int k = 0; k < K; ++k) {
    const float4 e0 = *reinterpret_cast<const float4*>(T + k * ts + 8 * i);
    const float4 e1 = *reinterpret_cast<const float4*>(T + k * ts + 8 * i + 4);
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

// rows [r0, r0 + 32) of x [n, D] -> T[k][row] (zero for rows >= n)
__device__ __forceinline__ void load_tile_T(const float* __restrict__ x, long long n, int D, long long r0, float* T, int ts,
                                            int lane) {
  const int d4 = D >> 2;
  for (int q = lane; q < 32 * d4; q += 32) {
    const int row = q / d4, c4 = q - row * d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + row < n) v = __ldg(reinterpret_cast<const float4*>(x + (r0 + row) * D) + c4);
    T[(4 * c4 + 0) * ts + row] = v.x;
    T[(4 * c4 + 1) * ts + row] = v.y;
    T[(4 * c4 + 2) * ts + row] = v.z;
    T[(4 * c4 + 3) * ts + row] = v.w;
  }
}

template <int KOUT>
__global__ void __launch_bounds__(kMlpThreads, 2)
mlp3_fwd_kernel(const float* __restrict__ x, long long n, int D, const float* __restrict__ w1,
                const float* __restrict__ w2, const float* __restrict__ w3, float* __restrict__ y) {
  extern __shared__ __align__(16) float sm[];
  float* wp1 = sm;                       // [D][64]
  float* wp2 = wp1 + D * kHidden;        // [64][64]
  float* sw3 = wp2 + kHidden * kHidden;  // [KOUT][64]
  const int rows_t = D > kHidden ? D : kHidden;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, li = lane >> 3, lj = lane & 7;
  float* T = sw3 + KOUT * kHidden + warp * rows_t * kTs;
  stage_perm_out_in(wp1, w1, D, tid);
  stage_perm_out_in(wp2, w2, kHidden, tid);
  for (int i = tid; i < KOUT * kHidden; i += kMlpThreads) sw3[i] = w3[i];
  __syncthreads();
  const long long n_tiles = (n + kTile - 1) / kTile;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long r0 = tile * kTile + warp * 32;
    __syncwarp();
    load_tile_T(x, n, D, r0, T, kTs, lane);
    __syncwarp();
    float acc[8][8];
    tile_mm(T, kTs, wp1, D, li, lj, acc);
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      float* row = T + hidden_of(lj, b) * kTs + 8 * li;
      *reinterpret_cast<float4*>(row) = make_float4(fmaxf(acc[0][b], 0.f), fmaxf(acc[1][b], 0.f), fmaxf(acc[2][b], 0.f),
                                                    fmaxf(acc[3][b], 0.f));
      *reinterpret_cast<float4*>(row + 4) = make_float4(fmaxf(acc[4][b], 0.f), fmaxf(acc[5][b], 0.f),
                                                        fmaxf(acc[6][b], 0.f), fmaxf(acc[7][b], 0.f));
    }
    __syncwarp();
    tile_mm(T, kTs, wp2, kHidden, li, lj, acc);
    float o[KOUT];
#pragma unroll
    for (int c = 0; c < KOUT; ++c) {
      float part[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        float s = 0.f;
#pragma unroll
        for (int b = 0; b < 8; ++b) s = fmaf(sw3[c * kHidden + hidden_of(lj, b)], fmaxf(acc[a][b], 0.f), s);
        part[a] = s;
      }
      o[c] = reduce_scatter8(part, lj);
    }
    if (r0 + lane < n) {
#pragma unroll
      for (int c = 0; c < KOUT; ++c) y[(r0 + lane) * KOUT + c] = o[c];
    }
  }
}

template <int KOUT>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp3_bwd_kernel(const float* __restrict__ x, long long n, int D, const float* __restrict__ w1,
                const float* __restrict__ w2, const float* __restrict__ w3, const float* __restrict__ dy,
                float* __restrict__ dx, int accumulate, float* __restrict__ g_w1, float* __restrict__ g_w2,
                float* __restrict__ g_w3) {
  extern __shared__ __align__(16) float sm[];
  float* wp1 = sm;                        // [D][64] permuted (layer-1 recompute)
  float* w1r = wp1 + D * kHidden;         // [64][D] row-major (dX = dH1 W1)
  float* wp2 = w1r + kHidden * D;         // [64 in][64 out] permuted (layer-2 recompute)
  float* wq2 = wp2 + kHidden * kHidden;   // [64 out][64 in] permuted (dH1 = dH2 W2)
  float* sw3 = wq2 + kHidden * kHidden;   // [KOUT][64]
  float* et = sw3 + KOUT * kHidden;       // [4][D][kTs]
  float* h1t = et + 4 * D * kTs;          // [64][kCs]
  float* dht = h1t + kHidden * kCs;       // [64][kCs]  dH2^T, later dH1^T
  float* sdy = dht + kHidden * kCs;       // [KOUT][128]
  float* sg3 = sdy + KOUT * kTile;        // [KOUT][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, li = lane >> 3, lj = lane & 7;
  stage_perm_out_in(wp1, w1, D, tid);
  stage_perm_out_in(wp2, w2, kHidden, tid);
  stage_perm_in_out(wq2, w2, tid);
  for (int i = tid; i < kHidden * D; i += kMlpThreads) w1r[i] = w1[i];
  for (int i = tid; i < KOUT * kHidden; i += kMlpThreads) sw3[i] = w3[i], sg3[i] = 0.f;
  __syncthreads();

  const int hg = tid >> 3, eg = tid & 7;  // weight-gradient tiles: 4 units {hg + 16 a} x {eg + 8 b}
  const int e_per = D >> 3;               // columns of dW1 per thread (<= 12)
  float gw1[4][12], gw2[4][8], gw3[KOUT][8];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 12; ++b) gw1[a][b] = 0.f;
#pragma unroll
    for (int b = 0; b < 8; ++b) gw2[a][b] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < KOUT; ++c)
#pragma unroll
    for (int b = 0; b < 8; ++b) gw3[c][b] = 0.f;

  float* T = et + warp * D * kTs;
  const int c0 = warp * 32 + 8 * li;  // this lane's first column in the CTA-wide tiles
  const long long n_tiles = (n + kTile - 1) / kTile;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long r0 = tile * kTile + warp * 32;
    load_tile_T(x, n, D, r0, T, kTs, lane);
    {
      const long long r = tile * kTile + tid;
#pragma unroll
      for (int c = 0; c < KOUT; ++c) sdy[c * kTile + tid] = r < n ? dy[r * KOUT + c] : 0.f;
    }
    __syncthreads();
    // ---- layer 1 recompute: mask1 + H1^T
    float acc[8][8];
    tile_mm(T, kTs, wp1, D, li, lj, acc);
    unsigned long long mask1 = 0ull;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        if (acc[a][b] > 0.f) mask1 |= 1ull << (b * 8 + a);
        acc[a][b] = fmaxf(acc[a][b], 0.f);
      }
      float* row = h1t + hidden_of(lj, b) * kCs + c0;
      *reinterpret_cast<float4*>(row) = make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
      *reinterpret_cast<float4*>(row + 4) = make_float4(acc[4][b], acc[5][b], acc[6][b], acc[7][b]);
    }
    __syncwarp();
    // ---- layer 2 recompute, dW3 partials, dH2^T
    tile_mm(h1t + warp * 32, kCs, wp2, kHidden, li, lj, acc);
    {
      float w3r[KOUT][8];
#pragma unroll
      for (int c = 0; c < KOUT; ++c)
#pragma unroll
        for (int b = 0; b < 8; ++b) w3r[c][b] = sw3[c * kHidden + hidden_of(lj, b)];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        float d[KOUT];
#pragma unroll
        for (int c = 0; c < KOUT; ++c) d[c] = sdy[c * kTile + c0 + a];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const float h = fmaxf(acc[a][b], 0.f);
          float v = 0.f;
#pragma unroll
          for (int c = 0; c < KOUT; ++c) {
            gw3[c][b] = fmaf(h, d[c], gw3[c][b]);
            v = fmaf(w3r[c][b], d[c], v);
          }
          acc[a][b] = acc[a][b] > 0.f ? v : 0.f;
        }
      }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      float* row = dht + hidden_of(lj, b) * kCs + c0;
      *reinterpret_cast<float4*>(row) = make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
      *reinterpret_cast<float4*>(row + 4) = make_float4(acc[4][b], acc[5][b], acc[6][b], acc[7][b]);
    }
    __syncwarp();
    // ---- dH1 = mask1 * (dH2 W2), kept in registers until dW2 has consumed dH2^T
    tile_mm(dht + warp * 32, kCs, wq2, kHidden, li, lj, acc);
#pragma unroll
    for (int b = 0; b < 8; ++b)
#pragma unroll
      for (int a = 0; a < 8; ++a)
        if (!((mask1 >> (b * 8 + a)) & 1ull)) acc[a][b] = 0.f;
    __syncthreads();
    // ---- dW2[o][i] += sum_s dH2^T[o][s] H1^T[i][s]   (thread: o = hg + 16 a, i = eg + 8 b)
#pragma unroll 1
    for (int s4 = 0; s4 < kTile / 4; ++s4) {
      float4 dv[4], hv[8];
#pragma unroll
      for (int a = 0; a < 4; ++a) dv[a] = *reinterpret_cast<const float4*>(dht + (hg + 16 * a) * kCs + s4 * 4);
#pragma unroll
      for (int b = 0; b < 8; ++b) hv[b] = *reinterpret_cast<const float4*>(h1t + (eg + 8 * b) * kCs + s4 * 4);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          float v = gw2[a][b];
          v = fmaf(dv[a].x, hv[b].x, v);
          v = fmaf(dv[a].y, hv[b].y, v);
          v = fmaf(dv[a].z, hv[b].z, v);
          v = fmaf(dv[a].w, hv[b].w, v);
          gw2[a][b] = v;
        }
    }
    __syncthreads();
    // ---- dH1^T over dH2^T
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      float* row = dht + hidden_of(lj, b) * kCs + c0;
      *reinterpret_cast<float4*>(row) = make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
      *reinterpret_cast<float4*>(row + 4) = make_float4(acc[4][b], acc[5][b], acc[6][b], acc[7][b]);
    }
    __syncthreads();
    // ---- dW1[h][e] += sum_s dH1^T[h][s] X^T[e][s]   (thread: h = hg + 16 a, e = eg + 8 b)
#pragma unroll 1
    for (int sub = 0; sub < 4; ++sub) {
      const float* es = et + sub * D * kTs;
#pragma unroll 1
      for (int s4 = 0; s4 < 8; ++s4) {
        float4 dv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
          dv[a] = *reinterpret_cast<const float4*>(dht + (hg + 16 * a) * kCs + sub * 32 + s4 * 4);
#pragma unroll
        for (int b = 0; b < 12; ++b) {
          if (b >= e_per) break;
          const float4 ev = *reinterpret_cast<const float4*>(es + (eg + 8 * b) * kTs + s4 * 4);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            float v = gw1[a][b];
            v = fmaf(dv[a].x, ev.x, v);
            v = fmaf(dv[a].y, ev.y, v);
            v = fmaf(dv[a].z, ev.z, v);
            v = fmaf(dv[a].w, ev.w, v);
            gw1[a][b] = v;
          }
        }
      }
    }
    // ---- dX[s][e] = sum_h dH1^T[h][s] W1[h][e], 32 columns per pass (lane: rows 8 li + a, columns 32 p + 4 lj + c)
    if (dx) {
      for (int pass = 0; pass * 32 < D; ++pass) {
        const int e0 = pass * 32 + 4 * lj;
        if (e0 >= D) continue;
        float de[8][4];
#pragma unroll
        for (int a = 0; a < 8; ++a) de[a][0] = de[a][1] = de[a][2] = de[a][3] = 0.f;
#pragma unroll 2
        for (int h = 0; h < kHidden; ++h) {
          const float4 d0 = *reinterpret_cast<const float4*>(dht + h * kCs + c0);
          const float4 d1 = *reinterpret_cast<const float4*>(dht + h * kCs + c0 + 4);
          const float4 wv = *reinterpret_cast<const float4*>(w1r + h * D + e0);
          const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            de[a][0] = fmaf(d[a], wv.x, de[a][0]);
            de[a][1] = fmaf(d[a], wv.y, de[a][1]);
            de[a][2] = fmaf(d[a], wv.z, de[a][2]);
            de[a][3] = fmaf(d[a], wv.w, de[a][3]);
          }
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const long long r = tile * kTile + c0 + a;
          if (r >= n) continue;
          float4* dst = reinterpret_cast<float4*>(dx + r * D + e0);
          float4 v = make_float4(de[a][0], de[a][1], de[a][2], de[a][3]);
          if (accumulate) {
            const float4 o = *dst;
            v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
          }
          *dst = v;
        }
      }
    }
    __syncthreads();  // the tiles are rewritten by the next iteration
  }

  // ---- flush weight gradients
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 12; ++b)
      if (b < e_per) atomicAdd(g_w1 + (hg + 16 * a) * D + eg + 8 * b, gw1[a][b]);
#pragma unroll
    for (int b = 0; b < 8; ++b) atomicAdd(g_w2 + (hg + 16 * a) * kHidden + eg + 8 * b, gw2[a][b]);
  }
#pragma unroll
  for (int c = 0; c < KOUT; ++c)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      float v = gw3[c][b];
      v += __shfl_xor_sync(kFullMask, v, 8);
      v += __shfl_xor_sync(kFullMask, v, 16);
      if (li == 0) atomicAdd(&sg3[c * kHidden + hidden_of(lj, b)], v);
    }
  __syncthreads();
  for (int i = tid; i < KOUT * kHidden; i += kMlpThreads) atomicAdd(g_w3 + i, sg3[i]);
}

size_t fwd_smem(int D, int k) {
  const int rows_t = D > kHidden ? D : kHidden;
  return sizeof(float) * ((size_t)D * kHidden + kHidden * kHidden + k * kHidden + 4 * (size_t)rows_t * kTs);
}
size_t bwd_smem(int D, int k) {
  return sizeof(float) * (2 * (size_t)D * kHidden + 2 * kHidden * kHidden + 2 * k * kHidden + 4 * (size_t)D * kTs +
                          2 * (size_t)kHidden * kCs + (size_t)k * kTile);
}

}  // namespace

// mlp3_tc.cu: the same forward on the tensor cores (3xTF32), the default
int launch_mlp3_fwd_tc(const float* x, long long n, int D, const float* w1, const float* w2, const float* w3, int n_out,
                       float* y, cudaStream_t s);
int launch_mlp3_bwd_tc(const float* x, long long n, int D, const float* w1, const float* w2, const float* w3, int n_out,
                       const float* dy, float* dx, int accumulate, float* g_w1, float* g_w2, float* g_w3, cudaStream_t s);

extern "C" {

int sdb_mlp3_forward(const float* x, long long n, int d_in, const float* w1, const float* w2, const float* w3,
                     int n_out, float* y, void* stream) {
  SDB_CHECK_ARG(w1 && w2 && w3 && n >= 0 && (n == 0 || (x && y)), "mlp3_forward: bad arguments");
  SDB_CHECK_ARG(d_in >= 8 && d_in <= 96 && d_in % 8 == 0, "mlp3: d_in must be a multiple of 8 in [8, 96]");
  SDB_CHECK_ARG(n_out == 1 || n_out == 3, "mlp3: n_out must be 1 or 3");
  if (n == 0) return SDB_OK;
  static const int use_tc = (getenv("SDB_MLP3_TC") && atoi(getenv("SDB_MLP3_TC")) == 0) ? 0 : 1;
  if (use_tc) return launch_mlp3_fwd_tc(x, n, d_in, w1, w2, w3, n_out, y, (cudaStream_t)stream);
  const size_t smem = fwd_smem(d_in, n_out);
  const long long tiles = (n + kTile - 1) / kTile;
  const int grid = (int)(tiles < 2LL * kNumSMs ? tiles : 2LL * kNumSMs);
  cudaStream_t s = (cudaStream_t)stream;
  if (n_out == 1) {
    cudaFuncSetAttribute(mlp3_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mlp3_fwd_kernel<1><<<grid, kMlpThreads, smem, s>>>(x, n, d_in, w1, w2, w3, y);
  } else {
    cudaFuncSetAttribute(mlp3_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mlp3_fwd_kernel<3><<<grid, kMlpThreads, smem, s>>>(x, n, d_in, w1, w2, w3, y);
  }
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("mlp3_forward");
  return SDB_OK;
}

int sdb_mlp3_backward(const float* x, long long n, int d_in, const float* w1, const float* w2, const float* w3,
                      int n_out, const float* d_y, float* d_x, int accumulate_dx, float* g_w1, float* g_w2,
                      float* g_w3, void* stream) {
  SDB_CHECK_ARG(w1 && w2 && w3 && g_w1 && g_w2 && g_w3 && n >= 0 && (n == 0 || (x && d_y)),
                "mlp3_backward: bad arguments");
  SDB_CHECK_ARG(d_in >= 8 && d_in <= 96 && d_in % 8 == 0, "mlp3: d_in must be a multiple of 8 in [8, 96]");
  SDB_CHECK_ARG(n_out == 1 || n_out == 3, "mlp3: n_out must be 1 or 3");
  if (n == 0) return SDB_OK;
  static const int use_tc = (getenv("SDB_MLP3_TC") && atoi(getenv("SDB_MLP3_TC")) == 0) ? 0 : 1;
  if (use_tc)
    return launch_mlp3_bwd_tc(x, n, d_in, w1, w2, w3, n_out, d_y, d_x, accumulate_dx, g_w1, g_w2, g_w3, (cudaStream_t)stream);
  const size_t smem = bwd_smem(d_in, n_out);
  const long long tiles = (n + kTile - 1) / kTile;
  const int grid = (int)(tiles < (long long)kNumSMs ? tiles : (long long)kNumSMs);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e;
  if (n_out == 1) {
    e = cudaFuncSetAttribute(mlp3_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      mlp3_bwd_kernel<1><<<grid, kMlpThreads, smem, s>>>(x, n, d_in, w1, w2, w3, d_y, d_x, accumulate_dx, g_w1, g_w2, g_w3);
  } else {
    e = cudaFuncSetAttribute(mlp3_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      mlp3_bwd_kernel<3><<<grid, kMlpThreads, smem, s>>>(x, n, d_in, w1, w2, w3, d_y, d_x, accumulate_dx, g_w1, g_w2, g_w3);
  }
  if (e != cudaSuccess) {
    sdb_set_error("mlp3_backward: smem attribute (%zu bytes): %s", smem, cudaGetErrorString(e));
    return SDB_ERR_CUDA;
  }
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("mlp3_backward");
  return SDB_OK;
}

}  // extern "C"
