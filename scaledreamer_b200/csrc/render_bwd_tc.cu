// Field backward of the taped NeRF renderer with the 32-64-{1,3} MLP contractions on the tensor cores
// (mma.sync.m16n8k8 tf32, fp32 accumulation) -- same interface, inputs and outputs as render_field_bwd_kernel
// (render_bwd2.cu), which keeps the fp32 CUDA-core version as the cross-check (SDB_FB_TC=0).
//
// Why: ncu on the fp32 kernel (profiles/r1g_ncu_full_render_bwd.csv) shows 255 registers, 12.5 % active warps and the
// FMA pipe at 41 %; counting LDS.128 per FFMA shows the register-tiled contractions are shared-memory-bandwidth bound
// (dE = dH W1: 3 LDS.128 per 32 FFMA; dW1 += dH^T E: 8 per 64). 12.3 k FMA per kept sample x 4.8 M samples is 2.9 ms of
// the 4.3 ms kernel. An m16n8k8 fragment needs 6 LDS.32 per 1024 FMA.
//
// Per 128-sample tile and network (density, feature), warp w owns samples 32 w .. 32 w + 31:
//   (1) H = E W1^T           M = 32 samples, N = 64 hidden, K = 32 features; 3xTF32 (hi/lo split of both operands): the
//                            ReLU mask must equal the forward's fp32 one, a plain tf32 recompute would flip ~0.2 % of it
//   (2) relu(H) -> smem;     dW2^T[o][h] += dout^T relu(H)   M = 16 (1 or 3 used), N = this warp's 16 hidden units,
//                            K = the tile's 128 samples
//   (3) dH = mask * (dout W2) in the accumulator registers; dE += dH W1: the C fragment of an 8-unit block IS the A
//       fragment of the next product once the k index is permuted (k position t <-> unit 2t, t+4 <-> 2t+1), so dH never
//       leaves the registers for this product
//   (4) dH -> smem;          dW1^T[e][h] += E^T dH          M = 32 features, N = this warp's 16 hidden units, K = 128
//   (5) dE (both networks) -> smem, then the trilinear red.v2 scatter of render_field_bwd_kernel (lane = 8 consecutive
//       samples x 2 levels, runs inside one cell merged).
// tf32 rounds operands to 11 significant bits (2^-12 relative); products are exact in fp32 and accumulate in fp32. The
// gradients are smooth in those operands, so the error stays two orders below the 5e-3 parity bound (measured in
// tests/test_render_gpu.py against the oracle and against the fp32 kernel).
#include <cstdlib>

#include "render_tape.cuh"

namespace {

constexpr int kTcThreads = 128;
constexpr int kTcTile = 128;
constexpr int kTcEt = 40;   // floats per feature row of a 32-sample encoding tile: 8 t + g hits 32 distinct banks
constexpr int kTcW1 = 36;   // floats per W1 row: 4 g + t and 8 t + g both hit 32 distinct banks
constexpr int kTcDh = 72;   // floats per sample row of the hidden-gradient tile: 8 t + g again

struct TcSmem {
  float et[4][kEncDim * kTcEt];          // encodings, feature-major per 32-sample sub-tile (cp.async from the tape)
  float w1[2][kHidden * kTcW1];          // W1 [hidden][feature] of both networks
  float dh[kTcTile * kTcDh];             // relu(H), then dH, then dE: [sample][hidden or feature]
  float pos[2][3][kTcTile];
  float dout[2][4][kTcTile];             // d raw, d o0..2 (double-buffered with pos)
  float w2d[kHidden];
  float w2f[3 * kHidden];
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
// D += A B, m16n8k8, A row-major 16x8, B column-major 8x8 (tf32), D fp32
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Coarse levels are scattered into one of `n_rep` private copies (copy = CTA index mod n_rep) and summed afterwards:
// every ray crosses the same few hundred coarse cells in front of the object, and fp32 atomics on ONE address retire at
// about one per 100 ns chip-wide -- levels 0-2 alone (6 % of the atomics) cost 2.7 ms of a 7.3 ms launch before this.
constexpr int kRepLevels = 4;  // levels 0..3: 175 k entries (1.4 MB per copy) for the 16-level C2 grid
constexpr int kRepCopies = 16;

__global__ void __launch_bounds__(256)
fold_replicas_kernel(float2* __restrict__ g_table, float2* __restrict__ rep, int rep_entries, int n_rep) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rep_entries; i += gridDim.x * blockDim.x) {
    float sx = 0.f, sy = 0.f;
    for (int r = 0; r < n_rep; ++r) {  // fixed order
      float2* q = rep + (size_t)r * rep_entries + i;
      const float2 v = *q;
      sx += v.x;
      sy += v.y;
      *q = make_float2(0.f, 0.f);  // leaves the copies cleared for the next launch
    }
    g_table[i].x += sx;
    g_table[i].y += sy;
  }
}

__global__ void __launch_bounds__(kTcThreads, 2)
render_field_bwd_tc_kernel(const __grid_constant__ FieldMeta f, const FieldPtrs p, const FieldGrads g,
                           const RenderTape tape, const int scatter_on, float2* __restrict__ rep, const int rep_entries,
                           const int n_rep) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TcSmem& s = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;  // fragment coordinates: group id, thread in group
  for (int i = tid; i < kHidden * kEncDim; i += kTcThreads) {
    const int h = i / kEncDim, e = i - h * kEncDim;
    s.w1[0][h * kTcW1 + e] = p.w1d[i];
    s.w1[1][h * kTcW1 + e] = p.w1f[i];
  }
  for (int i = tid; i < kHidden; i += kTcThreads) s.w2d[i] = p.w2d[i];
  for (int i = tid; i < 3 * kHidden; i += kTcThreads) s.w2f[i] = p.w2f[i];
  __syncthreads();

  const int n = min(__ldg(tape.counter), tape.capacity);
  const int n_tiles = (n + kTcTile - 1) / kTcTile;
  const size_t cap = (size_t)tape.capacity;
  float2* g_table = reinterpret_cast<float2*>(g.table);

  // persistent accumulators (this warp's 16 hidden units hb .. hb + 15 of both networks)
  const int hb = warp * 16;
  float accw1[2][2][2][4];  // [net][feature block of 16][hidden block of 8][c]: dW1^T[e][h]
  float accw2[2][2][4];     // [net][hidden block of 8][c]: dW2^T[o][h], rows o = gq (and gq + 8: always zero)
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) accw1[q][a][b][c] = 0.f;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) accw2[q][b][c] = 0.f;

  // scatter mapping of render_field_bwd_kernel: lane (li, lj) owns samples 8 li + a of its warp's 32, levels 2 lj, 2 lj + 1
  const int li = lane >> 3, lj = lane & 7;
  float lv_scale[2];
  uint32_t lv_res[2], lv_size[2], lv_off[2], lv_hashed[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int l = 2 * lj + q;
    lv_scale[q] = f.grid.scale[l];
    lv_res[q] = f.grid.res[l];
    lv_size[q] = f.grid.size[l];
    lv_off[q] = f.grid.offset[l];
    lv_hashed[q] = f.grid.hashed[l];
  }

  auto issue_tile = [&](int tile, int buf) {
    const int base = tile * kTcTile;
    const float* src = tape.enc + (size_t)base * kEncDim;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int q = tid + kTcThreads * r;  // float4 index within the 4 x 32 x 32 block
      const int sub = q >> 8, k = (q & 255) >> 3, s4 = q & 7;
      const int slot = base + sub * 32 + s4 * 4;
      const int valid = min(max(n - slot, 0), 4) * 4;  // bytes
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.et[sub][k * kTcEt + s4 * 4]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src + (size_t)q * 4), "r"(valid)
                   : "memory");
    }
    const int slot = base + tid;
    const int valid = slot < n ? 4 : 0;
    const size_t off = slot < n ? (size_t)slot : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.pos[buf][c][tid]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(tape.pos + c * cap + off), "r"(valid)
                   : "memory");
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.dout[buf][c][tid]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(tape.sample + c * cap + off),
                   "r"(valid)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if ((int)blockIdx.x < n_tiles) issue_tile(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    const int base = tile * kTcTile;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    float dE[2][4][4];  // [sample block of 16][feature block of 8][c]: d enc of this warp's 32 samples
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int c = 0; c < 4; ++c) dE[mb][nb][c] = 0.f;

    const float* etw = s.et[warp];
#pragma unroll
    for (int net = 0; net < 2; ++net) {  // unrolled: the persistent accumulators are indexed by `net`
      const float* w1 = s.w1[net];
      // ---- (1) H = E W1^T for this warp's samples: A[s][k] = et[k][s], B[k][h] = W1[h][k]; 3xTF32
      float H[2][8][4];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
#pragma unroll
          for (int c = 0; c < 4; ++c) H[mb][nb][c] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const float* a = etw + (8 * ks + tq) * kTcEt + 16 * mb + gq;
          split_tf32(a[0], ah[mb][0], al[mb][0]);
          split_tf32(a[8], ah[mb][1], al[mb][1]);
          split_tf32(a[4 * kTcEt], ah[mb][2], al[mb][2]);
          split_tf32(a[4 * kTcEt + 8], ah[mb][3], al[mb][3]);
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          const float* b = w1 + (8 * nb + gq) * kTcW1 + 8 * ks + tq;
          uint32_t bh[2], bl[2];
          split_tf32(b[0], bh[0], bl[0]);
          split_tf32(b[4], bh[1], bl[1]);
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            mma_tf32(H[mb][nb], al[mb], bh);
            mma_tf32(H[mb][nb], ah[mb], bl);
            mma_tf32(H[mb][nb], ah[mb], bh);
          }
        }
      }
      // ---- (2) relu(H) -> smem [sample][hidden]; dW2^T += dout^T relu(H) over the whole tile
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          float* d0 = s.dh + (warp * 32 + 16 * mb + gq) * kTcDh + 8 * nb + 2 * tq;
          *reinterpret_cast<float2*>(d0) = make_float2(fmaxf(H[mb][nb][0], 0.f), fmaxf(H[mb][nb][1], 0.f));
          *reinterpret_cast<float2*>(d0 + 8 * kTcDh) = make_float2(fmaxf(H[mb][nb][2], 0.f), fmaxf(H[mb][nb][3], 0.f));
        }
      __syncthreads();
      {
        const int n_out = net == 0 ? 1 : 3;
        const float* dsrc = &s.dout[buf][net == 0 ? 0 : 1][0];  // row o of dout^T: dout[buf][first + o][sample]
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
          uint32_t a[4];
          // A[o][s]: rows gq (a0, a2) and gq + 8 (a1, a3: never an output)
          a[0] = gq < n_out ? to_tf32(dsrc[gq * kTcTile + 8 * ks + tq]) : 0u;
          a[2] = gq < n_out ? to_tf32(dsrc[gq * kTcTile + 8 * ks + tq + 4]) : 0u;
          a[1] = 0u;
          a[3] = 0u;
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            const float* b = s.dh + (8 * ks + tq) * kTcDh + hb + 8 * nb + gq;
            uint32_t bb[2] = {to_tf32(b[0]), to_tf32(b[4 * kTcDh])};
            mma_tf32(accw2[net][nb], a, bb);
          }
        }
      }
      __syncthreads();
      // ---- (3) dH in the accumulator registers; dE += dH W1 with the permuted k index
      {
        float dr[2][2][3];  // d out of rows (mb, half) for up to three outputs
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int sl = warp * 32 + 16 * mb + 8 * hf + gq;
            if (net == 0) {
              dr[mb][hf][0] = s.dout[buf][0][sl];
              dr[mb][hf][1] = dr[mb][hf][2] = 0.f;
            } else {
              dr[mb][hf][0] = s.dout[buf][1][sl];
              dr[mb][hf][1] = s.dout[buf][2][sl];
              dr[mb][hf][2] = s.dout[buf][3][sl];
            }
          }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          float wv[2][3];  // W2[o][h] for h = 8 nb + 2 tq, + 1
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int h = 8 * nb + 2 * tq + c;
            if (net == 0) {
              wv[c][0] = s.w2d[h];
              wv[c][1] = wv[c][2] = 0.f;
            } else {
              wv[c][0] = s.w2f[h];
              wv[c][1] = s.w2f[kHidden + h];
              wv[c][2] = s.w2f[2 * kHidden + h];
            }
          }
#pragma unroll
          for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int hf = c >> 1, cc = c & 1;
              const float up = fmaf(wv[cc][0], dr[mb][hf][0], fmaf(wv[cc][1], dr[mb][hf][1], wv[cc][2] * dr[mb][hf][2]));
              H[mb][nb][c] = H[mb][nb][c] > 0.f ? up : 0.f;
            }
        }
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {  // hidden units 8 ks .. 8 ks + 7; k position t <-> unit 2t, t + 4 <-> 2t + 1
        uint32_t a[2][4];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          a[mb][0] = to_tf32(H[mb][ks][0]);
          a[mb][1] = to_tf32(H[mb][ks][2]);
          a[mb][2] = to_tf32(H[mb][ks][1]);
          a[mb][3] = to_tf32(H[mb][ks][3]);
        }
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          const float* b = w1 + (8 * ks + 2 * tq) * kTcW1 + 8 * nb + gq;
          uint32_t bb[2] = {to_tf32(b[0]), to_tf32(b[kTcW1])};
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) mma_tf32(dE[mb][nb], a[mb], bb);
        }
      }
      // ---- (4) dH -> smem [sample][hidden]; dW1^T += E^T dH over the whole tile
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          float* d0 = s.dh + (warp * 32 + 16 * mb + gq) * kTcDh + 8 * nb + 2 * tq;
          *reinterpret_cast<float2*>(d0) = make_float2(H[mb][nb][0], H[mb][nb][1]);
          *reinterpret_cast<float2*>(d0 + 8 * kTcDh) = make_float2(H[mb][nb][2], H[mb][nb][3]);
        }
      __syncthreads();
#pragma unroll 4
      for (int ks = 0; ks < 16; ++ks) {  // samples 8 ks .. 8 ks + 7 of the tile (sub-tile ks >> 2)
        const float* ets = s.et[ks >> 2] + 8 * (ks & 3) + tq;
        uint32_t a[2][4];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {  // A[e][s] = et[e][s]: rows e = 16 mb + gq (+ 8), columns s = tq (+ 4)
          const float* ap = ets + (16 * mb + gq) * kTcEt;
          a[mb][0] = to_tf32(ap[0]);
          a[mb][1] = to_tf32(ap[8 * kTcEt]);
          a[mb][2] = to_tf32(ap[4]);
          a[mb][3] = to_tf32(ap[8 * kTcEt + 4]);
        }
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          const float* b = s.dh + (8 * ks + tq) * kTcDh + hb + 8 * nb + gq;
          uint32_t bb[2] = {to_tf32(b[0]), to_tf32(b[4 * kTcDh])};
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) mma_tf32(accw1[net][mb][nb], a[mb], bb);
        }
      }
      __syncthreads();  // dh is rewritten by the next network / by dE below
    }

    // ---- (5) dE -> smem [sample][feature] (reusing dh), then the scatter
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        float* d0 = s.dh + (warp * 32 + 16 * mb + gq) * kTcDh + 8 * nb + 2 * tq;
        *reinterpret_cast<float2*>(d0) = make_float2(dE[mb][nb][0], dE[mb][nb][1]);
        *reinterpret_cast<float2*>(d0 + 8 * kTcDh) = make_float2(dE[mb][nb][2], dE[mb][nb][3]);
      }
    __syncwarp();  // a warp scatters the samples it wrote

    // the encoding tiles are free (every thread passed the last barrier): stream the next tile in behind the scatter
    if (tile + (int)gridDim.x < n_tiles) issue_tile(tile + gridDim.x, buf ^ 1);

    if (scatter_on) {
      // Lane PAIRS: both lanes of a pair walk the same 8 consecutive samples x 4 levels, one takes the four corners with
      // x = cx, the other those with x = cx + 1. The two entries are neighbours in the table three times out of four
      // (dense levels: idx, idx + 1; hashed levels: the hash only XORs x in), and two lanes of ONE instruction on one
      // 32-byte sector cost one sector operation: 327 G lane-ops/s instead of 193 G/s (tools/red_probe.cu, mode 3).
      const int xp = lane & 1, pr = lane >> 1;
      const int sg = pr >> 2, lq = pr & 3;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int lvl = 4 * lq + q;
        if (!((scatter_on >> lvl) & 1)) continue;  // diagnostics: SDB_FB_LEVELS masks levels out
        const float sc = f.grid.scale[lvl];
        const uint32_t res = f.grid.res[lvl], size = f.grid.size[lvl], hashed = f.grid.hashed[lvl];
        float2* tl = (rep && lvl < kRepLevels ? rep + (size_t)(blockIdx.x % n_rep) * rep_entries : g_table) + f.grid.offset[lvl];
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
          const int sl = warp * 32 + 8 * sg + a;
          const float2 gxy = *reinterpret_cast<const float2*>(s.dh + sl * kTcDh + 2 * lvl);
          if (base + sl >= n || (gxy.x == 0.f && gxy.y == 0.f)) continue;
          const LevelCell c = level_cell(sc, s.pos[buf][0][sl], s.pos[buf][1][sl], s.pos[buf][2][sl]);
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
            ax[j] = fmaf(w, gxy.x, ax[j]);
            ay[j] = fmaf(w, gxy.y, ay[j]);
          }
        }
        if (open) flush();
      }
    }
    // (the loop-top barrier orders this tile's reads of dh / pos against the next tile's writes)
  }

  // ---- flush weight gradients: dW1^T[e][h] -> g.w1[h][e]; dW2^T[o][h] -> g.w2[o][h]
#pragma unroll
  for (int net = 0; net < 2; ++net) {
    float* gw1 = net == 0 ? g.w1d : g.w1f;
    float* gw2 = net == 0 ? g.w2d : g.w2f;
    const int n_out = net == 0 ? 1 : 3;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int e = 16 * mb + gq + 8 * (c >> 1), h = hb + 8 * nb + 2 * tq + (c & 1);
          atomicAdd(gw1 + h * kEncDim + e, accw1[net][mb][nb][c]);
        }
    if (gq < n_out) {
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int c = 0; c < 2; ++c) atomicAdd(gw2 + gq * kHidden + hb + 8 * nb + 2 * tq + c, accw2[net][nb][c]);
    }
  }
}

}  // namespace

int launch_render_field_bwd_tc(const FieldMeta& f, const FieldPtrs& p, const FieldGrads& g, const RenderTape& tape,
                               int scatter_on, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(render_field_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(TcSmem));
    if (e != cudaSuccess) {
      sdb_set_error("render_field_bwd_tc: smem attribute: %s", cudaGetErrorString(e));
      return SDB_ERR_CUDA;
    }
    attr_set = true;
  }
  const int max_tiles = (tape.capacity + kTcTile - 1) / kTcTile;
  const int grid = max(1, min(kNumSMs * 2, max_tiles));
  // replica buffer for the coarse levels: allocated (and cleared) once per device, kept cleared by the fold kernel
  static float2* rep_buf[16] = {};
  static int rep_cap[16] = {};
  static const bool rep_on = !(getenv("SDB_FB_REPLICAS") && atoi(getenv("SDB_FB_REPLICAS")) == 0);
  float2* rep = nullptr;
  int rep_entries = 0;
  if (rep_on && f.grid.n_levels > kRepLevels) {
    int dev = 0;
    cudaGetDevice(&dev);
    rep_entries = (int)f.grid.offset[kRepLevels];
    if (dev < 16 && rep_entries > 0) {
      if (rep_cap[dev] < rep_entries) {
        if (rep_buf[dev]) cudaFree(rep_buf[dev]);
        const size_t bytes = (size_t)kRepCopies * rep_entries * sizeof(float2);
        if (cudaMalloc(&rep_buf[dev], bytes) != cudaSuccess || cudaMemset(rep_buf[dev], 0, bytes) != cudaSuccess) {
          sdb_set_error("render_field_bwd_tc: replica buffer: %s", cudaGetErrorString(cudaGetLastError()));
          return SDB_ERR_CUDA;
        }
        rep_cap[dev] = rep_entries;
      }
      rep = rep_buf[dev];
    }
  }
  render_field_bwd_tc_kernel<<<grid, kTcThreads, sizeof(TcSmem), stream>>>(f, p, g, tape, scatter_on, rep, rep_entries,
                                                                          kRepCopies);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("render_field_bwd_tc");
  if (rep) {
    fold_replicas_kernel<<<min(kNumSMs * 4, (rep_entries + 255) / 256), 256, 0, stream>>>(reinterpret_cast<float2*>(g.table), rep,
                                                                                         rep_entries, kRepCopies);
    SDB_COUNT_LAUNCH();
    SDB_CHECK_LAUNCH("fold_replicas");
  }
  return SDB_OK;
}
