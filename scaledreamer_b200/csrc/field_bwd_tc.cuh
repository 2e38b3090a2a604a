// Shared pieces of the tensor-core field-backward kernels (render_bwd_tc.cu: the NeRF renderer's tape; hyper_field.cu:
// the prompt-conditioned field of the amortized generators): the shared-memory tile layout, the mma.sync tf32 helpers, the
// per-tile contraction core and the lane-pair scatter that reads dE back from shared memory.
//
// Per 128-sample tile and network (head 0: 64 -> 1, head 1: 64 -> 3), warp w owns samples 32 w .. 32 w + 31:
//   (1) H = E W1^T           M = 32 samples, N = 64 hidden, K = 32 features; 3xTF32 (hi/lo split of both operands): the
//                            ReLU mask must equal the forward's fp32 one, a plain tf32 recompute would flip ~0.2 % of it
//   (2) relu(H) -> smem;     dW2^T[o][h] += dout^T relu(H)   M = 16 (1 or 3 used), N = this warp's 16 hidden units,
//                            K = the tile's 128 samples
//   (3) dH = mask * (dout W2) in the accumulator registers; dE += dH W1: the C fragment of an 8-unit block IS the A
//       fragment of the next product once the k index is permuted (k position t <-> unit 2t, t+4 <-> 2t+1), so dH never
//       leaves the registers for this product
//   (4) dH -> smem;          dW1^T[e][h] += E^T dH          M = 32 features, N = this warp's 16 hidden units, K = 128
//   (5) dE (both networks) -> smem [sample][feature]
#pragma once
#include "render_tape.cuh"

namespace fbtc {

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

// Steps (1)-(5) for the tile staged in s.et / s.dout[buf]; s.w1 holds W1 [hidden][feature] (row stride kTcW1), s.w2d /
// s.w2f the output layers ([64] and [3][64]). accw1[net][feature block of 16][hidden block of 8][c] = dW1^T[e][h] and
// accw2[net][hidden block of 8][c] = dW2^T[o][h] (rows o = gq) accumulate across tiles for this warp's 16 hidden units.
// A head that is absent (has0 / has1, uniform across the CTA) is skipped. Ends with dE of this warp's samples in s.dh.
__device__ __forceinline__ void contract_tile(TcSmem& s, const int buf, const int warp, const int lane, const bool has0,
                                              const bool has1, float (&accw1)[2][2][2][4], float (&accw2)[2][2][4]) {
  const int gq = lane >> 2, tq = lane & 3;  // fragment coordinates: group id, thread in group
  const int hb = warp * 16;
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
      if (!(net == 0 ? has0 : has1)) continue;
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
}

// Lane PAIRS: both lanes of a pair walk the same 8 consecutive samples x 4 levels of the dE tile in s.dh, one takes the
// four corners with x = cx, the other those with x = cx + 1. The two entries are neighbours in the table three times out
// of four (dense levels: idx, idx + 1; hashed levels: the hash only XORs x in), and two lanes of ONE instruction on one
// 32-byte sector cost one sector operation: 327 G lane-ops/s instead of 193 G/s (tools/red_probe.cu, mode 3). Samples
// that stay in one cell are summed per corner in registers first. n_valid = samples of the tile that exist; level_mask
// bit l = send level l; rep (optional): private copies of levels < rep_levels, copy chosen by CTA.
__device__ __forceinline__ void scatter_tile(const GridMeta& gm, float2* __restrict__ g_table, const TcSmem& s, const int buf,
                                             const int warp, const int lane, const int n_valid, const int level_mask,
                                             float2* __restrict__ rep, const int rep_levels, const int rep_entries,
                                             const int n_rep) {
      const int xp = lane & 1, pr = lane >> 1;
      const int sg = pr >> 2, lq = pr & 3;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int lvl = 4 * lq + q;
        if (!((level_mask >> lvl) & 1)) continue;  // diagnostics: SDB_FB_LEVELS masks levels out
        const float sc = gm.scale[lvl];
        const uint32_t res = gm.res[lvl], size = gm.size[lvl], hashed = gm.hashed[lvl];
        float2* tl = (rep && lvl < rep_levels ? rep + (size_t)(blockIdx.x % n_rep) * rep_entries : g_table) + gm.offset[lvl];
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
          if (sl >= n_valid || (gxy.x == 0.f && gxy.y == 0.f)) continue;
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

}  // namespace fbtc
