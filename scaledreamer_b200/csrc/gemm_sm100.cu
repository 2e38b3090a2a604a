// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
// One CTA computes a 128 x BN tile of D = A * B^T with fp16 operands and fp32 accumulation in tensor memory:
//   warp 0   : TMA producer   (cp.async.bulk.tensor into a ring of 128-byte-swizzled K-major smem tiles)
//   warp 1   : MMA issuer     (one elected thread issues tcgen05.mma 128xBNx16, commits stages back to the ring)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue       (tcgen05.ld accumulator rows -> bias / timestep-embedding / activation / residual
//                              -> fp16 or fp32 global stores)
// The A operand is addressed through the tensor map in one of three ways (dense.h::OperandMode): a plain
// (batched) matrix, an NHWC activation read as shifted 4-D boxes (the 9 taps of a 3x3 convolution, zero
// padding supplied by TMA out-of-bounds fill), or a per-head view of a [B, L, heads*64] tensor.
//
// Replaces on the reference path: cuDNN convolutions and cuBLAS GEMMs reached through
// extern/mvdream/ldm/modules/diffusionmodules/openaimodel.py:255-275 (ResBlock), ldm/modules/attention.py:163-194
// (CrossAttention), ldm/modules/diffusionmodules/model.py:129-203 (VAE ResnetBlock / AttnBlock) and their
// diffusers equivalents (stable_diffusion_asd_guidance.py:170-178, 318-331).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dense.h"
#include "ptx_sm100.cuh"

namespace dense {
namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;               // 64 fp16 = 128 bytes = one swizzle row
constexpr int kATileBytes = kBM * kBK * 2;

// DEEP: at most one CTA per SM is in flight (<= 148 tiles), so the ring takes the whole shared memory instead.
template <int BN, bool DEEP = false>
struct Cfg {
  static constexpr int kBTileBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  // two CTAs share an SM (one CTA's epilogue overlaps the other's main loop): <= ~110 KB of ring per CTA
  static constexpr int kStages = DEEP ? (BN <= 80 ? 8 : (BN <= 160 ? 6 : 4)) : (BN <= 80 ? 4 : 3);
  // DEEP: two accumulators in tensor memory (the MMA warp fills one while the epilogue drains the other) and two sets
  // of four epilogue warps that take alternate 32-column chunks of a tile.
  // (round 2) also with two CTAs per SM when both CTAs' accumulator pairs fit the 512 tensor-memory columns (BN <= 128):
  // short-K layers spend longer in the epilogue than in the main loop, and a single accumulator serialises the two
  static constexpr int kAccBufs = (DEEP || BN <= 128) ? 2 : 1;
  static constexpr int kEpiSets = DEEP ? 2 : 1;
  static constexpr int kThreads = 128 + 128 * kEpiSets;
  static constexpr int kAccStride = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
  static constexpr int kTmemCols = kAccStride * kAccBufs;
  // per epilogue warp: the tile's BN bias values (fp16), fetched once per tile before the accumulator is ready instead
  // of once per 32-column chunk with the load latency exposed (ncu source view, round 2: the bias conversion was the
  // hottest stall of the epilogue warps)
  static constexpr int kBiasBytes = (BN == 64 || BN == 128 || BN == 256) ? BN * 2 * 4 * kEpiSets : 0;
  // per epilogue warp: two 32-row x 32-byte staging tiles for the TMA stores of the fp16 results (one 32-byte sector
  // per row and store instead of a 16-byte piece per thread: the per-thread stores cost one L2 request per piece and
  // bounded short-K layers at 14 GB/s per SM)
  static constexpr int kStoreBytes = BN == 80 ? 0 : 2048 * 4 * kEpiSets;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 1024 /*barriers + pad*/ + kBiasBytes + kStoreBytes;
};

__device__ __forceinline__ void load_operand(const CUtensorMap* tm, const OperandGeom& g, void* smem, uint64_t* bar,
                                             int kb, int r0, int z) {
  if (g.mode == kMatrix) {
    ptx::tma_load_3d(smem, tm, bar, kb * kBK, r0, g.batched ? z : 0);
  } else if (g.mode == kHeads) {
    if (g.mn_major)  // tile = [64 K-rows (sequence)][64 MN (head dim)]
      ptx::tma_load_4d(smem, tm, bar, r0, z % g.heads, kb * kBK, z / g.heads);
    else
      ptx::tma_load_4d(smem, tm, bar, kb * kBK, z % g.heads, r0, z / g.heads);
  } else {  // kConv3x3
    const int tap = kb / g.cin_blocks, cb = kb - tap * g.cin_blocks;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int hw = g.H * g.W;
    const int n0 = r0 / hw, rem = r0 - n0 * hw;
    const int h0 = rem / g.W, w0 = rem - h0 * g.W;
    // two 64-pixel boxes (8 KB each) instead of one of 128: the copy engine works on them concurrently, which is what
    // keeps a lone CTA per SM fed (measured: 425 -> see profiles/r2i_conv_split_timeline.log ns per k-block)
    const int hw_ = g.split_dim == 1 ? g.bw / 2 : 0, hh_ = g.split_dim == 2 ? g.bh / 2 : 0, hn_ = g.split_dim == 3 ? g.bn / 2 : 0;
    ptx::tma_load_4d(smem, tm, bar, cb * kBK, w0 + dx, h0 + dy, n0);
    ptx::tma_load_4d(static_cast<uint8_t*>(smem) + kATileBytes / 2, tm, bar, cb * kBK, w0 + dx + hw_, h0 + dy + hh_, n0 + hn_);
  }
}

// The common epilogue (no split-K partials, no GEGLU pairing, no row softmax) in 16-column steps with the tensor-memory
// read of the NEXT step in flight while the current one is converted and stored: the ncu source view of round 2 put the
// wait on tcgen05.ld first among the epilogue warps' stalls, and short-K layers are bounded by these four warps, not
// by the main loop. v = act(alpha * acc + bias + rowbias) + residual, one output row per thread.
template <int BN>
__device__ __forceinline__ void epilogue_tile_pipelined(const GemmParams& p, uint32_t tm_row, int lane, int m, int n0,
                                                        long long zoff, int set, int nsets, const __half* sbias,
                                                        const CUtensorMap* tmC, uint8_t* stage) {
  const bool m_ok = m < p.M;
  const bool tma = stage != nullptr;  // launch-uniform: fp16 output through shared memory + TMA stores
  const int m_row0 = m - lane;        // first row of this warp's 32
  int nstore = 0;                     // stores issued by lane 0 so far (staging tile = nstore & 1)
  const float* rowb = p.rowbias ? p.rowbias + (long long)(m_ok ? m / p.rows_per_group : 0) * p.rowbias_ld : nullptr;
  const bool res_vec = p.residual && (p.ldr & 7) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0;
  const bool bias_vec = p.bias && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
  const bool rowb_vec = rowb && (reinterpret_cast<uintptr_t>(rowb) & 15) == 0 && (p.rowbias_ld & 3) == 0;
  const __half* rrow = p.residual ? p.residual + (long long)m * p.ldr : nullptr;
  const int cstep = 16 * nsets;
  auto load_res = [&](int c, uint4(&rv)[2]) {
    const bool on = res_vec && m_ok && n0 + c + 16 <= p.N;
#pragma unroll
    for (int q = 0; q < 2; ++q)
      rv[q] = on ? __ldg(reinterpret_cast<const uint4*>(rrow + n0 + c) + q) : make_uint4(0, 0, 0, 0);
  };
  auto finish = [&](const uint32_t(&v)[16], const uint4(&res)[2], int c) {
    const int nb = n0 + c;
    if (nb >= p.N) return;         // warp-uniform
    if (!tma && !m_ok) return;     // rows past M: nothing to store (the TMA path clips them and stays warp-collective)
    const bool full = nb + 16 <= p.N;
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = 0.f;
    if (p.bias) {
      if (sbias || (bias_vec && full)) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint4 bv = sbias ? *(reinterpret_cast<const uint4*>(sbias + c) + q)
                                 : __ldg(reinterpret_cast<const uint4*>(p.bias + nb) + q);
          const __half2* h2 = reinterpret_cast<const __half2*>(&bv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 t = __half22float2(h2[e]);
            f[q * 8 + 2 * e] = t.x;
            f[q * 8 + 2 * e + 1] = t.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (nb + j < p.N) f[j] = __half2float(__ldg(p.bias + nb + j));
      }
    }
    if (rowb) {
      if (rowb_vec && full) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(rowb + nb) + q);
          f[4 * q] += t.x, f[4 * q + 1] += t.y, f[4 * q + 2] += t.z, f[4 * q + 3] += t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (nb + j < p.N) f[j] += __ldg(rowb + nb + j);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaf(__uint_as_float(v[j]), p.alpha, f[j]);
    if (p.act == kActSilu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = f[j] / (1.f + __expf(-f[j]));
    } else if (p.act == kActGelu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = 0.5f * f[j] * (1.f + erff(f[j] * 0.70710678118654752f));
    }
    if (p.residual) {
      if (res_vec && full) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const __half2* h2 = reinterpret_cast<const __half2*>(&res[q]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 t = __half22float2(h2[e]);
            f[q * 8 + 2 * e] += t.x;
            f[q * 8 + 2 * e + 1] += t.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (nb + j < p.N) f[j] += __half2float(__ldg(rrow + nb + j));
      }
    }
    if (p.out_fp32) {
      float* o = reinterpret_cast<float*>(p.out) + zoff + (long long)m * p.ldc + nb;
      if (full && (p.ldc & 3) == 0 && (zoff & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<float4*>(o)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (nb + j < p.N) o[j] = f[j];
      }
    } else if (tma) {
      // 32 rows x 32 bytes, 16-byte halves swapped on rows with bit 2 set (the 32-byte swizzle of the tensor map)
      uint8_t* tile = stage + (nstore & 1) * 1024;
      if (nstore >= 2) {  // the store that last read this tile must have drained it
        if (lane == 0) ptx::tma_store_wait_read<1>();
        __syncwarp();
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint4 ov;
        __half2* h2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(f[q * 8 + 2 * e], f[q * 8 + 2 * e + 1]);
        *reinterpret_cast<uint4*>(tile + lane * 32 + ((q ^ ((lane >> 2) & 1)) << 4)) = ov;
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d(tmC, tile, nb, m_row0);
        ptx::tma_store_commit();
      }
      ++nstore;
    } else {
      __half* o = reinterpret_cast<__half*>(p.out) + zoff + (long long)m * p.ldc + nb;
      if (full && (p.ldc & 7) == 0 && (zoff & 7) == 0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint4 ov;
          __half2* h2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(f[q * 8 + 2 * e], f[q * 8 + 2 * e + 1]);
          reinterpret_cast<uint4*>(o)[q] = ov;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (nb + j < p.N) o[j] = __float2half_rn(f[j]);
      }
    }
  };
  uint32_t va[16], vb[16];
  uint4 ra[2], rb[2];
  int c = 16 * set;
  if (c >= BN) return;
  ptx::tmem_ld_32x16(tm_row + (uint32_t)c, va);
  load_res(c, ra);
#pragma unroll 1
  for (; c < BN; c += 2 * cstep) {
    ptx::tmem_ld_wait();  // va
    const bool more1 = c + cstep < BN;
    if (more1) {
      ptx::tmem_ld_32x16(tm_row + (uint32_t)(c + cstep), vb);
      load_res(c + cstep, rb);
    }
    finish(va, ra, c);
    if (!more1) break;
    ptx::tmem_ld_wait();  // vb
    if (c + 2 * cstep < BN) {
      ptx::tmem_ld_32x16(tm_row + (uint32_t)(c + 2 * cstep), va);
      load_res(c + 2 * cstep, ra);
    }
    finish(vb, rb, c + cstep);
  }
  if (tma && nstore > 0) {  // the next tile starts with both staging tiles free
    if (lane == 0) ptx::tma_store_wait_read<0>();
    __syncwarp();
  }
}

// Epilogue of one 128 x BN tile for the 32 rows of TMEM lane quadrant wq (one row per thread).
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tmem_base, int wq, int lane, int m0, int n0,
                                              int z, int zsplit, int set, int nsets, const __half* sbias,
                                              const CUtensorMap* tmC, uint8_t* stage) {
  const int m = m0 + wq * 32 + lane;
  const bool m_ok = m < p.M;
  const long long zoff = (long long)(z / p.out_zdiv) * p.out_zs_hi + (long long)(z % p.out_zdiv) * p.out_zs_lo;
  const float* rowb = p.rowbias ? p.rowbias + (long long)(m_ok ? m / p.rows_per_group : 0) * p.rowbias_ld : nullptr;
  if constexpr (BN == 80) {
    if (p.row_softmax && set != 0) return;  // the whole row belongs to one thread: the first warp set does it
    if (p.row_softmax) {
      // whole score row in this tile (N <= 80): softmax(alpha * acc) in registers, padding columns zeroed
      uint32_t v[96];
      uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
      uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
      uint32_t(&v2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[64]);
      const uint32_t tb = tmem_base + ((uint32_t)(wq * 32) << 16);
      ptx::tmem_ld_32x32(tb, v0);
      ptx::tmem_ld_32x32(tb + 32u, v1);
      ptx::tmem_ld_32x32(tb + 64u, v2);  // columns 80..95 are unallocated-by-MMA garbage and are masked below
      ptx::tmem_ld_wait();
      if (m_ok) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 80; ++j)
          if (j < p.N) mx = fmaxf(mx, __uint_as_float(v[j]) * p.alpha);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 80; ++j) {
          const float e = j < p.N ? __expf(__uint_as_float(v[j]) * p.alpha - mx) : 0.f;
          v[j] = __float_as_uint(e);
          sum += e;
        }
        const float inv = 1.f / sum;
        __half* o = reinterpret_cast<__half*>(p.out) + zoff + (long long)m * p.ldc;
#pragma unroll
        for (int q = 0; q < 10; ++q) {
          if (q * 8 >= p.ldc) break;
          uint4 ov;
          __half2* h2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            h2[e] = __floats2half2_rn(__uint_as_float(v[q * 8 + 2 * e]) * inv, __uint_as_float(v[q * 8 + 2 * e + 1]) * inv);
          reinterpret_cast<uint4*>(o)[q] = ov;
        }
      }
      return;
    }
  }
  if (p.splits == 1 && p.act != kActGeglu) {
    uint32_t tm = tmem_base + ((uint32_t)(wq * 32) << 16);
    asm volatile("" : "+r"(tm));  // keep it in a register: the compiler otherwise re-derives it from %tid every step
    epilogue_tile_pipelined<BN>(p, tm, lane, m, n0, zoff, set, nsets, sbias, tmC,
                                (p.tma_store && !p.out_fp32) ? stage : nullptr);
    return;
  }
  // Left for the 32-column loop: split-K partial planes (raw fp32 accumulators, the epilogue runs in
  // splitk_finalize_kernel) and GEGLU projections (bias only; a chunk holds 16 values followed by their 16 gates).
  const bool bias_vec = p.bias && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
  uint32_t tm_row = tmem_base + ((uint32_t)(wq * 32) << 16);
  asm volatile("" : "+r"(tm_row));
#pragma unroll 1
  for (int c = 32 * set; c < BN; c += 32 * nsets) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(tm_row + (uint32_t)c, v);
    ptx::tmem_ld_wait();
    const int nb = n0 + c;
    if (!m_ok || nb >= p.N) continue;
    const bool full_chunk = nb + 32 <= p.N;
    if (p.splits > 1) {
      float* o = p.ws + ((long long)zsplit * p.M + m) * p.N + nb;
      if (full_chunk && (p.N & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          reinterpret_cast<uint4*>(o)[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nb + j < p.N) o[j] = __uint_as_float(v[j]);
      }
      continue;
    }
    // GEGLU: value * gelu(gate) -> 16 outputs
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = 0.f;
    if (p.bias) {
      if (sbias || (bias_vec && full_chunk)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 bv = sbias ? *(reinterpret_cast<const uint4*>(sbias + c) + q)
                                 : __ldg(reinterpret_cast<const uint4*>(p.bias + nb) + q);
          const __half2* h2 = reinterpret_cast<const __half2*>(&bv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 t = __half22float2(h2[e]);
            f[q * 8 + 2 * e] = t.x;
            f[q * 8 + 2 * e + 1] = t.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nb + j < p.N) f[j] = __half2float(__ldg(p.bias + nb + j));
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaf(__uint_as_float(v[j]), p.alpha, f[j]);
    __half* o = reinterpret_cast<__half*>(p.out) + zoff + (long long)m * p.ldc + (nb >> 1);
    uint4 ov[2];
    __half2* h2 = reinterpret_cast<__half2*>(ov);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float g0 = f[16 + 2 * e], g1 = f[17 + 2 * e];
      h2[e] = __floats2half2_rn(f[2 * e] * (0.5f * g0 * (1.f + erff(g0 * 0.70710678118654752f))),
                                f[2 * e + 1] * (0.5f * g1 * (1.f + erff(g1 * 0.70710678118654752f))));
    }
    reinterpret_cast<uint4*>(o)[0] = ov[0];
    reinterpret_cast<uint4*>(o)[1] = ov[1];
  }
}

// Persistent: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (m fastest, then n, then z / K split).
// The TMA ring keeps running across tile boundaries, so the next tile's operands stream in while the epilogue
// warps drain the accumulator; two CTAs share an SM, so one CTA's epilogue also overlaps the other's main loop.
template <int BN, bool DEEP>
__global__ void __launch_bounds__(Cfg<BN, DEEP>::kThreads, DEEP ? 1 : 2)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ GemmParams p) {
  using C = Cfg<BN, DEEP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(base + C::kStages * C::kStageBytes);
  uint64_t* empty = full + C::kStages;
  uint64_t* accum_full = empty + C::kStages;        // [kAccBufs]
  uint64_t* accum_empty = accum_full + C::kAccBufs;  // [kAccBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_empty + C::kAccBufs);
  uint8_t* store_smem = base + C::kStages * C::kStageBytes + 1024;  // [epilogue warp][2][1 KB] (kStoreBytes), 1 KB aligned
  uint8_t* bias_smem = store_smem + C::kStoreBytes;                 // [epilogue warp][BN] fp16 (kBiasBytes)
  (void)bias_smem;
  (void)store_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto stamp = [&](int slot) {
    if (p.dbg && threadIdx.x == 128) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[(size_t)blockIdx.x * 4 + slot] = t;
    }
  };
  stamp(0);
  pdl_launch_dependents();  // the next kernel may start its own set-up; its data dependency is its pdl_wait()

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (p.tma_store) ptx::prefetch_tmap(&tmC);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < C::kAccBufs; ++b) {
      ptx::mbar_init(&accum_full[b], 1);
      ptx::mbar_init(&accum_empty[b], 128 * C::kEpiSets);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
    ptx::tmem_relinquish();
  }
  pdl_wait();  // everything above touched only shared / tensor memory and the kernel's own parameters
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  stamp(1);
  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int total = tiles_mn * p.tiles_z;
  const int per_split = (p.num_k_blocks + p.splits - 1) / p.splits;

  // tile -> (m0, n0, batch index z, split index zs, K-block range)
#define SDB_DECODE_TILE(tile)                                                        \
  const int t_mn = (tile) % tiles_mn, t_z = (tile) / tiles_mn;                       \
  const int m0 = (t_mn % p.tiles_m) * kBM, n0 = (t_mn / p.tiles_m) * BN;             \
  const int z = p.splits > 1 ? 0 : t_z, zs = p.splits > 1 ? t_z : 0;                 \
  const int kb0 = zs * per_split;                                                    \
  const int nkb = p.splits > 1 ? min(per_split, p.num_k_blocks - kb0) : p.num_k_blocks;

  if (warp == 0) {
    if (ptx::elect_one()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        SDB_DECODE_TILE(tile)
        (void)zs;
        if (p.a.mode == kConv3x3 && p.b.mode == kMatrix) {
          // implicit 3x3 convolution: everything that depends on the tile only (image / row / column of its first pixel)
          // is worked out once, the (tap, channel block) walk is two counters -- the producer is ONE thread, and four
          // integer divisions per k-block were enough to make it the slowest stage of a single-CTA-per-SM launch
          const OperandGeom& g = p.a;
          const int hw = g.H * g.W;
          const int img = m0 / hw, rem = m0 - img * hw;
          const int h0 = rem / g.W, w0 = rem - h0 * g.W;
          const int hw_ = g.split_dim == 1 ? g.bw / 2 : 0, hh_ = g.split_dim == 2 ? g.bh / 2 : 0,
                    hn_ = g.split_dim == 3 ? g.bn / 2 : 0;
          int tap = kb0 / g.cin_blocks, cb = kb0 - tap * g.cin_blocks;
          int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
          int s = it % C::kStages;
          uint32_t ph = (it / C::kStages) & 1;
          for (int kb = 0; kb < nkb; ++kb) {
            ptx::mbar_wait(&empty[s], ph ^ 1u);
            ptx::mbar_arrive_expect_tx(&full[s], C::kStageBytes);
            uint8_t* sa = base + s * C::kStageBytes;
            ptx::tma_load_4d(sa, &tmA, &full[s], cb * kBK, w0 + dx, h0 + dy, img);
            ptx::tma_load_4d(sa + kATileBytes / 2, &tmA, &full[s], cb * kBK, w0 + dx + hw_, h0 + dy + hh_, img + hn_);
            ptx::tma_load_3d(sa + kATileBytes, &tmB, &full[s], (kb0 + kb) * kBK, n0, 0);
            if (++cb == g.cin_blocks) {
              cb = 0;
              if (++dx == 2) dx = -1, ++dy;
            }
            if (++s == C::kStages) s = 0, ph ^= 1u;
          }
          it += nkb;
          continue;
        }
        if (p.a.mode == kMatrix && p.b.mode == kMatrix) {  // plain (batched) matrices: no per-k-block decisions at all
          const int za = p.a.batched ? z : 0, zb = p.b.batched ? z : 0;
          int s = it % C::kStages;
          uint32_t ph = (it / C::kStages) & 1;
          for (int kb = 0; kb < nkb; ++kb) {
            ptx::mbar_wait(&empty[s], ph ^ 1u);
            ptx::mbar_arrive_expect_tx(&full[s], C::kStageBytes);
            uint8_t* sa = base + s * C::kStageBytes;
            ptx::tma_load_3d(sa, &tmA, &full[s], (kb0 + kb) * kBK, m0, za);
            ptx::tma_load_3d(sa + kATileBytes, &tmB, &full[s], (kb0 + kb) * kBK, n0, zb);
            if (++s == C::kStages) s = 0, ph ^= 1u;
          }
          it += nkb;
          continue;
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          ptx::mbar_wait(&empty[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&full[s], C::kStageBytes);
          uint8_t* sa = base + s * C::kStageBytes;
          load_operand(&tmA, p.a, sa, &full[s], kb0 + kb, m0, z);
          load_operand(&tmB, p.b, sa + kATileBytes, &full[s], kb0 + kb, n0, z);
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_f16(kBM, BN, 0, 0, p.b.mn_major);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
        SDB_DECODE_TILE(tile)
        (void)m0; (void)n0; (void)z; (void)zs; (void)kb0;
        const uint32_t buf = tcount % C::kAccBufs, use = tcount / C::kAccBufs;
        const uint32_t tmem_acc = tmem_base + buf * C::kAccStride;
        ptx::mbar_wait(&accum_empty[buf], (use & 1) ^ 1u);  // the epilogue warps drained this accumulator
        ptx::tc_fence_after();
        int s = it % C::kStages;
        uint32_t ph = (it / C::kStages) & 1;
        for (int kb = 0; kb < nkb; ++kb, ++it, s = (s + 1 == C::kStages ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(base + s * C::kStageBytes);
          const uint32_t sb = sa + kATileBytes;
          const uint64_t da = ptx::smem_desc_k_sw128(sa);
          const uint64_t db = p.b.mn_major ? ptx::smem_desc_mn_sw128(sb, 8192) : ptx::smem_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // K-major: 16 elements = 32 bytes further along the swizzled row; MN-major: 16 K-rows = 2048 bytes
            const uint64_t a_adv = (uint64_t)((k * 32) >> 4);
            const uint64_t b_adv = p.b.mn_major ? (uint64_t)((k * 2048) >> 4) : (uint64_t)((k * 32) >> 4);
            ptx::umma_f16(tmem_acc, da + a_adv, db + b_adv, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty[s]);
        }
        ptx::umma_commit(&accum_full[buf]);
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3, set = (warp - 4) >> 2;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
      SDB_DECODE_TILE(tile)
      (void)kb0; (void)nkb;
      const uint32_t buf = tcount % C::kAccBufs, use = tcount / C::kAccBufs;
      const __half* sbias = nullptr;
      if constexpr (C::kBiasBytes > 0) {
        if (p.bias && p.splits == 1 && !p.row_softmax && n0 + BN <= p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
          __half* mine = reinterpret_cast<__half*>(bias_smem) + (warp - 4) * BN;
          __syncwarp();  // the previous tile's readers are done
          constexpr int kPer = BN / 32;  // halfs per lane: 2, 4 or 8
          if constexpr (kPer == 2)
            *reinterpret_cast<uint32_t*>(mine + 2 * lane) = __ldg(reinterpret_cast<const uint32_t*>(p.bias + n0) + lane);
          else if constexpr (kPer == 4)
            *reinterpret_cast<uint2*>(mine + 4 * lane) = __ldg(reinterpret_cast<const uint2*>(p.bias + n0) + lane);
          else
            *reinterpret_cast<uint4*>(mine + 8 * lane) = __ldg(reinterpret_cast<const uint4*>(p.bias + n0) + lane);
          __syncwarp();
          sbias = mine;
        }
      }
      ptx::mbar_wait(&accum_full[buf], use & 1);
      ptx::tc_fence_after();
      if (tcount == 0) stamp(2);
      epilogue_tile<BN>(p, tmem_base + buf * C::kAccStride, wq, lane, m0, n0, z, zs, set, C::kEpiSets, sbias, &tmC,
                        C::kStoreBytes > 0 ? store_smem + (warp - 4) * 2048 : nullptr);
      ptx::tc_fence_before();
      ptx::mbar_arrive(&accum_empty[buf]);
    }
  }
#undef SDB_DECODE_TILE
  stamp(3);
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

__global__ void splitk_finalize_kernel(const GemmParams p);

// Optional per-launch timing (bench.py roofline): CUDA events on the launching stream around every GEMM launch.
struct ProfRec {
  cudaEvent_t a, b;
  double flops;
  int M, N, K, z, splits, bn, mode;
};
char g_prof_dump[512] = "";
bool g_prof = false;
unsigned long long* g_dbg = nullptr;
std::vector<ProfRec> g_recs;

template <int BN, bool DEEP>
int launch_v(const GemmPlan& plan, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel<BN, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<BN, DEEP>::kSmemBytes);
    if (e != cudaSuccess) {
      sdb_set_error("gemm: smem attribute: %s", cudaGetErrorString(e));
      return SDB_ERR_CUDA;
    }
    attr = true;
  }
  ProfRec rec;
  if (g_prof) {
    cudaEventCreate(&rec.a);
    cudaEventCreate(&rec.b);
    rec.flops = 2.0 * plan.p.M * plan.p.N * plan.p.K * (plan.p.splits > 1 ? 1 : plan.grid.z);
    rec.M = plan.p.M, rec.N = plan.p.N, rec.K = plan.p.K, rec.z = plan.p.splits > 1 ? 1 : (int)plan.grid.z;
    rec.splits = plan.p.splits, rec.bn = BN, rec.mode = plan.p.a.mode;
    cudaEventRecord(rec.a, stream);
  }
  GemmParams prm = plan.p;
  prm.dbg = g_dbg;
  prm.tiles_m = (int)plan.grid.x, prm.tiles_n = (int)plan.grid.y, prm.tiles_z = (int)plan.grid.z;
  const long long total_tiles = (long long)plan.grid.x * plan.grid.y * plan.grid.z;
  const int ctas = (int)std::min<long long>(total_tiles, DEEP ? (long long)kNumSMs : 2LL * kNumSMs);
  sdb_launch(gemm_f16_kernel<BN, DEEP>, ctas, Cfg<BN, DEEP>::kThreads, Cfg<BN, DEEP>::kSmemBytes, stream, plan.ta, plan.tb, plan.tc, prm);
  if (plan.p.splits > 1) {
    const long long total = (long long)plan.p.M * ((plan.p.N + 3) / 4);
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 8);
    sdb_launch(splitk_finalize_kernel, grid, 256, 0, stream, prm);
    SDB_COUNT_LAUNCH();
  }
  if (g_prof) {
    cudaEventRecord(rec.b, stream);
    g_recs.push_back(rec);
  }
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("gemm_f16");
  return SDB_OK;
}

template <int BN>
int launch(const GemmPlan& plan, cudaStream_t stream) {
  // SDB_GEMM_DEEP = all | none | auto (default) selects the one-CTA-per-SM variant (diagnostics)
  static const int policy = [] {
    const char* e = getenv("SDB_GEMM_DEEP");
    return !e ? 2 : (!strcmp(e, "all") ? 1 : (!strcmp(e, "none") ? 0 : 2));
  }();
  const long long total_tiles = (long long)plan.grid.x * plan.grid.y * plan.grid.z;
  const bool deep = policy == 1 || (policy == 2 && total_tiles <= kNumSMs);
  return deep ? launch_v<BN, true>(plan, stream) : launch_v<BN, false>(plan, stream);
}

}  // namespace

// 128 x 256 tiles move 48 KB per K-block for 4.2 MFLOP (87 FLOP per operand byte, against 71 for 128 x 160 and 64 for
// 128 x 128): the main loop is bound by L2 -> shared-memory delivery, so wide-N layers with a long K run on the wider
// tile (one CTA per SM, two accumulators = all 512 tensor-memory columns).
bool wide_tile(int M, int N, int K) {
  static const bool on = !(getenv("SDB_GEMM_BN256") && atoi(getenv("SDB_GEMM_BN256")) == 0);
  // only when the wider tiles still fill every SM: with fewer the extra CTAs of the narrow tiling win
  // (measured: 1280 x 1280 x 1280, 50 wide tiles, 0.54 ms vs 0.45 ms; 65536 x 256 x 2304, 512 tiles, 0.46 vs 0.53 ms)
  static const int min_k = getenv("SDB_GEMM_WIDE_MINK") ? atoi(getenv("SDB_GEMM_WIDE_MINK")) : 256;
  return on && N % 256 == 0 && K >= min_k && (long long)((M + kBM - 1) / kBM) * (N / 256) >= kNumSMs;
}

int gemm_splits(int M, int N, int K) {
  const bool wide = wide_tile(M, N, K);
  const int bn = wide ? 256 : (N <= 64 ? 64 : 128);
  const int slots = wide ? kNumSMs : 2 * kNumSMs;  // the wide tile runs one CTA per SM
  const int tiles = ((M + kBM - 1) / kBM) * ((N + bn - 1) / bn);
  const int nkb = (K + kBK - 1) / kBK;
  // A lone CTA per SM streams a k-block in ~200 ns since the producer thread stopped dividing (round 2), so a launch of
  // 80-128 tiles is better left whole (1280 x 1280 x 11520: 45 us whole, 59 us as three splits + finalize); only
  // launches that would leave three quarters of the SMs idle are split along K.
  if (tiles * 4 > kNumSMs || nkb < 48) return 1;
  int sp = std::min(std::min(slots / tiles, nkb / 4), 16);
  if (sp <= 1) return 1;
  const int per = (nkb + sp - 1) / sp;
  return (nkb + per - 1) / per;  // every split owns at least one K block
}

namespace {

int pick_bn(int M, int N, int K) {
  if (const char* e = getenv("SDB_GEMM_BN")) {  // diagnostics: force a tile width
    const int bn = atoi(e);
    if (bn == 64 || bn == 80 || bn == 128 || bn == 160) return bn;
  }
  if (wide_tile(M, N, K)) return 256;
  // 160-wide tiles divide the UNet's channel counts exactly, but only 128-wide ones leave room for two accumulators per
  // CTA with two CTAs per SM (the epilogue of one tile then overlaps the main loop of the next): measured faster on
  // every UNet / VAE shape despite the padded last tile of N = 320 / 960 (profiles/r2l_gemm_shapes_acc2_bn128.log).
  // SDB_GEMM_BN=160 still selects the old tiling.
  if (N <= 64) return 64;
  return 128;
}

void fill_epilogue(GemmParams& p, const Epilogue& ep) {
  p.out = ep.out;
  p.out_fp32 = ep.out_fp32;
  p.ldc = ep.ldc;
  p.out_zdiv = 1;
  p.out_zs_hi = 0;
  p.out_zs_lo = 0;
  p.bias = ep.bias;
  p.rowbias = ep.rowbias;
  p.rows_per_group = ep.rows_per_group > 0 ? ep.rows_per_group : 1;
  p.rowbias_ld = ep.rowbias_ld > 0 ? ep.rowbias_ld : p.N;
  p.residual = ep.residual;
  p.ldr = ep.ldr;
  p.alpha = ep.alpha;
  p.act = ep.act;
  p.splits = 1;
  p.dbg = nullptr;
  p.ws = nullptr;
  p.row_softmax = 0;
}

// thread per 4 output columns: sums the split planes in fixed order (bitwise reproducible), then the GEMM epilogue
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const GemmParams p) {
  pdl_prologue();
  const int n4 = (p.N + 3) / 4;
  const long long total = (long long)p.M * n4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / n4), nb = (int)(i % n4) * 4;
    float f[4] = {0.f, 0.f, 0.f, 0.f};
    const long long plane = (long long)p.M * p.N;
    const float* w0 = p.ws + (long long)m * p.N + nb;
    if (nb + 4 <= p.N && (p.N & 3) == 0) {
      // four planes in flight (the loop is latency-bound otherwise); the additions keep the plane order
      int s = 0;
      for (; s + 4 <= p.splits; s += 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(w0 + (s + u) * plane);
#pragma unroll
        for (int u = 0; u < 4; ++u) f[0] += v[u].x, f[1] += v[u].y, f[2] += v[u].z, f[3] += v[u].w;
      }
      for (; s < p.splits; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(w0 + s * plane);
        f[0] += v.x, f[1] += v.y, f[2] += v.z, f[3] += v.w;
      }
    } else {
      for (int s = 0; s < p.splits; ++s) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < p.N) f[j] += w0[s * plane + j];
      }
    }
    const float* rowb = p.rowbias ? p.rowbias + (long long)(m / p.rows_per_group) * p.rowbias_ld : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (nb + j >= p.N) continue;
      float v = f[j] * p.alpha;
      if (p.bias) v += __half2float(p.bias[nb + j]);
      if (rowb) v += rowb[nb + j];
      if (p.act == kActSilu) v = v / (1.f + __expf(-v));
      else if (p.act == kActGelu) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
      if (p.residual) v += __half2float(p.residual[(long long)m * p.ldr + nb + j]);
      if (p.out_fp32) reinterpret_cast<float*>(p.out)[(long long)m * p.ldc + nb + j] = v;
      else reinterpret_cast<__half*>(p.out)[(long long)m * p.ldc + nb + j] = __float2half_rn(v);
    }
  }
}

// Output tensor map for the TMA-store epilogue: plain [M, N] fp16 with 16-byte-aligned rows, one launch plane.
int plan_output_map(GemmPlan* plan) {
  static const bool on = !(getenv("SDB_GEMM_TMA_STORE") && atoi(getenv("SDB_GEMM_TMA_STORE")) == 0);
  GemmParams& p = plan->p;
  p.tma_store = 0;
  memset(&plan->tc, 0, sizeof(plan->tc));
  if (!on || p.out_fp32 || p.splits > 1 || p.row_softmax || p.act == kActGeglu || plan->grid.z != 1 || plan->bn == 80 ||
      (p.ldc & 7) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0 || p.out_zs_hi != 0 || p.out_zs_lo != 0)
    return SDB_OK;
  uint64_t dc[2] = {(uint64_t)p.N, (uint64_t)p.M};
  uint64_t sc[1] = {(uint64_t)p.ldc * 2};
  uint32_t bc[2] = {16, 32};
  int rc = make_tmap(&plan->tc, p.out, 2, dc, sc, bc, 32);
  if (rc) return rc;
  p.tma_store = 1;
  return SDB_OK;
}

void apply_splitk(GemmPlan* plan, const Epilogue& ep) {
  if (!ep.splitk_ws || plan->grid.z != 1) return;
  const int sp = gemm_splits(plan->p.M, plan->p.N, plan->p.K);
  if (sp <= 1) return;
  plan->p.splits = sp;
  plan->p.ws = ep.splitk_ws;
  plan->grid.z = sp;
}

}  // namespace

int make_tmap(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int swizzle_bytes, int dtype) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    sdb_set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)");
    return SDB_ERR_CUDA;
  }
  cuuint64_t d[5];
  cuuint64_t st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
    if (i > 0) st[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) {
    sdb_set_error("tensor map: base pointer must be 16-byte aligned");
    return SDB_ERR_ARG;
  }
  for (int i = 0; i + 1 < rank; ++i)
    if (st[i] % 16 != 0) {
      sdb_set_error("tensor map: stride %d (%llu bytes) must be a multiple of 16", i, (unsigned long long)st[i]);
      return SDB_ERR_ARG;
    }
  CUresult r = enc(out, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), d, st, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                   : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                   : swizzle_bytes == 12832 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B  /* 128-byte swizzle, 32-byte atoms */
                                         : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sdb_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r,
                  rank, (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)(rank > 2 ? d[2] : 0),
                  b[0], b[1], rank > 2 ? b[2] : 0);
    return SDB_ERR_CUDA;
  }
  return SDB_OK;
}

int plan_gemm(GemmPlan* plan, const __half* A, long long lda, const __half* B, long long ldb, int M, int N, int K,
              const Epilogue& ep, int batch, long long a_zs, long long b_zs, long long out_zs) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0 || (lda & 7) || (ldb & 7) || (a_zs & 7) || (b_zs & 7)) {
    sdb_set_error("gemm: bad shape M=%d N=%d K=%d batch=%d lda=%lld ldb=%lld (leading dims must be multiples of 8)", M,
                  N, K, batch, lda, ldb);
    return SDB_ERR_ARG;
  }
  memset(plan, 0, sizeof(*plan));
  plan->bn = pick_bn(M, N, K);
  GemmParams& p = plan->p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_k_blocks = (K + kBK - 1) / kBK;
  p.a.mode = kMatrix;
  p.a.batched = batch > 1 && a_zs != 0;
  p.b.mode = kMatrix;
  p.b.batched = batch > 1 && b_zs != 0;
  fill_epilogue(p, ep);
  p.out_zs_hi = out_zs;
  uint64_t da[3] = {(uint64_t)K, (uint64_t)M, (uint64_t)(p.a.batched ? batch : 1)};
  uint64_t sa[2] = {(uint64_t)lda * 2, (uint64_t)(p.a.batched ? a_zs : (long long)M * lda) * 2};
  uint32_t ba[3] = {kBK, kBM, 1};
  int rc = make_tmap(&plan->ta, A, 3, da, sa, ba);
  if (rc) return rc;
  uint64_t db[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)(p.b.batched ? batch : 1)};
  uint64_t sb[2] = {(uint64_t)ldb * 2, (uint64_t)(p.b.batched ? b_zs : (long long)N * ldb) * 2};
  uint32_t bb[3] = {kBK, (uint32_t)plan->bn, 1};
  rc = make_tmap(&plan->tb, B, 3, db, sb, bb);
  if (rc) return rc;
  plan->grid = dim3((M + kBM - 1) / kBM, (N + plan->bn - 1) / plan->bn, batch);
  apply_splitk(plan, ep);
  return plan_output_map(plan);
}

int plan_conv3x3(GemmPlan* plan, const __half* x, int N, int H, int W, int Cin, const __half* w, int Cout,
                 const Epilogue& ep) {
  if (Cin % kBK != 0) {
    sdb_set_error("conv3x3: Cin=%d must be a multiple of %d for the tensor-core path", Cin, kBK);
    return SDB_ERR_UNSUPPORTED;
  }
  int bw, bh, bn;
  if (W >= kBM) {
    if (W % kBM) {
      sdb_set_error("conv3x3: W=%d must be a multiple of %d", W, kBM);
      return SDB_ERR_UNSUPPORTED;
    }
    bw = kBM, bh = 1, bn = 1;
  } else {
    if (kBM % W) {
      sdb_set_error("conv3x3: W=%d must divide %d", W, kBM);
      return SDB_ERR_UNSUPPORTED;
    }
    bw = W;
    if (H * W >= kBM) {
      bh = kBM / W, bn = 1;
      if (H % bh) {
        sdb_set_error("conv3x3: H=%d must be a multiple of %d", H, bh);
        return SDB_ERR_UNSUPPORTED;
      }
    } else {
      if (kBM % (H * W)) {
        sdb_set_error("conv3x3: H*W=%d must divide %d", H * W, kBM);
        return SDB_ERR_UNSUPPORTED;
      }
      bh = H, bn = kBM / (H * W);
    }
  }
  memset(plan, 0, sizeof(*plan));
  plan->bn = pick_bn(N * H * W, Cout, 9 * Cin);
  GemmParams& p = plan->p;
  p.M = N * H * W;
  p.N = Cout;
  p.K = 9 * Cin;
  p.num_k_blocks = 9 * (Cin / kBK);
  p.a.mode = kConv3x3;
  p.a.W = W;
  p.a.H = H;
  p.a.bw = bw;
  p.a.bh = bh;
  p.a.bn = bn;
  p.a.cin_blocks = Cin / kBK;
  p.b.mode = kMatrix;
  fill_epilogue(p, ep);
  uint64_t da[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t sa[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
  p.a.split_dim = bn >= 2 ? 3 : (bh >= 2 ? 2 : 1);
  uint32_t ba[4] = {kBK, (uint32_t)(p.a.split_dim == 1 ? bw / 2 : bw), (uint32_t)(p.a.split_dim == 2 ? bh / 2 : bh),
                    (uint32_t)(p.a.split_dim == 3 ? bn / 2 : bn)};
  int rc = make_tmap(&plan->ta, x, 4, da, sa, ba);
  if (rc) return rc;
  uint64_t db[3] = {(uint64_t)p.K, (uint64_t)Cout, 1};
  uint64_t sb[2] = {(uint64_t)p.K * 2, (uint64_t)Cout * p.K * 2};
  uint32_t bb[3] = {kBK, (uint32_t)plan->bn, 1};
  rc = make_tmap(&plan->tb, w, 3, db, sb, bb);
  if (rc) return rc;
  plan->grid = dim3((p.M + kBM - 1) / kBM, (Cout + plan->bn - 1) / plan->bn, 1);
  apply_splitk(plan, ep);
  return plan_output_map(plan);
}

int plan_attn_scores(GemmPlan* plan, const __half* Q, long long ldq, const __half* K, long long ldk, int B, int heads,
                     int head_dim, int Lq, int Lk, __half* S, long long lds, float alpha, int fuse_softmax) {
  if (fuse_softmax && (Lk > 80 || lds > 80 || (lds & 7))) {
    sdb_set_error("attention: fused softmax needs Lk <= 80 and lds <= 80 (multiple of 8); got %d / %lld", Lk, lds);
    return SDB_ERR_UNSUPPORTED;
  }
  if (head_dim % 64) {
    sdb_set_error("attention: head_dim=%d must be a multiple of 64", head_dim);
    return SDB_ERR_UNSUPPORTED;
  }
  memset(plan, 0, sizeof(*plan));
  plan->bn = Lk <= 80 ? 80 : 128;
  GemmParams& p = plan->p;
  p.M = Lq;
  p.N = Lk;
  p.K = head_dim;
  p.num_k_blocks = head_dim / kBK;
  p.a.mode = kHeads;
  p.a.heads = heads;
  p.b.mode = kHeads;
  p.b.heads = heads;
  Epilogue ep;
  ep.out = S;
  ep.ldc = lds;
  ep.alpha = alpha;
  fill_epilogue(p, ep);
  p.out_zs_hi = (long long)Lq * lds;  // S is [B*heads, Lq, lds]
  p.row_softmax = fuse_softmax ? 1 : 0;
  uint64_t dq[4] = {(uint64_t)head_dim, (uint64_t)heads, (uint64_t)Lq, (uint64_t)B};
  uint64_t sq[3] = {(uint64_t)head_dim * 2, (uint64_t)ldq * 2, (uint64_t)Lq * ldq * 2};
  uint32_t bq[4] = {64, 1, kBM, 1};
  int rc = make_tmap(&plan->ta, Q, 4, dq, sq, bq);
  if (rc) return rc;
  uint64_t dk[4] = {(uint64_t)head_dim, (uint64_t)heads, (uint64_t)Lk, (uint64_t)B};
  uint64_t sk[3] = {(uint64_t)head_dim * 2, (uint64_t)ldk * 2, (uint64_t)Lk * ldk * 2};
  uint32_t bk[4] = {64, 1, (uint32_t)plan->bn, 1};
  rc = make_tmap(&plan->tb, K, 4, dk, sk, bk);
  if (rc) return rc;
  plan->grid = dim3((Lq + kBM - 1) / kBM, (Lk + plan->bn - 1) / plan->bn, B * heads);
  return SDB_OK;
}

int plan_attn_apply(GemmPlan* plan, const __half* P, long long ldp, const __half* V, long long ldv, int B, int heads,
                    int head_dim, int Lq, int Lk, __half* O, long long ldo, float alpha) {
  if (head_dim % 64) {
    sdb_set_error("attention: head_dim=%d must be a multiple of 64", head_dim);
    return SDB_ERR_UNSUPPORTED;
  }
  memset(plan, 0, sizeof(*plan));
  plan->bn = 64;
  GemmParams& p = plan->p;
  p.M = Lq;
  p.N = head_dim;
  p.K = Lk;
  p.num_k_blocks = (Lk + kBK - 1) / kBK;
  p.a.mode = kMatrix;
  p.a.batched = 1;
  p.b.mode = kHeads;
  p.b.heads = heads;
  p.b.mn_major = 1;
  Epilogue ep;
  ep.out = O;
  ep.ldc = ldo;
  ep.alpha = alpha;
  fill_epilogue(p, ep);
  p.out_zdiv = heads;
  p.out_zs_hi = (long long)Lq * ldo;
  p.out_zs_lo = head_dim;
  uint64_t dp[3] = {(uint64_t)Lk, (uint64_t)Lq, (uint64_t)B * heads};
  uint64_t sp[2] = {(uint64_t)ldp * 2, (uint64_t)Lq * ldp * 2};
  uint32_t bp[3] = {kBK, kBM, 1};
  int rc = make_tmap(&plan->ta, P, 3, dp, sp, bp);
  if (rc) return rc;
  uint64_t dv[4] = {(uint64_t)head_dim, (uint64_t)heads, (uint64_t)Lk, (uint64_t)B};
  uint64_t sv[3] = {(uint64_t)head_dim * 2, (uint64_t)ldv * 2, (uint64_t)Lk * ldv * 2};
  uint32_t bv[4] = {64, 1, kBK, 1};
  rc = make_tmap(&plan->tb, V, 4, dv, sv, bv);
  if (rc) return rc;
  plan->grid = dim3((Lq + kBM - 1) / kBM, head_dim / 64, B * heads);
  return SDB_OK;
}

// Brackets a non-GEMM tensor-core launch (flash attention) with the same event records.
int profile_mark_begin(double flops, int M, int N, int K, int z, cudaStream_t stream) {
  if (!g_prof) return -1;
  ProfRec rec;
  cudaEventCreate(&rec.a);
  cudaEventCreate(&rec.b);
  rec.flops = flops;
  rec.M = M, rec.N = N, rec.K = K, rec.z = z, rec.splits = 1, rec.bn = 0, rec.mode = 9;
  cudaEventRecord(rec.a, stream);
  g_recs.push_back(rec);
  return (int)g_recs.size() - 1;
}
void profile_mark_end(int idx, cudaStream_t stream) {
  if (idx >= 0 && idx < (int)g_recs.size()) cudaEventRecord(g_recs[idx].b, stream);
}

void debug_timeline(unsigned long long* device_buf) { g_dbg = device_buf; }

void profile_dump_to(const char* path) {
  snprintf(g_prof_dump, sizeof(g_prof_dump), "%s", path ? path : "");
}

void profile_begin() {
  g_recs.clear();
  g_prof = true;
}

int profile_end(double* ms, double* flops, int* launches) {
  g_prof = false;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    sdb_set_error("profile_end: %s", cudaGetErrorString(e));
    return SDB_ERR_CUDA;
  }
  double t = 0.0, f = 0.0;
  FILE* fp = g_prof_dump[0] ? fopen(g_prof_dump, "w") : nullptr;
  if (fp) fprintf(fp, "M,N,K,batch,splits,bn,a_mode,ms,tflops\n");
  for (ProfRec& r : g_recs) {
    float m = 0.f;
    cudaEventElapsedTime(&m, r.a, r.b);
    if (fp)
      fprintf(fp, "%d,%d,%d,%d,%d,%d,%d,%.5f,%.1f\n", r.M, r.N, r.K, r.z, r.splits, r.bn, r.mode, m,
              r.flops / (m * 1e-3) / 1e12);
    t += m;
    f += r.flops;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  if (fp) fclose(fp);
  g_prof_dump[0] = 0;
  *ms = t;
  *flops = f;
  *launches = (int)g_recs.size();
  g_recs.clear();
  return SDB_OK;
}

int run_gemm(const GemmPlan& plan, cudaStream_t stream) {
  switch (plan.bn) {
    case 64: return launch<64>(plan, stream);
    case 80: return launch<80>(plan, stream);
    case 128: return launch<128>(plan, stream);
    case 160: return launch<160>(plan, stream);
    case 256: return launch_v<256, true>(plan, stream);
  }
  sdb_set_error("gemm: unsupported BN %d", plan.bn);
  return SDB_ERR_UNSUPPORTED;
}

}  // namespace dense
