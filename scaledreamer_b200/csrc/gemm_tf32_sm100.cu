// fp32-in / fp32-out GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, fp32 accumulation in tensor
// memory) for the one TRAINED dense network on the path, the Triplane-Transformer generator
// (custom/amortized/extern/triplane_transformer_modules.py:33-71,115-187; the reference trains it at `precision: 32`,
// configs/multi-prompt_benchmark/asd_mv_triplane_transformer_10k.yaml:127, i.e. cuBLAS fp32 / tf32 GEMMs).
//
//   out[z][M, N] = act(alpha * A[z][M, K] * B[z][N, K]^T + bias[N]) + residual[z][M, N]
//
// Same structure as the fp16 kernel (gemm_sm100.cu), one persistent CTA per SM:
//   warp 0   : TMA producer  (4-D maps {K, rows, z_lo, z_hi}, 128-byte swizzle: a row of a tile is 32 floats)
//   warp 1   : MMA issuer    (128 x BN x 8 per instruction, four per 32-float k-block; two accumulators in tensor memory)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      (tcgen05.ld 32 columns at a time, next read in flight while this one leaves: each warp stages
//                             its 32 rows x 128 bytes in shared memory (128-byte swizzle) and one lane issues a TMA store,
//                             three staging tiles per warp so that 48 KB of stores are in flight per SM. Per-thread
//                             16-byte stores cost one L2 request each and held the 3072 x 3072 x 48 score products at
//                             2.3 TB/s of output; two 2 KB stores in flight per warp at 3.1 TB/s: the store latency)
// A partial last k-block and rows past M / N are zero-filled by TMA (out-of-bounds fill), so K needs no padding:
// the attention products run with K = head_dim = 48 and K = 77 text tokens.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "dense.h"
#include "ptx_sm100.cuh"

namespace dense {
namespace {

constexpr int kBM = 128;
constexpr int kBKf = 32;  // floats per k-block row = 128 bytes
constexpr int kATile = kBM * 128;

template <int BN>
struct TCfg {
  static constexpr int kBTile = BN * 128;
  static constexpr int kStage = kATile + kBTile;
  static constexpr int kStages = BN <= 64 ? 6 : 5;
  static constexpr int kAccStride = BN <= 64 ? 64 : 128;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kThreads = 256;
  static constexpr int kStoreTiles = 3;                      // staging tiles per epilogue warp
  static constexpr int kStoreBytes = 4 * kStoreTiles * 4096;  // each 32 rows x 128 bytes
  static constexpr int kSmem = kStages * kStage + kStoreBytes + 1024 /*align*/ + 256 /*barriers*/;
};

struct Tf32Params {
  int M, N, K, num_k_blocks;
  int zdiv;
  int a_hi, a_lo, b_hi, b_lo;  // 1: the operand is indexed along that batch coordinate
  int a_mn;                    // A is stored [K][M] (M contiguous): MN-major tiles, four 32-wide M chunks per k-block
  float* out;
  long long ldc, out_zs_hi, out_zs_lo;
  const float* bias;
  const float* residual;
  long long ldr, res_zs_hi, res_zs_lo;
  float alpha;
  int act;
  int round_out;  // results are rounded to tf32: the output only feeds further GEMMs
  int tma_store;  // results leave through shared memory + cp.async.bulk.tensor (tmC); c_hi / c_lo as for the operands
  int c_hi, c_lo;
  int tiles_m, tiles_n, tiles_z;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ Tf32Params p) {
  using C = TCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* store_smem = base + C::kStages * C::kStage;  // 1024-byte aligned
  uint64_t* full = reinterpret_cast<uint64_t*>(store_smem + C::kStoreBytes);
  uint64_t* empty = full + C::kStages;
  uint64_t* accum_full = empty + C::kStages;  // [2]
  uint64_t* accum_empty = accum_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

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
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&accum_full[b], 1);
      ptx::mbar_init(&accum_empty[b], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
    ptx::tmem_relinquish();
  }
  pdl_wait();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int total = tiles_mn * p.tiles_z;
  const int nkb = p.num_k_blocks;

  if (warp == 0) {
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int t_mn = tile % tiles_mn, z = tile / tiles_mn;
        const int m0 = (t_mn % p.tiles_m) * kBM, n0 = (t_mn / p.tiles_m) * BN;
        const int zh = z / p.zdiv, zl = z - zh * p.zdiv;
        const int a3 = p.a_hi ? zh : 0, a2 = p.a_lo ? zl : 0, b3 = p.b_hi ? zh : 0, b2 = p.b_lo ? zl : 0;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&empty[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&full[s], C::kStage);
          uint8_t* sa = base + s * C::kStage;
          if (p.a_mn) {  // [32 K rows][32 M floats] boxes: chunk c holds M = m0 + 32 c .. + 31, 4 KB each
#pragma unroll
            for (int c = 0; c < 4; ++c) ptx::tma_load_4d(sa + c * 4096, &tmA, &full[s], m0 + 32 * c, kb * kBKf, a2, a3);
          } else {
            ptx::tma_load_4d(sa, &tmA, &full[s], kb * kBKf, m0, a2, a3);
          }
          ptx::tma_load_4d(sa + kATile, &tmB, &full[s], kb * kBKf, n0, b2, b3);
          if (++s == C::kStages) s = 0, ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_tf32(kBM, BN, p.a_mn);
      int s = 0;
      uint32_t ph = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1u, use = tcount >> 1;
        const uint32_t tmem_acc = tmem_base + buf * C::kAccStride;
        ptx::mbar_wait(&accum_empty[buf], (use & 1u) ^ 1u);
        ptx::tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(base + s * C::kStage);
          // MN-major A (stored [K][M]): 32-bit operands transpose only in the "128-byte swizzle, 32-byte atom" layout
          // (TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, descriptor layout type 1): atoms of 4 K rows x 128 bytes, 512 bytes
          // apart along K (SBO), the 32-wide M chunks 4096 bytes apart (LBO). With the plain 128-byte swizzle (type 2) the
          // product came back as zeros.
          const uint64_t da = p.a_mn ? ptx::smem_desc_mn_sw128(sa, 4096, 1, 512) : ptx::smem_desc_k_sw128(sa);
          const uint64_t db = ptx::smem_desc_k_sw128(sa + kATile);
#pragma unroll
          for (int k = 0; k < kBKf / 8; ++k)  // 8 tf32 = 32 bytes further along the swizzled row
            ptx::umma_tf32(tmem_acc, da + (uint64_t)(p.a_mn ? k * 64 : k * 2), db + (uint64_t)(k * 2), idesc,
                           (kb | k) != 0 ? 1u : 0u);  // 8 K rows further: 1024 bytes (MN-major) or 32 bytes along the row
          ptx::umma_commit(&empty[s]);
          if (++s == C::kStages) s = 0, ph ^= 1u;
        }
        ptx::umma_commit(&accum_full[buf]);
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;
    uint32_t tcount = 0;
    uint8_t* stage = store_smem + wq * (C::kStoreTiles * 4096);
    int nstore = 0;  // TMA stores issued by this warp's lane 0 so far (staging tile = nstore % kStoreTiles)
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
      const int t_mn = tile % tiles_mn, z = tile / tiles_mn;
      const int m0 = (t_mn % p.tiles_m) * kBM, n0 = (t_mn / p.tiles_m) * BN;
      const int zh = z / p.zdiv, zl = z - zh * p.zdiv;
      const uint32_t buf = tcount & 1u, use = tcount >> 1;
      const int m = m0 + wq * 32 + lane;
      const bool m_ok = m < p.M;
      float* orow = p.out + (long long)zh * p.out_zs_hi + (long long)zl * p.out_zs_lo + (long long)m * p.ldc;
      const float* rrow =
          p.residual ? p.residual + (long long)zh * p.res_zs_hi + (long long)zl * p.res_zs_lo + (long long)m * p.ldr : nullptr;
      const bool vec_ok = ((reinterpret_cast<uintptr_t>(orow) & 15) == 0) &&
                          (!rrow || (reinterpret_cast<uintptr_t>(rrow) & 15) == 0);
      const int c3 = p.c_hi ? zh : 0, c2 = p.c_lo ? zl : 0;
      ptx::mbar_wait(&accum_full[buf], use & 1u);
      ptx::tc_fence_after();
      uint32_t tm_row = tmem_base + buf * C::kAccStride + ((uint32_t)(wq * 32) << 16);
      auto finish = [&](const uint32_t(&v)[32], int c) {
        const int nb = n0 + c;
        if (nb >= p.N) return;              // warp-uniform
        if (!p.tma_store && !m_ok) return;  // rows past M: nothing to store (the TMA path clips them, stays collective)
        const bool fullc = nb + 32 <= p.N;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
        if (p.bias) {
          if (fullc && (reinterpret_cast<uintptr_t>(p.bias + nb) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + q);
              f[4 * q] += t.x, f[4 * q + 1] += t.y, f[4 * q + 2] += t.z, f[4 * q + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < p.N) f[j] += __ldg(p.bias + nb + j);
          }
        }
        if (p.act == kActGelu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = 0.5f * f[j] * (1.f + erff(f[j] * 0.70710678118654752f));
        }
        if (rrow && m_ok) {
          if (fullc && vec_ok && (nb & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t = *(reinterpret_cast<const float4*>(rrow + nb) + q);
              f[4 * q] += t.x, f[4 * q + 1] += t.y, f[4 * q + 2] += t.z, f[4 * q + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < p.N) f[j] += rrow[nb + j];
          }
        }
        if (p.round_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = round_tf32(f[j]);
        }
        if (p.tma_store) {
          // 32 rows x 128 bytes; the 16-byte piece q of row r sits at piece q ^ (r & 7) (128-byte swizzle of tmC)
          uint8_t* tile_s = stage + (nstore % C::kStoreTiles) * 4096;
          if (nstore >= C::kStoreTiles) {  // the store that last read this staging tile must have drained it
            if (lane == 0) ptx::tma_store_wait_read<C::kStoreTiles - 1>();
            __syncwarp();
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(tile_s + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_4d(&tmC, tile_s, nb, m - lane, c2, c3);  // columns past N / rows past M are clipped
            ptx::tma_store_commit();
          }
          ++nstore;
        } else if (fullc && vec_ok && (nb & 3) == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            reinterpret_cast<float4*>(orow + nb)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) orow[nb + j] = f[j];
        }
      };
      uint32_t va[32], vb[32];
      ptx::tmem_ld_32x32(tm_row, va);
#pragma unroll 1
      for (int c = 0; c < BN; c += 64) {
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32(tm_row + (uint32_t)(c + 32), vb);
        finish(va, c);
        ptx::tmem_ld_wait();
        if (c + 64 < BN) ptx::tmem_ld_32x32(tm_row + (uint32_t)(c + 64), va);
        finish(vb, c + 32);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&accum_empty[buf]);
    }
    if (p.tma_store && nstore > 0) {
      if (lane == 0) ptx::tma_store_wait_read<0>();
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

int operand_map(CUtensorMap* tm, const Tf32Operand& o, int rows, int K, int z_hi, int z_lo, int box_rows, int* use_hi,
                int* use_lo, const char* name) {
  if ((o.ld & 3) || (o.zs_hi & 3) || (o.zs_lo & 3)) {
    sdb_set_error("gemm_tf32: %s strides (ld %lld, zs_hi %lld, zs_lo %lld) must be multiples of 4 floats", name, o.ld,
                  o.zs_hi, o.zs_lo);
    return SDB_ERR_ARG;
  }
  *use_hi = (z_hi > 1 && o.zs_hi != 0) ? 1 : 0;
  *use_lo = (z_lo > 1 && o.zs_lo != 0) ? 1 : 0;
  // an index the operand is shared along becomes a dimension of extent 1 (any valid stride)
  const uint64_t dflt = (uint64_t)o.ld * 4 * (uint64_t)rows;
  uint64_t dims[4] = {(uint64_t)K, (uint64_t)rows, (uint64_t)(*use_lo ? z_lo : 1), (uint64_t)(*use_hi ? z_hi : 1)};
  uint64_t strides[3] = {(uint64_t)o.ld * 4, *use_lo ? (uint64_t)o.zs_lo * 4 : dflt, *use_hi ? (uint64_t)o.zs_hi * 4 : dflt};
  uint32_t box[4] = {(uint32_t)kBKf, (uint32_t)box_rows, 1, 1};
  return make_tmap(tm, o.ptr, 4, dims, strides, box, o.mn_major ? 12832 : 128, 1);
}

template <int BN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const Tf32Params& p, int grid,
           cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCfg<BN>::kSmem);
    if (e != cudaSuccess) {
      sdb_set_error("gemm_tf32: smem attribute: %s", cudaGetErrorString(e));
      return SDB_ERR_CUDA;
    }
    attr = true;
  }
  sdb_launch(gemm_tf32_kernel<BN>, dim3(grid), dim3(256), (size_t)TCfg<BN>::kSmem, stream, ta, tb, tc, p);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("gemm_tf32");
  return SDB_OK;
}

}  // namespace

int gemm_tf32(const Tf32Operand& A, const Tf32Operand& B, int M, int N, int K, float* out, long long ldc, int batch,
              int zdiv, long long out_zs_hi, long long out_zs_lo, const Tf32Epilogue& ep, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0 || zdiv <= 0 || batch % zdiv != 0) {
    sdb_set_error("gemm_tf32: bad shape M=%d N=%d K=%d batch=%d zdiv=%d", M, N, K, batch, zdiv);
    return SDB_ERR_ARG;
  }
  if (ep.act != kActNone && ep.act != kActGelu) {
    sdb_set_error("gemm_tf32: activation %d is not wired (none | gelu)", ep.act);
    return SDB_ERR_UNSUPPORTED;
  }
  const int bn = N <= 64 ? 64 : 128;
  Tf32Params p;
  memset(&p, 0, sizeof(p));
  p.M = M, p.N = N, p.K = K;
  p.num_k_blocks = (K + kBKf - 1) / kBKf;
  p.zdiv = zdiv;
  p.out = out, p.ldc = ldc, p.out_zs_hi = out_zs_hi, p.out_zs_lo = out_zs_lo;
  p.bias = ep.bias, p.residual = ep.residual, p.ldr = ep.ldr, p.res_zs_hi = ep.res_zs_hi, p.res_zs_lo = ep.res_zs_lo;
  p.alpha = ep.alpha, p.act = ep.act, p.round_out = ep.round_out;
  p.tiles_m = (M + kBM - 1) / kBM, p.tiles_n = (N + bn - 1) / bn, p.tiles_z = batch;
  CUtensorMap ta, tb;
  p.a_mn = A.mn_major ? 1 : 0;
  int rc = A.mn_major ? operand_map(&ta, A, K, M, batch / zdiv, zdiv, kBKf, &p.a_hi, &p.a_lo, "A (MN-major)")
                      : operand_map(&ta, A, M, K, batch / zdiv, zdiv, kBM, &p.a_hi, &p.a_lo, "A");
  if (rc) return rc;
  rc = operand_map(&tb, B, N, K, batch / zdiv, zdiv, bn, &p.b_hi, &p.b_lo, "B");
  if (rc) return rc;
  // output map {N, M, z_lo, z_hi}, 32-column x 32-row boxes, 128-byte swizzle
  CUtensorMap tc;
  memset(&tc, 0, sizeof(tc));
  static const int tma_store_on = getenv("SDB_TF32_TMA_STORE") ? atoi(getenv("SDB_TF32_TMA_STORE")) : 1;
  // (N % 4 != 0 keeps the per-thread stores: the bulk store writes whole 16-byte granules, i.e. it would also write the
  // columns between N and the next multiple of 4 -- measured with N = 5, ldc = 8)
  if (tma_store_on && (N & 3) == 0 && (ldc & 3) == 0 && (out_zs_hi & 3) == 0 && (out_zs_lo & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int z_hi = batch / zdiv, z_lo = zdiv;
    p.c_hi = (z_hi > 1 && out_zs_hi != 0) ? 1 : 0;
    p.c_lo = (z_lo > 1 && out_zs_lo != 0) ? 1 : 0;
    const uint64_t dflt = (uint64_t)ldc * 4 * (uint64_t)M;
    uint64_t dims[4] = {(uint64_t)N, (uint64_t)M, (uint64_t)(p.c_lo ? z_lo : 1), (uint64_t)(p.c_hi ? z_hi : 1)};
    uint64_t strides[3] = {(uint64_t)ldc * 4, p.c_lo ? (uint64_t)out_zs_lo * 4 : dflt, p.c_hi ? (uint64_t)out_zs_hi * 4 : dflt};
    uint32_t box[4] = {32, 32, 1, 1};
    rc = make_tmap(&tc, out, 4, dims, strides, box, 128, 1);
    if (rc) return rc;
    p.tma_store = 1;
  }
  const long long total = (long long)p.tiles_m * p.tiles_n * p.tiles_z;
  const int grid = (int)(total < kNumSMs ? total : kNumSMs);
  return bn == 64 ? launch<64>(ta, tb, tc, p, grid, stream) : launch<128>(ta, tb, tc, p, grid, stream);
}

}  // namespace dense
