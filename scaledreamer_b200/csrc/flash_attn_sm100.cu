// Fused multi-head attention forward for sm_100a (head_dim 64): O = softmax(alpha Q K^T) V without materialising the
// score matrix. One CTA owns 128 query rows of one (batch, head) and streams the keys in blocks of 128:
//   warp 0      : TMA producer   (Q once; K double-buffered, V single-buffered; 128-byte-swizzled tiles)
//   warp 1      : tcgen05 issuer (S = Q K^T into tensor memory, then PV = P V with P read from shared memory)
//   warps 2..5  : softmax, one thread per query row: tcgen05.ld of the S row, online max / sum (exp2 with the scale
//                 folded in), P written as fp16 into a K-major swizzled tile, running O kept in registers and
//                 updated from the PV accumulator of the PREVIOUS block (so the PV MMA overlaps the next softmax).
// Two CTAs share an SM; the MUFU exp2 throughput is the roofline of this kernel.
//
// Replaces CrossAttention.forward for self-attention (extern/mvdream/ldm/modules/attention.py:163-194: einsum QK^T,
// softmax, einsum PV) / diffusers AttnProcessor2_0 -> SDPA (stable_diffusion_asd_guidance.py:318-331).
#include <cstring>

#include "dense.h"
#include "ptx_sm100.cuh"

namespace dense {
namespace {

constexpr int kQ = 128;    // query rows per CTA
constexpr int kKV = 128;   // keys per block
constexpr int kD = 64;     // head dim
constexpr int kFaThreads = 192;
constexpr int kTileBytes = 128 * kD * 2;  // 16 KB: 128 rows x 128 bytes
constexpr int kFaSmem = kTileBytes * (1 + 2 + 1 + 2) + 1024 + 256;

struct FlashParams {
  int Lq, Lk, heads;
  float scale_log2;  // alpha * log2(e)
  __half* out;
  long long ldo;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kFaThreads, 2)
flash_attn_f16_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const FlashParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = base;
  uint8_t* sK = sQ + kTileBytes;       // 2 stages
  uint8_t* sV = sK + 2 * kTileBytes;
  uint8_t* sP = sV + kTileBytes;       // 2 K-blocks of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;   // 128 arrivals
  uint64_t* pv_full = bars + 9;
  uint64_t* pv_empty = bars + 10;  // 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQ;
  const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
  const int nblk = (p.Lk + kKV - 1) / kKV;

  pdl_launch_dependents();
  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], 1);
    }
    ptx::mbar_init(v_full, 1);
    ptx::mbar_init(v_empty, 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(pv_full, 1);
    ptx::mbar_init(pv_empty, 128);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<256>(tmem_slot);
    ptx::tmem_relinquish();
  }
  pdl_wait();  // set-up above touched no global memory; q / k / v come from the preceding kernel
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_PV = tmem_base + 128u;

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(q_full, kTileBytes);
      ptx::tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        ptx::mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(&k_full[s], kTileBytes);
        ptx::tma_load_4d(sK + s * kTileBytes, &tmK, &k_full[s], 0, h, j * kKV, b);
        ptx::mbar_wait(v_empty, (j & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(v_full, kTileBytes);
        ptx::tma_load_4d(sV, &tmV, v_full, 0, h, j * kKV, b);
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc_s = ptx::make_idesc_f16(kQ, kKV, 0, 0, 0);
      const uint32_t idesc_pv = ptx::make_idesc_f16(kQ, kD, 0, 0, 1);
      const uint64_t dq = ptx::smem_desc_k_sw128(ptx::smem_u32(sQ));
      const uint64_t dp0 = ptx::smem_desc_k_sw128(ptx::smem_u32(sP));
      const uint64_t dp1 = ptx::smem_desc_k_sw128(ptx::smem_u32(sP + kTileBytes));
      const uint64_t dv = ptx::smem_desc_mn_sw128(ptx::smem_u32(sV), 8192);
      ptx::mbar_wait(q_full, 0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        // S_j = Q K_j^T (S_{j-1} is free: this thread saw p_full_{j-1} before it issued PV_{j-1})
        ptx::mbar_wait(&k_full[s], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint64_t dk = ptx::smem_desc_k_sw128(ptx::smem_u32(sK + s * kTileBytes));
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          ptx::umma_f16(tmem_S, dq + (uint64_t)((k * 32) >> 4), dk + (uint64_t)((k * 32) >> 4), idesc_s, k != 0 ? 1u : 0u);
        ptx::umma_commit(&k_empty[s]);
        ptx::umma_commit(s_full);
        // PV_j = P_j V_j once P_j is in shared memory and the previous PV accumulator has been consumed
        ptx::mbar_wait(p_full, j & 1);
        ptx::mbar_wait(v_full, j & 1);
        if (j > 0) ptx::mbar_wait(pv_empty, (j - 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < kKV / 16; ++k) {
          const uint64_t da = (k < 4 ? dp0 : dp1) + (uint64_t)(((k & 3) * 32) >> 4);
          ptx::umma_f16(tmem_PV, da, dv + (uint64_t)((k * 2048) >> 4), idesc_pv, k != 0 ? 1u : 0u);
        }
        ptx::umma_commit(v_empty);
        ptx::umma_commit(pv_full);
      }
    }
  } else {
    // ---- softmax / output: thread per query row; TMEM lane quadrant = warp % 4
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    float O[kD];
#pragma unroll
    for (int d = 0; d < kD; ++d) O[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = sP + row * 128;
    const int sw = row & 7;
    for (int j = 0; j < nblk; ++j) {
      ptx::mbar_wait(s_full, j & 1);
      ptx::tc_fence_after();
      const int kv0 = j * kKV;
      const bool tail = kv0 + kKV > p.Lk;
      // pass 1: block max
      float mx = m_run;
#pragma unroll 1
      for (int c = 0; c < kKV; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_S + lane_off + (uint32_t)c, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = __uint_as_float(v[i]);
          if (!tail || kv0 + c + i < p.Lk) mx = fmaxf(mx, sv);
        }
      }
      const float m_new = mx;  // finite: every block holds at least one valid key
      const float corr = fast_exp2((m_run - m_new) * p.scale_log2);  // exp2(-inf) = 0 on the first block
      const float moff = m_new * p.scale_log2;
      // previous block's PV accumulator -> O, then rescale to the new max
      if (j > 0) {
        ptx::mbar_wait(pv_full, (j - 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < kD; c += 32) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_PV + lane_off + (uint32_t)c, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) O[c + i] += __uint_as_float(v[i]);
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(pv_empty);
      }
#pragma unroll
      for (int d = 0; d < kD; ++d) O[d] *= corr;
      // pass 2: p = exp2(s * scale - m * scale), row sum, fp16 P into the swizzled K-major tile
      float lsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < kKV; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_S + lane_off + (uint32_t)c, v);
        ptx::tmem_ld_wait();
        uint8_t* pblk = prow + (c >> 6) * kTileBytes;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 ov;
          __half2* h2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = q * 8 + 2 * e;
            float p0 = fast_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2, -moff));
            float p1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -moff));
            if (tail) {
              if (kv0 + c + i >= p.Lk) p0 = 0.f;
              if (kv0 + c + i + 1 >= p.Lk) p1 = 0.f;
            }
            lsum += p0 + p1;
            h2[e] = __floats2half2_rn(p0, p1);
          }
          const int chunk = ((c & 63) >> 3) + q;  // 16-byte chunk within the 128-byte row
          *reinterpret_cast<uint4*>(pblk + ((chunk ^ sw) << 4)) = ov;
        }
      }
      l_run = l_run * corr + lsum;
      m_run = m_new;
      // make the generic-proxy writes of P visible to the tensor core (async proxy), release S
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
    }
    // last PV block
    ptx::mbar_wait(pv_full, (nblk - 1) & 1);
    ptx::tc_fence_after();
#pragma unroll
    for (int c = 0; c < kD; c += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_PV + lane_off + (uint32_t)c, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) O[c + i] += __uint_as_float(v[i]);
    }
    ptx::tc_fence_before();
    const int qrow = q0 + row;
    if (qrow < p.Lq) {
      const float inv = 1.f / l_run;
      __half* o = p.out + ((long long)b * p.Lq + qrow) * p.ldo + h * kD;
#pragma unroll
      for (int q = 0; q < kD / 8; ++q) {
        uint4 ov;
        __half2* h2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(O[q * 8 + 2 * e] * inv, O[q * 8 + 2 * e + 1] * inv);
        reinterpret_cast<uint4*>(o)[q] = ov;
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace

int plan_flash_attn(FlashPlan* plan, const __half* Q, long long ldq, const __half* K, long long ldk, const __half* V,
                    long long ldv, int B, int heads, int head_dim, int Lq, int Lk, __half* O, long long ldo,
                    float alpha) {
  if (head_dim != kD) {
    sdb_set_error("flash attention: head_dim must be %d (got %d)", kD, head_dim);
    return SDB_ERR_UNSUPPORTED;
  }
  if ((ldo & 7) || (reinterpret_cast<uintptr_t>(O) & 15)) {
    sdb_set_error("flash attention: output rows must be 16-byte aligned");
    return SDB_ERR_ARG;
  }
  memset(plan, 0, sizeof(*plan));
  const __half* ptrs[3] = {Q, K, V};
  const long long lds[3] = {ldq, ldk, ldv};
  const int Ls[3] = {Lq, Lk, Lk};
  CUtensorMap* maps[3] = {&plan->tq, &plan->tk, &plan->tv};
  for (int i = 0; i < 3; ++i) {
    uint64_t d[4] = {(uint64_t)kD, (uint64_t)heads, (uint64_t)Ls[i], (uint64_t)B};
    uint64_t st[3] = {(uint64_t)kD * 2, (uint64_t)lds[i] * 2, (uint64_t)Ls[i] * lds[i] * 2};
    uint32_t box[4] = {64, 1, 128, 1};
    int rc = make_tmap(maps[i], ptrs[i], 4, d, st, box);
    if (rc) return rc;
  }
  plan->Lq = Lq;
  plan->Lk = Lk;
  plan->heads = heads;
  plan->scale_log2 = alpha * 1.4426950408889634f;
  plan->out = O;
  plan->ldo = ldo;
  plan->grid = dim3((Lq + kQ - 1) / kQ, B * heads, 1);
  plan->flops = 4.0 * B * heads * (double)Lq * Lk * kD;
  return SDB_OK;
}

int run_flash_attn(const FlashPlan& plan, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem);
    if (e != cudaSuccess) {
      sdb_set_error("flash attention: smem attribute: %s", cudaGetErrorString(e));
      return SDB_ERR_CUDA;
    }
    attr = true;
  }
  FlashParams p;
  p.Lq = plan.Lq, p.Lk = plan.Lk, p.heads = plan.heads, p.scale_log2 = plan.scale_log2, p.out = plan.out, p.ldo = plan.ldo;
  const int rec = profile_mark_begin(plan.flops, plan.Lq, plan.Lk, kD, (int)plan.grid.y, stream);
  sdb_launch(flash_attn_f16_kernel, plan.grid, kFaThreads, kFaSmem, stream, plan.tq, plan.tk, plan.tv, p);
  profile_mark_end(rec, stream);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("flash_attn_f16");
  return SDB_OK;
}

}  // namespace dense
