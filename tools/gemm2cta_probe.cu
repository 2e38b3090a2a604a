// Probe for DESIGN.md §9 item 1: does a tcgen05 CTA-PAIR tile (cta_group::2, 256 x BN per pair, each CTA fetching half of
// B) lift the L2->SM operand-delivery bound measured on the 128 x BN single-CTA tiles?
//
// STATUS: compiles for sm_100a, NEVER RUN (written after the round-1 GPU budget was spent). It is a stand-alone program,
// not part of libsdb200.so. Every barrier wait is bounded (clock64 budget) and reports a time-out instead of spinning, so
// a protocol mistake shows up as "TIMEOUT at <site>" and exit code 2, not as a hung GPU.
//
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -I scaledreamer_b200/csrc \
//        tools/gemm2cta_probe.cu -o gpurun_out/gemm2cta_probe
//   timeout 60 gpurun_out/gemm2cta_probe [M N K]          (default 8192 2560 2880)
//
// D[M,N] (fp16) = A[M,K] * B[N,K]^T, fp16 operands, fp32 accumulation. Two kernels over the same operands:
//   single<BN>  one CTA per 128 x BN tile  - the main loop of csrc/gemm_sm100.cu reduced to one tile per CTA
//   pair<BN>    one 2-CTA cluster per 256 x BN tile:
//                 CTA r loads A rows m0 + 128 r and B rows n0 + (BN/2) r; both TMAs complete on the LEADER's full barrier
//                 (the leader arms it with both CTAs' bytes); the leader's elected thread issues
//                 tcgen05.mma.cta_group::2 (M = 256); stage release and accumulator-ready are multicast commits
//                 (mask 0b11) so each CTA waits on its own barriers; each CTA drains its own 128 TMEM lanes.
// Both are checked against a CUDA-core dot product on sampled elements and timed with CUDA events (20 launches).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx_sm100.cuh"

namespace {

constexpr int kBK = 64;
constexpr int kATile = 128 * kBK * 2;  // 16 KB
constexpr long long kSpinBudget = 2000000000LL;  // ~1 s of SM clocks

__device__ int g_timeout_site = 0;

__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity, int site) {
  const long long t0 = clock64();
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(ptx::smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((spins & 1023u) != 1023u) continue;  // the budget / abort flag are looked at once per 1024 polls
    if (clock64() - t0 > kSpinBudget || *(volatile int*)&g_timeout_site != 0) {
      atomicCAS(&g_timeout_site, 0, site);
      return false;
    }
  }
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (an address in this CTA's shared memory) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(const void* p, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(ptx::smem_u32(p)), "r"(rank));
  return out;
}
// TMA load whose completion bytes are credited to a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(ptx::smem_u32(smem)),
      "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          ptx::smem_u32(smem)),
      "l"(m), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in every CTA of `mask` once the issued MMAs have completed
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   ptx::smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// one row per thread: 128 TMEM lanes x BN fp32 columns -> fp16 global
template <int BN>
__device__ __forceinline__ void drain(uint32_t tmem_base, int warp_q, int lane, __half* out, long long ldc, int row, int n0) {
  __half* o = out + (long long)row * ldc + n0;
#pragma unroll 1
  for (int c = 0; c < BN; c += 32) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp_q * 32) << 16) + (uint32_t)c, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 pk;
      __half2* h = reinterpret_cast<__half2*>(&pk);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        h[e] = __floats2half2_rn(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1]));
      reinterpret_cast<uint4*>(o + c)[q] = pk;
    }
  }
}

template <int BN, int STAGES>
struct Smem {
  static constexpr int kBTile = BN * kBK * 2;
  static constexpr int kStage = kATile + kBTile;
  static constexpr int kBytes = STAGES * kStage + 1024 + 256;
};

// ---------------------------------------------------------------------------------------------- single CTA per tile
template <int BN, int STAGES>
__global__ void __launch_bounds__(256, 1)
single_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, __half* out, int M, int N, int K) {
  using S = Smem<BN, STAGES>;
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * S::kStage);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = M / 128;
  const int m0 = (blockIdx.x % tiles_m) * 128, n0 = (blockIdx.x / tiles_m) * BN;
  const int nkb = K / kBK;
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < STAGES; ++s) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
    ptx::mbar_init(acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<BN>(slot);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    if (ptx::elect_one())
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        if (!wait_bounded(&empty[s], ((kb / STAGES) & 1) ^ 1u, 11)) break;
        ptx::mbar_arrive_expect_tx(&full[s], S::kStage);
        tma_load_2d(base + s * S::kStage, &tmA, &full[s], kb * kBK, m0);
        tma_load_2d(base + s * S::kStage + kATile, &tmB, &full[s], kb * kBK, n0);
      }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_f16(128, BN, 0, 0, 0);
      bool ok = true;
      for (int kb = 0; kb < nkb && ok; ++kb) {
        const int s = kb % STAGES;
        ok = wait_bounded(&full[s], (kb / STAGES) & 1, 12);
        if (!ok) break;
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(base + s * S::kStage);
        const uint64_t da = ptx::smem_desc_k_sw128(sa), db = ptx::smem_desc_k_sw128(sa + kATile);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          ptx::umma_f16(tmem, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), idesc, (kb | k) != 0);
        ptx::umma_commit(&empty[s]);
      }
      ptx::umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    if (wait_bounded(acc_full, 0, 13)) {
      ptx::tc_fence_after();
      drain<BN>(tmem, warp & 3, lane, out, N, m0 + (warp & 3) * 32 + lane, n0);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<BN>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------- CTA pair per tile
template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhalf, __half* out, int M, int N,
            int K) {
  constexpr int kBHalf = (BN / 2) * kBK * 2;   // this CTA's half of the B tile
  constexpr int kStage = kATile + kBHalf;      // per CTA
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * kStage);  // used in the leader only
  uint64_t* empty = full + STAGES;                                        // one per CTA, armed by the multicast commit
  uint64_t* acc_full = empty + STAGES;                                    // one per CTA
  uint32_t* slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int tiles_m = M / 256;
  const int m0 = (pair % tiles_m) * 256, n0 = (pair / tiles_m) * BN;
  const int nkb = K / kBK;
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < STAGES; ++s) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
    ptx::mbar_init(acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) tmem_alloc2<BN>(slot);  // both CTAs of the pair execute the paired allocation (as CUTLASS / DeepGEMM do)
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    if (ptx::elect_one())
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        if (!wait_bounded(&empty[s], ((kb / STAGES) & 1) ^ 1u, 21)) break;
        if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * kStage);  // both CTAs' bytes land on the leader's barrier
        const uint32_t leader_full = mapa(&full[s], 0);
        tma_load_2d_pair(base + s * kStage, &tmA, leader_full, kb * kBK, m0 + 128 * (int)rank);
        tma_load_2d_pair(base + s * kStage + kATile, &tmBhalf, leader_full, kb * kBK, n0 + (BN / 2) * (int)rank);
      }
  } else if (warp == 1 && rank == 0) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_f16(256, BN, 0, 0, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        if (!wait_bounded(&full[s], (kb / STAGES) & 1, 22)) break;
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(base + s * kStage);  // same offsets in the peer CTA
        const uint64_t da = ptx::smem_desc_k_sw128(sa), db = ptx::smem_desc_k_sw128(sa + kATile);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          umma2_f16(tmem, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), idesc, (kb | k) != 0);
        umma2_commit_mc(&empty[s], 0b11);
      }
      umma2_commit_mc(acc_full, 0b11);
    }
  } else if (warp >= 4) {
    if (wait_bounded(acc_full, 0, 23)) {
      ptx::tc_fence_after();
      drain<BN>(tmem, warp & 3, lane, out, N, m0 + 128 * (int)rank + (warp & 3) * 32 + lane, n0);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while its partner can still touch it
  if (warp == 2) {
    ptx::tc_fence_after();
    tmem_dealloc2<BN>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------- host
__global__ void fill_kernel(__half* p, long long n, uint32_t seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u ^ seed;
    x ^= x >> 15, x *= 2246822519u, x ^= x >> 13;
    p[i] = __float2half(((x & 0xffff) / 65535.0f - 0.5f) * 0.25f);
  }
}
__global__ void check_kernel(const __half* A, const __half* B, const __half* D, int M, int N, int K, int samples, float* max_err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= samples) return;
  const uint32_t h = (uint32_t)i * 2654435761u;
  const int m = (int)(h % (uint32_t)M), n = (int)((h >> 7) % (uint32_t)N);
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __half2float(A[(long long)m * K + k]) * __half2float(B[(long long)n * K + k]);
  const float err = fabsf(acc - __half2float(D[(long long)m * N + n])) / (fabsf(acc) + 1e-2f);
  atomicMax(reinterpret_cast<int*>(max_err), __float_as_int(err));  // err >= 0: int order == float order
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_map(CUtensorMap* out, EncodeTiledFn enc, void* ptr, int rows, int K, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

template <typename Launch>
int run(const char* name, Launch launch, const __half* A, const __half* B, __half* D, int M, int N, int K) {
  CK(cudaMemset(D, 0, (size_t)M * N * 2));
  int zero = 0;
  CK(cudaMemcpyToSymbol(g_timeout_site, &zero, sizeof(int)));
  launch();
  CK(cudaDeviceSynchronize());
  int site = 0;
  CK(cudaMemcpyFromSymbol(&site, g_timeout_site, sizeof(int)));
  if (site) {
    printf("%-14s TIMEOUT at wait site %d (1x: single, 2x: pair; x1 producer/empty, x2 mma/full, x3 epilogue/accumulator)\n",
           name, site);
    return 2;
  }
  float* err;
  CK(cudaMalloc(&err, 4));
  CK(cudaMemset(err, 0, 4));
  check_kernel<<<(4096 + 127) / 128, 128>>>(A, B, D, M, N, K, 4096, err);
  float h = 0;
  CK(cudaMemcpy(&h, err, 4, cudaMemcpyDeviceToHost));
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) launch();
  cudaEventRecord(a);
  for (int i = 0; i < 20; ++i) launch();
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 20;
  printf("%-14s max rel err %.3e (%s)  %.1f us  %.0f TFLOP/s\n", name, h, h < 2e-2f ? "ok" : "WRONG", ms * 1e3,
         2.0 * M * N * K / (ms * 1e-3) / 1e12);
  return h < 2e-2f ? 0 : 3;
}

}  // namespace

int main(int argc, char** argv) {
  const int M = argc > 3 ? atoi(argv[1]) : 8192, N = argc > 3 ? atoi(argv[2]) : 2560, K = argc > 3 ? atoi(argv[3]) : 2880;
  constexpr int BN = 256;
  if (M % 256 || N % BN || K % kBK) {
    printf("M %% 256, N %% %d, K %% 64 must be 0\n", BN);
    return 1;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  __half *A, *B, *D;
  CK(cudaMalloc(&A, (size_t)M * K * 2));
  CK(cudaMalloc(&B, (size_t)N * K * 2));
  CK(cudaMalloc(&D, (size_t)M * N * 2));
  fill_kernel<<<1024, 256>>>(A, (long long)M * K, 1u);
  fill_kernel<<<1024, 256>>>(B, (long long)N * K, 2u);
  CUtensorMap tmA, tmB, tmBh;
  if (!make_map(&tmA, enc, A, M, K, 128) || !make_map(&tmB, enc, B, N, K, BN) || !make_map(&tmBh, enc, B, N, K, BN / 2)) {
    printf("cuTensorMapEncodeTiled failed\n");
    return 1;
  }
  constexpr int S1 = 4, S2 = 6;  // 48 KB x 4 = 192 KB ; 32 KB x 6 = 192 KB per CTA
  CK(cudaFuncSetAttribute(single_kernel<BN, S1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN, S1>::kBytes));
  constexpr int kPairSmem = S2 * (kATile + (BN / 2) * kBK * 2) + 1024 + 256;
  CK(cudaFuncSetAttribute(pair_kernel<BN, S2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
  printf("D[%d,%d] = A[%d,%d] B^T, fp16, tiles 128x%d (single) vs 256x%d (pair)\n", M, N, M, K, BN, BN);
  int rc = run("single 128xBN", [&] { single_kernel<BN, S1><<<(M / 128) * (N / BN), 256, Smem<BN, S1>::kBytes>>>(tmA, tmB, D, M, N, K); },
               A, B, D, M, N, K);
  if (rc) return rc;
  rc = run("pair 256xBN", [&] { pair_kernel<BN, S2><<<(M / 256) * (N / BN) * 2, 256, kPairSmem>>>(tmA, tmBh, D, M, N, K); }, A, B, D,
           M, N, K);
  return rc;
}
