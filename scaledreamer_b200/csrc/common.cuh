// Shared device helpers for the sm_100a ASD-step kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define SDB_OK 0
#define SDB_ERR_ARG -1
#define SDB_ERR_CUDA -2
#define SDB_ERR_UNSUPPORTED -3

// Sets the thread-local last-error string (defined in capi.cu).
void sdb_set_error(const char* fmt, ...);

#define SDB_CHECK_ARG(cond, ...)                \
  do {                                          \
    if (!(cond)) {                              \
      sdb_set_error(__VA_ARGS__);               \
      return SDB_ERR_ARG;                       \
    }                                           \
  } while (0)

#define SDB_CHECK_LAUNCH(name)                                               \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      sdb_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return SDB_ERR_CUDA;                                                   \
    }                                                                        \
  } while (0)

// Counts kernels launched by this library (bench.py reports it as gpu_launches).
extern unsigned long long g_sdb_launch_count;
#define SDB_COUNT_LAUNCH() (++g_sdb_launch_count)

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// A training step is ~700 dependent launches of 5-100 us each. With programmatic stream serialisation the CTAs of
// kernel i+1 are scheduled (and run their set-up: barrier / tensor-memory allocation, descriptor prefetch, index
// arithmetic) while kernel i drains; `griddepcontrol.wait` then blocks until kernel i has completed and its writes are
// visible. Every kernel launched through sdb_launch() therefore calls pdl_wait() before its first access to global
// memory and pdl_launch_dependents() as early as it can. Both instructions are no-ops in a kernel launched the ordinary
// way; SDB_PDL=0 launches everything the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}
bool sdb_pdl_enabled();  // capi.cu

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t sdb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = sdb_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// fp32 -> nearest tf32 (10 mantissa bits), returned as fp32 bits. tcgen05 kind::tf32 IGNORES the low 13 mantissa bits of
// its fp32 operands (truncation: a bias of ~3e-4 per operand that compounds through chained GEMMs); every tensor that
// only feeds GEMMs is therefore written already rounded by its producer.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

static constexpr int kNumSMs = 148;
static constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// Inclusive warp prefix sum.
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x / (1.f + __expf(-x)); }
