// How does the cost of scattered fp32 reductions (red.global.add.v2.f32) scale on B200: with the number of warp-level
// INSTRUCTIONS, with the number of active LANES, or with the number of distinct 32-byte SECTORS touched?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/red_probe.cu -o /tmp/red_probe && /tmp/red_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// mode 0: every active lane a random entry; mode 1: active lanes in groups of 4 consecutive entries (one 32-B sector
// per group of 4... entries are 8 B, so 4 entries = one sector); mode 2: all active lanes the same random entry;
// mode 3: lane pairs on adjacent entries e, e + 1 (the two x-corners of a grid cell: same sector three times out of four)
template <int VEC>
__global__ void red_kernel(float2* table, uint32_t n_entries, int iters, int active, int mode) {
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (lane >= active) return;
  for (int i = 0; i < iters; ++i) {
    uint32_t base = hash32(gw * 9781u + i * 6271u);
    uint32_t idx;
    if (mode == 0) idx = hash32(base + lane * 7919u) % n_entries;
    else if (mode == 1) idx = ((hash32(base + (lane >> 2) * 7919u) % (n_entries / 4)) * 4 + (lane & 3));
    else if (mode == 3) idx = (hash32(base + (lane >> 1) * 7919u) % (n_entries - 1)) + (lane & 1);  // lane pairs on entries e, e + 1
    else idx = base % n_entries;
    if (VEC == 2) atomicAdd(table + idx, make_float2(1.f, 2.f));
    else atomicAdd(reinterpret_cast<float*>(table + idx), 1.f);
  }
}

int main() {
  const uint32_t n_entries = 6299960;  // the C2 hash table
  float2* table;
  cudaMalloc(&table, (size_t)n_entries * 8);
  cudaMemset(table, 0, (size_t)n_entries * 8);
  const int iters = 2000;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  printf("warps/SM  active  mode  ns per warp-instruction per SM   G lane-ops/s\n");
  for (int wps : {8, 32})
    for (int mode : {0, 1, 3, 2})
      for (int active : {4, 8, 16, 32}) {
        const int blocks = 148 * wps / 4;  // 4 warps per block
        red_kernel<2><<<blocks, 128>>>(table, n_entries, 50, active, mode);
        cudaEventRecord(a);
        red_kernel<2><<<blocks, 128>>>(table, n_entries, iters, active, mode);
        cudaEventRecord(b);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double instr_per_sm = (double)wps * iters;
        printf("%8d  %6d  %4d  %10.1f ns (%5.0f clk)   %8.1f\n", wps, active, mode, ms * 1e6 / instr_per_sm,
               ms * 1e6 / instr_per_sm * 1.965, (double)148 * wps * iters * active / (ms * 1e-3) / 1e9);
      }
  return 0;
}
