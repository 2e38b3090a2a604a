// Field backward of the taped NeRF renderer with the 32-64-{1,3} MLP contractions on the tensor cores
// (mma.sync.m16n8k8 tf32, fp32 accumulation) -- same interface, inputs and outputs as render_field_bwd_kernel
// (render_bwd2.cu), which keeps the fp32 CUDA-core version as the cross-check (SDB_FB_TC=0). The contraction core and the
// lane-pair scatter are shared with the amortized field (field_bwd_tc.cuh).
//
// Why: ncu on the fp32 kernel (profiles/r1g_ncu_full_render_bwd.csv) shows 255 registers, 12.5 % active warps and the
// FMA pipe at 41 %; counting LDS.128 per FFMA shows the register-tiled contractions are shared-memory-bandwidth bound
// (dE = dH W1: 3 LDS.128 per 32 FFMA; dW1 += dH^T E: 8 per 64). 12.3 k FMA per kept sample x 4.8 M samples is 2.9 ms of
// the 4.3 ms kernel. An m16n8k8 fragment needs 6 LDS.32 per 1024 FMA.
// tf32 rounds operands to 11 significant bits (2^-12 relative); products are exact in fp32 and accumulate in fp32. The
// gradients are smooth in those operands, so the error stays two orders below the 5e-3 parity bound (measured in
// tests/test_render_gpu.py against the oracle and against the fp32 kernel).
#include <cstdlib>

#include "field_bwd_tc.cuh"

namespace {

using namespace fbtc;

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

    contract_tile(s, buf, warp, lane, true, true, accw1, accw2);

    // the encoding tiles are free (every thread passed the last barrier): stream the next tile in behind the scatter
    if (tile + (int)gridDim.x < n_tiles) issue_tile(tile + gridDim.x, buf ^ 1);

    if (scatter_on)  // scatter_on doubles as the level mask (diagnostics: SDB_FB_LEVELS)
      scatter_tile(f.grid, g_table, s, buf, warp, lane, n - base, scatter_on, rep, kRepLevels, rep_entries, n_rep);
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
