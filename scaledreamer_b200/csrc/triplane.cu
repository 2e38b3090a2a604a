// Triplane feature lookup for the Triplane-Transformer generator (amortized path, C5): bilinear samples of three
// axis-aligned feature planes, concatenated plane-major into a 96-wide encoding, forward and backward.
// Replaces sample_from_planes / project_onto_planes (custom/amortized/models/geometry/utils.py:67-97: inverse plane
// matrices + torch.bmm + F.grid_sample(bilinear, zeros padding, align_corners=False)) and its autograd.
//
// Layout: planes are taken CHANNELS-LAST, [B, 3, H, W, C] (the generator's [B, 3, C, H, W] output is permuted once per
// step), so each bilinear tap is one contiguous C*4-byte read. Forward: one thread per (point, plane, channel quad): the
// eight threads of a tap read 128 contiguous bytes and the 24 threads of a point write 384 contiguous bytes. Backward:
// see triplane_bwd_kernel (runs of points on one texel merged in registers, red.v4).
#include "../../include/sdb200.h"
#include "common.cuh"

namespace {

// plane 0 samples (x, y), plane 1 (x, z), plane 2 (z, y): (grid x -> W, grid y -> H)
__device__ __forceinline__ void plane_uv(int plane, float x, float y, float z, float* u, float* v) {
  if (plane == 0) { *u = x; *v = y; }
  else if (plane == 1) { *u = x; *v = z; }
  else { *u = z; *v = y; }
}

struct Taps {
  int x0, y0;
  float wx, wy;
};
__device__ __forceinline__ Taps make_taps(float u, float v, int W, int H) {
  Taps t;
  const float px = ((u + 1.f) * (float)W - 1.f) * 0.5f, py = ((v + 1.f) * (float)H - 1.f) * 0.5f;
  const float fx = floorf(px), fy = floorf(py);
  t.x0 = (int)fx;
  t.y0 = (int)fy;
  t.wx = px - fx;
  t.wy = py - fy;
  return t;
}

__global__ void __launch_bounds__(256)
triplane_fwd_kernel(const float* __restrict__ planes, const float* __restrict__ pts, int B, int N, int H, int W, int C4,
                    float* __restrict__ enc) {
  const long long total = (long long)B * N * 3 * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    long long r = i / C4;
    const int plane = (int)(r % 3);
    r /= 3;  // point index b*N + n
    const int b = (int)(r / N);
    const float x = pts[r * 3], y = pts[r * 3 + 1], z = pts[r * 3 + 2];
    float u, v;
    plane_uv(plane, x, y, z, &u, &v);
    const Taps t = make_taps(u, v, W, H);
    const float4* P = reinterpret_cast<const float4*>(planes) + ((long long)(b * 3 + plane) * H * W) * C4 + q;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
      if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;  // zeros padding
      const float w = ((k & 1) ? t.wx : 1.f - t.wx) * ((k >> 1) ? t.wy : 1.f - t.wy);
      const float4 f = __ldg(P + ((long long)yy * W + xx) * C4);
      acc.x = fmaf(w, f.x, acc.x);
      acc.y = fmaf(w, f.y, acc.y);
      acc.z = fmaf(w, f.z, acc.z);
      acc.w = fmaf(w, f.w, acc.w);
    }
    reinterpret_cast<float4*>(enc)[i] = acc;  // [B, N, 3, C] == [B, N, 3*C] plane-major
  }
}

// Backward: thread = (segment of kSeg consecutive points, plane, channel quad). Consecutive points are consecutive samples
// of one ray in every caller (VolSDF renderer: [rays, samples, 3]); on a 64 x 64 plane they move about a third of a texel
// per step, so the four bilinear taps of a run of points land on the same texels: their contributions are summed in
// registers and sent when the texel changes, as ONE red.global.add.v4.f32 per tap (16 bytes per lane, eight lanes = the
// 128 contiguous bytes of a texel). The first version sent 4 taps x 2 red.v2 per point and quad: 192 lane-ops per point,
// all 236 M points of a C5 step onto 393 k plane entries (217 ms). Incoherent points cost what they cost before.
constexpr int kSeg = 16;

__device__ __forceinline__ void red_add_v4(float* dst, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256)
triplane_bwd_kernel(const float* __restrict__ d_enc, const float* __restrict__ pts, int B, int N, int H, int W, int C4,
                    float* __restrict__ d_planes) {
  const int segs_per_b = (N + kSeg - 1) / kSeg;  // segments never straddle prompts
  const long long total = (long long)B * segs_per_b * 3 * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    long long r = i / C4;
    const int plane = (int)(r % 3);
    r /= 3;  // segment index b * segs_per_b + s
    const int b = (int)(r / segs_per_b);
    const int n0 = (int)(r - (long long)b * segs_per_b) * kSeg, n1 = min(N, n0 + kSeg);
    float* P = d_planes + (((long long)(b * 3 + plane) * H * W) * C4 + q) * 4;
    float4 acc[4];
    int cx = 0, cy = 0;
    bool open = false;
    auto flush = [&]() {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int xx = cx + (k & 1), yy = cy + (k >> 1);
        if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;  // zeros padding
        red_add_v4(P + ((long long)yy * W + xx) * C4 * 4, acc[k]);
      }
    };
    for (int n = n0; n < n1; ++n) {
      const long long pt = (long long)b * N + n;
      const float4 g = reinterpret_cast<const float4*>(d_enc)[(pt * 3 + plane) * C4 + q];
      if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
      float u, v;
      plane_uv(plane, pts[pt * 3], pts[pt * 3 + 1], pts[pt * 3 + 2], &u, &v);
      const Taps t = make_taps(u, v, W, H);
      if (open && (t.x0 != cx || t.y0 != cy)) {
        flush();
        open = false;
      }
      if (!open) {
        cx = t.x0, cy = t.y0;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        open = true;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float w = ((k & 1) ? t.wx : 1.f - t.wx) * ((k >> 1) ? t.wy : 1.f - t.wy);
        acc[k].x = fmaf(w, g.x, acc[k].x);
        acc[k].y = fmaf(w, g.y, acc[k].y);
        acc[k].z = fmaf(w, g.z, acc[k].z);
        acc[k].w = fmaf(w, g.w, acc[k].w);
      }
    }
    if (open) flush();
  }
}

}  // namespace

extern "C" {

int sdb_triplane_sample_forward(const float* planes_cl, const float* points, int n_prompts, int n_points, int height,
                                int width, int channels, float* enc, void* stream) {
  SDB_CHECK_ARG(planes_cl && points && enc && n_prompts > 0 && n_points >= 0 && height > 0 && width > 0,
                "triplane_sample_forward: bad arguments");
  SDB_CHECK_ARG(channels > 0 && channels % 4 == 0, "triplane_sample: channels must be a multiple of 4");
  if (n_points == 0) return SDB_OK;
  const long long total = (long long)n_prompts * n_points * 3 * (channels / 4);
  const int grid = (int)((total + 255) / 256 < (long long)kNumSMs * 16 ? (total + 255) / 256 : (long long)kNumSMs * 16);
  triplane_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(planes_cl, points, n_prompts, n_points, height, width,
                                                              channels / 4, enc);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("triplane_sample_forward");
  return SDB_OK;
}

int sdb_triplane_sample_backward(const float* d_enc, const float* points, int n_prompts, int n_points, int height,
                                 int width, int channels, float* d_planes_cl, void* stream) {
  SDB_CHECK_ARG(d_enc && points && d_planes_cl && n_prompts > 0 && n_points >= 0 && height > 0 && width > 0,
                "triplane_sample_backward: bad arguments");
  SDB_CHECK_ARG(channels > 0 && channels % 4 == 0, "triplane_sample: channels must be a multiple of 4");
  if (n_points == 0) return SDB_OK;
  const long long total = (long long)n_prompts * ((n_points + kSeg - 1) / kSeg) * 3 * (channels / 4);
  const int grid = (int)((total + 255) / 256 < (long long)kNumSMs * 16 ? (total + 255) / 256 : (long long)kNumSMs * 16);
  triplane_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_enc, points, n_prompts, n_points, height, width,
                                                              channels / 4, d_planes_cl);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("triplane_sample_backward");
  return SDB_OK;
}

}  // extern "C"
