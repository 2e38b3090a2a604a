// Asynchronous-Score-Distillation glue kernels (include/sdb200_asd.h). HBM-streaming / tiny reductions.
// Reference: threestudio/models/guidance/stable_diffusion_asd_guidance.py:196-316, 333-428;
// threestudio/models/guidance/mvdream_asd_guidance.py:141-304; threestudio/models/prompt_processors/base.py:53-167;
// threestudio/utils/ops.py:493-511; extern/mvdream/ldm/modules/distributions/distributions.py:24-37.
#include <cmath>

#include "../../include/sdb200_asd.h"
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------- bilinear resize
__device__ __forceinline__ void src_index(int dst, float ratio, int in_size, int* i0, int* i1, float* lam) {
  // ATen area_pixel_compute_source_index, align_corners=False: src = ratio*(dst+0.5)-0.5, clamped at 0
  float s = ratio * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  const int a = (int)s;
  *i0 = a;
  *i1 = a + (a < in_size - 1 ? 1 : 0);
  *lam = s - (float)a;
}

__global__ void resize_fwd_kernel(const float* __restrict__ x, int B, int h, int w, int C, float* __restrict__ y, int H,
                                  int W, float scale, float shift) {
  const long long total = (long long)B * H * W * C;
  const float ry = (float)h / (float)H, rx = (float)w / (float)W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ox = (int)(r % W);
    r /= W;
    const int oy = (int)(r % H);
    const int b = (int)(r / H);
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(oy, ry, h, &y0, &y1, &ly);
    src_index(ox, rx, w, &x0, &x1, &lx);
    const float* p = x + (long long)b * h * w * C + c;
    const float v00 = p[((long long)y0 * w + x0) * C], v01 = p[((long long)y0 * w + x1) * C];
    const float v10 = p[((long long)y1 * w + x0) * C], v11 = p[((long long)y1 * w + x1) * C];
    const float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    y[i] = fmaf(v, scale, shift);
  }
}

__global__ void resize_bwd_kernel(const float* __restrict__ dy, int B, int h, int w, int C, float* __restrict__ dx,
                                  int H, int W, float scale) {
  const long long total = (long long)B * H * W * C;
  const float ry = (float)h / (float)H, rx = (float)w / (float)W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ox = (int)(r % W);
    r /= W;
    const int oy = (int)(r % H);
    const int b = (int)(r / H);
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(oy, ry, h, &y0, &y1, &ly);
    src_index(ox, rx, w, &x0, &x1, &lx);
    const float g = dy[i] * scale;
    float* p = dx + (long long)b * h * w * C + c;
    atomicAdd(p + ((long long)y0 * w + x0) * C, g * (1.f - ly) * (1.f - lx));
    atomicAdd(p + ((long long)y0 * w + x1) * C, g * (1.f - ly) * lx);
    atomicAdd(p + ((long long)y1 * w + x0) * C, g * ly * (1.f - lx));
    atomicAdd(p + ((long long)y1 * w + x1) * C, g * ly * lx);
  }
}

// ------------------------------------------------------------------------------------------- text embeddings
struct PromptPlan {  // per sample, computed by thread 0 of each block row
  int dir;           // 0 side, 1 front, 2 back, 3 overhead
  int pos_a, pos_b;  // positive embedding = wa*emb[pos_a] + (1-wa)*emb[pos_b]
  float wa;
  int neg0, neg1;    // indices into emb_vd, or -1 = uncond[dir]
  float w0, w1;
};

__device__ __forceinline__ float shift_az(float a) {  // (a + 180) % 360 - 180 with python's modulo
  float m = fmodf(a + 180.f, 360.f);
  if (m < 0.f) m += 360.f;
  return m - 180.f;
}

__device__ PromptPlan make_plan(const sdb_prompt_cfg& c, float ele, float azi_raw) {
  PromptPlan p;
  const float azs = shift_az(azi_raw);
  int dir = 0;
  if (c.view_dependent) {
    // later directions overwrite earlier ones (prompt_processors/base.py:66-69): side, front, back, overhead
    if (azs > -c.front_threshold && azs < c.front_threshold) dir = 1;
    if (azs > 180.f - c.back_threshold || azs < -180.f + c.back_threshold) dir = 2;
    if (ele > c.overhead_threshold) dir = 3;
  }
  p.dir = dir;
  p.pos_a = p.pos_b = dir;
  p.wa = 1.f;
  p.neg0 = p.neg1 = -1;
  p.w0 = p.w1 = 0.f;
  if (c.perp_neg) {
    const float aa = fabsf(azs);
    if (dir == 3) {
      p.pos_a = p.pos_b = 3;
    } else if (aa < 90.f) {
      const float r = 1.f - aa / 90.f;  // front-side interpolation
      p.pos_a = 1, p.pos_b = 0, p.wa = r;
      p.neg0 = 1, p.neg1 = 0;
      p.w0 = -(c.f_fs[0] * expf(-c.f_fs[1] * r) + c.f_fs[2]);
      p.w1 = -(c.f_sf[0] * expf(-c.f_sf[1] * (1.f - r)) + c.f_sf[2]);
    } else {
      const float r = 2.f - aa / 90.f;  // side-back interpolation
      p.pos_a = 0, p.pos_b = 2, p.wa = r;
      p.neg0 = 0, p.neg1 = 1;
      p.w0 = -(c.f_sb[0] * expf(-c.f_sb[1] * r) + c.f_sb[2]);
      p.w1 = -(c.f_fsb[0] * expf(-c.f_fsb[1] * r) + c.f_fsb[2]);
    }
  }
  return p;
}

// grid: (blocks over tokens*dim/8, B). Each thread moves 8 halfs of every output row of its sample.
__global__ void text_embeddings_kernel(const __grid_constant__ sdb_prompt_cfg c, const __half* __restrict__ emb,
                                       const __half* __restrict__ unc, const float* __restrict__ elevation,
                                       const float* __restrict__ azimuth, int B, long long row8,
                                       __half* __restrict__ ctx, float* __restrict__ neg_w,
                                       const int* __restrict__ prompt_idx, long long table_stride8) {
  const int b = blockIdx.y;
  const PromptPlan p = make_plan(c, elevation[b], azimuth[b]);
  if (neg_w && blockIdx.x == 0 && threadIdx.x == 0) {
    neg_w[2 * b] = p.w0 * c.neg_scale;
    neg_w[2 * b + 1] = p.w1 * c.neg_scale;
  }
  // multi-prompt: every sample reads the table of its own prompt out of a stacked [P, n_dir, tokens, dim] tensor
  const uint4* E = reinterpret_cast<const uint4*>(emb) + (prompt_idx ? (long long)prompt_idx[b] * table_stride8 : 0LL);
  const uint4* U = reinterpret_cast<const uint4*>(unc);
  uint4* O = reinterpret_cast<uint4*>(ctx);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row8;
       i += (long long)gridDim.x * blockDim.x) {
    uint4 pos;
    if (p.pos_a == p.pos_b) {
      pos = E[p.pos_a * row8 + i];
    } else {
      const uint4 a = E[p.pos_a * row8 + i], bb = E[p.pos_b * row8 + i];
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
      const __half2* bh = reinterpret_cast<const __half2*>(&bb);
      __half2* ph = reinterpret_cast<__half2*>(&pos);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fa = __half22float2(ah[e]), fb = __half22float2(bh[e]);
        ph[e] = __floats2half2_rn(p.wa * fa.x + (1.f - p.wa) * fb.x, p.wa * fa.y + (1.f - p.wa) * fb.y);
      }
    }
    const uint4 un = U[p.dir * row8 + i];
    O[(0LL * B + b) * row8 + i] = pos;
    O[(1LL * B + b) * row8 + i] = un;
    if (c.perp_neg) {
      O[(2LL * B + 2 * b) * row8 + i] = p.neg0 >= 0 ? E[p.neg0 * row8 + i] : un;
      O[(2LL * B + 2 * b + 1) * row8 + i] = p.neg1 >= 0 ? E[p.neg1 * row8 + i] : un;
      O[(4LL * B + b) * row8 + i] = pos;
    } else {
      O[(2LL * B + b) * row8 + i] = pos;
    }
  }
}

// ------------------------------------------------------------------------------------------- prologue
__device__ __forceinline__ void moments_of(const float* __restrict__ h8, const float* __restrict__ qw,
                                           const float* __restrict__ qb, float* m) {
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    float a = qb[o];
#pragma unroll
    for (int i = 0; i < 8; ++i) a = fmaf(qw[o * 8 + i], h8[i], a);
    m[o] = a;
  }
}

__global__ void asd_prologue_kernel(const float* __restrict__ h, const float* __restrict__ qw,
                                    const float* __restrict__ qb, const float* __restrict__ eps_post,
                                    const float* __restrict__ noise, const int* __restrict__ t,
                                    const int* __restrict__ t_plus, const float* __restrict__ ac, float sf, int B,
                                    int HW, int R, float* __restrict__ latents, __half* __restrict__ ux,
                                    float* __restrict__ ut) {
  __shared__ float sw[64], sb[8];
  if (threadIdx.x < 64) sw[threadIdx.x] = qw[threadIdx.x];
  if (threadIdx.x < 8) sb[threadIdx.x] = qb[threadIdx.x];
  __syncthreads();
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    float hv[8], m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) hv[k] = h[i * 8 + k];
    moments_of(hv, sw, sb, m);
    const float a1 = ac[t[b]], a2 = ac[t_plus[b]];
    const float s1 = sqrtf(a1), n1 = sqrtf(1.f - a1), s2 = sqrtf(a2), n2 = sqrtf(1.f - a2);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float lv = fminf(fmaxf(m[4 + c], -30.f), 20.f);
      const float z = (m[c] + expf(0.5f * lv) * eps_post[i * 4 + c]) * sf;
      latents[i * 4 + c] = z;
      const float nz = noise[i * 4 + c];
      const __half x1 = __float2half_rn(s1 * z + n1 * nz), x2 = __float2half_rn(s2 * z + n2 * nz);
      for (int r = 0; r < R; ++r) ux[((long long)r * B * HW + i) * 4 + c] = x1;
      ux[((long long)R * B * HW + i) * 4 + c] = x2;
    }
    if (i % HW == 0) {
      for (int r = 0; r < R; ++r) ut[r * B + b] = (float)t[b];
      ut[R * B + b] = (float)t_plus[b];
    }
  }
}

// ------------------------------------------------------------------------------------------- epilogue
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  return s;
}

// one block per sample
__global__ void __launch_bounds__(1024)
asd_epilogue_kernel(const float* __restrict__ eps, const float* __restrict__ h, const float* __restrict__ qw,
                    const float* __restrict__ qb, const float* __restrict__ eps_post, const int* __restrict__ t,
                    const float* __restrict__ ac, const float* __restrict__ neg_w, float gs, int weighting, float clip,
                    float sf, float loss_scale, int B, int HW, int R, float* __restrict__ grad,
                    float* __restrict__ d_h, float* __restrict__ loss, float* __restrict__ grad_norm_sq) {
  __shared__ float red[32];
  __shared__ float sw[64], sb[8];
  if (threadIdx.x < 64) sw[threadIdx.x] = qw[threadIdx.x];
  if (threadIdx.x < 8) sb[threadIdx.x] = qb[threadIdx.x];
  const int b = blockIdx.x;
  const int n = HW * 4;
  const float* e_c = eps + (long long)(0 * B + b) * n;
  const float* e_u = eps + (long long)(1 * B + b) * n;
  const float* e_n0 = neg_w ? eps + (long long)(2 * B + 2 * b) * n : nullptr;
  const float* e_n1 = neg_w ? eps + (long long)(2 * B + 2 * b + 1) * n : nullptr;
  const float* e_s = eps + (long long)(R * B + b) * n;
  float k0 = 0.f, k1 = 0.f;  // perp coefficients <v_i, p> / max(<p,p>, 1e-6)
  if (neg_w) {
    float d0 = 0.f, d1 = 0.f, pp = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float u = e_u[i], p = e_c[i] - u;
      d0 = fmaf(e_n0[i] - u, p, d0);
      d1 = fmaf(e_n1[i] - u, p, d1);
      pp = fmaf(p, p, pp);
    }
    d0 = block_sum(d0, red);
    d1 = block_sum(d1, red);
    pp = block_sum(pp, red);
    const float den = fmaxf(pp, 1e-6f);
    k0 = d0 / den;
    k1 = d1 / den;
  }
  __syncthreads();
  const float a = ac[t[b]];
  const float w = weighting == 0 ? (1.f - a) : (weighting == 1 ? 1.f : sqrtf(a) * (1.f - a));
  const float w0 = neg_w ? neg_w[2 * b] : 0.f, w1 = neg_w ? neg_w[2 * b + 1] : 0.f;
  float lsum = 0.f;
  for (int px = threadIdx.x; px < HW; px += blockDim.x) {
    float hv[8], m[8], dm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) hv[k] = h[((long long)b * HW + px) * 8 + k];
    moments_of(hv, sw, sb, m);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = px * 4 + c;
      const float u = e_u[i], p = e_c[i] - u;
      float acc = p;
      if (neg_w) {
        const float v0 = e_n0[i] - u, v1 = e_n1[i] - u;
        acc += w0 * (v0 - k0 * p) + w1 * (v1 - k1 * p);
      }
      float g = w * (u + gs * acc - e_s[i]);
      if (isnan(g)) g = 0.f;                       // torch.nan_to_num defaults
      else if (isinf(g)) g = g > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
      if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
      grad[(long long)b * n + i] = g;
      lsum = fmaf(g, g, lsum);
      // backward of z = sf * (mean + exp(0.5*clamp(logvar)) * eps): dz = g / B
      const float dz = g / (float)B * loss_scale;
      const float lv = m[4 + c];
      dm[c] = sf * dz;
      dm[4 + c] = (lv > -30.f && lv < 20.f) ? sf * dz * eps_post[((long long)b * HW + px) * 4 + c] * 0.5f * expf(0.5f * lv)
                                            : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int o = 0; o < 8; ++o) acc = fmaf(sw[o * 8 + j], dm[o], acc);
      d_h[((long long)b * HW + px) * 8 + j] = acc;
    }
  }
  lsum = block_sum(lsum, red);
  if (threadIdx.x == 0) {
    atomicAdd(loss, 0.5f * lsum / (float)B);
    atomicAdd(grad_norm_sq, lsum);
  }
}

__global__ void sqrt_kernel(float* v) { *v = sqrtf(*v); }

__global__ void t_plus_kernel(const int* __restrict__ t, const float* __restrict__ u, int B, float ratio, int min_step,
                              int T, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float tp = ratio * (float)(t[i] - min_step);
  tp = fminf(fmaxf(tp, 0.f), (float)(T - t[i] - 1));
  if (u) tp *= u[i];
  int r = t[i] + (int)tp;  // .to(torch.long) truncates
  r = min(max(r, 1), T - 1);
  out[i] = r;
}

inline int ew_grid(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int sdb_resize_bilinear_forward(const float* x, int batch, int h, int w, int c, float* y, int H, int W, float scale,
                                float shift, void* stream) {
  SDB_CHECK_ARG(x && y && batch > 0 && h > 0 && w > 0 && c > 0 && H > 0 && W > 0, "resize_forward: bad arguments");
  resize_fwd_kernel<<<ew_grid((long long)batch * H * W * c), 256, 0, (cudaStream_t)stream>>>(x, batch, h, w, c, y, H, W,
                                                                                             scale, shift);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("resize_bilinear_forward");
  return SDB_OK;
}

int sdb_resize_bilinear_backward(const float* d_y, int batch, int h, int w, int c, float* d_x, int H, int W,
                                 float scale, void* stream) {
  SDB_CHECK_ARG(d_y && d_x && batch > 0 && h > 0 && w > 0 && c > 0 && H > 0 && W > 0, "resize_backward: bad arguments");
  cudaMemsetAsync(d_x, 0, sizeof(float) * (size_t)batch * h * w * c, (cudaStream_t)stream);
  resize_bwd_kernel<<<ew_grid((long long)batch * H * W * c), 256, 0, (cudaStream_t)stream>>>(d_y, batch, h, w, c, d_x, H,
                                                                                             W, scale);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("resize_bilinear_backward");
  return SDB_OK;
}

int sdb_asd_text_embeddings(const sdb_prompt_cfg* cfg, const void* emb_vd, const void* uncond_vd,
                            const float* elevation, const float* azimuth, int batch, int tokens, int dim, void* ctx,
                            float* neg_weights, void* stream) {
  SDB_CHECK_ARG(cfg && emb_vd && uncond_vd && elevation && azimuth && ctx && batch > 0, "text_embeddings: bad arguments");
  SDB_CHECK_ARG(((long long)tokens * dim) % 8 == 0, "text_embeddings: tokens*dim must be a multiple of 8");
  SDB_CHECK_ARG(!cfg->perp_neg || cfg->view_dependent, "Perp-Neg only works with view-dependent prompting");
  const long long row8 = (long long)tokens * dim / 8;
  dim3 grid((unsigned)((row8 + 255) / 256), batch);
  text_embeddings_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      *cfg, reinterpret_cast<const __half*>(emb_vd), reinterpret_cast<const __half*>(uncond_vd), elevation, azimuth,
      batch, row8, reinterpret_cast<__half*>(ctx), cfg->perp_neg ? neg_weights : nullptr, nullptr, 0);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("asd_text_embeddings");
  return SDB_OK;
}

int sdb_asd_text_embeddings_multi(const sdb_prompt_cfg* cfg, const void* emb_tables, const void* uncond_vd,
                                  const int* prompt_idx, int n_prompts, const float* elevation, const float* azimuth,
                                  int batch, int tokens, int dim, void* ctx, float* neg_weights, void* stream) {
  SDB_CHECK_ARG(cfg && emb_tables && uncond_vd && prompt_idx && elevation && azimuth && ctx && batch > 0 && n_prompts > 0,
                "text_embeddings_multi: bad arguments");
  SDB_CHECK_ARG(((long long)tokens * dim) % 8 == 0, "text_embeddings: tokens*dim must be a multiple of 8");
  SDB_CHECK_ARG(!cfg->perp_neg || cfg->view_dependent, "Perp-Neg only works with view-dependent prompting");
  const long long row8 = (long long)tokens * dim / 8;
  const long long stride8 = (cfg->view_dependent ? 4 : 1) * row8;
  dim3 grid((unsigned)((row8 + 255) / 256), batch);
  text_embeddings_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      *cfg, reinterpret_cast<const __half*>(emb_tables), reinterpret_cast<const __half*>(uncond_vd), elevation, azimuth,
      batch, row8, reinterpret_cast<__half*>(ctx), cfg->perp_neg ? neg_weights : nullptr, prompt_idx, stride8);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("asd_text_embeddings_multi");
  return SDB_OK;
}

int sdb_asd_prologue(const float* h, const float* quant_w, const float* quant_b, const float* eps_post,
                     const float* noise, const int* t, const int* t_plus, const float* alphas_cumprod,
                     float scaling_factor, int batch, int hw, int num_repeats, float* latents, void* unet_x,
                     float* unet_t, void* stream) {
  SDB_CHECK_ARG(h && quant_w && quant_b && eps_post && noise && t && t_plus && alphas_cumprod && latents && unet_x &&
                    unet_t && batch > 0 && hw > 0 && num_repeats >= 1,
                "asd_prologue: bad arguments");
  asd_prologue_kernel<<<ew_grid((long long)batch * hw), 256, 0, (cudaStream_t)stream>>>(
      h, quant_w, quant_b, eps_post, noise, t, t_plus, alphas_cumprod, scaling_factor, batch, hw, num_repeats, latents,
      reinterpret_cast<__half*>(unet_x), unet_t);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("asd_prologue");
  return SDB_OK;
}

int sdb_asd_epilogue(const float* eps, const float* h, const float* quant_w, const float* quant_b,
                     const float* eps_post, const int* t, const float* alphas_cumprod, const float* neg_weights,
                     float guidance_scale, int weighting, float grad_clip, float scaling_factor, float loss_scale,
                     int batch, int hw, int num_repeats, float* grad, float* d_h, float* loss, float* grad_norm,
                     void* stream) {
  SDB_CHECK_ARG(eps && h && quant_w && quant_b && eps_post && t && alphas_cumprod && grad && d_h && loss && grad_norm &&
                    batch > 0 && hw > 0,
                "asd_epilogue: bad arguments");
  SDB_CHECK_ARG(!neg_weights || num_repeats == 4, "asd_epilogue: Perp-Neg needs the 5B batch layout (num_repeats 4)");
  SDB_CHECK_ARG(weighting >= 0 && weighting <= 2, "asd_epilogue: unknown weighting strategy");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(loss, 0, sizeof(float), s);
  cudaMemsetAsync(grad_norm, 0, sizeof(float), s);
  asd_epilogue_kernel<<<batch, 1024, 0, s>>>(eps, h, quant_w, quant_b, eps_post, t, alphas_cumprod, neg_weights,
                                             guidance_scale, weighting, grad_clip, scaling_factor, loss_scale, batch, hw,
                                             num_repeats, grad, d_h, loss, grad_norm);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("asd_epilogue");
  sqrt_kernel<<<1, 1, 0, s>>>(grad_norm);
  SDB_COUNT_LAUNCH();
  return SDB_OK;
}

int sdb_asd_t_plus(const int* t, const float* u, int batch, float plus_ratio, int min_step, int num_train_timesteps,
                   int* t_plus, void* stream) {
  SDB_CHECK_ARG(t && t_plus && batch > 0 && plus_ratio >= 0.f, "asd_t_plus: bad arguments");
  t_plus_kernel<<<(batch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(t, u, batch, plus_ratio, min_step,
                                                                   num_train_timesteps, t_plus);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("asd_t_plus");
  return SDB_OK;
}

}  // extern "C"
