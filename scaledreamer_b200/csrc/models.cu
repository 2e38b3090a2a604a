// Executors of the frozen latent-diffusion UNet and KL-VAE encoder (see models.h).
//
// Architecture restated from the reference's vendored LDM (paths relative to /root/reference/extern/mvdream):
//   UNetModel / MultiViewUNetModel  ldm/modules/diffusionmodules/openaimodel.py:422-808, 811-1213
//   ResBlock                        openaimodel.py:163-275          Downsample / Upsample  openaimodel.py:91-160
//   SpatialTransformer(3D)          ldm/modules/attention.py:289-412  BasicTransformerBlock(3D)  :245-354
//   CrossAttention / GEGLU          attention.py:49-76, 143-194
//   VAE Encoder / ResnetBlock / AttnBlock / Downsample   ldm/modules/diffusionmodules/model.py:452-543, 90-203, 67-87
// which is the same network as diffusers' UNet2DConditionModel / AutoencoderKL for stable-diffusion-2-1-base
// (threestudio/models/guidance/stable_diffusion_asd_guidance.py:68-104).
#include "models.h"

#include <cstring>

using namespace dense;

namespace nn {

static inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

// --------------------------------------------------------------------------------------------------- Net
int Net::bind(void* weights, void* work) {
  if (!weights || !work) {
    sdb_set_error("net: bind() needs both arenas");
    return SDB_ERR_ARG;
  }
  wbase_ = reinterpret_cast<char*>(weights);
  kbase_ = reinterpret_cast<char*>(work);
  dry_ = false;
  woff_ = 0;
  koff_ = scratch_max_;  // scratch region sits at the start of the work arena
  scratch_off_ = 0;
  fwd_.clear();
  bwd_.clear();
  post_.clear();
  build_rc_ = 0;
  int rc = build();
  if (rc) return rc;
  return build_rc_;
}

__half* Net::param(const std::string& name, int ndim, int d0, int d1, int d2, int d3) {
  const long long numel = (long long)d0 * d1 * d2 * d3;
  const long long off = woff_;
  woff_ += align_up(numel * 2, 256);
  if (dry_) {
    ParamInfo pi;
    pi.name = name;
    pi.ndim = ndim;
    pi.shape[0] = d0, pi.shape[1] = d1, pi.shape[2] = d2, pi.shape[3] = d3;
    pi.numel = numel;
    pi.ptr = nullptr;
    pi.loaded = false;
    index_[name] = (int)params_.size();
    params_.push_back(pi);
    return nullptr;
  }
  ParamInfo& pi = params_[index_.at(name)];
  pi.ptr = reinterpret_cast<__half*>(wbase_ + off);
  return pi.ptr;
}

__half* Net::derived(long long numel) {
  const long long off = woff_;
  woff_ += align_up(numel * 2, 256);
  return dry_ ? nullptr : reinterpret_cast<__half*>(wbase_ + off);
}

void* Net::work(long long bytes) {
  const long long off = koff_;
  koff_ += align_up(bytes, 256);
  return dry_ ? nullptr : kbase_ + off;
}

T Net::act(int n, int h, int w, int c) {
  T t;
  t.n = n, t.h = h, t.w = w, t.c = c;
  t.p = reinterpret_cast<__half*>(work(t.numel() * 2));
  return t;
}

void* Net::scratch(long long bytes) {
  const long long off = scratch_off_;
  scratch_off_ += align_up(bytes, 256);
  if (scratch_off_ > scratch_max_) {
    if (!dry_) {
      sdb_set_error("net: scratch overflow (internal planning error)");
      build_rc_ = SDB_ERR_ARG;
    } else {
      scratch_max_ = scratch_off_;
    }
  }
  return dry_ ? nullptr : kbase_ + off;
}

int Net::load_param(const char* name, const void* src, long long numel, cudaStream_t s) {
  auto it = index_.find(name);
  if (it == index_.end()) {
    sdb_set_error("net: unknown parameter '%s'", name);
    return SDB_ERR_ARG;
  }
  ParamInfo& pi = params_[it->second];
  if (!pi.ptr) {
    sdb_set_error("net: bind() before load_param()");
    return SDB_ERR_ARG;
  }
  if (numel != pi.numel) {
    sdb_set_error("net: parameter '%s' has %lld elements, got %lld", name, pi.numel, numel);
    return SDB_ERR_ARG;
  }
  cudaError_t e = cudaMemcpyAsync(pi.ptr, src, numel * 2, cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) {
    sdb_set_error("net: copy of '%s' failed: %s", name, cudaGetErrorString(e));
    return SDB_ERR_CUDA;
  }
  pi.loaded = true;
  return SDB_OK;
}

int Net::finalize(cudaStream_t s) {
  for (const ParamInfo& pi : params_)
    if (!pi.loaded) {
      sdb_set_error("net: parameter '%s' was never loaded", pi.name.c_str());
      return SDB_ERR_ARG;
    }
  return run(post_, s);
}

int Net::run(const std::vector<Op>& ops, cudaStream_t s) {
  for (const Op& op : ops) {
    int rc = op(s);
    if (rc) return rc;
  }
  return SDB_OK;
}

T Net::gn(std::vector<Op>* ops, const T& x, const std::string& name, float eps, bool silu, float** stats_out) {
  const __half* g = param(name + ".weight", 1, x.c);
  const __half* b = param(name + ".bias", 1, x.c);
  float* stats = reinterpret_cast<float*>(work(groupnorm_workspace_floats(x.n, x.h * x.w, x.c, 32) * 4));
  T y = act(x.n, x.h, x.w, x.c);
  if (stats_out) *stats_out = stats;
  if (!dry_) {
    const T xx = x;
    ops->push_back([=](cudaStream_t s) {
      return groupnorm_forward(xx.p, g, b, y.p, stats, xx.n, xx.h * xx.w, xx.c, 32, eps, silu ? 1 : 0, s);
    });
  }
  return y;
}

T Net::conv3(std::vector<Op>* ops, const T& x, const __half* w, const __half* bias, int cout, const float* rowbias,
             long long rowbias_ld, const T* residual) {
  T y = act(x.n, x.h, x.w, cout);
  const int sp = gemm_splits((int)x.rows(), cout, 9 * x.c);  // few output tiles, long K: split K over more SMs
  float* ws = sp > 1 ? reinterpret_cast<float*>(work((long long)sp * x.rows() * cout * 4)) : nullptr;
  if (!dry_) {
    Epilogue ep;
    ep.out = y.p;
    ep.ldc = cout;
    ep.bias = bias;
    ep.rowbias = rowbias;
    ep.rows_per_group = x.h * x.w;
    ep.rowbias_ld = rowbias_ld;
    ep.splitk_ws = ws;
    if (residual) {
      ep.residual = residual->p;
      ep.ldr = cout;
    }
    GemmPlan plan;
    if (fail(plan_conv3x3(&plan, x.p, x.n, x.h, x.w, x.c, w, cout, ep))) return y;
    ops->push_back([plan](cudaStream_t s) { return run_gemm(plan, s); });
  }
  return y;
}

T Net::linear(std::vector<Op>* ops, const T& x, const __half* w, const __half* bias, int cout, const T* residual,
              int act_kind) {
  T y = act(x.n, x.h, x.w, cout);
  const int sp = gemm_splits((int)x.rows(), cout, x.c);
  float* ws = sp > 1 ? reinterpret_cast<float*>(work((long long)sp * x.rows() * cout * 4)) : nullptr;
  if (!dry_) {
    Epilogue ep;
    ep.out = y.p;
    ep.ldc = cout;
    ep.bias = bias;
    ep.act = act_kind;
    ep.splitk_ws = ws;
    if (residual) {
      ep.residual = residual->p;
      ep.ldr = cout;
    }
    GemmPlan plan;
    if (fail(plan_gemm(&plan, x.p, x.c, w, x.c, (int)x.rows(), cout, x.c, ep))) return y;
    ops->push_back([plan](cudaStream_t s) { return run_gemm(plan, s); });
  }
  return y;
}

T Net::conv3_s2(std::vector<Op>* ops, const T& x, const __half* w, const __half* bias, int cout, int pad_lo) {
  T y = act(x.n, x.h / 2, x.w / 2, cout);
  reset_scratch();
  __half* col = reinterpret_cast<__half*>(scratch(y.rows() * 9 * x.c * 2));
  const int sp = gemm_splits((int)y.rows(), cout, 9 * x.c);
  float* ws = sp > 1 ? reinterpret_cast<float*>(work((long long)sp * y.rows() * cout * 4)) : nullptr;
  if (!dry_) {
    const T xx = x;
    ops->push_back([=](cudaStream_t s) { return im2col_3x3_s2(xx.p, col, xx.n, xx.h, xx.w, xx.c, pad_lo, s); });
    Epilogue ep;
    ep.out = y.p;
    ep.ldc = cout;
    ep.bias = bias;
    ep.splitk_ws = ws;
    GemmPlan plan;
    if (fail(plan_gemm(&plan, col, 9 * x.c, w, 9 * x.c, (int)y.rows(), cout, 9 * x.c, ep))) return y;
    ops->push_back([plan](cudaStream_t s) { return run_gemm(plan, s); });
  }
  return y;
}

T Net::attention(std::vector<Op>* ops, const __half* q, long long ldq, const __half* k, long long ldk, const __half* v,
                 long long ldv, int B, int heads, int head_dim, int Lq, int Lk, __half** probs_out) {
  const int inner = heads * head_dim;
  if (head_dim == 64 && Lk > 80 && !probs_out) {  // self-attention: fused, the score matrix never reaches HBM
    T o = act(B, 1, Lq, inner);
    if (!dry_) {
      FlashPlan fp;
      if (fail(plan_flash_attn(&fp, q, ldq, k, ldk, v, ldv, B, heads, head_dim, Lq, Lk, o.p, inner,
                               1.f / sqrtf((float)head_dim))))
        return o;
      ops->push_back([fp](cudaStream_t s) { return run_flash_attn(fp, s); });
    }
    return o;
  }
  const long long lds = align_up(Lk, 8);
  const long long sbytes = (long long)B * heads * Lq * lds * 2;
  __half* S = reinterpret_cast<__half*>(probs_out ? work(sbytes) : scratch(sbytes));
  if (probs_out) *probs_out = S;
  T o = act(B, 1, Lq, inner);
  if (!dry_) {
    GemmPlan ps, pa;
    const int fuse = Lk <= 80 ? 1 : 0;  // cross-attention: the whole score row sits in one tile, softmax in the epilogue
    if (fail(plan_attn_scores(&ps, q, ldq, k, ldk, B, heads, head_dim, Lq, Lk, S, lds, 1.f / sqrtf((float)head_dim),
                              fuse)))
      return o;
    if (fail(plan_attn_apply(&pa, S, lds, v, ldv, B, heads, head_dim, Lq, Lk, o.p, inner))) return o;
    const long long rows = (long long)B * heads * Lq;
    ops->push_back([ps](cudaStream_t s) { return run_gemm(ps, s); });
    if (!fuse) ops->push_back([=](cudaStream_t s) { return softmax_rows(S, rows, Lk, lds, s); });
    ops->push_back([pa](cudaStream_t s) { return run_gemm(pa, s); });
  }
  return o;
}

// -------------------------------------------------------------------------------------------------- UNet
UNet::UNet(const UNetCfg& cfg, int batch, int h, int w) : cfg_(cfg), B_(batch), H_(h), W_(w) {
  // two counting passes: the first learns how many timestep-projection rows the ResBlocks need in total
  dry_ = true;
  build();
  emb_rows_ = emb_cursor_;
  params_.clear();
  index_.clear();
  woff_ = koff_ = scratch_off_ = scratch_max_ = 0;
  build();
  weight_bytes_ = woff_;
  work_bytes_ = koff_ + scratch_max_;
}

T UNet::resblock(const T& x, const std::string& name, int cout) {
  const int ted = 4 * cfg_.model_channels;
  T a = gn(&fwd_, x, name + ".in_layers.0", 1e-5f, true);
  const __half* w1 = param(name + ".in_layers.2.weight", 4, cout, 3, 3, x.c);
  const __half* b1 = param(name + ".in_layers.2.bias", 1, cout);
  // timestep projection rows live in the shared [emb_rows, ted] matrix
  const int row0 = emb_cursor_;
  emb_cursor_ += cout;
  if (dry_) {
    ParamInfo pw, pb;
    pw.name = name + ".emb_layers.1.weight";
    pw.ndim = 2, pw.shape[0] = cout, pw.shape[1] = ted, pw.shape[2] = pw.shape[3] = 1;
    pw.numel = (long long)cout * ted, pw.ptr = nullptr, pw.loaded = false;
    pb.name = name + ".emb_layers.1.bias";
    pb.ndim = 1, pb.shape[0] = cout, pb.shape[1] = pb.shape[2] = pb.shape[3] = 1;
    pb.numel = cout, pb.ptr = nullptr, pb.loaded = false;
    index_[pw.name] = (int)params_.size();
    params_.push_back(pw);
    index_[pb.name] = (int)params_.size();
    params_.push_back(pb);
  } else {
    params_[index_.at(name + ".emb_layers.1.weight")].ptr = emb_w_ + (long long)row0 * ted;
    params_[index_.at(name + ".emb_layers.1.bias")].ptr = emb_b_ + row0;
  }
  T h = conv3(&fwd_, a, w1, b1, cout, dry_ ? nullptr : emb_out_ + row0, emb_rows_);
  T b = gn(&fwd_, h, name + ".out_layers.0", 1e-5f, true);
  const __half* w2 = param(name + ".out_layers.3.weight", 4, cout, 3, 3, cout);
  const __half* b2 = param(name + ".out_layers.3.bias", 1, cout);
  T skip = x;
  if (x.c != cout) {
    const __half* ws = param(name + ".skip_connection.weight", 2, cout, x.c);
    const __half* bs = param(name + ".skip_connection.bias", 1, cout);
    skip = linear(&fwd_, x, ws, bs, cout);
  }
  return conv3(&fwd_, b, w2, b2, cout, nullptr, 0, &skip);
}

T UNet::transformer(const T& x, const std::string& name) {
  const int C = x.c, hd = cfg_.head_dim, heads = C / hd;
  const int HW = x.h * x.w, F = cfg_.num_frames;
  const std::string blk = name + ".transformer_blocks.0";
  T xn = gn(&fwd_, x, name + ".norm", 1e-6f, false);
  const __half* wpi = param(name + ".proj_in.weight", 2, C, C);
  const __half* bpi = param(name + ".proj_in.bias", 1, C);
  T h = linear(&fwd_, xn, wpi, bpi, C);

  auto ln = [&](const T& in, const std::string& nm) {
    const __half* g = param(nm + ".weight", 1, C);
    const __half* b = param(nm + ".bias", 1, C);
    T y = act(in.n, in.h, in.w, C);
    if (!dry_) {
      const T xx = in;
      fwd_.push_back([=](cudaStream_t s) { return layernorm_forward(xx.p, g, b, y.p, (int)xx.rows(), C, 1e-5f, s); });
    }
    return y;
  };

  // self-attention (over all frames of one object for the multi-view model)
  T n1 = ln(h, blk + ".norm1");
  const __half* wq = param(blk + ".attn1.to_q.weight", 2, C, C);
  param(blk + ".attn1.to_k.weight", 2, C, C);
  param(blk + ".attn1.to_v.weight", 2, C, C);  // q,k,v rows are contiguous: one [3C, C] matrix
  T qkv = linear(&fwd_, n1, wq, nullptr, 3 * C);
  reset_scratch();
  T a1 = attention(&fwd_, qkv.p, 3 * C, qkv.p + C, 3 * C, qkv.p + 2 * C, 3 * C, x.n / F, heads, hd, F * HW, F * HW);
  const __half* wo = param(blk + ".attn1.to_out.0.weight", 2, C, C);
  const __half* bo = param(blk + ".attn1.to_out.0.bias", 1, C);
  a1.n = x.n, a1.h = x.h, a1.w = x.w;
  h = linear(&fwd_, a1, wo, bo, C, &h);

  // cross-attention to the text tokens
  T n2 = ln(h, blk + ".norm2");
  const __half* wq2 = param(blk + ".attn2.to_q.weight", 2, C, C);
  T q2 = linear(&fwd_, n2, wq2, nullptr, C);
  const __half* wk2 = param(blk + ".attn2.to_k.weight", 2, C, cfg_.context_dim);
  param(blk + ".attn2.to_v.weight", 2, C, cfg_.context_dim);
  T ctx;
  ctx.p = in_ctx_, ctx.n = x.n, ctx.h = 1, ctx.w = cfg_.context_len, ctx.c = cfg_.context_dim;
  T kv = linear(&fwd_, ctx, wk2, nullptr, 2 * C);
  reset_scratch();
  T a2 = attention(&fwd_, q2.p, C, kv.p, 2 * C, kv.p + C, 2 * C, x.n, heads, hd, HW, cfg_.context_len);
  const __half* wo2 = param(blk + ".attn2.to_out.0.weight", 2, C, C);
  const __half* bo2 = param(blk + ".attn2.to_out.0.bias", 1, C);
  a2.n = x.n, a2.h = x.h, a2.w = x.w;
  h = linear(&fwd_, a2, wo2, bo2, C, &h);

  // GEGLU feed-forward
  T n3 = ln(h, blk + ".norm3");
  const __half* wf1 = param(blk + ".ff.net.0.proj.weight", 2, 8 * C, C);
  const __half* bf1 = param(blk + ".ff.net.0.proj.bias", 1, 8 * C);
  // GEGLU fused into the projection's epilogue: value / gate weight rows interleaved per 32-column chunk (derived copy)
  __half* wf1i = derived((long long)8 * C * C);
  __half* bf1i = derived(8 * C);
  post([=](cudaStream_t s) { return interleave_geglu_rows(wf1, wf1i, 4 * C, C, s); });
  post([=](cudaStream_t s) { return interleave_geglu_rows(bf1, bf1i, 4 * C, 1, s); });
  T gg = act(x.n, x.h, x.w, 4 * C);
  if (!dry_) {
    Epilogue ep;
    ep.out = gg.p;
    ep.ldc = 4 * C;
    ep.bias = bf1i;
    ep.act = kActGeglu;
    GemmPlan plan;
    if (fail(plan_gemm(&plan, n3.p, C, wf1i, C, (int)n3.rows(), 8 * C, C, ep))) return gg;
    fwd_.push_back([plan](cudaStream_t s) { return run_gemm(plan, s); });
  }
  const __half* wf2 = param(blk + ".ff.net.2.weight", 2, C, 4 * C);
  const __half* bf2 = param(blk + ".ff.net.2.bias", 1, C);
  h = linear(&fwd_, gg, wf2, bf2, C, &h);

  const __half* wpo = param(name + ".proj_out.weight", 2, C, C);
  const __half* bpo = param(name + ".proj_out.bias", 1, C);
  return linear(&fwd_, h, wpo, bpo, C, &x);
}

int UNet::build() {
  const int mc = cfg_.model_channels, ted = 4 * mc;
  emb_cursor_ = 0;
  // ---- staging + embeddings
  in_x_ = reinterpret_cast<__half*>(work((long long)B_ * H_ * W_ * 8 * 2));  // channel-padded to 8 is not needed; keep slack
  in_t_ = reinterpret_cast<float*>(work(B_ * 4));
  in_ctx_ = reinterpret_cast<__half*>(work((long long)B_ * cfg_.context_len * cfg_.context_dim * 2));
  in_cam_ = cfg_.camera_dim ? reinterpret_cast<__half*>(work(B_ * cfg_.camera_dim * 2)) : nullptr;
  out_ = reinterpret_cast<float*>(work((long long)B_ * H_ * W_ * cfg_.out_channels * 4));
  emb_w_ = derived((long long)emb_rows_ * ted);
  emb_b_ = derived(emb_rows_);
  emb_out_ = reinterpret_cast<float*>(work((long long)B_ * (emb_rows_ > 0 ? emb_rows_ : 1) * 4));
  emb_act_ = reinterpret_cast<__half*>(work((long long)B_ * ted * 2));

  __half* temb = reinterpret_cast<__half*>(work(B_ * mc * 2));
  __half* e1 = reinterpret_cast<__half*>(work(B_ * ted * 2));
  float* e2 = reinterpret_cast<float*>(work(B_ * ted * 4));
  const __half* tw0 = param("time_embed.0.weight", 2, ted, mc);
  const __half* tb0 = param("time_embed.0.bias", 1, ted);
  const __half* tw2 = param("time_embed.2.weight", 2, ted, ted);
  const __half* tb2 = param("time_embed.2.bias", 1, ted);
  float* ec = nullptr;
  if (!dry_) {
    const int B = B_;
    float* in_t = in_t_;
    fwd_.push_back([=](cudaStream_t s) { return timestep_embedding(in_t, temb, B, mc, 10000.f, s); });
    fwd_.push_back([=](cudaStream_t s) { return linear_small(temb, tw0, tb0, e1, 0, B, ted, mc, 0, s); });
    fwd_.push_back([=](cudaStream_t s) { return linear_small(e1, tw2, tb2, e2, 1, B, ted, ted, 1, s); });
  }
  if (cfg_.camera_dim) {
    __half* c1 = reinterpret_cast<__half*>(work(B_ * ted * 2));
    ec = reinterpret_cast<float*>(work(B_ * ted * 4));
    const __half* cw0 = param("camera_embed.0.weight", 2, ted, cfg_.camera_dim);
    const __half* cb0 = param("camera_embed.0.bias", 1, ted);
    const __half* cw2 = param("camera_embed.2.weight", 2, ted, ted);
    const __half* cb2 = param("camera_embed.2.bias", 1, ted);
    if (!dry_) {
      const int B = B_, cd = cfg_.camera_dim;
      __half* cam = in_cam_;
      fwd_.push_back([=](cudaStream_t s) { return linear_small(cam, cw0, cb0, c1, 0, B, ted, cd, 0, s); });
      fwd_.push_back([=](cudaStream_t s) { return linear_small(c1, cw2, cb2, ec, 1, B, ted, ted, 1, s); });
    }
  }
  if (!dry_) {
    const int B = B_, rows = emb_rows_;
    __half* ea = emb_act_;
    float* eo = emb_out_;
    const __half *ew = emb_w_, *eb = emb_b_;
    fwd_.push_back([=](cudaStream_t s) { return add_silu_f32_to_f16(e2, ec, ea, (long long)B * ted, s); });
    fwd_.push_back([=](cudaStream_t s) { return linear_small(ea, ew, eb, eo, 1, B, rows, ted, 0, s); });
  }

  // ---- input blocks
  const __half* w_in = param("input_blocks.0.0.weight", 4, mc, 3, 3, cfg_.in_channels);
  const __half* b_in = param("input_blocks.0.0.bias", 1, mc);
  T h = act(B_, H_, W_, mc);
  if (!dry_) {
    const int B = B_, H = H_, W = W_, ci = cfg_.in_channels;
    __half* xin = in_x_;
    fwd_.push_back([=](cudaStream_t s) { return conv3x3_small(xin, 0, w_in, b_in, h.p, 0, B, H, W, ci, mc, s); });
  }
  std::vector<T> hs;
  hs.push_back(h);
  int ib = 1;
  for (int level = 0; level < cfg_.num_levels; ++level) {
    const int cout = cfg_.channel_mult[level] * mc;
    for (int r = 0; r < cfg_.num_res_blocks; ++r) {
      const std::string nm = "input_blocks." + std::to_string(ib);
      h = resblock(h, nm + ".0", cout);
      if (level < cfg_.attn_levels) h = transformer(h, nm + ".1");
      hs.push_back(h);
      ++ib;
    }
    if (level != cfg_.num_levels - 1) {
      const std::string nm = "input_blocks." + std::to_string(ib) + ".0.op";
      const __half* w = param(nm + ".weight", 4, h.c, 3, 3, h.c);
      const __half* b = param(nm + ".bias", 1, h.c);
      h = conv3_s2(&fwd_, h, w, b, h.c, 1);
      hs.push_back(h);
      ++ib;
    }
  }
  // ---- middle
  h = resblock(h, "middle_block.0", h.c);
  h = transformer(h, "middle_block.1");
  h = resblock(h, "middle_block.2", h.c);
  // ---- output blocks
  int ob = 0;
  for (int level = cfg_.num_levels - 1; level >= 0; --level) {
    const int cout = cfg_.channel_mult[level] * mc;
    for (int i = 0; i <= cfg_.num_res_blocks; ++i) {
      const T skip = hs.back();
      hs.pop_back();
      T cat = act(h.n, h.h, h.w, h.c + skip.c);
      if (!dry_) {
        const T hh = h;
        fwd_.push_back([=](cudaStream_t s) { return concat_channels(hh.p, hh.c, skip.p, skip.c, cat.p, hh.rows(), s); });
      }
      const std::string nm = "output_blocks." + std::to_string(ob);
      h = resblock(cat, nm + ".0", cout);
      int sub = 1;
      if (level < cfg_.attn_levels) {
        h = transformer(h, nm + ".1");
        sub = 2;
      }
      if (level && i == cfg_.num_res_blocks) {
        const std::string un = nm + "." + std::to_string(sub) + ".conv";
        const __half* w = param(un + ".weight", 4, h.c, 3, 3, h.c);
        const __half* b = param(un + ".bias", 1, h.c);
        T up = act(h.n, 2 * h.h, 2 * h.w, h.c);
        if (!dry_) {
          const T hh = h;
          fwd_.push_back([=](cudaStream_t s) { return upsample_nearest2x(hh.p, up.p, hh.n, hh.h, hh.w, hh.c, s); });
        }
        h = conv3(&fwd_, up, w, b, h.c);
      }
      ++ob;
    }
  }
  // ---- out
  T a = gn(&fwd_, h, "out.0", 1e-5f, true);
  const __half* w_out = param("out.2.weight", 4, cfg_.out_channels, 3, 3, mc);
  const __half* b_out = param("out.2.bias", 1, cfg_.out_channels);
  if (!dry_) {
    const int B = B_, H = H_, W = W_, co = cfg_.out_channels;
    float* out = out_;
    fwd_.push_back([=](cudaStream_t s) { return conv3x3_small(a.p, 0, w_out, b_out, out, 1, B, H, W, mc, co, s); });
  }
  return SDB_OK;
}

int UNet::forward(const void* x, const float* t, const void* ctx, const void* camera, float* out, cudaStream_t s) {
  if (dry_) {
    sdb_set_error("unet: bind() first");
    return SDB_ERR_ARG;
  }
  if (!x || !t || !ctx || !out || (cfg_.camera_dim && !camera)) {
    sdb_set_error("unet: NULL input");
    return SDB_ERR_ARG;
  }
  const size_t nx = (size_t)B_ * H_ * W_ * cfg_.in_channels * 2;
  cudaMemcpyAsync(in_x_, x, nx, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(in_t_, t, B_ * 4, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(in_ctx_, ctx, (size_t)B_ * cfg_.context_len * cfg_.context_dim * 2, cudaMemcpyDeviceToDevice, s);
  if (cfg_.camera_dim) cudaMemcpyAsync(in_cam_, camera, B_ * cfg_.camera_dim * 2, cudaMemcpyDeviceToDevice, s);
  int rc = run(fwd_, s);
  if (rc) return rc;
  cudaMemcpyAsync(out, out_, (size_t)B_ * H_ * W_ * cfg_.out_channels * 4, cudaMemcpyDeviceToDevice, s);
  return SDB_OK;
}

// ---------------------------------------------------------------------------------------------------- VAE
VaeEncoder::VaeEncoder(const VaeCfg& cfg, int batch, int h, int w) : cfg_(cfg), B_(batch), H_(h), W_(w) {
  dry_ = true;
  build();
  weight_bytes_ = woff_;
  work_bytes_ = koff_ + scratch_max_;
}

int VaeEncoder::build() {
  typedef std::function<T(const T&)> BwdFn;  // dy -> dx, appends to bwd_
  std::vector<BwdFn> tape;
  const int ch = cfg_.ch;

  in_x_ = reinterpret_cast<float*>(work((long long)B_ * H_ * W_ * cfg_.in_channels * 4));
  out_dx_ = reinterpret_cast<float*>(work((long long)B_ * H_ * W_ * cfg_.in_channels * 4));
  const int zc2 = 2 * cfg_.z_channels;
  out_ = reinterpret_cast<float*>(work((long long)B_ * (H_ / 8) * (W_ / 8) * zc2 * 4));
  in_dy_ = reinterpret_cast<float*>(work((long long)B_ * (H_ / 8) * (W_ / 8) * zc2 * 4));

  // y = GN(+SiLU)(x)
  auto gn_fb = [&](const T& x, const std::string& name, bool silu) {
    float* stats = nullptr;
    T y = gn(&fwd_, x, name, 1e-6f, silu, &stats);
    const __half* g = dry_ ? nullptr : params_[index_.at(name + ".weight")].ptr;
    const __half* b = dry_ ? nullptr : params_[index_.at(name + ".bias")].ptr;
    tape.push_back([=](const T& dy) {
      T dx = act(x.n, x.h, x.w, x.c);
      float* sc = reinterpret_cast<float*>(work(groupnorm_workspace_floats(x.n, x.h * x.w, x.c, 32) * 4));
      bwd([=](cudaStream_t s) {
        return groupnorm_backward(x.p, g, b, stats, dy.p, dx.p, sc, x.n, x.h * x.w, x.c, 32, 1e-6f, silu ? 1 : 0, s);
      });
      return dx;
    });
    return y;
  };
  // y = conv3x3(x) (+ residual)
  auto conv_fb = [&](const T& x, const std::string& name, int cout, const T* residual) {
    const __half* w = param(name + ".weight", 4, cout, 3, 3, x.c);
    const __half* b = param(name + ".bias", 1, cout);
    __half* wr = derived((long long)cout * 9 * x.c);
    const int cin = x.c;
    post([=](cudaStream_t s) { return rotate_w3x3(w, wr, cout, cin, s); });
    T y = conv3(&fwd_, x, w, b, cout, nullptr, 0, residual);
    tape.push_back([=](const T& dy) { return conv3(&bwd_, dy, wr, nullptr, cin); });
    return y;
  };

  struct Saved { T x; };
  // ResnetBlock (model.py:90-149): out = shortcut(x) + conv2(silu(gn2(conv1(silu(gn1(x))))))
  auto resblock_fb = [&](const T& x, const std::string& name, int cout) {
    const size_t mark = tape.size();
    T a = gn_fb(x, name + ".norm1", true);
    T h = conv_fb(a, name + ".conv1", cout, nullptr);
    T b = gn_fb(h, name + ".norm2", true);
    T skip = x;
    const __half* wn = nullptr;
    __half* wnt = nullptr;
    const int cin = x.c;
    if (cin != cout) {
      wn = param(name + ".nin_shortcut.weight", 2, cout, cin);
      const __half* bn = param(name + ".nin_shortcut.bias", 1, cout);
      wnt = derived((long long)cout * cin);
      post([=](cudaStream_t s) { return transpose_f16(wn, wnt, cout, cin, s); });
      skip = linear(&fwd_, x, wn, bn, cout);
    }
    T y = conv_fb(b, name + ".conv2", cout, &skip);
    // fold the four branch emitters into one block emitter
    std::vector<BwdFn> branch(tape.begin() + mark, tape.end());
    tape.resize(mark);
    tape.push_back([=](const T& dy) {
      T d = dy;
      for (auto it = branch.rbegin(); it != branch.rend(); ++it) d = (*it)(d);
      if (cin != cout) return linear(&bwd_, dy, wnt, nullptr, cin, &d);  // d_x = dy W_nin + d_branch
      T dx = act(d.n, d.h, d.w, d.c);
      bwd([=](cudaStream_t s) { return add_f16(d.p, dy.p, dx.p, d.numel(), s); });
      return dx;
    });
    return y;
  };
  // Downsample (model.py:67-87): pad (0,1,0,1) then 3x3 stride 2
  auto down_fb = [&](const T& x, const std::string& name) {
    const int C = x.c;
    const __half* w = param(name + ".conv.weight", 4, C, 3, 3, C);
    const __half* b = param(name + ".conv.bias", 1, C);
    __half* wt = derived((long long)C * 9 * C);
    post([=](cudaStream_t s) { return transpose_f16(w, wt, C, 9 * C, s); });
    T y = conv3_s2(&fwd_, x, w, b, C, 0);
    tape.push_back([=](const T& dy) {
      reset_scratch();
      T dcol;
      dcol.n = dy.n, dcol.h = dy.h, dcol.w = dy.w, dcol.c = 9 * C;
      dcol.p = reinterpret_cast<__half*>(scratch(dcol.numel() * 2));
      if (!dry_) {
        Epilogue ep;
        ep.out = dcol.p;
        ep.ldc = 9 * C;
        GemmPlan plan;
        if (!fail(plan_gemm(&plan, dy.p, C, wt, C, (int)dy.rows(), 9 * C, C, ep)))
          bwd([plan](cudaStream_t s) { return run_gemm(plan, s); });
      }
      T dx = act(x.n, x.h, x.w, C);
      bwd([=](cudaStream_t s) { return col2im_3x3_s2(dcol.p, dx.p, x.n, x.h, x.w, C, 0, s); });
      return dx;
    });
    return y;
  };
  // AttnBlock (model.py:152-203): single head over all pixels, head_dim = C
  auto attn_fb = [&](const T& x, const std::string& name) {
    const int C = x.c, L = x.h * x.w, B = x.n;
    const size_t mark = tape.size();
    T xn = gn_fb(x, name + ".norm", false);
    BwdFn gn_bwd = tape.back();
    tape.resize(mark);
    const __half* wq = param(name + ".q.weight", 2, C, C);
    param(name + ".k.weight", 2, C, C);
    param(name + ".v.weight", 2, C, C);
    const __half* bq = param(name + ".q.bias", 1, C);
    param(name + ".k.bias", 1, C);
    param(name + ".v.bias", 1, C);
    __half* wqkv_t = derived((long long)3 * C * C);
    post([=](cudaStream_t s) { return transpose_f16(wq, wqkv_t, 3 * C, C, s); });
    if ((C * 2) % 256) fail(SDB_ERR_UNSUPPORTED);  // q/k/v biases must stay contiguous
    T qkv = linear(&fwd_, xn, wq, bq, 3 * C);
    __half* P = nullptr;
    T o = attention(&fwd_, qkv.p, 3 * C, qkv.p + C, 3 * C, qkv.p + 2 * C, 3 * C, B, 1, C, L, L, &P);
    o.n = x.n, o.h = x.h, o.w = x.w;
    const __half* wp = param(name + ".proj_out.weight", 2, C, C);
    const __half* bp = param(name + ".proj_out.bias", 1, C);
    __half* wpt = derived((long long)C * C);
    post([=](cudaStream_t s) { return transpose_f16(wp, wpt, C, C, s); });
    T y = linear(&fwd_, o, wp, bp, C, &x);
    const float scale = 1.f / sqrtf((float)C);
    tape.push_back([=](const T& dy) {
      T d_o = linear(&bwd_, dy, wpt, nullptr, C);
      T dqkv = act(x.n, x.h, x.w, 3 * C);
      const long long lds = align_up(L, 8);
      __half* Pt = reinterpret_cast<__half*>(work((long long)B * L * lds * 2));
      __half* dP = reinterpret_cast<__half*>(work((long long)B * L * lds * 2));
      __half* dSt = reinterpret_cast<__half*>(work((long long)B * L * lds * 2));
      if (!dry_) {
        GemmPlan p_dv, p_dp, p_dq, p_dk;
        // dV = P^T dO ; dP = dO V^T ; dS = P o (dP - rowsum(dP o P)) * scale ; dQ = dS K ; dK = dS^T Q
        if (fail(plan_attn_apply(&p_dv, Pt, lds, d_o.p, C, B, 1, C, L, L, dqkv.p + 2 * C, 3 * C))) return dqkv;
        if (fail(plan_attn_scores(&p_dp, d_o.p, C, qkv.p + 2 * C, 3 * C, B, 1, C, L, L, dP, lds, 1.f))) return dqkv;
        if (fail(plan_attn_apply(&p_dq, dP, lds, qkv.p + C, 3 * C, B, 1, C, L, L, dqkv.p, 3 * C))) return dqkv;
        if (fail(plan_attn_apply(&p_dk, dSt, lds, qkv.p, 3 * C, B, 1, C, L, L, dqkv.p + C, 3 * C))) return dqkv;
        for (int b = 0; b < B; ++b) {
          const __half* Pb = P + (long long)b * L * lds;
          __half* Ptb = Pt + (long long)b * L * lds;
          bwd([=](cudaStream_t s) { return transpose_f16(Pb, Ptb, L, (int)lds, s); });
        }
        bwd([p_dv](cudaStream_t s) { return run_gemm(p_dv, s); });
        bwd([p_dp](cudaStream_t s) { return run_gemm(p_dp, s); });
        bwd([=](cudaStream_t s) { return softmax_rows_backward(P, dP, (long long)B * L, L, lds, scale, s); });
        bwd([p_dq](cudaStream_t s) { return run_gemm(p_dq, s); });
        for (int b = 0; b < B; ++b) {
          const __half* dSb = dP + (long long)b * L * lds;
          __half* dStb = dSt + (long long)b * L * lds;
          bwd([=](cudaStream_t s) { return transpose_f16(dSb, dStb, L, (int)lds, s); });
        }
        bwd([p_dk](cudaStream_t s) { return run_gemm(p_dk, s); });
      }
      T d_xn = linear(&bwd_, dqkv, wqkv_t, nullptr, C);
      T d_x = gn_bwd(d_xn);
      T dx = act(x.n, x.h, x.w, C);
      bwd([=](cudaStream_t s) { return add_f16(d_x.p, dy.p, dx.p, dx.numel(), s); });
      return dx;
    });
    return y;
  };

  // ---- conv_in (3 -> ch, CUDA cores; fp32 image in)
  const __half* w_in = param("encoder.conv_in.weight", 4, ch, 3, 3, cfg_.in_channels);
  const __half* b_in = param("encoder.conv_in.bias", 1, ch);
  __half* w_in_r = derived((long long)ch * 9 * cfg_.in_channels);
  {
    const int ci = cfg_.in_channels;
    post([=](cudaStream_t s) { return rotate_w3x3(w_in, w_in_r, ch, ci, s); });
  }
  T h = act(B_, H_, W_, ch);
  if (!dry_) {
    const int B = B_, H = H_, W = W_, ci = cfg_.in_channels;
    float* xin = in_x_;
    fwd_.push_back([=](cudaStream_t s) { return conv3x3_small(xin, 1, w_in, b_in, h.p, 0, B, H, W, ci, ch, s); });
  }
  int cin = ch;
  for (int level = 0; level < cfg_.num_levels; ++level) {
    const int cout = ch * cfg_.ch_mult[level];
    for (int r = 0; r < cfg_.num_res_blocks; ++r) {
      h = resblock_fb(h, "encoder.down." + std::to_string(level) + ".block." + std::to_string(r), cout);
      cin = cout;
    }
    if (level != cfg_.num_levels - 1) h = down_fb(h, "encoder.down." + std::to_string(level) + ".downsample");
  }
  h = resblock_fb(h, "encoder.mid.block_1", cin);
  h = attn_fb(h, "encoder.mid.attn_1");
  h = resblock_fb(h, "encoder.mid.block_2", cin);
  T a = gn_fb(h, "encoder.norm_out", true);
  const __half* w_out = param("encoder.conv_out.weight", 4, zc2, 3, 3, cin);
  const __half* b_out = param("encoder.conv_out.bias", 1, zc2);
  __half* w_out_r = derived((long long)zc2 * 9 * cin);
  post([=](cudaStream_t s) { return rotate_w3x3(w_out, w_out_r, zc2, cin, s); });
  if (!dry_) {
    float* out = out_;
    fwd_.push_back([=](cudaStream_t s) { return conv3x3_small(a.p, 0, w_out, b_out, out, 1, a.n, a.h, a.w, cin, zc2, s); });
  }

  // ---- backward list: conv_out dgrad, then the tape in reverse, then conv_in dgrad
  T d = act(a.n, a.h, a.w, cin);
  if (!dry_) {
    float* dy = in_dy_;
    bwd_.push_back([=](cudaStream_t s) { return conv3x3_small(dy, 1, w_out_r, nullptr, d.p, 0, d.n, d.h, d.w, zc2, cin, s); });
  }
  for (auto it = tape.rbegin(); it != tape.rend(); ++it) d = (*it)(d);
  if (!dry_) {
    float* dx = out_dx_;
    const int B = B_, H = H_, W = W_, ci = cfg_.in_channels;
    bwd_.push_back([=](cudaStream_t s) { return conv3x3_small(d.p, 0, w_in_r, nullptr, dx, 1, B, H, W, ch, ci, s); });
  }
  return SDB_OK;
}

int VaeEncoder::forward(const float* x, float* moments_pre, cudaStream_t s) {
  if (dry_ || !x || !moments_pre) {
    sdb_set_error("vae: bind() first / NULL argument");
    return SDB_ERR_ARG;
  }
  cudaMemcpyAsync(in_x_, x, (size_t)B_ * H_ * W_ * cfg_.in_channels * 4, cudaMemcpyDeviceToDevice, s);
  int rc = run(fwd_, s);
  if (rc) return rc;
  cudaMemcpyAsync(moments_pre, out_, (size_t)B_ * (H_ / 8) * (W_ / 8) * 2 * cfg_.z_channels * 4,
                  cudaMemcpyDeviceToDevice, s);
  return SDB_OK;
}

int VaeEncoder::backward(const float* d_moments_pre, float* d_x, cudaStream_t s) {
  if (dry_ || !d_moments_pre || !d_x) {
    sdb_set_error("vae: bind() first / NULL argument");
    return SDB_ERR_ARG;
  }
  cudaMemcpyAsync(in_dy_, d_moments_pre, (size_t)B_ * (H_ / 8) * (W_ / 8) * 2 * cfg_.z_channels * 4,
                  cudaMemcpyDeviceToDevice, s);
  int rc = run(bwd_, s);
  if (rc) return rc;
  cudaMemcpyAsync(d_x, out_dx_, (size_t)B_ * H_ * W_ * cfg_.in_channels * 4, cudaMemcpyDeviceToDevice, s);
  return SDB_OK;
}

}  // namespace nn
