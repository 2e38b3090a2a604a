// Native executors of the two frozen networks of the ASD guidance step: the latent-diffusion UNet (single-view
// SD-2.1 shape or MVDream multi-view) and the KL-VAE encoder (forward + data-gradient backward).
// A Net is built once for a fixed (batch, height, width): it enumerates its parameters under the reference's
// state-dict names, owns a packed fp16 copy of them inside a caller-provided arena, pre-builds every TMA
// descriptor / kernel plan over caller-provided workspace, and replays the op list on a stream.
#pragma once
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "dense.h"

namespace nn {

struct T {  // channels-last activation: [n, h, w, c] fp16 == [n*h*w, c]
  __half* p = nullptr;
  int n = 0, h = 0, w = 0, c = 0;
  long long rows() const { return (long long)n * h * w; }
  long long numel() const { return rows() * c; }
};

struct ParamInfo {
  std::string name;
  int ndim;
  int shape[4];      // layout the loader must supply (fp16): conv3x3 [Cout,3,3,Cin], linear / 1x1 conv [out,in], vec [n]
  long long numel;
  __half* ptr;
  bool loaded;
};

struct UNetCfg {
  int in_channels = 4, out_channels = 4, model_channels = 320;
  int num_levels = 4;
  int channel_mult[4] = {1, 2, 4, 4};
  int num_res_blocks = 2;
  int attn_levels = 3;       // levels [0, attn_levels) carry a SpatialTransformer (attention_resolutions 4,2,1)
  int head_dim = 64;
  int context_dim = 1024;
  int context_len = 77;
  int camera_dim = 0;        // 16 for MVDream
  int num_frames = 1;        // 4 for MVDream: self-attention spans the frames of one object
};

struct VaeCfg {
  int in_channels = 3, ch = 128, num_levels = 4;
  int ch_mult[4] = {1, 2, 4, 4};
  int num_res_blocks = 2;
  int z_channels = 4;        // encoder emits 2*z_channels moments
};

class Net {
 public:
  virtual ~Net() {}
  // sizes of the two arenas the caller must provide (bytes)
  long long weight_bytes() const { return weight_bytes_; }
  long long work_bytes() const { return work_bytes_; }
  int bind(void* weights, void* work);  // builds the op list; returns 0 or error
  const std::vector<ParamInfo>& params() const { return params_; }
  int load_param(const char* name, const void* src_fp16, long long numel, cudaStream_t s);
  int finalize(cudaStream_t s);  // derived weights (fused / rotated copies); checks every parameter was loaded
  int launches_per_forward() const { return (int)fwd_.size(); }

 protected:
  typedef std::function<int(cudaStream_t)> Op;
  virtual int build() = 0;  // called twice: dry (counting) and bound
  int run(const std::vector<Op>& ops, cudaStream_t s);

  // ---- builder helpers (valid inside build()) ----
  bool dry_ = true;
  __half* param(const std::string& name, int ndim, int d0, int d1 = 1, int d2 = 1, int d3 = 1);
  __half* derived(long long numel);                  // weight-arena space filled by finalize()
  void* work(long long bytes);                       // persistent activation / scratch
  T act(int n, int h, int w, int c);
  void reset_scratch() { scratch_off_ = 0; }
  void* scratch(long long bytes);                    // reused region (attention scores, im2col buffers)
  void fwd(Op op) { if (!dry_) fwd_.push_back(std::move(op)); }
  void bwd(Op op) { if (!dry_) bwd_.push_back(std::move(op)); }
  void post(Op op) { if (!dry_) post_.push_back(std::move(op)); }
  int fail(int rc) { if (rc && !build_rc_) build_rc_ = rc; return rc; }

  // layer emitters; `ops` selects forward or backward list
  T gn(std::vector<Op>* ops, const T& x, const std::string& name, float eps, bool silu, float** stats_out = nullptr);
  T conv3(std::vector<Op>* ops, const T& x, const __half* w, const __half* bias, int cout, const float* rowbias = nullptr,
          long long rowbias_ld = 0, const T* residual = nullptr);
  T linear(std::vector<Op>* ops, const T& x, const __half* w, const __half* bias, int cout, const T* residual = nullptr,
           int act = 0);
  T conv3_s2(std::vector<Op>* ops, const T& x, const __half* w, const __half* bias, int cout, int pad_lo);
  // softmax(alpha q k^T) v with q/k/v given as (pointer, row stride); heads*head_dim channels
  T attention(std::vector<Op>* ops, const __half* q, long long ldq, const __half* k, long long ldk, const __half* v,
              long long ldv, int B, int heads, int head_dim, int Lq, int Lk, __half** probs_out = nullptr);

  std::vector<ParamInfo> params_;
  std::unordered_map<std::string, int> index_;
  std::vector<Op> fwd_, bwd_, post_;
  long long weight_bytes_ = 0, work_bytes_ = 0;
  char* wbase_ = nullptr;
  char* kbase_ = nullptr;
  long long woff_ = 0, koff_ = 0, scratch_off_ = 0, scratch_max_ = 0, scratch_base_off_ = 0;
  int build_rc_ = 0;
  friend class UNet;
  friend class VaeEncoder;
};

class UNet : public Net {
 public:
  UNet(const UNetCfg& cfg, int batch, int h, int w);
  // x: fp16 [B,H,W,in_ch]; t: fp32 [B]; ctx: fp16 [B,ctx_len,ctx_dim]; camera: fp16 [B,camera_dim] or null;
  // out: fp32 [B,H,W,out_ch]
  int forward(const void* x, const float* t, const void* ctx, const void* camera, float* out, cudaStream_t s);

 protected:
  int build() override;
  T resblock(const T& x, const std::string& name, int cout);
  T transformer(const T& x, const std::string& name);
  UNetCfg cfg_;
  int B_, H_, W_;
  // staging buffers the forward copies its inputs into (fixed addresses inside the plans)
  __half* in_x_ = nullptr;
  float* in_t_ = nullptr;
  __half* in_ctx_ = nullptr;
  __half* in_cam_ = nullptr;
  float* out_ = nullptr;
  // all ResBlock timestep projections as one [emb_rows, 4*mc] matrix
  __half* emb_w_ = nullptr;
  __half* emb_b_ = nullptr;
  float* emb_out_ = nullptr;
  __half* emb_act_ = nullptr;
  int emb_rows_ = 0, emb_cursor_ = 0;
};

class VaeEncoder : public Net {
 public:
  VaeEncoder(const VaeCfg& cfg, int batch, int h, int w);
  // x: fp32 [B,H,W,3] in [-1,1]; moments_pre: fp32 [B,H/8,W/8,2*z] (conv_out output, BEFORE quant_conv)
  int forward(const float* x, float* moments_pre, cudaStream_t s);
  // d_moments_pre: fp32 [B,H/8,W/8,2*z] -> d_x fp32 [B,H,W,3]. Must follow a forward() (uses its saved tensors).
  int backward(const float* d_moments_pre, float* d_x, cudaStream_t s);
  int launches_per_backward() const { return (int)bwd_.size(); }

 protected:
  int build() override;
  VaeCfg cfg_;
  int B_, H_, W_;
  float* in_x_ = nullptr;
  float* out_ = nullptr;
  float* in_dy_ = nullptr;
  float* out_dx_ = nullptr;
};

}  // namespace nn
