// extern "C" entry points of the network executors (include/sdb200_nn.h, second half).
#include "../../include/sdb200_nn.h"
#include "models.h"

struct sdb_net {
  int kind;  // 0 UNet, 1 VAE encoder
  nn::Net* net;
};

extern "C" {

int sdb_unet_create(const sdb_unet_cfg* c, int batch, int height, int width, sdb_net** out) {
  SDB_CHECK_ARG(c && out && batch > 0 && height > 0 && width > 0, "unet_create: bad arguments");
  SDB_CHECK_ARG(c->num_levels >= 1 && c->num_levels <= 4 && c->model_channels % 64 == 0 && c->head_dim % 64 == 0,
                "unet_create: unsupported configuration");
  SDB_CHECK_ARG(c->num_frames >= 1 && batch % c->num_frames == 0, "unet_create: batch must be a multiple of num_frames");
  SDB_CHECK_ARG(batch <= 16, "unet_create: batch must be <= 16");
  SDB_CHECK_ARG(height % (1 << (c->num_levels - 1)) == 0 && width % (1 << (c->num_levels - 1)) == 0,
                "unet_create: latent size must be divisible by 2^(levels-1)");
  nn::UNetCfg u;
  u.in_channels = c->in_channels;
  u.out_channels = c->out_channels;
  u.model_channels = c->model_channels;
  u.num_levels = c->num_levels;
  for (int i = 0; i < 4; ++i) u.channel_mult[i] = c->channel_mult[i];
  u.num_res_blocks = c->num_res_blocks;
  u.attn_levels = c->attn_levels;
  u.head_dim = c->head_dim;
  u.context_dim = c->context_dim;
  u.context_len = c->context_len;
  u.camera_dim = c->camera_dim;
  u.num_frames = c->num_frames;
  sdb_net* h = new sdb_net;
  h->kind = 0;
  h->net = new nn::UNet(u, batch, height, width);
  *out = h;
  return SDB_OK;
}

int sdb_unet_forward(sdb_net* net, const void* x, const float* t, const void* ctx, const void* camera, float* out,
                     void* stream) {
  SDB_CHECK_ARG(net && net->kind == 0, "unet_forward: not a UNet handle");
  return static_cast<nn::UNet*>(net->net)->forward(x, t, ctx, camera, out, (cudaStream_t)stream);
}

int sdb_vae_encoder_create(const sdb_vae_cfg* c, int batch, int height, int width, sdb_net** out) {
  SDB_CHECK_ARG(c && out && batch > 0 && height > 0 && width > 0, "vae_create: bad arguments");
  SDB_CHECK_ARG(c->num_levels >= 1 && c->num_levels <= 4 && c->ch % 64 == 0, "vae_create: unsupported configuration");
  SDB_CHECK_ARG(height % 8 == 0 && width % 8 == 0, "vae_create: image size must be divisible by 8");
  nn::VaeCfg v;
  v.in_channels = c->in_channels;
  v.ch = c->ch;
  v.num_levels = c->num_levels;
  for (int i = 0; i < 4; ++i) v.ch_mult[i] = c->ch_mult[i];
  v.num_res_blocks = c->num_res_blocks;
  v.z_channels = c->z_channels;
  sdb_net* h = new sdb_net;
  h->kind = 1;
  h->net = new nn::VaeEncoder(v, batch, height, width);
  *out = h;
  return SDB_OK;
}

int sdb_vae_encoder_forward(sdb_net* net, const float* x, float* h, void* stream) {
  SDB_CHECK_ARG(net && net->kind == 1, "vae_forward: not a VAE handle");
  return static_cast<nn::VaeEncoder*>(net->net)->forward(x, h, (cudaStream_t)stream);
}

int sdb_vae_encoder_backward(sdb_net* net, const float* d_h, float* d_x, void* stream) {
  SDB_CHECK_ARG(net && net->kind == 1, "vae_backward: not a VAE handle");
  return static_cast<nn::VaeEncoder*>(net->net)->backward(d_h, d_x, (cudaStream_t)stream);
}

void sdb_net_destroy(sdb_net* net) {
  if (!net) return;
  delete net->net;
  delete net;
}

int sdb_net_sizes(sdb_net* net, long long* weight_bytes, long long* work_bytes) {
  SDB_CHECK_ARG(net && weight_bytes && work_bytes, "net_sizes: NULL argument");
  *weight_bytes = net->net->weight_bytes();
  *work_bytes = net->net->work_bytes();
  return SDB_OK;
}

int sdb_net_bind(sdb_net* net, void* weights, void* work) {
  SDB_CHECK_ARG(net, "net_bind: NULL handle");
  return net->net->bind(weights, work);
}

int sdb_net_num_params(sdb_net* net) { return net ? (int)net->net->params().size() : -1; }

int sdb_net_param(sdb_net* net, int index, const char** name, int* ndim, int* shape4) {
  SDB_CHECK_ARG(net && index >= 0 && index < (int)net->net->params().size() && name && ndim && shape4,
                "net_param: bad arguments");
  const nn::ParamInfo& p = net->net->params()[index];
  *name = p.name.c_str();
  *ndim = p.ndim;
  for (int i = 0; i < 4; ++i) shape4[i] = p.shape[i];
  return SDB_OK;
}

int sdb_net_load_param(sdb_net* net, const char* name, const void* src, long long numel, void* stream) {
  SDB_CHECK_ARG(net && name && src, "net_load_param: NULL argument");
  return net->net->load_param(name, src, numel, (cudaStream_t)stream);
}

int sdb_net_finalize(sdb_net* net, void* stream) {
  SDB_CHECK_ARG(net, "net_finalize: NULL handle");
  return net->net->finalize((cudaStream_t)stream);
}

int sdb_net_num_launches(sdb_net* net, int backward) {
  if (!net) return -1;
  if (backward) return net->kind == 1 ? static_cast<nn::VaeEncoder*>(net->net)->launches_per_backward() : 0;
  return net->net->launches_per_forward();
}

}  // extern "C"
