// extern "C" entry points of include/sdb200.h: argument validation, hash-grid geometry resolution and
// dispatch to the sm_100a kernels. No torch types cross this boundary.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <cstdlib>

#include "../../include/sdb200.h"
#include "render_tape.cuh"

unsigned long long g_sdb_launch_count = 0ull;
bool sdb_pdl_enabled() {
  static const bool on = !(getenv("SDB_PDL") && atoi(getenv("SDB_PDL")) == 0);
  return on;
}
static thread_local char g_err[512] = "";

void sdb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int launch_render_fwd(const FieldMeta&, const FieldPtrs&, const MarchMeta&, const RayIO&, const PackedOut&,
                      cudaStream_t);
int launch_render_bwd(const FieldMeta&, const FieldPtrs&, const FieldGrads&, const MarchMeta&, const RayIO&,
                      cudaStream_t);
int launch_field_eval(const FieldMeta&, const FieldPtrs&, const float*, int, float*, float*, float*, cudaStream_t);
int launch_hashgrid_fwd(const GridMeta&, const float*, const float*, int, float*, cudaStream_t);
int launch_hashgrid_bwd(const GridMeta&, const float*, const float*, int, float*, cudaStream_t);
int launch_occ_update(const FieldMeta&, const FieldPtrs&, const int*, const float*, int, int, float, float, float,
                      float*, uint32_t*, float*, cudaStream_t);

// tiny-cuda-nn GridEncoding level geometry (restated; see oracle/render_oracle.py::grid_meta).
static int resolve_grid(const sdb_grid_cfg* c, GridMeta* gm) {
  if (!c || c->n_levels < 1 || c->n_levels > kMaxLevels) {
    sdb_set_error("grid: n_levels must be in [1,%d]", kMaxLevels);
    return SDB_ERR_ARG;
  }
  if (c->n_features_per_level != 2) {
    sdb_set_error("grid: only n_features_per_level=2 is implemented (got %d)", c->n_features_per_level);
    return SDB_ERR_UNSUPPORTED;
  }
  if (c->log2_hashmap_size < 1 || c->log2_hashmap_size > 28) {
    sdb_set_error("grid: log2_hashmap_size out of range");
    return SDB_ERR_ARG;
  }
  memset(gm, 0, sizeof(*gm));
  gm->n_levels = c->n_levels;
  const uint32_t max_params = 1u << c->log2_hashmap_size;
  uint32_t offset = 0;
  const double log2_pls = std::log2((double)c->per_level_scale);
  for (int l = 0; l < c->n_levels; ++l) {
    const float scale = (float)(std::exp2((double)l * log2_pls) * (double)c->base_resolution - 1.0);
    const uint32_t res = (uint32_t)std::ceil(scale) + 1u;
    const unsigned long long dense = (unsigned long long)res * res * res;
    unsigned long long n = (dense + 7ull) / 8ull * 8ull;  // tcnn: next_multiple(res^3, 8) then min(., 2^log2)
    if (n > max_params) n = max_params;
    gm->scale[l] = scale;
    gm->res[l] = res;
    gm->size[l] = (uint32_t)n;
    gm->offset[l] = offset;
    gm->hashed[l] = dense > n ? 1u : 0u;
    offset += (uint32_t)n;
  }
  return SDB_OK;
}

static int resolve_field(const sdb_field* f, FieldMeta* fm, FieldPtrs* fp, bool need_bg) {
  if (!f) {
    sdb_set_error("field is NULL");
    return SDB_ERR_ARG;
  }
  int rc = resolve_grid(&f->grid, &fm->grid);
  if (rc) return rc;
  if (f->grid.n_levels * f->grid.n_features_per_level != kEncDim) {
    sdb_set_error("field: encoding width must be %d (n_levels*n_features), got %d", kEncDim,
                  f->grid.n_levels * f->grid.n_features_per_level);
    return SDB_ERR_UNSUPPORTED;
  }
  if (need_bg) {
    rc = resolve_grid(&f->bg_grid, &fm->bg_grid);
    if (rc) return rc;
    if (f->bg_grid.n_levels != 4) {
      sdb_set_error("field: environment-map grid must have 4 levels (got %d)", f->bg_grid.n_levels);
      return SDB_ERR_UNSUPPORTED;
    }
    if (!f->bg_table || !f->bg_w1 || !f->bg_w2 || !f->bg_w3) {
      sdb_set_error("field: background pointers are NULL");
      return SDB_ERR_ARG;
    }
  } else {
    memset(&fm->bg_grid, 0, sizeof(fm->bg_grid));
  }
  if (!f->table || !f->w1_density || !f->w2_density || !f->w1_feature || !f->w2_feature) {
    sdb_set_error("field: geometry pointers are NULL");
    return SDB_ERR_ARG;
  }
  if (!(f->radius > 0.f)) {
    sdb_set_error("field: radius must be > 0");
    return SDB_ERR_ARG;
  }
  if (f->density_bias_type < 0 || f->density_bias_type > 2 || f->density_activation < 0 ||
      f->density_activation > 2 || f->color_activation < 0 || f->color_activation > 1) {
    sdb_set_error("field: unknown bias/activation enum");
    return SDB_ERR_UNSUPPORTED;
  }
  fm->radius = f->radius;
  fm->bias_type = f->density_bias_type;
  fm->bias_const = f->density_bias_const;
  fm->blob_scale = f->density_blob_scale;
  fm->blob_std = f->density_blob_std;
  fm->density_act = f->density_activation;
  fm->color_act = f->color_activation;
  fm->bg_color_act = f->bg_color_activation;
  fm->fd_eps = f->fd_normal_eps > 0.f ? f->fd_normal_eps : 0.01f;
  fp->table = f->table;
  fp->w1d = f->w1_density;
  fp->w2d = f->w2_density;
  fp->w1f = f->w1_feature;
  fp->w2f = f->w2_feature;
  fp->bg_table = f->bg_table;
  fp->bg_w1 = f->bg_w1;
  fp->bg_w2 = f->bg_w2;
  fp->bg_w3 = f->bg_w3;
  return SDB_OK;
}

static int resolve_march(const sdb_march_cfg* c, MarchMeta* m) {
  if (!c || !(c->render_step_size > 0.f)) {
    sdb_set_error("march: render_step_size must be > 0");
    return SDB_ERR_ARG;
  }
  if (c->grid_resolution < 1 || c->grid_resolution > 32) {
    sdb_set_error("march: grid_resolution must be in [1,32]");
    return SDB_ERR_UNSUPPORTED;
  }
  m->step = c->render_step_size;
  m->near_plane = c->near_plane;
  m->far_plane = c->far_plane;
  m->prune = c->prune;
  m->alpha_thre = c->alpha_thre;
  m->early_stop_eps = c->early_stop_eps;
  m->grid_res = c->grid_resolution;
  m->output_normal = c->output_normal;
  return SDB_OK;
}

int launch_hyper_field_fwd(const GridMeta&, const float*, const float*, int, int, const float*, const float*, const float*,
                           const float*, float*, float*, float*, cudaStream_t);
int launch_hyper_field_bwd(const GridMeta&, const float*, int, int, const float*, const float*, const float*,
                           const float*, const float*, const float*, const float*, float*, float*, float*, float*,
                           float*, cudaStream_t);

namespace {

__global__ void raygen_kernel(const float* __restrict__ c2w, const float* __restrict__ fovy, int B, int H, int W,
                              float* __restrict__ rays_o, float* __restrict__ rays_d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H * W) return;
  const int b = i / (H * W), rem = i % (H * W), y = rem / W, x = rem % W;
  const float focal = 0.5f * (float)H / tanf(0.5f * fovy[b]);
  const float dx = ((float)x + 0.5f - 0.5f * (float)W) / focal;
  const float dy = -((float)y + 0.5f - 0.5f * (float)H) / focal;
  const float dz = -1.f;
  const float* m = c2w + 16 * b;
  float rx = dx * m[0] + dy * m[1] + dz * m[2];
  float ry = dx * m[4] + dy * m[5] + dz * m[6];
  float rz = dx * m[8] + dy * m[9] + dz * m[10];
  const float inv = 1.f / fmaxf(sqrtf(rx * rx + ry * ry + rz * rz), 1e-12f);
  rays_d[3 * i + 0] = rx * inv;
  rays_d[3 * i + 1] = ry * inv;
  rays_d[3 * i + 2] = rz * inv;
  rays_o[3 * i + 0] = m[3];
  rays_o[3 * i + 1] = m[7];
  rays_o[3 * i + 2] = m[11];
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2_sqrt, float gscale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    float pi = p[i];
    pi *= (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// Adan (threestudio/systems/optimizers.py:200-250, _single_tensor_adan) fused into one pass; prev_g holds the previous
// (clipped) gradient, i.e. minus the reference's neg_pre_grad.
__global__ void adan_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ n, float* __restrict__ d, float* __restrict__ prev_g, long long count,
                            float lr, float b1, float b2, float b3, float eps, float wd, float bc1, float bc2,
                            float bc3_sqrt, float gscale, int first, int no_prox) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float diff = first ? 0.f : gi - prev_g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float di = b2 * d[i] + (1.f - b2) * diff;
    const float u = b2 * diff + gi;
    const float ni = b3 * n[i] + (1.f - b3) * u * u;
    m[i] = mi, d[i] = di, n[i] = ni, prev_g[i] = gi;
    const float denom = sqrtf(ni) / bc3_sqrt + eps;
    float pi = p[i];
    if (no_prox) pi *= 1.f - lr * wd;
    pi -= (lr / bc1) * (mi / denom) + (lr * b2 / bc2) * (di / denom);
    if (!no_prox) pi /= 1.f + lr * wd;
    p[i] = pi;
  }
}

}  // namespace

extern "C" {

const char* sdb_last_error(void) { return g_err; }
int sdb_abi_version(void) { return 2; }
unsigned long long sdb_launch_count(void) { return g_sdb_launch_count; }

long long sdb_grid_num_entries(const sdb_grid_cfg* cfg) {
  GridMeta gm;
  if (resolve_grid(cfg, &gm)) return -1;
  return (long long)gm.offset[gm.n_levels - 1] + gm.size[gm.n_levels - 1];
}

int sdb_grid_describe(const sdb_grid_cfg* cfg, float* scale, uint32_t* res, uint32_t* size, uint32_t* offset,
                      uint32_t* hashed) {
  GridMeta gm;
  int rc = resolve_grid(cfg, &gm);
  if (rc) return rc;
  for (int l = 0; l < gm.n_levels; ++l) {
    if (scale) scale[l] = gm.scale[l];
    if (res) res[l] = gm.res[l];
    if (size) size[l] = gm.size[l];
    if (offset) offset[l] = gm.offset[l];
    if (hashed) hashed[l] = gm.hashed[l];
  }
  return SDB_OK;
}

int sdb_hashgrid_forward(const sdb_grid_cfg* cfg, const float* table, const float* x01, int n, float* out,
                         void* stream) {
  GridMeta gm;
  int rc = resolve_grid(cfg, &gm);
  if (rc) return rc;
  SDB_CHECK_ARG(gm.n_levels == 16 || gm.n_levels == 4, "hashgrid: n_levels must be 16 or 4");
  if (n == 0) return SDB_OK;
  SDB_CHECK_ARG(table && x01 && out && n > 0, "hashgrid_forward: bad arguments");
  return launch_hashgrid_fwd(gm, table, x01, n, out, (cudaStream_t)stream);
}

int sdb_hashgrid_backward(const sdb_grid_cfg* cfg, const float* x01, const float* g_out, int n, float* g_table,
                          void* stream) {
  GridMeta gm;
  int rc = resolve_grid(cfg, &gm);
  if (rc) return rc;
  SDB_CHECK_ARG(gm.n_levels == 16 || gm.n_levels == 4, "hashgrid: n_levels must be 16 or 4");
  if (n == 0) return SDB_OK;
  SDB_CHECK_ARG(x01 && g_out && g_table && n > 0, "hashgrid_backward: bad arguments");
  return launch_hashgrid_bwd(gm, x01, g_out, n, g_table, (cudaStream_t)stream);
}

int sdb_field_forward(const sdb_field* field, const float* points, int n, float* density, float* features,
                      float* normal, void* stream) {
  FieldMeta fm;
  FieldPtrs fp;
  int rc = resolve_field(field, &fm, &fp, false);
  if (rc) return rc;
  SDB_CHECK_ARG(points && density && n >= 0, "field_forward: bad arguments");
  if (n == 0) return SDB_OK;
  return launch_field_eval(fm, fp, points, n, density, features, normal, (cudaStream_t)stream);
}

int sdb_occgrid_update(const sdb_field* field, const int* cell_idx, const float* cell_rand, int n_cells,
                       int resolution, float render_step_size, float ema_decay, float occ_thre, float* occs,
                       uint32_t* occ_bits, float* occ_mean, void* stream) {
  FieldMeta fm;
  FieldPtrs fp;
  int rc = resolve_field(field, &fm, &fp, false);
  if (rc) return rc;
  SDB_CHECK_ARG(resolution >= 1 && resolution <= 32, "occgrid: resolution must be in [1,32]");
  SDB_CHECK_ARG(occs && occ_bits && occ_mean && n_cells >= 0, "occgrid: bad arguments");
  SDB_CHECK_ARG(n_cells == 0 || (cell_idx && cell_rand), "occgrid: cell list is NULL");
  return launch_occ_update(fm, fp, cell_idx, cell_rand, n_cells, resolution, render_step_size, ema_decay, occ_thre,
                           occs, occ_bits, occ_mean, (cudaStream_t)stream);
}

int sdb_render_nerf_forward(const sdb_field* field, const sdb_march_cfg* march, const uint32_t* occ_bits,
                            const float* occ_mean, const float* rays_o, const float* rays_d, const float* jitter,
                            const float* bg_override, int n_rays, int rays_per_image, float* comp_rgb,
                            float* comp_rgb_fg, float* comp_rgb_bg, float* opacity, float* depth, float* z_variance,
                            const sdb_packed_samples* packed, int* work, void* stream) {
  FieldMeta fm;
  FieldPtrs fp;
  MarchMeta mm;
  int rc = resolve_field(field, &fm, &fp, true);
  if (rc) return rc;
  rc = resolve_march(march, &mm);
  if (rc) return rc;
  SDB_CHECK_ARG(occ_bits && rays_o && rays_d && work, "render_forward: NULL input");
  SDB_CHECK_ARG(comp_rgb && comp_rgb_fg && comp_rgb_bg && opacity && depth && z_variance,
                "render_forward: NULL output");
  SDB_CHECK_ARG(n_rays >= 0 && rays_per_image > 0, "render_forward: bad ray counts");
  if (n_rays == 0) return SDB_OK;
  RayIO io;
  memset(&io, 0, sizeof(io));
  io.rays_o = rays_o;
  io.rays_d = rays_d;
  io.jitter = jitter;
  io.bg_override = bg_override;
  io.occ_bits = occ_bits;
  io.occ_mean = occ_mean;
  io.n_rays = n_rays;
  io.rays_per_image = rays_per_image;
  io.comp_rgb = comp_rgb;
  io.comp_rgb_fg = comp_rgb_fg;
  io.comp_rgb_bg = comp_rgb_bg;
  io.opacity = opacity;
  io.depth = depth;
  io.z_variance = z_variance;
  io.work_counter = work;
  PackedOut pk;
  memset(&pk, 0, sizeof(pk));
  if (packed && packed->counter) {
    SDB_CHECK_ARG(packed->capacity >= 0 && packed->ray_indices && packed->t_starts && packed->t_ends &&
                      packed->weights && packed->density && packed->rgb,
                  "render_forward: packed sample buffers are NULL");
    pk.counter = packed->counter;
    pk.capacity = packed->capacity;
    pk.ray_idx = packed->ray_indices;
    pk.t_start = packed->t_starts;
    pk.t_end = packed->t_ends;
    pk.weight = packed->weights;
    pk.density = packed->density;
    pk.rgb = packed->rgb;
    pk.normal = packed->normal;
  }
  return launch_render_fwd(fm, fp, mm, io, pk, (cudaStream_t)stream);
}

int sdb_render_nerf_backward(const sdb_field* field, const sdb_field_grads* grads, const sdb_march_cfg* march,
                             const uint32_t* occ_bits, const float* occ_mean, const float* rays_o,
                             const float* rays_d, const float* jitter, const float* bg_override, int n_rays,
                             int rays_per_image, const float* comp_rgb_fg, const float* comp_rgb_bg,
                             const float* opacity, const float* depth, const float* g_comp_rgb,
                             const float* g_opacity, const float* g_depth, int* work, void* stream) {
  FieldMeta fm;
  FieldPtrs fp;
  MarchMeta mm;
  int rc = resolve_field(field, &fm, &fp, true);
  if (rc) return rc;
  rc = resolve_march(march, &mm);
  if (rc) return rc;
  SDB_CHECK_ARG(grads && grads->table && grads->w1_density && grads->w2_density && grads->w1_feature &&
                    grads->w2_feature && grads->bg_table && grads->bg_w1 && grads->bg_w2 && grads->bg_w3,
                "render_backward: NULL gradient buffer");
  SDB_CHECK_ARG(occ_bits && rays_o && rays_d && work && comp_rgb_fg && comp_rgb_bg && opacity && depth && g_comp_rgb,
                "render_backward: NULL input");
  SDB_CHECK_ARG(n_rays >= 0 && rays_per_image > 0, "render_backward: bad ray counts");
  if (n_rays == 0) return SDB_OK;
  RayIO io;
  memset(&io, 0, sizeof(io));
  io.rays_o = rays_o;
  io.rays_d = rays_d;
  io.jitter = jitter;
  io.bg_override = bg_override;
  io.occ_bits = occ_bits;
  io.occ_mean = occ_mean;
  io.n_rays = n_rays;
  io.rays_per_image = rays_per_image;
  io.comp_rgb_fg = const_cast<float*>(comp_rgb_fg);
  io.comp_rgb_bg = const_cast<float*>(comp_rgb_bg);
  io.opacity = const_cast<float*>(opacity);
  io.depth = const_cast<float*>(depth);
  io.g_comp_rgb = g_comp_rgb;
  io.g_opacity = g_opacity;
  io.g_depth = g_depth;
  io.work_counter = work;
  FieldGrads fg;
  fg.table = grads->table;
  fg.w1d = grads->w1_density;
  fg.w2d = grads->w2_density;
  fg.w1f = grads->w1_feature;
  fg.w2f = grads->w2_feature;
  fg.bg_table = grads->bg_table;
  fg.bg_w1 = grads->bg_w1;
  fg.bg_w2 = grads->bg_w2;
  fg.bg_w3 = grads->bg_w3;
  return launch_render_bwd(fm, fp, fg, mm, io, (cudaStream_t)stream);
}

static int resolve_tape(const sdb_render_tape* t, int n_rays, RenderTape* out) {
  SDB_CHECK_ARG(t->capacity > 0 && t->capacity % 128 == 0, "tape: capacity must be a positive multiple of 128");
  SDB_CHECK_ARG(t->capacity <= (1 << 27), "tape: capacity must be <= 2^27 slots");
  SDB_CHECK_ARG(t->max_chunks > 0, "tape: max_chunks must be > 0");
  SDB_CHECK_ARG(t->counter && t->enc && t->pos && t->sample && t->ray_chunks && t->ray_nchunks,
                "tape: NULL buffer");
  (void)n_rays;
  out->capacity = t->capacity;
  out->max_chunks = t->max_chunks;
  out->counter = t->counter;
  out->enc = t->enc;
  out->pos = t->pos;
  out->sample = t->sample;
  out->ray_chunks = t->ray_chunks;
  out->ray_nchunks = t->ray_nchunks;
  return SDB_OK;
}

int sdb_render_tape_geometry(const sdb_march_cfg* march, float radius, int n_rays, long long* capacity,
                             int* max_chunks) {
  SDB_CHECK_ARG(march && march->render_step_size > 0.f && radius > 0.f && n_rays >= 0 && capacity && max_chunks,
                "tape_geometry: bad arguments");
  // longest chord of the box, clipped by the near / far planes, on the fixed-step lattice (+ the marcher's slack)
  double chord = 2.0 * std::sqrt(3.0) * (double)radius;
  const double span = (double)march->far_plane - (double)march->near_plane;
  if (span < chord) chord = span > 0.0 ? span : 0.0;
  const long long per_ray = (long long)std::ceil(chord / (double)march->render_step_size) + 4;
  long long cap = per_ray * (long long)n_rays;
  cap = (cap + 127) / 128 * 128;
  if (cap < 128) cap = 128;
  *capacity = cap;
  *max_chunks = (int)((per_ray + 31) / 32) + 1;
  return SDB_OK;
}

int sdb_render_nerf_forward_v2(const sdb_field* field, const sdb_march_cfg* march, const uint32_t* occ_bits,
                               const float* occ_mean, const float* rays_o, const float* rays_d, const float* jitter,
                               const float* bg_override, int n_rays, int rays_per_image, float* comp_rgb,
                               float* comp_rgb_fg, float* comp_rgb_bg, float* opacity, float* depth,
                               float* z_variance, const sdb_render_tape* tape, int* work, void* stream) {
  FieldMeta fm;
  FieldPtrs fp;
  MarchMeta mm;
  int rc = resolve_field(field, &fm, &fp, true);
  if (rc) return rc;
  rc = resolve_march(march, &mm);
  if (rc) return rc;
  SDB_CHECK_ARG(occ_bits && rays_o && rays_d && work, "render_forward_v2: NULL input");
  SDB_CHECK_ARG(comp_rgb && comp_rgb_fg && comp_rgb_bg && opacity && depth && z_variance,
                "render_forward_v2: NULL output");
  SDB_CHECK_ARG(n_rays >= 0 && rays_per_image > 0, "render_forward_v2: bad ray counts");
  if (n_rays == 0) return SDB_OK;
  RayIO io;
  memset(&io, 0, sizeof(io));
  io.rays_o = rays_o;
  io.rays_d = rays_d;
  io.jitter = jitter;
  io.bg_override = bg_override;
  io.occ_bits = occ_bits;
  io.occ_mean = occ_mean;
  io.n_rays = n_rays;
  io.rays_per_image = rays_per_image;
  io.comp_rgb = comp_rgb;
  io.comp_rgb_fg = comp_rgb_fg;
  io.comp_rgb_bg = comp_rgb_bg;
  io.opacity = opacity;
  io.depth = depth;
  io.z_variance = z_variance;
  io.work_counter = work;
  RenderTape t;
  if (tape) {
    rc = resolve_tape(tape, n_rays, &t);
    if (rc) return rc;
  }
  return launch_render_fwd2(fm, fp, mm, io, tape ? &t : nullptr, (cudaStream_t)stream);
}

int sdb_render_nerf_backward_tape(const sdb_field* field, const sdb_field_grads* grads, const sdb_march_cfg* march,
                                  const float* rays_d, const float* bg_override, int n_rays, int rays_per_image,
                                  const float* comp_rgb_fg, const float* comp_rgb_bg, const float* opacity,
                                  const float* depth, const float* g_comp_rgb, const float* g_opacity,
                                  const float* g_depth, const sdb_render_tape* tape, void* stream) {
  return sdb_render_nerf_backward_tape_zv(field, grads, march, rays_d, bg_override, n_rays, rays_per_image, comp_rgb_fg,
                                          comp_rgb_bg, opacity, depth, nullptr, g_comp_rgb, g_opacity, g_depth, nullptr,
                                          tape, stream);
}

int sdb_render_nerf_backward_tape_zv(const sdb_field* field, const sdb_field_grads* grads, const sdb_march_cfg* march,
                                     const float* rays_d, const float* bg_override, int n_rays, int rays_per_image,
                                     const float* comp_rgb_fg, const float* comp_rgb_bg, const float* opacity,
                                     const float* depth, const float* z_variance, const float* g_comp_rgb,
                                     const float* g_opacity, const float* g_depth, const float* g_z_variance,
                                     const sdb_render_tape* tape, void* stream) {
  SDB_CHECK_ARG(!g_z_variance || z_variance, "render_backward_tape: g_z_variance needs the forward's z_variance");
  FieldMeta fm;
  FieldPtrs fp;
  MarchMeta mm;
  int rc = resolve_field(field, &fm, &fp, true);
  if (rc) return rc;
  rc = resolve_march(march, &mm);
  if (rc) return rc;
  SDB_CHECK_ARG(grads && grads->table && grads->w1_density && grads->w2_density && grads->w1_feature &&
                    grads->w2_feature && grads->bg_table && grads->bg_w1 && grads->bg_w2 && grads->bg_w3,
                "render_backward_tape: NULL gradient buffer");
  SDB_CHECK_ARG(rays_d && comp_rgb_fg && comp_rgb_bg && opacity && depth && g_comp_rgb && tape,
                "render_backward_tape: NULL input");
  SDB_CHECK_ARG(n_rays >= 0 && rays_per_image > 0, "render_backward_tape: bad ray counts");
  if (n_rays == 0) return SDB_OK;
  RenderTape t;
  rc = resolve_tape(tape, n_rays, &t);
  if (rc) return rc;
  RayIO io;
  memset(&io, 0, sizeof(io));
  io.rays_d = rays_d;
  io.bg_override = bg_override;
  io.n_rays = n_rays;
  io.rays_per_image = rays_per_image;
  io.comp_rgb_fg = const_cast<float*>(comp_rgb_fg);
  io.comp_rgb_bg = const_cast<float*>(comp_rgb_bg);
  io.opacity = const_cast<float*>(opacity);
  io.depth = const_cast<float*>(depth);
  io.g_comp_rgb = g_comp_rgb;
  io.g_opacity = g_opacity;
  io.g_depth = g_depth;
  io.z_variance = const_cast<float*>(z_variance);
  io.g_z_variance = g_z_variance;
  FieldGrads fg;
  fg.table = grads->table;
  fg.w1d = grads->w1_density;
  fg.w2d = grads->w2_density;
  fg.w1f = grads->w1_feature;
  fg.w2f = grads->w2_feature;
  fg.bg_table = grads->bg_table;
  fg.bg_w1 = grads->bg_w1;
  fg.bg_w2 = grads->bg_w2;
  fg.bg_w3 = grads->bg_w3;
  return launch_render_bwd2(fm, fp, fg, mm, io, t, (cudaStream_t)stream);
}

static int orient_common(const sdb_field* field, const sdb_render_tape* tape, int n_rays, FieldMeta* fm, FieldPtrs* fp,
                         RenderTape* t) {
  int rc = resolve_field(field, fm, fp, true);
  if (rc) return rc;
  SDB_CHECK_ARG(tape && n_rays >= 0, "render_orient: bad arguments");
  SDB_CHECK_ARG(fm->fd_eps > 0.f, "render_orient: finite_difference_normal_eps must be > 0");
  return resolve_tape(tape, n_rays, t);
}

int sdb_render_orient_forward(const sdb_field* field, const float* rays_d, int n_rays, const sdb_render_tape* tape,
                              float* orient, float* og, void* stream) {
  FieldMeta fm;
  FieldPtrs fp;
  RenderTape t;
  int rc = orient_common(field, tape, n_rays, &fm, &fp, &t);
  if (rc) return rc;
  SDB_CHECK_ARG(rays_d && orient && og, "render_orient_forward: NULL buffer");
  if (n_rays == 0) return SDB_OK;
  return launch_render_orient_fwd(fm, fp, rays_d, n_rays, t, orient, og, (cudaStream_t)stream);
}

int sdb_render_orient_backward(const sdb_field* field, const sdb_field_grads* grads, int n_rays,
                               const sdb_render_tape* tape, float* og, const float* g_orient, void* stream) {
  FieldMeta fm;
  FieldPtrs fp;
  RenderTape t;
  int rc = orient_common(field, tape, n_rays, &fm, &fp, &t);
  if (rc) return rc;
  SDB_CHECK_ARG(grads && grads->table && grads->w1_density && grads->w2_density && og && g_orient,
                "render_orient_backward: NULL buffer");
  if (n_rays == 0) return SDB_OK;
  FieldGrads fg;
  memset(&fg, 0, sizeof(fg));
  fg.table = grads->table;
  fg.w1d = grads->w1_density;
  fg.w2d = grads->w2_density;
  return launch_render_orient_bwd(fm, fp, fg, n_rays, t, og, g_orient, (cudaStream_t)stream);
}

long long sdb_hyper_field_tape_floats(int n_prompts, int n_points) {
  const long long n_pad = ((long long)n_points + 127) / 128 * 128;
  return (long long)n_prompts * n_pad * kEncDim;
}

int sdb_hyper_field_forward(const sdb_grid_cfg* grid, const float* table, const float* points01, int n_prompts,
                            int n_points, const float* w1_a, const float* w2_a, const float* w1_b, const float* w2_b,
                            float* out_a, float* out_b, float* tape, void* stream) {
  GridMeta gm;
  int rc = resolve_grid(grid, &gm);
  if (rc) return rc;
  SDB_CHECK_ARG(gm.n_levels * 2 == kEncDim, "hyper_field: the encoding must be 16 levels x 2 features");
  SDB_CHECK_ARG(table && points01 && n_prompts > 0 && n_points >= 0, "hyper_field_forward: bad arguments");
  SDB_CHECK_ARG((w1_a && w2_a && out_a) || (w1_b && w2_b && out_b), "hyper_field_forward: no head requested");
  SDB_CHECK_ARG((!w1_a || (w2_a && out_a)) && (!w1_b || (w2_b && out_b)), "hyper_field_forward: incomplete head");
  if (n_points == 0) return SDB_OK;
  return launch_hyper_field_fwd(gm, table, points01, n_prompts, n_points, w1_a, w2_a, w1_b, w2_b, out_a, out_b, tape,
                                (cudaStream_t)stream);
}

int sdb_hyper_field_backward(const sdb_grid_cfg* grid, const float* points01, int n_prompts, int n_points,
                             const float* w1_a, const float* w2_a, const float* w1_b, const float* w2_b,
                             const float* tape, const float* d_out_a, const float* d_out_b, float* g_table,
                             float* g_w1_a, float* g_w2_a, float* g_w1_b, float* g_w2_b, void* stream) {
  GridMeta gm;
  int rc = resolve_grid(grid, &gm);
  if (rc) return rc;
  SDB_CHECK_ARG(gm.n_levels * 2 == kEncDim, "hyper_field: the encoding must be 16 levels x 2 features");
  SDB_CHECK_ARG(points01 && tape && g_table && n_prompts > 0 && n_points >= 0, "hyper_field_backward: bad arguments");
  SDB_CHECK_ARG(!d_out_a || (w1_a && w2_a && g_w1_a && g_w2_a), "hyper_field_backward: head a is incomplete");
  SDB_CHECK_ARG(!d_out_b || (w1_b && w2_b && g_w1_b && g_w2_b), "hyper_field_backward: head b is incomplete");
  if (n_points == 0 || (!d_out_a && !d_out_b)) return SDB_OK;
  return launch_hyper_field_bwd(gm, points01, n_prompts, n_points, w1_a, w2_a, w1_b, w2_b, tape, d_out_a, d_out_b,
                                g_table, g_w1_a, g_w2_a, g_w1_b, g_w2_b, (cudaStream_t)stream);
}

int sdb_raygen(const float* c2w, const float* fovy, int n_images, int height, int width, float* rays_o,
               float* rays_d, void* stream) {
  SDB_CHECK_ARG(c2w && fovy && rays_o && rays_d && n_images >= 0 && height > 0 && width > 0, "raygen: bad arguments");
  const int n = n_images * height * width;
  if (n == 0) return SDB_OK;
  raygen_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c2w, fovy, n_images, height, width, rays_o, rays_d);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("raygen");
  return SDB_OK;
}

int sdb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                   void* stream) {
  SDB_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "adamw_step: bad arguments");
  if (n == 0) return SDB_OK;
  const float bc1 = 1.f - (float)std::pow((double)beta1, (double)step);
  const float bc2 = 1.f - (float)std::pow((double)beta2, (double)step);
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)kNumSMs * 8);
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                       weight_decay, bc1, std::sqrt(bc2), grad_scale);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("adamw_step");
  return SDB_OK;
}

int sdb_adan_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* exp_avg_diff,
                  float* prev_grad, long long n, float lr, float beta1, float beta2, float beta3, float eps,
                  float weight_decay, int step, float grad_scale, int no_prox, void* stream) {
  SDB_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && exp_avg_diff && prev_grad && n >= 0 && step >= 1,
                "adan_step: bad arguments");
  if (n == 0) return SDB_OK;
  const float bc1 = 1.f - (float)std::pow((double)beta1, (double)step);
  const float bc2 = 1.f - (float)std::pow((double)beta2, (double)step);
  const float bc3 = 1.f - (float)std::pow((double)beta3, (double)step);
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)kNumSMs * 8);
  adan_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, exp_avg_diff, prev_grad, n, lr,
                                                      beta1, beta2, beta3, eps, weight_decay, bc1, bc2, std::sqrt(bc3),
                                                      grad_scale, step == 1 ? 1 : 0, no_prox);
  SDB_COUNT_LAUNCH();
  SDB_CHECK_LAUNCH("adan_step");
  return SDB_OK;
}

}  // extern "C"
