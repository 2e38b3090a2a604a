// Parameter blocks shared by the fused NeRF render kernels and their host wrappers.
#pragma once
#include <stdint.h>
#include "common.cuh"

static constexpr int kMaxLevels = 16;
static constexpr int kEncDim = 32;     // n_levels * n_features of the field encoding (16 x 2)
static constexpr int kHidden = 64;     // VanillaMLP n_neurons for density / feature nets
static constexpr int kBgEncDim = 8;    // background grid: 4 levels x 2 features
static constexpr int kBgHidden = 16;   // background MLP: 8 -> 16 -> 16 -> 3

// Resolved multiresolution-hash-grid geometry (tiny-cuda-nn "HashGrid" semantics, see
// oracle/render_oracle.py::grid_meta for the restated rules).
struct GridMeta {
  int n_levels;
  float scale[kMaxLevels];
  uint32_t res[kMaxLevels];
  uint32_t size[kMaxLevels];    // entries in this level
  uint32_t offset[kMaxLevels];  // first entry of this level
  uint32_t hashed[kMaxLevels];  // 1 = spatial hash, 0 = dense
};

struct FieldMeta {
  GridMeta grid;
  GridMeta bg_grid;
  float radius;
  int bias_type;        // 0 const, 1 blob_magic3d, 2 blob_dreamfusion
  float bias_const;
  float blob_scale;
  float blob_std;
  int density_act;      // 0 softplus, 1 exp, 2 trunc_exp
  int color_act;        // 0 sigmoid, 1 sigmoid-mipnerf
  int bg_color_act;     // 0 sigmoid, 1 sigmoid-mipnerf
  float fd_eps;         // finite-difference normal epsilon
};

struct FieldPtrs {
  const float* table;     // [entries, 2]
  const float* w1d;       // [64, 32]
  const float* w2d;       // [1, 64]
  const float* w1f;       // [64, 32]
  const float* w2f;       // [3, 64]
  const float* bg_table;  // [bg_entries, 2]
  const float* bg_w1;     // [16, 8]
  const float* bg_w2;     // [16, 16]
  const float* bg_w3;     // [3, 16]
};

struct FieldGrads {
  float* table;
  float* w1d;
  float* w2d;
  float* w1f;
  float* w2f;
  float* bg_table;
  float* bg_w1;
  float* bg_w2;
  float* bg_w3;
};

struct MarchMeta {
  float step;            // render_step_size
  float near_plane;
  float far_plane;
  int prune;             // 1: alpha / transmittance visibility pruning (sigma_fn path)
  float alpha_thre;      // 0.01 in the reference; clipped by *occ_mean on device
  float early_stop_eps;  // 1e-4
  int grid_res;          // occupancy grid resolution (32)
  int output_normal;     // compute finite-difference normals for packed output
};

struct PackedOut {
  int* counter;      // [1] number of kept samples (may exceed capacity; writes are clipped)
  int capacity;
  int* ray_idx;      // [cap]
  float* t_start;    // [cap]
  float* t_end;      // [cap]
  float* weight;     // [cap]
  float* density;    // [cap]
  float* rgb;        // [cap,3]  colour after activation
  float* normal;     // [cap,3]  (only if output_normal)
};

struct RayIO {
  const float* rays_o;       // [Nr,3]
  const float* rays_d;       // [Nr,3]
  const float* jitter;       // [Nr] or null
  const float* bg_override;  // [B,3] or null
  const uint32_t* occ_bits;  // [res^3/32]
  const float* occ_mean;     // [1] or null
  int n_rays;
  int rays_per_image;
  // forward outputs / backward saved tensors
  float* comp_rgb;      // [Nr,3]
  float* comp_rgb_fg;   // [Nr,3]
  float* comp_rgb_bg;   // [Nr,3]
  float* opacity;       // [Nr]
  float* depth;         // [Nr]
  float* z_variance;    // [Nr]
  // backward inputs
  const float* g_comp_rgb;  // [Nr,3]
  const float* g_opacity;   // [Nr] or null
  const float* g_depth;     // [Nr] or null
  const float* g_z_variance;  // [Nr] or null (tape backward only; needs z_variance)
  int* work_counter;        // [1] zeroed before launch
};
