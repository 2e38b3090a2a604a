/*
 * sdb200.h — C ABI of the B200-native Asynchronous-Score-Distillation step library (libsdb200.so).
 *
 * The reference (theEricMa/ScaleDreamer) has no FFI of its own: its hot path calls un-vendored Python
 * extensions (nerfacc, tiny-cuda-nn, diffusers/cuDNN). Each entry point below is what a threestudio plugin
 * would bind instead of those calls; the reference call site it replaces is cited as file:line relative to
 * the reference tree. All pointers are DEVICE pointers unless stated otherwise; all buffers are owned by the
 * caller (torch-allocated); every function enqueues work on `stream` (a cudaStream_t passed as void*) and
 * returns 0 on success or a negative error code (message via sdb_last_error()). No function synchronises.
 */
#ifndef SDB200_H
#define SDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_OK 0
#define SDB_ERR_ARG -1
#define SDB_ERR_CUDA -2
#define SDB_ERR_UNSUPPORTED -3

/* ---- library ---------------------------------------------------------------------------------------- */
const char* sdb_last_error(void);            /* thread-local message of the last failing call */
int sdb_abi_version(void);                   /* bumped on any signature change */
unsigned long long sdb_launch_count(void);   /* kernels launched by this library so far (host counter) */

/* ---- multiresolution hash grid (tiny-cuda-nn "HashGrid"; threestudio/models/networks.py:55-64) ------- */
typedef struct {
  int n_levels;              /* 16 for the field, 4 for the environment map */
  int n_features_per_level;  /* must be 2 */
  int log2_hashmap_size;
  int base_resolution;
  float per_level_scale;
} sdb_grid_cfg;

/* Number of (n_features-wide) entries the grid owns; host-only helper. */
long long sdb_grid_num_entries(const sdb_grid_cfg* cfg);
/* Per-level resolved geometry, host-only helper: fills res/size/offset/hashed[n_levels] and scale[n_levels]. */
int sdb_grid_describe(const sdb_grid_cfg* cfg, float* scale, uint32_t* res, uint32_t* size, uint32_t* offset,
                      uint32_t* hashed);

/* out[n, 2*n_levels] = encode(x01[n,3]); replaces tcnn.Encoding.forward (networks.py:63-64). */
int sdb_hashgrid_forward(const sdb_grid_cfg* cfg, const float* table, const float* x01, int n, float* out,
                         void* stream);
/* g_table += scatter(g_out); replaces tcnn's kernel_grid_backward. g_table must be pre-zeroed by the caller. */
int sdb_hashgrid_backward(const sdb_grid_cfg* cfg, const float* x01, const float* g_out, int n, float* g_table,
                          void* stream);

/* ---- iNGP field = hash grid + two bias-free 32-64-{1,3} ReLU MLPs + environment map ------------------ */
typedef struct {
  /* geometry "implicit-volume" (threestudio/models/geometry/implicit_volume.py:19-207) */
  sdb_grid_cfg grid;
  const float* table;        /* geometry.encoding params, [entries, 2] fp32 */
  const float* w1_density;   /* geometry.density_network.layers.0.weight [64,32] */
  const float* w2_density;   /* geometry.density_network.layers.2.weight [1,64]  */
  const float* w1_feature;   /* geometry.feature_network.layers.0.weight [64,32] */
  const float* w2_feature;   /* geometry.feature_network.layers.2.weight [3,64]  */
  float radius;
  int density_bias_type;     /* 0 const, 1 blob_magic3d, 2 blob_dreamfusion */
  float density_bias_const;
  float density_blob_scale;
  float density_blob_std;
  int density_activation;    /* 0 softplus, 1 exp, 2 trunc_exp */
  float fd_normal_eps;       /* finite_difference_normal_eps */
  /* material "no-material" (threestudio/models/materials/no_material.py:41-54) */
  int color_activation;      /* 0 sigmoid, 1 sigmoid-mipnerf */
  /* background "neural-environment-map-background" (…/background/neural_environment_map_background.py:46-67) */
  sdb_grid_cfg bg_grid;
  const float* bg_table;     /* background.encoding params */
  const float* bg_w1;        /* background.network.layers.0.weight [16,8]  */
  const float* bg_w2;        /* background.network.layers.2.weight [16,16] */
  const float* bg_w3;        /* background.network.layers.4.weight [3,16]  */
  int bg_color_activation;
} sdb_field;

typedef struct {
  float* table;
  float* w1_density;
  float* w2_density;
  float* w1_feature;
  float* w2_feature;
  float* bg_table;
  float* bg_w1;
  float* bg_w2;
  float* bg_w3;
} sdb_field_grads;            /* all accumulate (+=); caller zeroes */

/* geometry.forward / forward_density on arbitrary points (implicit_volume.py:109-207).
 * features / normal may be NULL. features are pre-activation (the material applies the colour activation). */
int sdb_field_forward(const sdb_field* field, const float* points, int n, float* density, float* features,
                      float* normal, void* stream);

/* ---- occupancy grid (nerfacc.OccGridEstimator, nerf_volume_renderer.py:60-65,430-444) ---------------- */
/* occs[cell] = max(occs[cell]*decay, sigma(x_cell)*step) for the listed cells (x = cell corner + rand*cell),
 * then bits = occs > min(mean(occs), occ_thre) and *occ_mean = mean(occs). n_cells may be 0 (binarize only). */
int sdb_occgrid_update(const sdb_field* field, const int* cell_idx, const float* cell_rand, int n_cells,
                       int resolution, float render_step_size, float ema_decay, float occ_thre, float* occs,
                       uint32_t* occ_bits, float* occ_mean, void* stream);

/* ---- fused renderer (NeRFVolumeRenderer.forward, nerf_volume_renderer.py:118-428) -------------------- */
typedef struct {
  float render_step_size;   /* 1.732*2*radius/num_samples_per_ray (:66-68) */
  float near_plane;
  float far_plane;
  int prune;                /* grid_prune && prune_alpha_threshold: visibility pruning by the sigma pass */
  float alpha_thre;         /* 0.01 (:177); clipped on device by *occ_mean as nerfacc does */
  float early_stop_eps;     /* nerfacc default 1e-4 */
  int grid_resolution;      /* 32 */
  int output_normal;        /* material.requires_normal: fill packed normal */
} sdb_march_cfg;

typedef struct {
  int* counter;             /* [1] out: number of kept samples */
  int capacity;
  int* ray_indices;         /* [cap] */
  float* t_starts;          /* [cap] */
  float* t_ends;            /* [cap] */
  float* weights;           /* [cap] */
  float* density;           /* [cap] */
  float* rgb;               /* [cap,3] */
  float* normal;            /* [cap,3] or NULL */
} sdb_packed_samples;         /* optional training extras (:375-386); order is not sorted by ray */

/* rays_o/rays_d [n_rays,3]; jitter [n_rays] in [0,1) or NULL (eval); bg_override [n_images,3] or NULL;
 * occ_bits [res^3/32]; occ_mean [1] or NULL; work [1] int scratch.
 * Outputs comp_rgb/comp_rgb_fg/comp_rgb_bg [n_rays,3], opacity/depth/z_variance [n_rays]. packed may be NULL. */
int sdb_render_nerf_forward(const sdb_field* field, const sdb_march_cfg* march, const uint32_t* occ_bits,
                            const float* occ_mean, const float* rays_o, const float* rays_d, const float* jitter,
                            const float* bg_override, int n_rays, int rays_per_image, float* comp_rgb,
                            float* comp_rgb_fg, float* comp_rgb_bg, float* opacity, float* depth, float* z_variance,
                            const sdb_packed_samples* packed, int* work, void* stream);

/* Backward of the above w.r.t. every field parameter. g_opacity / g_depth may be NULL.
 * comp_rgb_fg / comp_rgb_bg / opacity / depth are the tensors the forward produced. */
int sdb_render_nerf_backward(const sdb_field* field, const sdb_field_grads* grads, const sdb_march_cfg* march,
                             const uint32_t* occ_bits, const float* occ_mean, const float* rays_o,
                             const float* rays_d, const float* jitter, const float* bg_override, int n_rays,
                             int rays_per_image, const float* comp_rgb_fg, const float* comp_rgb_bg,
                             const float* opacity, const float* depth, const float* g_comp_rgb,
                             const float* g_opacity, const float* g_depth, int* work, void* stream);

/* ---- tape-based renderer (v2): same maths, the forward records every kept sample so the backward neither
 * re-marches nor re-gathers the hash grid. ------------------------------------------------------------- */
typedef struct {
  int capacity;             /* sample slots, multiple of 128 (worst case: sdb_render_tape_geometry) */
  int max_chunks;           /* stride of ray_chunks */
  int* counter;             /* [2] out: kept samples, overflow flag (non-zero: enlarge the tape and re-run) */
  float* enc;               /* [capacity*32] tile-transposed encodings: [capacity/32][32 features][32 slots] */
  float* pos;               /* [3*capacity] x01,y01,z01 planes */
  float* sample;            /* [8*capacity] raw, o0, o1, o2, w, T*exp(-sigma*delta), t_mid, delta planes;
                               the backward overwrites planes 0..3 with d raw, d o0..2 */
  uint32_t* ray_chunks;     /* [n_rays*max_chunks] (slot0 << 5) | (count-1), front to back */
  int* ray_nchunks;         /* [n_rays] */
} sdb_render_tape;

/* Host-only: worst-case tape geometry for n_rays rays through the [-radius, radius]^3 box. */
int sdb_render_tape_geometry(const sdb_march_cfg* march, float radius, int n_rays, long long* capacity,
                             int* max_chunks);

/* Same contract as sdb_render_nerf_forward (no packed extras); tape may be NULL (no gradient wanted). */
int sdb_render_nerf_forward_v2(const sdb_field* field, const sdb_march_cfg* march, const uint32_t* occ_bits,
                               const float* occ_mean, const float* rays_o, const float* rays_d, const float* jitter,
                               const float* bg_override, int n_rays, int rays_per_image, float* comp_rgb,
                               float* comp_rgb_fg, float* comp_rgb_bg, float* opacity, float* depth,
                               float* z_variance, const sdb_render_tape* tape, int* work, void* stream);

/* Backward of sdb_render_nerf_forward_v2 from its tape (two launches: per-ray compositing gradient +
 * environment map, then per-sample MLP backward + hash-grid scatter). Gradients accumulate (+=). */
int sdb_render_nerf_backward_tape(const sdb_field* field, const sdb_field_grads* grads, const sdb_march_cfg* march,
                                  const float* rays_d, const float* bg_override, int n_rays, int rays_per_image,
                                  const float* comp_rgb_fg, const float* comp_rgb_bg, const float* opacity,
                                  const float* depth, const float* g_comp_rgb, const float* g_opacity,
                                  const float* g_depth, const sdb_render_tape* tape, void* stream);

/* Same, with the gradient of the z-variance output (HiFA loss, threestudio/systems/scaledreamer.py:93-102 on
 * nerf_volume_renderer.py:335-349): d z_variance / d w_i = ((t_i - zbar)^2 - z_variance) / opacity on rays with
 * opacity > 0.5. z_variance / g_z_variance [n_rays] may both be NULL (then identical to the call above). */
int sdb_render_nerf_backward_tape_zv(const sdb_field* field, const sdb_field_grads* grads, const sdb_march_cfg* march,
                                     const float* rays_d, const float* bg_override, int n_rays, int rays_per_image,
                                     const float* comp_rgb_fg, const float* comp_rgb_bg, const float* opacity,
                                     const float* depth, const float* z_variance, const float* g_comp_rgb,
                                     const float* g_opacity, const float* g_depth, const float* g_z_variance,
                                     const sdb_render_tape* tape, void* stream);

/* Orientation term on the taped samples of sdb_render_nerf_forward_v2 (call it before the backward, which overwrites
 * tape plane 0): orient[ray] = sum_i w_i relu(n_i . d_ray)^2 with the finite-difference normals
 * n = normalize(-(sigma(clamp(x + eps e_k)) - sigma(x)) / eps) of threestudio/models/geometry/implicit_volume.py:137-177;
 * the numerator of loss_orient (threestudio/systems/scaledreamer.py:70-80; the weights are detached there, so only the
 * density network and the table receive a gradient). og [4*capacity] receives d term / d raw density at the sample
 * and its three offset points. */
int sdb_render_orient_forward(const sdb_field* field, const float* rays_d, int n_rays, const sdb_render_tape* tape,
                              float* orient, float* og, void* stream);

/* Backward of the orientation term: og is scaled by g_orient[ray] in place, then the density-network backward and the
 * hash-grid scatter run on the four points of every sample. Accumulates into grads->table / w1_density / w2_density
 * (the other members are ignored). */
int sdb_render_orient_backward(const sdb_field* field, const sdb_field_grads* grads, int n_rays,
                               const sdb_render_tape* tape, float* og, const float* g_orient, void* stream);

/* ---- hypernetwork of the amortized generators ----------------------------------------------------------------
 * LinearHyperNetwork (custom/amortized/models/geometry/hyper_iNGP.py:18-111, n_hidden_layers 1, n_neurons 64; also the
 * environment map's, custom/amortized/models/background/multiprompt_neural_environment_hashgrid_map_background.py:82-99):
 *   out[b] = w1 silu(LayerNorm(w0 x[b])) + b1,  x [n_prompts, c_dim], w0 [64, c_dim] (no bias), w1 [n_out, 64], fp32.
 * `hidden` [n_prompts, 64] keeps w0 x for the backward. n_prompts <= 64. */
int sdb_hypernet_forward(const float* x, int n_prompts, int c_dim, const float* w0, const float* ln_weight,
                         const float* ln_bias, float ln_eps, const float* w1, const float* b1, int n_out,
                         float* hidden, float* out, void* stream);
/* Host-only: floats of the backward's scratch buffer. */
long long sdb_hypernet_scratch_floats(int n_prompts, int n_out);
/* Gradients of every parameter (WRITTEN, not accumulated; fixed summation order). g_b1 may be NULL. */
int sdb_hypernet_backward(const float* x, int n_prompts, int c_dim, const float* ln_weight, const float* ln_bias,
                          float ln_eps, const float* w1, int n_out, const float* hidden, const float* d_out,
                          float* g_w0, float* g_ln_weight, float* g_ln_bias, float* g_w1, float* g_b1, float* scratch,
                          void* stream);

/* ---- prompt-conditioned hash-grid field (amortized generators) -------------------------------------------
 * out = relu(enc(x) W1[b]) W2[b], weights per prompt b from a hypernetwork, two optional heads on one encoding:
 *   head a 32->64->1: SDF of "Hyper-iNGP" (custom/amortized/models/geometry/hyper_iNGP.py:261-349, torch.bmm);
 *   head b 32->64->3: its features, or the colour of "multiprompt-neural-hashgrid-environment-map-background"
 *                     (custom/amortized/models/background/multiprompt_..._background.py:83-101).
 * points01 [n_prompts, n_points, 3] already contracted to [0,1]; w1 [n_prompts,32,64], w2 [n_prompts,64,{1,3}]
 * (the `enc @ W` layout the reference hypernetwork emits). A head is skipped when its w1 is NULL.
 * tape (optional, sdb_hyper_field_tape_floats floats) keeps the encodings for the backward. */
long long sdb_hyper_field_tape_floats(int n_prompts, int n_points);
int sdb_hyper_field_forward(const sdb_grid_cfg* grid, const float* table, const float* points01, int n_prompts,
                            int n_points, const float* w1_a, const float* w2_a, const float* w1_b, const float* w2_b,
                            float* out_a, float* out_b, float* tape, void* stream);
/* Gradients accumulate (+=) into g_table [entries,2], g_w1_* [n_prompts,32,64], g_w2_* [n_prompts,64,{1,3}].
 * d_out_a [n_prompts,n_points] / d_out_b [n_prompts,n_points,3] may be NULL (head without gradient). */
int sdb_hyper_field_backward(const sdb_grid_cfg* grid, const float* points01, int n_prompts, int n_points,
                             const float* w1_a, const float* w2_a, const float* w1_b, const float* w2_b,
                             const float* tape, const float* d_out_a, const float* d_out_b, float* g_table,
                             float* g_w1_a, float* g_w2_a, float* g_w1_b, float* g_w2_b, void* stream);

/* ---- VolSDF importance renderer (dense [n_rays, S] samples; amortized path) -------------------------------
 * Replaces ImportanceEstimator.sampling (threestudio/models/estimators.py:23-101), volsdf_density / get_alpha
 * (renderers/neus_volume_renderer.py:19-23,93-96) and the nerfacc compositing calls of
 * custom/amortized/models/renderers/generative_space_volsdf_volume_renderer.py:356-424. */
/* points[n_rays, n_coarse, 3]: mid-points of the proposal intervals with edges near + (far-near) (j + u)/(n_coarse+1). */
int sdb_volsdf_coarse_points(const float* rays_o, const float* rays_d, const float* u_coarse, int n_rays,
                             int n_coarse, float near_plane, float far_plane, float* points, void* stream);
/* sdf[n_rays, n_coarse] at those points -> t_all[n_rays, n_coarse + n_fine + 2]: proposal edges merged with the
 * n_fine + 1 inverse-CDF edges drawn at (j + u_fine)/(n_fine+1), sorted. */
int sdb_volsdf_resample(const float* sdf, const float* u_coarse, const float* u_fine, int n_rays, int n_coarse,
                        int n_fine, float near_plane, float far_plane, float inv_std, float* t_all, void* stream);
/* alpha = |delta| sigma_volsdf(sdf), w = alpha prod_{j<i}(1-alpha_j); rgb = sigmoid(features).
 * color_activation: 0 sigmoid, 1 sigmoid-mipnerf (sigmoid * 1.002 - 0.001; materials/no_material.py:41-54).
 * Outputs weights [n_rays,S], opacity / depth / z_variance [n_rays], comp_rgb_fg / comp_normal [n_rays,3]. */
int sdb_volsdf_composite_forward(const float* sdf, const float* features, const float* normal, const float* t_mid,
                                 const float* delta, int n_rays, int n_samples, float inv_std, int color_activation,
                                 float* weights,
                                 float* opacity, float* depth, float* comp_rgb_fg, float* z_variance,
                                 float* comp_normal, void* stream);
/* d_sdf [n_rays,S], d_features [n_rays,S,3] from the gradients of comp_rgb_fg / opacity / depth. */
int sdb_volsdf_composite_backward(const float* sdf, const float* features, const float* t_mid, const float* delta,
                                  const float* weights, const float* opacity, const float* depth,
                                  const float* comp_rgb_fg, const float* g_comp_rgb_fg, const float* g_opacity,
                                  const float* g_depth, int n_rays, int n_samples, float inv_std, int color_activation,
                                  float* d_sdf, float* d_features, void* stream);

/* ---- triplane feature lookup ("Triplane-transformer-sdf", custom/amortized/models/geometry/utils.py:67-97) ---------
 * planes_cl [n_prompts, 3, H, W, C] fp32 CHANNELS-LAST; points [n_prompts, n_points, 3] in [-1,1];
 * enc [n_prompts, n_points, 3*C] plane-major: plane 0 samples (x,y), plane 1 (x,z), plane 2 (z,y) with
 * F.grid_sample(bilinear, zeros padding, align_corners=False) semantics. */
int sdb_triplane_sample_forward(const float* planes_cl, const float* points, int n_prompts, int n_points, int height,
                                int width, int channels, float* enc, void* stream);
/* d_planes_cl += scatter(d_enc) (caller zeroes). */
int sdb_triplane_sample_backward(const float* d_enc, const float* points, int n_prompts, int n_points, int height,
                                 int width, int channels, float* d_planes_cl, void* stream);

/* ---- unfused ("packed") renderer stages (nerf_volume_renderer.py:139-180, 313-373) ----------------------------
 * For geometries the fused renderer does not evaluate itself (C1: frequency encoding + VanillaMLP). Samples are
 * packed and sorted by (ray, t); offsets [n_rays+1] (int64) delimit each ray's samples.
 *   sdb_march_count : counts[ray] = candidate lattice samples in occupied cells (nerfacc OccGridEstimator.sampling,
 *                     cone_angle 0, stratified jitter per ray)
 *   sdb_march_fill  : ray_indices / t_starts / t_ends / positions [n,3] at offsets = exclusive cumsum of counts
 *   sdb_packed_visibility : keep[s] = alpha >= min(alpha_thre, *occ_mean) && T >= early_stop_eps  (sigma_fn pruning,
 *                     :153-180; occ_mean may be NULL), kept_counts[ray] = kept samples of the ray
 *   sdb_packed_composite_forward : weights = T (1 - exp(-sigma dt)) (nerfacc render_weight_from_density) and the
 *                     per-ray sums opacity, depth, comp_rgb_fg [n_rays,3], z_variance (:321-373); trans [n] is kept
 *                     for the backward
 *   sdb_packed_composite_backward : d_sigma [n], d_rgb [n,3] from g_opacity / g_depth / g_comp_rgb_fg (any may be
 *                     NULL = zero) */
int sdb_march_count(const sdb_march_cfg* march, float radius, const uint32_t* occ_bits, const float* rays_o,
                    const float* rays_d, const float* jitter, int n_rays, int* counts, void* stream);
int sdb_march_fill(const sdb_march_cfg* march, float radius, const uint32_t* occ_bits, const float* rays_o,
                   const float* rays_d, const float* jitter, int n_rays, const long long* offsets, int* ray_indices,
                   float* t_starts, float* t_ends, float* positions, void* stream);
int sdb_packed_visibility(const float* sigma, const float* t_starts, const float* t_ends, const long long* offsets,
                          int n_rays, float alpha_thre, const float* occ_mean, float early_stop_eps,
                          unsigned char* keep, int* kept_counts, void* stream);
int sdb_packed_composite_forward(const float* sigma, const float* rgb, const float* t_starts, const float* t_ends,
                                 const long long* offsets, int n_rays, float* weights, float* trans, float* opacity,
                                 float* depth, float* comp_rgb_fg, float* z_variance, void* stream);
int sdb_packed_composite_backward(const float* rgb, const float* t_starts, const float* t_ends,
                                  const long long* offsets, int n_rays, const float* weights, const float* trans,
                                  const float* g_opacity, const float* g_depth, const float* g_comp_rgb_fg,
                                  float* d_sigma, float* d_rgb, void* stream);
/* ProgressiveBandFrequency (threestudio/models/networks.py:16-52) under CompositeEncoding (:170-190):
 * out[i, lead + (2f+fn)*3 + c] = fn(2^f x01[i,c]) * mask[f], fn = sin, cos; lead = 3 columns of x01*2-1 when
 * include_xyz; columns up to out_stride are zero (row padding for the MLP kernels). mask [n_frequencies] on device. */
int sdb_freq_encode(const float* x01, long long n, int n_frequencies, const float* mask, int include_xyz,
                    int out_stride, float* out, void* stream);
/* Occupancy refresh from caller-evaluated values (= sigma * step at the jittered cell points):
 * occs[cell] = max(occs[cell]*decay, value), then binarise as sdb_occgrid_update does. */
int sdb_occgrid_update_values(const int* cell_idx, const float* values, int n_cells, int resolution, float ema_decay,
                              float occ_thre, float* occs, uint32_t* occ_bits, float* occ_mean, void* stream);

/* ---- bias-free ReLU MLP d_in -> 64 -> 64 -> n_out (threestudio/models/networks.py:214-251 VanillaMLP with
 * n_neurons 64, n_hidden_layers 2: the sdf / feature heads of "Triplane-transformer-sdf",
 * custom/amortized/models/geometry/triplane_transformer_sdf.py:150-170). fp32; weights in nn.Linear layout
 * (w1 [64,d_in], w2 [64,64], w3 [n_out,64]); x [n,d_in] row-major; d_in a multiple of 8 in [8,96]; n_out 1 or 3.
 * The backward recomputes the hidden activations from x (nothing is saved by the forward). */
int sdb_mlp3_forward(const float* x, long long n, int d_in, const float* w1, const float* w2, const float* w3,
                     int n_out, float* y, void* stream);
/* g_w1/g_w2/g_w3 += weight gradients (caller zeroes); d_x [n,d_in] is written, or accumulated into when
 * accumulate_dx != 0 (two heads sharing one input); d_x may be NULL to skip the input gradient. */
int sdb_mlp3_backward(const float* x, long long n, int d_in, const float* w1, const float* w2, const float* w3,
                      int n_out, const float* d_y, float* d_x, int accumulate_dx, float* g_w1, float* g_w2,
                      float* g_w3, void* stream);

/* rays from cameras (threestudio/utils/ops.py:183-269 get_ray_directions + get_rays):
 * c2w [B,4,4], fovy [B] (radians) -> rays_o, rays_d [B,H,W,3] (normalised). */
int sdb_raygen(const float* c2w, const float* fovy, int n_images, int height, int width, float* rays_o,
               float* rays_d, void* stream);

/* ---- optimizer (threestudio/systems/utils.py:34-53 -> torch.optim.AdamW / Adam) ----------------------- */
/* One fused step over a flat fp32 parameter block. grad_scale multiplies the gradient (1/world_size after
 * the all-reduce). step is the 1-based step count for bias correction. */
int sdb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                   void* stream);

/* Adan (threestudio/systems/optimizers.py:23-315; C5's optimizer), one fused pass. prev_grad keeps the previous
 * gradient (= -neg_pre_grad of the reference); step is 1-based. */
int sdb_adan_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* exp_avg_diff,
                  float* prev_grad, long long n, float lr, float beta1, float beta2, float beta3, float eps,
                  float weight_decay, int step, float grad_scale, int no_prox, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDB200_H */
