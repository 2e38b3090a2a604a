/*
 * sdb200_asd.h — C ABI of the Asynchronous-Score-Distillation glue kernels of libsdb200.so: image resize into
 * the VAE, view-dependent / Perp-Neg text-embedding assembly, posterior sampling + q-sample (prologue) and the
 * CFG / Perp-Neg / w(t) score-gradient with its hand-derived backward into the VAE moments (epilogue).
 * Each entry point cites the reference lines it replaces (paths relative to the reference tree). Device pointers,
 * caller-owned buffers, cudaStream_t `stream`, 0 / negative return as in sdb200.h. Images and latents are
 * channels-last fp32: [B, H, W, C].
 */
#ifndef SDB200_ASD_H
#define SDB200_ASD_H

#ifdef __cplusplus
extern "C" {
#endif

/* y = F.interpolate(x, (H,W), mode="bilinear", align_corners=False) * scale + shift
 * (stable_diffusion_asd_guidance.py:204-206 + :175 `imgs * 2 - 1`; mvdream_asd_guidance.py:133-135 + :108). */
int sdb_resize_bilinear_forward(const float* x, int batch, int h, int w, int c, float* y, int H, int W, float scale,
                                float shift, void* stream);
/* d_x = scale * interpolate^T(d_y) (the autograd of the above); d_x is overwritten. */
int sdb_resize_bilinear_backward(const float* d_y, int batch, int h, int w, int c, float* d_x, int H, int W,
                                 float scale, void* stream);

typedef struct {
  int view_dependent;        /* 0: row 0 of the tables for every sample (view_dependent_prompting=False) */
  int perp_neg;              /* 1: PromptProcessorOutput.get_text_embeddings_perp_neg */
  float front_threshold, back_threshold, overhead_threshold; /* degrees (prompt_processors/base.py:189-191) */
  float f_sb[3], f_fsb[3], f_fs[3], f_sf[3];                 /* a*exp(-b*r)+c coefficients (:199-207) */
  float neg_scale;           /* multiplies the negative-prompt weights: -1 * guidance_perp_neg (asd guidance :357) */
} sdb_prompt_cfg;

/* Builds the UNet context batch on device with no host sync.
 * emb_vd / uncond_vd: fp16 [4, tokens, dim] in direction order side, front, back, overhead
 * (prompt_processors/base.py:107-110). elevation / azimuth: fp32 [B] degrees.
 * ctx out, fp16: perp_neg ? [vd(B), uncond(B), neg(2B, sample-major), vd(B)] : [vd(B), uncond(B), vd(B)]
 * (stable_diffusion_asd_guidance.py:376-382). neg_weights out: fp32 [B,2] (already times neg_scale) or NULL.
 * Replaces prompt_processors/base.py:53-167 (per-sample Python loop with .item() syncs). */
int sdb_asd_text_embeddings(const sdb_prompt_cfg* cfg, const void* emb_vd, const void* uncond_vd,
                            const float* elevation, const float* azimuth, int batch, int tokens, int dim, void* ctx,
                            float* neg_weights, void* stream);
/* Multi-prompt batches (custom/amortized/models/prompt_processors/base.py:434-568): emb_tables is the stacked
 * [n_prompts, 4 (or 1 when not view dependent), tokens, dim] table of this rank's prompt library and prompt_idx [B]
 * (device, int32) names the prompt of every sample; everything else as above. */
int sdb_asd_text_embeddings_multi(const sdb_prompt_cfg* cfg, const void* emb_tables, const void* uncond_vd,
                                  const int* prompt_idx, int n_prompts, const float* elevation, const float* azimuth,
                                  int batch, int tokens, int dim, void* ctx, float* neg_weights, void* stream);

/* moments = quant_conv(h) ; z = (mean + exp(0.5*clamp(logvar,-30,20)) * eps_post) * scaling_factor
 * (autoencoder.py:81-85, distributions.py:24-37, interface.py:108-111 / vae.config.scaling_factor);
 * x_t = sqrt(ac[t]) z + sqrt(1-ac[t]) noise for t and t_plus (scheduler.add_noise :242-246 / q_sample
 * interface.py:91-94); unet_x = [x_t] * num_repeats ++ [x_{t+}] as fp16, unet_t likewise as fp32.
 * h: fp32 [B,HW,8]; quant_w fp32 [8,8] (out,in), quant_b [8]; eps_post, noise: fp32 [B,HW,4]; t, t_plus: int32 [B]. */
int sdb_asd_prologue(const float* h, const float* quant_w, const float* quant_b, const float* eps_post,
                     const float* noise, const int* t, const int* t_plus, const float* alphas_cumprod,
                     float scaling_factor, int batch, int hw, int num_repeats, float* latents, void* unet_x,
                     float* unet_t, void* stream);

/* eps: fp32 [(num_repeats+1)*B, HW, 4] UNet output in the prologue's batch order.
 * eps_hat = e_u + gs * ((e_c - e_u) + sum_i w_i perp(e_neg_i - e_u, e_c - e_u))   (:405-428; perp: utils/ops.py:501-511)
 * grad = w(t) * (eps_hat - e_second), nan_to_num, optional clamp (:261-277); loss = 0.5 * sum(grad^2) / B (:283);
 * d_h = loss_scale * dLoss/dh through z (dz = grad / B) -- the hand-written backward of the prologue.
 * weighting: 0 sds (1 - ac[t]), 1 uniform, 2 fantasia3d. neg_weights NULL = plain CFG (MVDream :279-283).
 * Outputs: grad fp32 [B,HW,4], d_h fp32 [B,HW,8], loss [1], grad_norm [1]. */
int sdb_asd_epilogue(const float* eps, const float* h, const float* quant_w, const float* quant_b,
                     const float* eps_post, const int* t, const float* alphas_cumprod, const float* neg_weights,
                     float guidance_scale, int weighting, float grad_clip, float scaling_factor, float loss_scale,
                     int batch, int hw, int num_repeats, float* grad, float* d_h, float* loss, float* grad_norm,
                     void* stream);

/* t_plus = clamp(t + floor(clamp(plus_ratio*(t - min_step), 0, T-1-t) * u), 1, T-1) (second get_t_plus definition,
 * stable_diffusion_asd_guidance.py:294-316; mvdream_asd_guidance.py:141-164). u: fp32 [B] in [0,1) or NULL (=1). */
int sdb_asd_t_plus(const int* t, const float* u, int batch, float plus_ratio, int min_step, int num_train_timesteps,
                   int* t_plus, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDB200_ASD_H */
