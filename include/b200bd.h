/* libb200bd -- C-ABI of the B200-native BadDiffusion hot path (sm_100a).
 *
 * The reference (IBM/BadDiffusion) is pure Python and has no FFI of its own: the path sits behind the
 * duck-typed Python objects of SURVEY.md section 8(b).  This header is the boundary those objects'
 * drop-in replacements (package `baddiffusion_b200`) bind with ctypes; every entry point names the
 * reference code it replaces (paths relative to the reference root, D/ = diffusers/src/diffusers/).
 *
 * Conventions
 *   - every function returns BD_OK (0) or a negative BD_ERR_* code; bd_last_error() gives the message
 *     (thread-local).  Nothing here allocates device memory, synchronises, or touches a CPU fallback:
 *     work is enqueued on `stream` (a cudaStream_t passed as void*), so calls are CUDA-graph capturable.
 *   - all data pointers are DEVICE pointers owned by the caller (PyTorch allocations in practice).
 *   - images at the model boundary are fp32 NCHW exactly as in the reference; internal activations are
 *     fp16 NHWC "views": (pointer, ld) where ld = elements between consecutive pixels (>= channels), so
 *     a channel slice of a wider buffer is a view (this is how torch.cat of D/models/unet_2d_blocks.py
 *     :1726,1924 disappears: producers write straight into the concat buffer).
 *   - activation gradients are fp16 scaled by the loss scale (the reference trains under accelerate's
 *     fp16 GradScaler, baddiffusion.py:116,608); weight gradients are fp32 (scaled).
 */
#ifndef B200BD_H
#define B200BD_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BD_OK 0
#define BD_ERR_INVALID (-1)     /* bad argument / unsupported shape (ValueError in the Python layer) */
#define BD_ERR_CUDA (-2)        /* CUDA runtime / driver error                                       */
#define BD_ERR_UNSUPPORTED (-3) /* valid request this build cannot serve (NotImplementedError)       */

int bd_version(void);
/* one-time per-process/device setup (allocates the 4-byte device error word used by the tcgen05 kernels'
 * bounded mbarrier waits).  Must be called once before any CUDA-graph capture. */
int bd_init(void);
const char* bd_last_error(void);
/* non-zero if a tcgen05 kernel reported a pipeline time-out since the last call (synchronises; debug/tests) */
int bd_umma_error(void);
/* 1 if the current device is compute capability 10.x (tcgen05/TMEM/TMA paths usable) */
int bd_device_supported(void);
/* kernel launches issued by this library in this process (bench.py's `gpu_launches`) */
uint64_t bd_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * K11  batch-prep: poison blend + forward diffusion + loss target in ONE pass.
 * Replaces dataset.py:275-276,288-315 (mask/blend), D/schedulers/scheduling_ddpm.py:422-443 (add_noise)
 * and loss.py:257-285 (q_sample_diffuser).  Bit-exact with the reference's fp32 op order.
 *   img (B,C,H,W) f32, is_poison (B) u8 [nullable: then R/x0 are given explicitly via img=x0, R],
 *   trigger/target (C,H,W) f32, noise (B,C,H,W) f32 [nullable -> Philox4x32-10 N(0,1) from seed/offset],
 *   t (B) i64, alphas / alphas_cumprod (T) f32.
 *   outputs: x_noisy, eps_target (B,C,H,W) f32; noise_out nullable (the noise that was used).
 * ---------------------------------------------------------------------------------------------- */
int bd_batch_prep(const float* img, const uint8_t* is_poison, const float* trigger, const float* target,
                  const float* R_explicit, const float* noise, const int64_t* t, const float* alphas,
                  const float* alphas_cumprod, float* x_noisy, float* eps_target, float* noise_out, int B, int C,
                  int H, int W, int T, uint64_t seed, uint64_t offset,
                  const int* noise_counter /* device int added to `offset` (nullable): new noise per graph replay */,
                  void* stream);

/* K11 with the data path of SURVEY 8f n2 folded in: the batch arrives as decoded uint8 NHWC pixels (B,H,W,C);
 * transforms.ToTensor (/255), util.normalize (util.py:83-111 with vmin_in 0, vmax_in 1 -> [-1, 1-2e-5], quirk Q7) and
 * RandomHorizontalFlip (dataset.py:120-136; the per-sample coin `flip` (B) u8 is drawn by the host, nullable = never)
 * run on load, then the poison blend / add_noise / loss target of bd_batch_prep.  Bit-exact with the reference's
 * CPU transform chain.  image_out (B,C,H,W) f32 nullable: the normalised, flipped image (the DataLoader's `image`). */
int bd_batch_prep_u8(const uint8_t* img_nhwc, const uint8_t* flip, const uint8_t* is_poison, const float* trigger,
                     const float* target, const float* noise, const int64_t* t, const float* alphas,
                     const float* alphas_cumprod, float* x_noisy, float* eps_target, float* noise_out, float* image_out,
                     int B, int C, int H, int W, int T, uint64_t seed, uint64_t offset, const int* noise_counter,
                     void* stream);

/* K12  loss.py:301 F.mse_loss(target, eps_hat) and its gradient 2(eps_hat-target)/n * loss_scale.
 *   partial: workspace of >= bd_mse_workspace_floats() floats; loss: 1 float; grad nullable (f32, same
 *   layout as eps_hat); loss_scale: device pointer to 1 float (nullable -> 1). Deterministic 2-stage sum. */
size_t bd_mse_workspace_floats(void);
int bd_mse_fwd_bwd(const float* eps_hat, const float* target, float* loss, float* grad, float* partial,
                   const float* loss_scale, size_t n, void* stream);

/* K13  D/schedulers/scheduling_ddpm.py:324-420 (DDPMScheduler.step, epsilon prediction).
 * coef: DEVICE table of rows of 8 floats {sqrt_beta_prod_t, sqrt_alpha_prod_t, c0, ct, sigma, clip(<=0 off),
 * clip_defense(<=0 off), has_noise}; the host fills it with the reference's own 0-d fp32 torch expressions.
 * step_index: device int (nullable -> row 0) so one captured graph serves all 1000 steps.
 * z nullable -> Philox noise (seed, offset + (id << 32) + *step_index), id = (uint32)(has_noise - 1): a caller that
 * writes has_noise = 1 + id (id < 2^23, exact in fp32) into the table selects a fresh noise stream per pipeline call
 * without re-capturing.  In-place (x_prev == x) allowed.                                               */
int bd_ddpm_step(const float* x, const float* eps_hat, const float* z, float* x_prev, const float* coef,
                 const int* step_index, size_t n, uint64_t seed, uint64_t offset, void* stream);
/* D/schedulers/scheduling_ddim.py:261-381.  coef row: {sqrt_beta_prod_t, sqrt_alpha_prod_t,
 * sqrt_alpha_prod_prev, dir_coef=(1-a_prev-std^2)^0.5, std, clip(<=0 off), use_clipped, noise-stream id (0 default,
 * < 2^23; Philox offset += id << 32)}.                                                                 */
int bd_ddim_step(const float* x, const float* eps_hat, const float* z, float* x_prev, const float* coef,
                 const int* step_index, size_t n, uint64_t seed, uint64_t offset, void* stream);
/* PNDM step (D/schedulers/scheduling_pndm.py:215-400 = step_prk / step_plms / _get_prev_sample, plus the clamp of the
 * reference's patched PNDMPipeline): the sampler model.py:598-630 runs for every --sched other than DDPM / DDIM.
 * state: 6*n floats on the device {cur_model_output, cur_sample, ets[0..3]}; coef: rows of 16 floats
 * {mode, sample_coeff, a_prev - a_t, denom, clip(<=0 off), push slot (-1 none), slots of ets[-1..-4], 0...};
 * x_prev may alias x.  Bit-exact with the reference's fp32 CPU arithmetic. */
int bd_pndm_step(const float* x, const float* eps_hat, float* x_prev, float* state, const float* coef,
                 const int* step_index, size_t n, void* stream);
/* advances *step_index and broadcasts timesteps[*step_index] into t_vec (B) i64 for the next UNet call */
int bd_sampler_advance(int* step_index, const int64_t* timesteps, int64_t* t_vec, int B, int first, void* stream);

/* K14  D/pipelines/ddpm/pipeline_ddpm.py:115-116 + model.py:499: (x/2+0.5).clamp(0,1) -> NHWC f32 and/or
 * round(.*255) u8 (round-half-even like numpy).                                                       */
int bd_finalize_images(const float* x, float* nhwc01, uint8_t* nhwc_u8, int B, int C, int H, int W, void* stream);

/* Measurement tail on the device: replaces save_imgs -> ImagePathDataset -> nn.MSELoss +
 * torchmetrics StructuralSimilarityIndexMeasure(data_range=1.0) (baddiffusion.py:533-546, model.py:496-529).
 *   img (B,H,W,C) u8 = bd_finalize_images' nhwc_u8 (what the PNG would hold), target (C,H,W) f32 in [-1,1];
 *   acc[0] += sum (x - y)^2 over B*C*H*W elements, acc[1] += sum of the SSIM map (11x11 Gaussian window, sigma 1.5,
 *   k1 0.01, k2 0.03) over B*C*(H-10)*(W-10) window centres; x = img/255, y = (target/2+0.5).clamp(0,1).
 *   acc: 2 doubles on the device, zeroed by the caller (several calls / ranks accumulate); H, W > 10. */
int bd_image_metrics(const uint8_t* img_nhwc_u8, const float* target_chw, double* acc, int B, int C, int H, int W,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * K8  D/models/embeddings.py:22-62,155-212: sinusoidal Timesteps + TimestepEmbedding MLP, fp32.
 *   t (B) i64 -> emb (B,temb) f32 and silu(emb) as f16 (B,temb) (the A operand of the per-resnet
 *   time_emb_proj linears, D/models/resnet.py:574-577).  sin_out (B,dim) f32 and h1 (B,temb) f32
 *   (pre-activation of linear_1) are saved for the backward pass.
 * ---------------------------------------------------------------------------------------------- */
int bd_temb_mlp(const int64_t* t, const float* w1, const float* b1, const float* w2, const float* b2, float* sin_out,
                float* h1, float* emb, void* silu_emb_f16, int B, int dim, int temb, int flip_sin_to_cos,
                const float* freqs /* (dim/2) f32: exp(-ln(1e4) * i / (dim/2 - freq_shift)), host-evaluated */,
                void* stream);

/* generic small fp32 GEMM (strided, optionally batched): C[m,n] (+)= sum_k A[m,k]*B[k,n] (+ bias[n]).
 * Used for the tiny linears of the timestep path and their gradients (exact fp32, SIMT).              */
int bd_sgemm(const float* A, int64_t sam, int64_t sak, const float* Bm, int64_t sbk, int64_t sbn, float* C,
             int64_t scm, int64_t scn, const float* bias, int M, int N, int K, int accumulate,
             int act /* 0 none, 1 silu(A), 2 silu(B) */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4/K5  torch.nn.GroupNorm (+SiLU) on fp16 NHWC views; statistics in fp32 (two-stage, fixed order).
 * Call sites replaced: D/models/resnet.py:491,510,553-559,588-591; attention.py:64,126; unet_2d.py:312-313.
 *   x: (B,HW,C) view with ld_x; y likewise.  stats (B,G,2) f32 = {mean, rstd}.
 *   work: bd_gn_workspace_floats(B, C) floats of scratch (per-split partials, group terms, and B arrival counters for
 *   samples that are split over several CTAs: the CTA that finishes a sample last folds its partials).  One GroupNorm
 *   call at a time per workspace; contents need no initialisation and are not preserved.
 * ---------------------------------------------------------------------------------------------- */
size_t bd_gn_workspace_floats(int B, int C);
int bd_groupnorm_fwd(const void* x, int64_t ld_x, void* y, int64_t ld_y, const float* gamma, const float* beta,
                     float* stats, float* work, int B, int HW, int C, int G, float eps, int apply_silu, void* stream);
/* GroupNorm (+SiLU) forward as a pure streaming pass over statistics that the producing conv accumulated (`sums`, see
 * bd_conv_args.gn_sums; channel c of x at sums[b*ld_sums + 2c]): D/models/resnet.py:553-559,588-591 without the
 * reduction pass.  stats (B,G,2) = {mean, rstd} out (nullable), as bd_groupnorm_fwd writes them for the backward. */
int bd_groupnorm_apply_sums(const void* x, int64_t ld_x, void* y, int64_t ld_y, const float* gamma, const float* beta,
                            const float* sums, int64_t ld_sums, float* stats, int B, int HW, int C, int G, float eps,
                            int apply_silu, void* stream);
/* backward: dx = GN'(dy (* SiLU')) [+ add_dx [+ add_dx2]] (gradient fan-in of up to two more branches: the residual
 * path and an earlier consumer's contribution, resnet.py:597-601 / the skip connections); dgamma / dbeta f32 are
 * ACCUMULATED into (zero them first).  work: bd_gn_workspace_floats(B, C) floats.                                 */
int bd_groupnorm_bwd(const void* x, int64_t ld_x, const void* dy, int64_t ld_dy, const void* add_dx, int64_t ld_add,
                     const void* add_dx2, int64_t ld_add2,
                     void* dx, int64_t ld_dx, const float* gamma, const float* beta, const float* stats, float* dgamma,
                     float* dbeta, float* dgb_work,
                     float* gsum /* nullable: (B, C) f32 with row stride ld_gsum, OVERWRITTEN with the per-sample
                                    channel sums of dx = the bias / time_emb_proj gradient of x's producer */,
                     int64_t ld_gsum,
                     float* dgb_parts /* nullable: (B, 2C) f32, OVERWRITTEN with this sample's {dbeta | dgamma} terms
                                         instead of atomically accumulating into dgamma / dbeta (which may then be
                                         null); the caller sums over the batch with bd_bias_from_gsum */,
                     int B, int HW, int C, int G, int apply_silu, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K1/K2/K3/K6/K9/K10  convolution / linear as (implicit) GEMM on fp16 NHWC views, fp32 accumulate.
 * Replaces torch conv2d / F.linear at D/models/resnet.py:118,185,493,514,547-549; attention.py:67-72;
 * unet_2d.py:124,217 and their autograd backward.
 * ---------------------------------------------------------------------------------------------- */
enum {
  BD_CONV_S1 = 0,      /* kxk stride 1, pad k/2                                   */
  BD_CONV_S2_PAD01 = 1 /* 3x3 stride 2 after F.pad(0,1,0,1) (Downsample2D, padding=0) or pad 1 (pad field) */
};
enum { BD_OUT_F16 = 0, BD_OUT_F32 = 1 };
enum {
  BD_IMPL_AUTO = 0,      /* tcgen05 when the shape tiles, CUDA cores otherwise                           */
  BD_IMPL_SIMT = 1,      /* CUDA-core kernels                                                            */
  BD_IMPL_UMMA = 2,      /* tcgen05 (halo-reuse 3x3 kernel where it applies, per-tile kernel otherwise)  */
  BD_IMPL_UMMA_TILE = 3  /* tcgen05 per-tile implicit-GEMM kernel only (A re-fetched per tap)            */
};

typedef struct bd_conv_args {
  /* geometry: input (B,H,W,Cin) -> output (B,Ho,Wo,Cout); ksize 1 or 3 */
  int B, H, W, Cin, Cout, ksize, mode, pad; /* pad: only for BD_CONV_S2_PAD01: 0 -> (0,1,0,1), 1 -> symmetric 1 */
  const void* x;  int64_t ld_x;     /* f16 NHWC view                                         */
  const void* w;                    /* packed f16 [tap][Cout][Cin] (the native layout of the flat
                                       parameter buffer; fwd AND dgrad read it untransposed)  */
  /* optional second K segment fused into the same accumulation (1x1 shortcut, resnet.py:596-597): */
  const void* x2; int64_t ld_x2; int Cin2; const void* w2; /* w2 packed [1][Cout][Cin2]      */
  const float* bias;                /* (Cout) f32, nullable                                  */
  const float* bias2;               /* second bias (the fused shortcut's), nullable          */
  const float* rowbias; int64_t ld_rowbias; /* (B, Cout) f32 added per sample (temb), nullable */
  const void* residual; int64_t ld_res;     /* f16 NHWC view added in the epilogue, nullable */
  float out_scale;                  /* multiplies the result (1/output_scale_factor)          */
  void* y; int64_t ld_y; int out_dtype;     /* BD_OUT_F16 / BD_OUT_F32 NHWC view             */
  int impl;                         /* BD_IMPL_*                                             */
  /* GroupNorm statistics of the OUTPUT accumulated in the epilogue (bd_conv_fwd only, nullable): gn_sums[b*ld_sums +
   * 2*c + {0,1}] += sum / sum of squares over the pixels of sample b of the fp16-rounded outputs of channel c (c counts
   * from this call's first output channel; for a channel slice of a wider tensor pass the slice's address).  The caller
   * zeroes the buffer; bd_groupnorm_apply_sums consumes it.  Only where bd_conv_fwd_gn_sums_supported() says so. */
  float* gn_sums; int64_t ld_sums;
} bd_conv_args;

int bd_conv_fwd(const bd_conv_args* a, void* stream);
/* 1 if bd_conv_fwd would run `a` on a kernel whose epilogue can accumulate gn_sums (the persistent / cluster 3x3
 * tcgen05 kernels: stride 1, H %% 32 == 0, fp16 output), else 0.  Plan-time query, launches nothing. */
int bd_conv_fwd_gn_sums_supported(const bd_conv_args* a);
/* dgrad: dx (B,H,W,Cin) = conv^T(dy (B,Ho,Wo,Cout)) with the SAME packed forward weight (read as an
 * MN-major operand).  `residual` is added (gradient fan-in), x2/w2 unused.  Uses the same struct:
 * x:=dy, y:=dx, Cin/Cout keep their FORWARD meaning.                                                    */
int bd_conv_dgrad(const bd_conv_args* a, void* stream);
/* wgrad: dw f32 in the packed layout [tap][Cout][Cin] (+)= sum over pixels (split-K; the partial tiles meet in dw through
 * vector reductions, red.global.add.v4.f32 -- order noise in the last bits, no scalar atomics);
 * dbias (Cout) f32 nullable.  x (B,H,W,Cin) f16 view, dy (B,Ho,Wo,Cout) f16 view.                      */
int bd_conv_wgrad(const void* x, int64_t ld_x, const void* dy, int64_t ld_dy, float* dw, float* dbias,
                  int B, int H, int W, int Cin, int Cout, int ksize, int mode, int pad, int accumulate, int impl,
                  void* stream);

/* torch OIHW f32 weight -> packed [tap][O][I] as f32 and/or f16 (load_state_dict of foreign checkpoints, tests) */
int bd_pack_conv_weight(const float* w_oihw, float* w_packed_f32, void* w_packed_f16, int O, int I, int ksize, void* stream);
/* f32 -> f16 copy of the flat parameter buffer (one launch per optimizer step) */
int bd_cast_f32_to_f16(const float* src, void* dst, size_t n, void* stream);
/* out[b][c] (+)= sum over the rows of sample b of x (f16 view): per-sample temb gradients and bias gradients */
int bd_colsum_f16(const void* x, int64_t ld_x, float* out, int64_t ld_out, int B, int64_t rows_per_b, int C,
                  int accumulate, void* stream);
/* every bias gradient of the plan in one launch: jobs = DEVICE array of njobs x 4 int64 {gsum pointer (B rows of f32,
 * row stride ld), ld, dst pointer (C f32), C}; dst[c] += sum_b gsum[b][c].  gsum rows come from bd_groupnorm_bwd.   */
int bd_bias_from_gsum(const void* jobs, int njobs, int max_c, int B, void* stream);
/* dx = dy * silu'(x) (f32), n elements: backward of the SiLUs in the timestep MLP */
int bd_silu_bwd_f32(const float* dy, const float* x, float* dx, size_t n, void* stream);
/* y_f16 = silu(x_f32) */
int bd_silu_f32_to_f16(const float* x, void* y, size_t n, void* stream);

/* conv_in (Cin=3, f32 NCHW image in -> f16 NHWC out), unet_2d.py:124,283.  Weights / weight gradients
 * of these two layers are f32 in the packed [tap][Cout][Cin] layout.                                  */
int bd_conv_in_fwd(const float* x_nchw, const float* w_packed, const float* bias, void* y, int64_t ld_y, int B, int Cin,
                   int H, int W, int Cout, void* stream);
/* conv_in whose epilogue also accumulates the GroupNorm statistics of y (layout as bd_conv_args.gn_sums): the first
 * ResnetBlock2D's norm1 and the last up-block concat then need no reduction pass either.  Query first. */
int bd_conv_in_fwd_gn_sums_supported(int Cin, int H, int W, int Cout);
int bd_conv_in_fwd_sums(const float* x_nchw, const float* w_packed, const float* bias, void* y, int64_t ld_y, float* gn_sums,
                        int64_t ld_sums, int B, int Cin, int H, int W, int Cout, void* stream);
int bd_conv_in_wgrad(const float* x_nchw, const void* dy, int64_t ld_dy, float* dw, float* dbias, int B, int Cin,
                     int H, int W, int Cout, int accumulate, void* stream);
/* conv_out (f16 NHWC in -> Cout=3 f32 NCHW eps_hat), unet_2d.py:217,314; and its backward.  bd_conv_out_bwd: dx nullable
 * (parameter gradients only) or dw nullable (data gradient only): the two halves may run on different streams.          */
int bd_conv_out_fwd(const void* x, int64_t ld_x, const float* w_packed, const float* bias, float* y_nchw, int B, int Cin,
                    int H, int W, int Cout, void* stream);
int bd_conv_out_bwd(const void* x, int64_t ld_x, const float* w_packed, const float* dy_nchw, void* dx, int64_t ld_dx,
                    float* dw, float* dbias, int B, int Cin, int H, int W, int Cout, int accumulate, void* stream);

/* K9  F.interpolate(scale_factor=2, nearest) (D/models/resnet.py:146) and its adjoint (2x2 sum) */
int bd_upsample2x(const void* x, int64_t ld_x, void* y, int64_t ld_y, int B, int H, int W, int C, void* stream);
int bd_upsample2x_bwd(const void* dy, int64_t ld_dy, void* dx, int64_t ld_dx, int B, int H, int W, int C, void* stream);
/* y = a (+ b) on f16 views (gradient fan-in where no epilogue is available) */
int bd_add_f16(const void* a, int64_t ld_a, const void* b, int64_t ld_b, void* y, int64_t ld_y, int64_t rows, int C,
               void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7  D/models/attention.py:135-162: softmax_fp32(q k^T * scale) v on the fused-QKV buffer
 * qkv: (B, S, 3C) f16 with ld; heads split C as in reshape_heads_to_batch_dim (:77-82).
 * probs (B*heads, S, S) f16 is saved for backward (nullable in inference).  out: (B,S,C) f16 view.
 * One head, S in {128, 256}, C in {64, 128, 256}: ONE tcgen05 kernel forward (QK^T in TMEM -> softmax in registers -> P as
 * the shared-memory A operand of PV; csrc/umma_attn.cu), two backward (dP -> dS -> dQ; dK + dV); S = 256 runs one CTA per
 * image.  Other single-head tcgen05-tileable shapes: GEMM + softmax + GEMM launches; everything else: CUDA cores.
 * ---------------------------------------------------------------------------------------------- */
size_t bd_attention_fwd_workspace_bytes(int B, int S, int C, int heads);
int bd_attention_fwd(const void* qkv, int64_t ld_qkv, void* probs, void* out, int64_t ld_out, void* work, int B, int S,
                     int C, int heads, float scale, int impl, void* stream);
/* d_qkv (B,S,3C) from d_out; needs probs and qkv from forward. */
int bd_attention_bwd(const void* qkv, int64_t ld_qkv, const void* probs, const void* d_out, int64_t ld_dout,
                     void* d_qkv, int64_t ld_dqkv, void* work, int B, int S, int C, int heads, float scale, int impl,
                     void* stream);
size_t bd_attention_bwd_workspace_bytes(int B, int S, int C, int heads);

/* ------------------------------------------------------------------------------------------------
 * Optimizer tail: baddiffusion.py:608-615 -- unscale, clip_grad_norm_(1.0), Adam, on the
 * flat fp32 parameter / gradient buffers.  state: device floats {loss_scale, growth_tracker, found_inf,
 * grad_norm, step_skipped}.  lr passed by device pointer so the captured graph follows the LR schedule.
 * ---------------------------------------------------------------------------------------------- */
size_t bd_gradnorm_workspace_floats(void);
int bd_grad_norm(const float* grad, size_t n, float* partial, float* state, void* stream);
int bd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                 const float* lr /* device table of lr_len entries, indexed by *step (cosine schedule) */, int lr_len,
                 float beta1, float beta2, float eps, float weight_decay, float max_norm, const int* step,
                 float* state, void* stream);
int bd_scaler_update(float* state, int* step, float growth, float backoff, int growth_interval, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200BD_H */
