/* dm_abi.h -- C ABI of the B200-native typicality / DIFT engine (libdm_b200.so).
 *
 * The reference (ysig/diff-mining) has no FFI of its own: its hot path is plain Python method calls into
 * diffusers/xformers/torch.  Each entry point below states the reference interface it replaces
 * (paths relative to /root/reference/).  Plain pointers and sizes only; every device pointer is owned by the
 * caller (torch on the host side); all work is enqueued on the passed cudaStream_t (as void*); the engine owns
 * only its packed weights, context K/V caches and workspace arena.
 *
 * Return value: 0 = OK, negative = error; dm_last_error() returns the message of the calling thread's last
 * failure.  There is no CPU fallback: unsupported shapes / missing weights fail loudly.
 */
#ifndef DM_ABI_H_
#define DM_ABI_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dm_engine dm_engine;

enum { DM_OK = 0, DM_ERR = -1 };
enum { DM_F32 = 0, DM_F16 = 1, DM_BF16 = 2 }; /* host dtype tags for dm_load_tensor */

const char* dm_last_error(void);
int dm_abi_version(void);

/* ---- lifetime ------------------------------------------------------------------------------------------
 * replaces: SD.__init__ building StableDiffusionPipeline.from_pretrained(...).to(device)
 *           (diffmining/typicality/compute.py:57-79) and SDFeaturizer.__init__ (diffmining/typicality/dift.py:195-211). */
int dm_create(int device, dm_engine** out);
int dm_destroy(dm_engine* e);

/* One call per state-dict entry, diffusers key schema: "unet.<key>" (686 tensors), "vae.encoder.<key>",
 * "vae.quant_conv.{weight,bias}".  host_ptr is HOST memory in `dtype`; values are rounded to fp16 exactly
 * as the reference's torch_dtype=float16 load does (compute.py:65-70).  The engine repacks to its own layouts
 * (OIHW -> [O][kh][kw][I] K-major, fused q|k|v, interleaved GEGLU rows). */
int dm_load_tensor(dm_engine* e, const char* key, const void* host_ptr, int dtype, int ndim, const int64_t* shape);
int dm_finalize_weights(dm_engine* e); /* checks that every SD-1.5 tensor arrived; packs; frees staging */

/* Packed-weight cache: dm_save_packed writes the packed device buffers of a finalized engine to one
 * file; dm_load_packed on a FRESH engine replaces the whole dm_load_tensor... + dm_finalize_weights sequence (the host
 * keys the file by a hash of the checkpoint it came from: diff-mining_b200/typicality.py packed_cache_path). */
int dm_save_packed(dm_engine* e, const char* path);
int dm_load_packed(dm_engine* e, const char* path);

/* alphas_cumprod-derived tables of scheduler.add_noise (compute.py:99; dift.py:190): 1000 fp32 each, HOST. */
int dm_set_schedule(dm_engine* e, const float* sqrt_acp, const float* sqrt_one_minus_acp, int n);

/* Text context slot <- encoder_hidden_states [77,768] fp32 HOST (CategoryFeatures.embed output,
 * compute.py:39-51).  Precomputes the 16 cross-attention K/V projections for the slot. */
int dm_set_context(dm_engine* e, int slot, const float* ctx_77x768, void* stream);

/* ---- hot path ------------------------------------------------------------------------------------------ */

/* replaces SD.encode_vae (compute.py:91-93): vae.encode(x).latent_dist.sample() * 0.18215 under autocast.
 * img: DEVICE fp32 [B,3,H,W] in [-1,1]; eps: DEVICE fp32 [B,4,H/8,W/8] posterior draw (or NULL -> mean);
 * outputs (any may be NULL): z, mean, logvar DEVICE fp32 [B,4,H/8,W/8]. */
int dm_vae_encode(dm_engine* e, const float* img, const float* eps, int B, int H, int W, float* z, float* mean,
                  float* logvar, void* stream);

/* replaces self.model.unet(noisy, t, ctx).sample (compute.py:100): x_noisy DEVICE fp32 [Bf,4,h,w],
 * t DEVICE int64 [Bf], ctx_slots HOST int32 [Bf]; eps_out DEVICE fp32 [Bf,4,h,w] (fp16-rounded values). */
int dm_unet_eps(dm_engine* e, const float* x_noisy, const int64_t* t, const int32_t* ctx_slots, int Bf, int h, int w,
                float* eps_out, void* stream);

/* General row form of SD.compute_loss (compute.py:95-102): row i uses latent x[x_index[i]], draw
 * noise[noise_index[i]] with timestep t[noise_index[i]], and context slot ctx_slots[i] (index arrays are HOST
 * int32 [M]; NULL index = identity).  noise == NULL means x rows are already noisy and t is indexed like x.
 * loss_out / eps_out (either may be NULL): DEVICE fp32 [M,4,h,w]. */
int dm_unet_rows(dm_engine* e, const float* x, const float* noise, const int64_t* t, const int32_t* x_index,
                 const int32_t* noise_index, const int32_t* ctx_slots, int M, int h, int w, float* loss_out,
                 float* eps_out, int max_forwards, void* stream);

/* replaces SD.compute_loss (compute.py:95-102) for one micro-batch: rows are condition-major
 * [cond0 x S ; cond1 x S ; ...] exactly as D.compute_losses builds them (compute.py:150-152).
 * x0 DEVICE fp32 [1,4,h,w]; noise DEVICE fp32 [S,4,h,w]; t DEVICE int64 [S]; ctx_slots HOST int32 [n_cond];
 * loss_out DEVICE fp32 [n_cond*S,4,h,w]. */
int dm_compute_loss(dm_engine* e, const float* x0, const float* noise, const int64_t* t, const int32_t* ctx_slots,
                    int S, int n_cond, int h, int w, float* loss_out, void* stream);

/* replaces D.compute_losses' Monte-Carlo loop (compute.py:134-160) for Bi images at once plus the consumers'
 * reduction into T(x|c) (diffmining/typicality/cluster.py:112-123; utils.py:122-134):
 * x0 DEVICE fp32 [Bi,4,h,w]; noise DEVICE fp32 [N,4,h,w] and t DEVICE int64 [N] shared by all images
 * (the reference re-seeds per image, compute.py:139); ctx_slots HOST int32 [n_cond], LAST = unconditional.
 * grid_out (or NULL): DEVICE fp16 [Bi,N,n_cond,4,h,w] -- the .npy payload (compute.py:155-160,192).
 * T_out (or NULL): DEVICE fp32 [Bi,n_cond-1,h,w] = mean_s[mean_ch L(uncond) - mean_ch L(cond_k)].
 * max_forwards bounds the U-Net micro-batch (0 = engine default). */
int dm_typicality(dm_engine* e, const float* x0, const float* noise, const int64_t* t, const int32_t* ctx_slots, int Bi,
                  int N, int n_cond, int h, int w, void* grid_out, float* T_out, int max_forwards, void* stream);

/* replaces MyUNet2DConditionModel.forward with up_ft_indices=[up_ft_index] + the ensemble mean of
 * SDFeaturizer.forward (dift.py:24-169, 229-231): latents DEVICE fp32 [B*E,4,h,w] (clean), noise DEVICE fp32
 * same shape, scalar timestep t; feat_out DEVICE fp32 [B, C_up, h_up, w_up] = mean over the E members. */
int dm_dift(dm_engine* e, const float* latents, const float* noise, int64_t t, int ctx_slot, int B, int E, int h, int w,
            int up_ft_index, float* feat_out, void* stream);
int dm_dift_shape(int h, int w, int up_ft_index, int* C, int* ho, int* wo);

/* ---- introspection -------------------------------------------------------------------------------------- */
/* kernels launched by this engine since creation (bench.py's gpu_launches) */
int64_t dm_launch_count(dm_engine* e);
/* algorithmic FLOPs of the tensor-core ops enqueued since creation */
double dm_flop_count(dm_engine* e);
/* copy a named intermediate activation of the LAST U-Net forward to `out` as fp32 NCHW; returns element count or <0.
 * Only recorded when dm_debug_keep(e, 1) was set before the forward (disables buffer reuse). */
int dm_debug_keep(dm_engine* e, int on);
int64_t dm_debug_fetch(dm_engine* e, const char* name, float* out_dev, int64_t capacity, int* dims4, void* stream);
/* per-kernel timing of one U-Net micro-batch (ms by op class) for roofline reporting; arrays of n entries */
int dm_profile_unet(dm_engine* e, int Bf, int h, int w, int iters, double* ms_igemm, double* ms_attn, double* ms_other,
                    double* flops_igemm, double* flops_attn);

/* same for any cached plan: kind 0 = U-Net eps (aux = rows per shared-prefix group, 0/1 = unshared), 1 = DIFT partial
 * forward (aux = up_ft_index), 2 = VAE encoder (h, w = image size).  ms_by_class[3] = {implicit GEMM, attention, other},
 * flops_by_class[2] = {implicit GEMM, attention} algorithmic FLOPs of one replay. */
int dm_profile_plan(dm_engine* e, int kind, int Bf, int h, int w, int aux, int iters, double* ms_by_class,
                    double* flops_by_class);

/* ---- T-map consumer: replaces Cluster.load_typicality + df_D + get_non_overlapping
 * (cluster.py:125-137,183-205; utils.py:74-102) for the engine's own T maps.  T: DEVICE fp32 [B,h,w] (dm_typicality
 * output, one condition); the map is resized bilinearly (align_corners=False) to H x W, average-pooled with a kx x ky
 * window (stride 1, valid) and the k best mutually non-overlapping windows are kept (descending != 0: highest scores
 * first).  work: DEVICE fp32 scratch of dm_patch_topk_work_floats() elements.  Outputs (DEVICE): boxes int32 [B,k,4] =
 * (x_start, y_start, x_end, y_end) in the reference's (row, column) convention with x_end = x_start + kx, scores fp32
 * [B,k], count int32 [B] (windows actually found). */
int dm_patch_topk(const float* T, int B, int h, int w, int H, int W, int kx, int ky, int k, int descending, float* work,
                  int32_t* boxes, float* scores, int32_t* count, void* stream);
int64_t dm_patch_topk_work_floats(int B, int H, int W, int ky);

/* ---- operator-level entry points (unit tests / debugging; same kernels the engine uses) ------------------ */
/* conv / linear as implicit GEMM.  x,x2: DEVICE fp16 NHWC [N,H,W,C0|C1] (x2 may be NULL); w: DEVICE fp16
 * [Cout, ks*ks*(C0+C1)] tap-major; bias DEVICE fp32 [Cout] or NULL; rowbias DEVICE fp16 [N,Cout] or NULL;
 * residual DEVICE fp16 [N*Ho*Wo, Cout] or NULL; out DEVICE fp16 (fp32 if out_f32) [N*Ho*Wo, Cout (/2 if geglu)].
 * stride 2 uses pad 1 (U-Net) or the VAE's (0,1,0,1) pad when vae_pad != 0. */
int dm_op_conv(const void* x, const void* x2, int N, int H, int W, int C0, int C1, const void* w, int Cout, int ks,
               int stride, int vae_pad, const float* bias, const void* rowbias, const void* residual, void* out,
               int out_f32, int geglu, int act_silu, int bn, void* stream);
/* 3x3 / 1x1 stride-1 conv over ONE source whose epilogue also forms the GroupNorm(32 groups) statistics of its output
 * (igemm.cuh: IgGn), followed by the fold + apply kernel (norm.cuh: gn_fold_apply_kernel) -- the pair the engine uses for
 * conv1 -> norm2 and conv2 -> Transformer2DModel.norm.  out = conv output fp16 [N*H*W, Cout]; gn_out =
 * GroupNorm(+SiLU)(out) fp16.  Errors when H*W is not made of whole 128-pixel tiles or Cout not of whole N-tiles. */
int dm_op_conv_gn(const void* x, int N, int H, int W, int C0, const void* w, int Cout, int ks, const float* bias,
                  const void* rowbias, const void* residual, const float* gamma, const float* beta, float eps, int silu,
                  void* out, void* gn_out, void* stream);
int dm_op_attention(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_k, int64_t ld_v, int64_t bs_q,
                    int64_t bs_k, int64_t bs_v, int B, int heads, int D, int Tq, int Tk, int kv_batches,
                    const int32_t* kv_index_dev, void* out, int64_t ld_out, void* stream);
int dm_op_groupnorm(const void* x, const void* x2, int N, int HW, int C0, int C1, const float* gamma, const float* beta,
                    float eps, int silu, void* out, void* stream);
/* LayerNorm over the last dim.  At C = 320 / 640 / 1280 gamma and beta are rounded to fp16 inside the kernel (they are fp16
 * weights in the engine, where the fp32 copies convert back exactly). */
int dm_op_layernorm(const void* x, int64_t rows, int C, const float* gamma, const float* beta, float eps, void* out,
                    void* stream);
/* kernel-variant switches for tests / A-B timing (affect ops prepared afterwards; -1 = built-in default):
 *   igemm_pair  1 = CTA pairs (tcgen05 cta_group::2, 256-row tiles) where the problem is large enough, 0 = never,
 *               2 = wherever the tile shape allows (N-tile >= 128, fp16 output)
 *   igemm_ng4   1 = four epilogue warpgroups for short-K (epilogue-bound) layers, 0 = always two, 2 = wherever possible
 *   gn_fused    1 = cluster-fused single-pass GroupNorm for images that fit in L2, 0 = two-kernel path
 *   xattn       2 = persistent short-key-set cross-attention kernel at head_dim 40 (default), 3 = also at head_dim 80,
 *               1 = the same arithmetic with one CTA per query block (bit-identical), 0 = generic flash kernel
 *   igemm_ws    1 = weight-stationary CTA pairs for K <= 320 Linears with many M-tiles (measured slower: default 0),
 *               2 = wherever the shape allows (tests)
 *   attn3       persistent self-attention kernel (attention3.cuh): 1 = two softmax threads per query row, 2 = one thread per
 *               row; 0 = one CTA per 256-query block (attention2.cuh)
 *   gn_epilogue bit mask: 1 = 3x3 convs, 2 = 1x1 convs / Linears form the GroupNorm statistics of their output in the
 *               epilogue (igemm.cuh: IgGn) and the GroupNorm becomes one fold + apply pass; 0 = stand-alone GroupNorm
 *               kernels everywhere.  Default 3.  Takes effect for plans built afterwards.
 *   prefix_share 1 = dm_typicality computes the context-free U-Net prefix once per (eps,t) draw (bit-identical) */
int dm_op_set_variant(const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* DM_ABI_H_ */
