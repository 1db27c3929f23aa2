/* rovitkan.h -- C ABI of the B200 (sm_100a) implementation of the RoViT-KAN forward/backward path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every function
 *   - takes DEVICE pointers unless the parameter name ends in `_host`,
 *   - enqueues work on `stream` (a cudaStream_t passed as void*) and returns immediately,
 *   - never allocates device memory (the caller owns outputs and workspaces) and never synchronises,
 *   - returns 0 on success or an RvkStatus code; rvk_strerror()/rvk_last_error() describe it.
 * The Python nn.Module mirror of the reference (rovitkan_b200.models.*) is a thin ctypes client of
 * exactly these symbols; tests/test_abi.py checks that the built library exports every one of them.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference repo).
 */
#ifndef ROVITKAN_H_
#define ROVITKAN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVK_ABI_VERSION 1

/* ---- status ------------------------------------------------------------------------------------- */
int rvk_abi_version(void);
const char* rvk_strerror(int status);
const char* rvk_last_error(void);            /* detail of the last failure on the calling thread */
int rvk_device_check(void);                  /* 0 iff the current device is compute capability 10.x */

/* Every entry point only enqueues work; a fault inside a kernel (every mbarrier wait of this library is bounded: a protocol
 * error ends in a device-side trap after ~4 s, never in a hung GPU) is asynchronous.  rvk_stream_check synchronises the
 * stream and returns RVK_ERR_CUDA (detail in rvk_last_error) if anything enqueued on it faulted.
 * rvk_debug_mbar_timeout launches a kernel that provokes exactly that trap (tests). */
int rvk_stream_check(void* stream);
/* The trunk backward runs its weight-gradient GEMMs on a second, lower-priority stream (joined before it returns);
 * 0 keeps everything on the caller's stream (per-kernel timing, debugging), 1 forces the side stream, -1 = default
 * (on unless RVK_TN_SIDE_STREAM=0). */
void rvk_set_side_stream(int on);
int rvk_debug_mbar_timeout(void* stream);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------------
 * rvk_launch_count: kernels this library has launched in this process so far.
 * rvk_gemm_timing_enable(1): bracket every tensor-core GEMM launch with CUDA events on its own stream;
 * rvk_gemm_timing_collect (after a stream sync) returns the launch count and sums their device time and
 * algorithmic FLOPs (2*M*N*K) since the previous collect; rvk_gemm_timing_kind then gives the same three numbers
 * of that collect for one kernel: kind 0 = gemm_nt_kernel, 1 = gemm_tn_kernel, 2 = mlp_fused_kernel. */
int64_t rvk_launch_count(void);
void rvk_gemm_timing_enable(int on);
int rvk_gemm_timing_collect(double* total_ms_host, double* total_flops_host);
int rvk_gemm_timing_kind(int kind, double* ms_host, double* flops_host);
/* The same for every kernel family of the path: rvk_timing_collect (after a device sync) returns the number of timed launches
 * since the previous collect; rvk_timing_kind gives launches (return value), summed device ms, algorithmic FLOPs and algorithmic
 * HBM bytes of one family; rvk_timing_kind_name its kernel name (NULL past the last kind). */
void rvk_timing_enable(int on);
int rvk_timing_collect(void);
int rvk_timing_kind(int kind, double* ms_host, double* flops_host, double* bytes_host);
const char* rvk_timing_kind_name(int kind);

/* ---- KAN severity path ----------------------------------------------------------------------------
 * Replaces KANLayer.forward (models/kan.py:70-95) incl. BSplineBasis.compute_basis (:10-44), and the
 * inter-layer ReLU / final 3*sigmoid of KANSeverityModule.forward (:138-149) through `act`
 * (0 none, 1 relu, 2 3*sigmoid).  spline [in,out,7], lin_w [out,in], lin_b [out] are the reference
 * parameters unchanged; knots_host are the 11 values of the reference's `knots` buffer.
 * Only the reference configuration (num_knots=5, degree=3 -> 11 knots, 7 basis functions) is built. */
int64_t rvk_kan_layer_workspace_floats(int in_features, int out_features, int with_backward);
/* BSplineBasis.compute_basis (models/kan.py:10-44): t holds n already-normalised inputs, out is [n,7]. */
int rvk_kan_basis(const float* t, const float* knots_host, int num_knots_total, int64_t n, float* out,
                  void* stream);
/* with_backward: bit 0 = the matching backward call will follow (its operands are prepared too); bit 1 = `workspace`
 * already holds the packed / split weights of exactly these parameters from an earlier call with the same bit 0
 * (callers key this on the parameter versions), so the prepare launches are skipped.  Layers with <= 16 outputs run a
 * shuffle-reduction kernel that needs no prepared weights at all. */
int rvk_kan_layer_forward(const float* x, const float* spline, const float* lin_w, const float* lin_b,
                          const float* knots_host, int num_knots_total, int batch, int in_features,
                          int out_features, int act, float* y, float* workspace, int with_backward,
                          void* stream);
/* Gradient of the same layer (what autograd derives from kan.py:70-95).  gy = dL/dy (activated output).
 * dx [batch,in] is overwritten; dspline/dlin_w/dlin_b are accumulated into (+=); any of them may be NULL.
 * `workspace` must be the one the matching forward call (with_backward=1) filled. */
int rvk_kan_layer_backward(const float* x, const float* y, const float* gy, const float* spline,
                           const float* lin_w, const float* knots_host, int num_knots_total, int batch,
                           int in_features, int out_features, int act, float* dx, float* dspline,
                           float* dlin_w, float* dlin_b, float* workspace, void* stream);

/* ---- MLP heads ----------------------------------------------------------------------------------
 * y = epilogue(x W^T + b): one Linear of ClassificationHead / OrdinalHead / UncertaintyHead
 * (models/heads.py:17-22, 38-43, 91-102).  relu!=0 applies ReLU; drop_p>0 applies inverted dropout with
 * a Philox stream keyed by (seed, offset); clamp_lo<clamp_hi clamps (the log-variance head, heads.py:100). */
int rvk_linear_forward(const float* x, const float* w, const float* b, int batch, int in_features,
                       int out_features, int relu, float drop_p, uint64_t seed, uint64_t offset,
                       float clamp_lo, float clamp_hi, float* y, void* stream);
/* y is the forward output (its zeros/saturation encode the ReLU, dropout and clamp masks).
 * dx overwritten (or accumulated when accumulate_dx!=0); dw, db accumulated (+=); gpre_ws [batch,out] scratch. */
int rvk_linear_backward(const float* x, const float* w, const float* y, const float* gy, int batch,
                        int in_features, int out_features, int relu, float drop_p, float clamp_lo,
                        float clamp_hi, float* dx, int accumulate_dx, float* dw, float* db, float* gpre_ws,
                        void* stream);

/* ---- joint loss -----------------------------------------------------------------------------------
 * Replaces JointLoss.forward (training/losses.py:139-181) = FocalLoss (:15-38) + OrdinalBCELoss (:48-72)
 * + UncertaintyLoss (:80-101) + KANRegressionLoss (:109-114).  NULL outputs are the stage-gated heads.
 * out5 = {cls, ord, unc, kan, total}; d_* receive the LOCAL gradients d(term)/d(head output).
 * severity_targets are fp32 (the reference casts them with .float(), losses.py:92,112; ordinal targets are
 * [severity > k]); a class target outside [0, num_classes) makes cls/total NaN and contributes no gradient. */
int rvk_joint_loss_forward(const float* cls_logits, int num_classes, const float* ord_logits, const float* mu,
                           const float* log_var, const float* kan, const int64_t* class_targets,
                           const float* severity_targets, const float* alpha, float gamma, float lambda_ord,
                           float mu_unc, float nu_kan, int batch, float* sums_ws4, float* out5, float* d_cls,
                           float* d_ord, float* d_mu, float* d_lv, float* d_kan, void* stream);
/* dst = local * (upstream5[term] + w_total * upstream5[4]); upstream5 lives on the device (no host sync). */
int rvk_joint_loss_backward(const float* local, const float* upstream5, int term, float w_total, float* dst,
                            int n, void* stream);

/* ---- DeiT-Tiny trunk ------------------------------------------------------------------------------
 * Replaces DeiTTinyBackbone.forward (models/backbone.py:23-25), i.e. the forward of
 * timm.create_model('deit_tiny_patch16_224', num_classes=0) (models/backbone.py:12-16), and its autograd
 * backward.  `params` / `grads` are HOST arrays of 150 device pointers in timm state_dict order:
 *   0 cls_token, 1 pos_embed, 2 patch_embed.proj.weight, 3 patch_embed.proj.bias,
 *   4+12*i.. : blocks.i.{norm1.weight,norm1.bias,attn.qkv.weight,attn.qkv.bias,attn.proj.weight,
 *              attn.proj.bias,norm2.weight,norm2.bias,mlp.fc1.weight,mlp.fc1.bias,mlp.fc2.weight,mlp.fc2.bias},
 *   148 norm.weight, 149 norm.bias                                     (all fp32, reference shapes)
 * prepare_weights casts the GEMM weights to bf16 (plus transposed copies when training) into `wbuf`
 * and folds cls_token/pos_embed/patch bias into the token table; call it whenever parameters change.
 * images: fp32 NCHW [batch,3,224,224]; features: fp32 [batch,192].
 * chunk_images: images per pass through the 12 blocks (0 = whole batch); sized so that a chunk's
 * activations stay L2-resident between kernels. */
#define RVK_ENCODER_NUM_PARAMS 150
int64_t rvk_encoder_weight_bytes(int training);
int64_t rvk_encoder_workspace_bytes(int batch, int training, int chunk_images);
int rvk_encoder_prepare_weights(const void* const* params_host, void* wbuf, int training, void* stream);
int rvk_encoder_forward(const void* const* params_host, const void* wbuf, const float* images, int batch,
                        int training, int chunk_images, void* workspace, float* features, void* stream);
/* Needs the workspace of the matching forward (training=1).  Gradients are accumulated (+=). */
/* Same with bf16 images (B,3,224,224): halves the host->device copy of a serving loop; the result is bit-identical
 * to rvk_encoder_forward on the fp32 images these were rounded from (the trunk rounds pixels to bf16 first). */
int rvk_encoder_forward_bf16(const void* const* params_host, const void* wbuf, const void* images_bf16, int batch,
                             int training, int chunk_images, void* workspace, float* features, void* stream);
/* Same with uint8 NCHW pixels (what an image decoder produces: a quarter of the fp32 host->device bytes).  The
 * reference's transform ToTensor + Normalize(mean, std) is folded into the patch gather: pixel * scale[c] + shift[c]
 * with scale = 1 / (255 * std[c]), shift = -mean[c] / std[c] (host arrays of 3). */
int rvk_encoder_forward_u8(const void* const* params_host, const void* wbuf, const uint8_t* images_u8,
                           const float* scale3_host, const float* shift3_host, int batch, int training,
                           int chunk_images, void* workspace, float* features, void* stream);
int rvk_encoder_backward(const void* const* params_host, const void* wbuf, void* workspace,
                         const float* dfeatures, int batch, int chunk_images, void* const* grads_host,
                         void* stream);

/* Hook / explainability support (reference models/backbone.py:36-62, explainability/attention_maps.py:16-34).
 * rvk_encoder_saved_offset: byte offset, inside the workspace of a training-mode forward over `batch` images (one chunk), of a
 * tensor that forward saved for block `block` -- which: 0 x_in fp32 [M,192] (block input), 1 ln1 bf16 [M,192] (norm1 output),
 * 2 qkv bf16 [M,576], 3 ctx bf16 [M,192] (attention output before proj), 4 x_mid fp32 [M,192] (after the attention
 * residual), 5 ln2 bf16 [M,192], 6 z bf16 [M,768] (fc1 output), 7 h bf16 [M,768] (GELU output); block 12 / which 0 = the
 * stream after the last block.  M = batch * 197.  Returns -1 for an invalid request.
 * rvk_attention_probs: softmax(q k^T / 8) of one block as an explicit fp32 [batch, 3, 197, 197] tensor (timm Attention's
 * `attn` before dropout): the fused attention kernels keep it on chip. */
int64_t rvk_encoder_saved_offset(int batch, int block, int which);
int rvk_attention_probs(const void* qkv_bf16, float* probs, int batch, void* stream);

/* The same backward pass in pieces, so that a data-parallel caller can start the gradient all-reduce of finished blocks
 * while earlier blocks are still being differentiated ("overlapped with the encoder backward in 2-3
 * buckets: heads + KAN first, then blocks 11 -> 0").  Stages: 0 = final LayerNorm, 1 + j = block 11 - j, 13 = patch / class
 * token / position embedding; call with consecutive [stage_begin, stage_end) ranges covering 0..14 in order, whole batch in
 * one chunk.  After stage 1 + j the gradients of blocks 11 - j .. 11 and of norm.* are final, plus mlp.fc2.bias of block
 * 10 - j (it is produced by the LayerNorm backward of the block above it). */
#define RVK_ENCODER_BACKWARD_STAGES 14
int rvk_encoder_backward_range(const void* const* params_host, const void* wbuf, void* workspace,
                               const float* dfeatures, int batch, int chunk_images, void* const* grads_host,
                               int stage_begin, int stage_end, void* stream);

/* RoViTKAN.predict epilogue (models/rovit_kan.py:126-161; OrdinalHead.predict_probabilities / predict_severity,
 * models/heads.py:45-77): class_probs = softmax(cls_logits) [batch,C], class_index = argmax (int64), ordinal_probs [batch,C]
 * from the C-1 cumulative logits, ordinal_severity [batch] = sum_k k * p_k, uncertainty_std [batch] = exp(log_var / 2).
 * ordinal_logits / log_var may be NULL (stage-gated heads). */
int rvk_predict_decode(const float* cls_logits, int num_classes, const float* ordinal_logits, const float* log_var,
                       int batch, int64_t* class_index, float* class_probs, float* ordinal_probs,
                       float* ordinal_severity, float* uncertainty_std, void* stream);

/* The same tail for the TRAINING step (north_star (c): "heads and their losses become one fused epilogue", forward and
 * backward): one forward launch and -- after rvk_joint_loss_forward / _backward produced d(loss)/d(head outputs) -- one
 * backward launch (+ a memset and an unpack launch) replace ~50 per-layer launches.  Forward: Dropout(drop_p) on the three
 * hidden layers with a Philox keep mask (seed, offset); h_save [batch,384], a1_save [batch,64], a2_save [batch,16] receive
 * the activations the backward needs.  Backward: d_* = gradients of the five outputs (NULL = that head is stage-gated
 * off: its parameters get no gradient), dfeatures [batch,192] is overwritten, dws = scratch of
 * rvk_heads_fused_workspace_floats() floats, grads23_host = host array of 23 device pointers in the parameter order of
 * rvk_heads_fused_prepare (NULL entries are skipped); gradients are OVERWRITTEN, not accumulated. */
int rvk_heads_train_forward(const float* features, const float* ws, const float* knots_host, int batch, float drop_p,
                            uint64_t seed, uint64_t offset, float* cls_logits, float* ordinal_logits, float* mu,
                            float* log_var, float* kan_severity, float* h_save, float* a1_save, float* a2_save,
                            void* stream);
int rvk_heads_train_backward(const float* features, const float* ws, const float* knots_host, int batch, float drop_p,
                             const float* h_save, const float* a1_save, const float* a2_save, const float* log_var,
                             const float* kan_severity, const float* d_cls, const float* d_ord, const float* d_mu,
                             const float* d_log_var, const float* d_kan, float* dfeatures, float* dws,
                             float* const* grads23_host, void* stream);

/* ---- fused optimizer tail ---------------------------------------------------------------
 * Replaces what the reference's trainer runs after loss.backward() (training/trainer.py:118-129): GradScaler.unscale_ and
 * its inf check, clip_grad_norm_(parameters, max_norm), AdamW.step with the two learning-rate groups of
 * training/optimizer.py:18-25 -- three launches over all tensors.  `params_host` / `grads_host`: host arrays of n DEVICE
 * pointers (fp32; a NULL gradient skips that tensor as torch does); `group_host[i]` selects lr_host[group]; exp_avg /
 * exp_avg_sq: rvk_optimizer_state_floats(n, numel) floats each, zero-initialised by the caller; state4 (device, zeroed
 * once): {scratch, step count, last gradient norm (after grad_mult and unscaling), 1 if the last step ran}.
 * Gradients are used as g * grad_mult / *grad_scale_dev (grad_scale_dev may be NULL); when the resulting global norm is
 * non-finite or *found_inf_dev != 0 the whole update is skipped and the step count is not advanced (GradScaler.step).
 * max_grad_norm <= 0 disables clipping.  AdamW arithmetic follows torch's fused kernel (decoupled weight decay,
 * lerp moment update, bias corrections in double). */
int64_t rvk_optimizer_state_floats(int n_tensors, const int64_t* numel_host);
int rvk_optimizer_step(int n_tensors, void* const* params_host, const void* const* grads_host, const int64_t* numel_host,
                       const int* group_host, float* exp_avg, float* exp_avg_sq, float* state4, const double* lr_host,
                       int n_groups, double beta1, double beta2, double eps, double weight_decay, float max_grad_norm,
                       float grad_mult, const float* grad_scale_dev, const float* found_inf_dev, void* stream);

/* ---- fused inference tail: all four heads of RoViTKAN.forward (models/rovit_kan.py:96-124 in eval mode) in ONE kernel:
 * the three Linear-ReLU-Linear heads (models/heads.py:17-22, 38-43, 91-102, log_var clamped to +-10) and the KAN severity
 * stack KAN(192,64)-ReLU-KAN(64,16)-ReLU-KAN(16,1)-3*sigmoid (models/kan.py:138-149).  Fixed reference architecture
 * (embed 192, hidden 128, 4 classes, kan_layers [192,64,16,1]).  `params23_host`: host array of 23 DEVICE pointers in the
 * order cls.fc1.{weight,bias}, cls.fc2.{weight,bias}, ord.fc1.*, ord.fc2.*, unc.fc1.*, unc.fc_mu.*, unc.fc_logvar.*, then
 * {spline_weights, linear.weight, linear.bias} of kan_layers 0..2.  rvk_heads_fused_prepare repacks them into `ws`
 * (rvk_heads_fused_workspace_floats() floats; call again whenever a parameter changes). */
int64_t rvk_heads_fused_workspace_floats(void);
int rvk_heads_fused_prepare(const void* const* params23_host, float* ws, void* stream);
int rvk_heads_fused(const float* features, const float* ws, const float* knots_host, int batch, float* cls_logits,
                    float* ordinal_logits, float* mu, float* log_var, float* kan_severity, void* stream);

/* ---- individual encoder kernels (exposed for parity tests and microbenchmarks) ---------------------- */
/* C = epilogue(A[M,K] B[N,K]^T), bf16 operands, tcgen05/TMEM.  mode: 0 bf16(acc+bias), 1 gelu (+z in out2),
 * 2 acc*gelu'(aux z), 3 fp32(acc+bias), 4 fp32 acc+bias+residual(aux or table) with optional fused LayerNorm
 * (out2 bf16, gamma/beta, mean/rstd).  N must be a multiple of 192 (<= 768), K a multiple of 64. */
int rvk_gemm_nt(int mode, const void* a_bf16, int64_t lda, const void* b_bf16, int64_t ldb, void* out, int64_t ldo,
                void* out2, int64_t ldo2, const void* aux, int64_t ldaux, int m, int n, int k, const float* bias,
                const float* gamma, const float* beta, const float* res_table, int table_rows, float ln_eps,
                float* mean_out, float* rstd_out, void* stream);
/* Fused MLP half of a DeiT block (timm Block.forward: x + mlp(norm2(x)), plus the NEXT block's norm1), inference path:
 *   x_out = x_in + fc2(gelu(fc1(LayerNorm(x_in)*gamma2 + beta2) + b1)) + b2;   ln_out = LayerNorm(x_out)*gamma + beta
 * (ln_out may be NULL).  Neither the normalised input nor the hidden activation is written to memory.
 * x_in / x_out fp32 in the TILED token-stream layout (element (r,c) at
 * (((r/32)*6 + c/32)*8 + (c%32)/4)*128 + (r%32)*4 + c%4, rows padded to 128; x_out may alias x_in); w1 bf16 [768,192];
 * w2 FP16 [192,768] (the hidden activation stays on chip in fp16: GELU runs as packed half2 arithmetic); ln_out bf16
 * [m,192].  cta_group: 2 = CTA pairs sharing one tcgen05.mma.cta_group::2, 1 = single CTAs (4: see below, needs the projection). */
int rvk_mlp_fused(const float* x_in_tiled, float* x_out_tiled, const float* gamma2, const float* beta2,
                  const void* w1_bf16, const float* b1, const void* w2_f16, const float* b2, const float* gamma,
                  const float* beta, float eps, void* ln_out_bf16, int m, int cta_group, void* stream);
/* The same kernel with the attention output projection folded in (timm Block.forward first half: x + proj(attn),
 * Attention.proj): before the MLP half it computes  x_in += ctx . wproj^T + bproj  on the tensor cores (accumulator in
 * the TMEM columns that are idle between two row tiles), so the projected residual stream never makes its own round
 * trip through memory.  ctx bf16 [m,192] (rvk_attention_forward output), wproj bf16 [192,192], bproj fp32 [192].
 * cta_group 4 = CTA pairs with TWO row tiles in flight (two groups of epilogue warps, the projected row parked in TMEM under
 * fc2; csrc/mlp_fused2.cuh) -- what the inference trunk runs by default; RVK_ERR_BAD_ARG in rvk_mlp_fused (no projection). */
int rvk_attn_proj_mlp_fused(const float* x_in_tiled, float* x_out_tiled, const void* ctx_bf16, const void* wproj_bf16,
                            const float* bproj, const float* gamma2, const float* beta2, const void* w1_bf16, const float* b1,
                            const void* w2_f16, const float* b2, const float* gamma, const float* beta, float eps,
                            void* ln_out_bf16, int m, int cta_group, void* stream);
/* Debugging aid: device buffer of 4*512 int64 into which the following rvk_mlp_fused launches log clock64 events
 * of CTA 0 (NULL switches it off, the default). */
void rvk_debug_set_mlp_trace(void* device_buf);
void rvk_debug_set_attn_trace(void* device_buf);   /* same for rvk_attention_forward */
/* C[P,Q] (fp32) += scale * A[M,P]^T B[M,Q], bf16 operands (weight gradients).  a_colsum (optional, fp32 [P], needs
 * Q <= 192): += scale * sum_m A[m,p] -- the bias gradient of the same Linear layer, computed by the same kernel through a
 * constant column of ones appended to B. */
int rvk_gemm_tn(const void* a_bf16, int64_t lda, const void* b_bf16, int64_t ldb, float* c, int64_t ldc, int m,
                int p, int q, float scale, float* a_colsum, void* stream);
/* softmax(q k^T / 8) v per (image, head); qkv bf16 [batch*197,576], ctx bf16 [batch*197,192], lse fp32
 * [batch,3,197] (log2 domain) or NULL. */
int rvk_attention_forward(const void* qkv, void* ctx, float* lse, int batch, void* stream);
int rvk_attention_backward(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                           int batch, void* stream);
int rvk_layernorm_forward(const float* x, int64_t x_row_stride, const float* gamma, const float* beta, float eps,
                          void* y, int y_is_bf16, int64_t y_row_stride, float* mean, float* rstd, int rows,
                          void* stream);
int rvk_layernorm_backward(const void* g, int g_is_bf16, int64_t g_row_stride, const float* x, int64_t x_row_stride,
                           const float* mean, const float* rstd, const float* gamma, const float* dx_in,
                           float* dx_out, int64_t dx_row_stride, void* dx_out_bf16, float* dgamma, float* dbeta,
                           float* dcolsum, int rows, void* stream);
int rvk_im2col(const float* images, void* patches_bf16, int batch, void* stream);
int rvk_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ROVITKAN_H_ */
