// Internal (C++) launch interface of the sm_100a kernels.  Everything takes raw device pointers,
// sizes and a stream, never allocates or synchronises, and returns an RvkStatus.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_nt.cuh"
#include "gemm_tn.cuh"
#include "mlp_fused.cuh"

// ---- tcgen05 GEMMs -----------------------------------------------------------------------------
struct GemmNtArgs {
  int mode = EPI_BF16;            // GemmEpilogue
  const void* A = nullptr;        // bf16 [M, K], row stride lda
  int64_t lda = 0;
  const void* B = nullptr;        // bf16 [N, K], row stride ldb
  int64_t ldb = 0;
  void* out = nullptr;            // bf16 or fp32 [M, N], row stride ldo
  int64_t ldo = 0;
  void* out2 = nullptr;           // EPI_GELU: pre-activation z (bf16); EPI_RES_LN: LayerNorm output (bf16)
  int64_t ldo2 = 0;
  const void* aux = nullptr;      // EPI_DGELU: z (bf16 [M,N]); EPI_RES_LN: residual (fp32 [M,192])
  int64_t ldaux = 0;
  GemmNtParams p{};
};
int rvk_gemm_nt_launch(const GemmNtArgs& a, cudaStream_t stream);

// C[P,Q] (fp32, ldc) += scale * A[M,P]^T B[M,Q]; A, B bf16 row-major; optional a_colsum[P] += scale * column sums of A
// (the bias gradient when A is an output gradient; needs Q <= 192)
int rvk_gemm_tn_launch(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int M, int P,
                       int Q, float scale, float* a_colsum, cudaStream_t stream);

// fused MLP block of the inference path: x_out = x_in + fc2(gelu(fc1(LayerNorm2(x_in)))) (+ LayerNorm -> ln_out); the
// fp32 token stream x uses the tiled layout of common.cuh (xt_offset).  cta_group: 2 = CTA pairs (default), 1 = single CTAs.
struct MlpFusedArgs {
  const void* w1 = nullptr;       // bf16 [768,192]
  const void* w2_f16 = nullptr;   // fp16 [192,768] (the hidden activation is kept in fp16)
  void* ln_out = nullptr;         // bf16 [M,192] (p.has_ln)
  const void* ctx = nullptr;      // bf16 [M,192] attention output  (p.has_proj: x_in += ctx . wproj^T + p.bp first)
  const void* wproj = nullptr;    // bf16 [192,192]
  int cta_group = 2;
  MlpFusedParams p{};
};
int rvk_mlp_fused_launch(const MlpFusedArgs& a, cudaStream_t stream);

// programmatic dependent launch of the big trunk kernels (gemm_nt, attention forward, fused MLP); RVK_PDL=0 switches it off
bool rvk_pdl_enabled();

// ---- attention -----------------------------------------------------------------------------------
int rvk_attention_fwd_launch(const void* qkv, void* ctx, float* lse, int batch, cudaStream_t stream);
int rvk_attention_probs_launch(const void* qkv, float* probs, int batch, cudaStream_t stream);
int rvk_attention_bwd_launch(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                             int batch, cudaStream_t stream);
int rvk_attention_bwd_tc_launch(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                                int batch, cudaStream_t stream);

// ---- token-stream kernels (encoder_kernels.cu) -----------------------------------------------------
// fused multi-task tail of the training step (kan.cu: heads_fused.cuh<true>, heads_train.cuh)
int rvk_heads_train_fwd_launch(const float* features, const float* ws, const float* knots_host, int batch, float drop_p,
                               unsigned long long seed, unsigned long long offset, float* cls, float* ord, float* mu,
                               float* log_var, float* kan, float* h_save, float* a1_save, float* a2_save, cudaStream_t stream);
int rvk_heads_train_bwd_launch(const float* features, const float* ws, const float* knots_host, int batch, float drop_p,
                               const float* h_save, const float* a1_save, const float* a2_save, const float* lv_out,
                               const float* kan_out, const float* d_cls, const float* d_ord, const float* d_mu, const float* d_lv,
                               const float* d_kan, float* dfeat, float* dws, float* const* grads23_host, cudaStream_t stream);
// fused optimizer tail (optimizer.cu)
int64_t rvk_optimizer_state_floats_impl(int n, const int64_t* numel_host);
int rvk_optimizer_step_impl(int n, void* const* params_host, const void* const* grads_host, const int64_t* numel_host,
                            const int* group_host, float* exp_avg, float* exp_avg_sq, float* state4, const double* lr_host,
                            int n_groups, double beta1, double beta2, double eps, double weight_decay, float max_grad_norm,
                            float grad_mult, const float* grad_scale_dev, const float* found_inf_dev, cudaStream_t stream);
// several fp32 -> bf16 / bf16-transposed / fp16 matrix casts in one launch (tile_start is filled in by the launcher)
struct RvkCastJob { const float* src; void* dst; int rows, cols, mode, tile_start; };      // mode 0 bf16, 1 bf16 transposed, 2 fp16
constexpr int kRvkMaxCastJobs = 64;
struct RvkCastTable { RvkCastJob job[kRvkMaxCastJobs]; int n; };
int rvk_cast_multi_launch(RvkCastTable& T, cudaStream_t stream);
int rvk_debug_mbar_timeout_launch(cudaStream_t stream);
// fmt: 0 fp32, 1 bf16, 2 uint8 (+ norm6_host = {scale[3], shift[3]}: pixel * scale[c] + shift[c])
int rvk_im2col_launch(const void* images, int fmt, void* patches_bf16, int batch, const float* norm6_host, cudaStream_t stream);
int rvk_token_table_launch(const float* cls_token, const float* pos_embed, const float* patch_bias, float* table,
                           cudaStream_t stream);
int rvk_cast_bf16_launch(const float* src, void* dst, int64_t n, cudaStream_t stream);
int rvk_cast_f16_launch(const float* src, void* dst, int64_t n, cudaStream_t stream);
int rvk_cast_transpose_bf16_launch(const float* src, void* dst, int rows, int cols, cudaStream_t stream);
int rvk_layernorm_fwd_launch(const float* x, int64_t x_row_stride, const float* gamma, const float* beta, float eps,
                             void* y, int y_is_bf16, int64_t y_row_stride, float* mean, float* rstd, int rows,
                             cudaStream_t stream);
// LayerNorm of rows r * token_row_stride (r < rows) of the TILED fp32 token stream (common.cuh xt_offset) -> fp32 y
int rvk_layernorm_fwd_tiled_launch(const float* x_tiled, int64_t token_row_stride, const float* gamma, const float* beta,
                                   float eps, float* y, int64_t y_row_stride, int rows, cudaStream_t stream);
// dx_out = dx_in + LN'(g) (dx_in may be null; dx_out may alias dx_in); g is bf16 or fp32 with row stride g_stride
int rvk_layernorm_bwd_launch(const void* g, int g_is_bf16, int64_t g_row_stride, const float* x, int64_t x_row_stride,
                             const float* mean, const float* rstd, const float* gamma, const float* dx_in,
                             float* dx_out, int64_t dx_row_stride, void* dx_out_bf16, float* dgamma, float* dbeta,
                             float* dcolsum, int rows, cudaStream_t stream);
int rvk_colsum_launch(const void* src, int src_is_bf16, int64_t ld, int rows, int cols, float* out, float scale,
                      cudaStream_t stream);
int rvk_token_grad_reduce_launch(const float* dx0, int batch, float* dpos, float* dcls, float* dpatch_bias,
                                 cudaStream_t stream);

// ---- KAN ---------------------------------------------------------------------------------------------
struct KanLayerDesc {
  const float* spline = nullptr;   // [in, out, 7]   (reference KANLayer.spline_weights)
  const float* lin_w = nullptr;    // [out, in]
  const float* lin_b = nullptr;    // [out]
  float knots_host[11] = {};       // values of the reference's `knots` buffer (host copy)
  int num_knots = 11, num_basis = 7;
  int in_features = 0, out_features = 0;
};
int64_t rvk_kan_workspace_floats(int n_in, int n_out, int with_backward);
int rvk_kan_basis_launch(const float* t, const float* knots_host, float* out, int64_t n, cudaStream_t stream);
// act: 0 none, 1 relu, 2 3*sigmoid.  `workspace` holds the packed weights (see rvk_kan_workspace_floats).
int rvk_kan_layer_fwd_launch(const KanLayerDesc& L, const float* x, float* y, int act, int batch, float* workspace,
                             int with_backward, cudaStream_t stream);
// gy: gradient w.r.t. the activated output y.  dspline/dlin_w/dlin_b are accumulated into (+=); dx is overwritten.
// Needs the workspace of the matching forward launch (with_backward = 1).
int rvk_kan_layer_bwd_launch(const KanLayerDesc& L, const float* x, const float* y, const float* gy, int act,
                             float* dx, float* dspline, float* dlin_w, float* dlin_b, int batch, float* workspace,
                             cudaStream_t stream);

// fused inference tail: all four heads in one kernel (heads_fused.cuh).  p23: 23 device pointers in the order
// cls.fc1.{w,b}, cls.fc2.{w,b}, ord.fc1.{w,b}, ord.fc2.{w,b}, unc.fc1.{w,b}, unc.fc_mu.{w,b}, unc.fc_logvar.{w,b}, then
// (spline, lin_w, lin_b) of the three KAN layers; fixed architecture 192 -> 128 -> {4,3,1,1} and KAN [192,64,16,1].
int64_t rvk_heads_fused_workspace_floats_impl();
int rvk_heads_fused_prepare_launch(const void* const* p23, float* ws, cudaStream_t stream);
int rvk_heads_fused_launch(const float* features, const float* ws, const float* knots_host, int batch, float* cls,
                           float* ord, float* mu, float* log_var, float* kan, cudaStream_t stream);

// ---- heads (fp32 SIMT GEMM with fused epilogues) and joint loss -------------------------------------------
struct SgemmArgs {
  const float* A = nullptr; int64_t lda = 0; int transA = 0;
  const float* B = nullptr; int64_t ldb = 0; int transB = 0;
  float* C = nullptr; int64_t ldc = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;
  int relu = 0;
  float clamp_lo = 0.0f, clamp_hi = 0.0f;
  float drop_p = 0.0f;
  unsigned long long seed = 0, offset = 0;
  const float* mask_src = nullptr; int64_t ld_mask = 0; float mask_scale = 1.0f;
  int accumulate = 0;
  int split_k = 0;
};
int rvk_sgemm_launch(const SgemmArgs& a, cudaStream_t stream);
int rvk_colsum_small_launch(const float* src, int64_t ld, int rows, int cols, float* out, cudaStream_t stream);
int rvk_epilogue_grad_launch(const float* y, const float* g, float* out, int relu, float scale, float lo, float hi,
                             int n, cudaStream_t stream);

struct JointLossArgs {
  const float* cls_logits = nullptr; int num_classes = 4;
  const float* ord_logits = nullptr;
  const float* mu = nullptr; const float* log_var = nullptr;
  const float* kan = nullptr;
  const int64_t* class_t = nullptr; const float* sev_t = nullptr;
  const float* alpha = nullptr;
  float gamma = 2.0f, lambda_ord = 1.0f, mu_unc = 0.5f, nu_kan = 0.5f;
  int batch = 0;
  float* sums_ws = nullptr;    // [4] scratch
  float* out = nullptr;        // [5] cls, ord, unc, kan, total
  float* d_cls = nullptr; float* d_ord = nullptr; float* d_mu = nullptr; float* d_lv = nullptr; float* d_kan = nullptr;
};
int rvk_joint_loss_launch(const JointLossArgs& a, cudaStream_t stream);
int rvk_predict_decode_launch(const float* cls, int num_classes, const float* ordl, const float* log_var, int batch,
                              long long* cls_idx, float* probs, float* ord_probs, float* ord_sev, float* unc_std,
                              cudaStream_t stream);
int rvk_loss_scale_grad_launch(const float* local, const float* upstream5, int term, float w_total, float* dst, int n,
                               cudaStream_t stream);
