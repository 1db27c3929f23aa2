// extern "C" surface declared in include/rovitkan.h: argument checking + dispatch to the launchers.
#include "../../include/rovitkan.h"

#include <cmath>
#include <cstdlib>

#include "encoder.h"
#include "kernels.h"

const char* rvk_last_error_cstr();
long long rvk_launch_count_impl();
void rvk_timing_enable_impl(int on);
int rvk_timing_collect_impl();
int rvk_timing_kind_impl(int kind, double* ms, double* flops, double* bytes);
const char* rvk_timing_kind_name_impl(int kind);

void rvk_set_last_error_text(const char* what);

namespace {
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

// The KAN kernels are written for the reference's knot buffer, linspace(-1, 1, 11) (models/kan.py:47-48: never trained, never
// changed by the reference): closed uniform-knot cubic segments, no clamp of tanh(x) to the knot range (it covers tanh's
// image), interval index from 5 (tanh x + 1).  The reference's Cox-de Boor recursion would also accept another knot vector;
// these kernels would evaluate a different spline for it -- refuse it here.
int check_knots(const float* k, int n) {
  if (k == nullptr) return RVK_ERR_BAD_ARG;
  if (n != 11) return RVK_ERR_UNSUPPORTED_SHAPE;   // reference config: 5 + 2*3 knots, 7 basis functions
  for (int i = 0; i < 11; ++i) {
    if (!(fabsf(k[i] - (-1.0f + 0.2f * static_cast<float>(i))) <= 2e-6f)) {
      rvk_set_last_error_text("KAN knots must be the reference's linspace(-1, 1, 11) (closed-form uniform cubic B-spline kernels)");
      return RVK_ERR_UNSUPPORTED_SHAPE;
    }
  }
  return RVK_OK;
}

int fill_kan_desc(KanLayerDesc& L, const float* spline, const float* lin_w, const float* lin_b,
                  const float* knots_host, int num_knots_total, int in_features, int out_features) {
  if (knots_host == nullptr || in_features <= 0 || out_features <= 0) return RVK_ERR_BAD_ARG;
  RVK_TRY(check_knots(knots_host, num_knots_total));
  L.spline = spline; L.lin_w = lin_w; L.lin_b = lin_b;
  for (int i = 0; i < 11; ++i) L.knots_host[i] = knots_host[i];
  L.num_knots = 11; L.num_basis = 7;
  L.in_features = in_features; L.out_features = out_features;
  return RVK_OK;
}
}  // namespace

static long long* g_mlp_trace = nullptr;
void rvk_debug_set_attn_trace_impl(void* buf);

#pragma GCC visibility push(default)
extern "C" {

int rvk_abi_version(void) { return RVK_ABI_VERSION; }

const char* rvk_strerror(int status) {
  switch (status) {
    case RVK_OK: return "ok";
    case RVK_ERR_BAD_ARG: return "bad argument (null pointer or non-positive size)";
    case RVK_ERR_CUDA: return "CUDA runtime error (see rvk_last_error)";
    case RVK_ERR_UNSUPPORTED_SHAPE: return "shape not supported by the sm_100a kernels";
    case RVK_ERR_TMA_ENCODE: return "cuTensorMapEncodeTiled failed (see rvk_last_error)";
    case RVK_ERR_NO_DRIVER: return "CUDA driver entry point cuTensorMapEncodeTiled unavailable (no GPU driver?)";
    case RVK_ERR_WORKSPACE: return "workspace too small";
    case RVK_ERR_ALIGNMENT: return "pointer or leading dimension not 16-byte aligned";
    default: return "unknown status";
  }
}
const char* rvk_last_error(void) { return rvk_last_error_cstr(); }

int rvk_device_check(void) {
  int dev = 0;
  RVK_CUDA_TRY(cudaGetDevice(&dev));
  int major = 0;
  RVK_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? RVK_OK : RVK_ERR_UNSUPPORTED_SHAPE;
}

int64_t rvk_launch_count(void) { return rvk_launch_count_impl(); }
int rvk_stream_check(void* stream) {
  cudaError_t e = cudaStreamSynchronize(S(stream));
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    rvk_set_last_cuda_error(static_cast<int>(e), "asynchronous kernel fault");
    return RVK_ERR_CUDA;
  }
  return RVK_OK;
}
void rvk_set_side_stream(int on) { rvk_set_side_stream_impl(on); }
int rvk_debug_mbar_timeout(void* stream) { return rvk_debug_mbar_timeout_launch(S(stream)); }
void rvk_timing_enable(int on) { rvk_timing_enable_impl(on); }
int rvk_timing_collect(void) { return rvk_timing_collect_impl(); }
int rvk_timing_kind(int kind, double* ms_host, double* flops_host, double* bytes_host) {
  return rvk_timing_kind_impl(kind, ms_host, flops_host, bytes_host);
}
const char* rvk_timing_kind_name(int kind) { return rvk_timing_kind_name_impl(kind); }
// round-1 names: the three tcgen05 GEMM kernels only (kinds 0..2)
void rvk_gemm_timing_enable(int on) { rvk_timing_enable_impl(on); }
int rvk_gemm_timing_kind(int kind, double* ms_host, double* flops_host) {
  return (kind >= 0 && kind <= 2) ? rvk_timing_kind_impl(kind, ms_host, flops_host, nullptr) : -1;
}
int rvk_gemm_timing_collect(double* total_ms_host, double* total_flops_host) {
  rvk_timing_collect_impl();
  double ms = 0.0, fl = 0.0;
  int n = 0;
  for (int k = 0; k <= 2; ++k) {
    double m = 0.0, f = 0.0;
    n += rvk_timing_kind_impl(k, &m, &f, nullptr);
    ms += m; fl += f;
  }
  if (total_ms_host) *total_ms_host = ms;
  if (total_flops_host) *total_flops_host = fl;
  return n;
}

// ---- KAN
int64_t rvk_kan_layer_workspace_floats(int in_features, int out_features, int with_backward) {
  return rvk_kan_workspace_floats(in_features, out_features, with_backward & 1);
}
int rvk_kan_layer_forward(const float* x, const float* spline, const float* lin_w, const float* lin_b,
                          const float* knots_host, int num_knots_total, int batch, int in_features,
                          int out_features, int act, float* y, float* workspace, int with_backward, void* stream) {
  if (batch == 0) return RVK_OK;
  if (x == nullptr || spline == nullptr || lin_w == nullptr || lin_b == nullptr || y == nullptr ||
      workspace == nullptr || batch < 0 || act < 0 || act > 2)
    return RVK_ERR_BAD_ARG;
  KanLayerDesc L;
  RVK_TRY(fill_kan_desc(L, spline, lin_w, lin_b, knots_host, num_knots_total, in_features, out_features));
  return rvk_kan_layer_fwd_launch(L, x, y, act, batch, workspace, with_backward, S(stream));
}
int rvk_kan_layer_backward(const float* x, const float* y, const float* gy, const float* spline, const float* lin_w,
                           const float* knots_host, int num_knots_total, int batch, int in_features,
                           int out_features, int act, float* dx, float* dspline, float* dlin_w, float* dlin_b,
                           float* workspace, void* stream) {
  if (batch == 0) return RVK_OK;
  if (x == nullptr || y == nullptr || gy == nullptr || workspace == nullptr || batch < 0 || act < 0 || act > 2)
    return RVK_ERR_BAD_ARG;
  if ((dspline == nullptr) != (dlin_w == nullptr) || (dspline == nullptr) != (dlin_b == nullptr)) return RVK_ERR_BAD_ARG;
  KanLayerDesc L;
  RVK_TRY(fill_kan_desc(L, spline, lin_w, nullptr, knots_host, num_knots_total, in_features, out_features));
  return rvk_kan_layer_bwd_launch(L, x, y, gy, act, dx, dspline, dlin_w, dlin_b, batch, workspace, S(stream));
}

int rvk_kan_basis(const float* t, const float* knots_host, int num_knots_total, int64_t n, float* out, void* stream) {
  if (n == 0) return RVK_OK;
  if (t == nullptr || knots_host == nullptr || out == nullptr || n < 0) return RVK_ERR_BAD_ARG;
  RVK_TRY(check_knots(knots_host, num_knots_total));
  return rvk_kan_basis_launch(t, knots_host, out, n, S(stream));
}

// ---- heads
int rvk_linear_forward(const float* x, const float* w, const float* b, int batch, int in_features, int out_features,
                       int relu, float drop_p, uint64_t seed, uint64_t offset, float clamp_lo, float clamp_hi,
                       float* y, void* stream) {
  if (batch == 0) return RVK_OK;
  if (x == nullptr || w == nullptr || y == nullptr || batch < 0 || in_features <= 0 || out_features <= 0 ||
      drop_p < 0.0f || drop_p >= 1.0f)
    return RVK_ERR_BAD_ARG;
  SgemmArgs a;
  a.A = x; a.lda = in_features; a.transA = 0;
  a.B = w; a.ldb = in_features; a.transB = 1;
  a.C = y; a.ldc = out_features;
  a.M = batch; a.N = out_features; a.K = in_features;
  a.bias = b; a.relu = relu;
  a.clamp_lo = clamp_lo; a.clamp_hi = clamp_hi;
  a.drop_p = drop_p; a.seed = seed; a.offset = offset;
  return rvk_sgemm_launch(a, S(stream));
}
int rvk_linear_backward(const float* x, const float* w, const float* y, const float* gy, int batch, int in_features,
                        int out_features, int relu, float drop_p, float clamp_lo, float clamp_hi, float* dx,
                        int accumulate_dx, float* dw, float* db, float* gpre_ws, void* stream) {
  if (batch == 0) return RVK_OK;
  if (x == nullptr || w == nullptr || y == nullptr || gy == nullptr || gpre_ws == nullptr || batch < 0 ||
      in_features <= 0 || out_features <= 0 || drop_p < 0.0f || drop_p >= 1.0f)
    return RVK_ERR_BAD_ARG;
  cudaStream_t s = S(stream);
  const float* gpre = gy;
  if (relu || clamp_lo < clamp_hi) {
    RVK_TRY(rvk_epilogue_grad_launch(y, gy, gpre_ws, relu, 1.0f / (1.0f - drop_p), clamp_lo, clamp_hi,
                                     batch * out_features, s));
    gpre = gpre_ws;
  }
  if (dx != nullptr) {   // dx = gpre W
    SgemmArgs a;
    a.A = gpre; a.lda = out_features; a.B = w; a.ldb = in_features; a.C = dx; a.ldc = in_features;
    a.M = batch; a.N = in_features; a.K = out_features; a.accumulate = accumulate_dx;
    RVK_TRY(rvk_sgemm_launch(a, s));
  }
  if (dw != nullptr) {   // dW += gpre^T x   (reduction over the batch, split across CTAs)
    SgemmArgs a;
    a.A = gpre; a.lda = out_features; a.transA = 1; a.B = x; a.ldb = in_features; a.C = dw; a.ldc = in_features;
    a.M = out_features; a.N = in_features; a.K = batch; a.accumulate = 1; a.split_k = 1;
    RVK_TRY(rvk_sgemm_launch(a, s));
  }
  if (db != nullptr) RVK_TRY(rvk_colsum_small_launch(gpre, out_features, batch, out_features, db, s));
  return RVK_OK;
}

// ---- fused inference tail
int64_t rvk_heads_fused_workspace_floats(void) { return rvk_heads_fused_workspace_floats_impl(); }
int rvk_heads_fused_prepare(const void* const* params23_host, float* ws, void* stream) {
  return rvk_heads_fused_prepare_launch(params23_host, ws, S(stream));
}
int rvk_heads_fused(const float* features, const float* ws, const float* knots_host, int batch, float* cls_logits,
                    float* ordinal_logits, float* mu, float* log_var, float* kan_severity, void* stream) {
  if (batch < 0) return RVK_ERR_BAD_ARG;
  RVK_TRY(check_knots(knots_host, 11));
  return rvk_heads_fused_launch(features, ws, knots_host, batch, cls_logits, ordinal_logits, mu, log_var, kan_severity,
                                S(stream));
}

// ---- loss
int rvk_joint_loss_forward(const float* cls_logits, int num_classes, const float* ord_logits, const float* mu,
                           const float* log_var, const float* kan, const int64_t* class_targets,
                           const float* severity_targets, const float* alpha, float gamma, float lambda_ord,
                           float mu_unc, float nu_kan, int batch, float* sums_ws4, float* out5, float* d_cls,
                           float* d_ord, float* d_mu, float* d_lv, float* d_kan, void* stream) {
  JointLossArgs a;
  a.cls_logits = cls_logits; a.num_classes = num_classes; a.ord_logits = ord_logits;
  a.mu = mu; a.log_var = log_var; a.kan = kan;
  a.class_t = class_targets; a.sev_t = severity_targets; a.alpha = alpha;
  a.gamma = gamma; a.lambda_ord = lambda_ord; a.mu_unc = mu_unc; a.nu_kan = nu_kan;
  a.batch = batch; a.sums_ws = sums_ws4; a.out = out5;
  a.d_cls = d_cls; a.d_ord = d_ord; a.d_mu = d_mu; a.d_lv = d_lv; a.d_kan = d_kan;
  return rvk_joint_loss_launch(a, S(stream));
}
int rvk_joint_loss_backward(const float* local, const float* upstream5, int term, float w_total, float* dst, int n,
                            void* stream) {
  if (n == 0) return RVK_OK;
  if (local == nullptr || upstream5 == nullptr || dst == nullptr || term < 0 || term > 3 || n < 0) return RVK_ERR_BAD_ARG;
  return rvk_loss_scale_grad_launch(local, upstream5, term, w_total, dst, n, S(stream));
}

// ---- encoder
int64_t rvk_encoder_weight_bytes(int training) { return rvk_encoder_weight_bytes_impl(training); }
int64_t rvk_encoder_workspace_bytes(int batch, int training, int chunk_images) {
  return rvk_encoder_workspace_bytes_impl(batch, training, chunk_images);
}
int rvk_encoder_prepare_weights(const void* const* params_host, void* wbuf, int training, void* stream) {
  return rvk_encoder_prepare_weights_impl(params_host, wbuf, training, S(stream));
}
int rvk_encoder_forward(const void* const* params_host, const void* wbuf, const float* images, int batch, int training,
                        int chunk_images, void* workspace, float* features, void* stream) {
  if (batch < 0) return RVK_ERR_BAD_ARG;
  return rvk_encoder_forward_impl(params_host, wbuf, images, 0, nullptr, batch, training, chunk_images, workspace, features, S(stream));
}
int rvk_encoder_forward_bf16(const void* const* params_host, const void* wbuf, const void* images_bf16, int batch,
                             int training, int chunk_images, void* workspace, float* features, void* stream) {
  if (batch < 0) return RVK_ERR_BAD_ARG;
  return rvk_encoder_forward_impl(params_host, wbuf, images_bf16, 1, nullptr, batch, training, chunk_images, workspace, features,
                                  S(stream));
}
int rvk_encoder_forward_u8(const void* const* params_host, const void* wbuf, const uint8_t* images_u8, const float* scale3_host,
                           const float* shift3_host, int batch, int training, int chunk_images, void* workspace,
                           float* features, void* stream) {
  if (batch < 0 || scale3_host == nullptr || shift3_host == nullptr) return RVK_ERR_BAD_ARG;
  const float norm6[6] = {scale3_host[0], scale3_host[1], scale3_host[2], shift3_host[0], shift3_host[1], shift3_host[2]};
  return rvk_encoder_forward_impl(params_host, wbuf, images_u8, 2, norm6, batch, training, chunk_images, workspace, features,
                                  S(stream));
}
int rvk_encoder_backward(const void* const* params_host, const void* wbuf, void* workspace, const float* dfeatures,
                         int batch, int chunk_images, void* const* grads_host, void* stream) {
  if (batch < 0) return RVK_ERR_BAD_ARG;
  return rvk_encoder_backward_impl(params_host, wbuf, workspace, dfeatures, batch, chunk_images, grads_host, 0,
                                   RVK_ENCODER_BACKWARD_STAGES, S(stream));
}
int rvk_encoder_backward_range(const void* const* params_host, const void* wbuf, void* workspace, const float* dfeatures,
                               int batch, int chunk_images, void* const* grads_host, int stage_begin, int stage_end,
                               void* stream) {
  if (batch < 0) return RVK_ERR_BAD_ARG;
  return rvk_encoder_backward_impl(params_host, wbuf, workspace, dfeatures, batch, chunk_images, grads_host, stage_begin,
                                   stage_end, S(stream));
}

int64_t rvk_encoder_saved_offset(int batch, int block, int which) { return rvk_encoder_saved_offset_impl(batch, block, which); }
int rvk_attention_probs(const void* qkv_bf16, float* probs, int batch, void* stream) {
  if (batch < 0 || (batch > 0 && (qkv_bf16 == nullptr || probs == nullptr))) return RVK_ERR_BAD_ARG;
  return rvk_attention_probs_launch(qkv_bf16, probs, batch, S(stream));
}

int rvk_predict_decode(const float* cls_logits, int num_classes, const float* ordinal_logits, const float* log_var, int batch,
                       int64_t* class_index, float* class_probs, float* ordinal_probs, float* ordinal_severity,
                       float* uncertainty_std, void* stream) {
  if (batch < 0) return RVK_ERR_BAD_ARG;
  return rvk_predict_decode_launch(cls_logits, num_classes, ordinal_logits, log_var, batch,
                                   reinterpret_cast<long long*>(class_index), class_probs, ordinal_probs, ordinal_severity,
                                   uncertainty_std, S(stream));
}

// ---- fused multi-task tail, training
int rvk_heads_train_forward(const float* features, const float* ws, const float* knots_host, int batch, float drop_p,
                            uint64_t seed, uint64_t offset, float* cls_logits, float* ordinal_logits, float* mu, float* log_var,
                            float* kan_severity, float* h_save, float* a1_save, float* a2_save, void* stream) {
  if (batch < 0) return RVK_ERR_BAD_ARG;
  RVK_TRY(check_knots(knots_host, 11));
  return rvk_heads_train_fwd_launch(features, ws, knots_host, batch, drop_p, seed, offset, cls_logits, ordinal_logits, mu, log_var,
                                    kan_severity, h_save, a1_save, a2_save, S(stream));
}
int rvk_heads_train_backward(const float* features, const float* ws, const float* knots_host, int batch, float drop_p,
                             const float* h_save, const float* a1_save, const float* a2_save, const float* log_var,
                             const float* kan_severity, const float* d_cls, const float* d_ord, const float* d_mu,
                             const float* d_log_var, const float* d_kan, float* dfeatures, float* dws,
                             float* const* grads23_host, void* stream) {
  if (batch < 0) return RVK_ERR_BAD_ARG;
  RVK_TRY(check_knots(knots_host, 11));
  return rvk_heads_train_bwd_launch(features, ws, knots_host, batch, drop_p, h_save, a1_save, a2_save, log_var, kan_severity, d_cls,
                                    d_ord, d_mu, d_log_var, d_kan, dfeatures, dws, grads23_host, S(stream));
}

// ---- fused optimizer tail
int64_t rvk_optimizer_state_floats(int n_tensors, const int64_t* numel_host) {
  if (n_tensors < 0 || (n_tensors > 0 && numel_host == nullptr)) return -1;
  return rvk_optimizer_state_floats_impl(n_tensors, numel_host);
}
int rvk_optimizer_step(int n_tensors, void* const* params_host, const void* const* grads_host, const int64_t* numel_host,
                       const int* group_host, float* exp_avg, float* exp_avg_sq, float* state4, const double* lr_host,
                       int n_groups, double beta1, double beta2, double eps, double weight_decay, float max_grad_norm,
                       float grad_mult, const float* grad_scale_dev, const float* found_inf_dev, void* stream) {
  if (n_tensors < 0) return RVK_ERR_BAD_ARG;
  return rvk_optimizer_step_impl(n_tensors, params_host, grads_host, numel_host, group_host, exp_avg, exp_avg_sq, state4,
                                 lr_host, n_groups, beta1, beta2, eps, weight_decay, max_grad_norm, grad_mult, grad_scale_dev,
                                 found_inf_dev, S(stream));
}

// ---- individual kernels
int rvk_gemm_nt(int mode, const void* a_bf16, int64_t lda, const void* b_bf16, int64_t ldb, void* out, int64_t ldo,
                void* out2, int64_t ldo2, const void* aux, int64_t ldaux, int m, int n, int k, const float* bias,
                const float* gamma, const float* beta, const float* res_table, int table_rows, float ln_eps,
                float* mean_out, float* rstd_out, void* stream) {
  if (m < 0) return RVK_ERR_BAD_ARG;
  GemmNtArgs a;
  a.mode = mode;
  a.A = a_bf16; a.lda = lda; a.B = b_bf16; a.ldb = ldb; a.out = out; a.ldo = ldo;
  a.out2 = out2; a.ldo2 = ldo2; a.aux = aux; a.ldaux = ldaux;
  a.p = GemmNtParams{};
  a.p.M = m; a.p.N = n; a.p.K = k; a.p.bias = bias; a.p.gamma = gamma; a.p.beta = beta;
  a.p.res_table = res_table; a.p.table_rows = table_rows; a.p.ln_eps = ln_eps;
  a.p.mean_out = mean_out; a.p.rstd_out = rstd_out;
  a.p.has_out2 = out2 != nullptr ? 1 : 0;
  a.p.has_res = (aux != nullptr || res_table != nullptr) ? 1 : 0;
  return rvk_gemm_nt_launch(a, S(stream));
}
static int mlp_fused_common(const float* x_in_tiled, float* x_out_tiled, const void* ctx_bf16, const void* wproj_bf16,
                            const float* bproj, const float* gamma2, const float* beta2, const void* w1_bf16, const float* b1,
                            const void* w2_f16, const float* b2, const float* gamma, const float* beta, float eps,
                            void* ln_out_bf16, int m, int cta_group, void* stream) {
  if (m < 0) return RVK_ERR_BAD_ARG;
  MlpFusedArgs a;
  a.w1 = w1_bf16; a.w2_f16 = w2_f16; a.ln_out = ln_out_bf16;
  a.ctx = ctx_bf16; a.wproj = wproj_bf16;
  a.cta_group = cta_group;
  a.p.M = m; a.p.x_in = x_in_tiled; a.p.x_out = x_out_tiled; a.p.gamma2 = gamma2; a.p.beta2 = beta2;
  a.p.b1 = b1; a.p.b2 = b2; a.p.gamma = gamma; a.p.beta = beta; a.p.eps = eps;
  a.p.bp = bproj;
  a.p.has_proj = ctx_bf16 != nullptr ? 1 : 0;
  a.p.has_ln = ln_out_bf16 != nullptr ? 1 : 0;
  a.p.trace = g_mlp_trace;
  return rvk_mlp_fused_launch(a, S(stream));
}
int rvk_mlp_fused(const float* x_in_tiled, float* x_out_tiled, const float* gamma2, const float* beta2,
                  const void* w1_bf16, const float* b1, const void* w2_f16, const float* b2, const float* gamma,
                  const float* beta, float eps, void* ln_out_bf16, int m, int cta_group, void* stream) {
  return mlp_fused_common(x_in_tiled, x_out_tiled, nullptr, nullptr, nullptr, gamma2, beta2, w1_bf16, b1, w2_f16, b2, gamma,
                          beta, eps, ln_out_bf16, m, cta_group, stream);
}
int rvk_attn_proj_mlp_fused(const float* x_in_tiled, float* x_out_tiled, const void* ctx_bf16, const void* wproj_bf16,
                            const float* bproj, const float* gamma2, const float* beta2, const void* w1_bf16, const float* b1,
                            const void* w2_f16, const float* b2, const float* gamma, const float* beta, float eps,
                            void* ln_out_bf16, int m, int cta_group, void* stream) {
  if (m > 0 && (ctx_bf16 == nullptr || wproj_bf16 == nullptr || bproj == nullptr)) return RVK_ERR_BAD_ARG;
  return mlp_fused_common(x_in_tiled, x_out_tiled, ctx_bf16, wproj_bf16, bproj, gamma2, beta2, w1_bf16, b1, w2_f16, b2, gamma,
                          beta, eps, ln_out_bf16, m, cta_group, stream);
}
/* debugging aid (not part of the product path): device buffer of 4*512 int64 that the next rvk_mlp_fused launches log clock events into */
void rvk_debug_set_mlp_trace(void* buf) { g_mlp_trace = static_cast<long long*>(buf); }
void rvk_debug_set_attn_trace(void* buf) { rvk_debug_set_attn_trace_impl(buf); }
int rvk_gemm_tn(const void* a_bf16, int64_t lda, const void* b_bf16, int64_t ldb, float* c, int64_t ldc, int m, int p,
                int q, float scale, float* a_colsum, void* stream) {
  if (m < 0) return RVK_ERR_BAD_ARG;
  return rvk_gemm_tn_launch(a_bf16, lda, b_bf16, ldb, c, ldc, m, p, q, scale, a_colsum, S(stream));
}
int rvk_attention_forward(const void* qkv, void* ctx, float* lse, int batch, void* stream) {
  if (batch < 0 || (batch > 0 && (qkv == nullptr || ctx == nullptr))) return RVK_ERR_BAD_ARG;
  return rvk_attention_fwd_launch(qkv, ctx, lse, batch, S(stream));
}
int rvk_attention_backward(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv, int batch,
                           void* stream) {
  if (batch < 0 || (batch > 0 && (qkv == nullptr || ctx == nullptr || dctx == nullptr || lse == nullptr || dqkv == nullptr)))
    return RVK_ERR_BAD_ARG;
  return rvk_attention_bwd_launch(qkv, ctx, dctx, lse, dqkv, batch, S(stream));
}
int rvk_layernorm_forward(const float* x, int64_t x_row_stride, const float* gamma, const float* beta, float eps,
                          void* y, int y_is_bf16, int64_t y_row_stride, float* mean, float* rstd, int rows,
                          void* stream) {
  if (rows < 0 || (rows > 0 && (x == nullptr || gamma == nullptr || beta == nullptr || y == nullptr))) return RVK_ERR_BAD_ARG;
  return rvk_layernorm_fwd_launch(x, x_row_stride, gamma, beta, eps, y, y_is_bf16, y_row_stride, mean, rstd, rows, S(stream));
}
int rvk_layernorm_backward(const void* g, int g_is_bf16, int64_t g_row_stride, const float* x, int64_t x_row_stride,
                           const float* mean, const float* rstd, const float* gamma, const float* dx_in, float* dx_out,
                           int64_t dx_row_stride, void* dx_out_bf16, float* dgamma, float* dbeta, float* dcolsum,
                           int rows, void* stream) {
  if (rows < 0 || (rows > 0 && (g == nullptr || x == nullptr || mean == nullptr || rstd == nullptr || gamma == nullptr ||
                                dx_out == nullptr)))
    return RVK_ERR_BAD_ARG;
  if ((dgamma == nullptr) != (dbeta == nullptr)) return RVK_ERR_BAD_ARG;
  return rvk_layernorm_bwd_launch(g, g_is_bf16, g_row_stride, x, x_row_stride, mean, rstd, gamma, dx_in, dx_out,
                                  dx_row_stride, dx_out_bf16, dgamma, dbeta, dcolsum, rows, S(stream));
}
int rvk_im2col(const float* images, void* patches_bf16, int batch, void* stream) {
  if (batch < 0 || (batch > 0 && (images == nullptr || patches_bf16 == nullptr))) return RVK_ERR_BAD_ARG;
  return rvk_im2col_launch(images, 0, patches_bf16, batch, nullptr, S(stream));
}
int rvk_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (src == nullptr || dst_bf16 == nullptr))) return RVK_ERR_BAD_ARG;
  return rvk_cast_bf16_launch(src, dst_bf16, n, S(stream));
}

}  // extern "C"
#pragma GCC visibility pop
