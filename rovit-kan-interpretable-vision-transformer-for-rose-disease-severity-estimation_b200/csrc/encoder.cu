// DeiT-Tiny trunk (timm deit_tiny_patch16_224, num_classes=0 -- what reference models/backbone.py:12-25
// runs) as a fixed launch sequence over the sm_100a kernels.  Host code only: lays out the workspace,
// walks the 12 blocks, never allocates, never synchronises.
//
// Token stream layout in HBM (M = images * 197 rows):
//   x      fp32 [M,192]  residual stream (kept fp32: it is the sum of 25 bf16-GEMM outputs)
//   ln     bf16 [M,192]  LayerNorm output = next GEMM's A operand (produced by the previous GEMM's epilogue)
//   qkv    bf16 [M,576]  columns [3][head][64] as timm's reshape expects
//   ctx    bf16 [M,192]  attention output, columns [head][64]
//   h / z  bf16 [M,768]  MLP hidden after / before GELU
// Inference walks the batch in chunks of `chunk_images` so that a chunk's inter-kernel tensors
// (<= 58 MB at 192 images) stay in the 126 MB L2 between producer and consumer; training keeps every
// chunk's activations for the backward pass.
#include "encoder.h"

#include "kernels.h"

#include <cstdlib>

namespace {

constexpr int kD = 192, kTok = 197, kDepth = 12, kMlp = 768, kPatchK = 768, kQkv = 576;
constexpr float kLnEps = 1e-6f;

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// ---- bf16 weight buffer ----------------------------------------------------------------------------
struct WeightLayout {
  size_t patch_w;                    // [192,768]
  size_t qkv[kDepth], proj[kDepth], fc1[kDepth], fc2[kDepth];          // torch layout [out,in]
  size_t qkvT[kDepth], projT[kDepth], fc1T[kDepth], fc2T[kDepth];      // transposed [in,out] (training only)
  size_t fc2h[kDepth];               // fp16 copy of fc2 for the fused inference MLP (hidden activation in fp16)
  size_t table;                      // fp32 [197,192]
  size_t total;
};
WeightLayout weight_layout(bool training) {
  WeightLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  L.patch_w = take(size_t(kD) * kPatchK * 2);
  for (int i = 0; i < kDepth; ++i) {
    L.qkv[i] = take(size_t(kQkv) * kD * 2);
    L.proj[i] = take(size_t(kD) * kD * 2);
    L.fc1[i] = take(size_t(kMlp) * kD * 2);
    L.fc2[i] = take(size_t(kD) * kMlp * 2);
    if (!training) L.fc2h[i] = take(size_t(kD) * kMlp * 2);
    if (training) {
      L.qkvT[i] = take(size_t(kQkv) * kD * 2);
      L.projT[i] = take(size_t(kD) * kD * 2);
      L.fc1T[i] = take(size_t(kMlp) * kD * 2);
      L.fc2T[i] = take(size_t(kD) * kMlp * 2);
    }
  }
  L.table = take(size_t(kTok) * kD * 4);
  L.total = off;
  return L;
}

// ---- activation workspace ------------------------------------------------------------------------------
struct BlockSaved {
  size_t x_in, ln1, qkv, ctx, x_mid, ln2, z, h, mean1, rstd1, mean2, rstd2, lse;
};
struct ActLayout {
  // training: per-chunk-independent offsets for the whole batch; inference: one chunk's worth, reused
  size_t patches;
  BlockSaved blk[kDepth];
  size_t x_final, mean_f, rstd_f;
  // backward temporaries (training only)
  size_t dx, dxb, g, dz, dctx, dqkv;
  size_t dxb2, dz2, dqkv2;     // second copies: the weight-gradient GEMMs on the side stream still read the first ones
  size_t total;
};
ActLayout act_layout(int64_t rows, int64_t images, bool training) {
  ActLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  const size_t R = static_cast<size_t>(rows);
  L.patches = take(R * kPatchK * 2);
  if (training) {
    for (int i = 0; i < kDepth; ++i) {
      BlockSaved& b = L.blk[i];
      b.x_in = take(R * kD * 4);
      b.ln1 = take(R * kD * 2);
      b.qkv = take(R * kQkv * 2);
      b.ctx = take(R * kD * 2);
      b.x_mid = take(R * kD * 4);
      b.ln2 = take(R * kD * 2);
      b.z = take(R * kMlp * 2);
      b.h = take(R * kMlp * 2);
      b.mean1 = take(R * 4);
      b.rstd1 = take(R * 4);
      b.mean2 = take(R * 4);
      b.rstd2 = take(R * 4);
      b.lse = take(static_cast<size_t>(images) * 3 * kTok * 4);
    }
    L.x_final = take(R * kD * 4);
    L.mean_f = take(static_cast<size_t>(images) * 4);
    L.rstd_f = take(static_cast<size_t>(images) * 4);
    L.dx = take(R * kD * 4);
    L.dxb = take(R * kD * 2);
    L.g = take(R * kD * 2);
    L.dz = take(R * kMlp * 2);
    L.dctx = take(R * kD * 2);
    L.dqkv = take(R * kQkv * 2);
    L.dxb2 = take(R * kD * 2);
    L.dz2 = take(R * kMlp * 2);
    L.dqkv2 = take(R * kQkv * 2);
  } else {
    BlockSaved b{};
    b.x_in = take((R + 127) / 128 * 128 * kD * 4);      // tiled layout (xt_offset): rows padded to 128
    b.x_mid = b.x_in;            // in-place residual updates
    b.ln1 = take(R * kD * 2);
    b.ln2 = b.ln1;
    b.qkv = take(R * kQkv * 2);
    b.ctx = take(R * kD * 2);
    b.h = take(R * kMlp * 2);
    b.z = 0;
    for (int i = 0; i < kDepth; ++i) L.blk[i] = b;
    L.x_final = b.x_in;
  }
  L.total = off;
  return L;
}

inline uint8_t* at(void* base, size_t off) { return static_cast<uint8_t*>(base) + off; }
inline const uint8_t* at(const void* base, size_t off) { return static_cast<const uint8_t*>(base) + off; }
inline const float* P(const void* const* params, int idx) { return static_cast<const float*>(params[idx]); }
inline float* G(void* const* grads, int idx) { return static_cast<float*>(grads[idx]); }

// parameter table indices (timm state_dict order)
enum { P_CLS = 0, P_POS = 1, P_PATCH_W = 2, P_PATCH_B = 3, P_BLOCK0 = 4, P_NORM_W = 4 + 12 * kDepth, P_NORM_B = P_NORM_W + 1 };
enum { B_N1W = 0, B_N1B, B_QKVW, B_QKVB, B_PROJW, B_PROJB, B_N2W, B_N2B, B_FC1W, B_FC1B, B_FC2W, B_FC2B };
inline int bp(int blk, int which) { return P_BLOCK0 + blk * 12 + which; }

int gemm_plain(const void* A, int64_t lda, const void* B, int64_t ldb, void* out, int64_t ldo, int M, int N, int K,
               const float* bias, cudaStream_t s) {
  GemmNtArgs a;
  a.mode = EPI_BF16;
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.out = out; a.ldo = ldo;
  a.p = GemmNtParams{};
  a.p.M = M; a.p.N = N; a.p.K = K; a.p.bias = bias;
  return rvk_gemm_nt_launch(a, s);
}

int gemm_res_ln(const void* A, int64_t lda, int K, const void* B, const float* bias, const float* residual,
                const float* table, float* x_out, void* ln_out, const float* gamma, const float* beta, float* mean,
                float* rstd, int M, cudaStream_t s, bool tiled = false) {
  GemmNtArgs a;
  a.mode = EPI_RES_LN;
  a.A = A; a.lda = lda; a.B = B; a.ldb = K; a.out = x_out; a.ldo = kD;
  a.out2 = ln_out; a.ldo2 = kD;
  a.aux = residual; a.ldaux = kD;
  a.p = GemmNtParams{};
  a.p.M = M; a.p.N = kD; a.p.K = K; a.p.bias = bias;
  a.p.gamma = gamma; a.p.beta = beta; a.p.ln_eps = kLnEps;
  a.p.res_table = table; a.p.table_rows = kTok;
  a.p.mean_out = mean; a.p.rstd_out = rstd;
  a.p.has_out2 = ln_out != nullptr ? 1 : 0;
  a.p.has_res = (residual != nullptr || table != nullptr) ? 1 : 0;
  if (tiled) {      // fp32 token stream in the tiled layout: no TMA for x
    a.out = nullptr; a.aux = nullptr;
    a.p.out_tiled = x_out;
    a.p.res_tiled = residual;
  }
  return rvk_gemm_nt_launch(a, s);
}

// inference uses the fused MLP kernel and the tiled token stream unless RVK_UNFUSED=1 (A/B switch for measurements)
bool fused_inference() {
  static const bool on = [] {
    const char* e = getenv("RVK_UNFUSED");
    return !(e != nullptr && e[0] == '1');
  }();
  return on;
}
bool fuse_proj() {        // RVK_FUSE_PROJ=0: keep the attention output projection as its own GEMM launch (A/B measurements)
  static const bool on = [] {
    const char* e = getenv("RVK_FUSE_PROJ");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}
// The weight-gradient GEMMs (off the critical path: nothing in the backward pass reads a weight gradient) run on a second,
// lower-priority stream next to the dgrad / LayerNorm / attention kernels of the main stream: 40.0 k -> 42.1 k img/s on
// the batch-256 train step (A/B in one gpurun call).  RVK_TN_SIDE_STREAM=0 keeps everything on the caller's stream.
int g_side_stream_override = -1;      // rvk_set_side_stream: -1 = environment default, 0 / 1 = forced
bool tn_side_stream() {
  static const bool on = [] { const char* e = getenv("RVK_TN_SIDE_STREAM"); return !(e != nullptr && e[0] == '0'); }();
  return g_side_stream_override < 0 ? on : g_side_stream_override != 0;
}
struct SideStream {
  cudaStream_t stream = nullptr;
  static constexpr int kEvents = 64;
  cudaEvent_t ev[kEvents];
  int next = 0;
  bool ok = false;
  cudaEvent_t take() { cudaEvent_t e = ev[next]; next = (next + 1) % kEvents; return e; }
};
SideStream* side_stream_for_current_device() {
  static SideStream per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& S = per_dev[dev];
  if (!S.ok) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);          // lo = least priority
    if (cudaStreamCreateWithPriority(&S.stream, cudaStreamNonBlocking, lo) != cudaSuccess) return nullptr;
    for (int i = 0; i < SideStream::kEvents; ++i)
      if (cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    S.ok = true;
  }
  return &S;
}

int mlp_cta_group() {
  static const int g = [] {
    const char* e = getenv("RVK_MLP_CTA_GROUP");
    return (e != nullptr && e[0] == '1') ? 1 : (e != nullptr && e[0] == '2') ? 2 : 4;   // 4 = CTA pairs, two row tiles in flight
  }();
  return g;
}

}  // namespace

void rvk_set_side_stream_impl(int on) { g_side_stream_override = on; }

int64_t rvk_encoder_weight_bytes_impl(int training) { return static_cast<int64_t>(weight_layout(training != 0).total); }

int64_t rvk_encoder_workspace_bytes_impl(int batch, int training, int chunk_images) {
  if (batch <= 0) return 0;
  const int chunk = (chunk_images <= 0 || chunk_images > batch) ? batch : chunk_images;
  const int64_t imgs = training ? batch : chunk;
  return static_cast<int64_t>(act_layout(imgs * kTok, imgs, training != 0).total);
}

// byte offset, inside a training workspace of `batch` images, of one tensor the forward saved for block `block`
// (which: 0 x_in fp32 [M,192], 1 ln1 bf16 [M,192], 2 qkv bf16 [M,576], 3 ctx bf16 [M,192], 4 x_mid fp32 [M,192], 5 ln2 bf16
// [M,192], 6 z bf16 [M,768], 7 h bf16 [M,768]); block == 12: which 0 = x_final fp32 [M,192].  -1 if out of range.
int64_t rvk_encoder_saved_offset_impl(int batch, int block, int which) {
  if (batch <= 0 || block < 0 || block > kDepth || which < 0 || which > 7) return -1;
  const ActLayout A = act_layout(int64_t(batch) * kTok, batch, true);
  if (block == kDepth) return which == 0 ? static_cast<int64_t>(A.x_final) : -1;
  const BlockSaved& B = A.blk[block];
  const size_t off[8] = {B.x_in, B.ln1, B.qkv, B.ctx, B.x_mid, B.ln2, B.z, B.h};
  return static_cast<int64_t>(off[which]);
}

int rvk_encoder_prepare_weights_impl(const void* const* params, void* wbuf, int training, cudaStream_t s) {
  if (params == nullptr || wbuf == nullptr) return RVK_ERR_BAD_ARG;
  const WeightLayout W = weight_layout(training != 0);
  RvkCastTable T{};
  auto add = [&](const float* src, void* dst, int rows, int cols, int mode) -> int {
    if (T.n == kRvkMaxCastJobs) {
      RVK_TRY(rvk_cast_multi_launch(T, s));
      T.n = 0;
    }
    T.job[T.n++] = RvkCastJob{src, dst, rows, cols, mode, 0};
    return RVK_OK;
  };
  RVK_TRY(add(P(params, P_PATCH_W), at(wbuf, W.patch_w), kD, kPatchK, 0));
  for (int i = 0; i < kDepth; ++i) {
    RVK_TRY(add(P(params, bp(i, B_QKVW)), at(wbuf, W.qkv[i]), kQkv, kD, 0));
    RVK_TRY(add(P(params, bp(i, B_PROJW)), at(wbuf, W.proj[i]), kD, kD, 0));
    RVK_TRY(add(P(params, bp(i, B_FC1W)), at(wbuf, W.fc1[i]), kMlp, kD, 0));
    RVK_TRY(add(P(params, bp(i, B_FC2W)), at(wbuf, W.fc2[i]), kD, kMlp, 0));
    if (!training) RVK_TRY(add(P(params, bp(i, B_FC2W)), at(wbuf, W.fc2h[i]), kD, kMlp, 2));
    if (training) {
      RVK_TRY(add(P(params, bp(i, B_QKVW)), at(wbuf, W.qkvT[i]), kQkv, kD, 1));
      RVK_TRY(add(P(params, bp(i, B_PROJW)), at(wbuf, W.projT[i]), kD, kD, 1));
      RVK_TRY(add(P(params, bp(i, B_FC1W)), at(wbuf, W.fc1T[i]), kMlp, kD, 1));
      RVK_TRY(add(P(params, bp(i, B_FC2W)), at(wbuf, W.fc2T[i]), kD, kMlp, 1));
    }
  }
  RVK_TRY(rvk_cast_multi_launch(T, s));
  return rvk_token_table_launch(P(params, P_CLS), P(params, P_POS), P(params, P_PATCH_B),
                                reinterpret_cast<float*>(at(wbuf, W.table)), s);
}

int rvk_encoder_forward_impl(const void* const* params, const void* wbuf, const void* images, int image_fmt, const float* norm6_host, int batch, int training,
                             int chunk_images, void* workspace, float* features, cudaStream_t s) {
  if (batch <= 0) return RVK_OK;
  if (params == nullptr || wbuf == nullptr || images == nullptr || workspace == nullptr || features == nullptr)
    return RVK_ERR_BAD_ARG;
  const bool train = training != 0;
  const WeightLayout W = weight_layout(train);
  const int chunk = (chunk_images <= 0 || chunk_images > batch) ? batch : chunk_images;
  const int64_t lay_imgs = train ? batch : chunk;
  const ActLayout A = act_layout(lay_imgs * kTok, lay_imgs, train);
  const float* table = reinterpret_cast<const float*>(at(wbuf, W.table));

  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int nb = (batch - b0 < chunk) ? batch - b0 : chunk;
    const int M = nb * kTok;
    // row offset of this chunk inside whole-batch buffers (training) or 0 (inference: buffers are per chunk)
    const size_t r0 = train ? static_cast<size_t>(b0) * kTok : 0;
    const size_t i0 = train ? static_cast<size_t>(b0) : 0;
    auto f32 = [&](size_t off, size_t width) { return reinterpret_cast<float*>(at(workspace, off) + r0 * width * 4); };
    auto b16 = [&](size_t off, size_t width) { return at(workspace, off) + r0 * width * 2; };
    auto stat = [&](size_t off) { return train ? reinterpret_cast<float*>(at(workspace, off)) + r0 : nullptr; };

    uint8_t* patches = b16(A.patches, kPatchK);
    RVK_TRY(rvk_im2col_launch(static_cast<const uint8_t*>(images) + static_cast<size_t>(b0) * 3 * 224 * 224 * (image_fmt == 0 ? 4 : (image_fmt == 1 ? 2 : 1)),
                              image_fmt, patches, nb, norm6_host, s));
    if (!train && fused_inference()) {
      // ---- inference: tiled fp32 token stream (updated in place), fused MLP block
      float* x = f32(A.blk[0].x_in, kD);
      uint8_t* ln = b16(A.blk[0].ln1, kD);
      RVK_TRY(gemm_res_ln(patches, kPatchK, kPatchK, at(wbuf, W.patch_w), nullptr, nullptr, table, x, ln,
                          P(params, bp(0, B_N1W)), P(params, bp(0, B_N1B)), nullptr, nullptr, M, s, true));
      for (int i = 0; i < kDepth; ++i) {
        const BlockSaved& B = A.blk[i];
        RVK_TRY(gemm_plain(ln, kD, at(wbuf, W.qkv[i]), kD, b16(B.qkv, kQkv), kQkv, M, kQkv, kD, P(params, bp(i, B_QKVB)), s));
        RVK_TRY(rvk_attention_fwd_launch(b16(B.qkv, kQkv), b16(B.ctx, kD), nullptr, nb, s));
        // x += proj(ctx): folded into the MLP kernel (default) or as its own GEMM; LayerNorm2 is applied on load by the
        // MLP kernel, so no normalised copy is written
        if (!fuse_proj())
          RVK_TRY(gemm_res_ln(b16(B.ctx, kD), kD, kD, at(wbuf, W.proj[i]), P(params, bp(i, B_PROJB)), x, nullptr, x, nullptr,
                              nullptr, nullptr, nullptr, nullptr, M, s, true));
        const bool last = (i == kDepth - 1);
        MlpFusedArgs m;
        m.w1 = at(wbuf, W.fc1[i]); m.w2_f16 = at(wbuf, W.fc2h[i]);
        m.ln_out = last ? nullptr : ln;
        m.cta_group = (mlp_cta_group() == 4 && !fuse_proj()) ? 2 : mlp_cta_group();   // 4 (two tiles in flight) needs the folded projection
        m.p.M = M; m.p.x_in = x; m.p.x_out = x;
        m.p.gamma2 = P(params, bp(i, B_N2W)); m.p.beta2 = P(params, bp(i, B_N2B));
        m.p.b1 = P(params, bp(i, B_FC1B)); m.p.b2 = P(params, bp(i, B_FC2B));
        m.p.gamma = last ? nullptr : P(params, bp(i + 1, B_N1W));
        m.p.beta = last ? nullptr : P(params, bp(i + 1, B_N1B));
        m.p.eps = kLnEps; m.p.has_ln = last ? 0 : 1; m.p.trace = nullptr;
        if (fuse_proj()) {
          m.ctx = b16(B.ctx, kD); m.wproj = at(wbuf, W.proj[i]);
          m.p.bp = P(params, bp(i, B_PROJB)); m.p.has_proj = 1;
        }
        RVK_TRY(rvk_mlp_fused_launch(m, s));
      }
      RVK_TRY(rvk_layernorm_fwd_tiled_launch(x, kTok, P(params, P_NORM_W), P(params, P_NORM_B), kLnEps,
                                             features + static_cast<size_t>(b0) * kD, kD, nb, s));
      continue;
    }
    // patch embedding + cls/pos/bias table -> x0, fused LayerNorm (block 0 norm1) -> ln1
    RVK_TRY(gemm_res_ln(patches, kPatchK, kPatchK, at(wbuf, W.patch_w), nullptr, nullptr, table, f32(A.blk[0].x_in, kD),
                        b16(A.blk[0].ln1, kD), P(params, bp(0, B_N1W)), P(params, bp(0, B_N1B)), stat(A.blk[0].mean1),
                        stat(A.blk[0].rstd1), M, s));
    for (int i = 0; i < kDepth; ++i) {
      const BlockSaved& B = A.blk[i];
      // qkv = ln1 * Wqkv^T + b
      RVK_TRY(gemm_plain(b16(B.ln1, kD), kD, at(wbuf, W.qkv[i]), kD, b16(B.qkv, kQkv), kQkv, M, kQkv, kD,
                         P(params, bp(i, B_QKVB)), s));
      float* lse = train ? reinterpret_cast<float*>(at(workspace, B.lse)) + i0 * 3 * kTok : nullptr;
      RVK_TRY(rvk_attention_fwd_launch(b16(B.qkv, kQkv), b16(B.ctx, kD), lse, nb, s));
      // x_mid = x_in + ctx * Wproj^T + b ; ln2 = LN2(x_mid)
      RVK_TRY(gemm_res_ln(b16(B.ctx, kD), kD, kD, at(wbuf, W.proj[i]), P(params, bp(i, B_PROJB)), f32(B.x_in, kD),
                          nullptr, f32(B.x_mid, kD), b16(B.ln2, kD), P(params, bp(i, B_N2W)), P(params, bp(i, B_N2B)),
                          stat(B.mean2), stat(B.rstd2), M, s));
      // h = gelu(ln2 * W1^T + b) (z kept for the backward pass)
      {
        GemmNtArgs a;
        a.mode = EPI_GELU;
        a.A = b16(B.ln2, kD); a.lda = kD; a.B = at(wbuf, W.fc1[i]); a.ldb = kD;
        a.out = b16(B.h, kMlp); a.ldo = kMlp;
        a.out2 = train ? b16(B.z, kMlp) : nullptr; a.ldo2 = kMlp;
        a.p = GemmNtParams{};
        a.p.M = M; a.p.N = kMlp; a.p.K = kD; a.p.bias = P(params, bp(i, B_FC1B));
        a.p.has_out2 = train ? 1 : 0;
        RVK_TRY(rvk_gemm_nt_launch(a, s));
      }
      // x_next = x_mid + h * W2^T + b ; ln1(next block) fused, except after the last block
      const bool last = (i == kDepth - 1);
      float* x_next = last ? f32(A.x_final, kD) : f32(A.blk[i + 1].x_in, kD);
      RVK_TRY(gemm_res_ln(b16(B.h, kMlp), kMlp, kMlp, at(wbuf, W.fc2[i]), P(params, bp(i, B_FC2B)), f32(B.x_mid, kD),
                          nullptr, x_next, last ? nullptr : b16(A.blk[i + 1].ln1, kD),
                          last ? nullptr : P(params, bp(i + 1, B_N1W)), last ? nullptr : P(params, bp(i + 1, B_N1B)),
                          last ? nullptr : stat(A.blk[i + 1].mean1), last ? nullptr : stat(A.blk[i + 1].rstd1), M, s));
    }
    // final LayerNorm on the class-token rows only -> features (fp32)
    float* mean_f = train ? reinterpret_cast<float*>(at(workspace, A.mean_f)) + i0 : nullptr;
    float* rstd_f = train ? reinterpret_cast<float*>(at(workspace, A.rstd_f)) + i0 : nullptr;
    RVK_TRY(rvk_layernorm_fwd_launch(f32(A.x_final, kD), int64_t(kTok) * kD, P(params, P_NORM_W), P(params, P_NORM_B),
                                     kLnEps, features + static_cast<size_t>(b0) * kD, 0, kD, mean_f, rstd_f, nb, s));
  }
  return RVK_OK;
}

// Stages of the backward pass: 0 = final LayerNorm (class-token rows), 1 + j = block 11 - j (j = 0..11), 13 = patch embedding /
// class token / position embedding.  [stage_begin, stage_end) lets the caller issue the gradient all-reduce of the blocks that
// are already complete while the remaining ones are still being computed (data-parallel training, dist.py).  Gradients
// written by stage s: its own tensors, except that mlp.fc2.bias of block i is written by the stage of block i + 1 (by stage 0
// for block 11): after stage 1 + j every tensor of blocks 11 - j .. 11, norm.* and mlp.fc2.bias of block 10 - j are final.
int rvk_encoder_backward_impl(const void* const* params, const void* wbuf, void* workspace, const float* dfeatures,
                              int batch, int chunk_images, void* const* grads, int stage_begin, int stage_end, cudaStream_t s) {
  if (batch <= 0) return RVK_OK;
  if (stage_begin < 0 || stage_end > kDepth + 2 || stage_begin >= stage_end) return RVK_ERR_BAD_ARG;
  if (params == nullptr || wbuf == nullptr || workspace == nullptr || dfeatures == nullptr || grads == nullptr)
    return RVK_ERR_BAD_ARG;
  const WeightLayout W = weight_layout(true);
  const int chunk = (chunk_images <= 0 || chunk_images > batch) ? batch : chunk_images;
  const ActLayout A = act_layout(int64_t(batch) * kTok, batch, true);

  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int nb = (batch - b0 < chunk) ? batch - b0 : chunk;
    const int M = nb * kTok;
    const size_t r0 = static_cast<size_t>(b0) * kTok;
    auto f32 = [&](size_t off, size_t width) { return reinterpret_cast<float*>(at(workspace, off) + r0 * width * 4); };
    auto b16 = [&](size_t off, size_t width) { return at(workspace, off) + r0 * width * 2; };
    auto stat = [&](size_t off) { return reinterpret_cast<float*>(at(workspace, off)) + r0; };
    float* dx = f32(A.dx, kD);
    uint8_t* dxb = b16(A.dxb, kD);
    uint8_t* g = b16(A.g, kD);
    uint8_t* dz = b16(A.dz, kMlp);
    uint8_t* dctx = b16(A.dctx, kD);
    uint8_t* dqkv = b16(A.dqkv, kQkv);
    // ---- optional side stream for the weight-gradient GEMMs.  `dxb` ping-pongs between two buffers at every LayerNorm
    // backward, dz / dqkv between two buffers per block, so the side stream's reads (launched up to one block earlier)
    // never race with the main stream's writes; `side_done_*` events order the remaining reuse.
    SideStream* side = tn_side_stream() ? side_stream_for_current_device() : nullptr;
    uint8_t* dxb_pp[2] = {dxb, b16(A.dxb2, kD)};
    uint8_t* dz_pp[2] = {dz, b16(A.dz2, kMlp)};
    uint8_t* dqkv_pp[2] = {dqkv, b16(A.dqkv2, kQkv)};
    int dxb_cur = 0;                                  // which dxb buffer holds the current gradient (bf16 copy)
    cudaEvent_t dxb_reader[2] = {nullptr, nullptr};   // last side-stream reader of each dxb buffer
    cudaEvent_t dz_reader[2] = {nullptr, nullptr}, dqkv_reader[2] = {nullptr, nullptr};
    cudaEvent_t last_side = nullptr;
    // weight gradient C += A^T B on the side stream (after everything enqueued on `s` so far), or inline
    auto tn = [&](const void* Ab, int64_t lda, const void* Bb, int64_t ldb, float* C, int64_t ldc, int Mr, int Pp, int Qq,
                  float* colsum, cudaEvent_t* reader_a) -> int {
      if (side == nullptr) return rvk_gemm_tn_launch(Ab, lda, Bb, ldb, C, ldc, Mr, Pp, Qq, 1.0f, colsum, s);
      cudaEvent_t ready = side->take();
      RVK_CUDA_TRY(cudaEventRecord(ready, s));
      RVK_CUDA_TRY(cudaStreamWaitEvent(side->stream, ready, 0));
      RVK_TRY(rvk_gemm_tn_launch(Ab, lda, Bb, ldb, C, ldc, Mr, Pp, Qq, 1.0f, colsum, side->stream));
      cudaEvent_t done = side->take();
      RVK_CUDA_TRY(cudaEventRecord(done, side->stream));
      if (reader_a != nullptr) *reader_a = done;
      last_side = done;
      return RVK_OK;
    };
    auto wait_reader = [&](cudaEvent_t e) -> int {     // main stream must not overwrite a buffer the side stream still reads
      if (side != nullptr && e != nullptr) RVK_CUDA_TRY(cudaStreamWaitEvent(s, e, 0));
      return RVK_OK;
    };

    if (chunk < batch && !(stage_begin == 0 && stage_end == kDepth + 2)) return RVK_ERR_UNSUPPORTED_SHAPE;   // ranges: one chunk only
    // final LayerNorm backward: only the class-token rows carry gradient
    if (stage_begin == 0) {
    RVK_CUDA_TRY(cudaMemsetAsync(dx, 0, size_t(M) * kD * 4, s));
    RVK_CUDA_TRY(cudaMemsetAsync(dxb, 0, size_t(M) * kD * 2, s));
    RVK_TRY(rvk_layernorm_bwd_launch(dfeatures + size_t(b0) * kD, 0, kD, f32(A.x_final, kD), int64_t(kTok) * kD,
                                     reinterpret_cast<float*>(at(workspace, A.mean_f)) + b0,
                                     reinterpret_cast<float*>(at(workspace, A.rstd_f)) + b0, P(params, P_NORM_W),
                                     nullptr, dx, int64_t(kTok) * kD, dxb, G(grads, P_NORM_W), G(grads, P_NORM_B),
                                     G(grads, bp(kDepth - 1, B_FC2B)), nb, s));
    }

    for (int i = kDepth - 1; i >= 0; --i) {
      const int stage = 1 + (kDepth - 1 - i);
      if (stage < stage_begin || stage >= stage_end) continue;
      const BlockSaved& B = A.blk[i];
      if (side != nullptr) {      // (stage ranges: every range starts with dxb in buffer 0, see the end of the block)
        dz = dz_pp[i & 1];
        dqkv = dqkv_pp[i & 1];
        RVK_TRY(wait_reader(dz_reader[i & 1]));
        RVK_TRY(wait_reader(dqkv_reader[i & 1]));
      }
      dxb = dxb_pp[dxb_cur];
      // ---- MLP: x_next = x_mid + fc2(gelu(fc1(ln2)))
      RVK_TRY(tn(dxb, kD, b16(B.h, kMlp), kMlp, G(grads, bp(i, B_FC2W)), kMlp, M, kD, kMlp, nullptr, &dxb_reader[dxb_cur]));
      // (fc2 bias gradient = column sums of dx: accumulated by the LayerNorm backward that wrote dx)
      {   // dz = (dx * W2) o gelu'(z)
        GemmNtArgs a;
        a.mode = EPI_DGELU;
        a.A = dxb; a.lda = kD; a.B = at(wbuf, W.fc2T[i]); a.ldb = kD;
        a.out = dz; a.ldo = kMlp; a.aux = b16(B.z, kMlp); a.ldaux = kMlp;
        a.p = GemmNtParams{};
        a.p.M = M; a.p.N = kMlp; a.p.K = kD;
        RVK_TRY(rvk_gemm_nt_launch(a, s));
      }
      // (fc1 bias gradient = column sums of dz: the ones column of the same GEMM)
      RVK_TRY(tn(dz, kMlp, b16(B.ln2, kD), kD, G(grads, bp(i, B_FC1W)), kD, M, kMlp, kD, G(grads, bp(i, B_FC1B)),
                 &dz_reader[i & 1]));
      RVK_TRY(gemm_plain(dz, kMlp, at(wbuf, W.fc1T[i]), kMlp, g, kD, M, kD, kMlp, nullptr, s));
      if (side != nullptr) {      // the LayerNorm backward writes the OTHER dxb buffer
        dxb_cur ^= 1;
        dxb = dxb_pp[dxb_cur];
        RVK_TRY(wait_reader(dxb_reader[dxb_cur]));
      }
      RVK_TRY(rvk_layernorm_bwd_launch(g, 1, kD, f32(B.x_mid, kD), kD, stat(B.mean2), stat(B.rstd2),
                                       P(params, bp(i, B_N2W)), dx, dx, kD, dxb, G(grads, bp(i, B_N2W)),
                                       G(grads, bp(i, B_N2B)), G(grads, bp(i, B_PROJB)), M, s));
      // ---- attention: x_mid = x_in + proj(attn(qkv(ln1)))
      RVK_TRY(tn(dxb, kD, b16(B.ctx, kD), kD, G(grads, bp(i, B_PROJW)), kD, M, kD, kD, nullptr, &dxb_reader[dxb_cur]));
      RVK_TRY(gemm_plain(dxb, kD, at(wbuf, W.projT[i]), kD, dctx, kD, M, kD, kD, nullptr, s));
      RVK_TRY(rvk_attention_bwd_launch(b16(B.qkv, kQkv), b16(B.ctx, kD), dctx,
                                       reinterpret_cast<float*>(at(workspace, B.lse)) + size_t(b0) * 3 * kTok, dqkv, nb, s));
      RVK_TRY(tn(dqkv, kQkv, b16(B.ln1, kD), kD, G(grads, bp(i, B_QKVW)), kD, M, kQkv, kD, G(grads, bp(i, B_QKVB)),
                 &dqkv_reader[i & 1]));
      RVK_TRY(gemm_plain(dqkv, kQkv, at(wbuf, W.qkvT[i]), kQkv, g, kD, M, kD, kQkv, nullptr, s));
      if (side != nullptr) {
        dxb_cur ^= 1;                         // back to buffer 0: every block (and stage range) starts and ends there
        dxb = dxb_pp[dxb_cur];
        RVK_TRY(wait_reader(dxb_reader[dxb_cur]));
      }
      RVK_TRY(rvk_layernorm_bwd_launch(g, 1, kD, f32(B.x_in, kD), kD, stat(B.mean1), stat(B.rstd1),
                                       P(params, bp(i, B_N1W)), dx, dx, kD, dxb, G(grads, bp(i, B_N1W)),
                                       G(grads, bp(i, B_N1B)), i > 0 ? G(grads, bp(i - 1, B_FC2B)) : nullptr, M, s));
    }
    // ---- patch embedding, class token, position embedding
    if (stage_end == kDepth + 2) {
    dxb = dxb_pp[0];
    RVK_TRY(rvk_gemm_tn_launch(dxb, kD, b16(A.patches, kPatchK), kPatchK, G(grads, P_PATCH_W), kPatchK, M, kD, kPatchK,
                               1.0f, nullptr, s));
    RVK_TRY(rvk_token_grad_reduce_launch(dx, nb, G(grads, P_POS), G(grads, P_CLS), G(grads, P_PATCH_B), s));
    }
    // every weight gradient of this range is complete before the caller (optimizer, all-reduce bucket) continues on `s`
    if (side != nullptr && last_side != nullptr) RVK_CUDA_TRY(cudaStreamWaitEvent(s, last_side, 0));
  }
  return RVK_OK;
}
