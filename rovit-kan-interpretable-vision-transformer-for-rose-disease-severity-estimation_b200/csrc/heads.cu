// Multi-task heads and joint loss (reference models/heads.py, training/losses.py), fp32.
//
// The heads act on the 192-d CLS feature: three Linear(192,128)->ReLU->Dropout->Linear(128,{4,3,1+1})
// MLPs.  At B <= a few thousand rows these are tiny fp32 GEMMs (75 kMAC/sample), far below anything a
// tensor-core tile can use, so they run on one register-tiled SIMT GEMM with the elementwise work
// (bias, ReLU, dropout keep-mask, log-variance clamp, ReLU/dropout backward mask) fused into its
// epilogue.  fp32 keeps the heads bit-comparable with the reference given the same features.
//
// The joint loss (focal CE + ordinal BCE + heteroscedastic NLL + MSE) is one kernel that produces the
// four batch means AND the local gradients w.r.t. every head output in the same pass.
#include "kernels.h"

namespace {

// ------------------------------------------------------------------ register-tiled fp32 GEMM
// C[M,N] = epilogue(op(A)[M,K] * op(B)[K,N]); 64x64 tile, 4x4 per thread, K chunks of 16.
struct SgemmParams {
  const float* A; long long lda; int transA;     // op(A)[m][k] = transA ? A[k*lda+m] : A[m*lda+k]
  const float* B; long long ldb; int transB;     // op(B)[k][n] = transB ? B[n*ldb+k] : B[k*ldb+n]
  float* C; long long ldc;
  int M, N, K;
  const float* bias;          // [N] or null
  int relu;
  float clamp_lo, clamp_hi;   // applied when clamp_lo < clamp_hi
  float drop_p;               // > 0: multiply by Philox keep mask / (1-p)
  unsigned long long seed, offset;
  const float* mask_src;      // non-null: C = acc * mask_scale * (mask_src[m*ld_mask+n] > 0)
  long long ld_mask;
  float mask_scale;
  int accumulate;             // C += result (atomicAdd when split over K)
  int k_per_split;
};

// RM rows per thread: 64-row tiles (RM = 4) for large M, 16-row tiles (RM = 1) so that a 1024-sample inference batch
// still spreads over >= 128 CTAs instead of 32
template <int RM>
__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmParams p) {
  constexpr int TM = 16 * RM;
  __shared__ __align__(16) float sA[16][TM];
  __shared__ __align__(16) float sB[16][64];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * 64;
  const int k_begin = blockIdx.z * p.k_per_split, k_end = min(p.K, k_begin + p.k_per_split);
  float acc[RM][4];
#pragma unroll
  for (int a = 0; a < RM; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    for (int idx = tid; idx < 16 * TM; idx += 256) {
      int kl, ml;
      if (p.transA) { kl = idx / TM; ml = idx % TM; } else { ml = idx >> 4; kl = idx & 15; }
      const int m = m0 + ml, k = k0 + kl;
      float v = 0.0f;
      if (m < p.M && k < k_end) v = p.transA ? p.A[static_cast<long long>(k) * p.lda + m] : p.A[static_cast<long long>(m) * p.lda + k];
      sA[kl][ml] = v;
    }
    for (int idx = tid; idx < 16 * 64; idx += 256) {
      int kl, nl;
      if (p.transB) { nl = idx >> 4; kl = idx & 15; } else { kl = idx >> 6; nl = idx & 63; }
      const int n = n0 + nl, k = k0 + kl;
      float v = 0.0f;
      if (n < p.N && k < k_end) v = p.transB ? p.B[static_cast<long long>(n) * p.ldb + k] : p.B[static_cast<long long>(k) * p.ldb + n];
      sB[kl][nl] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a4[RM];
#pragma unroll
      for (int a = 0; a < RM; ++a) a4[a] = sA[kk][ty * RM + a];
      const float4 bv = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int a = 0; a < RM; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int a = 0; a < RM; ++a) {
    const int m = m0 + ty * RM + a;
    if (m >= p.M) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int n = n0 + tx * 4 + b;
      if (n >= p.N) continue;
      float v = acc[a][b];
      float* dst = p.C + static_cast<long long>(m) * p.ldc + n;
      if (split) { atomicAdd(dst, v); continue; }
      if (p.bias != nullptr) v += p.bias[n];
      if (p.relu) v = fmaxf(v, 0.0f);
      if (p.clamp_lo < p.clamp_hi) v = fminf(fmaxf(v, p.clamp_lo), p.clamp_hi);
      if (p.drop_p > 0.0f) {
        const float u = uniform01(p.seed, p.offset, static_cast<unsigned long long>(m) * p.N + n);
        v = (u >= p.drop_p) ? v * (1.0f / (1.0f - p.drop_p)) : 0.0f;
      }
      if (p.mask_src != nullptr) v = (p.mask_src[static_cast<long long>(m) * p.ld_mask + n] > 0.0f) ? v * p.mask_scale : 0.0f;
      if (p.accumulate) v += *dst;
      *dst = v;
    }
  }
}

// out[c] += sum_r src[r*ld + c] for narrow matrices (bias gradients of the heads; any column count)
__global__ void colsum_small_kernel(const float* __restrict__ src, long long ld, int rows, int cols,
                                    float* __restrict__ out) {
  const int c = blockIdx.x;
  float acc = 0.0f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) acc += src[static_cast<long long>(r) * ld + c];
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += part[i];
    out[c] += t;
  }
}

// gradient through the fused epilogue, recovered from the forward OUTPUT y:
//   relu/dropout: pass (scaled by 1/(1-p)) where y > 0;  clamp: pass where lo < y < hi
__global__ void epilogue_grad_kernel(const float* __restrict__ y, const float* __restrict__ g, float* __restrict__ out,
                                     int relu, float scale, float lo, float hi, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = g[i];
  if (relu) v = (y[i] > 0.0f) ? v * scale : 0.0f;
  if (lo < hi) v = (y[i] > lo && y[i] < hi) ? v : 0.0f;
  out[i] = v;
}

// ------------------------------------------------------------------ joint loss
struct LossParams {
  const float* cls_logits; int num_classes;    // [B, C]
  const float* ord_logits;                     // [B, C-1] or null
  const float* mu; const float* log_var;       // [B] or null
  const float* kan;                            // [B] or null
  const long long* class_t; const float* sev_t;   // severities are fp32: the reference casts them with .float() (losses.py:92,112)
  const float* alpha;                          // [C] or null
  float gamma;
  int batch;
  float* sums;                                 // [4] running sums of per-sample terms (already / denominators)
  float* d_cls; float* d_ord; float* d_mu; float* d_lv; float* d_kan;   // local gradients (may be null)
};

// The four per-sample loss terms of sample b (each already divided by its denominator) and, where the d_* pointers are set,
// their local gradients.  A plain function of (p, b) so that tests/test_kernel_constants.py can compile it for the host and
// check it against the reference-generated vectors without a GPU.
__device__ __forceinline__ void joint_loss_sample(const LossParams& p, int b, float& l_cls, float& l_ord, float& l_unc,
                                                  float& l_kan) {
  {
    const float inv_b = 1.0f / static_cast<float>(p.batch);
    const int C = p.num_classes;
    {   // focal cross-entropy (losses.py:15-38)
      const float* z = p.cls_logits + static_cast<size_t>(b) * C;
      const long long t_raw = p.class_t[b];
      // an out-of-range label (torch raises a device-side assert in gather, losses.py:27): never index with it; the
      // sample contributes NaN to the loss (loud) and a zero gradient
      const bool t_ok = t_raw >= 0 && t_raw < C;
      const int t = t_ok ? static_cast<int>(t_raw) : 0;
      float mx = z[0];
      for (int j = 1; j < C; ++j) mx = fmaxf(mx, z[j]);
      float se = 0.0f;
      for (int j = 0; j < C; ++j) se += expf(z[j] - mx);
      const float lse = mx + logf(se);
      const float ce = lse - z[t];
      const float pt = expf(z[t] - lse);
      const float a = (p.alpha != nullptr) ? p.alpha[t] : 1.0f;
      const float om = fmaxf(1.0f - pt, 0.0f);        // rounding can push pt above 1: powf(negative, non-integer) is NaN
      const float f = (p.gamma == 2.0f) ? om * om : powf(om, p.gamma);
      l_cls = t_ok ? a * f * ce * inv_b : __int_as_float(0x7fc00000);
      if (p.d_cls != nullptr) {
        const float fm1 = (p.gamma == 2.0f) ? om : ((om > 0.0f) ? powf(om, p.gamma - 1.0f) : 0.0f);
        const float coef = t_ok ? a * (p.gamma * fm1 * pt * ce + f) * inv_b : 0.0f;
        for (int j = 0; j < C; ++j) {
          const float pj = expf(z[j] - lse);
          p.d_cls[static_cast<size_t>(b) * C + j] = coef * (pj - (j == t ? 1.0f : 0.0f));
        }
      }
    }
    const float y = p.sev_t[b];
    if (p.ord_logits != nullptr) {   // BCE-with-logits on [y > k] (losses.py:48-72)
      const int K = C - 1;
      const float inv = inv_b / static_cast<float>(K);
      for (int k = 0; k < K; ++k) {
        const float z = p.ord_logits[static_cast<size_t>(b) * K + k];
        const float t = (y > static_cast<float>(k)) ? 1.0f : 0.0f;
        l_ord += (fmaxf(z, 0.0f) - z * t + log1pf(expf(-fabsf(z)))) * inv;
        if (p.d_ord != nullptr) p.d_ord[static_cast<size_t>(b) * K + k] = (1.0f / (1.0f + expf(-z)) - t) * inv;
      }
    }
    if (p.mu != nullptr && p.log_var != nullptr) {   // heteroscedastic NLL (losses.py:80-101)
      const float d = y - p.mu[b], lv = p.log_var[b];
      const float prec = expf(-lv);
      l_unc = 0.5f * (d * d * prec + lv) * inv_b;
      if (p.d_mu != nullptr) p.d_mu[b] = -d * prec * inv_b;
      if (p.d_lv != nullptr) p.d_lv[b] = 0.5f * (1.0f - d * d * prec) * inv_b;
    }
    if (p.kan != nullptr) {   // MSE (losses.py:109-114)
      const float d = p.kan[b] - y;
      l_kan = d * d * inv_b;
      if (p.d_kan != nullptr) p.d_kan[b] = 2.0f * d * inv_b;
    }
  }
}

__global__ void __launch_bounds__(256) joint_loss_kernel(const LossParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float l_cls = 0.0f, l_ord = 0.0f, l_unc = 0.0f, l_kan = 0.0f;
  if (b < p.batch) joint_loss_sample(p, b, l_cls, l_ord, l_unc, l_kan);
  __shared__ float part[4][8];
  const float v[4] = {warp_sum(l_cls), warp_sum(l_ord), warp_sum(l_unc), warp_sum(l_kan)};
  if ((threadIdx.x & 31) == 0)
    for (int i = 0; i < 4; ++i) part[i][threadIdx.x >> 5] = v[i];
  __syncthreads();
  if (threadIdx.x < 4) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += part[threadIdx.x][i];
    atomicAdd(&p.sums[threadIdx.x], t);
  }
}

// RoViTKAN.predict epilogue (rovit_kan.py:126-161 with heads.py:45-77): softmax + argmax of the class logits, the ordinal
// decode c = sigmoid(logits), p0 = c0, pk = ck - ck-1, pK-1 = 1 - cK-2 and its expected value, std = exp(log_var / 2) --
// one launch instead of ~12 elementwise / reduction kernels (and the reference runs the ordinal head three times for it)
__global__ void predict_decode_kernel(const float* __restrict__ cls, int C, const float* __restrict__ ordl,
                                      const float* __restrict__ log_var, int batch, long long* __restrict__ cls_idx,
                                      float* __restrict__ probs, float* __restrict__ ord_probs, float* __restrict__ ord_sev,
                                      float* __restrict__ unc_std) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  {
    const float* z = cls + static_cast<size_t>(b) * C;
    float mx = z[0];
    for (int j = 1; j < C; ++j) mx = fmaxf(mx, z[j]);
    float se = 0.0f;
    for (int j = 0; j < C; ++j) se += expf(z[j] - mx);
    const float inv = 1.0f / se;
    int best = 0;
    float bp = -1.0f;
    for (int j = 0; j < C; ++j) {
      const float pj = expf(z[j] - mx) * inv;
      probs[static_cast<size_t>(b) * C + j] = pj;
      if (pj > bp) { bp = pj; best = j; }            // first maximum, like torch.argmax
    }
    cls_idx[b] = best;
  }
  if (ordl != nullptr) {
    const int K = C - 1;
    const float* z = ordl + static_cast<size_t>(b) * K;
    float prev = 0.0f, sev = 0.0f;
    for (int k = 0; k < K; ++k) {
      const float c = 1.0f / (1.0f + expf(-z[k]));
      const float pk = (k == 0) ? c : c - prev;
      ord_probs[static_cast<size_t>(b) * C + k] = pk;
      sev = fmaf(static_cast<float>(k), pk, sev);
      prev = c;
    }
    const float pl = 1.0f - prev;
    ord_probs[static_cast<size_t>(b) * C + K] = pl;
    ord_sev[b] = fmaf(static_cast<float>(K), pl, sev);
  }
  if (log_var != nullptr) unc_std[b] = expf(0.5f * log_var[b]);
}

// out = {cls, ord, unc, kan, cls + l_ord*ord + m_unc*unc + n_kan*kan}
__global__ void loss_finalize_kernel(const float* sums, float l_ord, float m_unc, float n_kan, float* out) {
  if (threadIdx.x == 0) {
    const float c = sums[0], o = sums[1], u = sums[2], k = sums[3];
    out[0] = c; out[1] = o; out[2] = u; out[3] = k;
    out[4] = c + l_ord * o + m_unc * u + n_kan * k;
  }
}

// dst[i] = src[i] * (g[term] + w_total * g[4])
__global__ void loss_scale_grad_kernel(const float* __restrict__ src, const float* __restrict__ g, int term,
                                       float w_total, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * (g[term] + w_total * g[4]);
}

}  // namespace

int rvk_sgemm_launch(const SgemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return RVK_OK;
  if (a.K <= 0 || a.A == nullptr || a.B == nullptr || a.C == nullptr) return RVK_ERR_BAD_ARG;
  SgemmParams p;
  p.A = a.A; p.lda = a.lda; p.transA = a.transA;
  p.B = a.B; p.ldb = a.ldb; p.transB = a.transB;
  p.C = a.C; p.ldc = a.ldc;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias; p.relu = a.relu;
  p.clamp_lo = a.clamp_lo; p.clamp_hi = a.clamp_hi;
  p.drop_p = a.drop_p; p.seed = a.seed; p.offset = a.offset;
  p.mask_src = a.mask_src; p.ld_mask = a.ld_mask; p.mask_scale = a.mask_scale;
  p.accumulate = a.accumulate;
  int splits = 1;
  if (a.split_k) {
    if (!a.accumulate || a.bias != nullptr || a.relu || a.drop_p > 0.0f || a.mask_src != nullptr) return RVK_ERR_BAD_ARG;
    const int tiles = ((a.M + 63) / 64) * ((a.N + 63) / 64);
    splits = (kNumSMsB200 + tiles - 1) / tiles;
    const int max_splits = (a.K + 63) / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  p.k_per_split = ((a.K + splits - 1) / splits + 15) / 16 * 16;
  splits = (a.K + p.k_per_split - 1) / p.k_per_split;
  if (a.M <= 4096) {
    dim3 grid((a.M + 15) / 16, (a.N + 63) / 64, splits);
    sgemm_kernel<1><<<grid, 256, 0, stream>>>(p);
  } else {
    dim3 grid((a.M + 63) / 64, (a.N + 63) / 64, splits);
    sgemm_kernel<4><<<grid, 256, 0, stream>>>(p);
  }
  return rvk_launch_check();
}

int rvk_colsum_small_launch(const float* src, int64_t ld, int rows, int cols, float* out, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return RVK_OK;
  colsum_small_kernel<<<cols, 256, 0, stream>>>(src, ld, rows, cols, out);
  return rvk_launch_check();
}

int rvk_epilogue_grad_launch(const float* y, const float* g, float* out, int relu, float scale, float lo, float hi,
                             int n, cudaStream_t stream) {
  if (n <= 0) return RVK_OK;
  epilogue_grad_kernel<<<(n + 255) / 256, 256, 0, stream>>>(y, g, out, relu, scale, lo, hi, n);
  return rvk_launch_check();
}

int rvk_joint_loss_launch(const JointLossArgs& a, cudaStream_t stream) {
  if (a.batch <= 0 || a.cls_logits == nullptr || a.class_t == nullptr || a.sev_t == nullptr || a.out == nullptr ||
      a.sums_ws == nullptr)
    return RVK_ERR_BAD_ARG;
  if (a.num_classes < 2 || a.num_classes > 64) return RVK_ERR_UNSUPPORTED_SHAPE;
  RVK_CUDA_TRY(cudaMemsetAsync(a.sums_ws, 0, 4 * sizeof(float), stream));
  LossParams p;
  p.cls_logits = a.cls_logits; p.num_classes = a.num_classes;
  p.ord_logits = a.ord_logits; p.mu = a.mu; p.log_var = a.log_var; p.kan = a.kan;
  p.class_t = reinterpret_cast<const long long*>(a.class_t);
  p.sev_t = static_cast<const float*>(a.sev_t);
  p.alpha = a.alpha; p.gamma = a.gamma; p.batch = a.batch; p.sums = a.sums_ws;
  p.d_cls = a.d_cls; p.d_ord = a.d_ord; p.d_mu = a.d_mu; p.d_lv = a.d_lv; p.d_kan = a.d_kan;
  joint_loss_kernel<<<(a.batch + 255) / 256, 256, 0, stream>>>(p);
  RVK_TRY(rvk_launch_check());
  loss_finalize_kernel<<<1, 32, 0, stream>>>(a.sums_ws, a.ord_logits ? a.lambda_ord : 0.0f,
                                             (a.mu && a.log_var) ? a.mu_unc : 0.0f, a.kan ? a.nu_kan : 0.0f, a.out);
  return rvk_launch_check();
}

int rvk_loss_scale_grad_launch(const float* local, const float* upstream5, int term, float w_total, float* dst, int n,
                               cudaStream_t stream) {
  if (n <= 0) return RVK_OK;
  loss_scale_grad_kernel<<<(n + 255) / 256, 256, 0, stream>>>(local, upstream5, term, w_total, dst, n);
  return rvk_launch_check();
}

int rvk_predict_decode_launch(const float* cls, int num_classes, const float* ordl, const float* log_var, int batch,
                              long long* cls_idx, float* probs, float* ord_probs, float* ord_sev, float* unc_std,
                              cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  if (cls == nullptr || cls_idx == nullptr || probs == nullptr || num_classes < 2 || num_classes > 64 ||
      (ordl != nullptr && (ord_probs == nullptr || ord_sev == nullptr)) || (log_var != nullptr && unc_std == nullptr))
    return RVK_ERR_BAD_ARG;
  predict_decode_kernel<<<(batch + 255) / 256, 256, 0, stream>>>(cls, num_classes, ordl, log_var, batch, cls_idx, probs, ord_probs,
                                                                 ord_sev, unc_std);
  return rvk_launch_check();
}
