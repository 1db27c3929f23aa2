// Attention helpers that are not on the tensor-core hot path (the forward / backward kernels live in attention_tc.cu):
//   * rvk_attention_bwd_launch   dispatch to the tcgen05 backward (the round-1 mma.sync backward that used to live in this
//                                file was only reachable through an environment switch and has been removed)
//   * rvk_attention_probs_launch softmax(q k^T / 8) of one block as an explicit [batch, 3, 197, 197] fp32 tensor, for the
//                                reference's explainability code (attention rollout, explainability/attention_maps.py:46-80):
//                                the fused kernels keep P on chip and never materialise it.
#include "kernels.h"

namespace {

constexpr int kTok = 197, kHeads = 3, kHd = 64, kQkvLd = 576;

// one warp per (image, head, query row); lane l owns keys l, l+32, ... (7 per lane)
__global__ void __launch_bounds__(256) attn_probs_kernel(const __nv_bfloat16* __restrict__ qkv, float* __restrict__ probs, int batch) {
  const int lane = threadIdx.x & 31;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long total = static_cast<long long>(batch) * kHeads * kTok;
  if (row >= total) return;
  const int qi = static_cast<int>(row % kTok);
  const int h = static_cast<int>((row / kTok) % kHeads);
  const long long b = row / (static_cast<long long>(kTok) * kHeads);
  const __nv_bfloat16* base = qkv + b * kTok * kQkvLd;
  const __nv_bfloat16* q = base + static_cast<long long>(qi) * kQkvLd + h * kHd;
  float qv[kHd];
#pragma unroll
  for (int d = 0; d < kHd; ++d) qv[d] = __bfloat162float(q[d]);
  constexpr int kPer = (kTok + 31) / 32;
  float s[kPer];
  float mx = -INFINITY;
#pragma unroll
  for (int r = 0; r < kPer; ++r) {
    const int j = lane + 32 * r;
    float acc = -INFINITY;
    if (j < kTok) {
      const __nv_bfloat16* k = base + static_cast<long long>(j) * kQkvLd + 192 + h * kHd;
      acc = 0.0f;
#pragma unroll
      for (int d = 0; d < kHd; ++d) acc = fmaf(qv[d], __bfloat162float(k[d]), acc);
      acc *= 0.125f;                                   // head_dim^-0.5 (timm Attention.scale)
    }
    s[r] = acc;
    mx = fmaxf(mx, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.0f;
#pragma unroll
  for (int r = 0; r < kPer; ++r) {
    s[r] = (lane + 32 * r < kTok) ? expf(s[r] - mx) : 0.0f;
    sum += s[r];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  float* out = probs + row * kTok;
#pragma unroll
  for (int r = 0; r < kPer; ++r)
    if (lane + 32 * r < kTok) out[lane + 32 * r] = s[r] * inv;
}

}  // namespace

int rvk_attention_bwd_tc_launch(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv, int batch,
                                cudaStream_t stream);      // attention_tc.cu

int rvk_attention_bwd_launch(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                             int batch, cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  return rvk_attention_bwd_tc_launch(qkv, ctx, dctx, lse, dqkv, batch, stream);
}

int rvk_attention_probs_launch(const void* qkv, float* probs, int batch, cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  const long long warps = static_cast<long long>(batch) * kHeads * kTok;
  attn_probs_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(qkv), probs, batch);
  return rvk_launch_check();
}
