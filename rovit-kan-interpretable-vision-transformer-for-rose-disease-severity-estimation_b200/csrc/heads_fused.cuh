// One fused kernel for the whole multi-task tail in inference (included by kan.cu inside its anonymous namespace:
// uses Knots and kan_expand, i.e. exactly the arithmetic of the per-layer KAN kernel):
//
//   features [B,192] -> classification logits [B,4]      Linear(192,128) -> ReLU -> Linear(128,4)       heads.py:17-22
//                    -> cumulative ordinal logits [B,3]  Linear(192,128) -> ReLU -> Linear(128,3)       heads.py:38-43
//                    -> mu, log_var [B,1]                shared Linear(192,128) -> ReLU; two Linear(128,1); clamp +-10   heads.py:91-102
//                    -> KAN severity [B,1]               KAN(192->64) -> ReLU -> KAN(64->16) -> ReLU -> KAN(16->1) -> 3*sigmoid   kan.py:138-149
// (eval mode: Dropout is the identity).  Unfused this tail is 7 GEMM launches + 3 KAN launches + 3 weight packs, each
// on a grid of a few dozen CTAs: 0.3 ms of a 5.5 ms forward at batch 1024.  Here a CTA of 512 threads owns 8 samples end
// to end (128 CTAs at batch 1024): the feature tile, the 3 x 128 hidden activations and the expanded KAN activations live
// in shared memory; the two big weight matrices (fc1 of the heads, KAN layer 0) stream from a prepacked L2-resident
// buffer (transposed so that consecutive threads read consecutive outputs) with the contraction dimension SPLIT across
// warps (8 resp. 4 ways, partial sums reduced through shared memory) so that the chain of dependent L2 round trips per
// thread is 24 + 6 batches instead of 192 + 96; the small late-layer weights are prefetched into shared memory with
// cp.async while the first phase runs.  fp32 throughout: argmax / ordinal decisions stay bit-comparable with the
// reference given the same features.
#pragma once

constexpr int kHfS = 8;                     // samples per CTA
constexpr int kHfThreads = 512;
constexpr int kHfD = 192, kHfH = 128, kHfU = 3 * kHfH;
constexpr int kHfK0 = 192, kHfO0 = 64, kHfO1 = 16;          // KAN stack 192 -> 64 -> 16 -> 1
// prepacked weight buffer (floats)
constexpr int kHfOffW1T = 0;                                  // [192][384]  fc1 of the three heads, transposed
constexpr int kHfOffB1 = kHfOffW1T + kHfD * kHfU;             // [384]
constexpr int kHfOffW2 = kHfOffB1 + kHfU;                     // [9][128]  cls0..3, ord0..2, mu, log_var
constexpr int kHfOffB2 = kHfOffW2 + 9 * kHfH;                 // [16] (9 used)
constexpr int kHfOffWp0 = kHfOffB2 + 16;                      // [192*8][64]
constexpr int kHfOffKb0 = kHfOffWp0 + kHfK0 * 8 * kHfO0;      // [64]
constexpr int kHfOffWp1 = kHfOffKb0 + kHfO0;                  // [64*8][16]
constexpr int kHfOffKb1 = kHfOffWp1 + kHfO0 * 8 * kHfO1;      // [16]
constexpr int kHfOffWp2 = kHfOffKb1 + kHfO1;                  // [16*8]
constexpr int kHfOffKb2 = kHfOffWp2 + kHfO1 * 8;              // [1] (+3 pad)
constexpr int kHfWsFloats = kHfOffKb2 + 4;
// shared memory (floats)
constexpr int kHfAStride = 72;               // expanded activations [input i][basis j][sample s], i-stride padded (bank spread)
constexpr int kHfHStride = kHfU + 4;         // hidden activations [sample][384], padded
constexpr int kHfG0 = 8;                     // K split of KAN layer 0 (1536 / 8 = 192 per thread)
constexpr int kHfG1 = 4;                     // K split of the heads' fc1 (192 / 4 = 48 per thread) and of KAN layer 1
constexpr int kHfTailFloats = kHfWsFloats - kHfOffWp1;        // Wp1, kb1, Wp2, kb2
constexpr int kHfW2Floats = kHfOffWp0 - kHfOffW2;             // fc2 weights + biases
constexpr int kHfSmF = 0;                                     // [192][8] features, transposed
constexpr int kHfSmA = kHfSmF + kHfD * kHfS;                  // [192][72] expanded KAN activations (layer 1 reuses [64][72])
constexpr int kHfSmP0 = kHfSmA + kHfK0 * kHfAStride;          // [8][64][8] KAN layer-0 partial sums (layer 1 reuses [4][128])
constexpr int kHfSmP1 = kHfSmP0 + kHfG0 * kHfO0 * kHfS;       // [4][384][8] fc1 partial sums
constexpr int kHfSmH = kHfSmP1 + kHfG1 * kHfU * kHfS;         // [8][388] hidden activations
constexpr int kHfSmW2 = kHfSmH + kHfS * kHfHStride;           // fc2 weights + biases
constexpr int kHfSmTail = kHfSmW2 + kHfW2Floats;              // Wp1, kb1, Wp2, kb2
constexpr int kHfSmemBytes = (kHfSmTail + kHfTailFloats) * 4;
static_assert(kHfOffW2 % 4 == 0 && kHfOffWp1 % 4 == 0 && kHfW2Floats % 4 == 0 && kHfTailFloats % 4 == 0, "16-byte cp.async");
static_assert(kHfSmW2 % 4 == 0 && kHfSmTail % 4 == 0 && kHfSmH % 4 == 0 && kHfSmA % 4 == 0, "float4 alignment");

struct HeadsFusedParams {          // device pointers, reference parameter layouts
  const float* fc1_w[3]; const float* fc1_b[3];      // cls, ord, unc: [128,192], [128]
  const float* fc2_w[4]; const float* fc2_b[4];      // cls [4,128], ord [3,128], mu [1,128], log_var [1,128]
  const float* spline[3]; const float* lin_w[3]; const float* lin_b[3];   // KAN layers: [in,out,7], [out,in], [out]
};

__global__ void heads_fused_pack_kernel(const HeadsFusedParams p, float* __restrict__ ws) {
  const int n = kHfWsFloats;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    float v = 0.0f;
    if (idx < kHfOffB1) {
      const int k = idx / kHfU, u = idx % kHfU;
      v = p.fc1_w[u / kHfH][(u % kHfH) * kHfD + k];
    } else if (idx < kHfOffW2) {
      const int u = idx - kHfOffB1;
      v = p.fc1_b[u / kHfH][u % kHfH];
    } else if (idx < kHfOffB2) {
      const int o = (idx - kHfOffW2) / kHfH, j = (idx - kHfOffW2) % kHfH;
      v = (o < 4) ? p.fc2_w[0][o * kHfH + j] : (o < 7) ? p.fc2_w[1][(o - 4) * kHfH + j] : p.fc2_w[o - 5][j];
    } else if (idx < kHfOffWp0) {
      const int o = idx - kHfOffB2;
      if (o < 4) v = p.fc2_b[0][o]; else if (o < 7) v = p.fc2_b[1][o - 4]; else if (o < 9) v = p.fc2_b[o - 5][0];
    } else if (idx < kHfOffKb0) {
      const int kk = (idx - kHfOffWp0) / kHfO0, o = (idx - kHfOffWp0) % kHfO0, i = kk >> 3, k = kk & 7;
      v = (k < 7) ? p.spline[0][(i * kHfO0 + o) * 7 + k] : p.lin_w[0][o * kHfK0 + i];
    } else if (idx < kHfOffWp1) {
      v = p.lin_b[0][idx - kHfOffKb0];
    } else if (idx < kHfOffKb1) {
      const int kk = (idx - kHfOffWp1) / kHfO1, o = (idx - kHfOffWp1) % kHfO1, i = kk >> 3, k = kk & 7;
      v = (k < 7) ? p.spline[1][(i * kHfO1 + o) * 7 + k] : p.lin_w[1][o * kHfO0 + i];
    } else if (idx < kHfOffWp2) {
      v = p.lin_b[1][idx - kHfOffKb1];
    } else if (idx < kHfOffKb2) {
      const int kk = idx - kHfOffWp2, i = kk >> 3, k = kk & 7;
      v = (k < 7) ? p.spline[2][i * 7 + k] : p.lin_w[2][i];
    } else if (idx == kHfOffKb2) {
      v = p.lin_b[2][0];
    }
    ws[idx] = v;
  }
}

__device__ __forceinline__ void hf_cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc) : "memory");
}

// TRAIN: the same kernel as the forward of the training step -- dropout on the three hidden layers (Philox keep mask, scaled
// by 1 / (1 - p), heads.py:20,41,94) and the activations the backward kernel (heads_train.cuh) needs are written out: the
// hidden layers after ReLU + dropout [B,384], the KAN hidden activations after ReLU [B,64] and [B,16].
struct HeadsTrainSave {
  float* h;  float* a1;  float* a2;
  float drop_p;  unsigned long long seed, offset;
};

template <bool TRAIN>
__global__ void __launch_bounds__(kHfThreads, 1)
heads_fused_kernel(const float* __restrict__ feat, const float* __restrict__ ws, Knots kn, int batch,
                   float* __restrict__ cls, float* __restrict__ ord, float* __restrict__ mu, float* __restrict__ log_var,
                   float* __restrict__ kan, const HeadsTrainSave sv) {
  extern __shared__ __align__(16) float hsm[];
  float* sF = hsm + kHfSmF;
  float* sA = hsm + kHfSmA;
  float* sP0 = hsm + kHfSmP0;
  float* sP1 = hsm + kHfSmP1;
  float* sH = hsm + kHfSmH;
  float* sW2 = hsm + kHfSmW2;
  float* sTail = hsm + kHfSmTail;
  const float* sWp1 = sTail;                                   // [512][16]
  const float* sKb1 = sTail + (kHfOffKb1 - kHfOffWp1);
  const float* sWp2 = sTail + (kHfOffWp2 - kHfOffWp1);         // [16][8]
  const float* sKb2 = sTail + (kHfOffKb2 - kHfOffWp1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s0 = blockIdx.x * kHfS;

  // late-layer weights -> shared memory, asynchronously (needed after the second barrier at the earliest)
  for (int i = tid; i < kHfW2Floats / 4; i += kHfThreads) hf_cp_async16(sW2 + 4 * i, ws + kHfOffW2 + 4 * i);
  for (int i = tid; i < kHfTailFloats / 4; i += kHfThreads) hf_cp_async16(sTail + 4 * i, ws + kHfOffWp1 + 4 * i);
  asm volatile("cp.async.commit_group;" ::: "memory");

  for (int idx = tid; idx < kHfS * kHfD; idx += kHfThreads) {
    const int s = idx / kHfD, k = idx % kHfD;
    sF[k * kHfS + s] = (s0 + s < batch) ? feat[static_cast<size_t>(s0 + s) * kHfD + k] : 0.0f;
  }
  __syncthreads();
  // ---- expanded KAN activations of layer 0 (same closed form / tanhf as the per-layer kernel)
  for (int idx = tid; idx < kHfS * kHfK0; idx += kHfThreads) {
    const int s = idx & (kHfS - 1), i = idx >> 3;
    float a[kKW], da[kKW], dt;
    kan_expand<false>(sF[i * kHfS + s], kn, a, da, dt);
#pragma unroll
    for (int k = 0; k < kKW; ++k) sA[i * kHfAStride + k * kHfS + s] = a[k];
  }
  // ---- fc1 of the three heads: thread = (unit j of every head, quarter kq of the 192 inputs)
  {
    const int j = tid & (kHfH - 1), kq = tid >> 7;
    float acc[3][kHfS];
#pragma unroll
    for (int h = 0; h < 3; ++h)
#pragma unroll
      for (int s = 0; s < kHfS; ++s) acc[h][s] = 0.0f;
    const float* w = ws + kHfOffW1T + static_cast<size_t>(kq * (kHfD / kHfG1)) * kHfU + j;
    const float* f = sF + kq * (kHfD / kHfG1) * kHfS;
#pragma unroll 8
    for (int k = 0; k < kHfD / kHfG1; ++k) {
      const float w0 = __ldg(w + k * kHfU), w1 = __ldg(w + k * kHfU + kHfH), w2 = __ldg(w + k * kHfU + 2 * kHfH);
      const float4 fa = *reinterpret_cast<const float4*>(f + k * kHfS);
      const float4 fb = *reinterpret_cast<const float4*>(f + k * kHfS + 4);
      const float fv[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll
      for (int s = 0; s < kHfS; ++s) {
        acc[0][s] = fmaf(fv[s], w0, acc[0][s]);
        acc[1][s] = fmaf(fv[s], w1, acc[1][s]);
        acc[2][s] = fmaf(fv[s], w2, acc[2][s]);
      }
    }
#pragma unroll
    for (int h = 0; h < 3; ++h) {
      float* dst = sP1 + (kq * kHfU + h * kHfH + j) * kHfS;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[h][4], acc[h][5], acc[h][6], acc[h][7]);
    }
  }
  __syncthreads();          // sA complete
  // ---- KAN layer 0: thread = (output o, eighth g of the 1536 expanded inputs)
  {
    const int o = tid & (kHfO0 - 1), g = tid >> 6;
    constexpr int kIPer = kHfK0 / kHfG0;      // 24 inputs = 192 expanded rows
    float acc[kHfS];
#pragma unroll
    for (int s = 0; s < kHfS; ++s) acc[s] = 0.0f;
    const float* w = ws + kHfOffWp0 + static_cast<size_t>(g * kIPer * 8) * kHfO0 + o;
    const float* a = sA + g * kIPer * kHfAStride;
#pragma unroll 2
    for (int i = 0; i < kIPer; ++i) {
      float wv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) wv[k] = __ldg(w + (i * 8 + k) * kHfO0);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 aa = *reinterpret_cast<const float4*>(a + i * kHfAStride + k * kHfS);
        const float4 ab = *reinterpret_cast<const float4*>(a + i * kHfAStride + k * kHfS + 4);
        acc[0] = fmaf(aa.x, wv[k], acc[0]); acc[1] = fmaf(aa.y, wv[k], acc[1]);
        acc[2] = fmaf(aa.z, wv[k], acc[2]); acc[3] = fmaf(aa.w, wv[k], acc[3]);
        acc[4] = fmaf(ab.x, wv[k], acc[4]); acc[5] = fmaf(ab.y, wv[k], acc[5]);
        acc[6] = fmaf(ab.z, wv[k], acc[6]); acc[7] = fmaf(ab.w, wv[k], acc[7]);
      }
    }
    float* dst = sP0 + (g * kHfO0 + o) * kHfS;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();          // partial sums + prefetched weights visible; every thread is done reading sA
  // ---- reduce: hidden activations of the heads (bias, ReLU) ...
  for (int idx = tid; idx < kHfU * kHfS; idx += kHfThreads) {
    const int s = idx & (kHfS - 1), u = idx >> 3;
    float v = __ldg(ws + kHfOffB1 + u);
#pragma unroll
    for (int q = 0; q < kHfG1; ++q) v += sP1[(q * kHfU + u) * kHfS + s];
    v = fmaxf(v, 0.0f);
    if (TRAIN) {
      if (sv.drop_p > 0.0f) {
        const float r = uniform01(sv.seed, sv.offset, static_cast<unsigned long long>(s0 + s) * kHfU + u);
        v = (r >= sv.drop_p) ? v * (1.0f / (1.0f - sv.drop_p)) : 0.0f;
      }
      if (s0 + s < batch) sv.h[static_cast<size_t>(s0 + s) * kHfU + u] = v;
    }
    sH[s * kHfHStride + u] = v;
  }
  // ---- ... and KAN layer-0 outputs (bias, ReLU), expanded straight into the layer-1 activations
  {
    const int s = tid & (kHfS - 1), o = tid >> 3;
    float v = __ldg(ws + kHfOffKb0 + o);
#pragma unroll
    for (int g = 0; g < kHfG0; ++g) v += sP0[(g * kHfO0 + o) * kHfS + s];
    float a[kKW], da[kKW], dt;
    v = fmaxf(v, 0.0f);
    if (TRAIN && s0 + s < batch) sv.a1[static_cast<size_t>(s0 + s) * kHfO0 + o] = v;
    kan_expand<false>(v, kn, a, da, dt);
#pragma unroll
    for (int k = 0; k < kKW; ++k) sA[o * kHfAStride + k * kHfS + s] = a[k];
  }
  __syncthreads();
  // ---- KAN layer 1 (64 -> 16): thread = (output o, sample s, quarter q of the 512 expanded inputs)
  {
    const int o = tid & (kHfO1 - 1), s = (tid >> 4) & (kHfS - 1), q = tid >> 7;
    float acc = 0.0f;
#pragma unroll 4
    for (int i = q * 16; i < q * 16 + 16; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(sA[i * kHfAStride + k * kHfS + s], sWp1[(i * 8 + k) * kHfO1 + o], acc);
    sP0[q * 128 + s * kHfO1 + o] = acc;       // sP0 was consumed before the previous barrier
  }
  __syncthreads();
  if (warp < 4) {
    // ---- KAN layer-1 epilogue, layer 2 (16 -> 1) and 3*sigmoid: thread = (sample s, layer-2 input i), 16-lane reduction
    const int i = tid & (kHfO1 - 1), s = tid >> 4;
    float v = sKb1[i];
#pragma unroll
    for (int q = 0; q < kHfG1; ++q) v += sP0[q * 128 + s * kHfO1 + i];
    float a[kKW], da[kKW], dt;
    v = fmaxf(v, 0.0f);
    if (TRAIN && s0 + s < batch) sv.a2[static_cast<size_t>(s0 + s) * kHfO1 + i] = v;
    kan_expand<false>(v, kn, a, da, dt);
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < kKW; ++k) acc = fmaf(a[k], sWp2[i * 8 + k], acc);
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (i == 0 && s0 + s < batch) kan[s0 + s] = 3.0f / (1.0f + expf(-(acc + sKb2[0])));
  } else if (warp < 13) {
    // ---- second layers of the heads: warp = one of the 9 outputs, lanes split the 128 hidden units
    const int o = warp - 4;
    const int head = (o < 4) ? 0 : (o < 7) ? 1 : 2;
    const float4 wv = *reinterpret_cast<const float4*>(sW2 + o * kHfH + lane * 4);
    const float bias = sW2[9 * kHfH + o];
#pragma unroll
    for (int s = 0; s < kHfS; ++s) {
      const float4 hv = *reinterpret_cast<const float4*>(sH + s * kHfHStride + head * kHfH + lane * 4);
      float acc = fmaf(hv.x, wv.x, fmaf(hv.y, wv.y, fmaf(hv.z, wv.z, hv.w * wv.w)));
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
      acc += bias;
      const int sg = s0 + s;
      if (lane == 0 && sg < batch) {
        if (o < 4) cls[static_cast<size_t>(sg) * 4 + o] = acc;
        else if (o < 7) ord[static_cast<size_t>(sg) * 3 + (o - 4)] = acc;
        else if (o == 7) mu[sg] = acc;
        else log_var[sg] = fminf(fmaxf(acc, -10.0f), 10.0f);
      }
    }
  }
}
