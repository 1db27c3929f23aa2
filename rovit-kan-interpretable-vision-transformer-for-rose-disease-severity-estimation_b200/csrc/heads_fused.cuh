// One fused kernel for the whole multi-task tail in inference (included by kan.cu inside its anonymous namespace:
// uses Knots and kan_expand, i.e. exactly the arithmetic of the per-layer KAN kernel):
//
//   features [B,192] -> classification logits [B,4]      Linear(192,128) -> ReLU -> Linear(128,4)       heads.py:17-22
//                    -> cumulative ordinal logits [B,3]  Linear(192,128) -> ReLU -> Linear(128,3)       heads.py:38-43
//                    -> mu, log_var [B,1]                shared Linear(192,128) -> ReLU; two Linear(128,1); clamp +-10   heads.py:91-102
//                    -> KAN severity [B,1]               KAN(192->64) -> ReLU -> KAN(64->16) -> ReLU -> KAN(16->1) -> 3*sigmoid   kan.py:138-149
// (eval mode: Dropout is the identity).  Unfused this tail is 7 GEMM launches + 3 KAN launches + 3 weight packs, each
// on a grid of a few dozen CTAs: 0.3 ms of a 5.5 ms forward at batch 1024.  Here a CTA owns 16 samples end to end: the
// feature tile, the 3 x 128 hidden activations and the expanded KAN activations live in shared memory, all weights
// stream from a prepacked L2-resident buffer (transposed so that consecutive threads read consecutive outputs).
// fp32 throughout: argmax / ordinal decisions stay bit-comparable with the reference given the same features.
#pragma once

constexpr int kHfS = 16;                    // samples per CTA
constexpr int kHfThreads = 256;
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
constexpr int kHfSmemBytes = (kHfD * kHfS + kHfS * kHfU + kHfK0 * 8 * kHfS + kHfS * kHfO0 + kHfS * kHfO1) * 4;

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

__global__ void __launch_bounds__(kHfThreads)
heads_fused_kernel(const float* __restrict__ feat, const float* __restrict__ ws, Knots kn, int batch,
                   float* __restrict__ cls, float* __restrict__ ord, float* __restrict__ mu, float* __restrict__ log_var,
                   float* __restrict__ kan) {
  extern __shared__ __align__(16) float hsm[];
  float* sF = hsm;                          // [192][16]   features, transposed: sF[k][s]
  float* sH = sF + kHfD * kHfS;             // [16][384]   hidden activations of the three heads
  float* sA = sH + kHfS * kHfU;             // [1536][16]  expanded KAN activations sA[i*8+k][s]
  float* sK1 = sA + kHfK0 * 8 * kHfS;       // [64][16]    KAN layer-0 output (transposed)
  float* sK2 = sK1 + kHfS * kHfO0;          // [16][16]    KAN layer-1 output (transposed)
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * kHfS;

  for (int idx = tid; idx < kHfS * kHfD; idx += kHfThreads) {
    const int s = idx / kHfD, k = idx % kHfD;
    sF[k * kHfS + s] = (s0 + s < batch) ? feat[static_cast<size_t>(s0 + s) * kHfD + k] : 0.0f;
  }
  __syncthreads();

  // ---- hidden layers of the three heads: unit u = head*128 + j; thread owns units tid (and tid + 256 < 384)
  for (int u = tid; u < kHfU; u += kHfThreads) {
    float acc[kHfS];
#pragma unroll
    for (int s = 0; s < kHfS; ++s) acc[s] = 0.0f;
    const float* w = ws + kHfOffW1T + u;
#pragma unroll 4
    for (int k = 0; k < kHfD; ++k) {
      const float wk = w[k * kHfU];
      const float4* f4 = reinterpret_cast<const float4*>(sF + k * kHfS);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 f = f4[q];
        acc[4 * q + 0] = fmaf(f.x, wk, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(f.y, wk, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(f.z, wk, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(f.w, wk, acc[4 * q + 3]);
      }
    }
    const float b = ws[kHfOffB1 + u];
#pragma unroll
    for (int s = 0; s < kHfS; ++s) sH[s * kHfU + u] = fmaxf(acc[s] + b, 0.0f);
  }
  // ---- expanded KAN activations of layer 0 (same closed form / tanhf as the per-layer kernel)
  for (int pr = tid; pr < kHfS * kHfK0; pr += kHfThreads) {
    const int s = pr / kHfK0, i = pr % kHfK0;
    float a[kKW], da[kKW], dt;
    kan_expand<false>(sF[i * kHfS + s], kn, a, da, dt);
#pragma unroll
    for (int k = 0; k < kKW; ++k) sA[(i * 8 + k) * kHfS + s] = a[k];
  }
  __syncthreads();

  // ---- second layers of the heads: 9 outputs x 16 samples
  if (tid < 9 * kHfS) {
    const int o = tid % 9, s = tid / 9;
    const int head = (o < 4) ? 0 : (o < 7) ? 1 : 2;
    const float* w = ws + kHfOffW2 + o * kHfH;
    const float* h = sH + s * kHfU + head * kHfH;
    float acc = 0.0f;
#pragma unroll 8
    for (int j = 0; j < kHfH; ++j) acc = fmaf(h[j], w[j], acc);
    acc += ws[kHfOffB2 + o];
    const int sg = s0 + s;
    if (sg < batch) {
      if (o < 4) cls[static_cast<size_t>(sg) * 4 + o] = acc;
      else if (o < 7) ord[static_cast<size_t>(sg) * 3 + (o - 4)] = acc;
      else if (o == 7) mu[sg] = acc;
      else log_var[sg] = fminf(fmaxf(acc, -10.0f), 10.0f);
    }
  }
  // ---- KAN layer 0: thread (o = tid % 64, samples 4*(tid/64) .. +3)
  {
    const int o = tid & 63, sg4 = (tid >> 6) * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* w = ws + kHfOffWp0 + o;
#pragma unroll 8
    for (int kk = 0; kk < kHfK0 * 8; ++kk) {
      const float wk = w[kk * kHfO0];
      const float4 a = *reinterpret_cast<const float4*>(sA + kk * kHfS + sg4);
      acc[0] = fmaf(a.x, wk, acc[0]); acc[1] = fmaf(a.y, wk, acc[1]);
      acc[2] = fmaf(a.z, wk, acc[2]); acc[3] = fmaf(a.w, wk, acc[3]);
    }
    const float b = ws[kHfOffKb0 + o];
#pragma unroll
    for (int q = 0; q < 4; ++q) sK1[o * kHfS + sg4 + q] = fmaxf(acc[q] + b, 0.0f);
  }
  __syncthreads();
  // ---- KAN layer 1 (64 -> 16): expand, then thread (o = tid % 16, sample tid / 16)
  for (int pr = tid; pr < kHfS * kHfO0; pr += kHfThreads) {
    const int s = pr / kHfO0, i = pr % kHfO0;
    float a[kKW], da[kKW], dt;
    kan_expand<false>(sK1[i * kHfS + s], kn, a, da, dt);
#pragma unroll
    for (int k = 0; k < kKW; ++k) sA[(i * 8 + k) * kHfS + s] = a[k];
  }
  __syncthreads();
  {
    const int o = tid & 15, s = tid >> 4;
    float acc = 0.0f;
    const float* w = ws + kHfOffWp1 + o;
#pragma unroll 8
    for (int kk = 0; kk < kHfO0 * 8; ++kk) acc = fmaf(sA[kk * kHfS + s], w[kk * kHfO1], acc);
    sK2[o * kHfS + s] = fmaxf(acc + ws[kHfOffKb1 + o], 0.0f);
  }
  __syncthreads();
  // ---- KAN layer 2 (16 -> 1) and 3*sigmoid: one thread per sample
  if (tid < kHfS) {
    const int s = tid;
    float acc = 0.0f;
    for (int i = 0; i < kHfO1; ++i) {
      float a[kKW], da[kKW], dt;
      kan_expand<false>(sK2[i * kHfS + s], kn, a, da, dt);
#pragma unroll
      for (int k = 0; k < kKW; ++k) acc = fmaf(a[k], ws[kHfOffWp2 + i * 8 + k], acc);
    }
    acc += ws[kHfOffKb2];
    if (s0 + s < batch) kan[s0 + s] = 3.0f / (1.0f + expf(-acc));
  }
}
