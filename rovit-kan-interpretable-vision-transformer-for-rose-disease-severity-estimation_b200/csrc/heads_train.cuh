// Backward of the fused multi-task tail (forward: heads_fused_kernel<true> in heads_fused.cuh), included by kan.cu inside its
// anonymous namespace.  One kernel takes the gradients of the five head outputs (what the joint-loss kernel produced,
// scaled by autograd) and returns d(loss)/d(features) plus the gradients of all 23 head / KAN parameters:
//
//   classification / ordinal / uncertainty heads (heads.py:17-22, 38-43, 91-102): Linear(128,{4,3,1,1}) backward, the
//   ReLU + dropout mask recovered from the saved hidden activations (zero <=> masked), Linear(192,128) backward;
//   KAN stack (kan.py:138-149): 3*sigmoid, KAN(16,1), ReLU, KAN(64,16), ReLU, KAN(192,64) backward with the closed-form
//   basis derivatives of kan_expand<true>.
//
// A CTA of 512 threads owns 8 samples, like the forward.  Every weight gradient is a [rows x 8 samples] x [8 x cols] product:
// the CTA reduces its 8 samples in registers and adds the result into a packed gradient buffer (layout = the packed weight
// buffer `ws`) with one atomicAdd per element; an unpack kernel then writes the 23 gradients in the reference's layouts.
// The two large input-gradient contractions (g0 . Wp0^T over 64 outputs, dH . W1 over 384 hidden units) stream the weights
// from L2 row by row, one warp per row, lanes over the contraction index (coalesced), 8 samples at a time.
#pragma once

constexpr int kHtThreads = 512;
constexpr int kHtSmF = 0;                                   // [192][8]   features, transposed
constexpr int kHtSmA0 = kHtSmF + kHfD * kHfS;               // [192][72]  expanded layer-0 activations; later T [1536][8]
constexpr int kHtSmA1e = kHtSmA0 + kHfK0 * kHfAStride;      // [64][72]
constexpr int kHtSmA2e = kHtSmA1e + kHfO0 * kHfAStride;     // [16][72]
constexpr int kHtSmA1 = kHtSmA2e + kHfO1 * kHfAStride;      // [64][8]    a1 = relu(KAN layer 0)
constexpr int kHtSmA2 = kHtSmA1 + kHfO0 * kHfS;             // [16][8]
constexpr int kHtSmG0 = kHtSmA2 + kHfO1 * kHfS;             // [8][64]    d loss / d (layer-0 pre-activation)
constexpr int kHtSmG1 = kHtSmG0 + kHfS * kHfO0;             // [8][16]
constexpr int kHtSmG2 = kHtSmG1 + kHfS * kHfO1;             // [8]
constexpr int kHtSmGO = kHtSmG2 + kHfS;                     // [9][8]     d loss / d (cls0..3, ord0..2, mu, log_var pre-clamp)
constexpr int kHtSmH = kHtSmGO + 9 * kHfS + 8;              // [8][388]   hidden activations after ReLU + dropout
constexpr int kHtSmDH = kHtSmH + kHfS * kHfHStride;         // [8][388]   d loss / d (hidden pre-activation)
constexpr int kHtSmW2 = kHtSmDH + kHfS * kHfHStride;        // fc2 weights + biases
constexpr int kHtSmTail = kHtSmW2 + kHfW2Floats;            // Wp1, kb1, Wp2, kb2
constexpr int kHtSmDK = kHtSmTail + kHfTailFloats;          // [8][192]   d loss / d features through the KAN branch
constexpr int kHtSmemBytes = (kHtSmDK + kHfS * kHfD) * 4;
static_assert(kHtSmH % 4 == 0 && kHtSmDH % 4 == 0 && kHtSmW2 % 4 == 0 && kHtSmTail % 4 == 0 && kHtSmA0 % 4 == 0, "float4 alignment");
static_assert(kHfK0 * 8 * kHfS <= kHfK0 * kHfAStride, "T aliases the expanded layer-0 activations");

struct HeadsTrainBwdArgs {
  const float* feat; const float* ws;
  const float* h; const float* a1; const float* a2;          // saved by the forward
  const float* lv_out; const float* kan_out;                 // forward outputs (clamp mask, sigmoid derivative)
  const float* d_cls; const float* d_ord; const float* d_mu; const float* d_lv; const float* d_kan;   // any may be null
  float* dfeat; float* dws;
  float drop_p; int batch;
};

// Sum eight per-lane values over the warp in 9 shuffles (recursive halving: each exchange hands the partner the half of the
// values it keeps); on return every lane holds the warp total of p[(lane >> 2) & 7].
__device__ __forceinline__ float ht_warp_sum8(const float (&p)[kHfS], int lane) {
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
  float q[4], r[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = (b4 ? p[i + 4] : p[i]) + __shfl_xor_sync(0xffffffffu, b4 ? p[i] : p[i + 4], 16);
#pragma unroll
  for (int i = 0; i < 2; ++i) r[i] = (b3 ? q[i + 2] : q[i]) + __shfl_xor_sync(0xffffffffu, b3 ? q[i] : q[i + 2], 8);
  float t = (b2 ? r[1] : r[0]) + __shfl_xor_sync(0xffffffffu, b2 ? r[0] : r[1], 4);
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  return t;
}

__global__ void __launch_bounds__(kHtThreads, 1) heads_train_bwd_kernel(const HeadsTrainBwdArgs a, Knots kn) {
  extern __shared__ __align__(16) float hsm[];
  float* sF = hsm + kHtSmF;
  float* sA0 = hsm + kHtSmA0;
  float* sT = sA0;                                   // alias: valid after the layer-0 weight gradient has consumed sA0
  float* sA1e = hsm + kHtSmA1e;
  float* sA2e = hsm + kHtSmA2e;
  float* sA1 = hsm + kHtSmA1;
  float* sA2 = hsm + kHtSmA2;
  float* sG0 = hsm + kHtSmG0;
  float* sG1 = hsm + kHtSmG1;
  float* sG2 = hsm + kHtSmG2;
  float* sGO = hsm + kHtSmGO;
  float* sH = hsm + kHtSmH;
  float* sDH = hsm + kHtSmDH;
  float* sW2 = hsm + kHtSmW2;
  float* sTail = hsm + kHtSmTail;
  float* sDK = hsm + kHtSmDK;
  const float* sWp1 = sTail;                                   // [512][16]
  const float* sWp2 = sTail + (kHfOffWp2 - kHfOffWp1);         // [16][8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s0 = blockIdx.x * kHfS;
  const float inv_keep = (a.drop_p > 0.0f) ? 1.0f / (1.0f - a.drop_p) : 1.0f;

  for (int i = tid; i < kHfW2Floats / 4; i += kHtThreads) hf_cp_async16(sW2 + 4 * i, a.ws + kHfOffW2 + 4 * i);
  for (int i = tid; i < kHfTailFloats / 4; i += kHtThreads) hf_cp_async16(sTail + 4 * i, a.ws + kHfOffWp1 + 4 * i);
  asm volatile("cp.async.commit_group;" ::: "memory");

  // ---- phase 0: inputs of this CTA's 8 samples
  for (int idx = tid; idx < kHfS * kHfD; idx += kHtThreads) {
    const int s = idx / kHfD, k = idx % kHfD;
    sF[k * kHfS + s] = (s0 + s < a.batch) ? a.feat[static_cast<size_t>(s0 + s) * kHfD + k] : 0.0f;
  }
  for (int idx = tid; idx < kHfS * kHfU; idx += kHtThreads) {
    const int s = idx / kHfU, u = idx % kHfU;
    sH[s * kHfHStride + u] = (s0 + s < a.batch) ? a.h[static_cast<size_t>(s0 + s) * kHfU + u] : 0.0f;
  }
  for (int idx = tid; idx < kHfS * kHfO0; idx += kHtThreads) {
    const int s = idx / kHfO0, o = idx % kHfO0;
    sA1[o * kHfS + s] = (s0 + s < a.batch) ? a.a1[static_cast<size_t>(s0 + s) * kHfO0 + o] : 0.0f;
  }
  if (tid < kHfS * kHfO1) {
    const int s = tid / kHfO1, i = tid % kHfO1;
    sA2[i * kHfS + s] = (s0 + s < a.batch) ? a.a2[static_cast<size_t>(s0 + s) * kHfO1 + i] : 0.0f;
  }
  if (tid >= 128 && tid < 128 + 9 * kHfS) {
    const int o = (tid - 128) / kHfS, s = (tid - 128) % kHfS;
    const int sg = s0 + s;
    float g = 0.0f;
    if (sg < a.batch) {
      if (o < 4) { if (a.d_cls != nullptr) g = a.d_cls[static_cast<size_t>(sg) * 4 + o]; }
      else if (o < 7) { if (a.d_ord != nullptr) g = a.d_ord[static_cast<size_t>(sg) * 3 + (o - 4)]; }
      else if (o == 7) { if (a.d_mu != nullptr) g = a.d_mu[sg]; }
      else if (a.d_lv != nullptr) {
        const float lv = a.lv_out[sg];                       // clamp(-10, 10): the gradient passes strictly inside (heads.py:100)
        g = (lv > -10.0f && lv < 10.0f) ? a.d_lv[sg] : 0.0f;
      }
    }
    sGO[o * kHfS + s] = g;
  }
  if (tid >= 256 && tid < 256 + kHfS) {
    const int s = tid - 256, sg = s0 + s;
    float g = 0.0f;
    if (sg < a.batch && a.d_kan != nullptr) {
      const float y = a.kan_out[sg];                         // y = 3 * sigmoid(z): dy/dz = y * (1 - y / 3)
      g = a.d_kan[sg] * y * (3.0f - y) * (1.0f / 3.0f);
    }
    sG2[s] = g;
  }
  __syncthreads();
  // expanded activations of the three KAN layers (values only; derivatives are re-evaluated where they are needed)
  for (int idx = tid; idx < kHfS * kHfK0; idx += kHtThreads) {
    const int s = idx & (kHfS - 1), i = idx >> 3;
    float av[kKW], da[kKW], dt;
    kan_expand<false>(sF[i * kHfS + s], kn, av, da, dt);
#pragma unroll
    for (int k = 0; k < kKW; ++k) sA0[i * kHfAStride + k * kHfS + s] = av[k];
  }
  {
    const int s = tid & (kHfS - 1), i = tid >> 3;             // 512 threads = 64 inputs x 8 samples
    float av[kKW], da[kKW], dt;
    kan_expand<false>(sA1[i * kHfS + s], kn, av, da, dt);
#pragma unroll
    for (int k = 0; k < kKW; ++k) sA1e[i * kHfAStride + k * kHfS + s] = av[k];
    if (tid < kHfS * kHfO1) {
      kan_expand<false>(sA2[i * kHfS + s], kn, av, da, dt);
#pragma unroll
      for (int k = 0; k < kKW; ++k) sA2e[i * kHfAStride + k * kHfS + s] = av[k];
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const bool has_kan = a.d_kan != nullptr;          // uniform over the grid: curriculum stages 1-3 skip the KAN branch
  if (!has_kan) {
    for (int idx = tid; idx < kHfS * kHfD; idx += kHtThreads) sDK[idx] = 0.0f;
  }
  if (has_kan) {
  // ---- phase 1: KAN layer 2 (16 -> 1)
  if (tid < kHfS * kHfO1) {
    const int s = tid & (kHfS - 1), i = tid >> 3;
    float av[kKW], da[kKW], dt;
    const float x2 = sA2[i * kHfS + s];
    kan_expand<true>(x2, kn, av, da, dt);
    float sp = 0.0f;
#pragma unroll
    for (int k = 0; k < kNB; ++k) sp = fmaf(sWp2[i * 8 + k], da[k], sp);
    const float dx = sG2[s] * fmaf(dt, sp, sWp2[i * 8 + 7]);
    sG1[s * kHfO1 + i] = (x2 > 0.0f) ? dx : 0.0f;             // through the ReLU that produced a2
  } else if (tid < 2 * kHfS * kHfO1) {
    const int r = tid - kHfS * kHfO1;                          // packed row i*8 + k of Wp2
    const int i = r >> 3, k = r & 7;
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < kHfS; ++s) acc = fmaf(sA2e[i * kHfAStride + k * kHfS + s], sG2[s], acc);
    if (acc != 0.0f) atomicAdd(a.dws + kHfOffWp2 + r, acc);
  } else if (tid == 2 * kHfS * kHfO1) {
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < kHfS; ++s) acc += sG2[s];
    atomicAdd(a.dws + kHfOffKb2, acc);
  }
  __syncthreads();

  // ---- phase 2: KAN layer 1 (64 -> 16)
  {
    const int s = tid & (kHfS - 1), i = tid >> 3;
    float av[kKW], da[kKW], dt;
    const float x1 = sA1[i * kHfS + s];
    kan_expand<true>(x1, kn, av, da, dt);
    float dx = 0.0f;
#pragma unroll 4
    for (int o = 0; o < kHfO1; ++o) {
      float sp = 0.0f;
#pragma unroll
      for (int k = 0; k < kNB; ++k) sp = fmaf(sWp1[(i * 8 + k) * kHfO1 + o], da[k], sp);
      dx = fmaf(sG1[s * kHfO1 + o], fmaf(dt, sp, sWp1[(i * 8 + 7) * kHfO1 + o]), dx);
    }
    sG0[s * kHfO0 + i] = (x1 > 0.0f) ? dx : 0.0f;
  }
  for (int idx = tid; idx < kHfO0 * 8 * kHfO1; idx += kHtThreads) {
    const int o = idx & (kHfO1 - 1), r = idx >> 4;
    const float* arow = sA1e + (r >> 3) * kHfAStride + (r & 7) * kHfS;
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < kHfS; ++s) acc = fmaf(arow[s], sG1[s * kHfO1 + o], acc);
    if (acc != 0.0f) atomicAdd(a.dws + kHfOffWp1 + idx, acc);
  }
  if (tid < kHfO1) {
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < kHfS; ++s) acc += sG1[s * kHfO1 + tid];
    atomicAdd(a.dws + kHfOffKb1 + tid, acc);
  }
  __syncthreads();

  // ---- phase 3: KAN layer 0 (192 -> 64).  Weight gradient: thread = (output o, row group), g0[.][o] in registers
  {
    const int o = tid & (kHfO0 - 1), rg = tid >> 6;
    float g[kHfS];
#pragma unroll
    for (int s = 0; s < kHfS; ++s) g[s] = sG0[s * kHfO0 + o];
    for (int r = rg; r < kHfK0 * 8; r += kHtThreads / kHfO0) {
      const float* arow = sA0 + (r >> 3) * kHfAStride + (r & 7) * kHfS;
      const float4 aa = *reinterpret_cast<const float4*>(arow);
      const float4 ab = *reinterpret_cast<const float4*>(arow + 4);
      float acc = aa.x * g[0];
      acc = fmaf(aa.y, g[1], acc); acc = fmaf(aa.z, g[2], acc); acc = fmaf(aa.w, g[3], acc);
      acc = fmaf(ab.x, g[4], acc); acc = fmaf(ab.y, g[5], acc); acc = fmaf(ab.z, g[6], acc); acc = fmaf(ab.w, g[7], acc);
      if (acc != 0.0f) atomicAdd(a.dws + kHfOffWp0 + static_cast<size_t>(r) * kHfO0 + o, acc);
    }
    if (tid < kHfO0) {
      float acc = 0.0f;
#pragma unroll
      for (int s = 0; s < kHfS; ++s) acc += g[s];
      atomicAdd(a.dws + kHfOffKb0 + tid, acc);
    }
  }
  __syncthreads();          // sA0 is free: T may overwrite it
  // T[r][s] = sum_o g0[s][o] * Wp0[r][o]: one warp per packed row, lanes over the 64 outputs
  {
    float g0a[kHfS], g0b[kHfS];                       // this lane's two output columns of g0, for the 8 samples
#pragma unroll
    for (int s = 0; s < kHfS; ++s) { g0a[s] = sG0[s * kHfO0 + lane]; g0b[s] = sG0[s * kHfO0 + lane + 32]; }
    constexpr int kWarps = kHtThreads / 32, kRows = kHfK0 * 8 / kWarps;      // 96 rows per warp
    static_assert(kRows % 4 == 0, "row batches of 4");
    for (int rb = 0; rb < kRows; rb += 4) {           // four rows per batch: eight L2 loads in flight per lane
      float w0[4], w1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = warp + (rb + q) * kWarps;
        w0[q] = __ldg(a.ws + kHfOffWp0 + static_cast<size_t>(r) * kHfO0 + lane);
        w1[q] = __ldg(a.ws + kHfOffWp0 + static_cast<size_t>(r) * kHfO0 + lane + 32);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = warp + (rb + q) * kWarps;
        float p[kHfS];
#pragma unroll
        for (int s = 0; s < kHfS; ++s) p[s] = fmaf(w0[q], g0a[s], w1[q] * g0b[s]);
        const float tot = ht_warp_sum8(p, lane);
        if ((lane & 3) == 0) sT[r * kHfS + (lane >> 2)] = tot;
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < kHfS * kHfK0; idx += kHtThreads) {
    const int s = idx & (kHfS - 1), i = idx >> 3;
    float av[kKW], da[kKW], dt;
    kan_expand<true>(sF[i * kHfS + s], kn, av, da, dt);
    float sp = 0.0f;
#pragma unroll
    for (int k = 0; k < kNB; ++k) sp = fmaf(sT[(i * 8 + k) * kHfS + s], da[k], sp);
    sDK[s * kHfD + i] = fmaf(dt, sp, sT[(i * 8 + 7) * kHfS + s]);
  }
  }   // has_kan

  // ---- phase 4: second layers of the three heads
  for (int idx = tid; idx < kHfS * kHfU; idx += kHtThreads) {
    const int s = idx / kHfU, u = idx % kHfU;
    const int head = u >> 7, j = u & (kHfH - 1);
    float acc;
    if (head == 0) {
      acc = sGO[0 * kHfS + s] * sW2[0 * kHfH + j];
#pragma unroll
      for (int o = 1; o < 4; ++o) acc = fmaf(sGO[o * kHfS + s], sW2[o * kHfH + j], acc);
    } else if (head == 1) {
      acc = sGO[4 * kHfS + s] * sW2[4 * kHfH + j];
#pragma unroll
      for (int o = 5; o < 7; ++o) acc = fmaf(sGO[o * kHfS + s], sW2[o * kHfH + j], acc);
    } else {
      acc = fmaf(sGO[7 * kHfS + s], sW2[7 * kHfH + j], sGO[8 * kHfS + s] * sW2[8 * kHfH + j]);
    }
    sDH[s * kHfHStride + u] = (sH[s * kHfHStride + u] > 0.0f) ? acc * inv_keep : 0.0f;     // ReLU and dropout mask in one
  }
  for (int idx = tid; idx < 9 * kHfH; idx += kHtThreads) {
    const int o = idx >> 7, j = idx & (kHfH - 1);
    const int head = (o < 4) ? 0 : (o < 7) ? 1 : 2;
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < kHfS; ++s) acc = fmaf(sGO[o * kHfS + s], sH[s * kHfHStride + head * kHfH + j], acc);
    if (acc != 0.0f) atomicAdd(a.dws + kHfOffW2 + idx, acc);
  }
  if (tid < 9) {
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < kHfS; ++s) acc += sGO[tid * kHfS + s];
    atomicAdd(a.dws + kHfOffB2 + tid, acc);
  }
  __syncthreads();

  // ---- phase 5: first layers of the heads
  for (int idx = tid; idx < kHfD * kHfU; idx += kHtThreads) {
    const int k = idx / kHfU, u = idx % kHfU;
    const float4 fa = *reinterpret_cast<const float4*>(sF + k * kHfS);
    const float4 fb = *reinterpret_cast<const float4*>(sF + k * kHfS + 4);
    float acc = fa.x * sDH[0 * kHfHStride + u];
    acc = fmaf(fa.y, sDH[1 * kHfHStride + u], acc); acc = fmaf(fa.z, sDH[2 * kHfHStride + u], acc);
    acc = fmaf(fa.w, sDH[3 * kHfHStride + u], acc); acc = fmaf(fb.x, sDH[4 * kHfHStride + u], acc);
    acc = fmaf(fb.y, sDH[5 * kHfHStride + u], acc); acc = fmaf(fb.z, sDH[6 * kHfHStride + u], acc);
    acc = fmaf(fb.w, sDH[7 * kHfHStride + u], acc);
    if (acc != 0.0f) atomicAdd(a.dws + kHfOffW1T + idx, acc);
  }
  if (tid < kHfU) {
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < kHfS; ++s) acc += sDH[s * kHfHStride + tid];
    atomicAdd(a.dws + kHfOffB1 + tid, acc);
  }
  // d features through the heads: one warp per feature k, lanes over the 384 hidden units
  for (int k = warp; k < kHfD; k += kHtThreads / 32) {
    float p[kHfS];
#pragma unroll
    for (int s = 0; s < kHfS; ++s) p[s] = 0.0f;
    float w[kHfU / 32];                               // the whole row slice of this lane first: 12 L2 loads in flight
#pragma unroll
    for (int m = 0; m < kHfU / 32; ++m) w[m] = __ldg(a.ws + kHfOffW1T + static_cast<size_t>(k) * kHfU + lane + 32 * m);
#pragma unroll
    for (int m = 0; m < kHfU / 32; ++m) {
      const int u = lane + 32 * m;
#pragma unroll
      for (int s = 0; s < kHfS; ++s) p[s] = fmaf(w[m], sDH[s * kHfHStride + u], p[s]);
    }
    const float tot = ht_warp_sum8(p, lane);
    if ((lane & 3) == 0) sDK[(lane >> 2) * kHfD + k] += tot;      // (this warp owns column k; sDK was completed before the last barrier)
  }
  __syncthreads();
  for (int idx = tid; idx < kHfS * kHfD; idx += kHtThreads) {
    const int s = idx / kHfD;
    if (s0 + s < a.batch) a.dfeat[static_cast<size_t>(s0) * kHfD + idx] = sDK[idx];
  }
}

// packed gradient buffer (layout of `ws`) -> the 23 parameter gradients in the reference's layouts (overwritten)
struct HeadsGradPtrs {
  float* fc1_w[3]; float* fc1_b[3];
  float* fc2_w[4]; float* fc2_b[4];
  float* spline[3]; float* lin_w[3]; float* lin_b[3];
};

__global__ void heads_fused_unpack_grad_kernel(const float* __restrict__ dws, const HeadsGradPtrs p) {
  const int n = kHfWsFloats;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const float v = dws[idx];
    float* dst = nullptr;
    if (idx < kHfOffB1) {
      const int k = idx / kHfU, u = idx % kHfU;
      if (p.fc1_w[u / kHfH] != nullptr) dst = p.fc1_w[u / kHfH] + (u % kHfH) * kHfD + k;
    } else if (idx < kHfOffW2) {
      const int u = idx - kHfOffB1;
      if (p.fc1_b[u / kHfH] != nullptr) dst = p.fc1_b[u / kHfH] + u % kHfH;
    } else if (idx < kHfOffB2) {
      const int o = (idx - kHfOffW2) / kHfH, j = (idx - kHfOffW2) % kHfH;
      if (o < 4) { if (p.fc2_w[0] != nullptr) dst = p.fc2_w[0] + o * kHfH + j; }
      else if (o < 7) { if (p.fc2_w[1] != nullptr) dst = p.fc2_w[1] + (o - 4) * kHfH + j; }
      else if (p.fc2_w[o - 5] != nullptr) dst = p.fc2_w[o - 5] + j;
    } else if (idx < kHfOffWp0) {
      const int o = idx - kHfOffB2;
      if (o < 4) { if (p.fc2_b[0] != nullptr) dst = p.fc2_b[0] + o; }
      else if (o < 7) { if (p.fc2_b[1] != nullptr) dst = p.fc2_b[1] + (o - 4); }
      else if (o < 9 && p.fc2_b[o - 5] != nullptr) dst = p.fc2_b[o - 5];
    } else if (idx < kHfOffKb0) {
      const int kk = (idx - kHfOffWp0) / kHfO0, o = (idx - kHfOffWp0) % kHfO0, i = kk >> 3, k = kk & 7;
      if (p.spline[0] != nullptr) dst = (k < 7) ? p.spline[0] + (i * kHfO0 + o) * 7 + k : p.lin_w[0] + o * kHfK0 + i;
    } else if (idx < kHfOffWp1) {
      if (p.lin_b[0] != nullptr) dst = p.lin_b[0] + (idx - kHfOffKb0);
    } else if (idx < kHfOffKb1) {
      const int kk = (idx - kHfOffWp1) / kHfO1, o = (idx - kHfOffWp1) % kHfO1, i = kk >> 3, k = kk & 7;
      if (p.spline[1] != nullptr) dst = (k < 7) ? p.spline[1] + (i * kHfO1 + o) * 7 + k : p.lin_w[1] + o * kHfO0 + i;
    } else if (idx < kHfOffWp2) {
      if (p.lin_b[1] != nullptr) dst = p.lin_b[1] + (idx - kHfOffKb1);
    } else if (idx < kHfOffKb2) {
      const int kk = idx - kHfOffWp2, i = kk >> 3, k = kk & 7;
      if (p.spline[2] != nullptr) dst = (k < 7) ? p.spline[2] + i * 7 + k : p.lin_w[2] + i;
    } else if (idx == kHfOffKb2) {
      if (p.lin_b[2] != nullptr) dst = p.lin_b[2];
    }
    if (dst != nullptr) *dst = v;
  }
}
