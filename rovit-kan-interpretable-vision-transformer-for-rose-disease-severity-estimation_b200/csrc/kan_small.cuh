// KAN layers with few outputs (n_out <= 16: the 64 -> 16 and 16 -> 1 layers of the production stack, the 64 -> 1 layer of
// the BASELINE microbenchmark), fp32, included by kan.cu inside its anonymous namespace.
//
// The general kernels tile 64 outputs per CTA: for n_out = 1 they spend 63/64 of their FMAs and shared memory on padding
// (the 64 -> 1 layer cost as much as the 192 -> 64 tensor-core layer).  Here only the <= 4 live basis functions of an
// input are touched (support-4 instead of dense-7 contraction), the packed weights of the whole layer sit in shared
// memory ([in][8 slots][NOUT], slot 7 = the linear branch), and nothing is packed or unpacked by separate launches:
//
//   forward   one warp per sample: lane = input (strided), NOUT partial sums per lane, warp-shuffle reduction.
//   backward  ONE kernel for dx, dW, dWl and db: thread (input i, output quad q) keeps its 8 x OPT weight-gradient
//             accumulators in REGISTERS across all samples of its CTA (no atomics in the sample loop); the expansion
//             (tanh, interval, cubics and their derivatives) is evaluated once per (sample, input) and serves both dx and
//             dW; dx is reduced over the output quads by shuffles; per-CTA results leave through one atomicAdd each.
#pragma once

constexpr int kSmThreads = 256;

// interval j of t on the fp32 knot buffer and the four cubic segments v[m] = N_{j-m}(t) (optionally d/dt): the arithmetic
// of kan_basis_at, without the one-hot placement.  Returns false in the dead zone (j >= 7) and left of the first knot.
template <bool DERIV>
__device__ __forceinline__ bool kan_segment(float t, const Knots& kn, int& j, float (&v)[4], float (&d)[4]) {
  j = 0;
#pragma unroll
  for (int m = 1; m < kKnots; ++m) j += (t >= kn.k[m]) ? 1 : 0;
  if (!(j < kNB && t >= kn.k[0])) return false;
  float k0 = kn.k[0], k1 = kn.k[1];
#pragma unroll
  for (int m = 1; m < kNB; ++m)
    if (j == m) { k0 = kn.k[m]; k1 = kn.k[m + 1]; }
  const float ih = 1.0f / (k1 - k0);
  const float u = (t - k0) * ih;
  const float u2 = u * u, u3 = u2 * u;
  const float om = 1.0f - u;
  v[0] = u3 * (1.0f / 6.0f);
  v[1] = (1.0f + 3.0f * u + 3.0f * u2 - 3.0f * u3) * (1.0f / 6.0f);
  v[2] = (4.0f - 6.0f * u2 + 3.0f * u3) * (1.0f / 6.0f);
  v[3] = om * om * om * (1.0f / 6.0f);
  if (DERIV) {
    d[0] = 0.5f * u2 * ih;
    d[1] = (3.0f + 6.0f * u - 9.0f * u2) * (1.0f / 6.0f) * ih;
    d[2] = (-12.0f * u + 9.0f * u2) * (1.0f / 6.0f) * ih;
    d[3] = -0.5f * om * om * ih;
  }
  return true;
}

// sW[(i*8 + k) * NOUT + o] = W[i,o,k] (k < 7), Wl[o,i] (k = 7); zero for o >= n_out
template <int NOUT>
__device__ __forceinline__ void kan_small_load_weights(float* sW, const float* __restrict__ spline,
                                                       const float* __restrict__ lin_w, int n_in, int n_out) {
  for (int idx = threadIdx.x; idx < n_in * kKW * NOUT; idx += blockDim.x) {
    const int o = idx % NOUT, kk = idx / NOUT, i = kk >> 3, k = kk & 7;
    float w = 0.0f;
    if (o < n_out) w = (k < kNB) ? spline[(static_cast<size_t>(i) * n_out + o) * kNB + k] : lin_w[static_cast<size_t>(o) * n_in + i];
    sW[idx] = w;
  }
}

__device__ __forceinline__ float kan_act(int act, float v) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return 3.0f / (1.0f + expf(-v));
  return v;
}

template <int NOUT>
__global__ void __launch_bounds__(kSmThreads)
kan_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ spline, const float* __restrict__ lin_w,
                     const float* __restrict__ bias, Knots kn, float* __restrict__ y, int act, int batch, int n_in, int n_out) {
  extern __shared__ __align__(16) float sm_small[];
  float* sW = sm_small;
  kan_small_load_weights<NOUT>(sW, spline, lin_w, n_in, n_out);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kSmThreads / 32;
  const int stride = gridDim.x * kWarps;
  float x_pre = 0.0f;                                          // first input of the NEXT sample of this warp, loaded one sample ahead
  {
    const int b0 = blockIdx.x * kWarps + warp;
    if (b0 < batch && lane < n_in) x_pre = x[static_cast<size_t>(b0) * n_in + lane];
  }
  for (int b = blockIdx.x * kWarps + warp; b < batch; b += stride) {
    float acc[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) acc[o] = 0.0f;
    const float* xr = x + static_cast<size_t>(b) * n_in;
    const float x_first = x_pre;
    if (b + stride < batch && lane < n_in) x_pre = x[static_cast<size_t>(b + stride) * n_in + lane];
    for (int i = lane; i < n_in; i += 32) {
      const float xv = (i == lane) ? x_first : xr[i];
      const float* row = sW + i * kKW * NOUT;
#pragma unroll
      for (int o = 0; o < NOUT; ++o) acc[o] = fmaf(xv, row[7 * NOUT + o], acc[o]);
      int j;
      float v[4], d[4];
      if (kan_segment<false>(tanhf(xv), kn, j, v, d)) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int slot = j - m;
          if (slot >= 0) {
            const float* wr = row + slot * NOUT;
#pragma unroll
            for (int o = 0; o < NOUT; ++o) acc[o] = fmaf(v[m], wr[o], acc[o]);
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
    }
    float mine = 0.0f;
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
      if (lane == o) mine = acc[o];
    if (lane < n_out) y[static_cast<size_t>(b) * n_out + lane] = kan_act(act, mine + bias[lane]);
  }
}

// Backward.  NOUT = padded outputs (1, 2, 4, 8, 16); OPT = min(NOUT, 4) outputs per thread, OQ = NOUT / OPT output quads;
// a "stream" of TPS = n_in * OQ threads walks samples; G = 256 / TPS streams per CTA.  Thread (i, q) of a stream owns
// dWp[i][0..7][4q .. 4q+OPT) in registers.  Neighbouring lanes q = 0..OQ-1 share input i (OQ divides 32): lane q == 0
// evaluates the expansion and the others receive it by shuffle.
template <int NOUT>
__global__ void __launch_bounds__(kSmThreads)
kan_small_bwd_kernel(const float* __restrict__ x, const float* __restrict__ yv, const float* __restrict__ gy,
                     const float* __restrict__ spline, const float* __restrict__ lin_w, Knots kn, int act,
                     float* __restrict__ dx, float* __restrict__ dspline, float* __restrict__ dlin_w, float* __restrict__ dlin_b,
                     int batch, int n_in, int n_out, int samples_per_cta) {
  constexpr int OPT = NOUT < 4 ? NOUT : 4;
  constexpr int OQ = NOUT / OPT;
  extern __shared__ __align__(16) float sm_small[];
  float* sW = sm_small;                                   // [n_in*8][NOUT]
  kan_small_load_weights<NOUT>(sW, spline, lin_w, n_in, n_out);
  const int tps = n_in * OQ;
  const int G = kSmThreads / tps;
  const int stream = threadIdx.x / tps, within = threadIdx.x % tps;
  const bool active = stream < G;
  const int i = within / OQ, q = within % OQ;
  const int lane = threadIdx.x & 31;
  const int lead = lane - (lane % OQ);                    // lane of q == 0 for this input (same warp: OQ divides 32, tps % OQ == 0)
  // weight-gradient accumulators: PRIVATE per thread, but in shared memory ([slot][c][thread]: conflict-free) so that the
  // slot j - m can index them directly; with register accumulators the one-hot placement cost ~110 select instructions per
  // (sample, input) (ncu: 320 instructions per pair, ALU pipe 44 %)
  float* sAcc = sW + n_in * kKW * NOUT;                   // [8][OPT][256]
#pragma unroll
  for (int k = 0; k < kKW; ++k)
#pragma unroll
    for (int c = 0; c < OPT; ++c) sAcc[(k * OPT + c) * kSmThreads + threadIdx.x] = 0.0f;
  float db[OPT];
#pragma unroll
  for (int c = 0; c < OPT; ++c) db[c] = 0.0f;
  const int b_begin = blockIdx.x * samples_per_cta;
  const int b_end = min(batch, b_begin + samples_per_cta);
  __syncthreads();
  // all lanes of a warp run the same number of rounds (the shuffles below are warp-wide); no block-level barrier in the loop
  const int rounds = (b_end - b_begin + G - 1) / (G > 0 ? G : 1);
  const float* row = sW + i * kKW * NOUT + q * OPT;
  // the operands of round r + 1 are loaded before round r is evaluated (ncu on the first version: 49 % of all warp samples
  // waited for these loads)
  float x_nxt = 0.0f, gy_nxt[OPT], y_nxt[OPT];
  auto fetch = [&](int r) {
    const int b = b_begin + r * G + stream;
    const bool ok = active && r < rounds && b < b_end;
    x_nxt = (ok && q == 0) ? x[static_cast<size_t>(b) * n_in + i] : 0.0f;
#pragma unroll
    for (int c = 0; c < OPT; ++c) {
      const int o = q * OPT + c;
      const bool oo = ok && o < n_out;
      const size_t off = static_cast<size_t>(oo ? b : 0) * n_out + (oo ? o : 0);
      gy_nxt[c] = oo ? gy[off] : 0.0f;                    // the same few words for every input of the sample: L1 broadcast
      y_nxt[c] = oo ? yv[off] : 0.0f;
    }
  };
  fetch(0);
  for (int r = 0; r < rounds; ++r) {
    const int b = b_begin + r * G + stream;
    const bool live_b = active && b < b_end;
    float g[OPT];
#pragma unroll
    for (int c = 0; c < OPT; ++c) g[c] = gy_nxt[c] * act_grad(act, y_nxt[c]);
    const float x_cur = x_nxt;
    fetch(r + 1);
    float xv = 0.0f, dtdx = 0.0f;
    int j = 8;
    float v[4] = {0.f, 0.f, 0.f, 0.f}, d[4] = {0.f, 0.f, 0.f, 0.f};
    if (live_b && q == 0) {
      xv = x_cur;
      const float t = tanhf(xv);
      dtdx = 1.0f - t * t;
      if (!kan_segment<true>(t, kn, j, v, d)) j = 8;
    }
    if (OQ > 1) {
      xv = __shfl_sync(0xffffffffu, xv, lead);
      dtdx = __shfl_sync(0xffffffffu, dtdx, lead);
      j = __shfl_sync(0xffffffffu, j, lead);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        v[m] = __shfl_sync(0xffffffffu, v[m], lead);
        d[m] = __shfl_sync(0xffffffffu, d[m], lead);
      }
    }
    float dxp = 0.0f;
    if (live_b) {
      // linear branch (slot 7)
#pragma unroll
      for (int c = 0; c < OPT; ++c) {
        sAcc[(7 * OPT + c) * kSmThreads + threadIdx.x] += xv * g[c];
        dxp = fmaf(g[c], row[7 * NOUT + c], dxp);
        db[c] += g[c];
      }
      // spline branch: segment m lives in slot j - m (j = 8 in the dead zone: skipped)
      float sp = 0.0f;
      if (j < kNB) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int slot = j - m;
          if (slot >= 0) {
#pragma unroll
            for (int c = 0; c < OPT; ++c) {
              sAcc[(slot * OPT + c) * kSmThreads + threadIdx.x] += v[m] * g[c];
              sp = fmaf(g[c] * row[slot * NOUT + c], d[m], sp);
            }
          }
        }
      }
      dxp = fmaf(dtdx, sp, dxp);
    }
    if (OQ > 1) {
#pragma unroll
      for (int s = 1; s < OQ; s <<= 1) dxp += __shfl_xor_sync(0xffffffffu, dxp, s);
    }
    if (live_b && q == 0 && dx != nullptr) dx[static_cast<size_t>(b) * n_in + i] = dxp;
  }
  if (dspline == nullptr) return;
  // reduce the G streams of this CTA, then ONE atomicAdd per gradient element and CTA: with one flush per thread the 513
  // addresses of a 64 -> 1 layer took millions of same-address atomics (measured: 232 us for the kernel, 76 us without)
  constexpr int kPer = kKW * OPT + OPT;
  float* sDb = sAcc + kKW * OPT * kSmThreads;             // [OPT][256]
#pragma unroll
  for (int c = 0; c < OPT; ++c) sDb[c * kSmThreads + threadIdx.x] = db[c];
  __syncthreads();
  if (stream != 0 || !active) return;
  float tot[kPer];
#pragma unroll
  for (int e = 0; e < kPer; ++e) tot[e] = 0.0f;
  for (int g = 0; g < G; ++g) {
    const int t = g * tps + within;                       // the thread of stream g that owns the same (input, output quad)
#pragma unroll
    for (int e = 0; e < kKW * OPT; ++e) tot[e] += sAcc[e * kSmThreads + t];
#pragma unroll
    for (int c = 0; c < OPT; ++c) tot[kKW * OPT + c] += sDb[c * kSmThreads + t];
  }
  // per-CTA results -> global, straight into the reference layouts (dspline [in][out][7], dlin_w [out][in])
#pragma unroll
  for (int c = 0; c < OPT; ++c) {
    const int o = q * OPT + c;
    if (o >= n_out) continue;
#pragma unroll
    for (int k = 0; k < kNB; ++k)
      if (tot[k * OPT + c] != 0.0f) atomicAdd(&dspline[(static_cast<size_t>(i) * n_out + o) * kNB + k], tot[k * OPT + c]);
    atomicAdd(&dlin_w[static_cast<size_t>(o) * n_in + i], tot[7 * OPT + c]);
  }
  // bias gradient: every input thread of an output quad holds the same sum; input 0 writes it
  if (i == 0) {
#pragma unroll
    for (int c = 0; c < OPT; ++c)
      if (q * OPT + c < n_out) atomicAdd(&dlin_b[q * OPT + c], tot[kKW * OPT + c]);
  }
}

// few outputs: always; up to 16 outputs: only where the tensor-core formulation is not available (small batches, odd
// shapes) -- at 16 outputs the support-4 gather costs 5 x 16 FMAs + loads per (sample, input) on the CUDA cores, more than
// the dense-7 product costs on the tensor pipe (measured: 64 -> 16 at batch 65536, 170 us against 45 us)
bool kan_small_ok(int n_in, int n_out, bool tc_available) {
  if (n_out > 16 || n_in < 1) return false;
  if (n_out > 4 && tc_available) return false;
  const int nout = n_out <= 1 ? 1 : n_out <= 2 ? 2 : n_out <= 4 ? 4 : n_out <= 8 ? 8 : 16;
  const int oq = nout <= 4 ? 1 : nout / 4;
  return n_in * oq <= kSmThreads && static_cast<size_t>(n_in) * kKW * nout * 4 <= 60 * 1024;
}
bool kan_small_disabled() {
  static const bool off = [] { const char* e = getenv("RVK_KAN_NO_SMALL"); return e != nullptr && e[0] == '1'; }();
  return off;
}

template <int NOUT>
int kan_small_fwd_launch_t(const KanLayerDesc& L, const float* x, float* y, int act, int batch, const Knots& kn, cudaStream_t stream) {
  const int smem = L.in_features * kKW * NOUT * 4;
  auto kernel = kan_small_fwd_kernel<NOUT>;
  if (smem > 48 * 1024) RVK_SET_MAX_SMEM(kernel, 96 * 1024);
  const int warps = kSmThreads / 32;
  int grid = (batch + warps - 1) / warps;
  if (grid > kNumSMsB200 * 4) grid = kNumSMsB200 * 4;
  kernel<<<grid, kSmThreads, smem, stream>>>(x, L.spline, L.lin_w, L.lin_b, kn, y, act, batch, L.in_features, L.out_features);
  return rvk_launch_check();
}

template <int NOUT>
int kan_small_bwd_launch_t(const KanLayerDesc& L, const float* x, const float* y, const float* gy, int act, float* dx, float* dspline,
                           float* dlin_w, float* dlin_b, int batch, const Knots& kn, cudaStream_t stream) {
  constexpr int OQ = NOUT <= 4 ? 1 : NOUT / 4;
  constexpr int OPT = NOUT < 4 ? NOUT : 4;
  const int G = kSmThreads / (L.in_features * OQ);
  (void)G;
  const int smem = (L.in_features * kKW * NOUT + (kKW * OPT + OPT) * kSmThreads) * 4;    // weights + per-thread accumulators
  auto kernel = kan_small_bwd_kernel<NOUT>;
  if (smem > 48 * 1024) RVK_SET_MAX_SMEM(kernel, 100 * 1024);
  // enough resident warps to hide the load latency of the sample loop (8 CTAs per SM), each CTA with at least a few rounds of
  // work: its register accumulators are flushed once, n_in * 8 * n_out atomics per CTA
  static const int per_sm = [] { const char* e = getenv("RVK_KAN_SMALL_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 4; }();
  int ctas = kNumSMsB200 * per_sm;
  const int min_per_cta = G * 8;
  if (static_cast<long long>(ctas) * min_per_cta > batch) ctas = (batch + min_per_cta - 1) / min_per_cta;
  if (ctas < 1) ctas = 1;
  const int spc = (batch + ctas - 1) / ctas;
  ctas = (batch + spc - 1) / spc;
  kernel<<<ctas, kSmThreads, smem, stream>>>(x, y, gy, L.spline, L.lin_w, kn, act, dx, dspline, dlin_w, dlin_b, batch,
                                            L.in_features, L.out_features, spc);
  return rvk_launch_check();
}

int kan_small_fwd_launch(const KanLayerDesc& L, const float* x, float* y, int act, int batch, const Knots& kn, cudaStream_t stream) {
  const int n = L.out_features;
  if (n <= 1) return kan_small_fwd_launch_t<1>(L, x, y, act, batch, kn, stream);
  if (n <= 2) return kan_small_fwd_launch_t<2>(L, x, y, act, batch, kn, stream);
  if (n <= 4) return kan_small_fwd_launch_t<4>(L, x, y, act, batch, kn, stream);
  if (n <= 8) return kan_small_fwd_launch_t<8>(L, x, y, act, batch, kn, stream);
  return kan_small_fwd_launch_t<16>(L, x, y, act, batch, kn, stream);
}

int kan_small_bwd_launch(const KanLayerDesc& L, const float* x, const float* y, const float* gy, int act, float* dx, float* dspline,
                         float* dlin_w, float* dlin_b, int batch, const Knots& kn, cudaStream_t stream) {
  const int n = L.out_features;
  if (n <= 1) return kan_small_bwd_launch_t<1>(L, x, y, gy, act, dx, dspline, dlin_w, dlin_b, batch, kn, stream);
  if (n <= 2) return kan_small_bwd_launch_t<2>(L, x, y, gy, act, dx, dspline, dlin_w, dlin_b, batch, kn, stream);
  if (n <= 4) return kan_small_bwd_launch_t<4>(L, x, y, gy, act, dx, dspline, dlin_w, dlin_b, batch, kn, stream);
  if (n <= 8) return kan_small_bwd_launch_t<8>(L, x, y, gy, act, dx, dspline, dlin_w, dlin_b, batch, kn, stream);
  return kan_small_bwd_launch_t<16>(L, x, y, gy, act, dx, dspline, dlin_w, dlin_b, batch, kn, stream);
}
