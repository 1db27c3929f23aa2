// Fused KAN layer (reference models/kan.py:70-95), forward and backward, fp32.
//
//   y[b,o] = bias[o] + sum_i x[b,i]*Wl[o,i] + sum_i sum_k N_k(tanh x[b,i]) * W[i,o,k]
//
// The reference builds a (B,in,7) basis tensor with ~150 elementwise launches and contracts it in
// a Python double loop.  Here the truncated cubic B-spline is evaluated in closed form per
// (sample, input) -- interval j on the fp32 knot buffer, u = (t - knot_j)/h, four live cubics, all
// zero for j >= 7 (the reference's missing degree-0 seeds, kan.py:12-13,23-25,39-40) -- and the
// basis values never leave the SM: each (sample, input) expands to 8 "activations"
// [N_0..N_6, x] written to shared memory and contracted immediately against a packed weight
//   Wp[i*8 + k][o] = W[i,o,k] (k < 7),  Wp[i*8 + 7][o] = Wl[o,i]
// with a register-tiled outer-product loop (4 samples x 4 outputs per thread).
//
// Backward: gpre = gy * act'(y);   dWp = A^T gpre  (batch-split, register tile 8 x 4, atomics);
//           G = gpre * Wp^T,  dx[b,i] = G[b,i,7] + (1 - t^2) * sum_k G[b,i,k] * N'_k(t).
#include "kernels.h"
#include "tma_host.h"

#include <cmath>
#include <cstdlib>

namespace {

constexpr int kNB = 7;          // basis functions
constexpr int kKW = 8;          // packed width per input: 7 basis + raw x
constexpr int kKnots = 11;
constexpr int kIC = 8;          // inputs per chunk -> 64 packed rows
constexpr int kKC = kIC * kKW;  // 64
constexpr int kTS = 64;         // samples per CTA tile
constexpr int kTO = 64;         // outputs per CTA tile

struct Knots { float k[kKnots]; };

// basis values (and optionally d/dt) of the truncated cubic B-spline family at normalised input t;
// a[7] / da[7] are left untouched
template <bool DERIV>
__device__ __forceinline__ void kan_basis_at(float t, const Knots& kn, float (&a)[kKW], float (&da)[kKW]) {
#pragma unroll
  for (int k = 0; k < kNB; ++k) { a[k] = 0.0f; if (DERIV) da[k] = 0.0f; }
  int j = 0;
#pragma unroll
  for (int m = 1; m < kKnots; ++m) j += (t >= kn.k[m]) ? 1 : 0;
  if (j < kNB && t >= kn.k[0]) {
    const float k0 = kn.k[j];
    const float h = kn.k[j + 1] - k0;
    const float u = (t - k0) / h;
    const float u2 = u * u, u3 = u2 * u;
    const float om = 1.0f - u;
    float v[4], d[4];
    v[0] = u3 * (1.0f / 6.0f);
    v[1] = (1.0f + 3.0f * u + 3.0f * u2 - 3.0f * u3) * (1.0f / 6.0f);
    v[2] = (4.0f - 6.0f * u2 + 3.0f * u3) * (1.0f / 6.0f);
    v[3] = om * om * om * (1.0f / 6.0f);
    if (DERIV) {
      const float ih = 1.0f / h;
      d[0] = 0.5f * u2 * ih;
      d[1] = (3.0f + 6.0f * u - 9.0f * u2) * (1.0f / 6.0f) * ih;
      d[2] = (-12.0f * u + 9.0f * u2) * (1.0f / 6.0f) * ih;
      d[3] = -0.5f * om * om * ih;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int idx = j - m;
#pragma unroll
      for (int k = 0; k < kNB; ++k)
        if (idx == k) { a[k] = v[m]; if (DERIV) da[k] = d[m]; }
    }
  }
}

// the 8 packed activations [N_0(tanh x) .. N_6(tanh x), x] of scalar input x
template <bool DERIV>
__device__ __forceinline__ void kan_expand(float x, const Knots& kn, float (&a)[kKW], float (&da)[kKW], float& dtdx) {
  const float t = tanhf(x);                 // precise tanhf: a 1-ulp move across 0.4 flips a term by O(0.1)
  kan_basis_at<DERIV>(t, kn, a, da);
  a[7] = x;
  if (DERIV) { da[7] = 0.0f; dtdx = 1.0f - t * t; }
}

// BSplineBasis.compute_basis (kan.py:10-44) on already-normalised inputs: out[n,7]
__global__ void kan_basis_kernel(const float* __restrict__ t, Knots kn, float* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[kKW], da[kKW];
  const float tc = fminf(fmaxf(t[i], kn.k[0]), kn.k[kKnots - 1]);     // kan.py:16
  kan_basis_at<false>(tc, kn, a, da);
#pragma unroll
  for (int k = 0; k < kNB; ++k) out[i * kNB + k] = a[k];
}

__device__ __forceinline__ float act_grad(int act, float y) {
  if (act == 1) return y > 0.0f ? 1.0f : 0.0f;            // relu
  if (act == 2) return y * (3.0f - y) * (1.0f / 3.0f);    // y = 3*sigmoid(z): dy/dz = y*(1 - y/3)
  return 1.0f;
}

// ------------------------------------------------------------------ weight packing
// Wp[(i*8+k) * out_pad + o], WpT[o * (in_pad*8) + i*8+k]; pads are zero.
__global__ void kan_pack_kernel(const float* __restrict__ spline, const float* __restrict__ lin_w, int n_in, int n_out,
                                int in_pad, int out_pad, float* __restrict__ Wp, float* __restrict__ WpT) {
  const long long total = static_cast<long long>(in_pad) * kKW * out_pad;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(idx % out_pad);
    const int kk = static_cast<int>(idx / out_pad);
    const int i = kk / kKW, k = kk % kKW;
    float v = 0.0f;
    if (i < n_in && o < n_out)
      v = (k < kNB) ? spline[(static_cast<size_t>(i) * n_out + o) * kNB + k] : lin_w[static_cast<size_t>(o) * n_in + i];
    Wp[idx] = v;
    if (WpT != nullptr) WpT[static_cast<size_t>(o) * (in_pad * kKW) + kk] = v;
  }
}

// dspline[i,o,k] += dWp[i*8+k][o]; dlin_w[o,i] += dWp[i*8+7][o]
__global__ void kan_unpack_grad_kernel(const float* __restrict__ dWp, int n_in, int n_out, int out_pad,
                                       float* __restrict__ dspline, float* __restrict__ dlin_w) {
  const long long total = static_cast<long long>(n_in) * n_out * kKW;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx % kKW);
    const int o = static_cast<int>((idx / kKW) % n_out);
    const int i = static_cast<int>(idx / (static_cast<long long>(kKW) * n_out));
    const float v = dWp[(static_cast<size_t>(i) * kKW + k) * out_pad + o];
    if (k < kNB) dspline[(static_cast<size_t>(i) * n_out + o) * kNB + k] += v;
    else dlin_w[static_cast<size_t>(o) * n_in + i] += v;
  }
}

// ------------------------------------------------------------------ forward
// grid (sample tiles of 16*SPT, output tiles of 64); 256 threads; thread (ty, tx) owns samples ty*SPT..+SPT-1,
// outputs tx*4..+3.  Per chunk of 8 inputs: expand activations into sA, copy the packed weight rows
// into sW, then 64 rank-1 updates of the SPT x 4 register tile.  SPT = 4 (64-sample tiles) for large batches,
// SPT = 1 (16-sample tiles) so that an inference batch of ~1000 samples still fills the machine.
template <int SPT>
__global__ void __launch_bounds__(256)
kan_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Wp, const float* __restrict__ bias, Knots kn,
               float* __restrict__ y, int act, int batch, int n_in, int n_out, int in_pad, int out_pad) {
  constexpr int TS = 16 * SPT;
  __shared__ __align__(16) float sA[kKC][TS];    // [packed row][sample]
  __shared__ __align__(16) float sW[kKC][kTO];   // [packed row][output]   16 KB
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int s0 = blockIdx.x * TS, o0 = blockIdx.y * kTO;
  float acc[SPT][4];
#pragma unroll
  for (int a = 0; a < SPT; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;

  for (int i0 = 0; i0 < in_pad; i0 += kIC) {
    // expand: TS samples x 8 inputs per chunk.  SPT = 4: thread handles (sample tid/4, inputs (tid%4)*2, +1);
    // SPT = 1: threads 0..127 handle (sample tid/8, input tid%8)
    {
      constexpr int PER = (SPT == 4) ? 2 : 1;
      const int sl = (SPT == 4) ? (tid >> 2) : (tid >> 3);
      const int sg = s0 + sl;
#pragma unroll
      for (int e = 0; e < PER; ++e) {
        const int il = (SPT == 4) ? ((tid & 3) * 2 + e) : (tid & 7);
        const int ig = i0 + il;
        if (sl < TS) {
          float a[kKW], da[kKW], dt;
          if (sg < batch && ig < n_in) {
            kan_expand<false>(x[static_cast<size_t>(sg) * n_in + ig], kn, a, da, dt);
          } else {
#pragma unroll
            for (int k = 0; k < kKW; ++k) a[k] = 0.0f;
          }
#pragma unroll
          for (int k = 0; k < kKW; ++k) sA[il * kKW + k][sl] = a[k];
        }
      }
    }
    // weights: rows i0*8 .. +64 of Wp, columns o0 .. o0+64 (pads are zero-filled by the packer)
    for (int idx = tid; idx < kKC * (kTO / 4); idx += 256) {
      const int r = idx >> 4, c4 = idx & 15;
      *reinterpret_cast<float4*>(&sW[r][c4 * 4]) =
          *reinterpret_cast<const float4*>(Wp + static_cast<size_t>(i0 * kKW + r) * out_pad + o0 + c4 * 4);
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kKC; ++kk) {
      float a4[SPT];
#pragma unroll
      for (int a = 0; a < SPT; ++a) a4[a] = sA[kk][ty * SPT + a];
      const float4 wv = *reinterpret_cast<const float4*>(&sW[kk][tx * 4]);
      const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int a = 0; a < SPT; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], w4[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < SPT; ++a) {
    const int sg = s0 + ty * SPT + a;
    if (sg >= batch) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int og = o0 + tx * 4 + b;
      if (og >= n_out) continue;
      float v = acc[a][b] + bias[og];
      if (act == 1) v = fmaxf(v, 0.0f);
      else if (act == 2) v = 3.0f / (1.0f + expf(-v));
      y[static_cast<size_t>(sg) * n_out + og] = v;
    }
  }
}

// ------------------------------------------------------------------ backward: weight gradient
// grid (input chunks of 8 -> 64 packed rows, output tiles of 64, batch splits); thread (ty, tx) owns packed rows
// ty*4..+3 and outputs tx*4..+3; reduction over samples in sub-tiles of 64.
__global__ void __launch_bounds__(256)
kan_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ yv, const float* __restrict__ gy, int act,
                 Knots kn, float* __restrict__ dWp, float* __restrict__ dbias, int batch, int n_in, int n_out,
                 int out_pad, int samples_per_split) {
  __shared__ __align__(16) float sA[kTS][kKC];   // [sample][packed row]
  __shared__ __align__(16) float sG[kTS][kTO];   // [sample][output]
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int i0 = blockIdx.x * kIC, o0 = blockIdx.y * kTO;
  const int b_begin = blockIdx.z * samples_per_split;
  const int b_end = min(batch, b_begin + samples_per_split);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
  float bsum = 0.0f;   // thread tid < 64 accumulates dbias[o0 + tid] when blockIdx.x == 0

  for (int s0 = b_begin; s0 < b_end; s0 += kTS) {
    {
      const int sl = tid >> 2, q = tid & 3;
      const int sg = s0 + sl;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int il = q * 2 + e, ig = i0 + il;
        float a[kKW], da[kKW], dt;
        if (sg < b_end && ig < n_in) {
          kan_expand<false>(x[static_cast<size_t>(sg) * n_in + ig], kn, a, da, dt);
        } else {
#pragma unroll
          for (int k = 0; k < kKW; ++k) a[k] = 0.0f;
        }
        *reinterpret_cast<float4*>(&sA[sl][il * kKW]) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(&sA[sl][il * kKW + 4]) = make_float4(a[4], a[5], a[6], a[7]);
      }
    }
    for (int idx = tid; idx < kTS * kTO; idx += 256) {
      const int sl = idx >> 6, ol = idx & 63;
      const int sg = s0 + sl, og = o0 + ol;
      float g = 0.0f;
      if (sg < b_end && og < n_out) {
        const size_t off = static_cast<size_t>(sg) * n_out + og;
        g = gy[off] * act_grad(act, yv[off]);
      }
      sG[sl][ol] = g;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid < kTO) {
      for (int sl = 0; sl < kTS; ++sl) bsum += sG[sl][tid];
    }
#pragma unroll 8
    for (int sl = 0; sl < kTS; ++sl) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[sl][ty * 4]);
      const float4 gv = *reinterpret_cast<const float4*>(&sG[sl][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, g4[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], g4[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      atomicAdd(&dWp[static_cast<size_t>(i0 * kKW + ty * 4 + a) * out_pad + o0 + tx * 4 + b], acc[a][b]);
  if (blockIdx.x == 0 && tid < kTO && o0 + tid < n_out) atomicAdd(&dbias[o0 + tid], bsum);
}

// ------------------------------------------------------------------ backward: input gradient
// grid (sample tiles of 64); per chunk of 16 inputs thread (ty, tx) owns samples ty*4..+3 and input tx;
// G[s, i, 0..7] = sum_o gpre[s,o] * Wp[i*8+k][o], contracted in output sub-tiles of 64.
constexpr int kDxIC = 16;
__global__ void __launch_bounds__(256)
kan_bwd_x_kernel(const float* __restrict__ x, const float* __restrict__ yv, const float* __restrict__ gy, int act,
                 Knots kn, const float* __restrict__ WpT, float* __restrict__ dx, int batch, int n_in, int n_out,
                 int in_pad, int out_pad) {
  __shared__ __align__(16) float sG[kTO][kTS];              // [output][sample]        16 KB
  __shared__ __align__(16) float sWT[kTO][kDxIC * kKW];     // [output][packed row]    32 KB
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int s0 = blockIdx.x * kTS;
  const int ldT = in_pad * kKW;

  {
    const int i0 = blockIdx.y * kDxIC;          // one 16-input chunk per CTA: small batches still fill the machine
    float acc[4][kKW];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int k = 0; k < kKW; ++k) acc[a][k] = 0.0f;
    for (int o0 = 0; o0 < n_out; o0 += kTO) {
      for (int idx = tid; idx < kTO * kTS; idx += 256) {
        const int sl = idx >> 6, ol = idx & 63;       // coalesced over outputs in global memory
        const int sg = s0 + sl, og = o0 + ol;
        float g = 0.0f;
        if (sg < batch && og < n_out) {
          const size_t off = static_cast<size_t>(sg) * n_out + og;
          g = gy[off] * act_grad(act, yv[off]);
        }
        sG[ol][sl] = g;
      }
      for (int idx = tid; idx < kTO * (kDxIC * kKW / 4); idx += 256) {
        const int ol = idx >> 5, c4 = idx & 31;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o0 + ol < out_pad && i0 * kKW + c4 * 4 < ldT)
          w = *reinterpret_cast<const float4*>(WpT + static_cast<size_t>(o0 + ol) * ldT + i0 * kKW + c4 * 4);
        *reinterpret_cast<float4*>(&sWT[ol][c4 * 4]) = w;
      }
      __syncthreads();
#pragma unroll 4
      for (int ol = 0; ol < kTO; ++ol) {
        const float4 gv = *reinterpret_cast<const float4*>(&sG[ol][ty * 4]);
        const float4 w0 = *reinterpret_cast<const float4*>(&sWT[ol][tx * kKW]);
        const float4 w1 = *reinterpret_cast<const float4*>(&sWT[ol][tx * kKW + 4]);
        const float g4[4] = {gv.x, gv.y, gv.z, gv.w};
        const float w8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int k = 0; k < kKW; ++k) acc[a][k] = fmaf(g4[a], w8[k], acc[a][k]);
      }
      __syncthreads();
    }
    const int ig = i0 + tx;
    if (ig < n_in) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int sg = s0 + ty * 4 + a;
        if (sg >= batch) continue;
        float av[kKW], da[kKW], dt;
        kan_expand<true>(x[static_cast<size_t>(sg) * n_in + ig], kn, av, da, dt);
        float sp = 0.0f;
#pragma unroll
        for (int k = 0; k < kNB; ++k) sp = fmaf(acc[a][k], da[k], sp);
        dx[static_cast<size_t>(sg) * n_in + ig] = acc[a][7] + dt * sp;
      }
    }
  }
}

int pad_to(int v, int m) { return (v + m - 1) / m * m; }

constexpr int kKanTcMinBatch = 8192;
// RVK_KAN_SIMT=1 keeps the fp32 CUDA-core kernels for every batch size (A/B measurements)
bool kan_tc_disabled() {
  static const bool off = [] { const char* e = getenv("RVK_KAN_SIMT"); return e != nullptr && e[0] == '1'; }();
  return off;
}
bool kan_use_tc(int batch, int n_in, int n_out) {
  return batch >= kKanTcMinBatch && n_in % 8 == 0 && n_out <= 64 && n_in >= 64 && !kan_tc_disabled();
}

#include "kan_tc.cuh"
#include "kan_small.cuh"
#include "heads_fused.cuh"
#include "heads_train.cuh"

}  // namespace

// workspace floats needed by forward (packed weights) and backward (transposed pack + packed gradient)
int64_t rvk_kan_workspace_floats(int n_in, int n_out, int with_backward) {
  const int64_t in_pad = pad_to(n_in, 16), out_pad = pad_to(n_out, 64);
  const int64_t wp = in_pad * 8 * out_pad;
  // + one wp-sized block for the bf16 hi / lo split of the packed weights (tensor-core path, large batches)
  // (+ 64 floats: interval thresholds of the tensor-core path)
  // (+ with backward: a second split in packed-row-major order, the B operand of the tensor-core dx kernel)
  return (with_backward ? 3 * wp : wp) + wp + 64 + (with_backward ? wp : 0);
}

int rvk_kan_basis_launch(const float* t, const float* knots_host, float* out, int64_t n, cudaStream_t stream) {
  if (n <= 0) return RVK_OK;
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = knots_host[i];
  kan_basis_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(t, kn, out, n);
  return rvk_launch_check();
}

int rvk_kan_layer_fwd_launch(const KanLayerDesc& L, const float* x, float* y, int act, int batch, float* workspace,
                             int with_backward_flags, cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  // bit 0: the backward pass will follow (also prepare its operands); bit 1: this workspace already holds the packed / split
  // weights of these parameters (the caller keys that on the parameter versions), skip the prepare launches
  const int with_backward = with_backward_flags & 1;
  const bool prepared = (with_backward_flags & 2) != 0;
  if (L.num_basis != kNB || L.num_knots != kKnots) return RVK_ERR_UNSUPPORTED_SHAPE;
  const int in_pad = pad_to(L.in_features, 16), out_pad = pad_to(L.out_features, 64);
  const int64_t wp = static_cast<int64_t>(in_pad) * 8 * out_pad;
  float* Wp = workspace;
  float* WpT = with_backward ? workspace + wp : nullptr;
  // algorithmic: dense-7 contraction + linear branch; x read, y written
  RvkScopedTimer timer(stream, 2.0 * batch * L.in_features * 8.0 * L.out_features,
                       4.0 * batch * (L.in_features + L.out_features), RVK_T_KAN_FWD);
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = L.knots_host[i];
  if (kan_small_ok(L.in_features, L.out_features, kan_use_tc(batch, L.in_features, L.out_features)) && !kan_small_disabled())      // few outputs: no packing
    return kan_small_fwd_launch(L, x, y, act, batch, kn, stream);
  const int pack_blocks = static_cast<int>((wp + 255) / 256 < 1184 ? (wp + 255) / 256 : 1184);
  if (!prepared) {
    kan_pack_kernel<<<pack_blocks, 256, 0, stream>>>(L.spline, L.lin_w, L.in_features, L.out_features, in_pad, out_pad, Wp, WpT);
    RVK_TRY(rvk_launch_check());
  }
  if (kan_use_tc(batch, L.in_features, L.out_features)) {
    // tensor-core path: operands split hi + lo in bf16, activations generated on the fly (kan_tc.cuh)
    const int kp = in_pad * 8;
    auto* w_hi = reinterpret_cast<__nv_bfloat16*>(workspace + (with_backward ? 3 : 1) * wp);
    auto* w_lo = w_hi + static_cast<size_t>(64) * kp;
    if (!prepared) {
      kan_split_weights_kernel<<<pack_blocks, 256, 0, stream>>>(L.spline, L.lin_w, L.in_features, L.out_features, kp, w_hi, w_lo);
      RVK_TRY(rvk_launch_check());
    }
    if (with_backward && !prepared) {      // packed-row-major split for the tensor-core dx kernel (after the 64 threshold floats)
      auto* w2_hi = reinterpret_cast<__nv_bfloat16*>(workspace + 4 * wp + 64);
      kan_split_weights_rows_kernel<<<pack_blocks, 256, 0, stream>>>(L.spline, L.lin_w, L.in_features, L.out_features, kp, w2_hi,
                                                                    w2_hi + static_cast<size_t>(64) * kp);
      RVK_TRY(rvk_launch_check());
    }
    CUtensorMap tmWhi, tmWlo;
    RVK_TRY(rvk_make_tmap_2d(&tmWhi, w_hi, RVK_BF16, 64, kp, kp, 64, 64));
    RVK_TRY(rvk_make_tmap_2d(&tmWlo, w_lo, RVK_BF16, 64, kp, kp, 64, 64));
    const int tiles = (batch + 127) / 128;
    const int grid = tiles < kNumSMsB200 ? tiles : kNumSMsB200;
    KanTcTables tb;
    float* xthr = reinterpret_cast<float*>(w_lo + static_cast<size_t>(64) * kp);     // 16 floats after the split weights
    if (!prepared) {
      kan_tc_thresholds_kernel<<<1, 32, 0, stream>>>(kn, xthr);
      RVK_TRY(rvk_launch_check());
    }
    tb.xthr = xthr;
    RVK_SET_MAX_SMEM(kan_fwd_tc_kernel, kTcSmemBytes);
    kan_fwd_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, stream>>>(tmWhi, tmWlo, x, L.lin_b, tb, y, act, batch, L.in_features,
                                                                 L.out_features, L.in_features / 8);
    return rvk_launch_check();
  }
  if (batch <= 4096) {
    dim3 grid((batch + 15) / 16, out_pad / kTO);
    kan_fwd_kernel<1><<<grid, 256, 0, stream>>>(x, Wp, L.lin_b, kn, y, act, batch, L.in_features, L.out_features, in_pad, out_pad);
  } else {
    dim3 grid((batch + kTS - 1) / kTS, out_pad / kTO);
    kan_fwd_kernel<4><<<grid, 256, 0, stream>>>(x, Wp, L.lin_b, kn, y, act, batch, L.in_features, L.out_features, in_pad, out_pad);
  }
  return rvk_launch_check();
}

int rvk_kan_layer_bwd_launch(const KanLayerDesc& L, const float* x, const float* y, const float* gy, int act,
                             float* dx, float* dspline, float* dlin_w, float* dlin_b, int batch, float* workspace,
                             cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  if (L.num_basis != kNB || L.num_knots != kKnots) return RVK_ERR_UNSUPPORTED_SHAPE;
  const int in_pad = pad_to(L.in_features, 16), out_pad = pad_to(L.out_features, 64);
  const int64_t wp = static_cast<int64_t>(in_pad) * 8 * out_pad;
  const float* WpT = workspace + wp;      // written by the forward launch
  float* dWp = workspace + 2 * wp;
  // algorithmic: dW and dx contractions; x, y, gy read, dx written
  RvkScopedTimer timer(stream, 2.0 * batch * L.in_features * 8.0 * L.out_features * ((dspline ? 1 : 0) + (dx ? 1 : 0)),
                       4.0 * batch * (L.in_features * (dx ? 2.0 : 1.0) + 2.0 * L.out_features), RVK_T_KAN_BWD);
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = L.knots_host[i];
  if (kan_small_ok(L.in_features, L.out_features, kan_use_tc(batch, L.in_features, L.out_features)) && !kan_small_disabled())      // one kernel: dx, dW, dWl, db (+=)
    return kan_small_bwd_launch(L, x, y, gy, act, dx, dspline, dlin_w, dlin_b, batch, kn, stream);
  if (dspline != nullptr) {
    RVK_CUDA_TRY(cudaMemsetAsync(dWp, 0, wp * sizeof(float), stream));
    const int tiles = (in_pad / kIC) * (out_pad / kTO);
    int splits = (2 * kNumSMsB200 + tiles - 1) / tiles;
    const int sub_tiles = (batch + kTS - 1) / kTS;
    if (splits > sub_tiles) splits = sub_tiles;
    const int sps = ((sub_tiles + splits - 1) / splits) * kTS;
    splits = (batch + sps - 1) / sps;
    if (kan_use_tc(batch, L.in_features, L.out_features) && L.in_features % 64 == 0) {
      // tensor-core weight gradient (kan_tc.cuh): grid = (batch slices) x (groups of 64 inputs)
      KanTcTables tb;
      tb.xthr = workspace + 4 * wp;          // written by the forward launch
      const int groups = L.in_features / 64;
      const int tiles128 = (batch + 127) / 128;
      int slices = kNumSMsB200 / groups;
      if (slices > tiles128) slices = tiles128;
      RVK_SET_MAX_SMEM(kan_bwd_w_tc_kernel, kTcWgSmemBytes);
      kan_bwd_w_tc_kernel<<<dim3(slices, groups), kTcThreads, kTcWgSmemBytes, stream>>>(x, y, gy, tb, dWp, dlin_b, act, batch,
                                                                                        L.in_features, L.out_features);
    } else {
      dim3 grid(in_pad / kIC, out_pad / kTO, splits);
      kan_bwd_w_kernel<<<grid, 256, 0, stream>>>(x, y, gy, act, kn, dWp, dlin_b, batch, L.in_features, L.out_features,
                                                 out_pad, sps);
    }
    RVK_TRY(rvk_launch_check());
    const int64_t tot = static_cast<int64_t>(L.in_features) * L.out_features * 8;
    const int ub = static_cast<int>((tot + 255) / 256 < 1184 ? (tot + 255) / 256 : 1184);
    kan_unpack_grad_kernel<<<ub, 256, 0, stream>>>(dWp, L.in_features, L.out_features, out_pad, dspline, dlin_w);
    RVK_TRY(rvk_launch_check());
  }
  if (dx != nullptr && kan_use_tc(batch, L.in_features, L.out_features)) {
    // tensor-core dx (kan_tc.cuh); split weights and thresholds were written by the forward launch into the workspace
    const int kp = in_pad * 8;
    auto* w2_hi = reinterpret_cast<__nv_bfloat16*>(workspace + 4 * wp + 64);
    auto* w2_lo = w2_hi + static_cast<size_t>(64) * kp;
    CUtensorMap tmWhi, tmWlo;
    RVK_TRY(rvk_make_tmap_2d(&tmWhi, w2_hi, RVK_BF16, kp, 64, 64, 64, 64));
    RVK_TRY(rvk_make_tmap_2d(&tmWlo, w2_lo, RVK_BF16, kp, 64, 64, 64, 64));
    KanTcTables tb;
    tb.xthr = workspace + 4 * wp;            // = the 16 floats after the forward's split weights
    const int tiles = (batch + 127) / 128;
    const int grid = tiles < kNumSMsB200 ? tiles : kNumSMsB200;
    RVK_SET_MAX_SMEM(kan_bwd_x_tc_kernel, kTcBxSmemBytes);
    kan_bwd_x_tc_kernel<<<grid, kTcThreads, kTcBxSmemBytes, stream>>>(tmWhi, tmWlo, x, y, gy, tb, dx, act, batch, L.in_features,
                                                                     L.out_features, L.in_features / 8);
    RVK_TRY(rvk_launch_check());
  } else if (dx != nullptr) {
    kan_bwd_x_kernel<<<dim3((batch + kTS - 1) / kTS, (L.in_features + kDxIC - 1) / kDxIC), 256, 0, stream>>>(
        x, y, gy, act, kn, WpT, dx, batch, L.in_features, L.out_features, in_pad, out_pad);
    RVK_TRY(rvk_launch_check());
  }
  return RVK_OK;
}

// ---- fused inference tail (heads_fused.cuh) ---------------------------------------------------------------
int64_t rvk_heads_fused_workspace_floats_impl() { return kHfWsFloats; }

int rvk_heads_fused_prepare_launch(const void* const* p23, float* ws, cudaStream_t stream) {
  if (p23 == nullptr || ws == nullptr) return RVK_ERR_BAD_ARG;
  for (int i = 0; i < 23; ++i)
    if (p23[i] == nullptr) return RVK_ERR_BAD_ARG;
  HeadsFusedParams p;
  auto F = [&](int i) { return static_cast<const float*>(p23[i]); };
  // order: cls.fc1.{w,b}, cls.fc2.{w,b}, ord.fc1.{w,b}, ord.fc2.{w,b}, unc.fc1.{w,b}, unc.fc_mu.{w,b}, unc.fc_logvar.{w,b},
  //        kan layer l: spline, lin_w, lin_b  (l = 0, 1, 2)
  p.fc1_w[0] = F(0); p.fc1_b[0] = F(1); p.fc2_w[0] = F(2); p.fc2_b[0] = F(3);
  p.fc1_w[1] = F(4); p.fc1_b[1] = F(5); p.fc2_w[1] = F(6); p.fc2_b[1] = F(7);
  p.fc1_w[2] = F(8); p.fc1_b[2] = F(9); p.fc2_w[2] = F(10); p.fc2_b[2] = F(11); p.fc2_w[3] = F(12); p.fc2_b[3] = F(13);
  for (int l = 0; l < 3; ++l) { p.spline[l] = F(14 + 3 * l); p.lin_w[l] = F(15 + 3 * l); p.lin_b[l] = F(16 + 3 * l); }
  heads_fused_pack_kernel<<<(kHfWsFloats + 255) / 256, 256, 0, stream>>>(p, ws);
  return rvk_launch_check();
}

int rvk_heads_fused_launch(const float* features, const float* ws, const float* knots_host, int batch, float* cls,
                           float* ord, float* mu, float* log_var, float* kan, cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  if (features == nullptr || ws == nullptr || knots_host == nullptr || cls == nullptr || ord == nullptr || mu == nullptr ||
      log_var == nullptr || kan == nullptr)
    return RVK_ERR_BAD_ARG;
  RVK_SET_MAX_SMEM(heads_fused_kernel<false>, kHfSmemBytes);
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = knots_host[i];
  RvkScopedTimer timer(stream, 2.0 * batch * (3.0 * 192 * 128 + 128.0 * 9 + 8.0 * (192 * 64 + 64 * 16 + 16)), 4.0 * batch * (192 + 10),
                       RVK_T_HEADS_FUSED);
  heads_fused_kernel<false><<<(batch + kHfS - 1) / kHfS, kHfThreads, kHfSmemBytes, stream>>>(features, ws, kn, batch, cls, ord, mu,
                                                                                            log_var, kan, HeadsTrainSave{});
  return rvk_launch_check();
}

// ---- fused multi-task tail of the TRAINING step (heads_fused.cuh<true>, heads_train.cuh) -----------------------
int rvk_heads_train_fwd_launch(const float* features, const float* ws, const float* knots_host, int batch, float drop_p,
                               unsigned long long seed, unsigned long long offset, float* cls, float* ord, float* mu,
                               float* log_var, float* kan, float* h_save, float* a1_save, float* a2_save, cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  if (features == nullptr || ws == nullptr || knots_host == nullptr || cls == nullptr || ord == nullptr || mu == nullptr ||
      log_var == nullptr || kan == nullptr || h_save == nullptr || a1_save == nullptr || a2_save == nullptr || drop_p < 0.0f ||
      drop_p >= 1.0f)
    return RVK_ERR_BAD_ARG;
  RVK_SET_MAX_SMEM(heads_fused_kernel<true>, kHfSmemBytes);
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = knots_host[i];
  HeadsTrainSave sv{h_save, a1_save, a2_save, drop_p, seed, offset};
  RvkScopedTimer timer(stream, 2.0 * batch * (3.0 * 192 * 128 + 128.0 * 9 + 8.0 * (192 * 64 + 64 * 16 + 16)),
                       4.0 * batch * (192 + 10 + 384 + 80), RVK_T_HEADS_TRAIN);
  heads_fused_kernel<true><<<(batch + kHfS - 1) / kHfS, kHfThreads, kHfSmemBytes, stream>>>(features, ws, kn, batch, cls, ord, mu,
                                                                                           log_var, kan, sv);
  return rvk_launch_check();
}

int rvk_heads_train_bwd_launch(const float* features, const float* ws, const float* knots_host, int batch, float drop_p,
                               const float* h_save, const float* a1_save, const float* a2_save, const float* lv_out,
                               const float* kan_out, const float* d_cls, const float* d_ord, const float* d_mu, const float* d_lv,
                               const float* d_kan, float* dfeat, float* dws, float* const* grads23_host, cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  if (features == nullptr || ws == nullptr || knots_host == nullptr || h_save == nullptr || a1_save == nullptr ||
      a2_save == nullptr || lv_out == nullptr || kan_out == nullptr || dfeat == nullptr || dws == nullptr || grads23_host == nullptr)
    return RVK_ERR_BAD_ARG;
  RVK_SET_MAX_SMEM(heads_train_bwd_kernel, kHtSmemBytes);
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = knots_host[i];
  RvkScopedTimer timer(stream, 4.0 * batch * (3.0 * 192 * 128 + 128.0 * 9 + 8.0 * (192 * 64 + 64 * 16 + 16)),
                       4.0 * batch * (2 * 192 + 10 + 384 + 80) + 8.0 * kHfWsFloats, RVK_T_HEADS_TRAIN);
  RVK_CUDA_TRY(cudaMemsetAsync(dws, 0, sizeof(float) * kHfWsFloats, stream));
  HeadsTrainBwdArgs a{features, ws, h_save, a1_save, a2_save, lv_out, kan_out, d_cls, d_ord, d_mu, d_lv, d_kan, dfeat, dws, drop_p, batch};
  heads_train_bwd_kernel<<<(batch + kHfS - 1) / kHfS, kHtThreads, kHtSmemBytes, stream>>>(a, kn);
  RVK_TRY(rvk_launch_check());
  HeadsGradPtrs g;
  auto G = [&](int i) { return grads23_host[i]; };
  g.fc1_w[0] = G(0); g.fc1_b[0] = G(1); g.fc2_w[0] = G(2); g.fc2_b[0] = G(3);
  g.fc1_w[1] = G(4); g.fc1_b[1] = G(5); g.fc2_w[1] = G(6); g.fc2_b[1] = G(7);
  g.fc1_w[2] = G(8); g.fc1_b[2] = G(9); g.fc2_w[2] = G(10); g.fc2_b[2] = G(11); g.fc2_w[3] = G(12); g.fc2_b[3] = G(13);
  for (int l = 0; l < 3; ++l) { g.spline[l] = G(14 + 3 * l); g.lin_w[l] = G(15 + 3 * l); g.lin_b[l] = G(16 + 3 * l); }
  heads_fused_unpack_grad_kernel<<<(kHfWsFloats + 255) / 256, 256, 0, stream>>>(dws, g);
  return rvk_launch_check();
}
