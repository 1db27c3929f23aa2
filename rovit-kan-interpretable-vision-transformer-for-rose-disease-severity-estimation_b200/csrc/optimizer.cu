// Fused optimizer tail of the train step: what the reference's trainer does after backward --
// GradScaler.unscale_ + inf check, clip_grad_norm_(max_norm), AdamW with two learning-rate groups
// (training/trainer.py:118-129, training/optimizer.py:18-25) -- as three launches over all parameter tensors
// instead of ~150 per-tensor kernels:
//
//   1. optim_norm_kernel    sum of squares of every gradient (one atomicAdd per block into state[0]); a non-finite
//                           gradient makes the sum non-finite, which IS the inf check
//   2. optim_update_kernel  g' = g * grad_mult * clip,  clip = min(1, max_norm / (||g * grad_mult|| + 1e-6));
//                           AdamW update of p, exp_avg, exp_avg_sq (torch's fused-kernel formulas); skipped entirely when
//                           the norm is non-finite or *found_inf != 0 (what GradScaler.step does)
//   3. optim_finish_kernel  step += 1 if the update ran; state[0] = 0 for the next call; state[2] keeps the last norm
//
// Tensors are addressed through a table passed as a kernel parameter (pointers change every step: autograd hands out
// a fresh gradient buffer).  The moment buffers live in a PADDED flat index space in which every tensor starts at a
// multiple of kChunk, so a block finds its tensor with one binary search and never straddles two tensors.
#include "kernels.h"

namespace {

constexpr int kChunk = 4096;        // elements per block
constexpr int kOptThreads = 256;
constexpr int kMaxTensors = 128;     // table = 3.2 KB: stays inside the classic 4 KB kernel-parameter space

struct OptTable {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  int chunk_start[kMaxTensors + 1];   // first chunk of tensor t in the padded space of THIS launch
  int numel[kMaxTensors];
  unsigned char group[kMaxTensors];
  int n;
};

struct OptHyper {
  float lr[4];
  float decay[4];                                    // 1 - lr * weight_decay, evaluated in double on the host as torch does
  float beta1, beta2, eps, weight_decay, max_norm;   // max_norm <= 0: no clipping
  float om_beta1, om_beta2;                          // 1 - beta, rounded from double (torch passes them as Python floats)
  double lr_d[4], beta1_d, beta2_d;                  // bias corrections and step size are evaluated in double
  float grad_mult;                                   // e.g. 1 / world_size
  const float* grad_scale;                           // device, nullable: gradients are divided by *grad_scale
  const float* found_inf;                            // device, nullable: != 0 skips the update
};

__device__ __forceinline__ int find_tensor(const OptTable& T, int chunk) {
  int lo = 0, hi = T.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (T.chunk_start[mid] <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kOptThreads)
optim_norm_kernel(const __grid_constant__ OptTable T, float* __restrict__ state) {
  __shared__ int s_t;
  __shared__ float s_part[kOptThreads / 32];
  if (threadIdx.x == 0) s_t = find_tensor(T, blockIdx.x);
  __syncthreads();
  const int t = s_t;
  const float* g = T.g[t];
  const int base = (static_cast<int>(blockIdx.x) - T.chunk_start[t]) * kChunk;
  const int n = T.numel[t];
  float acc = 0.0f;
  if (g != nullptr) {
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      for (int i = base + threadIdx.x * 4; i < base + kChunk && i < n; i += kOptThreads * 4) {
        if (i + 3 < n) {
          const float4 v = *reinterpret_cast<const float4*>(g + i);
          acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
        } else {
          for (int j = i; j < n; ++j) acc = fmaf(g[j], g[j], acc);
        }
      }
    } else {
      for (int i = base + threadIdx.x; i < base + kChunk && i < n; i += kOptThreads) acc = fmaf(g[i], g[i], acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.0f;
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w) tot += s_part[w];
    atomicAdd(&state[0], tot);
  }
}

// state: [0] running sum of squares (this call), [1] step count (as float, exact up to 2^24), [2] last total norm,
//        [3] 1 if the last update ran, 0 if it was skipped
__global__ void __launch_bounds__(kOptThreads)
optim_update_kernel(const __grid_constant__ OptTable T, const OptHyper H, float* __restrict__ exp_avg,
                    float* __restrict__ exp_avg_sq, const float* __restrict__ state, int chunk_offset) {
  __shared__ int s_t;
  if (threadIdx.x == 0) s_t = find_tensor(T, blockIdx.x);
  __syncthreads();
  const int t = s_t;
  const float* g = T.g[t];
  if (g == nullptr) return;                                   // parameter without gradient: torch skips it too
  const float inv_scale = (H.grad_scale != nullptr) ? 1.0f / *H.grad_scale : 1.0f;
  const float mult = H.grad_mult * inv_scale;
  const float norm = sqrtf(state[0]) * fabsf(mult);
  if (!isfinite(norm) || (H.found_inf != nullptr && *H.found_inf != 0.0f)) return;
  float clip = 1.0f;
  if (H.max_norm > 0.0f) clip = fminf(1.0f, H.max_norm / (norm + 1e-6f));      // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max=1)
  const float gm = mult * clip;
  const float step = state[1] + 1.0f;
  // torch/aten fused_adam_utils.cuh (adamw): bias corrections in double, the rest in fp32
  const double bc1 = 1.0 - pow(H.beta1_d, static_cast<double>(step));
  const double bc2 = 1.0 - pow(H.beta2_d, static_cast<double>(step));
  const float step_size = static_cast<float>(H.lr_d[T.group[t]] / bc1);
  const float bc2_sqrt = static_cast<float>(sqrt(bc2));
  const float decay = H.decay[T.group[t]];
  float* p = T.p[t];
  const int base = (static_cast<int>(blockIdx.x) - T.chunk_start[t]) * kChunk;
  const int n = T.numel[t];
  const size_t sbase = (static_cast<size_t>(chunk_offset) + blockIdx.x) * kChunk - base;   // padded index of element 0 of the tensor
  for (int i = base + threadIdx.x; i < base + kChunk && i < n; i += kOptThreads) {
    const float gr = g[i] * gm;
    float w = p[i] * decay;
    float m = exp_avg[sbase + i];
    float v = exp_avg_sq[sbase + i];
    m = m + H.om_beta1 * (gr - m);                            // lerp(exp_avg, grad, 1 - beta1)
    v = H.beta2 * v + H.om_beta2 * gr * gr;
    const float denom = sqrtf(v) / bc2_sqrt + H.eps;
    w -= step_size * m / denom;
    p[i] = w;
    exp_avg[sbase + i] = m;
    exp_avg_sq[sbase + i] = v;
  }
}

__global__ void optim_finish_kernel(const OptHyper H, float* __restrict__ state) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float inv_scale = (H.grad_scale != nullptr) ? 1.0f / *H.grad_scale : 1.0f;
  const float norm = sqrtf(state[0]) * fabsf(H.grad_mult * inv_scale);
  const bool ran = isfinite(norm) && !(H.found_inf != nullptr && *H.found_inf != 0.0f);
  if (ran) state[1] += 1.0f;
  state[2] = norm;
  state[3] = ran ? 1.0f : 0.0f;
  state[0] = 0.0f;
}

int chunks_of(int64_t numel) { return static_cast<int>((numel + kChunk - 1) / kChunk); }

}  // namespace

int64_t rvk_optimizer_state_floats_impl(int n, const int64_t* numel_host) {
  int64_t chunks = 0;
  for (int i = 0; i < n; ++i) chunks += chunks_of(numel_host[i] > 0 ? numel_host[i] : 0);
  return chunks * kChunk;
}

int rvk_optimizer_step_impl(int n, void* const* params_host, const void* const* grads_host, const int64_t* numel_host,
                            const int* group_host, float* exp_avg, float* exp_avg_sq, float* state4, const double* lr_host,
                            int n_groups, double beta1, double beta2, double eps, double weight_decay, float max_grad_norm,
                            float grad_mult, const float* grad_scale_dev, const float* found_inf_dev, cudaStream_t stream) {
  if (n <= 0) return RVK_OK;
  if (params_host == nullptr || grads_host == nullptr || numel_host == nullptr || group_host == nullptr || exp_avg == nullptr ||
      exp_avg_sq == nullptr || state4 == nullptr || lr_host == nullptr || n_groups < 1 || n_groups > 4)
    return RVK_ERR_BAD_ARG;
  OptHyper H{};
  for (int i = 0; i < n_groups; ++i) {
    H.lr[i] = static_cast<float>(lr_host[i]);
    H.lr_d[i] = lr_host[i];
    H.decay[i] = static_cast<float>(1.0 - lr_host[i] * weight_decay);
  }
  H.om_beta1 = static_cast<float>(1.0 - beta1);
  H.om_beta2 = static_cast<float>(1.0 - beta2);
  H.beta1_d = beta1; H.beta2_d = beta2;
  H.beta1 = static_cast<float>(beta1); H.beta2 = static_cast<float>(beta2); H.eps = static_cast<float>(eps);
  H.weight_decay = static_cast<float>(weight_decay); H.max_norm = max_grad_norm;
  H.grad_mult = grad_mult; H.grad_scale = grad_scale_dev; H.found_inf = found_inf_dev;
  double bytes = 0.0;
  for (int i = 0; i < n; ++i)
    if (grads_host[i] != nullptr) bytes += 4.0 * numel_host[i] * 8.0;      // norm: g; update: g, p, m, v read + p, m, v written
  RvkScopedTimer timer(stream, 0.0, bytes, RVK_T_OPTIMIZER);
  // phase 1 over all tensors (table by table), then phase 2: the clip coefficient needs the global norm
  for (int phase = 0; phase < 2; ++phase) {
    int chunk_offset = 0;
    for (int t0 = 0; t0 < n; t0 += kMaxTensors) {
      const int cnt = (n - t0 < kMaxTensors) ? n - t0 : kMaxTensors;
      OptTable T{};
      T.n = cnt;
      int chunks = 0;
      for (int i = 0; i < cnt; ++i) {
        if (numel_host[t0 + i] <= 0 || numel_host[t0 + i] > (1LL << 30) || params_host[t0 + i] == nullptr ||
            group_host[t0 + i] < 0 || group_host[t0 + i] >= n_groups)
          return RVK_ERR_BAD_ARG;
        T.p[i] = static_cast<float*>(params_host[t0 + i]);
        T.g[i] = static_cast<const float*>(grads_host[t0 + i]);
        T.chunk_start[i] = chunks;
        T.numel[i] = static_cast<int>(numel_host[t0 + i]);
        T.group[i] = static_cast<unsigned char>(group_host[t0 + i]);
        chunks += chunks_of(numel_host[t0 + i]);
      }
      T.chunk_start[cnt] = chunks;
      if (phase == 0) optim_norm_kernel<<<chunks, kOptThreads, 0, stream>>>(T, state4);
      else optim_update_kernel<<<chunks, kOptThreads, 0, stream>>>(T, H, exp_avg, exp_avg_sq, state4, chunk_offset);
      RVK_TRY(rvk_launch_check());
      chunk_offset += chunks;
    }
  }
  optim_finish_kernel<<<1, 32, 0, stream>>>(H, state4);
  return rvk_launch_check();
}
