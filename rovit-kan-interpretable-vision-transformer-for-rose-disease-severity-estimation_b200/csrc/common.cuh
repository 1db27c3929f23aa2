// Shared device helpers for the sm_100a kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (UMMA + TMEM) wrappers written as inline PTX, plus small math utilities.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor"
// tables (same fields CUTLASS names SmemDescriptor / InstrDescriptor).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// ----------------------------------------------------------------------------- status codes
enum RvkStatus : int {
  RVK_OK = 0,
  RVK_ERR_BAD_ARG = 1,
  RVK_ERR_CUDA = 2,
  RVK_ERR_UNSUPPORTED_SHAPE = 3,
  RVK_ERR_TMA_ENCODE = 4,
  RVK_ERR_NO_DRIVER = 5,
  RVK_ERR_WORKSPACE = 6,
  RVK_ERR_ALIGNMENT = 7,
};

#define RVK_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) {                                 \
      rvk_set_last_cuda_error((int)_e, #expr);               \
      return RVK_ERR_CUDA;                                   \
    }                                                        \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: do it once per device and kernel, not once per process
// (a process may drive more than one GPU even though the benchmark uses one process per GPU)
#define RVK_SET_MAX_SMEM(kernel, bytes)                                                                        \
  do {                                                                                                         \
    static bool rvk_done_[64] = {};                                                                            \
    int rvk_dev_ = 0;                                                                                          \
    RVK_CUDA_TRY(cudaGetDevice(&rvk_dev_));                                                                    \
    if (rvk_dev_ < 0 || rvk_dev_ >= 64 || !rvk_done_[rvk_dev_]) {                                              \
      RVK_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));          \
      if (rvk_dev_ >= 0 && rvk_dev_ < 64) rvk_done_[rvk_dev_] = true;                                          \
    }                                                                                                          \
  } while (0)

#define RVK_TRY(expr)                 \
  do {                                \
    int _s = (expr);                  \
    if (_s != RVK_OK) return _s;      \
  } while (0)

void rvk_set_last_cuda_error(int code, const char* what);
void rvk_count_launch();   // every kernel launch of this library is counted (bench.py reports it)

// ---- optional in-situ timing of kernel launches (bench.py roofline): CUDA events on the launch stream around one launch
enum RvkTimedKind {
  RVK_T_GEMM_NT = 0, RVK_T_GEMM_TN = 1, RVK_T_MLP_FUSED = 2, RVK_T_ATTN_FWD = 3, RVK_T_ATTN_BWD = 4, RVK_T_LN_BWD = 5,
  RVK_T_KAN_FWD = 6, RVK_T_KAN_BWD = 7, RVK_T_HEADS_FUSED = 8, RVK_T_IM2COL = 9, RVK_T_OPTIMIZER = 10, RVK_T_HEADS_TRAIN = 11,
  RVK_T_COUNT = 12
};
void* rvk_timer_begin(cudaStream_t s, double flops, double bytes, int kind);   // nullptr when timing is off
void rvk_timer_end(void* handle, cudaStream_t s);
struct RvkScopedTimer {
  cudaStream_t s;
  void* h;
  RvkScopedTimer(cudaStream_t stream, double flops, double bytes, int kind) : s(stream), h(rvk_timer_begin(stream, flops, bytes, kind)) {}
  ~RvkScopedTimer() { if (h != nullptr) rvk_timer_end(h, s); }
  RvkScopedTimer(const RvkScopedTimer&) = delete;
  RvkScopedTimer& operator=(const RvkScopedTimer&) = delete;
};

static inline int rvk_launch_check() {
  rvk_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    rvk_set_last_cuda_error((int)e, "kernel launch");
    return RVK_ERR_CUDA;
  }
  return RVK_OK;
}

constexpr int kNumSMsB200 = 148;

// ----------------------------------------------------------------------------- small device utils
#ifdef __CUDACC__

// ------------------------------------------------------------------ Philox4x32-10 (counter-based RNG)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long offset, unsigned long long idx) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32),
                                           static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32)),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  return (r.x >> 8) * (1.0f / 16777216.0f);
}


__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}

// Exact-form GELU (timm nn.GELU default, approximate='none') with ONE MUFU op and no division:
//     gelu(x) = x*Phi(x) = relu(x) - |x| * Phi(-|x|),     Phi(-a) = 2^P(a),
// P = degree-6 minimax fit of log2(Phi(-a)) on [0, 5.5] (max |dP| = 3.6e-5, i.e. 2.5e-5 relative in Phi;
// |x| beyond 5.5 is clamped: the absolute error there is < 3e-7).  Resulting gelu error: <= 4e-6 absolute,
// <= 2.7e-5 relative for |x| < 5.5 -- two orders of magnitude below the bf16 rounding (2^-9) of the stored
// result.  libdevice erff costs ~40 instructions and the rational A&S form two MUFU ops per element, which
// made the XU pipe the bound of the fc1 GEMM (ncu: profiles/).
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// Phi(-a) for a >= 0
__device__ __forceinline__ float norm_cdf_neg(float a) {
  const float t = fminf(a, 5.5f);
  float p = fmaf(2.615375161e-05f, t, -6.609812262e-04f);
  p = fmaf(p, t, 7.488309406e-03f);
  p = fmaf(p, t, -5.197040364e-02f);
  p = fmaf(p, t, -4.603294730e-01f);
  p = fmaf(p, t, -1.150583982e+00f);
  p = fmaf(p, t, -1.000036120e+00f);
  return ex2_approx(p);
}
__device__ __forceinline__ float gelu_erf(float x) {
  return fmaf(-fabsf(x), norm_cdf_neg(fabsf(x)), fmaxf(x, 0.0f));
}
// d/dx gelu = Phi(x) + x*phi(x).  With a = |x| and D(a) = Phi(-a) - a phi(a):  gelu'(x) = x >= 0 ? 1 - D : D, and
// D(a) = 2^(-a^2 log2(e) / 2) * R(a), R = (Mills ratio - a) / sqrt(2 pi) as a degree-6 polynomial on [0, 5.5] (weighted minimax
// fit, tools/fit_gelu_grad.py: max |error| of gelu' 1.6e-5 -- two orders below the bf16 rounding of the stored product; beyond
// 5.5, |D| < 6e-7).  ONE exponential and 14 instructions per element; the two-exponential form Phi + x phi (21 instructions,
// two MUFU) made the fc2 dgrad GEMM's epilogue its bound (56 us against 42 us for the forward GELU GEMM of the same traffic).
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float t = fminf(fabsf(x), 5.5f);
  float r = fmaf(7.042132202e-04f, t, -8.041755296e-03f);
  r = fmaf(r, t, 3.967186809e-02f);
  r = fmaf(r, t, -1.169406101e-01f);
  r = fmaf(r, t, 2.444232404e-01f);
  r = fmaf(r, t, -7.971517444e-01f);
  r = fmaf(r, t, 4.999842942e-01f);
  const float d = ex2_approx(t * -0.72134752044448170368f * t) * r;
  return (x >= 0.0f) ? 1.0f - d : d;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or the hint
// (ns) expires, so waiting warps do not burn issue slots that the epilogue warps need
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must end in a trap (sticky CUDA error reported to the
// caller), never in a hung GPU.  The clock is consulted only every 1024 failed probes.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  for (uint32_t spins = 1;; ++spins) {
    if (mbar_try_wait(bar, parity)) return;
    if ((spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) {   // ~4 s
        printf("rvk: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
        __trap();
      }
    }
  }
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// generic-proxy writes to smem -> visible to the async proxy (TMA store / UMMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------- programmatic dependent launch
// Kernels launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while their predecessor in the stream is
// still draining: everything before griddep_wait() (barrier init, TMEM allocation, descriptor prefetch, parameter staging)
// overlaps the predecessor's tail; no thread may touch data the predecessor produces before it.  Without the attribute the
// instructions are no-ops.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// bring a box into L2 ahead of the load that will need it
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// the box at (c0, c1) -> L2 only (no shared memory, no barrier): takes HBM latency off a later tma_load_2d of the same box
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
      "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
      "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
      "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])),
      "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
      "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])),
      "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])),
      "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (64-bit).  Fields (PTX ISA, tcgen05 smem descriptor):
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset>>4 [46,48) version = 1 on sm_100
//   [49,52) base offset (0: tile bases are 1024 B aligned)
//   [61,64) swizzle: 0 none, 1 128B(32B atom), 2 128B, 4 64B, 6 32B
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}

// Instruction descriptor (32-bit) for kind::f16 with bf16 A/B and fp32 D.
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format (1 = bf16)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a [rows x 128 B] panel laid out
// with the TMA/UMMA 128-byte swizzle (panel base 1024 B aligned)
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}


// ----------------------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster of 2 share one UMMA: the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 with
// M = 256 (each CTA owns 128 accumulator rows in its own TMEM) and every CTA stages its own A rows and HALF of
// the B rows in its own shared memory, at the same offsets.  Barriers the leader's MMA thread waits on live in
// the leader's shared memory; the peer reaches them through mapa + cluster-scope arrives / TMA complete_tx.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) inside CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
  return r;
}
// Arrive on a barrier of the pair's leader.  Default semantics (release at CTA scope): a cluster-scope release
// compiles to MEMBAR.ALL.GPU + ERRBAR and waits for every outstanding global access of the warp (measured: 20% of
// the fused MLP kernel's stall samples).  What the leader's UMMA needs -- this warp's shared-memory writes visible
// to the async proxy -- is established by fence.proxy.async.shared::cta BEFORE the arrive, as CUTLASS' 2-SM
// pipelines do.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
// bounded wait whose acquire covers arrivals made by the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  long long t0 = 0;
  for (uint32_t spins = 1;; ++spins) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    if ((spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) {
        printf("rvk: cluster mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
        __trap();
      }
    }
  }
}
// generic-proxy writes (any state space) -> async proxy, for smem that the PEER's UMMA will read
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// TMA load into this CTA's smem whose bytes complete on the LEADER's mbarrier (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Descriptor halves for a K-major, 128-byte-swizzled operand tile (LBO 16 B, SBO 1024 B): the high word is constant
// and the low word is (smem address >> 4) | LBO field, so stepping K by 16 bf16 (32 bytes) is `lo + 2`.  Splitting the
// descriptor keeps the UMMA issue loop at a few uniform-datapath instructions per MMA.
constexpr uint32_t kUmmaDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes = 16) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
template <int G>
__device__ __forceinline__ void umma_f16_split(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  if (G == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHiSw128)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHiSw128)
        : "memory");
  }
}

// arrives on the barrier at this smem offset in BOTH CTAs of the pair once all prior UMMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}

// ----------------------------------------------------------------------------- tiled fp32 token stream
// The fused inference kernels keep the fp32 residual stream x[M,192] in a register-friendly tiling so that an
// epilogue warp (one accumulator row per lane) reads and writes it with fully coalesced 512-byte requests and
// no shared-memory staging:   element (row r, col c) lives at
//     (((r/32)*6 + c/32)*8 + (c%32)/4)*128 + (r%32)*4 + c%4          (floats)
// i.e. [32-row block][32-col panel][float4 index j][lane][4].  Rows are padded to a multiple of 128.
__host__ __device__ __forceinline__ size_t xt_offset(int row, int panel, int j) {
  return ((static_cast<size_t>(row >> 5) * 6 + panel) * 8 + j) * 128 + static_cast<size_t>(row & 31) * 4;
}
__host__ __device__ __forceinline__ size_t xt_elem_offset(int row, int col) {
  return xt_offset(row, col >> 5, (col & 31) >> 2) + (col & 3);
}

#endif  // __CUDACC__
