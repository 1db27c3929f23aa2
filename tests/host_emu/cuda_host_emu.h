// Thread-level CUDA emulation for the CPU test-suite (TEST INFRASTRUCTURE ONLY, used by tests/test_kernels_on_host.py).
//
// A kernel cut verbatim out of a .cu file is compiled by g++ against this header and run block by block, every CUDA thread of a
// block as one FIBER (ucontext) of a single OS thread, scheduled round-robin: a fiber runs until it reaches a barrier and then
// hands the core to the next one.  `__shared__` variables are statics of the (single) running block, `__syncthreads()` is a
// barrier over the live threads of the block and the warp shuffles exchange through a per-warp barrier, so kernels that reduce
// with `__shfl_xor_sync`, stage through shared memory and finish with `atomicAdd` run with their real control flow -- and
// deterministically.  Threads that return early are dropped from the barriers (a CUDA block does not wait for exited threads
// either).  Tensor cores, TMA and mbarriers are modelled functionally by tcgen05_host_emu.h (single-CTA kernels); inline PTX is not.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <ucontext.h>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

using std::max;
using std::min;
using std::isfinite;

struct EmuDim { unsigned x = 1, y = 1, z = 1; };
struct EmuIdx { unsigned x = 0, y = 0, z = 0; };
static EmuIdx threadIdx, blockIdx;          // of the fiber that is running
static EmuDim blockDim, gridDim;

// ---- vector types / conversions the kernels use
// (CUDA's alignment requirements: an 8- / 16-byte vector access must be naturally aligned -- UBSan's alignment check sees these)
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float rsqrtf(float v) { return 1.0f / std::sqrt(v); }
static inline uint16_t emu_bf16_rn(float f) {          // cvt.rn.bf16.f32 for finite inputs
  uint32_t u; std::memcpy(&u, &f, 4);
  return static_cast<uint16_t>((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
struct __nv_bfloat16 { uint16_t v; };
struct __nv_bfloat162 { uint16_t x, y; };              // .x = low half of the 32-bit word
static inline __nv_bfloat16 __float2bfloat16(float f) { return __nv_bfloat16{emu_bf16_rn(f)}; }
static inline float __bfloat162float(__nv_bfloat16 h) { return __uint_as_float(static_cast<uint32_t>(h.v) << 16); }
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{emu_bf16_rn(a), emu_bf16_rn(b)}; }
static inline uint32_t pack_bf16x2(float lo, float hi) { return emu_bf16_rn(lo) | (static_cast<uint32_t>(emu_bf16_rn(hi)) << 16); }
static inline float2 unpack_bf16x2(uint32_t v) { return float2{__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)}; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
static inline uint16_t emu_f16_rn(float f) {           // cvt.rn.f16.f32
  uint32_t u; std::memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const int32_t e = static_cast<int32_t>((u >> 23) & 0xffu) - 127 + 15;
  uint32_t m = u & 0x7fffffu;
  if (((u >> 23) & 0xffu) == 0xffu) return static_cast<uint16_t>(sign | 0x7c00u | (m ? 0x200u : 0u));
  if (e >= 31) return static_cast<uint16_t>(sign | 0x7c00u);
  if (e <= 0) {                                         // subnormal half (or zero)
    if (e < -10) return static_cast<uint16_t>(sign);
    m |= 0x800000u;
    const int shift = 14 - e;                           // 14 .. 24
    const uint32_t half = m >> shift, rem = m & ((1u << shift) - 1u), mid = 1u << (shift - 1);
    return static_cast<uint16_t>(sign | (half + ((rem > mid || (rem == mid && (half & 1u))) ? 1u : 0u)));
  }
  const uint32_t half = (static_cast<uint32_t>(e) << 10) | (m >> 13), rem = m & 0x1fffu;
  return static_cast<uint16_t>(sign | (half + ((rem > 0x1000u || (rem == 0x1000u && (half & 1u))) ? 1u : 0u)));
}
struct __half { uint16_t v; };
static inline __half __float2half(float f) { return __half{emu_f16_rn(f)}; }
static inline __half __float2half_rn(float f) { return __half{emu_f16_rn(f)}; }
static inline void griddep_wait() {}                   // programmatic dependent launch: nothing to wait for here
static inline void griddep_launch_dependents() {}

// ---- fibers and barriers that tolerate threads leaving
struct EmuFiber { ucontext_t ctx; EmuIdx tidx; bool done = false; };
static ucontext_t emu_sched_ctx;
static EmuFiber* emu_cur = nullptr;
static std::function<void()> emu_kernel_call;
static inline void emu_yield() { swapcontext(&emu_cur->ctx, &emu_sched_ctx); }
static std::vector<std::function<void()>> emu_deferred;   // asynchronous completions: applied at the start of the next scheduler pass
static const void* emu_wait_on[1024];         // what each thread is parked at (diagnostics of the deadlock message)
static unsigned long emu_progress = 0;       // barrier arrivals / releases + finished fibers: a scheduler pass without any is a deadlock

class EmuBarrier {
  int expected_ = 0, waiting_ = 0;
  unsigned long gen_ = 0;
 public:
  void reset(int n) { expected_ = n; waiting_ = 0; }
  void arrive_and_wait() {
    const unsigned long g = gen_;
    ++emu_progress;                                    // an arrival is a state change, released or not
    if (++waiting_ >= expected_) { waiting_ = 0; ++gen_; return; }
    emu_wait_on[(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 1023u] = this;
    while (gen_ == g) emu_yield();
    ++emu_progress;                                    // so is leaving a wait
  }
  void drop() {
    --expected_;
    if (expected_ > 0 && waiting_ >= expected_) { waiting_ = 0; ++gen_; ++emu_progress; }
  }
};
static EmuBarrier emu_block_bar;
static EmuBarrier emu_warp_bar[32];
static uint32_t emu_xchg[1024];
static std::vector<unsigned char> emu_dyn_smem;

static inline unsigned emu_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
static inline void __syncthreads() { emu_block_bar.arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_bar[emu_tid() >> 5].arrive_and_wait(); }
template <class T>
static inline T emu_shfl(T v, unsigned src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  const unsigned tid = emu_tid(), w = tid >> 5;
  std::memcpy(&emu_xchg[tid], &v, 4);
  emu_warp_bar[w].arrive_and_wait();
  T out;
  std::memcpy(&out, &emu_xchg[(w << 5) | (src_lane & 31u)], 4);
  emu_warp_bar[w].arrive_and_wait();
  return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu_shfl(v, (emu_tid() & 31u) ^ static_cast<unsigned>(lane_mask)); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, static_cast<unsigned>(src)); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  const unsigned lane = emu_tid() & 31u;
  const T r = emu_shfl(v, lane + d < 32u ? lane + d : lane);
  return r;
}
static inline float atomicAdd(float* p, float v) {
  const float old = *p;
  *p = old + v;
  return old;
}
static inline void* emu_dynamic_smem() { return emu_dyn_smem.data(); }

// ---- launch: blocks one after the other, the threads of a block as round-robin fibers
static constexpr size_t kEmuStackBytes = 256 * 1024;
static std::vector<unsigned char*> emu_stacks;
static void emu_trampoline() {
  emu_kernel_call();
  emu_cur->done = true;
  ++emu_progress;
  const unsigned t = emu_cur->tidx.x + blockDim.x * (emu_cur->tidx.y + blockDim.y * emu_cur->tidx.z);
  emu_warp_bar[t >> 5].drop();
  emu_block_bar.drop();
  swapcontext(&emu_cur->ctx, &emu_sched_ctx);
}
template <class F>
static void emu_launch(EmuDim grid, EmuDim block, size_t dyn_smem_bytes, F kernel_call) {
  gridDim = grid;
  blockDim = block;
  emu_kernel_call = kernel_call;
  emu_dyn_smem.assign(dyn_smem_bytes, 0xCD);           // exact size (bounds-checked under ASan); "uninitialised" is not zero
  const unsigned nthreads = block.x * block.y * block.z;
  while (emu_stacks.size() < nthreads) emu_stacks.push_back(new unsigned char[kEmuStackBytes]);
  std::vector<EmuFiber> fibers(nthreads);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        emu_block_bar.reset(static_cast<int>(nthreads));
        for (unsigned w = 0; w * 32 < nthreads; ++w) emu_warp_bar[w].reset(static_cast<int>(std::min(32u, nthreads - w * 32)));
        for (unsigned t = 0; t < nthreads; ++t) {
          EmuFiber& f = fibers[t];
          f.done = false;
          f.tidx.x = t % block.x; f.tidx.y = (t / block.x) % block.y; f.tidx.z = t / (block.x * block.y);
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = emu_stacks[t];
          f.ctx.uc_stack.ss_size = kEmuStackBytes;
          f.ctx.uc_link = &emu_sched_ctx;
          makecontext(&f.ctx, emu_trampoline, 0);
        }
        // ROVITKAN_EMU_ORDER=reverse resumes the fibers from the last thread to the first: a kernel whose result depended on the
        // order in which threads reach a barrier-free stretch (a missing __syncthreads / __syncwarp) would differ between the orders
        static const bool reverse = [] { const char* e = std::getenv("ROVITKAN_EMU_ORDER"); return e != nullptr && e[0] == 'r'; }();
        unsigned remaining = nthreads;
        while (remaining > 0) {
          const unsigned long before = emu_progress;
          if (!emu_deferred.empty()) {
            std::vector<std::function<void()>> due;
            due.swap(emu_deferred);
            for (auto& fn : due) fn();
            ++emu_progress;
          }
          for (unsigned i = 0; i < nthreads; ++i) {
            const unsigned t = reverse ? nthreads - 1 - i : i;
            EmuFiber& f = fibers[t];
            if (f.done) continue;
            emu_cur = &f;
            threadIdx = f.tidx;
            swapcontext(&emu_sched_ctx, &f.ctx);
            if (f.done) --remaining;
          }
          if (remaining > 0 && emu_progress == before) {        // every live fiber is parked at a barrier that cannot complete
            std::fprintf(stderr, "cuda_host_emu: deadlock in block (%u,%u,%u): %u threads wait at a barrier the others never reach "
                                 "(divergent __syncthreads / partial-warp shuffle)\n", bx, by, bz, remaining);
            for (unsigned t = 0; t < nthreads && t < 1024; ++t)
              if (!fibers[t].done && (t % 32 == 0 || emu_wait_on[t] != emu_wait_on[t - 1])) std::fprintf(stderr, "  thread %u waits on %p\n", t, emu_wait_on[t]);
            std::abort();
          }
        }
      }
}
