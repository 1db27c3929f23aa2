#include "tma_host.h"

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn g_encode = nullptr;
std::once_flag g_once;

void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<EncodeFn>(fn);
}

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

}  // namespace

void rvk_set_last_cuda_error(int code, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(static_cast<cudaError_t>(code));
}
void rvk_set_last_error_text(const char* what) { g_last_error = what; }
const char* rvk_last_error_cstr() { return g_last_error.c_str(); }
void rvk_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- in-situ launch timing ------------------------------------------------------------------------------------
namespace {
struct TimedLaunch { cudaEvent_t beg, end; double flops, bytes; int kind; };
bool g_timing = false;
std::mutex g_timing_mu;
std::vector<TimedLaunch*> g_timed;
std::vector<TimedLaunch*> g_timer_pool;
double g_kind_ms[RVK_T_COUNT], g_kind_flops[RVK_T_COUNT], g_kind_bytes[RVK_T_COUNT];
int g_kind_n[RVK_T_COUNT];
const char* const kKindNames[RVK_T_COUNT] = {"gemm_nt_kernel", "gemm_tn_kernel", "mlp_fused_kernel", "attn_fwd_tc_kernel",
                                             "attn_bwd_tc_kernel", "layernorm_bwd_kernel", "kan_fwd", "kan_bwd", "heads_fused_kernel",
                                             "im2col_kernel", "optimizer_tail", "heads_train"};
}  // namespace

void* rvk_timer_begin(cudaStream_t s, double flops, double bytes, int kind) {
  if (!g_timing || kind < 0 || kind >= RVK_T_COUNT) return nullptr;
  TimedLaunch* t = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    if (!g_timer_pool.empty()) { t = g_timer_pool.back(); g_timer_pool.pop_back(); }
  }
  if (t == nullptr) {
    t = new TimedLaunch();
    if (cudaEventCreate(&t->beg) != cudaSuccess || cudaEventCreate(&t->end) != cudaSuccess) { delete t; return nullptr; }
  }
  t->flops = flops; t->bytes = bytes; t->kind = kind;
  cudaEventRecord(t->beg, s);
  return t;
}
void rvk_timer_end(void* handle, cudaStream_t s) {
  TimedLaunch* t = static_cast<TimedLaunch*>(handle);
  cudaEventRecord(t->end, s);
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_timed.push_back(t);
}
void rvk_timing_enable_impl(int on) { g_timing = on != 0; }
// Sums the device time of every timed launch since the last collect (caller must have synchronised the device).
int rvk_timing_collect_impl() {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (int k = 0; k < RVK_T_COUNT; ++k) { g_kind_ms[k] = g_kind_flops[k] = g_kind_bytes[k] = 0.0; g_kind_n[k] = 0; }
  int n = 0;
  for (TimedLaunch* t : g_timed) {
    float e = 0.0f;
    if (cudaEventElapsedTime(&e, t->beg, t->end) == cudaSuccess) {
      g_kind_ms[t->kind] += e; g_kind_flops[t->kind] += t->flops; g_kind_bytes[t->kind] += t->bytes; ++g_kind_n[t->kind]; ++n;
    }
    g_timer_pool.push_back(t);
  }
  g_timed.clear();
  return n;
}
int rvk_timing_kind_impl(int kind, double* ms, double* flops, double* bytes) {
  if (kind < 0 || kind >= RVK_T_COUNT) return -1;
  if (ms) *ms = g_kind_ms[kind];
  if (flops) *flops = g_kind_flops[kind];
  if (bytes) *bytes = g_kind_bytes[kind];
  return g_kind_n[kind];
}
const char* rvk_timing_kind_name_impl(int kind) { return (kind >= 0 && kind < RVK_T_COUNT) ? kKindNames[kind] : nullptr; }
long long rvk_launch_count_impl() { return g_launches.load(std::memory_order_relaxed); }

int rvk_make_tmap_2d(CUtensorMap* out, const void* base, int dtype, int64_t rows, int64_t cols, int64_t ld,
                     int box_rows, int box_cols) {
  std::call_once(g_once, resolve_encode);
  if (g_encode == nullptr) return RVK_ERR_NO_DRIVER;
  const int esz = (dtype == RVK_BF16) ? 2 : 4;
  if (box_cols * esz != 128 || box_rows > 256 || box_rows < 1) return RVK_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * esz) & 15) != 0) return RVK_ERR_ALIGNMENT;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = g_encode(out, dtype == RVK_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                        2, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_error = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
    return RVK_ERR_TMA_ENCODE;
  }
  return RVK_OK;
}

int rvk_make_tmap_3d(CUtensorMap* out, const void* base, int dtype, int64_t d0, int64_t d1, int64_t d2,
                     int64_t stride1, int64_t stride2, int box_d0, int box_d1) {
  std::call_once(g_once, resolve_encode);
  if (g_encode == nullptr) return RVK_ERR_NO_DRIVER;
  const int esz = (dtype == RVK_BF16) ? 2 : 4;
  if (box_d0 * esz != 128 || box_d1 > 256 || box_d1 < 1) return RVK_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((stride1 * esz) & 15) != 0 || ((stride2 * esz) & 15) != 0)
    return RVK_ERR_ALIGNMENT;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(stride1) * esz, static_cast<cuuint64_t>(stride2) * esz};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_d0), static_cast<cuuint32_t>(box_d1), 1};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = g_encode(out, dtype == RVK_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                        3, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_error = "cuTensorMapEncodeTiled(3d) failed with CUresult " + std::to_string(static_cast<int>(r));
    return RVK_ERR_TMA_ENCODE;
  }
  return RVK_OK;
}
