// Host launchers for the tcgen05 GEMM kernels: build the TMA tensor maps, pick the grid, launch.
#include "kernels.h"
#include "mlp_fused.cuh"
#include "mlp_fused2.cuh"
#include "tma_host.h"

#include <cstdlib>
#include <vector>

namespace {

constexpr int kBN = 192;       // every N on this path (192, 576, 768) is a multiple of 192
constexpr int kNtStages = 3;
constexpr int kBQ = 192;
constexpr int kTnStages = 4;

template <int MODE, int kStages>
int launch_nt_stages(const GemmNtArgs& a, cudaStream_t stream) {
  using L = GemmNtSmem<kBN, kStages>;
  auto kernel = gemm_nt_kernel<kBN, MODE, kStages>;
  RVK_SET_MAX_SMEM(kernel, L::kTotal);
  const GemmNtParams& p = a.p;
  CUtensorMap tmA, tmB, tmOut, tmOut2, tmAux;
  RVK_TRY(rvk_make_tmap_2d(&tmA, a.A, RVK_BF16, p.M, p.K, a.lda, 128, 64));
  RVK_TRY(rvk_make_tmap_2d(&tmB, a.B, RVK_BF16, p.N, p.K, a.ldb, kBN, 64));
  const bool out_f32 = (MODE == EPI_F32 || MODE == EPI_RES_LN);
  // stores go out per epilogue warp: 32-row boxes
  if (MODE == EPI_RES_LN && p.out_tiled != nullptr) tmOut = tmA;   // x' leaves through registers, not TMA
  else RVK_TRY(rvk_make_tmap_2d(&tmOut, a.out, out_f32 ? RVK_F32 : RVK_BF16, p.M, p.N, a.ldo, 32, out_f32 ? 32 : 64));
  tmOut2 = tmOut;
  tmAux = tmOut;
  if (p.has_out2) {
    if (a.out2 == nullptr) return RVK_ERR_BAD_ARG;
    RVK_TRY(rvk_make_tmap_2d(&tmOut2, a.out2, RVK_BF16, p.M, p.N, a.ldo2, 32, 64));
  }
  if (MODE == EPI_DGELU) {
    if (a.aux == nullptr) return RVK_ERR_BAD_ARG;
    RVK_TRY(rvk_make_tmap_2d(&tmAux, a.aux, RVK_BF16, p.M, p.N, a.ldaux, 128, 64));
  } else if (MODE == EPI_RES_LN && p.has_res && p.res_table == nullptr && p.out_tiled == nullptr) {
    if (a.aux == nullptr) return RVK_ERR_BAD_ARG;
    RVK_TRY(rvk_make_tmap_2d(&tmAux, a.aux, RVK_F32, p.M, p.N, a.ldaux, 128, 32));
  }
  const int tiles = ((p.M + 127) / 128) * (p.N / kBN);
  const int grid = tiles < kNumSMsB200 ? tiles : kNumSMsB200;
  RvkScopedTimer timer(stream, 2.0 * p.M * p.N * p.K, 0.0, RVK_T_GEMM_NT);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = rvk_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RVK_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmOut, tmOut2, tmAux, p));
  return rvk_launch_check();
}

// The aux-loading DGELU mode holds an epilogue panel slot from the load of the saved z panel until the dz store out of the same
// slot has drained; with six slots its epilogue warps spent 23 % of their samples waiting for a panel (ncu).  K = 192 (three K
// blocks per tile) needs little operand prefetch, so this mode runs two ring stages and the 40 KB they free hold two more
// slots: 54.5 -> 44.5 us per launch at 256 images.  The other K = 192 modes do not hold slots across a load and measured
// SLOWER on that variant (train 42.7 k -> 41.9 k img/s, inference 213.5 k -> 203.9 k): they keep three stages.
// RVK_NT_STAGES=2 / 3 forces one variant for every K <= 192 launch (experiments).
template <int MODE>
int launch_nt(const GemmNtArgs& a, cudaStream_t stream) {
  static const int force = [] { const char* e = getenv("RVK_NT_STAGES"); return e != nullptr ? atoi(e) : 0; }();
  const bool two = a.p.K <= 192 && force != 3 && (force == 2 || MODE == EPI_DGELU);
  return two ? launch_nt_stages<MODE, 2>(a, stream) : launch_nt_stages<MODE, kNtStages>(a, stream);
}

}  // namespace

bool rvk_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("RVK_PDL"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}

int rvk_gemm_nt_launch(const GemmNtArgs& a, cudaStream_t stream) {
  const GemmNtParams& p = a.p;
  if (p.M <= 0) return RVK_OK;
  const bool tiled_out = a.mode == EPI_RES_LN && p.out_tiled != nullptr;
  if (p.N <= 0 || p.K <= 0 || a.A == nullptr || a.B == nullptr || (a.out == nullptr && !tiled_out)) return RVK_ERR_BAD_ARG;
  if (!tiled_out && p.res_tiled != nullptr) return RVK_ERR_BAD_ARG;
  if (tiled_out && p.has_res && p.res_table == nullptr && p.res_tiled == nullptr) return RVK_ERR_BAD_ARG;
  if (p.N % kBN != 0 || p.N > 768 || p.K % 64 != 0) return RVK_ERR_UNSUPPORTED_SHAPE;
  switch (a.mode) {
    case EPI_BF16: return launch_nt<EPI_BF16>(a, stream);
    case EPI_GELU: return launch_nt<EPI_GELU>(a, stream);
    case EPI_DGELU: return launch_nt<EPI_DGELU>(a, stream);
    case EPI_F32: return launch_nt<EPI_F32>(a, stream);
    case EPI_RES_LN:
      if (p.N != 192) return RVK_ERR_UNSUPPORTED_SHAPE;
      if (p.res_table != nullptr && p.table_rows <= 0) return RVK_ERR_BAD_ARG;
      return launch_nt<EPI_RES_LN>(a, stream);
    default: return RVK_ERR_BAD_ARG;
  }
}

int rvk_gemm_tn_launch(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int M, int P,
                       int Q, float scale, float* a_colsum, cudaStream_t stream) {
  if (M <= 0 || P <= 0 || Q <= 0) return RVK_OK;
  if (a_colsum != nullptr && Q > kBQ) return RVK_ERR_UNSUPPORTED_SHAPE;    // the ones column rides on a single Q tile
  if (A == nullptr || B == nullptr || C == nullptr) return RVK_ERR_BAD_ARG;
  if (P % 64 != 0 || Q % 64 != 0) return RVK_ERR_UNSUPPORTED_SHAPE;
  using L = GemmTnSmem<kBQ, kTnStages>;
  auto kernel = gemm_tn_kernel<kBQ, kTnStages>;
  RVK_SET_MAX_SMEM(kernel, L::kTotal);
  CUtensorMap tmA, tmB;
  RVK_TRY(rvk_make_tmap_2d(&tmA, A, RVK_BF16, M, P, lda, 64, 64));
  RVK_TRY(rvk_make_tmap_2d(&tmB, B, RVK_BF16, M, Q, ldb, 64, 64));
  const int tiles = ((P + 127) / 128) * ((Q + kBQ - 1) / kBQ);
  const int total_chunks = (M + 63) / 64;
  // one CTA per SM and launch: a CTA pays ~7 us of fixed cost (TMEM allocation, pipeline fill, the red.global.add epilogue of its
  // 128 x 192 tile), so a single full wave beats two half-length ones (measured: 31.4k vs 30.2k img/s on the train step).
  // RVK_TN_WAVES=2..4 restores finer splits for A/B measurements.
  static const int waves = [] { const char* e = getenv("RVK_TN_WAVES"); return (e != nullptr && e[0] >= '1' && e[0] <= '4') ? e[0] - '0' : 1; }();
  int splits = waves * kNumSMsB200 / tiles;
  if (splits > total_chunks) splits = total_chunks;
  if (splits < 1) splits = 1;
  GemmTnParams p;
  p.M = M; p.P = P; p.Q = Q;
  p.ldc = static_cast<int>(ldc);
  p.chunks_per_split = (total_chunks + splits - 1) / splits;
  p.C = C;
  p.scale = scale;
  p.colsum = a_colsum;
  splits = (total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
  RvkScopedTimer timer(stream, 2.0 * M * P * Q, 0.0, RVK_T_GEMM_TN);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles, splits);
  cfg.blockDim = dim3(kTnThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = rvk_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RVK_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, p));
  return rvk_launch_check();
}

// ---- fused MLP block (fc1 + GELU + fc2 + residual + LayerNorm), inference path ---------------------------
template <int G>
static int launch_mlp_fused(const MlpFusedArgs& a, cudaStream_t stream) {
  using L = MlpSmem<G>;
  auto kernel = mlp_fused_kernel<G>;
  RVK_SET_MAX_SMEM(kernel, L::kTotal);
  const MlpFusedParams& p = a.p;
  CUtensorMap tmW1, tmW2, tmLn, tmCtx, tmWp;
  RVK_TRY(rvk_make_tmap_2d(&tmW1, a.w1, RVK_BF16, 768, 192, 192, 128 / G, 64));
  RVK_TRY(rvk_make_tmap_2d(&tmW2, a.w2_f16, RVK_BF16 /* 2-byte elements */, 192, 768, 768, 192 / G, 64));
  tmLn = tmW1;
  if (p.has_ln) RVK_TRY(rvk_make_tmap_2d(&tmLn, a.ln_out, RVK_BF16, p.M, 192, 192, 32, 64));
  tmCtx = tmW1;
  tmWp = tmW1;
  if (p.has_proj) {
    RVK_TRY(rvk_make_tmap_2d(&tmCtx, a.ctx, RVK_BF16, p.M, 192, 192, 128, 64));
    RVK_TRY(rvk_make_tmap_2d(&tmWp, a.wproj, RVK_BF16, 192, 192, 192, 32, 64));
  }
  const int tiles = (p.M + 127) / 128;
  const int units = (tiles + G - 1) / G;
  const int max_clusters = kNumSMsB200 / G;
  const int clusters = units < max_clusters ? units : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * G);
  cfg.blockDim = dim3(kMlpThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = rvk_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  RvkScopedTimer timer(stream, 2.0 * p.M * 192.0 * (768.0 * 2.0 + (p.has_proj ? 192.0 : 0.0)), 0.0, RVK_T_MLP_FUSED);
  RVK_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, tmW1, tmW2, tmLn, tmCtx, tmWp, p));
  return rvk_launch_check();
}

// two row tiles in flight per CTA pair (mlp_fused2.cuh); needs the folded attention projection
template <int kProjQ, bool kTrace>
static int launch_mlp_fused2(const MlpFusedArgs& a, cudaStream_t stream) {
  using L = Mlp2Smem;
  auto kernel = mlp_fused2_kernel<kProjQ, kTrace>;
  RVK_SET_MAX_SMEM(kernel, L::kTotal);
  const MlpFusedParams& p = a.p;
  CUtensorMap tmW1, tmW2, tmLn, tmCtx, tmWp;
  RVK_TRY(rvk_make_tmap_2d(&tmW1, a.w1, RVK_BF16, 768, 192, 192, 32, 64));
  RVK_TRY(rvk_make_tmap_2d(&tmW2, a.w2_f16, RVK_BF16 /* 2-byte elements */, 192, 768, 768, 96, 64));
  tmLn = tmW1;
  if (p.has_ln) RVK_TRY(rvk_make_tmap_2d(&tmLn, a.ln_out, RVK_BF16, p.M, 192, 192, 32, 64));
  RVK_TRY(rvk_make_tmap_2d(&tmCtx, a.ctx, RVK_BF16, p.M, 192, 192, 128, 64));
  RVK_TRY(rvk_make_tmap_2d(&tmWp, a.wproj, RVK_BF16, 192, 192, 192, 32, 64));
  const int tiles = (p.M + 127) / 128;
  const int units = (tiles + 1) / 2;
  const int max_clusters = kNumSMsB200 / 2;
  const int clusters = units < max_clusters ? units : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(kMlp2Threads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = rvk_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  RvkScopedTimer timer(stream, 2.0 * p.M * 192.0 * (768.0 * 2.0 + 192.0), 0.0, RVK_T_MLP_FUSED);
  RVK_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, tmW1, tmW2, tmLn, tmCtx, tmWp, p));
  return rvk_launch_check();
}

// A/B switch of the two-tile kernel: RVK_MLP2_PROJQ=1..5 (half-chunk of tile i at which the projection of tile i+1 is issued)
static int mlp2_proj_q() {
  static const int q = [] {
    const char* e = getenv("RVK_MLP2_PROJQ");
    return (e != nullptr && e[0] >= '1' && e[0] <= '5') ? e[0] - '0' : 3;
  }();
  return q;
}

int rvk_mlp_fused_launch(const MlpFusedArgs& a, cudaStream_t stream) {
  const MlpFusedParams& p = a.p;
  if (p.M <= 0) return RVK_OK;
  if (a.w1 == nullptr || a.w2_f16 == nullptr || p.x_in == nullptr || p.x_out == nullptr || p.b1 == nullptr ||
      p.b2 == nullptr || p.gamma2 == nullptr || p.beta2 == nullptr)
    return RVK_ERR_BAD_ARG;
  if (p.has_ln && (a.ln_out == nullptr || p.gamma == nullptr || p.beta == nullptr)) return RVK_ERR_BAD_ARG;
  if (p.has_proj && (a.ctx == nullptr || a.wproj == nullptr || p.bp == nullptr)) return RVK_ERR_BAD_ARG;
  if (a.cta_group == 1) return launch_mlp_fused<1>(a, stream);
  if (a.cta_group == 2) return launch_mlp_fused<2>(a, stream);
  if (a.cta_group == 4) {           // CTA pairs, two row tiles in flight
    if (!p.has_proj) return RVK_ERR_BAD_ARG;
    if (p.trace != nullptr) return launch_mlp_fused2<3, true>(a, stream);     // debugging build with the clock64 event log
    switch (mlp2_proj_q()) {
      case 1: return launch_mlp_fused2<1, false>(a, stream);
      case 2: return launch_mlp_fused2<2, false>(a, stream);
      case 4: return launch_mlp_fused2<4, false>(a, stream);
      case 5: return launch_mlp_fused2<5, false>(a, stream);
      default: return launch_mlp_fused2<3, false>(a, stream);
    }
  }
  return RVK_ERR_BAD_ARG;
}
