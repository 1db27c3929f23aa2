// Weight-gradient GEMM  C[P,Q] += A[M,P]^T * B[M,Q]  on tcgen05/TMEM, sm_100a.
//
// Both operands are activations stored row-major with the reduction index m outermost, so both are
// "MN-major" UMMA operands: a TMA box of [64 rows(m) x 64 cols] lands in shared memory as the
// canonical 128-byte-swizzled MN-major atom stack (64 contiguous MN elements per 128 B row, 8 rows
// per 1024 B swizzle atom).  Descriptor strides: LBO = distance between 64-column boxes (8 KB),
// SBO = distance between 8-row groups (1 KB).
//
// Bias gradients for free: when the launcher passes `colsum` (and Q fits one tile), a constant panel of ONES is appended
// to the B operand of every stage and the UMMA runs with N = BQ + 16: accumulator column BQ then holds sum_m A[m, p],
// the column sums of A = the bias gradient of the Linear layer whose output gradient A is.  8 % more tensor work instead
// of a separate pass over A.
//
// Grid = (P tiles of 128) x (Q tiles of BQ) x splits over M; every CTA reduces its M-range into one
// TMEM accumulator and adds it into the fp32 gradient with coalesced red.global.add (split-K).
// Warp roles (192 threads): w0 TMA producer, w1 UMMA issuer + TMEM owner, w2..w5 epilogue.
#pragma once

#include "common.cuh"

struct GemmTnParams {
  int M, P, Q;        // C is [P, Q]
  int ldc;
  int chunks_per_split;   // in units of 64 rows of M
  float* C;
  float scale;
  float* colsum;          // optional [P]: += scale * sum_m A[m, p]  (requires Q <= BQ)
};

constexpr int kTnThreads = 192;

template <int BQ, int STAGES>
struct GemmTnSmem {
  static constexpr int kABytes = 2 * 64 * 128;          // two [64 x 64] bf16 boxes
  static constexpr int kBBytes = (BQ / 64) * 64 * 128;
  static constexpr int kOnesBytes = 64 * 128;           // constant [64 x 64] panel of 1.0 right after the B boxes (colsum)
  static constexpr int kStageBytes = kABytes + kBBytes + kOnesBytes;
  static constexpr int kTxBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = 128 * 33 * 4;
  static constexpr int kTotal = 1024 + STAGES * kStageBytes + kStagingBytes + 256;
};

#ifdef __CUDACC__

template <int BQ, int STAGES>
__global__ void __launch_bounds__(kTnThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmTnParams p) {
  using L = GemmTnSmem<BQ, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset form keeps the shared address space (LDS/STS)
  uint8_t* sOperands = smem;
  float* sStage = reinterpret_cast<float*>(smem + STAGES * L::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStage) + L::kStagingBytes);
  uint64_t* full = bars;
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tiles = (p.Q + BQ - 1) / BQ;
  const int p0 = (blockIdx.x / q_tiles) * 128;
  const int q0 = (blockIdx.x % q_tiles) * BQ;
  const int total_chunks = (p.M + 63) / 64;
  const int c_begin = blockIdx.y * p.chunks_per_split;
  const int c_end = min(total_chunks, c_begin + p.chunks_per_split);
  const int n_chunks = max(0, c_end - c_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  if (p.colsum != nullptr) {      // the ones panels are never touched by TMA: written once
    for (int i = threadIdx.x; i < STAGES * (L::kOnesBytes / 16); i += blockDim.x) {
      const int st = i / (L::kOnesBytes / 16), o = i % (L::kOnesBytes / 16);
      *reinterpret_cast<uint4*>(sOperands + st * L::kStageBytes + L::kABytes + L::kBBytes + o * 16) =
          make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();                    // (programmatic dependent launch: the setup above overlapped the previous kernel's tail)
  griddep_launch_dependents();

  if (n_chunks > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int c = c_begin; c < c_end; ++c) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], L::kTxBytes);
          uint8_t* a = sOperands + s * L::kStageBytes;
          tma_load_2d(a, &tmA, &full[s], p0, c * 64);
          tma_load_2d(a + 8192, &tmA, &full[s], p0 + 64, c * 64);
#pragma unroll
          for (int b = 0; b < BQ / 64; ++b)
            tma_load_2d(a + L::kABytes + b * 8192, &tmB, &full[s], q0 + b * 64, c * 64);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = (p.colsum != nullptr) ? umma_idesc_bf16(128, BQ + 16, 1, 1) : umma_idesc_bf16(128, BQ, 1, 1);
        int s = 0;
        uint32_t ph = 0;
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sOperands + s * L::kStageBytes);
          const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 16 rows of m per UMMA = two 8-row swizzle atoms
            const uint64_t ad = umma_smem_desc(a_addr + k * 2048, 8192, 1024);
            const uint64_t bd = umma_smem_desc(b_addr + k * 2048, 8192, 1024);
            umma_bf16(tmem_base, ad, bd, idesc, (c | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(acc_full);
      }
    } else {
      const int quad = warp & 3;
      const int row = quad * 32 + lane;
      const int ew = warp - 2;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BQ / 32; ++c) {
        float v[32];
        tmem_ld32(tacc + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) sStage[row * 33 + i] = v[i];
        named_bar_sync(1, 128);
        // coalesced split-K reduction: one warp per output row, 32 consecutive columns per warp op
        for (int rr = ew * 32; rr < ew * 32 + 32; ++rr) {
          const int gp = p0 + rr;
          const int gq = q0 + c * 32 + lane;
          if (gp < p.P && gq < p.Q)
            atomicAdd(p.C + static_cast<size_t>(gp) * p.ldc + gq, sStage[rr * 33 + lane] * p.scale);
        }
        named_bar_sync(1, 128);
      }
      if (p.colsum != nullptr) {
        float v[32];
        tmem_ld32(tacc + BQ, v);             // column BQ = A^T . 1  (columns beyond BQ + 16 are unused TMEM)
        const int gp = p0 + row;
        if (gp < p.P) atomicAdd(p.colsum + gp, v[0] * p.scale);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

#endif  // __CUDACC__
