// Attention forward for DeiT-Tiny (197 tokens, 3 heads of 64) on tcgen05 / TMEM, sm_100a.
//
// Persistent CTAs walk (image, head) items.  Per item, with queries split in two 128-row M tiles t:
//   S_t = Q_t K^T      UMMA M=128 N=208 K=64   (Q, K: K-major bf16 tiles loaded by 3-D TMA; rows >= 197 are
//                                               zero-filled by the tensor map, so padding never needs masking
//                                               on the load side)
//   P_t = softmax      one thread per query row reads its whole score row from TMEM (tcgen05.ld), so the row
//                      max / sum need no shuffles; ONE pass over TMEM (shift by the first chunk's max, exact-max
//                      fallback behind an overflow guard); P_t is written as the K-major 128-byte-swizzled A
//                      operand of the next UMMA
//   O_t = P_t V        UMMA M=128 N=64 K=208, V consumed MN-major exactly as it lies in the qkv row (no
//                      transpose); O_t aliases the first 64 TMEM columns of S_t
//   ctx rows           O_t / rowsum -> bf16 -> swizzled staging (P_t's first panel) -> per-warp 3-D TMA store
//                      (the tensor map clips rows >= 197 of the image)
// Warp roles (320 threads): w0-3 softmax/epilogue of tile 0, w4-7 of tile 1 (TMEM lane quadrant = warp % 4),
// w8 TMA producer, w9 UMMA issuer + TMEM owner.  Q/K are released to the producer as soon as both S tiles
// are issued and V as soon as both PV products are, so the next item's loads overlap this item's softmax.
// Restates timm Attention.forward: softmax(q k^T * 64^-0.5) v (oracle/vit.py::_Attention).
#include "kernels.h"
#include "tma_host.h"

namespace {

constexpr int kTok = 197, kHeads = 3, kHd = 64;
constexpr int kKeysPad = 208;                  // 13 UMMA K-steps of 16 keys
constexpr int kThreads = 320;
constexpr int kQBytes = 256 * 128;             // two M tiles
constexpr int kKVBytes = kKeysPad * 128;
constexpr int kPBytes = 4 * 128 * 128;         // four 64-key panels of [128 x 128 B]
constexpr int kSmemBytes = 1024 + kQBytes + 2 * kKVBytes + 2 * kPBytes + 256;
constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmCtx, float* __restrict__ lse, int num_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;
  uint8_t* sV = sK + kKVBytes;
  uint8_t* sP = sV + kKVBytes;                 // [2][kPBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* qk_full = bars;
  uint64_t* v_full = bars + 1;
  uint64_t* qk_empty = bars + 2;
  uint64_t* v_empty = bars + 3;
  uint64_t* s_full = bars + 4;       // [2]
  uint64_t* p_full = bars + 6;       // [2]
  uint64_t* o_full = bars + 8;       // [2]
  uint64_t* tmem_free = bars + 10;   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmCtx);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_empty, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_full[t], 1);
      mbar_init(&tmem_free[t], 128);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    // ================================================================= TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int b = item / kHeads, h = item % kHeads;
        mbar_wait(qk_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(qk_full, kQBytes + kKVBytes);
        tma_load_3d(sQ, &tmQ, qk_full, h * kHd, 0, b);
        tma_load_3d(sK, &tmKV, qk_full, 192 + h * kHd, 0, b);
        mbar_wait(v_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(v_full, kKVBytes);
        tma_load_3d(sV, &tmKV, v_full, 384 + h * kHd, 0, b);
      }
    }
  } else if (warp == 9) {
    // ================================================================= UMMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKeysPad, 0, 0);   // S = Q K^T
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHd, 0, 1);        // O = P V, V MN-major
      uint32_t it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        mbar_wait(qk_full, ph);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&tmem_free[t], ph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + t * 256, umma_smem_desc(q_addr + t * 16384 + k * 32, 16, 1024),
                      umma_smem_desc(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
        }
        umma_commit(qk_empty);            // Q and K may be overwritten once both S tiles are done
        mbar_wait(v_full, ph);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], ph);
          tc_fence_after();
          const uint32_t p_addr = smem_u32(sP + t * kPBytes);
#pragma unroll
          for (int j = 0; j < 13; ++j)
            umma_bf16(tmem_base + t * 256, umma_smem_desc(p_addr + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                      umma_smem_desc(v_addr + j * 2048, 8192, 1024), idesc_o, j != 0 ? 1u : 0u);
          umma_commit(&o_full[t]);
        }
        umma_commit(v_empty);
      }
    }
  } else {
    // ================================================================= softmax + epilogue warps
    const int t = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;                 // TMEM lane
    const int qrow = t * 128 + row;                   // query index inside the image
    const bool warp_valid = (t * 128 + quad * 32) < kTok;
    const uint32_t tS = tmem_base + t * 256 + (static_cast<uint32_t>(quad * 32) << 16);
    uint8_t* myP = sP + t * kPBytes;
    // exp2 of one 32-column chunk of scaled scores, shifted by `shift`; returns packed bf16 P in the panel and
    // accumulates the row sum and the largest exponent argument seen (overflow guard)
    auto softmax_chunk = [&](int c, float (&v)[32], float shift, float& sum, float& amax) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (c * 32 + i < kTok) {
          const float a = fmaf(v[i], kScaleLog2e, -shift);
          amax = fmaxf(amax, a);
          v[i] = ex2_approx(a);
          sum += v[i];
        } else {
          v[i] = 0.0f;
        }
      }
      uint8_t* panel = myP + (c >> 1) * 16384;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q = make_uint4(pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                             pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]));
        *reinterpret_cast<uint4*>(panel + sw128_offset(row, (c & 1) * 4 + j)) = q;
      }
    };

    uint32_t it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int b = item / kHeads, h = item % kHeads;
      mbar_wait(&s_full[t], ph);
      tc_fence_after();
      float shift = 0.0f, sum = 0.0f;
      if (warp_valid) {
        // the previous item's ctx store read this warp's rows of P_t panel 0: it must be done before P is rewritten
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        // Single pass over TMEM (its read bandwidth, not MUFU, bounds this kernel): softmax is shift-invariant, so
        // the row is shifted by the max of its FIRST 32 scores instead of the full-row max.  The true max is tracked
        // on the fly; only if it exceeds the provisional shift by more than 2^64 (never for sane logits) is the row
        // redone with the exact max.  sum >= 1 always (the provisional max itself contributes 2^0).
        float v[32];
        tmem_ld32(tS, v);
        float m = v[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) m = fmaxf(m, v[i]);
        shift = m * kScaleLog2e;
        float amax = 0.0f;
        softmax_chunk(0, v, shift, sum, amax);
#pragma unroll 1
        for (int c = 1; c < 7; ++c) {
          tmem_ld32(tS + c * 32, v);
          softmax_chunk(c, v, shift, sum, amax);
        }
        if (__any_sync(0xffffffffu, amax > 64.0f)) {      // rare exact path: redo the warp's rows with the true max
          shift += amax;
          sum = 0.0f;
          float dummy = 0.0f;
#pragma unroll 1
          for (int c = 0; c < 7; ++c) {
            tmem_ld32(tS + c * 32, v);
            softmax_chunk(c, v, shift, sum, dummy);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_full[t]);

      mbar_wait(&o_full[t], ph);
      tc_fence_after();
      float o[2][32];
      if (warp_valid) {
        tmem_ld32(tS, o[0]);
        tmem_ld32(tS + 32, o[1]);
      }
      tc_fence_before();
      mbar_arrive(&tmem_free[t]);          // S_t / O_t columns are free: the next item's Q K^T may start now
      if (warp_valid) {
        const float inv = 1.0f / sum;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float* sv = &o[j >> 2][(j & 3) * 8];
          uint4 q = make_uint4(pack_bf16x2(sv[0] * inv, sv[1] * inv), pack_bf16x2(sv[2] * inv, sv[3] * inv),
                               pack_bf16x2(sv[4] * inv, sv[5] * inv), pack_bf16x2(sv[6] * inv, sv[7] * inv));
          *reinterpret_cast<uint4*>(myP + sw128_offset(row, j)) = q;     // P_t is dead: reuse its first panel
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmCtx, myP + quad * 32 * 128, h * kHd, t * 128 + quad * 32, b);
          tma_store_commit();
        }
        if (lse != nullptr && qrow < kTok) lse[static_cast<size_t>(item) * kTok + qrow] = shift + log2f(sum);
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int rvk_attention_fwd_launch(const void* qkv, void* ctx, float* lse, int batch, cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  static bool configured = false;
  if (!configured) {
    RVK_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  CUtensorMap tmQ, tmKV, tmCtx;
  RVK_TRY(rvk_make_tmap_3d(&tmQ, qkv, RVK_BF16, 576, kTok, batch, 576, int64_t(kTok) * 576, 64, 256));
  RVK_TRY(rvk_make_tmap_3d(&tmKV, qkv, RVK_BF16, 576, kTok, batch, 576, int64_t(kTok) * 576, 64, kKeysPad));
  RVK_TRY(rvk_make_tmap_3d(&tmCtx, ctx, RVK_BF16, 192, kTok, batch, 192, int64_t(kTok) * 192, 64, 32));
  const int items = batch * kHeads;
  const int grid = items < kNumSMsB200 ? items : kNumSMsB200;
  attn_fwd_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tmQ, tmKV, tmCtx, lse, items);
  return rvk_launch_check();
}
