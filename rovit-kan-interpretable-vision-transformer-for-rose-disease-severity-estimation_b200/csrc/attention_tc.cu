// Attention forward for DeiT-Tiny (197 tokens, 3 heads of 64) on tcgen05 / TMEM, sm_100a.
//
// Persistent CTAs walk (image, head) items.  Per item, with queries split in two 128-row M tiles t:
//   S_t = Q_t K^T      UMMA M=128 N=208 K=64   (Q, K: K-major bf16 tiles loaded by 3-D TMA; rows >= 197 are
//                                               zero-filled by the tensor map, so padding never needs masking
//                                               on the load side)
//   P_t = softmax      SIXTEEN softmax warps: warp (quad, cg) owns the 32 query rows of TMEM lane quadrant `quad`
//                      (of both tiles) and the key-column group cg (64 / 48 / 48 / 48 of the 208 padded keys).  A
//                      thread's slice of its score row fits in registers, so TMEM is read exactly once; the exact
//                      row maximum and the row sum are combined over the four column groups through shared memory
//                      (one 128-thread named barrier per tile).  P_t is written as the K-major 128-byte-swizzled A
//                      operand of the next UMMA.  (The first version used 8 warps with a whole 208-score row per
//                      thread: two warps per SM sub-partition could not hide the TMEM / MUFU latencies.)
//   O_t = P_t V        UMMA M=128 N=64 K=208, V consumed MN-major exactly as it lies in the qkv row (no
//                      transpose); O_t aliases the first 64 TMEM columns of S_t
//   ctx rows           O_t / rowsum -> bf16 (16 columns per warp) -> swizzled staging (P_t's first panel) ->
//                      per-quadrant 3-D TMA store (the tensor map clips rows >= 197 of the image)
// Warp roles (576 threads): w0-15 softmax / epilogue, w16 TMA producer, w17 UMMA issuer + TMEM owner.  Q/K are
// released to the producer as soon as both S tiles are issued and V as soon as both PV products are, so the next
// item's loads overlap this item's softmax.
// Restates timm Attention.forward: softmax(q k^T * 64^-0.5) v (oracle/vit.py::_Attention).
#include "kernels.h"
#include "tma_host.h"

namespace {

constexpr int kTok = 197, kHeads = 3, kHd = 64;
constexpr int kKeysPad = 208;                  // 13 UMMA K-steps of 16 keys
constexpr int kSmWarps = 16;
constexpr int kThreads = (kSmWarps + 2) * 32;  // 576
constexpr int kQBytes = 256 * 128;             // two M tiles
constexpr int kKVBytes = kKeysPad * 128;
constexpr int kPBytes = 4 * 128 * 128;         // four 64-key panels of [128 x 128 B]
constexpr int kXchgBytes = 2 * 2 * 128 * 4 * 4;   // [max | sum][tile][row][column group] fp32
constexpr int kSmemBytes = 1024 + kQBytes + 2 * kKVBytes + 2 * kPBytes + kXchgBytes + 256;
constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;

__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmCtx, float* __restrict__ lse, int num_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;
  uint8_t* sV = sK + kKVBytes;
  uint8_t* sP = sV + kKVBytes;                 // [2][kPBytes]
  float* sMax = reinterpret_cast<float*>(sP + 2 * kPBytes);   // [2][128][4]
  float* sSum = sMax + 2 * 128 * 4;                            // [2][128][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 2 * 128 * 4);
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
  if (warp == kSmWarps && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmCtx);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_empty, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], kSmWarps);
      mbar_init(&o_full[t], 1);
      mbar_init(&tmem_free[t], kSmWarps);
    }
    fence_mbar_init();
  }
  if (warp == kSmWarps + 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == kSmWarps) {
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
  } else if (warp == kSmWarps + 1) {
    // ================================================================= UMMA issuer
    // The whole warp walks the loop with warp-uniform values (descriptors stay on the uniform datapath); only the
    // tcgen05 instructions are issued by one elected lane.  Per-lane descriptor arithmetic cost ~100 cycles per MMA --
    // three times the 32 cycles a 128x64x16 MMA takes -- and serialised the S -> softmax -> PV -> O chain of an item.
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKeysPad, 0, 0);   // S = Q K^T
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHd, 0, 1);        // O = P V, V MN-major
      const bool issuer = elect_one();
      const uint32_t q_lo = umma_desc_lo(smem_u32(sQ)), k_lo = umma_desc_lo(smem_u32(sK));
      const uint32_t p_lo = umma_desc_lo(smem_u32(sP)), v_lo = umma_desc_lo(smem_u32(sV), 8192);
      uint32_t it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        mbar_wait(qk_full, ph);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&tmem_free[t], ph ^ 1);
          tc_fence_after();
          if (issuer) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_split<1>(tmem_base + t * 256, q_lo + t * (16384 >> 4) + 2 * k, k_lo + 2 * k, idesc_s, k != 0 ? 1u : 0u);
            umma_commit(&s_full[t]);
          }
          __syncwarp();
        }
        if (issuer) umma_commit(qk_empty);            // Q and K may be overwritten once both S tiles are done
        mbar_wait(v_full, ph);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], ph);
          tc_fence_after();
          if (issuer) {
#pragma unroll
            for (int j = 0; j < 13; ++j)
              umma_f16_split<1>(tmem_base + t * 256, p_lo + t * (kPBytes >> 4) + (j >> 2) * (16384 >> 4) + (j & 3) * 2,
                                v_lo + j * (2048 >> 4), idesc_o, j != 0 ? 1u : 0u);
            umma_commit(&o_full[t]);
          }
          __syncwarp();
        }
        if (issuer) umma_commit(v_empty);
      }
    }
  } else {
    // ================================================================= softmax + epilogue warps
    const int quad = warp & 3, cg = warp >> 2;
    const int row = quad * 32 + lane;                 // TMEM lane = query row inside the tile
    const int g0 = (cg == 0) ? 0 : 1 + 3 * cg;        // first 16-key group: groups 0-3 | 4-6 | 7-9 | 10-12
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;

    uint32_t it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int b = item / kHeads, h = item % kHeads;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const bool valid = (t * 128 + quad * 32) < kTok;       // uniform over the four warps of a quadrant
        uint8_t* myP = sP + t * kPBytes;
        mbar_wait(&s_full[t], ph);
        tc_fence_after();
        if (valid) {
          // the previous item's ctx store read this quadrant's rows of P_t panel 0: done before P is rewritten
          if (cg == 0 && lane == 0) tma_store_wait_read<0>();
          // 48 scores per thread stay in registers (column group 0 has 16 more: read twice, 7% extra TMEM traffic);
          // only column group 3 contains padded keys (197..207 = its elements 37..47)
          float v[48], w[16];
          const uint32_t tS = tmem_base + t * 256 + g0 * 16 + lane_sel;
          tmem_ld32(tS, *reinterpret_cast<float(*)[32]>(&v[0]));
          tmem_ld16f(tS + 32, &v[32]);
          float m = v[0];
#pragma unroll
          for (int i = 1; i < 32; ++i) m = fmaxf(m, v[i]);
          if (cg == 3) {
#pragma unroll
            for (int i = 32; i < 37; ++i) m = fmaxf(m, v[i]);
          } else {
#pragma unroll
            for (int i = 32; i < 48; ++i) m = fmaxf(m, v[i]);
          }
          if (cg == 0) {
            tmem_ld16f(tS + 48, w);
#pragma unroll
            for (int i = 0; i < 16; ++i) m = fmaxf(m, w[i]);
          }
          sMax[(t * 128 + row) * 4 + cg] = m;
          named_bar_sync(1 + quad, 128);
          const float4 m4 = *reinterpret_cast<const float4*>(&sMax[(t * 128 + row) * 4]);
          const float shift = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w)) * kScaleLog2e;
          float sum = 0.0f;
#pragma unroll
          for (int i = 0; i < 32; ++i) { v[i] = ex2_approx(fmaf(v[i], kScaleLog2e, -shift)); sum += v[i]; }
          if (cg == 3) {
#pragma unroll
            for (int i = 32; i < 48; ++i) {
              if (i < 37) { v[i] = ex2_approx(fmaf(v[i], kScaleLog2e, -shift)); sum += v[i]; } else { v[i] = 0.0f; }
            }
          } else {
#pragma unroll
            for (int i = 32; i < 48; ++i) { v[i] = ex2_approx(fmaf(v[i], kScaleLog2e, -shift)); sum += v[i]; }
          }
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int chunk = g0 * 2 + j;                      // 16-byte chunk (8 keys) of the 416-byte P row
            uint4 q = make_uint4(pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                                 pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]));
            *reinterpret_cast<uint4*>(myP + (chunk >> 3) * 16384 + sw128_offset(row, chunk & 7)) = q;
          }
          if (cg == 0) {
            tmem_ld16f(tS + 48, w);
#pragma unroll
            for (int i = 0; i < 16; ++i) { w[i] = ex2_approx(fmaf(w[i], kScaleLog2e, -shift)); sum += w[i]; }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint4 q = make_uint4(pack_bf16x2(w[j * 8 + 0], w[j * 8 + 1]), pack_bf16x2(w[j * 8 + 2], w[j * 8 + 3]),
                                   pack_bf16x2(w[j * 8 + 4], w[j * 8 + 5]), pack_bf16x2(w[j * 8 + 6], w[j * 8 + 7]));
              *reinterpret_cast<uint4*>(myP + sw128_offset(row, 6 + j)) = q;     // keys 48..63: panel 0, chunks 6, 7
            }
          }
          sSum[(t * 128 + row) * 4 + cg] = sum;
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const bool valid = (t * 128 + quad * 32) < kTok;
        uint8_t* myP = sP + t * kPBytes;
        mbar_wait(&o_full[t], ph);             // also orders the other column groups' partial sums before us
        tc_fence_after();
        float o[16];
        if (valid) tmem_ld16f(tmem_base + t * 256 + cg * 16 + lane_sel, o);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_free[t]);   // S_t / O_t columns are free: the next item's Q K^T may start
        if (valid) {
          const float4 s4 = *reinterpret_cast<const float4*>(&sSum[(t * 128 + row) * 4]);
          const float tot = (s4.x + s4.y) + (s4.z + s4.w);
          const float inv = 1.0f / tot;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint4 q = make_uint4(pack_bf16x2(o[j * 8 + 0] * inv, o[j * 8 + 1] * inv), pack_bf16x2(o[j * 8 + 2] * inv, o[j * 8 + 3] * inv),
                                 pack_bf16x2(o[j * 8 + 4] * inv, o[j * 8 + 5] * inv), pack_bf16x2(o[j * 8 + 6] * inv, o[j * 8 + 7] * inv));
            *reinterpret_cast<uint4*>(myP + sw128_offset(row, cg * 2 + j)) = q;     // P_t is dead: reuse its first panel
          }
          fence_proxy_async_smem();
          named_bar_sync(1 + quad, 128);
          if (cg == 0) {
            if (lane == 0) {
              tma_store_3d(&tmCtx, myP + quad * 32 * 128, h * kHd, t * 128 + quad * 32, b);
              tma_store_commit();
            }
            if (lse != nullptr && t * 128 + row < kTok) {
              const float4 m4 = *reinterpret_cast<const float4*>(&sMax[(t * 128 + row) * 4]);
              const float shift = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w)) * kScaleLog2e;
              lse[static_cast<size_t>(item) * kTok + t * 128 + row] = shift + log2f(tot);
            }
          }
        }
      }
    }
    if (cg == 0 && lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kSmWarps + 1) tmem_dealloc(tmem_base, 512);
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
