// Attention forward for DeiT-Tiny (197 tokens, 3 heads of 64) on tcgen05 / TMEM, sm_100a.
//
// Persistent CTAs walk (image, head) items; an item is two work UNITS (the 128-row query tiles t = 0, 1), and unit u
// lives in TMEM region u & 1 (256 columns each):
//   S = Q_t K^T        UMMA M=128 N=208 K=64 into region columns [0,208)   (Q, K: K-major bf16 tiles loaded by 3-D TMA;
//                      rows >= 197 are zero-filled by the tensor map)
//   P = softmax        SIXTEEN softmax warps: warp (quad, cg) owns the 32 query rows of TMEM lane quadrant `quad` and the
//                      key-column group cg (52 of the 208 padded keys each).  The slice of a score row fits in
//                      registers: TMEM is read once, the exact row maximum and the row sum are combined over the four
//                      column groups through shared memory.  P is written back INTO TMEM as packed bf16 (columns
//                      [104,208) of the region, over scores that every warp already holds in registers) with tcgen05.st ...
//   O = P V            ... and consumed from there: tcgen05.mma with the A operand in TMEM (M=128 N=64 K=208, V MN-major
//                      exactly as it lies in the qkv row) into region columns [0,64).  No shared memory for P.
//   ctx rows           O / rowsum -> bf16, 16 columns (32 bytes) per thread straight to global memory
// The loop is software-pipelined so that the tensor pipe never waits for the softmax warps and vice versa: the issuer
// runs  PV(u), then S(u+2) as soon as O(u) has been read; the softmax warps read O(u-1) in the middle of unit u (after
// their max exchange, when PV(u-1) has long finished), so S(u+1) is always ready when softmax(u) ends.  Q/K/V are
// double-buffered (168 KB) and prefetched a whole item ahead.
// Warp roles (576 threads): w0-15 softmax / epilogue, w16 TMA producer, w17 UMMA issuer (whole warp, uniform datapath).
// Restates timm Attention.forward: softmax(q k^T * 64^-0.5) v (oracle/vit.py::_Attention).
#include "kernels.h"
#include "tma_host.h"

#include <type_traits>

namespace {

constexpr int kTok = 197, kHeads = 3, kHd = 64;
constexpr int kKeysPad = 208;                  // 13 UMMA K-steps of 16 keys
constexpr int kSmWarps = 16;
constexpr int kThreads = (kSmWarps + 2) * 32;  // 576
constexpr int kQBytes = 256 * 128;             // two M tiles
constexpr int kKVBytes = kKeysPad * 128;
constexpr int kXchgBytes = 2 * 2 * 128 * 4 * 4;   // [max | sum][region][row][column group] fp32
constexpr int kSmemBytes = 1024 + 2 * kQBytes + 4 * kKVBytes + kXchgBytes + 256;
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
// 8 packed words (16 bf16 = one UMMA K step) of this thread's row -> 8 TMEM columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* w) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld4f(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* w) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
               "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t w0, uint32_t w1) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(w0), "r"(w1) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
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
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// D[tmem] (+)= A[tmem, K-major bf16 pairs] * B[smem desc]
__device__ __forceinline__ void umma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHiSw128)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse, int num_items, long long* trace_buf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                          // [2][kQBytes]
  uint8_t* sK = sQ + 2 * kQBytes;              // [2][kKVBytes]
  uint8_t* sV = sK + 2 * kKVBytes;             // [2][kKVBytes]
  float* sMax = reinterpret_cast<float*>(sV + 2 * kKVBytes);          // [2][128][4]
  float* sSum = sMax + 2 * 128 * 4;                                    // [2][128][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 2 * 128 * 4);
  uint64_t* qk_full = bars;          // [2] item buffers
  uint64_t* v_full = bars + 2;       // [2]
  uint64_t* qk_empty = bars + 4;     // [2]
  uint64_t* v_empty = bars + 6;      // [2]
  uint64_t* s_full = bars + 8;       // [2] TMEM regions
  uint64_t* p_full = bars + 10;      // [2]
  uint64_t* o_full = bars + 12;      // [2]
  uint64_t* tmem_free = bars + 14;   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int U = 2 * n_items;                    // work units of this CTA
  if (warp == kSmWarps && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&qk_empty[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], kSmWarps);
      mbar_init(&o_full[i], 1);
      mbar_init(&tmem_free[i], kSmWarps);
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
  griddep_wait();                    // (programmatic dependent launch: the setup above overlapped the previous kernel's tail)
  griddep_launch_dependents();
  // debugging: clock64 event log of CTA 0 (role 0 = UMMA issuer, 1 = softmax warp 0); entry = (tag << 48) | clock
  int trace_n = 0;
  auto trace = [&](int role, int tag) {
    if (trace_buf != nullptr && blockIdx.x == 0 && lane == 0 && trace_n < 512)
      trace_buf[role * 512 + trace_n++] = (static_cast<long long>(tag) << 48) | (clock64() & 0xFFFFFFFFFFFFLL);
  };

  if (warp == kSmWarps) {
    // ================================================================= TMA producer (one item ahead)
    if (lane == 0) {
      for (int ii = 0; ii < n_items; ++ii) {
        const int item = blockIdx.x + ii * gridDim.x;
        const int b = item / kHeads, h = item % kHeads;
        const int qb = ii & 1;
        const uint32_t n = static_cast<uint32_t>(ii >> 1);
        mbar_wait(&qk_empty[qb], (n & 1) ^ 1);
        mbar_arrive_expect_tx(&qk_full[qb], kQBytes + kKVBytes);
        tma_load_3d(sQ + qb * kQBytes, &tmQ, &qk_full[qb], h * kHd, 0, b);
        tma_load_3d(sK + qb * kKVBytes, &tmKV, &qk_full[qb], 192 + h * kHd, 0, b);
        mbar_wait(&v_empty[qb], (n & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[qb], kKVBytes);
        tma_load_3d(sV + qb * kKVBytes, &tmKV, &v_full[qb], 384 + h * kHd, 0, b);
      }
    }
  } else if (warp == kSmWarps + 1) {
    // ================================================================= UMMA issuer (whole warp; one elected lane issues)
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKeysPad, 0, 0);   // S = Q K^T
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHd, 0, 1);        // O = P V, V MN-major
    const bool issuer = elect_one();
    const uint32_t q_lo = umma_desc_lo(smem_u32(sQ)), k_lo = umma_desc_lo(smem_u32(sK));
    const uint32_t v_lo = umma_desc_lo(smem_u32(sV), 8192);
    auto issue_s = [&](int u) {
      const int ii = u >> 1, t = u & 1, qb = ii & 1;
      if (t == 0) {
        mbar_wait(&qk_full[qb], static_cast<uint32_t>(ii >> 1) & 1);
        tc_fence_after();
      }
      if (issuer) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_split<1>(tmem_base + t * 256, q_lo + qb * (kQBytes >> 4) + t * (16384 >> 4) + 2 * k,
                            k_lo + qb * (kKVBytes >> 4) + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
        if (t == 1) umma_commit(&qk_empty[qb]);      // Q and K of the item may be overwritten once both S tiles are done
      }
      __syncwarp();
    };
    auto issue_pv = [&](int u) {
      const int ii = u >> 1, t = u & 1, qb = ii & 1;
      if (t == 0) mbar_wait(&v_full[qb], static_cast<uint32_t>(ii >> 1) & 1);
      mbar_wait(&p_full[t], ii & 1);
      tc_fence_after();
      if (issuer) {
#pragma unroll
        for (int j = 0; j < 13; ++j)
          umma_ts_bf16(tmem_base + t * 256, tmem_base + t * 256 + 104 + j * 8, v_lo + qb * (kKVBytes >> 4) + j * (2048 >> 4),
                       idesc_o, j != 0 ? 1u : 0u);
        umma_commit(&o_full[t]);
        if (t == 1) umma_commit(&v_empty[qb]);
      }
      __syncwarp();
    };
    if (U > 0) {
      issue_s(0);
      issue_s(1);
      for (int u = 0; u < U; ++u) {
        issue_pv(u);
        trace(0, 1);
        if (u + 2 < U) {
          mbar_wait(&tmem_free[u & 1], (u >> 1) & 1);     // O(u) has been read: the region may take S(u+2)
          tc_fence_after();
          trace(0, 2);
          issue_s(u + 2);
          trace(0, 3);
        }
      }
    }
  } else {
    // ================================================================= softmax + epilogue warps
    const int quad = warp & 3, cg = warp >> 2;
    const int row = quad * 32 + lane;                 // TMEM lane = query row inside the tile
    const int key0 = 52 * cg;                         // this warp's 52 of the 208 padded keys (13 x 4 per column group)
    // (column group 3 ends with the 11 padded keys 197..207 = its elements 41..51)
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;

    float o[16];                                      // O slice of the previous unit, between its read and its store
    float o_shift = 0.0f;                             // its softmax shift (for the log-sum-exp output)
    // first half of finishing unit `u`: read its O slice (frees the TMEM region for S(u+2))
    auto read_o = [&](int u) {
      const int ii = u >> 1, t = u & 1;
      mbar_wait(&o_full[t], ii & 1);          // also orders the other column groups' partial sums before us
      tc_fence_after();
      if ((t * 128 + quad * 32) < kTok) {
        tmem_ld16f(tmem_base + t * 256 + cg * 16 + lane_sel, o);
        // sMax of this region is rewritten by unit u+2, which cannot start before every warp has arrived on tmem_free below
        const float4 m4 = *reinterpret_cast<const float4*>(&sMax[(t * 128 + row) * 4]);
        o_shift = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w)) * kScaleLog2e;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_free[t]);
    };
    // second half: normalise and store 16 ctx columns per thread straight to global memory (32 bytes per row: no
    // staging, no barriers -- the two named barriers and the TMA-store wait of a staged store cost 1.2k cycles per unit)
    auto store_o = [&](int u) {
      const int ii = u >> 1, t = u & 1;
      const int q = t * 128 + row;
      if (q < kTok) {
        const int item = blockIdx.x + ii * gridDim.x;
        const int b = item / kHeads, h = item % kHeads;
        const float4 s4 = *reinterpret_cast<const float4*>(&sSum[(t * 128 + row) * 4]);
        const float tot = (s4.x + s4.y) + (s4.z + s4.w);
        const float inv = 1.0f / tot;
        uint4* dst = reinterpret_cast<uint4*>(ctx + (static_cast<size_t>(b) * kTok + q) * 192 + h * kHd + cg * 16);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          dst[j] = make_uint4(pack_bf16x2(o[j * 8 + 0] * inv, o[j * 8 + 1] * inv), pack_bf16x2(o[j * 8 + 2] * inv, o[j * 8 + 3] * inv),
                              pack_bf16x2(o[j * 8 + 4] * inv, o[j * 8 + 5] * inv), pack_bf16x2(o[j * 8 + 6] * inv, o[j * 8 + 7] * inv));
        if (lse != nullptr && cg == 0) lse[static_cast<size_t>(item) * kTok + q] = o_shift + log2f(tot);
      }
    };

#pragma unroll 1
    for (int u = 0; u < U; ++u) {
      const int ii = u >> 1, t = u & 1;
      const bool valid = (t * 128 + quad * 32) < kTok;       // uniform over the four warps of a quadrant
      if (warp == 0) trace(1, 10);
      mbar_wait(&s_full[t], ii & 1);
      tc_fence_after();
      if (warp == 0) trace(1, 11);
      float v[52];                                    // this thread's slice of its score row: TMEM is read exactly once
      float shift = 0.0f, sum = 0.0f;
      const uint32_t tS = tmem_base + t * 256 + key0 + lane_sel;
      const uint32_t tP = tmem_base + t * 256 + 104 + 26 * cg + lane_sel;   // packed bf16 pairs: key pair c -> column 104 + c
      if (valid) {
        tmem_ld32(tS, *reinterpret_cast<float(*)[32]>(&v[0]));
        tmem_ld16f(tS + 32, &v[32]);
        tmem_ld4f(tS + 48, &v[48]);
        float m = v[0];
#pragma unroll
        for (int i = 1; i < 52; ++i)
          if (i < 41 || cg != 3) m = fmaxf(m, v[i]);
        // sMax / sSum of this region were last read while unit u-2 was finished (during unit u-1): this quadrant's
        // barrier of that unit orders those reads before these writes
        sMax[(t * 128 + row) * 4 + cg] = m;
        named_bar_sync(1 + quad, 128);           // every column group has its scores in registers and its max published
        const float4 m4 = *reinterpret_cast<const float4*>(&sMax[(t * 128 + row) * 4]);
        shift = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w)) * kScaleLog2e;
        // first 32 exponentials (the MUFU pipe bounds this phase) ...
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i] = ex2_approx(fmaf(v[i], kScaleLog2e, -shift)); sum += v[i]; }
        uint32_t pw[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pw[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
        tmem_st16(tP, pw);
      }
      if (warp == 0) trace(1, 12);
      // ... meanwhile the previous unit's P V product has finished: take its O out of TMEM now, so that S(u+1) is issued
      // while the rest of this unit's exponentials run
      if (u > 0) read_o(u - 1);
      if (warp == 0) trace(1, 13);
      if (valid) {
#pragma unroll
        for (int i = 32; i < 52; ++i) {
          if (i < 41 || cg != 3) { v[i] = ex2_approx(fmaf(v[i], kScaleLog2e, -shift)); sum += v[i]; } else { v[i] = 0.0f; }
        }
        uint32_t pw[10];
#pragma unroll
        for (int e = 0; e < 10; ++e) pw[e] = pack_bf16x2(v[32 + 2 * e], v[32 + 2 * e + 1]);
        tmem_st8(tP + 16, pw);
        tmem_st2(tP + 24, pw[8], pw[9]);
        sSum[(t * 128 + row) * 4 + cg] = sum;
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (u > 0) store_o(u - 1);
      if (warp == 0) trace(1, 14);
    }
    if (U > 0) { read_o(U - 1); store_o(U - 1); }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kSmWarps + 1) tmem_dealloc(tmem_base, 512);
}


// =====================================================================================================================
// Attention BACKWARD on tcgen05 (recompute from the saved log-sum-exp; no atomics, no stored probabilities, every
// (key, query) pair visited ONCE).  Per (image, head) four work units (key tile kt) x (query tile qt), rows = keys:
//     S^T = K_kt Q_qt^T,  dP^T = V_kt dO_qt^T         UMMA 128 x Nq x 64, Nq = 128 (qt = 0) or 80 (qt = 1: queries 128..207)
//     P^T = exp2(S^T c - lse[q]),  dS^T = P^T o (dP^T - delta[q])               sixteen warps, one 32- or 16-column piece each
//     dV_kt += P^T dO_qt                               A operand (bf16 pairs) in TMEM, B = dO tile read MN-major
//     dK_kt += dS^T Q_qt,  dQ_qt += dS K_kt            A = the dS^T tile in SHARED memory, read K-major (M = keys) for dK and
//                                                      MN-major (M = queries) for dQ; B = Q / K tile read MN-major
// with delta[q] = sum_d dO[q,d] O[q,d].  TMEM: S^T [0,128), dP^T [128,256), dV [256,320), dK [320,384), dQ_0 [384,448),
// dQ_1 [448,512) -- all 512 columns.  A warp (quad, cg) owns 32 key rows and a column slice that is a whole number of
// UMMA K steps; it writes the packed P pairs over the very dP columns it has just read (no warp ever overwrites values
// another warp still needs) and the dS values into a swizzled shared-memory tile (double-buffered).  With no operand
// left in the S region, S^T of the NEXT unit is issued right behind dV, and dK / dQ of this unit run while the
// elementwise warps already work on the next one.  Q, K, V, dO tiles are 256 rows (rows >= 197 zero-filled by TMA) so that each serves as M tile,
// N operand and MN-major B operand alike.
constexpr int kBwdTileBytes = 256 * 128;
constexpr int kBwdDsBytes = 2 * 16384;          // dS^T tile: two 64-query panels of [128 key rows x 128 B]
constexpr int kBwdSmemBytes = 1024 + 4 * kBwdTileBytes + 2 * kBwdDsBytes + 4 * 256 * 4 + 256;
// slice of column group cg: qt = 0: 32 columns at 32*cg;  qt = 1: cg 0 -> [0,32), cg 1..3 -> 16 columns at 16 + 16*cg.
// K step j (queries 16j .. 16j+15 of the tile) -> column (inside the dP region) of its packed P pairs
__device__ constexpr int kBwdPCol0[8] = {0, 8, 32, 40, 64, 72, 96, 104};
__device__ constexpr int kBwdPCol1[5] = {0, 8, 32, 48, 64};

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __nv_bfloat16* __restrict__ ctx, const __nv_bfloat16* __restrict__ dctx, const float* __restrict__ lse,
                   __nv_bfloat16* __restrict__ dqkv, int num_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kBwdTileBytes;
  uint8_t* sV = sK + kBwdTileBytes;
  uint8_t* sDO = sV + kBwdTileBytes;
  uint8_t* sDS = sDO + kBwdTileBytes;
  float* sLse = reinterpret_cast<float*>(sDS + 2 * kBwdDsBytes);   // [2][256] per item parity (+inf beyond 197: P = 0 there)
  float* sDelta = sLse + 512;                                      // [2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 512);
  // The four tiles are loaded as HALVES, grouped by the units that read them, so that the next item's halves arrive while this
  // item is still being processed: A = {K, V rows 0..127} (units 0, 1), B = {K, V rows 128..255} (units 2, 3),
  // C = {Q, dO rows 0..127} (units 0, 2), D = {Q, dO rows 128..255} (units 1, 3)
  uint64_t* grp_full = bars;         // [4]
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* o_full = bars + 6;       // the unit's three products have completed (accumulators, dS buffer, and its tile halves)
  uint64_t* kv_read = bars + 7;      // dV / dK of a key tile are in registers
  uint64_t* q_read = bars + 8;       // dQ of the item is in registers
  uint64_t* unit_done = bars + 9;    // [4] unit u of the item has completed (same event as o_full, but one barrier per u, so the
                                     // producer's parity wait stays valid even if it ever lagged more than one unit behind)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  if (warp == kSmWarps && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    for (int i = 0; i < 4; ++i) mbar_init(&grp_full[i], 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, kSmWarps);
    mbar_init(o_full, 1);
    mbar_init(kv_read, kSmWarps);
    mbar_init(q_read, kSmWarps);
    for (int i = 0; i < 4; ++i) mbar_init(&unit_done[i], 1);
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
  constexpr uint32_t kColDP = 128, kColDV = 256, kColDK = 320, kColDQ = 384;
  griddep_wait();                    // (programmatic dependent launch: the setup above overlapped the previous kernel's tail)
  griddep_launch_dependents();

  if (warp == kSmWarps) {
    // ================================================================= TMA producer
    if (lane == 0 && n_items > 0) {
      auto load_grp = [&](int grp, int ii) {
        const int item = blockIdx.x + ii * gridDim.x;
        const int b = item / kHeads, h = item % kHeads;
        const int r0 = (grp & 1) * 128;
        mbar_arrive_expect_tx(&grp_full[grp], 2 * 16384);
        if (grp < 2) {
          tma_load_3d(sK + r0 * 128, &tmQKV, &grp_full[grp], 192 + h * kHd, r0, b);
          tma_load_3d(sV + r0 * 128, &tmQKV, &grp_full[grp], 384 + h * kHd, r0, b);
        } else {
          tma_load_3d(sQ + r0 * 128, &tmQKV, &grp_full[grp], h * kHd, r0, b);
          tma_load_3d(sDO + r0 * 128, &tmDO, &grp_full[grp], h * kHd, r0, b);
        }
      };
      load_grp(0, 0); load_grp(2, 0); load_grp(3, 0); load_grp(1, 0);
      const int total = n_items * 4;
      for (int n = 0; n < total; ++n) {
        const int ii = n >> 2, u = n & 3;
        mbar_wait(&unit_done[u], ii & 1);  // unit n's products have completed: the halves only it (and earlier units) read are free
        if (ii + 1 < n_items) {
          if (u == 1) load_grp(0, ii + 1);
          else if (u == 2) load_grp(2, ii + 1);
          else if (u == 3) { load_grp(1, ii + 1); load_grp(3, ii + 1); }
        }
      }
    }
  } else if (warp == kSmWarps + 1) {
    // ================================================================= UMMA issuer (whole warp; one elected lane issues)
    constexpr uint32_t idesc_s0 = umma_idesc_bf16(128, 128, 0, 0);       // [128 x Nq] = A_tile B^T, both K-major
    constexpr uint32_t idesc_s1 = umma_idesc_bf16(128, 80, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHd, 0, 1);        // [128 x 64] = A B, A K-major (TMEM or smem), B MN-major
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, kHd, 1, 1);        // [128 x 64] = A^T B, both MN-major (dQ)
    const bool issuer = elect_one();
    const uint32_t q_lo = umma_desc_lo(smem_u32(sQ)), k_lo = umma_desc_lo(smem_u32(sK));
    const uint32_t v_lo = umma_desc_lo(smem_u32(sV)), do_lo = umma_desc_lo(smem_u32(sDO));
    const uint32_t q_mn = umma_desc_lo(smem_u32(sQ), 8192), k_mn = umma_desc_lo(smem_u32(sK), 8192);
    const uint32_t do_mn = umma_desc_lo(smem_u32(sDO), 8192);
    const uint32_t ds_k0 = umma_desc_lo(smem_u32(sDS));                  // K-major view (dK): rows = keys
    const uint32_t ds_mn0 = umma_desc_lo(smem_u32(sDS), 16384);          // MN-major view (dQ): two 64-query panels, 16 KB apart
    const int total = n_items * 4;
    // Everything an MMA needs is computed here in warp-uniform code (uniform datapath); the elected lane only executes the
    // tcgen05 instructions.  With the address arithmetic inside the divergent region every MMA costs ~20 dependent
    // instructions of the single issuing thread (fused MLP kernel).
    auto issue_s_dp = [&](uint32_t ak, uint32_t bq, uint32_t av, uint32_t bd, uint32_t idesc_s) {
      if (issuer) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_split<1>(tmem_base, ak + 2 * k, bq + 2 * k, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_split<1>(tmem_base + kColDP, av + 2 * k, bd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    if (n_items > 0) {
      mbar_wait(&grp_full[0], 0);
      mbar_wait(&grp_full[2], 0);
      tc_fence_after();
      issue_s_dp(k_lo, q_lo, v_lo, do_lo, idesc_s0);
    }
#pragma unroll 1
    for (int n = 0; n < total; ++n) {
      const int ii = n >> 2, u = n & 3;
      const int kt = u >> 1, qt = u & 1;
      const uint32_t row0 = static_cast<uint32_t>(qt * 128) * (128 >> 4);     // B tiles: first query row of the tile
      const uint32_t dsb = static_cast<uint32_t>(n & 1) * (kBwdDsBytes >> 4);
      const uint32_t d_dv = tmem_base + kColDV, d_dk = tmem_base + kColDK, d_dq = tmem_base + kColDQ + qt * 64;
      const uint32_t p_base = tmem_base + kColDP;
      const uint32_t do_b = do_mn + row0, q_b = q_mn + row0, k_b = k_mn + kt * (128 * 128 >> 4);
      const uint32_t ds_k = ds_k0 + dsb, ds_m = ds_mn0 + dsb;
      const uint32_t first_kv = qt != 0 ? 1u : 0u, first_q = kt != 0 ? 1u : 0u;   // accumulate flag of a product's first MMA
      const bool same_item = u < 3;             // the next unit works on the tiles that are already in shared memory
      const int kt1 = (u + 1) >> 1, qt1 = (u + 1) & 1;
      const uint32_t ak1 = k_lo + kt1 * (16384 >> 4), bq1 = q_lo + qt1 * (16384 >> 4);
      const uint32_t av1 = v_lo + kt1 * (16384 >> 4), bd1 = do_lo + qt1 * (16384 >> 4);
      const uint32_t idesc_s1n = qt1 == 0 ? idesc_s0 : idesc_s1;
      // the accumulators this unit starts must have been read out by the epilogue of the previous key tile / item
      if (qt == 0 && n >= 2) { mbar_wait(kv_read, ((n >> 1) - 1) & 1); }
      if (u == 0 && ii > 0) { mbar_wait(q_read, (ii - 1) & 1); }
      mbar_wait(p_full, n & 1);
      tc_fence_after();
      if (issuer) {
        // dV += P^T dO   (P pairs sit in the dP region)
        if (qt == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) umma_ts_bf16(d_dv, p_base + kBwdPCol0[j], do_b + j * (2048 >> 4), idesc_o, j != 0 ? 1u : first_kv);
        } else {
#pragma unroll
          for (int j = 0; j < 5; ++j) umma_ts_bf16(d_dv, p_base + kBwdPCol1[j], do_b + j * (2048 >> 4), idesc_o, j != 0 ? 1u : first_kv);
        }
      }
      __syncwarp();
      if (same_item) {
        // the next unit's tile halves (loaded long ago): unit 1 adds D = {Q, dO rows 128..}, unit 2 adds B = {K, V rows 128..}
        if (u == 0) mbar_wait(&grp_full[3], ii & 1);
        else if (u == 1) mbar_wait(&grp_full[1], ii & 1);
        tc_fence_after();
      }
      if (issuer) {
        if (same_item) {
          // next unit's S^T / dP^T right behind dV: tcgen05.mma instructions of one thread execute in issue order
          // (pipelined), so dP^T cannot overwrite the P pairs before dV has read them; the S region holds no operand
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_split<1>(tmem_base, ak1 + 2 * k, bq1 + 2 * k, idesc_s1n, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_split<1>(p_base, av1 + 2 * k, bd1 + 2 * k, idesc_s1n, k != 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        // dK += dS^T Q   (dS^T tile K-major: 16 queries = 32 bytes of a row, 64 per panel)
        if (qt == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_f16_split<1>(d_dk, ds_k + (j >> 2) * (16384 >> 4) + (j & 3) * 2, q_b + j * (2048 >> 4), idesc_o, j != 0 ? 1u : first_kv);
        } else {
#pragma unroll
          for (int j = 0; j < 5; ++j)
            umma_f16_split<1>(d_dk, ds_k + (j >> 2) * (16384 >> 4) + (j & 3) * 2, q_b + j * (2048 >> 4), idesc_o, j != 0 ? 1u : first_kv);
        }
        // dQ_qt += dS K_kt: both operands MN-major, 16 key rows per step
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_f16_split<1>(d_dq, ds_m + k * (2048 >> 4), k_b + k * (2048 >> 4), idesc_q, k != 0 ? 1u : first_q);
        umma_commit(o_full);
        umma_commit(&unit_done[u]);
      }
      __syncwarp();
      if (!same_item && ii + 1 < n_items) {
        // next item: halves A and C of its tiles have been in shared memory since units 1 and 2 of this item finished
        mbar_wait(&grp_full[0], (ii + 1) & 1);
        mbar_wait(&grp_full[2], (ii + 1) & 1);
        tc_fence_after();
        issue_s_dp(k_lo, q_lo, v_lo, do_lo, idesc_s0);
      }
    }
    if (total > 0) mbar_wait(o_full, (total - 1) & 1);
  } else {
    // ================================================================= elementwise + epilogue warps
    const int quad = warp & 3, cg = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel, tB = tmem_base + kColDP + lane_sel;
    const int tid = warp * 32 + lane;                 // 0..511
    auto store16 = [&](__nv_bfloat16* dst, const float* v, float sc) {
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        d4[j] = make_uint4(pack_bf16x2(v[j * 8 + 0] * sc, v[j * 8 + 1] * sc), pack_bf16x2(v[j * 8 + 2] * sc, v[j * 8 + 3] * sc),
                           pack_bf16x2(v[j * 8 + 4] * sc, v[j * 8 + 5] * sc), pack_bf16x2(v[j * 8 + 6] * sc, v[j * 8 + 7] * sc));
    };
    // delta[q] = dO[q,:] . O[q,:] and lse[q] of item `ii` -> shared buffer ii & 1 (two threads per row, straight from global
    // memory: independent of the tile loads, so it runs while the tensor pipe still works on the previous item)
    auto prepass = [&](int ii) {
      const int item = blockIdx.x + ii * gridDim.x;
      const int b = item / kHeads, h = item % kHeads;
      const int r = tid >> 1, half = tid & 1;
      float acc = 0.0f;
      if (r < kTok) {
        const size_t off = (static_cast<size_t>(b) * kTok + r) * 192 + h * kHd + half * 32;
        const uint4* op = reinterpret_cast<const uint4*>(ctx + off);
        const uint4* dp = reinterpret_cast<const uint4*>(dctx + off);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 o4 = op[c], d4 = dp[c];
          const uint32_t ow[4] = {o4.x, o4.y, o4.z, o4.w}, dw[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 of = unpack_bf16x2(ow[e]), df = unpack_bf16x2(dw[e]);
            acc = fmaf(of.x, df.x, acc);
            acc = fmaf(of.y, df.y, acc);
          }
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (half == 0) {
        sDelta[(ii & 1) * 256 + r] = acc;
        sLse[(ii & 1) * 256 + r] = (r < kTok) ? lse[static_cast<size_t>(item) * kTok + r] : INFINITY;
      }
    };
    // accumulators of the key tile that unit m (qt = 1) completed: dV, dK; after the item's last unit also dQ of both query tiles
    auto epilogue = [&](int m) {
      const int ii = m >> 2, kt = (m & 3) >> 1;
      const int item = blockIdx.x + ii * gridDim.x;
      const int b = item / kHeads, h = item % kHeads;
      const int key = kt * 128 + row;
      const bool warp_live = (kt * 128 + quad * 32) < kTok;
      tc_fence_after();
      float o0[16];
      if (warp_live) tmem_ld16f(tmem_base + kColDV + cg * 16 + lane_sel, o0);
      __nv_bfloat16* base = dqkv + (static_cast<size_t>(b) * kTok + (key < kTok ? key : 0)) * 576 + h * kHd + cg * 16;
      if (warp_live && key < kTok) store16(base + 384, o0, 1.0f);
      if (warp_live) tmem_ld16f(tmem_base + kColDK + cg * 16 + lane_sel, o0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(kv_read);
      if (warp_live && key < kTok) store16(base + 192, o0, 0.125f);
      if (kt == 1) {
        // dQ of both query tiles (rows = queries now)
#pragma unroll 1
        for (int t2 = 0; t2 < 2; ++t2) {
          const int q = t2 * 128 + row;
          const bool live_q = (t2 * 128 + quad * 32) < kTok;
          if (live_q) tmem_ld16f(tmem_base + kColDQ + t2 * 64 + cg * 16 + lane_sel, o0);
          if (t2 == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_read);
          }
          if (live_q && q < kTok)
            store16(dqkv + (static_cast<size_t>(b) * kTok + q) * 576 + h * kHd + cg * 16, o0, 0.125f);
        }
      }
    };

    const int total = n_items * 4;
    if (n_items > 0) {
      prepass(0);
      named_bar_sync(1, kSmWarps * 32);
    }
#pragma unroll 1
    for (int n = 0; n < total; ++n) {
      const int ii = n >> 2, u = n & 3;
      const int kt = u >> 1, qt = u & 1;
      const bool warp_live = (kt * 128 + quad * 32) < kTok;   // uniform over the warp: some key row of the warp is real
      // this warp's query columns of the unit: [a, a + W) of the tile
      const int a = (qt == 0) ? 32 * cg : (cg == 0 ? 0 : 16 + 16 * cg);
      const bool wide = (qt == 0) || (cg == 0);
      uint8_t* ds_buf = sDS + (n & 1) * kBwdDsBytes;
      const float* lse_b = sLse + (ii & 1) * 256;
      const float* delta_b = sDelta + (ii & 1) * 256;
      mbar_wait(s_full, n & 1);     // (also: every earlier product but dK / dQ of unit n-1 has completed -> this dS buffer is free)
      tc_fence_after();
      if (warp_live) {
        // P / dS of columns [a, a+W): packed P pairs over dP[a, a+W/2) (columns this warp has just read), dS into the smem tile
        auto piece = [&](auto wtag) {
          constexpr int W = decltype(wtag)::value;
          float sv[W], dv[W];
          if (W == 32) {
            tmem_ld32_nowait(tS + a, *reinterpret_cast<float(*)[32]>(&sv[0]));
            tmem_ld32_nowait(tB + a, *reinterpret_cast<float(*)[32]>(&dv[0]));
          } else {
            tmem_ld16_nowait(tS + a, sv);
            tmem_ld16_nowait(tB + a, dv);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          uint32_t pw[W / 2], dw[W / 2];
          const int q0 = qt * 128 + a;
#pragma unroll
          for (int e = 0; e < W / 4; ++e) {
            const float4 l4 = *reinterpret_cast<const float4*>(&lse_b[q0 + 4 * e]);
            const float4 d4 = *reinterpret_cast<const float4*>(&delta_b[q0 + 4 * e]);
            const float p0 = ex2_approx(fmaf(sv[4 * e + 0], kScaleLog2e, -l4.x));
            const float p1 = ex2_approx(fmaf(sv[4 * e + 1], kScaleLog2e, -l4.y));
            const float p2 = ex2_approx(fmaf(sv[4 * e + 2], kScaleLog2e, -l4.z));
            const float p3 = ex2_approx(fmaf(sv[4 * e + 3], kScaleLog2e, -l4.w));
            pw[2 * e] = pack_bf16x2(p0, p1);
            pw[2 * e + 1] = pack_bf16x2(p2, p3);
            dw[2 * e] = pack_bf16x2(p0 * (dv[4 * e + 0] - d4.x), p1 * (dv[4 * e + 1] - d4.y));
            dw[2 * e + 1] = pack_bf16x2(p2 * (dv[4 * e + 2] - d4.z), p3 * (dv[4 * e + 3] - d4.w));
          }
          if (W == 32) tmem_st16(tB + a, pw); else tmem_st8(tB + a, pw);
          // dS^T tile (this unit's buffer): row = key, 16-byte chunks of 8 queries, panel = 64 queries
#pragma unroll
          for (int c = 0; c < W / 8; ++c) {
            const int qc = a + 8 * c;                 // first query (within the tile) of this chunk
            *reinterpret_cast<uint4*>(ds_buf + (qc >> 6) * 16384 + sw128_offset(row, (qc >> 3) & 7)) =
                make_uint4(dw[4 * c], dw[4 * c + 1], dw[4 * c + 2], dw[4 * c + 3]);
          }
        };
        if (wide) piece(std::integral_constant<int, 32>{});
        else piece(std::integral_constant<int, 16>{});
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else {
        // no real key in these 32 rows: their dS rows must still be finite zeros for the dQ product (K rows are zero, 0 * NaN is not)
        const int W = wide ? 32 : 16;
        for (int c = 0; c < W / 8; ++c) {
          const int qc = a + 8 * c;
          *reinterpret_cast<uint4*>(ds_buf + (qc >> 6) * 16384 + sw128_offset(row, (qc >> 3) & 7)) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      // consume o_full phase by phase (a parity wait must never skip one): dK / dQ of the previous unit ran while this unit was
      // processed and have long finished; the barrier cannot be further than phase n before this warp arrives below
      if (n > 0) mbar_wait(o_full, (n - 1) & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // the key tile the PREVIOUS unit completed is stored now, while the tensor pipe works on this unit's products
      if (n > 0 && qt == 0) epilogue(n - 1);
      if (u == 3 && ii + 1 < n_items) {
        prepass(ii + 1);
        named_bar_sync(1, kSmWarps * 32);
      }
    }
    if (total > 0) {
      mbar_wait(o_full, (total - 1) & 1);
      epilogue(total - 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kSmWarps + 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

static long long* g_attn_trace = nullptr;
void rvk_debug_set_attn_trace_impl(void* buf) { g_attn_trace = static_cast<long long*>(buf); }

int rvk_attention_fwd_launch(const void* qkv, void* ctx, float* lse, int batch, cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  RVK_SET_MAX_SMEM(attn_fwd_tc_kernel, kSmemBytes);
  CUtensorMap tmQ, tmKV;
  RVK_TRY(rvk_make_tmap_3d(&tmQ, qkv, RVK_BF16, 576, kTok, batch, 576, int64_t(kTok) * 576, 64, 256));
  RVK_TRY(rvk_make_tmap_3d(&tmKV, qkv, RVK_BF16, 576, kTok, batch, 576, int64_t(kTok) * 576, 64, kKeysPad));
  const int items = batch * kHeads;
  const int grid = items < kNumSMsB200 ? items : kNumSMsB200;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = rvk_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // algorithmic: QK^T and PV over 197x197x64 per (image, head); qkv read once, ctx written once
  RvkScopedTimer timer(stream, 4.0 * items * 197.0 * 197.0 * 64.0, double(batch) * 197.0 * (576.0 + 192.0) * 2.0, RVK_T_ATTN_FWD);
  RVK_CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_fwd_tc_kernel, tmQ, tmKV, static_cast<__nv_bfloat16*>(ctx), lse, items, g_attn_trace));
  return rvk_launch_check();
}

int rvk_attention_bwd_tc_launch(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv, int batch,
                                cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  RVK_SET_MAX_SMEM(attn_bwd_tc_kernel, kBwdSmemBytes);
  CUtensorMap tmQKV, tmDO;
  RVK_TRY(rvk_make_tmap_3d(&tmQKV, qkv, RVK_BF16, 576, kTok, batch, 576, int64_t(kTok) * 576, 64, 128));
  RVK_TRY(rvk_make_tmap_3d(&tmDO, dctx, RVK_BF16, 192, kTok, batch, 192, int64_t(kTok) * 192, 64, 128));
  const int items = batch * kHeads;
  const int grid = items < kNumSMsB200 ? items : kNumSMsB200;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kBwdSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = rvk_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // algorithmic: S, dP, dV, dK, dQ = five 197x197x64 products per (image, head); qkv + dctx + ctx read, dqkv written
  RvkScopedTimer timer(stream, 10.0 * items * 197.0 * 197.0 * 64.0, double(batch) * 197.0 * (576.0 * 2 + 192.0 * 2) * 2.0, RVK_T_ATTN_BWD);
  RVK_CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_bwd_tc_kernel, tmQKV, tmDO, static_cast<const __nv_bfloat16*>(ctx),
                                  static_cast<const __nv_bfloat16*>(dctx), lse, static_cast<__nv_bfloat16*>(dqkv), items));
  return rvk_launch_check();
}
