// Fused transformer block tail, TWO ROW TILES IN FLIGHT per CTA pair (inference path, sm_100a).  Same arithmetic as
// mlp_fused.cuh (timm Block.forward: x + proj(attn), x + mlp(norm2(x)), then the NEXT block's norm1; oracle/vit.py::_Block),
//
//     x_mid = x + ctx . Wproj^T + bp ;  a = LayerNorm2(x_mid) ;  x_out = x_mid + fc2(gelu(fc1(a) + b1)) + b2 ;  ln_out = LayerNorm1'(x_out)
//
// restructured after the clock traces of that kernel: there ONE set of sixteen epilogue warps runs
// GELU x6 -> LayerNorm-on-load of the next tile -> final epilogue back to back, each phase bound by its own latency chain
// (13.5 k + 7.4 k + 7.5 k cycles per tile), while the tensor pipe needs 10.4 k and idles through both LayerNorm phases.
// Here the sixteen warps form TWO GROUPS of eight; group g owns every second tile of the CTA (slot = tile & 1 = g) from its
// LayerNorm-on-load to its final epilogue, and the two groups run half a tile period apart: while group 0 feeds the tensor
// pipe with GELU chunks of tile i, group 1 drains tile i-1 and prepares tile i+1.
//
// What makes two tiles fit:
//   TMEM (512 columns): D1 [0,128) = two fc1 accumulators of 64 columns (the hidden dimension is walked in twelve half-chunks
//         of 64, ping-pong), D2[slot] [128+192*slot, +192) = per-tile accumulator that carries, in turn, the attention
//         projection (read by the LayerNorm-on-load), then the projected residual row x_mid PARKED by tcgen05.st, then fc2
//         accumulating on top of it: the final epilogue reads x_mid + fc2(..) in one piece, so neither x_mid nor a residual
//         re-read touches memory (HBM/L2 traffic per row: x in, ctx in, x out, ln out only).
//   The hidden activation never leaves TMEM: a GELU warp reads its 32 fp32 columns of D1[b] and writes the packed fp16 pairs back
//         over the first 16 of the SAME columns; fc2 takes them from there as its A operand (tcgen05.mma.cta_group::2 with a
//         TMEM A operand: each CTA of the pair supplies its own 128 rows), and the next fc1 into D1[b] is issued right behind
//         that fc2 (tcgen05.mma of one thread execute in order).  No swizzled shared-memory stores / proxy fence per half-chunk.
//   smem: A[slot] 2 x 48 KB (ctx tile, then the normalised rows), one 48 KB staging buffer for ln_out (handed from tile to tile
//         through the stage_free barrier), a 5-stage ring of 12 KB weight stages (one stage = the three K panels of a W1
//         half-chunk, or one K panel of W2 / Wproj).
//   registers: a thread owns 96 columns of a row in the LayerNorm phases; it walks them as two pieces of 48 and re-reads the
//         pieces (from TMEM on load, from its own x_out rows in the final epilogue) for the normalising pass.
// The tensor-pipe program is static and identical in the producer and the issuer (tiles in order; inside tile i, per
// half-chunk q: fc2(q) | fc1(q+2) | [projection of tile i+1 at q = kProjQ]; the first two fc1 of tile i+1 take the place of
// fc1(12), fc1(13)).
//
// Warp roles (576 threads, 96 registers each: ptxas budgets registers for 640 threads): w0 TMA producer, w1 UMMA issuer (leader
// CTA) + TMEM owner, w2..w9 group 0, w10..w17 group 1; inside a group: team = 32-column half of a hidden half-chunk / 96-column
// half of a token row, quad = TMEM lane quadrant.
#pragma once

#include <cuda_fp16.h>

#include "mlp_fused.cuh"

constexpr int kMlp2Threads = (2 + 16) * 32;   // 576

struct Mlp2Smem {
  static constexpr int kABytes = 3 * 16384;                 // per slot: three [128 x 64] K panels
  static constexpr int kStageBytes = 3 * 16384;             // ln_out staging (bf16 [128 x 192] as three swizzled panels), shared by the groups
  static constexpr int kWStage = 12288;                     // this CTA's half of a W2 / Wproj panel [96 x 64], or of three W1 panels [32 x 64]
  static constexpr int kWStages = 5;
  static constexpr int kVecBytes = 768 * 2 + 6 * 192 * 4;   // b1 (fp16); b2, gamma2, beta2, gamma, beta, bp (fp32)
  static constexpr int kPartBytes = 2 * 2 * 2 * 128 * 8;    // LayerNorm partial (sum, sumsq): [use parity][group][team][row]
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = 1024 + 2 * kABytes + kStageBytes + kWStages * kWStage + kVecBytes + kPartBytes + kBarBytes;
};

#ifdef __CUDACC__

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
      "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
      "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
      "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem: this CTA's 128 rows, K-major packed 16-bit pairs, 8 columns per K = 16] * B[smem desc], CTA pair
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHiSw128)
      : "memory");
}

// kTrace: build with the clock64 event log (p.trace); the production instantiation carries no trace code.
// Measured and dropped (A/B in one gpurun call, 191 us either way or slower): the final epilogue's normalising pass re-reading D2
// from TMEM instead of the x_out rows from L2 (D2 released later), and the two 48-column pieces of the LayerNorm passes
// fully unrolled (spills at the 96-register cap).
// kProjQ: the hidden half-chunk (0..11) of tile i at which the projection of tile i+1 is issued (its D2 slot must have been drained by the
// other group's final epilogue of tile i-1 by then, and the LayerNorm-on-load of tile i+1 must fit behind it)
template <int kProjQ, bool kTrace>
__global__ void __launch_bounds__(kMlp2Threads, 1)
mlp_fused2_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                  const __grid_constant__ CUtensorMap tmLn, const __grid_constant__ CUtensorMap tmCtx,
                  const __grid_constant__ CUtensorMap tmWp, const MlpFusedParams p) {
  using L = Mlp2Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                  // [2][kABytes]
  uint8_t* sStage = sA + 2 * L::kABytes;               // [kStageBytes]
  uint8_t* sW = sStage + L::kStageBytes;
  __half* sB1 = reinterpret_cast<__half*>(sW + L::kWStages * L::kWStage);
  float* sB2 = reinterpret_cast<float*>(sB1 + 768);
  float* sGamma2 = sB2 + 192;
  float* sBeta2 = sGamma2 + 192;
  float* sGamma = sBeta2 + 192;
  float* sBeta = sGamma + 192;
  float* sBp = sBeta + 192;
  float2* sPart = reinterpret_cast<float2*>(sBp + 192);          // [2][2][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPart + 2 * 2 * 2 * 128);
  // per slot [2]:
  uint64_t* ctx_full = bars + 0;      // leader: ctx tiles of both CTAs landed in A[slot]
  uint64_t* proj_full = bars + 2;     // projection accumulator complete (commit, both CTAs)
  uint64_t* a_full = bars + 4;        // leader: A operand written (and x_mid parked in D2[slot]) by the group's warps of the pair
  uint64_t* a_empty = bars + 6;       // the tile's last fc1 has read A[slot] (commit)
  uint64_t* d2_full = bars + 8;       // the tile's last fc2 complete (commit)
  uint64_t* d2_empty = bars + 10;     // leader: D2[slot] read out by the final epilogue
  // per slot and D1 buffer [2][2] (index slot * 2 + buffer; each group counts the uses of a buffer by its own tiles):
  uint64_t* d1_full = bars + 12;      // fc1 half-chunk complete in D1[buffer] (commit)
  uint64_t* h_full = bars + 16;       // leader: the hidden half-chunk sits in D1[buffer] as packed fp16 (group's warps of the pair)
  uint64_t* stage_free = bars + 20;   // this CTA: the ln_out staging buffer has been read by the previous tile's TMA stores
  uint64_t* w_full = bars + 21;       // [kWStages]
  uint64_t* w_empty = w_full + L::kWStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + L::kWStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x / 2;
  const int num_clusters = gridDim.x / 2;
  const int num_tiles = (p.M + 127) / 128;
  const int num_units = (num_tiles + 1) / 2;               // a unit = 2 consecutive 128-row tiles, one per CTA of the pair
  const int n_my = (cluster_id < num_units) ? (num_units - cluster_id + num_clusters - 1) / num_clusters : 0;
  auto tile_row0 = [&](int it) { return ((cluster_id + it * num_clusters) * 2 + static_cast<int>(rank)) * 128; };
  constexpr int kGroupArrivals = 8 * 2;                    // eight warps per group, two CTAs

  // ---- one-time setup
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    sB1[i] = __float2half_rn(p.b1[i]);
    if (i < 192) {
      sB2[i] = p.b2[i];
      sGamma2[i] = p.gamma2[i];
      sBeta2[i] = p.beta2[i];
      sGamma[i] = p.has_ln ? p.gamma[i] : 1.0f;
      sBeta[i] = p.has_ln ? p.beta[i] : 0.0f;
      sBp[i] = p.bp[i];
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmLn);
    tma_prefetch_desc(&tmCtx);
    tma_prefetch_desc(&tmWp);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctx_full[s], 1);
      mbar_init(&proj_full[s], 1);
      mbar_init(&a_full[s], kGroupArrivals);
      mbar_init(&a_empty[s], 1);
      mbar_init(&d2_full[s], 1);
      mbar_init(&d2_empty[s], kGroupArrivals);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&d1_full[i], 1);
      mbar_init(&h_full[i], kGroupArrivals);
    }
    mbar_init(stage_free, 4);
    for (int i = 0; i < L::kWStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();
  griddep_launch_dependents();
  // event log: role 0 = UMMA issuer, 1 = first warp of group 0, 2 = first warp of group 1; entry = (tag << 48) | clock
  int trace_n = 0;
  auto trace = [&](int role, int tag) {
    if (kTrace && p.trace != nullptr && blockIdx.x == 0 && trace_n < 512) {
      p.trace[role * 512 + trace_n++] = (static_cast<long long>(tag) << 48) | (clock64() & 0xFFFFFFFFFFFFLL);
    }
  };

  if (warp == 0) {
    // ================================================================= TMA producer (every CTA loads its own share)
    if (lane == 0 && n_my > 0) {
      int ws = 0;
      uint32_t wph = 0;
      auto load_w1 = [&](int q) {        // W1 rows [q*64, q*64+64), all of K: this CTA stages 32 of them, one stage = three K panels
        mbar_wait(&w_empty[ws], wph ^ 1);
        if (leader) mbar_arrive_expect_tx(&w_full[ws], 2 * L::kWStage);
        uint8_t* dst = sW + ws * L::kWStage;
        for (int kp = 0; kp < 3; ++kp)
          tma_load_2d_pair(dst + kp * 4096, &tmW1, mapa_u32(smem_u32(&w_full[ws]), 0), kp * 64, q * 64 + static_cast<int>(rank) * 32);
        if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
      };
      auto load_w2 = [&](int q) {        // W2 columns [q*64, q*64+64) of all 192 rows: this CTA stages 96 rows
        mbar_wait(&w_empty[ws], wph ^ 1);
        if (leader) mbar_arrive_expect_tx(&w_full[ws], 2 * L::kWStage);
        tma_load_2d_pair(sW + ws * L::kWStage, &tmW2, mapa_u32(smem_u32(&w_full[ws]), 0), q * 64, static_cast<int>(rank) * 96);
        if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
      };
      auto load_wp = [&]() {             // Wproj [192 out, 192 in], three K panels, this CTA's 96 rows as three 32-row boxes
        for (int kp = 0; kp < 3; ++kp) {
          mbar_wait(&w_empty[ws], wph ^ 1);
          if (leader) mbar_arrive_expect_tx(&w_full[ws], 2 * L::kWStage);
          uint8_t* dst = sW + ws * L::kWStage;
          for (int j = 0; j < 3; ++j)
            tma_load_2d_pair(dst + j * 4096, &tmWp, mapa_u32(smem_u32(&w_full[ws]), 0), kp * 64, static_cast<int>(rank) * 96 + j * 32);
          if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
        }
      };
      auto load_ctx = [&](int it) {      // attention output rows of tile `it` into A[slot] once the slot's previous tile is through fc1
        const int s = it & 1, k = it >> 1;
        mbar_wait(&a_empty[s], (k & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(&ctx_full[s], 2 * L::kABytes);
        const int m0 = tile_row0(it);
        for (int kp = 0; kp < 3; ++kp)
          tma_load_2d_pair(sA + s * L::kABytes + kp * 16384, &tmCtx, mapa_u32(smem_u32(&ctx_full[s]), 0), kp * 64, m0);
      };
      load_ctx(0);
      load_wp();
      load_w1(0);
      load_w1(1);
      for (int it = 0; it < n_my; ++it) {
        for (int q = 0; q < 12; ++q) {
          // the ctx tile does not go through the ring: its buffer has been free since the last fc1 of tile it-1
          if (q == 0 && it + 1 < n_my) load_ctx(it + 1);
          load_w2(q);
          if (q < 10) load_w1(q + 2);
          else if (it + 1 < n_my) load_w1(q - 10);
          if (q == kProjQ && it + 1 < n_my) load_wp();
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= UMMA issuer (leader CTA only); the whole warp walks
    // the loop with warp-uniform values, one elected lane issues (see mlp_fused.cuh)
    if (leader && n_my > 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(256, 64, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_f16(256, 192);
      constexpr uint32_t idescP = umma_idesc_bf16(256, 192, 0, 0);
      const bool issuer = elect_one();
      int ws = 0;
      uint32_t wph = 0;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA));
      const uint32_t w_lo0 = umma_desc_lo(smem_u32(sW));
      auto commit = [&](uint64_t* bar) {
        if (issuer) umma_commit_pair(bar);
      };
      // one ring stage: NP K panels (4 MMAs of K = 16 each); A panels 16 KB apart, B panels b_step bytes apart inside the stage
      auto stage_mmas = [&](uint32_t d, uint32_t a_lo, uint32_t idesc, int np, uint32_t b_step, bool acc_first) {
        mbar_wait(&w_full[ws], wph);
        tc_fence_after();
        const uint32_t b_lo = w_lo0 + ws * (L::kWStage >> 4);
        if (issuer) {
          for (int kp = 0; kp < np; ++kp) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_split<2>(d, a_lo + kp * (16384 >> 4) + 2 * k, b_lo + kp * (b_step >> 4) + 2 * k, idesc, (acc_first || (kp | k) != 0) ? 1u : 0u);
          }
        }
        __syncwarp();
        commit(&w_empty[ws]);
        if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
      };
      auto fc1 = [&](int it, int q) {     // half-chunk q of tile it -> D1[q & 1] (64 columns)
        const int s = it & 1, b = q & 1;
        stage_mmas(tmem_base + 64 * b, a_lo0 + (s * L::kABytes) / 16, idesc1, 3, 4096, false);
        commit(&d1_full[s * 2 + b]);
        if (q == 11) commit(&a_empty[s]);
      };
      auto fc2 = [&](int it, int q) {     // K panel q of fc2: A = the packed hidden half-chunk in D1[q & 1] (this team layout: hidden
        const int s = it & 1, b = q & 1;  // columns 32t..32t+31 in TMEM columns 64b+32t .. +16), accumulating on the parked residual row
        mbar_wait(&w_full[ws], wph);
        tc_fence_after();
        const uint32_t b_lo = w_lo0 + ws * (L::kWStage >> 4);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ts_pair(tmem_base + 128 + s * 192, tmem_base + 64 * b + 32 * (k >> 1) + 8 * (k & 1), b_lo + 2 * k, idesc2, 1u);
        }
        __syncwarp();
        commit(&w_empty[ws]);
        if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
        if (q == 11) commit(&d2_full[s]);
      };
      auto proj = [&](int it) {
        const int s = it & 1, k = it >> 1;
        mbar_wait(&ctx_full[s], k & 1);
        if (k > 0) mbar_wait(&d2_empty[s], (k - 1) & 1);
        tc_fence_after();
        if (kTrace && lane == 0) trace(0, 7);
        for (int kp = 0; kp < 3; ++kp)
          stage_mmas(tmem_base + 128 + s * 192, a_lo0 + (s * L::kABytes + kp * 16384) / 16, idescP, 1, 0, kp != 0);
        commit(&proj_full[s]);
      };
      proj(0);
      mbar_wait(&a_full[0], 0);
      tc_fence_after();
      fc1(0, 0);
      fc1(0, 1);
      for (int it = 0; it < n_my; ++it) {
        const int s = it & 1, k = it >> 1;
#pragma unroll 1
        for (int q = 0; q < 12; ++q) {
          const int b = q & 1;
          const uint32_t u = static_cast<uint32_t>(k * 6 + (q >> 1));     // use count of buffer b by this slot
          mbar_wait(&h_full[s * 2 + b], u & 1);                           // hidden half-chunk q sits in D1[b]
          tc_fence_after();
          if (kTrace && lane == 0) trace(0, 1);
          fc2(it, q);
          // D1[b] is free again behind fc2(q) (tcgen05.mma of one thread execute in order)
          if (q < 10) {
            fc1(it, q + 2);
          } else if (it + 1 < n_my) {
            if (q == 10) {
              mbar_wait(&a_full[s ^ 1], ((it + 1) >> 1) & 1);
              tc_fence_after();
              if (kTrace && lane == 0) trace(0, 2);
            }
            fc1(it + 1, q - 10);
          }
          if (q == kProjQ && it + 1 < n_my) {
            proj(it + 1);
            if (kTrace && lane == 0) trace(0, 4);
          }
          if (kTrace && lane == 0) trace(0, 6);
        }
      }
    }
  } else {
    // ================================================================= epilogue warps: two groups, one per tile slot
    const int ew = warp - 2;
    const int grp = ew >> 3;                            // == slot of the tiles this group owns
    const int team = (ew >> 2) & 1;
    const int quad = warp & 3;                          // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;                   // token row of the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t bar_id = 2 + grp * 4 + quad;         // the two warps (teams) of this group that share the 32 rows
    const bool tr = kTrace && (ew == grp * 8 && lane == 0);
    const int trole = 1 + grp;
    const uint32_t af_l = mapa_u32(smem_u32(&a_full[grp]), 0);
    const uint32_t hf_l[2] = {mapa_u32(smem_u32(&h_full[grp * 2 + 0]), 0), mapa_u32(smem_u32(&h_full[grp * 2 + 1]), 0)};
    const uint32_t d2e_l = mapa_u32(smem_u32(&d2_empty[grp]), 0);
    uint8_t* sAg = sA + grp * L::kABytes;
    const uint32_t tD1 = tmem_base + lane_sel + 32 * team;
    const uint32_t tD2 = tmem_base + lane_sel + 128 + grp * 192 + 96 * team;
    uint32_t part_use = 0;

    // (sum, sumsq) of this thread's 96 columns -> (mean, rstd) of the row
    auto row_stats = [&](float s, float ss, float& mean, float& rstd) {
      float2* part = sPart + ((part_use & 1) * 2 + grp) * 2 * 128;
      ++part_use;
      part[team * 128 + row] = make_float2(s, ss);
      named_bar_sync(bar_id, 64);
      const float2 o = part[(team ^ 1) * 128 + row];
      const float ts = s + o.x, tss = ss + o.y;
      mean = ts * (1.0f / 192.0f);
      rstd = rsqrtf(fmaxf(tss * (1.0f / 192.0f) - mean * mean, 0.0f) + p.eps);
    };
    // LayerNorm of N8*8 consecutive columns starting at column `col` (a multiple of 8) -> bf16 -> the K-major swizzled
    // panel tile at `dst` (panel = 64 columns = 16 KB); panel_shift re-bases the panel index (ln_out staging, second round)
    auto store_ln = [&](uint8_t* dst, const float* x, int n8, int col, int panel_shift, float mean, float rstd, const float* gam,
                        const float* bet) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (j < n8) {
          const int g = (col >> 3) + j;
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaf((x[j * 8 + e] - mean) * rstd, gam[col + j * 8 + e], bet[col + j * 8 + e]);
          *reinterpret_cast<uint4*>(dst + ((g >> 3) - panel_shift) * 16384 + sw128_offset(row, g & 7)) =
              make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        }
      }
    };
    auto ld48 = [&](uint32_t taddr, float (&x)[48]) {
      float v[32];
      tmem_ld32(taddr, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = v[i];
      float w[16];
      tmem_ld16(taddr + 32, w);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[32 + i] = w[i];
    };
    auto prefetch_rows = [&](int it) {               // this thread's row of tile `it`: one 128-byte line per 8 lanes
      const int grow = tile_row0(it) + row;
      if (grow < p.M && (lane & 7) == 0) {
        const float* xp = p.x_in + xt_offset(grow, 0, 0) + 24 * team * 128;
#pragma unroll
        for (int j = 0; j < 24; ++j) prefetch_l2(xp + j * 128);
      }
    };

    // LayerNorm-on-load of tile `it`: x_mid = x + projection + bp parked in D2[slot], LayerNorm2(x_mid) -> A[slot]
    auto produce_a = [&](int it) {
      const int k = it >> 1;
      const int grow = tile_row0(it) + row;
      const bool valid = grow < p.M;
      const float* src = p.x_in + xt_offset(valid ? grow : 0, 0, 0) + 24 * team * 128;
      if (tr) trace(trole, 30);
      float s = 0.0f, ss = 0.0f;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float x[48];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) r = *reinterpret_cast<const float4*>(src + (12 * h + i) * 128);
          x[i * 4 + 0] = r.x; x[i * 4 + 1] = r.y; x[i * 4 + 2] = r.z; x[i * 4 + 3] = r.w;
        }
        if (h == 0) {
          mbar_wait(&proj_full[grp], k & 1);
          tc_fence_after();
          if (tr) trace(trole, 31);
        }
        const int cb = 96 * team + 48 * h;
        {
          float v[32];
          tmem_ld32(tD2 + 48 * h, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] += v[i] + sBp[cb + i];
        }
        {
          float v[16];
          tmem_ld16(tD2 + 48 * h + 32, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) x[32 + i] += v[i] + sBp[cb + 32 + i];
        }
        {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = x[i];
          tmem_st32(tD2 + 48 * h, v);
        }
        tmem_st16(tD2 + 48 * h + 32, &x[32]);
#pragma unroll
        for (int i = 0; i < 48; ++i) { s += x[i]; ss = fmaf(x[i], x[i], ss); }
      }
      if (tr) trace(trole, 32);
      float mean, rstd;
      row_stats(s, ss, mean, rstd);
      if (tr) trace(trole, 33);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float x[48];
        ld48(tD2 + 48 * h, x);
        store_ln(sAg, x, 6, 96 * team + 48 * h, 0, mean, rstd, sGamma2, sBeta2);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(af_l);
      if (tr) trace(trole, 34);
    };

    // final epilogue of tile `it`: x_out = D2 + b2 (D2 = x_mid + fc2), ln_out = LayerNorm(x_out) through H[slot] and TMA.
    // D2[slot] is released after the FIRST pass (the projection of the slot's next tile is waiting for it); the normalising
    // pass re-reads the rows this thread has just written to x_out (L2 hits)
    auto final_tile = [&](int it) {
      const int k = it >> 1;
      const int m0 = tile_row0(it);
      const int grow = m0 + row;
      const bool valid = grow < p.M;
      float* dstx = p.x_out + xt_offset(valid ? grow : 0, 0, 0) + 24 * team * 128;
      if (tr) trace(trole, 40);
      mbar_wait(&d2_full[grp], k & 1);
      tc_fence_after();
      if (tr) trace(trole, 41);
      float s = 0.0f, ss = 0.0f;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float x[48];
        ld48(tD2 + 48 * h, x);
        const int cb = 96 * team + 48 * h;
#pragma unroll
        for (int i = 0; i < 48; ++i) { x[i] += sB2[cb + i]; s += x[i]; ss = fmaf(x[i], x[i], ss); }
        if (valid) {
#pragma unroll
          for (int i = 0; i < 12; ++i)
            *reinterpret_cast<float4*>(dstx + (12 * h + i) * 128) = make_float4(x[i * 4], x[i * 4 + 1], x[i * 4 + 2], x[i * 4 + 3]);
        }
      }
      auto release_d2 = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(d2e_l);
      };
      release_d2();
      if (tr) trace(trole, 42);
      if (!p.has_ln) return;
      float mean, rstd;
      row_stats(s, ss, mean, rstd);
      if (tr) trace(trole, 43);
      // n4 float4 slots (12 or 8) of this thread's row of x_out, starting at slot f0 of its 24 (L2 hits)
      auto reload = [&](float* x, int f0, int n4) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          if (i < n4) {
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) r = *reinterpret_cast<const float4*>(dstx + (f0 + i) * 128);
            x[i * 4 + 0] = r.x; x[i * 4 + 1] = r.y; x[i * 4 + 2] = r.z; x[i * 4 + 3] = r.w;
          }
        }
      };
      // ln_out leaves through the staging buffer and TMA (three swizzled panels; team 0: columns 0..95, team 1: 96..191), once
      // the previous tile's stores have read it
      mbar_wait(stage_free, (it & 1) ^ 1);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float x[48];
        reload(x, 12 * h, 12);
        store_ln(sStage, x, 6, 96 * team + 48 * h, 0, mean, rstd, sGamma, sBeta);
      }
      fence_proxy_async_smem();
      named_bar_sync(bar_id, 64);
      if (team == 0 && lane == 0) {
        if (m0 + quad * 32 < p.M) {
#pragma unroll
          for (int pn = 0; pn < 3; ++pn) tma_store_2d(&tmLn, sStage + pn * 16384 + quad * 4096, pn * 64, m0 + quad * 32);
        }
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(stage_free);
      }
      __syncwarp();
      if (tr) trace(trole, 44);
    };

    // one hidden half-chunk (64 columns): D1[q & 1] (this team's 32 fp32 columns) -> GELU -> packed fp16 pairs written back over
    // the first 16 of the SAME columns, from where fc2 takes them as its A operand: no shared-memory round trip, no proxy fence
    auto gelu_chunk = [&](int it, int q) {
      const int b = q & 1;
      const uint32_t u = static_cast<uint32_t>((it >> 1) * 6 + (q >> 1));
      if (tr) trace(trole, 10);
      mbar_wait(&d1_full[grp * 2 + b], u & 1);
      tc_fence_after();
      if (tr) trace(trole, 12);
      float v[32];
      tmem_ld32(tD1 + 64 * b, v);
      const uint4* bb = reinterpret_cast<const uint4*>(sB1 + q * 64 + team * 32);
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 bq = bb[j];                              // 8 fp16 biases
        const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __half2 hx = __floats2half2_rn(v[j * 8 + 2 * e], v[j * 8 + 2 * e + 1]);
          hx = __hadd2(hx, *reinterpret_cast<const __half2*>(&bw[e]));
          const __half2 g = gelu_erf_h2(hx);
          o[j * 4 + e] = *reinterpret_cast<const uint32_t*>(&g);
        }
      }
      if (tr) trace(trole, 11);
      tmem_st16u(tD1 + 64 * b, o);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(hf_l[b]);
      if (tr) trace(trole, 13);
    };

    if (grp < n_my) {
      prefetch_rows(grp);
      produce_a(grp);
    }
    for (int it = grp; it < n_my; it += 2) {
      if (it + 2 < n_my) prefetch_rows(it + 2);
#pragma unroll 1
      for (int q = 0; q < 12; ++q) gelu_chunk(it, q);
      final_tile(it);
      if (it + 2 < n_my) produce_a(it + 2);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  // ---- teardown
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

#endif  // __CUDACC__
