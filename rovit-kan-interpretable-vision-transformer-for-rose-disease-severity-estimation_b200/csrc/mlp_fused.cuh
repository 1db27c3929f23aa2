// Fused transformer block tail on tcgen05 / TMEM, sm_100a (inference path of the DeiT-Tiny trunk): the attention output
// projection and the whole MLP half in ONE launch per block,
//
//     x     = x + ctx . Wproj^T + bp                              (has_proj; timm Attention.proj + first residual)
//     a     = LayerNorm2(x) * gamma2 + beta2                      (computed ON LOAD, never written to HBM)
//     x_out = x + fc2( gelu( fc1(a) + b1 ) ) + b2 ;     ln_out = LayerNorm1'(x_out) * gamma1 + beta1  (next block)
//
// (timm Block.forward: x + attn(norm1(x)) [the projection part], x + mlp(norm2(x)), followed by the NEXT block's norm1;
// restated in oracle/vit.py::_Block.)  Neither the normalised input nor the 768-wide hidden activation touches HBM: per
// 128-row tile the epilogue warps read the fp32 token rows, add the projection out of TMEM, normalise the rows into the
// K-major swizzled A operand in shared memory, and the hidden dimension is walked in six chunks of 128 columns,
//     D1[c&1] = a . W1[c]^T          UMMA  M=128*G  N=128  K=192   bf16 x bf16  (TMEM columns   0..255, two buffers)
//     H[c&1]  = f16(gelu(D1 + b1))   epilogue warps: tcgen05.ld -> registers -> K-major swizzled smem operand
//     D2     += H[c&1] . W2[:,c]^T   UMMA  M=128*G  N=192  K=128   f16 x f16    (TMEM columns 256..447)
// and the final epilogue adds bias + residual, writes the fp32 token stream and (through the then idle H buffers
// and TMA) the LayerNorm'ed bf16 operand of the next block's qkv GEMM.
// The projection of the NEXT tile (A = its ctx tile, TMA-loaded into the A buffer as soon as the tile's last fc1 has read
// it; B = Wproj panels from the weight ring) runs as two accumulators -- outputs 0..127 in D1[0] once GELU of chunk 4
// has drained it, outputs 128..191 in the 64 spare TMEM columns 448..511 -- so it is issued before the tile boundary.
// HBM traffic per row: 768 B (x in) + 384 B (ctx in) + 768 B (x out) + 384 B (ln out) [+ 768 B + 768 B through L2 for the
// projected rows]; the unfused proj / fc1 / fc2 GEMMs moved 7.7 KB.
//
// The kernel is bound by the epilogue warps, not by the tensor pipe (ncu + clock traces: profiles/), so GELU
// runs as packed half2 arithmetic (two elements per instruction) and the hidden activation is kept in fp16 (11-bit
// significand: one rounding of 2^-11 instead of bf16's 2^-9; the fp16 polynomial evaluation costs about that
// difference back, see gelu_erf_h2).  fc2 therefore takes an fp16 copy of W2.
//
// G = 2 (default): the two CTAs of a cluster drive ONE tcgen05.mma.cta_group::2 with M = 256 -- each CTA stages
// its own 128 rows of A / H and only HALF of every weight panel, so the weight stream out of L2 and the shared
// memory operand reads per SM are halved against G = 1 (kept as a single-CTA variant for testing).
//
// Warp roles (608 threads): w0 TMA producer (ring of weight panels, ctx tiles), w1 UMMA issuer (leader CTA only) + TMEM
// owner, w2 idle, w3..w18 sixteen epilogue warps = four teams x four TMEM lane quadrants (one token row per
// thread).  Team t owns columns [32t, 32t+32) of every hidden chunk and columns [48t, 48t+48) of the token row in
// the LayerNorm-on-load and in the final epilogue, so all teams carry the same load in every phase.  The fp32
// token stream uses the tiled layout of common.cuh (xt_offset): coalesced 512-byte global accesses straight
// from/to registers.  Per tile the epilogue warps run  GELU x6 -> A operand of the NEXT tile -> final epilogue,
// so the tensor pipe already works on the next tile's fc1 while the final epilogue drains D2.  Launched with
// programmatic stream serialization: everything before griddep_wait() overlaps the previous kernel's tail.
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"

struct MlpFusedParams {
  int M;
  const float* x_in;    // tiled fp32 [M_pad, 192] token stream (LayerNorm2 input and residual)
  float* x_out;         // tiled fp32 (may alias x_in)
  const float* gamma2;  // [192] LayerNorm applied on load (this block's norm2)
  const float* beta2;
  const float* b1;      // [768]
  const float* b2;      // [192]
  const float* gamma;   // [192] next LayerNorm (has_ln)
  const float* beta;
  const float* bp;      // [192] attention output-projection bias (has_proj)
  int has_proj;         // the kernel also does  x += ctx . Wproj^T + bp  (timm Attention.proj + residual) on load
  float eps;
  int has_ln;           // also write ln_out = LayerNorm(x_out) (bf16 [M,192], through the tmLn tensor map)
  long long* trace;     // debugging: optional [4][512] clock64 event log of CTA 0 (nullptr = off)
};

constexpr int kMlpTeams = 4;
constexpr int kMlpEpiWarps = 4 * kMlpTeams;
constexpr int kMlpThreads = (3 + kMlpEpiWarps) * 32;    // 608

template <int G>
struct MlpSmem {
  static constexpr int kABytes = 3 * 16384;                 // A tile: three [128 x 64] K panels
  static constexpr int kHBytes = 2 * 2 * 16384;             // two hidden-chunk buffers of two K panels
  static constexpr int kWStage = 24576 / G;                 // one W2 panel [192/G x 64]; W1 panels [128/G x 64] use 2/3
  static constexpr int kWStages = (G == 2) ? 8 : 4;
  static constexpr int kW1Bytes = 16384 / G;
  static constexpr int kW2Bytes = 24576 / G;
  static constexpr int kVecBytes = 768 * 2 + 6 * 192 * 4;   // b1 (fp16); b2, gamma2, beta2, gamma, beta, bp (fp32)
  static constexpr int kPartBytes = 2 * kMlpTeams * 128 * 8;   // LayerNorm partial (sum, sumsq): [on-load | final][team][row]
  static constexpr int kBarBytes = 512;
  static constexpr int kTotal = 1024 + kABytes + kHBytes + kWStages * kWStage + kVecBytes + kPartBytes + kBarBytes;
};

#ifdef __CUDACC__

// 16 fp32 columns of 32 lanes
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
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

// instruction descriptor for kind::f16 with fp16 A/B (format code 0) and fp32 D
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// Exact-form (erf) GELU on a half2 pair: gelu(x) = relu(x) - a * Phi(-a), a = |x|, Phi(-a) = 2^P(a) with P the
// degree-5 minimax fit of log2(Phi(-a)) on [0, 5.5] (fit error 2.5e-4 relative in Phi; a beyond 5.5 is clamped,
// a * Phi(-a) < 1.1e-7 there).  Evaluated in fp16 the Horner chain is good to ~3e-3 relative on Phi(-a) for
// a <= 3, i.e. an absolute error <= 4e-4 on a term that is at most 0.17 -- of the order of ONE bf16 rounding of a
// typical hidden value, which is what the unfused path spends on storing the activation.
// ex2.approx.f16x2 WITHOUT .ftz maps to two MUFU.EX2.F16; the .ftz form (and cuda_fp16's h2exp2) expands to an
// fp32 round trip of 7 instructions per pair.
__device__ __forceinline__ __half2 gelu_erf_h2(__half2 x) {
  const __half2 t = __hmin2(__habs2(x), __float2half2_rn(5.5f));
  __half2 pl = __hfma2(__float2half2_rn(-2.33248135e-04f), t, __float2half2_rn(4.86966228e-03f));
  pl = __hfma2(pl, t, __float2half2_rn(-4.45980624e-02f));
  pl = __hfma2(pl, t, __float2half2_rn(-4.69684135e-01f));
  pl = __hfma2(pl, t, __float2half2_rn(-1.14625716e+00f));
  pl = __hfma2(pl, t, __float2half2_rn(-1.00036022e+00f));
  __half2 phi;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(*reinterpret_cast<uint32_t*>(&phi)) : "r"(*reinterpret_cast<const uint32_t*>(&pl)));
  return __hfma2(__hneg2(__habs2(x)), phi, __hmax2(x, __float2half2_rn(0.0f)));
}

template <int G>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmLn, const __grid_constant__ CUtensorMap tmCtx,
                 const __grid_constant__ CUtensorMap tmWp, const MlpFusedParams p) {
  using L = MlpSmem<G>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offset (keeps the shared address space visible to the compiler: LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sH = sA + L::kABytes;
  uint8_t* sW = sH + L::kHBytes;
  __half* sB1 = reinterpret_cast<__half*>(sW + L::kWStages * L::kWStage);
  float* sB2 = reinterpret_cast<float*>(sB1 + 768);
  float* sGamma2 = sB2 + 192;
  float* sBeta2 = sGamma2 + 192;
  float* sGamma = sBeta2 + 192;
  float* sBeta = sGamma + 192;
  float* sBp = sBeta + 192;
  float2* sPartA = reinterpret_cast<float2*>(sBp + 192);
  float2* sPartF = sPartA + kMlpTeams * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPartF + kMlpTeams * 128);
  uint64_t* a_full = bars;            // leader: A operand of both CTAs written
  uint64_t* a_empty = bars + 1;       // fc1 of the tile's last chunk done
  uint64_t* d1_full = bars + 2;       // [2]
  uint64_t* gelu_done = bars + 4;     // [2] leader: D1[b] drained and H[b] written by every epilogue warp of the pair
  uint64_t* h_empty = bars + 6;       // [2]
  uint64_t* d2_full = bars + 8;
  uint64_t* d2_empty = bars + 9;      // leader
  uint64_t* ctx_full = bars + 10;     // leader: ctx tiles of both CTAs landed in the A buffer (has_proj)
  uint64_t* proj_full = bars + 11;    // projection accumulator complete (has_proj)
  uint64_t* w_full = bars + 12;       // [kWStages]
  uint64_t* w_empty = w_full + L::kWStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + L::kWStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (G == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x / G;
  const int num_clusters = gridDim.x / G;
  const int num_tiles = (p.M + 127) / 128;
  const int num_units = (num_tiles + G - 1) / G;            // a unit = G consecutive 128-row tiles
  const int n_my = (cluster_id < num_units) ? (num_units - cluster_id + num_clusters - 1) / num_clusters : 0;
  auto tile_row0 = [&](int it) { return ((cluster_id + it * num_clusters) * G + static_cast<int>(rank)) * 128; };

  // ---- one-time setup
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    sB1[i] = __float2half_rn(p.b1[i]);
    if (i < 192) {
      sB2[i] = p.b2[i];
      sGamma2[i] = p.gamma2[i];
      sBeta2[i] = p.beta2[i];
      sGamma[i] = p.has_ln ? p.gamma[i] : 1.0f;
      sBeta[i] = p.has_ln ? p.beta[i] : 0.0f;
      sBp[i] = p.has_proj ? p.bp[i] : 0.0f;
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmLn);
    tma_prefetch_desc(&tmCtx);
    tma_prefetch_desc(&tmWp);
    mbar_init(a_full, kMlpEpiWarps * G);
    mbar_init(a_empty, 1);
    for (int i = 0; i < L::kWStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d1_full[i], 1);
      mbar_init(&gelu_done[i], kMlpEpiWarps * G);
      mbar_init(&h_empty[i], 1);
    }
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, kMlpEpiWarps * G);
    mbar_init(ctx_full, 1);
    mbar_init(proj_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (G == 2) { tmem_alloc_pair(tmem_ptr, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (G == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();                    // the setup above overlapped the previous kernel's tail; its outputs are needed from here on
  griddep_launch_dependents();
  // event log: role 0 = UMMA issuer, 1 = epilogue warp 3; entry = (tag << 48) | clock
  int trace_n = 0;
  auto trace = [&](int role, int tag) {
    if (p.trace != nullptr && blockIdx.x == 0 && trace_n < 512) {
      p.trace[role * 512 + trace_n++] = (static_cast<long long>(tag) << 48) | (clock64() & 0xFFFFFFFFFFFFLL);
    }
  };

  // Weight panels are consumed in this order (producer and issuer walk the same sequence):
  //   [Wp] W1(0) W1(1) | per tile:  c = 0..3: W1(c+2) W2(c);  c = 4: W2(4) [Wp of the next tile];  c = 5: W2(5) [then W1(0) W1(1) of the next tile]
  if (warp == 0) {
    // ================================================================= TMA producer (every CTA loads its own share)
    if (lane == 0 && n_my > 0) {
      int ws = 0;
      uint32_t wph = 0;
      auto load_panel = [&](const CUtensorMap* tm, int bytes, int c0, int c1) {
        mbar_wait(&w_empty[ws], wph ^ 1);
        if (leader) mbar_arrive_expect_tx(&w_full[ws], G * bytes);
        if (G == 2) tma_load_2d_pair(sW + ws * L::kWStage, tm, mapa_u32(smem_u32(&w_full[ws]), 0), c0, c1);
        else tma_load_2d(sW + ws * L::kWStage, tm, &w_full[ws], c0, c1);
        if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
      };
      auto load_w1 = [&](int c) {        // W1 rows [c*128, c*128+128): this CTA stages 128/G of them
        for (int kp = 0; kp < 3; ++kp) load_panel(&tmW1, L::kW1Bytes, kp * 64, c * 128 + static_cast<int>(rank) * (128 / G));
      };
      auto load_w2 = [&](int c) {        // W2 columns [c*128, c*128+128) of all 192 rows: this CTA stages 192/G rows
        for (int kp = 0; kp < 2; ++kp) load_panel(&tmW2, L::kW2Bytes, c * 128 + kp * 64, static_cast<int>(rank) * (192 / G));
      };
      // Wproj [192 out, 192 in], three K panels.  The projection runs as TWO accumulators (see the issuer): outputs 0..127
      // and 128..191, so a stage holds [this CTA's 128/G rows of the first | its 64/G rows of the second], in 32-row boxes
      auto load_wp = [&]() {
        constexpr int nA = 128 / G / 32, nB = 64 / G / 32;
        for (int kp = 0; kp < 3; ++kp) {
          mbar_wait(&w_empty[ws], wph ^ 1);
          if (leader) mbar_arrive_expect_tx(&w_full[ws], G * L::kW2Bytes);
          uint8_t* dst = sW + ws * L::kWStage;
          for (int j = 0; j < nA + nB; ++j) {
            const int r0 = (j < nA) ? static_cast<int>(rank) * (128 / G) + j * 32 : 128 + static_cast<int>(rank) * (64 / G) + (j - nA) * 32;
            if (G == 2) tma_load_2d_pair(dst + j * 4096, &tmWp, mapa_u32(smem_u32(&w_full[ws]), 0), kp * 64, r0);
            else tma_load_2d(dst + j * 4096, &tmWp, &w_full[ws], kp * 64, r0);
          }
          if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
        }
      };
      auto load_ctx = [&](int it) {      // attention output rows of tile `it` into the (free) A buffer
        mbar_wait(a_empty, (it & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(ctx_full, G * L::kABytes);
        const int m0 = tile_row0(it);
        for (int kp = 0; kp < 3; ++kp) {
          if (G == 2) tma_load_2d_pair(sA + kp * 16384, &tmCtx, mapa_u32(smem_u32(ctx_full), 0), kp * 64, m0);
          else tma_load_2d(sA + kp * 16384, &tmCtx, ctx_full, kp * 64, m0);
        }
      };
      if (p.has_proj) { load_ctx(0); load_wp(); }
      load_w1(0);
      load_w1(1);
      for (int it = 0; it < n_my; ++it) {
        for (int c = 0; c < 6; ++c) {
          if (c <= 3) load_w1(c + 2);
          load_w2(c);
          // the A buffer is free once the tile's last fc1 (issued at c = 3) has completed
          if (c == 4 && p.has_proj && it + 1 < n_my) { load_ctx(it + 1); load_wp(); }
        }
        if (it + 1 < n_my) {
          load_w1(0);
          load_w1(1);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= UMMA issuer (leader CTA only)
    // Barriers the peer CTA also arrives on are waited for with plain CTA-scope probes, as CUTLASS' 2-SM pipelines do
    // (a cluster-scope acquire would add an L1 invalidate per probe).
    // The WHOLE warp walks the loop with warp-uniform values (so the compiler keeps addresses, descriptors and phase
    // bits on the uniform datapath); only the tcgen05 instructions themselves are issued by one elected lane.  With
    // per-lane code the issue thread needed ~19 dependent instructions per MMA and ~70 per panel -- more than the
    // 64 cycles an M=256 N=128 K=16 MMA takes -- and the issuer, not the tensor pipe, paced the kernel (clock traces).
    if (leader && n_my > 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(128 * G, 128, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_f16(128 * G, 192);
      const bool issuer = elect_one();
      int ws = 0;
      uint32_t wph = 0;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA));
      const uint32_t h_lo0 = umma_desc_lo(smem_u32(sH));
      const uint32_t w_lo0 = umma_desc_lo(smem_u32(sW));
      auto commit = [&](uint64_t* bar) {
        if (issuer) {
          if (G == 2) umma_commit_pair(bar); else umma_commit(bar);
        }
      };
      // one weight panel = 4 MMAs of K = 16; a_lo: descriptor low word of the A panel
      auto panel_mmas = [&](uint32_t d, uint32_t a_lo, uint32_t idesc, bool first_acc) {
        mbar_wait(&w_full[ws], wph);
        tc_fence_after();
        const uint32_t b_lo = w_lo0 + ws * (L::kWStage >> 4);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_split<G>(d, a_lo + 2 * k, b_lo + 2 * k, idesc, (first_acc || k != 0) ? 1u : 0u);
        }
        __syncwarp();
        commit(&w_empty[ws]);
        if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
      };
      auto fc1 = [&](int c) {
        const int b = c & 1;
        const uint32_t d = tmem_base + b * 128;
#pragma unroll
        for (int kp = 0; kp < 3; ++kp) panel_mmas(d, a_lo0 + kp * (16384 >> 4), idesc1, kp != 0);
        commit(&d1_full[b]);
        if (c == 5) commit(a_empty);
      };
      auto fc2 = [&](int c) {
        const int b = c & 1;
        const uint32_t d = tmem_base + 256;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) panel_mmas(d, h_lo0 + (b * 32768 + kp * 16384) / 16, idesc2, (c | kp) != 0);
        commit(&h_empty[b]);
        if (c == 5) commit(d2_full);
      };
      // attention output projection of tile `it`, A = ctx tile in the A buffer, as two accumulators so that it can be
      // issued BEFORE the tile boundary: outputs 0..127 -> D1[0] (idle once GELU of chunk 4 has drained it), outputs
      // 128..191 -> the 64 TMEM columns 448..511 nobody else uses.  By the time the epilogue warps have finished GELU(5)
      // the projection is complete, so the next tile's LayerNorm-on-load never waits for the tensor pipe.
      auto proj = [&](int it) {
        constexpr uint32_t idescA = umma_idesc_bf16(128 * G, 128, 0, 0);
        constexpr uint32_t idescB = umma_idesc_bf16(128 * G, 64, 0, 0);
        mbar_wait(ctx_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kp = 0; kp < 3; ++kp) {
          mbar_wait(&w_full[ws], wph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + kp * (16384 >> 4);
          const uint32_t b_lo = w_lo0 + ws * (L::kWStage >> 4);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_split<G>(tmem_base, a_lo + 2 * k, b_lo + 2 * k, idescA, (kp | k) != 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_split<G>(tmem_base + 448, a_lo + 2 * k, b_lo + ((128 / G) * 128 >> 4) + 2 * k, idescB, (kp | k) != 0 ? 1u : 0u);
          }
          __syncwarp();
          commit(&w_empty[ws]);
          if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
        }
        commit(proj_full);
      };
      if (p.has_proj) proj(0);
      mbar_wait(a_full, 0);
      tc_fence_after();
      fc1(0);
      fc1(1);
      for (int it = 0; it < n_my; ++it) {
#pragma unroll 1
        for (int c = 0; c < 6; ++c) {
          const int q = it * 6 + c;
          mbar_wait(&gelu_done[q & 1], (q >> 1) & 1);
          tc_fence_after();
          if (lane == 0) trace(0, 1);
          if (c <= 3) fc1(c + 2);
          if (c == 0 && it > 0) {
            mbar_wait(d2_empty, (it - 1) & 1);
            tc_fence_after();
          }
          fc2(c);
          if (c == 4 && p.has_proj && it + 1 < n_my) proj(it + 1);
          if (lane == 0) trace(0, 3);
        }
        if (it + 1 < n_my) {
          mbar_wait(a_full, (it + 1) & 1);
          tc_fence_after();
          fc1(0);
          fc1(1);
        }
      }
    }
  } else if (warp >= 3) {
    // ================================================================= epilogue warps
    const int quad = warp & 3;
    const int team = (warp - 3) >> 2;
    const int row = quad * 32 + lane;                   // token row of the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const bool tr = (warp == 3 && lane == 0);
    const uint32_t gd_l[2] = {(G == 2) ? mapa_u32(smem_u32(&gelu_done[0]), 0) : 0u,
                              (G == 2) ? mapa_u32(smem_u32(&gelu_done[1]), 0) : 0u};
    const uint32_t d2e_l = (G == 2) ? mapa_u32(smem_u32(d2_empty), 0) : 0u;
    const uint32_t af_l = (G == 2) ? mapa_u32(smem_u32(a_full), 0) : 0u;

    // this thread's 48 columns [48*team, 48*team+48) of token row `grow` = float4 slots f = 12*team + i of the tiled stream
    auto load_row48_from = [&](const float* xbase, int grow, float (&x)[48]) {
      const bool valid = grow < p.M;
      const float* src = xbase + xt_offset(valid ? grow : 0, 0, 0) + 12 * team * 128;
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) r = *reinterpret_cast<const float4*>(src + i * 128);
        x[i * 4 + 0] = r.x; x[i * 4 + 1] = r.y; x[i * 4 + 2] = r.z; x[i * 4 + 3] = r.w;
      }
    };
    auto load_row48 = [&](int grow, float (&x)[48]) { load_row48_from(p.x_in, grow, x); };
    auto prefetch_row48 = [&](int grow) {       // one 128-byte line per 8 lanes
      if (grow < p.M && (lane & 7) == 0) {
        const float* xp = p.x_in + xt_offset(grow, 0, 0) + 12 * team * 128;
#pragma unroll
        for (int j = 0; j < 12; ++j) prefetch_l2(xp + j * 128);
      }
    };
    // row statistics over the four teams that share a row; (sum, sumsq) -> (mean, rstd)
    auto row_stats = [&](float2* part, const float (&x)[48], float& mean, float& rstd) {
      float s = 0.0f, ss = 0.0f;
#pragma unroll
      for (int i = 0; i < 48; ++i) { s += x[i]; ss = fmaf(x[i], x[i], ss); }
      part[team * 128 + row] = make_float2(s, ss);
      named_bar_sync(2 + quad, 32 * kMlpTeams);
      float ts = 0.0f, tss = 0.0f;
#pragma unroll
      for (int t = 0; t < kMlpTeams; ++t) { const float2 v = part[t * 128 + row]; ts += v.x; tss += v.y; }
      mean = ts * (1.0f / 192.0f);
      rstd = rsqrtf(fmaxf(tss * (1.0f / 192.0f) - mean * mean, 0.0f) + p.eps);
    };
    // LayerNorm of this thread's 48 columns -> bf16 -> three-panel K-major swizzled tile at `dst` (48 KB)
    auto store_ln48 = [&](uint8_t* dst, const float (&x)[48], float mean, float rstd, const float* gam, const float* bet) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int g = 6 * team + j;                       // 16-byte chunk of the 384-byte bf16 row
        const int col = g * 8;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf((x[j * 8 + e] - mean) * rstd, gam[col + e], bet[col + e]);
        *reinterpret_cast<uint4*>(dst + (g >> 3) * 16384 + sw128_offset(row, g & 7)) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
    };

    // A operand of tile `it`: LayerNorm2 of the token rows, straight into shared memory
    auto produce_a = [&](int it) {
      float x[48];
      const int grow = tile_row0(it) + row;
      load_row48(grow, x);
      if (tr) trace(1, 30);
      if (p.has_proj) {
        // x += ctx . Wproj^T + bp: the accumulator sits in the idle D1 columns; the new row goes back to the token stream
        // (the final epilogue re-reads it as the MLP residual) and feeds LayerNorm2 below
        mbar_wait(proj_full, it & 1);
        tc_fence_after();
        if (tr) trace(1, 31);
        // columns [48*team, 48*team+48) of the projection: outputs 0..127 sit in TMEM columns 0..127, outputs 128..191 in 448..511
        const uint32_t tP0 = tmem_base + lane_sel + (team < 3 ? team * 48 : 448 + 16);
        const uint32_t tP1 = tmem_base + lane_sel + (team < 2 ? team * 48 + 32 : team == 2 ? 448 : 448 + 48);
        {
          float v[32];
          tmem_ld32(tP0, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] += v[i] + sBp[team * 48 + i];
        }
        {
          float v[16];
          tmem_ld16(tP1, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) x[32 + i] += v[i] + sBp[team * 48 + 32 + i];
        }
        tc_fence_before();
        if (grow < p.M) {
          float* dstx = p.x_out + xt_offset(grow, 0, 0) + 12 * team * 128;
#pragma unroll
          for (int i = 0; i < 12; ++i)
            *reinterpret_cast<float4*>(dstx + i * 128) = make_float4(x[i * 4], x[i * 4 + 1], x[i * 4 + 2], x[i * 4 + 3]);
        }
      } else {
        mbar_wait(a_empty, (it & 1) ^ 1);         // fc1 of the previous tile has read the buffer
      }
      if (tr) trace(1, 32);
      float mean, rstd;
      row_stats(sPartA, x, mean, rstd);
      if (tr) trace(1, 33);
      store_ln48(sA, x, mean, rstd, sGamma2, sBeta2);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (G == 2) mbar_arrive_cluster(af_l); else mbar_arrive(a_full);
      }
    };

    // final epilogue of tile `it`
    auto final_tile = [&](int it) {
      const int m0 = tile_row0(it);
      const int grow = m0 + row;
      const bool valid = grow < p.M;
      float x[48];
      load_row48_from(p.has_proj ? p.x_out : p.x_in, grow, x);     // residual (L2 hit: the rows produce_a read / wrote)
      if (tr) trace(1, 40);
      mbar_wait(d2_full, it & 1);
      tc_fence_after();
      if (tr) trace(1, 41);
      const uint32_t tD2 = tmem_base + 256 + team * 48 + lane_sel;
      {
        float v[32];
        tmem_ld32(tD2, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] += v[i] + sB2[team * 48 + i];
      }
      {
        float v[16];
        tmem_ld16(tD2 + 32, v);
        tc_fence_before();                       // D2 is in registers: the next tile's fc2 may overwrite it
        __syncwarp();
        if (lane == 0) {
          if (G == 2) mbar_arrive_cluster(d2e_l); else mbar_arrive(d2_empty);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) x[32 + i] += v[i] + sB2[team * 48 + 32 + i];
      }
      if (tr) trace(1, 42);
      if (valid) {
        float* dstx = p.x_out + xt_offset(grow, 0, 0) + 12 * team * 128;
#pragma unroll
        for (int i = 0; i < 12; ++i)
          *reinterpret_cast<float4*>(dstx + i * 128) = make_float4(x[i * 4], x[i * 4 + 1], x[i * 4 + 2], x[i * 4 + 3]);
      }
      if (tr) trace(1, 43);
      if (p.has_ln) {
        float mean, rstd;
        row_stats(sPartF, x, mean, rstd);
        if (tr) trace(1, 44);
        // both H buffers are idle between the last fc2 of this tile (d2_full) and the next tile's first GELU chunk
        store_ln48(sH, x, mean, rstd, sGamma, sBeta);
        fence_proxy_async_smem();
        if (tr) trace(1, 45);
        named_bar_sync(2 + quad, 32 * kMlpTeams);
        if (team == 0) {
          if (lane == 0) {
            if (m0 + quad * 32 < p.M) {
#pragma unroll
              for (int pn = 0; pn < 3; ++pn) tma_store_2d(&tmLn, sH + pn * 16384 + quad * 4096, pn * 64, m0 + quad * 32);
            }
            tma_store_commit();
            tma_store_wait_read<0>();            // H may be overwritten by the next GELU chunk
          }
          __syncwarp();
        }
        named_bar_sync(2 + quad, 32 * kMlpTeams);
      }
    };

    if (n_my > 0) {
      prefetch_row48(tile_row0(0) + row);
      produce_a(0);
    }
    for (int it = 0; it < n_my; ++it) {
      if (it + 1 < n_my) prefetch_row48(tile_row0(it + 1) + row);     // read again ~6 chunks later
#pragma unroll 1
      for (int c = 0; c < 6; ++c) {
        const int q = it * 6 + c, b = q & 1;
        const uint32_t n = static_cast<uint32_t>(q >> 1);
        if (tr) trace(1, 10);
        mbar_wait(&d1_full[b], n & 1);
        tc_fence_after();
        if (tr) trace(1, 12);
        {
          float v[32];
          tmem_ld32(tmem_base + b * 128 + team * 32 + lane_sel, v);
          const uint4* bb = reinterpret_cast<const uint4*>(sB1 + c * 128 + team * 32);
          uint8_t* panel = sH + b * 32768 + (team >> 1) * 16384;
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
          // the hidden-chunk buffer is only needed now: fc2 of chunk c-2 (its last reader) had the whole GELU to finish
          mbar_wait(&h_empty[b], (n & 1) ^ 1);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(panel + sw128_offset(row, (team & 1) * 4 + j)) =
                make_uint4(o[j * 4], o[j * 4 + 1], o[j * 4 + 2], o[j * 4 + 3]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (G == 2) mbar_arrive_cluster(gd_l[b]); else mbar_arrive(&gelu_done[b]);
        }
        if (tr) trace(1, 13);
      }
      if (it + 1 < n_my) produce_a(it + 1);
      if (tr) trace(1, 20);
      final_tile(it);
      if (tr) trace(1, 14);
    }
    if (warp >= 3 && team == 0 && lane == 0) tma_store_wait_all<0>();
  }

  // ---- teardown
  tc_fence_before();
  if (G == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if (G == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

#endif  // __CUDACC__
