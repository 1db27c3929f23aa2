// Persistent tcgen05/TMEM GEMM  C[M,N] = A[M,K] * B[N,K]^T  with fused epilogues, sm_100a.
//
//   A  activations  bf16 row-major [M, K]   (K-major UMMA operand, TMA box 64 x 128, 128B swizzle)
//   B  weights      bf16 row-major [N, K]   (= torch Linear.weight layout; TMA box 64 x BN)
//   D  fp32 accumulators in TMEM, two stages of 256 columns so the epilogue of tile t overlaps
//      the MMAs of tile t+1.
//
// Warp roles (480 threads): w0 operand TMA producer, w1 UMMA issuer (one elected lane) + TMEM
// owner, w2 epilogue-panel producer (TMA loads of residual / pre-activation panels), w3..w14
// epilogue: three teams of four warps (TMEM lane quadrant = warp_id % 4, one accumulator row per
// thread); the teams take the 128-byte column panels round-robin, so every SM sub-partition has
// three epilogue warps to overlap TMEM loads, MUFU latency and shared-memory traffic (the epilogues
// are bound by instruction issue, not by the tensor pipe: ncu in profiles/).
//
// All epilogue I/O moves through a ring of six [128 rows x 128 B] shared-memory panels in the TMA
// 128-byte swizzle: auxiliary inputs arrive by TMA load (up to four panels ahead of their use), are
// rewritten in place by the row owner, and leave by TMA store (tail rows are clipped by the tensor
// map).  Each epilogue warp owns its 32 rows end to end -- it stores its own [32 x 128 B] sub-panel
// and retires it with its own bulk-group -- so the epilogue has no CTA-wide barrier.
//
// Epilogue modes (the encoder's fused ops):
//   EPI_BF16     out = bf16(acc + bias)                               qkv, generic dgrad
//   EPI_GELU     h = bf16(gelu(acc + bias)), optionally z = bf16(acc + bias)     fc1
//   EPI_DGELU    out = bf16(acc * gelu'(z))                           dgrad through fc2's input
//   EPI_F32      out = fp32(acc + bias)
//   EPI_RES_LN   x' = acc + bias + residual (fp32, BN = 192 = full row); optionally
//                y = bf16(LayerNorm(x') * gamma + beta) and per-row mean / rstd.
//                The residual is a TMA-loaded tile of the token stream, or a per-position table
//                (row % table_rows) for the patch-embedding GEMM (cls/pos/bias folded in).
#pragma once

#include "common.cuh"

enum GemmEpilogue : int { EPI_BF16 = 0, EPI_GELU = 1, EPI_DGELU = 2, EPI_F32 = 3, EPI_RES_LN = 4 };

struct GemmNtParams {
  int M, N, K;
  const float* bias;        // [N] or nullptr
  const float* gamma;       // [192] (EPI_RES_LN with LN)
  const float* beta;        // [192]
  const float* res_table;   // [table_rows, 192] or nullptr (then residual comes by TMA)
  int table_rows;
  float ln_eps;
  float* mean_out;          // [M] or nullptr
  float* rstd_out;          // [M] or nullptr
  int has_out2;             // EPI_GELU: also store z;  EPI_RES_LN: also store LN output
  int has_res;              // EPI_RES_LN: 0 = no residual at all
  // EPI_RES_LN, inference path: the fp32 token stream in the tiled layout of common.cuh (xt_offset).  With
  // out_tiled set, x' is written straight from registers (coalesced, no smem panels / TMA) and the residual
  // comes from res_tiled (or the table); `out` / `aux` are then unused.
  const float* res_tiled;
  float* out_tiled;
};

constexpr int kGemmThreads = 480;
constexpr int kNumTeams = 3;
constexpr int kPanelBytes = 128 * 128;   // 16 KB

template <int BN, int STAGES>
struct GemmNtSmem {
  static constexpr int kABytes = 128 * 64 * 2;
  static constexpr int kBBytes = BN * 64 * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOperandBytes = STAGES * kStageBytes;
  // epilogue staging slots (16 KB panels): 6 next to a 3-stage operand ring; a 2-stage ring (K = 192 needs little operand
  // prefetch) leaves room for 8 -- used by the aux-loading DGELU mode, whose slots are held from the z load to the dz store
  static constexpr int kNumSlots = STAGES >= 3 ? 6 : 8;
  static constexpr int kSlotBytes = kNumSlots * kPanelBytes;
  static constexpr int kVecBytes = (768 + 192 + 192) * 4;   // bias / gamma / beta
  static constexpr int kBarBytes = 256;
  static constexpr int kPartBytes = 2 * kNumTeams * 128 * 4;   // LayerNorm partial sums [pass][team][row]
  static constexpr int kTotal = 1024 /*align slack*/ + kOperandBytes + kSlotBytes + kVecBytes + kBarBytes + kPartBytes;
};

#ifdef __CUDACC__

// number of epilogue panels per output tile and whether panel `i` needs a TMA aux load
template <int BN, int MODE>
__device__ __forceinline__ int panels_per_tile(const GemmNtParams& p) {
  if (MODE == EPI_BF16 || MODE == EPI_DGELU) return BN / 64;
  if (MODE == EPI_GELU) return (BN / 64) * (p.has_out2 ? 2 : 1);
  if (MODE == EPI_F32) return BN / 32;
  return (p.out_tiled != nullptr ? 0 : 6) + (p.has_out2 ? 3 : 0);   // EPI_RES_LN
}
template <int BN, int MODE>
__device__ __forceinline__ bool panel_has_aux(const GemmNtParams& p, int i) {
  if (MODE == EPI_DGELU) return true;
  if (MODE == EPI_RES_LN) return p.out_tiled == nullptr && i < 6 && p.has_res && p.res_table == nullptr;
  return false;
}

template <int BN, int MODE, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
               const __grid_constant__ CUtensorMap tmAux, const GemmNtParams p) {
  using L = GemmNtSmem<BN, STAGES>;
  static_assert(BN % 64 == 0 && BN <= 256, "BN must be a multiple of 64 up to 256");
  static_assert(MODE != EPI_RES_LN || BN == 192, "fused LayerNorm needs the whole 192-wide row in one tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset form keeps the shared address space (LDS/STS)
  uint8_t* sOperands = smem;
  uint8_t* sSlots = smem + L::kOperandBytes;
  float* sBias = reinterpret_cast<float*>(sSlots + L::kSlotBytes);
  float* sGamma = sBias + 768;
  float* sBeta = sGamma + 192;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBeta + 192);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = full + STAGES;             // [STAGES]
  uint64_t* tmem_full = empty + STAGES;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint64_t* slot_full = tmem_empty + 2;        // [L::kNumSlots]
  uint64_t* slot_empty = slot_full + L::kNumSlots;  // [L::kNumSlots]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(slot_empty + L::kNumSlots);
  float* sPart = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + L::kBarBytes);   // [2][2][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m_tiles = (p.M + 127) / 128;
  const int num_n_tiles = p.N / BN;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int num_kb = p.K / 64;

  // ---- one-time setup
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    sBias[i] = (p.bias != nullptr && i < p.N) ? p.bias[i] : 0.0f;
    if (i < 192) {
      sGamma[i] = (p.gamma != nullptr) ? p.gamma[i] : 1.0f;
      sBeta[i] = (p.beta != nullptr) ? p.beta[i] : 0.0f;
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128 * kNumTeams);
    }
    for (int i = 0; i < L::kNumSlots; ++i) {
      mbar_init(&slot_full[i], 1);
      mbar_init(&slot_empty[i], 4);     // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();                    // (programmatic dependent launch: the setup above overlapped the previous kernel's tail)
  griddep_launch_dependents();

  if (warp == 0) {
    // ================================================================= operand producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / num_n_tiles) * 128;
        const int n0 = (t % num_n_tiles) * BN;
        // (an L2 prefetch of the next tile's A rows from here measured 2-8 % SLOWER in every epilogue mode: the loads are
        // not what the tiles wait for)
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], L::kStageBytes);
          uint8_t* a = sOperands + s * L::kStageBytes;
          tma_load_2d(a, &tmA, &full[s], kb * 64, m0);
          tma_load_2d(a + L::kABytes, &tmB, &full[s], kb * 64, n0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= UMMA issuer
    // whole warp, warp-uniform values (descriptor words stay on the uniform datapath); one elected lane issues
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      const bool issuer = elect_one();
      const uint32_t op_lo = umma_desc_lo(smem_u32(sOperands));
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_lo = op_lo + s * (L::kStageBytes >> 4);
          const uint32_t b_lo = a_lo + (L::kABytes >> 4);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_split<1>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty[s]);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (issuer) umma_commit(&tmem_full[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ================================================================= epilogue-panel producer
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / num_n_tiles) * 128;
        const int n0 = (t % num_n_tiles) * BN;
        const int np = panels_per_tile<BN, MODE>(p);
        if constexpr (MODE == EPI_DGELU) {
          // A slot is held from the aux load until its output store has drained, so the loads cannot run far ahead of the
          // epilogue: with the z panels coming from HBM the epilogue warps spent 23 % of their samples waiting for slot_full
          // (ncu source view).  Pull the panels of the tile after the next one into L2 now; the slot load then hits L2.
          const int t2 = t + 2 * static_cast<int>(gridDim.x);
          if (t2 < num_tiles) {
            for (int i = 0; i < np; ++i)
              tma_prefetch_2d(&tmAux, (t2 % num_n_tiles) * BN + i * 64, (t2 / num_n_tiles) * 128);
          }
        }
        for (int i = 0; i < np; ++i, ++cnt) {
          const int slot = cnt % L::kNumSlots;
          const uint32_t par = (cnt / L::kNumSlots) & 1;
          mbar_wait(&slot_empty[slot], par ^ 1);
          if (panel_has_aux<BN, MODE>(p, i)) {
            mbar_arrive_expect_tx(&slot_full[slot], kPanelBytes);
            const int c0 = (MODE == EPI_DGELU) ? n0 + i * 64 : n0 + i * 32;
            tma_load_2d(sSlots + slot * kPanelBytes, &tmAux, &slot_full[slot], c0, m0);
          } else {
            mbar_arrive(&slot_full[slot]);
          }
        }
      }
    }
  } else {
    // ================================================================= epilogue warps
    const int quad = warp & 3;
    const int team = (warp - 3) >> 2;                 // 0: warps 3-6, 1: warps 7-10, 2: warps 11-14
    const int row = quad * 32 + lane;                 // accumulator row == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    uint32_t cnt = 0;                                 // global panel counter (all panels, all teams)
    int tile_iter = 0;
    int prev_slot = -1;
    int acc = 0;
    uint32_t acc_ph = 0;

    // panels go round-robin over the teams; `cnt` walks the CTA's whole panel sequence
    auto mine = [&]() -> bool { return cnt % kNumTeams == static_cast<uint32_t>(team); };
    auto acquire = [&](int& slot) -> uint8_t* {
      slot = cnt % L::kNumSlots;
      mbar_wait(&slot_full[slot], (cnt / L::kNumSlots) & 1);
      ++cnt;
      return sSlots + slot * kPanelBytes;
    };
    // this warp's 32 rows of the panel are final: store them, then retire the previous panel's sub-store
    // (one panel of slack keeps the wait off the critical path) and hand that slot back to the producer
    auto publish = [&](const CUtensorMap* tm, int slot, int c0, int r0) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (r0 + quad * 32 < p.M)
          tma_store_2d(tm, sSlots + slot * kPanelBytes + quad * 32 * 128, c0, r0 + quad * 32);
        tma_store_commit();
        if (prev_slot >= 0) {
          tma_store_wait_read<1>();
          mbar_arrive(&slot_empty[prev_slot]);
        }
      }
      prev_slot = slot;
    };
    // the warps that share a TMEM quadrant (one per team) meet here
    auto quad_sync = [&]() {
      tc_fence_before();
      named_bar_sync(2 + quad, 32 * kNumTeams);
      tc_fence_after();
    };
    // 32 accumulator columns -> registers, plus bias
    auto load32 = [&](uint32_t taddr, int col0, float (&v)[32]) {
      tmem_ld32(taddr, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = *reinterpret_cast<const float4*>(&sBias[col0 + i * 4]);
        v[i * 4 + 0] += b4.x; v[i * 4 + 1] += b4.y; v[i * 4 + 2] += b4.z; v[i * 4 + 3] += b4.w;
      }
    };
    // 32 fp32 values -> 32 bf16 = chunks [4*half, 4*half+4) of this thread's 128-byte panel row
    auto store_bf16_half = [&](uint8_t* panel, int half, const float (&v)[32]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q = make_uint4(pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                             pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]));
        *reinterpret_cast<uint4*>(panel + sw128_offset(row, half * 4 + j)) = q;
      }
    };

    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / num_n_tiles) * 128;
      const int n0 = (t % num_n_tiles) * BN;
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * 256 + lane_sel;

      if constexpr (MODE == EPI_BF16 || MODE == EPI_GELU || MODE == EPI_DGELU) {
        // panel sequence of a tile: one bf16 panel per 64-column chunk, or (EPI_GELU storing z too) a z panel
        // and an h panel per chunk.  Which of the pair comes first flips with chunk and tile parity so that
        // all teams get the same mix of cheap (z) and GELU (h) panels.
        const bool two = (MODE == EPI_GELU) && p.has_out2;
        const int np = two ? 2 * (BN / 64) : BN / 64;
#pragma unroll 1
        for (int pos = 0; pos < np; ++pos) {
          if (!mine()) { ++cnt; continue; }
          const int c = two ? (pos >> 1) : pos;
          const bool is_z = two && ((((pos & 1) + c + tile_iter) & 1) == 0);
          int slot;
          uint8_t* po = acquire(slot);
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            float v[32];
            load32(tacc + c * 64 + half * 32, n0 + c * 64 + half * 32, v);
            if constexpr (MODE == EPI_GELU) {
              if (!is_z) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
              }
            }
            if constexpr (MODE == EPI_DGELU) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 zq = *reinterpret_cast<const uint4*>(po + sw128_offset(row, half * 4 + j));
                const uint32_t zw[4] = {zq.x, zq.y, zq.z, zq.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 z = unpack_bf16x2(zw[e]);
                  v[j * 8 + 2 * e] *= gelu_erf_grad(z.x);
                  v[j * 8 + 2 * e + 1] *= gelu_erf_grad(z.y);
                }
              }
            }
            store_bf16_half(po, half, v);
          }
          publish(is_z ? &tmOut2 : &tmOut, slot, n0 + c * 64, m0);
        }
      } else if constexpr (MODE == EPI_F32) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          if (!mine()) { ++cnt; continue; }
          float v[32];
          load32(tacc + c * 32, n0 + c * 32, v);
          int slot;
          uint8_t* po = acquire(slot);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(po + sw128_offset(row, j)) =
                make_float4(v[j * 4 + 0], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
          publish(&tmOut, slot, n0 + c * 32, m0);
        }
      } else {   // EPI_RES_LN
        float sum = 0.0f;
        const bool tiled = p.out_tiled != nullptr;
        const bool tma_res = !tiled && p.has_res && p.res_table == nullptr;
        const float* trow =
            (p.res_table != nullptr) ? p.res_table + static_cast<size_t>((m0 + row) % p.table_rows) * 192 : nullptr;
        const uint32_t cnt0 = cnt;          // panel c of this tile belongs to team panel_team(c)
        auto panel_team = [&](int c) -> uint32_t {
          return tiled ? static_cast<uint32_t>(c) % kNumTeams : (cnt0 + c) % kNumTeams;   // fixed: batch-invariant sums
        };
        const int grow = m0 + row;
        const bool rvalid = grow < p.M;
        if (tiled && p.res_tiled != nullptr && (lane & 7) == 0) {
          // next tile's residual rows -> L2 (one 128-byte line per 8 lanes, this team's two panels)
          const int nrow = grow + static_cast<int>(gridDim.x) * 128;
          if (t + static_cast<int>(gridDim.x) < num_tiles && nrow < p.M) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const float* xp = p.res_tiled + xt_offset(nrow, team + 3 * k, 0);
#pragma unroll
              for (int j = 0; j < 8; ++j) prefetch_l2(xp + j * 128);
            }
          }
        }
#pragma unroll 1
        for (int c = 0; c < 6; ++c) {
          if (panel_team(c) != static_cast<uint32_t>(team)) { if (!tiled) ++cnt; continue; }
          float4 res4[8];
          if (tiled) {   // residual loads (L2-prefetched one tile ahead) are in flight while the accumulator is read
            const size_t xo = xt_offset(rvalid ? grow : 0, c, 0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              res4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.res_tiled != nullptr) { if (rvalid) res4[j] = *reinterpret_cast<const float4*>(p.res_tiled + xo + j * 128); }
              else if (trow != nullptr) res4[j] = *reinterpret_cast<const float4*>(trow + c * 32 + j * 4);
            }
          }
          float v[32];
          load32(tacc + c * 32, c * 32, v);
          if (tiled) {
            const size_t xo = xt_offset(rvalid ? grow : 0, c, 0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 r = res4[j];
              float4 q = make_float4(v[j * 4 + 0] + r.x, v[j * 4 + 1] + r.y, v[j * 4 + 2] + r.z, v[j * 4 + 3] + r.w);
              v[j * 4 + 0] = q.x; v[j * 4 + 1] = q.y; v[j * 4 + 2] = q.z; v[j * 4 + 3] = q.w;
              sum += (q.x + q.y) + (q.z + q.w);
              if (rvalid) *reinterpret_cast<float4*>(p.out_tiled + xo + j * 128) = q;
            }
            if (p.has_out2) tmem_st32(tacc + c * 32, v);
            continue;
          }
          int slot;
          uint8_t* po = acquire(slot);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tma_res) r = *reinterpret_cast<const float4*>(po + sw128_offset(row, j));
            else if (trow != nullptr) r = *reinterpret_cast<const float4*>(trow + c * 32 + j * 4);
            float4 q = make_float4(v[j * 4 + 0] + r.x, v[j * 4 + 1] + r.y, v[j * 4 + 2] + r.z, v[j * 4 + 3] + r.w);
            v[j * 4 + 0] = q.x; v[j * 4 + 1] = q.y; v[j * 4 + 2] = q.z; v[j * 4 + 3] = q.w;
            sum += (q.x + q.y) + (q.z + q.w);
            *reinterpret_cast<float4*>(po + sw128_offset(row, j)) = q;
          }
          if (p.has_out2) tmem_st32(tacc + c * 32, v);   // keep x' on chip for the LayerNorm passes
          publish(&tmOut, slot, c * 32, m0);
        }
        if (p.has_out2) {
          // row statistics: every team holds a third of the row's columns
          sPart[team * 128 + row] = sum;
          quad_sync();
          float tot = 0.0f;
#pragma unroll
          for (int tm = 0; tm < kNumTeams; ++tm) tot += sPart[tm * 128 + row];
          const float mean = tot * (1.0f / 192.0f);
          float var = 0.0f;
#pragma unroll 1
          for (int c = 0; c < 6; ++c) {
            if (panel_team(c) != static_cast<uint32_t>(team)) continue;
            float v[32];
            tmem_ld32(tacc + c * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float d = v[i] - mean; var = fmaf(d, d, var); }
          }
          sPart[(kNumTeams + team) * 128 + row] = var;
          quad_sync();
          float vtot = 0.0f;
#pragma unroll
          for (int tm = 0; tm < kNumTeams; ++tm) vtot += sPart[(kNumTeams + tm) * 128 + row];
          const float rstd = rsqrtf(vtot * (1.0f / 192.0f) + p.ln_eps);
          if (team == 0 && m0 + row < p.M) {
            if (p.mean_out != nullptr) p.mean_out[m0 + row] = mean;
            if (p.rstd_out != nullptr) p.rstd_out[m0 + row] = rstd;
          }
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            if (!mine()) { ++cnt; continue; }
            int slot;
            uint8_t* po = acquire(slot);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              float v[32];
              tmem_ld32(tacc + c * 64 + half * 32, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int col = c * 64 + half * 32 + i;
                v[i] = (v[i] - mean) * rstd * sGamma[col] + sBeta[col];
              }
              store_bf16_half(po, half, v);
            }
            publish(&tmOut2, slot, c * 64, m0);
          }
        }
      }

      // accumulator stage drained: hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      ++tile_iter;
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

#endif  // __CUDACC__
