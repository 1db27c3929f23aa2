// Fused transformer MLP block on tcgen05 / TMEM, sm_100a (inference path of the DeiT-Tiny trunk):
//
//     x_out = x_in + fc2( gelu( fc1( ln2 ) + b1 ) ) + b2 ;     ln_out = LayerNorm(x_out) * gamma + beta
//
// (timm Block.forward second half: x + mlp(norm2(x)), followed by the NEXT block's norm1; restated in
// oracle/vit.py::_Block.)  The 768-wide hidden activation never leaves the SM: per 128-row tile the hidden dimension is
// walked in six chunks of 128 columns,
//     D1[c&1] = ln2 . W1[c]^T        UMMA  M=128*G  N=128  K=192   (TMEM columns   0..255, two buffers)
//     H[c&1]  = bf16(gelu(D1 + b1))  epilogue warps: tcgen05.ld -> registers -> K-major swizzled smem operand
//     D2     += H[c&1] . W2[:,c]^T   UMMA  M=128*G  N=192  K=128   (TMEM columns 256..447)
// and the final epilogue adds bias + residual, writes the fp32 token stream and the LayerNorm'ed bf16 operand of
// the next GEMM.  HBM traffic per row: 384 B (ln2) + 768 B (x in) + 768 B (x out) + 384 B (ln out) instead of the
// additional 2 x 1536 B round trip of the hidden activation in the unfused fc1 / fc2 pair.
//
// G = 2 (default): the two CTAs of a cluster drive ONE tcgen05.mma.cta_group::2 with M = 256 -- each CTA stages
// its own 128 rows of A / H and only HALF of every weight panel, so the weight stream out of L2 and the shared
// memory operand reads per SM are halved against G = 1 (kept as a single-CTA variant for testing).
//
// Warp roles (480 threads): w0 TMA producer (A tile + a ring of weight panels), w1 UMMA issuer (leader CTA only) +
// TMEM owner, w2 idle, w3..w14 twelve epilogue warps = three teams x four TMEM lane quadrants (one accumulator row
// per thread).  GELU column groups rotate over the teams; in the final epilogue team t owns columns [64t, 64t+64).
// The fp32 token stream uses the tiled layout of common.cuh (xt_offset): coalesced 512-byte global accesses
// straight from/to registers, no staging.  The final epilogue of tile i runs after the first GELU chunk of tile
// i+1, so the epilogue warps never wait for the last fc2 of a tile.
#pragma once

#include "common.cuh"

struct MlpFusedParams {
  int M;
  const float* x_in;    // tiled fp32 [M_pad, 192] residual
  float* x_out;         // tiled fp32 (may alias x_in)
  const float* b1;      // [768]
  const float* b2;      // [192]
  const float* gamma;   // [192] next LayerNorm (has_ln)
  const float* beta;
  float eps;
  int has_ln;
};

constexpr int kMlpThreads = 480;

template <int G>
struct MlpSmem {
  static constexpr int kABytes = 3 * 16384;                 // ln2 tile: three [128 x 64] K panels
  static constexpr int kHBytes = 2 * 2 * 16384;             // two hidden-chunk buffers of two K panels
  static constexpr int kWStage = 24576 / G;                 // one W2 panel [192/G x 64]; W1 panels [128/G x 64] use 2/3
  static constexpr int kWStages = (G == 2) ? 4 : 2;
  static constexpr int kW1Bytes = 16384 / G;
  static constexpr int kW2Bytes = 24576 / G;
  static constexpr int kLnBytes = 12 * 4096;                // per epilogue warp: [32 rows x 128 B] bf16 staging
  static constexpr int kVecBytes = (768 + 3 * 192) * 4;
  static constexpr int kPartBytes = 2 * 3 * 128 * 8;        // LayerNorm partial (sum, sumsq) [parity][team][row]
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = 1024 + kABytes + kHBytes + kWStages * kWStage + kLnBytes + kVecBytes + kPartBytes + kBarBytes;
};

#ifdef __CUDACC__

template <int G>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmLn,
                 const MlpFusedParams p) {
  using L = MlpSmem<G>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sH = sA + L::kABytes;
  uint8_t* sW = sH + L::kHBytes;
  uint8_t* sLn = sW + L::kWStages * L::kWStage;
  float* sB1 = reinterpret_cast<float*>(sLn + L::kLnBytes);
  float* sB2 = sB1 + 768;
  float* sGamma = sB2 + 192;
  float* sBeta = sGamma + 192;
  float2* sPart = reinterpret_cast<float2*>(sBeta + 192);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sPart) + L::kPartBytes);
  uint64_t* a_full = bars;            // leader: A tiles of both CTAs landed
  uint64_t* a_empty = bars + 1;       // fc1 of the tile's last chunk done
  uint64_t* w_full = bars + 2;        // [4]
  uint64_t* w_empty = bars + 6;       // [4]
  uint64_t* d1_full = bars + 10;      // [2]
  uint64_t* gelu_done = bars + 12;    // [2] leader: D1[b] drained and H[b] written by every epilogue warp of the pair
  uint64_t* h_empty = bars + 14;      // [2]
  uint64_t* d2_full = bars + 16;
  uint64_t* d2_empty = bars + 17;     // leader
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (G == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x / G;
  const int num_clusters = gridDim.x / G;
  const int num_tiles = (p.M + 127) / 128;
  const int num_units = (num_tiles + G - 1) / G;            // a unit = G consecutive 128-row tiles
  const int n_my = (cluster_id < num_units) ? (num_units - cluster_id + num_clusters - 1) / num_clusters : 0;
  const int Q = 6 * n_my;                                   // hidden chunks this cluster walks
  auto tile_row0 = [&](int it) { return ((cluster_id + it * num_clusters) * G + static_cast<int>(rank)) * 128; };

  // ---- one-time setup
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    sB1[i] = p.b1[i];
    if (i < 192) {
      sB2[i] = p.b2[i];
      sGamma[i] = p.has_ln ? p.gamma[i] : 1.0f;
      sBeta[i] = p.has_ln ? p.beta[i] : 0.0f;
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmLn);
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < L::kWStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d1_full[i], 1);
      mbar_init(&gelu_done[i], 12 * G);
      mbar_init(&h_empty[i], 1);
    }
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, 12 * G);
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

  auto wait_leader = [&](uint64_t* bar, uint32_t parity) {
    if (G == 2) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity);
  };

  if (warp == 0) {
    // ================================================================= TMA producer (every CTA loads its own share)
    if (lane == 0 && n_my > 0) {
      const uint32_t a_full_l = (G == 2) ? mapa_u32(smem_u32(a_full), 0) : 0u;
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
      auto load_a = [&](int it) {
        mbar_wait(a_empty, (it & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(a_full, G * L::kABytes);
        const int m0 = tile_row0(it);
        for (int kp = 0; kp < 3; ++kp) {
          if (G == 2) tma_load_2d_pair(sA + kp * 16384, &tmA, a_full_l, kp * 64, m0);
          else tma_load_2d(sA + kp * 16384, &tmA, a_full, kp * 64, m0);
        }
      };
      load_a(0);
      load_w1(0);
      load_w1(1);
      for (int q = 0; q < Q; ++q) {
        const int c = q % 6, it = q / 6;
        if (q + 2 < Q) {
          const int c2 = (q + 2) % 6;
          if (c2 != 0) { load_w1(c2); load_w2(c); }
          else { load_w2(c); load_a(it + 1); load_w1(0); }
        } else {
          load_w2(c);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= UMMA issuer (leader CTA only)
    if (lane == 0 && leader && n_my > 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(128 * G, 128, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(128 * G, 192, 0, 0);
      int ws = 0;
      uint32_t wph = 0;
      auto mma = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool acc) {
        const uint64_t ad = umma_smem_desc(a_addr, 16, 1024);
        const uint64_t bd = umma_smem_desc(b_addr, 16, 1024);
        if (G == 2) umma_bf16_pair(d, ad, bd, idesc, acc ? 1u : 0u);
        else umma_bf16(d, ad, bd, idesc, acc ? 1u : 0u);
      };
      auto commit = [&](uint64_t* bar) {
        if (G == 2) umma_commit_pair(bar); else umma_commit(bar);
      };
      auto fc1 = [&](int q) {
        const int b = q & 1;
        const uint32_t d = tmem_base + b * 128;
        for (int kp = 0; kp < 3; ++kp) {
          wait_leader(&w_full[ws], wph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + kp * 16384);
          const uint32_t b_addr = smem_u32(sW + ws * L::kWStage);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma(d, a_addr + k * 32, b_addr + k * 32, idesc1, (kp | k) != 0);
          commit(&w_empty[ws]);
          if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
        }
        commit(&d1_full[b]);
        if (q % 6 == 5) commit(a_empty);
      };
      auto fc2 = [&](int q) {
        const int b = q & 1, c = q % 6;
        const uint32_t d = tmem_base + 256;
        for (int kp = 0; kp < 2; ++kp) {
          wait_leader(&w_full[ws], wph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sH + b * 32768 + kp * 16384);
          const uint32_t b_addr = smem_u32(sW + ws * L::kWStage);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma(d, a_addr + k * 32, b_addr + k * 32, idesc2, (c | kp | k) != 0);
          commit(&w_empty[ws]);
          if (++ws == L::kWStages) { ws = 0; wph ^= 1; }
        }
        commit(&h_empty[b]);
        if (c == 5) commit(d2_full);
      };
      wait_leader(a_full, 0);
      tc_fence_after();
      fc1(0);
      fc1(1);
      for (int q = 0; q < Q; ++q) {
        const int c = q % 6, it = q / 6;
        wait_leader(&gelu_done[q & 1], (q >> 1) & 1);
        tc_fence_after();
        const bool more = q + 2 < Q;
        const int c2 = (q + 2) % 6;
        if (more && c2 != 0) fc1(q + 2);
        if (c == 0 && it > 0) {
          wait_leader(d2_empty, (it - 1) & 1);
          tc_fence_after();
        }
        fc2(q);
        if (more && c2 == 0) {
          wait_leader(a_full, (it + 1) & 1);
          tc_fence_after();
          fc1(q + 2);
        }
      }
    }
  } else if (warp >= 3) {
    // ================================================================= epilogue warps
    const int quad = warp & 3;
    const int ew = warp - 3;
    const int team = ew >> 2;
    const int row = quad * 32 + lane;                   // accumulator row == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    uint8_t* myLn = sLn + ew * 4096;
    const uint32_t gd_l[2] = {(G == 2) ? mapa_u32(smem_u32(&gelu_done[0]), 0) : 0u,
                              (G == 2) ? mapa_u32(smem_u32(&gelu_done[1]), 0) : 0u};
    const uint32_t d2e_l = (G == 2) ? mapa_u32(smem_u32(d2_empty), 0) : 0u;

    auto final_tile = [&](int itf) {
      const int m0 = tile_row0(itf);
      const int grow = m0 + row;
      const bool valid = grow < p.M;
      const float* xin = p.x_in + xt_offset(valid ? grow : 0, 2 * team, 0);
      float* xout = p.x_out + xt_offset(valid ? grow : 0, 2 * team, 0);
      float x0[32], x1[32];
      // residual of the first 32 columns: in flight while the last fc2 of the tile completes
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) r = *reinterpret_cast<const float4*>(xin + j * 128);
        x0[j * 4 + 0] = r.x; x0[j * 4 + 1] = r.y; x0[j * 4 + 2] = r.z; x0[j * 4 + 3] = r.w;
      }
      mbar_wait(d2_full, itf & 1);
      tc_fence_after();
      const uint32_t tD2 = tmem_base + 256 + team * 64 + lane_sel;
      {
        float v[32];
        tmem_ld32(tD2, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) x0[i] += v[i] + sB2[team * 64 + i];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) r = *reinterpret_cast<const float4*>(xin + 1024 + j * 128);
        x1[j * 4 + 0] = r.x; x1[j * 4 + 1] = r.y; x1[j * 4 + 2] = r.z; x1[j * 4 + 3] = r.w;
      }
      {
        float v[32];
        tmem_ld32(tD2 + 32, v);
        // D2 is in registers: the next tile's fc2 may overwrite it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (G == 2) mbar_arrive_cluster(d2e_l); else mbar_arrive(d2_empty);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) x1[i] += v[i] + sB2[team * 64 + 32 + i];
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<float4*>(xout + j * 128) = make_float4(x0[j * 4], x0[j * 4 + 1], x0[j * 4 + 2], x0[j * 4 + 3]);
          *reinterpret_cast<float4*>(xout + 1024 + j * 128) = make_float4(x1[j * 4], x1[j * 4 + 1], x1[j * 4 + 2], x1[j * 4 + 3]);
        }
      }
      if (p.has_ln) {
        float s = 0.0f, ss = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) { s += x0[i] + x1[i]; ss = fmaf(x0[i], x0[i], ss); ss = fmaf(x1[i], x1[i], ss); }
        float2* part = sPart + (itf & 1) * 384;
        part[team * 128 + row] = make_float2(s, ss);
        named_bar_sync(2 + quad, 96);               // the three warps that share this TMEM quadrant
        float ts = 0.0f, tss = 0.0f;
#pragma unroll
        for (int t = 0; t < 3; ++t) { const float2 v = part[t * 128 + row]; ts += v.x; tss += v.y; }
        const float mean = ts * (1.0f / 192.0f);
        const float var = fmaxf(tss * (1.0f / 192.0f) - mean * mean, 0.0f);
        const float rstd = rsqrtf(var + p.eps);
        if (lane == 0) tma_store_wait_read<0>();    // the previous tile's store out of myLn has been read
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float* xs = (j < 4) ? &x0[j * 8] : &x1[(j - 4) * 8];
          const int col = team * 64 + j * 8;
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaf((xs[e] - mean) * rstd, sGamma[col + e], sBeta[col + e]);
          *reinterpret_cast<uint4*>(myLn + sw128_offset(lane, j)) =
              make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (m0 + quad * 32 < p.M) tma_store_2d(&tmLn, myLn, team * 64, m0 + quad * 32);
          tma_store_commit();
        }
      }
    };

    for (int it = 0; it < n_my; ++it) {
#pragma unroll 1
      for (int c = 0; c < 6; ++c) {
        const int q = it * 6 + c, b = q & 1;
        const uint32_t n = static_cast<uint32_t>(q >> 1);
        mbar_wait(&d1_full[b], n & 1);
        mbar_wait(&h_empty[b], (n & 1) ^ 1);
        tc_fence_after();
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          if ((q + g) % 3 != team) continue;
          float v[32];
          tmem_ld32(tmem_base + b * 128 + g * 32 + lane_sel, v);
          const float* bb = sB1 + c * 128 + g * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = *reinterpret_cast<const float4*>(bb + i * 4);
            v[i * 4 + 0] = gelu_erf(v[i * 4 + 0] + b4.x);
            v[i * 4 + 1] = gelu_erf(v[i * 4 + 1] + b4.y);
            v[i * 4 + 2] = gelu_erf(v[i * 4 + 2] + b4.z);
            v[i * 4 + 3] = gelu_erf(v[i * 4 + 3] + b4.w);
          }
          uint8_t* panel = sH + b * 32768 + (g >> 1) * 16384;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 qv = make_uint4(pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                                  pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]));
            *reinterpret_cast<uint4*>(panel + sw128_offset(row, (g & 1) * 4 + j)) = qv;
          }
        }
        tc_fence_before();
        if (G == 2) fence_proxy_async_all(); else fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (G == 2) mbar_arrive_cluster(gd_l[b]); else mbar_arrive(&gelu_done[b]);
        }
        if (c == 0 && it > 0) final_tile(it - 1);
      }
    }
    if (n_my > 0) final_tile(n_my - 1);
    if (lane == 0) tma_store_wait_all<0>();
  }

  // ---- teardown
  tc_fence_before();
  if (G == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    if (G == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

#endif  // __CUDACC__
