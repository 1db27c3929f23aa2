// Attention BACKWARD for DeiT-Tiny (197 tokens, 3 heads of 64); the forward lives in attention_tc.cu
// (tcgen05).  One CTA per (image, head).  Q, K, V, dO of that head (197 x 64 bf16, padded to 208 rows) live in
// shared memory with a 16-byte-chunk XOR swizzle and feed mma.sync m16n8k16 tiles.
//
// Backward (recompute from the saved log-sum-exp, no atomics, no stored probabilities):
//   phase K  (warp = 16 keys)   S^T = K Q^T, P^T = exp(S^T*scale - lse), dV = P^T dO,
//                               dP^T = V dO^T, dS^T = P^T o (dP^T - delta), dK = scale * dS^T Q
//   phase Q  (warp = 16 queries) S, P, dP = dO V^T, dS = P o (dP - delta), dQ = scale * dS K
//   with delta[q] = sum_d dO[q,d] * O[q,d].
#include "common.cuh"

#include <cstdlib>

namespace {

constexpr int kTok = 197;
constexpr int kPad = 208;      // 13 tiles of 16
constexpr int kHd = 64;
constexpr int kHeads = 3;
constexpr int kQkvLd = 576;
constexpr int kCtxLd = 192;
constexpr int kAttnWarps = 13;      // one 16-row tile per warp and phase (7 warps x 2 tiles left the SM at 7 resident warps)
constexpr int kAttnThreads = kAttnWarps * 32;
constexpr float kScale = 0.125f;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {   // bytes, [rows][64] bf16
  return static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A-operand fragments (16 x 64) of row tile `rt` from a swizzled [rows][64] tile
__device__ __forceinline__ void load_a_frags(uint32_t base, int rt, int lane, uint32_t (&f)[4][4]) {
  const int r = rt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(base + tile_off(r, ks * 2 + (lane >> 4)), f[ks]);
}

// acc[nt] (16 x 8 per n-tile, NT n-tiles) += A(16 x 64) * T^T where T is a swizzled [rows][64] tile
// (row = output column index): "A times rows of T".
template <int NT>
__device__ __forceinline__ void mma_a_times_rows(float (&acc)[NT][4], const uint32_t (&a)[4][4], uint32_t tbase,
                                                 int lane) {
  static_assert(NT % 2 == 0, "n-tiles come in pairs");
#pragma unroll
  for (int np = 0; np < NT / 2; ++np) {
    const int r = np * 16 + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t b[4];
      ldsm_x4(tbase + tile_off(r, ks * 2 + ((lane >> 3) & 1)), b);
      mma_bf16(acc[2 * np], a[ks], b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}

// out[dn] (16 x 8 per n-tile, 8 n-tiles = 64 columns) += P(16 x 208, register fragments) * T where T is a
// swizzled [208][64] tile (row = reduction index): "P times columns of T".
__device__ __forceinline__ void mma_p_times_cols(float (&out)[8][4], const uint32_t (&pf)[13][4], uint32_t tbase,
                                                 int lane) {
#pragma unroll
  for (int j = 0; j < 13; ++j) {
    const int r = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldsm_x4_t(tbase + tile_off(r, dp * 2 + (lane >> 4)), b);
      mma_bf16(out[2 * dp], pf[j], b[0], b[1]);
      mma_bf16(out[2 * dp + 1], pf[j], b[2], b[3]);
    }
  }
}

// copy a head's [197 x 64] slice (row stride ld elements) into a swizzled [208][64] tile with cp.async: every
// 16-byte request is in flight at once (the pad rows 197..207 are zeroed once per buffer by zero_pad_rows)
__device__ __forceinline__ void load_head_tile_async(uint8_t* dst, const __nv_bfloat16* src, int ld, int tid,
                                                     int nthreads) {
  const uint32_t base = smem_u32(dst);
  for (int i = tid; i < kTok * 8; i += nthreads) {
    const int r = i >> 3, c = i & 7;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + tile_off(r, c)),
                 "l"(src + static_cast<size_t>(r) * ld + c * 8)
                 : "memory");
  }
}
__device__ __forceinline__ void zero_pad_rows(uint8_t* dst, int tid, int nthreads) {
  for (int i = kTok * 8 + tid; i < kPad * 8; i += nthreads)
    *reinterpret_cast<uint4*>(dst + tile_off(i >> 3, i & 7)) = make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one 16-column chunk `j` of "A times rows of T": acc2 (16 x 16) = A(16 x 64) * T[j*16 .. j*16+16, :]^T
__device__ __forceinline__ void mma_a_times_rows_chunk(float (&acc2)[2][4], const uint32_t (&a)[4][4],
                                                       uint32_t tbase, int j, int lane) {
  const int r = j * 16 + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t b[4];
    ldsm_x4(tbase + tile_off(r, ks * 2 + ((lane >> 3) & 1)), b);
    mma_bf16(acc2[0], a[ks], b[0], b[1]);
    mma_bf16(acc2[1], a[ks], b[2], b[3]);
  }
}
// one 16-deep reduction step `j` of "P times columns of T": out (16 x 64) += Pj(16 x 16) * T[j*16 .. +16, :]
__device__ __forceinline__ void mma_p_chunk_times_cols(float (&out)[8][4], const uint32_t (&pa)[4], uint32_t tbase,
                                                       int j, int lane) {
  const int r = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int dp = 0; dp < 4; ++dp) {
    uint32_t b[4];
    ldsm_x4_t(tbase + tile_off(r, dp * 2 + (lane >> 4)), b);
    mma_bf16(out[2 * dp], pa, b[0], b[1]);
    mma_bf16(out[2 * dp + 1], pa, b[2], b[3]);
  }
}

// write a warp's 16 x 64 fp32 fragment tile as bf16 rows [rt*16, rt*16+16) of a global matrix, staging
// through a warp-private [16][64] swizzled smem tile so the global stores are 128 B per row
__device__ __forceinline__ void store_tile_bf16(uint8_t* stage, const float (&o)[8][4], int rt, int lane,
                                                __nv_bfloat16* dst, int ld, float mul0, float mul1) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) {
    *reinterpret_cast<uint32_t*>(stage + tile_off(g, dn) + t * 4) = pack_bf16x2(o[dn][0] * mul0, o[dn][1] * mul0);
    *reinterpret_cast<uint32_t*>(stage + tile_off(g + 8, dn) + t * 4) = pack_bf16x2(o[dn][2] * mul1, o[dn][3] * mul1);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = i * 32 + lane;
    const int lr = idx >> 3, c = idx & 7;
    const int r = rt * 16 + lr;
    if (r < kTok)
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(r) * ld + c * 8) =
          *reinterpret_cast<const uint4*>(stage + tile_off(lr, c));
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------- backward
// smem tiles: Q, K, V, dO (bf16 [208][64] swizzled), per-warp staging, lse[208], delta[208] (fp32)
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ ctx,
                const __nv_bfloat16* __restrict__ dctx, const float* __restrict__ lse,
                __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kPad * 128;
  uint8_t* sV = sK + kPad * 128;
  uint8_t* sDO = sV + kPad * 128;
  uint8_t* sStage = sDO + kPad * 128;
  float* sLse = reinterpret_cast<float*>(sStage + kAttnWarps * 2048);
  float* sDelta = sLse + kPad;
  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t tok0 = static_cast<size_t>(b) * kTok;
  const __nv_bfloat16* base = qkv + tok0 * kQkvLd + h * kHd;
  const __nv_bfloat16* dO = dctx + tok0 * kCtxLd + h * kHd;
  const __nv_bfloat16* O = ctx + tok0 * kCtxLd + h * kHd;
  load_head_tile_async(sQ, base, kQkvLd, threadIdx.x, kAttnThreads);
  load_head_tile_async(sK, base + 192, kQkvLd, threadIdx.x, kAttnThreads);
  load_head_tile_async(sV, base + 384, kQkvLd, threadIdx.x, kAttnThreads);
  load_head_tile_async(sDO, dO, kCtxLd, threadIdx.x, kAttnThreads);
  cp_async_commit();
  zero_pad_rows(sQ, threadIdx.x, kAttnThreads);
  zero_pad_rows(sK, threadIdx.x, kAttnThreads);
  zero_pad_rows(sV, threadIdx.x, kAttnThreads);
  zero_pad_rows(sDO, threadIdx.x, kAttnThreads);
  // delta[q] = <dO[q,:], O[q,:]>; one 8-lane group per row
  for (int r = threadIdx.x >> 3; r < kPad; r += kAttnThreads >> 3) {
    float acc = 0.0f;
    if (r < kTok) {
      const int c = threadIdx.x & 7;
      const uint4 a = *reinterpret_cast<const uint4*>(dO + static_cast<size_t>(r) * kCtxLd + c * 8);
      const uint4 o = *reinterpret_cast<const uint4*>(O + static_cast<size_t>(r) * kCtxLd + c * 8);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(ow[e]);
        acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if ((threadIdx.x & 7) == 0) {
      sDelta[r] = acc;
      sLse[r] = (r < kTok) ? lse[static_cast<size_t>(blockIdx.x) * kTok + r] : 0.0f;
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  uint8_t* stage = sStage + warp * 2048;
  __nv_bfloat16* dq_out = dqkv + tok0 * kQkvLd + h * kHd;
  constexpr float kS2 = kScale * kLog2e;
  const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aDO = smem_u32(sDO);

  // ---- phase K: this warp owns 16 keys; it streams over 16-query chunks of the transposed score tile
  for (int kt = warp; kt < 13; kt += kAttnWarps) {
    uint32_t kf[4][4], vf[4][4];
    load_a_frags(aK, kt, lane, kf);
    load_a_frags(aV, kt, lane, vf);
    const int key0 = kt * 16 + g, key1 = key0 + 8;
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.0f;
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.0f;
    }
#pragma unroll 1
    for (int j = 0; j < 13; ++j) {
      float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_a_times_rows_chunk(st, kf, aQ, j, lane);      // S^T  = K Q^T
      mma_a_times_rows_chunk(dpt, vf, aDO, j, lane);    // dP^T = V dO^T
      uint32_t pa[4], dsa[4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        float pv[4], dsv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = j * 16 + n * 8 + 2 * t + (e & 1);
          const int key = (e < 2) ? key0 : key1;
          const bool ok = (q < kTok) && (key < kTok);
          pv[e] = ok ? exp2f(st[n][e] * kS2 - sLse[q]) : 0.0f;
          dsv[e] = pv[e] * (dpt[n][e] - sDelta[q]);
        }
        pa[n * 2 + 0] = pack_bf16x2(pv[0], pv[1]);
        pa[n * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
        dsa[n * 2 + 0] = pack_bf16x2(dsv[0], dsv[1]);
        dsa[n * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
      }
      mma_p_chunk_times_cols(dv, pa, aDO, j, lane);     // dV += P^T  dO
      mma_p_chunk_times_cols(dk, dsa, aQ, j, lane);     // dK += dS^T Q
    }
    store_tile_bf16(stage, dv, kt, lane, dq_out + 384, kQkvLd, 1.0f, 1.0f);
    store_tile_bf16(stage, dk, kt, lane, dq_out + 192, kQkvLd, kScale, kScale);
  }

  // ---- phase Q: this warp owns 16 queries; it streams over 16-key chunks
  for (int qt = warp; qt < 13; qt += kAttnWarps) {
    uint32_t qf[4][4], dof[4][4];
    load_a_frags(aQ, qt, lane, qf);
    load_a_frags(aDO, qt, lane, dof);
    const int q0 = qt * 16 + g, q1 = q0 + 8;
    const float lse0 = sLse[q0], lse1 = sLse[q1], dl0 = sDelta[q0], dl1 = sDelta[q1];
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.0f;
#pragma unroll 1
    for (int j = 0; j < 13; ++j) {
      float s2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_a_times_rows_chunk(s2, qf, aK, j, lane);      // S  = Q K^T
      mma_a_times_rows_chunk(dp, dof, aV, j, lane);     // dP = dO V^T
      uint32_t dsa[4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        float dsv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = j * 16 + n * 8 + 2 * t + (e & 1);
          const int q = (e < 2) ? q0 : q1;
          const bool ok = (q < kTok) && (key < kTok);
          const float pv = ok ? exp2f(s2[n][e] * kS2 - ((e < 2) ? lse0 : lse1)) : 0.0f;
          dsv[e] = pv * (dp[n][e] - ((e < 2) ? dl0 : dl1));
        }
        dsa[n * 2 + 0] = pack_bf16x2(dsv[0], dsv[1]);
        dsa[n * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
      }
      mma_p_chunk_times_cols(dq, dsa, aK, j, lane);     // dQ += dS K
    }
    store_tile_bf16(stage, dq, qt, lane, dq_out, kQkvLd, kScale, kScale);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------- launchers
int rvk_attention_bwd_tc_launch(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv, int batch,
                                cudaStream_t stream);      // attention_tc.cu
constexpr int kAttnBwdSmem = 4 * kPad * 128 + kAttnWarps * 2048 + 2 * kPad * 4;

int rvk_attention_bwd_launch(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                             int batch, cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  // default: the tcgen05 kernel (attention_tc.cu); RVK_ATTN_BWD_SIMT=1 keeps this file's mma.sync kernel (A/B measurements)
  static const bool simt = [] { const char* e = getenv("RVK_ATTN_BWD_SIMT"); return e != nullptr && e[0] == '1'; }();
  if (!simt) return rvk_attention_bwd_tc_launch(qkv, ctx, dctx, lse, dqkv, batch, stream);
  RVK_SET_MAX_SMEM(attn_bwd_kernel, kAttnBwdSmem);
  attn_bwd_kernel<<<batch * kHeads, kAttnThreads, kAttnBwdSmem, stream>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(ctx),
      static_cast<const __nv_bfloat16*>(dctx), lse, static_cast<__nv_bfloat16*>(dqkv));
  return rvk_launch_check();
}
