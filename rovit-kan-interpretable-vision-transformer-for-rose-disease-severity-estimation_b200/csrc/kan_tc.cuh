// Tensor-core formulation of the fused KAN layer forward for large batches (included by kan.cu inside its anonymous
// namespace: uses Knots, kan_expand, kKW).
//
//   y[b, o] = act( bias[o] + sum_{i,k} A[b, i*8+k] * Wp[i*8+k, o] ),   A[b, i*8+k] = N_k(tanh x[b,i]) (k < 7), x[b,i] (k = 7)
//
// is a GEMM [B x 8*in] x [8*in x out] whose left operand does not exist in memory: sixteen producer warps evaluate
// the truncated cubic B-spline (closed form, precise tanhf: the reference's basis is discontinuous at tanh x = 0.4)
// for 128 samples x 8 inputs at a time and write the 64 packed activations per sample straight into a K-major,
// 128-byte-swizzled UMMA operand tile in shared memory.  fp32 parity (1e-3 relative on the layer output) rules out
// plain bf16 operands (2^-9) and TF32 (2^-11 per operand); both operands are therefore split hi + lo in bf16 and
// the product is accumulated as  A_hi W_hi + A_hi W_lo + A_lo W_hi  (relative error ~2^-16 per term) in fp32 TMEM.
//
// Warp roles (576 threads): w0 TMA producer (weight chunks [64 out x 64 k] hi / lo, 4-stage ring), w1 UMMA issuer
// (whole warp, uniform datapath), w2..w17 activation producers + epilogue (TMEM -> bias -> activation -> y).
// Per 128-sample tile: in/8 chunks x (3 products x 4 UMMA of M=128 N=64 K=16); two TMEM accumulators so that the
// epilogue of tile t overlaps the chunks of tile t+1.  The kernel is bound by the producers (~100 instructions per
// (sample, input): tanhf, interval search, four cubics, hi/lo split), not by the tensor pipe or HBM.
#pragma once

constexpr int kTcThreads = 18 * 32;
constexpr int kTcAStages = 3;
constexpr int kTcWStages = 4;
constexpr int kTcAStageBytes = 2 * 16384;            // A_hi | A_lo, each [128 x 64] bf16
constexpr int kTcWStageBytes = 2 * 8192;             // W_hi | W_lo, each [64 x 64] bf16
constexpr int kTcSmemBytes = 1024 + kTcAStages * kTcAStageBytes + kTcWStages * kTcWStageBytes + 96 * 4 + 256;

// Interval search in x-space: tanh is monotone, so  tanh(x) >= knot_m  <=>  x >= atanh(knot_m).  Deciding the
// knot interval (and with it the reference's discontinuity at tanh x = 0.4) from thresholds in x removes the
// accuracy requirement from tanh itself: the VALUES of the cubics are continuous in t, so a branch-free
// tanh = 1 - 2 / (2^(2 log2(e) |x|) + 1) on MUFU.EX2 / MUFU.RCP (absolute error ~3e-7) is enough for them.
struct KanTcTables {
  const float* xthr; // device [9]: xthr[m] = smallest float x with tanhf(x) >= knot_m (m = 1..7); xthr[0] = -inf, xthr[8] = +inf
};

// The thresholds are calibrated against the SAME tanhf the CUDA-core kernels (and the parity tests) use, so that
// both paths take identical interval decisions even for inputs sitting exactly on a knot: scan +-32 ulps around
// atanh(knot_m) for the smallest float whose tanhf reaches the knot.
__global__ void kan_tc_thresholds_kernel(Knots kn, float* __restrict__ xthr) {
  const int m = threadIdx.x;
  if (m > 8) return;
  if (m == 0) { xthr[0] = -INFINITY; return; }
  if (m == 8) { xthr[8] = INFINITY; return; }
  const float km = kn.k[m];
  float xf = atanhf(km);
  for (int d = 0; d < 32; ++d) xf = nextafterf(xf, -INFINITY);
  for (int d = 0; d < 64 && tanhf(xf) < km; ++d) xf = nextafterf(xf, INFINITY);
  xthr[m] = xf;
}

// WpT (fp32 [out_pad=64][kp]) -> bf16 hi / lo, row-major [64][kp] each
__global__ void kan_split_weights_kernel(const float* __restrict__ spline, const float* __restrict__ lin_w, int n_in,
                                         int n_out, int kp, __nv_bfloat16* __restrict__ w_hi,
                                         __nv_bfloat16* __restrict__ w_lo) {
  const long long total = 64LL * kp;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(idx / kp), kk = static_cast<int>(idx % kp);
    const int i = kk >> 3, k = kk & 7;
    float v = 0.0f;
    if (i < n_in && o < n_out)
      v = (k < 7) ? spline[(static_cast<size_t>(i) * n_out + o) * 7 + k] : lin_w[static_cast<size_t>(o) * n_in + i];
    const __nv_bfloat16 hi = __float2bfloat16(v);
    w_hi[idx] = hi;
    w_lo[idx] = __float2bfloat16(v - __bfloat162float(hi));
  }
}

// The 8 packed activations [N_0(tanh x) .. N_6(tanh x), x] of one (sample, input), split hi + lo in bf16, written as one
// 16-byte K chunk each at byte offset `off` of the hi / lo operand tiles.
//
// Round-2 version (ncu on the round-1 one: 125 instructions per (sample, input), 8.7 M shared-memory bank conflicts per launch
// from eight predicated 2-byte stores per pair, 10 % of all warp samples waiting for the threshold / knot table loads):
//   * interval from s = 5 (tanh x + 1) directly; the calibrated x-space thresholds (exact decisions, incl. the reference's
//     jump at tanh x = 0.4) are only consulted when s is within 1e-4 of an integer (fast-tanh error: 2e-6 in s);
//   * the local coordinate is u = s - j: the knot vector is the reference's linspace(-1, 1, 11) (api.cu::check_knots);
//   * the four live cubics in Horner form, converted to bf16 hi / lo two at a time (cvt.rn.bf16x2.f32);
//   * one-hot placement = a 128-bit shift of the packed 4 x bf16 group by 16 (j - 3) bits, one 16-byte store per tile.
// Knot interval j of x and float(j) for the reference's knot vector linspace(-1, 1, 11) (the only one the C ABI accepts,
// api.cu::check_knots): j = floor(s), s = 5 (tanh x + 1) in [0, 10]; 0..6 live, >= 7 dead zone (a negative j can only come
// out of a NaN input: callers test `unsigned(j) < 7`).  Taken WITHOUT F2I / I2F (XU-pipe conversions in the middle of the
// dependent chain): a round-down add of 1.5 * 2^23 leaves floor(s) in the low mantissa bits and, minus the constant, as an
// exact float.  The calibrated x-space thresholds are consulted only within 1e-4 of a knot (fast-tanh error in s: 2e-6);
// sXthr[9..11] = +inf, so no range check on j is needed.  No clamps here: the callers mask dead lanes.
__device__ __forceinline__ void kan_tc_interval(float xe, float s, const float* sXthr, int& j, float& jfl) {
  const float sm = __fadd_rd(s, 12582912.0f);
  j = __float_as_int(sm) - 0x4B400000;
  jfl = sm - 12582912.0f;
  if (fabsf((s - jfl) - 0.5f) > 0.5f - 1e-4f) {           // rare (2e-4 of the inputs): exact decision
    if (xe < sXthr[j]) { --j; jfl -= 1.0f; }
    else if (xe >= sXthr[j + 1]) { ++j; jfl += 1.0f; }
  }
}

// tanh x = 1 - 2 / (2^(2 log2(e) x) + 1), branch-free on MUFU.EX2 / MUFU.RCP (absolute error ~3e-7) for either sign of x
// (no |x| / copysign: one ALU-pipe instruction less); rc = 1 / (e^(2x) + 1), so 1 - tanh^2 = 4 rc (1 - rc)
__device__ __forceinline__ float kan_tc_tanh(float xe, float& rc) {
  const float ex = ex2_approx(xe * 2.8853900817779268f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(ex + 1.0f));
  return fmaf(-2.0f, rc, 1.0f);
}

__device__ __forceinline__ void kan_tc_expand_store(float xe, uint32_t off, uint8_t* ahi, uint8_t* alo, const float* sXthr) {
  float rc;
  const float tt = kan_tc_tanh(xe, rc);                    // values only; the interval comes from x-space thresholds where it matters
  const float s = fmaf(tt, 5.0f, 5.0f);                    // in [0, 10]
  int j;
  float jfl;
  kan_tc_interval(xe, s, sXthr, j, jfl);
  const bool live = static_cast<unsigned>(j) < 7u;         // dead zone: zeros (the mask rides in the polynomial scale)
  const int jc = static_cast<int>(min(static_cast<unsigned>(j), 6u));
  const float c6 = live ? (1.0f / 6.0f) : 0.0f;
  const float u = s - jfl;
  const float u2 = u * u, om = 1.0f - u;
  const float v0 = u2 * u * c6;                                                        // slot j
  const float v1 = fmaf(fmaf(fmaf(-3.0f, u, 3.0f), u, 3.0f), u, 1.0f) * c6;            // slot j-1
  const float v2 = fmaf(fmaf(3.0f, u, -6.0f), u2, 4.0f) * c6;                          // slot j-2
  const float v3 = om * om * om * c6;                                                  // slot j-3
  // hi / lo split, two values per conversion: word = [second : first] as bf16 pairs
  const __nv_bfloat162 h32 = __floats2bfloat162_rn(v3, v2), h10 = __floats2bfloat162_rn(v1, v0);
  const uint32_t w32 = *reinterpret_cast<const uint32_t*>(&h32), w10 = *reinterpret_cast<const uint32_t*>(&h10);
  const __nv_bfloat162 l32 = __floats2bfloat162_rn(v3 - __uint_as_float(w32 << 16), v2 - __uint_as_float(w32 & 0xffff0000u));
  const __nv_bfloat162 l10 = __floats2bfloat162_rn(v1 - __uint_as_float(w10 << 16), v0 - __uint_as_float(w10 & 0xffff0000u));
  const uint32_t q32 = *reinterpret_cast<const uint32_t*>(&l32), q10 = *reinterpret_cast<const uint32_t*>(&l10);
  // slots j-3 .. j <- (v3, v2, v1, v0): shift the 64-bit group by 16 * (j - 3) bits inside the 128-bit row.
  // (A branch-free form -- three funnel shifts + a two-level select on (7 - jc) / 2 -- measured 4-6 % SLOWER than the shift
  // pairs + short branches below.)
  const int sh = (jc - 3) * 16;                            // -48 .. 48
  const unsigned long long vh = (static_cast<unsigned long long>(w10) << 32) | w32;
  const unsigned long long vl = (static_cast<unsigned long long>(q10) << 32) | q32;
  unsigned long long h_lo, h_hi, l_lo, l_hi;
  if (sh >= 0) {
    h_lo = vh << sh; l_lo = vl << sh;
    h_hi = sh ? (vh >> (64 - sh)) : 0ull;
    l_hi = sh ? (vl >> (64 - sh)) : 0ull;
  } else {
    h_lo = vh >> (-sh); l_lo = vl >> (-sh);
    h_hi = 0ull; l_hi = 0ull;
  }
  // slot 7 = raw x (hi / lo)
  const uint32_t xh = pack_bf16x2(0.0f, xe) & 0xffff0000u;             // packed conversions run on the ALU pipe (F2F: XU)
  const uint32_t xl = pack_bf16x2(0.0f, xe - __uint_as_float(xh)) & 0xffff0000u;
  h_hi |= static_cast<unsigned long long>(xh) << 32;
  l_hi |= static_cast<unsigned long long>(xl) << 32;
  *reinterpret_cast<uint4*>(ahi + off) = make_uint4(static_cast<uint32_t>(h_lo), static_cast<uint32_t>(h_lo >> 32),
                                                     static_cast<uint32_t>(h_hi), static_cast<uint32_t>(h_hi >> 32));
  *reinterpret_cast<uint4*>(alo + off) = make_uint4(static_cast<uint32_t>(l_lo), static_cast<uint32_t>(l_lo >> 32),
                                                     static_cast<uint32_t>(l_hi), static_cast<uint32_t>(l_hi >> 32));
}

__global__ void __launch_bounds__(kTcThreads, 1)
kan_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                  const float* __restrict__ x, const float* __restrict__ bias, const KanTcTables tb, float* __restrict__ y,
                  int act, int batch, int n_in, int n_out, int num_chunks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sW = sA + kTcAStages * kTcAStageBytes;
  float* sBias = reinterpret_cast<float*>(sW + kTcWStages * kTcWStageBytes);
  float* sXthr = sBias + 64;     // [12]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXthr + 32);
  uint64_t* a_full = bars;                       // [3]
  uint64_t* a_empty = a_full + kTcAStages;       // [3]
  uint64_t* w_full = a_empty + kTcAStages;       // [4]
  uint64_t* w_empty = w_full + kTcWStages;       // [4]
  uint64_t* d_full = w_empty + kTcWStages;       // [2]
  uint64_t* d_empty = d_full + 2;                // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (batch + 127) / 128;

  if (threadIdx.x < 64) sBias[threadIdx.x] = (threadIdx.x < n_out) ? bias[threadIdx.x] : 0.0f;
  if (threadIdx.x < 12) sXthr[threadIdx.x] = threadIdx.x < 9 ? tb.xthr[threadIdx.x] : INFINITY;     // written by kan_tc_thresholds_kernel on this stream
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmWhi);
    tma_prefetch_desc(&tmWlo);
    for (int i = 0; i < kTcAStages; ++i) { mbar_init(&a_full[i], 16); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kTcWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 16); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================================================================= weight producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        for (int c = 0; c < num_chunks; ++c) {
          mbar_wait(&w_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&w_full[s], kTcWStageBytes);
          tma_load_2d(sW + s * kTcWStageBytes, &tmWhi, &w_full[s], c * 64, 0);
          tma_load_2d(sW + s * kTcWStageBytes + 8192, &tmWlo, &w_full[s], c * 64, 0);
          if (++s == kTcWStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= UMMA issuer (whole warp, one elected lane issues)
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const bool issuer = elect_one();
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA)), w_lo0 = umma_desc_lo(smem_u32(sW));
    int sa = 0, sw = 0;
    uint32_t pha = 0, phw = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&d_empty[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * 64;
      for (int c = 0; c < num_chunks; ++c) {
        mbar_wait(&a_full[sa], pha);
        mbar_wait(&w_full[sw], phw);
        tc_fence_after();
        const uint32_t ahi = a_lo0 + sa * (kTcAStageBytes >> 4), alo = ahi + (16384 >> 4);
        const uint32_t whi = w_lo0 + sw * (kTcWStageBytes >> 4), wlo = whi + (8192 >> 4);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_split<1>(d, ahi + 2 * k, whi + 2 * k, idesc, (c | k) != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_split<1>(d, ahi + 2 * k, wlo + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_split<1>(d, alo + 2 * k, whi + 2 * k, idesc, 1u);
          umma_commit(&a_empty[sa]);
          umma_commit(&w_empty[sw]);
          if (c == num_chunks - 1) umma_commit(&d_full[acc]);
        }
        __syncwarp();
        if (++sa == kTcAStages) { sa = 0; pha ^= 1; }
        if (++sw == kTcWStages) { sw = 0; phw ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  } else {
    // ================================================================= activation producers + epilogue
    const int pw = warp - 2;                              // 0..15
    const int pt = pw * 32 + lane;                        // 0..511
    const int srow = pt >> 2;                             // sample row of the tile: 4 threads share a sample
    const int ip = pt & 3;                                // inputs 2*ip, 2*ip+1 of the chunk (adjacent lanes: one 32-byte sector)
    const int quad = warp & 3, cg = pw >> 2;              // epilogue: TMEM lane quadrant (= warp % 4) and 16-column group
    const int erow = quad * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    int sa = 0;
    uint32_t pha = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int sg = t * 128 + srow;
      const float* xr = x + static_cast<size_t>(sg < batch ? sg : 0) * n_in + 2 * ip;
      auto load_x = [&](int c) -> float2 {
        float2 xv = make_float2(0.0f, 0.0f);
        const int i0 = c * 8 + 2 * ip;
        if (sg < batch && c < num_chunks) {
          if (i0 + 1 < n_in) xv = *reinterpret_cast<const float2*>(xr + c * 8);
          else if (i0 < n_in) xv.x = xr[c * 8];
        }
        return xv;
      };
      {   // pull the NEXT tile's 128 input rows (contiguous in x) into L2 while this tile is expanded: with a single
          // 8-byte load in flight per thread the x stream was latency-bound at 0.38 TB/s (ncu: 17 % of all samples on the first use)
        const long long tn = static_cast<long long>(t) + gridDim.x;
        if (tn < num_tiles) {
          const char* base = reinterpret_cast<const char*>(x) + tn * 128 * static_cast<long long>(n_in) * 4;
          const long long bytes = min(128LL, static_cast<long long>(batch) - tn * 128) * n_in * 4;
          for (long long o = static_cast<long long>(pt) * 128; o < bytes; o += 512 * 128) prefetch_l2(base + o);
        }
      }
      // the inputs of the next four-chunk GROUP are in flight while one group is expanded.  Two register sets with fixed
      // roles (xa: chunks c0..c0+3, xb: c0+4..c0+7): a rotating queue `xq[k] = load(c + 4)` made the compiler funnel every load
      // through one temporary pair and copy it into place at the end of the SAME chunk, i.e. one chunk of distance and a
      // long-scoreboard stall per chunk (ncu source view: 13 % of the producers' samples on the first use of x)
      float2 xa[4], xb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) xa[k] = load_x(k);
      auto expand_group = [&](const float2 (&xg)[4], int cbase) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = cbase + k;
          if (c < num_chunks) {                        // warp-uniform
            mbar_wait(&a_empty[sa], pha ^ 1);
            uint8_t* ahi = sA + sa * kTcAStageBytes;
            uint8_t* alo = ahi + 16384;
            kan_tc_expand_store(xg[k].x, sw128_offset(srow, 2 * ip), ahi, alo, sXthr);
            kan_tc_expand_store(xg[k].y, sw128_offset(srow, 2 * ip + 1), ahi, alo, sXthr);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[sa]);
            if (++sa == kTcAStages) { sa = 0; pha ^= 1; }
          }
        }
      };
#pragma unroll 1
      for (int c0 = 0; c0 < num_chunks; c0 += 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k) xb[k] = load_x(c0 + 4 + k);
        expand_group(xa, c0);
#pragma unroll
        for (int k = 0; k < 4; ++k) xa[k] = load_x(c0 + 8 + k);
        expand_group(xb, c0 + 4);
      }
      // ---- epilogue of this tile
      mbar_wait(&d_full[acc], acc_ph);
      tc_fence_after();
      float v[16];
      {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem_base + acc * 64 + cg * 16 + lane_sel)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_empty[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      const int eg = t * 128 + erow;
      if (eg < batch) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float u = v[i] + sBias[cg * 16 + i];
          if (act == 1) u = fmaxf(u, 0.0f);
          else if (act == 2) u = 3.0f / (1.0f + expf(-u));
          v[i] = u;
        }
        float* yr = y + static_cast<size_t>(eg) * n_out + cg * 16;
        if (n_out == 64) {
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(yr + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cg * 16 + i < n_out) yr[i] = v[i];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

// =====================================================================================================================
// Tensor-core backward w.r.t. the layer input:
//     T[b, i*8+k] = sum_o g[b,o] * Wp[i*8+k, o],   g = gy * act'(y)         GEMM [B x out] x [out x 8*in], split hi + lo
//     dx[b, i]    = T[b, i*8+7] + (1 - t^2) * sum_k N'_k(t) * T[b, i*8+k],   t = tanh x[b,i]
// Per 128-sample tile the sixteen worker warps build the K-major swizzled operand tile of g (hi | lo) once; the issuer
// then walks the input chunks (8 inputs = 64 packed rows of Wp = one UMMA N tile) with two TMEM accumulators, and the
// workers contract each finished 128 x 64 block of T against the basis derivatives straight out of TMEM (thread = one
// sample row x two inputs), so T never exists in memory either.  Wp2 hi / lo: bf16 [8*in_pad][64] (packed row major).
__device__ __forceinline__ void kan_tmem_ld16(uint32_t taddr, float (&v)[16]) {
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

constexpr int kTcGBufBytes = 2 * 16384;             // g_hi | g_lo, each [128 x 64] bf16
constexpr int kTcBxSmemBytes = 1024 + 2 * kTcGBufBytes + kTcWStages * kTcWStageBytes + 32 * 4 + 256;

// Wp (fp32 [kp][out_pad=64]) -> bf16 hi / lo, row-major [kp][64] each
__global__ void kan_split_weights_rows_kernel(const float* __restrict__ spline, const float* __restrict__ lin_w, int n_in,
                                              int n_out, int kp, __nv_bfloat16* __restrict__ w_hi,
                                              __nv_bfloat16* __restrict__ w_lo) {
  const long long total = 64LL * kp;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kk = static_cast<int>(idx >> 6), o = static_cast<int>(idx & 63);
    const int i = kk >> 3, k = kk & 7;
    float v = 0.0f;
    if (i < n_in && o < n_out)
      v = (k < 7) ? spline[(static_cast<size_t>(i) * n_out + o) * 7 + k] : lin_w[static_cast<size_t>(o) * n_in + i];
    const __nv_bfloat16 hi = __float2bfloat16(v);
    w_hi[idx] = hi;
    w_lo[idx] = __float2bfloat16(v - __bfloat162float(hi));
  }
}

__global__ void __launch_bounds__(kTcThreads, 1)
kan_bwd_x_tc_kernel(const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                    const float* __restrict__ x, const float* __restrict__ yv, const float* __restrict__ gy, const KanTcTables tb,
                    float* __restrict__ dx, int act, int batch, int n_in, int n_out, int num_chunks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sG = smem;
  uint8_t* sW = sG + 2 * kTcGBufBytes;
  float* sXthr = reinterpret_cast<float*>(sW + kTcWStages * kTcWStageBytes);     // [12]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXthr + 32);
  uint64_t* g_full = bars;                       // [2]
  uint64_t* g_empty = g_full + 2;                // [2]
  uint64_t* w_full = g_empty + 2;                // [4]
  uint64_t* w_empty = w_full + kTcWStages;       // [4]
  uint64_t* d_full = w_empty + kTcWStages;       // [2]
  uint64_t* d_empty = d_full + 2;                // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (batch + 127) / 128;

  if (threadIdx.x < 12) sXthr[threadIdx.x] = threadIdx.x < 9 ? tb.xthr[threadIdx.x] : INFINITY;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmWhi);
    tma_prefetch_desc(&tmWlo);
    for (int i = 0; i < 2; ++i) { mbar_init(&g_full[i], 16); mbar_init(&g_empty[i], 1); }
    for (int i = 0; i < kTcWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 16); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================================================================= weight producer: Wp rows [c*64, c*64+64), hi | lo
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        for (int c = 0; c < num_chunks; ++c) {
          mbar_wait(&w_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&w_full[s], kTcWStageBytes);
          tma_load_2d(sW + s * kTcWStageBytes, &tmWhi, &w_full[s], 0, c * 64);
          tma_load_2d(sW + s * kTcWStageBytes + 8192, &tmWlo, &w_full[s], 0, c * 64);
          if (++s == kTcWStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= UMMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const bool issuer = elect_one();
    const uint32_t g_lo0 = umma_desc_lo(smem_u32(sG)), w_lo0 = umma_desc_lo(smem_u32(sW));
    int sw = 0, acc = 0, it = 0;
    uint32_t phw = 0, acc_ph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int gb = it & 1;
      mbar_wait(&g_full[gb], (it >> 1) & 1);
      const uint32_t ghi = g_lo0 + gb * (kTcGBufBytes >> 4), glo = ghi + (16384 >> 4);
      for (int c = 0; c < num_chunks; ++c) {
        mbar_wait(&w_full[sw], phw);
        mbar_wait(&d_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 64;
        const uint32_t whi = w_lo0 + sw * (kTcWStageBytes >> 4), wlo = whi + (8192 >> 4);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_split<1>(d, ghi + 2 * k, whi + 2 * k, idesc, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_split<1>(d, ghi + 2 * k, wlo + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_split<1>(d, glo + 2 * k, whi + 2 * k, idesc, 1u);
          umma_commit(&w_empty[sw]);
          umma_commit(&d_full[acc]);
          if (c == num_chunks - 1) umma_commit(&g_empty[gb]);
        }
        __syncwarp();
        if (++sw == kTcWStages) { sw = 0; phw ^= 1; }
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // ================================================================= workers: g tile builder + derivative contraction
    const int pw = warp - 2;                              // 0..15
    const int pt = pw * 32 + lane;                        // 0..511
    const int srow = pt >> 2, gq = pt & 3;                // g tile: sample row, 16-column quarter
    const int quad = warp & 3, cg = pw >> 2;              // T blocks: TMEM lane quadrant (= warp % 4), inputs 2cg, 2cg+1 of a chunk
    const int erow = quad * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    int acc = 0, it = 0;
    uint32_t acc_ph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      // ---- g = gy * act'(y) -> hi | lo operand tile
      {
        const int gb = it & 1;
        const int sg = t * 128 + srow;
        float g[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) g[i] = 0.0f;
        if (sg < batch) {
          const size_t off = static_cast<size_t>(sg) * n_out + gq * 16;
          if (n_out == 64) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 a = *reinterpret_cast<const float4*>(gy + off + 4 * q);
              const float4 b = *reinterpret_cast<const float4*>(yv + off + 4 * q);
              g[4 * q + 0] = a.x * act_grad(act, b.x); g[4 * q + 1] = a.y * act_grad(act, b.y);
              g[4 * q + 2] = a.z * act_grad(act, b.z); g[4 * q + 3] = a.w * act_grad(act, b.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (gq * 16 + i < n_out) g[i] = gy[off + i] * act_grad(act, yv[off + i]);
          }
        }
        mbar_wait(&g_empty[gb], ((it >> 1) & 1) ^ 1);
        uint8_t* ghi = sG + gb * kTcGBufBytes;
        uint8_t* glo = ghi + 16384;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t wh[4], wl[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = g[h * 8 + 2 * e], b = g[h * 8 + 2 * e + 1];      // two values per (ALU-pipe) packed conversion
            wh[e] = pack_bf16x2(a, b);
            wl[e] = pack_bf16x2(a - __uint_as_float(wh[e] << 16), b - __uint_as_float(wh[e] & 0xffff0000u));
          }
          const uint32_t off = sw128_offset(srow, gq * 2 + h);
          *reinterpret_cast<uint4*>(ghi + off) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
          *reinterpret_cast<uint4*>(glo + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&g_full[gb]);
      }
      // ---- per chunk: T block out of TMEM, contracted with the basis derivatives
      const int eg = t * 128 + erow;
      const bool live = eg < batch;
      const float* xr = x + static_cast<size_t>(live ? eg : 0) * n_in + 2 * cg;
      float* dxr = dx + static_cast<size_t>(live ? eg : 0) * n_in + 2 * cg;
      {   // next tile's x rows -> L2
        const long long tn = static_cast<long long>(t) + gridDim.x;
        if (tn < num_tiles) {
          const char* base = reinterpret_cast<const char*>(x) + tn * 128 * static_cast<long long>(n_in) * 4;
          const long long bytes = min(128LL, static_cast<long long>(batch) - tn * 128) * n_in * 4;
          for (long long o = static_cast<long long>(pt) * 128; o < bytes; o += 512 * 128) prefetch_l2(base + o);
        }
      }
      // the inputs of the next four-chunk group are in flight while one group is contracted (two register sets with fixed
      // roles, see kan_fwd_tc_kernel: a rotating queue degenerates into one chunk of prefetch distance)
      auto load_x = [&](int c) -> float2 {
        return (live && c < num_chunks) ? *reinterpret_cast<const float2*>(xr + c * 8) : make_float2(0.f, 0.f);
      };
      float2 xa[4], xb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) xa[k] = load_x(k);
      auto contract_group = [&](const float2 (&xg)[4], int cbase) {
#pragma unroll
      for (int kq = 0; kq < 4; ++kq) {
        const int c = cbase + kq;
        if (c >= num_chunks) break;                    // warp-uniform
        const float2 xv = xg[kq];
        mbar_wait(&d_full[acc], acc_ph);
        tc_fence_after();
        float T[16];
        kan_tmem_ld16(tmem_base + acc * 64 + cg * 16 + lane_sel, T);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_empty[acc]);
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        float out[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float xe = e == 0 ? xv.x : xv.y;
          float rc;
          const float tt = kan_tc_tanh(xe, rc);
          const float rc4 = 4.0f * rc;
          const float dt = fmaf(-rc4, rc, rc4);                        // 1 - tanh^2 = 4 rc (1 - rc), without cancellation
          const float s5 = fmaf(tt, 5.0f, 5.0f);
          int j;
          float jfl;
          kan_tc_interval(xe, s5, sXthr, j, jfl);
          // basis derivatives d/dt of the four live cubics (knot spacing 1/5 folded into the coefficients)
          const float u = s5 - jfl, om = 1.0f - u;
          const float d0 = 2.5f * u * u;                               // slot j
          const float d1 = fmaf(fmaf(-7.5f, u, 5.0f), u, 2.5f);        // slot j-1
          const float d2 = u * fmaf(7.5f, u, -10.0f);                  // slot j-2
          const float d3 = -2.5f * om * om;                            // slot j-3
          // the four T columns j-3 .. j (zero below slot 0) through a 3-level select on the bits of j -- a barrel shifter in
          // registers: 16 selects instead of 28 compares + 28 predicated FMAs for the `slot == k` form (49 of ~125
          // instructions per (sample, input) in the round-2 profile)
          const float* Te = T + e * 8;
          const bool b2 = (j & 4) != 0, b1 = (j & 2) != 0, b0 = (j & 1) != 0;
          const float E[11] = {0.0f, 0.0f, 0.0f, Te[0], Te[1], Te[2], Te[3], Te[4], Te[5], Te[6], 0.0f};
          float F[7], G[5], H[4];
#pragma unroll
          for (int i = 0; i < 7; ++i) F[i] = b2 ? E[i + 4] : E[i];
#pragma unroll
          for (int i = 0; i < 5; ++i) G[i] = b1 ? F[i + 2] : F[i];
#pragma unroll
          for (int i = 0; i < 4; ++i) H[i] = b0 ? G[i + 1] : G[i];       // H[i] = T column j - 3 + i
          float sp = fmaf(d0, H[3], fmaf(d1, H[2], fmaf(d2, H[1], d3 * H[0])));
          sp = static_cast<unsigned>(j) < 7u ? sp : 0.0f;              // dead zone (and NaN inputs): linear branch only
          out[e] = fmaf(dt, sp, T[e * 8 + 7]);
        }
        if (live) *reinterpret_cast<float2*>(dxr + c * 8) = make_float2(out[0], out[1]);
      }
      };
#pragma unroll 1
      for (int c0 = 0; c0 < num_chunks; c0 += 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k) xb[k] = load_x(c0 + 4 + k);
        contract_group(xa, c0);
#pragma unroll
        for (int k = 0; k < 4; ++k) xa[k] = load_x(c0 + 8 + k);
        contract_group(xb, c0 + 4);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

// =====================================================================================================================
// Tensor-core weight gradient:
//     dWp[i*8+k, o] += sum_b A[b, i*8+k] * g[b, o],   dlin_b[o] += sum_b g[b, o],      g = gy * act'(y)
// a GEMM whose reduction index is the SAMPLE: both operands are MN-major UMMA operands (gemm_tn.cuh), and the tiles the
// forward kernel generates -- [128 samples x 64 packed rows], 128-byte rows, 8-row swizzle atoms -- are exactly the
// canonical MN-major atom stack, so the same expansion code feeds this kernel.  A CTA owns a group of 64 inputs (four
// M = 128 blocks of packed rows = four TMEM accumulators that live for the whole kernel) and every (grid.x)-th 128-sample
// tile; per tile it expands its 64 inputs (two chunks = one M block per stage, hi | lo) and the g tile (hi | lo), and
// the issuer adds  A_hi g_hi + A_hi g_lo + A_lo g_hi  over the 8 K steps of 16 samples.  At the end the accumulators
// are added into the fp32 packed gradient with red.global.add (split over the batch).
constexpr int kTcWgAStageBytes = 4 * 16384;           // [hi | lo] x two 64-row panels of [128 samples x 64] bf16
constexpr int kTcWgSmemBytes = 1024 + 2 * kTcWgAStageBytes + 2 * kTcGBufBytes + 64 * 4 + 32 * 4 + 256;

__global__ void __launch_bounds__(kTcThreads, 1)
kan_bwd_w_tc_kernel(const float* __restrict__ x, const float* __restrict__ yv, const float* __restrict__ gy, const KanTcTables tb,
                    float* __restrict__ dWp, float* __restrict__ dlin_b, int act, int batch, int n_in, int n_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sG = sA + 2 * kTcWgAStageBytes;
  float* sDb = reinterpret_cast<float*>(sG + 2 * kTcGBufBytes);     // [64]
  float* sXthr = sDb + 64;       // [12]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXthr + 32);
  uint64_t* a_full = bars;                       // [2]
  uint64_t* a_empty = a_full + 2;                // [2]
  uint64_t* g_full = a_empty + 2;                // [2]
  uint64_t* g_empty = g_full + 2;                // [2]
  uint64_t* acc_full = g_empty + 2;              // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (batch + 127) / 128;
  const int group = blockIdx.y;                  // inputs [64*group, 64*group + 64)
  const int n_my = (blockIdx.x < num_tiles) ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x < 64) sDb[threadIdx.x] = 0.0f;
  if (threadIdx.x < 12) sXthr[threadIdx.x] = threadIdx.x < 9 ? tb.xthr[threadIdx.x] : INFINITY;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 16); mbar_init(&a_empty[i], 1);
      mbar_init(&g_full[i], 16); mbar_init(&g_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 1) {
    // ================================================================= UMMA issuer (both operands MN-major)
    if (lane == 0 && n_my > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      int sa = 0;
      uint32_t pha = 0;
      for (int it = 0; it < n_my; ++it) {
        const int gb = it & 1;
        mbar_wait(&g_full[gb], (it >> 1) & 1);
        const uint32_t g_hi = smem_u32(sG + gb * kTcGBufBytes), g_lo = g_hi + 16384;
        for (int pr = 0; pr < 4; ++pr) {
          mbar_wait(&a_full[sa], pha);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(sA + sa * kTcWgAStageBytes), a_lo = a_hi + 32768;
          const uint32_t d = tmem_base + pr * 64;
#pragma unroll
          for (int k = 0; k < 8; ++k) {           // 16 samples per UMMA = two 8-row swizzle atoms
            const uint64_t ah = umma_smem_desc(a_hi + k * 2048, 16384, 1024);
            const uint64_t al = umma_smem_desc(a_lo + k * 2048, 16384, 1024);
            const uint64_t gh = umma_smem_desc(g_hi + k * 2048, 16384, 1024);
            const uint64_t gl = umma_smem_desc(g_lo + k * 2048, 16384, 1024);
            umma_bf16(d, ah, gh, idesc, (it | k) != 0 ? 1u : 0u);
            umma_bf16(d, ah, gl, idesc, 1u);
            umma_bf16(d, al, gh, idesc, 1u);
          }
          umma_commit(&a_empty[sa]);
          if (pr == 3) umma_commit(&g_empty[gb]);
          if (++sa == 2) { sa = 0; pha ^= 1; }
        }
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 2) {
    // ================================================================= workers: operand generation, final reduction
    const int pw = warp - 2;                              // 0..15
    const int pt = pw * 32 + lane;                        // 0..511
    const int srow = pt >> 2, ip = pt & 3;                // sample row; A: inputs 2ip, 2ip+1 of a chunk; g: 16-column quarter ip
    int sa = 0;
    uint32_t pha = 0;
    float db[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) db[i] = 0.0f;
    // this thread's inputs of all eight chunks of a tile; the NEXT tile's are loaded while the current tile is expanded
    auto load_tile_x = [&](int it2, float2 (&xn)[8]) {
      const int sg2 = (blockIdx.x + it2 * static_cast<int>(gridDim.x)) * 128 + srow;
      const bool ok = it2 < n_my && sg2 < batch;
      const float* xr2 = x + static_cast<size_t>(ok ? sg2 : 0) * n_in + group * 64 + 2 * ip;
#pragma unroll
      for (int cn = 0; cn < 8; ++cn) xn[cn] = ok ? *reinterpret_cast<const float2*>(xr2 + cn * 8) : make_float2(0.f, 0.f);
    };
    float2 xq[8], xnext[8];
    load_tile_x(0, xq);
    for (int it = 0; it < n_my; ++it) {
      const int t = blockIdx.x + it * gridDim.x;
      const int sg = t * 128 + srow;
      const bool live = sg < batch;
      load_tile_x(it + 1, xnext);
      // ---- g tile (hi | lo), and this thread's share of the bias gradient
      {
        const int gb = it & 1;
        float g[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) g[i] = 0.0f;
        if (live) {
          const size_t off = static_cast<size_t>(sg) * n_out + ip * 16;
          if (n_out == 64) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 a = *reinterpret_cast<const float4*>(gy + off + 4 * q);
              const float4 b = *reinterpret_cast<const float4*>(yv + off + 4 * q);
              g[4 * q + 0] = a.x * act_grad(act, b.x); g[4 * q + 1] = a.y * act_grad(act, b.y);
              g[4 * q + 2] = a.z * act_grad(act, b.z); g[4 * q + 3] = a.w * act_grad(act, b.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (ip * 16 + i < n_out) g[i] = gy[off + i] * act_grad(act, yv[off + i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) db[i] += g[i];
        mbar_wait(&g_empty[gb], ((it >> 1) & 1) ^ 1);
        uint8_t* ghi = sG + gb * kTcGBufBytes;
        uint8_t* glo = ghi + 16384;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t wh[4], wl[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = g[h * 8 + 2 * e], b = g[h * 8 + 2 * e + 1];      // two values per (ALU-pipe) packed conversion
            wh[e] = pack_bf16x2(a, b);
            wl[e] = pack_bf16x2(a - __uint_as_float(wh[e] << 16), b - __uint_as_float(wh[e] & 0xffff0000u));
          }
          const uint32_t off = sw128_offset(srow, ip * 2 + h);
          *reinterpret_cast<uint4*>(ghi + off) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
          *reinterpret_cast<uint4*>(glo + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&g_full[gb]);
      }
      // ---- expanded activations of this CTA's 64 inputs: four stages of two chunks
      {   // the x rows of the tile after the next one (this group's 256-byte slice of each) -> L2
        const int tn = t + 2 * gridDim.x;
        if (tn < num_tiles && ip == 0 && tn * 128 + srow < batch)
          prefetch_l2(x + static_cast<size_t>(tn * 128 + srow) * n_in + group * 64);
      }
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        mbar_wait(&a_empty[sa], pha ^ 1);
        uint8_t* stage = sA + sa * kTcWgAStageBytes;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 xv = xq[pr * 2 + h];
          uint8_t* ahi = stage + h * 16384;
          uint8_t* alo = ahi + 32768;
          if (live) {
            kan_tc_expand_store(xv.x, sw128_offset(srow, 2 * ip), ahi, alo, sXthr);
            kan_tc_expand_store(xv.y, sw128_offset(srow, 2 * ip + 1), ahi, alo, sXthr);
          } else {                                  // rows past the batch must not contribute (raw-x slot of x = 0 is 0 anyway,
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);   // but the spline slots of tanh(0) are not)
            *reinterpret_cast<uint4*>(ahi + sw128_offset(srow, 2 * ip)) = z;
            *reinterpret_cast<uint4*>(alo + sw128_offset(srow, 2 * ip)) = z;
            *reinterpret_cast<uint4*>(ahi + sw128_offset(srow, 2 * ip + 1)) = z;
            *reinterpret_cast<uint4*>(alo + sw128_offset(srow, 2 * ip + 1)) = z;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[sa]);
        if (++sa == 2) { sa = 0; pha ^= 1; }
      }
#pragma unroll
      for (int cn = 0; cn < 8; ++cn) xq[cn] = xnext[cn];
    }
    // ---- bias gradient: columns ip*16 .. +15, summed over this thread's rows -> shared -> global (group 0 only)
    if (group == 0 && dlin_b != nullptr) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float v = db[i];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (lane < 4) atomicAdd(&sDb[ip * 16 + i], v);
      }
    }
    // ---- accumulators -> fp32 packed gradient (split over the batch: red.global.add)
    if (n_my > 0) {
      const int quad = warp & 3, cg = pw >> 2;
      const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
      mbar_wait(acc_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int pr = 0; pr < 4; ++pr) {
        float v[16];
        kan_tmem_ld16(tmem_base + pr * 64 + cg * 16 + lane_sel, v);
        float* dst = dWp + (static_cast<size_t>(group) * 512 + pr * 128 + quad * 32 + lane) * 64 + cg * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 4)      // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the L2 transactions
          atomicAdd(reinterpret_cast<float4*>(dst + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
      }
    }
    named_bar_sync(1, 16 * 32);
    if (group == 0 && dlin_b != nullptr && pt < n_out) atomicAdd(&dlin_b[pt], sDb[pt]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}
