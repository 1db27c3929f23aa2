// HBM-bound kernels of the DeiT-Tiny token stream: patch extraction, the folded cls/pos/bias table,
// LayerNorm forward/backward (warp per 192-wide row), bf16 casts, column sums for bias gradients and
// the batch reduction of the token-0 gradients.  All rows are 192 fp32 = 768 B = 6 coalesced 128 B
// lines; a warp owns a row and each lane 6 columns (lane + 32*i), so every access is a full line.
#include <cuda_fp16.h>

#include "kernels.h"

namespace {

constexpr int kD = 192;
constexpr int kTok = 197;
constexpr int kPatchK = 768;

// ------------------------------------------------------------------ patch extraction (im2col)
// One warp per (image, token, channel): reads 16 image-row segments of 64 B, writes 512 contiguous
// bytes of the bf16 patch matrix [B*197, 768] (column = c*256 + ky*16 + kx, the flattening of the
// Conv2d(3,192,16,16) weight).  Token 0 (cls slot) is an all-zero row: the class token enters
// through the additive token table, so the patch GEMM can emit the [B*197,192] stream directly.
// IN_BF16: the images are already bf16 (a serving path that halves the host->device copy; bit-identical result, the
// fp32 path rounds the pixels to bf16 here anyway)
// FMT 2: uint8 NCHW pixels (what an image decoder produces: a quarter of the host->device bytes); the per-channel
// normalisation (pixel / 255 - mean) / std is applied here as pixel * scale[c] + shift[c]
struct PixelNorm { float scale[3]; float shift[3]; };
template <int FMT>
__global__ void im2col_kernel(const void* __restrict__ img_v, __nv_bfloat16* __restrict__ out, int batch, PixelNorm nrm) {
  const int lane = threadIdx.x & 31;
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long total = static_cast<long long>(batch) * kTok * 3;
  if (w >= total) return;
  const int c = static_cast<int>(w % 3);
  const int tok = static_cast<int>((w / 3) % kTok);
  const int b = static_cast<int>(w / (3 * kTok));
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  const int ky = lane >> 1, half = lane & 1;
  if (tok > 0) {
    const int p = tok - 1, py = p / 14, px = p % 14;
    const size_t off = ((static_cast<size_t>(b) * 3 + c) * 224 + (py * 16 + ky)) * 224 + px * 16 + half * 8;
    if (FMT == 1) {
      v = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(img_v) + off);
    } else if (FMT == 2) {
      const uint2 q = *reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(img_v) + off);
      const float sc = nrm.scale[c], sh = nrm.shift[c];
      auto px = [&](uint32_t word, int byte) { return fmaf(static_cast<float>((word >> (8 * byte)) & 0xffu), sc, sh); };
      v = make_uint4(pack_bf16x2(px(q.x, 0), px(q.x, 1)), pack_bf16x2(px(q.x, 2), px(q.x, 3)),
                     pack_bf16x2(px(q.y, 0), px(q.y, 1)), pack_bf16x2(px(q.y, 2), px(q.y, 3)));
    } else {
      const float* src = static_cast<const float*>(img_v) + off;
      const float4 a = *reinterpret_cast<const float4*>(src);
      const float4 d = *reinterpret_cast<const float4*>(src + 4);
      v = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
    }
  }
  __nv_bfloat16* dst = out + (static_cast<size_t>(b) * kTok + tok) * kPatchK + c * 256 + ky * 16 + half * 8;
  *reinterpret_cast<uint4*>(dst) = v;
}

// table[0] = cls_token + pos[0]; table[i] = patch_bias + pos[i]
__global__ void token_table_kernel(const float* __restrict__ cls, const float* __restrict__ pos,
                                   const float* __restrict__ pbias, float* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kTok * kD) return;
  const int tok = i / kD, c = i % kD;
  table[i] = pos[i] + (tok == 0 ? cls[c] : pbias[c]);
}

// All weight shadows of the trunk in one launch (was ~97 per optimizer step): every job is a [rows x cols] fp32 matrix cast
// tile by tile (32 x 32) to bf16, to bf16 transposed [cols x rows] (dgrad operands), or to fp16 (fc2 of the fused MLP kernel).
__global__ void __launch_bounds__(256) cast_multi_kernel(const __grid_constant__ RvkCastTable T) {
  __shared__ float tile[32][33];
  __shared__ int s_j;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    int lo = 0, hi = T.n - 1;
    const int b = static_cast<int>(blockIdx.x);
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (T.job[mid].tile_start <= b) lo = mid; else hi = mid - 1;
    }
    s_j = lo;
  }
  __syncthreads();
  const RvkCastJob& J = T.job[s_j];
  const int local = static_cast<int>(blockIdx.x) - J.tile_start;
  const int tiles_x = (J.cols + 31) / 32;
  const int r0 = (local / tiles_x) * 32, c0 = (local % tiles_x) * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;          // block (32, 8)
  if (J.mode == 1) {
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      tile[i][tx] = (r < J.rows && c < J.cols) ? J.src[static_cast<size_t>(r) * J.cols + c] : 0.0f;
    }
    __syncthreads();
    auto* dst = static_cast<__nv_bfloat16*>(J.dst);
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;
      if (r < J.rows && c < J.cols) dst[static_cast<size_t>(c) * J.rows + r] = __float2bfloat16(tile[tx][i]);
    }
  } else {
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      if (r < J.rows && c < J.cols) {
        const size_t o = static_cast<size_t>(r) * J.cols + c;
        if (J.mode == 0) static_cast<__nv_bfloat16*>(J.dst)[o] = __float2bfloat16(J.src[o]);
        else static_cast<__half*>(J.dst)[o] = __float2half(J.src[o]);
      }
    }
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
  }
}

__global__ void cast_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(src[i]);
}

// dst[c][r] = bf16(src[r][c]) through a padded 32x32 tile
__global__ void cast_transpose_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows,
                                      int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[static_cast<size_t>(r) * cols + c] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[static_cast<size_t>(c) * rows + r] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

// ------------------------------------------------------------------ LayerNorm forward
// TILED: x is the tiled fp32 token stream (common.cuh xt_offset) and row r of this launch is token row r * xs
template <bool OUT_BF16, bool TILED = false>
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, long long xs, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, void* __restrict__ y, long long ys,
                                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    const float* xr = x + static_cast<size_t>(r) * xs;
    float v[6];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      v[i] = TILED ? x[xt_elem_offset(static_cast<int>(r * xs), lane + 32 * i)] : xr[lane + 32 * i];
      s += v[i];
    }
    const float mean = warp_sum(s) * (1.0f / kD);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 6; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + eps);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = lane + 32 * i;
      const float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
      if (OUT_BF16) static_cast<__nv_bfloat16*>(y)[static_cast<size_t>(r) * ys + c] = __float2bfloat16(o);
      else static_cast<float*>(y)[static_cast<size_t>(r) * ys + c] = o;
    }
    if (lane == 0) {
      if (mean_out != nullptr) mean_out[r] = mean;
      if (rstd_out != nullptr) rstd_out[r] = rstd;
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)) (+ dx_in); dgamma += g*xhat; dbeta += g;
// dcolsum += column sums of the dx written (the bias gradient of the Linear layer whose output gradient dx is: fc2 / proj
// bias -- saves a separate pass over dx).
// A streaming kernel: 16 lanes per row, 16-byte accesses (lane l owns columns 4l + 64j .. +3, j = 0..2), two rows per warp
// per iteration, all loads of a row issued before the first reduction.
template <bool G_BF16>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const void* __restrict__ g, long long gs, const float* __restrict__ x, long long xs,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const float* dx_in, float* dx_out, long long dxs, __nv_bfloat16* __restrict__ dx_bf16,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcolsum, int rows) {
  __shared__ float s_part[8][3][kD];   // per-warp column sums (blockDim.x == 256)
  griddep_wait();                    // (programmatic dependent launch)
  griddep_launch_dependents();
  const int hl = threadIdx.x & 15;                                   // lane within the half-warp that shares a row
  const int half = threadIdx.x >> 4;                                 // row slot within the block
  const int slots = blockDim.x >> 4;
  float4 gam[3];
  float dg[12], db[12], ds[12];
#pragma unroll
  for (int j = 0; j < 3; ++j) gam[j] = *reinterpret_cast<const float4*>(gamma + 4 * hl + 64 * j);
#pragma unroll
  for (int i = 0; i < 12; ++i) { dg[i] = 0.0f; db[i] = 0.0f; ds[i] = 0.0f; }
  const int n_iter = (rows + gridDim.x * slots - 1) / (gridDim.x * slots);      // uniform trip count: shuffles stay converged
  for (int itn = 0; itn < n_iter; ++itn) {
    const int r = (itn * gridDim.x + blockIdx.x) * slots + half;
    const bool live = r < rows;
    const size_t rr = live ? static_cast<size_t>(r) : 0;
    float gv[12], xh[12], din[12];
    const float mu = mean[rr], rs = rstd[rr];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int c = 4 * hl + 64 * j;
      if (G_BF16) {
        const uint2 w = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(g) + rr * gs + c);
        gv[4 * j + 0] = __uint_as_float(w.x << 16); gv[4 * j + 1] = __uint_as_float(w.x & 0xffff0000u);
        gv[4 * j + 2] = __uint_as_float(w.y << 16); gv[4 * j + 3] = __uint_as_float(w.y & 0xffff0000u);
      } else {
        const float4 w = *reinterpret_cast<const float4*>(static_cast<const float*>(g) + rr * gs + c);
        gv[4 * j + 0] = w.x; gv[4 * j + 1] = w.y; gv[4 * j + 2] = w.z; gv[4 * j + 3] = w.w;
      }
      const float4 xv = *reinterpret_cast<const float4*>(x + rr * xs + c);
      xh[4 * j + 0] = xv.x; xh[4 * j + 1] = xv.y; xh[4 * j + 2] = xv.z; xh[4 * j + 3] = xv.w;
      float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dx_in != nullptr) dv = *reinterpret_cast<const float4*>(dx_in + rr * dxs + c);
      din[4 * j + 0] = dv.x; din[4 * j + 1] = dv.y; din[4 * j + 2] = dv.z; din[4 * j + 3] = dv.w;
    }
    float c1 = 0.0f, c2 = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float gm[4] = {gam[j].x, gam[j].y, gam[j].z, gam[j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = 4 * j + e;
        if (!live) gv[i] = 0.0f;
        xh[i] = (xh[i] - mu) * rs;
        const float gg = gv[i] * gm[e];
        c1 += gg;
        c2 = fmaf(gg, xh[i], c2);
        dg[i] = fmaf(gv[i], xh[i], dg[i]);
        db[i] += gv[i];
      }
    }
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) {
      c1 += __shfl_xor_sync(0xffffffffu, c1, d);
      c2 += __shfl_xor_sync(0xffffffffu, c2, d);
    }
    c1 *= (1.0f / kD);
    c2 *= (1.0f / kD);
    if (live) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int c = 4 * hl + 64 * j;
        const float gm[4] = {gam[j].x, gam[j].y, gam[j].z, gam[j].w};
        float d[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 4 * j + e;
          d[e] = rs * (gv[i] * gm[e] - c1 - xh[i] * c2) + din[i];
          ds[i] += d[e];
        }
        *reinterpret_cast<float4*>(dx_out + rr * dxs + c) = make_float4(d[0], d[1], d[2], d[3]);
        if (dx_bf16 != nullptr) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(d[0], d[1]), hi = __floats2bfloat162_rn(d[2], d[3]);
          *reinterpret_cast<uint2*>(dx_bf16 + rr * dxs + c) =
              make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
      }
    }
  }
  // column sums over the rows of this block: the two half-warps of a warp by one shuffle, the 8 warps through shared memory
  // (round 1 used shared-memory atomicAdd here: a compare-and-swap loop on sm_100, 16 row slots contending for every column)
  const int wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    dg[i] += __shfl_xor_sync(0xffffffffu, dg[i], 16);
    db[i] += __shfl_xor_sync(0xffffffffu, db[i], 16);
    ds[i] += __shfl_xor_sync(0xffffffffu, ds[i], 16);
  }
  if ((threadIdx.x & 16) == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int c = 4 * hl + 64 * j;
      *reinterpret_cast<float4*>(&s_part[wid][0][c]) = make_float4(dg[4 * j], dg[4 * j + 1], dg[4 * j + 2], dg[4 * j + 3]);
      *reinterpret_cast<float4*>(&s_part[wid][1][c]) = make_float4(db[4 * j], db[4 * j + 1], db[4 * j + 2], db[4 * j + 3]);
      *reinterpret_cast<float4*>(&s_part[wid][2][c]) = make_float4(ds[4 * j], ds[4 * j + 1], ds[4 * j + 2], ds[4 * j + 3]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * kD; i += blockDim.x) {
    const int which = i / kD, c = i % kD;
    if (which < 2 ? dgamma == nullptr : dcolsum == nullptr) continue;
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_part[w][which][c];
    atomicAdd(which == 0 ? &dgamma[c] : which == 1 ? &dbeta[c] : &dcolsum[c], t);
  }
}

// ------------------------------------------------------------------ column sums (bias gradients)
// out[c] += scale * sum_r src[r, c].  A pure streaming reduction: 16-byte loads, `tpr` threads across a row and
// blockDim.x / tpr rows in flight per block iteration (4 independent loads per thread), partial sums combined through
// shared memory, one atomicAdd per column and block.  cols * elem_size must be a multiple of 16 (768 / 576 bf16, 192 fp32).
template <bool SRC_BF16>
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ src, long long ld, int rows, int cols, float* __restrict__ out, float scale,
              int rows_per_block) {
  constexpr int VE = SRC_BF16 ? 8 : 4;                 // elements per 16-byte load
  __shared__ float sPart[256 * VE];
  const int tpr = cols / VE;                           // threads across a row
  const int ry = blockDim.x / tpr;                     // rows in flight
  const int tx = threadIdx.x % tpr, ty = threadIdx.x / tpr;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float acc[VE];
#pragma unroll
  for (int e = 0; e < VE; ++e) acc[e] = 0.0f;
  if (ty < ry) {
    const size_t esz = SRC_BF16 ? 2 : 4;
    const uint8_t* base = static_cast<const uint8_t*>(src) + static_cast<size_t>(tx) * 16;
    auto add_row = [&](int r) {
      const uint4 q = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(r) * ld * esz);
      if (SRC_BF16) {
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 v = unpack_bf16x2(w[e]); acc[(2 * e) % VE] += v.x; acc[(2 * e + 1) % VE] += v.y; }
      } else {
        acc[0] += __uint_as_float(q.x); acc[1] += __uint_as_float(q.y);
        acc[2 % VE] += __uint_as_float(q.z); acc[3 % VE] += __uint_as_float(q.w);
      }
    };
    int r = r0 + ty;
    for (; r + 3 * ry < r1; r += 4 * ry) { add_row(r); add_row(r + ry); add_row(r + 2 * ry); add_row(r + 3 * ry); }
    for (; r < r1; r += ry) add_row(r);
  }
#pragma unroll
  for (int e = 0; e < VE; ++e) sPart[threadIdx.x * VE + e] = acc[e];
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const int cx = c / VE, ce = c % VE;
    float t = 0.0f;
    for (int y = 0; y < ry; ++y) t += sPart[(y * tpr + cx) * VE + ce];
    atomicAdd(&out[c], t * scale);
  }
}

// ------------------------------------------------------------------ gradients of cls_token / pos_embed / patch bias
// S[tok, c] = sum_b dx0[b, tok, c];  dpos += S, dcls += S[0], dpatch_bias += sum_{tok>=1} S[tok]
__global__ void token_grad_reduce_kernel(const float* __restrict__ dx0, int batch, int b_per_block,
                                         float* __restrict__ dpos, float* __restrict__ dcls,
                                         float* __restrict__ dpbias) {
  const int tok = blockIdx.x;
  const int b0 = blockIdx.y * b_per_block, b1 = min(batch, b0 + b_per_block);
  const int c = threadIdx.x;
  float acc = 0.0f;
  for (int b = b0; b < b1; ++b) acc += dx0[(static_cast<size_t>(b) * kTok + tok) * kD + c];
  atomicAdd(&dpos[tok * kD + c], acc);
  // 196 tokens fold into one bias vector: reduce over the tokens of this block column first
  if (tok == 0) atomicAdd(&dcls[c], acc);
  else atomicAdd(&dpbias[c], acc);
}

}  // namespace

int rvk_im2col_launch(const void* images, int fmt, void* patches_bf16, int batch, const float* norm6_host, cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  const long long warps = static_cast<long long>(batch) * kTok * 3;
  const int threads = 256;
  const long long blocks = (warps * 32 + threads - 1) / threads;
  if (fmt < 0 || fmt > 2 || (fmt == 2 && norm6_host == nullptr)) return RVK_ERR_BAD_ARG;
  const int px_bytes = fmt == 0 ? 4 : (fmt == 1 ? 2 : 1);
  RvkScopedTimer timer(stream, 0.0, double(batch) * (3.0 * 224 * 224 * px_bytes + 197.0 * 768 * 2), RVK_T_IM2COL);
  PixelNorm nrm{};
  if (fmt == 2)
    for (int i = 0; i < 3; ++i) { nrm.scale[i] = norm6_host[i]; nrm.shift[i] = norm6_host[3 + i]; }
  auto* out = static_cast<__nv_bfloat16*>(patches_bf16);
  if (fmt == 1) im2col_kernel<1><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(images, out, batch, nrm);
  else if (fmt == 2) im2col_kernel<2><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(images, out, batch, nrm);
  else im2col_kernel<0><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(images, out, batch, nrm);
  return rvk_launch_check();
}

int rvk_token_table_launch(const float* cls_token, const float* pos_embed, const float* patch_bias, float* table,
                           cudaStream_t stream) {
  token_table_kernel<<<(kTok * kD + 255) / 256, 256, 0, stream>>>(cls_token, pos_embed, patch_bias, table);
  return rvk_launch_check();
}

// Debugging / test aid: a kernel that waits on an mbarrier nobody arrives on, i.e. the protocol error every bounded wait of
// this library guards against (common.cuh::mbar_wait: ~4 s, device printf, trap).
namespace {
__global__ void debug_mbar_timeout_kernel() {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) mbar_wait(&bar, 0);
}
}  // namespace
int rvk_debug_mbar_timeout_launch(cudaStream_t stream) {
  debug_mbar_timeout_kernel<<<1, 32, 0, stream>>>();
  return rvk_launch_check();
}

int rvk_cast_multi_launch(RvkCastTable& T, cudaStream_t stream) {
  if (T.n <= 0) return RVK_OK;
  int tiles = 0;
  for (int i = 0; i < T.n; ++i) {
    const RvkCastJob& J = T.job[i];
    if (J.src == nullptr || J.dst == nullptr || J.rows <= 0 || J.cols <= 0 || J.mode < 0 || J.mode > 2) return RVK_ERR_BAD_ARG;
    T.job[i].tile_start = tiles;
    tiles += ((J.rows + 31) / 32) * ((J.cols + 31) / 32);
  }
  cast_multi_kernel<<<tiles, dim3(32, 8), 0, stream>>>(T);
  return rvk_launch_check();
}

int rvk_cast_bf16_launch(const float* src, void* dst, int64_t n, cudaStream_t stream) {
  if (n <= 0) return RVK_OK;
  const long long quads = (n + 3) / 4;
  cast_bf16_kernel<<<static_cast<unsigned>((quads + 255) / 256), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), n);
  return rvk_launch_check();
}

int rvk_cast_f16_launch(const float* src, void* dst, int64_t n, cudaStream_t stream) {
  if (n <= 0) return RVK_OK;
  cast_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(src, static_cast<__half*>(dst), n);
  return rvk_launch_check();
}

int rvk_cast_transpose_bf16_launch(const float* src, void* dst, int rows, int cols, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return RVK_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  cast_transpose_kernel<<<grid, block, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), rows, cols);
  return rvk_launch_check();
}

int rvk_layernorm_fwd_launch(const float* x, int64_t x_row_stride, const float* gamma, const float* beta, float eps,
                             void* y, int y_is_bf16, int64_t y_row_stride, float* mean, float* rstd, int rows,
                             cudaStream_t stream) {
  if (rows <= 0) return RVK_OK;
  const int blocks = min((rows + 7) / 8, kNumSMsB200 * 8);
  if (y_is_bf16)
    layernorm_fwd_kernel<true><<<blocks, 256, 0, stream>>>(x, x_row_stride, gamma, beta, eps, y, y_row_stride, mean, rstd, rows);
  else
    layernorm_fwd_kernel<false><<<blocks, 256, 0, stream>>>(x, x_row_stride, gamma, beta, eps, y, y_row_stride, mean, rstd, rows);
  return rvk_launch_check();
}

int rvk_layernorm_fwd_tiled_launch(const float* x_tiled, int64_t token_row_stride, const float* gamma, const float* beta,
                                   float eps, float* y, int64_t y_row_stride, int rows, cudaStream_t stream) {
  if (rows <= 0) return RVK_OK;
  const int blocks = min((rows + 7) / 8, kNumSMsB200 * 8);
  layernorm_fwd_kernel<false, true><<<blocks, 256, 0, stream>>>(x_tiled, token_row_stride, gamma, beta, eps, y,
                                                                y_row_stride, nullptr, nullptr, rows);
  return rvk_launch_check();
}

int rvk_layernorm_bwd_launch(const void* g, int g_is_bf16, int64_t g_row_stride, const float* x, int64_t x_row_stride,
                             const float* mean, const float* rstd, const float* gamma, const float* dx_in,
                             float* dx_out, int64_t dx_row_stride, void* dx_out_bf16, float* dgamma, float* dbeta,
                             float* dcolsum, int rows, cudaStream_t stream) {
  if (rows <= 0) return RVK_OK;
  if ((g_row_stride | x_row_stride | dx_row_stride) % 4 != 0 ||
      ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx_in) |
        reinterpret_cast<uintptr_t>(dx_out) | reinterpret_cast<uintptr_t>(dx_out_bf16) | reinterpret_cast<uintptr_t>(gamma)) & 15) != 0)
    return RVK_ERR_UNSUPPORTED_SHAPE;
  const int blocks = min((rows + 15) / 16, kNumSMsB200 * 6);
  auto* dxb = static_cast<__nv_bfloat16*>(dx_out_bf16);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = rvk_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const long long gs = g_row_stride, xs = x_row_stride, dxs = dx_row_stride;
  // algorithmic bytes per row of 192: g (bf16 or fp32) + x fp32 + dx_in fp32 (if any) read; dx fp32 + dx bf16 written
  RvkScopedTimer timer(stream, 0.0, double(rows) * 192.0 * ((g_is_bf16 ? 2.0 : 4.0) + 4.0 + (dx_in ? 4.0 : 0.0) + 4.0 + (dxb ? 2.0 : 0.0)),
                       RVK_T_LN_BWD);
  if (g_is_bf16)
    RVK_CUDA_TRY(cudaLaunchKernelEx(&cfg, layernorm_bwd_kernel<true>, g, gs, x, xs, mean, rstd, gamma, dx_in, dx_out, dxs, dxb, dgamma,
                                    dbeta, dcolsum, rows));
  else
    RVK_CUDA_TRY(cudaLaunchKernelEx(&cfg, layernorm_bwd_kernel<false>, g, gs, x, xs, mean, rstd, gamma, dx_in, dx_out, dxs, dxb, dgamma,
                                    dbeta, dcolsum, rows));
  return rvk_launch_check();
}

int rvk_colsum_launch(const void* src, int src_is_bf16, int64_t ld, int rows, int cols, float* out, float scale,
                      cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return RVK_OK;
  const int ve = src_is_bf16 ? 8 : 4;
  if (cols % ve != 0 || ld % ve != 0 || cols / ve > 256 || (reinterpret_cast<uintptr_t>(src) & 15) != 0)
    return RVK_ERR_UNSUPPORTED_SHAPE;
  int blocks = kNumSMsB200 * 4;
  int rpb = (rows + blocks - 1) / blocks;
  if (rpb < 16) rpb = 16;
  blocks = (rows + rpb - 1) / rpb;
  if (src_is_bf16) colsum_kernel<true><<<blocks, 256, 0, stream>>>(src, ld, rows, cols, out, scale, rpb);
  else colsum_kernel<false><<<blocks, 256, 0, stream>>>(src, ld, rows, cols, out, scale, rpb);
  return rvk_launch_check();
}

int rvk_token_grad_reduce_launch(const float* dx0, int batch, float* dpos, float* dcls, float* dpatch_bias,
                                 cudaStream_t stream) {
  if (batch <= 0) return RVK_OK;
  const int bpb = 32;
  dim3 grid(kTok, (batch + bpb - 1) / bpb);
  token_grad_reduce_kernel<<<grid, kD, 0, stream>>>(dx0, batch, bpb, dpos, dcls, dpatch_bias);
  return rvk_launch_check();
}
