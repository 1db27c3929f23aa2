// Host-side construction of TMA tensor maps (cuTensorMapEncodeTiled looked up through the
// runtime's driver entry point so the library does not link libcuda directly).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

enum RvkDType : int { RVK_BF16 = 0, RVK_F32 = 1 };

// 2-D row-major tensor [rows, cols] with leading dimension `ld` (elements); box = [box_rows, box_cols];
// the inner box extent must be 128 bytes (64 bf16 / 32 fp32): every map uses the 128-byte swizzle.
int rvk_make_tmap_2d(CUtensorMap* out, const void* base, int dtype, int64_t rows, int64_t cols, int64_t ld,
                     int box_rows, int box_cols);

// 3-D view [d2][d1][d0] of a row-major tensor (d0 innermost, strides in elements); box = [1][box_d1][box_d0].
// Used for per-image token tiles: rows beyond an image's 197 tokens are zero-filled on load / clipped on store.
int rvk_make_tmap_3d(CUtensorMap* out, const void* base, int dtype, int64_t d0, int64_t d1, int64_t d2,
                     int64_t stride1, int64_t stride2, int box_d0, int box_d1);
