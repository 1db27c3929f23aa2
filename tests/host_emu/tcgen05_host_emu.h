// Functional emulation of the Blackwell pieces the single-CTA tcgen05 kernels use (TEST INFRASTRUCTURE ONLY; include after
// cuda_host_emu.h): mbarriers with transaction counts, TMA 2-D tile loads with the 128-byte swizzle, tensor memory, and
// `tcgen05.mma kind::f16` on bf16 operands read through shared-memory matrix descriptors.  The wrappers have the names and
// signatures of csrc/common.cuh, so a kernel cut verbatim out of a .cuh file compiles against them; the descriptor BUILDERS
// (umma_smem_desc, umma_idesc_bf16, sw128_offset) are NOT re-implemented here -- the test takes them verbatim from common.cuh and
// this file only DECODES what they produce, following the PTX ISA layouts the header documents:
//   shared-memory descriptor  [0,14) start >> 4, [16,30) leading-dimension byte offset >> 4, [32,46) stride byte offset >> 4,
//                             [61,64) swizzle (2 = 128 B)
//   instruction descriptor    [15] A major (1 = MN), [16] B major, [17,23) N >> 3, [24,29) M >> 4, bf16 x bf16 -> f32
//   128-byte swizzle          the 16-byte chunk index (address bits 4-6) is XORed with address bits 7-9
//   K-major operand           element (r, k): (r / 8) * SBO + (r % 8) * 128 + k * 2 inside a 128-byte row (K = 64 per row)
//   MN-major operand          element (mn, k): (mn / 64) * LBO + (k / 8) * SBO + (k % 8) * 128 + (mn % 64) * 2
// TMA loads and MMAs execute at issue time, tcgen05.commit arrives one scheduler pass later: one legal schedule of the
// asynchronous hardware; waits are fiber yields.
#pragma once
#include <map>

struct CUtensorMap {            // what rvk_make_tmap_2d encodes: a 2-D row-major tensor, box of `box_rows` x 128 bytes, SWIZZLE_128B
  const void* base; int elem_bytes; long long cols, rows, ld; int box_cols, box_rows;
};
static inline CUtensorMap emu_make_tmap_2d(const void* base, int elem_bytes, long long rows, long long cols, long long ld, int box_rows, int box_cols) {
  return CUtensorMap{base, elem_bytes, cols, rows, ld, box_cols, box_rows};
}

// ---- shared-memory addresses: offsets into the block's dynamic shared memory
static inline uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(static_cast<const unsigned char*>(p) - emu_dyn_smem.data()); }
static inline unsigned char* emu_smem_ptr(uint32_t addr) { return emu_dyn_smem.data() + addr; }
static inline uint32_t emu_sw128(uint32_t addr) { return addr ^ (((addr >> 7) & 7u) << 4); }
static inline uint32_t lane_id() { return emu_tid() & 31u; }
static inline bool elect_one() { return lane_id() == 0; }

// ---- mbarriers (state kept beside the 8-byte object the kernel reserves)
struct EmuMbar { int count = 0, pending = 0; long long tx = 0; unsigned phase = 0; };
static std::map<const void*, EmuMbar> emu_mbars;
static inline void emu_mbar_check(EmuMbar& b) {
  ++emu_progress;
  if (b.pending <= 0 && b.tx <= 0) { b.phase ^= 1u; b.pending = b.count; b.tx = 0; }
}
static inline void mbar_init(uint64_t* bar, uint32_t count) { EmuMbar b; b.count = b.pending = static_cast<int>(count); emu_mbars[bar] = b; }
static inline void fence_mbar_init() {}
static inline void mbar_arrive(uint64_t* bar) { EmuMbar& b = emu_mbars.at(bar); --b.pending; emu_mbar_check(b); }
static inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { EmuMbar& b = emu_mbars.at(bar); b.tx += bytes; --b.pending; emu_mbar_check(b); }
static inline void emu_complete_tx(uint64_t* bar, uint32_t bytes) { EmuMbar& b = emu_mbars.at(bar); b.tx -= bytes; emu_mbar_check(b); }
static inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return emu_mbars.at(bar).phase != (parity & 1u); }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) { emu_wait_on[emu_tid() & 1023u] = bar; while (!mbar_try_wait(bar, parity)) emu_yield(); ++emu_progress; }
static inline void fence_proxy_async_smem() {}
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}

// ---- named barriers (bar.sync id, nthreads)
static EmuBarrier emu_named_bar[16];
static int emu_named_n[16];
static inline void named_bar_sync(uint32_t id, uint32_t nthreads) {
  if (emu_named_n[id] != static_cast<int>(nthreads)) { emu_named_bar[id].reset(static_cast<int>(nthreads)); emu_named_n[id] = static_cast<int>(nthreads); }
  emu_named_bar[id].arrive_and_wait();
}

// ---- TMA: box rows of 128 bytes, written through the swizzle; out-of-range elements are zero-filled, the whole box counts as bytes
static inline void tma_prefetch_desc(const CUtensorMap*) {}
static inline void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const uint32_t dst = smem_u32(smem_dst);
  for (int r = 0; r < m->box_rows; ++r)
    for (int e = 0; e < m->box_cols; ++e) {
      const long long gr = static_cast<long long>(c1) + r, gc = static_cast<long long>(c0) + e;
      unsigned char* d = emu_smem_ptr(emu_sw128(dst + static_cast<uint32_t>(r) * 128u + static_cast<uint32_t>(e * m->elem_bytes)));
      if (gr >= 0 && gr < m->rows && gc >= 0 && gc < m->cols)
        std::memcpy(d, static_cast<const unsigned char*>(m->base) + (gr * m->ld + gc) * m->elem_bytes, m->elem_bytes);
      else
        std::memset(d, 0, m->elem_bytes);
    }
  emu_complete_tx(bar, static_cast<uint32_t>(m->box_rows) * 128u);
}

// ---- tensor memory: 128 lanes x 512 columns of 32 bits; address = lane << 16 | column
static uint32_t emu_tmem[128][512];
static inline void tmem_alloc(uint32_t* smem_result, uint32_t) { if (lane_id() == 0) *smem_result = 0; std::memset(emu_tmem, 0xCD, sizeof(emu_tmem)); }
static inline void tmem_relinquish() {}
static inline void tmem_dealloc(uint32_t, uint32_t) {}
static inline void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  const uint32_t lane = (taddr >> 16) + lane_id(), col = taddr & 0xffffu;
  for (int i = 0; i < 32; ++i) std::memcpy(&v[i], &emu_tmem[lane][col + i], 4);
}

// ---- tcgen05.mma kind::f16, cta_group::1, bf16 operands from shared-memory descriptors, fp32 accumulator in tensor memory
static inline float emu_operand(uint64_t desc, int major_mn, int idx, int k) {      // idx = row (M or N index), k = 0..15
  const uint32_t start = static_cast<uint32_t>(desc & 0x3FFFu) << 4;
  const uint32_t lbo = static_cast<uint32_t>((desc >> 16) & 0x3FFFu) << 4, sbo = static_cast<uint32_t>((desc >> 32) & 0x3FFFu) << 4;
  uint32_t off;
  if (major_mn) off = start + static_cast<uint32_t>(idx / 64) * lbo + static_cast<uint32_t>(k / 8) * sbo + static_cast<uint32_t>(k % 8) * 128u + static_cast<uint32_t>(idx % 64) * 2u;
  else off = start + static_cast<uint32_t>(idx / 8) * sbo + static_cast<uint32_t>(idx % 8) * 128u + static_cast<uint32_t>(k) * 2u;
  if ((desc >> 61) == 2) off = emu_sw128(off);
  uint16_t h;
  std::memcpy(&h, emu_smem_ptr(off), 2);
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}
static inline void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const int M = static_cast<int>((idesc >> 24) & 31u) << 4, N = static_cast<int>((idesc >> 17) & 63u) << 3;
  const int a_mn = (idesc >> 15) & 1u, b_mn = (idesc >> 16) & 1u;
  const uint32_t lane0 = d_tmem >> 16, col0 = d_tmem & 0xffffu;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.0f;
      if (accumulate) std::memcpy(&acc, &emu_tmem[lane0 + m][col0 + n], 4);
      for (int k = 0; k < 16; ++k) acc = std::fmaf(emu_operand(a_desc, a_mn, m, k), emu_operand(b_desc, b_mn, n, k), acc);
      std::memcpy(&emu_tmem[lane0 + m][col0 + n], &acc, 4);
    }
  ++emu_progress;
}
// tcgen05.commit arrives when the MMAs issued so far have FINISHED, i.e. some time after the instruction: applied at the start of
// the next scheduler pass.  (Applied at once, the elected lane could release a ring stage, the producer refill it and the
// barrier's phase wrap before the other lanes of the issuer warp -- which the kernels keep within one iteration by __syncwarp --
// had looked at it; on the hardware a TMA refill cannot overtake lanes that are a few instructions behind.)
static inline void umma_commit(uint64_t* bar) { emu_deferred.push_back([bar] { mbar_arrive(bar); }); }

// ---- the rest of the single-CTA surface (gemm_nt.cuh): split descriptors, TMA stores, tensor-memory stores
// common.cuh passes the K-major SW128 descriptor as (low word, constant high word); the test takes umma_desc_lo / kUmmaDescHiSw128
// verbatim and this only glues the halves together
template <int G>
static inline void umma_f16_split(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate);
static inline void emu_umma_split(uint32_t hi, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  umma_bf16(d_tmem, (static_cast<uint64_t>(hi) << 32) | a_lo, (static_cast<uint64_t>(hi) << 32) | b_lo, idesc, accumulate);
}
static inline void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  const uint32_t src = smem_u32(smem_src);
  for (int r = 0; r < m->box_rows; ++r)
    for (int e = 0; e < m->box_cols; ++e) {
      const long long gr = static_cast<long long>(c1) + r, gc = static_cast<long long>(c0) + e;
      if (gr < 0 || gr >= m->rows || gc < 0 || gc >= m->cols) continue;                    // clipped by the tensor map
      std::memcpy(const_cast<unsigned char*>(static_cast<const unsigned char*>(m->base)) + (gr * m->ld + gc) * m->elem_bytes,
                  emu_smem_ptr(emu_sw128(src + static_cast<uint32_t>(r) * 128u + static_cast<uint32_t>(e * m->elem_bytes))), m->elem_bytes);
    }
  ++emu_progress;
}
static inline void tma_store_commit() {}
template <int N> static inline void tma_store_wait_read() {}
template <int N = 0> static inline void tma_store_wait_all() {}
static inline void tma_prefetch_2d(const CUtensorMap*, int, int) {}
static inline void prefetch_l2(const void*) {}
static inline void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t lane = (taddr >> 16) + lane_id(), col = taddr & 0xffffu;
  for (int i = 0; i < 32; ++i) std::memcpy(&emu_tmem[lane][col + i], &v[i], 4);
}
static inline float ex2_approx(float x) { return exp2f(x); }          // ex2.approx.ftz.f32 (2 ulp) modelled by the exact function
