// coup_incremental.cuh -- the incremental info-state contract: a persistent buffer with both players' rows of every env,
// kept current by rewriting only what a step changed.
#pragma once
#include "coup_encode.cuh"

namespace coup {

// ---- incremental info-state contract ------------------------------------------------------------------------------
// The caller keeps a PERSISTENT buffer T[n][2][stride] holding both players' info-state rows of every env (filled
// once with the dense encoder). Each fused step then rewrites only what changed: the 62-element head of both views,
// the history rows of the moves made in this step (1 player move + <= 3 deals, or the 4 deals of a re-dealt
// episode) and, when an episode was re-dealt in place, zeros over the rows the finished episode had used.
// ~0.9 KB of stores per env-step instead of 2 x 9 968 B; the buffer always equals what the dense encoder would write.
//
// Every store is a 16-byte unit and every run of units starts and ends on a 32-BYTE SECTOR boundary of the buffer. A
// store that covers part of a sector makes the L2 fetch the rest from DRAM before it can merge, and the store path backs up
// behind those fills: with 8/16-byte stores at their natural offsets ncu shows 0.25 GB of DRAM reads per step for a kernel
// that reads 0.08 GB, 9.6 long-scoreboard stall cycles per issued instruction and 30 % issue activity (0.52 ms per step,
// 2^20 envs, f32). Rows of the reference layout are 9 968 B apart, so every second row even starts in the middle of a
// sector. So a span that changed -- [0, 62) and [62 + 18 first, 62 + 18 len), or one span from 0 to the end of the finished
// episode after a re-deal -- is widened to whole sectors, and what the widening touches is recomputed, not read: the tail
// of the previous row (always zero: history rows >= 91 are never used), rows 0..3 next to the head, the two rows before
// the first new one, zeros past the last move.
//
// Mapping. The owner lane of an env steps it (history row in shared memory, as in the other fused kernels) and leaves, per
// view, the two spans as BITMAPS of their 0/1 content (192 bits each, bit t = element span_start + t) plus where they start
// and how many units they have. Then the warp walks its touched envs; for each, the two half-warps take the two views and
// every lane turns 4 / 8 / 16 bits of a bitmap into one 16-byte unit (the unit holding the raw coin counts is patched).
// Earlier mappings, measured on one B200 at 2^20 envs, fp32: natural-offset 8/16-byte stores 0.52 ms; whole-sector stores
// with the content recomputed per element inside the walk 0.75-1.06 ms (3.5x the instructions, no fills any more); whole
// warp per (env, view) with one store per row 0.556 ms; every thread storing its own env's units 0.950 ms.
constexpr int kIncViewWords = 18;                      // per view: 2 span headers, 6 + 6 bitmap words, the 4 words of the coin unit
constexpr int kIncRecWords = 1 + 2 * kIncViewWords;    // 37: an odd pitch, conflict-free per-lane access; 4 CTAs fit an SM
constexpr int kIncRowPitch = kHistoryWords + 1;
constexpr int kIncSmemWords = kBlockThreads * (kIncRecWords + kIncRowPitch);       // records + rows; the unit table follows
constexpr int kIncSmemBytesMax = kIncSmemWords * 4 + 4096;                          // dynamic (above the 48 KB static limit)

// Sixteen bytes of consecutive tensor elements: from 0/1 bits, or from small-integer values.
template <typename T> struct Pack16;
template <> struct Pack16<float> {
  static constexpr int kElems = 4;
  static __device__ __forceinline__ uint4 from_bits(uint32_t b) {
    return make_uint4((b & 1u) * 0x3F800000u, ((b >> 1) & 1u) * 0x3F800000u, ((b >> 2) & 1u) * 0x3F800000u, ((b >> 3) & 1u) * 0x3F800000u);
  }
  template <typename F> static __device__ __forceinline__ uint4 make(F value) {
    return make_uint4(__float_as_uint(static_cast<float>(value(0))), __float_as_uint(static_cast<float>(value(1))),
                      __float_as_uint(static_cast<float>(value(2))), __float_as_uint(static_cast<float>(value(3))));
  }
};
template <> struct Pack16<__nv_bfloat16> {
  static constexpr int kElems = 8;
  static __device__ __forceinline__ uint4 from_bits(uint32_t b) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = ((b >> (2 * k)) & 1u) * 0x3F80u + ((b >> (2 * k + 1)) & 1u) * 0x3F800000u;
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
  template <typename F> static __device__ __forceinline__ uint4 make(F value) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = Unit4<__nv_bfloat16>::bits(value(2 * k)) | (Unit4<__nv_bfloat16>::bits(value(2 * k + 1)) << 16);
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct Pack16<uint8_t> {
  static constexpr int kElems = 16;
  static __device__ __forceinline__ uint4 from_bits(uint32_t b) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = (((b >> (4 * k)) & 15u) * 0x00204081u) & 0x01010101u;   // four bits -> four bytes
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
  template <typename F> static __device__ __forceinline__ uint4 make(F value) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = value(4 * k) | (value(4 * k + 1) << 8) | (value(4 * k + 2) << 16) | (value(4 * k + 3) << 24);
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// bits -> 16-byte unit through a small table in shared memory (one or two loads instead of a dozen ALU instructions):
// 16 entries for f32 (4 elements per unit), 256 for bf16 (8), 256 eight-byte entries looked up twice for u8 (16).
template <typename T> struct UnitLut;
template <> struct UnitLut<float> {
  static constexpr int kBytes = 16 * 16;
  static __device__ __forceinline__ void init(void* lut, int tid) {
    if (tid < 16) reinterpret_cast<uint4*>(lut)[tid] = Pack16<float>::from_bits(tid);
  }
  static __device__ __forceinline__ uint4 lookup(const void* lut, uint32_t bits) { return reinterpret_cast<const uint4*>(lut)[bits]; }
};
template <> struct UnitLut<__nv_bfloat16> {
  static constexpr int kBytes = 256 * 16;
  static __device__ __forceinline__ void init(void* lut, int tid) {
    if (tid < 256) reinterpret_cast<uint4*>(lut)[tid] = Pack16<__nv_bfloat16>::from_bits(tid);
  }
  static __device__ __forceinline__ uint4 lookup(const void* lut, uint32_t bits) { return reinterpret_cast<const uint4*>(lut)[bits]; }
};
template <> struct UnitLut<uint8_t> {
  static constexpr int kBytes = 256 * 8;
  static __device__ __forceinline__ void init(void* lut, int tid) {
    if (tid < 256) {
      const uint4 v = Pack16<uint8_t>::from_bits(tid);     // the low 8 bits fill x, y
      reinterpret_cast<uint2*>(lut)[tid] = make_uint2(v.x, v.y);
    }
  }
  static __device__ __forceinline__ uint4 lookup(const void* lut, uint32_t bits) {
    const uint2 a = reinterpret_cast<const uint2*>(lut)[bits & 255u], b = reinterpret_cast<const uint2*>(lut)[bits >> 8];
    return make_uint4(a.x, a.y, b.x, b.y);
  }
};

// Sets bit t of a 192-bit bitmap kept as six words in shared memory; t outside [0, 192) is ignored.
__device__ __forceinline__ void bitmap_set(uint32_t* words, int t) {
  if (t >= 0 && t < 192) words[t >> 5] |= 1u << (t & 31);
}

#ifndef COUP_INC_BLOCKS
#define COUP_INC_BLOCKS 4   // resident CTAs per SM the kernel is compiled for (registers <= 64; shared memory allows 3-4)
#endif
template <typename T>
__global__ void __launch_bounds__(kBlockThreads, COUP_INC_BLOCKS)
k_rollout_incremental(EnvArrays A, uint64_t step, T* __restrict__ buf, uint32_t stride) {
  __shared__ uint32_t s_stats[COUP_STATS_LEN];
  extern __shared__ __align__(16) uint32_t s_dyn[];      // kIncSmemBytes: [256][kIncRecWords] records, [256][kIncRowPitch] rows
  uint32_t (*s_rec)[32][kIncRecWords] = reinterpret_cast<uint32_t (*)[32][kIncRecWords]>(s_dyn);
  uint32_t (*s_row)[32][kIncRowPitch] = reinterpret_cast<uint32_t (*)[32][kIncRowPitch]>(s_dyn + kBlockThreads * kIncRecWords);
  void* const lut = s_dyn + kIncSmemWords;               // 16-byte aligned: kIncSmemWords is a multiple of 4
  UnitLut<T>::init(lut, threadIdx.x);                    // made visible by the __syncthreads of st.init below
  constexpr uint32_t kEl = Pack16<T>::kElems;            // elements per 16-byte unit
  constexpr uint32_t kSector = 2u * kEl;                  // elements per 32-byte sector
  BlockStats st;
  st.init(s_stats);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = e < A.n;
  Env s = {};
  uint32_t* hist_row = A.history + static_cast<size_t>(e) * kHistoryWords;
  uint32_t* row_copy = s_row[warp][lane];
  if (active) s = load_env_and_row(A, e, row_copy);
  const uint32_t old_len = c_moves(s.c);
  const StepResult r = step_env<true>(s, HistRow{row_copy, hist_row}, 0, nullptr, A, e, step, active);
  if (active) {
    store_env(A.state + e, s);
    write_outputs(A, e, r);
    if (r.stepped) {
      const uint32_t new_len = c_moves(s.c);
      const bool redealt = r.finished && new_len < r.final_moves + 1 && (A.flags & COUP_FLAG_AUTO_RESET);
      const uint32_t first = redealt ? 0u : old_len;   // rows [first, new_len) are (re)written, at most 4
      const bool term = is_terminal(s);
      auto code_at = [&](uint32_t i) {                  // 31 = no such row
        const uint32_t w = min(i, 95u) / 6u;
        return i < new_len ? (row_copy[w] >> (5u * (i - 6u * w))) & 31u : 31u;
      };
      uint32_t* rec = s_rec[warp][lane];
      const uint32_t coin0 = pw_coins(s.p[0]), coin1 = pw_coins(s.p[1]);
      // span 0: the head -- or, after a re-deal, everything from element 0 to the end of the finished episode's rows;
      // span 1: the new rows (none after a re-deal: they are part of span 0)
      const uint32_t hi0 = redealt ? 62u + 18u * max(r.final_moves, new_len) : 62u;
      const uint32_t lo1 = 62u + 18u * first, hi1 = redealt ? lo1 : lo1 + 18u * (new_len - first);
#pragma unroll
      for (uint32_t view = 0; view < 2; ++view) {
        uint32_t* vr = rec + 1 + view * kIncViewWords;
        const size_t row_base = (static_cast<size_t>(e) * 2 + view) * stride;         // absolute index of element 0 of the row
        const size_t a0 = row_base / kSector * kSector, b0 = (row_base + hi0 + kSector - 1u) / kSector * kSector;
        const int back0 = static_cast<int>(row_base - a0);                             // span 0 starts `back0` elements early
        vr[0] = static_cast<uint32_t>(back0) | (static_cast<uint32_t>((b0 - a0) / kEl) << 8);
        const size_t a1 = (row_base + lo1) / kSector * kSector, b1 = (row_base + hi1 + kSector - 1u) / kSector * kSector;
        const int start1 = static_cast<int>(a1 - row_base);                            // row element where span 1 starts
        vr[1] = static_cast<uint32_t>(start1) | ((hi1 > lo1 ? static_cast<uint32_t>((b1 - a1) / kEl) : 0u) << 16);
        // bitmaps: bit t = element (span start + t)
        uint32_t* bm0 = vr + 2;
        uint32_t* bm1 = vr + 8;
        const uint64_t head = head_mask(s, view, term);
        const unsigned long long lo = head << back0, hi = back0 ? head >> (64 - back0) : 0ull;   // back0 <= 31
        bm0[0] = static_cast<uint32_t>(lo); bm0[1] = static_cast<uint32_t>(lo >> 32); bm0[2] = static_cast<uint32_t>(hi);
        bm0[3] = bm0[4] = bm0[5] = 0u;
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {                                             // rows 0..3, next to the head
          const uint32_t col = history_column(code_at(k), view);
          if (col != 31u) bitmap_set(bm0, 62 + 18 * static_cast<int>(k) + static_cast<int>(col) + back0);
        }
        // the one unit of span 0 that is not 0/1: it holds the raw coin counts (elements 60, 61; 207-213). Rows start on
        // multiples of four elements, so both counts are always in the same unit.
        const uint32_t uc = (60u + static_cast<uint32_t>(back0)) / kEl;
        const int pc = static_cast<int>(uc * kEl) - back0;                             // row element of the unit's first element
        const uint32_t cbits = (bm0[(uc * kEl) >> 5] >> ((uc * kEl) & 31u)) & ((1u << kEl) - 1u);
        const uint4 cu = Pack16<T>::make([&](int k) { return pc + k == 60 ? coin0 : pc + k == 61 ? coin1 : (cbits >> k) & 1u; });
        vr[14] = cu.x; vr[15] = cu.y; vr[16] = cu.z; vr[17] = cu.w;
#pragma unroll
        for (int k = 0; k < 6; ++k) bm1[k] = 0u;
#pragma unroll
        for (uint32_t k = 0; k < 6; ++k) {                                             // rows first-2 .. first+3, around the new rows
          const uint32_t i = first + k;
          const uint32_t col = i >= 2u ? history_column(code_at(i - 2u), view) : 31u;
          if (col != 31u) bitmap_set(bm1, 62 + 18 * (static_cast<int>(i) - 2) + static_cast<int>(col) - start1);
        }
      }
    }
  }
  account(st, r, active);
  uint32_t touched = __ballot_sync(0xffffffffu, active && r.stepped);
  __syncwarp();
  const uint32_t view = static_cast<uint32_t>(lane) >> 4, l = static_cast<uint32_t>(lane) & 15u;
  const uint32_t e0 = e - lane;
  while (touched) {
    const int j = __ffs(touched) - 1;
    touched &= touched - 1;
    const uint32_t* rec = s_rec[warp][j];
    const uint32_t* vr = rec + 1 + view * kIncViewWords;
    const uint32_t hdr0 = vr[0], hdr1 = vr[1], coin_unit = (60u + (hdr0 & 255u)) / kEl;
    const size_t row_base = (static_cast<size_t>(e0 + j) * 2 + view) * stride;
    // both spans as one list of units: [0, n0) span 0, [n0, n0 + n1) span 1
    const uint32_t n0 = hdr0 >> 8, n1 = hdr1 >> 16;
    uint4* const dst0 = reinterpret_cast<uint4*>(buf + (row_base - (hdr0 & 255u)));
    uint4* const dst1 = reinterpret_cast<uint4*>(buf + (row_base + (hdr1 & 0xFFFFu)));
    for (uint32_t k = l; k < n0 + n1; k += 16u) {
      const bool second = k >= n0;
      const uint32_t u = second ? k - n0 : k;
      const uint32_t o = u * kEl;                                                    // never straddles a 32-bit word
      const uint32_t word = o < 192u ? vr[(second ? 8u : 2u) + (o >> 5)] : 0u;
      const uint32_t bits = (word >> (o & 31u)) & ((1u << kEl) - 1u);
      uint4 v = UnitLut<T>::lookup(lut, bits);
      if (!second && u == coin_unit) v = make_uint4(vr[14], vr[15], vr[16], vr[17]);   // the unit with the raw coin counts
      (second ? dst1 : dst0)[u] = v;
    }
  }
  st.flush(A.stats);
}

}  // namespace coup
