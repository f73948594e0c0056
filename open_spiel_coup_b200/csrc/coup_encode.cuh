// coup_encode.cuh -- the tensor encoders (CoupObserver::WriteTensor, coup.cc:150-287): info-state rows composed in shared
// memory and written with bulk (TMA) stores, the plain vector-store variant, the 98-element observation rows, the sources
// they read (env slab, gather list, packed records / the finished-episode ring), and the row hash used for verification.
#pragma once
#include "coup_step.cuh"

namespace coup {

// ---- tensor element types ------------------------------------------------------------------------
template <typename T> struct Unit4;  // four consecutive tensor elements
template <> struct Unit4<float> {
  using type = float4;
  static __device__ __forceinline__ type make(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return make_float4(static_cast<float>(a), static_cast<float>(b), static_cast<float>(c), static_cast<float>(d));
  }
};
template <> struct Unit4<uint8_t> {
  using type = uint32_t;
  static __device__ __forceinline__ type make(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return a | (b << 8) | (c << 16) | (d << 24);
  }
};
template <> struct Unit4<__nv_bfloat16> {
  using type = uint2;
  // bf16 of a small non-negative integer = the top half of its fp32 encoding (exact for 0..255).
  static __device__ __forceinline__ uint32_t bits(uint32_t v) { return __float_as_uint(static_cast<float>(v)) >> 16; }
  static __device__ __forceinline__ type make(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return make_uint2(bits(a) | (bits(b) << 16), bits(c) | (bits(d) << 16));
  }
};

// ---- info-state encoder ----------------------------------------------------------------------------
// Shared-memory record of one env, filled by the lane that owns the env:
//   [0,16) history words  [16,18) head mask of view A  [18,20) head mask of view B
//   [20] len | coins0<<8 | coins1<<16 | viewA_observer<<24 | viewB_observer<<25
__device__ __forceinline__ void fill_record(uint32_t* rec, const Env& s, const uint32_t* hist_row,
                                            int player_sel) {
  if (hist_row != rec) {                      // the fused step kernels keep the row in the record all along
    const uint4* h4 = reinterpret_cast<const uint4*>(hist_row);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint4 v = h4[k];
      rec[4 * k + 0] = v.x; rec[4 * k + 1] = v.y; rec[4 * k + 2] = v.z; rec[4 * k + 3] = v.w;
    }
  }
  const bool term = is_terminal(s);
  const int who = player_sel & 7;                                   // COUP_PLAYER_*; bits 8.. = kVis* of the observer type
  const uint32_t vis = static_cast<uint32_t>(player_sel) >> 8;
  const uint32_t obs_a = who == COUP_PLAYER_1 ? 1u : who == COUP_PLAYER_CURRENT ? g_mover(s.g) : 0u;
  const uint32_t obs_b = 1u;
  const uint64_t ma = head_mask(s, obs_a, term, vis);
  rec[16] = static_cast<uint32_t>(ma); rec[17] = static_cast<uint32_t>(ma >> 32);
  if (who == COUP_PLAYER_BOTH) {
    const uint64_t mb = head_mask(s, obs_b, term, vis);
    rec[18] = static_cast<uint32_t>(mb); rec[19] = static_cast<uint32_t>(mb >> 32);
  }
  rec[20] = c_moves(s.c) | (pw_coins(s.p[0]) << 8) | (pw_coins(s.p[1]) << 16) | (obs_a << 24) | (obs_b << 25);
}

// Value of info-state element `p` (0..2491) of a record/view. Small non-negative integer.
__device__ __forceinline__ uint32_t info_value(const uint32_t* rec, uint64_t mask, uint32_t meta,
                                               uint32_t observer, uint32_t p) {
  if (p < 60u) return static_cast<uint32_t>(mask >> p) & 1u;
  if (p < 62u) return (meta >> (8u + 8u * (p - 60u))) & 255u;        // WriteCoins, 207-213
  const uint32_t i = (p - 62u) / 18u, a = (p - 62u) - 18u * i;       // WriteActionHistory, 230-245
  if (i >= (meta & 255u)) return 0u;
  const uint32_t w = i / 6u;
  const uint32_t code = (rec[w] >> (5u * (i - 6u * w))) & 31u;
  return history_column(code, observer) == a ? 1u : 0u;
}

// The warp writes `nrows` consecutive rows (row r of the warp -> record r>>both, view r&both) starting at
// out_row0. Rows are 623 units of four elements; within a row lane l handles units l, l+32, ...
template <typename T>
__device__ __forceinline__ void warp_encode_info(const uint32_t* recs, int nrec, bool both,
                                                 typename Unit4<T>::type* out_units, int lane, int row_units) {
  using U = typename Unit4<T>::type;
  const int nrows = both ? 2 * nrec : nrec;
  for (int r = 0; r < nrows; ++r) {
    const uint32_t* rec = recs + (both ? (r >> 1) : r) * kRecWords;
    const int view = both ? (r & 1) : 0;
    const uint32_t meta = rec[20];
    const uint32_t observer = (meta >> (24 + view)) & 1u;
    const uint64_t mask = static_cast<uint64_t>(rec[16 + 2 * view]) | (static_cast<uint64_t>(rec[17 + 2 * view]) << 32);
    const int len = static_cast<int>(meta & 255u);
    // units [0, nz_end) can hold non-zeros: the 62-float head plus `len` history rows of 18
    const int nz_end = min(kUnitsPerInfoRow, (62 + 18 * len + 3) >> 2);
    U* row = out_units + static_cast<size_t>(r) * row_units;  // row_units >= 623: padded row stride
    int q = lane;
    for (; q < nz_end; q += 32) {
      const uint32_t p0 = 4u * q;
      U v;
      if (p0 + 3u < 60u) {
        const uint32_t b = static_cast<uint32_t>(mask >> p0);
        v = Unit4<T>::make(b & 1u, (b >> 1) & 1u, (b >> 2) & 1u, (b >> 3) & 1u);
      } else {
        v = Unit4<T>::make(info_value(rec, mask, meta, observer, p0), info_value(rec, mask, meta, observer, p0 + 1u),
                           info_value(rec, mask, meta, observer, p0 + 2u), info_value(rec, mask, meta, observer, p0 + 3u));
      }
      row[q] = v;
    }
    const U zero = Unit4<T>::make(0, 0, 0, 0);
#pragma unroll 4
    for (; q < row_units; q += 32) row[q] = zero;
  }
}

// Where the encoders find the (state, history) pair behind output row group `e`:
//   SlabSource   -- the env slab itself, optionally through a gather list of env ids;
//   RecordSource -- an array (or ring) of packed observation records (COUP_RECORD_WORDS each: 16 history words, 4 state
//                   words, 4 meta words), optionally through an index list, optionally limited to "the episodes that
//                   finished in the last step call" = ring positions [ctrl[1], ctrl[0]). The row count is then only known
//                   on the device: the grid is sized for the caller's capacity and surplus blocks exit.
struct SlabSource {
  const uint4* state;
  const uint32_t* history;
  const uint32_t* ids;
  uint32_t n;
  const uint32_t* count_ptr;   // optional: only the first *count_ptr rows (a level of a traversal)
  __device__ __forceinline__ uint32_t rows() const { return count_ptr ? min(n, *count_ptr) : n; }
  __device__ __forceinline__ void locate(uint32_t e, const uint4*& sp, const uint32_t*& hp, uint32_t& id, int& sel) const {
    id = ids ? ids[e] : e;
    sp = state + id;
    hp = history + static_cast<size_t>(id) * kHistoryWords;
  }
};
struct RecordSource {
  const uint32_t* records;
  const uint32_t* indices;           // optional
  const unsigned long long* ctrl;    // optional ring control words
  uint32_t index_mask;               // ring capacity - 1, or 0xFFFFFFFF for a plain array
  uint32_t n;                        // rows wanted (an upper bound when ctrl is given)
  __device__ __forceinline__ uint32_t rows() const {
    if (ctrl == nullptr) return n;
    const unsigned long long avail = ctrl[0] - ctrl[1];
    return avail < n ? static_cast<uint32_t>(avail) : n;
  }
  __device__ __forceinline__ void locate(uint32_t e, const uint4*& sp, const uint32_t*& hp, uint32_t& id, int& sel) const {
    const uint32_t idx = indices ? indices[e] : (ctrl ? static_cast<uint32_t>(ctrl[1]) + e : e);
    const uint32_t* rec = records + static_cast<size_t>(idx & index_mask) * kRecordWords;
    hp = rec;
    sp = reinterpret_cast<const uint4*>(rec + kHistoryWords);
    id = rec[20];
    if ((sel & 7) == COUP_PLAYER_FROM_RECORD) sel = (sel & ~7) | static_cast<int>(rec[21] >> 31);   // the record's seat
  }
};

// Loads the pair behind row group `e`, leaves its encoder record in `rec`, reports the env id the row describes.
template <typename Src>
__device__ __forceinline__ void load_and_fill(const Src& src, uint32_t e, int player_sel, uint32_t* rec, uint32_t* ids_out) {
  const uint4* sp; const uint32_t* hp; uint32_t id; int sel = player_sel;
  src.locate(e, sp, hp, id, sel);
  const Env s = load_env(sp);
  fill_record(rec, s, hp, sel);
  if (ids_out != nullptr) ids_out[e] = id;
}

template <typename T, typename Src>
__global__ void __launch_bounds__(kBlockThreads)
k_encode_info(Src src, int player_sel, T* __restrict__ out, uint32_t stride, uint32_t* __restrict__ ids_out,
              uint32_t* __restrict__ count_out) {
  __shared__ uint32_t s_rec[kWarpsPerBlock][32 * kRecWords];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = src.rows();
  if (count_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *count_out = n;
  const uint32_t e0 = (blockIdx.x * kWarpsPerBlock + warp) * 32u;
  if (e0 >= n) return;
  const uint32_t e = e0 + lane;
  if (e < n) load_and_fill(src, e, player_sel, &s_rec[warp][lane * kRecWords], ids_out);
  __syncwarp();
  const int nrec = static_cast<int>(min(32u, n - e0));
  const bool both = (player_sel & 7) == COUP_PLAYER_BOTH;
  using U = typename Unit4<T>::type;
  const int row_units = static_cast<int>(stride / 4);
  U* out_units = reinterpret_cast<U*>(out) + static_cast<size_t>(e0) * (both ? 2 : 1) * row_units;
  warp_encode_info<T>(s_rec[warp], nrec, both, out_units, lane, row_units);
}

// ---- info-state encoder, staged variant: rows are composed in shared memory and written with bulk
// (TMA) stores. A dense row is >97 % zeros, so instead of computing and storing 623 units per row the warp
// keeps an all-zero 9 968-byte staging buffer in shared memory, pokes the ~30 non-zeros of a row into it,
// hands the buffer to the TMA engine (cp.async.bulk shared -> global, 1 instruction, SASS UBLKCP), waits
// for the engine to have READ the buffer, and un-pokes the same positions. 9 968 B = one f32 row = two
// bf16 rows = four u8 rows, always a multiple of 16 B and 16-B aligned in the output.
constexpr int kStageBytes = 2496 * 4;   // 9984: room for the padded row stride 2496 (2492 -> 9968 used)
constexpr int kTmaWarpsPerBlock = 8;
constexpr int kTmaBlockThreads = kTmaWarpsPerBlock * 32;
constexpr int kTmaSmemPerWarp = kStageBytes + 32 * kRecWords * 4;  // staging buffer + 32 records
constexpr int kTmaSmemBytes = kTmaWarpsPerBlock * kTmaSmemPerWarp + COUP_STATS_LEN * 4;

template <typename T> struct Elem;
template <> struct Elem<float> { static __device__ __forceinline__ float from(uint32_t v) { return static_cast<float>(v); } };
template <> struct Elem<uint8_t> { static __device__ __forceinline__ uint8_t from(uint32_t v) { return static_cast<uint8_t>(v); } };
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ __nv_bfloat16 from(uint32_t v) { return __float2bfloat16(static_cast<float>(v)); }
};

__device__ __forceinline__ void tma_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_store(void* gptr, const void* smem, uint32_t bytes) {
  const uint32_t saddr = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gptr), "r"(saddr), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// What one lane pokes into one row, decided ahead of time (one word): bits 0-1 elements `lane` and `32 + lane` of the
// head are 1; bits 2-9 the coin count this lane writes (lanes 28, 29 -> elements 60, 61, raw counts 207-213); bits 10-24
// the columns of history rows lane, lane + 32, lane + 64 (5 bits each, 31 = nothing to write: a deal to the other
// player, or past the end); bits 25-31 the number of moves. Planning reads the records and does all the arithmetic;
// poking is then nothing but shared-memory stores, and the PLAN of the next group is computed while the TMA engine
// reads the buffer of the current one.
__device__ __forceinline__ uint32_t plan_row(const uint32_t* rec, int view, int lane) {
  const uint32_t meta = rec[20];
  const uint32_t observer = (meta >> (24 + view)) & 1u;
  const uint32_t lo = rec[16 + 2 * view], hi = rec[17 + 2 * view];
  const uint32_t len = meta & 127u;
  uint32_t plan = ((lo >> lane) & 1u) | (((hi >> lane) & 1u) << 1) | (len << 25);
  if (lane >= 28 && lane < 30) plan |= ((meta >> (8u + 8u * (lane - 28))) & 255u) << 2;
#pragma unroll
  for (uint32_t j = 0; j < 3; ++j) {
    const uint32_t i = lane + 32u * j;
    uint32_t col = 31u;
    if (i < len) {
      const uint32_t w = i / 6u;
      col = history_column((rec[w] >> (5u * (i - 6u * w))) & 31u, observer);   // WriteActionHistory, 230-245
    }
    plan |= col << (10u + 5u * j);
  }
  return plan;
}

template <typename T>
__device__ __forceinline__ void poke_plan(T* row, uint32_t plan, int lane) {
  const T one = Elem<T>::from(1u);
  if (plan & 1u) row[lane] = one;                                          // elements 0..31
  if (plan & 2u) row[32 + lane] = one;                                     // elements 32..59 (the mask has 60 bits)
  if (lane >= 28 && lane < 30) row[32 + lane] = Elem<T>::from((plan >> 2) & 255u);
#pragma unroll
  for (uint32_t j = 0; j < 3; ++j) {
    const uint32_t col = (plan >> (10u + 5u * j)) & 31u;
    if (col != 31u) row[62u + 18u * (lane + 32u * j) + col] = one;
  }
}

// Erases a row again: every non-zero lives in the first 62 + 18*len elements, so instead of recomputing the poked
// positions the warp zero-fills that prefix, widened to 16-byte boundaries, with uint4 stores (one or two store
// instructions per row for any element type). The widening can only touch the zero tail of the previous row of the
// same staging buffer or later elements of this row, all of which are zero once the group has been erased.
template <typename T>
__device__ __forceinline__ void clear_row(T* row, uint32_t len, int lane) {
  const uint32_t saddr = static_cast<uint32_t>(__cvta_generic_to_shared(row));
  const uint32_t lo = saddr & ~15u;
  const uint32_t hi = (saddr + (62u + 18u * len) * static_cast<uint32_t>(sizeof(T)) + 15u) & ~15u;
  uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(row) - (saddr - lo));
  const int units = static_cast<int>((hi - lo) >> 4);
  for (int q = lane; q < units; q += 32) p[q] = make_uint4(0u, 0u, 0u, 0u);
}

// Full block (8 warps x 32 records): the block's rows form one contiguous span of the output, written in groups
// of G = 4/sizeof(T) rows per bulk store. Groups are dealt to the warps ROUND-ROBIN (warp w takes groups w, w+8,
// ...), so at any moment the eight warps of a block are writing eight ADJACENT groups: that keeps the DRAM write
// stream sequential over ~80 KB windows and is worth ~5 % of HBM write bandwidth over each warp streaming its own
// 32 rows (scripts/store_bw_probe.cu: 7.26 vs 6.93 TB/s for pure bulk stores of this shape).
// Per group: poke the planned non-zeros, hand the buffer to the TMA engine (cp.async.bulk shared -> global, SASS UBLKCP),
// plan the NEXT group while the engine reads, wait for the read, erase.
template <typename T>
__device__ __forceinline__ void block_encode_info_tma(const uint32_t* block_recs, bool both, T* stage,
                                                      unsigned char* out_block_bytes, int warp, int lane, int stride,
                                                      int nwarps = kTmaWarpsPerBlock, bool wait_for_writes = true) {
  constexpr int G = 4 / static_cast<int>(sizeof(T));
  const uint32_t group_bytes = static_cast<uint32_t>(G * stride) * sizeof(T);  // 9968 (stride 2492) or 9984 (2496)
  const int ngroups = (both ? 2 : 1) * kTmaWarpsPerBlock * 32 / G;
  uint32_t plan[G], next[G];
  auto plan_group = [&](int g, uint32_t (&out)[G]) {
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const int r = g * G + k;
      out[k] = plan_row(block_recs + (both ? (r >> 1) : r) * kRecWords, both ? (r & 1) : 0, lane);
    }
  };
  if (warp < ngroups) plan_group(warp, plan);
  for (int g = warp; g < ngroups; g += nwarps) {
#pragma unroll
    for (int k = 0; k < G; ++k) poke_plan<T>(stage + k * stride, plan[k], lane);
    tma_store_fence();   // generic-proxy writes -> visible to the async proxy
    __syncwarp();
    if (lane == 0) tma_bulk_store(out_block_bytes + static_cast<size_t>(g) * group_bytes, stage, group_bytes);
    if (g + nwarps < ngroups) plan_group(g + nwarps, next);   // overlaps the engine's read of the buffer
    if (lane == 0) tma_wait_read_all();  // the engine has read the buffer (the global write itself is still in flight)
    __syncwarp();
#pragma unroll
    for (int k = 0; k < G; ++k) {
      clear_row<T>(stage + k * stride, plan[k] >> 25, lane);
      plan[k] = next[k];
    }
  }
  // Before the CTA exits the engine must be done with this warp's shared memory (a persistent caller defers this).
  if (wait_for_writes && lane == 0) tma_wait_all();
}

// Carves the dynamic shared memory of a staged kernel: [8 stage buffers of 9984 B][8 x 32 records][stats].
struct TmaSmem {
  unsigned char* stage;   // this warp's staging buffer
  uint32_t* block_recs;   // records of the whole block, indexed by env-in-block
  uint32_t* recs;         // this warp's 32 records
  uint32_t* stats;
  __device__ __forceinline__ TmaSmem(unsigned char* base, int warp) {
    stage = base + static_cast<size_t>(warp) * kStageBytes;
    block_recs = reinterpret_cast<uint32_t*>(base + static_cast<size_t>(kTmaWarpsPerBlock) * kStageBytes);
    recs = block_recs + warp * 32 * kRecWords;
    stats = reinterpret_cast<uint32_t*>(base + static_cast<size_t>(kTmaWarpsPerBlock) * kTmaSmemPerWarp);
  }
};
__device__ __forceinline__ void zero_stage(unsigned char* stage, int lane) {
  uint4* p = reinterpret_cast<uint4*>(stage);
  for (int q = lane; q < kStageBytes / 16; q += 32) p[q] = make_uint4(0, 0, 0, 0);
}

template <typename T, typename Src>
__global__ void __launch_bounds__(kTmaBlockThreads, 2)
k_encode_info_tma(Src src, int player_sel, T* __restrict__ out, uint32_t stride, uint32_t* __restrict__ ids_out,
                  uint32_t* __restrict__ count_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TmaSmem sm(smem_raw, warp);
  const uint32_t n = src.rows();
  if (count_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *count_out = n;
  const uint32_t b0 = blockIdx.x * (kTmaWarpsPerBlock * 32u);
  if (b0 >= n) return;                                        // uniform over the block (device-side row counts)
  const uint32_t e0 = b0 + warp * 32u;
  const uint32_t e = e0 + lane;
  const bool block_full = b0 + kTmaWarpsPerBlock * 32u <= n;  // uniform over the block
  const bool both = (player_sel & 7) == COUP_PLAYER_BOTH;
  if (block_full) zero_stage(sm.stage, lane);
  if (e < n) load_and_fill(src, e, player_sel, sm.recs + lane * kRecWords, ids_out);
  if (block_full) {
    __syncthreads();
    block_encode_info_tma<T>(sm.block_recs, both, reinterpret_cast<T*>(sm.stage),
                             reinterpret_cast<unsigned char*>(out) + static_cast<size_t>(b0) * (both ? 2 : 1) * stride * sizeof(T),
                             warp, lane, static_cast<int>(stride));
  } else if (e0 < n) {  // ragged last block: per-warp plain vector stores
    __syncwarp();
    using U = typename Unit4<T>::type;
    const size_t row0 = static_cast<size_t>(e0) * (both ? 2 : 1);
    warp_encode_info<T>(sm.recs, static_cast<int>(min(32u, n - e0)), both, reinterpret_cast<U*>(out) + row0 * (stride / 4), lane,
                        static_cast<int>(stride / 4));
  }
}

// ---- observation encoder (98 elements per row) ------------------------------------------------------------------------
// One warp per 32 consecutive envs: although one row is not a multiple of 16 bytes, the 32 (x2 views) rows of a warp are one
// contiguous, 16-byte aligned span of the output (3 136 / 6 272 / 12 544 B per view for u8 / bf16 / f32). Each lane composes
// its own env's row(s) in the warp's staging span in shared memory (49 independent pair stores per row), lane 0 hands the
// span to the TMA engine with one bulk store, and the next group overwrites it once the engine has read it. Persistent:
// warps stride over the 32-env groups. A ragged last group, or an output that is not 16-byte aligned, is copied out of the
// staging span element by element.
constexpr int kObsWarps = 4;
constexpr int kObsThreads = kObsWarps * 32;

// Two consecutive tensor elements as one store unit: a row of 98 (or 42) elements is 49 (21) such pairs, and a row starts on
// a pair boundary for every element type (98 and 42 are even).
template <typename T> struct Pair;
template <> struct Pair<uint8_t> {
  using type = uint16_t;
  static __device__ __forceinline__ type make(uint32_t a, uint32_t b) { return static_cast<uint16_t>(a | (b << 8)); }
};
template <> struct Pair<__nv_bfloat16> {
  using type = uint32_t;
  static __device__ __forceinline__ type make(uint32_t a, uint32_t b) {
    return Unit4<__nv_bfloat16>::bits(a) | (Unit4<__nv_bfloat16>::bits(b) << 16);
  }
};
template <> struct Pair<float> {
  using type = float2;
  static __device__ __forceinline__ type make(uint32_t a, uint32_t b) { return make_float2(static_cast<float>(a), static_cast<float>(b)); }
};

// Writes one observation row into the staging span: every pair of the row is computed from the head mask, the coin counts
// and the last-action mask and stored -- 49 independent stores, no chain of find-first-set steps, and nothing to erase
// afterwards because the next group overwrites every pair.
template <typename T>
__device__ __forceinline__ void compose_obs_row(T* row, uint64_t head, uint64_t last_action, uint32_t coins, bool pub) {
  using P = typename Pair<T>::type;
  P* out = reinterpret_cast<P*>(row);
  const uint32_t h0 = static_cast<uint32_t>(head), h1 = static_cast<uint32_t>(head >> 32);
#pragma unroll
  for (int k = 0; k < 21; ++k) {                                           // elements 0..41: observer, card values
    const uint32_t b = k < 16 ? (h0 >> (2 * k)) & 3u : (h1 >> (2 * (k - 16))) & 3u;
    out[k] = Pair<T>::make(b & 1u, b >> 1);
  }
  if (!pub) return;                                                        // no public info: the row ends after element 41
#pragma unroll
  for (int k = 21; k < 30; ++k) {                                          // elements 42..59: cur_move_player, cards_state
    const uint32_t b = (h1 >> (2 * (k - 16))) & 3u;
    out[k] = Pair<T>::make(b & 1u, b >> 1);
  }
  out[30] = Pair<T>::make(coins & 255u, coins >> 8);                       // WriteCoins, 207-213
  const uint32_t l0 = static_cast<uint32_t>(last_action), l1 = static_cast<uint32_t>(last_action >> 32);
#pragma unroll
  for (int k = 0; k < 18; ++k) {                                           // WriteLastAction, 217-225: elements 62..97
    const uint32_t b = k < 16 ? (l0 >> (2 * k)) & 3u : (l1 >> (2 * (k - 16))) & 3u;
    out[31 + k] = Pair<T>::make(b & 1u, b >> 1);
  }
}

// The sparse alternative: pokes (set = true) or erases (set = false) only the ~14 non-zeros of a row. Used for fp32 rows:
// with both views their lane pitch of 196 words makes the 49 pair stores of compose_obs_row four-way bank conflicts (139.7
// against 151.5 us per 2^20 envs), and with one view composing was no faster and less steady (80 us, at times 96). For the
// narrower element types composing wins (u8: 45.3 -> 31.0 us).
template <typename T>
__device__ __forceinline__ void poke_obs_row(T* row, uint64_t head, uint64_t last_action, uint32_t coins, bool set, bool pub) {
  const T one = Elem<T>::from(set ? 1u : 0u);
  while (head) {                                                        // elements 0..59
    const int b = __ffsll(static_cast<long long>(head)) - 1;
    head &= head - 1;
    row[b] = one;
  }
  if (!pub) return;                                                     // no public info: the row ends after element 41
  row[60] = Elem<T>::from(set ? coins & 255u : 0u);                     // WriteCoins, 207-213
  row[61] = Elem<T>::from(set ? coins >> 8 : 0u);
  while (last_action) {                                                 // WriteLastAction, 217-225
    const int b = __ffsll(static_cast<long long>(last_action)) - 1;
    last_action &= last_action - 1;
    row[62 + b] = one;
  }
}

// player_sel: bits 0-2 COUP_PLAYER_*, bits 8.. kVis* (the observer type); row_len = 98, or 42 without public info.
// kSparse: fp32 rows (see poke_obs_row); a compile-time switch so that the other element types carry no trace of it.
template <typename T, typename Src, bool kSparse>
__global__ void __launch_bounds__(kObsThreads)
k_encode_obs(Src src, int player_sel, T* __restrict__ out, int use_bulk, uint32_t row_len, uint32_t* __restrict__ ids_out,
             uint32_t* __restrict__ count_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t vis = static_cast<uint32_t>(player_sel) >> 8;
  const bool pub = (vis & kVisNoPublic) == 0;
  const bool both = (player_sel & 7) == COUP_PLAYER_BOTH;
  const uint32_t views = both ? 2u : 1u;
  constexpr bool sparse = kSparse;
  const uint32_t n = src.rows();
  if (count_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *count_out = n;
  const uint32_t span_elems = 32u * views * row_len;
  T* stage = reinterpret_cast<T*>(smem_raw) + static_cast<size_t>(warp) * span_elems;
  for (uint32_t q = lane; q < span_elems * sizeof(T) / 16u; q += 32u) reinterpret_cast<uint4*>(stage)[q] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
  const uint32_t n_groups = (n + 31u) / 32u;
  // The state word of the NEXT group is requested before this group's bulk store is waited for, so the load's round trip
  // (microseconds next to a saturated store stream) overlaps the engine's read of the buffer.
  auto fetch = [&](uint32_t g, uint4& sv, uint32_t& id, int& sel) {
    const uint32_t e = g * 32u + lane;
    sel = player_sel;
    if (g < n_groups && e < n) {
      const uint4* sp; const uint32_t* hp;
      src.locate(e, sp, hp, id, sel);
      sv = *sp;
    }
  };
  uint4 sv_next = make_uint4(0u, 0u, 0u, 0u);
  uint32_t id_next = 0;
  int sel_next = player_sel;
  const uint32_t g_first = blockIdx.x * kObsWarps + warp, g_step = gridDim.x * kObsWarps;
  fetch(g_first, sv_next, id_next, sel_next);
  for (uint32_t g = g_first; g < n_groups; g += g_step) {
    const uint32_t e0 = g * 32u, e = e0 + lane;
    const uint32_t nrec = min(32u, n - e0);
    const uint4 sv = sv_next;
    const uint32_t id = id_next;
    const int sel = sel_next;
    uint64_t head_a = 0, head_b = 0, la = 0;
    uint32_t coins = 0;
    if (e < n) {
      Env s;
      s.p[0] = sv.x; s.p[1] = sv.y; s.g = sv.z; s.c = sv.w;
      const bool term = is_terminal(s);
      const int who = sel & 7;
      const uint32_t obs_a = who == COUP_PLAYER_1 ? 1u : who == COUP_PLAYER_CURRENT ? g_mover(s.g) : 0u;
      head_a = head_mask(s, obs_a, term, vis);
      if (both) head_b = head_mask(s, 1u, term, vis);
      la = last_action_mask(s);
      coins = pw_coins(s.p[0]) | (pw_coins(s.p[1]) << 8);
      if (ids_out != nullptr) ids_out[e] = id;
      T* row = stage + static_cast<size_t>(lane) * views * row_len;
      if (sparse) {
        poke_obs_row<T>(row, head_a, la, coins, true, pub);
        if (both) poke_obs_row<T>(row + row_len, head_b, la, coins, true, pub);
      } else {
        compose_obs_row<T>(row, head_a, la, coins, pub);
        if (both) compose_obs_row<T>(row + row_len, head_b, la, coins, pub);
      }
    }
    T* dst = out + static_cast<size_t>(e0) * views * row_len;
    const bool bulk = use_bulk && nrec == 32u;
    if (bulk) {
      tma_store_fence();
      __syncwarp();
      if (lane == 0) tma_bulk_store(dst, stage, span_elems * static_cast<uint32_t>(sizeof(T)));
    } else {
      __syncwarp();
      const uint32_t total = nrec * views * row_len;
      for (uint32_t i = lane; i < total; i += 32u) dst[i] = stage[i];
    }
    fetch(g + g_step, sv_next, id_next, sel_next);
    if (bulk && lane == 0) tma_wait_read_all();     // the engine has read the span: the next group may overwrite it
    __syncwarp();
    if (sparse && e < n) {                          // sparse rows are erased again
      T* row = stage + static_cast<size_t>(lane) * views * row_len;
      poke_obs_row<T>(row, head_a, la, coins, false, pub);
      if (both) poke_obs_row<T>(row + row_len, head_b, la, coins, false, pub);
    }
  }
  if (lane == 0) tma_wait_all();   // the engine must be done with this warp's shared memory before the CTA exits
}

// ---- verification: position-keyed 64-bit hash of every row of a dense tensor -------------------------
template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<uint8_t>(uint8_t v) { return static_cast<float>(v); }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(kBlockThreads)
k_row_hash(const T* __restrict__ t, uint32_t rows, uint32_t row_len, uint64_t* __restrict__ out) {
  const uint32_t row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* p = t + static_cast<size_t>(row) * row_len;
  uint64_t h = 0;
  for (uint32_t i = lane; i < row_len; i += 32) {
    const uint32_t bits = __float_as_uint(to_float<T>(p[i]));
    if (bits != 0) h += mix64((static_cast<uint64_t>(i) << 32) | bits);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
  if (lane == 0) out[row] = h;
}

}  // namespace coup
