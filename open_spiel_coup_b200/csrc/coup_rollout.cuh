// coup_rollout.cuh -- the fused random-rollout step (sample -> step -> chance -> auto-reset -> ring -> outputs -> encode):
// the persistent warp-specialised kernel bench.py times (k_rollout_ws), the CTA-per-256-envs and plain-store variants, and
// the env-only multi-step kernel.
#pragma once
#include "coup_encode.cuh"

namespace coup {

// ---- fused random rollout step: sample -> step -> chance -> [auto-reset] -> outputs -> encode -------
template <typename T, bool kEncode>
__global__ void __launch_bounds__(kBlockThreads, 4)
k_rollout(EnvArrays A, uint64_t step, int player_sel, T* __restrict__ out, uint32_t stride) {
  __shared__ uint32_t s_stats[COUP_STATS_LEN];
  __shared__ uint32_t s_rec[kEncode ? kWarpsPerBlock : 1][kEncode ? 32 * kRecWords : 1];
  BlockStats st;
  st.init(s_stats);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t e0 = (blockIdx.x * kWarpsPerBlock + warp) * 32u;
  const uint32_t e = e0 + lane;
  const bool active = e < A.n;
  Env s = {};
  uint32_t* hist_row = A.history + static_cast<size_t>(e) * kHistoryWords;
  uint32_t* rec = kEncode ? &s_rec[warp][lane * kRecWords] : hist_row;
  if (active) s = kEncode ? load_env_and_row(A, e, rec) : load_env(A.state + e);
  const StepResult r = step_env<true>(s, kEncode ? HistRow{rec, hist_row} : global_row(hist_row), 0, nullptr, A, e, step, active);
  if (active) {
    store_env(A.state + e, s);
    write_outputs(A, e, r);
    if (kEncode) fill_record(rec, s, rec, player_sel);
  }
  account(st, r, active);
  if (kEncode && e0 < A.n) {
    __syncwarp();
    const int nrec = static_cast<int>(min(32u, A.n - e0));
    const bool both = player_sel == COUP_PLAYER_BOTH;
    using U = typename Unit4<T>::type;
    const int row_units = static_cast<int>(stride / 4);
    U* out_units = reinterpret_cast<U*>(out) + static_cast<size_t>(e0) * (both ? 2 : 1) * row_units;
    warp_encode_info<T>(s_rec[warp], nrec, both, out_units, lane, row_units);
  }
  st.flush(A.stats);
}

// Env-only rollout of `n_steps` steps in ONE launch: envs are independent, so a thread keeps its env in registers and its
// history row in shared memory across steps (one load and one store of each per launch, one set of outputs, one launch
// instead of n_steps of each); the statistics are updated every step exactly as k_rollout<T, false> does, and step k uses
// Philox counter step + k. With the row in shared memory no step waits for a global load: the merge of new move codes
// into the current word and the copy of a finished episode's log to the ring read it there.
constexpr int kEnvRowPitch = 20;   // words: rows stay 16-byte aligned (vector access), quarter-warps conflict-free
__global__ void __launch_bounds__(kBlockThreads, COUP_ENV_BLOCKS)
k_rollout_env_multi(EnvArrays A, uint64_t step, int n_steps) {
  __shared__ uint32_t s_stats[COUP_STATS_LEN];
  __shared__ __align__(16) uint32_t s_row[kBlockThreads][kEnvRowPitch];
  BlockStats st;
  st.init(s_stats);
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = e < A.n;
  Env s = {};
  uint4* const hist_row = reinterpret_cast<uint4*>(A.history + static_cast<size_t>(active ? e : 0) * kHistoryWords);
  uint4* const row = reinterpret_cast<uint4*>(s_row[threadIdx.x]);
  if (active) {
    const uint4 sv = A.state[e];
    const uint4 h0 = hist_row[0], h1 = hist_row[1], h2 = hist_row[2], h3 = hist_row[3];
    row[0] = h0; row[1] = h1; row[2] = h2; row[3] = h3;
    s.p[0] = sv.x; s.p[1] = sv.y; s.g = sv.z; s.c = sv.w;
  }
  StepResult r = {};
  // the mask of the loaded state; from then on every step hands the next one its mask
  r.legal = is_terminal(s) ? 0u : (g_chance(s.g) ? legal_mask_chance(s) : legal_mask_decision(s));
  StatAcc acc;
  acc.clear();
  for (int k = 0; k < n_steps; ++k) {                       // n_steps <= StatAcc::kMaxAdds (the host launches in chunks of 64)
    r = step_env<true, true>(s, global_row(s_row[threadIdx.x]), 0, nullptr, A, e, step + static_cast<uint64_t>(k), active, r.legal);
    acc.add(r, active);
  }
  acc.flush(st);
  if (active) {
    hist_row[0] = row[0]; hist_row[1] = row[1]; hist_row[2] = row[2]; hist_row[3] = row[3];
    store_env(A.state + e, s);
    write_outputs(A, e, r);
  }
  st.flush(A.stats);
}

// Same fused step, with the staged (shared memory + bulk store) encoder.
template <typename T>
__global__ void __launch_bounds__(kTmaBlockThreads, 2)
k_rollout_tma(EnvArrays A, uint64_t step, int player_sel, T* __restrict__ out, uint32_t stride, uint32_t env_base) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TmaSmem sm(smem_raw, warp);
  BlockStats st;
  st.init(sm.stats);
  const uint32_t b0 = env_base + blockIdx.x * (kTmaWarpsPerBlock * 32u);
  const uint32_t e0 = b0 + warp * 32u;
  const uint32_t e = e0 + lane;
  const bool active = e < A.n;
  const bool block_full = b0 + kTmaWarpsPerBlock * 32u <= A.n;  // uniform over the block
  const bool both = player_sel == COUP_PLAYER_BOTH;
  if (block_full) zero_stage(sm.stage, lane);
  Env s = {};
  uint32_t* hist_row = A.history + static_cast<size_t>(e) * kHistoryWords;
  uint32_t* rec = sm.recs + lane * kRecWords;
  if (active) s = load_env_and_row(A, e, rec);
  const StepResult r = step_env<true>(s, HistRow{rec, hist_row}, 0, nullptr, A, e, step, active);
  if (active) {
    store_env(A.state + e, s);
    write_outputs(A, e, r);
    fill_record(rec, s, rec, player_sel);
  }
  account(st, r, active);
  if (block_full) {
    __syncthreads();   // every warp's records are in shared memory
    block_encode_info_tma<T>(sm.block_recs, both, reinterpret_cast<T*>(sm.stage),
                             reinterpret_cast<unsigned char*>(out) + static_cast<size_t>(b0) * (both ? 2 : 1) * stride * sizeof(T),
                             warp, lane, static_cast<int>(stride));
  } else if (e0 < A.n) {
    __syncwarp();
    using U = typename Unit4<T>::type;
    const size_t row0 = static_cast<size_t>(e0) * (both ? 2 : 1);
    warp_encode_info<T>(sm.recs, static_cast<int>(min(32u, A.n - e0)), both, reinterpret_cast<U*>(out) + row0 * (stride / 4), lane,
                        static_cast<int>(stride / 4));
  }
  st.flush(A.stats);
}

// ---- warp-specialised persistent variant of the fused step --------------------------------------------------
// One 24-warp CTA per SM loops over 256-env batches. Warps 0..7 run the RULES for batch i+1 (thread per env,
// one 32-env group each) and leave the encoder records in one half of a double-buffered shared-memory area,
// while warps 8..23 ENCODE batch i with bulk stores (one staging buffer each, row groups dealt round-robin).
// The two roles hand batches over through named barriers (full/empty per buffer), so the bulk-store stream never
// pauses for a rules phase -- in k_rollout_tma every warp of a CTA stops storing while it steps its envs.
// The record area is double-buffered (kWsBufs). In-kernel cycle counters (-DCOUP_WS_DEBUG, scripts/ws_debug_probe.py)
// show neither role ever waiting for the other: next to the saturated store stream the rules warps' loads queue
// behind it, a rules phase stretches to one batch time (15 / 28 / 54 us for u8 / bf16 / f32) and finishes as the
// encoders do; a third record buffer changed nothing (1.538 / 0.806 / 0.430 ms per step either way).
constexpr int kWsWarps = 24, kWsRulesWarps = 8, kWsEncWarps = kWsWarps - kWsRulesWarps;
constexpr int kWsThreads = kWsWarps * 32;
constexpr int kWsBatch = kTmaWarpsPerBlock * 32;                       // 256 envs
constexpr int kWsRecBytes = kWsBatch * kRecWords * 4;                  // 21 504 B per record buffer
constexpr int kWsBufs = 2;
constexpr int kWsSmemBytes = kWsEncWarps * kStageBytes + kWsBufs * kWsRecBytes + COUP_STATS_LEN * 4;
enum { kBarFull0 = 1, kBarEmpty0 = kBarFull0 + kWsBufs, kBarRules = kBarEmpty0 + kWsBufs };   // named barriers 1..5

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <typename T>
__global__ void __launch_bounds__(kWsThreads, 1)
k_rollout_ws(EnvArrays A, uint64_t step, int player_sel, T* __restrict__ out, uint32_t stride, uint32_t n_batches,
             unsigned int* __restrict__ batch_counter) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_batch[kWsBufs];   // batch held by each record buffer, -1 = no more work
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* stage_base = smem_raw;
  uint32_t* const rec0 = reinterpret_cast<uint32_t*>(smem_raw + kWsEncWarps * kStageBytes);
  BlockStats st;
  st.init(reinterpret_cast<uint32_t*>(smem_raw + kWsEncWarps * kStageBytes + kWsBufs * kWsRecBytes));
  const bool both = player_sel == COUP_PLAYER_BOTH;
  const bool rules = warp < kWsRulesWarps;
  unsigned char* stage = rules ? nullptr : stage_base + static_cast<size_t>(warp - kWsRulesWarps) * kStageBytes;
  if (!rules) zero_stage(stage, lane);
  // Batches are handed out dynamically (global counter): SMs differ by ~20 % in achieved store bandwidth, so a
  // static split would leave the fast ones idle at the end.
#ifdef COUP_WS_DEBUG
  // Cycle counters of one rules warp and one encoder warp per CTA, summed into the spare statistics slots 24..29:
  // rules busy / rules waiting for a free record buffer / encoder waiting for records / encoder busy / CTA total /
  // batches. Build with COUP_B200_NVCC_EXTRA=-DCOUP_WS_DEBUG; read with coup_vec_stats.
  long long dbg_busy = 0, dbg_wait = 0, dbg_t, dbg_t0 = clock64();
  int dbg_batches = 0;
#define WS_DBG_MARK() (dbg_t = clock64())
#define WS_DBG_ADD(var) ((var) += clock64() - dbg_t)
#else
#define WS_DBG_MARK()
#define WS_DBG_ADD(var)
#endif
  unsigned int next_batch = 0;
  for (int it = 0;; ++it) {
    const int buf = it % kWsBufs;
    uint32_t* recs = rec0 + buf * (kWsRecBytes / 4);
    if (rules) {
      WS_DBG_MARK();
      if (it >= kWsBufs) named_bar_sync(kBarEmpty0 + buf, kWsThreads);  // the encoders are done with this buffer
      WS_DBG_ADD(dbg_wait);
      WS_DBG_MARK();
      if (warp == 0 && lane == 0) {
        // The batch number was requested one iteration ago (the first one here): next to the saturated store stream
        // a global atomic takes microseconds to come back, and nothing of this batch can start before it does.
        const unsigned int b = it == 0 ? atomicAdd(batch_counter, 1u) : next_batch;
        s_batch[buf] = b < n_batches ? static_cast<int>(b) : -1;
        next_batch = atomicAdd(batch_counter, 1u);
      }
      named_bar_sync(kBarRules, kWsRulesWarps * 32);
      const int b = s_batch[buf];
      if (b >= 0) {
        for (int sub = warp; sub < kWsBatch / 32; sub += kWsRulesWarps) {
          const uint32_t e = static_cast<uint32_t>(b) * kWsBatch + sub * 32 + lane;
          uint32_t* rec = recs + (sub * 32 + lane) * kRecWords;
          Env s = load_env_and_row(A, e, rec);
          uint32_t* hist_row = A.history + static_cast<size_t>(e) * kHistoryWords;
          const StepResult r = step_env<true>(s, HistRow{rec, hist_row}, 0, nullptr, A, e, step, true);
          store_env(A.state + e, s);
          write_outputs(A, e, r);
          fill_record(rec, s, rec, player_sel);
          account(st, r, true);
        }
      }
      __threadfence_block();
      named_bar_arrive(kBarFull0 + buf, kWsThreads);                    // records (or the stop mark) are ready
      WS_DBG_ADD(dbg_busy);
      if (b < 0) {
        // drain the hand-backs nobody will wait for any more (the last kWsBufs - 1 encoded batches)
        for (int back = 1; back < kWsBufs; ++back)
          if (it - back >= 0) named_bar_sync(kBarEmpty0 + (it - back) % kWsBufs, kWsThreads);
        break;
      }
    } else {
      WS_DBG_MARK();
      named_bar_sync(kBarFull0 + buf, kWsThreads);
      WS_DBG_ADD(dbg_wait);
      const int b = s_batch[buf];
      if (b < 0) break;
      WS_DBG_MARK();
      block_encode_info_tma<T>(recs, both, reinterpret_cast<T*>(stage),
                               reinterpret_cast<unsigned char*>(out) +
                                   static_cast<size_t>(b) * kWsBatch * (both ? 2 : 1) * stride * sizeof(T),
                               warp - kWsRulesWarps, lane, static_cast<int>(stride), kWsEncWarps, /*wait_for_writes=*/false);
      named_bar_arrive(kBarEmpty0 + buf, kWsThreads);                   // hand the record buffer back
      WS_DBG_ADD(dbg_busy);
#ifdef COUP_WS_DEBUG
      ++dbg_batches;
#endif
    }
  }
  if (!rules && lane == 0) tma_wait_all();
  st.flush(A.stats);
#ifdef COUP_WS_DEBUG
  if (lane == 0 && (warp == 0 || warp == kWsRulesWarps)) {
    const int base = warp == 0 ? 24 : 26;   // rules: busy, wait | encoder: wait, busy
    atomicAdd(&A.stats[base], static_cast<unsigned long long>(warp == 0 ? dbg_busy : dbg_wait));
    atomicAdd(&A.stats[base + 1], static_cast<unsigned long long>(warp == 0 ? dbg_wait : dbg_busy));
    if (warp == 0) atomicAdd(&A.stats[28], static_cast<unsigned long long>(clock64() - dbg_t0));
    else atomicAdd(&A.stats[29], static_cast<unsigned long long>(dbg_batches));
  }
#endif
}

}  // namespace coup
