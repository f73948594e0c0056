// coup_record.cuh -- self-play recording fused into the step (coup_vec_step_record): NFSP reservoir and DQN replay kept by
// the thread that steps the env, as packed 96-byte observation records.
#pragma once
#include "coup_step.cuh"

namespace coup {

// ---- self-play recording fused into the step (coup_vec_step_record) -----------------------------------------------------
// What the reference's agents keep per decision (python/algorithms/nfsp.py:226-242 `Transition(info_state, action_probs,
// legal_actions_mask)` into a reservoir, :322-371; python/algorithms/dqn.py:30-32,223-246 `Transition(info_state, action,
// reward, next_info_state, is_final_step, legal_actions_mask)` into a circular replay buffer) is recorded by the thread
// that steps the env, as PACKED observation records: 96 bytes (history + state + meta) instead of a 2492-element row. A
// record decodes into exactly the row the dense encoder writes (k_encode_info* with a RecordSource), so a learner
// materialises rows only for the batch it samples.
struct RecorderArrays {
  uint32_t* res_records;            // [res_capacity][24]  NFSP reservoir, or nullptr
  float* res_probs;                 // [res_capacity][18]
  unsigned long long* res_winner;   // [res_capacity] running index + 1 of the element that owns the slot
  unsigned long long res_capacity;
  unsigned long long res_base;      // elements offered before this step; env e of this step is element res_base + e
  uint32_t* transitions;            // [rb_capacity][2][24]  DQN replay (record of s, record of s'), or nullptr
  unsigned long long rb_capacity;
  unsigned long long* rb_total;     // transitions ever written (device counter)
  uint32_t* pending;                // [n][2][24] the previous decision of each seat; bit 30 of meta word 1 = valid
};

// Reservoir slot of the element with running index t (nfsp.py:340-356): t itself while the buffer fills, afterwards
// randint(0, t) if that is below the capacity. ~0ull = not stored.
__device__ __forceinline__ unsigned long long reservoir_slot(const EnvArrays& A, const RecorderArrays& R, uint32_t e, uint64_t step) {
  const unsigned long long t = R.res_base + e;
  if (t < R.res_capacity) return t;
  const uint4 rnd = env_random(A.seed, A.global_env_offset + e, step, 5);
  const unsigned long long u = (static_cast<unsigned long long>(rnd.x) << 32) | rnd.y;
  const unsigned long long draw = __umul64hi(u, t + 1ull);
  return draw < R.res_capacity ? draw : ~0ull;
}

// Pass 1: every env offers its decision; of the elements that draw the same slot in one step the LATER one must win, as it
// would sequentially, so slots are claimed with an atomic max of the running index before anything is written.
__global__ void __launch_bounds__(kBlockThreads)
k_reservoir_claim(EnvArrays A, RecorderArrays R, const uint8_t* __restrict__ actions, uint64_t step) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n || actions[e] == 0xFFu) return;
  const unsigned long long slot = reservoir_slot(A, R, e, step);
  if (slot != ~0ull) atomicMax(&R.res_winner[slot], R.res_base + e + 1ull);
}

__device__ __forceinline__ void store_record(uint32_t* dst, const uint4 (&h)[4], const Env& s, uint32_t m0, uint32_t m1,
                                             uint32_t m2, uint32_t m3) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int k = 0; k < 4; ++k) d[k] = h[k];
  d[4] = make_uint4(s.p[0], s.p[1], s.g, s.c);
  d[5] = make_uint4(m0, m1, m2, m3);
}

// One replay transition: the seat's pending record becomes `info_state` (its meta word 1 gains reward and is_final), the
// given observation becomes `next_info_state`. Slots from one atomic cursor bumped once per group of emitting lanes.
__device__ __forceinline__ void emit_transition(const RecorderArrays& R, uint32_t* pend, int reward, const uint4 (&h)[4],
                                                const Env& next, uint32_t e, uint32_t seat, uint32_t is_final,
                                                uint32_t legal_next) {
  const uint32_t peers = __activemask();
  const uint32_t lane = threadIdx.x & 31u;
  const int leader = __ffs(peers) - 1;
  unsigned long long base = 0;
  if (static_cast<int>(lane) == leader) base = atomicAdd(R.rb_total, static_cast<unsigned long long>(__popc(peers)));
  base = __shfl_sync(peers, base, leader);
  const unsigned long long ticket = base + __popc(peers & ((1u << lane) - 1u));
  uint32_t* dst = R.transitions + (ticket % R.rb_capacity) * (2 * kRecordWords);
  const uint4* p4 = reinterpret_cast<const uint4*>(pend);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int k = 0; k < 5; ++k) d4[k] = p4[k];
  const uint4 pm = p4[5];   // env | seat<<31, valid<<30, action | - | -
  d4[5] = make_uint4(pm.x, (pm.y & 0x8000001Fu) | (static_cast<uint32_t>(reward + 2) << 5) | (is_final << 8),
                     static_cast<uint32_t>(ticket), static_cast<uint32_t>(ticket >> 32));
  store_record(dst + kRecordWords, h, next, e, (seat << 31) | legal_next, 0u, 0u);
}

// Pass 2: reservoir commit, replay bookkeeping, and the step itself (same semantics as k_step).
__global__ void __launch_bounds__(kBlockThreads)   // more resident CTAs make it slower (69 / 74 / 80 / 90 us at 1 / 3 / 4 / 5 per SM)
k_step_record(EnvArrays A, RecorderArrays R, const uint8_t* __restrict__ actions, const float* __restrict__ probs,
              uint64_t step) {
  __shared__ uint32_t s_stats[COUP_STATS_LEN];
  BlockStats st;
  st.init(s_stats);
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t action = e < A.n ? actions[e] : 0xFFu;
  const bool active = action != 0xFFu;
  Env s = {};
  uint32_t* hist_row = A.history + static_cast<size_t>(e) * kHistoryWords;
  uint4 h[4] = {};
  if (active) {
    s = load_env(A.state + e);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = reinterpret_cast<const uint4*>(hist_row)[k];
  }
  const bool deciding = active && !is_terminal(s) && !g_chance(s.g);
  const uint32_t seat = g_mover(s.g);
  const uint32_t legal0 = legal_mask_decision(s);
  const uint32_t m0 = c_moves(s.c);
  if (deciding && R.res_records != nullptr) {                      // nfsp.py:226-242
    const unsigned long long slot = reservoir_slot(A, R, e, step);
    const unsigned long long t = R.res_base + e;
    if (slot != ~0ull && R.res_winner[slot] == t + 1ull) {
      store_record(R.res_records + slot * kRecordWords, h, s, e, (seat << 31) | legal0, static_cast<uint32_t>(t),
                   static_cast<uint32_t>(t >> 32));
      float* dst = R.res_probs + slot * kNumActions;
      const float* src = probs + static_cast<size_t>(e) * kNumActions;
#pragma unroll
      for (int a = 0; a < kNumActions; ++a) dst[a] = src[a];
    }
  }
  if (deciding && R.transitions != nullptr) {                      // dqn.py:223-246: the seat acts again
    uint32_t* pend = R.pending + (static_cast<size_t>(e) * 2 + seat) * kRecordWords;
    const int rew0 = c_reward0(s.c);
    if ((pend[21] >> 30) & 1u) emit_transition(R, pend, seat == 0u ? rew0 : -rew0, h, s, e, seat, 0u, legal0);
    store_record(pend, h, s, e, (seat << 31) | (1u << 30) | action, 0u, 0u);
  }
  const StepResult r = step_env<false>(s, global_row(hist_row), action, nullptr, A, e, step, active);
  if (active) {
    store_env(A.state + e, s);
    write_outputs(A, e, r);
  }
  if (r.finished && R.transitions != nullptr) {
    // Every agent is stepped with the final time step (coup_experiments/scripts/nfsp.py:141-143). The finished episode's
    // row: word 0 from before the step (a re-deal rewrites only that word), patched if the last action landed in it.
    uint4 ht[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ht[k] = reinterpret_cast<const uint4*>(hist_row)[k];
    ht[0].x = m0 < 6u ? h[0].x | (action << (5u * m0)) : h[0].x;
#pragma unroll
    for (uint32_t p = 0; p < 2; ++p) {
      uint32_t* pend = R.pending + (static_cast<size_t>(e) * 2 + p) * kRecordWords;
      if ((pend[21] >> 30) & 1u) {
        emit_transition(R, pend, p == 0u ? r.reward0 : -r.reward0, ht, r.final_state, e, p, 1u, 0u);
        pend[21] = 0u;
      }
    }
  }
  account(st, r, active);
  st.flush(A.stats);
}

}  // namespace coup
