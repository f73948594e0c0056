// coup_step.cuh -- the env slab, the per-env decision step and the kernels that only step: reset, step with caller-provided
// actions, single moves with explicit chance nodes, clone, and the batched state.child(action) of the tree traversals.
//
// Thread mapping: one thread per environment (16-byte state load/store, coalesced); the step is one instruction stream for
// the 32 envs of a warp (coup_device.cuh). No tensor cores: there is no contraction anywhere on this path.
#pragma once
#include <cuda_bf16.h>

#include "coup_device.cuh"
#include "../../include/coup_b200.h"

namespace coup {

constexpr int kBlockThreads = 256;
#ifndef COUP_ENV_BLOCKS
#define COUP_ENV_BLOCKS 3   // resident CTAs per SM the env-only rollout is compiled for (3 -> 80 registers, no spills)
#endif
constexpr int kWarpsPerBlock = kBlockThreads / 32;
constexpr int kUnitsPerInfoRow = kInfoStateSize / 4;  // 623 four-element units (16 B in fp32)
constexpr int kRecWords = 21;                         // per-env encoder record in shared memory

struct EnvArrays {
  uint4* state;        // [n]
  uint32_t* history;   // [n][16]
  uint32_t* legal;     // [n]
  int8_t* cur_player;  // [n]
  uint8_t* done;       // [n]
  int8_t* rewards;     // [n][2]
  int8_t* returns;     // [n][2]
  uint32_t* step_word; // [n] everything a host-side policy needs in one word (COUP_WORD_* in coup_b200.h)
  unsigned long long* stats;  // [COUP_STATS_LEN]
  uint32_t n;
  uint32_t flags;
  uint64_t seed;
  uint64_t global_env_offset;
  // finished-episode ring (coup_vec_finished_ring_enable): [ring_mask + 1][COUP_RECORD_WORDS], or nullptr
  uint32_t* ring;
  unsigned long long* ring_ctrl;  // [0] records ever appended, [1] value of [0] when the last step/rollout call began
  uint32_t ring_mask;
};

constexpr int kRecordWords = COUP_RECORD_WORDS;   // packed observation record: 16 history words, 4 state words, 4 meta words

__device__ __forceinline__ Env load_env(const uint4* p) {
  uint4 v = *p;
  Env s;
  s.p[0] = v.x; s.p[1] = v.y; s.g = v.z; s.c = v.w;
  return s;
}
__device__ __forceinline__ void store_env(uint4* p, const Env& s) { *p = make_uint4(s.p[0], s.p[1], s.g, s.c); }

// The fused step kernels issue ALL the global loads of an env at once -- its state word and its 64-byte history row,
// the row straight into the env's encoder record in shared memory -- and never load again: the step updates the row in
// the record and writes the changed words through to HBM. Next to a saturated store stream every dependent global round
// trip of the rules costs microseconds (scripts/ws_debug_probe.py), so the rules phase is ONE round trip, not four.
__device__ __forceinline__ Env load_env_and_row(const EnvArrays& A, uint32_t e, uint32_t* rec) {
  const uint4 sv = A.state[e];
  const uint4* g4 = reinterpret_cast<const uint4*>(A.history + static_cast<size_t>(e) * 16u);
  const uint4 h0 = g4[0], h1 = g4[1], h2 = g4[2], h3 = g4[3];
  rec[0] = h0.x; rec[1] = h0.y; rec[2] = h0.z; rec[3] = h0.w; rec[4] = h1.x; rec[5] = h1.y; rec[6] = h1.z; rec[7] = h1.w;
  rec[8] = h2.x; rec[9] = h2.y; rec[10] = h2.z; rec[11] = h2.w; rec[12] = h3.x; rec[13] = h3.y; rec[14] = h3.z; rec[15] = h3.w;
  Env s;
  s.p[0] = sv.x; s.p[1] = sv.y; s.g = sv.z; s.c = sv.w;
  return s;
}


// ---- statistics: warp ballots -> shared counters -> one global atomic per counter per block ------
struct BlockStats {
  uint32_t* sm;  // [COUP_STATS_LEN] in shared memory
  __device__ __forceinline__ void init(uint32_t* shared) {
    sm = shared;
    for (int i = threadIdx.x; i < COUP_STATS_LEN; i += blockDim.x) sm[i] = 0;
    __syncthreads();
  }
  // All 32 lanes of the warp must call these (inactive envs pass pred=false / value 0).
  __device__ __forceinline__ void count(int idx, bool pred) {
    uint32_t b = __ballot_sync(0xffffffffu, pred);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&sm[idx], __popc(b));
  }
  __device__ __forceinline__ void sum(int idx, uint32_t v) {
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sm[idx], v);
  }
  // Histogram over `nb` bins of a value known to be < nb for lanes with pred set.
  __device__ __forceinline__ void hist(int base, int nb, uint32_t value, bool pred) {
    for (int b = 0; b < nb; ++b) count(base + b, pred && value == static_cast<uint32_t>(b));
  }
  __device__ __forceinline__ void flush(unsigned long long* global) {
    __syncthreads();
    for (int i = threadIdx.x; i < COUP_STATS_LEN; i += blockDim.x)
      if (sm[i]) atomicAdd(&global[i], static_cast<unsigned long long>(sm[i]));
  }
};

// ---- the per-env step, shared by k_step and k_rollout --------------------------------------------
struct StepResult {
  uint32_t legal;     // legal mask of the state left in `s`
  int cur_player;     // 0/1/-4
  bool done;          // the state reached by this step was terminal (reported even if auto-reset)
  int reward0;        // Rewards()[0] of the stepped state
  int return0;        // Returns()[0] of the stepped state
  bool stepped;       // a player action was applied
  bool illegal;
  uint32_t n_legal_before;
  uint32_t chance_moves;
  bool finished;      // an episode ended in this call
  bool truncated;
  uint32_t final_moves;
  Env final_state;    // valid when `finished`: the terminal state, before any re-deal (unused fields are optimised away)
};

// Re-deal a fresh episode into `s` (CoupState ctor + the 4 initial chance nodes). With forced outcomes (known-answer
// replay) the deals run through the generic chance loop, else through the closed form; both give the same state for the
// same Philox words. Writes history word 0 and returns the number of deals made.
// The history row a step works on: `work` is read and updated (the env's row in HBM itself, or a copy the kernel holds in
// shared memory), `mirror` is the HBM row when `work` is a copy (written through, never read) and nullptr otherwise.
struct HistRow {
  uint32_t* work;
  uint32_t* mirror;
};
__device__ __forceinline__ HistRow global_row(uint32_t* row) { return HistRow{row, nullptr}; }

__device__ __forceinline__ uint32_t deal_new_episode(Env& s, HistRow row, const uint4& rnd,
                                                     const uint8_t* forced) {
  uint32_t codes = 0, n_codes = 0;
  if (forced != nullptr) {
    s = initial_state();
    resolve_chance(s, rnd, 0, forced, codes, n_codes);
  } else {
    s = dealt_initial_state(rnd, codes);
    n_codes = 4;
  }
  row.work[0] = codes;
  if (row.mirror != nullptr) row.mirror[0] = codes;
  return n_codes;
}

// The episode of env `e` has just ended in terminal state `s`: its trajectory log, terminal state and outcome go to the
// finished-episode ring BEFORE an auto-reset re-deals the env in place. This is what SyncVectorEnv.step hands back as
// `unreset_time_steps` (python/vector_env.py:52-66) and what every agent is stepped with at episode end
// (coup_experiments/scripts/nfsp.py:141-143): from the record, the terminal info-state rows of both players are encoded on
// demand (k_encode_info* with a RecordSource) and the whole episode replays through the reference.
// Slots come from ONE atomic cursor, bumped once per group of lanes that finish together (opportunistic aggregation).
// Two halves, so that the round trip of the atomic hides behind the rest of the step: ring_reserve issues it where the
// episode ends, ring_write -- called by the same lanes once the deals of the step are done -- consumes the slot.
struct RingTicket {
  unsigned long long base;   // the leader's atomic result
  uint32_t peers;
};
__device__ __forceinline__ RingTicket ring_reserve(const EnvArrays& A) {
  RingTicket t{0ull, 0u};
  if (A.ring == nullptr) return t;
  t.peers = __activemask();
  const uint32_t lane = threadIdx.x & 31u;
  if (static_cast<int>(lane) == __ffs(t.peers) - 1) t.base = atomicAdd(A.ring_ctrl, static_cast<unsigned long long>(__popc(t.peers)));
  return t;
}
__device__ __forceinline__ void ring_write(const EnvArrays& A, const RingTicket& t, uint32_t e, const Env& s,
                                           const uint32_t* hist_row, uint64_t step, bool truncated) {
  if (A.ring == nullptr) return;       // hist_row: the working copy of the finished episode's row, any alignment
  const uint32_t lane = threadIdx.x & 31u;
  const unsigned long long base = __shfl_sync(t.peers, t.base, __ffs(t.peers) - 1);
  const uint32_t slot = static_cast<uint32_t>(base + __popc(t.peers & ((1u << lane) - 1u))) & A.ring_mask;
  uint4* dst = reinterpret_cast<uint4*>(A.ring + static_cast<size_t>(slot) * kRecordWords);
  if ((reinterpret_cast<uintptr_t>(hist_row) & 15u) == 0) {      // the env's own row in HBM
#pragma unroll
    for (int k = 0; k < kHistoryWords / 4; ++k) dst[k] = reinterpret_cast<const uint4*>(hist_row)[k];
  } else {                                                        // a copy in a shared-memory record (odd pitch)
#pragma unroll
    for (int k = 0; k < kHistoryWords / 4; ++k)
      dst[k] = make_uint4(hist_row[4 * k], hist_row[4 * k + 1], hist_row[4 * k + 2], hist_row[4 * k + 3]);
  }
  dst[4] = make_uint4(s.p[0], s.p[1], s.g, s.c);
  const uint32_t meta = c_moves(s.c) | (static_cast<uint32_t>(returns_p0(s) + 2) << 8) |
                        (static_cast<uint32_t>(c_reward0(s.c) + 2) << 12) | (truncated ? 1u << 16 : 0u);
  dst[5] = make_uint4(e, meta, static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32));
}

// Runs (one thread) in front of every step/rollout launch: remembers where the ring stood, so that "the episodes that
// finished in the last call" is the range [ctrl[1], ctrl[0]), and re-arms the persistent kernel's batch counter.
__global__ void k_step_prologue(unsigned long long* ring_ctrl, unsigned int* batch_counter) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (ring_ctrl != nullptr) ring_ctrl[1] = ring_ctrl[0];
    if (batch_counter != nullptr) *batch_counter = 0u;
  }
}

// One decision step of one env per lane: sample or take the action, apply it, resolve the chance nodes that follow,
// log, hand a finished episode to the ring and (auto-reset) re-deal it. CONVERGENT: all 32 lanes of the warp call this
// together (`active` = this lane has an env to step), so that the step is one instruction stream under warp-uniform
// guards -- "does any lane finish an episode", "does any lane still have a deal pending" -- with selects inside.
// Per-lane branches remain only around memory side effects (the history / ring writes of the ~2 lanes in 32 that end
// an episode) and the once-in-10^6-episodes move cap that falls in the middle of a deal sequence.
template <bool kSample, bool kLegalKnown = false>
__device__ __forceinline__ StepResult step_env(Env& s, HistRow row, uint32_t action_in,
                                               const uint8_t* forced, const EnvArrays& A, uint32_t e,
                                               uint64_t step, bool active, uint32_t legal_known = 0u) {
  constexpr uint32_t kFull = 0xffffffffu;
  StepResult r;
  r.chance_moves = 0; r.truncated = false; r.final_moves = 0;
  const uint64_t genv = A.global_env_offset + e;
  const bool auto_reset = (A.flags & COUP_FLAG_AUTO_RESET) != 0;
  // kLegalKnown: the caller still holds the mask the previous step returned for this very state; that mask is empty
  // exactly when the state is terminal (a decision node always has a legal action, a chance node a card to deal)
  const bool term0 = kLegalKnown ? legal_known == 0u : is_terminal(s);
  const bool chance0 = g_chance(s.g) != 0;
  const uint32_t legal0 = kLegalKnown ? legal_known : legal_mask_decision(s);
  const uint4 rnd = env_random(A.seed, genv, step, 0);
  const uint32_t a = kSample ? sample_action(legal0, rnd.x) : action_in;
  // go: a player action is applied. An env left at an explicit chance node (coup_vec_new_initial_state /
  // coup_vec_apply_move) has no player to move and is refused like an illegal action.
  const bool go = active && !term0 && !chance0 && a < 18u && ((legal0 >> a) & 1u);
  r.stepped = go;
  r.illegal = active && !term0 && !go;
  r.n_legal_before = go ? popc32(legal0) : 0u;
  s.g |= r.illegal ? kBitError : 0u;        // sticky; the reference would SpielFatalError / raise (rl_environment.py:270-280)
  const uint32_t m0 = c_moves(s.c);
  bool fin = false;       // an episode ended in this call
  bool term = term0;      // the state left in `s` is terminal
  // Rewards() / Returns() of the stepped state: deals change neither, and an env that does not step keeps its own.
  if (!__any_sync(kFull, go)) {
    r.reward0 = c_reward0(s.c);
    r.return0 = returns_p0(s);
  } else {
    Env t = s;
    apply_player_action(t, a);
    s.p[0] = go ? t.p[0] : s.p[0]; s.p[1] = go ? t.p[1] : s.p[1]; s.g = go ? t.g : s.g; s.c = go ? t.c : s.c;
    r.reward0 = c_reward0(s.c);
    r.return0 = returns_p0(s);
    fin = go && is_terminal(s);
    term = go ? fin : term0;
    // pending history codes of this step: `n_codes` codes that become moves first .. of the row
    uint32_t codes = a, n_codes = go ? 1u : 0u, first = m0;
    RingTicket ticket{0ull, 0u};
    const bool any_fin = __any_sync(kFull, fin);
    if (any_fin) {
      if (fin) {                                   // memory side effects of the lanes that end an episode
        r.final_state = s;
        r.final_moves = c_moves(s.c);
        r.truncated = r.final_moves > kMaxGameLength;
        history_commit(row.work, m0, a, 1u, row.mirror);
        ticket = ring_reserve(A);
      }
      // Re-deal in place with the closed-form deal, computed by the whole warp. An episode that ends AT the action has
      // used none of the three deal words of its step block, so the four cards come from them: y serves two draws
      // (floor(y * 15 / 2^32), then its remainder y * 15 mod 2^32, again uniform), z and w one each.
      const uint4 rr = make_uint4(rnd.y, rnd.y * 15u, rnd.z, rnd.w);
      uint32_t fresh_codes;
      const Env fresh = dealt_initial_state(rr, fresh_codes);
      const bool redeal = fin && auto_reset;
      s.p[0] = redeal ? fresh.p[0] : s.p[0]; s.p[1] = redeal ? fresh.p[1] : s.p[1];
      s.g = redeal ? fresh.g : s.g; s.c = redeal ? fresh.c : s.c;
      codes = redeal ? fresh_codes : codes;
      n_codes = redeal ? 4u : (fin ? 0u : n_codes);      // a finished episode's last move is already in the row
      first = redeal ? 0u : first;
      r.chance_moves = redeal ? 4u : 0u;
      term = redeal ? false : term;
    }
    // The deals that follow the action (at most three: Exchange after a lost challenge). Deals never change who is
    // alive, so inside the loop only the move cap (coup.cc:990) can end the game; a freshly dealt or finished env has
    // nothing pending.
    // All deals queued by a player action go to ONE player (apply_player_action: bit 28), so the loop works on that
    // player's hand and the deck only; queue count, chance flag, move number and the player word are settled once after
    // it. nd = deals this lane makes: the whole queue, cut short by the move cap.
    const uint32_t target = (s.g >> 28) & 1u, qn = g_qn(s.g);
    const uint32_t room = static_cast<uint32_t>(kMaxGameLength + 1) - umin32(c_moves(s.c), kMaxGameLength + 1);
    const uint32_t nd = (go && g_chance(s.g)) ? umin32(qn, room) : 0u;
    const uint32_t tw = get_p(s, target);
    uint32_t hand = pw_hand(tw), g = s.g;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const bool pend = static_cast<uint32_t>(k) < nd;
      if (!__any_sync(kFull, pend)) break;
      uint32_t card = sample_card_g(g, k == 0 ? rnd.y : k == 1 ? rnd.z : rnd.w);
      if (forced != nullptr) {
        const uint32_t f = pend ? forced[k] : 0xFFu;   // lanes without an env must not touch the array
        card = (f < 5u && g_deck(g, f) != 0) ? f : card;
      }
      g -= pend ? 1u << (4u * card) : 0u;                                  // deck_[card] -= 1 (coup.cc:491-520)
      hand = pend ? hand_insert(hand, card << 1) : hand;
      // a lane that deals has logged exactly its action so far: deal k is code 1 + k of the step
      codes |= pend ? (18u + 5u * target + card) << (5u * (k + 1)) : 0u;
    }
    g -= nd << 24;                                                         // pop
    g &= (nd != 0u && nd == qn) ? ~(kBitChance | kBitQInitial) : ~0u;      // queue empty: is_chance_ = false (520)
    s.g = g;
    set_p(s, target, nd ? pw_set_hand(tw, hand) : tw);
    s.c += nd;                                                             // ++move_number_ per deal
    n_codes += nd;
    r.chance_moves += nd;
    if (any_fin && fin) ring_write(A, ticket, e, r.final_state, row.work, step, r.truncated);   // before word 0 is re-dealt
    if (n_codes) history_commit(row.work, first, codes, n_codes, row.mirror);
    if (go && !fin && c_moves(s.c) > kMaxGameLength) {
      // the move cap fell in the middle of a deal sequence: once in ~10^6 episodes, a slow path of its own
      fin = true;
      term = true;
      r.final_state = s;
      r.final_moves = c_moves(s.c);
      r.truncated = true;
      ring_write(A, ring_reserve(A), e, s, row.work, step, true);
      if (auto_reset) {
        const uint4 rr = env_random(A.seed, genv, step, 1);
        r.chance_moves += deal_new_episode(s, row, rr, nullptr);
        term = false;
      }
    }
  }
  r.done = fin || (!go && term0);
  r.finished = fin;
  const bool chance = !term && g_chance(s.g);   // only an env that was refused at an explicit chance node
  r.legal = term ? 0u : (chance ? legal_mask_chance(s) : legal_mask_decision(s));
  r.cur_player = term ? COUP_TERMINAL_PLAYER_ID : (chance ? COUP_CHANCE_PLAYER_ID : static_cast<int>(g_mover(s.g)));
  return r;
}

__device__ __forceinline__ void write_outputs(const EnvArrays& A, uint32_t e, const StepResult& r) {
  A.legal[e] = r.legal;
  A.cur_player[e] = static_cast<int8_t>(r.cur_player);
  A.done[e] = r.done ? 1 : 0;
  reinterpret_cast<char2*>(A.rewards)[e] = make_char2(static_cast<signed char>(r.reward0), static_cast<signed char>(-r.reward0));
  reinterpret_cast<char2*>(A.returns)[e] = make_char2(static_cast<signed char>(r.return0), static_cast<signed char>(-r.return0));
  A.step_word[e] = r.legal | (r.cur_player == 1 ? 1u << 18 : 0u) | (r.cur_player == COUP_TERMINAL_PLAYER_ID ? 1u << 19 : 0u) |
                   (r.done ? 1u << 20 : 0u) | (static_cast<uint32_t>(r.reward0 + 2) << 21) |
                   (static_cast<uint32_t>(r.return0 + 2) << 24);
}

// Statistics of the steps one lane makes, kept in registers as packed 8-bit (16-bit) per-lane counters and turned into
// warp sums only when flushed: the accounting of a step is ~20 ALU instructions, and the 11 warp reductions + shared-memory
// atomics are paid once per launch (or every kMaxAdds steps), not once per step. A lane may add() at most kMaxAdds times
// between flushes (8-bit fields: one count per add).
struct StatAcc {
  static constexpr int kMaxAdds = 64;
  uint32_t misc;                 // stepped | finished << 8 | truncated << 16 | illegal << 24
  uint32_t chance;               // chance moves (<= 7 per step)
  uint32_t moves;                // sum of final move numbers (<= 91 per step)
  unsigned long long returns;    // Returns()[0] histogram of finished episodes, 5 bins x 8 bits
  unsigned long long legal;      // legal-count histogram of the steps made, 8 bins x 8 bits
  __device__ __forceinline__ void clear() { misc = chance = moves = 0u; returns = legal = 0ull; }
  __device__ __forceinline__ void add(const StepResult& r, bool active) {
    const bool stepped = active && r.stepped, finished = active && r.finished;
    misc += (stepped ? 1u : 0u) | (finished ? 1u << 8 : 0u) | ((active && r.truncated) ? 1u << 16 : 0u) |
            ((active && r.illegal) ? 1u << 24 : 0u);
    chance += active ? r.chance_moves : 0u;
    moves += finished ? r.final_moves : 0u;
    returns += finished ? 1ull << (8 * (r.return0 + 2)) : 0ull;
    legal += stepped ? 1ull << (8u * min(r.n_legal_before, 7u)) : 0ull;
  }
  // All 32 lanes together. Even/odd fields are summed as 16-bit pairs (64 adds x 32 lanes < 2^16).
  __device__ __forceinline__ void flush(BlockStats& st) {
    constexpr uint32_t kFull = 0xffffffffu, kEven = 0x00FF00FFu;
    const uint32_t m0 = __reduce_add_sync(kFull, misc & kEven), m1 = __reduce_add_sync(kFull, (misc >> 8) & kEven);
    const uint32_t ch = __reduce_add_sync(kFull, chance), mv = __reduce_add_sync(kFull, moves);
    const uint32_t rl = static_cast<uint32_t>(returns), rh = static_cast<uint32_t>(returns >> 32);
    const uint32_t r0 = __reduce_add_sync(kFull, rl & kEven), r1 = __reduce_add_sync(kFull, (rl >> 8) & kEven);
    const uint32_t r2 = __reduce_add_sync(kFull, rh & 0xFFu);
    const uint32_t ll = static_cast<uint32_t>(legal), lh = static_cast<uint32_t>(legal >> 32);
    const uint32_t l0 = __reduce_add_sync(kFull, ll & kEven), l1 = __reduce_add_sync(kFull, (ll >> 8) & kEven);
    const uint32_t l2 = __reduce_add_sync(kFull, lh & kEven), l3 = __reduce_add_sync(kFull, (lh >> 8) & kEven);
    const int lane = threadIdx.x & 31;
    uint32_t val = 0;
    int idx = 0;
    switch (lane) {
      case 0: val = m0 & 0xFFFFu; idx = COUP_STAT_DECISION_STEPS; break;
      case 1: val = m1 & 0xFFFFu; idx = COUP_STAT_EPISODES; break;
      case 2: val = m0 >> 16; idx = COUP_STAT_TRUNCATED; break;
      case 3: val = m1 >> 16; idx = COUP_STAT_ILLEGAL; break;
      case 4: val = ch; idx = COUP_STAT_CHANCE_MOVES; break;
      case 5: val = mv; idx = COUP_STAT_EPISODE_MOVES; break;
      case 6: val = r0 & 0xFFFFu; idx = COUP_STAT_RETURN_HIST + 0; break;
      case 7: val = r1 & 0xFFFFu; idx = COUP_STAT_RETURN_HIST + 1; break;
      case 8: val = r0 >> 16; idx = COUP_STAT_RETURN_HIST + 2; break;
      case 9: val = r1 >> 16; idx = COUP_STAT_RETURN_HIST + 3; break;
      case 10: val = r2; idx = COUP_STAT_RETURN_HIST + 4; break;
      case 11: val = l0 & 0xFFFFu; idx = COUP_STAT_LEGAL_HIST + 0; break;
      case 12: val = l1 & 0xFFFFu; idx = COUP_STAT_LEGAL_HIST + 1; break;
      case 13: val = l0 >> 16; idx = COUP_STAT_LEGAL_HIST + 2; break;
      case 14: val = l1 >> 16; idx = COUP_STAT_LEGAL_HIST + 3; break;
      case 15: val = l2 & 0xFFFFu; idx = COUP_STAT_LEGAL_HIST + 4; break;
      case 16: val = l3 & 0xFFFFu; idx = COUP_STAT_LEGAL_HIST + 5; break;
      case 17: val = l2 >> 16; idx = COUP_STAT_LEGAL_HIST + 6; break;
      case 18: val = l3 >> 16; idx = COUP_STAT_LEGAL_HIST + 7; break;
      default: break;
    }
    if (val) atomicAdd(&st.sm[idx], val);
    clear();
  }
};

// One step of one warp, accounted at once (the single-step kernels): small fields packed side by side (a count over
// 32 lanes fits 6 bits), four warp reductions, and lanes 0..18 each add one counter to shared memory.
__device__ __forceinline__ void account(BlockStats& st, const StepResult& r, bool active) {
  const uint32_t nl = min(r.n_legal_before, 7u);
  const bool stepped = active && r.stepped, finished = active && r.finished;
  // A: stepped | finished<<6 | truncated<<12 | illegal<<18 | chance moves<<24 (<= 7 per lane)
  uint32_t a = !active ? 0u : (r.stepped ? 1u : 0u) | (r.finished ? 1u << 6 : 0u) | (r.truncated ? 1u << 12 : 0u) |
                                  (r.illegal ? 1u << 18 : 0u) | (r.chance_moves << 24);
  // B: Returns()[0] histogram of finished episodes, 5 bins x 6 bits
  uint32_t b = finished ? 1u << (6 * (r.return0 + 2)) : 0u;
  // C: sum of final move numbers (12 bits, <= 32 x 91) | legal-count bins 0..2 ; D: legal-count bins 3..7
  uint32_t c = (finished ? r.final_moves : 0u) | ((stepped && nl < 3u) ? 1u << (12u + 6u * nl) : 0u);
  uint32_t d = (stepped && nl >= 3u) ? 1u << (6u * (nl - 3u)) : 0u;
  a = __reduce_add_sync(0xffffffffu, a);
  b = __reduce_add_sync(0xffffffffu, b);
  c = __reduce_add_sync(0xffffffffu, c);
  d = __reduce_add_sync(0xffffffffu, d);
  const int lane = threadIdx.x & 31;
  uint32_t val = 0;
  int idx = 0;
  if (lane == 0) { val = a & 63u; idx = COUP_STAT_DECISION_STEPS; }
  else if (lane == 1) { val = (a >> 6) & 63u; idx = COUP_STAT_EPISODES; }
  else if (lane == 2) { val = (a >> 12) & 63u; idx = COUP_STAT_TRUNCATED; }
  else if (lane == 3) { val = (a >> 18) & 63u; idx = COUP_STAT_ILLEGAL; }
  else if (lane == 4) { val = a >> 24; idx = COUP_STAT_CHANCE_MOVES; }
  else if (lane == 5) { val = c & 4095u; idx = COUP_STAT_EPISODE_MOVES; }
  else if (lane < 11) { val = (b >> (6 * (lane - 6))) & 63u; idx = COUP_STAT_RETURN_HIST + lane - 6; }
  else if (lane < 14) { val = (c >> (12 + 6 * (lane - 11))) & 63u; idx = COUP_STAT_LEGAL_HIST + lane - 11; }
  else if (lane < 19) { val = (d >> (6 * (lane - 14))) & 63u; idx = COUP_STAT_LEGAL_HIST + 3 + lane - 14; }
  if (val) atomicAdd(&st.sm[idx], val);
}

// ---- reset -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads)
k_reset(EnvArrays A, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ forced, uint64_t step) {
  __shared__ uint32_t s_stats[COUP_STATS_LEN];
  BlockStats st;
  st.init(s_stats);
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = e < A.n && (mask == nullptr || mask[e] != 0);
  uint32_t dealt = 0;
  if (active) {
    Env s;
    const uint4 rnd = env_random(A.seed, A.global_env_offset + e, step, 1);
    dealt = deal_new_episode(s, global_row(A.history + static_cast<size_t>(e) * kHistoryWords), rnd,
                             forced ? forced + static_cast<size_t>(e) * 4 : nullptr);
    store_env(A.state + e, s);
    StepResult r;
    r.legal = legal_mask_decision(s);
    r.cur_player = static_cast<int>(g_mover(s.g));
    r.done = false; r.reward0 = 0; r.return0 = 0;
    write_outputs(A, e, r);
  }
  st.sum(COUP_STAT_CHANCE_MOVES, dealt);
  st.flush(A.stats);
}

// ---- step with caller-provided actions --------------------------------------------------------------
#ifndef COUP_STEP_BLOCKS
#define COUP_STEP_BLOCKS 5   // resident CTAs per SM: 61.5 / 57.4 / 55.5 / 57.5 us per 2^20 envs (with k_sample_uniform) at 3 / 4 / 5 / 6
#endif
__global__ void __launch_bounds__(kBlockThreads, COUP_STEP_BLOCKS)
k_step(EnvArrays A, const uint8_t* __restrict__ actions, const uint8_t* __restrict__ forced, uint64_t step) {
  __shared__ uint32_t s_stats[COUP_STATS_LEN];
  BlockStats st;
  st.init(s_stats);
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  // the env's history row is loaded with its state word, in one round trip, into shared memory (odd pitch: conflict-free);
  // the step updates it there and writes the changed words through to HBM
  __shared__ uint32_t s_row[kBlockThreads][kHistoryWords + 1];
  const uint32_t action = e < A.n ? actions[e] : 0xFFu;
  const bool active = action != 0xFFu;   // 0xFF: this env sits the step out, outputs keep their values
  Env s = {};
  uint32_t* const row = s_row[threadIdx.x];
  if (active) s = load_env_and_row(A, e, row);
  const StepResult r = step_env<false>(s, HistRow{row, A.history + static_cast<size_t>(e) * kHistoryWords}, action,
                                       forced ? forced + static_cast<size_t>(e) * 4 : nullptr, A, e, step, active);
  if (active) {
    store_env(A.state + e, s);
    write_outputs(A, e, r);
  }
  account(st, r, active);
  st.flush(A.stats);
}

// ---- single moves with explicit chance nodes (the OpenSpiel State surface: State::ApplyAction at any
// node, spiel.cc:322-332, without the rl_environment-style chance resolution of k_step) ----------------
__device__ __forceinline__ void write_outputs_any_node(const EnvArrays& A, uint32_t e, const Env& s) {
  StepResult r = {};
  const bool term = is_terminal(s);
  const bool chance = !term && g_chance(s.g);
  r.legal = term ? 0u : chance ? legal_mask_chance(s) : legal_mask_decision(s);           // coup.cc:824-938
  r.cur_player = term ? COUP_TERMINAL_PLAYER_ID : chance ? COUP_CHANCE_PLAYER_ID : static_cast<int>(g_mover(s.g));
  r.done = term;
  r.reward0 = c_reward0(s.c);
  r.return0 = returns_p0(s);
  write_outputs(A, e, r);
  if (chance) A.step_word[e] |= 1u << 27;
}

// mode 0: CoupState ctor only (env left at its first chance node); mode 1: apply one move per env
// (0xFF = leave untouched). An illegal move sets the sticky error bit and changes nothing else.
__global__ void __launch_bounds__(kBlockThreads)
k_single_move(EnvArrays A, const uint8_t* __restrict__ moves_or_mask, int mode) {
  __shared__ uint32_t s_stats[COUP_STATS_LEN];
  BlockStats st;
  st.init(s_stats);
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  bool illegal = false;
  if (e < A.n) {
    uint32_t* hist_row = A.history + static_cast<size_t>(e) * kHistoryWords;
    if (mode == 0) {
      if (moves_or_mask == nullptr || moves_or_mask[e] != 0) {
        const Env s = initial_state();
        store_env(A.state + e, s);
        write_outputs_any_node(A, e, s);
      }
    } else {
      const uint32_t mv = moves_or_mask[e];
      if (mv != 0xFFu) {
        Env s = load_env(A.state + e);
        const bool term = is_terminal(s);
        const bool chance = !term && g_chance(s.g);
        const uint32_t legal = term ? 0u : chance ? legal_mask_chance(s) : legal_mask_decision(s);
        if (mv < 18u && ((legal >> mv) & 1u)) {
          const uint32_t at = c_moves(s.c);
          uint32_t code = mv;
          if (chance) code = apply_chance(s, mv); else apply_player_action(s, mv);
          history_commit(hist_row, at, code, 1u);
        } else {
          s.g |= kBitError;
          illegal = true;
        }
        store_env(A.state + e, s);
        write_outputs_any_node(A, e, s);
      }
    }
  }
  st.count(COUP_STAT_ILLEGAL, illegal);
  st.flush(A.stats);
}

// One move on ONE env (mode as in k_single_move); *illegal_out is set to 1 when the move was rejected.
__global__ void k_single_move_one(EnvArrays A, uint32_t slot, uint32_t mv, int mode, uint32_t* illegal_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t* hist_row = A.history + static_cast<size_t>(slot) * kHistoryWords;
  Env s;
  if (mode == 0) {
    s = initial_state();
  } else {
    s = load_env(A.state + slot);
    const bool term = is_terminal(s);
    const bool chance = !term && g_chance(s.g);
    const uint32_t legal = term ? 0u : chance ? legal_mask_chance(s) : legal_mask_decision(s);
    if (mv < 18u && ((legal >> mv) & 1u)) {
      const uint32_t at = c_moves(s.c);
      uint32_t code = mv;
      if (chance) code = apply_chance(s, mv); else apply_player_action(s, mv);
      history_commit(hist_row, at, code, 1u);
      *illegal_out = 0;
    } else {
      *illegal_out = 1;  // nothing changes: the caller raises, as ApplyAction would (spiel_utils.cc:119-137)
      return;
    }
  }
  store_env(A.state + slot, s);
  write_outputs_any_node(A, slot, s);
}

// Copies env `src` onto env `dst` (State::Clone, coup.cc:1058-1060): state, history and outputs.
__global__ void k_copy_env(EnvArrays A, uint32_t src, uint32_t dst) {
  const int t = threadIdx.x;
  if (t < kHistoryWords) A.history[static_cast<size_t>(dst) * kHistoryWords + t] = A.history[static_cast<size_t>(src) * kHistoryWords + t];
  if (t == 0) {
    A.state[dst] = A.state[src];
    A.legal[dst] = A.legal[src];
    A.cur_player[dst] = A.cur_player[src];
    A.done[dst] = A.done[src];
    A.rewards[2 * dst] = A.rewards[2 * src]; A.rewards[2 * dst + 1] = A.rewards[2 * src + 1];
    A.returns[2 * dst] = A.returns[2 * src]; A.returns[2 * dst + 1] = A.returns[2 * src + 1];
    A.step_word[dst] = A.step_word[src];
  }
}

// ---- batched state.child(action): dst[i] = step(copy of src[parent[i]], action[i]) without auto-reset ------
// One thread per child: 16 B state + 64 B history row gathered from the parent slab (four 16 B loads), stepped in
// registers, written to the child's own row. D.flags arrives with COUP_FLAG_AUTO_RESET cleared.
#ifndef COUP_FORK_BLOCKS
#define COUP_FORK_BLOCKS 4   // resident CTAs per SM (64 registers): 59.5 -> 51.6 us per 2^20 children against 3
#endif
__global__ void __launch_bounds__(kBlockThreads, COUP_FORK_BLOCKS)
k_fork(EnvArrays D, const uint4* __restrict__ src_state, const uint32_t* __restrict__ src_history, uint32_t src_n,
       const uint32_t* __restrict__ parent, const uint8_t* __restrict__ actions, const uint8_t* __restrict__ forced,
       uint32_t count, uint64_t step, const uint32_t* __restrict__ count_ptr) {
  __shared__ uint32_t s_stats[COUP_STATS_LEN];
  // The child's history row is built in shared memory (odd pitch: conflict-free per-lane access) from the parent's row and
  // written to the child slab once, after the step: no read-modify-write of global memory inside the step.
  __shared__ uint32_t s_row[kBlockThreads][kHistoryWords + 1];
  BlockStats st;
  st.init(s_stats);
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = e < (count_ptr ? min(count, *count_ptr) : count);   // device-side child count of a traversal level
  const uint32_t p = active ? parent[e] : 0xFFFFFFFFu;
  const bool valid = active && p < src_n;                  // out-of-range parents set the child's error bit
  uint32_t* const row = s_row[threadIdx.x];
  Env s = initial_state();
  bool parent_terminal = false;
  if (valid) {
    const uint4* src_row = reinterpret_cast<const uint4*>(src_history + static_cast<size_t>(p) * kHistoryWords);
    const uint4 sv = src_state[p];
    const uint4 h0 = src_row[0], h1 = src_row[1], h2 = src_row[2], h3 = src_row[3];
    row[0] = h0.x; row[1] = h0.y; row[2] = h0.z; row[3] = h0.w; row[4] = h1.x; row[5] = h1.y; row[6] = h1.z; row[7] = h1.w;
    row[8] = h2.x; row[9] = h2.y; row[10] = h2.z; row[11] = h2.w; row[12] = h3.x; row[13] = h3.y; row[14] = h3.z; row[15] = h3.w;
    s.p[0] = sv.x; s.p[1] = sv.y; s.g = sv.z; s.c = sv.w;
    parent_terminal = is_terminal(s);
  }
  StepResult r = step_env<false>(s, global_row(row), valid ? actions[e] : 0xFFu, forced ? forced + static_cast<size_t>(e) * 4 : nullptr,
                                 D, e, step, valid);
  if (active) {
    if (parent_terminal) { s.g |= kBitError; r.illegal = true; }   // a terminal state has no children
    if (!valid) {
      s.g |= kBitError;
      r.illegal = true; r.done = false; r.legal = 0; r.cur_player = COUP_CHANCE_PLAYER_ID;
    }
    if (valid) {
      uint4* dst_row = reinterpret_cast<uint4*>(D.history + static_cast<size_t>(e) * kHistoryWords);
#pragma unroll
      for (int k = 0; k < kHistoryWords / 4; ++k) dst_row[k] = make_uint4(row[4 * k], row[4 * k + 1], row[4 * k + 2], row[4 * k + 3]);
    }
    store_env(D.state + e, s);
    write_outputs(D, e, r);
  }
  account(st, r, active);
  st.flush(D.stats);
}

}  // namespace coup
