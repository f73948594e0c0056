// coup_kernels.cuh -- sm_100a kernels of the batched Coup environment.
//
// Thread mapping: one thread per environment for the rules (16-byte state load/store, coalesced),
// one warp per 32 consecutive environments for the tensor encoders (the 32 rows of a warp are one
// contiguous span of the output, written with full-width vector stores). No tensor cores: there is no
// contraction anywhere on this path; every kernel is bounded by HBM traffic or instruction issue.
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

// ---- one level of a sampled CFR traversal (python/algorithms/deep_cfr.py:415-525), thread per node ------------------
// From the advantage-network outputs of the player to move: regret matching (positive parts over the legal actions,
// normalised; if none is positive, probability one on the legal action with the largest raw advantage, :499-525),
// then which children to expand: at the traverser's nodes every legal action (external sampling, :438-441) or
// min(n_legal, k) actions drawn without replacement from expl * uniform + (1 - expl) * strategy (outcome sampling,
// :442-466; k = outcome_factor, or per node outcome_factor with probability e_outcome and 1 otherwise); at the
// opponent's nodes one action drawn from the strategy (:482-487). Sampling without replacement is the Gumbel-top-k
// order of the log-probabilities, i.e. the sequential renormalised draw of np.random.choice(replace=False).
__device__ __forceinline__ float u01(uint32_t r) { return (static_cast<float>(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }
__device__ __forceinline__ uint32_t cfr_expand_node(const float* __restrict__ adv_row, uint32_t word, uint32_t i, int traverser,
                                                    int external, uint32_t outcome_factor, float e_outcome, float expl,
                                                    uint64_t seed, uint64_t counter, float* __restrict__ strategy_row);

__global__ void __launch_bounds__(kBlockThreads)
k_cfr_expand(const float* __restrict__ advantages, const uint32_t* __restrict__ step_words, uint32_t count,
             int traverser, int external, uint32_t outcome_factor, float e_outcome, float expl, uint64_t seed,
             uint64_t counter, float* __restrict__ strategy_out, uint32_t* __restrict__ expand_out,
             uint32_t* __restrict__ count_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t expand = cfr_expand_node(advantages + static_cast<size_t>(i) * 18, step_words[i], i, traverser, external,
                                          outcome_factor, e_outcome, expl, seed, counter, strategy_out + static_cast<size_t>(i) * 18);
  expand_out[i] = expand;
  count_out[i] = __popc(expand);
}

// Children of a level in parent order: child j of node i (j-th set bit of expand[i]) lands at offsets[i] + j, where
// offsets is the exclusive prefix sum of the child counts.
__global__ void __launch_bounds__(kBlockThreads)
k_cfr_children(const uint32_t* __restrict__ expand, const int64_t* __restrict__ offsets, uint32_t count,
               uint32_t* __restrict__ parent_out, uint8_t* __restrict__ action_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint32_t bits = expand[i];
  int64_t pos = offsets[i];
  while (bits) {
    const int a = __ffs(bits) - 1;
    bits &= bits - 1;
    parent_out[pos] = i;
    action_out[pos] = static_cast<uint8_t>(a);
    ++pos;
  }
}

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

// ---- a whole level of a sampled CFR traversal with DEVICE-side node counts ------------------------------------------------
// The level-by-level expansion above, without the host in the loop: the number of nodes of a level lives in device memory
// (levels shrink and grow with the sampling), every kernel is launched for the capacity of the level buffers and works on
// the first *count nodes, and one single-CTA kernel per level does regret matching, child selection, the prefix sum of the
// child counts and the (parent, action) lists, and leaves the next level's count. The host only checks "is the frontier
// empty" every few levels.
constexpr int kCfrLevelThreads = 1024;

// Regret matching + child selection of ONE node (the body of k_cfr_expand as a function).
__device__ __forceinline__ uint32_t cfr_expand_node(const float* __restrict__ adv_row, uint32_t word, uint32_t i, int traverser,
                                                    int external, uint32_t outcome_factor, float e_outcome, float expl,
                                                    uint64_t seed, uint64_t counter, float* __restrict__ strategy_row) {
  const uint32_t legal = word & 0x3FFFFu;
  const int player = (word >> 18) & 1u;
  const int n_legal = __popc(legal);
  float adv[18];
#pragma unroll
  for (int a = 0; a < 18; ++a) adv[a] = adv_row[a];
  float total = 0.f, best = -INFINITY;
  int best_a = 0;
#pragma unroll
  for (int a = 0; a < 18; ++a) {
    if ((legal >> a) & 1u) {
      total += fmaxf(adv[a], 0.f);
      if (adv[a] > best) { best = adv[a]; best_a = a; }
    }
  }
  float strat[18];
#pragma unroll
  for (int a = 0; a < 18; ++a) {
    const bool ok = (legal >> a) & 1u;
    strat[a] = !ok ? 0.f : total > 0.f ? fmaxf(adv[a], 0.f) / total : (a == best_a ? 1.f : 0.f);
    strategy_row[a] = strat[a];
  }
  uint32_t expand = 0;
  if (n_legal > 0) {
    const uint4 r0 = env_random(seed, i, counter, 2), r1 = env_random(seed, i, counter, 3), r2 = env_random(seed, i, counter, 4);
    if (player != traverser) {
      float sum = 0.f;
#pragma unroll
      for (int a = 0; a < 18; ++a) sum += strat[a];
      const float target = u01(r0.x) * sum;
      float acc = 0.f;
      int pick = best_a;
#pragma unroll
      for (int a = 17; a >= 0; --a) if (strat[a] > 0.f) pick = a;           // fall-back: first action with mass
      bool done = false;
#pragma unroll
      for (int a = 0; a < 18; ++a) {
        if (!done && strat[a] > 0.f) { acc += strat[a]; pick = a; if (target < acc) done = true; }
      }
      expand = 1u << pick;
    } else if (external) {
      expand = legal;
    } else {
      uint32_t k = outcome_factor;
      if (e_outcome >= 0.f) k = u01(r0.y) < e_outcome ? outcome_factor : 1u;
      k = min(k, static_cast<uint32_t>(n_legal));
      float key[18];
      int slot = 0;                                                          // legal actions draw r0.z, r0.w, r1.*, r2.* in order
#pragma unroll
      for (int a = 0; a < 18; ++a) {
        key[a] = -INFINITY;
        if ((legal >> a) & 1u) {
          const uint32_t r = slot == 0 ? r0.z : slot == 1 ? r0.w : slot == 2 ? r1.x : slot == 3 ? r1.y : slot == 4 ? r1.z
                           : slot == 5 ? r1.w : slot == 6 ? r2.x : slot == 7 ? r2.y : slot == 8 ? r2.z : r2.w;
          ++slot;
          const float p = expl / n_legal + (1.f - expl) * strat[a];
          if (p > 0.f) key[a] = logf(p) - logf(-logf(u01(r)));
        }
      }
      for (uint32_t t = 0; t < k; ++t) {
        int arg = -1;
        float m = -INFINITY;
#pragma unroll
        for (int a = 0; a < 18; ++a) if (!((expand >> a) & 1u) && key[a] > m) { m = key[a]; arg = a; }
        if (arg < 0) break;
        expand |= 1u << arg;
      }
    }
  }
  return expand;
}

// One CTA. Nodes [0, *count) of the level: terminal nodes (bit 19 of the step word) expand nothing. Writes, per node,
// strategy [18], expand mask and the exclusive prefix `offset` of its children; per child (parent order, ascending action)
// parent index and action; *next_count = number of children, clipped to `capacity` (then *overflow is set: the level is
// truncated -- the caller sized the buffers too small).
__global__ void __launch_bounds__(kCfrLevelThreads)
k_cfr_level(const float* __restrict__ advantages, const uint32_t* __restrict__ step_words, const uint32_t* __restrict__ count_ptr,
            uint32_t capacity, int traverser, int external, uint32_t outcome_factor, float e_outcome, float expl, uint64_t seed,
            uint64_t counter, float* __restrict__ strategy_out, uint32_t* __restrict__ expand_out, uint32_t* __restrict__ offset_out,
            uint32_t* __restrict__ parent_out, uint8_t* __restrict__ action_out, uint32_t* __restrict__ next_count,
            uint32_t* __restrict__ overflow) {
  __shared__ uint32_t s_warp[kCfrLevelThreads / 32];
  __shared__ uint32_t s_base;
  const uint32_t count = min(*count_ptr, capacity);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (uint32_t start = 0; start < count; start += kCfrLevelThreads) {     // uniform trip count
    const uint32_t i = start + threadIdx.x;
    uint32_t expand = 0;
    if (i < count) {
      const uint32_t word = step_words[i];
      if (((word >> 19) & 1u) == 0)
        expand = cfr_expand_node(advantages + static_cast<size_t>(i) * 18, word, i, traverser, external, outcome_factor,
                                 e_outcome, expl, seed, counter, strategy_out + static_cast<size_t>(i) * 18);
      expand_out[i] = expand;
    }
    // block-wide exclusive scan of the child counts of this chunk
    const uint32_t c = __popc(expand);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      s_warp[lane] = w;                      // inclusive over warps
    }
    __syncthreads();
    const uint32_t base = s_base + (warp ? s_warp[warp - 1] : 0u) + incl - c;
    const uint32_t chunk_total = s_warp[kCfrLevelThreads / 32 - 1];
    if (i < count) {
      offset_out[i] = base;
      uint32_t bits = expand, pos = base;
      while (bits) {
        const int a = __ffs(bits) - 1;
        bits &= bits - 1;
        if (pos < capacity) { parent_out[pos] = i; action_out[pos] = static_cast<uint8_t>(a); }
        ++pos;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const uint32_t total = s_base;
    *next_count = min(total, capacity);
    if (total > capacity) *overflow = 1u;
  }
}

// Backward pass of one level (deep_cfr.py:468-480, 492-497), thread per node: a terminal node's value is the traverser's
// return; an opponent node's value is its sampled child's; a traverser node's is cfv = sum_a strategy[a] * payoff[a] over
// its expanded children (unsampled actions count as payoff 0, as in the reference), and its sampled regrets are
// payoff[a] - cfv on the legal actions. `child_value` are the values of the next level (node offset[i] + j = j-th child).
__global__ void __launch_bounds__(kBlockThreads)
k_cfr_backward(const uint32_t* __restrict__ step_words, const uint32_t* __restrict__ count_ptr, uint32_t capacity, int traverser,
               const float* __restrict__ strategy, const uint32_t* __restrict__ expand, const uint32_t* __restrict__ offset,
               const double* __restrict__ child_value, double* __restrict__ value_out, float* __restrict__ regret_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= min(*count_ptr, capacity)) return;
  const uint32_t word = step_words[i];
  const double sign = traverser == 0 ? 1.0 : -1.0;
  if ((word >> 19) & 1u) {                                     // terminal: Returns()[traverser]
    value_out[i] = sign * (static_cast<int>((word >> 24) & 7u) - 2);
    return;
  }
  const uint32_t legal = word & 0x3FFFFu;
  const bool is_trav = static_cast<int>((word >> 18) & 1u) == traverser;
  double payoff[18];
#pragma unroll
  for (int a = 0; a < 18; ++a) payoff[a] = 0.0;
  uint32_t bits = expand[i], pos = offset[i];
  double sum = 0.0;
  while (bits) {
    const int a = __ffs(bits) - 1;
    bits &= bits - 1;
    const double v = pos < capacity ? child_value[pos] : 0.0;
#pragma unroll
    for (int b = 0; b < 18; ++b) if (b == a) payoff[b] = v;
    sum += v;
    ++pos;
  }
  if (!is_trav) { value_out[i] = sum; return; }
  double cfv = 0.0;
#pragma unroll
  for (int a = 0; a < 18; ++a) if ((legal >> a) & 1u) cfv += static_cast<double>(strategy[static_cast<size_t>(i) * 18 + a]) * payoff[a];
  value_out[i] = cfv;
#pragma unroll
  for (int a = 0; a < 18; ++a)
    regret_out[static_cast<size_t>(i) * 18 + a] = ((legal >> a) & 1u) ? static_cast<float>(payoff[a] - cfv) : 0.f;
}

// The nodes [0, *count) of a slab as packed records (history, state, meta: node index, seat<<31 | step word bits 0-26).
__global__ void __launch_bounds__(kBlockThreads)
k_pack_records(EnvArrays A, const uint32_t* __restrict__ count_ptr, uint32_t* __restrict__ records) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= min(*count_ptr, A.n)) return;
  uint4* dst = reinterpret_cast<uint4*>(records + static_cast<size_t>(e) * kRecordWords);
  const uint4* h4 = reinterpret_cast<const uint4*>(A.history + static_cast<size_t>(e) * kHistoryWords);
#pragma unroll
  for (int k = 0; k < 4; ++k) dst[k] = h4[k];
  dst[4] = A.state[e];
  const uint32_t word = A.step_word[e];
  dst[5] = make_uint4(e, (((word >> 18) & 1u) << 31) | (word & 0x7FFFFFFu), 0u, 0u);
}

// ---- uniform-random legal action (same draw the fused rollout would use at this step counter) ------
__global__ void __launch_bounds__(kBlockThreads)
k_sample_uniform(EnvArrays A, uint8_t* __restrict__ actions_out, uint64_t step) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n) return;
  const Env s = load_env(A.state + e);
  uint32_t a = 0xFFu;
  if (!is_terminal(s)) {
    const uint4 rnd = env_random(A.seed, A.global_env_offset + e, step, 0);
    a = sample_action(legal_mask_decision(s), rnd.x);
  }
  actions_out[e] = static_cast<uint8_t>(a);
}

// ---- masked policy sampling: the acting rule of the reference's agents (python/algorithms/nfsp.py:154-167)
// fused on the device: probs = softmax(logits); illegal -> 0; renormalise; action ~ probs. One thread per
// env; the draw is the x word of the step's Philox block (the slot coup_vec_sample_uniform uses).
template <typename T> __device__ __forceinline__ float logit_to_float(T v);
template <> __device__ __forceinline__ float logit_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float logit_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// Rows of 18 elements are 72 (36) bytes apart: a lane reading its own row touches a different cache line than its
// neighbour for every element. The warp therefore moves its 32 rows -- one contiguous 2 304-byte span -- with coalesced
// loads/stores through shared memory (row pitch 19 words: conflict-free) and each lane works on its row there.
constexpr int kRowPitch = kNumActions + 1;

#ifndef COUP_POLICY_BLOCKS
#define COUP_POLICY_BLOCKS 5   // resident CTAs per SM: 56.7 -> 47.2 us per 2^20 envs with probabilities, 37.7 -> 29.7 without
#endif
template <typename T>
__global__ void __launch_bounds__(kBlockThreads, COUP_POLICY_BLOCKS)
k_sample_policy(EnvArrays A, const T* __restrict__ logits, float* __restrict__ probs_out,
                uint8_t* __restrict__ actions_out, uint64_t step) {
  __shared__ float s_rows[kWarpsPerBlock][32 * kRowPitch];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t e0 = (blockIdx.x * kWarpsPerBlock + warp) * 32u;
  if (e0 >= A.n) return;
  const uint32_t e = e0 + lane;
  const uint32_t span = min(32u, A.n - e0) * kNumActions;
  float* rows = s_rows[warp];
  const T* src = logits + static_cast<size_t>(e0) * kNumActions;
#pragma unroll
  for (uint32_t i = 0; i < kNumActions; ++i) {
    const uint32_t j = lane + 32u * i;
    if (j < span) {
      const uint32_t r = (j * 3641u) >> 16;                    // j / 18 for j < 576
      rows[r * kRowPitch + (j - r * kNumActions)] = logit_to_float<T>(src[j]);
    }
  }
  __syncwarp();
  const uint32_t legal = e < A.n ? A.legal[e] : 0u;
  float p[kNumActions];
  float mx = -INFINITY;
#pragma unroll
  for (int a = 0; a < kNumActions; ++a) {
    p[a] = rows[lane * kRowPitch + a];
    if ((legal >> a) & 1u) mx = fmaxf(mx, p[a]);
  }
  // softmax over all actions followed by masking and renormalising == softmax over the legal ones;
  // subtracting the legal maximum keeps it finite.
  float sum = 0.f;
#pragma unroll
  for (int a = 0; a < kNumActions; ++a) {
    p[a] = ((legal >> a) & 1u) ? expf(p[a] - mx) : 0.f;
    sum += p[a];
  }
  uint32_t action = 0xFFu;
  if (legal != 0) {
    const float inv = 1.f / sum;
    const uint4 rnd = env_random(A.seed, A.global_env_offset + e, step, 0);
    const float u = static_cast<float>(rnd.x >> 8) * (1.0f / 16777216.0f);  // 24-bit uniform in [0,1)
    float cdf = 0.f;
    action = 31u - __clz(legal);  // falls back to the last legal action if rounding leaves u >= cdf
    bool found = false;
#pragma unroll
    for (int a = 0; a < kNumActions; ++a) {
      p[a] *= inv;
      cdf += p[a];
      if (!found && ((legal >> a) & 1u) && u < cdf) { action = a; found = true; }
    }
  }
  if (e < A.n) actions_out[e] = static_cast<uint8_t>(action);
  if (probs_out != nullptr) {
    __syncwarp();
#pragma unroll
    for (int a = 0; a < kNumActions; ++a) rows[lane * kRowPitch + a] = legal ? p[a] : 0.f;
    __syncwarp();
    float* dst = probs_out + static_cast<size_t>(e0) * kNumActions;
#pragma unroll
    for (uint32_t i = 0; i < kNumActions; ++i) {
      const uint32_t j = lane + 32u * i;
      if (j < span) {
        const uint32_t r = (j * 3641u) >> 16;
        dst[j] = rows[r * kRowPitch + (j - r * kNumActions)];
      }
    }
  }
}

// ---- dense legal mask: uint8[n][18] (State::LegalActionsMask, spiel.cc:371-377) ----------------------
// A warp expands the masks of 32 envs into one contiguous 576-byte span: 36 sixteen-byte stores, each byte's mask fetched
// from the lane that holds it.
__global__ void __launch_bounds__(kBlockThreads)
k_legal_actions_mask(const uint32_t* __restrict__ legal, uint8_t* __restrict__ out, uint32_t n) {
  const int lane = threadIdx.x & 31;
  const uint32_t e0 = (blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * 32u;
  if (e0 >= n) return;
  const uint32_t mine = e0 + lane < n ? legal[e0 + lane] : 0u;
  const uint32_t span = min(32u, n - e0) * kNumActions;                 // bytes of this warp
  uint8_t* dst = out + static_cast<size_t>(e0) * kNumActions;
  const bool vector_ok = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && span == 32u * kNumActions;
#pragma unroll
  for (uint32_t i = 0; i < 2; ++i) {
    const uint32_t unit = lane + 32u * i;                               // 16-byte unit of the span (36 of them)
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (uint32_t b = 0; b < 16; ++b) {
      const uint32_t j = min(unit * 16u + b, 32u * kNumActions - 1u);
      const uint32_t r = (j * 3641u) >> 16;
      const uint32_t bit = (__shfl_sync(0xffffffffu, mine, static_cast<int>(r)) >> (j - r * kNumActions)) & 1u;
      w[b >> 2] |= bit << (8u * (b & 3u));
    }
    if (unit < 36u) {
      if (vector_ok) {
        reinterpret_cast<uint4*>(dst)[unit] = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        for (uint32_t b = 0; b < 16; ++b)
          if (unit * 16u + b < span) dst[unit * 16u + b] = static_cast<uint8_t>((w[b >> 2] >> (8u * (b & 3u))) & 1u);
      }
    }
  }
}

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
// Same staging idea as the info-state encoder, one warp per 32 consecutive envs: a row is 98 elements with ~14
// non-zeros, and although one row is not a multiple of 16 bytes, the 32 (x2 views) rows of a warp are one contiguous,
// 16-byte aligned span of the output (3 136 / 6 272 / 12 544 B per view for u8 / bf16 / f32). The warp keeps that span
// zeroed in shared memory; each lane pokes the non-zeros of its own env's row(s), lane 0 hands the span to the TMA
// engine with one bulk store, and the lanes erase what they poked. Persistent: warps stride over the 32-env groups.
// A ragged last group, or an output that is not 16-byte aligned, is copied out of the staging span element by element.
constexpr int kObsWarps = 4;
constexpr int kObsThreads = kObsWarps * 32;

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
template <typename T, typename Src>
__global__ void __launch_bounds__(kObsThreads)
k_encode_obs(Src src, int player_sel, T* __restrict__ out, int use_bulk, uint32_t row_len, uint32_t* __restrict__ ids_out,
             uint32_t* __restrict__ count_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t vis = static_cast<uint32_t>(player_sel) >> 8;
  const bool pub = (vis & kVisNoPublic) == 0;
  const bool both = (player_sel & 7) == COUP_PLAYER_BOTH;
  const uint32_t views = both ? 2u : 1u;
  const uint32_t n = src.rows();
  if (count_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *count_out = n;
  const uint32_t span_elems = 32u * views * row_len;
  T* stage = reinterpret_cast<T*>(smem_raw) + static_cast<size_t>(warp) * span_elems;
  for (uint32_t q = lane; q < span_elems * sizeof(T) / 16u; q += 32u) reinterpret_cast<uint4*>(stage)[q] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
  const uint32_t n_groups = (n + 31u) / 32u;
  // The state word of the NEXT group is requested before this group's bulk store is waited for, so the load's round trip
  // (microseconds next to a saturated store stream) overlaps the engine's read of the buffer and the erase.
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
      poke_obs_row<T>(row, head_a, la, coins, true, pub);
      if (both) poke_obs_row<T>(row + row_len, head_b, la, coins, true, pub);
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
    if (bulk && lane == 0) tma_wait_read_all();
    __syncwarp();
    if (e < n) {
      T* row = stage + static_cast<size_t>(lane) * views * row_len;
      poke_obs_row<T>(row, head_a, la, coins, false, pub);
      if (both) poke_obs_row<T>(row + row_len, head_b, la, coins, false, pub);
    }
  }
  if (lane == 0) tma_wait_all();   // the engine must be done with this warp's shared memory before the CTA exits
}

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
