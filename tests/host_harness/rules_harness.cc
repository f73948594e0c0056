// rules_harness.cc -- TEST TOOLING, not part of the product. Compiles open_spiel_coup_b200/csrc/coup_device.cuh for the
// host (COUP_RULES_HOST_TEST) so that the CPU test-suite (-m "not gpu") can diff the branch-free device rules -- event
// table, legal masks, chance sampling, closed-form initial deal, history packing -- against the oracle / the compiled
// reference without a GPU. libcoup_b200.so never defines COUP_RULES_HOST_TEST and contains no host copy of the rules.
#define COUP_RULES_HOST_TEST 1
#include "../../open_spiel_coup_b200/csrc/coup_device.cuh"

#include <cstring>

using namespace coup;

namespace {

// Same layout as oracle/bindings.py TRACE_DTYPE (48 bytes); the tensor hashes stay 0 here.
struct TraceRec {
  int8_t cur_player;
  uint8_t is_terminal, is_chance, move_number;
  uint32_t legal_mask;
  int8_t rewards[2], returns[2];
  uint8_t coins[2], ncards[2];
  uint64_t hash_info[2], hash_obs[2];
};
static_assert(sizeof(TraceRec) == 48, "TRACE_DTYPE layout");

void record(const Env& s, TraceRec* r) {
  std::memset(r, 0, sizeof(*r));
  const bool term = is_terminal(s);
  const bool chance = !term && g_chance(s.g);
  r->cur_player = term ? -4 : chance ? -1 : static_cast<int8_t>(g_mover(s.g));
  r->is_terminal = term;
  r->is_chance = chance;
  r->move_number = static_cast<uint8_t>(c_moves(s.c));
  r->legal_mask = term ? 0u : chance ? legal_mask_chance(s) : legal_mask_decision(s);
  const int rew = c_reward0(s.c), ret = returns_p0(s);
  r->rewards[0] = static_cast<int8_t>(rew); r->rewards[1] = static_cast<int8_t>(-rew);
  r->returns[0] = static_cast<int8_t>(ret); r->returns[1] = static_cast<int8_t>(-ret);
  for (int p = 0; p < 2; ++p) {
    r->coins[p] = static_cast<uint8_t>(pw_coins(s.p[p]));
    r->ncards[p] = static_cast<uint8_t>(hand_count(pw_hand(s.p[p])));
  }
}

// Applies one move (card id at a chance node, action id otherwise) if it is legal; logs it into `row`.
bool apply_move(Env& s, uint32_t mv, uint32_t* row) {
  const bool term = is_terminal(s);
  const bool chance = !term && g_chance(s.g);
  const uint32_t legal = term ? 0u : chance ? legal_mask_chance(s) : legal_mask_decision(s);
  if (mv >= 18u || !((legal >> mv) & 1u)) return false;
  const uint32_t at = c_moves(s.c);
  uint32_t code = mv;
  if (chance) code = apply_chance(s, mv); else apply_player_action(s, mv);
  history_commit(row, at, code, 1u);
  return true;
}

uint64_t splitmix(uint64_t& st) {
  uint64_t z = (st += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

uint32_t naive_kth(uint32_t mask, uint32_t k) {
  for (uint32_t b = 0; b < 32; ++b)
    if ((mask >> b) & 1u) { if (k == 0) return b; --k; }
  return 32;
}

uint32_t naive_sample_card(uint32_t g, uint32_t u) {
  uint32_t d[5], total = 0;
  for (int c = 0; c < 5; ++c) { d[c] = (g >> (4 * c)) & 15u; total += d[c]; }
  const uint32_t r = static_cast<uint32_t>((static_cast<uint64_t>(u) * total) >> 32);
  uint32_t run = 0;
  for (uint32_t c = 0; c < 5; ++c) { run += d[c]; if (r < run) return c; }
  return 4;
}

}  // namespace

extern "C" {

// Records the initial state and the state after every move; returns n or -(index+1) of the first rejected move.
// *history_ok is cleared if the packed history row does not decode back to the moves made.
int hh_trace(const uint8_t* actions, int n, TraceRec* out, int* history_ok) {
  Env s = initial_state();
  uint32_t row[kHistoryWords];
  std::memset(row, 0xAB, sizeof(row));   // stale garbage: only words that receive a move may be relied upon
  record(s, &out[0]);
  for (int i = 0; i < n; ++i) {
    // whose deal this is, for the expected history code
    const bool chance = !is_terminal(s) && g_chance(s.g);
    const uint32_t qn = g_qn(s.g);
    const uint32_t target = (s.g & kBitQInitial) ? (qn & 1u) : ((s.g >> 28) & 1u);
    if (!apply_move(s, actions[i], row)) return -(i + 1);
    const uint32_t want = chance ? 18u + 5u * target + actions[i] : actions[i];
    const uint32_t got = (row[i / 6] >> (5 * (i % 6))) & 31u;
    if (got != want && history_ok) *history_ok = 0;
    record(s, &out[i + 1]);
  }
  return n;
}

int hh_trace_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, TraceRec* out, int* history_ok) {
  int bad = 0;
  *history_ok = 1;
  for (int t = 0; t < n_traj; ++t) {
    const int n = static_cast<int>(offsets[t + 1] - offsets[t]);
    if (hh_trace(actions + offsets[t], n, out + offsets[t] + t, history_ok) < 0) ++bad;
  }
  return bad;
}

// Plays `n_games` games with the harness's own rules: uniform-random legal moves (chance: generic sample_card), or --
// steer != 0 -- preferring Pass / Exchange / ExchangeReturn to reach the 91-move cap. Moves are appended to actions_out,
// offsets_out[g] .. offsets_out[g+1]. Returns the number of moves, or -1 if `cap` is too small.
long hh_random_games(int n_games, uint64_t seed, int steer, uint8_t* actions_out, int64_t* offsets_out, long cap) {
  uint64_t st = seed * 0x9E3779B97F4A7C15ULL + 12345;
  long pos = 0;
  offsets_out[0] = 0;
  for (int gidx = 0; gidx < n_games; ++gidx) {
    Env s = initial_state();
    uint32_t row[kHistoryWords] = {0};
    while (!is_terminal(s)) {
      if (pos >= cap) return -1;
      uint32_t mv;
      if (g_chance(s.g)) {
        mv = sample_card(s, static_cast<uint32_t>(splitmix(st) >> 32));
      } else {
        const uint32_t legal = legal_mask_decision(s);
        mv = sample_action(legal, static_cast<uint32_t>(splitmix(st) >> 32));
        if (steer) {
          const uint32_t prefer[5] = {15, 16, 17, 5, 9};
          for (uint32_t a : prefer) if ((legal >> a) & 1u) mv = a;
        }
      }
      if (!apply_move(s, mv, row)) return -2;
      actions_out[pos++] = static_cast<uint8_t>(mv);
    }
    offsets_out[gidx + 1] = pos;
  }
  return pos;
}

// kth_set_bit against a bit loop, for every mask of `bits` bits and every valid k. Returns mismatches.
long hh_check_kth_set_bit(int bits) {
  long bad = 0;
  for (uint32_t mask = 1; mask < (1u << bits); ++mask)
    for (uint32_t k = 0; k < popc32(mask) && k <= 7; ++k)
      if (kth_set_bit(mask, k) != naive_kth(mask, k)) ++bad;
  return bad;
}

// hand_insert (all four slots compared at once) against the plain definition, over every sorted hand with a free slot.
long hh_check_hand_insert(void) {
  long bad = 0;
  for (uint32_t h = 0; h < 0x10000u; ++h) {
    uint32_t n[4];
    bool ok = true;
    for (int i = 0; i < 4; ++i) { n[i] = (h >> (4 * i)) & 15u; if (n[i] != 15u && n[i] > 9u) ok = false; }
    for (int i = 0; i < 3; ++i) if (n[i] > n[i + 1]) ok = false;
    if (!ok || n[3] != 15u) continue;
    for (uint32_t key = 0; key <= 9; ++key) {
      uint32_t want[5], m = 0;
      bool placed = false;
      for (int i = 0; i < 4; ++i) {
        if (!placed && n[i] > key) { want[m++] = key; placed = true; }
        want[m++] = n[i];
      }
      const uint32_t w = want[0] | (want[1] << 4) | (want[2] << 8) | (want[3] << 12);
      if (hand_insert(h, key) != w) ++bad;
    }
  }
  return bad;
}

// sample_card (packed running counts) against the plain loop, on random decks whose counts sum to 1..15.
long hh_check_sample_card(long trials, uint64_t seed) {
  long bad = 0;
  uint64_t st = seed;
  for (long i = 0; i < trials; ++i) {
    uint32_t g = 0, total = 0;
    for (int c = 0; c < 5; ++c) {
      uint32_t d = static_cast<uint32_t>(splitmix(st) % 8);
      if (total + d > 15) d = 15 - total;
      total += d;
      g |= d << (4 * c);
    }
    if (total == 0) { g |= 1u << 8; }
    g |= static_cast<uint32_t>(splitmix(st)) & 0xFFF00000u;      // the other fields of the word must not matter
    Env s = initial_state();
    s.g = g;
    const uint32_t u = static_cast<uint32_t>(splitmix(st) >> 32);
    const uint32_t edge = (i & 7) == 0 ? 0xFFFFFFFFu : (i & 7) == 1 ? 0u : u;
    if (sample_card(s, edge) != naive_sample_card(g, edge)) ++bad;
  }
  return bad;
}

// dealt_initial_state (closed form) against initial_state() + four generic sample_card / apply_chance rounds.
long hh_check_closed_form_deal(long trials, uint64_t seed) {
  long bad = 0;
  uint64_t st = seed;
  for (long i = 0; i < trials; ++i) {
    uint4 r = make_uint4(static_cast<uint32_t>(splitmix(st)), static_cast<uint32_t>(splitmix(st)),
                         static_cast<uint32_t>(splitmix(st)), static_cast<uint32_t>(splitmix(st)));
    if ((i & 15) == 0) r.x = 0xFFFFFFFFu;
    if ((i & 15) == 1) r.y = 0xFFFFFFFFu;
    if ((i & 15) == 2) r.z = r.w = 0xFFFFFFFFu;
    if ((i & 15) == 3) r = make_uint4(0, 0, 0, 0);
    uint32_t codes_a = 0;
    const Env a = dealt_initial_state(r, codes_a);
    Env b = initial_state();
    uint32_t codes_b = 0, n = 0;
    resolve_chance(b, r, 0, nullptr, codes_b, n);
    if (n != 4 || a.p[0] != b.p[0] || a.p[1] != b.p[1] || a.g != b.g || a.c != b.c || codes_a != codes_b) ++bad;
  }
  return bad;
}

// history_commit with several codes at once (what a step does) against one code at a time.
long hh_check_history_commit(long trials, uint64_t seed) {
  long bad = 0;
  uint64_t st = seed;
  for (long i = 0; i < trials; ++i) {
    uint32_t a[kHistoryWords], b[kHistoryWords];
    for (int w = 0; w < kHistoryWords; ++w) a[w] = b[w] = static_cast<uint32_t>(splitmix(st));
    const uint32_t first = static_cast<uint32_t>(splitmix(st) % 88);
    const uint32_t n = 1 + static_cast<uint32_t>(splitmix(st) % 5);
    // both rows hold `first` valid moves: words below are (valid | zeros) as the writer guarantees
    for (uint32_t m = 0; m < first; ++m) { /* keep random codes */ }
    if (first % 6) { const uint32_t w = first / 6; a[w] &= (1u << (5 * (first % 6))) - 1u; b[w] = a[w]; }
    uint32_t codes = 0, c[5];
    for (uint32_t k = 0; k < n; ++k) { c[k] = static_cast<uint32_t>(splitmix(st) % 28); codes |= c[k] << (5 * k); }
    history_commit(a, first, codes, n);
    for (uint32_t k = 0; k < n; ++k) history_commit(b, first + k, c[k], 1u);
    for (uint32_t m = 0; m < first + n; ++m)
      if (((a[m / 6] >> (5 * (m % 6))) & 31u) != ((b[m / 6] >> (5 * (m % 6))) & 31u)) { ++bad; break; }
  }
  return bad;
}

}  // extern "C"
