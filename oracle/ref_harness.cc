// ORACLE TEST TOOLING ONLY -- never linked into, imported by, or shipped with the product library.
//
// extern "C" harness over the UNMODIFIED reference implementation of the Coup path
// (/root/reference/open_spiel/games/coup.{h,cc} + the OpenSpiel core it needs), compiled where the
// sources lie by oracle/Makefile into oracle/_ref/libcoup_ref.so (git-ignored). It makes exactly the
// calls pyspiel forwards one-to-one (python/pybind11/pyspiel.cc:263-347): NewInitialState,
// ApplyAction, CurrentPlayer, IsTerminal, LegalActions, ChanceOutcomes, Rewards, Returns,
// InformationStateTensor, ObservationTensor, the string observers, Serialize.
//
// Used by: tests/ (to validate the C restatement oracle/coup_oracle.c and the CUDA path),
// oracle/gen_golden.py (fixture generation), bench.py's cpu_baseline / --impl reference arm
// (ref_bench below follows open_spiel/examples/benchmark_game.cc:32-140).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "open_spiel/games/coup.h"
#include "open_spiel/spiel.h"
#include "open_spiel/spiel_utils.h"

namespace {

using open_spiel::Action;
using open_spiel::Game;
using open_spiel::State;

struct RefError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// The reference aborts the process on SpielFatalError (spiel_utils.cc:119-137); like pyspiel
// (pyspiel.cc:620-626) we swap the handler for one that throws so illegal moves become error codes.
void ThrowingHandler(const std::string& msg) { throw RefError(msg); }

std::shared_ptr<const Game> TheGame() {
  static std::shared_ptr<const Game> game = [] {
    open_spiel::SetErrorHandler(ThrowingHandler);
    return open_spiel::LoadGame("coup");
  }();
  return game;
}

thread_local std::string g_last_error;

inline State* S(void* h) { return static_cast<State*>(h); }
inline const State* S(const void* h) { return static_cast<const State*>(h); }

// 64-bit position-keyed additive hash of a float tensor (only non-zero entries contribute, so it can
// be evaluated from sparse or dense form, in any order). Restated identically in
// oracle/coup_oracle.c (oc_tensor_hash) and in the CUDA verification kernel.
inline uint64_t Mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}
uint64_t TensorHash(const float* t, int n) {
  uint64_t h = 0;
  for (int i = 0; i < n; ++i) {
    uint32_t bits;
    std::memcpy(&bits, &t[i], 4);
    if (bits != 0) h += Mix64((static_cast<uint64_t>(i) << 32) | bits);
  }
  return h;
}

}  // namespace

extern "C" {

// One record per visited state of a trajectory (prefix of the action list).
struct RefTraceRec {
  int8_t cur_player;    // 0/1, -1 chance, -4 terminal (spiel_globals.h:26-36)
  uint8_t is_terminal;
  uint8_t is_chance;
  uint8_t move_number;
  uint32_t legal_mask;  // bit a set iff a in LegalActions() (card ids at chance nodes)
  int8_t rewards[2];
  int8_t returns[2];
  uint8_t coins[2];
  uint8_t ncards[2];
  uint64_t hash_info[2];
  uint64_t hash_obs[2];
};

const char* ref_last_error() { return g_last_error.c_str(); }

// [0] NumDistinctActions [1] MaxChanceOutcomes [2] NumPlayers [3] MinUtility [4] MaxUtility
// [5] InformationStateTensorSize [6] ObservationTensorSize [7] MaxGameLength
// [8] MaxChanceNodesInHistory [9] MaxMoveNumber [10] UtilitySum
void ref_game_constants(double* out) {
  auto g = TheGame();
  out[0] = g->NumDistinctActions();
  out[1] = g->MaxChanceOutcomes();
  out[2] = g->NumPlayers();
  out[3] = g->MinUtility();
  out[4] = g->MaxUtility();
  out[5] = g->InformationStateTensorSize();
  out[6] = g->ObservationTensorSize();
  out[7] = g->MaxGameLength();
  out[8] = g->MaxChanceNodesInHistory();
  out[9] = g->MaxMoveNumber();
  out[10] = g->UtilitySum().value_or(-999);
}

void* ref_new_state() { return TheGame()->NewInitialState().release(); }
void ref_free_state(void* h) { delete S(h); }
void* ref_clone_state(const void* h) { return S(h)->Clone().release(); }

int ref_apply_action(void* h, long action) {
  try {
    // Mirror pyspiel's legality behaviour: ApplyAction itself does not validate membership in
    // LegalActions() (spiel.cc:322-332), the game only SPIEL_CHECKs some preconditions.
    S(h)->ApplyAction(action);
    return 0;
  } catch (const RefError& e) {
    g_last_error = e.what();
    return -1;
  }
}

int ref_current_player(const void* h) { return S(h)->CurrentPlayer(); }
int ref_is_terminal(const void* h) { return S(h)->IsTerminal() ? 1 : 0; }
int ref_is_chance(const void* h) { return S(h)->IsChanceNode() ? 1 : 0; }
int ref_move_number(const void* h) { return S(h)->MoveNumber(); }

int ref_legal_actions(const void* h, long* out) {
  try {
    std::vector<Action> la = S(h)->LegalActions();
    for (size_t i = 0; i < la.size(); ++i) out[i] = la[i];
    return static_cast<int>(la.size());
  } catch (const RefError& e) {
    g_last_error = e.what();
    return -1;
  }
}

int ref_chance_outcomes(const void* h, long* actions, double* probs) {
  try {
    auto oc = S(h)->ChanceOutcomes();
    for (size_t i = 0; i < oc.size(); ++i) { actions[i] = oc[i].first; probs[i] = oc[i].second; }
    return static_cast<int>(oc.size());
  } catch (const RefError& e) {
    g_last_error = e.what();
    return -1;
  }
}

void ref_returns(const void* h, double* out) {
  auto r = S(h)->Returns();
  out[0] = r[0]; out[1] = r[1];
}
void ref_rewards(const void* h, double* out) {
  auto r = S(h)->Rewards();
  out[0] = r[0]; out[1] = r[1];
}

void ref_information_state_tensor(const void* h, int player, float* out, int length) {
  S(h)->InformationStateTensor(player, absl::MakeSpan(out, length));
}
void ref_observation_tensor(const void* h, int player, float* out, int length) {
  S(h)->ObservationTensor(player, absl::MakeSpan(out, length));
}

// Generic observer (coup.cc:1132-1141): public_info / perfect_recall / private_info
// (0 none, 1 single player, 2 all players). Returns the number of floats written.
int ref_observer_tensor(const void* h, int player, int public_info, int perfect_recall,
                        int private_info, float* out, int cap) {
  open_spiel::IIGObservationType t{public_info != 0, perfect_recall != 0,
                                   static_cast<open_spiel::PrivateInfoType>(private_info)};
  auto obs = TheGame()->MakeObserver(t, {});
  open_spiel::Observation o(*TheGame(), obs);
  o.SetFrom(*S(h), player);
  auto t_span = o.Tensor();
  int n = static_cast<int>(t_span.size());
  if (n > cap) return -n;
  std::memcpy(out, t_span.data(), n * sizeof(float));
  return n;
}

static int CopyString(const std::string& s, char* buf, int cap) {
  int n = static_cast<int>(s.size());
  if (n + 1 > cap) return -(n + 1);
  std::memcpy(buf, s.data(), n);
  buf[n] = 0;
  return n;
}
int ref_to_string(const void* h, char* buf, int cap) { return CopyString(S(h)->ToString(), buf, cap); }
int ref_information_state_string(const void* h, int player, char* buf, int cap) {
  return CopyString(S(h)->InformationStateString(player), buf, cap);
}
int ref_observation_string(const void* h, int player, char* buf, int cap) {
  return CopyString(S(h)->ObservationString(player), buf, cap);
}
int ref_observer_string(const void* h, int player, int public_info, int perfect_recall,
                        int private_info, char* buf, int cap) {
  open_spiel::IIGObservationType t{public_info != 0, perfect_recall != 0,
                                   static_cast<open_spiel::PrivateInfoType>(private_info)};
  auto obs = TheGame()->MakeObserver(t, {});
  return CopyString(obs->StringFrom(*S(h), player), buf, cap);
}
int ref_serialize(const void* h, char* buf, int cap) { return CopyString(S(h)->Serialize(), buf, cap); }
int ref_action_to_string(int player, long action, char* buf, int cap) {
  return CopyString(TheGame()->ActionToString(player, action), buf, cap);
}

int ref_history(const void* h, long* actions, int* players, int cap) {
  const auto& hist = S(h)->FullHistory();
  int n = static_cast<int>(hist.size());
  for (int i = 0; i < n && i < cap; ++i) { actions[i] = hist[i].action; players[i] = hist[i].player; }
  return n;
}

// Game::DeserializeState (spiel.cc:393-432): the reference parses a serialized state. nullptr on error.
void* ref_deserialize_state(const char* text) {
  try {
    return TheGame()->DeserializeState(text).release();
  } catch (const RefError& e) {
    g_last_error = e.what();
    return nullptr;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return nullptr;
  }
}

// CoupState convenience accessors (coup.h:139-143).
int ref_get_cards(const void* h, int player, int* values, int* states) {
  auto* cs = static_cast<const open_spiel::coup::CoupState*>(S(h));
  auto v = cs->GetCardsValue(player);
  auto s = cs->GetCardsState(player);
  for (size_t i = 0; i < v.size(); ++i) { values[i] = static_cast<int>(v[i]); states[i] = static_cast<int>(s[i]); }
  return static_cast<int>(v.size());
}
int ref_get_coins(const void* h, int player) {
  return static_cast<const open_spiel::coup::CoupState*>(S(h))->GetCoins(player);
}
int ref_get_last_action(const void* h, int player) {
  return static_cast<int>(static_cast<const open_spiel::coup::CoupState*>(S(h))->GetLastAction(player));
}

uint64_t ref_tensor_hash(const float* t, int n) { return TensorHash(t, n); }

static void FillRec(const State& st, RefTraceRec* r, std::vector<float>& info, std::vector<float>& obs) {
  auto* cs = static_cast<const open_spiel::coup::CoupState*>(&st);
  std::memset(r, 0, sizeof(*r));
  r->cur_player = static_cast<int8_t>(st.CurrentPlayer());
  r->is_terminal = st.IsTerminal();
  r->is_chance = st.IsChanceNode();
  r->move_number = static_cast<uint8_t>(st.MoveNumber());
  uint32_t m = 0;
  if (!r->is_terminal) for (Action a : st.LegalActions()) m |= 1u << a;
  r->legal_mask = m;
  auto rew = st.Rewards();
  auto ret = st.Returns();
  for (int p = 0; p < 2; ++p) {
    r->rewards[p] = static_cast<int8_t>(rew[p]);
    r->returns[p] = static_cast<int8_t>(ret[p]);
    r->coins[p] = static_cast<uint8_t>(cs->GetCoins(p));
    r->ncards[p] = static_cast<uint8_t>(cs->GetCardsValue(p).size());
    st.InformationStateTensor(p, absl::MakeSpan(info));
    st.ObservationTensor(p, absl::MakeSpan(obs));
    r->hash_info[p] = TensorHash(info.data(), static_cast<int>(info.size()));
    r->hash_obs[p] = TensorHash(obs.data(), static_cast<int>(obs.size()));
  }
}

// Replays one action list from the initial state; writes n_actions+1 records (state before any
// move, then after each move). Returns the number of records written, or -(i+1) if move i raised.
int ref_trace(const uint8_t* actions, int n_actions, RefTraceRec* out) {
  auto g = TheGame();
  std::vector<float> info(g->InformationStateTensorSize()), obs(g->ObservationTensorSize());
  auto st = g->NewInitialState();
  FillRec(*st, &out[0], info, obs);
  for (int i = 0; i < n_actions; ++i) {
    try {
      if (st->IsTerminal()) throw RefError("ApplyAction on terminal state");
      st->ApplyAction(actions[i]);
    } catch (const RefError& e) {
      g_last_error = e.what();
      return -(i + 1);
    }
    FillRec(*st, &out[i + 1], info, obs);
  }
  return n_actions + 1;
}

// Batched ref_trace over n_traj trajectories; actions are concatenated, trajectory t occupies
// [offsets[t], offsets[t+1]); its records go to out[offsets[t] + t ...]. Returns 0 or the number of
// trajectories that raised.
int ref_trace_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, RefTraceRec* out,
                    int threads) {
  TheGame();
  std::atomic<int> next{0}, bad{0};
  auto work = [&] {
    while (true) {
      int t = next.fetch_add(1);
      if (t >= n_traj) break;
      int n = static_cast<int>(offsets[t + 1] - offsets[t]);
      if (ref_trace(actions + offsets[t], n, out + offsets[t] + t) < 0) bad.fetch_add(1);
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < std::max(1, threads); ++i) th.emplace_back(work);
  for (auto& x : th) x.join();
  return bad.load();
}

// State reached by an action list (for full-tensor / string comparisons). nullptr on error.
void* ref_state_from_actions(const uint8_t* actions, int n_actions) {
  auto st = TheGame()->NewInitialState();
  try {
    for (int i = 0; i < n_actions; ++i) st->ApplyAction(actions[i]);
  } catch (const RefError& e) {
    g_last_error = e.what();
    return nullptr;
  }
  return st.release();
}

// ---- replay digests (BASELINE config 5: bit-exactness replay of GPU trajectories) -------------------
// For one trajectory (action ids incl. chance outcomes) walks the reference and, at every state a
// batched env REPORTS -- the first decision node after the initial deals, then the decision/terminal
// node reached after each player action once the following chance nodes are resolved -- folds
//   legal mask, current player (& 0xFF), is_terminal, rewards+2, returns+2, hash(info-state P0), hash(info-state P1),
//   hash(observation P0), hash(observation P1)
// into D = D * 0x9E3779B97F4A7C15 + v + 1 (mod 2^64). The GPU side folds the same fields in the same order
// (scripts/replay_check.py), so equal digests <=> identical legal sets, tensors, terminal flags, rewards, returns
// at every step of the trajectory. Returns the number of reported states, or -(i+1) if move i raised.
static inline void Fold(uint64_t& d, uint64_t v) { d = d * 0x9E3779B97F4A7C15ULL + v + 1; }

static void FoldState(const State& st, uint64_t& d, std::vector<float>& info, std::vector<float>& obs) {
  uint32_t m = 0;
  const bool term = st.IsTerminal();
  if (!term) for (Action a : st.LegalActions()) m |= 1u << a;
  Fold(d, m);
  Fold(d, static_cast<uint64_t>(st.CurrentPlayer()) & 0xFF);
  Fold(d, term ? 1 : 0);
  auto rew = st.Rewards();
  auto ret = st.Returns();
  Fold(d, static_cast<uint64_t>(static_cast<int>(rew[0]) + 2));
  Fold(d, static_cast<uint64_t>(static_cast<int>(rew[1]) + 2));
  Fold(d, static_cast<uint64_t>(static_cast<int>(ret[0]) + 2));
  Fold(d, static_cast<uint64_t>(static_cast<int>(ret[1]) + 2));
  for (int p = 0; p < 2; ++p) {
    st.InformationStateTensor(p, absl::MakeSpan(info));
    Fold(d, TensorHash(info.data(), static_cast<int>(info.size())));
  }
  for (int p = 0; p < 2; ++p) {
    st.ObservationTensor(p, absl::MakeSpan(obs));
    Fold(d, TensorHash(obs.data(), static_cast<int>(obs.size())));
  }
}

int ref_replay_digest(const uint8_t* actions, int n_actions, uint64_t* digest_out) {
  auto g = TheGame();
  std::vector<float> info(g->InformationStateTensorSize()), obs(g->ObservationTensorSize());
  auto st = g->NewInitialState();
  uint64_t d = 0;
  int reports = 0;
  for (int i = 0; i < n_actions; ++i) {
    try {
      if (st->IsTerminal()) throw RefError("ApplyAction on terminal state");
      st->ApplyAction(actions[i]);
    } catch (const RefError& e) {
      g_last_error = e.what();
      return -(i + 1);
    }
    if (!st->IsChanceNode()) {
      FoldState(*st, d, info, obs);
      ++reports;
    }
  }
  *digest_out = d;
  return reports;
}

// Batched over trajectories; digests[t], reports[t] (negative = rejected). Returns #rejected.
int ref_replay_digest_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, uint64_t* digests,
                            int32_t* reports, int threads) {
  TheGame();
  std::atomic<int> next{0}, bad{0};
  auto work = [&] {
    while (true) {
      int t0 = next.fetch_add(256);
      if (t0 >= n_traj) break;
      for (int t = t0; t < std::min(n_traj, t0 + 256); ++t) {
        reports[t] = ref_replay_digest(actions + offsets[t], static_cast<int>(offsets[t + 1] - offsets[t]), &digests[t]);
        if (reports[t] < 0) bad.fetch_add(1);
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < std::max(1, threads); ++i) th.emplace_back(work);
  for (auto& x : th) x.join();
  return bad.load();
}

// CPU baseline: the reference's own rollout benchmark protocol (examples/benchmark_game.cc:32-140):
// uniform-random legal actions, SampleAction on ChanceOutcomes(), one Game-independent State and one
// std::mt19937 per thread. mode 0: step + LegalActions only; 1: + InformationStateTensor(current
// player) at every decision node (= benchmark_game.cc); 2: + both players' info-state tensors
// (= rl_environment.Environment.get_time_step, python/rl_environment.py:219-268).
// out: [0] seconds [1] moves [2] decisions [3] chance moves [4] episodes [5] truncated
//      [6..10] P0 return histogram -2..+2 [11..18] legal-count histogram 0..7 [19] max moves/episode
//      [20] max coins
struct BenchAcc {
  long moves = 0, dec = 0, chance = 0, eps = 0, trunc = 0, ret[5] = {0, 0, 0, 0, 0}, legal[8] = {0};
  long max_moves = 0, max_coins = 0;
};
static void BenchThread(uint32_t seed, long episodes, int mode, BenchAcc* acc) {
  auto game = TheGame();
  std::mt19937 rng(seed);
  std::vector<float> info(game->InformationStateTensorSize());
  for (long e = 0; e < episodes; ++e) {
    auto state = game->NewInitialState();
    long moves = 0;
    while (!state->IsTerminal()) {
      if (state->IsChanceNode()) {
        auto outcomes = state->ChanceOutcomes();
        Action a = open_spiel::SampleAction(outcomes, std::uniform_real_distribution<double>(0, 1)(rng)).first;
        state->ApplyAction(a);
        acc->chance++;
      } else {
        int p = state->CurrentPlayer();
        if (mode >= 1) state->InformationStateTensor(p, absl::MakeSpan(info));
        if (mode >= 2) state->InformationStateTensor(1 - p, absl::MakeSpan(info));
        std::vector<Action> legal = state->LegalActions();
        acc->legal[std::min<size_t>(legal.size(), 7)]++;
        std::uniform_int_distribution<int> dis(0, static_cast<int>(legal.size()) - 1);
        state->ApplyAction(legal[dis(rng)]);
        acc->dec++;
      }
      ++moves;
    }
    auto* cs = static_cast<const open_spiel::coup::CoupState*>(state.get());
    acc->max_coins = std::max<long>(acc->max_coins, std::max(cs->GetCoins(0), cs->GetCoins(1)));
    acc->moves += moves;
    acc->eps++;
    acc->max_moves = std::max(acc->max_moves, moves);
    if (moves > 90) acc->trunc++;
    acc->ret[static_cast<int>(state->Returns()[0]) + 2]++;
  }
}

int ref_bench(int mode, int threads, long episodes_total, uint32_t seed, double* out) {
  TheGame();
  threads = std::max(1, threads);
  std::vector<BenchAcc> acc(threads);
  std::vector<std::thread> th;
  long per = (episodes_total + threads - 1) / threads;
  auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < threads; ++t) th.emplace_back(BenchThread, seed + 7919u * t, per, mode, &acc[t]);
  for (auto& x : th) x.join();
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  BenchAcc a;
  for (auto& x : acc) {
    a.moves += x.moves; a.dec += x.dec; a.chance += x.chance; a.eps += x.eps; a.trunc += x.trunc;
    for (int i = 0; i < 5; ++i) a.ret[i] += x.ret[i];
    for (int i = 0; i < 8; ++i) a.legal[i] += x.legal[i];
    a.max_moves = std::max(a.max_moves, x.max_moves);
    a.max_coins = std::max(a.max_coins, x.max_coins);
  }
  out[0] = secs; out[1] = a.moves; out[2] = a.dec; out[3] = a.chance; out[4] = a.eps; out[5] = a.trunc;
  for (int i = 0; i < 5; ++i) out[6 + i] = a.ret[i];
  for (int i = 0; i < 8; ++i) out[11 + i] = a.legal[i];
  out[19] = a.max_moves;
  out[20] = a.max_coins;
  return 0;
}

}  // extern "C"
