// coup_b200_plugin.cc -- see coup_b200_plugin.h. Built only where the reference headers exist
// (plugin/Makefile); links libcoup_b200.so. No game rule is evaluated in this file.
#include "coup_b200_plugin.h"

#include <cstdlib>
#include <mutex>

#include "open_spiel/abseil-cpp/absl/strings/str_cat.h"
#include "open_spiel/spiel_utils.h"

extern "C" {
#include "coup_b200.h"
}

namespace open_spiel {
namespace coup {
namespace {

const GameType kGameType{/*short_name=*/"coup",
                         /*long_name=*/"Coup",
                         GameType::Dynamics::kSequential,
                         GameType::ChanceMode::kExplicitStochastic,
                         GameType::Information::kImperfectInformation,
                         GameType::Utility::kZeroSum,
                         GameType::RewardModel::kRewards,
                         /*max_num_players=*/2,
                         /*min_num_players=*/2,
                         /*provides_information_state_string=*/true,
                         /*provides_information_state_tensor=*/true,
                         /*provides_observation_string=*/true,
                         /*provides_observation_tensor=*/true,
                         /*parameter_specification=*/{}};

std::shared_ptr<const Game> Factory(const GameParameters& params) {
  return std::shared_ptr<const Game>(new CoupGame(params));
}
REGISTER_SPIEL_GAME(kGameType, Factory);

std::shared_ptr<Observer> MakeSingleTensorObserver(const Game& game, absl::optional<IIGObservationType> iig_obs_type,
                                                   const GameParameters& params) {
  return std::shared_ptr<Observer>(game.MakeBuiltInObserver(iig_obs_type));
}
ObserverRegisterer single_tensor(kGameType.short_name, "single_tensor", MakeSingleTensorObserver);

void Check(int rc, const char* what) {
  if (rc != COUP_OK) SpielFatalError(absl::StrCat(what, ": ", coup_last_error()));
}

// The device slab all live states of this process share: one env slot per State object.
class Slab {
 public:
  static Slab& Get() {
    static Slab slab;
    return slab;
  }
  int Alloc() {
    std::lock_guard<std::mutex> lk(mu_);
    if (free_.empty()) SpielFatalError("coup_b200 plugin: state capacity exhausted (COUP_B200_CAPACITY)");
    int s = free_.back();
    free_.pop_back();
    return s;
  }
  void Free(int slot) {
    std::lock_guard<std::mutex> lk(mu_);
    free_.push_back(slot);
  }
  coup_vec_env* env() { return env_; }

 private:
  Slab() {
    const char* cap = std::getenv("COUP_B200_CAPACITY");
    const char* dev = std::getenv("COUP_B200_DEVICE");
    coup_vec_opts opts{};
    opts.num_envs = cap ? static_cast<uint32_t>(std::atoi(cap)) : 4096u;
    opts.device = dev ? std::atoi(dev) : 0;
    opts.seed = 0;
    Check(coup_vec_create(&opts, &env_), "coup_vec_create");
    for (int s = static_cast<int>(opts.num_envs) - 1; s >= 0; --s) free_.push_back(s);
  }
  ~Slab() { coup_vec_destroy(env_); }
  coup_vec_env* env_ = nullptr;
  std::vector<int> free_;
  std::mutex mu_;
};

const char* const kCardNames[5] = {"Assassin", "Ambassador", "Captain", "Contessa", "Duke"};
const char* const kActionNames[18] = {"Income", "ForeignAid", "Coup", "Tax", "Assassinate", "Exchange", "Steal",
                                      "LoseCard1", "LoseCard2", "Pass", "Block", "Challenge", "ExchangeReturn12",
                                      "ExchangeReturn13", "ExchangeReturn14", "ExchangeReturn23", "ExchangeReturn24",
                                      "ExchangeReturn34"};
std::string CardName(int card) { return card < 0 ? "-" : kCardNames[card]; }
std::string ActionName(int action) { return action < 0 ? "None" : kActionNames[action]; }
std::string StateName(int state) { return state < 0 ? "None" : state == 0 ? "FaceDown" : "FaceUp"; }

// "P<k>" header plus one line per card; `shown(state)` decides whether the card's value is printed.
template <typename Shown>
void AppendCards(std::string* out, const HostView& v, int p, Shown shown, bool show_state) {
  absl::StrAppend(out, "P", p + 1, "\n        Card         State\n");
  for (size_t c = 0; c < v.cards[p].size(); ++c) {
    const auto& card = v.cards[p][c];
    const std::string name = CardName(shown(card.face_up) ? card.value : -1);
    absl::StrAppend(out, "Card ", c + 1, ": ", name, std::string(11 - name.size(), ' '), "| ",
                    StateName(show_state ? card.face_up : -1), "\n");
  }
}

class CoupObserver : public Observer {
 public:
  explicit CoupObserver(IIGObservationType t) : Observer(/*has_string=*/true, /*has_tensor=*/true), type_(t) {}

  void WriteTensor(const State& observed_state, int player, Allocator* allocator) const override {
    const auto& state = down_cast<const CoupState&>(observed_state);
    SPIEL_CHECK_GE(player, 0);
    SPIEL_CHECK_LT(player, kNumPlayers);
    const bool single = type_.private_info == PrivateInfoType::kSinglePlayer;
    if (type_.public_info && single) {
      // The two tensors the kernels produce: fetch the dense row and hand it out piece by piece in the
      // order the named tensors are laid out.
      const bool recall = type_.perfect_recall;
      std::vector<float> dense(recall ? COUP_INFO_STATE_SIZE : COUP_OBSERVATION_SIZE);
      if (recall) {
        Check(coup_env_information_state_tensor(Slab::Get().env(), state.slot(), player, dense.data(), COUP_INFO_STATE_SIZE),
              "coup_env_information_state_tensor");
      } else {
        Check(coup_env_observation_tensor(Slab::Get().env(), state.slot(), player, dense.data(), COUP_OBSERVATION_SIZE),
              "coup_env_observation_tensor");
      }
      const float* src = dense.data();
      auto emit = [&](const char* name, std::initializer_list<int> shape, int count) {
        auto out = allocator->Get(name, std::vector<int>(shape));
        for (int i = 0; i < count; ++i) out.data()[i] = src[i];
        src += count;
      };
      emit("player", {2}, 2);
      emit("p1_cards", {kMaxCardsInHand, kNumCardTypes}, 20);
      emit("p2_cards", {kMaxCardsInHand, kNumCardTypes}, 20);
      emit("cur_move_player", {2}, 2);
      emit("cards_state", {2, kMaxCardsInHand, 2}, 16);
      emit("coins", {2}, 2);
      if (recall) emit("history", {135, 18}, 135 * 18); else emit("last_action", {2, 18}, 36);
      return;
    }
    // Other observation types (public-only, private-only, all-players): laid out on the host from the
    // decoded device state. Pure data layout, single state, not a hot path.
    const HostView v = state.View();
    allocator->Get("player", {2}).at(player) = 1;
    for (int p = 0; p < 2; ++p) {
      auto out = allocator->Get(p == 0 ? "p1_cards" : "p2_cards", {kMaxCardsInHand, kNumCardTypes});
      const bool priv = type_.private_info == PrivateInfoType::kAllPlayers || (single && p == player);
      for (size_t i = 0; i < v.cards[p].size(); ++i) {
        const auto& c = v.cards[p][i];
        if ((priv && !c.face_up) || (type_.public_info && c.face_up)) out.at(static_cast<int>(i), c.value) = 1;
      }
    }
    if (!type_.public_info) return;
    auto cur = allocator->Get("cur_move_player", {2});
    if (!state.IsTerminal()) cur.at(v.cur_player_move) = 1;
    auto cs = allocator->Get("cards_state", {2, kMaxCardsInHand, 2});
    for (int p = 0; p < 2; ++p)
      for (size_t i = 0; i < v.cards[p].size(); ++i) cs.at(p, static_cast<int>(i), v.cards[p][i].face_up) = 1;
    auto coins = allocator->Get("coins", {2});
    coins.at(0) = v.coins[0];
    coins.at(1) = v.coins[1];
    if (type_.perfect_recall) {
      auto hist = allocator->Get("history", {135, 18});
      const auto& h = state.FullHistory();
      for (size_t i = 0; i < h.size(); ++i)
        if (h[i].player >= 0 || v.deal_target[i] == player) hist.at(static_cast<int>(i), static_cast<int>(h[i].action)) = 1;
    } else {
      auto last = allocator->Get("last_action", {2, 18});
      for (int p = 0; p < 2; ++p)
        if (v.last_action[p] >= 0) last.at(p, v.last_action[p]) = 1;
    }
  }

  std::string StringFrom(const State& observed_state, int player) const override {
    const auto& state = down_cast<const CoupState&>(observed_state);
    SPIEL_CHECK_GE(player, 0);
    SPIEL_CHECK_LT(player, kNumPlayers);
    const HostView v = state.View();
    const bool pub = type_.public_info, recall = type_.perfect_recall;
    const bool single = type_.private_info == PrivateInfoType::kSinglePlayer;
    const bool all = type_.private_info == PrivateInfoType::kAllPlayers;
    std::string out = absl::StrCat("Observer: P", player + 1, "\n");
    if (pub) absl::StrAppend(&out, "Turn: ", v.turn_number, "\nMove: P", v.cur_player_move + 1, "\n");
    for (int p = 0; p < 2; ++p) {
      if (pub || all || (single && player == p)) {
        AppendCards(&out, v, p,
                    [&](int up) { return (pub && up) || (single && p == player && !up) || (all && !up); }, pub);
      }
      if (pub) {
        absl::StrAppend(&out, "Coins: ", v.coins[p], "\n");
        if (!recall) absl::StrAppend(&out, "Last Action: ", ActionName(v.last_action[p]), "\n\n");
        else absl::StrAppend(&out, "\n");
      }
    }
    if (pub && recall) {
      absl::StrAppend(&out, "Action Sequence: ");
      const auto& h = state.FullHistory();
      for (size_t i = 0; i < h.size(); ++i) {
        const bool is_last = i + 1 == h.size();
        if (h[i].player == kChancePlayerId) {
          if (v.deal_target[i] == player) absl::StrAppend(&out, "PC-", CardName(static_cast<int>(h[i].action)), is_last ? "" : ", ");
        } else {
          absl::StrAppend(&out, "P", h[i].player + 1, "-", ActionName(static_cast<int>(h[i].action)), is_last ? "" : ", ");
        }
      }
      absl::StrAppend(&out, "\n");
    }
    return out;
  }

 private:
  IIGObservationType type_;
};

}  // namespace

// ---- CoupState ----------------------------------------------------------------------------------
CoupState::CoupState(std::shared_ptr<const Game> game) : State(game), slot_(Slab::Get().Alloc()) {
  Check(coup_env_new_initial_state(Slab::Get().env(), slot_), "coup_env_new_initial_state");
}

CoupState::CoupState(const CoupState& other) : State(other), slot_(Slab::Get().Alloc()) {
  Check(coup_env_clone(Slab::Get().env(), other.slot_, slot_), "coup_env_clone");
}

CoupState::~CoupState() { Slab::Get().Free(slot_); }

uint32_t CoupState::StepWord() const {
  uint32_t w = 0;
  Check(coup_env_read(Slab::Get().env(), slot_, nullptr, nullptr, &w), "coup_env_read");
  return w;
}

Player CoupState::CurrentPlayer() const {
  const uint32_t w = StepWord();
  if ((w >> 19) & 1u) return kTerminalPlayerId;
  if ((w >> 27) & 1u) return kChancePlayerId;
  return static_cast<Player>((w >> 18) & 1u);
}

bool CoupState::IsTerminal() const { return (StepWord() >> 19) & 1u; }

std::vector<Action> CoupState::LegalActions() const {
  const uint32_t w = StepWord();
  std::vector<Action> legal;
  for (int a = 0; a < 18; ++a)
    if ((w >> a) & 1u) legal.push_back(a);
  return legal;
}

void CoupState::DoApplyAction(Action move) {
  if (coup_env_apply_action(Slab::Get().env(), slot_, static_cast<int>(move)) != COUP_OK)
    SpielFatalError(absl::StrCat("Invalid player action: ", coup_last_error()));
}

std::vector<double> CoupState::Rewards() const {
  const double r0 = static_cast<int>((StepWord() >> 21) & 7u) - 2;
  return {r0, -r0};
}

std::vector<double> CoupState::Returns() const {
  const double r0 = static_cast<int>((StepWord() >> 24) & 7u) - 2;
  return {r0, -r0};
}

std::vector<std::pair<Action, double>> CoupState::ChanceOutcomes() const {
  SPIEL_CHECK_TRUE(IsChanceNode());
  const HostView v = View();
  double deck_size = 0;
  for (int c = 0; c < 5; ++c) deck_size += v.deck[c];
  std::vector<std::pair<Action, double>> outcomes;
  for (int c = 0; c < 5; ++c)
    if (v.deck[c] > 0) outcomes.push_back({c, v.deck[c] / deck_size});
  return outcomes;
}

HostView CoupState::View() const {
  uint32_t st[4], hist[16];
  Check(coup_env_read(Slab::Get().env(), slot_, st, hist, nullptr), "coup_env_read");
  HostView v;
  for (int p = 0; p < 2; ++p) {
    for (int i = 0; i < 4; ++i) {
      const uint32_t key = (st[p] >> (4 * i)) & 15u;
      if (key != 15u) v.cards[p].push_back({static_cast<int>(key >> 1), static_cast<int>(key & 1u)});
    }
    v.coins[p] = (st[p] >> 16) & 31u;
    const uint32_t last = (st[p] >> 21) & 31u;
    v.last_action[p] = last == 31u ? -1 : static_cast<int>(last);
  }
  for (int c = 0; c < 5; ++c) v.deck[c] = (st[2] >> (4 * c)) & 15u;
  v.cur_player_move = (st[2] >> 21) & 1u;
  v.turn_number = (st[3] >> 7) & 127u;
  v.deal_target.assign(history_.size(), -1);
  for (size_t i = 0; i < history_.size(); ++i) {
    const uint32_t code = (hist[i / 6] >> (5 * (i % 6))) & 31u;
    if (code >= 18u) v.deal_target[i] = static_cast<int>((code - 18u) / 5u);
  }
  return v;
}

std::string CoupState::ActionToString(Player player, Action move) const { return GetGame()->ActionToString(player, move); }

std::string CoupState::ToString() const {
  const HostView v = View();
  std::string out = absl::StrCat("Turn: ", v.turn_number, "\nMove: P", v.cur_player_move + 1, "\n");
  for (int p = 0; p < 2; ++p) {
    AppendCards(&out, v, p, [](int) { return true; }, true);
    absl::StrAppend(&out, "Coins: ", v.coins[p], "\nLast Action: ", ActionName(v.last_action[p]), "\n\n");
  }
  absl::StrAppend(&out, "Action Sequence: ");
  for (size_t i = 0; i < history_.size(); ++i) {
    const auto& pa = history_[i];
    if (pa.player == kChancePlayerId) absl::StrAppend(&out, "PC-", CardName(static_cast<int>(pa.action)));
    else absl::StrAppend(&out, "P", pa.player + 1, "-", ActionName(static_cast<int>(pa.action)));
    if (i + 1 < history_.size()) absl::StrAppend(&out, ", ");
  }
  absl::StrAppend(&out, "\n");
  return out;
}

std::string CoupState::InformationStateString(Player player) const {
  return CoupObserver(kInfoStateObsType).StringFrom(*this, player);
}
std::string CoupState::ObservationString(Player player) const {
  return CoupObserver(kDefaultObsType).StringFrom(*this, player);
}

void CoupState::InformationStateTensor(Player player, absl::Span<float> values) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, kNumPlayers);
  SPIEL_CHECK_EQ(values.size(), COUP_INFO_STATE_SIZE);
  Check(coup_env_information_state_tensor(Slab::Get().env(), slot_, player, values.data(), COUP_INFO_STATE_SIZE),
        "coup_env_information_state_tensor");
}

void CoupState::ObservationTensor(Player player, absl::Span<float> values) const {
  SPIEL_CHECK_GE(player, 0);
  SPIEL_CHECK_LT(player, kNumPlayers);
  SPIEL_CHECK_EQ(values.size(), COUP_OBSERVATION_SIZE);
  Check(coup_env_observation_tensor(Slab::Get().env(), slot_, player, values.data(), COUP_OBSERVATION_SIZE),
        "coup_env_observation_tensor");
}

std::unique_ptr<State> CoupState::Clone() const { return std::unique_ptr<State>(new CoupState(*this)); }

std::vector<CardType> CoupState::GetCardsValue(Player player) const {
  SPIEL_CHECK_LT(player, NumPlayers());
  std::vector<CardType> out;
  for (const auto& c : View().cards[player]) out.push_back(static_cast<CardType>(c.value));
  return out;
}
std::vector<CardStateType> CoupState::GetCardsState(Player player) const {
  SPIEL_CHECK_LT(player, NumPlayers());
  std::vector<CardStateType> out;
  for (const auto& c : View().cards[player]) out.push_back(static_cast<CardStateType>(c.face_up));
  return out;
}
int CoupState::GetCoins(Player player) const {
  SPIEL_CHECK_LT(player, NumPlayers());
  return View().coins[player];
}
Action CoupState::GetLastAction(Player player) const {
  SPIEL_CHECK_LT(player, NumPlayers());
  return View().last_action[player];
}

// ---- CoupGame -----------------------------------------------------------------------------------
CoupGame::CoupGame(const GameParameters& params) : Game(kGameType, params) {}

std::unique_ptr<State> CoupGame::NewInitialState() const {
  return std::unique_ptr<State>(new CoupState(shared_from_this()));
}

std::string CoupGame::ActionToString(Player player, Action action) const {
  if (player == kChancePlayerId) return absl::StrCat("Chance drawn card:", CardName(static_cast<int>(action)));
  return ActionName(static_cast<int>(action));
}

std::shared_ptr<Observer> CoupGame::MakeObserver(absl::optional<IIGObservationType> iig_obs_type,
                                                 const GameParameters& params) const {
  if (params.empty()) return std::make_shared<CoupObserver>(iig_obs_type.value_or(kDefaultObsType));
  return MakeRegisteredObserver(iig_obs_type, params);
}

}  // namespace coup
}  // namespace open_spiel
