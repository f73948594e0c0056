// coup_b200_plugin.h -- OpenSpiel Game/State plugin for "coup" backed by libcoup_b200.so.
//
// This is the reference-side translation unit INTEGRATION.md describes: it is compiled AGAINST the
// reference's own headers (open_spiel/spiel.h, observer.h) and registers a game named "coup", so the
// reference's tests (games/coup_test.cc, tests/basic_tests.cc) and any OpenSpiel algorithm run against the
// GPU implementation unchanged. It provides the same public names as open_spiel/games/coup.h
// (namespace open_spiel::coup: CardType, CardStateType, ActionType, CoupState, CoupGame and the
// GetCardsValue / GetCardsState / GetCoins / GetLastAction accessors the reference test uses); nothing else
// is shared with the reference implementation: the state lives in a device slab and every rule is evaluated
// by CUDA kernels through the C ABI (include/coup_b200.h). Host-side work here is what the reference keeps
// on the host as well: history bookkeeping of the State base class, string formatting, and copying tensors
// into caller-owned spans.
#ifndef COUP_B200_PLUGIN_H_
#define COUP_B200_PLUGIN_H_

#include <memory>
#include <string>
#include <vector>

#include "open_spiel/observer.h"
#include "open_spiel/spiel.h"

namespace open_spiel {
namespace coup {

inline constexpr int kNumPlayers = 2;
inline constexpr int kMaxCardsInHand = 4;
inline constexpr int kNumCardTypes = 5;
inline constexpr int kNumEachCardInDeck = 3;

// Ids are part of the game's public contract (action ids, chance outcome ids).
enum class CardType { kNone = -1, kAssassin = 0, kAmbassador = 1, kCaptain = 2, kContessa = 3, kDuke = 4 };
enum class CardStateType { kNone = -1, kFaceDown = 0, kFaceUp = 1 };
enum class ActionType : Action {
  kNone = -1, kIncome = 0, kForeignAid = 1, kCoup = 2, kTax = 3, kAssassinate = 4, kExchange = 5, kSteal = 6,
  kLoseCard1 = 7, kLoseCard2 = 8, kPass = 9, kBlock = 10, kChallenge = 11, kExchangeReturn12 = 12,
  kExchangeReturn13 = 13, kExchangeReturn14 = 14, kExchangeReturn23 = 15, kExchangeReturn24 = 16,
  kExchangeReturn34 = 17
};

class CoupGame;

// Decoded copy of one env's packed device state (see include/coup_b200.h for the bit layout).
struct HostView {
  struct Card { int value; int face_up; };
  std::vector<Card> cards[2];
  int coins[2];
  int last_action[2];
  int deck[5];
  int turn_number;
  int cur_player_move;
  std::vector<int> deal_target;  // per history entry: receiving player of a chance deal, -1 for player moves
};

class CoupState : public State {
 public:
  explicit CoupState(std::shared_ptr<const Game> game);
  CoupState(const CoupState& other);  // clone: copies the device slot
  ~CoupState() override;

  Player CurrentPlayer() const override;
  std::string ActionToString(Player player, Action move) const override;
  std::string ToString() const override;
  bool IsTerminal() const override;
  std::vector<double> Rewards() const override;
  std::vector<double> Returns() const override;
  std::string InformationStateString(Player player) const override;
  std::string ObservationString(Player player) const override;
  void InformationStateTensor(Player player, absl::Span<float> values) const override;
  void ObservationTensor(Player player, absl::Span<float> values) const override;
  std::unique_ptr<State> Clone() const override;
  std::vector<std::pair<Action, double>> ChanceOutcomes() const override;
  std::vector<Action> LegalActions() const override;
  std::vector<Action> ActionsConsistentWithInformationFrom(Action action) const override { return {action}; }

  std::vector<CardType> GetCardsValue(Player player) const;
  std::vector<CardStateType> GetCardsState(Player player) const;
  int GetCoins(Player player) const;
  Action GetLastAction(Player player) const;

  HostView View() const;
  int slot() const { return slot_; }

 protected:
  void DoApplyAction(Action move) override;

 private:
  uint32_t StepWord() const;
  int slot_;
};

class CoupGame : public Game {
 public:
  explicit CoupGame(const GameParameters& params);
  int NumDistinctActions() const override { return 18; }
  std::unique_ptr<State> NewInitialState() const override;
  int MaxChanceOutcomes() const override { return kNumCardTypes; }
  int NumPlayers() const override { return kNumPlayers; }
  double MinUtility() const override { return -2; }
  double MaxUtility() const override { return 2; }
  absl::optional<double> UtilitySum() const override { return 0; }
  std::vector<int> InformationStateTensorShape() const override { return {2492}; }
  std::vector<int> ObservationTensorShape() const override { return {98}; }
  int MaxGameLength() const override { return 90; }
  int MaxChanceNodesInHistory() const override { return 45; }
  std::string ActionToString(Player player, Action action) const override;
  std::shared_ptr<Observer> MakeObserver(absl::optional<IIGObservationType> iig_obs_type,
                                         const GameParameters& params) const override;
};

}  // namespace coup
}  // namespace open_spiel

#endif  // COUP_B200_PLUGIN_H_
