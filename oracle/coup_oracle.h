/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's Coup path (open_spiel/games/coup.{h,cc} of
 * BStarcheus/open_spiel_coup plus State::ApplyAction, spiel.cc:322-332). It exists to CHECK the CUDA
 * path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it. The product
 * library (open_spiel_coup_b200/csrc) never includes, links or calls anything in oracle/.
 *
 * Parity is PINNED: this file is differential-tested against the unmodified reference compiled into
 * oracle/_ref/libcoup_ref.so (tests/test_oracle_vs_reference.py), against the reference's 14 scenario
 * known-answer tests (coup_test.cc:41-556) and against its golden playthrough
 * (integration_tests/playthroughs/coup.txt), via the fixtures in tests/golden/.
 *
 * Every function cites the reference file:line it restates.
 */
#ifndef COUP_ORACLE_H_
#define COUP_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  OC_NUM_PLAYERS = 2,          /* coup.h:42 */
  OC_MAX_CARDS_IN_HAND = 4,    /* coup.h:43 */
  OC_NUM_CARD_TYPES = 5,       /* coup.h:44 */
  OC_NUM_EACH_CARD = 3,        /* coup.h:45 */
  OC_NUM_ACTIONS = 18,         /* coup.h:203 */
  OC_MAX_GAME_LENGTH = 90,     /* coup.h:219 */
  OC_MAX_CHANCE_NODES = 45,    /* coup.h:220 */
  OC_MAX_MOVE_NUMBER = 135,    /* spiel.h:888-890: MaxGameLength + MaxChanceNodesInHistory */
  OC_INFO_STATE_SIZE = 2492,   /* coup.cc:1104-1116 */
  OC_OBSERVATION_SIZE = 98,    /* coup.cc:1118-1130 */
  OC_HIST_CAP = 96             /* > 91 = the most moves a non-terminal game can have applied */
};

/* Player ids, spiel_globals.h:26-36 */
enum { OC_CHANCE_PLAYER = -1, OC_TERMINAL_PLAYER = -4 };

/* ActionType, coup.h:65-85 */
enum {
  OC_NONE = -1, OC_INCOME = 0, OC_FOREIGN_AID = 1, OC_COUP = 2, OC_TAX = 3, OC_ASSASSINATE = 4,
  OC_EXCHANGE = 5, OC_STEAL = 6, OC_LOSE_CARD_1 = 7, OC_LOSE_CARD_2 = 8, OC_PASS = 9, OC_BLOCK = 10,
  OC_CHALLENGE = 11, OC_EXCHANGE_RETURN_12 = 12, OC_EXCHANGE_RETURN_13 = 13,
  OC_EXCHANGE_RETURN_14 = 14, OC_EXCHANGE_RETURN_23 = 15, OC_EXCHANGE_RETURN_24 = 16,
  OC_EXCHANGE_RETURN_34 = 17
};
/* CardType, coup.h:50-57 */
enum { OC_ASSASSIN = 0, OC_AMBASSADOR = 1, OC_CAPTAIN = 2, OC_CONTESSA = 3, OC_DUKE = 4 };
/* CardStateType, coup.h:59-63 */
enum { OC_FACE_DOWN = 0, OC_FACE_UP = 1 };

typedef struct {          /* coup.h:87-95 */
  int32_t value;
  int32_t state;
} oc_card;

typedef struct {          /* coup.h:97-109 */
  oc_card cards[OC_MAX_CARDS_IN_HAND];
  int32_t num_cards;
  int32_t coins;
  int32_t last_action;
  int32_t lost_challenge;
} oc_player;

typedef struct {          /* coup.h:163-196 + State base (spiel.h:733-734) */
  int32_t deck[OC_NUM_CARD_TYPES];
  oc_player players[OC_NUM_PLAYERS];
  int32_t deal_queue[8];  /* deal_card_to_ (std::queue<Player>); never holds more than 4 */
  int32_t deal_head, deal_tail;
  int32_t cur_player_turn;
  int32_t cur_player_move;
  int32_t opp_player;
  int32_t is_turn_begin;
  int32_t turn_number;
  int32_t is_chance;
  int32_t cur_rewards[OC_NUM_PLAYERS];
  int32_t move_number;                     /* State::move_number_ */
  int32_t history_len;                     /* State::history_.size() */
  int8_t history_player[OC_HIST_CAP];      /* PlayerAction.player (-1 for chance) */
  int8_t history_action[OC_HIST_CAP];      /* PlayerAction.action */
  int8_t history_deal_player[OC_HIST_CAP]; /* history_chance_deal_player_ (-1 where not chance) */
  int32_t error;                           /* sticky: set where the reference would SpielFatalError */
} oc_state;

/* Error codes (the reference aborts; the oracle reports). */
enum { OC_OK = 0, OC_ERR_ILLEGAL = -1, OC_ERR_TERMINAL = -2, OC_ERR_INTERNAL = -3 };

int oc_sizeof_state(void);
void oc_init(oc_state* s);                                   /* CoupState ctor, coup.cc:393-428 */
int oc_is_terminal(const oc_state* s);                       /* coup.cc:989-1010 */
int oc_current_player(const oc_state* s);                    /* coup.cc:458-466 */
int oc_is_chance_node(const oc_state* s);                    /* spiel.h IsChanceNode */
int oc_apply_action(oc_state* s, int action);                /* spiel.cc:322-332 + coup.cc:490-809 */
int oc_apply_action_checked(oc_state* s, int action);        /* as above, rejects a not in LegalActions */
int oc_legal_actions(const oc_state* s, int32_t* out);       /* coup.cc:811-938; returns count */
uint32_t oc_legal_mask(const oc_state* s);                   /* spiel.cc:371-377 as a bitmask */
int oc_chance_outcomes(const oc_state* s, int32_t* actions, double* probs); /* coup.cc:1062-1077 */
void oc_returns(const oc_state* s, double* out2);            /* coup.cc:1016-1032 */
void oc_rewards(const oc_state* s, double* out2);            /* coup.cc:1012-1014 */
void oc_information_state_tensor(const oc_state* s, int player, float* out); /* coup.cc:1044-1049 */
void oc_observation_tensor(const oc_state* s, int player, float* out);       /* coup.cc:1051-1056 */
/* General observer, coup.cc:248-287. private_info: 0 none, 1 single player, 2 all players.
 * Returns the number of floats written (42, 98 or 2492). */
int oc_observer_tensor(const oc_state* s, int player, int public_info, int perfect_recall,
                       int private_info, float* out);
uint64_t oc_tensor_hash(const float* t, int n);

/* One record per visited state, identical layout to RefTraceRec in oracle/ref_harness.cc. */
typedef struct {
  int8_t cur_player;
  uint8_t is_terminal;
  uint8_t is_chance;
  uint8_t move_number;
  uint32_t legal_mask;
  int8_t rewards[2];
  int8_t returns[2];
  uint8_t coins[2];
  uint8_t ncards[2];
  uint64_t hash_info[2];
  uint64_t hash_obs[2];
} oc_trace_rec;

int oc_trace(const uint8_t* actions, int n_actions, oc_trace_rec* out);
int oc_trace_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, oc_trace_rec* out);
int oc_state_from_actions(oc_state* s, const uint8_t* actions, int n_actions);
int oc_final_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, oc_trace_rec* out);
int oc_final_tensors_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, float* info, float* obs);

int oc_replay_digest(const uint8_t* actions, int n_actions, uint64_t* digest_out);
int oc_replay_digest_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, uint64_t* digests,
                           int32_t* reports);

/* Batched helpers over an array of states (used by the GPU parity tests). */
void oc_batch_init(oc_state* s, int n);
/* Applies actions[i] to state i when do_mask is NULL or do_mask[i] != 0. Returns #errors. */
int oc_batch_apply(oc_state* s, int n, const uint8_t* actions, const uint8_t* do_mask);
void oc_batch_info_state(const oc_state* s, int n, const int8_t* players, float* out);
void oc_batch_observation(const oc_state* s, int n, const int8_t* players, float* out);

/* CPU baseline "port": uniform-random rollouts with the benchmark_game.cc protocol
 * (examples/benchmark_game.cc:32-140), multithreaded. Same out[] layout as ref_bench. */
int oc_bench(int mode, int threads, long episodes_total, uint32_t seed, double* out);

#ifdef __cplusplus
}
#endif
#endif  /* COUP_ORACLE_H_ */
