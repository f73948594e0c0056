/* coup_b200.h -- C ABI of libcoup_b200.so, the B200-native batched Coup environment.
 *
 * This is the drop-in boundary for the Coup game of BStarcheus/open_spiel_coup
 * (open_spiel/games/coup.{h,cc}). Conventions follow the reference's own pure-C API
 * (open_spiel/rust/src/rust_open_spiel.h): opaque handles, scalar getters return int, tensors are
 * written into caller-provided buffers with an explicit length. Differences, all forced by batching
 * on a GPU: big buffers are DEVICE pointers, every call takes an explicit cudaStream_t (passed as
 * void*), and instead of aborting the process on an illegal move (spiel_utils.cc:119-137) calls return
 * a status code and the environment records a sticky per-env error bit.
 *
 * There is no CPU implementation behind this header: every entry point launches sm_100a kernels (or
 * copies device memory). COUP_ERR_NO_DEVICE is returned when no CUDA device is usable.
 *
 * Action ids (coup.h:65-85):  0 Income 1 ForeignAid 2 Coup 3 Tax 4 Assassinate 5 Exchange 6 Steal
 *   7 LoseCard1 8 LoseCard2 9 Pass 10 Block 11 Challenge 12..17 ExchangeReturn{12,13,14,23,24,34}
 * Chance outcome ids = card types (coup.h:50-57): 0 Assassin 1 Ambassador 2 Captain 3 Contessa 4 Duke
 */
#ifndef COUP_B200_H_
#define COUP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- game constants (CoupGame, coup.h:199-231, coup.cc:1104-1130) ------------------------------- */
#define COUP_NUM_PLAYERS 2
#define COUP_NUM_DISTINCT_ACTIONS 18
#define COUP_MAX_CHANCE_OUTCOMES 5
#define COUP_MAX_GAME_LENGTH 90
#define COUP_MAX_CHANCE_NODES_IN_HISTORY 45
#define COUP_INFO_STATE_SIZE 2492   /* player2 p1_cards20 p2_cards20 cur_move_player2 cards_state16 coins2 history135x18 */
/* The LIVE PREFIX of an info-state row. A game has at most 91 moves, chance nodes included (MaxGameLength 90, coup.h:219;
 * terminal at move_number_ > 90, coup.cc:990), while the tensor reserves 135 history rows (coup.cc:1104-1116): elements
 * >= 62 + 18 * 91 = 1700 are zero in every reachable state. Passing this value as `row_stride` to a *_strided / gather /
 * records entry point writes rows of 1728 elements (1700 rounded up to 64): the same values as elements [0, 1728) of the
 * full row, 31 % fewer bytes, nothing lost. */
#define COUP_LIVE_INFO_STATE_SIZE 1728
#define COUP_OBSERVATION_SIZE 98    /* ... same 62-float head ... last_action2x18 */
#define COUP_MIN_UTILITY (-2)
#define COUP_MAX_UTILITY 2
#define COUP_CHANCE_PLAYER_ID (-1)   /* spiel_globals.h:28 */
#define COUP_TERMINAL_PLAYER_ID (-4) /* spiel_globals.h:34 */

/* ---- status codes ------------------------------------------------------------------------------ */
enum {
  COUP_OK = 0,
  COUP_ERR_INVALID_ARG = 1,
  COUP_ERR_CUDA = 2,        /* a CUDA runtime call failed; see coup_last_error() */
  COUP_ERR_NO_DEVICE = 3,
  COUP_ERR_ILLEGAL_ACTION = 4 /* at least one env was handed an action outside LegalActions() */
};

/* ---- packed device layout (documented so that callers can snapshot / inspect it) -----------------
 * state: uint32[num_envs][4] (one 16-byte word per env)
 *   w0, w1 = player 0 / player 1:
 *     bits  0-15 hand: four 4-bit slots, slot = (card_value << 1) | face_up, 0xF = empty; always
 *                sorted ascending, i.e. by (value, FaceDown < FaceUp) as CoupPlayer::SortCards
 *                (coup.h:91-94, coup.cc:389-391)
 *     bits 16-20 coins        bits 21-25 last_action (31 = kNone)        bit 26 lost_challenge
 *   w2: bits 0-19 deck counts, 4 bits per card type (deck_, coup.h:164)
 *       bit 20 cur_player_turn_  bit 21 cur_player_move_  bit 22 is_turn_begin_  bit 23 is_chance_
 *       bits 24-26 number of queued deals (deal_card_to_)  bit 27 queue is the initial 0,1,0,1 deal
 *       bit 28 player the queued deals go to (when bit 27 is clear)  bit 29 sticky illegal-action flag
 *   w3: bits 0-6 move_number_  bits 7-13 turn_number_  bits 14-16 cur_rewards_[0] + 2
 *       (cur_rewards_[1] == -cur_rewards_[0] always)
 * history: uint32[num_envs][16]; move i of the episode is the 5-bit code at word i/6, bit 5*(i%6):
 *   0..17 player action id, 18+c card c dealt to player 0, 23+c card c dealt to player 1.
 *   Entries at index >= move_number_ are unspecified.
 */
#define COUP_STATE_WORDS 4
#define COUP_HISTORY_WORDS 16
/* Packed observation record (finished-episode ring, compact replay / reservoir records): uint32[24], 16-byte aligned:
 *   [0,16) history words   [16,20) state words   [20,24) meta words (meaning depends on who wrote the record).
 * Everything both tensor observers can show about a state is a function of these 80 bytes + the observer id, so a record
 * decodes into the same 2492-element row the dense encoder would have written (coup_records_information_state_tensor). */
#define COUP_RECORD_WORDS 24

/* ---- options ----------------------------------------------------------------------------------- */
enum {
  COUP_FLAG_AUTO_RESET = 1u << 0 /* vector_env.SyncVectorEnv.step(reset_if_done=True) semantics
                                    (python/vector_env.py:40-66): an env whose step ends the episode
                                    reports done/rewards/returns of the finished episode and is
                                    re-dealt in the same call; legal mask / current player / tensors
                                    then describe the first decision of the new episode */
  , COUP_FLAG_PLAIN_STORE_ENCODER = 1u << 1 /* info-state rows written with per-lane vector stores instead
                                    of the default shared-memory staging + bulk (TMA) stores; same bytes */
  , COUP_FLAG_NO_WARP_SPECIALISATION = 1u << 2 /* fused rollout: one CTA per 256 envs instead of the default
                                    persistent kernel whose rules warps and encoder warps overlap; same bytes */
  , COUP_FLAG_BLOCKING_SYNC = 1u << 3 /* the host-buffer calls (coup_vec_step_host*) sleep instead of spinning while
                                    they wait for the device: for hosts with fewer cores than threads */
};

typedef struct coup_vec_opts {
  uint32_t num_envs;          /* environments owned by this handle (one handle per GPU) */
  int32_t device;             /* CUDA device ordinal */
  uint64_t seed;              /* Philox4x32-10 seed */
  uint64_t global_env_offset; /* global id of env 0; Philox key = (seed, global_env_offset + i), so
                                 trajectories do not depend on how envs are sharded over GPUs */
  uint32_t flags;             /* COUP_FLAG_* */
  uint32_t reserved;
} coup_vec_opts;

typedef struct coup_vec_env coup_vec_env; /* opaque */

/* Which observer's view to encode. */
enum {
  COUP_PLAYER_0 = 0,
  COUP_PLAYER_1 = 1,
  COUP_PLAYER_CURRENT = 2, /* cur_player_move_ of each env (benchmark_game.cc:53-60 protocol) */
  COUP_PLAYER_BOTH = 3,    /* rows [env][player] (rl_environment.get_time_step, rl_environment.py:243-249) */
  COUP_PLAYER_FROM_RECORD = 4 /* record decoders only: the seat stored in bit 31 of meta word 1 of each record */
};
/* Element type of encoded tensors. Values are 0, 1 and coin counts <= 12: exact in all three. */
enum { COUP_DTYPE_F32 = 0, COUP_DTYPE_U8 = 1, COUP_DTYPE_BF16 = 2 };

/* Index of each counter in the statistics vector (uint64[COUP_STATS_LEN]). */
enum {
  COUP_STAT_DECISION_STEPS = 0, /* player actions applied */
  COUP_STAT_CHANCE_MOVES = 1,   /* chance deals applied (incl. the 4 initial deals of every episode) */
  COUP_STAT_EPISODES = 2,       /* episodes that reached a terminal state */
  COUP_STAT_TRUNCATED = 3,      /* ... of which ended by move_number_ > MaxGameLength (coup.cc:990) */
  COUP_STAT_EPISODE_MOVES = 4,  /* sum of move_number_ over finished episodes */
  COUP_STAT_ILLEGAL = 5,        /* illegal actions rejected */
  COUP_STAT_RETURN_HIST = 8,    /* [8..12]: finished episodes by Returns()[0] = -2..+2 */
  COUP_STAT_LEGAL_HIST = 16,    /* [16..23]: decision steps by number of legal actions 0..7 */
  COUP_STATS_LEN = 32
};

const char* coup_last_error(void);
int coup_device_count(void);

/* ---- lifecycle (GameNewInitialState / DeleteState, rust_open_spiel.h:41,48, batched) ----------- */
int coup_vec_create(const coup_vec_opts* opts, coup_vec_env** out);
int coup_vec_destroy(coup_vec_env* env);
uint32_t coup_vec_num_envs(const coup_vec_env* env);

/* ---- reset (CoupState ctor coup.cc:393-428 + the four initial deals; rl_environment.reset,
 * python/rl_environment.py:324-367). d_reset_mask: uint8[num_envs] on the device or NULL for all.
 * d_forced_deals: uint8[num_envs][4] card ids to deal instead of sampling (known-answer replay), or
 * NULL; an entry 0xFF means "sample". */
int coup_vec_reset(coup_vec_env* env, const uint8_t* d_reset_mask, const uint8_t* d_forced_deals,
                   void* stream);

/* ---- step (State::ApplyAction spiel.cc:322-332 + CoupState::DoApplyAction coup.cc:490-809, then
 * every following chance node resolved as rl_environment._sample_external_events does,
 * rl_environment.py:369-382). d_actions: uint8[num_envs] action ids on the device. d_forced_chance:
 * uint8[num_envs][4] outcomes for the chance nodes that follow (0xFF = sample), or NULL.
 * Envs that are already terminal ignore their action (rl_environment.py:301-302); the action id 0xFF makes an
 * env sit the step out (state and outputs untouched), for callers that advance only part of a slab. After the call the
 * per-env outputs below describe the new state. Returns COUP_ERR_ILLEGAL_ACTION only from
 * coup_vec_check_errors (the launch itself is asynchronous). */
int coup_vec_step(coup_vec_env* env, const uint8_t* d_actions, const uint8_t* d_forced_chance,
                  void* stream);

/* ---- the OpenSpiel State surface with EXPLICIT chance nodes (one move at a time), used by the single-state
 * Game/State mirror (open_spiel_coup_b200/spiel.py). Envs driven this way can sit at chance nodes; then
 * coup_vec_current_player() is COUP_CHANCE_PLAYER_ID, the legal mask holds the card ids still in the deck
 * (coup.cc:828-836) and bit 27 of the step word is set. Do not mix with coup_vec_step/rollout on one handle.
 *   coup_vec_new_initial_state: CoupState ctor (coup.cc:393-428) WITHOUT dealing; d_mask as in coup_vec_reset.
 *   coup_vec_apply_move: State::ApplyAction (spiel.cc:322-332): d_moves uint8[num_envs], a card id at a
 *     chance node / an action id at a decision node, 0xFF = leave the env untouched.
 *   coup_vec_copy_env: State::Clone (coup.cc:1058-1060) from slot src to slot dst. */
int coup_vec_new_initial_state(coup_vec_env* env, const uint8_t* d_mask, void* stream);
int coup_vec_apply_move(coup_vec_env* env, const uint8_t* d_moves, void* stream);
int coup_vec_copy_env(coup_vec_env* env, uint32_t src, uint32_t dst, void* stream);

/* Batched `state.child(action)` (spiel.h:565-570 Clone + ApplyAction), the tree-expansion step of the sampled
 * CFR traversals (python/algorithms/deep_cfr.py:415-497; every recursive call there is one child). For
 * i < count: env i of `dst` becomes a copy of env d_parent[i] of `src` (state + history) with player action
 * d_actions[i] applied and every following chance node resolved, exactly like coup_vec_step but NEVER
 * auto-resetting (a terminal child stays terminal so that its Returns() can be read). The chance draws come from
 * dst's Philox stream (dst seed, dst global env id of slot i, dst step counter, which advances by one), so
 * siblings forked from one parent draw independently, as np.random.choice does per recursive call
 * (deep_cfr.py:432-434). d_forced_chance as in coup_vec_step (uint8[count][4] or NULL). Envs >= count of `dst`
 * are left untouched; all per-env outputs of `dst` are refreshed for i < count. A terminal parent or an illegal
 * action sets the child's sticky error bit. `dst` and `src` must be different handles on the same device with
 * count <= num_envs(dst) and every d_parent[i] < num_envs(src) (out-of-range parents set the error bit). */
int coup_vec_fork(coup_vec_env* dst, const coup_vec_env* src, const uint32_t* d_parent, const uint8_t* d_actions,
                  const uint8_t* d_forced_chance, uint32_t count, void* stream);

/* One level of a sampled CFR traversal, for `count` nodes at once (python/algorithms/deep_cfr.py:415-525). Inputs:
 * the advantage-network outputs of the player to move (float[count][18]) and the nodes' step words
 * (coup_vec_step_word: legal mask + current player). coup_cfr_expand writes the regret-matched strategy
 * (`_sample_action_from_advantage`, :499-525) and, as a bit mask per node, the children to expand: every legal action
 * of the traverser with `external` != 0 (:438-441), else min(n_legal, k) actions drawn without replacement from
 * expl * uniform + (1 - expl) * strategy with k = outcome_factor, or -- when e_outcome >= 0 -- outcome_factor with
 * probability e_outcome and 1 otherwise (:442-466); one action drawn from the strategy at the opponent's nodes
 * (:482-487). Draws come from Philox4x32-10 keyed by (seed, node index) at `counter`. d_child_count_out[i] is the
 * number of bits set. coup_cfr_children turns the masks into the (parent, action) lists coup_vec_fork consumes, given
 * the exclusive prefix sum d_offsets of the child counts (int64[count]). Device pointers; not tied to a handle. */
int coup_cfr_expand(const float* d_advantages, const uint32_t* d_step_words, uint32_t count, int traverser, int external,
                    uint32_t outcome_factor, float e_outcome, float expl, uint64_t seed, uint64_t counter,
                    float* d_strategy_out, uint32_t* d_expand_out, uint32_t* d_child_count_out, void* stream);
int coup_cfr_children(const uint32_t* d_expand, const int64_t* d_offsets, uint32_t count, uint32_t* d_parent_out,
                      uint8_t* d_action_out, void* stream);

/* The same traversal level WITHOUT the host in the loop: the number of nodes of a level lives in device memory (*d_count),
 * every call is launched for a capacity and works on the first min(*d_count, capacity) nodes.
 *   coup_vec_information_state_tensor_prefix: info-state rows of envs [0, *d_count) of a slab.
 *   coup_cfr_level: coup_cfr_expand + the prefix sum of the child counts + coup_cfr_children in one single-CTA kernel:
 *     strategy [capacity][18], expand masks, d_offset_out[i] = index of node i's first child in the next level, the
 *     children's (parent, action) lists in parent order, *d_next_count = number of children (clipped to capacity, then
 *     *d_overflow = 1). Terminal nodes (bit 19 of the step word) have no children.
 *   coup_vec_fork_counted: coup_vec_fork for the first min(*d_count, max_count) children.
 *   coup_vec_pack_records: the nodes [0, *d_count) of a slab as packed records (meta: node index, seat<<31 | step word).
 *   coup_cfr_backward: the return path of `_traverse_game_tree` (deep_cfr.py:468-480, 492-497) for one level: values
 *     (double) from the next level's values, and the sampled regrets [capacity][18] of the traverser's nodes. */
int coup_vec_information_state_tensor_prefix(coup_vec_env* env, const uint32_t* d_count, uint32_t max_count, int player, int dtype,
                                             void* d_out, uint32_t row_stride, void* stream);
int coup_cfr_level(const float* d_advantages, const uint32_t* d_step_words, const uint32_t* d_count, uint32_t capacity,
                   int traverser, int external, uint32_t outcome_factor, float e_outcome, float expl, uint64_t seed,
                   uint64_t counter, float* d_strategy_out, uint32_t* d_expand_out, uint32_t* d_offset_out,
                   uint32_t* d_parent_out, uint8_t* d_action_out, uint32_t* d_next_count, uint32_t* d_overflow, void* stream);
int coup_vec_fork_counted(coup_vec_env* dst, const coup_vec_env* src, const uint32_t* d_parent, const uint8_t* d_actions,
                          const uint32_t* d_count, uint32_t max_count, void* stream);
int coup_vec_pack_records(coup_vec_env* env, const uint32_t* d_count, uint32_t* d_records_out, void* stream);
int coup_cfr_backward(const uint32_t* d_step_words, const uint32_t* d_count, uint32_t capacity, int traverser,
                      const float* d_strategy, const uint32_t* d_expand, const uint32_t* d_offset, const double* d_child_value,
                      double* d_value_out, float* d_regret_out, void* stream);

/* Single-env accessors with HOST buffers, following rust_open_spiel.h one to one (GameNewInitialState :41,
 * StateApplyAction :62, StateClone :49, StateInformationStateTensor / StateObservationTensor :73-76: tensors
 * are written into a caller-provided buffer of explicit length). `slot` indexes an env of the handle; these
 * calls synchronise. coup_env_apply_action returns COUP_ERR_ILLEGAL_ACTION and changes nothing when the move is
 * not in LegalActions() (card ids at chance nodes). coup_env_read copies the packed state (uint32[4]), the
 * history (uint32[16]) and the step word of one env; any pointer may be NULL. */
int coup_env_new_initial_state(coup_vec_env* env, uint32_t slot);
int coup_env_apply_action(coup_vec_env* env, uint32_t slot, int action);
int coup_env_clone(coup_vec_env* env, uint32_t src, uint32_t dst);
int coup_env_read(coup_vec_env* env, uint32_t slot, uint32_t* h_state4, uint32_t* h_history16, uint32_t* h_step_word);
int coup_env_information_state_tensor(coup_vec_env* env, uint32_t slot, int player, float* h_buf, int length);
int coup_env_observation_tensor(coup_vec_env* env, uint32_t slot, int player, float* h_buf, int length);

/* Uniform-random legal action per env (the policy of benchmark_game.cc:96-99), Philox-driven:
 * writes uint8[num_envs] to d_actions_out (terminal envs get 0xFF). */
int coup_vec_sample_uniform(coup_vec_env* env, uint8_t* d_actions_out, void* stream);

/* Masked policy sampling, the acting rule of the reference's agents (python/algorithms/nfsp.py:154-167:
 * softmax, zero the illegal actions, renormalise, sample): d_logits is dtype[num_envs][18] (COUP_DTYPE_F32 or
 * COUP_DTYPE_BF16) on the device; writes the sampled action ids to d_actions_out (uint8[num_envs], 0xFF for
 * terminal envs) and, if d_probs_out is not NULL, the renormalised probabilities float[num_envs][18]. */
int coup_vec_sample_policy(coup_vec_env* env, const void* d_logits, int dtype, float* d_probs_out,
                           uint8_t* d_actions_out, void* stream);

/* Fused random rollout: n_steps x (sample uniform legal action, step, resolve chance, [auto-reset],
 * encode). encode_player: COUP_PLAYER_* or -1 for no tensor. d_tensor_out: dtype[rows][2492] where
 * rows = num_envs (x2 for COUP_PLAYER_BOTH); it is overwritten at every step (the consumer reads it
 * between steps when n_steps == 1). */
int coup_vec_rollout(coup_vec_env* env, int n_steps, int encode_player, int dtype, void* d_tensor_out,
                     void* stream);

/* Fused random rollout with the INCREMENTAL tensor contract (contract I): d_buf is a
 * persistent dtype[num_envs][2][row_stride] buffer that already holds both players' info-state rows of every env
 * (fill it once with coup_vec_information_state_tensor_strided(COUP_PLAYER_BOTH)); every step rewrites only the
 * elements that changed (head, the new history rows, and zeros over a finished episode's rows), so the buffer always
 * equals what the dense encoder would write. Re-fill it after coup_vec_reset / coup_vec_restore. */
int coup_vec_rollout_incremental(coup_vec_env* env, int n_steps, int dtype, void* d_buf, uint32_t row_stride, void* stream);

/* ---- per-env outputs of the last reset/step (device pointers owned by the handle) --------------- */
const uint32_t* coup_vec_legal_mask(const coup_vec_env* env);  /* LegalActions() as bit a, coup.cc:824-938 */
const int8_t* coup_vec_current_player(const coup_vec_env* env); /* 0, 1 or -4 terminal, coup.cc:458-466 */
const uint8_t* coup_vec_done(const coup_vec_env* env);          /* IsTerminal() of the state just stepped */
const int8_t* coup_vec_rewards(const coup_vec_env* env);        /* [num_envs][2] Rewards(), coup.cc:1012-1014 */
const int8_t* coup_vec_returns(const coup_vec_env* env);        /* [num_envs][2] Returns(), coup.cc:1016-1032 */
/* One word per env with everything a host-side policy needs: bits 0-17 legal mask, 18 current player,
 * 19 state is terminal, 20 done (an episode ended in this step), 21-23 Rewards()[0]+2, 24-26 Returns()[0]+2. */
const uint32_t* coup_vec_step_word(const coup_vec_env* env);
#define COUP_WORD_LEGAL(w) ((w) & 0x3FFFFu)
#define COUP_WORD_CURRENT_PLAYER(w) (((w) >> 19) & 1u ? COUP_TERMINAL_PLAYER_ID : (int)(((w) >> 18) & 1u))
#define COUP_WORD_DONE(w) (((w) >> 20) & 1u)
#define COUP_WORD_REWARD0(w) ((int)(((w) >> 21) & 7u) - 2)
#define COUP_WORD_RETURN0(w) ((int)(((w) >> 24) & 7u) - 2)
uint32_t* coup_vec_state(coup_vec_env* env);                    /* [num_envs][4], layout above */
uint32_t* coup_vec_history(coup_vec_env* env);                  /* [num_envs][16], layout above */

/* ---- legal actions as a dense mask (State::LegalActionsMask, spiel.cc:371-377):
 * uint8[num_envs][18] on the device, 1 where legal. */
int coup_vec_legal_actions_mask(coup_vec_env* env, uint8_t* d_out, void* stream);

/* ---- tensors (CoupState::InformationStateTensor / ObservationTensor, coup.cc:1044-1056, i.e.
 * CoupObserver::WriteTensor coup.cc:248-287). d_out: dtype[rows][2492 or 98] on the device, fully
 * overwritten (the reference's ContiguousAllocator zero-fills first, observer.h:175-177).
 * Alignment: every tensor output (here, in the rollout calls and in the record decoders) must be aligned to the
 * four-element store unit of its dtype -- 16 bytes for f32, 8 for bf16, 4 for u8 -- else COUP_ERR_INVALID_ARG. Outputs
 * that are also 16-byte aligned (any fresh allocation) get the bulk-store encoder; others the plain-store one. */
int coup_vec_information_state_tensor(coup_vec_env* env, int player, int dtype, void* d_out, void* stream);
int coup_vec_observation_tensor(coup_vec_env* env, int player, int dtype, void* d_out, void* stream);
/* The general observer, CoupGame::MakeObserver(IIGObservationType{public_info, perfect_recall, private_info})
 * (coup.cc:1132-1141, 248-287; observer.h:270-315), for every env: private_info 0 kNone / 1 kSinglePlayer / 2 kAllPlayers
 * decides whose face-down cards show, public_info adds face-up cards, cur_move_player, cards_state, coins and then the
 * history (perfect_recall) or the last actions. Rows are contiguous: 2492 elements with public_info and perfect_recall,
 * 98 with public_info only, 42 without public_info. (1, 1, 1) is the info-state tensor, (1, 0, 1) the observation tensor. */
int coup_vec_observer_tensor(coup_vec_env* env, int player, int public_info, int perfect_recall, int private_info, int dtype,
                             void* d_out, void* stream);
/* Observation rows of a SUBSET of envs (row i, or rows 2i, 2i+1, describe env d_env_ids[i]). */
int coup_vec_observation_tensor_gather(coup_vec_env* env, const uint32_t* d_env_ids, uint32_t count, int player, int dtype,
                                       void* d_out, void* stream);
/* Same tensors with a padded row stride (in elements, a multiple of 4, >= 2492): elements
 * [2492, row_stride) of every row are written as zeros. row_stride 2496 keeps the consumer's first GEMM on
 * its aligned (fast) path: K = 2492 is not a multiple of 8 and costs a bf16 cuBLAS GEMM 5.6x on B200.
 * row_stride == COUP_LIVE_INFO_STATE_SIZE writes the live prefix of every row instead (see above); every entry point
 * that takes a row_stride accepts it, except coup_vec_rollout_incremental. */
int coup_vec_information_state_tensor_strided(coup_vec_env* env, int player, int dtype, void* d_out,
                                              uint32_t row_stride, void* stream);
int coup_vec_rollout_strided(coup_vec_env* env, int n_steps, int encode_player, int dtype, void* d_tensor_out,
                             uint32_t row_stride, void* stream);
/* Info-state tensors of a SUBSET of envs: output row i (rows 2i, 2i+1 for COUP_PLAYER_BOTH) describes env
 * d_env_ids[i] (uint32[count] on the device). Used to fetch, e.g., both players' terminal observations of the
 * envs that just finished (what every agent is stepped with at episode end, coup_experiments/scripts/nfsp.py:141-143). */
int coup_vec_information_state_tensor_gather(coup_vec_env* env, const uint32_t* d_env_ids, uint32_t count, int player,
                                             int dtype, void* d_out, uint32_t row_stride, void* stream);

/* ---- finished episodes (the `unreset_time_steps` of vector_env.SyncVectorEnv.step, python/vector_env.py:52-66, and
 * north_star's "every GPU trajectory's action and chance log", in the auto-reset mode that re-deals finished envs in
 * place). When enabled, every step/rollout kernel appends one packed record per episode that ends -- BEFORE the env is
 * re-dealt -- to a device ring of `capacity_records` (a power of two; 0 disables and frees) records:
 *   [0,16)  the finished episode's history (move i = 5-bit code at word i/6, bit 5*(i%6), as in coup_vec_history)
 *   [16,20) the TERMINAL packed state (layout above: Returns() = face-up counts, Rewards()[0] in w3)
 *   [20] env slot  [21] move_number_ | (Returns()[0]+2)<<8 | (Rewards()[0]+2)<<12 | truncated<<16
 *   [22],[23] low / high half of the Philox step counter of the step that ended the episode.
 * Slots are handed out by one atomic cursor, so records of one step appear in unspecified order. Works with and without
 * COUP_FLAG_AUTO_RESET (without it an episode is appended once, when it ends). Not part of snapshots.
 *   coup_vec_finished_ring / _ctrl: device pointers. ctrl is uint64[4]: [0] records ever appended, [1] the value of [0]
 *     when the most recent step/rollout call began, i.e. that call's episodes are ring positions [ctrl[1], ctrl[0]).
 *   coup_vec_finished_drain: copies the records not handed out yet (oldest first, at most max_records) to `out_records`
 *     (host or device memory) and advances the consumer cursor; *h_dropped_out = records lost so far because the
 *     producer lapped the consumer. Synchronises the stream.
 *   coup_vec_finished_information_state_tensor: the terminal info-state rows (what every agent is stepped with at
 *     episode end, coup_experiments/scripts/nfsp.py:141-143; python/algorithms/dqn.py:223-246) of the episodes that
 *     ended in the most recent step/rollout call, in ring order: row i (rows 2i, 2i+1 for COUP_PLAYER_BOTH) describes
 *     the i-th of them, d_env_ids_out[i] (may be NULL) is its env slot, *d_count_out (device, may be NULL) their number,
 *     clipped to max_episodes. The count never visits the host: surplus blocks of the launch exit. */
int coup_vec_finished_ring_enable(coup_vec_env* env, uint32_t capacity_records);
const uint32_t* coup_vec_finished_ring(const coup_vec_env* env);
const uint64_t* coup_vec_finished_ring_ctrl(const coup_vec_env* env);
uint32_t coup_vec_finished_ring_capacity(const coup_vec_env* env);
int coup_vec_finished_drain(coup_vec_env* env, void* out_records, uint32_t max_records, uint32_t* h_count_out,
                            uint64_t* h_dropped_out, void* stream);
int coup_vec_finished_information_state_tensor(coup_vec_env* env, int player, int dtype, void* d_out, uint32_t row_stride,
                                               uint32_t max_episodes, uint32_t* d_env_ids_out, uint32_t* d_count_out,
                                               void* stream);
int coup_vec_finished_observation_tensor(coup_vec_env* env, int player, int dtype, void* d_out, uint32_t max_episodes,
                                         uint32_t* d_env_ids_out, uint32_t* d_count_out, void* stream);
/* Info-state rows of ANY array of packed records on the handle's device: output row i decodes record d_indices[i]
 * (d_indices NULL: record i). `player` may be COUP_PLAYER_FROM_RECORD. This is how compact replay / reservoir records
 * (80 bytes instead of a 2492-element row) turn back into network inputs when a batch is sampled. */
int coup_records_information_state_tensor(coup_vec_env* env, const uint32_t* d_records, const uint32_t* d_indices,
                                          uint32_t count, int player, int dtype, void* d_out, uint32_t row_stride, void* stream);
int coup_records_observation_tensor(coup_vec_env* env, const uint32_t* d_records, const uint32_t* d_indices, uint32_t count,
                                    int player, int dtype, void* d_out, void* stream);

/* ---- self-play recording fused into the step ------------------------------------------------------------------------
 * coup_vec_step_record = coup_vec_step that also keeps what the reference's learning agents keep per decision, without a
 * host round trip and without materialising tensors (the caller owns all buffers, device memory, zero-initialised):
 *   NFSP supervised-learning reservoir (python/algorithms/nfsp.py:226-242,322-371): every env that is about to act offers
 *     Transition(info_state, action_probs, legal_actions_mask). The element with running index t (= reservoir_offered +
 *     env slot; the caller adds num_envs to reservoir_offered after every call) lands in slot t while the buffer fills,
 *     afterwards in slot randint(0, t) if that is < capacity; of several elements drawing one slot in the same call the
 *     later one wins, as it would sequentially. d_reservoir_records[slot] = packed record of the state (meta: env slot,
 *     seat<<31 | legal mask, t lo, t hi), d_reservoir_probs[slot] = the 18 probabilities given in d_action_probs.
 *   DQN replay (python/algorithms/dqn.py:30-32,223-246; the loop of coup_experiments/scripts/nfsp.py:134-144): per
 *     (env, seat) the previous decision is kept in d_pending; when that seat acts again, and for BOTH seats when the
 *     episode ends, Transition(info_state, action, reward, next_info_state, is_final_step, legal_actions_mask) is appended
 *     at position (*d_replay_total)++ % replay_capacity of d_transitions as two packed records: [0] the earlier state
 *     (meta: env slot, seat<<31 | action | (reward+2)<<5 | is_final<<8, ticket lo, ticket hi), [1] the later state (meta:
 *     env slot, seat<<31 | legal mask of the later state). Rewards that fall between a seat's own turns are dropped, as
 *     in the reference. Order of the transitions of one call is unspecified.
 * Records turn into info-state rows with coup_records_information_state_tensor(..., COUP_PLAYER_FROM_RECORD, ...).
 * Either half may be disabled with NULL pointers. Works with COUP_FLAG_AUTO_RESET (finished envs are re-dealt in place;
 * their terminal observations go into the final transitions first). */
typedef struct coup_recorder_buffers {
  uint32_t* d_reservoir_records;  /* [reservoir_capacity][COUP_RECORD_WORDS] or NULL */
  float* d_reservoir_probs;       /* [reservoir_capacity][18] */
  uint64_t* d_reservoir_winner;   /* [reservoir_capacity] scratch, zero-initialised once */
  uint64_t reservoir_capacity;
  uint64_t reservoir_offered;     /* elements offered by earlier calls */
  uint32_t* d_transitions;        /* [replay_capacity][2][COUP_RECORD_WORDS] or NULL */
  uint64_t replay_capacity;
  uint64_t* d_replay_total;       /* one device counter */
  uint32_t* d_pending;            /* [num_envs][2][COUP_RECORD_WORDS], zero-initialised once */
} coup_recorder_buffers;
int coup_vec_step_record(coup_vec_env* env, const uint8_t* d_actions, const float* d_action_probs,
                         const coup_recorder_buffers* buffers, void* stream);

/* ---- host-buffer convenience path (what a host-driven caller such as rl_environment would use):
 * copies uint8[num_envs] actions from (pinned) host memory, steps, optionally encodes the current
 * player's info-state into d_tensor_out (device; may be NULL) and copies legal_mask / current_player /
 * done / rewards back into the host buffers (each may be NULL). Returns once the host buffers are filled;
 * the tensor is complete in stream order (work queued on `stream` afterwards sees it), so the host can
 * choose the next actions while the encoder is still writing. */
int coup_vec_step_host(coup_vec_env* env, const uint8_t* h_actions, uint32_t* h_legal_mask,
                       int8_t* h_current_player, uint8_t* h_done, int8_t* h_rewards, int dtype,
                       void* d_tensor_out, void* stream);

/* Same, with ONE device->host copy: h_step_words receives coup_vec_step_word() (uint32[num_envs]). */
int coup_vec_step_host_packed(coup_vec_env* env, const uint8_t* h_actions, uint32_t* h_step_words, int dtype,
                              void* d_tensor_out, void* stream);

/* The same call in two halves, for callers that drive several handles (sub-slabs, one stream each) from one host thread:
 * _async queues the copies, the step and the encoder and returns at once; coup_vec_host_outputs_wait blocks until the
 * h_step_words of the handle's LAST _async call are filled. Waiting for a sub-slab only right before its next actions are
 * chosen keeps the host off the device's critical path (bench.py's e2e leg). The host buffers must stay valid until then. */
int coup_vec_step_host_packed_async(coup_vec_env* env, const uint8_t* h_actions, uint32_t* h_step_words, int dtype,
                                    void* d_tensor_out, void* stream);
int coup_vec_host_outputs_wait(coup_vec_env* env);

/* Host-side uniform-random POLICY for callers that keep their policy on the host (the host analogue of
 * benchmark_game.cc:96-99): for every env picks the k-th set bit of h_legal_mask[i], k drawn from the same
 * Philox stream coup_vec_sample_uniform would use at step counter `step`; 0xFF where the mask is 0.
 * Only bits 0-17 of each word are read, so packed step words can be passed directly.
 * Pure bit selection on host memory with `threads` host threads -- no game rule is evaluated. */
int coup_host_sample_uniform(const uint32_t* h_legal_mask, uint32_t n, uint64_t seed, uint64_t global_env_offset,
                             uint64_t step, uint8_t* h_actions_out, int threads);

/* ---- statistics, errors, verification ---------------------------------------------------------- */
int coup_vec_stats(coup_vec_env* env, uint64_t* h_out /* [COUP_STATS_LEN] */, void* stream);
uint64_t* coup_vec_stats_device(coup_vec_env* env); /* uint64[COUP_STATS_LEN] for an NCCL all-reduce */
int coup_vec_clear_stats(coup_vec_env* env, void* stream);
int coup_vec_check_errors(coup_vec_env* env, void* stream); /* COUP_OK or COUP_ERR_ILLEGAL_ACTION */
/* 64-bit position-keyed hash of every row of a dense float tensor already in device memory
 * (h = sum over non-zero t[i] of mix64(i << 32 | bits(t[i]))); used to compare 10^6-trajectory runs
 * against the oracle without moving the tensors. d_hash_out: uint64[rows]. */
int coup_tensor_row_hash(const void* d_tensor, int dtype, uint32_t rows, uint32_t row_len,
                         uint64_t* d_hash_out, void* stream);

/* Checkpoint / resume of a whole slab (the reference checkpoints a game as its action history,
 * State::Serialize spiel.cc:297-311; here a snapshot is the raw packed state + history + outputs + statistics +
 * the Philox step counter, so a restored handle continues bit-identically). Host buffers. */
size_t coup_vec_snapshot_size(const coup_vec_env* env);
int coup_vec_snapshot(coup_vec_env* env, void* h_buf, size_t bytes, void* stream);
int coup_vec_restore(coup_vec_env* env, const void* h_buf, size_t bytes, void* stream);

/* Global step counter that keys the Philox streams (snapshot = state + history + this counter). */
uint64_t coup_vec_step_counter(const coup_vec_env* env);
int coup_vec_set_step_counter(coup_vec_env* env, uint64_t value);

#ifdef __cplusplus
}
#endif
#endif /* COUP_B200_H_ */
