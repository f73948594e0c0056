/* ORACLE -- TEST INFRASTRUCTURE ONLY (see coup_oracle.h). Plain-C CPU restatement of the reference's
 * Coup path; each function cites the reference file:line it follows. Never used by the product. */
#define _POSIX_C_SOURCE 200809L
#include "coup_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

int oc_sizeof_state(void) { return (int)sizeof(oc_state); }

/* ------------------------------------------------------------------------------------------------
 * Hand helpers
 * ---------------------------------------------------------------------------------------------- */

/* CoupCard::operator<, coup.h:91-94: order by value, then state with FaceDown(0) < FaceUp(1). */
static int card_less(const oc_card* a, const oc_card* b) {
  return a->value < b->value || (a->value == b->value && a->state < b->state);
}

/* CoupPlayer::SortCards, coup.cc:389-391 (std::sort; equal cards are indistinguishable, so any
 * correct sort gives the same sequence). Insertion sort over <= 4 cards. */
static void sort_cards(oc_player* p) {
  for (int i = 1; i < p->num_cards; ++i) {
    oc_card c = p->cards[i];
    int j = i - 1;
    while (j >= 0 && card_less(&c, &p->cards[j])) {
      p->cards[j + 1] = p->cards[j];
      --j;
    }
    p->cards[j + 1] = c;
  }
}

/* CoupPlayer::HasFaceDownCard, coup.cc:379-387 */
static int has_face_down_card(const oc_player* p, int card) {
  for (int i = 0; i < p->num_cards; ++i)
    if (p->cards[i].value == card && p->cards[i].state == OC_FACE_DOWN) return 1;
  return 0;
}

/* std::vector::erase(begin()+idx) */
static void erase_card(oc_player* p, int idx) {
  for (int k = idx; k + 1 < p->num_cards; ++k) p->cards[k] = p->cards[k + 1];
  p->num_cards--;
}

static void queue_push(oc_state* s, int player) {
  if (s->deal_tail >= 8) { s->error = OC_ERR_INTERNAL; return; }
  s->deal_queue[s->deal_tail++] = player;
}
static int queue_size(const oc_state* s) { return s->deal_tail - s->deal_head; }

/* ------------------------------------------------------------------------------------------------
 * Construction and scalar queries
 * ---------------------------------------------------------------------------------------------- */

/* CoupState::CoupState, coup.cc:393-428: deck 3 of each of 5 cards; P1 has 1 coin and P2 has 2;
 * no last action; deal queue [0,1,0,1]; starts at a chance node with P1 to move afterwards. */
void oc_init(oc_state* s) {
  memset(s, 0, sizeof(*s));
  for (int i = 0; i < OC_NUM_CARD_TYPES; ++i) s->deck[i] = OC_NUM_EACH_CARD;
  s->players[0].coins = 1;
  s->players[1].coins = 2;
  for (int p = 0; p < OC_NUM_PLAYERS; ++p) {
    s->players[p].last_action = OC_NONE;
    s->players[p].lost_challenge = 0;
    s->players[p].num_cards = 0;
  }
  s->cur_player_turn = 0;
  s->cur_player_move = 0;
  s->opp_player = 1;
  s->is_turn_begin = 1;
  s->turn_number = 0;
  s->is_chance = 1;
  queue_push(s, 0);
  queue_push(s, 1);
  queue_push(s, 0);
  queue_push(s, 1);
  memset(s->history_deal_player, -1, sizeof(s->history_deal_player));
}

/* CoupState::IsTerminal, coup.cc:989-1010: over the move cap (strictly greater than
 * MaxGameLength), or fewer than two players alive where a player holding < 2 cards (mid
 * replacement) counts as alive and otherwise a player is alive iff some card is face down. */
int oc_is_terminal(const oc_state* s) {
  if (s->move_number > OC_MAX_GAME_LENGTH) return 1;
  int alive = 0;
  for (int p = 0; p < OC_NUM_PLAYERS; ++p) {
    const oc_player* pl = &s->players[p];
    if (pl->num_cards < 2) { alive += 1; continue; }
    for (int i = 0; i < pl->num_cards; ++i)
      if (pl->cards[i].state == OC_FACE_DOWN) { alive += 1; break; }
  }
  return alive > 1 ? 0 : 1;
}

/* CoupState::CurrentPlayer, coup.cc:458-466 */
int oc_current_player(const oc_state* s) {
  if (oc_is_terminal(s)) return OC_TERMINAL_PLAYER;
  if (s->is_chance) return OC_CHANCE_PLAYER;
  return s->cur_player_move;
}

int oc_is_chance_node(const oc_state* s) { return oc_current_player(s) == OC_CHANCE_PLAYER; }

/* CoupState::NextPlayerTurn, coup.cc:1079-1086 */
static void next_player_turn(oc_state* s) {
  s->cur_player_turn = 1 - s->cur_player_turn;
  s->cur_player_move = s->cur_player_turn;
  s->opp_player = 1 - s->cur_player_move;
  s->turn_number += 1;
  s->is_turn_begin = 1;
}

/* CoupState::NextPlayerMove, coup.cc:1088-1092 */
static void next_player_move(oc_state* s) {
  s->cur_player_move = 1 - s->cur_player_move;
  s->opp_player = 1 - s->cur_player_move;
  s->is_turn_begin = 0;
}

/* CoupState::ChallengeFailReplaceCard, coup.cc:468-486: the FIRST face-down card of that type in the
 * opponent's (sorted) hand goes back to the deck and a replacement deal is queued for them. */
static void challenge_fail_replace_card(oc_state* s, int card) {
  oc_player* op = &s->players[s->opp_player];
  for (int i = 0; i < op->num_cards; ++i) {
    if (op->cards[i].value == card && op->cards[i].state == OC_FACE_DOWN) {
      s->deck[card] += 1;
      erase_card(op, i);
      queue_push(s, s->opp_player);
      s->is_chance = 1;
      return;
    }
  }
  s->error = OC_ERR_INTERNAL; /* "Tried to replace card which was not found in hand" */
}

/* ------------------------------------------------------------------------------------------------
 * Transition
 * ---------------------------------------------------------------------------------------------- */

/* CoupState::DoApplyAction, coup.cc:490-809. Recursive exactly where the reference is (628, 718). */
static void do_apply_action(oc_state* s, int move) {
  if (oc_is_chance_node(s)) {
    /* coup.cc:491-520 */
    if (move < 0 || move >= OC_NUM_CARD_TYPES || s->deck[move] <= 0 || queue_size(s) <= 0) {
      s->error = OC_ERR_ILLEGAL;
      return;
    }
    int deal_to = s->deal_queue[s->deal_head++];
    if (s->deal_head == s->deal_tail) s->deal_head = s->deal_tail = 0;
    /* history_chance_deal_player_.insert({history_.size(), dealToPlayer}), coup.cc:504 */
    s->history_deal_player[s->history_len] = (int8_t)deal_to;
    s->deck[move] -= 1;
    oc_player* pl = &s->players[deal_to];
    if (pl->num_cards >= OC_MAX_CARDS_IN_HAND) { s->error = OC_ERR_INTERNAL; return; }
    pl->cards[pl->num_cards].value = move;
    pl->cards[pl->num_cards].state = OC_FACE_DOWN;
    pl->num_cards++;
    sort_cards(pl);
    if (queue_size(s) == 0) s->is_chance = 0;
    return;
  }

  /* coup.cc:522-529 */
  oc_player* cp = &s->players[s->cur_player_move];
  oc_player* op = &s->players[s->opp_player];
  const int mover = s->cur_player_move;
  const int other = s->opp_player;
  s->cur_rewards[0] = 0;
  s->cur_rewards[1] = 0;
  const int action = move;

  if (action == OC_INCOME) {                                    /* 531-534 */
    cp->last_action = action;
    cp->coins += 1;
    next_player_turn(s);

  } else if (action == OC_FOREIGN_AID) {                        /* 536-546 */
    if (s->is_turn_begin) {
      cp->last_action = action;
      next_player_move(s);
    } else {
      cp->coins += 2;
      next_player_turn(s);
    }

  } else if (action == OC_COUP) {                               /* 548-553 */
    if (cp->coins < 7) { s->error = OC_ERR_ILLEGAL; return; }
    cp->last_action = action;
    cp->coins -= 7;
    next_player_move(s);

  } else if (action == OC_TAX) {                                /* 555-565 */
    if (s->is_turn_begin) {
      cp->last_action = action;
      next_player_move(s);
    } else {
      cp->coins += 3;
      next_player_turn(s);
    }

  } else if (action == OC_ASSASSINATE) {                        /* 567-573 */
    if (cp->coins < 3) { s->error = OC_ERR_ILLEGAL; return; }
    cp->last_action = action;
    cp->coins -= 3; /* paid whether or not the action is blocked/challenged */
    next_player_move(s);

  } else if (action == OC_EXCHANGE) {                           /* 575-587 */
    if (s->is_turn_begin) {
      cp->last_action = action;
      next_player_move(s);
    } else {
      queue_push(s, s->cur_player_move);
      queue_push(s, s->cur_player_move);
      s->is_chance = 1;
    }

  } else if (action == OC_STEAL) {                              /* 589-603 */
    if (op->coins < 1) { s->error = OC_ERR_ILLEGAL; return; }
    if (s->is_turn_begin) {
      cp->last_action = action;
      next_player_move(s);
    } else {
      int num_steal = (op->coins > 1) ? 2 : 1;
      cp->coins += num_steal;
      op->coins -= num_steal;
      next_player_turn(s);
    }

  } else if (action == OC_LOSE_CARD_1 || action == OC_LOSE_CARD_2) { /* 605-616 */
    int card_to_lose = move - OC_LOSE_CARD_1;
    if (card_to_lose >= cp->num_cards || cp->cards[card_to_lose].state != OC_FACE_DOWN) {
      s->error = OC_ERR_ILLEGAL;
      return;
    }
    cp->last_action = action;
    cp->cards[card_to_lose].state = OC_FACE_UP;
    cp->lost_challenge = 0;
    sort_cards(cp);
    s->cur_rewards[mover] -= 1;
    s->cur_rewards[other] += 1;
    next_player_turn(s);

  } else if (action == OC_PASS) {                               /* 618-629 */
    cp->last_action = action;
    int n_act = op->last_action;
    if (n_act == OC_BLOCK) {
      next_player_turn(s);
    } else {
      next_player_move(s);
      do_apply_action(s, n_act); /* Pass, so complete their action */
    }

  } else if (action == OC_BLOCK) {                              /* 631-633 */
    cp->last_action = action;
    next_player_move(s);

  } else if (action == OC_CHALLENGE) {                          /* 635-771 */
    if (op->last_action == OC_BLOCK) {
      if (cp->last_action == OC_FOREIGN_AID) {                  /* 637-649 */
        cp->last_action = action;
        if (has_face_down_card(op, OC_DUKE)) {
          cp->lost_challenge = 1;
          challenge_fail_replace_card(s, OC_DUKE);
        } else {
          op->lost_challenge = 1;
          cp->coins += 2;
          next_player_move(s);
        }
      } else if (cp->last_action == OC_ASSASSINATE) {           /* 650-670 */
        cp->last_action = action;
        if (has_face_down_card(op, OC_CONTESSA)) {
          cp->lost_challenge = 1;
          challenge_fail_replace_card(s, OC_CONTESSA);
        } else {
          /* op loses the game: one card for the assassination, one for the lost challenge */
          for (int k = 0; k < 2; ++k) {
            if (k < op->num_cards && op->cards[k].state == OC_FACE_DOWN) {
              op->cards[k].state = OC_FACE_UP;
              s->cur_rewards[mover] += 1;
              s->cur_rewards[other] -= 1;
            }
          }
        }
      } else if (cp->last_action == OC_STEAL) {                 /* 671-690 */
        cp->last_action = action;
        if (has_face_down_card(op, OC_CAPTAIN)) {
          cp->lost_challenge = 1;
          challenge_fail_replace_card(s, OC_CAPTAIN);
        } else if (has_face_down_card(op, OC_AMBASSADOR)) {
          cp->lost_challenge = 1;
          challenge_fail_replace_card(s, OC_AMBASSADOR);
        } else {
          op->lost_challenge = 1;
          int num_steal = (op->coins > 1) ? 2 : 1;
          cp->coins += num_steal;
          op->coins -= num_steal;
          next_player_move(s);
        }
      } else {
        s->error = OC_ERR_ILLEGAL; /* "Invalid player action", 692 */
      }
    } else if (op->last_action == OC_TAX) {                     /* 694-706 */
      cp->last_action = action;
      if (has_face_down_card(op, OC_DUKE)) {
        cp->lost_challenge = 1;
        challenge_fail_replace_card(s, OC_DUKE);
        op->coins += 3; /* complete the action now */
      } else {
        op->lost_challenge = 1;
        next_player_move(s);
      }
    } else if (op->last_action == OC_EXCHANGE) {                /* 708-725 */
      cp->last_action = action;
      if (has_face_down_card(op, OC_AMBASSADOR)) {
        cp->lost_challenge = 1;
        challenge_fail_replace_card(s, OC_AMBASSADOR);
        s->is_chance = 0; /* momentarily, so that the Exchange below runs as a player move */
        next_player_move(s);
        do_apply_action(s, OC_EXCHANGE);
      } else {
        op->lost_challenge = 1;
        next_player_move(s);
      }
    } else if (op->last_action == OC_ASSASSINATE) {             /* 727-749 */
      cp->last_action = action;
      if (has_face_down_card(op, OC_ASSASSIN)) {
        /* cp loses the game: one card for the assassination, one for the lost challenge */
        for (int k = 0; k < 2; ++k) {
          if (k < cp->num_cards && cp->cards[k].state == OC_FACE_DOWN) {
            cp->cards[k].state = OC_FACE_UP;
            s->cur_rewards[mover] -= 1;
            s->cur_rewards[other] += 1;
          }
        }
      } else {
        op->lost_challenge = 1;
        op->coins += 3; /* coins spent are returned in this one case */
        next_player_move(s);
      }
    } else if (op->last_action == OC_STEAL) {                   /* 751-767 */
      cp->last_action = action;
      if (has_face_down_card(op, OC_CAPTAIN)) {
        cp->lost_challenge = 1;
        challenge_fail_replace_card(s, OC_CAPTAIN);
        int num_steal = (cp->coins > 1) ? 2 : 1; /* the steal completes now, against the challenger */
        op->coins += num_steal;
        cp->coins -= num_steal;
      } else {
        op->lost_challenge = 1;
        next_player_move(s);
      }
    } else {
      s->error = OC_ERR_ILLEGAL; /* "Invalid player action", 770 */
    }

  } else if (move >= OC_EXCHANGE_RETURN_12 && move <= OC_EXCHANGE_RETURN_34) { /* 773-803 */
    cp->last_action = action;
    int card_ind[4];
    int n_ind = 0;
    if (move <= OC_EXCHANGE_RETURN_14) card_ind[n_ind++] = 0;
    if (action == OC_EXCHANGE_RETURN_12 || action == OC_EXCHANGE_RETURN_23 ||
        action == OC_EXCHANGE_RETURN_24) card_ind[n_ind++] = 1;
    if (action == OC_EXCHANGE_RETURN_13 || action == OC_EXCHANGE_RETURN_23 ||
        action == OC_EXCHANGE_RETURN_34) card_ind[n_ind++] = 2;
    if (action == OC_EXCHANGE_RETURN_14 || action == OC_EXCHANGE_RETURN_24 ||
        action == OC_EXCHANGE_RETURN_34) card_ind[n_ind++] = 3;
    if (n_ind != 2 || cp->num_cards != 4) { s->error = OC_ERR_ILLEGAL; return; }
    for (int i = 1; i >= 0; --i) {
      int c = card_ind[i];
      erase_card(cp, c);
      /* REFERENCE QUIRK, coup.cc:789-795: deck_ is indexed by the HAND SLOT c, not by the value
       * of the card that was in that slot. Reproduced on purpose (bit-exact replay). */
      s->deck[c] += 1;
    }
    if (op->lost_challenge) {
      next_player_move(s);
    } else {
      next_player_turn(s);
    }

  } else {
    s->error = OC_ERR_ILLEGAL; /* "Invalid player action", 806 */
  }
}

/* State::ApplyAction, spiel.cc:322-332: player = CurrentPlayer(); DoApplyAction(a);
 * history_.push_back({player, a}); ++move_number_. */
int oc_apply_action(oc_state* s, int action) {
  if (s->error) return s->error;
  int player = oc_current_player(s);
  if (player == OC_TERMINAL_PLAYER) return OC_ERR_TERMINAL;
  if (s->history_len >= OC_HIST_CAP) { s->error = OC_ERR_INTERNAL; return s->error; }
  do_apply_action(s, action);
  if (s->error) return s->error;
  s->history_player[s->history_len] = (int8_t)player;
  s->history_action[s->history_len] = (int8_t)action;
  s->history_len++;
  s->move_number++;
  return OC_OK;
}

int oc_apply_action_checked(oc_state* s, int action) {
  if (s->error) return s->error;
  if (oc_is_terminal(s)) return OC_ERR_TERMINAL;
  if (action < 0 || action >= 32 || !((oc_legal_mask(s) >> action) & 1u)) return OC_ERR_ILLEGAL;
  return oc_apply_action(s, action);
}

/* ------------------------------------------------------------------------------------------------
 * Legal actions
 * ---------------------------------------------------------------------------------------------- */

/* CoupState::LegalLoseCardActions, coup.cc:811-822 */
static int legal_lose_card_actions(const oc_state* s, int32_t* out) {
  int n = 0;
  const oc_player* p = &s->players[s->cur_player_move];
  if (p->num_cards > 0 && p->cards[0].state == OC_FACE_DOWN) out[n++] = OC_LOSE_CARD_1;
  if (p->num_cards > 1 && p->cards[1].state == OC_FACE_DOWN) out[n++] = OC_LOSE_CARD_2;
  return n;
}

/* CoupState::LegalActions, coup.cc:824-938. Returns the count; ids are ascending as in the
 * reference. -1 where the reference would SpielFatalError. */
int oc_legal_actions(const oc_state* s, int32_t* out) {
  int n = 0;
  if (oc_is_terminal(s)) return 0;
  if (oc_is_chance_node(s)) {                                   /* 828-836 */
    for (int i = 0; i < OC_NUM_CARD_TYPES; ++i)
      if (s->deck[i] > 0) out[n++] = i;
    return n;
  }
  const oc_player* cp = &s->players[s->cur_player_move];
  const oc_player* op = &s->players[s->opp_player];
  if (s->is_turn_begin) {                                       /* 841-854 */
    if (cp->coins >= 10) { out[n++] = OC_COUP; return n; }
    out[n++] = OC_INCOME;
    out[n++] = OC_FOREIGN_AID;
    if (cp->coins >= 7) out[n++] = OC_COUP;
    out[n++] = OC_TAX;
    if (cp->coins >= 3) out[n++] = OC_ASSASSINATE;
    out[n++] = OC_EXCHANGE;
    if (op->coins > 0) out[n++] = OC_STEAL;
    return n;
  } else if (cp->lost_challenge) {                              /* 856-858 */
    return legal_lose_card_actions(s, out);
  } else if (s->cur_player_move != s->cur_player_turn) {        /* 860-887 */
    if (op->last_action == OC_FOREIGN_AID) {
      out[n++] = OC_PASS; out[n++] = OC_BLOCK;
      return n;
    } else if (op->last_action == OC_TAX || op->last_action == OC_EXCHANGE) {
      out[n++] = OC_PASS; out[n++] = OC_CHALLENGE;
      return n;
    } else if (op->last_action == OC_STEAL) {
      out[n++] = OC_PASS; out[n++] = OC_BLOCK; out[n++] = OC_CHALLENGE;
      return n;
    } else if (op->last_action == OC_ASSASSINATE) {
      n = legal_lose_card_actions(s, out);
      out[n++] = OC_BLOCK; out[n++] = OC_CHALLENGE;
      return n;
    } else if (op->last_action == OC_COUP) {
      return legal_lose_card_actions(s, out);
    }
    return -1;
  } else if (cp->last_action == OC_EXCHANGE) {                  /* 889-928 */
    if (cp->num_cards < 4) return -1;
    int face_up_ind = -1;
    for (int i = 0; i < cp->num_cards; ++i)
      if (cp->cards[i].state == OC_FACE_UP) { face_up_ind = i; break; }
    static const int32_t kAll[6] = {12, 13, 14, 15, 16, 17};
    static const int32_t kUp0[3] = {15, 16, 17};  /* 23, 24, 34 */
    static const int32_t kUp1[3] = {13, 14, 17};  /* 13, 14, 34 */
    static const int32_t kUp2[3] = {12, 14, 16};  /* 12, 14, 24 */
    static const int32_t kUp3[3] = {12, 13, 15};  /* 12, 13, 23 */
    const int32_t* src = kAll;
    int cnt = 6;
    if (face_up_ind == 0) { src = kUp0; cnt = 3; }
    else if (face_up_ind == 1) { src = kUp1; cnt = 3; }
    else if (face_up_ind == 2) { src = kUp2; cnt = 3; }
    else if (face_up_ind == 3) { src = kUp3; cnt = 3; }
    for (int i = 0; i < cnt; ++i) out[n++] = src[i];
    return n;
  } else if (op->last_action == OC_BLOCK) {                     /* 930-933 */
    out[n++] = OC_PASS; out[n++] = OC_CHALLENGE;
    return n;
  }
  return -1;                                                    /* 935-937 */
}

/* State::LegalActionsMask, spiel.cc:371-377, folded into one bitmask (bit a <=> a is legal). */
uint32_t oc_legal_mask(const oc_state* s) {
  int32_t la[OC_NUM_ACTIONS];
  int n = oc_legal_actions(s, la);
  uint32_t m = 0;
  for (int i = 0; i < n; ++i) m |= 1u << la[i];
  return m;
}

/* CoupState::ChanceOutcomes, coup.cc:1062-1077 */
int oc_chance_outcomes(const oc_state* s, int32_t* actions, double* probs) {
  if (!oc_is_chance_node(s)) return -1;
  double deck_size = 0;
  for (int i = 0; i < OC_NUM_CARD_TYPES; ++i) deck_size += s->deck[i];
  int n = 0;
  for (int i = 0; i < OC_NUM_CARD_TYPES; ++i) {
    if (s->deck[i] > 0) {
      actions[n] = i;
      probs[n] = s->deck[i] / deck_size;
      n++;
    }
  }
  return n;
}

/* CoupState::Returns, coup.cc:1016-1032 */
void oc_returns(const oc_state* s, double* out) {
  int face_up[2] = {0, 0};
  for (int p = 0; p < OC_NUM_PLAYERS; ++p)
    for (int i = 0; i < s->players[p].num_cards; ++i)
      if (s->players[p].cards[i].state == OC_FACE_UP) face_up[p] += 1;
  out[0] = face_up[1] - face_up[0];
  out[1] = face_up[0] - face_up[1];
}

/* CoupState::Rewards, coup.cc:1012-1014 */
void oc_rewards(const oc_state* s, double* out) {
  out[0] = s->cur_rewards[0];
  out[1] = s->cur_rewards[1];
}

/* ------------------------------------------------------------------------------------------------
 * Tensor observers
 * ---------------------------------------------------------------------------------------------- */

/* CoupObserver::WriteTensor, coup.cc:248-287, with the ContiguousAllocator laying the named
 * tensors out back to back in Get() order (observer.h:173-184, observer.cc:28-36) after zero-filling
 * the whole span (observer.h:175-177). */
int oc_observer_tensor(const oc_state* s, int player, int public_info, int perfect_recall,
                       int private_info, float* out) {
  int size = 2 + 2 * OC_MAX_CARDS_IN_HAND * OC_NUM_CARD_TYPES;
  if (public_info) {
    size += 2 + 2 * OC_MAX_CARDS_IN_HAND * 2 + 2;
    size += perfect_recall ? OC_MAX_MOVE_NUMBER * OC_NUM_ACTIONS : 2 * OC_NUM_ACTIONS;
  }
  memset(out, 0, sizeof(float) * (size_t)size);
  int off = 0;
  /* WritePlayer(state, player, "", ...), coup.cc:160-165, 255 */
  out[off + player] = 1;
  off += 2;
  /* WritePlayerCardsValue for p1 then p2, coup.cc:178-191, 258-265 */
  for (int p = 0; p < OC_NUM_PLAYERS; ++p) {
    int priv = (private_info == 2) || (private_info == 1 && p == player);
    const oc_player* pl = &s->players[p];
    for (int i = 0; i < pl->num_cards; ++i) {
      const oc_card* c = &pl->cards[i];
      if (c->value != -1 && ((priv && c->state == OC_FACE_DOWN) || (public_info && c->state == OC_FACE_UP)))
        out[off + i * OC_NUM_CARD_TYPES + c->value] = 1;
    }
    off += OC_MAX_CARDS_IN_HAND * OC_NUM_CARD_TYPES;
  }
  if (public_info) {
    /* cur_move_player, coup.cc:269-276: all zero at terminal, else one-hot(cur_player_move_)
     * (also at chance nodes). */
    if (!oc_is_terminal(s)) out[off + s->cur_player_move] = 1;
    off += 2;
    /* WriteCardsState, coup.cc:194-204: [player][slot][FaceDown, FaceUp] */
    for (int p = 0; p < OC_NUM_PLAYERS; ++p)
      for (int i = 0; i < s->players[p].num_cards; ++i)
        out[off + (p * OC_MAX_CARDS_IN_HAND + i) * 2 + s->players[p].cards[i].state] = 1;
    off += 2 * OC_MAX_CARDS_IN_HAND * 2;
    /* WriteCoins, coup.cc:207-213: raw counts, not one-hot */
    for (int p = 0; p < OC_NUM_PLAYERS; ++p) out[off + p] = (float)s->players[p].coins;
    off += 2;
    if (perfect_recall) {
      /* WriteActionHistory, coup.cc:230-245 */
      for (int i = 0; i < s->history_len; ++i) {
        int p = s->history_player[i];
        if (p >= 0 || (p == OC_CHANCE_PLAYER && s->history_deal_player[i] == player)) {
          int a = s->history_action[i];
          if (a != OC_NONE) out[off + i * OC_NUM_ACTIONS + a] = 1;
        }
      }
      off += OC_MAX_MOVE_NUMBER * OC_NUM_ACTIONS;
    } else {
      /* WriteLastAction, coup.cc:217-225 */
      for (int p = 0; p < OC_NUM_PLAYERS; ++p) {
        int a = s->players[p].last_action;
        if (a != OC_NONE) out[off + p * OC_NUM_ACTIONS + a] = 1;
      }
      off += 2 * OC_NUM_ACTIONS;
    }
  }
  return off;
}

/* CoupState::InformationStateTensor, coup.cc:1044-1049 with kInfoStateObsType
 * (observer.h:287-291: public_info, perfect_recall, private_info = single player). */
void oc_information_state_tensor(const oc_state* s, int player, float* out) {
  oc_observer_tensor(s, player, 1, 1, 1, out);
}

/* CoupState::ObservationTensor, coup.cc:1051-1056 with kDefaultObsType
 * (observer.h:281-285: public_info, no perfect recall, private_info = single player). */
void oc_observation_tensor(const oc_state* s, int player, float* out) {
  oc_observer_tensor(s, player, 1, 0, 1, out);
}

/* Same hash as TensorHash in oracle/ref_harness.cc. */
static uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}
uint64_t oc_tensor_hash(const float* t, int n) {
  uint64_t h = 0;
  for (int i = 0; i < n; ++i) {
    uint32_t bits;
    memcpy(&bits, &t[i], 4);
    if (bits != 0) h += mix64(((uint64_t)i << 32) | bits);
  }
  return h;
}

/* ------------------------------------------------------------------------------------------------
 * Trajectory tracing (same record layout as ref_trace in ref_harness.cc)
 * ---------------------------------------------------------------------------------------------- */

static void fill_rec(const oc_state* s, oc_trace_rec* r) {
  float info[OC_INFO_STATE_SIZE];
  float obs[OC_OBSERVATION_SIZE];
  double d[2];
  memset(r, 0, sizeof(*r));
  r->cur_player = (int8_t)oc_current_player(s);
  r->is_terminal = (uint8_t)oc_is_terminal(s);
  r->is_chance = (uint8_t)oc_is_chance_node(s);
  r->move_number = (uint8_t)s->move_number;
  r->legal_mask = r->is_terminal ? 0u : oc_legal_mask(s);
  oc_rewards(s, d);
  r->rewards[0] = (int8_t)d[0]; r->rewards[1] = (int8_t)d[1];
  oc_returns(s, d);
  r->returns[0] = (int8_t)d[0]; r->returns[1] = (int8_t)d[1];
  for (int p = 0; p < 2; ++p) {
    r->coins[p] = (uint8_t)s->players[p].coins;
    r->ncards[p] = (uint8_t)s->players[p].num_cards;
    oc_information_state_tensor(s, p, info);
    oc_observation_tensor(s, p, obs);
    r->hash_info[p] = oc_tensor_hash(info, OC_INFO_STATE_SIZE);
    r->hash_obs[p] = oc_tensor_hash(obs, OC_OBSERVATION_SIZE);
  }
}

int oc_trace(const uint8_t* actions, int n_actions, oc_trace_rec* out) {
  oc_state s;
  oc_init(&s);
  fill_rec(&s, &out[0]);
  for (int i = 0; i < n_actions; ++i) {
    if (oc_apply_action(&s, actions[i]) != OC_OK) return -(i + 1);
    fill_rec(&s, &out[i + 1]);
  }
  return n_actions + 1;
}

int oc_trace_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, oc_trace_rec* out) {
  int bad = 0;
  for (int t = 0; t < n_traj; ++t) {
    int n = (int)(offsets[t + 1] - offsets[t]);
    if (oc_trace(actions + offsets[t], n, out + offsets[t] + t) < 0) bad++;
  }
  return bad;
}

/* Only the record of the LAST state of each trajectory (out[t]); a rejected trajectory gets
 * cur_player = 127. Returns the number of rejected trajectories. */
int oc_final_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, oc_trace_rec* out) {
  int bad = 0;
  for (int t = 0; t < n_traj; ++t) {
    oc_state s;
    int n = (int)(offsets[t + 1] - offsets[t]);
    if (oc_state_from_actions(&s, actions + offsets[t], n) != OC_OK) {
      memset(&out[t], 0, sizeof(out[t]));
      out[t].cur_player = 127;
      bad++;
    } else {
      fill_rec(&s, &out[t]);
    }
  }
  return bad;
}

/* Dense tensors of the last state of each trajectory: info[t][2][2492], obs[t][2][98] (either may be NULL). */
int oc_final_tensors_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, float* info, float* obs) {
  int bad = 0;
  for (int t = 0; t < n_traj; ++t) {
    oc_state s;
    int n = (int)(offsets[t + 1] - offsets[t]);
    if (oc_state_from_actions(&s, actions + offsets[t], n) != OC_OK) { bad++; continue; }
    for (int p = 0; p < 2; ++p) {
      if (info) oc_information_state_tensor(&s, p, info + ((size_t)t * 2 + p) * OC_INFO_STATE_SIZE);
      if (obs) oc_observation_tensor(&s, p, obs + ((size_t)t * 2 + p) * OC_OBSERVATION_SIZE);
    }
  }
  return bad;
}

int oc_state_from_actions(oc_state* s, const uint8_t* actions, int n_actions) {
  oc_init(s);
  for (int i = 0; i < n_actions; ++i) {
    int rc = oc_apply_action(s, actions[i]);
    if (rc != OC_OK) return rc;
  }
  return OC_OK;
}

/* Replay digest: same definition as ref_replay_digest in oracle/ref_harness.cc (see there). */
static void fold(uint64_t* d, uint64_t v) { *d = *d * 0x9E3779B97F4A7C15ULL + v + 1; }

int oc_replay_digest(const uint8_t* actions, int n_actions, uint64_t* digest_out) {
  float info[OC_INFO_STATE_SIZE];
  float obs[OC_OBSERVATION_SIZE];
  oc_state s;
  uint64_t d = 0;
  int reports = 0;
  oc_init(&s);
  for (int i = 0; i < n_actions; ++i) {
    if (oc_apply_action(&s, actions[i]) != OC_OK) return -(i + 1);
    if (!oc_is_chance_node(&s)) {
      int term = oc_is_terminal(&s);
      double rew[2], ret[2];
      fold(&d, term ? 0u : oc_legal_mask(&s));
      fold(&d, (uint64_t)oc_current_player(&s) & 0xFF);
      fold(&d, (uint64_t)term);
      oc_rewards(&s, rew);
      oc_returns(&s, ret);
      fold(&d, (uint64_t)((int)rew[0] + 2));
      fold(&d, (uint64_t)((int)rew[1] + 2));
      fold(&d, (uint64_t)((int)ret[0] + 2));
      fold(&d, (uint64_t)((int)ret[1] + 2));
      for (int p = 0; p < 2; ++p) {
        oc_information_state_tensor(&s, p, info);
        fold(&d, oc_tensor_hash(info, OC_INFO_STATE_SIZE));
      }
      for (int p = 0; p < 2; ++p) {
        oc_observation_tensor(&s, p, obs);
        fold(&d, oc_tensor_hash(obs, OC_OBSERVATION_SIZE));
      }
      reports++;
    }
  }
  *digest_out = d;
  return reports;
}

int oc_replay_digest_batch(const uint8_t* actions, const int64_t* offsets, int n_traj, uint64_t* digests,
                           int32_t* reports) {
  int bad = 0;
  for (int t = 0; t < n_traj; ++t) {
    reports[t] = oc_replay_digest(actions + offsets[t], (int)(offsets[t + 1] - offsets[t]), &digests[t]);
    if (reports[t] < 0) bad++;
  }
  return bad;
}

/* ------------------------------------------------------------------------------------------------
 * Batched helpers
 * ---------------------------------------------------------------------------------------------- */

void oc_batch_init(oc_state* s, int n) {
  for (int i = 0; i < n; ++i) oc_init(&s[i]);
}

int oc_batch_apply(oc_state* s, int n, const uint8_t* actions, const uint8_t* do_mask) {
  int bad = 0;
  for (int i = 0; i < n; ++i) {
    if (do_mask && !do_mask[i]) continue;
    if (oc_apply_action_checked(&s[i], actions[i]) != OC_OK) bad++;
  }
  return bad;
}

void oc_batch_info_state(const oc_state* s, int n, const int8_t* players, float* out) {
  for (int i = 0; i < n; ++i)
    oc_information_state_tensor(&s[i], players[i], out + (size_t)i * OC_INFO_STATE_SIZE);
}

void oc_batch_observation(const oc_state* s, int n, const int8_t* players, float* out) {
  for (int i = 0; i < n; ++i)
    oc_observation_tensor(&s[i], players[i], out + (size_t)i * OC_OBSERVATION_SIZE);
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline ("port"): the reference's rollout benchmark protocol, examples/benchmark_game.cc:32-140
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
  uint64_t rng;
  long episodes;
  int mode;
  long moves, dec, chance, eps, trunc, ret[5], legal[8], max_moves, max_coins;
} bench_arg;

static uint64_t splitmix(uint64_t* st) {
  uint64_t z = (*st += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
static uint32_t bounded(uint64_t* st, uint32_t n) {
  return (uint32_t)(((splitmix(st) >> 32) * (uint64_t)n) >> 32);
}

static void* bench_thread(void* p) {
  bench_arg* a = (bench_arg*)p;
  float info[OC_INFO_STATE_SIZE];
  for (long e = 0; e < a->episodes; ++e) {
    oc_state s;
    oc_init(&s);
    long moves = 0;
    while (!oc_is_terminal(&s)) {
      if (oc_is_chance_node(&s)) {
        int total = 0;
        for (int i = 0; i < OC_NUM_CARD_TYPES; ++i) total += s.deck[i];
        int r = (int)bounded(&a->rng, (uint32_t)total);
        int c = 0;
        while (r >= s.deck[c]) { r -= s.deck[c]; c++; }
        oc_apply_action(&s, c);
        a->chance++;
      } else {
        int pl = oc_current_player(&s);
        if (a->mode >= 1) oc_information_state_tensor(&s, pl, info);
        if (a->mode >= 2) oc_information_state_tensor(&s, 1 - pl, info);
        int32_t la[OC_NUM_ACTIONS];
        int n = oc_legal_actions(&s, la);
        a->legal[n > 7 ? 7 : n]++;
        oc_apply_action(&s, la[bounded(&a->rng, (uint32_t)n)]);
        a->dec++;
      }
      moves++;
    }
    long mc = s.players[0].coins > s.players[1].coins ? s.players[0].coins : s.players[1].coins;
    if (mc > a->max_coins) a->max_coins = mc;
    a->moves += moves;
    a->eps++;
    if (moves > a->max_moves) a->max_moves = moves;
    if (moves > OC_MAX_GAME_LENGTH) a->trunc++;
    double ret[2];
    oc_returns(&s, ret);
    a->ret[(int)ret[0] + 2]++;
  }
  /* keep the tensor alive so the compiler cannot drop the encode */
  if (info[0] == 12345.0f) a->moves++;
  return 0;
}

int oc_bench(int mode, int threads, long episodes_total, uint32_t seed, double* out) {
  if (threads < 1) threads = 1;
  bench_arg* args = (bench_arg*)calloc((size_t)threads, sizeof(bench_arg));
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  long per = (episodes_total + threads - 1) / threads;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; ++t) {
    args[t].rng = 0x1234567ULL * (seed + 1) + 7919ULL * (uint64_t)t;
    args[t].episodes = per;
    args[t].mode = mode;
    pthread_create(&th[t], 0, bench_thread, &args[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], 0);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  for (int i = 0; i < 21; ++i) out[i] = 0;
  out[0] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  for (int t = 0; t < threads; ++t) {
    out[1] += (double)args[t].moves; out[2] += (double)args[t].dec; out[3] += (double)args[t].chance;
    out[4] += (double)args[t].eps; out[5] += (double)args[t].trunc;
    for (int i = 0; i < 5; ++i) out[6 + i] += (double)args[t].ret[i];
    for (int i = 0; i < 8; ++i) out[11 + i] += (double)args[t].legal[i];
    if ((double)args[t].max_moves > out[19]) out[19] = (double)args[t].max_moves;
    if ((double)args[t].max_coins > out[20]) out[20] = (double)args[t].max_coins;
  }
  free(args);
  free(th);
  return 0;
}
