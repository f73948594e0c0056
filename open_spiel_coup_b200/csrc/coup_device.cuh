// coup_device.cuh -- device-side Coup rules on the packed per-env state (sm_100a).
//
// One environment = one 128-bit word (layout in include/coup_b200.h). Everything here is
// __device__-only: there is no host instantiation of the rules in the product (the CPU restatement
// lives in oracle/ and is test tooling). Reference citations are to
// /root/reference/open_spiel/games/coup.cc unless another file is named.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace coup {

// ---- action / card ids (coup.h:50-85) ---------------------------------------------------------
enum : uint32_t {
  kIncome = 0, kForeignAid = 1, kCoup = 2, kTax = 3, kAssassinate = 4, kExchange = 5, kSteal = 6,
  kLoseCard1 = 7, kLoseCard2 = 8, kPass = 9, kBlock = 10, kChallenge = 11,
  kExchangeReturn12 = 12, kExchangeReturn34 = 17, kNoAction = 31
};
enum : uint32_t { kAssassin = 0, kAmbassador = 1, kCaptain = 2, kContessa = 3, kDuke = 4 };

constexpr int kMaxGameLength = 90;       // coup.h:219
constexpr int kInfoStateSize = 2492;     // coup.cc:1104-1116
constexpr int kObservationSize = 98;     // coup.cc:1118-1130
constexpr int kHeadSize = 62;            // floats shared by both tensors
constexpr int kNumActions = 18;
constexpr int kHistoryWords = 16;

// ---- packed state -----------------------------------------------------------------------------
struct Env {
  uint32_t p[2];  // per-player word
  uint32_t g;     // deck + phase + deal queue
  uint32_t c;     // counters + reward
};

// Player words are always picked with selects, never with a run-time array index, so that the state stays
// in registers (a dynamically indexed p[] would be demoted to local memory).
__device__ __forceinline__ uint32_t get_p(const Env& s, uint32_t i) { return i ? s.p[1] : s.p[0]; }
__device__ __forceinline__ void set_p(Env& s, uint32_t i, uint32_t v) {
  s.p[0] = i ? s.p[0] : v;
  s.p[1] = i ? v : s.p[1];
}

// player word
__device__ __forceinline__ uint32_t pw_hand(uint32_t w) { return w & 0xFFFFu; }
__device__ __forceinline__ uint32_t pw_coins(uint32_t w) { return (w >> 16) & 31u; }
__device__ __forceinline__ uint32_t pw_last(uint32_t w) { return (w >> 21) & 31u; }
__device__ __forceinline__ uint32_t pw_lost(uint32_t w) { return (w >> 26) & 1u; }
__device__ __forceinline__ uint32_t pw_set_hand(uint32_t w, uint32_t h) { return (w & ~0xFFFFu) | (h & 0xFFFFu); }
__device__ __forceinline__ uint32_t pw_add_coins(uint32_t w, int d) { return w + (static_cast<uint32_t>(d) << 16); }
__device__ __forceinline__ uint32_t pw_set_last(uint32_t w, uint32_t a) { return (w & ~(31u << 21)) | (a << 21); }
__device__ __forceinline__ uint32_t pw_set_lost(uint32_t w, uint32_t v) { return (w & ~(1u << 26)) | (v << 26); }

// global word
constexpr uint32_t kBitTurn = 1u << 20, kBitMover = 1u << 21, kBitTurnBegin = 1u << 22,
                   kBitChance = 1u << 23, kBitQInitial = 1u << 27, kBitQPlayer = 1u << 28,
                   kBitError = 1u << 29;
__device__ __forceinline__ uint32_t g_deck(uint32_t g, uint32_t c) { return (g >> (4 * c)) & 15u; }
__device__ __forceinline__ uint32_t g_turn(uint32_t g) { return (g >> 20) & 1u; }
__device__ __forceinline__ uint32_t g_mover(uint32_t g) { return (g >> 21) & 1u; }
__device__ __forceinline__ uint32_t g_turn_begin(uint32_t g) { return (g >> 22) & 1u; }
__device__ __forceinline__ uint32_t g_chance(uint32_t g) { return (g >> 23) & 1u; }
__device__ __forceinline__ uint32_t g_qn(uint32_t g) { return (g >> 24) & 7u; }

// counter word
__device__ __forceinline__ uint32_t c_moves(uint32_t c) { return c & 127u; }
__device__ __forceinline__ uint32_t c_turns(uint32_t c) { return (c >> 7) & 127u; }
__device__ __forceinline__ int c_reward0(uint32_t c) { return static_cast<int>((c >> 14) & 7u) - 2; }
__device__ __forceinline__ uint32_t c_set_reward0(uint32_t c, int r) {
  return (c & ~(7u << 14)) | (static_cast<uint32_t>(r + 2) << 14);
}

// ---- hand arithmetic: four sorted 4-bit slots, slot = value<<1 | face_up, 0xF empty ----------------
// Sorted ascending == CoupCard::operator< order (coup.h:91-94), so a hand is always what
// CoupPlayer::SortCards (389-391) would leave.
__device__ __forceinline__ uint32_t hand_slot(uint32_t h, uint32_t i) { return (h >> (4 * i)) & 15u; }
__device__ __forceinline__ uint32_t hand_empty_mask(uint32_t h) { return (h >> 3) & (h >> 2) & 0x1111u; }
__device__ __forceinline__ uint32_t hand_count(uint32_t h) { return 4u - __popc(hand_empty_mask(h)); }
__device__ __forceinline__ uint32_t hand_down_mask(uint32_t h) { return ~h & 0x1111u; }  // empties have bit0 set
__device__ __forceinline__ uint32_t hand_face_up_count(uint32_t h) {
  return __popc(h & 0x1111u) - __popc(hand_empty_mask(h));
}
// Insert a card keeping the order. Requires a free slot.
__device__ __forceinline__ uint32_t hand_insert(uint32_t h, uint32_t key) {
  uint32_t pos = (hand_slot(h, 0) <= key) + (hand_slot(h, 1) <= key) + (hand_slot(h, 2) <= key) +
                 (hand_slot(h, 3) <= key);
  uint32_t sh = 4 * pos;
  uint32_t low = h & ((1u << sh) - 1u);
  uint32_t high = (h >> sh) << (sh + 4);
  return (low | (key << sh) | high) & 0xFFFFu;
}
// vector::erase(begin()+slot)
__device__ __forceinline__ uint32_t hand_remove(uint32_t h, uint32_t slot) {
  uint32_t sh = 4 * slot;
  uint32_t low = h & ((1u << sh) - 1u);
  uint32_t high = (h >> (sh + 4)) << sh;
  return (low | high | 0xF000u) & 0xFFFFu;
}
// Index of the first slot equal to `key`, or 4. (HasFaceDownCard 379-387 / the search in
// ChallengeFailReplaceCard 471-483 with key = card<<1, i.e. face down.)
__device__ __forceinline__ uint32_t hand_find(uint32_t h, uint32_t key) {
  uint32_t x = h ^ (key * 0x1111u);
  uint32_t z = (x - 0x1111u) & ~x & 0x8888u;  // lowest flagged nibble is exact
  return z ? static_cast<uint32_t>(__ffs(z) - 1) >> 2 : 4u;
}

// ---- phase queries ----------------------------------------------------------------------------
// CoupState::IsTerminal, 989-1010.
__device__ __forceinline__ bool is_terminal(const Env& s) {
  if (c_moves(s.c) > kMaxGameLength) return true;
  uint32_t h0 = pw_hand(s.p[0]), h1 = pw_hand(s.p[1]);
  bool alive0 = hand_count(h0) < 2 || hand_down_mask(h0) != 0;
  bool alive1 = hand_count(h1) < 2 || hand_down_mask(h1) != 0;
  return !(alive0 && alive1);
}

// LegalLoseCardActions, 811-822.
__device__ __forceinline__ uint32_t lose_card_mask(uint32_t hand) {
  uint32_t m = 0;
  if ((hand & 0x1u) == 0) m |= 1u << kLoseCard1;
  if ((hand & 0x10u) == 0) m |= 1u << kLoseCard2;
  return m;
}

// CoupState::LegalActions at a decision node, 838-937, as a bitmask. Caller guarantees the state is
// neither terminal nor a chance node.
__device__ __forceinline__ uint32_t legal_mask_decision(const Env& s) {
  const uint32_t m = g_mover(s.g);
  const uint32_t cp = get_p(s, m), op = get_p(s, m ^ 1u);
  if (g_turn_begin(s.g)) {                                              // 841-854
    const uint32_t coins = pw_coins(cp);
    if (coins >= 10) return 1u << kCoup;
    uint32_t mask = (1u << kIncome) | (1u << kForeignAid) | (1u << kTax) | (1u << kExchange);
    if (coins >= 7) mask |= 1u << kCoup;
    if (coins >= 3) mask |= 1u << kAssassinate;
    if (pw_coins(op) > 0) mask |= 1u << kSteal;
    return mask;
  }
  if (pw_lost(cp)) return lose_card_mask(pw_hand(cp));                  // 856-858
  const uint32_t ol = pw_last(op);
  if (m != g_turn(s.g)) {                                               // 860-887
    if (ol == kForeignAid) return (1u << kPass) | (1u << kBlock);
    if (ol == kTax || ol == kExchange) return (1u << kPass) | (1u << kChallenge);
    if (ol == kSteal) return (1u << kPass) | (1u << kBlock) | (1u << kChallenge);
    if (ol == kAssassinate) return lose_card_mask(pw_hand(cp)) | (1u << kBlock) | (1u << kChallenge);
    if (ol == kCoup) return lose_card_mask(pw_hand(cp));
    return 0;  // unreachable by legal play (reference: SpielFatalError, 886)
  }
  if (pw_last(cp) == kExchange) {                                       // 889-928
    // first face-up slot f (or none) -> which ExchangeReturn pairs avoid it; 6 bits per case,
    // bit k <=> action 12+k: none 111111, f=0 111000, f=1 100110, f=2 010101, f=3 001011
    const uint32_t up = pw_hand(cp) & 0x1111u;
    const uint32_t f1 = up ? ((static_cast<uint32_t>(__ffs(up)) - 1u) >> 2) + 1u : 0u;
    const uint32_t table = 0x3Fu | (0x38u << 6) | (0x26u << 12) | (0x15u << 18) | (0x0Bu << 24);
    return ((table >> (6 * f1)) & 0x3Fu) << kExchangeReturn12;
  }
  if (ol == kBlock) return (1u << kPass) | (1u << kChallenge);          // 930-933
  return 0;  // unreachable by legal play (reference: SpielFatalError, 936)
}

// Chance node legal set, 828-836: card ids still in the deck.
__device__ __forceinline__ uint32_t legal_mask_chance(const Env& s) {
  uint32_t m = 0;
#pragma unroll
  for (uint32_t c = 0; c < 5; ++c) m |= (g_deck(s.g, c) != 0 ? 1u : 0u) << c;
  return m;
}

// CoupState::Returns, 1016-1032: returns[0] = faceUp(P2) - faceUp(P1).
__device__ __forceinline__ int returns_p0(const Env& s) {
  return static_cast<int>(hand_face_up_count(pw_hand(s.p[1]))) -
         static_cast<int>(hand_face_up_count(pw_hand(s.p[0])));
}

// ---- history log --------------------------------------------------------------------------------
// Six 5-bit move codes per 32-bit word. The writer keeps the word being filled in a register and
// flushes it when it moves on, so a word is only ever stored as (valid codes | zeros).
struct HistoryWriter {
  uint32_t* base;   // this env's 16 words
  uint32_t word;    // cached value
  int index;        // cached word index, -1 = none
  __device__ __forceinline__ explicit HistoryWriter(uint32_t* b) : base(b), word(0), index(-1) {}
  __device__ __forceinline__ void append(uint32_t move_index, uint32_t code) {
    const int wi = static_cast<int>(move_index / 6u);
    const uint32_t sh = 5u * (move_index - 6u * static_cast<uint32_t>(wi));
    if (wi != index) {
      flush();
      index = wi;
      word = (sh == 0) ? 0u : base[wi];
    }
    word |= code << sh;
  }
  __device__ __forceinline__ void flush() {
    if (index >= 0) base[index] = word;
  }
};

// ---- transitions --------------------------------------------------------------------------------
// NextPlayerTurn, 1079-1086.
__device__ __forceinline__ void next_turn(Env& s) {
  uint32_t t = g_turn(s.g) ^ 1u;
  s.g = (s.g & ~(kBitTurn | kBitMover)) | (t << 20) | (t << 21) | kBitTurnBegin;
  s.c += 1u << 7;
}
// NextPlayerMove, 1088-1092.
__device__ __forceinline__ void next_move(Env& s) { s.g = (s.g ^ kBitMover) & ~kBitTurnBegin; }

// Queue `n` deals to `player` (deal_card_to_.push, 480 / 583-584) and enter the chance phase.
__device__ __forceinline__ void queue_deals(Env& s, uint32_t player, uint32_t n) {
  uint32_t qn = g_qn(s.g) + n;
  s.g = (s.g & ~((7u << 24) | kBitQInitial | kBitQPlayer)) | (qn << 24) | (player << 28) | kBitChance;
}

// ChallengeFailReplaceCard, 468-486, for player `who` (the reference's opp_player_) whose word is `w`.
__device__ __forceinline__ void replace_card(Env& s, uint32_t& w, uint32_t who, uint32_t card) {
  uint32_t h = pw_hand(w);
  uint32_t slot = hand_find(h, card << 1);
  s.g += 1u << (4 * card);  // deck_[card] += 1
  w = pw_set_hand(w, hand_remove(h, slot));
  queue_deals(s, who, 1);
}

// Turn every face-down card among slots 0 and 1 of a player face up (660-669 / 733-742); returns the
// number flipped. No re-sort: the reference does not sort here, and the order cannot change.
__device__ __forceinline__ int flip_two(uint32_t& w) {
  uint32_t h = pw_hand(w);
  uint32_t down = ~h & 0x11u;
  w = pw_set_hand(w, h | down);
  return __popc(down);
}

// (`to` steals from `from`) 599-601 / 685-687 / 758-760
__device__ __forceinline__ void steal(uint32_t& to, uint32_t& from) {
  int k = pw_coins(from) > 1 ? 2 : 1;
  to = pw_add_coins(to, k);
  from = pw_add_coins(from, -k);
}

// One PLAYER move: State::ApplyAction (spiel.cc:322-332) + the non-chance branch of
// CoupState::DoApplyAction (522-807) with its two recursions (628, 718) flattened. The caller has
// checked that `a` is in the legal mask. Appends the move to the history and bumps move_number_.
__device__ __forceinline__ void apply_player_action(Env& s, uint32_t a, HistoryWriter& hist) {
  const uint32_t m = g_mover(s.g), o = m ^ 1u;
  hist.append(c_moves(s.c), a);
  int rew_m = 0;  // reward of the mover; the other player's is the negation
  uint32_t cp = get_p(s, m), op = get_p(s, o);  // mover / other player words, written back at the end
  const uint32_t prev_last = pw_last(cp);
  // Every branch of the reference sets cp.last_action = action before anything else, except the
  // "complete the action" else-branches that only run through the recursion.
  cp = pw_set_last(cp, a);
  switch (a) {
    case kIncome:                                                      // 531-534
      cp = pw_add_coins(cp, 1);
      next_turn(s);
      break;
    case kForeignAid: case kTax: case kExchange: case kSteal:          // declared: 536-541 etc.
    case kBlock:                                                       // 631-633
      next_move(s);
      break;
    case kCoup:                                                        // 548-553
      cp = pw_add_coins(cp, -7);
      next_move(s);
      break;
    case kAssassinate:                                                 // 567-573
      cp = pw_add_coins(cp, -3);
      next_move(s);
      break;
    case kLoseCard1: case kLoseCard2: {                                // 605-616
      const uint32_t k = a - kLoseCard1;
      uint32_t h = pw_hand(cp);
      const uint32_t key = hand_slot(h, k);
      h = hand_insert(hand_remove(h, k), key | 1u);  // FaceUp, then SortCards
      cp = pw_set_lost(pw_set_hand(cp, h), 0);
      rew_m = -1;
      next_turn(s);
      break;
    }
    case kPass: {                                                      // 618-629
      const uint32_t n = pw_last(op);
      if (n == kBlock) {
        next_turn(s);
      } else if (n == kExchange) {
        // NextPlayerMove, then Exchange's else-branch (582-586): two deals to the actor.
        next_move(s);
        queue_deals(s, o, 2);
      } else {
        // NextPlayerMove then the completing else-branch, which ends in NextPlayerTurn.
        if (n == kForeignAid) op = pw_add_coins(op, 2);        // 544
        else if (n == kTax) op = pw_add_coins(op, 3);          // 563
        else steal(op, cp);                                           // 599-601 (n == kSteal)
        next_turn(s);
      }
      break;
    }
    case kChallenge: {                                                 // 635-771
      const uint32_t ol = pw_last(op);
      const uint32_t oh = pw_hand(op);
      if (ol == kBlock) {
        if (prev_last == kForeignAid) {                                // 637-649
          if (hand_find(oh, kDuke << 1) < 4) {
            cp = pw_set_lost(cp, 1);
            replace_card(s, op, o, kDuke);
          } else {
            op = pw_set_lost(op, 1);
            cp = pw_add_coins(cp, 2);
            next_move(s);
          }
        } else if (prev_last == kAssassinate) {                        // 650-670
          if (hand_find(oh, kContessa << 1) < 4) {
            cp = pw_set_lost(cp, 1);
            replace_card(s, op, o, kContessa);
          } else {
            rew_m = flip_two(op);
          }
        } else {                                                       // 671-690 (prev == kSteal)
          if (hand_find(oh, kCaptain << 1) < 4) {
            cp = pw_set_lost(cp, 1);
            replace_card(s, op, o, kCaptain);
          } else if (hand_find(oh, kAmbassador << 1) < 4) {
            cp = pw_set_lost(cp, 1);
            replace_card(s, op, o, kAmbassador);
          } else {
            op = pw_set_lost(op, 1);
            steal(cp, op);
            next_move(s);
          }
        }
      } else if (ol == kTax) {                                         // 694-706
        if (hand_find(oh, kDuke << 1) < 4) {
          cp = pw_set_lost(cp, 1);
          replace_card(s, op, o, kDuke);
          op = pw_add_coins(op, 3);
        } else {
          op = pw_set_lost(op, 1);
          next_move(s);
        }
      } else if (ol == kExchange) {                                    // 708-725
        if (hand_find(oh, kAmbassador << 1) < 4) {
          cp = pw_set_lost(cp, 1);
          replace_card(s, op, o, kAmbassador);
          next_move(s);
          queue_deals(s, o, 2);  // the recursive Exchange (718): queue is now [o, o, o]
        } else {
          op = pw_set_lost(op, 1);
          next_move(s);
        }
      } else if (ol == kAssassinate) {                                 // 727-749
        if (hand_find(oh, kAssassin << 1) < 4) {
          rew_m = -flip_two(cp);
        } else {
          op = pw_add_coins(pw_set_lost(op, 1), 3);
          next_move(s);
        }
      } else {                                                         // 751-767 (ol == kSteal)
        if (hand_find(oh, kCaptain << 1) < 4) {
          cp = pw_set_lost(cp, 1);
          replace_card(s, op, o, kCaptain);
          steal(op, cp);
        } else {
          op = pw_set_lost(op, 1);
          next_move(s);
        }
      }
      break;
    }
    default: {                                                         // 773-803 ExchangeReturnXY
      // slot pairs (i<j) for 12..17: (0,1)(0,2)(0,3)(1,2)(1,3)(2,3), packed 2 bits each
      const uint32_t k = a - kExchangeReturn12;
      const uint32_t i = (0x940u >> (2 * k)) & 3u;    // 0,0,0,1,1,2
      const uint32_t j = (0xFB9u >> (2 * k)) & 3u;    // 1,2,3,2,3,3
      uint32_t h = pw_hand(cp);
      h = hand_remove(hand_remove(h, j), i);
      cp = pw_set_hand(cp, h);
      // REFERENCE QUIRK (789-795): the deck count that grows is indexed by the hand SLOT, not by the
      // value of the card that was returned. Reproduced for bit-exact replay.
      s.g += (1u << (4 * j)) + (1u << (4 * i));
      if (pw_lost(op)) next_move(s); else next_turn(s);
      break;
    }
  }
  set_p(s, m, cp);
  set_p(s, o, op);
  // cur_rewards_ (527, 614-615, 662-668, 735-741): zero-sum, stored from player 0's point of view.
  s.c = c_set_reward0(s.c, m == 0 ? rew_m : -rew_m) + 1u;  // ++move_number_
}

// One CHANCE move: the chance branch of DoApplyAction, 491-520. `card` must be in the deck.
__device__ __forceinline__ void apply_chance(Env& s, uint32_t card, HistoryWriter& hist) {
  const uint32_t qn = g_qn(s.g);
  const uint32_t target = (s.g & kBitQInitial) ? (qn & 1u) : ((s.g >> 28) & 1u);  // queue 0,1,0,1
  hist.append(c_moves(s.c), 18u + 5u * target + card);
  s.g -= 1u << (4 * card);                                             // deck_[card] -= 1
  const uint32_t tw = get_p(s, target);
  set_p(s, target, pw_set_hand(tw, hand_insert(pw_hand(tw), card << 1)));
  s.g -= 1u << 24;                                                     // pop
  if (qn == 1) s.g &= ~(kBitChance | kBitQInitial);                    // 520
  s.c += 1u;                                                           // ++move_number_
}

// Draw a card with probability deck_[c] / sum(deck_) (ChanceOutcomes, 1062-1077) from one uniform
// 32-bit word: r = floor(u * total / 2^32), then the first c whose running count exceeds r.
__device__ __forceinline__ uint32_t sample_card(const Env& s, uint32_t u) {
  uint32_t d0 = g_deck(s.g, 0), d1 = g_deck(s.g, 1), d2 = g_deck(s.g, 2), d3 = g_deck(s.g, 3),
           d4 = g_deck(s.g, 4);
  uint32_t r = __umulhi(u, d0 + d1 + d2 + d3 + d4);
  uint32_t c0 = d0, c1 = c0 + d1, c2 = c1 + d2, c3 = c2 + d3;
  return (r >= c0) + (r >= c1) + (r >= c2) + (r >= c3);
}

// CoupState::CoupState, 393-428.
__device__ __forceinline__ Env initial_state() {
  Env s;
  s.p[0] = 0xFFFFu | (1u << 16) | (kNoAction << 21);
  s.p[1] = 0xFFFFu | (2u << 16) | (kNoAction << 21);
  s.g = 0x33333u | kBitTurnBegin | kBitChance | (4u << 24) | kBitQInitial;
  s.c = 2u << 14;
  return s;
}

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based: no per-env RNG state ---------------------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
}
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = mulhi32(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// Stream layout: key = (global env id lo, global env id hi ^ seed lo);
// counter = (step lo, step hi, purpose, seed hi). purpose 0: x = action choice, y/z/w = the up to
// three chance draws that can follow one player action; purpose 1: the four deals of a reset.
// (__host__ too: the host-side uniform policy of the host-buffer path draws from the same stream.)
__host__ __device__ __forceinline__ uint4 env_random(uint64_t seed, uint64_t global_env, uint64_t step, uint32_t purpose) {
  uint2 key = make_uint2(static_cast<uint32_t>(global_env),
                         static_cast<uint32_t>(global_env >> 32) ^ static_cast<uint32_t>(seed));
  uint4 ctr = make_uint4(static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32), purpose,
                         static_cast<uint32_t>(seed >> 32));
  return philox4x32_10(ctr, key);
}

__device__ __forceinline__ uint32_t pick(const uint4& r, int k) {
  return k == 0 ? r.x : k == 1 ? r.y : k == 2 ? r.z : r.w;
}

// k-th (0-based) set bit of a non-zero mask.
__device__ __forceinline__ uint32_t kth_set_bit(uint32_t mask, uint32_t k) {
  return __fns(mask, 0, static_cast<int>(k) + 1);
}

// Uniform legal action (benchmark_game.cc:96-99) from one uniform 32-bit word.
__device__ __forceinline__ uint32_t sample_action(uint32_t legal, uint32_t u) {
  return kth_set_bit(legal, __umulhi(u, static_cast<uint32_t>(__popc(legal))));
}

// Resolve every pending chance node (rl_environment._sample_external_events, rl_environment.py:369-382).
// `rnd` supplies draws rnd[first..]; forced[k] != 0xFF overrides the k-th draw. Stops at terminal
// (the move cap can hit in the middle of a deal sequence, coup.cc:990). Returns #chance moves.
__device__ __forceinline__ int resolve_chance(Env& s, const uint4& rnd, int first, const uint8_t* forced,
                                              HistoryWriter& hist) {
  int k = 0;
  while (g_chance(s.g) && !is_terminal(s) && k < 4) {
    uint32_t card = sample_card(s, pick(rnd, first + k));
    if (forced != nullptr) {
      uint32_t f = forced[k];
      if (f < 5u && g_deck(s.g, f) != 0) card = f;
    }
    apply_chance(s, card, hist);
    ++k;
  }
  return k;
}

// ---- observer head --------------------------------------------------------------------------------
// The first 62 floats of both tensors (CoupObserver::WriteTensor 248-279): bit q of the returned mask
// is 1 where float q is 1; floats 60/61 are the raw coin counts (207-213), returned separately.
//   [0,2) observer  [2,22) p1_cards[4][5]  [22,42) p2_cards[4][5]  [42,44) cur_move_player
//   [44,60) cards_state[2][4][2]  [60,62) coins
__device__ __forceinline__ uint64_t head_mask(const Env& s, uint32_t observer, bool terminal) {
  uint64_t mask = 1ull << observer;                                    // WritePlayer, 160-165
#pragma unroll
  for (uint32_t pl = 0; pl < 2; ++pl) {
    const uint32_t h = pw_hand(s.p[pl]);
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
      const uint32_t key = hand_slot(h, i);
      if (key != 15u) {
        const uint32_t up = key & 1u, value = key >> 1;
        // WritePlayerCardsValue 178-191 with kSinglePlayer private info (258-265): own face-down
        // cards and every face-up card are visible.
        if (up || pl == observer) mask |= 1ull << (2u + 20u * pl + 5u * i + value);
        mask |= 1ull << (44u + (pl * 4u + i) * 2u + up);               // WriteCardsState, 194-204
      }
    }
  }
  if (!terminal) mask |= 1ull << (42u + g_mover(s.g));                 // 268-276
  return mask;
}

// Observation tensor tail [62,98): last_action[2][18] one-hot (WriteLastAction, 217-225), as bits
// 0..35 of the returned mask.
__device__ __forceinline__ uint64_t last_action_mask(const Env& s) {
  uint64_t m = 0;
  uint32_t l0 = pw_last(s.p[0]), l1 = pw_last(s.p[1]);
  if (l0 != kNoAction) m |= 1ull << l0;
  if (l1 != kNoAction) m |= 1ull << (18u + l1);
  return m;
}

// Column of history row code `code` as seen by `observer` (WriteActionHistory, 230-245): player moves
// are public, a deal is visible only to the player who received it. 31 = all-zero row.
__device__ __forceinline__ uint32_t history_column(uint32_t code, uint32_t observer) {
  if (code < 18u) return code;
  const uint32_t base = 18u + 5u * observer;
  return (code >= base && code < base + 5u) ? code - base : 31u;
}

// 64-bit finaliser used by the position-keyed tensor hash (see coup_tensor_row_hash).
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

}  // namespace coup
