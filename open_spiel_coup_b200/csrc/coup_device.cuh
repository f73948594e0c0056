// coup_device.cuh -- device-side Coup rules on the packed per-env state (sm_100a).
//
// One environment = one 128-bit word (layout in include/coup_b200.h). In the product everything here is
// __device__-only: libcoup_b200.so holds no host instantiation of the rules (the CPU restatement lives in oracle/ and is
// test tooling). Reference citations are to /root/reference/open_spiel/games/coup.cc unless another file is named.
//
// The rules are BRANCH-FREE: a player action is first resolved to one of 25 events (the action id, refined for Pass by
// the action being passed on and for Challenge by what is challenged and whether the challenged player holds the card
// he claimed), the event selects a 32-bit effect descriptor from a table (coin deltas, lost-challenge flags, turn/move
// advance, card to replace, deals to queue, cards to flip), and the descriptor is applied with selects. A warp whose 32
// envs are in 32 different phases of the game executes one instruction stream, not the union of the reference's
// 18-way if-chain (coup.cc:531-807).
//
// COUP_RULES_HOST_TEST (defined only by tests/host_harness, never by the library build) compiles the same functions for
// the host, so that the CPU test-suite can diff them against the oracle without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(COUP_RULES_HOST_TEST) && !defined(__CUDACC__)
#define COUP_FN static inline
#define COUP_TABLE static const
#else
#define COUP_FN __device__ __forceinline__
#define COUP_TABLE __device__ const
#endif

namespace coup {

#if defined(COUP_RULES_HOST_TEST) && !defined(__CUDACC__)
COUP_FN uint32_t popc32(uint32_t x) { return static_cast<uint32_t>(__builtin_popcount(x)); }
COUP_FN uint32_t ffs32(uint32_t x) { return static_cast<uint32_t>(__builtin_ffs(static_cast<int>(x))); }
COUP_FN uint32_t umulhi32(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
COUP_FN uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
COUP_FN uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }
#else
COUP_FN uint32_t popc32(uint32_t x) { return static_cast<uint32_t>(__popc(x)); }
COUP_FN uint32_t ffs32(uint32_t x) { return static_cast<uint32_t>(__ffs(static_cast<int>(x))); }
COUP_FN uint32_t umulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
COUP_FN uint32_t umin32(uint32_t a, uint32_t b) { return min(a, b); }
COUP_FN uint32_t umax32(uint32_t a, uint32_t b) { return max(a, b); }
#endif

// x >> C for a compile-time C. (Issuing these as the high half of a multiply by 2^(32-C), to move them off the busy ALU
// pipe, was measured and is slower: env-only 33.7 -> 35.2 us per 2^20-env step.)
template <uint32_t C>
COUP_FN uint32_t rsh(uint32_t x) { return x >> C; }

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
COUP_FN uint32_t get_p(const Env& s, uint32_t i) { return i ? s.p[1] : s.p[0]; }
COUP_FN void set_p(Env& s, uint32_t i, uint32_t v) {
  s.p[0] = i ? s.p[0] : v;
  s.p[1] = i ? v : s.p[1];
}

// player word
COUP_FN uint32_t pw_hand(uint32_t w) { return w & 0xFFFFu; }
COUP_FN uint32_t pw_coins(uint32_t w) { return rsh<16>(w) & 31u; }
COUP_FN uint32_t pw_last(uint32_t w) { return rsh<21>(w) & 31u; }
COUP_FN uint32_t pw_lost(uint32_t w) { return rsh<26>(w) & 1u; }
COUP_FN uint32_t pw_set_hand(uint32_t w, uint32_t h) { return (w & ~0xFFFFu) | (h & 0xFFFFu); }
COUP_FN uint32_t pw_add_coins(uint32_t w, int d) { return w + (static_cast<uint32_t>(d) << 16); }
COUP_FN uint32_t pw_set_last(uint32_t w, uint32_t a) { return (w & ~(31u << 21)) | (a << 21); }
COUP_FN uint32_t pw_set_lost(uint32_t w, uint32_t v) { return (w & ~(1u << 26)) | (v << 26); }

// global word
constexpr uint32_t kBitTurn = 1u << 20, kBitMover = 1u << 21, kBitTurnBegin = 1u << 22,
                   kBitChance = 1u << 23, kBitQInitial = 1u << 27, kBitQPlayer = 1u << 28,
                   kBitError = 1u << 29;
COUP_FN uint32_t g_deck(uint32_t g, uint32_t c) { return (g >> (4 * c)) & 15u; }
COUP_FN uint32_t g_turn(uint32_t g) { return rsh<20>(g) & 1u; }
COUP_FN uint32_t g_mover(uint32_t g) { return rsh<21>(g) & 1u; }
COUP_FN uint32_t g_turn_begin(uint32_t g) { return rsh<22>(g) & 1u; }
COUP_FN uint32_t g_chance(uint32_t g) { return rsh<23>(g) & 1u; }
COUP_FN uint32_t g_qn(uint32_t g) { return rsh<24>(g) & 7u; }

// counter word
COUP_FN uint32_t c_moves(uint32_t c) { return c & 127u; }
COUP_FN uint32_t c_turns(uint32_t c) { return rsh<7>(c) & 127u; }
COUP_FN int c_reward0(uint32_t c) { return static_cast<int>(rsh<14>(c) & 7u) - 2; }
COUP_FN uint32_t c_set_reward0(uint32_t c, int r) {
  return (c & ~(7u << 14)) | (static_cast<uint32_t>(r + 2) << 14);
}

// ---- hand arithmetic: four sorted 4-bit slots, slot = value<<1 | face_up, 0xF empty ----------------
// Sorted ascending == CoupCard::operator< order (coup.h:91-94), so a hand is always what
// CoupPlayer::SortCards (389-391) would leave.
COUP_FN uint32_t hand_slot(uint32_t h, uint32_t i) { return (h >> (4 * i)) & 15u; }
COUP_FN uint32_t hand_empty_mask(uint32_t h) { return rsh<3>(h) & rsh<2>(h) & 0x1111u; }
COUP_FN uint32_t hand_count(uint32_t h) { return 4u - popc32(hand_empty_mask(h)); }
COUP_FN uint32_t hand_down_mask(uint32_t h) { return ~h & 0x1111u; }  // empties have bit0 set
COUP_FN uint32_t hand_face_up_count(uint32_t h) {
  return popc32(h & 0x1111u) - popc32(hand_empty_mask(h));
}
// Insert a card keeping the order. Requires a free slot. The position is the number of slots <= key, counted for all
// four slots at once: the nibbles are spread to bytes, and (0x10 + key - slot) keeps bit 4 exactly when slot <= key.
COUP_FN uint32_t hand_insert(uint32_t h, uint32_t key) {
  uint32_t x = (h | (h << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;                                   // byte i = slot i
  const uint32_t le = ((key * 0x01010101u + 0x10101010u) - x) & 0x10101010u;
  const uint32_t sh = 4u * popc32(le);
  const uint32_t low = h & ((1u << sh) - 1u);
  const uint32_t high = (h >> sh) << (sh + 4u);
  return (low | (key << sh) | high) & 0xFFFFu;
}
// vector::erase(begin()+slot)
COUP_FN uint32_t hand_remove(uint32_t h, uint32_t slot) {
  uint32_t sh = 4 * slot;
  uint32_t low = h & ((1u << sh) - 1u);
  uint32_t high = (h >> (sh + 4)) << sh;
  return (low | high | 0xF000u) & 0xFFFFu;
}
// Index of the first slot equal to `key`, or 4. (HasFaceDownCard 379-387 / the search in
// ChallengeFailReplaceCard 471-483 with key = card<<1, i.e. face down.)
COUP_FN uint32_t hand_find(uint32_t h, uint32_t key) {
  uint32_t x = h ^ (key * 0x1111u);
  uint32_t z = (x - 0x1111u) & ~x & 0x8888u;  // lowest flagged nibble is exact
  return z ? (ffs32(z) - 1u) >> 2 : 4u;
}

// ---- phase queries ----------------------------------------------------------------------------
// CoupState::IsTerminal, 989-1010.
COUP_FN bool is_terminal(const Env& s) {
  // A player is out when he holds at least two cards and none of them is face down (995-999). Empty slots sort last and
  // have bit 0 set, so "no face-down card" is bit 0 of all four slots, and "at least two cards" is slot 1 not empty
  // (bits 2 and 3 of a card slot are never both set: values are 0..4).
  const bool out0 = (s.p[0] & 0x1111u) == 0x1111u && (s.p[0] & 0xC0u) != 0xC0u;
  const bool out1 = (s.p[1] & 0x1111u) == 0x1111u && (s.p[1] & 0xC0u) != 0xC0u;
  return c_moves(s.c) > kMaxGameLength || out0 || out1;
}

// LegalLoseCardActions, 811-822: slot 0 / slot 1 face down -> bit kLoseCard1 / kLoseCard2.
COUP_FN uint32_t lose_card_mask(uint32_t hand) {
  const uint32_t down = ~hand;
  return ((down & 1u) << kLoseCard1) | (((down >> 4) & 1u) << kLoseCard2);
}

// Small lookup tables packed into 64-bit immediates: entry i = (table >> (bits * i)) & mask.
COUP_FN uint32_t lut3(uint64_t table, uint32_t i) { return static_cast<uint32_t>(table >> (3u * i)) & 7u; }

// CoupState::LegalActions at a decision node, 838-937, as a bitmask, without branches: the candidate sets of the five
// phases are all cheap, so each is computed and the phase picks one. Caller guarantees the state is neither terminal
// nor a chance node.
COUP_FN uint32_t legal_mask_decision(const Env& s) {
  const uint32_t m = g_mover(s.g);
  const uint32_t cp = get_p(s, m), op = get_p(s, m ^ 1u);
  const uint32_t coins = pw_coins(cp), ol = pw_last(op);
  constexpr uint32_t kP = 1u << kPass, kB = 1u << kBlock, kC = 1u << kChallenge;
  // turn begin, 841-854
  uint32_t begin = (1u << kIncome) | (1u << kForeignAid) | (1u << kTax) | (1u << kExchange);
  begin |= (coins >= 7u ? 1u << kCoup : 0u) | (coins >= 3u ? 1u << kAssassinate : 0u) | (pw_coins(op) > 0u ? 1u << kSteal : 0u);
  begin = coins >= 10u ? 1u << kCoup : begin;
  // lost a challenge, 856-858; also part of the answers to Assassinate and Coup
  const uint32_t lose = lose_card_mask(pw_hand(cp));
  // answering the opponent's declared action, 860-887: three bits (Pass, Block, Challenge) per action id
  //   ForeignAid P B . | Tax P . C | Exchange P . C | Steal P B C | Assassinate . B C (+lose) | Coup (lose only)
  constexpr uint64_t kAnswer = (3ull << (3 * kForeignAid)) | (5ull << (3 * kTax)) | (5ull << (3 * kExchange)) |
                               (7ull << (3 * kSteal)) | (6ull << (3 * kAssassinate));
  const uint32_t ol_c = umin32(ol, 15u);
  const uint32_t answer = (lut3(kAnswer, ol_c) << kPass) | ((ol == kAssassinate || ol == kCoup) ? lose : 0u);
  // returning two of four cards, 889-928: first face-up slot f (or none) -> the pairs that avoid it; 6 bits per case,
  // bit k <=> action 12+k: none 111111, f=0 111000, f=1 100110, f=2 010101, f=3 001011
  const uint32_t up = pw_hand(cp) & 0x1111u;
  const uint32_t f1 = up ? ((ffs32(up) - 1u) >> 2) + 1u : 0u;
  constexpr uint32_t kPairs = 0x3Fu | (0x38u << 6) | (0x26u << 12) | (0x15u << 18) | (0x0Bu << 24);
  const uint32_t give_back = ((kPairs >> (6u * f1)) & 0x3Fu) << kExchangeReturn12;
  // own turn, the opponent blocked, 930-933
  const uint32_t blocked = ol == kBlock ? (kP | kC) : 0u;
  (void)kB;
  const uint32_t own = pw_last(cp) == kExchange ? give_back : blocked;
  const uint32_t later = pw_lost(cp) ? lose : (m != g_turn(s.g) ? answer : own);
  return g_turn_begin(s.g) ? begin : later;
}

// Chance node legal set, 828-836: card ids still in the deck.
COUP_FN uint32_t legal_mask_chance(const Env& s) {
  uint32_t m = 0;
#pragma unroll
  for (uint32_t c = 0; c < 5; ++c) m |= (g_deck(s.g, c) != 0 ? 1u : 0u) << c;
  return m;
}

// CoupState::Returns, 1016-1032: returns[0] = faceUp(P2) - faceUp(P1).
COUP_FN int returns_p0(const Env& s) {
  // both hands side by side in one word: bit 0 of a slot is set for face-up cards and for empty slots
  const uint32_t hh = pw_hand(s.p[0]) | (s.p[1] << 16);
  const uint32_t up = hh & ~(rsh<3>(hh) & rsh<2>(hh)) & 0x11111111u;
  return static_cast<int>(popc32(rsh<16>(up))) - static_cast<int>(popc32(up & 0xFFFFu));
}

// ---- history log --------------------------------------------------------------------------------
// Six 5-bit move codes per 32-bit word. A step produces at most five codes (one player action and up to four deals);
// they are collected in a register and merged into the row with one read-modify-write of at most two words.
// `codes` holds `n` codes, 5 bits each, the first in the low bits; they become moves first .. first+n-1 of the row.
// Words past the last move are left alone ("entries at index >= move_number_ are unspecified").
COUP_FN void history_commit(uint32_t* row, uint32_t first, uint32_t codes, uint32_t n, uint32_t* mirror = nullptr) {
  // `row` is read and updated; `mirror` (optional) receives the same words write-only -- the step kernels keep the row
  // they work on in shared memory and write through to HBM, so that a step never waits for a second global load.
  const uint32_t w = first / 6u;
  const uint32_t sh = 5u * (first - 6u * w);
  const uint64_t bits = static_cast<uint64_t>(codes) << sh;
  const uint32_t keep = sh ? row[w] & ((1u << sh) - 1u) : 0u;          // a word is only ever (valid codes | zeros)
  const uint32_t v0 = keep | (static_cast<uint32_t>(bits) & 0x3FFFFFFFu);
  row[w] = v0;
  if (mirror != nullptr) mirror[w] = v0;
  if (sh + 5u * n > 30u) {
    const uint32_t v1 = static_cast<uint32_t>(bits >> 30);
    row[w + 1u] = v1;
    if (mirror != nullptr) mirror[w + 1u] = v1;
  }
}

// ---- transitions --------------------------------------------------------------------------------
// Effect descriptor of one event. Fields:
//   bits 0-3  coin delta of the mover + 8        bits 4-5  coin delta of the other player
//   bits 6-7  coin transfer: 1 mover takes from the other player, 2 the other player takes from the mover
//             (1 or 2 coins: 2 if the victim has more than one, 599-601 / 685-687 / 758-760)
//   bit 8 mover.lost_challenge = 1   bit 9 other.lost_challenge = 1   bit 10 mover.lost_challenge = 0
//   bits 11-12 who moves next: 0 nobody changes, 1 NextPlayerMove (1088-1092), 2 NextPlayerTurn (1079-1086),
//             3 NextPlayerMove if the other player still has a card to lose, else NextPlayerTurn (798-802)
//   bit 13 ChallengeFailReplaceCard on the other player (468-486): the shown card goes back, one deal is queued
//   bit 14 two more deals queued for the other player (the Exchange draw, 583-584)
//   bits 15-16 flip the face-down cards among slots 0, 1: 1 of the other player (reward +n), 2 of the mover (reward -n)
//   bit 17 LoseCard (605-616)   bit 18 ExchangeReturn (773-803)
constexpr uint32_t desc(int mover_coins, int other_coins, int transfer, int mover_lost, int other_lost, int clear_lost,
                        int advance, int replace, int two_deals, int flip, int lose, int give_back) {
  return static_cast<uint32_t>(mover_coins + 8) | (other_coins << 4) | (transfer << 6) | (mover_lost << 8) | (other_lost << 9) |
         (clear_lost << 10) | (advance << 11) | (replace << 13) | (two_deals << 14) | (flip << 15) | (lose << 17) |
         (give_back << 18);
}
enum : uint32_t { kEvPass = 6, kEvChallenge = 11, kNumEvents = 25 };
//                                 coins:mv ot tr | lost:mv ot clr | adv rep 2d flip lose back
COUP_TABLE uint32_t kEventTable[32] = {
    /* 0 Income                531-534 */ desc(+1, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0),
    /* 1 declare FA/Tax/Exchange/Steal (536-541, 555-560, 575-580, 589-596), Block (631-633) */
                                          desc(0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0),
    /* 2 Coup                  548-553 */ desc(-7, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0),
    /* 3 Assassinate           567-573 */ desc(-3, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0),
    /* 4 LoseCard1/2           605-616 */ desc(0, 0, 0, 0, 0, 1, 2, 0, 0, 0, 1, 0),
    /* 5 ExchangeReturnXY      773-803 */ desc(0, 0, 0, 0, 0, 0, 3, 0, 0, 0, 0, 1),
    // Pass (618-629) completes the action it lets through, the recursion of 628 flattened:
    /* 6 .. on a Block                 */ desc(0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0),
    /* 7 .. on Exchange        582-586 */ desc(0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0),
    /* 8 .. on ForeignAid          544 */ desc(0, 2, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0),
    /* 9 .. on Tax                 563 */ desc(0, 3, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0),
    /* 10 .. on Steal          599-601 */ desc(0, 0, 2, 0, 0, 0, 2, 0, 0, 0, 0, 0),
    // Challenge (635-771): 11 + 2 * kind + (the challenged player holds the card he claimed)
    /* 11 Block of ForeignAid, no Duke       645-648 */ desc(+2, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0),
    /* 12 .. Duke shown                      639-643 */ desc(0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0),
    /* 13 Block of Assassinate, no Contessa  657-669 */ desc(0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0),
    /* 14 .. Contessa shown                  652-656 */ desc(0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0),
    /* 15 Block of Steal, neither card       682-689 */ desc(0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0),
    /* 16 .. Captain or Ambassador shown     673-681 */ desc(0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0),
    /* 17 Tax, no Duke                       702-705 */ desc(0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0),
    /* 18 .. Duke shown (tax is paid now)    696-701 */ desc(0, 3, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0),
    /* 19 Exchange, no Ambassador            721-724 */ desc(0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0),
    /* 20 .. Ambassador shown, draw (718)    710-720 */ desc(0, 0, 0, 1, 0, 0, 1, 1, 1, 0, 0, 0),
    /* 21 Assassinate, no Assassin: refund   744-748 */ desc(0, 3, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0),
    /* 22 .. Assassin shown: both cards go   729-743 */ desc(0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0),
    /* 23 Steal, no Captain                  763-766 */ desc(0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0),
    /* 24 .. Captain shown, steal completes  753-762 */ desc(0, 0, 2, 1, 0, 0, 0, 1, 0, 0, 0, 0),
    0, 0, 0, 0, 0, 0, 0};

// One PLAYER move: State::ApplyAction (spiel.cc:322-332) + the non-chance branch of CoupState::DoApplyAction
// (522-807), both recursions (628, 718) flattened. The caller has checked that `a` is in the legal mask and logs the
// move itself (the history code of a player move is its action id). Bumps move_number_. No branches: lanes of a warp
// that apply 32 different actions run the same instructions.
COUP_FN void apply_player_action(Env& s, uint32_t a) {
  const uint32_t m = g_mover(s.g), o = m ^ 1u;
  uint32_t cp = get_p(s, m), op = get_p(s, o);   // mover / other player words, written back at the end
  const uint32_t prev_last = umin32(pw_last(cp), 15u), ol = umin32(pw_last(op), 15u);
  const uint32_t ch = pw_hand(cp), oh = pw_hand(op);

  // ---- which event -------------------------------------------------------------------------------
  // by action id (Pass and Challenge are refined below): 0 Income 1 declare/Block 2 Coup 3 Assassinate 4 LoseCard 5 Return
  constexpr uint64_t kByAction = (0ull << (3 * kIncome)) | (1ull << (3 * kForeignAid)) | (2ull << (3 * kCoup)) |
                                 (1ull << (3 * kTax)) | (3ull << (3 * kAssassinate)) | (1ull << (3 * kExchange)) |
                                 (1ull << (3 * kSteal)) | (4ull << (3 * kLoseCard1)) | (4ull << (3 * kLoseCard2)) |
                                 (1ull << (3 * kBlock)) | (5ull << 36) | (5ull << 39) | (5ull << 42) | (5ull << 45) |
                                 (5ull << 48) | (5ull << 51);
  // Pass: what the opponent's last action was -> Block 0, Exchange 1, ForeignAid 2, Tax 3, Steal 4
  constexpr uint64_t kPassOn = (0ull << (3 * kBlock)) | (1ull << (3 * kExchange)) | (2ull << (3 * kForeignAid)) |
                               (3ull << (3 * kTax)) | (4ull << (3 * kSteal));
  // Challenge: kind of claim under challenge. Opponent's last action Tax 3, Exchange 4, Assassinate 5, Steal 6; if it
  // is a Block, the mover's own previous action says what was blocked: ForeignAid 0, Assassinate 1, Steal 2.
  constexpr uint64_t kKindDirect = (3ull << (3 * kTax)) | (4ull << (3 * kExchange)) | (5ull << (3 * kAssassinate)) |
                                   (6ull << (3 * kSteal));
  constexpr uint64_t kKindBlocked = (0ull << (3 * kForeignAid)) | (1ull << (3 * kAssassinate)) | (2ull << (3 * kSteal));
  const uint32_t kind = ol == kBlock ? lut3(kKindBlocked, prev_last) : lut3(kKindDirect, ol);
  // the card each kind claims: Duke, Contessa, Captain (or Ambassador, 673-681), Duke, Ambassador, Assassin, Captain
  constexpr uint64_t kClaim = (uint64_t(kDuke) << 0) | (uint64_t(kContessa) << 3) | (uint64_t(kCaptain) << 6) |
                              (uint64_t(kDuke) << 9) | (uint64_t(kAmbassador) << 12) | (uint64_t(kAssassin) << 15) |
                              (uint64_t(kCaptain) << 18);
  const uint32_t claim = lut3(kClaim, kind);
  const uint32_t slot_claim = hand_find(oh, claim << 1);                 // HasFaceDownCard, 379-387
  const uint32_t slot_alt = kind == 2u ? hand_find(oh, kAmbassador << 1) : 4u;
  const bool has_claim = slot_claim < 4u;
  const uint32_t has = (has_claim || slot_alt < 4u) ? 1u : 0u;
  const uint32_t shown = has_claim ? claim : kAmbassador;
  const uint32_t shown_slot = has_claim ? slot_claim : slot_alt;
  uint32_t ev = lut3(kByAction, umin32(a, 17u));
  ev = a == kPass ? kEvPass + lut3(kPassOn, ol) : ev;
  ev = a == kChallenge ? kEvChallenge + 2u * kind + has : ev;
  const uint32_t d = kEventTable[ev];

  // ---- coins ---------------------------------------------------------------------------------------
  const uint32_t transfer = rsh<6>(d) & 3u;
  const uint32_t victim_coins = transfer == 1u ? pw_coins(op) : pw_coins(cp);
  const int k = transfer ? (victim_coins > 1u ? 2 : 1) : 0;
  const int to_mover = transfer == 1u ? k : -k;
  cp = pw_add_coins(cp, static_cast<int>(d & 15u) - 8 + to_mover);
  op = pw_add_coins(op, static_cast<int>(rsh<4>(d) & 3u) - to_mover);

  // ---- cards ---------------------------------------------------------------------------------------
  const bool lose = (d >> 17) & 1u, give_back = (d >> 18) & 1u, replace = (d >> 13) & 1u;
  const uint32_t flip = rsh<15>(d) & 3u;
  // LoseCard k: slot k turns FaceUp, then SortCards (608-613). A player who has to lose a card holds exactly two
  // (a hand of four is always followed by ExchangeReturn first), so sorting is one min/max of the two slots.
  const uint32_t lk = (a - kLoseCard1) & 1u;
  const uint32_t s0 = (ch & 15u) | (lk ^ 1u), s1 = ((ch >> 4) & 15u) | lk;
  const uint32_t h_lose = 0xFF00u | umin32(s0, s1) | (umax32(s0, s1) << 4);
  // ExchangeReturn: slot pairs (i<j) for 12..17: (0,1)(0,2)(0,3)(1,2)(1,3)(2,3), packed 2 bits each; erase j, then i.
  const uint32_t rk = umin32(a - kExchangeReturn12, 5u);
  const uint32_t ri = (0x940u >> (2u * rk)) & 3u;    // 0,0,0,1,1,2
  const uint32_t rj = (0xFB9u >> (2u * rk)) & 3u;    // 1,2,3,2,3,3
  const uint32_t h_back = hand_remove(hand_remove(ch, rj), ri);
  // flips (660-669 / 733-742): every face-down card among slots 0 and 1 turns up; no re-sort, the order cannot change
  const uint32_t flip_hand = flip == 2u ? ch : oh;
  const uint32_t down = ~flip_hand & 0x11u;
  const int n_flipped = static_cast<int>(popc32(down));
  uint32_t ch_new = lose ? h_lose : ch;
  ch_new = give_back ? h_back : ch_new;
  ch_new = flip == 2u ? (ch | down) : ch_new;
  // ChallengeFailReplaceCard (468-486): the first slot holding the shown card face down leaves the other player's hand
  uint32_t oh_new = replace ? hand_remove(oh, shown_slot & 3u) : oh;
  oh_new = flip == 1u ? (oh | down) : oh_new;
  cp = pw_set_hand(cp, ch_new);
  op = pw_set_hand(op, oh_new);
  // deck_: the replaced card goes back by VALUE; ExchangeReturn grows deck_[hand SLOT] -- REFERENCE QUIRK (789-795),
  // reproduced for bit-exact replay.
  uint32_t g = s.g;
  g += replace ? 1u << (4u * shown) : 0u;
  g += give_back ? (1u << (4u * rj)) + (1u << (4u * ri)) : 0u;

  // ---- flags, last action ----------------------------------------------------------------------------
  cp = pw_set_last(cp, a);                                   // every branch of the reference does this first
  cp |= (d << 18) & (1u << 26);              // bit 8 -> bit 26
  cp &= ~((d << 16) & (1u << 26));           // bit 10
  op |= (d << 17) & (1u << 26);              // bit 9

  // ---- who moves next --------------------------------------------------------------------------------
  uint32_t advance = rsh<11>(d) & 3u;
  advance = advance == 3u ? (pw_lost(op) ? 1u : 2u) : advance;
  const uint32_t t = g_turn(g) ^ 1u;
  const uint32_t g_next_turn = (g & ~(kBitTurn | kBitMover)) | (t << 20) | (t << 21) | kBitTurnBegin;   // 1079-1086
  const uint32_t g_next_move = (g ^ kBitMover) & ~kBitTurnBegin;                                        // 1088-1092
  g = advance == 2u ? g_next_turn : (advance == 1u ? g_next_move : g);
  // deals: one per replaced card, two for an Exchange draw, all to the other player (480 / 583-584)
  const uint32_t deals = (replace ? 1u : 0u) + (rsh<13>(d) & 2u);
  const uint32_t g_queued = (g & ~((7u << 24) | kBitQInitial | kBitQPlayer)) | (deals << 24) | (o << 28) | kBitChance;
  g = deals ? g_queued : g;
  s.g = g;
  set_p(s, m, cp);
  set_p(s, o, op);
  // cur_rewards_ (527, 614-615, 662-668, 735-741): zero-sum, stored from player 0's point of view; ++turn_number_
  // with NextPlayerTurn; ++move_number_.
  const int rew_m = lose ? -1 : (flip == 1u ? n_flipped : (flip == 2u ? -n_flipped : 0));
  s.c = c_set_reward0(s.c, m == 0u ? rew_m : -rew_m) + 1u + (advance == 2u ? 1u << 7 : 0u);
}

// One CHANCE move: the chance branch of DoApplyAction, 491-520. `card` must be in the deck. Returns the history
// code of the move (18 + 5 * receiver + card).
COUP_FN uint32_t apply_chance(Env& s, uint32_t card) {
  const uint32_t qn = g_qn(s.g);
  const uint32_t target = (s.g & kBitQInitial) ? (qn & 1u) : ((s.g >> 28) & 1u);  // queue 0,1,0,1
  s.g -= 1u << (4 * card);                                             // deck_[card] -= 1
  const uint32_t tw = get_p(s, target);
  set_p(s, target, pw_set_hand(tw, hand_insert(pw_hand(tw), card << 1)));
  s.g -= 1u << 24;                                                     // pop
  s.g &= qn == 1u ? ~(kBitChance | kBitQInitial) : ~0u;                // 520
  s.c += 1u;                                                           // ++move_number_
  return 18u + 5u * target + card;
}

// Draw a card with probability deck_[c] / sum(deck_) (ChanceOutcomes, 1062-1077) from one uniform
// 32-bit word: r = floor(u * total / 2^32), then the first c whose running count exceeds r. The five 4-bit counts
// never sum past 15 (there are 15 cards in the game; the slot-indexed returns of 789-795 move counts between types but
// conserve the total), so one multiply by 0x11111 leaves all five running counts side by side in nibbles 0..4.
// The comparison runs on four byte lanes at once: the counts of types 0..3 are spread to bytes, one multiply by
// 0x01010101 leaves byte c = deck_[0] + .. + deck_[c], and (0x10 + r - run_c) keeps bit 4 exactly when run_c <= r
// (r <= 14, run_c <= 15: no borrow between bytes).
COUP_FN uint32_t sample_card_g(uint32_t g, uint32_t u) {
  uint32_t x = ((g & 0xFFFFu) | (g << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;                                   // byte c = deck_[c], c < 4
  const uint32_t run = x * 0x01010101u;
  const uint32_t r = umulhi32(u, (run >> 24) + ((g >> 16) & 15u));
  return popc32(((r * 0x01010101u + 0x10101010u) - run) & 0x10101010u);
}
COUP_FN uint32_t sample_card(const Env& s, uint32_t u) { return sample_card_g(s.g, u); }

// CoupState::CoupState, 393-428.
COUP_FN Env initial_state() {
  Env s;
  s.p[0] = 0xFFFFu | (1u << 16) | (kNoAction << 21);
  s.p[1] = 0xFFFFu | (2u << 16) | (kNoAction << 21);
  s.g = 0x33333u | kBitTurnBegin | kBitChance | (4u << 24) | kBitQInitial;
  s.c = 2u << 14;
  return s;
}

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based: no per-env RNG state ---------------------
COUP_FN uint32_t mulhi32(uint32_t a, uint32_t b) {
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
}
COUP_FN uint4 philox4x32_10(uint4 ctr, uint2 key) {
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
// three chance draws that can follow one player action -- or, when the action ends the episode and the env is re-dealt in
// place, the four initial deals (y, y * 15 mod 2^32, z, w); purpose 1: the four deals of an explicit reset (and of the
// re-deal after a game cut by the move cap in the middle of a deal sequence); purpose 2-4: CFR expansion; 5: reservoir.
COUP_FN uint4 env_random(uint64_t seed, uint64_t global_env, uint64_t step, uint32_t purpose) {
  uint2 key = make_uint2(static_cast<uint32_t>(global_env),
                         static_cast<uint32_t>(global_env >> 32) ^ static_cast<uint32_t>(seed));
  uint4 ctr = make_uint4(static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32), purpose,
                         static_cast<uint32_t>(seed >> 32));
  return philox4x32_10(ctr, key);
}

COUP_FN uint32_t pick(const uint4& r, int k) {
  return k == 0 ? r.x : k == 1 ? r.y : k == 2 ? r.z : r.w;
}

// k-th (0-based) set bit of a mask with more than k bits set, k <= 7 (a Coup state has at most 7 legal actions, a deck
// 5 card types): clear the lowest set bit k times, without a loop-carried branch.
COUP_FN uint32_t kth_set_bit(uint32_t mask, uint32_t k) {
#pragma unroll
  for (uint32_t i = 0; i < 7; ++i) mask = i < k ? mask & (mask - 1u) : mask;
  return ffs32(mask) - 1u;
}

// Uniform legal action (benchmark_game.cc:96-99) from one uniform 32-bit word.
COUP_FN uint32_t sample_action(uint32_t legal, uint32_t u) {
  return kth_set_bit(legal, umulhi32(u, popc32(legal)));
}

// The state after the four initial deals (CoupState ctor 393-428 + four chance moves 491-520, queue 0,1,0,1) in closed
// form, from the four words of the reset Philox block. Same outcome as four sample_card / apply_chance rounds on
// initial_state(): give the 15 cards ids 0..14 with type = id / 3; "the first type whose running count exceeds r"
// among the remaining cards IS the type of the r-th remaining card in id order, so each draw is the r-th id not yet
// taken. `codes` receives the four history codes, 5 bits each.
COUP_FN Env dealt_initial_state(const uint4& r, uint32_t& codes) {
  const uint32_t p1 = umulhi32(r.x, 15u);
  uint32_t p2 = umulhi32(r.y, 14u);
  p2 += p2 >= p1 ? 1u : 0u;
  const uint32_t lo = umin32(p1, p2), hi = umax32(p1, p2);
  uint32_t p3 = umulhi32(r.z, 13u);
  p3 += p3 >= lo ? 1u : 0u;
  p3 += p3 >= hi ? 1u : 0u;
  const uint32_t a = umin32(lo, p3), c = umax32(hi, p3), b = lo + hi + p3 - a - c;   // the three taken ids, ascending
  uint32_t p4 = umulhi32(r.w, 12u);
  p4 += p4 >= a ? 1u : 0u;
  p4 += p4 >= b ? 1u : 0u;
  p4 += p4 >= c ? 1u : 0u;
  const uint32_t t1 = (p1 * 11u) >> 5, t2 = (p2 * 11u) >> 5, t3 = (p3 * 11u) >> 5, t4 = (p4 * 11u) >> 5;   // id / 3
  Env s;
  // deals 1 and 3 go to player 0, deals 2 and 4 to player 1; hands sorted (all four cards face down)
  s.p[0] = 0xFF00u | (umin32(t1, t3) << 1) | (umax32(t1, t3) << 5) | (1u << 16) | (kNoAction << 21);
  s.p[1] = 0xFF00u | (umin32(t2, t4) << 1) | (umax32(t2, t4) << 5) | (2u << 16) | (kNoAction << 21);
  s.g = (0x33333u - (1u << (4u * t1)) - (1u << (4u * t2)) - (1u << (4u * t3)) - (1u << (4u * t4))) | kBitTurnBegin;
  s.c = (2u << 14) + 4u;
  codes = (18u + t1) | ((23u + t2) << 5) | ((18u + t3) << 10) | ((23u + t4) << 15);
  return s;
}

// Resolve every pending chance node (rl_environment._sample_external_events, rl_environment.py:369-382) of a state
// with the generic draw: `rnd` supplies draws rnd[first..]; forced[k] < 5 overrides the k-th draw. Stops at terminal
// (the move cap can hit in the middle of a deal sequence, coup.cc:990). `codes`/`n_codes` collect the history codes.
COUP_FN uint32_t resolve_chance(Env& s, const uint4& rnd, int first, const uint8_t* forced, uint32_t& codes,
                                uint32_t& n_codes) {
  uint32_t k = 0;
  while (g_chance(s.g) && !is_terminal(s) && k < 4u) {
    uint32_t card = sample_card(s, pick(rnd, first + static_cast<int>(k)));
    if (forced != nullptr) {
      const uint32_t f = forced[k];
      if (f < 5u && g_deck(s.g, f) != 0) card = f;
    }
    codes |= apply_chance(s, card) << (5u * n_codes);
    ++n_codes;
    ++k;
  }
  return k;
}

// ---- observer head --------------------------------------------------------------------------------
// The first 62 floats of both tensors (CoupObserver::WriteTensor 248-279): bit q of the returned mask
// is 1 where float q is 1; floats 60/61 are the raw coin counts (207-213), returned separately.
//   [0,2) observer  [2,22) p1_cards[4][5]  [22,42) p2_cards[4][5]  [42,44) cur_move_player
//   [44,60) cards_state[2][4][2]  [60,62) coins
// `vis` selects the IIGObservationType (observer.h:270-315) of the general observer, CoupGame::MakeObserver
// (1132-1141): 0 = private_info kSinglePlayer + public_info, the type of both built-in tensors; kVisPrivateNone /
// kVisPrivateAll change whose face-down cards show (185-188, 258-265); kVisNoPublic drops everything public: face-up
// cards, cur_move_player, cards_state (267-279) -- the tensor then ends after element 41.
enum : uint32_t { kVisPrivateNone = 1u, kVisPrivateAll = 2u, kVisNoPublic = 4u };
COUP_FN uint64_t head_mask(const Env& s, uint32_t observer, bool terminal, uint32_t vis = 0u) {
  uint64_t mask = 1ull << observer;                                    // WritePlayer, 160-165
  const bool pub = (vis & kVisNoPublic) == 0;
#pragma unroll
  for (uint32_t pl = 0; pl < 2; ++pl) {
    const uint32_t h = pw_hand(s.p[pl]);
    const bool priv = (vis & kVisPrivateAll) != 0 || ((vis & kVisPrivateNone) == 0 && pl == observer);
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
      const uint32_t key = hand_slot(h, i);
      if (key != 15u) {
        const uint32_t up = key & 1u, value = key >> 1;
        // WritePlayerCardsValue 178-191: a face-down card shows to whom private info is granted, a face-up one to
        // whoever sees public info
        if (up ? pub : priv) mask |= 1ull << (2u + 20u * pl + 5u * i + value);
        if (pub) mask |= 1ull << (44u + (pl * 4u + i) * 2u + up);     // WriteCardsState, 194-204
      }
    }
  }
  if (!terminal && pub) mask |= 1ull << (42u + g_mover(s.g));         // 268-276
  return mask;
}

// Observation tensor tail [62,98): last_action[2][18] one-hot (WriteLastAction, 217-225), as bits
// 0..35 of the returned mask.
COUP_FN uint64_t last_action_mask(const Env& s) {
  uint64_t m = 0;
  uint32_t l0 = pw_last(s.p[0]), l1 = pw_last(s.p[1]);
  if (l0 != kNoAction) m |= 1ull << l0;
  if (l1 != kNoAction) m |= 1ull << (18u + l1);
  return m;
}

// Column of history row code `code` as seen by `observer` (WriteActionHistory, 230-245): player moves
// are public, a deal is visible only to the player who received it. 31 = all-zero row.
COUP_FN uint32_t history_column(uint32_t code, uint32_t observer) {
  if (code < 18u) return code;
  const uint32_t base = 18u + 5u * observer;
  return (code >= base && code < base + 5u) ? code - base : 31u;
}

// 64-bit finaliser used by the position-keyed tensor hash (see coup_tensor_row_hash).
COUP_FN uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

}  // namespace coup
