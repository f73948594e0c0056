"""Single-state `Game` / `State` mirror of the reference plugin surface for "coup".

Same method names, argument meaning and error behaviour as the pyspiel objects the reference's tests
and experiments use (`open_spiel/python/pybind11/pyspiel.cc:263-405` binding
`CoupGame`/`CoupState`, `open_spiel/games/coup.h:111-231`), so code written against
`pyspiel.load_game("coup")` runs unchanged against `load_game("coup")` from this module.

Every rule still runs on the GPU: a `CoupGame` owns a small device slab (one slot per live `CoupState`)
behind the C ABI and each call launches the same kernels the batched environment uses, one move at a
time with explicit chance nodes (`coup_vec_apply_move`). What the reference keeps on the host stays on
the host here too: the `(player, action)` history of `State::ApplyAction` (`spiel.cc:322-332`), string
formatting (`coup.cc:60-135,290-373,945-987`), (de)serialisation (`spiel.cc:297-311,393-432`) and the
probability arithmetic of `ChanceOutcomes` (`coup.cc:1062-1077`) on the device's deck counts.
This path is for API parity and debugging, not for throughput (use `CoupVectorEnv` for that).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (CHANCE_PLAYER_ID, INFO_STATE_SIZE, MAX_CHANCE_NODES_IN_HISTORY, MAX_GAME_LENGTH, NUM_DISTINCT_ACTIONS,
                   OBSERVATION_SIZE, PLAYER_0, TERMINAL_PLAYER_ID, CoupError, check)
from .vector_env import CoupVectorEnv, _stream_ptr, decode_history, unpack_states

SpielError = CoupError  # pyspiel raises SpielError where the C++ library would abort (pyspiel.cc:620-626)

CARD_NAMES = ["Assassin", "Ambassador", "Captain", "Contessa", "Duke"]                      # coup.cc:60-77
CARD_STATE_NAMES = {-1: "None", 0: "FaceDown", 1: "FaceUp"}                                 # coup.cc:79-90
ACTION_NAMES = ["Income", "ForeignAid", "Coup", "Tax", "Assassinate", "Exchange", "Steal", "LoseCard1", "LoseCard2",
                "Pass", "Block", "Challenge", "ExchangeReturn12", "ExchangeReturn13", "ExchangeReturn14",
                "ExchangeReturn23", "ExchangeReturn24", "ExchangeReturn34"]               # coup.cc:92-135


def card_to_string(card):
    return "-" if card < 0 else CARD_NAMES[card]


def action_name(action):
    return "None" if action < 0 else ACTION_NAMES[action]


class IIGObservationType:
    """open_spiel/observer.h:270-315. private_info: 0 kNone, 1 kSinglePlayer, 2 kAllPlayers."""

    def __init__(self, public_info=True, perfect_recall=False, private_info=1):
        self.public_info, self.perfect_recall, self.private_info = bool(public_info), bool(perfect_recall), int(private_info)


DEFAULT_OBS_TYPE = IIGObservationType(True, False, 1)       # observer.h:287-290
INFO_STATE_OBS_TYPE = IIGObservationType(True, True, 1)     # observer.h:294-297
PUBLIC_OBS_TYPE = IIGObservationType(True, False, 0)        # observer.h:300-303
PRIVATE_OBS_TYPE = IIGObservationType(False, False, 1)      # observer.h:312-315


# ---- host-side string formatting (pure functions of a plain "view" of the state) -------------------
def make_view(turn_number, cur_player_move, cards, coins, last_action, history):
    """cards: per player list of (value, state); history: list of (player, action, deal_target)."""
    return {"turn_number": int(turn_number), "cur_player_move": int(cur_player_move),
            "cards": [[(int(v), int(s)) for v, s in c] for c in cards], "coins": [int(c) for c in coins],
            "last_action": [int(a) for a in last_action], "history": [(int(p), int(a), int(t)) for p, a, t in history]}


def _cards_block(view, p, show_value, show_state):
    out = "P%d\n        Card         State\n" % (p + 1)
    for c, (value, state) in enumerate(view["cards"][p]):
        name = card_to_string(value if show_value(state) else -1)
        out += "Card %d: %s%s| %s\n" % (c + 1, name, " " * (11 - len(name)), CARD_STATE_NAMES[state if show_state else -1])
    return out


def state_to_string(view):
    """CoupState::ToString, coup.cc:945-987."""
    out = "Turn: %d\nMove: P%d\n" % (view["turn_number"], view["cur_player_move"] + 1)
    for p in range(2):
        out += _cards_block(view, p, lambda s: True, True)
        out += "Coins: %d\nLast Action: %s\n\n" % (view["coins"][p], action_name(view["last_action"][p]))
    out += "Action Sequence: "
    h = view["history"]
    for i, (player, action, _) in enumerate(h):
        out += ("PC-" + card_to_string(action)) if player == CHANCE_PLAYER_ID else ("P%d-%s" % (player + 1, action_name(action)))
        if i < len(h) - 1:
            out += ", "
    return out + "\n"


def observer_string(view, player, obs_type):
    """CoupObserver::StringFrom, coup.cc:290-373."""
    pub, recall, priv = obs_type.public_info, obs_type.perfect_recall, obs_type.private_info
    out = "Observer: P%d\n" % (player + 1)
    if pub:
        out += "Turn: %d\nMove: P%d\n" % (view["turn_number"], view["cur_player_move"] + 1)
    for p in range(2):
        if pub or priv == 2 or (priv == 1 and player == p):
            def show(state, p=p):
                return (pub and state == 1) or (priv == 1 and p == player and state == 0) or (priv == 2 and state == 0)
            out += _cards_block(view, p, show, pub)
        if pub:
            out += "Coins: %d\n" % view["coins"][p]
            out += ("Last Action: %s\n\n" % action_name(view["last_action"][p])) if not recall else "\n"
    if pub and recall:
        out += "Action Sequence: "
        h = view["history"]
        for i, (pl, action, target) in enumerate(h):
            if pl == CHANCE_PLAYER_ID:
                if target == player:        # only deals to the observing player are shown
                    out += "PC-" + card_to_string(action)
                    if i < len(h) - 1:
                        out += ", "
            else:
                out += "P%d-%s" % (pl + 1, action_name(action))
                if i < len(h) - 1:
                    out += ", "
        out += "\n"
    return out


class GameType:
    """kGameType, coup.cc:38-52."""
    short_name = "coup"
    long_name = "Coup"
    dynamics = "SEQUENTIAL"
    chance_mode = "EXPLICIT_STOCHASTIC"
    information = "IMPERFECT_INFORMATION"
    utility = "ZERO_SUM"
    reward_model = "REWARDS"
    max_num_players = 2
    min_num_players = 2
    provides_information_state_string = True
    provides_information_state_tensor = True
    provides_observation_string = True
    provides_observation_tensor = True
    parameter_specification = {}


class CoupGame:
    """`CoupGame`, coup.h:199-231. `capacity` = number of simultaneously live states."""

    def __init__(self, params=None, device=0, capacity=256):
        if params:
            raise SpielError("Unknown parameter(s) for game coup: %s" % sorted(params))
        self._vec = CoupVectorEnv(capacity, seed=0, device=device)
        self._free = list(range(capacity - 1, -1, -1))
        self._moves = torch.full((capacity,), 0xFF, dtype=torch.uint8, device=self._vec.device)
        self._info = torch.empty((2 * capacity, INFO_STATE_SIZE), dtype=torch.float32, device=self._vec.device)
        self._obs = torch.empty((2 * capacity, OBSERVATION_SIZE), dtype=torch.float32, device=self._vec.device)

    # -- static facts ------------------------------------------------------------------------------
    def get_type(self):
        return GameType

    def get_parameters(self):
        return {}

    def num_distinct_actions(self):
        return NUM_DISTINCT_ACTIONS

    def policy_tensor_shape(self):
        return [NUM_DISTINCT_ACTIONS]

    def max_chance_outcomes(self):
        return _lib.MAX_CHANCE_OUTCOMES

    def num_players(self):
        return 2

    def min_utility(self):
        return -2.0

    def max_utility(self):
        return 2.0

    def utility_sum(self):
        return 0.0

    def information_state_tensor_shape(self):
        return [INFO_STATE_SIZE]

    def information_state_tensor_size(self):
        return INFO_STATE_SIZE

    def observation_tensor_shape(self):
        return [OBSERVATION_SIZE]

    def observation_tensor_size(self):
        return OBSERVATION_SIZE

    def max_game_length(self):
        return MAX_GAME_LENGTH

    def max_chance_nodes_in_history(self):
        return MAX_CHANCE_NODES_IN_HISTORY

    def max_move_number(self):
        return MAX_GAME_LENGTH + MAX_CHANCE_NODES_IN_HISTORY      # spiel.h:888-890

    def action_to_string(self, player, action):
        """CoupGame::ActionToString, coup.cc:1143-1149."""
        if player == CHANCE_PLAYER_ID:
            return "Chance drawn card:" + card_to_string(action)
        return action_name(action)

    def __str__(self):
        return "coup()"

    # -- states ------------------------------------------------------------------------------------
    def _alloc(self):
        if not self._free:
            raise SpielError("CoupGame: state capacity exhausted (raise `capacity` in load_game)")
        return self._free.pop()

    def _release(self, slot):
        self._free.append(slot)

    def new_initial_state(self):
        slot = self._alloc()
        mask = torch.zeros(self._vec.num_envs, dtype=torch.uint8, device=self._vec.device)
        mask[slot] = 1
        check(self._vec._lib.coup_vec_new_initial_state(self._vec._h, C.c_void_p(mask.data_ptr()), _stream_ptr(self._vec.device)))
        return CoupState(self, slot, [])

    def deserialize_state(self, text):
        """Game::DeserializeState, spiel.cc:393-432: one action id per line."""
        state = self.new_initial_state()
        for line in text.split("\n"):
            if line:
                state.apply_action(int(line))
        return state

    def make_observer(self, iig_obs_type=None, params=None):
        return CoupObserver(iig_obs_type or DEFAULT_OBS_TYPE)


class CoupObserver:
    """String side of CoupObserver (coup.cc:150-377) for any IIGObservationType; the two tensor types the
    kernels implement are the default (98) and info-state (2492) ones."""

    def __init__(self, obs_type):
        self.obs_type = obs_type

    def string_from(self, state, player):
        return observer_string(state._view(), player, self.obs_type)


class CoupState:
    """`CoupState`, coup.h:111-150, one slot of the game's device slab."""

    def __init__(self, game, slot, history):
        self._game, self._slot, self._history = game, slot, list(history)

    def __del__(self):
        try:
            self._game._release(self._slot)
        except Exception:
            pass

    # -- device reads ------------------------------------------------------------------------------
    def _vecenv(self):
        return self._game._vec

    def _word(self):
        return int(self._vecenv().step_word[self._slot]) & 0xFFFFFFFF

    def _packed(self):
        return unpack_states(self._vecenv().state[self._slot:self._slot + 1].cpu().numpy())

    def get_game(self):
        return self._game

    def num_players(self):
        return 2

    def num_distinct_actions(self):
        return NUM_DISTINCT_ACTIONS

    def current_player(self):
        """coup.cc:458-466."""
        w = self._word()
        if (w >> 19) & 1:
            return TERMINAL_PLAYER_ID
        if (w >> 27) & 1:
            return CHANCE_PLAYER_ID
        return (w >> 18) & 1

    def is_terminal(self):
        return self.current_player() == TERMINAL_PLAYER_ID

    def is_chance_node(self):
        return self.current_player() == CHANCE_PLAYER_ID

    def is_simultaneous_node(self):
        return False

    def is_player_node(self):
        return self.current_player() >= 0

    def move_number(self):
        return len(self._history)

    def legal_actions(self, player=None):
        """coup.cc:824-938; for a given player: empty unless it is that player's node (spiel.h:255-261)."""
        cur = self.current_player()
        if player is not None and player != cur:
            return []
        w = self._word()
        return [a for a in range(NUM_DISTINCT_ACTIONS) if (w >> a) & 1]

    def legal_actions_mask(self, player=None):
        """spiel.cc:371-377: length 18 at decision nodes, length MaxChanceOutcomes at chance nodes."""
        length = _lib.MAX_CHANCE_OUTCOMES if self.is_chance_node() else NUM_DISTINCT_ACTIONS
        legal = set(self.legal_actions(player))
        return [1 if a in legal else 0 for a in range(length)]

    def chance_outcomes(self):
        """coup.cc:1062-1077: deck_[i] / sum(deck_) for the card types still in the deck."""
        if not self.is_chance_node():
            raise SpielError("ChanceOutcomes() called on a non-chance node")
        deck = self._packed()["deck"][0]
        total = float(deck.sum())
        return [(int(i), int(deck[i]) / total) for i in range(5) if deck[i] > 0]

    def apply_action(self, action):
        """State::ApplyAction (spiel.cc:322-332) + CoupState::DoApplyAction (coup.cc:490-809) on the device."""
        action = int(action)
        player = self.current_player()
        if player == TERMINAL_PLAYER_ID:
            raise SpielError("ApplyAction called on a terminal state")
        if action not in self.legal_actions():
            raise SpielError("Invalid action %d at this node (legal: %s)" % (action, self.legal_actions()))
        g = self._game
        g._moves[self._slot] = action
        check(g._vec._lib.coup_vec_apply_move(g._vec._h, C.c_void_p(g._moves.data_ptr()), _stream_ptr(g._vec.device)))
        g._moves[self._slot] = 0xFF
        self._history.append((player, action))

    def child(self, action):
        c = self.clone()
        c.apply_action(action)
        return c

    def clone(self):
        """coup.cc:1058-1060."""
        g = self._game
        slot = g._alloc()
        check(g._vec._lib.coup_vec_copy_env(g._vec._h, self._slot, slot, _stream_ptr(g._vec.device)))
        return CoupState(g, slot, self._history)

    def rewards(self):
        r0 = ((self._word() >> 21) & 7) - 2
        return [float(r0), float(-r0)]

    def returns(self):
        r0 = ((self._word() >> 24) & 7) - 2
        return [float(r0), float(-r0)]

    def player_return(self, player):
        return self.returns()[player]

    def player_reward(self, player):
        return self.rewards()[player]

    # -- history / serialisation (host-side bookkeeping, as in State) --------------------------------
    def history(self):
        return [a for _, a in self._history]

    def full_history(self):
        return list(self._history)

    def history_str(self):
        return ", ".join(str(a) for _, a in self._history)

    def serialize(self):
        """State::Serialize, spiel.cc:297-311."""
        return "\n".join(str(a) for _, a in self._history) + "\n"

    def action_to_string(self, *args):
        player, action = args if len(args) == 2 else (self.current_player(), args[0])
        return self._game.action_to_string(player, action)

    # -- tensors -------------------------------------------------------------------------------------
    def _check_player(self, player):
        if player is None:
            player = self.current_player()
        if not 0 <= player < 2:
            raise SpielError("player %r out of range" % (player,))   # SPIEL_CHECK_GE/LT, coup.cc:251-252
        return player

    def information_state_tensor(self, player=None):
        """coup.cc:1044-1049, as a list of 2492 floats like pyspiel."""
        player = self._check_player(player)
        g = self._game
        g._vec.information_state_tensor(_lib.PLAYER_BOTH, out=g._info)
        return g._info[2 * self._slot + player].cpu().tolist()

    def observation_tensor(self, player=None):
        """coup.cc:1051-1056."""
        player = self._check_player(player)
        g = self._game
        g._vec.observation_tensor(_lib.PLAYER_BOTH, out=g._obs)
        return g._obs[2 * self._slot + player].cpu().tolist()

    # -- strings -------------------------------------------------------------------------------------
    def _view(self):
        st = self._packed()
        cards = []
        for p in range(2):
            hand = [int(k) for k in st["hands"][0, p] if k != 15]
            cards.append([(k >> 1, k & 1) for k in hand])
        hist_words = self._vecenv().history[self._slot:self._slot + 1].cpu().numpy().view(np.uint32)
        _, targets = decode_history(hist_words, np.array([len(self._history)]))[0]
        history = [(p, a, int(targets[i])) for i, (p, a) in enumerate(self._history)]
        return make_view(st["turn_number"][0], st["cur_player_move"][0], cards, st["coins"][0], st["last_action"][0], history)

    def to_string(self):
        return state_to_string(self._view())

    __str__ = to_string

    def information_state_string(self, player=None):
        """coup.cc:1034-1037."""
        return observer_string(self._view(), self._check_player(player), INFO_STATE_OBS_TYPE)

    def observation_string(self, player=None):
        """coup.cc:1039-1042."""
        return observer_string(self._view(), self._check_player(player), DEFAULT_OBS_TYPE)

    # -- CoupState convenience accessors (coup.h:139-143) -----------------------------------------------
    def get_cards_value(self, player):
        return [v for v, _ in self._view()["cards"][player]]

    def get_cards_state(self, player):
        return [s for _, s in self._view()["cards"][player]]

    def get_coins(self, player):
        return self._view()["coins"][player]

    def get_last_action(self, player):
        return self._view()["last_action"][player]


_games = {}


def load_game(name="coup", params=None, device=0, capacity=256):
    """`pyspiel.load_game("coup")` (pyspiel.cc:557-566). Accepts "coup" or "coup()"."""
    if name not in ("coup", "coup()"):
        raise SpielError("Unknown game '%s'. This library provides exactly one game: coup" % name)
    return CoupGame(params, device=device, capacity=capacity)


def serialize_game_and_state(game, state):
    """spiel.cc:413-432 (SerializeGameAndState)."""
    return ("# Automatically generated by OpenSpiel SerializeGameAndState\n[Meta]\nVersion: 1\n\n[Game]\n%s\n[State]\n%s\n"
            % (str(game), state.serialize()))
