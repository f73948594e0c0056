"""Single-game `Environment` with the interface of the reference's `open_spiel/python/rl_environment.py`,
on top of the GPU-backed `Game`/`State` mirror (`spiel.py`). This is what the reference's experiment scripts
drive (`coup_experiments/scripts/nfsp.py:100-144`, `algorithms/rl_response.py:65-91`): `reset()` /
`step([action])` returning `TimeStep(observations, rewards, discounts, step_type)`.

It exists for drop-in compatibility of batch-1 code. Anything throughput-sensitive should use
`vector_env.CoupVectorEnv` / `selfplay.SelfPlayDataGen`, which do the same work for 2^20 games per call.
"""
import collections
import enum

import numpy as np

from . import spiel

SIMULTANEOUS_PLAYER_ID = -2   # pyspiel.PlayerId.SIMULTANEOUS (spiel_globals.h:30)


class TimeStep(collections.namedtuple("TimeStep", ["observations", "rewards", "discounts", "step_type"])):
    """rl_environment.py:57-97."""
    __slots__ = ()

    def first(self):
        return self.step_type == StepType.FIRST

    def mid(self):
        return self.step_type == StepType.MID

    def last(self):
        return self.step_type == StepType.LAST

    def is_simultaneous_move(self):
        return self.observations["current_player"] == SIMULTANEOUS_PLAYER_ID

    def current_player(self):
        return self.observations["current_player"]


class StepType(enum.Enum):
    """rl_environment.py:100-114."""
    FIRST = 0
    MID = 1
    LAST = 2

    def first(self):
        return self is StepType.FIRST

    def mid(self):
        return self is StepType.MID

    def last(self):
        return self is StepType.LAST


class ChanceEventSampler:
    """rl_environment.py:119-131: samples a chance outcome with `np.random.RandomState.choice`."""

    def __init__(self, seed=None):
        self.seed(seed)

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def __call__(self, state):
        actions, probs = zip(*state.chance_outcomes())
        return self._rng.choice(actions, p=probs)


class ObservationType(enum.Enum):
    OBSERVATION = 0
    INFORMATION_STATE = 1


class Environment:
    """rl_environment.py:140-480 for game "coup" (turn-based, 2 players)."""

    def __init__(self, game="coup", discount=1.0, chance_event_sampler=None, observation_type=None,
                 include_full_state=False, enable_legality_check=False, device=0, **kwargs):
        self._game = spiel.load_game(game, device=device) if isinstance(game, str) else game
        self._chance_event_sampler = chance_event_sampler or ChanceEventSampler()
        self._include_full_state = include_full_state
        self._enable_legality_check = enable_legality_check
        self._num_players = self._game.num_players()
        self._state = None
        self._should_reset = True
        self._discounts = [discount] * self._num_players
        # default: information state when the game provides it (rl_environment.py:194-207)
        self._use_observation = observation_type == ObservationType.OBSERVATION

    def seed(self, seed=None):
        self._chance_event_sampler.seed(seed)

    def _observations(self):
        """The `observations` dict of a time step (rl_environment.py:224-262): per player the tensor and the legal
        actions (empty for the player who is not to move), then the mover and, on request, the serialised state."""
        state = self._state
        tensor = state.observation_tensor if self._use_observation else state.information_state_tensor
        players = range(self.num_players)
        obs = {"info_state": [tensor(p) for p in players], "legal_actions": [state.legal_actions(p) for p in players],
               "current_player": state.current_player(), "serialized_state": []}
        if self._include_full_state:
            obs["serialized_state"] = spiel.serialize_game_and_state(self._game, state)
        return obs

    def get_time_step(self):
        """rl_environment.py:219-268: MID, or LAST (discounts zeroed) once the state is terminal."""
        last = self._state.is_terminal()
        self._should_reset = last
        return TimeStep(observations=self._observations(), rewards=list(self._state.rewards()),
                        discounts=[0.0] * self.num_players if last else self._discounts,
                        step_type=StepType.LAST if last else StepType.MID)

    def _check_legality(self, actions):
        legal_actions = self._state.legal_actions()
        if actions[0] not in legal_actions:
            raise RuntimeError(f"step() called on illegal action {actions[0]}")

    def step(self, actions):
        """rl_environment.py:282-322: one player action, then every following chance node."""
        assert len(actions) == self.num_actions_per_step, "Invalid number of actions! Expected {}".format(self.num_actions_per_step)
        if self._should_reset:
            return self.reset()
        if self._enable_legality_check:
            self._check_legality(actions)
        self._state.apply_action(actions[0])
        self._sample_external_events()
        return self.get_time_step()

    def reset(self):
        """rl_environment.py:324-367: a new initial state with its chance nodes resolved; FIRST carries no rewards."""
        self._should_reset = False
        self._state = self._game.new_initial_state()
        self._sample_external_events()
        return TimeStep(observations=self._observations(), rewards=None, discounts=None, step_type=StepType.FIRST)

    def _sample_external_events(self):
        """rl_environment.py:369-382."""
        while self._state.is_chance_node():
            outcome = self._chance_event_sampler(self._state)
            self._state.apply_action(int(outcome))

    def observation_spec(self):
        """rl_environment.py:384-414."""
        size = (self._game.observation_tensor_size() if self._use_observation
                else self._game.information_state_tensor_size())
        return dict(info_state=tuple([size]), legal_actions=(self._game.num_distinct_actions(),), current_player=(),
                    serialized_state=())

    def action_spec(self):
        """rl_environment.py:416-432."""
        return dict(num_actions=self._game.num_distinct_actions(), min=0, max=self._game.num_distinct_actions() - 1,
                    dtype=int)

    @property
    def use_observation(self):
        return self._use_observation

    @property
    def name(self):
        return self._game.get_type().short_name

    @property
    def num_players(self):
        return self._num_players

    @property
    def num_actions_per_step(self):
        return 1   # turn-based (rl_environment.py:448-449)

    @property
    def is_turn_based(self):
        return True

    @property
    def max_game_length(self):
        return self._game.max_game_length()

    @property
    def is_chance_node(self):
        return self._state.is_chance_node()

    @property
    def game(self):
        return self._game

    def set_state(self, new_state):
        assert new_state.get_game() is self._game or str(new_state.get_game()) == str(self._game)
        self._state = new_state

    @property
    def get_state(self):
        return self._state
