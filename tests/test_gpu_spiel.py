"""GPU tests (-m gpu) of the single-state Game/State mirror (open_spiel_coup_b200/spiel.py), written the way
the reference's own tests read: the 14 scenario tests of coup_test.cc through `load_game("coup")`, the
golden playthrough coup.txt including every string, and the generic conformance checks of
tests/basic_tests.cc (legal actions sorted, empty for the non-acting player, clone equality,
serialize/deserialize round trip, sum of rewards == returns, utilities in range and zero-sum)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

from open_spiel_coup_b200 import spiel  # noqa: E402
from open_spiel_coup_b200.spiel import SpielError, load_game  # noqa: E402


@pytest.fixture(scope="module")
def game():
    return load_game("coup")


def test_load_game_and_static_facts(game, playthrough):
    h = playthrough["header"]
    assert str(game) == "coup()"
    assert game.num_distinct_actions() == int(h["NumDistinctActions"])
    assert game.policy_tensor_shape() == [18]
    assert game.max_chance_outcomes() == int(h["MaxChanceOutcomes"])
    assert game.get_parameters() == {}
    assert game.num_players() == int(h["NumPlayers"])
    assert game.min_utility() == float(h["MinUtility"]) and game.max_utility() == float(h["MaxUtility"])
    assert game.utility_sum() == float(h["UtilitySum"])
    assert game.information_state_tensor_size() == int(h["InformationStateTensorSize"])
    assert game.observation_tensor_size() == int(h["ObservationTensorSize"])
    assert game.max_game_length() == int(h["MaxGameLength"])
    t = game.get_type()
    assert (t.short_name, t.long_name, t.chance_mode, t.dynamics, t.information, t.utility, t.reward_model) == \
        ("coup", "Coup", "EXPLICIT_STOCHASTIC", "SEQUENTIAL", "IMPERFECT_INFORMATION", "ZERO_SUM", "REWARDS")
    with pytest.raises(SpielError):
        load_game("kuhn_poker")


def test_scenario_kats_like_coup_test_cc(game, kat_scenarios):
    for kat in kat_scenarios:
        state = game.new_initial_state()
        for step in kat["steps"]:
            if "apply" in step:
                state.apply_action(step["apply"])
            exp = step.get("expect", {})
            name = kat["name"]
            if "current_player" in exp:
                assert state.current_player() == exp["current_player"], name
            if "legal_actions" in exp:
                assert state.legal_actions() == exp["legal_actions"], name
            if "coins" in exp:
                for p, c in enumerate(exp["coins"]):
                    assert c is None or state.get_coins(p) == c, name
            if "num_cards" in exp:
                for p, c in enumerate(exp["num_cards"]):
                    assert c is None or len(state.get_cards_value(p)) == c, name
            if exp.get("all_face_down"):
                assert all(s == 0 for p in (0, 1) for s in state.get_cards_state(p)), name
            if "last_action" in exp:
                assert [state.get_last_action(0), state.get_last_action(1)] == exp["last_action"], name
            if "face_up" in exp:
                for p, slot in exp["face_up"]:
                    assert state.get_cards_state(p)[slot] == 1, name
            if "cards_p0" in exp:
                assert list(zip(state.get_cards_value(0), state.get_cards_state(0))) == [tuple(c) for c in exp["cards_p0"]], name
            if "is_terminal" in exp:
                assert state.is_terminal() == exp["is_terminal"], name
            if "rewards" in exp:
                assert state.rewards() == exp["rewards"], name
            if "returns" in exp:
                assert state.returns() == exp["returns"], name


def _dense(pairs, n):
    t = np.zeros(n, np.float32)
    for k, v in pairs:
        t[k] = v
    return t


def test_playthrough_everything(game, playthrough):
    """integration_tests/playthroughs/coup.txt replayed move by move (explicit chance nodes)."""
    state = game.new_initial_state()
    full = 0
    for st in playthrough["states"]:
        if "CurrentPlayer" in st:
            assert state.current_player() == st["CurrentPlayer"]
        if "IsTerminal" in st:
            assert state.is_terminal() == st["IsTerminal"]
        if "IsChanceNode" in st:
            assert state.is_chance_node() == st["IsChanceNode"]
        if "History" in st:
            assert state.history() == st["History"]
        if "LegalActions" in st:
            assert state.legal_actions() == st["LegalActions"]
        if "ChanceOutcomes" in st:
            got = state.chance_outcomes()
            assert [a for a, _ in got] == [a for a, _ in st["ChanceOutcomes"]]
            np.testing.assert_allclose([p for _, p in got], [p for _, p in st["ChanceOutcomes"]], atol=1e-12)
        if "Rewards" in st:
            assert state.rewards() == st["Rewards"]
        if "Returns" in st:
            assert state.returns() == st["Returns"]
        if "InformationStateString(0)" in st:
            full += 1
            assert [ln.rstrip() for ln in str(state).split("\n")] == [ln.rstrip() for ln in st["ToString"].split("\n")]
            for p in (0, 1):
                assert state.information_state_string(p) == st[f"InformationStateString({p})"]
                assert state.observation_string(p) == st[f"ObservationString({p})"]
                assert game.make_observer(spiel.PRIVATE_OBS_TYPE).string_from(state, p) == st[f"PrivateObservationString({p})"]
                np.testing.assert_array_equal(np.array(state.information_state_tensor(p), np.float32),
                                              _dense(st[f"InformationStateTensor({p})"], 2492))
                np.testing.assert_array_equal(np.array(state.observation_tensor(p), np.float32),
                                              _dense(st[f"ObservationTensor({p})"], 98))
            assert game.make_observer(spiel.PUBLIC_OBS_TYPE).string_from(state, 0) == st["PublicObservationString"]
        if st.get("action") is not None:
            state.apply_action(st["action"])
    assert full == 9 and state.is_terminal() and state.returns() == [1.0, -1.0]
    assert state.serialize() == "".join(f"{a}\n" for a in state.history())


def test_random_sims_conformance(game, oracle):
    """The invariants of tests/basic_tests.cc (RandomSimTest) + step-by-step agreement with the oracle."""
    rng = np.random.default_rng(3)
    for ep in range(25):
        state = game.new_initial_state()
        s = oracle.new_state()
        total = np.zeros(2)
        moves = 0
        while not state.is_terminal():
            cur = state.current_player()
            assert cur == oracle.current_player(s)
            la = state.legal_actions()
            assert la == sorted(la) and la == oracle.legal_actions(s)              # basic_tests.cc:740-747
            if cur >= 0:
                assert state.legal_actions(1 - cur) == []                            # basic_tests.cc:83
                assert state.legal_actions(cur) == la
                mask = state.legal_actions_mask()
                assert len(mask) == 18 and [a for a in range(18) if mask[a]] == la    # basic_tests.cc:107
                for p in (0, 1):
                    t = state.information_state_tensor(p)
                    assert len(t) == 2492 and np.isfinite(t).all()
                    np.testing.assert_array_equal(np.array(t, np.float32), oracle.info_state(s, p))
                    np.testing.assert_array_equal(np.array(state.observation_tensor(p), np.float32), oracle.observation(s, p))
                a = int(la[rng.integers(len(la))])
            else:
                oc = state.chance_outcomes()
                assert oc == oracle.chance_outcomes(s)                               # same double arithmetic
                assert abs(sum(p for _, p in oc) - 1) < 1e-12
                assert len(state.legal_actions_mask()) == 5
                a = int(rng.choice([x for x, _ in oc], p=[p for _, p in oc]))
            if moves % 7 == 3:                                                       # clone equality, basic_tests.cc:357-360
                c = state.clone()
                assert str(c) == str(state) and c.history() == state.history()
                assert c.information_state_tensor(0) == state.information_state_tensor(0)
                child = state.child(a)
                assert child.history() == state.history() + [a]
                del c, child
            state.apply_action(a)
            oracle.apply(s, a)
            moves += 1
            if cur >= 0:
                total += state.rewards()
            assert state.returns() == oracle.returns(s) and state.rewards() == oracle.rewards(s)
            assert list(total) == state.returns()                                    # basic_tests.cc:440-455
        assert moves <= 91 and state.legal_actions() == []
        r = state.returns()
        assert -2 <= min(r) and max(r) <= 2 and sum(r) == 0                          # basic_tests.cc:186-210
        # serialize / deserialize round trip (basic_tests.cc:167)
        again = game.deserialize_state(state.serialize())
        assert str(again) == str(state) and again.returns() == r and again.is_terminal()
        assert spiel.serialize_game_and_state(game, state).startswith("# Automatically generated by OpenSpiel SerializeGameAndState\n[Meta]\nVersion: 1\n\n[Game]\ncoup()\n[State]\n")
        with pytest.raises(SpielError):
            state.apply_action(0)
        del again


def test_illegal_moves_raise(game):
    state = game.new_initial_state()
    with pytest.raises(SpielError):
        state.apply_action(7)            # not a card id
    for c in (0, 1, 2, 3):
        state.apply_action(c)
    with pytest.raises(SpielError):
        state.apply_action(2)            # Coup with one coin (coup.cc:549)
    with pytest.raises(SpielError):
        state.information_state_tensor(2)
    with pytest.raises(SpielError):
        state.chance_outcomes()
    assert state.legal_actions() == [0, 1, 3, 5, 6]


def test_rl_environment_mirror(oracle):
    """open_spiel/python/rl_environment.py semantics on the GPU-backed game: FIRST/MID/LAST, rewards None at FIRST,
    discounts zeroed at LAST, legal actions empty for the non-acting player, step after LAST resets
    (rl_environment.py:219-367), and every time step equal to what the oracle produces for the same history."""
    from open_spiel_coup_b200 import rl_environment
    env = rl_environment.Environment("coup", discount=0.9)
    env.seed(7)
    assert env.num_players == 2 and env.is_turn_based and env.name == "coup"
    assert env.observation_spec()["info_state"] == (2492,) and env.action_spec()["num_actions"] == 18
    rng = np.random.default_rng(0)
    for ep in range(6):
        ts = env.reset()
        assert ts.first() and ts.rewards is None and ts.discounts is None
        n_steps = 0
        while not ts.last():
            cur = ts.observations["current_player"]
            hist = env.get_state.history()
            s = oracle.state_from_actions(hist)
            assert cur == oracle.current_player(s)
            assert ts.observations["legal_actions"][cur] == oracle.legal_actions(s)
            assert ts.observations["legal_actions"][1 - cur] == []
            for p in (0, 1):
                np.testing.assert_array_equal(np.array(ts.observations["info_state"][p], np.float32), oracle.info_state(s, p))
            la = ts.observations["legal_actions"][cur]
            ts = env.step([int(la[rng.integers(len(la))])])
            n_steps += 1
            if not ts.last():
                assert ts.mid() and ts.discounts == [0.9, 0.9]
            s2 = oracle.state_from_actions(env.get_state.history())
            assert ts.rewards == oracle.rewards(s2)
        assert ts.discounts == [0.0, 0.0] and ts.observations["current_player"] == -4
        assert ts.observations["legal_actions"] == [[], []]
        again = env.step([0])          # a step after LAST starts a new sequence and ignores the action
        assert again.first()
    obs_env = rl_environment.Environment("coup", observation_type=rl_environment.ObservationType.OBSERVATION,
                                         enable_legality_check=True, include_full_state=True)
    ts = obs_env.reset()
    assert len(ts.observations["info_state"][0]) == 98 and ts.observations["serialized_state"].startswith("# Automatically")
    with pytest.raises(RuntimeError):
        obs_env.step([9])
