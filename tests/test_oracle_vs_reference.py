"""CPU tests: differential test of the C restatement against the UNMODIFIED reference compiled into
oracle/_ref (skipped where that build is absent). This is what pins the oracle beyond the reference's
own golden vectors -- in particular the ExchangeReturn deck quirk, which no reference test covers."""
import numpy as np
import pytest


def _random_trajectories(ref, n, seed, steer_every=25):
    rng = np.random.default_rng(seed)
    trajs = []
    for ep in range(n):
        h = ref.new_state()
        acts = []
        steer = steer_every and ep % steer_every == 0
        while not ref.is_terminal(h):
            if ref.is_chance(h):
                oc = ref.chance_outcomes(h)
                a = int(rng.choice([x for x, _ in oc], p=[p for _, p in oc]))
            else:
                la = ref.legal_actions(h)
                a = int(la[rng.integers(len(la))])
                if steer:
                    for x in la:
                        if x in (5, 9, 17, 16, 15):
                            a = x
                            if x in (9, 5):
                                break
            ref.apply(h, a)
            acts.append(a)
        ref.free(h)
        trajs.append(np.array(acts, np.uint8))
    return trajs


def test_reference_own_constants(reference):
    c = reference.game_constants()
    assert c == {"NumDistinctActions": 18, "MaxChanceOutcomes": 5, "NumPlayers": 2, "MinUtility": -2,
                 "MaxUtility": 2, "InformationStateTensorSize": 2492, "ObservationTensorSize": 98,
                 "MaxGameLength": 90, "MaxChanceNodesInHistory": 45, "MaxMoveNumber": 135, "UtilitySum": 0}


def test_trace_records_identical(oracle, reference):
    trajs = _random_trajectories(reference, 1500, seed=123)
    flat = np.concatenate(trajs)
    off = np.concatenate([[0], np.cumsum([len(t) for t in trajs])]).astype(np.int64)
    a, bad_a = oracle.trace_batch(flat, off)
    b, bad_b = reference.trace_batch(flat, off, 8)
    assert bad_a == 0 and bad_b == 0
    assert a.tobytes() == b.tobytes()
    assert sum(len(t) > 90 for t in trajs) >= 30


def test_step_by_step_everything(oracle, reference):
    """Walks games move by move comparing every observable, incl. chance probabilities and the
    general observer tensors (coup.cc:1132-1141) for all IIGObservationType combinations."""
    rng = np.random.default_rng(5)
    combos = [(pub, rec, priv) for pub in (0, 1) for rec in (0, 1) for priv in (0, 1, 2)]
    for ep in range(60):
        h = reference.new_state()
        s = oracle.new_state()
        while True:
            assert oracle.is_terminal(s) == reference.is_terminal(h)
            assert oracle.current_player(s) == reference.current_player(h)
            assert oracle.returns(s) == reference.returns(h)
            assert oracle.rewards(s) == reference.rewards(h)
            for p in (0, 1):
                assert [s.players[p].cards[i].value for i in range(s.players[p].num_cards)] == reference.cards(h, p)[0]
                assert [s.players[p].cards[i].state for i in range(s.players[p].num_cards)] == reference.cards(h, p)[1]
                assert s.players[p].coins == reference.coins(h, p)
                assert s.players[p].last_action == reference.last_action(h, p)
                for pub, rec, priv in combos:
                    np.testing.assert_array_equal(oracle.observer_tensor(s, p, pub, rec, priv),
                                                  reference.observer_tensor(h, p, pub, rec, priv))
            if reference.is_terminal(h):
                break
            assert oracle.legal_actions(s) == reference.legal_actions(h)
            if reference.is_chance(h):
                oc = reference.chance_outcomes(h)
                got = oracle.chance_outcomes(s)
                assert [a for a, _ in got] == [a for a, _ in oc]
                assert [p for _, p in got] == [p for _, p in oc]      # same double arithmetic
                a = int(rng.choice([x for x, _ in oc], p=[p for _, p in oc]))
            else:
                la = reference.legal_actions(h)
                a = int(la[rng.integers(len(la))])
                if ep % 3 == 0:
                    for x in la:
                        if x in (5, 9, 17, 16, 15):
                            a = x
                            if x in (9, 5):
                                break
            assert reference.apply(h, a) == 0
            assert oracle.apply(s, a) == 0
        reference.free(h)


def test_reference_rejects_what_oracle_rejects(oracle, reference):
    # Coup with too few coins is one of the reference's own SPIEL_CHECKs (coup.cc:549).
    h = reference.state_from_actions([0, 1, 2, 3])
    assert reference.apply(h, 2) != 0
    s = oracle.state_from_actions([0, 1, 2, 3])
    assert oracle.apply(s, 2, checked=False) != 0
    reference.free(h)


def test_cpu_baseline_harness_statistics(oracle, reference):
    """The two CPU baselines follow the same protocol; their workload statistics must agree with each
    other and with SURVEY.md section 6 (21.2 moves/episode, 15.0 decisions, mean legal 3.59)."""
    r = reference.bench(0, 2, 20000, 1)
    o = oracle.bench(0, 2, 20000, 1)
    for b in (r, o):
        assert 20.6 < b["moves"] / b["episodes"] < 21.8
        assert 14.6 < b["decisions"] / b["episodes"] < 15.5
        tot = sum(b["legal_count_hist"])
        mean_legal = sum(i * c for i, c in enumerate(b["legal_count_hist"])) / tot
        assert 3.5 < mean_legal < 3.7
        assert b["returns_hist_p0"][2] == 0
        assert b["max_coins"] <= 12
