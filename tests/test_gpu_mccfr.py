"""GPU tests (-m gpu) of the batched outcome-sampling MCCFR (open_spiel_coup_b200/mccfr.py) against a literal
restatement of the reference recursion (open_spiel/python/algorithms/outcome_sampling_mccfr.py:62-145 with the table
of mccfr.py:70-131) replayed on the CPU oracle over the very episodes the device sampled. float64 on both sides;
tolerance 1e-9 relative (the only difference is the order in which increments of one batch are summed)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

from open_spiel_coup_b200 import mccfr  # noqa: E402


def _regret_matching(regrets):
    pos = np.maximum(regrets, 0.0)
    s = pos.sum()
    return np.ones(len(regrets)) / len(regrets) if s <= 0 else pos / s


class _RefTable:
    def __init__(self):
        self.t = {}

    def get(self, key, n):
        return self.t.get(key, [np.ones(n) / 1e6, np.ones(n) / 1e6])


def _replay_batch(oracle, table, update_player, histories, expl=0.6):
    """All episodes of one batch against the SAME table snapshot; returns {key: (legal, d_regret, d_avstrat)}."""
    delta = {}
    roots = []
    for actions, _ in histories:
        actions = [int(a) for a in actions]
        s = oracle.new_state()
        my_reach = opp_reach = sample_reach = 1.0
        i = 0
        frames = []
        while not oracle.is_terminal(s):
            if oracle.current_player(s) == -1:
                probs = dict(oracle.chance_outcomes(s))
                p = probs[actions[i]]
                oracle.apply(s, actions[i]); i += 1
                opp_reach *= p; sample_reach *= p                                    # :73-77
                continue
            cur = oracle.current_player(s)
            key = oracle.tensor_hash(oracle.info_state(s, cur))
            key = key - (1 << 64) if key >= (1 << 63) else key                       # as a signed int64, like the device table
            legal = oracle.legal_actions(s)
            entry = table.get(key, len(legal))
            policy = _regret_matching(entry[0])
            sample_policy = expl * np.ones(len(legal)) / len(legal) + (1 - expl) * policy if cur == update_player else policy
            aidx = legal.index(actions[i])
            frames.append((key, legal, cur, policy, sample_policy, aidx, my_reach, opp_reach, sample_reach))
            if cur == update_player:
                my_reach *= policy[aidx]
            else:
                opp_reach *= policy[aidx]
            sample_reach *= sample_policy[aidx]
            oracle.apply(s, actions[i]); i += 1
        assert i == len(actions)
        value = oracle.returns(s)[update_player]
        for key, legal, cur, policy, sample_policy, aidx, my_r, opp_r, smp_r in reversed(frames):
            child_values = np.zeros(len(legal))
            child_values[aidx] = value / sample_policy[aidx]                          # :50-60 with baseline 0
            value_estimate = float((policy * child_values).sum())
            d = delta.setdefault(key, (legal, np.zeros(len(legal)), np.zeros(len(legal))))
            assert d[0] == legal
            if cur == update_player:
                cf_value = value_estimate * opp_r / smp_r
                d[1][:] += child_values * opp_r / smp_r - cf_value                    # :127-137
                d[2][:] += my_r * policy / smp_r                                      # :139-142
            value = value_estimate
        roots.append((frames, value))
    return delta, roots


def test_batched_outcome_sampling_equals_reference_recursion(oracle):
    solver = mccfr.OutcomeSamplingSolver(num_envs=48, seed=5, record=True)
    ref = _RefTable()
    for it in range(4):
        solver.iteration()
        assert [b["update_player"] for b in solver.last_episodes] == [0, 1]
        for batch in solver.last_episodes:
            delta, roots = _replay_batch(oracle, ref, batch["update_player"], batch["histories"])
            # what the device used at every step equals what the restatement derives from the same table snapshot
            first = batch["steps"][0]
            for e, (frames, value) in enumerate(roots):
                key, legal, cur, policy, sample_policy, aidx, my_r, opp_r, smp_r = frames[0]
                assert int(first["keys"][e]) == key
                np.testing.assert_allclose(first["policy"][e].cpu().numpy()[legal], policy, rtol=1e-12, atol=0)
                np.testing.assert_allclose(first["sample_policy"][e].cpu().numpy()[legal], sample_policy, rtol=1e-12, atol=0)
                np.testing.assert_allclose(float(first["opp_reach"][e]), opp_r, rtol=1e-12)   # the four initial deals
                np.testing.assert_allclose(float(batch["root_value"][e]), value, rtol=1e-9, atol=1e-12)
            for key, (legal, d_reg, d_av) in delta.items():
                entry = ref.t.setdefault(key, [np.ones(len(legal)) / 1e6, np.ones(len(legal)) / 1e6, legal])
                entry[0] = entry[0] + d_reg
                entry[1] = entry[1] + d_av
        table = solver.infostates
        assert len(table) == len(ref.t)
        keys = table.keys.cpu().tolist()
        assert keys == sorted(ref.t)
        vals = table.values.cpu().numpy()
        bits = table.legal_bits.cpu().tolist()
        for k, v, b in zip(keys, vals, bits):
            reg, av, legal = ref.t[k]
            assert [a for a in range(18) if (b >> a) & 1] == legal
            np.testing.assert_allclose(v[0][legal], reg, rtol=1e-9, atol=1e-15)
            np.testing.assert_allclose(v[1][legal], av, rtol=1e-9, atol=1e-15)
            illegal = [a for a in range(18) if a not in legal]
            assert (v[:, illegal] == 0).all()
    assert len(ref.t) > 200
    # average policy: normalised cumulative strategy where visited, uniform elsewhere (mccfr.py:33-60)
    pol = solver.average_policy()
    from open_spiel_coup_b200.spiel import load_game
    state = load_game("coup").new_initial_state()
    for c in (0, 1, 2, 3):
        state.apply_action(c)
    probs = pol.action_probabilities(state)
    key = oracle.tensor_hash(np.array(state.information_state_tensor(0), np.float32))
    key = key - (1 << 64) if key >= (1 << 63) else key
    legal = state.legal_actions()
    if key in ref.t:
        np.testing.assert_allclose([probs[a] for a in legal], ref.t[key][1] / ref.t[key][1].sum(), rtol=1e-6)
    else:
        np.testing.assert_allclose([probs[a] for a in legal], 1 / len(legal), rtol=1e-6)
    unseen = pol.action_probs(torch.zeros((2, 2492), dtype=torch.uint8, device="cuda"), torch.tensor([0b111, 0b1001], dtype=torch.int32, device="cuda"))
    np.testing.assert_allclose(unseen.cpu().numpy()[0, :3], 1 / 3, rtol=1e-6)
    np.testing.assert_allclose(unseen.cpu().numpy()[1, [0, 3]], 0.5, rtol=1e-6)


def test_mccfr_average_policy_as_rl_resp_exploitee():
    """The loop of coup_experiments/scripts/mccfr.py:46-60 at a small scale: iterate, then measure the average policy
    with rl_resp."""
    from open_spiel_coup_b200 import agents
    solver = mccfr.OutcomeSamplingSolver(num_envs=2048, seed=1)
    for _ in range(5):
        solver.iteration()
    assert len(solver.infostates) > 5000
    recs = agents.rl_resp(exploitee=solver.average_policy(), num_train_episodes=512, eval_every=512, eval_episodes=256, num_envs=256)
    assert len(recs) == 1 and -4 <= recs[0]["value"] <= 4
