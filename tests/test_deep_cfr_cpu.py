"""CPU tests of the host-side pieces of open_spiel_coup_b200/deep_cfr.py: regret matching against a literal
restatement of `_sample_action_from_advantage` (open_spiel/python/algorithms/deep_cfr.py:499-525), the batched
reservoir buffer against the sequential rule (deep_cfr.py:43-99), and the network initialiser (simple_nets.py:44-52).
The traversal itself needs the CUDA library and is tested under -m gpu (tests/test_gpu_deep_cfr.py)."""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from open_spiel_coup_b200.deep_cfr import MLP, ReservoirBuffer, _legal_bool, regret_matching  # noqa: E402


def _reference_matching(advantages_full, legal_actions, num_actions=18):
    advantages = [max(0., a) for a in advantages_full]
    cumulative_regret = np.sum([advantages[a] for a in legal_actions])
    matched = np.array([0.] * num_actions)
    if cumulative_regret > 0.:
        for a in legal_actions:
            matched[a] = advantages[a] / cumulative_regret
    else:
        matched[max(legal_actions, key=lambda a: advantages_full[a])] = 1
    return matched


def test_regret_matching_equals_reference_rule():
    rng = np.random.default_rng(0)
    adv = rng.normal(size=(500, 18)).astype(np.float32)
    adv[:100] = -np.abs(adv[:100])                      # nothing positive: fallback branch
    adv[100:120, :] = -1.0                              # ties: the first legal action wins
    masks = rng.integers(1, 1 << 18, size=500).astype(np.int32)
    legal = _legal_bool(torch.as_tensor(masks))
    got = regret_matching(torch.as_tensor(adv), legal).numpy()
    for i in range(500):
        la = [a for a in range(18) if (masks[i] >> a) & 1]
        np.testing.assert_allclose(got[i], _reference_matching(adv[i].tolist(), la), rtol=1e-6, atol=1e-7)
    assert np.allclose(got.sum(1), 1, atol=1e-6) and (got[~legal.numpy()] == 0).all()


def test_reservoir_fill_then_uniform_replacement():
    fields = {"x": ((3,), torch.float32), "iteration": ((), torch.int32)}
    buf = ReservoirBuffer(100, "cpu", fields, seed=1)
    buf.add(x=torch.arange(60).float().view(-1, 1).expand(-1, 3), iteration=torch.ones(60))
    assert len(buf) == 60 and buf.all()["x"][:, 0].tolist() == list(range(60))
    buf.add(x=torch.arange(60, 100).float().view(-1, 1).expand(-1, 3), iteration=torch.ones(40))
    assert len(buf) == 100 and buf.all()["x"][:, 1].tolist() == list(range(100))
    with pytest.raises(ValueError):
        ReservoirBuffer(10, "cpu", fields).sample(1)
    # after 100 000 more elements every element ever added is kept with probability capacity / total
    hits = np.zeros(4, np.int64)
    for rep in range(30):
        buf = ReservoirBuffer(100, "cpu", fields, seed=rep)
        for start in range(0, 4000, 500):
            v = torch.arange(start, start + 500).float().view(-1, 1).expand(-1, 3)
            buf.add(x=v, iteration=torch.full((500,), 2))
        kept = buf.all()["x"][:, 0].long().numpy()
        assert len(np.unique(kept)) == 100
        hits += np.bincount(kept // 1000, minlength=4)
    assert abs(hits / hits.sum() - 0.25).max() < 0.04
    s = buf.sample(50)
    assert s["x"].shape == (50, 3) and len(np.unique(s["x"][:, 0].numpy())) == 50
    buf.clear()
    assert len(buf) == 0


def test_mlp_initialiser_and_shape():
    torch.manual_seed(0)
    net = MLP(2492, [64, 32], 18)
    first = net.net[0]
    std = 1 / math.sqrt(2492)
    assert float(first.weight.abs().max()) <= 2 * std + 1e-7 and float(first.bias.abs().max()) == 0
    assert abs(float(first.weight.std()) - 0.88 * std) < 0.05 * std          # std of a normal truncated at 2 sigma
    assert net(torch.zeros(5, 2492)).shape == (5, 18)
    before = first.weight.clone()
    net.reset_parameters()
    assert not torch.equal(before, first.weight)


def test_backward_sweep_on_a_hand_made_tree():
    """deep_cfr.py:468-480 on a three-level tree: traverser root with actions {0, 1, 3} all expanded (external
    sampling), child 0 terminal (+1), child 1 an opponent node whose sampled action leads to a terminal (-2), child 3 a
    traverser node with legal {2, 5} of which only 5 was sampled (outcome sampling), leading to a terminal (+2)."""
    from open_spiel_coup_b200.deep_cfr import backward_sweep
    T, F = True, False

    def legal(*acts):
        row = torch.zeros(18, dtype=torch.bool)
        row[list(acts)] = True
        return row

    def strat(d):
        row = torch.zeros(18)
        for a, p in d.items():
            row[a] = p
        return row

    l0 = {"m": 1, "terminal": torch.tensor([F]), "ret_p": torch.zeros(1, dtype=torch.float64), "nt": torch.tensor([0]), "children": 3,
          "strategy": strat({0: 0.5, 1: 0.25, 3: 0.25}).view(1, -1), "legal": legal(0, 1, 3).view(1, -1), "is_trav": torch.tensor([T]),
          "trav": torch.tensor([0]), "local": torch.tensor([0, 0, 0]), "action": torch.tensor([0, 1, 3])}
    l1 = {"m": 3, "terminal": torch.tensor([T, F, F]), "ret_p": torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64), "nt": torch.tensor([1, 2]),
          "children": 2, "strategy": torch.stack([strat({9: 0.7, 10: 0.3}), strat({2: 0.4, 5: 0.6})]),
          "legal": torch.stack([legal(9, 10), legal(2, 5)]), "is_trav": torch.tensor([F, T]), "trav": torch.tensor([1]),
          "local": torch.tensor([0, 1]), "action": torch.tensor([10, 5])}
    l2 = {"m": 2, "terminal": torch.tensor([T, T]), "ret_p": torch.tensor([-2.0, 2.0], dtype=torch.float64), "nt": torch.zeros(0, dtype=torch.long),
          "children": 0}
    out = list(backward_sweep([l0, l1, l2]))
    assert [id(lvl) for lvl, _ in out] == [id(l1), id(l0)]
    # level 1: opponent passes its sampled child's value through; traverser node: cfv = 0.6 * 2, unsampled action 2 counts as 0
    np.testing.assert_allclose(l1["value"].numpy(), [1.0, -2.0, 1.2])
    r1 = out[0][1].numpy()[0]
    np.testing.assert_allclose(r1[[2, 5]], [0.0 - 1.2, 2.0 - 1.2], rtol=1e-6)
    assert (np.delete(r1, [2, 5]) == 0).all()
    # root: cfv = 0.5 * 1 + 0.25 * (-2) + 0.25 * 1.2
    cfv = 0.5 - 0.5 + 0.3
    np.testing.assert_allclose(l0["value"].numpy(), [cfv])
    r0 = out[1][1].numpy()[0]
    np.testing.assert_allclose(r0[[0, 1, 3]], [1 - cfv, -2 - cfv, 1.2 - cfv], rtol=1e-6)
    np.testing.assert_allclose(l2["value"].numpy(), [-2.0, 2.0])
