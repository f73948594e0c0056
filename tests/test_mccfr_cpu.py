"""CPU tests of the table and regret matching of open_spiel_coup_b200/mccfr.py against a dict-based restatement of
`MCCFRSolverBase` (open_spiel/python/algorithms/mccfr.py:70-131)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from open_spiel_coup_b200.mccfr import AVG_POLICY_INDEX, REGRET_INDEX, InfostateTable, regret_matching  # noqa: E402
from open_spiel_coup_b200.deep_cfr import _legal_bool  # noqa: E402


def test_table_insert_lookup_accumulate():
    rng = np.random.default_rng(0)
    table = InfostateTable("cpu")
    ref = {}
    legal_of = {}
    for rnd in range(6):
        keys = rng.integers(-50, 50, size=200)
        for k in keys:
            legal_of.setdefault(int(k), int(rng.integers(1, 1 << 18)))
        bits = np.array([legal_of[int(k)] for k in keys], np.int32)
        inc = rng.normal(size=(200, 2, 18))
        legal = _legal_bool(torch.as_tensor(bits)).numpy()
        inc *= legal[:, None, :]
        # lookup before the merge sees the old table (initial value for unknown keys)
        got, found = table.lookup(torch.as_tensor(keys), torch.as_tensor(legal))
        for i, k in enumerate(keys):
            exp = ref.get(int(k), np.broadcast_to(legal[i] * 1e-6, (2, 18)))
            np.testing.assert_allclose(got[i].numpy(), exp, rtol=1e-12, atol=1e-18)
            assert bool(found[i]) == (int(k) in ref)
        table.add(torch.as_tensor(keys), torch.as_tensor(bits), torch.as_tensor(inc))
        for i, k in enumerate(keys):
            if int(k) not in ref:
                ref[int(k)] = np.broadcast_to(legal[i] * 1e-6, (2, 18)).copy()
            ref[int(k)] += inc[i]
        assert table.keys.tolist() == sorted(ref)
        for k, v, b in zip(table.keys.tolist(), table.values.numpy(), table.legal_bits.tolist()):
            np.testing.assert_allclose(v, ref[k], rtol=1e-12, atol=1e-15)
            assert b == legal_of[k]
    assert len(table) == len(ref) and REGRET_INDEX == 0 and AVG_POLICY_INDEX == 1


def test_regret_matching_rule():
    legal = _legal_bool(torch.tensor([0b1011, 0b110, 0b1], dtype=torch.int32))
    regrets = torch.zeros((3, 18), dtype=torch.float64)
    regrets[0, [0, 1, 3]] = torch.tensor([2.0, -1.0, 6.0], dtype=torch.float64)
    regrets[1, [1, 2]] = torch.tensor([-3.0, -0.5], dtype=torch.float64)          # nothing positive: uniform
    regrets[2, 0] = 1e-6
    p = regret_matching(regrets, legal).numpy()
    np.testing.assert_allclose(p[0, [0, 1, 3]], [0.25, 0.0, 0.75])
    np.testing.assert_allclose(p[1, [1, 2]], [0.5, 0.5])
    np.testing.assert_allclose(p[2, 0], 1.0)
    assert (p[~legal.numpy()] == 0).all()
