"""GPU test (-m gpu): the bit-exactness replay of BASELINE config 5 at a size that runs in seconds.
The full 10^6-trajectory run is scripts/replay_check.py (result committed under profiles/)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import replay_check  # noqa: E402


@pytest.mark.parametrize("use_reference", [False, True], ids=["oracle-c-port", "reference-build"])
def test_replay_digests_match(use_reference):
    from oracle.bindings import Reference
    if use_reference and not Reference.available():
        pytest.skip("oracle/_ref not built")
    res = replay_check.check(envs=30000, steered=1024, seed=99, device=0, threads=os.cpu_count() or 1,
                             use_reference=use_reference)
    assert res["total_mismatches"] == 0, res
    uni, steer = res["batches"]
    assert uni["trajectories"] == 30000 and uni["episodes_stat"] == 30000
    assert uni["reported_states"] > 30000 * 14
    assert uni["games_with_chance_after_exchange_return"] > 1000      # the deck quirk is exercised
    assert steer["truncated_91_move_games"] >= 10                      # scripted 91-move truncation games
    assert ("reference" in res["checker"]) == use_reference
