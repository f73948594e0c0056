"""GPU tests (-m gpu): the bit-exactness replay of BASELINE configs[4] at FULL size -- 2^20 uniform-random + 4 096
steered trajectories without auto-reset, and 2^20 finished episodes harvested from the ring of an AUTO-RESET fused
rollout (the mode bench.py times) -- replayed through the compiled unmodified reference (oracle/_ref). A small
configuration runs against the C port as well."""
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

THREADS = os.cpu_count() or 1


def _have_reference():
    from oracle.bindings import Reference
    return Reference.available()


def test_replay_digests_match_c_port():
    res = replay_check.check(envs=30000, steered=1024, seed=99, device=0, threads=THREADS, use_reference=False)
    assert res["total_mismatches"] == 0, res
    uni, steer = res["batches"]
    assert uni["trajectories"] == 30000 and uni["episodes_stat"] == 30000
    assert uni["reported_states"] > 30000 * 14
    assert uni["games_with_chance_after_exchange_return"] > 1000      # the deck quirk is exercised
    assert steer["truncated_91_move_games"] >= 10                      # scripted 91-move truncation games
    assert "reference" not in res["checker"]


def test_replay_full_size_against_reference():
    """BASELINE configs[4]: 10^6 trajectories -> 2^20 + 4096 steered, checker = the compiled reference."""
    if not _have_reference():
        pytest.skip("oracle/_ref not built")
    res = replay_check.check(envs=1 << 20, steered=4096, seed=1234, device=0, threads=THREADS, use_reference=True)
    assert res["total_mismatches"] == 0, res
    uni, steer = res["batches"]
    assert "reference" in res["checker"]
    assert uni["trajectories"] == 1 << 20 and uni["episodes_stat"] == 1 << 20
    assert uni["games_with_chance_after_exchange_return"] > 30000
    assert steer["truncated_91_move_games"] >= 40


@pytest.mark.parametrize("envs,episodes,use_reference", [(4096, 20000, False), (1 << 17, 1 << 20, True)],
                         ids=["small-c-port", "2^20-episodes-reference"])
def test_autoreset_ring_replay(envs, episodes, use_reference):
    """>= `episodes` finished episodes from an auto-reset coup_vec_rollout run, through the ring, 0 mismatches."""
    if use_reference and not _have_reference():
        pytest.skip("oracle/_ref not built")
    res = replay_check.check_autoreset(envs, episodes, seed=4321, device=0, threads=THREADS, use_reference=use_reference)
    assert res["total_mismatches"] == 0, res
    assert res["episodes"] >= episodes and res["reported_states"] > 14 * episodes
    assert ("reference" in res["checker"]) == use_reference


def test_sharding_invariance_full_size():
    """2^20 envs as one slab == the same global env ids as 8 slabs (global-id Philox keys), state + history words."""
    from open_spiel_coup_b200.vector_env import CoupVectorEnv
    n, parts = 1 << 20, 8
    whole = CoupVectorEnv(n, seed=77, device=0, auto_reset=True)
    whole.rollout(40)
    for r in range(parts):
        lo = r * (n // parts)
        part = CoupVectorEnv(n // parts, seed=77, device=0, global_env_offset=lo, auto_reset=True)
        part.rollout(40)
        assert bool((part.state == whole.state[lo:lo + n // parts]).all())
        assert bool((part.history == whole.history[lo:lo + n // parts]).all())
        assert bool((part.step_word == whole.step_word[lo:lo + n // parts]).all())
        part.close()
    whole.close()
