"""GPU tests (-m gpu): the general observer on the device (CoupGame::MakeObserver for any IIGObservationType,
coup.cc:1132-1141) against the oracle's restatement of CoupObserver::WriteTensor (coup.cc:248-287), all 12
(public_info, perfect_recall, private_info) combinations, both seats, all element types."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

from open_spiel_coup_b200 import _lib  # noqa: E402
from open_spiel_coup_b200.vector_env import CoupVectorEnv  # noqa: E402


@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8, torch.bfloat16], ids=["f32", "u8", "bf16"])
def test_general_observer_matches_oracle(oracle, dtype):
    n = 300                                        # ragged: 9 full warps + 12 envs
    env = CoupVectorEnv(n, seed=17, auto_reset=False)
    checked = 0
    for steps in (0, 7, 30):                       # fresh deals, mid-game (hands of 1-4 cards, lost cards), many terminal
        env.rollout(steps)
        states = [oracle.state_from_actions(a) for a, _ in env.trajectories()]
        for public, recall, private in itertools.product((0, 1), (0, 1), (0, 1, 2)):
            width = (2492 if recall else 98) if public else 42
            both = env.observer_tensor(_lib.PLAYER_BOTH, public, recall, private, dtype=dtype).float().cpu().numpy()
            assert both.shape == (2 * n, width)
            for e in range(0, n, 7):
                for p in (0, 1):
                    want = oracle.observer_tensor(states[e], p, public, recall, private)
                    assert want.shape == (width,)
                    assert (both[2 * e + p] == want).all(), (steps, public, recall, private, e, p)
                    checked += 1
            one = env.observer_tensor(_lib.PLAYER_1, public, recall, private, dtype=dtype).float().cpu().numpy()
            assert (one == both[1::2]).all()
    assert checked > 3000
    # the two built-in tensors are the (1,1,1) and (1,0,1) instances
    assert torch.equal(env.observer_tensor(_lib.PLAYER_BOTH, 1, 1, 1, dtype=dtype), env.information_state_tensor(_lib.PLAYER_BOTH, dtype=dtype))
    assert torch.equal(env.observer_tensor(_lib.PLAYER_BOTH, 1, 0, 1, dtype=dtype), env.observation_tensor(_lib.PLAYER_BOTH, dtype=dtype))
