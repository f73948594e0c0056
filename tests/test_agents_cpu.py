"""CPU tests of the host-side logic of open_spiel_coup_b200/agents.py on synthetic step batches (no env needed): the
pending-decision bookkeeping of dqn.py:175-248, the learning cadence, the epsilon schedule (dqn.py:296-304), NFSP's
per-episode mode (nfsp.py:146-150) and RollingAverage (rl_response.py:152-171)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from open_spiel_coup_b200 import agents as A  # noqa: E402


def _batch(env_ids, legal_bits, rewards, last, fill):
    k = len(env_ids)
    info = torch.full((k, 2492), fill, dtype=torch.uint8)
    return A.StepBatch(torch.tensor(env_ids), info, torch.tensor(legal_bits, dtype=torch.int32),
                       torch.tensor(rewards, dtype=torch.float32), torch.tensor(last))


def test_dqn_pending_decisions_and_cadence():
    dqn = A.DQN(0, 4, [8], replay_buffer_capacity=64, batch_size=2, min_buffer_size_to_learn=2, learn_every=3,
                update_target_network_every=5, epsilon_start=1.0, epsilon_end=0.0, epsilon_decay_duration=10, device="cpu", seed=1)
    assert dqn._get_epsilon(False) == 1.0 and dqn._get_epsilon(True) == 0.0
    a, p = dqn.step(_batch([0, 2], [0b1001, 0b0110], [0, 0], [False, False], 1))
    assert a[0] in (0, 3) and a[1] in (1, 2) and torch.allclose(p.sum(1), torch.ones(2))
    assert dqn.replay_buffer.total == 0 and dqn._prev_valid.tolist() == [True, False, True, False]
    assert dqn.step_counter == 2 and abs(dqn._get_epsilon(False) - 0.8) < 1e-12
    first_actions = dqn._prev_action[[0, 2]].tolist()
    # env 0 acts again, env 2 ends: two transitions
    dqn.step(_batch([0, 2], [0b11, 0], [1, -2], [False, True], 7))
    rb = dqn.replay_buffer
    assert rb.total == 2 and rb.action[:2].tolist() == first_actions
    assert rb.reward[:2].tolist() == [1, -2] and rb.is_final_step[:2].tolist() == [0, 1]
    assert rb.legal_actions_mask[:2].tolist() == [0b11, 0]
    assert int(rb.info_state[0, 0]) == 1 and int(rb.next_info_state[0, 0]) == 7
    assert dqn._prev_valid.tolist() == [True, False, False, False]
    assert dqn.step_counter == 4 and dqn.loss is None          # crossed 3 once, but the buffer was still empty then
    before = [w.clone() for w in dqn._target_q_network.parameters()]
    dqn.step(_batch([0, 1], [0b1, 0b10], [0, 0], [False, False], 9))      # counter 6: crossed 5 (target copy) and 6 (learn)
    assert dqn.loss is not None and np.isfinite(dqn.loss)
    assert all(torch.equal(t, q) for t, q in zip(dqn._target_q_network.parameters(), dqn._q_network.parameters()))
    assert any(not torch.equal(b, t) for b, t in zip(before, dqn._target_q_network.parameters()))
    # evaluation: greedy, nothing recorded
    total, counter = rb.total, dqn.step_counter
    a, p = dqn.step(_batch([3], [0b100100], [0], [False], 2), is_evaluation=True)
    assert a[0] in (2, 5) and p.max() == 1.0 and (rb.total, dqn.step_counter) == (total, counter)
    with pytest.raises(ValueError):
        A.DQN(0, 1, [8], loss_str="l1", device="cpu")
    with pytest.raises(ValueError):
        A.DQN(0, 1, [8], optimizer_str="rmsprop", device="cpu")


def test_nfsp_modes_and_buffers():
    nf = A.NFSP(1, 2000, [8], reservoir_buffer_capacity=5000, anticipatory_param=0.25, batch_size=4, min_buffer_size_to_learn=4,
                learn_every=1000, device="cpu", seed=3)
    frac = float(nf._best_response.float().mean())
    assert abs(frac - 0.25) < 0.04
    ids = list(range(2000))
    a, p = nf.step(_batch(ids, [0b110] * 2000, [0] * 2000, [False] * 2000, 1))
    assert set(a.tolist()) <= {1, 2} and torch.allclose(p.sum(1), torch.ones(2000))
    n_br = int(nf._best_response.sum())
    assert len(nf.reservoir_buffer) == n_br                       # only best-response decisions are recorded
    assert nf.rl_agent.step_counter == n_br and nf.get_step_counter() == 2000
    assert bool(nf.rl_agent._prev_valid.all())                    # both modes remember the decision for the DQN record
    sl, rl = nf.loss
    assert sl is not None                                         # counter passed 1000 and 2000
    mode_before = nf._best_response.clone()
    nf.step(_batch(ids, [0] * 2000, [1] * 2000, [True] * 2000, 5))
    assert nf.rl_agent.replay_buffer.total == 2000 and not bool(nf.rl_agent._prev_valid.any())
    assert not torch.equal(mode_before, nf._best_response)        # a new mode per env for the next episode
    with nf.temp_mode_as(A.MODE.average_policy):
        a, p = nf.step(_batch([0, 1], [0b1010, 0b1010], [0, 0], [False, False], 1), is_evaluation=True)
        assert (p[:, [1, 3]] > 0).all() and nf.get_step_counter() == 4000


def test_rolling_average_and_crossings():
    r = A.RollingAverage(3)
    assert r.mean() == 0
    for v in (1, 2, 3, 4):
        r.add(v)
    assert r.mean() == 3.0
    assert A._crossings(0, 64, 64) == 1 and A._crossings(63, 64, 64) == 1 and A._crossings(64, 127, 64) == 0
    assert A._crossings(10, 4106, 64) == 64
