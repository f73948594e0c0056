"""GPU tests (-m gpu) of the finished-episode ring and the SyncVectorEnv-shaped adapter: in auto-reset mode the step
kernels re-deal finished envs in place, and the ring is what keeps their action + chance logs, terminal states and
terminal observations (python/vector_env.py:52-66 `unreset_time_steps`; coup_experiments/scripts/nfsp.py:141-143)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

from open_spiel_coup_b200 import _lib  # noqa: E402
from open_spiel_coup_b200.vector_env import (CoupVectorEnv, SyncVectorEnv, decode_finished_records)  # noqa: E402


def _flat(trajs):
    flat = np.concatenate([a for a, _ in trajs]) if trajs else np.zeros(0, np.uint8)
    off = np.concatenate([[0], np.cumsum([len(a) for a, _ in trajs])]).astype(np.int64)
    return flat, off


@pytest.mark.parametrize("path", ["fused_ws_f32", "fused_cta_u8", "step_kernel", "env_only_multi", "incremental"])
def test_ring_records_replay_through_oracle(oracle, path):
    n = 4096 if path != "fused_ws_f32" else 8192
    env = CoupVectorEnv(n, seed=2024, auto_reset=True, finished_ring=1 << 16, warp_specialised=(path != "fused_cta_u8"))
    buf = None
    if path == "fused_ws_f32":
        buf = torch.empty((n, 2492), dtype=torch.float32, device=env.device)
    elif path == "fused_cta_u8":
        buf = torch.empty((n, 2492), dtype=torch.uint8, device=env.device)
    elif path == "incremental":
        buf = env.information_state_tensor(_lib.PLAYER_BOTH)
    seen = 0
    per_step_terminal = []
    for step in range(70):
        if path == "step_kernel":
            env.step(env.sample_uniform())
        elif path == "env_only_multi":
            env.rollout(3)
        elif path == "incremental":
            env.rollout_incremental(1, buf)
        else:
            env.rollout(1, _lib.PLAYER_CURRENT, out=buf)
        if path != "env_only_multi":
            rows, ids, cnt = env.finished_information_state_tensor(_lib.PLAYER_BOTH, max_episodes=n)
            orow, ids2, cnt2 = env.finished_observation_tensor(_lib.PLAYER_BOTH, max_episodes=n)
            k = int(cnt.item())
            assert k == int(cnt2.item()) == int(env.done.sum().item())
            assert bool((ids[:k] == ids2[:k]).all())
            assert sorted(ids[:k].tolist()) == torch.nonzero(env.done).view(-1).tolist()
            per_step_terminal.append((rows[:2 * k].cpu().numpy(), orow[:2 * k].cpu().numpy(), ids[:k].cpu().numpy()))
            seen += k
    recs, dropped = env.finished_drain()
    stats = env.stats()
    assert dropped == 0 and len(recs) == stats["episodes"] > n * 3
    d = decode_finished_records(recs)
    flat, off = _flat(d["trajectories"])
    rec, bad = oracle.final_batch(flat, off)
    assert bad == 0
    assert (rec["cur_player"] == _lib.TERMINAL_PLAYER_ID).all()            # every logged episode ends terminal
    assert (rec["returns"][:, 0] == d["return0"]).all() and (rec["rewards"][:, 0] == d["reward0"]).all()
    assert (d["moves"] == np.diff(off)).all()
    assert (d["truncated"] == (d["moves"] > 90)).all()
    assert int(d["truncated"].sum()) == stats["truncated"]
    hist = np.bincount(d["return0"] + 2, minlength=5)
    assert hist.tolist() == stats["returns_hist_p0"]
    if per_step_terminal:
        assert seen == len(recs)
        ti = np.concatenate([t[0] for t in per_step_terminal]).reshape(-1, 2, 2492)
        to = np.concatenate([t[1] for t in per_step_terminal]).reshape(-1, 2, 98)
        assert (np.concatenate([t[2] for t in per_step_terminal]) == d["env"]).all()
        m = min(len(recs), 600)                                             # dense compare on a prefix, hashes on all
        oi, oo, _ = oracle.final_tensors_batch(flat[: off[m]], off[: m + 1])
        assert (ti[:m] == oi).all() and (to[:m] == oo).all()
        hi = env.tensor_row_hash(torch.as_tensor(ti.reshape(-1, 2492)).to(env.device)).cpu().numpy().view(np.uint64).reshape(-1, 2)
        ho = env.tensor_row_hash(torch.as_tensor(to.reshape(-1, 98)).to(env.device)).cpu().numpy().view(np.uint64).reshape(-1, 2)
        assert (hi == rec["hash_info"]).all() and (ho == rec["hash_obs"]).all()
    # decoding the drained records as a plain record array gives the same rows
    dev_recs = torch.as_tensor(recs.view(np.int32)).to(env.device)
    idx = torch.arange(0, min(len(recs), 1000), 3, device=env.device)
    dec = env.records_information_state_tensor(dev_recs, idx, _lib.PLAYER_BOTH, dtype=torch.uint8).view(-1, 2, 2492).cpu().numpy()
    oi, _, _ = oracle.final_tensors_batch(flat[: off[1000 if len(recs) >= 1000 else len(recs)]], off[: min(len(recs), 1000) + 1], obs=False)
    assert (dec == oi[idx.cpu().numpy()]).all()


def test_ring_without_auto_reset_and_overflow(oracle):
    n = 2048
    env = CoupVectorEnv(n, seed=5, auto_reset=False, finished_ring=256)     # far too small on purpose
    env.rollout(100)
    assert bool(env.done.all())
    recs, dropped = env.finished_drain()
    assert len(recs) == 256 and dropped == n - 256                          # the newest 256 survive, the rest are counted
    d = decode_finished_records(recs)
    flat, off = _flat(d["trajectories"])
    rec, bad = oracle.final_batch(flat, off)
    assert bad == 0 and (rec["cur_player"] == _lib.TERMINAL_PLAYER_ID).all()
    # each surviving record equals the env's own (never re-dealt) final trajectory
    trajs = env.trajectories()
    for i, e in enumerate(d["env"]):
        assert (trajs[e][0] == d["trajectories"][i][0]).all()
    again, dropped2 = env.finished_drain()
    assert len(again) == 0 and dropped2 == dropped
    env.enable_finished_ring(0)
    env.reset()
    env.rollout(5)                                                           # no ring: must simply work


def test_misaligned_outputs_take_the_plain_path_or_are_rejected():
    n = 512
    env = CoupVectorEnv(n, seed=3, auto_reset=True)
    env.rollout(10)
    ref = env.information_state_tensor(_lib.PLAYER_CURRENT, dtype=torch.uint8)
    big = torch.empty((n + 1, 2492), dtype=torch.uint8, device=env.device)
    view = big[1:]                                                           # 2492-byte offset: 4-byte but not 16-byte aligned
    assert view.data_ptr() % 16 != 0 and view.data_ptr() % 4 == 0
    env.information_state_tensor(_lib.PLAYER_CURRENT, out=view)
    assert bool((view == ref).all())
    flat = torch.empty(n * 2492 + 8, dtype=torch.uint8, device=env.device)
    bad = flat[1:1 + n * 2492].view(n, 2492)
    with pytest.raises(ValueError):
        env.information_state_tensor(_lib.PLAYER_CURRENT, out=bad)
    obs_ref = env.observation_tensor(_lib.PLAYER_BOTH, dtype=torch.uint8)
    oflat = torch.empty(2 * n * 98 + 8, dtype=torch.uint8, device=env.device)
    oview = oflat[3:3 + 2 * n * 98].view(2 * n, 98)
    env.observation_tensor(_lib.PLAYER_BOTH, out=oview)
    assert bool((oview == obs_ref).all())
    torch.cuda.synchronize()


def test_step_at_explicit_chance_node_is_refused():
    env = CoupVectorEnv(64, seed=1)
    env._lib.coup_vec_new_initial_state(env._h, None, None)                  # every env at its first chance node
    env.step(torch.zeros(64, dtype=torch.uint8, device=env.device))
    assert env.stats()["illegal"] == 64
    assert bool((env.current_player == _lib.CHANCE_PLAYER_ID).all())
    with pytest.raises(_lib.CoupError):
        env.check_errors()


def test_sync_vector_env_adapter(oracle):
    """vector_env.SyncVectorEnv call shape; every time step checked against the oracle replay of the env's own log."""
    n = 256
    venv = SyncVectorEnv(n, seed=11)
    ts = venv.reset()
    assert bool(ts.first().all()) and len(ts) == n
    logs = [[] for _ in range(n)]           # (action lists) of the episode each env is in
    env = venv.env

    def check_live(ts, which):
        trajs = env.trajectories()
        sel = [i for i in range(n) if which[i]]
        flat, off = _flat([trajs[i] for i in sel])
        rec, bad = oracle.final_batch(flat, off)
        assert bad == 0
        ti, _, _ = oracle.final_tensors_batch(flat, off, obs=False)
        got = ts.observations["info_state"].cpu().numpy()[sel]
        assert (got == ti).all()
        cur = ts.observations["current_player"].cpu().numpy()[sel]
        assert (cur == rec["cur_player"]).all()
        mask = ts.observations["legal_actions_mask"].cpu().numpy()[sel]
        bits = (rec["legal_mask"][:, None] >> np.arange(18)) & 1
        for j in range(len(sel)):
            assert (mask[j, cur[j]] == bits[j]).all() and not mask[j, 1 - cur[j]].any()

    check_live(ts, np.ones(n, bool))
    finished_seen = 0
    for it in range(60):
        reset_if_done = it % 2 == 0
        actions = env.sample_uniform()
        ts, reward, done, unreset = venv.step(actions, reset_if_done=reset_if_done)
        done_np = done.cpu().numpy().astype(bool)
        finished_seen += int(done_np.sum())
        assert bool((unreset.step_type[done] == 2).all()) and bool((unreset.observations["current_player"][done] == -4).all())
        assert bool((unreset.observations["legal_actions_mask"][done] == 0).all()) and bool((unreset.discounts[done] == 0).all())
        if reset_if_done:
            assert bool((ts.step_type[done] == 0).all()) and bool((ts.rewards[done] == 0).all())
            check_live(ts, np.ones(n, bool))
        else:
            assert ts is unreset
            check_live(ts, ~done_np)
        # terminal time steps: replay the ring's log of exactly those episodes
        recs = env.finished_ring.cpu().numpy().view(np.uint32)
        ctrl = env.finished_ctrl.cpu().numpy()
        seg = recs[[(p & (recs.shape[0] - 1)) for p in range(int(ctrl[1]), int(ctrl[0]))]]
        d = decode_finished_records(seg)
        assert sorted(d["env"].tolist()) == np.nonzero(done_np)[0].tolist()
        if len(seg):
            flat, off = _flat(d["trajectories"])
            ti, _, _ = oracle.final_tensors_batch(flat, off, obs=False)
            rec, _ = oracle.final_batch(flat, off)
            got = unreset.observations["info_state"].cpu().numpy()[d["env"]]
            assert (got == ti).all()
            assert (reward.cpu().numpy()[d["env"]] == rec["rewards"]).all()
        one = unreset[int(np.nonzero(done_np)[0][0])] if done_np.any() else ts[0]
        assert isinstance(one.observations["legal_actions"], list) and len(one.observations["info_state"][0]) == 2492
    assert finished_seen > n
