"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against
the oracle (oracle/coup_oracle.c), against the committed golden fixtures generated from the reference,
and -- at larger sizes -- through size-independent properties. Everything is bit-exact: the path is
integer/bit work and the tensors only hold 0, 1 and small coin counts.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover - CPU container
    pytest.skip("no CUDA device", allow_module_level=True)

from open_spiel_coup_b200 import build as _build  # noqa: E402

_build.build()
from open_spiel_coup_b200 import _lib  # noqa: E402
from open_spiel_coup_b200.vector_env import CoupVectorEnv, unpack_states  # noqa: E402

INFO, OBS = 2492, 98


def _flatten(trajs):
    flat = np.concatenate([a for a, _ in trajs]) if trajs else np.zeros(0, np.uint8)
    off = np.concatenate([[0], np.cumsum([len(a) for a, _ in trajs])]).astype(np.int64)
    return flat, off


def _u64(t):
    return t.cpu().numpy().view(np.uint64)


def check_env_against_oracle(oracle, env, dense_rows=128, expect_done=True):
    """Replays every env's device history through the oracle and compares all observables."""
    trajs = env.trajectories()
    flat, off = _flatten(trajs)
    rec, bad = oracle.final_batch(flat, off)
    assert bad == 0, "oracle rejected a trajectory the GPU produced"
    n = env.num_envs
    st = unpack_states(env.state.cpu().numpy())
    assert (st["error"] == 0).all()
    np.testing.assert_array_equal(st["move_number"], rec["move_number"])
    np.testing.assert_array_equal(st["coins"], rec["coins"])
    np.testing.assert_array_equal(st["num_cards"], rec["ncards"])
    assert (rec["is_chance"] == 0).all(), "a step must leave every env at a decision or terminal node"
    np.testing.assert_array_equal(env.current_player.cpu().numpy(), rec["cur_player"])
    np.testing.assert_array_equal(env.legal_mask.cpu().numpy().view(np.uint32), rec["legal_mask"])
    if expect_done:
        np.testing.assert_array_equal(env.done.cpu().numpy(), rec["is_terminal"])
        np.testing.assert_array_equal(env.rewards.cpu().numpy(), rec["rewards"])
        np.testing.assert_array_equal(env.returns.cpu().numpy(), rec["returns"])
    # tensors: hash of what is actually in HBM, for both players
    info_both = env.information_state_tensor(_lib.PLAYER_BOTH)
    obs_both = env.observation_tensor(_lib.PLAYER_BOTH)
    np.testing.assert_array_equal(_u64(env.tensor_row_hash(info_both)).reshape(n, 2), rec["hash_info"])
    np.testing.assert_array_equal(_u64(env.tensor_row_hash(obs_both)).reshape(n, 2), rec["hash_obs"])
    # single-view encoders agree with the both-view one
    for sel in (_lib.PLAYER_0, _lib.PLAYER_1):
        assert torch.equal(env.information_state_tensor(sel), info_both.view(n, 2, INFO)[:, sel])
        assert torch.equal(env.observation_tensor(sel), obs_both.view(n, 2, OBS)[:, sel])
    mover = torch.as_tensor(st["cur_player_move"], device=env.device)
    idx = torch.arange(n, device=env.device)
    assert torch.equal(env.information_state_tensor(_lib.PLAYER_CURRENT), info_both.view(n, 2, INFO)[idx, mover])
    assert torch.equal(env.observation_tensor(_lib.PLAYER_CURRENT), obs_both.view(n, 2, OBS)[idx, mover])
    # dense comparison on a prefix
    k = min(dense_rows, n)
    ti, to, bad = oracle.final_tensors_batch(flat[: off[k]], off[: k + 1])
    assert bad == 0
    np.testing.assert_array_equal(info_both.view(n, 2, INFO)[:k].cpu().numpy(), ti)
    np.testing.assert_array_equal(obs_both.view(n, 2, OBS)[:k].cpu().numpy(), to)
    return rec


@pytest.mark.parametrize("plain", [False, True], ids=["staged-tma", "plain-stores"])
def test_sample_step_until_all_done(oracle, plain):
    env = CoupVectorEnv(2048, seed=7, plain_store_encoder=plain)
    check_env_against_oracle(oracle, env)
    for step in range(100):
        a = env.sample_uniform()
        env.step(a)
        if step < 30 or step % 7 == 0:
            check_env_against_oracle(oracle, env)
    rec = check_env_against_oracle(oracle, env, dense_rows=2048)
    assert rec["is_terminal"].all(), "every game ends within 91 moves"
    s = env.stats()
    assert s["illegal"] == 0 and s["episodes"] == 2048
    assert sum(s["returns_hist_p0"]) == 2048 and s["returns_hist_p0"][2] == 0
    # a done env ignores further actions (rl_environment.py:301-302)
    before = env.state.clone()
    env.step(torch.zeros(2048, dtype=torch.uint8, device=env.device))
    assert torch.equal(before, env.state)
    env.check_errors()


@pytest.mark.parametrize("plain,ws", [(False, True), (False, False), (True, False)],
                         ids=["warp-specialised", "staged-tma", "plain-stores"])
def test_fused_rollout_equals_sample_then_step(oracle, plain, ws):
    a_env = CoupVectorEnv(4096 + 37, seed=99, auto_reset=True, plain_store_encoder=not plain)   # ragged tail
    b_env = CoupVectorEnv(4096 + 37, seed=99, auto_reset=True, plain_store_encoder=plain, warp_specialised=ws)
    out = torch.empty((a_env.num_envs, INFO), dtype=torch.float32, device=a_env.device)
    for step in range(80):
        acts = a_env.sample_uniform()
        a_env.step(acts)
        b_env.rollout(1, _lib.PLAYER_CURRENT, out=out)
        assert torch.equal(a_env.state, b_env.state)
        assert torch.equal(a_env.legal_mask, b_env.legal_mask)
        assert torch.equal(a_env.done, b_env.done)
        assert torch.equal(a_env.rewards, b_env.rewards)
        assert torch.equal(a_env.returns, b_env.returns)
        assert torch.equal(a_env.current_player, b_env.current_player)
        assert torch.equal(a_env.information_state_tensor(_lib.PLAYER_CURRENT), out)
        if step % 10 == 0:
            check_env_against_oracle(oracle, b_env, expect_done=False)
    assert a_env.stats() == b_env.stats()
    both = b_env.rollout(1, _lib.PLAYER_BOTH)
    assert torch.equal(both, b_env.information_state_tensor(_lib.PLAYER_BOTH))


def test_auto_reset_reports_finished_episode(oracle):
    env = CoupVectorEnv(4096, seed=3, auto_reset=True)
    prev = env.trajectories()
    finished = 0
    for step in range(60):
        acts = env.sample_uniform()
        env.step(acts)
        a = acts.cpu().numpy()
        done = env.done.cpu().numpy().astype(bool)
        rew = env.rewards.cpu().numpy()
        ret = env.returns.cpu().numpy()
        cur = env.trajectories()
        # finished episodes: previous history + the action must be terminal in the oracle, with the
        # reported rewards/returns (a move-cap ending after trailing deals is skipped: needs the deals)
        idx = np.nonzero(done)[0]
        if len(idx):
            cand = [(np.concatenate([prev[e][0], [a[e]]]).astype(np.uint8), None) for e in idx]
            flat, off = _flatten(cand)
            rec, bad = oracle.final_batch(flat, off)
            assert bad == 0
            term = rec["is_terminal"].astype(bool)
            assert term.mean() > 0.99
            np.testing.assert_array_equal(rew[idx][term], rec["rewards"][term])
            np.testing.assert_array_equal(ret[idx][term], rec["returns"][term])
            # and the env now holds a freshly dealt episode
            for e in idx:
                assert len(cur[e][0]) == 4 and (cur[e][1] == [0, 1, 0, 1]).all()
            assert (env.current_player.cpu().numpy()[idx] == 0).all()
            finished += len(idx)
        # unfinished episodes extend their previous history
        for e in np.nonzero(~done)[0][:256]:
            assert (cur[e][0][: len(prev[e][0])] == prev[e][0]).all() and cur[e][0][len(prev[e][0])] == a[e]
        prev = cur
    assert finished > 4096
    s = env.stats()
    assert s["episodes"] == finished and s["decision_steps"] == 60 * 4096 and s["illegal"] == 0
    check_env_against_oracle(oracle, env, expect_done=False)


def _scenario_steps(oracle, applies):
    """Splits a flat apply list into (deals[4], [(action, chance[<=3])...]) using the oracle to tell
    chance nodes from decision nodes."""
    s = oracle.new_state()
    deals, steps = [], []
    for a in applies:
        if oracle.current_player(s) == -1:
            if len(steps) == 0:
                deals.append(a)
            else:
                steps[-1][1].append(a)
        else:
            steps.append((a, []))
        assert oracle.apply(s, a) == 0
    return deals, steps


def test_scenario_kats_forced_chance(oracle, kat_scenarios):
    """The 14 known-answer tests of coup_test.cc, one env each, chance outcomes forced as in the test."""
    n = len(kat_scenarios)
    plans = [_scenario_steps(oracle, [s["apply"] for s in k["steps"] if "apply" in s]) for k in kat_scenarios]
    env = CoupVectorEnv(n, seed=1)
    deals = np.array([p[0] for p in plans], np.uint8)
    env.reset(forced_deals=deals)
    max_steps = max(len(p[1]) for p in plans)
    for t in range(max_steps + 1):
        # compare every scenario that has executed exactly t decision steps with the oracle state
        st = unpack_states(env.state.cpu().numpy())
        for e, (k, (d, steps)) in enumerate(zip(kat_scenarios, plans)):
            if t > len(steps):
                continue
            applied = list(d) + [x for a, ch in steps[:t] for x in [a] + ch]
            s = oracle.state_from_actions(applied)
            assert int(env.current_player[e]) == oracle.current_player(s), (k["name"], t)
            assert int(env.legal_mask[e]) & 0xFFFFFFFF == oracle.legal_mask(s), (k["name"], t)
            assert [int(x) for x in env.rewards[e]] == oracle.rewards(s), (k["name"], t)
            assert [int(x) for x in env.returns[e]] == oracle.returns(s), (k["name"], t)
            assert list(st["coins"][e]) == [s.players[0].coins, s.players[1].coins], (k["name"], t)
            for p in (0, 1):
                hand = [(s.players[p].cards[i].value << 1) | s.players[p].cards[i].state for i in range(s.players[p].num_cards)]
                assert list(st["hands"][e, p][: len(hand)]) == hand and (st["hands"][e, p][len(hand):] == 15).all(), (k["name"], t)
                assert st["last_action"][e, p] == s.players[p].last_action
            # the scenario's own expectations, where they are attached to the end of a decision step
            if t == len(steps):
                exp = k["steps"][-1].get("expect", {})
                if "current_player" in exp and exp["current_player"] >= 0:
                    assert int(env.current_player[e]) == exp["current_player"], k["name"]
                if "legal_actions" in exp:
                    assert int(env.legal_mask[e]) == sum(1 << a for a in exp["legal_actions"]), k["name"]
                if "rewards" in exp:
                    assert [int(x) for x in env.rewards[e]] == exp["rewards"], k["name"]
                if "returns" in exp:
                    assert [int(x) for x in env.returns[e]] == exp["returns"], k["name"]
                if "is_terminal" in exp:
                    assert bool(env.done[e]) == exp["is_terminal"], k["name"]
                if "coins" in exp:
                    for p, c in enumerate(exp["coins"]):
                        assert c is None or st["coins"][e, p] == c, k["name"]
                if "face_up" in exp:
                    for p, slot in exp["face_up"]:
                        assert st["hands"][e, p, slot] & 1 == 1 and st["hands"][e, p, slot] != 15, k["name"]
        if t == max_steps:
            break
        acts = env.sample_uniform().cpu().numpy()  # scenarios that already ended keep playing legally
        forced = np.full((n, 4), 0xFF, np.uint8)
        for e, (d, steps) in enumerate(plans):
            if t < len(steps):
                acts[e] = steps[t][0]
                forced[e, : len(steps[t][1])] = steps[t][1]
        acts[acts == 0xFF] = 0
        env.step(acts, forced_chance=forced)
    env.check_errors()


def _replay_golden(env, actions, offsets, records):
    """Feeds reference trajectories (forced chance) through the GPU; yields (step, per-env prefix index)."""
    n = len(offsets) - 1
    lens = np.diff(offsets)
    # per trajectory: indices of player moves
    plans = []
    for t in range(n):
        rec = records[offsets[t] + t: offsets[t + 1] + t + 1]
        acts = actions[offsets[t]: offsets[t + 1]]
        is_chance = rec["is_chance"][:-1].astype(bool)
        dec = np.nonzero(~is_chance)[0]
        plans.append((acts, dec, rec))
    deals = np.stack([p[0][:4] for p in plans]).astype(np.uint8)
    env.reset(forced_deals=deals)
    max_steps = max(len(p[1]) for p in plans)
    yield -1, np.full(n, 4)
    for k in range(max_steps):
        a = np.zeros(n, np.uint8)
        forced = np.full((n, 4), 0xFF, np.uint8)
        prefix = np.zeros(n, np.int64)
        for t, (acts, dec, rec) in enumerate(plans):
            if k < len(dec):
                m = dec[k]
                a[t] = acts[m]
                nxt = dec[k + 1] if k + 1 < len(dec) else len(acts)
                ch = acts[m + 1: nxt]
                forced[t, : len(ch)] = ch
                prefix[t] = nxt
            else:
                prefix[t] = lens[t]
        env.step(a, forced_chance=forced)
        yield k, prefix


def test_reference_trajectories_forced_replay(ref_trajectories):
    """2 000 trajectories produced by the compiled reference (incl. 100 truncated 91-move games and
    post-ExchangeReturn deals) replayed on the GPU and compared with the REFERENCE's own records."""
    actions, offsets, records = ref_trajectories
    n = len(offsets) - 1
    env = CoupVectorEnv(n, seed=5)
    base = offsets[:-1] + np.arange(n)
    for k, prefix in _replay_golden(env, actions, offsets, records):
        rec = records[base + prefix]
        np.testing.assert_array_equal(env.current_player.cpu().numpy(), rec["cur_player"])
        np.testing.assert_array_equal(env.legal_mask.cpu().numpy().view(np.uint32), rec["legal_mask"])
        if k >= 0:
            np.testing.assert_array_equal(env.done.cpu().numpy(), rec["is_terminal"])
            np.testing.assert_array_equal(env.rewards.cpu().numpy(), rec["rewards"])
        np.testing.assert_array_equal(env.returns.cpu().numpy(), rec["returns"])
        if k < 12 or k % 5 == 0:
            info = env.information_state_tensor(_lib.PLAYER_BOTH)
            obs = env.observation_tensor(_lib.PLAYER_BOTH)
            np.testing.assert_array_equal(_u64(env.tensor_row_hash(info)).reshape(n, 2), rec["hash_info"])
            np.testing.assert_array_equal(_u64(env.tensor_row_hash(obs)).reshape(n, 2), rec["hash_obs"])
    assert env.done.all()
    env.check_errors()
    # the device log reproduces the reference action lists exactly
    got = env.trajectories()
    for t in range(n):
        assert (got[t][0] == actions[offsets[t]: offsets[t + 1]]).all()


def test_playthrough_forced_replay(oracle, playthrough):
    """The reference's golden playthrough (integration_tests/playthroughs/coup.txt)."""
    hist = [s["action"] for s in playthrough["states"] if s.get("action") is not None]
    states = {s["index"]: s for s in playthrough["states"]}
    env = CoupVectorEnv(32, seed=11)
    env.reset(forced_deals=np.tile(np.array(hist[:4], np.uint8), (32, 1)))
    # which indices of History() are chance moves (the oracle only classifies nodes here; every
    # expected value below comes from coup.txt)
    chance_positions = set()
    s = oracle.new_state()
    for idx, a in enumerate(hist):
        if oracle.current_player(s) == -1:
            chance_positions.add(idx)
        assert oracle.apply(s, a) == 0
    assert chance_positions == {0, 1, 2, 3, 10, 11}
    checked = 0
    i = 4
    while True:
        st = states.get(i)
        if st is not None and "CurrentPlayer" in st and st["CurrentPlayer"] != -1:
            assert int(env.current_player[0]) == st["CurrentPlayer"]
            if "LegalActions" in st:
                assert int(env.legal_mask[0]) == sum(1 << a for a in st["LegalActions"])
            if "Returns" in st:
                assert [int(x) for x in env.returns[0]] == st["Returns"]
            if "Rewards" in st:
                assert [int(x) for x in env.rewards[0]] == st["Rewards"]
            info = env.information_state_tensor(_lib.PLAYER_BOTH).view(32, 2, INFO)
            obs = env.observation_tensor(_lib.PLAYER_BOTH).view(32, 2, OBS)
            for p in (0, 1):
                for key, t, size in ((f"InformationStateTensor({p})", info, INFO), (f"ObservationTensor({p})", obs, OBS)):
                    if key in st:
                        dense = np.zeros(size, np.float32)
                        for k, v in st[key]:
                            dense[k] = v
                        np.testing.assert_array_equal(t[0, p].cpu().numpy(), dense)
                        np.testing.assert_array_equal(t[31, p].cpu().numpy(), dense)
                        checked += 1
        if i >= len(hist):
            break
        a = hist[i]
        j = i + 1
        forced = np.full((32, 4), 0xFF, np.uint8)
        while j < len(hist) and j in chance_positions:
            forced[:, j - i - 1] = hist[j]
            j += 1
        env.step(np.full(32, a, np.uint8), forced_chance=forced)
        i = j
    assert checked >= 24
    assert bool(env.done[0]) and [int(x) for x in env.returns[0]] == [1, -1]
    got = env.trajectories()[0][0]
    assert list(got) == hist
    env.check_errors()


@pytest.mark.parametrize("plain", [False, True], ids=["staged-tma", "plain-stores"])
def test_dtypes_agree(oracle, plain):
    env = CoupVectorEnv(1000, seed=21, auto_reset=True, plain_store_encoder=plain)
    env.rollout(25)
    for sel in (_lib.PLAYER_CURRENT, _lib.PLAYER_BOTH):
        f32 = env.information_state_tensor(sel)
        u8 = env.information_state_tensor(sel, dtype=torch.uint8)
        bf = env.information_state_tensor(sel, dtype=torch.bfloat16)
        assert torch.equal(u8.float(), f32) and torch.equal(bf.float(), f32)
        assert torch.equal(env.tensor_row_hash(u8), env.tensor_row_hash(f32))
        assert torch.equal(env.tensor_row_hash(bf), env.tensor_row_hash(f32))
        o32 = env.observation_tensor(sel)
        assert torch.equal(env.observation_tensor(sel, dtype=torch.uint8).float(), o32)
        assert torch.equal(env.observation_tensor(sel, dtype=torch.bfloat16).float(), o32)
    assert f32.max() > 1  # coin counts are raw values, not one-hot (coup.cc:207-213)
    # fused rollout with u8 / bf16 output
    e2 = CoupVectorEnv(1000, seed=21, auto_reset=True, plain_store_encoder=not plain)
    e2.rollout(25)
    a = env.rollout(1, _lib.PLAYER_CURRENT, dtype=torch.uint8)
    b = e2.rollout(1, _lib.PLAYER_CURRENT, dtype=torch.bfloat16)
    assert torch.equal(a.float(), b.float())
    assert torch.equal(a.float(), env.information_state_tensor(_lib.PLAYER_CURRENT))


@pytest.mark.parametrize("plain", [False, True], ids=["staged-tma", "plain-stores"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.uint8])
@pytest.mark.parametrize("width", [2496, 2560])
def test_padded_row_stride(plain, dtype, width):
    """Rows padded to a GEMM-friendly stride: same values, zero pad columns (2496 takes the staged path,
    2560 the plain-store path)."""
    n = 1000 + 13
    env = CoupVectorEnv(n, seed=8, auto_reset=True, plain_store_encoder=plain)
    env.rollout(19)
    for sel in (_lib.PLAYER_CURRENT, _lib.PLAYER_BOTH):
        rows = n * (2 if sel == _lib.PLAYER_BOTH else 1)
        dense = env.information_state_tensor(sel, dtype=dtype)
        padded = torch.full((rows, width), 7, dtype=dtype, device=env.device)
        env.information_state_tensor(sel, out=padded)
        assert torch.equal(padded[:, :INFO], dense) and (padded[:, INFO:] == 0).all()
    e2 = CoupVectorEnv(n, seed=8, auto_reset=True, plain_store_encoder=plain)
    e2.rollout(19)
    padded = torch.full((n, width), 7, dtype=dtype, device=env.device)
    e2.rollout(1, _lib.PLAYER_CURRENT, out=padded)
    env.rollout(1)
    assert torch.equal(padded[:, :INFO], env.information_state_tensor(_lib.PLAYER_CURRENT, dtype=dtype))
    assert (padded[:, INFO:] == 0).all()
    with pytest.raises(ValueError):
        env.information_state_tensor(_lib.PLAYER_0, out=torch.empty((n, 2400), dtype=dtype, device=env.device))
    with pytest.raises(_lib.CoupError):
        env.information_state_tensor(_lib.PLAYER_0, out=torch.empty((n, 2494), dtype=dtype, device=env.device))


@pytest.mark.parametrize("plain", [False, True], ids=["staged-tma", "plain-stores"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.uint8])
def test_live_prefix_rows(plain, dtype):
    """row_stride == COUP_LIVE_INFO_STATE_SIZE: rows of 1728 elements that equal elements [0, 1728) of the full rows, and the
    columns they drop are zero in every state -- steered 91-move games included. Every entry point that takes a stride:
    the encoders (all envs, gathered, from the finished-episode ring) and the fused rollout; guard bands around the
    narrower output; the incremental contract refuses it."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from replay_check import steering_policy
    LIVE = _lib.LIVE_INFO_STATE_SIZE
    assert LIVE >= 62 + 18 * 91 and LIVE % 64 == 0
    n = 4096 + 5
    env = CoupVectorEnv(n, seed=21, auto_reset=False, plain_store_encoder=plain)
    for _ in range(100):
        env.step(steering_policy(env))
    assert int(env.move_numbers().max()) == 91
    for sel in (_lib.PLAYER_CURRENT, _lib.PLAYER_BOTH):
        rows = n * (2 if sel == _lib.PLAYER_BOTH else 1)
        dense = env.information_state_tensor(sel, dtype=dtype)
        assert float(dense[:, 62 + 18 * 91:].float().abs().sum()) == 0.0
        assert float(dense[:, 62 + 18 * 90: 62 + 18 * 91].float().abs().sum()) > 0.0     # history row 90 is in use
        guarded = torch.full((rows + 2, LIVE), 7, dtype=dtype, device=env.device)
        live = guarded[1:rows + 1]
        env.information_state_tensor(sel, out=live)
        assert torch.equal(live, dense[:, :LIVE])
        assert (guarded[0] == 7).all() and (guarded[rows + 1] == 7).all()
    ids = torch.randperm(n, device=env.device)[:777].to(torch.int32)
    got = env.information_state_tensor_gather(ids, _lib.PLAYER_BOTH, out=torch.empty((2 * 777, LIVE), dtype=dtype, device=env.device))
    assert torch.equal(got, env.information_state_tensor_gather(ids, _lib.PLAYER_BOTH, dtype=dtype)[:, :LIVE])
    # fused rollout into live-prefix rows, auto-reset on
    a = CoupVectorEnv(n, seed=9, auto_reset=True, plain_store_encoder=plain)
    b = CoupVectorEnv(n, seed=9, auto_reset=True, plain_store_encoder=plain)
    a.rollout(30); b.rollout(30)
    live = torch.full((n, LIVE), 7, dtype=dtype, device=a.device)
    a.rollout(1, _lib.PLAYER_CURRENT, out=live)
    b.rollout(1)
    assert torch.equal(live, b.information_state_tensor(_lib.PLAYER_CURRENT, dtype=dtype)[:, :LIVE])
    # rows decoded from packed records (the finished-episode ring): same prefix
    c = CoupVectorEnv(2048, seed=4, auto_reset=True, finished_ring=1 << 14, plain_store_encoder=plain)
    c.rollout(40)
    recs, dropped = c.finished_drain()
    assert dropped == 0 and len(recs) > 1000
    dev_recs = torch.as_tensor(recs.view(np.int32)).to(c.device)
    full = c.records_information_state_tensor(dev_recs, None, _lib.PLAYER_BOTH, dtype=dtype)
    live = c.records_information_state_tensor(dev_recs, None, _lib.PLAYER_BOTH,
                                              out=torch.full((2 * len(recs), LIVE), 7, dtype=dtype, device=c.device))
    assert torch.equal(live, full[:, :LIVE]) and float(full[:, LIVE:].float().abs().sum()) == 0.0
    with pytest.raises(_lib.CoupError):
        a.rollout_incremental(1, torch.zeros((2 * n, LIVE), dtype=dtype, device=a.device))
    with pytest.raises(ValueError):
        a.information_state_tensor(_lib.PLAYER_0, out=torch.empty((n, 2496), dtype=dtype, device=a.device)[:, :LIVE])


@pytest.mark.parametrize("plain", [False, True], ids=["staged-tma", "plain-stores"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.uint8])
@pytest.mark.parametrize("n", [1, 3, 32, 35, 257, 1023])
def test_encoders_stay_inside_their_output(plain, dtype, n):
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are hunted with guard bands: the
    output sits between two canary regions that must survive every encoder variant, for ragged sizes that
    exercise the partial-warp and partial-group (bulk store) paths."""
    env = CoupVectorEnv(n, seed=n, auto_reset=True, plain_store_encoder=plain)
    env.rollout(9)
    canary = 113 if dtype == torch.uint8 else 113.0
    guard = 4096
    for sel, rows, width, fn in (
            (_lib.PLAYER_CURRENT, n, INFO, env.information_state_tensor), (_lib.PLAYER_BOTH, 2 * n, INFO, env.information_state_tensor),
            (_lib.PLAYER_BOTH, 2 * n, 2496, env.information_state_tensor), (_lib.PLAYER_BOTH, 2 * n, OBS, env.observation_tensor)):
        flat = torch.full((guard + rows * width + guard,), canary, dtype=dtype, device=env.device)
        out = flat[guard: guard + rows * width].view(rows, width)
        if out.data_ptr() % 16:       # the C ABI requires 16-byte aligned tensors; shift the window if needed
            continue
        fn(sel, out=out)
        torch.cuda.synchronize()
        assert (flat[:guard] == canary).all() and (flat[guard + rows * width:] == canary).all()
        assert (out != canary).all()
    flat = torch.full((guard + n * INFO + guard,), canary, dtype=dtype, device=env.device)
    out = flat[guard: guard + n * INFO].view(n, INFO)
    if out.data_ptr() % 16 == 0:
        env.rollout(1, _lib.PLAYER_CURRENT, out=out)
        assert (flat[:guard] == canary).all() and (flat[guard + n * INFO:] == canary).all() and (out != canary).all()


def test_legal_actions_mask_dense():
    env = CoupVectorEnv(3000, seed=2, auto_reset=True)
    env.rollout(17)
    dense = env.legal_actions_mask()
    bits = (env.legal_mask.view(-1, 1) >> torch.arange(18, device=env.device, dtype=torch.int32)) & 1
    assert torch.equal(dense.int(), bits)
    assert (dense.sum(1) >= 1).all()


def test_illegal_action_sets_error_and_leaves_state():
    env = CoupVectorEnv(64, seed=4)
    before = env.state.clone()
    acts = torch.full((64,), 9, dtype=torch.uint8, device=env.device)   # Pass at turn begin is illegal
    env.step(acts)
    st = unpack_states(env.state.cpu().numpy())
    assert (st["error"] == 1).all()
    assert torch.equal(env.state & ~(1 << 29), before)
    with pytest.raises(_lib.CoupError):
        env.check_errors()
    assert env.stats()["illegal"] == 64


def test_sharding_invariance():
    """Philox streams are keyed by the GLOBAL env id: any split of the envs over handles (GPUs) gives
    the same trajectories."""
    whole = CoupVectorEnv(1024, seed=77, auto_reset=True)
    lo = CoupVectorEnv(512, seed=77, auto_reset=True, global_env_offset=0)
    hi = CoupVectorEnv(512, seed=77, auto_reset=True, global_env_offset=512)
    for _ in range(40):
        whole.rollout(1)
        lo.rollout(1)
        hi.rollout(1)
    assert torch.equal(whole.state[:512], lo.state) and torch.equal(whole.state[512:], hi.state)
    assert torch.equal(whole.history[:512], lo.history) and torch.equal(whole.history[512:], hi.history)
    other = CoupVectorEnv(512, seed=78, auto_reset=True)
    other.rollout(40)
    assert not torch.equal(other.state, lo.state)


def test_step_host_path():
    n = 2048
    env = CoupVectorEnv(n, seed=13, auto_reset=True)
    ref = CoupVectorEnv(n, seed=13, auto_reset=True)
    h_act = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_legal = torch.empty(n, dtype=torch.int32).pin_memory()
    h_cur = torch.empty(n, dtype=torch.int8).pin_memory()
    h_done = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_rew = torch.empty((n, 2), dtype=torch.int8).pin_memory()
    tensor = torch.empty((n, INFO), dtype=torch.float32, device=env.device)
    for _ in range(20):
        acts = ref.sample_uniform()
        h_act.copy_(acts.cpu())
        env.step_host(h_act, h_legal, h_cur, h_done, h_rew, tensor_out=tensor)
        ref.step(acts)
        assert torch.equal(h_legal, ref.legal_mask.cpu()) and torch.equal(h_cur, ref.current_player.cpu())
        assert torch.equal(h_done, ref.done.cpu()) and torch.equal(h_rew, ref.rewards.cpu())
        assert torch.equal(tensor, ref.information_state_tensor(_lib.PLAYER_CURRENT))


@pytest.mark.parametrize("blocking_sync", [False, True], ids=["spin-wait", "blocking-sync"])
def test_step_host_packed_and_host_policy(blocking_sync):
    """One-copy host path: the packed step word carries legal mask / current player / done / rewards /
    returns, and the host-side uniform policy draws exactly what the device sampler draws. Same results whether the
    call spins or sleeps (COUP_FLAG_BLOCKING_SYNC) while it waits for the device."""
    import ctypes as C
    n = 5000
    lib = _lib.load()
    seed, offset = 4242, 1 << 20
    env = CoupVectorEnv(n, seed=seed, auto_reset=True, global_env_offset=offset, blocking_sync=blocking_sync)
    ref = CoupVectorEnv(n, seed=seed, auto_reset=True, global_env_offset=offset)
    h_act = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_words = torch.empty(n, dtype=torch.int32).pin_memory()
    h_words.copy_(env.step_word.cpu())
    tensor = torch.empty((n, INFO), dtype=torch.float32, device=env.device)
    side = torch.cuda.Stream()
    for it in range(40):
        dev_acts = ref.sample_uniform()
        rc = lib.coup_host_sample_uniform(C.c_void_p(h_words.data_ptr()), n, seed, offset, env.step_counter,
                                          C.c_void_p(h_act.data_ptr()), 3 if it % 2 else 1)
        assert rc == 0
        assert torch.equal(h_act, dev_acts.cpu()), "host policy and device sampler disagree"
        env.step_host_packed(h_act, h_words, tensor_out=tensor, stream=side if it % 3 == 0 else None)
        ref.step(dev_acts)
        w = h_words.numpy().view(np.uint32)
        np.testing.assert_array_equal(w & 0x3FFFF, ref.legal_mask.cpu().numpy().view(np.uint32))
        cur = np.where((w >> 19) & 1, -4, (w >> 18) & 1)
        np.testing.assert_array_equal(cur, ref.current_player.cpu().numpy())
        np.testing.assert_array_equal((w >> 20) & 1, ref.done.cpu().numpy())
        np.testing.assert_array_equal(((w >> 21) & 7).astype(np.int64) - 2, ref.rewards.cpu().numpy()[:, 0])
        np.testing.assert_array_equal(((w >> 24) & 7).astype(np.int64) - 2, ref.returns.cpu().numpy()[:, 0])
        side.synchronize()
        assert torch.equal(tensor, ref.information_state_tensor(_lib.PLAYER_CURRENT))
        assert torch.equal(env.step_word, ref.step_word)
    assert env.stats() == ref.stats()


def test_statistical_parity_with_reference_workload():
    """SURVEY.md section 6 (1.6 M reference episodes): 21.20 moves/episode = 15.03 decisions + 6.17 chance;
    mean legal actions 3.59; P0 returns -2/-1/0/+1/+2 = 30.4/20.5/0/20.0/29.1 %."""
    # exactly one episode per env (no auto-reset): an unbiased sample of episodes, unlike a fixed
    # window of an auto-resetting run, which under-counts long episodes
    n = 1 << 19
    env = CoupVectorEnv(n, seed=2024, auto_reset=False)
    env.rollout(95)
    assert env.done.all()
    s = env.stats()
    eps = s["episodes"]
    assert eps == n and s["illegal"] == 0
    assert abs(s["episode_moves"] / eps - 21.20) < 0.06
    assert abs(s["decision_steps"] / eps - 15.03) < 0.05
    assert s["chance_moves"] + s["decision_steps"] == s["episode_moves"]
    legal = np.array(s["legal_count_hist"], float)
    assert abs((legal * np.arange(8)).sum() / legal.sum() - 3.59) < 0.02
    np.testing.assert_allclose(legal / legal.sum(), [0, .058, .385, .104, .014, .258, .167, .015], atol=0.004)
    ret = np.array(s["returns_hist_p0"], float) / eps
    np.testing.assert_allclose(ret, [.304, .205, 0, .200, .291], atol=0.004)
    assert s["truncated"] <= 20


def _philox4x32_10(ctr, key):
    """NumPy Philox4x32-10 (Salmon et al. SC'11) on uint32 arrays [..,4] / [..,2]."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k = [key[..., i].astype(np.uint64) for i in range(2)]
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k[0]) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c[3] ^ k[1]) & mask, p0 & mask]
        k = [(k[0] + np.uint64(W0)) & mask, (k[1] + np.uint64(W1)) & mask]
    return np.stack(c, -1).astype(np.uint32)


def test_philox_stream_matches_specification():
    # Random123 known-answer vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = _philox4x32_10(np.array(c, np.uint32), np.array(k, np.uint32))
        assert tuple(int(x) for x in got) == want
    # the device's uniform action choice = k-th legal action with k = mulhi(x, popcount(legal))
    n, seed, offset = 4096, 0x1234_5678_9ABC_DEF0, (1 << 33) + 5
    env = CoupVectorEnv(n, seed=seed, global_env_offset=offset)
    for _ in range(3):
        step = env.step_counter
        legal = env.legal_mask.cpu().numpy().view(np.uint32)
        acts = env.sample_uniform().cpu().numpy()
        genv = offset + np.arange(n, dtype=np.uint64)
        key = np.stack([(genv & 0xFFFFFFFF), ((genv >> np.uint64(32)) ^ np.uint64(seed & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)], -1).astype(np.uint32)
        ctr = np.zeros((n, 4), np.uint32)
        ctr[:, 0] = step & 0xFFFFFFFF
        ctr[:, 1] = step >> 32
        ctr[:, 3] = seed >> 32
        x = _philox4x32_10(ctr, key)[:, 0].astype(np.uint64)
        for e in range(n):
            bits = [b for b in range(18) if (int(legal[e]) >> b) & 1]
            assert acts[e] == bits[int((x[e] * np.uint64(len(bits))) >> np.uint64(32))]
        env.step(torch.as_tensor(acts, device=env.device))


def test_device_trajectories_reproduced_from_philox_and_oracle(oracle):
    """Exact, not statistical: given (seed, global env id, step counter) the device's action choices, chance
    draws and auto-reset deals are those of the documented stream -- Philox4x32-10 words mapped with
    floor(u * n / 2^32) -- so NumPy Philox + the CPU oracle reproduce every trajectory move for move."""
    n, seed, offset = 160, 0xDEADBEEF12345678, 77
    env = CoupVectorEnv(n, seed=seed, global_env_offset=offset, auto_reset=True)

    def words(step, purpose):
        genv = offset + np.arange(n, dtype=np.uint64)
        key = np.stack([genv & np.uint64(0xFFFFFFFF), ((genv >> np.uint64(32)) ^ np.uint64(seed & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)], -1).astype(np.uint32)
        ctr = np.zeros((n, 4), np.uint32)
        ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3] = step & 0xFFFFFFFF, step >> 32, purpose, seed >> 32
        return _philox4x32_10(ctr, key)

    def draw_card(s, u):
        deck = [s.deck[i] for i in range(5)]
        r = (int(u) * sum(deck)) >> 32
        c = 0
        while r >= deck[c]:
            r -= deck[c]
            c += 1
        return c

    def deal(s, us):
        k = 0
        while oracle.current_player(s) == -1:
            assert oracle.apply(s, draw_card(s, us[k])) == 0
            k += 1

    states = []
    w = words(0, 1)                                   # the reset at creation used step counter 0
    for e in range(n):
        s = oracle.new_state()
        deal(s, w[e])
        states.append(s)
    for t in range(45):
        step = env.step_counter
        env.rollout(1)
        w0, w1 = words(step, 0), words(step, 1)
        for e in range(n):
            s = states[e]
            la = oracle.legal_actions(s)
            assert oracle.apply(s, la[(int(w0[e, 0]) * len(la)) >> 32]) == 0
            if oracle.is_terminal(s):
                # an episode that ends at the action has used none of the step block's deal words: the in-place re-deal
                # draws its four cards from them -- y twice (floor(y * 15 / 2^32), then the remainder y * 15 mod 2^32), z, w
                y = int(w0[e, 1])
                s = states[e] = oracle.new_state()
                deal(s, [y, (y * 15) & 0xFFFFFFFF, int(w0[e, 2]), int(w0[e, 3])])
                continue
            deal(s, w0[e, 1:])
            if oracle.is_terminal(s):          # move cap in the middle of a deal sequence: a reset block of its own
                s = states[e] = oracle.new_state()
                deal(s, w1[e])
        got = env.trajectories()
        for e in range(n):
            s = states[e]
            assert list(got[e][0]) == [s.history_action[i] for i in range(s.history_len)], (t, e)
    assert env.stats()["episodes"] > n


@pytest.mark.parametrize("n", [1, 31, 33, 255, 257])
def test_ragged_sizes(oracle, n):
    env = CoupVectorEnv(n, seed=n, auto_reset=True, plain_store_encoder=(n % 2 == 0))
    for _ in range(12):
        out = env.rollout(1, _lib.PLAYER_BOTH)
    assert out.shape == (2 * n, INFO)
    check_env_against_oracle(oracle, env, expect_done=False)


@pytest.mark.parametrize("n,ws", [(1 << 20, True), ((1 << 20) + 300, True), (1 << 20, False)],
                         ids=["warp-specialised", "warp-specialised-ragged", "cta-per-256-envs"])
def test_full_size_properties(n, ws):
    """BASELINE config 2 size (2^20 envs): properties that do not need the oracle."""
    env = CoupVectorEnv(n, seed=1234, auto_reset=True, warp_specialised=ws)
    out = torch.empty((n, INFO), dtype=torch.float32, device=env.device)
    env.rollout(30, _lib.PLAYER_CURRENT, out=out)
    s = env.stats()
    assert s["decision_steps"] == 30 * n and s["illegal"] == 0
    # fused tensor == standalone encoder, hashed on the device
    assert torch.equal(env.tensor_row_hash(out), env.tensor_row_hash(env.information_state_tensor(_lib.PLAYER_CURRENT)))
    # structural invariants of the info-state tensor
    assert torch.equal(out[:, 0:2].sum(1), torch.ones(n, device=env.device))             # observer one-hot
    assert torch.equal(out[:, 42:44].sum(1), torch.ones(n, device=env.device))           # nobody is terminal after auto-reset
    assert torch.equal(out[:, 0], out[:, 42])                                            # observer == player to move
    moves = env.move_numbers()
    hist_rows = out[:, 62:].view(n, 135, 18).sum(2)
    assert (hist_rows <= 1).all()
    assert torch.equal((hist_rows > 0).sum(1) <= moves, torch.ones(n, dtype=torch.bool, device=env.device))
    assert (hist_rows[:, 91:] == 0).all()
    deck = unpack_states(env.state.cpu().numpy())
    total = deck["deck"].sum(1) + deck["num_cards"].sum(1)
    assert (total == 15).all()                                                           # cards are conserved
    assert (deck["coins"] <= 12).all()


def test_snapshot_restore_resumes_bit_identically(oracle):
    n = 5000
    env = CoupVectorEnv(n, seed=17, auto_reset=True)
    env.rollout(20)
    snap = env.snapshot()
    counter = env.step_counter
    out_a = torch.empty((n, INFO), dtype=torch.float32, device=env.device)
    env.rollout(25, _lib.PLAYER_CURRENT, out=out_a)
    state_a, hist_a, stats_a, word_a = env.state.clone(), env.history.clone(), env.stats(), env.step_word.clone()
    # restore into the same handle ...
    env.restore(snap)
    assert env.step_counter == counter
    out_b = torch.empty_like(out_a)
    env.rollout(25, _lib.PLAYER_CURRENT, out=out_b)
    assert torch.equal(env.state, state_a) and torch.equal(env.history, hist_a) and torch.equal(out_a, out_b)
    assert env.stats() == stats_a and torch.equal(env.step_word, word_a)
    # ... and into a fresh handle created with other parameters (seed / offset come from the snapshot)
    other = CoupVectorEnv(n, seed=999, auto_reset=True, global_env_offset=123)
    other.restore(snap)
    other.rollout(25, _lib.PLAYER_CURRENT, out=out_b)
    assert torch.equal(other.state, state_a) and torch.equal(out_a, out_b) and other.stats() == stats_a
    with pytest.raises(_lib.CoupError):
        CoupVectorEnv(n + 1, seed=1).restore(snap)
    # the wire format of the current states replays through the oracle to the same states
    texts = env.serialized_states()
    for e in range(0, n, 250):
        acts = [int(x) for x in texts[e].split("\n") if x]
        s = oracle.state_from_actions(acts)
        assert oracle.legal_mask(s) == int(env.legal_mask[e]) & 0xFFFFFFFF


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.uint8])
@pytest.mark.parametrize("auto_reset,width,n", [(True, 2492, 3000), (True, 2496, 1000 + 7), (False, 2492, 2048)])
def test_incremental_contract_equals_dense_encoder(dtype, auto_reset, width, n):
    """SURVEY.md section 8(d) contract I: a persistent both-views buffer updated in place (head, new history rows,
    cleared rows after a re-deal) must equal the dense encoder's output after every step."""
    env = CoupVectorEnv(n, seed=5, auto_reset=auto_reset)
    twin = CoupVectorEnv(n, seed=5, auto_reset=auto_reset)
    buf = torch.full((2 * n, width), 9, dtype=dtype, device=env.device)
    env.information_state_tensor(_lib.PLAYER_BOTH, out=buf)
    dense = torch.empty_like(buf)
    for step in range(120 if not auto_reset else 70):
        env.rollout_incremental(1, buf)
        twin.rollout(1)
        assert torch.equal(env.state, twin.state) and torch.equal(env.step_word, twin.step_word)
        if step % 3 == 0 or step > 60:
            env.information_state_tensor(_lib.PLAYER_BOTH, out=dense)
            assert torch.equal(buf, dense), f"step {step}"
    env.rollout_incremental(25, buf)
    env.information_state_tensor(_lib.PLAYER_BOTH, out=dense)
    assert torch.equal(buf, dense)
    assert env.stats()["episodes"] >= n and env.stats()["illegal"] == 0


def test_partial_reset_only_touches_selected_envs(oracle):
    """SyncVectorEnv.reset(envs_to_reset) (vector_env.py:68-78): unselected envs keep their state and outputs."""
    n = 4096
    env = CoupVectorEnv(n, seed=41)
    for _ in range(12):
        env.step(env.sample_uniform())
    before_state, before_hist, before_word = env.state.clone(), env.history.clone(), env.step_word.clone()
    mask = (torch.arange(n, device=env.device) % 3 == 0).to(torch.uint8)
    env.reset(envs_to_reset=mask)
    keep = mask == 0
    assert torch.equal(env.state[keep], before_state[keep]) and torch.equal(env.history[keep], before_hist[keep])
    assert torch.equal(env.step_word[keep], before_word[keep])
    moves = env.move_numbers()
    assert (moves[mask == 1] == 4).all() and (env.current_player[mask == 1] == 0).all() and (env.done[mask == 1] == 0).all()
    assert (env.legal_mask[mask == 1] == sum(1 << a for a in (0, 1, 3, 5, 6))).all()
    check_env_against_oracle(oracle, env, expect_done=False)
    # forced deals go to P1, P2, P1, P2 in this order (coup.cc:424-427)
    deals = torch.tensor([[4, 0, 2, 3]], dtype=torch.uint8).repeat(n, 1)
    env.reset(forced_deals=deals)
    st = unpack_states(env.state.cpu().numpy())
    assert (st["hands"][:, 0, :2] == [2 << 1, 4 << 1]).all() and (st["hands"][:, 1, :2] == [0 << 1, 3 << 1]).all()
    assert (st["deck"] == [2, 3, 2, 2, 2]).all()


def test_chance_distribution_matches_chance_outcomes():
    """a-5: the device sampler must draw card c with probability deck_[c] / sum(deck_) (coup.cc:1062-1077). Over 2^20
    freshly dealt games: the first deal is uniform over the 5 card types (3 of 15 each), and the second, given the
    first, has probability 2/14 for the same type and 3/14 for each other type; 5-sigma bands on the counts."""
    n = 1 << 20
    env = CoupVectorEnv(n, seed=2024)
    hist = env.history[:, 0].cpu().numpy().view(np.uint32).astype(np.int64)
    c0 = (hist & 31) - 18                      # deal to player 0: codes 18..22
    c1 = ((hist >> 5) & 31) - 23               # deal to player 1: codes 23..27
    assert c0.min() >= 0 and c0.max() <= 4 and c1.min() >= 0 and c1.max() <= 4
    first = np.bincount(c0, minlength=5)
    sigma = np.sqrt(n * 0.2 * 0.8)
    assert np.abs(first - n * 0.2).max() < 5 * sigma
    for c in range(5):
        sel = c0 == c
        m = int(sel.sum())
        second = np.bincount(c1[sel], minlength=5)
        for d in range(5):
            p = (2 if d == c else 3) / 14
            assert abs(second[d] - m * p) < 5 * np.sqrt(m * p * (1 - p)), (c, d, second[d], m * p)
