"""GPU tests (-m gpu) of the on-device masked policy sampling and the self-play data generator. The sampling
kernel is floating point: it is compared with a plain PyTorch fp32 statement of the reference rule
(nfsp.py:154-167) with tolerance 1e-6 on the probabilities."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

from open_spiel_coup_b200 import _lib  # noqa: E402
from open_spiel_coup_b200.selfplay import MLPPolicy, SelfPlayDataGen, masked_action_probs  # noqa: E402
from open_spiel_coup_b200.vector_env import CoupVectorEnv  # noqa: E402


def test_sample_policy_matches_torch_reference():
    n = 8192
    env = CoupVectorEnv(n, seed=5, auto_reset=True)
    env.rollout(23)
    g = torch.Generator(device="cuda").manual_seed(1)
    logits = torch.randn(n, 18, device="cuda", generator=g) * 4
    probs = torch.empty(n, 18, device="cuda")
    acts = env.sample_policy(logits, probs_out=probs)
    ref = masked_action_probs(logits, env.legal_mask)
    assert torch.allclose(probs, ref, rtol=1e-5, atol=1e-6)            # tolerance 1e-6 (fp32 softmax)
    assert torch.allclose(probs.sum(1), torch.ones(n, device="cuda"), atol=1e-5)
    legal = ((env.legal_mask.view(-1, 1) >> torch.arange(18, device="cuda", dtype=torch.int32)) & 1).bool()
    assert legal.gather(1, acts.long().view(-1, 1)).all(), "sampled an illegal action"
    assert (probs[~legal] == 0).all()
    # deterministic in (seed, env id, step counter); bf16 logits accepted
    assert torch.equal(acts, env.sample_policy(logits))
    acts_bf = env.sample_policy(logits.bfloat16(), probs_out=probs)
    assert torch.allclose(probs, masked_action_probs(logits.bfloat16().float(), env.legal_mask), rtol=1e-5, atol=1e-6)
    assert legal.gather(1, acts_bf.long().view(-1, 1)).all()
    env.step(acts)
    env.check_errors()


def test_sample_policy_frequencies():
    """All envs are fresh episodes (legal set {0,1,3,5,6}); with shared logits the empirical action
    frequencies must match the masked softmax (chi-square, 4 d.o.f.)."""
    n = 1 << 18
    env = CoupVectorEnv(n, seed=77)
    logits = torch.tensor([0.3, -1.0, 5.0, 1.2, 0.0, 0.7, -0.4] + [0.0] * 11, device="cuda").repeat(n, 1)
    acts = env.sample_policy(logits)
    p = masked_action_probs(logits[:1], env.legal_mask[:1])[0].cpu().numpy()
    counts = np.bincount(acts.cpu().numpy(), minlength=18)
    assert counts[[2, 4] + list(range(7, 18))].sum() == 0
    chi2 = sum((counts[a] - n * p[a]) ** 2 / (n * p[a]) for a in (0, 1, 3, 5, 6))
    assert chi2 < 23.5, chi2      # p < 1e-4 for 4 degrees of freedom


def test_selfplay_datagen_runs_and_records():
    torch.manual_seed(0)
    gen = SelfPlayDataGen(num_envs=4096, policy=MLPPolicy(), seed=9, reservoir_capacity=10000)
    for _ in range(40):
        acts = gen.step()
        legal = (gen.legal_before >> acts.to(torch.int32)) & 1
        assert legal.all()
    s = gen.env.stats()
    assert s["illegal"] == 0 and s["decision_steps"] == 40 * 4096 and s["episodes"] > 4096
    r = gen.recorder
    assert r.reservoir_size == 10000 and r.offered == 40 * 4096
    info, probs, mask = r.sample_reservoir(256, dtype=torch.uint8)
    assert torch.allclose(probs.sum(1), torch.ones(256, device="cuda"), atol=1e-4)
    # a stored record is a real info state: exactly one observer bit, and it equals the player to move
    f = info.float()
    assert torch.equal(f[:, 0:2].sum(1), torch.ones(256, device="cuda"))
    assert torch.equal(f[:, 0], f[:, 42])
    # stored probabilities vanish outside the stored legal mask
    bits = ((mask.view(-1, 1) >> torch.arange(18, device="cuda", dtype=torch.int32)) & 1).bool()
    assert (probs[~bits] == 0).all()


def test_selfplay_datagen_torch_reservoir_still_works():
    gen = SelfPlayDataGen(num_envs=1024, policy=MLPPolicy(), seed=9, reservoir_capacity=3000, torch_reservoir=True)
    gen.run(10)
    assert gen.reservoir.size == 3000 and gen.recorder is None


def test_device_recorder_reservoir(oracle):
    """The reservoir kept by the step kernel: every stored record is the decision it claims to be (its running index
    t = step * num_envs + env names the step and the env; row, probabilities and legal mask are compared with what
    the env showed at that step), slot t holds element t while the buffer fills, and afterwards the stored indices
    are a uniform sample of everything offered (nfsp.py:322-371)."""
    from open_spiel_coup_b200.selfplay import DeviceRecorder
    n, cap, steps = 512, 4096, 64
    env = CoupVectorEnv(n, seed=21, auto_reset=True)
    rec = DeviceRecorder(env, reservoir_capacity=cap, seed=1)
    g = torch.Generator(device="cuda").manual_seed(3)
    rows, probs_log, legal_log = [], [], []
    for t in range(steps):
        info = env.information_state_tensor(_lib.PLAYER_CURRENT, dtype=torch.uint8)
        logits = torch.randn(n, 18, device="cuda", generator=g)
        probs = torch.empty(n, 18, device="cuda")
        acts = env.sample_policy(logits, probs_out=probs)
        rows.append(info.cpu().numpy()); probs_log.append(probs.cpu().numpy()); legal_log.append(env.legal_mask.cpu().numpy())
        rec.step(acts, probs)
        if (t + 1) * n == cap:                                   # just full: slot i holds element i
            assert torch.equal(rec.res_records[:, 22].long(), torch.arange(cap, device="cuda"))
    env.check_errors()
    assert rec.reservoir_size == cap and rec.offered == steps * n
    t_idx = (rec.res_records[:, 22].long() & 0xFFFFFFFF).cpu().numpy()
    assert len(np.unique(t_idx)) == cap                                           # an element is stored at most once
    dec = env.records_information_state_tensor(rec.res_records, None, _lib.PLAYER_FROM_RECORD, dtype=torch.uint8).cpu().numpy()
    st, e = t_idx // n, t_idx % n
    assert (rec.res_records[:, 20].cpu().numpy() == e).all()
    assert (dec == np.stack(rows)[st, e]).all()
    assert (rec.res_probs.cpu().numpy() == np.stack(probs_log)[st, e]).all()
    assert ((rec.res_records[:, 21].cpu().numpy() & 0x3FFFF) == np.stack(legal_log)[st, e]).all()
    # uniform over [0, offered): chi-square over 8 equal bins of the running index (7 d.o.f., p < 1e-4 at 29.9)
    counts = np.bincount(t_idx * 8 // (steps * n), minlength=8)
    chi2 = float(((counts - cap / 8) ** 2 / (cap / 8)).sum())
    assert chi2 < 29.9, (chi2, counts)
    info, p, mask = rec.sample_reservoir(64, dtype=torch.bfloat16)
    assert info.shape == (64, 2492) and torch.allclose(p.sum(1), torch.ones(64, device="cuda"), atol=1e-4)


def test_device_recorder_replay_matches_reference_bookkeeping(oracle):
    """The replay transitions written by the step kernel (auto-reset mode, packed records) against a sequential
    re-enactment of the reference loop (nfsp.py:134-144 driving DQN.step/add_transition, dqn.py:175-246) over the same
    games, with time steps produced by the CPU oracle. Finished games come from the ring."""
    from open_spiel_coup_b200.selfplay import DeviceRecorder
    from open_spiel_coup_b200.vector_env import decode_finished_records
    n, steps = 96, 70
    env = CoupVectorEnv(n, seed=31, auto_reset=True, finished_ring=1 << 12)
    rec = DeviceRecorder(env, replay_capacity=1 << 15)
    for _ in range(steps):
        rec.step(env.sample_uniform())
    env.check_errors()
    recs, dropped = env.finished_drain()
    assert dropped == 0
    fin = decode_finished_records(recs)
    open_hist = env.trajectories()
    expected = []

    def time_step(s):
        return {"info": [oracle.info_state(s, p).astype(np.uint8) for p in (0, 1)], "legal": oracle.legal_mask(s),
                "rewards": oracle.rewards(s), "cur": oracle.current_player(s), "last": oracle.is_terminal(s)}

    def play(actions):
        s = oracle.new_state()
        prev = [None, None]
        i = 0
        while True:
            while oracle.current_player(s) == -1:                       # rl_environment resolves chance nodes
                oracle.apply(s, actions[i]); i += 1
            ts = time_step(s)
            if ts["last"]:
                for p in (0, 1):                                        # every agent sees the final time step
                    if prev[p] is not None:
                        expected.append((prev[p][0][p].tobytes(), prev[p][1], ts["rewards"][p], ts["info"][p].tobytes(), 1, 0))
                return
            p = ts["cur"]
            if i >= len(actions):
                return                                                  # game still running on the device
            if prev[p] is not None:
                expected.append((prev[p][0][p].tobytes(), prev[p][1], ts["rewards"][p], ts["info"][p].tobytes(), 0, ts["legal"]))
            a = actions[i]; i += 1
            prev[p] = (ts["info"], a)
            oracle.apply(s, a)

    for acts, _ in fin["trajectories"]:
        play(list(acts))
    for e in range(n):
        play(list(open_hist[e][0]))
    total = int(rec.replay_total.item())
    assert total == len(expected) and total < rec.replay_capacity
    info, action, reward, nxt, final, legal = rec.decode_transitions(torch.arange(total, device="cuda"), dtype=torch.uint8)
    got = sorted(zip([bytes(x) for x in info.cpu().numpy()], action.cpu().tolist(), [float(x) for x in reward.cpu().tolist()],
                     [bytes(x) for x in nxt.cpu().numpy()], final.cpu().tolist(), legal.cpu().tolist()))
    assert got == sorted(expected)
    assert len(fin["env"]) > n and int(final.sum()) > n
    # tickets are a permutation of 0..total-1: one slot per transition
    tickets = rec.transitions[:total, 0, 22].long() & 0xFFFFFFFF
    assert torch.equal(torch.sort(tickets).values, torch.arange(total, device="cuda"))


def test_batched_evaluation_random_vs_random():
    """agent_cmp / eval_against_fixed_bots batched on the device. Uniform-random vs uniform-random must
    reproduce the reference's workload statistics (P0 returns -2/-1/+1/+2 =
    30.4/20.5/20.0/29.1 % => mean -0.031) and be exactly zero-sum."""
    from open_spiel_coup_b200.selfplay import UniformRandomPolicy, eval_against_fixed_bots, evaluate_policies
    n = 300_000
    mean, steps = evaluate_policies([UniformRandomPolicy(), UniformRandomPolicy()], n, num_envs=1 << 17, seed=5)
    assert mean[0] == -mean[1]
    assert abs(mean[0] - (-0.031)) < 0.02
    assert abs(steps / n - 15.03) < 0.08
    # an MLP policy in seat 0 against the random bot: runs, legal, zero-sum
    torch.manual_seed(1)
    seat = eval_against_fixed_bots([MLPPolicy(padded_input_size=2496), MLPPolicy(padded_input_size=2496)],
                                   [UniformRandomPolicy(), UniformRandomPolicy()], 20_000, num_envs=1 << 14, seed=6)
    assert len(seat) == 2 and all(-2 <= x <= 2 for x in seat)


def test_replay_recorder_matches_reference_bookkeeping(oracle):
    """Every transition the batched recorder emits is compared with a sequential re-enactment of the reference
    loop (nfsp.py:134-144 driving DQN.step/add_transition, dqn.py:175-246) over the same games, with time steps
    produced by the CPU oracle."""
    from open_spiel_coup_b200.selfplay import ReplayRecorder, UniformRandomPolicy
    from open_spiel_coup_b200.vector_env import decode_history
    n, steps = 96, 70
    games = [[] for _ in range(n)]        # finished games per env: full action lists

    def on_end(rec, ids):
        hist = rec.env.history[ids].cpu().numpy().view(np.uint32)
        lens = rec.env.move_numbers()[ids].cpu().numpy()
        for e, (acts, _) in zip(ids.cpu().tolist(), decode_history(hist, lens)):
            games[e].append(list(acts))

    rec = ReplayRecorder(n, policy=UniformRandomPolicy(), seed=31, replay_capacity=1 << 15, on_episode_end=on_end)
    for _ in range(steps):
        rec.step()
    # unfinished games: their pending decisions have emitted only the "acted again" transitions
    open_hist = rec.env.trajectories()
    expected = []

    def time_step(s):
        return {"info": [oracle.info_state(s, p).astype(np.uint8) for p in (0, 1)], "legal": oracle.legal_mask(s),
                "rewards": oracle.rewards(s), "cur": oracle.current_player(s), "last": oracle.is_terminal(s)}

    def play(actions, finished):
        s = oracle.new_state()
        prev = [None, None]
        i = 0
        while True:
            while oracle.current_player(s) == -1:                       # rl_environment resolves chance nodes
                oracle.apply(s, actions[i]); i += 1
            ts = time_step(s)
            if ts["last"]:
                for p in (0, 1):                                        # every agent sees the final time step
                    if prev[p] is not None:
                        expected.append((prev[p][0][p].tobytes(), prev[p][1], ts["rewards"][p], ts["info"][p].tobytes(), 1, 0))
                return
            p = ts["cur"]
            if i >= len(actions):
                return                                                  # game still running on the device
            if prev[p] is not None:
                expected.append((prev[p][0][p].tobytes(), prev[p][1], ts["rewards"][p], ts["info"][p].tobytes(), 0, ts["legal"]))
            a = actions[i]; i += 1
            prev[p] = (ts["info"], a)
            oracle.apply(s, a)

    for e in range(n):
        for g in games[e]:
            play(g, True)
        play(list(open_hist[e][0]), False)
    rb = rec.replay
    assert rb.total == len(expected) and rb.total < rb.capacity
    got = sorted(zip([bytes(x) for x in rb.info_state[:rb.total].cpu().numpy()], rb.action[:rb.total].cpu().tolist(),
                     [float(x) for x in rb.reward[:rb.total].cpu().tolist()],
                     [bytes(x) for x in rb.next_info_state[:rb.total].cpu().numpy()],
                     rb.is_final_step[:rb.total].cpu().tolist(), rb.legal_actions_mask[:rb.total].cpu().tolist()))
    assert got == sorted(expected)
    assert sum(len(g) for g in games) > 0 and rb.is_final_step[:rb.total].sum() > 0


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 6e-2)])
def test_bucketed_first_layer_equals_dense_forward(dtype, tol):
    """The K-truncated first layer (selfplay.bucketed_policy_forward) drops only exact zeros: its logits equal the
    dense forward of a plain PyTorch fp32 reference up to summation order (tolerance 2e-4 in fp32; bf16 inputs and
    weights are compared with the bf16 dense forward at 6e-2 on logits of magnitude ~1)."""
    from open_spiel_coup_b200.selfplay import PADDED_INFO_STATE_SIZE, _BUCKET_K, _BUCKET_MOVES, bucketed_policy_forward
    n = 20000
    env = CoupVectorEnv(n, seed=77, auto_reset=True)
    torch.manual_seed(0)
    policy = MLPPolicy(hidden_sizes=(256, 128), padded_input_size=PADDED_INFO_STATE_SIZE).to("cuda", dtype).eval()
    buf = torch.zeros((n, PADDED_INFO_STATE_SIZE), dtype=dtype, device="cuda")
    dense_in = torch.zeros_like(buf)
    for steps in (0, 3, 40):                       # fresh deals, early game, desynchronised mid-games
        env.rollout(steps)
        logits, perm, rows = bucketed_policy_forward(env, policy, buf)
        env.information_state_tensor(_lib.PLAYER_CURRENT, out=dense_in)
        assert torch.equal(rows, dense_in[perm])                                   # rows are the env rows, permuted
        assert torch.equal(torch.sort(perm).values, torch.arange(n, device="cuda"))
        moves = env.move_numbers()[perm]
        assert bool((torch.bucketize(moves, torch.tensor(_BUCKET_MOVES[:-1], device="cuda")).diff() >= 0).all())
        with torch.no_grad():
            ref = policy.float()(dense_in.float()) if dtype == torch.float32 else policy(dense_in).float()
        assert float((logits.float() - ref).abs().max()) < tol
        policy.to(dtype)
    # every bucket's K covers its move numbers
    assert all(k >= min(PADDED_INFO_STATE_SIZE, 62 + 18 * m) for k, m in zip(_BUCKET_K, _BUCKET_MOVES))
    # nothing non-zero beyond 62 + 18 * move_number (the property the truncation relies on)
    cols = torch.arange(PADDED_INFO_STATE_SIZE, device="cuda").view(1, -1)
    beyond = cols >= (62 + 18 * env.move_numbers()).view(-1, 1)
    assert float(dense_in.float()[beyond].abs().sum()) == 0.0


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 6e-2)])
def test_live_column_first_layer_equals_dense_forward(dtype, tol):
    """The inference path of MLPPolicy runs the first layer over the first LIVE_INFO_STATE_SIZE columns only. Everything
    beyond element 62 + 18 * 91 = 1700 is zero in every reachable state (steered 91-move games included), so the logits
    equal the dense nn.Sequential forward up to summation order (2e-4 in fp32, 6e-2 for bf16 logits of magnitude ~1)."""
    from open_spiel_coup_b200.selfplay import LIVE_INFO_STATE_SIZE, PADDED_INFO_STATE_SIZE
    assert LIVE_INFO_STATE_SIZE >= 62 + 18 * 91 and LIVE_INFO_STATE_SIZE % 8 == 0
    n = 8192
    env = CoupVectorEnv(n, seed=5, auto_reset=False)
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from replay_check import steering_policy      # prefers Exchange / Pass / ExchangeReturn: reaches the move cap
    for _ in range(100):
        env.step(steering_policy(env))
    assert int(env.move_numbers().max()) == 91
    x = torch.zeros((n, PADDED_INFO_STATE_SIZE), dtype=dtype, device="cuda")
    env.information_state_tensor(_lib.PLAYER_0, out=x)
    assert float(x[:, 62 + 18 * 91:].float().abs().sum()) == 0.0
    assert float(x[:, 62 + 18 * 90: 62 + 18 * 91].float().abs().sum()) > 0.0       # row 90 (move 91) is in use
    torch.manual_seed(1)
    policy = MLPPolicy(hidden_sizes=(256, 128), padded_input_size=PADDED_INFO_STATE_SIZE).to("cuda", dtype).eval()
    with torch.no_grad():
        fast = policy(x)                  # live columns, bias + ReLU in the GEMM epilogue
        ref = policy.net(x)               # plain nn.Sequential over all 2496 columns
    assert float((fast.float() - ref.float()).abs().max()) < tol
