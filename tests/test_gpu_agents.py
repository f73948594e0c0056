"""GPU tests (-m gpu) of the batched learning agents (open_spiel_coup_b200/agents.py): the transition bookkeeping
of DQN / NFSP is compared, transition by transition, with a sequential re-enactment of the reference loop
(coup_experiments/scripts/nfsp.py:134-144 driving dqn.py:175-248 and nfsp.py:177-247) over the same games with time
steps produced by the CPU oracle; the learning rules are compared with plain PyTorch restatements of the reference
losses (tolerance 1e-5, fp32)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

from open_spiel_coup_b200 import agents as A  # noqa: E402
from open_spiel_coup_b200.selfplay import UniformRandomPolicy  # noqa: E402
from open_spiel_coup_b200.vector_env import CoupVectorEnv, decode_history  # noqa: E402


def test_action_0xff_sits_the_step_out():
    env = CoupVectorEnv(256, seed=3)
    env.rollout(5)
    before = [t.clone() for t in (env.state, env.history, env.step_word, env.legal_mask, env.rewards, env.returns, env.done)]
    acts = env.sample_uniform()
    skip = torch.arange(256, device=env.device) % 3 == 0
    env.step(torch.where(skip, torch.full_like(acts, 0xFF), acts))
    env.check_errors()
    after = (env.state, env.history, env.step_word, env.legal_mask, env.rewards, env.returns, env.done)
    for b, a in zip(before, after):
        assert torch.equal(b[skip], a[skip])
    moved = ~skip & (before[6] == 0)
    assert bool((env.move_numbers()[moved] > (before[0][:, 3] & 127)[moved]).all())


def _expected_transitions(oracle, games):
    """Sequential re-enactment: per seat, the DQN transitions of complete games (dqn.py:175-248), and the decision
    states of each seat per game (for the reservoir check)."""
    expected = [[], []]
    decisions = []                          # per game: per seat list of (info bytes, legal mask)
    for actions in games:
        s = oracle.new_state()
        prev = [None, None]
        dec = [[], []]
        i = 0
        while True:
            while oracle.current_player(s) == -1:
                oracle.apply(s, actions[i]); i += 1
            info = [oracle.info_state(s, p).astype(np.uint8) for p in (0, 1)]
            rewards = oracle.rewards(s)
            if oracle.is_terminal(s):
                for p in (0, 1):
                    if prev[p] is not None:
                        expected[p].append((prev[p][0].tobytes(), prev[p][1], rewards[p], info[p].tobytes(), 1, 0))
                break
            p = oracle.current_player(s)
            legal = oracle.legal_mask(s)
            if prev[p] is not None:
                expected[p].append((prev[p][0].tobytes(), prev[p][1], rewards[p], info[p].tobytes(), 0, legal))
            a = actions[i]; i += 1
            dec[p].append((info[p].tobytes(), legal))
            prev[p] = (info[p], a)
            oracle.apply(s, a)
        assert i == len(actions)
        decisions.append(dec)
    return expected, decisions


def _buffer_tuples(rb):
    t = rb.total
    assert t < rb.capacity
    return sorted(zip([bytes(x) for x in rb.info_state[:t].cpu().numpy()], rb.action[:t].cpu().tolist(),
                      [float(x) for x in rb.reward[:t].cpu().tolist()], [bytes(x) for x in rb.next_info_state[:t].cpu().numpy()],
                      rb.is_final_step[:t].cpu().tolist(), rb.legal_actions_mask[:t].cpu().tolist()))


def _collect(games):
    def on_end(env, ids):
        hist = env.history[ids].cpu().numpy().view(np.uint32)
        lens = env.move_numbers()[ids].cpu().numpy()
        games.extend([int(a) for a in acts] for acts, _ in decode_history(hist, lens))
    return on_end


def test_dqn_transitions_match_sequential_reference_loop(oracle):
    n, episodes = 64, 150
    env = CoupVectorEnv(n, seed=12)
    agents = [A.DQN(p, n, [32], replay_buffer_capacity=1 << 14, batch_size=16, min_buffer_size_to_learn=64, learn_every=8,
                    update_target_network_every=200, epsilon_start=0.6, epsilon_end=0.2, epsilon_decay_duration=2000, seed=5)
              for p in (0, 1)]
    games = []
    totals, steps = A.run_episodes(env, agents, episodes, on_episode_end=_collect(games))
    env.check_errors()
    assert len(games) == episodes and float(totals.sum()) == 0.0
    expected, decisions = _expected_transitions(oracle, games)
    for p in (0, 1):
        assert _buffer_tuples(agents[p].replay_buffer) == sorted(expected[p])
        assert agents[p].step_counter == sum(len(d[p]) for d in decisions) + episodes      # every decision + every final step
        assert agents[p].loss is not None and np.isfinite(agents[p].loss)
        assert not bool(agents[p]._prev_valid.any())
    assert steps == sum(len(d[0]) + len(d[1]) for d in decisions)
    # evaluation leaves the agents untouched and is greedy
    snap = [agents[p].replay_buffer.total for p in (0, 1)], [agents[p].step_counter for p in (0, 1)]
    A.run_episodes(env, agents, 40, is_evaluation=True)
    assert snap == ([agents[p].replay_buffer.total for p in (0, 1)], [agents[p].step_counter for p in (0, 1)])


def test_nfsp_transitions_and_reservoir(oracle):
    n, episodes = 48, 120
    env = CoupVectorEnv(n, seed=2)
    kw = dict(replay_buffer_capacity=1 << 14, reservoir_buffer_capacity=1 << 14, anticipatory_param=0.4, batch_size=16,
              min_buffer_size_to_learn=32, learn_every=8, update_target_network_every=100, epsilon_start=0.3,
              epsilon_end=0.1, epsilon_decay_duration=1000)
    agents = [A.NFSP(p, n, [32], seed=9, **kw) for p in (0, 1)]
    games = []
    A.run_episodes(env, agents, episodes, on_episode_end=_collect(games))
    env.check_errors()
    expected, decisions = _expected_transitions(oracle, games)
    for p in (0, 1):
        ag = agents[p]
        # the inner DQN records every transition of the seat, in both modes (nfsp.py:189-204)
        assert _buffer_tuples(ag.rl_agent.replay_buffer) == sorted(expected[p])
        assert ag.get_step_counter() == sum(len(d[p]) for d in decisions) + episodes
        # reservoir: (info_state, behaviour probs, legal mask) of best-response decisions only (nfsp.py:191-192)
        res = ag.reservoir_buffer._buf.all()
        k = len(ag.reservoir_buffer)
        all_dec = {}
        for d in decisions:
            for info, legal in d[p]:
                all_dec[(info, legal)] = all_dec.get((info, legal), 0) + 1
        assert 0 < k < sum(all_dec.values())
        infos = [bytes(x) for x in res["info_state"].cpu().numpy()]
        legals = res["legal_actions_mask"].cpu().tolist()
        probs = res["action_probs"].cpu().numpy()
        for info, legal, pr in zip(infos, legals, probs):
            assert (info, legal) in all_dec
            la = [a for a in range(18) if (legal >> a) & 1]
            assert abs(pr.sum() - 1) < 1e-6 and all(pr[a] == 0 for a in range(18) if a not in la)
            assert pr.max() == 1.0 or np.allclose(pr[la], 1.0 / len(la))        # greedy one-hot or epsilon-uniform
        # the inner DQN's own counter only advances in best-response mode (its step() is not called otherwise)
        assert 0 < ag.rl_agent.step_counter < ag.get_step_counter()
        sl, rl = ag.loss
        assert sl is not None and np.isfinite(sl) and rl is not None and np.isfinite(rl)
    # joint average policy: batched and single-state forms agree
    joint = A.NFSPPolicies(agents, A.MODE.average_policy)
    from open_spiel_coup_b200.spiel import load_game
    state = load_game("coup").new_initial_state()
    for c in (0, 1, 2, 3):
        state.apply_action(c)
    pr = joint.action_probabilities(state)
    assert sorted(pr) == state.legal_actions() and abs(sum(pr.values()) - 1) < 1e-5
    with agents[0].temp_mode_as(A.MODE.best_response):
        assert agents[0]._forced_mode == A.MODE.best_response
    assert agents[0]._forced_mode is None


def test_learning_rules_equal_plain_pytorch_restatement():
    """One DQN and one NFSP gradient step against hand-written fp32 restatements of dqn.py:133-171 and
    nfsp.py:120-133 (tolerance 1e-5)."""
    n = 32
    env = CoupVectorEnv(n, seed=8)
    dqn = A.DQN(0, n, [16], replay_buffer_capacity=4096, batch_size=4096 // 8, min_buffer_size_to_learn=1, learn_every=10 ** 9,
                discount_factor=0.9, loss_str="huber", seed=1)
    other = A.PolicyAgent(1, UniformRandomPolicy(), seed=1)
    A.run_episodes(env, [dqn, other], 400)
    rb = dqn.replay_buffer
    assert rb.size >= 512
    # make the sampled batch the whole buffer so that the restatement sees the same rows
    dqn._batch_size = rb.size
    import copy
    q0, t0 = copy.deepcopy(dqn._q_network), copy.deepcopy(dqn._target_q_network)
    loss = dqn.learn()
    info, nxt = rb.info_state[:rb.size].float(), rb.next_info_state[:rb.size].float()
    legal = ((rb.legal_actions_mask[:rb.size].view(-1, 1) >> torch.arange(18, device="cuda", dtype=torch.int32)) & 1).float()
    with torch.no_grad():
        max_next = (t0(nxt) + (1 - legal) * -1e9).max(-1).values
        target = rb.reward[:rb.size].float() + (1 - rb.is_final_step[:rb.size].float()) * 0.9 * max_next
    pred = q0(info).gather(1, rb.action[:rb.size].long().view(-1, 1)).view(-1)
    err = (pred - target).abs()
    ref = torch.where(err <= 1, 0.5 * err ** 2, err - 0.5).mean()
    assert abs(loss - float(ref.detach())) < 1e-5 * max(1.0, abs(float(ref.detach())))
    ref.backward()
    for pn, p0 in zip(dqn._q_network.parameters(), q0.parameters()):
        assert torch.allclose(pn, p0 - 0.01 * p0.grad, atol=1e-6)                 # one SGD step, lr 0.01
    # NFSP supervised step: mean over the batch of -sum(p * log softmax(logits))
    nf = A.NFSP(0, n, [16], reservoir_buffer_capacity=4096, anticipatory_param=1.0, batch_size=8, min_buffer_size_to_learn=1,
                learn_every=10 ** 9, seed=2)
    A.run_episodes(env, [nf, other], 100)
    res = nf.reservoir_buffer._buf.all()
    nf._batch_size = len(nf.reservoir_buffer)
    a0 = copy.deepcopy(nf.avg_network)
    loss = nf._learn()
    ref = -(res["action_probs"] * torch.log_softmax(a0(res["info_state"].float()), -1)).sum(-1).mean()
    assert abs(loss - float(ref.detach())) < 1e-5


def test_rl_resp_exploits_the_first_action_bot():
    """rl_response.rl_resp against the `first` exploitee (always the lowest legal action: Income, never blocks or
    challenges): a few thousand training games are enough for the DQN best response to beat it clearly."""
    logs = []
    records = A.rl_resp(exploitee="first", seed=3, num_train_episodes=6144, eval_every=2048, eval_episodes=512,
                        replay_buffer_capacity=100000, batch_size=32, hidden_layers_sizes=[64, 64], num_envs=512, log=logs.append)
    assert len(records) == 3 and len(logs) == 3 and logs[0].startswith("[2048] Mean episode rewards")
    for rec in records:
        assert all(-2 <= r <= 2 for r in rec["r_mean"]) and abs(rec["value"] - sum(rec["r_mean"])) < 1e-12
    assert records[-1]["aval"] == pytest.approx(sum(r["value"] for r in records) / 3)
    print("rl_resp vs first:", [r["r_mean"] for r in records])
    assert records[-1]["value"] > 2.0          # observed: both seats reach +2.0 (the bot loses both cards every game)
    with pytest.raises(RuntimeError):
        A.rl_resp(exploitee="nobody", num_train_episodes=1)


def test_train_nfsp_runs_and_evaluates():
    seen = []
    agents = A.train_nfsp(1536, [64], num_envs=256, eval_every=512, eval_func=lambda pol, ep, losses: seen.append((ep, losses)),
                          replay_buffer_capacity=20000, reservoir_buffer_capacity=20000, anticipatory_param=0.1,
                          min_buffer_size_to_learn=500, batch_size=64, learn_every=64, seed=4)
    assert [ep for ep, _ in seen] == [512, 1024, 1536]
    assert all(len(a.reservoir_buffer) > 0 and a.rl_agent.replay_buffer.size > 1000 for a in agents)
    # the joint average policy can be handed to rl_resp as the exploitee, as the reference script does
    recs = A.rl_resp(exploitee=A.NFSPPolicies(agents), num_train_episodes=512, eval_every=512, eval_episodes=256, num_envs=256)
    assert len(recs) == 1 and -4 <= recs[0]["value"] <= 4


def test_agent_cmp_both_seats():
    """agent_cmp.py:123-149: the first-action bot loses to uniform-random play from either seat; a policy against
    itself scores ~0; episode lengths count chance nodes."""
    mean, length = A.agent_cmp(UniformRandomPolicy(), A.FirstActionPolicy(), 2000, num_envs=1024, seed=2)
    assert mean > 0.5 and 8 < length < 91
    mean2, length2 = A.agent_cmp(UniformRandomPolicy(), UniformRandomPolicy(), 4000, num_envs=2048, seed=3)
    assert abs(mean2) < 0.12 and abs(length2 - 21.2) < 1.0          # 21.2 moves per uniform-random episode
