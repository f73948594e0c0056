"""GPU tests (-m gpu) of the batched tree expansion (`coup_vec_fork`) and of the Deep CFR traversals built on it
(open_spiel_coup_b200/deep_cfr.py), against the oracle: every node of a recorded traversal tree is replayed
through the oracle from its action history, and the expected payoffs / sampled regrets of
`_traverse_game_tree` (open_spiel/python/algorithms/deep_cfr.py:415-497) are recomputed by a plain recursion."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

from open_spiel_coup_b200 import _lib  # noqa: E402
from open_spiel_coup_b200.deep_cfr import DeepCFRSolver  # noqa: E402
from open_spiel_coup_b200.vector_env import CoupVectorEnv, decode_history, unpack_states  # noqa: E402
from test_gpu_parity import check_env_against_oracle  # noqa: E402


@pytest.mark.parametrize("count", [4096, 1000, 1])
def test_fork_children_match_oracle(oracle, count):
    rng = np.random.default_rng(count)
    src = CoupVectorEnv(3000, seed=5)
    src.rollout(int(rng.integers(3, 12)))                       # parents at assorted depths, some already terminal
    alive = np.flatnonzero(src.done.cpu().numpy() == 0)
    assert len(alive) > 1000
    dst = CoupVectorEnv(4096, seed=9)
    before = dst.state.clone(), dst.history.clone()
    parents = alive[rng.integers(len(alive), size=count)]
    masks = src.legal_mask.cpu().numpy().view(np.uint32)[parents]
    actions = np.array([rng.choice([a for a in range(18) if (m >> a) & 1]) for m in masks], np.uint8)
    n = dst.fork_from(src, torch.as_tensor(parents), torch.as_tensor(actions))
    assert n == count
    dst.check_errors()
    check_env_against_oracle(oracle, dst)                       # every child is a state the oracle reaches the same way
    src_traj, dst_traj = src.trajectories(), dst.trajectories()
    for i in range(count):
        pa, _ = src_traj[parents[i]]
        ca, ct = dst_traj[i]
        assert len(ca) > len(pa) and (ca[: len(pa)] == pa).all() and ca[len(pa)] == actions[i]
        assert (ct[len(pa) + 1:] >= 0).all() and ct[len(pa)] == -1          # then only chance deals
    if count < 4096:                                            # slots >= count keep their contents
        assert torch.equal(dst.state[count:], before[0][count:]) and torch.equal(dst.history[count:], before[1][count:])


def test_fork_rejects_bad_children():
    src = CoupVectorEnv(64, seed=1)
    dst = CoupVectorEnv(64, seed=2)
    with pytest.raises(_lib.CoupError):
        dst.fork_from(dst, torch.zeros(4, dtype=torch.int32), torch.zeros(4, dtype=torch.uint8))   # dst == src
    with pytest.raises(_lib.CoupError):
        dst.fork_from(src, torch.zeros(65, dtype=torch.int32), torch.zeros(65, dtype=torch.uint8))  # count > num_envs
    dst.fork_from(src, torch.tensor([0, 1, 99]), torch.tensor([0, 2, 0], dtype=torch.uint8))   # Coup with 1 coin; no parent 99
    err = unpack_states(dst.state[:3].cpu().numpy())["error"]
    assert err.tolist() == [0, 1, 1]
    with pytest.raises(_lib.CoupError):
        dst.check_errors()
    # a terminal state has no children
    src.rollout(200)
    assert bool(src.done.all())
    dst2 = CoupVectorEnv(64, seed=3)
    dst2.fork_from(src, torch.arange(8), torch.zeros(8, dtype=torch.uint8))
    assert unpack_states(dst2.state[:8].cpu().numpy())["error"].all()
    assert dst.fork_from(src, torch.zeros(0, dtype=torch.int32), torch.zeros(0, dtype=torch.uint8)) == 0


def _host(rec):
    """One level of a recorded tree on the host: decoded action histories + numpy copies of the bookkeeping."""
    out = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in rec.items()}
    out["acts"] = [[int(a) for a in acts] for acts, _ in decode_history(out["history"].view(np.uint32), out["moves"])]
    return out


def _check_tree(oracle, player, tree, method, factor, rng, per_level=1500):
    """Replays recorded nodes through the oracle from their action histories and redoes, node by node, what
    `_traverse_game_tree` (deep_cfr.py:415-497) computes there from the values of its children. All nodes of a level
    when it has at most `per_level`, else a random sample (leaves are exact, so local consistency at every level
    is the whole recursion)."""
    levels = [_host(rec) for rec in tree]
    checked = {"trav": 0, "opp": 0, "term": 0}
    for depth, rec in enumerate(levels):
        m = rec["m"]
        nxt = levels[depth + 1] if depth + 1 < len(levels) else None
        nodes = np.arange(m) if m <= per_level else rng.choice(m, per_level, replace=False)
        nt_index = {int(e): k for k, e in enumerate(rec["nt"])}
        trav_index = {int(k): t for t, k in enumerate(rec["trav"])} if rec["children"] else {}
        for i in nodes:
            acts = rec["acts"][i]
            s = oracle.state_from_actions(acts)
            if oracle.is_terminal(s):
                assert rec["terminal"][i] and rec["value"][i] == oracle.returns(s)[player]
                checked["term"] += 1
                continue
            assert not rec["terminal"][i]
            k = nt_index[int(i)]
            cur, legal = oracle.current_player(s), oracle.legal_actions(s)
            assert cur >= 0, "a fork must resolve every chance node"
            assert [a for a in range(18) if rec["legal"][k, a]] == legal
            strategy = rec["strategy"][k].astype(np.float64)
            assert abs(strategy.sum() - 1) < 1e-5 and all(strategy[a] == 0 for a in range(18) if a not in legal)
            lo, hi = np.searchsorted(rec["local"], k, "left"), np.searchsorted(rec["local"], k, "right")
            kid_actions = [int(a) for a in rec["action"][lo:hi]]
            payoff = {}
            for j, a in zip(range(lo, hi), kid_actions):      # child j of this level = slot j of the next one
                assert nxt["acts"][j][: len(acts) + 1] == acts + [a]
                assert all(c < 5 for c in nxt["acts"][j][len(acts) + 1:])
                payoff[a] = float(nxt["value"][j])
            if cur == player:
                checked["trav"] += 1
                assert rec["is_trav"][k]
                if method == "external":
                    assert kid_actions == legal                                   # every legal action (:438-441)
                elif method == "outcome":
                    assert len(kid_actions) == min(len(legal), factor) and set(kid_actions) <= set(legal)
                else:
                    assert len(kid_actions) in (1, min(len(legal), factor)) and set(kid_actions) <= set(legal)
                cfv = sum(strategy[a] * payoff.get(a, 0.0) for a in legal)
                regrets = np.zeros(18)
                for a in legal:
                    regrets[a] = payoff.get(a, 0.0) - cfv
                t = trav_index[k]
                np.testing.assert_allclose(rec["regret"][t], regrets, atol=1e-6)
                if "rows" in rec:           # kept by record_tree for levels with at most 4096 traverser nodes
                    np.testing.assert_array_equal(rec["rows"][t], oracle.info_state(s, player).astype(np.uint8))
                assert rec["last_legal"][k] == legal[-1]
                assert abs(rec["value"][i] - cfv) < 1e-9
            else:
                checked["opp"] += 1
                assert not rec["is_trav"][k]
                assert len(kid_actions) == 1 and kid_actions[0] in legal and strategy[kid_actions[0]] > 0   # :482-487
                assert rec["value"][i] == payoff[kid_actions[0]]
    assert checked["term"] > 0 and checked["trav"] + checked["opp"] > 0
    return checked


def _check_batch(oracle, solver, player, roots, method, factor, rng, start=None):
    before = (solver.advantage_buffers[player].add_calls, solver.strategy_buffer.add_calls)
    _, nodes = solver._traverse_batch(player, roots, roots=start)
    tree = solver.last_tree
    assert nodes == sum(rec["m"] for rec in tree) and tree[0]["m"] == roots
    _check_tree(oracle, player, tree, method, factor, rng)
    n_trav = sum(int(rec["is_trav"].sum()) for rec in tree if rec["children"])
    n_opp = sum(int((~rec["is_trav"]).sum()) for rec in tree if rec["children"])
    assert solver.advantage_buffers[player].add_calls - before[0] == n_trav       # one AdvantageMemory per traverser node
    assert solver.strategy_buffer.add_calls - before[1] == n_opp                  # one StrategyMemory per opponent node
    for slab in solver._slabs:
        slab.check_errors()
    return tree


@pytest.mark.parametrize("method,factor,roots,fused", [("outcome", 1, 64, True), ("outcome", 1, 64, False), ("outcome", 2, 8, True),
                                                       ("outcome", 2, 8, False), ("e-outcome", 3, 8, True)])
def test_traversal_tree_matches_recursion_on_oracle(oracle, method, factor, roots, fused):
    solver = DeepCFRSolver(policy_network_layers=(32,), advantage_network_layers=(32,), num_traversals=roots,
                           sampling_method=method, outcome_factor=factor, e_outcome=0.25, memory_capacity=1 << 20,
                           max_nodes=1 << 12, max_tree_nodes=1 << 20, seed=11, record_tree=True, fused_expand=fused)
    rng = np.random.default_rng(0)
    done = 0
    for attempt in range(12):         # multi-outcome trees of a long game can outgrow the node budget: draw again
        try:
            _check_batch(oracle, solver, attempt & 1, roots, method, factor, rng)
            done += 1
        except RuntimeError as err:
            assert "max_tree_nodes" in str(err) and factor > 1
        if done == 2:
            break
    assert done == 2 and solver.get_environment_steps() > 0


def test_external_sampling_from_late_positions(oracle):
    """External sampling expands EVERY legal action of the traverser (deep_cfr.py:438-441); from the deal that is
    ~10^7 nodes per traversal in Coup, so the trees are grown from positions late in random games, one root at a
    time under a node budget, until three of them have been checked, at least one wider than a slab piece."""
    solver = DeepCFRSolver(policy_network_layers=(32,), advantage_network_layers=(32,), sampling_method="external",
                           memory_capacity=1 << 20, max_nodes=1 << 12, max_tree_nodes=150000, seed=4, record_tree=True)
    env = CoupVectorEnv(512, seed=21)
    env.rollout(24)
    alive = (env.done == 0).nonzero(as_tuple=True)[0]
    assert alive.numel() >= 20
    rng = np.random.default_rng(1)
    checked, widest = 0, 0
    for j, e in enumerate(alive.tolist()):
        start = env.state[e: e + 1], env.history[e: e + 1], env.step_word[e: e + 1]
        try:
            tree = _check_batch(oracle, solver, j & 1, 1, "external", 1, rng, start=start)
        except RuntimeError as err:                         # this position's tree is over the budget: next one
            assert "max_tree_nodes" in str(err)
            continue
        checked += 1
        widest = max(widest, max(rec["m"] for rec in tree))
        if checked >= 3 and widest > (1 << 12) // 7:
            break
    assert checked >= 3 and widest > (1 << 12) // 7, (checked, widest)


def test_tree_budget_raises_before_memory_runs_out():
    solver = DeepCFRSolver(policy_network_layers=(16,), advantage_network_layers=(16,), sampling_method="external",
                           max_nodes=1 << 12, max_tree_nodes=20000, seed=2)
    with pytest.raises(RuntimeError, match="max_tree_nodes"):
        solver.traverse(0, 4)


def test_opponent_sampling_follows_matched_regrets():
    """Many traversals from the same seed of the network: the empirical frequency of the opponent's first action
    matches the matched-regret strategy at the root info state (deep_cfr.py:482-487)."""
    solver = DeepCFRSolver(policy_network_layers=(16,), advantage_network_layers=(16,), sampling_method="outcome",
                           num_traversals=1 << 14, max_nodes=1 << 15, seed=3, record_tree=True)
    solver.traverse(1, 1 << 14)              # traverser is player 1, so player 0 (to move at every root) is sampled
    rec = solver.last_tree[0]
    assert not bool(rec["is_trav"].any())
    strategy = rec["strategy"].double()
    counts = torch.zeros(18, dtype=torch.float64, device=strategy.device).index_add_(0, rec["action"], torch.ones(rec["action"].numel(), dtype=torch.float64, device=strategy.device))
    expect = strategy.sum(0)
    z = (counts - expect) / expect.clamp_min(1.0).sqrt()
    assert float(z.abs().max()) < 5.0


def test_solver_runs_and_learns_something():
    solver = DeepCFRSolver(policy_network_layers=(64, 64), advantage_network_layers=(64, 64), num_iterations=3,
                           num_traversals=256, learning_rate=1e-3, batch_size_advantage=128, batch_size_strategy=256,
                           memory_capacity=100000, policy_network_train_steps=20, advantage_network_train_steps=20,
                           reinitialize_advantage_networks=False, sampling_method="outcome", outcome_factor=1,
                           max_nodes=1 << 16, seed=5)
    policy_net, adv_losses, policy_loss = solver.solve()
    assert solver.iteration == 4 and len(adv_losses[0]) == 3 and len(adv_losses[1]) == 3
    assert all(l is not None and np.isfinite(l) for p in (0, 1) for l in adv_losses[p])
    assert policy_loss is not None and np.isfinite(policy_loss)
    assert len(solver.strategy_buffer) > 1000 and min(len(b) for b in solver.advantage_buffers) > 1000
    its = solver.strategy_buffer.all()["iteration"]
    assert int(its.min()) == 1 and int(its.max()) == 3
    # acting with the averaged policy: batched and single-state forms agree, probabilities live on legal actions
    from open_spiel_coup_b200.spiel import load_game
    state = load_game("coup").new_initial_state()
    for c in (0, 1, 2, 3):
        state.apply_action(c)
    probs = solver.action_probabilities(state)
    assert sorted(probs) == state.legal_actions() and abs(sum(probs.values()) - 1) < 1e-5
    # and the trained policy network plugs into the batched evaluation loop (agent_cmp.py)
    import copy
    from open_spiel_coup_b200.selfplay import UniformRandomPolicy, evaluate_policies
    means, steps = evaluate_policies([copy.deepcopy(policy_net), UniformRandomPolicy()], 2000, device=0, seed=1)
    assert abs(means[0] + means[1]) < 1e-9 and -2 <= means[0] <= 2 and steps > 2000


def _inclusion_probability_top2(p):
    """P(action a is among the first two of a sequential draw without replacement with probabilities p)."""
    out = np.zeros_like(p)
    for a in range(len(p)):
        out[a] = p[a] + sum(p[b] * p[a] / (1 - p[b]) for b in range(len(p)) if b != a and p[b] < 1)
    return out


@pytest.mark.parametrize("fused", [True, False])
def test_outcome_sampling_without_replacement_distribution(fused):
    """The traverser's children at the root (outcome sampling, factor 2): inclusion frequencies against the analytic
    probabilities of np.random.choice(size=2, replace=False, p=0.6 uniform + 0.4 strategy) (deep_cfr.py:455-466),
    summed over the 2^14 roots with their individual strategies; 5-sigma band."""
    solver = DeepCFRSolver(policy_network_layers=(16,), advantage_network_layers=(16,), sampling_method="outcome", outcome_factor=2,
                           num_traversals=1 << 14, max_nodes=1 << 17, max_tree_nodes=1 << 24, seed=8, record_tree=True,
                           fused_expand=fused)
    try:
        solver.traverse(0, 1 << 14)                      # player 0 moves first: every root is a traverser node
    except RuntimeError as err:                          # deeper levels may outgrow the budget; the root level is recorded
        assert "max_tree_nodes" in str(err)
    rec = solver.last_tree[0]
    assert bool(rec["is_trav"].all())
    strategy = rec["strategy"].double().cpu().numpy()
    legal = rec["legal"].cpu().numpy()
    expect = np.zeros(18)
    for s, l in zip(strategy, legal):
        idx = np.flatnonzero(l)
        p = 0.6 / len(idx) + 0.4 * s[idx]
        expect[idx] += _inclusion_probability_top2(p / p.sum())
    counts = np.bincount(rec["action"].cpu().numpy(), minlength=18).astype(np.float64)
    per_parent = np.bincount(rec["local"].cpu().numpy(), minlength=len(legal))
    assert (per_parent == np.minimum(legal.sum(1), 2)).all()
    z = (counts - expect) / np.sqrt(np.maximum(expect, 1.0))
    assert np.abs(z).max() < 5.0, (counts, expect)
