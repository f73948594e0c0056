"""GPU tests (-m gpu) of the traversal engine that keeps the level sizes on the device (cfr_traversal.DeviceTreeTraverser):
every node of the recorded trees is replayed through the oracle and the recursion of `_traverse_game_tree`
(deep_cfr.py:415-497) is redone there, exactly as for the host-driven engine (tests/test_gpu_deep_cfr._check_tree)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("no CUDA device", allow_module_level=True)

from open_spiel_coup_b200.cfr_traversal import DeviceTreeTraverser  # noqa: E402
from open_spiel_coup_b200.deep_cfr import MLP  # noqa: E402
from test_gpu_deep_cfr import _check_tree  # noqa: E402


def _as_recorded_tree(result, player, traverser):
    """The engine's level buffers in the `last_tree` format of DeepCFRSolver (what _check_tree reads)."""
    tree = []
    env = traverser.slabs[0]
    for lv, m in zip(result["levels"], result["sizes"]):
        word = lv["words"][:m].cpu().numpy().view(np.uint32)
        terminal = ((word >> 19) & 1).astype(bool)
        nt = np.flatnonzero(~terminal)
        expand = lv["expand"][:m].cpu().numpy().view(np.uint32)
        legal = ((word[nt, None] >> np.arange(18)) & 1).astype(bool)
        is_trav = ((word[nt] >> 18) & 1) == player
        local, action = [], []
        for k, i in enumerate(nt):
            for a in range(18):
                if (expand[i] >> a) & 1:
                    local.append(k)
                    action.append(a)
        off = lv["offset"][:m].cpu().numpy()
        assert (off[nt] == np.concatenate([[0], np.cumsum([bin(int(x)).count("1") for x in expand[nt]])[:-1]])).all()
        trav = np.flatnonzero(is_trav)
        rec = {"m": m, "terminal": terminal, "nt": nt, "children": len(local), "history": lv["records"][:m, :16].cpu().numpy(),
               "moves": (lv["records"][:m, 19].cpu().numpy() & 127).astype(np.int64), "word": word,
               "value": lv["value"][:m].cpu().numpy()}
        if len(local):
            rec.update(local=np.array(local), action=np.array(action), strategy=lv["strategy"][:m].cpu().numpy()[nt], legal=legal,
                       is_trav=is_trav, trav=trav, last_legal=np.array([max(a for a in range(18) if l[a]) for l in legal]),
                       regret=lv["regret"][:m].cpu().numpy()[nt][trav])
            if len(trav):
                idx = torch.as_tensor(nt[trav], device=env.device)
                rec["rows"] = env.records_information_state_tensor(lv["records"][:m].contiguous(), idx, 2, dtype=torch.uint8).cpu().numpy()
        tree.append(rec)
    return tree


@pytest.mark.parametrize("method,factor,roots", [("outcome", 1, 200), ("outcome", 2, 3), ("e-outcome", 2, 64), ("e-outcome", 3, 16)])
def test_device_traversal_matches_recursion_on_oracle(oracle, method, factor, roots):
    torch.manual_seed(3)
    nets = [MLP(2492, [32], 18).cuda() for _ in range(2)]

    def advantages(rows, cur):
        x = rows.float()
        return torch.where((cur == 0).view(-1, 1), nets[0](x), nets[1](x))

    tr = DeviceTreeTraverser(1 << 16, advantages, seed=11, sampling_method=method, outcome_factor=factor, e_outcome=0.25)
    rng = np.random.default_rng(0)
    start = None
    if method == "outcome" and factor > 1:
        # two children at EVERY traverser node: 2^(traverser decisions) leaves per root, so these trees are grown from
        # positions late in random games (as the external-sampling test of the host-driven engine does)
        from open_spiel_coup_b200.vector_env import CoupVectorEnv
        late = CoupVectorEnv(256, seed=21)
        late.rollout(22)
        alive = (late.done == 0).nonzero(as_tuple=True)[0]
        assert alive.numel() >= 10 * roots
    done = 0
    for attempt in range(20):
        player = attempt & 1
        if method == "outcome" and factor > 1:
            pick = alive[attempt * roots:(attempt + 1) * roots]
            start = late.state[pick].clone(), late.history[pick].clone(), late.step_word[pick].clone()
        try:
            res = tr.traverse(player, roots, roots=start)
        except RuntimeError as err:               # a multi-outcome tree of a long game can outgrow the capacity: draw again
            assert "capacity" in str(err) and factor > 1
            continue
        assert res["sizes"][0] == roots and res["nodes"] == sum(res["sizes"])
        tree = _as_recorded_tree(res, player, tr)
        checked = _check_tree(oracle, player, tree, method, factor, rng)
        adv, strat = tr.memory_records(res, player)
        n_trav = sum(int(r["is_trav"].sum()) for r in tree if r["children"])
        n_opp = sum(int((~r["is_trav"]).sum()) for r in tree if r["children"])
        assert adv["info_state"].shape == (n_trav, 2492) and strat["info_state"].shape == (n_opp, 2492)
        assert adv["advantage"].shape == (n_trav, 18) and strat["strategy_action_probs"].shape == (n_opp, 18)
        # the records line up with the tree: rows and regrets of the traverser's nodes in level order
        rows = np.concatenate([r["rows"] for r in tree if r["children"] and "rows" in r])
        assert (adv["info_state"].cpu().numpy() == rows).all()
        regrets = np.concatenate([r["regret"] for r in tree if r["children"] and len(r["trav"])])
        assert (adv["advantage"].cpu().numpy() == regrets).all()
        last = np.concatenate([r["last_legal"][r["trav"]] for r in tree if r["children"] and len(r["trav"])])
        assert (adv["action"].cpu().numpy() == last).all()
        for slab in tr.slabs:
            slab.check_errors()
        done += 1
        if done == 2:
            break
    assert done == 2 and checked["trav"] > 0


def test_device_traversal_overflow_is_reported():
    nets = MLP(2492, [16], 18).cuda()
    tr = DeviceTreeTraverser(256, lambda rows, cur: nets(rows.float()), seed=1, sampling_method="external")
    with pytest.raises(RuntimeError, match="capacity"):
        tr.traverse(0, 200)
