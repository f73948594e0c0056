"""Deep CFR on Coup with the game-tree traversals batched on the device.

Mirror of the reference's `open_spiel/python/algorithms/deep_cfr.py` (`DeepCFRSolver`, as driven by
`coup_experiments/scripts/deep_cfr.py:86-124`): same constructor arguments, the same three sampling methods
(`external`, `outcome`, `e-outcome`), the same memory records and losses. What changes is HOW the tree is walked.
The reference recurses one `state.child(action)` at a time in Python and calls the advantage network with a batch
of one at every node (`deep_cfr.py:415-497,499-525`). Here `num_traversals` roots are expanded level by level:

  level L (M nodes resident in one env slab)
    -> info-state rows of the player to move             (coup_vec_information_state_tensor_gather, u8)
    -> advantage nets on all rows at once, regret matching (`regret_matching`, :499-525)
    -> traverser nodes: every legal action (external) or a sample without replacement (outcome, :441-466);
       opponent nodes: one action from the matched regrets (:482-492)
    -> coup_vec_fork: all children of the level in one launch into the other slab (chance nodes are resolved
       inside the fork from the child's own Philox stream, :430-434)
  then one backward sweep over the levels computes expected payoffs, counterfactual values and sampled regrets
  (:468-480) and appends `AdvantageMemory` / `StrategyMemory` records to device-resident reservoir buffers.

Only the per-level bookkeeping (a few dozen small torch ops) runs from Python; states never leave HBM.
"""
import collections
import dataclasses
import math
import os

import ctypes

import torch
from torch import nn

from ._lib import INFO_STATE_SIZE, NUM_DISTINCT_ACTIONS, PLAYER_CURRENT, check
from .vector_env import CoupVectorEnv

AdvantageMemory = collections.namedtuple("AdvantageMemory", "info_state iteration advantage action")   # deep_cfr.py:36-37
StrategyMemory = collections.namedtuple("StrategyMemory", "info_state iteration strategy_action_probs")  # deep_cfr.py:39-40


def _legal_bool(legal_bits):
    """uint32 legal masks [M] -> bool [M, 18]."""
    bits = torch.arange(NUM_DISTINCT_ACTIONS, device=legal_bits.device, dtype=torch.int32)
    return ((legal_bits.view(-1, 1).to(torch.int32) >> bits) & 1).to(torch.bool)


def regret_matching(advantages, legal):
    """`_sample_action_from_advantage` (deep_cfr.py:499-525) for a batch: positive parts of the advantages over the
    legal actions, normalised; when none is positive, probability one on the legal action with the largest raw
    advantage (the first such action, as Python's `max` over the ascending legal list returns)."""
    pos = advantages.clamp_min(0.0) * legal
    total = pos.sum(-1, keepdim=True)
    best = torch.where(legal, advantages, torch.full_like(advantages, -math.inf)).argmax(-1)
    fallback = torch.zeros_like(pos).scatter_(-1, best.view(-1, 1), 1.0)
    return torch.where(total > 0, pos / total.clamp_min(1e-38), fallback)


def backward_sweep(levels):
    """The return path of `_traverse_game_tree` (deep_cfr.py:468-480, 492-497) over a level-ordered tree. Each level is a
    dict with `terminal` [m] bool, `ret_p` [m] float64 (the traverser's return where terminal), `nt` (indices of the
    non-terminal nodes) and -- when it has `children` -- per non-terminal node `strategy` [k, 18], `legal` [k, 18] bool,
    `is_trav` [k] bool, `trav` (indices into the non-terminal list of the traverser's nodes) and the child lists
    `local` / `action` (child j of the level is node j of the next level). Sets `value` [m] on every level: the
    return at terminal nodes, the sampled child's value at the opponent's nodes, sum_a strategy[a] * payoff[a] at the
    traverser's (unsampled actions count as payoff 0, as in the reference). Yields (level, sampled regrets [t, 18]
    float32 of its traverser nodes: payoff[a] - value on the legal actions) from the deepest level up."""
    child_values = None
    for lvl in reversed(levels):
        value = torch.where(lvl["terminal"], lvl["ret_p"], torch.zeros_like(lvl["ret_p"]))
        regret = None
        if lvl["children"]:
            payoff = torch.zeros(lvl["legal"].shape, dtype=torch.float64, device=value.device)
            payoff[lvl["local"], lvl["action"]] = child_values
            cfv = (lvl["strategy"].double() * payoff * lvl["legal"]).sum(-1)
            value[lvl["nt"]] = torch.where(lvl["is_trav"], cfv, payoff.sum(-1))
            trav = lvl["trav"]
            if trav.numel():
                regret = ((payoff[trav] - cfv[trav].view(-1, 1)) * lvl["legal"][trav]).float()
        lvl["value"] = value
        child_values = value
        if regret is not None:
            yield lvl, regret


class MLP(nn.Module):
    """`simple_nets.MLP` (simple_nets.py:84-120): Linear+ReLU hidden layers, linear head, weights drawn from a
    normal truncated at two standard deviations with stddev 1/sqrt(fan_in), zero biases (simple_nets.py:44-52)."""

    def __init__(self, input_size, hidden_sizes, output_size):
        super().__init__()
        layers, prev = [], input_size
        for h in hidden_sizes:
            layers += [nn.Linear(prev, h), nn.ReLU()]
            prev = h
        layers.append(nn.Linear(prev, output_size))
        self.net = nn.Sequential(*layers)
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.net:
            if isinstance(m, nn.Linear):
                std = 1.0 / math.sqrt(m.in_features)
                nn.init.trunc_normal_(m.weight, mean=0.0, std=std, a=-2 * std, b=2 * std)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        return self.net(x)


class ReservoirBuffer:
    """`deep_cfr.ReservoirBuffer` (deep_cfr.py:43-99) on the device and batched: element number t (0-based count
    of everything ever added) fills slot t while t < capacity and afterwards replaces slot randint(0, t) when that
    is < capacity; inside one batch the later element wins a slot, as it would sequentially. Info-state rows are
    kept as uint8 (every value of the tensor is an integer in 0..12, coup.cc:150-287): 2.5 KB per record."""

    def __init__(self, capacity, device, fields, seed=0):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self._fields = dict(fields)
        self._data = {}          # allocated on first add, grown geometrically up to `capacity`
        self._allocated = 0
        self.add_calls = 0
        self.size = 0
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed)

    def _ensure(self, rows):
        if rows <= self._allocated:
            return
        new = min(self.capacity, max(rows, 2 * self._allocated, 1024))
        for name, (shape, dtype) in self._fields.items():
            buf = torch.zeros((new, *shape), dtype=dtype, device=self.device)
            if self._allocated:
                buf[: self._allocated] = self._data[name]
            self._data[name] = buf
        self._allocated = new

    def add(self, **columns):
        b = int(next(iter(columns.values())).shape[0])
        if b == 0:
            return
        self._ensure(min(self.capacity, self.add_calls + b))
        if self.add_calls + b <= self.capacity:        # still filling: element t goes to slot t
            for name, (_, dtype) in self._fields.items():
                self._data[name][self.add_calls: self.add_calls + b] = columns[name].to(dtype)
            self.add_calls += b
            self.size = self.add_calls
            return
        t = self.add_calls + torch.arange(b, device=self.device)
        draw = (torch.rand(b, device=self.device, generator=self._gen, dtype=torch.float64) * (t + 1).double()).long()
        slot = torch.where(t < self.capacity, t, draw)
        keep = slot < self.capacity
        slot, src = slot[keep], torch.arange(b, device=self.device)[keep]
        winner = torch.full((self.capacity,), -1, dtype=torch.long, device=self.device)
        winner.scatter_reduce_(0, slot, torch.arange(src.numel(), device=self.device), reduce="amax", include_self=True)
        chosen = winner[winner >= 0]
        dst, src = slot[chosen], src[chosen]
        for name, (_, dtype) in self._fields.items():
            self._data[name][dst] = columns[name][src].to(dtype)
        self.add_calls += b
        self.size = min(self.capacity, self.add_calls)

    def sample(self, num_samples):
        """`ReservoirBuffer.sample` (deep_cfr.py:78-92): `num_samples` distinct records, uniformly."""
        if num_samples > self.size:
            raise ValueError("{} elements could not be sampled from size {}".format(num_samples, self.size))
        j = torch.randperm(self.size, device=self.device, generator=self._gen)[:num_samples]
        return {name: buf[j] for name, buf in self._data.items()}

    def all(self):
        return {name: buf[: self.size] for name, buf in self._data.items()}

    def clear(self):
        self.add_calls = 0
        self.size = 0

    def __len__(self):
        return self.size


_ADV_FIELDS = {"info_state": ((INFO_STATE_SIZE,), torch.uint8), "iteration": ((), torch.int32),
               "advantage": ((NUM_DISTINCT_ACTIONS,), torch.float32), "action": ((), torch.uint8)}
_STRAT_FIELDS = {"info_state": ((INFO_STATE_SIZE,), torch.uint8), "iteration": ((), torch.int32),
                 "strategy_action_probs": ((NUM_DISTINCT_ACTIONS,), torch.float32)}


@dataclasses.dataclass
class DeepCFRConfig:
    """The keyword arguments of the reference's `DeepCFRSolver.__init__` (deep_cfr.py:118-188), same names and defaults
    (the thesis runs set them from `coup_experiments/scripts/flags/thesis_runs/deep_cfr-*.cfg`)."""
    policy_network_layers: tuple = (256, 256)
    advantage_network_layers: tuple = (128, 128)
    num_iterations: int = 100
    num_traversals: int = 20
    learning_rate: float = 1e-4
    batch_size_advantage: int = None
    batch_size_strategy: int = None
    memory_capacity: int = int(1e6)
    policy_network_train_steps: int = 1
    advantage_network_train_steps: int = 1
    reinitialize_advantage_networks: bool = True
    sampling_method: str = "external"
    outcome_samp_expl: float = 0.6
    outcome_factor: int = 1
    e_outcome: float = 0
    iter_net_train: bool = False
    adv_net_reinit_every: int = 1
    eval_func: object = None
    eval_every: int = 10
    eval_train_episodes: int = 10000
    eval_test_every: int = 5000
    eval_test_episodes: int = 1000
    use_checkpoints: bool = False
    checkpoint_dir: str = None
    save_every: int = None

    def __post_init__(self):
        if self.sampling_method not in ("external", "outcome", "e-outcome"):
            raise ValueError(f"Unknown sampling method '{self.sampling_method}'.")           # deep_cfr.py:229-230
        if self.outcome_factor <= 0:
            raise ValueError("Outcome factor must be greater than 0.")                      # deep_cfr.py:233-234


class DeepCFRSolver:
    """`deep_cfr.DeepCFRSolver` (deep_cfr.py:102-640) for game "coup". Arguments as in the reference (:118-188);
    additions: `device`, `seed`, `max_nodes` (size of each of the two scratch env slabs; a level with more nodes is
    expanded in slab-sized pieces, so the width of a tree is limited by HBM at ~200 B per node, not by the slabs),
    `fused_expand` (regret matching and child selection of a level in one CUDA kernel, `coup_cfr_expand`, instead of
    ~35 PyTorch ops), `roots_per_batch` (how many of the `num_traversals` roots are expanded together), `max_tree_nodes` (a batch whose
    trees grow past this many nodes raises instead of exhausting memory) and `record_tree` (keep the
    per-level bookkeeping of the last batch in `last_tree` for inspection and tests)."""

    def __init__(self, game=None, device=0, seed=0, max_nodes=1 << 18, roots_per_batch=None, max_tree_nodes=1 << 26,
                 record_tree=False, fused_expand=True, device_levels=True, **reference_args):
        self.cfg = cfg = DeepCFRConfig(**reference_args)          # the reference's keyword arguments, validated
        self._game = game
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self._num_players, self._num_actions, self._embedding_size = 2, NUM_DISTINCT_ACTIONS, INFO_STATE_SIZE
        self._iteration = 1
        self._environment_steps = 0
        self._record_tree = record_tree
        self._fused_expand = fused_expand       # regret matching + child selection in one CUDA kernel (coup_cfr_expand)
        # Outcome-sampling trees stay narrow: their levels are expanded by cfr_traversal.DeviceTreeTraverser, which keeps the
        # level sizes on the device (no host read-back per level). External sampling (levels wider than a slab) and
        # record_tree use the host-driven engine below.
        self._device_levels = bool(device_levels) and cfg.sampling_method != "external" and not record_tree
        self._traverser = None
        self._expand_seed, self._expand_counter = seed * 2654435761 % (1 << 63) + 17, 0
        self.last_tree = None
        policy_network_layers, advantage_network_layers = cfg.policy_network_layers, cfg.advantage_network_layers
        learning_rate, memory_capacity, num_traversals = cfg.learning_rate, cfg.memory_capacity, cfg.num_traversals

        torch.manual_seed(seed)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed)
        self._policy_network = MLP(self._embedding_size, list(policy_network_layers), self._num_actions).to(self.device)
        self._optimizer_policy = torch.optim.Adam(self._policy_network.parameters(), lr=learning_rate)
        self._strategy_memories = ReservoirBuffer(memory_capacity, self.device, _STRAT_FIELDS, seed=seed + 1)
        self._advantage_memories = [ReservoirBuffer(memory_capacity, self.device, _ADV_FIELDS, seed=seed + 2 + p)
                                    for p in range(self._num_players)]
        self._advantage_networks = [MLP(self._embedding_size, list(advantage_network_layers), self._num_actions).to(self.device)
                                    for _ in range(self._num_players)]
        # one optimiser per network for the whole run: re-initialising a network keeps its Adam moments, as
        # re-running only the variable initialisers does in the reference (deep_cfr.py:343-349)
        self._optimizer_advantages = [torch.optim.Adam(net.parameters(), lr=learning_rate) for net in self._advantage_networks]

        self._max_nodes = max(int(max_nodes), 7)
        self._max_tree_nodes = int(max_tree_nodes)
        self._roots_per_batch = int(roots_per_batch or min(num_traversals, self._max_nodes))
        dev_index = self.device.index or 0
        self._slabs = [CoupVectorEnv(self._max_nodes, seed=seed + 101 + i, device=dev_index, auto_reset=False)
                       for i in range(2)]
        self._arange = torch.arange(self._max_nodes, device=self.device)

    # ---- the reference's accessors -------------------------------------------------------------------
    @property
    def advantage_buffers(self):
        return self._advantage_memories

    @property
    def strategy_buffer(self):
        return self._strategy_memories

    @property
    def policy_network(self):
        return self._policy_network

    @property
    def advantage_networks(self):
        return self._advantage_networks

    @property
    def iteration(self):
        return self._iteration

    def clear_advantage_buffers(self):
        for m in self._advantage_memories:
            m.clear()

    def reinitialize_advantage_networks(self):
        for p in range(self._num_players):
            self.reinitialize_advantage_network(p)

    def reinitialize_advantage_network(self, player):
        self._advantage_networks[player].reset_parameters()

    def _reinitialize_policy_network(self):
        self._policy_network.reset_parameters()

    def get_environment_steps(self):
        return self._environment_steps

    def save_policy_network(self, checkpoint_dir, checkpoint_id):
        os.makedirs(checkpoint_dir, exist_ok=True)
        path = os.path.join(checkpoint_dir, "policy_network" + checkpoint_id + ".pt")
        torch.save(self._policy_network.state_dict(), path)
        return path

    def has_checkpoint(self, checkpoint_dir, checkpoint_id):
        return os.path.exists(os.path.join(checkpoint_dir, "policy_network" + checkpoint_id + ".pt"))

    def restore_policy_network(self, checkpoint_dir, checkpoint_id):
        path = os.path.join(checkpoint_dir, "policy_network" + checkpoint_id + ".pt")
        self._policy_network.load_state_dict(torch.load(path, map_location=self.device))

    # ---- solve: the schedule of deep_cfr.py:360-410 ----------------------------------------------------
    def _advantage_phase(self, losses):
        """Traversals for both players, each followed by a (possibly re-initialised) fit of that player's network."""
        cfg = self.cfg
        for player in range(self._num_players):
            self.traverse(player, cfg.num_traversals)
            if cfg.reinitialize_advantage_networks and self._iteration % cfg.adv_net_reinit_every == 0:
                self.reinitialize_advantage_network(player)
            losses[player].append(self._learn_advantage_network(player))

    def _evaluate(self):
        cfg = self.cfg
        cfg.eval_func(exploitee=self, num_train_episodes=cfg.eval_train_episodes, eval_every=cfg.eval_test_every,
                      eval_episodes=cfg.eval_test_episodes)

    def _checkpoint(self, iteration):
        self.save_policy_network(self.cfg.checkpoint_dir, f"iter{iteration}")

    def solve(self):
        """Returns (policy network, advantage losses per player, last policy loss). Per iteration: the advantage phase;
        a policy fit when `iter_net_train`; on evaluation iterations (every `eval_every`, never the last) the policy is
        fitted if it was not, evaluated, checkpointed every `save_every`, and -- if it is only trained on demand --
        discarded again. After the last iteration: final fit, evaluation, checkpoint."""
        cfg = self.cfg
        losses = collections.defaultdict(list)
        policy_loss = None
        for _ in range(cfg.num_iterations):
            it = self._iteration
            self._advantage_phase(losses)
            trained_now = cfg.iter_net_train
            if trained_now:
                policy_loss = self._learn_strategy_network()
            evaluating = cfg.eval_func is not None and it % cfg.eval_every == 0 and it != cfg.num_iterations
            if evaluating:
                if not trained_now:
                    policy_loss = self._learn_strategy_network()
                self._evaluate()
                if cfg.use_checkpoints and cfg.save_every and it % cfg.save_every == 0:
                    self._checkpoint(it)
                if not trained_now:
                    self._reinitialize_policy_network()
            self._iteration = it + 1
        policy_loss = self._learn_strategy_network()
        if cfg.eval_func is not None:
            self._evaluate()
        if cfg.use_checkpoints:
            self._checkpoint(self._iteration - 1)
        return self._policy_network, losses, policy_loss

    # ---- traversal -----------------------------------------------------------------------------------
    def traverse(self, player, num_traversals):
        """`num_traversals` calls of `_traverse_game_tree(root, player)` (deep_cfr.py:363-365). Returns the mean
        root value (the reference discards it) and the number of nodes expanded."""
        done, total_value, total_nodes = 0, 0.0, 0
        while done < num_traversals:
            r = min(self._roots_per_batch, num_traversals - done)
            values, nodes = (self._traverse_batch_device if self._device_levels else self._traverse_batch)(player, r)
            total_value += float(values.sum())
            total_nodes += nodes
            done += r
        return total_value / max(num_traversals, 1), total_nodes

    @torch.no_grad()
    def _advantages_both(self, rows_u8, cur_player):
        """Both networks on every row, then select: no data-dependent shapes, so nothing waits for the host."""
        x = rows_u8.float()
        return torch.where((cur_player == 0).view(-1, 1), self._advantage_networks[0](x), self._advantage_networks[1](x))

    def _traverse_batch_device(self, player, num_roots):
        """One batch of traversals with the level sizes kept on the device (cfr_traversal.DeviceTreeTraverser)."""
        from .cfr_traversal import DeviceTreeTraverser
        cfg = self.cfg
        if self._traverser is None:
            self._traverser = DeviceTreeTraverser(self._max_nodes, self._advantages_both, device=self.device.index or 0,
                                                  seed=self._expand_seed % (1 << 31), sampling_method=cfg.sampling_method,
                                                  outcome_factor=cfg.outcome_factor, e_outcome=cfg.e_outcome,
                                                  outcome_samp_expl=cfg.outcome_samp_expl)
        tr = self._traverser
        chance_before = [int(s.stats_device[1]) for s in tr.slabs]
        res = tr.traverse(player, num_roots)
        adv, strat = tr.memory_records(res, player)
        it = lambda k: torch.full((k,), self._iteration, device=self.device)
        self._advantage_memories[player].add(info_state=adv["info_state"], advantage=adv["advantage"], action=adv["action"],
                                             iteration=it(adv["action"].numel()))
        self._strategy_memories.add(info_state=strat["info_state"], strategy_action_probs=strat["strategy_action_probs"],
                                    iteration=it(strat["info_state"].shape[0]))
        chance = sum(int(s.stats_device[1]) - b for s, b in zip(tr.slabs, chance_before))
        self._environment_steps += res["nodes"] + chance
        self.last_level_widths = res["sizes"]
        return res["root_values"], res["nodes"]

    @torch.no_grad()
    def _advantages(self, rows_u8, cur_player):
        x = rows_u8.float()
        if x.shape[0] <= 8192:      # small levels are launch-bound: both networks on every row, then select
            return torch.where((cur_player == 0).view(-1, 1), self._advantage_networks[0](x), self._advantage_networks[1](x))
        out = torch.empty((x.shape[0], self._num_actions), dtype=torch.float32, device=self.device)
        for p in range(self._num_players):
            idx = (cur_player == p).nonzero(as_tuple=True)[0]
            if idx.numel():
                out[idx] = self._advantage_networks[p](x[idx])
        return out

    def _load(self, slab, state, history):
        k = state.shape[0]
        slab.state[:k] = state
        slab.history[:k] = history
        return k

    def _encode_saved(self, state, history):
        """Info-state rows (uint8, player to move) of packed states kept since the forward pass."""
        slab = self._slabs[0]
        out = torch.empty((state.shape[0], INFO_STATE_SIZE), dtype=torch.uint8, device=self.device)
        for a in range(0, state.shape[0], self._max_nodes):
            k = self._load(slab, state[a: a + self._max_nodes], history[a: a + self._max_nodes])
            slab.information_state_tensor_gather(self._arange[:k], player=PLAYER_CURRENT, out=out[a: a + k])
        return out

    def _children_of_traverser(self, strategy, legal):
        """bool [m, 18]: which actions of each traverser node are expanded (deep_cfr.py:438-466)."""
        if self.cfg.sampling_method == "external":
            return legal
        m = legal.shape[0]
        n_legal = legal.sum(-1)
        if self.cfg.sampling_method == "e-outcome":
            multi = torch.rand(m, device=self.device, generator=self._gen) < self.cfg.e_outcome
            factor = torch.where(multi, self.cfg.outcome_factor, 1)
        else:
            factor = torch.full((m,), self.cfg.outcome_factor, device=self.device)
        num_to_sample = torch.minimum(n_legal, factor)
        uniform = legal / n_legal.clamp_min(1).view(-1, 1)
        probs = self.cfg.outcome_samp_expl * uniform + (1.0 - self.cfg.outcome_samp_expl) * strategy
        probs = probs / probs.sum(-1, keepdim=True)
        # np.random.choice(size=k, replace=False, p=probs) draws sequentially without replacement, which is the
        # Plackett-Luce order of the Gumbel-perturbed log-probabilities: take the k largest keys
        u = torch.rand(probs.shape, device=self.device, generator=self._gen, dtype=torch.float64).clamp_min(1e-300)
        keys = torch.where(legal & (probs > 0), probs.double().log() - (-u.log()).log(), torch.full_like(u, -math.inf))
        rank = keys.argsort(-1, descending=True).argsort(-1)
        return legal & (rank < num_to_sample.view(-1, 1))

    def _select_children(self, player, adv, word):
        """Fused path (coup_cfr_expand + coup_cfr_children): strategy [k, 18] and the (parent, action) lists of the
        children, in parent order, from the advantages and step words of k nodes."""
        lib, k, dev = self._slabs[0]._lib, int(word.numel()), self.device
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        adv = adv.contiguous()
        word = word.contiguous()
        strategy = torch.empty((k, NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=dev)
        expand = torch.empty(k, dtype=torch.int32, device=dev)
        counts = torch.empty(k, dtype=torch.int32, device=dev)
        e_outcome = float(self.cfg.e_outcome) if self.cfg.sampling_method == "e-outcome" else -1.0
        check(lib.coup_cfr_expand(ptr(adv), ptr(word), k, player, int(self.cfg.sampling_method == "external"),
                                  int(self.cfg.outcome_factor), e_outcome, float(self.cfg.outcome_samp_expl), self._expand_seed,
                                  self._expand_counter, ptr(strategy), ptr(expand), ptr(counts), stream))
        self._expand_counter += 1
        ends = torch.cumsum(counts, 0, dtype=torch.int64)
        c = int(ends[-1])
        parent = torch.empty(c, dtype=torch.int32, device=dev)
        action = torch.empty(c, dtype=torch.uint8, device=dev)
        check(lib.coup_cfr_children(ptr(expand), ptr(ends - counts), k, ptr(parent), ptr(action), stream))
        return strategy, parent.long(), action.long()

    def _expand_chunk(self, player, state, history, word):
        """One slab-full of non-terminal nodes of a level: strategies, which children to expand, and the children
        themselves (packed state, history, step word), in parent order."""
        src, dst = self._slabs
        k = self._load(src, state, history)
        legal = _legal_bool(word & 0x3FFFF)
        cp = (word >> 18) & 1
        rows = src.information_state_tensor_gather(self._arange[:k], player=PLAYER_CURRENT, dtype=torch.uint8)
        adv = self._advantages(rows, cp)
        is_trav = cp == player
        if self._fused_expand:
            strategy, local, action = self._select_children(player, adv, word)
        else:
            strategy = regret_matching(adv, legal)
            # opponent nodes: one action from the matched regrets, renormalised (:482-492)
            probs = strategy / strategy.sum(-1, keepdim=True)
            sampled = torch.multinomial(probs, 1, generator=self._gen).view(-1)
            expand = torch.zeros_like(legal).scatter_(1, sampled.view(-1, 1), True)
            trav = is_trav.nonzero(as_tuple=True)[0]
            if trav.numel():
                expand[trav] = self._children_of_traverser(strategy[trav], legal[trav])
            local, action = expand.nonzero(as_tuple=True)          # row-major: children grouped by parent
        opp = (~is_trav).nonzero(as_tuple=True)[0]                 # a StrategyMemory record per opponent node
        if opp.numel():
            self._strategy_memories.add(info_state=rows[opp], strategy_action_probs=strategy[opp],
                                        iteration=torch.full((opp.numel(),), self._iteration, device=self.device))
        c = dst.fork_from(src, local, action.to(torch.uint8))
        # the reference's record keeps the loop variable `action` of its last `for`, i.e. the largest legal id (:479)
        last_legal = (NUM_DISTINCT_ACTIONS - 1) - legal.flip(-1).to(torch.int32).argmax(-1)
        info = {"local": local, "action": action, "strategy": strategy, "legal": legal, "is_trav": is_trav,
                "last_legal": last_legal}
        return info, dst.state[:c].clone(), dst.history[:c].clone(), dst.step_word[:c].clone()

    @torch.no_grad()
    def _traverse_batch(self, player, num_roots, roots=None):
        """One batch of traversals for `player`. `roots`: None for `num_roots` freshly dealt games (the reference's
        `_root_node` followed by its four chance nodes), or the (state, history, step_word) tensors of a
        CoupVectorEnv to start from arbitrary decision nodes."""
        stats_before = [int(s.stats_device[1]) for s in self._slabs]
        if roots is None:
            if num_roots > self._max_nodes:
                raise ValueError("roots_per_batch exceeds max_nodes")
            root = self._slabs[0]
            root.reset(envs_to_reset=(self._arange < num_roots).to(torch.uint8))     # root + its 4 deals (:430-434)
            roots = root.state[:num_roots], root.history[:num_roots], root.step_word[:num_roots]
        state, history, word = (t.clone() for t in roots)
        levels = []
        sign = 1.0 if player == 0 else -1.0
        chunk = self._max_nodes // 7            # a node has at most 7 legal actions (coup.cc:838-872), so the children fit
        total = 0
        while state.shape[0] > 0:
            m = state.shape[0]
            total += m
            if total > self._max_tree_nodes:
                # multi-outcome and external sampling grow exponentially with the length of a Coup game (5-7 actions
                # per turn, up to 91 moves): stop before the bookkeeping (~200 B per node) exhausts HBM
                if self._record_tree:
                    self.last_tree = levels          # forward part only (no values): what was expanded so far
                raise RuntimeError(f"traversal batch grew past max_tree_nodes={self._max_tree_nodes} at level "
                                   f"{len(levels)} ({m} nodes wide): use fewer roots per batch or a smaller outcome_factor")
            terminal = ((word >> 19) & 1).bool()
            ret_p = sign * (((word >> 24) & 7) - 2).double()
            nt = (~terminal).nonzero(as_tuple=True)[0]
            lvl = {"m": m, "terminal": terminal, "ret_p": ret_p, "nt": nt, "children": 0}
            if self._record_tree:
                lvl.update(history=history, moves=(state[:, 3] & 127).to(torch.int64), word=word)
            levels.append(lvl)
            if nt.numel() == 0:
                break
            infos, kids, offset = [], [], 0
            for a in range(0, nt.numel(), chunk):
                idx = nt[a: a + chunk]
                info, cs, ch, cw = self._expand_chunk(player, state[idx], history[idx], word[idx])
                info["local"] = info["local"] + a
                infos.append(info)
                kids.append((cs, ch, cw))
            lvl.update({k: torch.cat([i[k] for i in infos]) for k in infos[0]})
            trav = lvl["is_trav"].nonzero(as_tuple=True)[0]
            # traverser nodes wait for their regrets until the backward sweep: keep the packed state + history (80 B)
            # rather than the 2.5 KB row and encode again then
            lvl.update(trav=trav, trav_state=state[nt[trav]], trav_history=history[nt[trav]],
                       children=int(lvl["local"].numel()))
            state, history, word = (torch.cat([k[i] for k in kids]) for i in range(3))
        # ---- backward sweep (:468-480) ----
        for lvl, regret in backward_sweep(levels):
            trav = lvl["trav"]
            rows = self._encode_saved(lvl["trav_state"], lvl["trav_history"])
            self._advantage_memories[player].add(
                info_state=rows, advantage=regret, action=lvl["last_legal"][trav],
                iteration=torch.full((trav.numel(),), self._iteration, device=self.device))
            if self._record_tree:
                lvl["regret"] = regret
                if trav.numel() <= 4096:
                    lvl["rows"] = rows
        nodes = sum(l["m"] for l in levels)
        chance = sum(int(s.stats_device[1]) - b for s, b in zip(self._slabs, stats_before))
        self._environment_steps += nodes + chance        # every recursive call, chance nodes included (:427)
        self.last_level_widths = [l["m"] for l in levels]
        if self._record_tree:
            self.last_tree = levels
        return levels[0]["value"], nodes

    # ---- learning (deep_cfr.py:549-616) -----------------------------------------------------------------
    def _learn_advantage_network(self, player):
        memory, net, opt = self._advantage_memories[player], self._advantage_networks[player], self._optimizer_advantages[player]
        loss = None
        for _ in range(self.cfg.advantage_network_train_steps):
            if self.cfg.batch_size_advantage:
                if self.cfg.batch_size_advantage > len(memory):
                    return None                                                     # not enough samples (:562-564)
                batch = memory.sample(self.cfg.batch_size_advantage)
            else:
                batch = memory.all()
            if batch["info_state"].shape[0] == 0:
                return None
            w = batch["iteration"].float().sqrt().view(-1, 1)
            pred = net(batch["info_state"].float())
            loss = torch.mean((w * batch["advantage"] - w * pred) ** 2)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        return None if loss is None else float(loss.detach())

    def _learn_strategy_network(self):
        memory = self._strategy_memories
        loss = None
        for _ in range(self.cfg.policy_network_train_steps):
            if self.cfg.batch_size_strategy:
                if self.cfg.batch_size_strategy > len(memory):
                    return None
                batch = memory.sample(self.cfg.batch_size_strategy)
            else:
                batch = memory.all()
            if batch["info_state"].shape[0] == 0:
                return None
            w = batch["iteration"].float().sqrt().view(-1, 1)
            probs = torch.softmax(self._policy_network(batch["info_state"].float()), dim=-1)
            loss = torch.mean((w * batch["strategy_action_probs"] - w * probs) ** 2)
            self._optimizer_policy.zero_grad(set_to_none=True)
            loss.backward()
            self._optimizer_policy.step()
        return None if loss is None else float(loss.detach())

    # ---- acting (deep_cfr.py:527-547) -------------------------------------------------------------------
    @torch.no_grad()
    def action_probs(self, info_state, legal_bits):
        """Batched `action_probabilities`: softmax of the policy network, illegal actions removed, renormalised."""
        probs = torch.softmax(self._policy_network(info_state.float()), dim=-1) * _legal_bool(legal_bits)
        return probs / probs.sum(-1, keepdim=True)

    def action_probabilities(self, state, player_id=None):
        """Single-state form for a `spiel.CoupState`: {action: probability} over the legal actions."""
        cur_player = state.current_player()
        legal_actions = state.legal_actions(cur_player)
        info = torch.tensor(state.information_state_tensor(), dtype=torch.float32, device=self.device).view(1, -1)
        bits = torch.tensor([sum(1 << a for a in legal_actions)], dtype=torch.int32, device=self.device)
        probs = self.action_probs(info, bits)[0].tolist()
        return {a: probs[a] for a in legal_actions}
