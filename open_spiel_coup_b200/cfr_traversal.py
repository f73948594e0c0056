"""Sampled CFR traversals (the tree walk of `open_spiel/python/algorithms/deep_cfr.py:415-497`) expanded level by level
WITHOUT the host in the loop.

`DeepCFRSolver._traverse_batch` reads sizes back to the host several times per level (children of a level, traverser /
opponent nodes, terminal nodes) and runs ~60 launches per level. Here the number of nodes of a level stays in device
memory: every kernel is launched for the capacity of the level buffers and works on the first `*count` nodes
(include/coup_b200.h, "the same traversal level WITHOUT the host in the loop"):

  per level   coup_vec_pack_records                      keep the nodes as 96-byte packed records
              coup_vec_information_state_tensor_prefix   uint8 rows of the player to move
              advantage networks (PyTorch, both on every row)
              coup_cfr_level                             regret matching, child selection, prefix sum, child lists, next count
              coup_vec_fork_counted                      all children into the other slab, chance nodes resolved
  afterwards  coup_cfr_backward, one launch per level    values and sampled regrets
              one host read of the level sizes; memory records are decoded from the packed records of the levels

The host never waits for the level it has just queued: the size of level l is copied to pinned memory as soon as it
exists and is looked at one level later, when it has long arrived. It bounds the next level (a node has at most
`outcome_factor` children, or 7 with external sampling), which is all the host needs to size the network's GEMMs, and an
empty level ends the walk.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import INFO_STATE_SIZE, NUM_DISTINCT_ACTIONS, PLAYER_CURRENT, RECORD_WORDS, check
from .vector_env import CoupVectorEnv


class DeviceTreeTraverser:
    """Expands `num_roots` freshly dealt games (or given roots) for `player` with the sampling rule of the reference
    (`external`, `outcome` or `e-outcome`). `advantages(rows_u8 [m, 2492], cur_player [m]) -> float [m, 18]` supplies the
    advantage-network outputs of the player to move. A level may hold at most `capacity` nodes; a wider level raises."""

    def __init__(self, capacity, advantages, device=0, seed=0, sampling_method="outcome", outcome_factor=1, e_outcome=0.0,
                 outcome_samp_expl=0.6, max_levels=96):
        if sampling_method not in ("external", "outcome", "e-outcome"):
            raise ValueError(f"Unknown sampling method '{sampling_method}'.")
        self.capacity, self.advantages = int(capacity), advantages
        self.method, self.factor, self.e_outcome, self.expl = sampling_method, int(outcome_factor), float(e_outcome), float(outcome_samp_expl)
        self.max_levels = int(max_levels)
        self.slabs = [CoupVectorEnv(self.capacity, seed=seed + 101 + i, device=device, auto_reset=False) for i in range(2)]
        self.device = self.slabs[0].device
        self._lib = self.slabs[0]._lib
        self._seed, self._counter = seed * 2654435761 % (1 << 63) + 17, 0
        w, dev = self.capacity, self.device
        self.counts = torch.zeros(self.max_levels + 2, dtype=torch.int32, device=dev)
        self.counts_host = torch.zeros(self.max_levels + 2, dtype=torch.int32).pin_memory()
        self._count_ready = [torch.cuda.Event() for _ in range(self.max_levels + 2)]
        # children per node: outcome_factor at the traverser's nodes (7 = every legal action with external sampling), 1 at the opponent's
        self._growth = 7 if sampling_method == "external" else max(1, int(outcome_factor))
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.parent = torch.empty(w, dtype=torch.int32, device=dev)
        self.action = torch.empty(w, dtype=torch.uint8, device=dev)
        self.rows = torch.empty((w, INFO_STATE_SIZE), dtype=torch.uint8, device=dev)
        self._adv = torch.zeros((w, NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=dev)
        self._levels = []          # per-level buffers, allocated on first use and kept
        self.launches_per_level = 6   # ours (pack, copy, encode, level, count copy, fork); the networks' GEMMs are PyTorch's

    def _level(self, l):
        while len(self._levels) <= l:
            w, dev = self.capacity, self.device
            self._levels.append({
                "records": torch.empty((w, RECORD_WORDS), dtype=torch.int32, device=dev),
                "words": torch.empty(w, dtype=torch.int32, device=dev),
                "strategy": torch.empty((w, NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=dev),
                "expand": torch.empty(w, dtype=torch.int32, device=dev),
                "offset": torch.empty(w, dtype=torch.int32, device=dev),
                "value": torch.zeros(w, dtype=torch.float64, device=dev),
                "regret": torch.empty((w, NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=dev),
            })
        return self._levels[l]

    @torch.no_grad()
    def traverse(self, player, num_roots, roots=None):
        """Returns a dict: `sizes` (nodes per level, host list), `levels` (per-level device buffers: records, words,
        strategy, expand, offset, value, regret -- the first sizes[l] entries are valid), `nodes`, `root_values`."""
        lib, w, dev = self._lib, self.capacity, self.device
        ptr = lambda t: C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if num_roots > w:
            raise ValueError("more roots than the level capacity")
        cur, nxt = self.slabs
        if roots is None:
            cur.reset(envs_to_reset=(torch.arange(w, device=dev) < num_roots).to(torch.uint8))   # root + its 4 deals (:430-434)
        else:
            state, history, word = roots
            cur.state[:num_roots], cur.history[:num_roots], cur.step_word[:num_roots] = state, history, word
        self.counts.zero_()
        self.counts[0] = num_roots
        self.overflow.zero_()
        e_outcome = self.e_outcome if self.method == "e-outcome" else -1.0
        depth = 0
        bound = num_roots                      # upper bound of the size of the level about to be expanded
        self.counts_host[0] = num_roots
        for l in range(self.max_levels):
            if l >= 2:                         # the size of level l - 1 was queued for copy a whole level ago
                self._count_ready[l - 1].synchronize()
                prev = int(self.counts_host[l - 1])
                if prev == 0:
                    break                      # level l - 1 was already empty: nothing to expand
                bound = min(w, prev * self._growth)
            elif l == 1:
                bound = min(w, num_roots * self._growth)
            rows_now = min(w, -(-bound // 8) * 8)
            lvl, cnt = self._level(l), self.counts[l:l + 1]
            check(lib.coup_vec_pack_records(cur._h, ptr(cnt), ptr(lvl["records"]), stream))
            lvl["words"].copy_(cur.step_word)
            check(lib.coup_vec_information_state_tensor_prefix(cur._h, ptr(cnt), rows_now, PLAYER_CURRENT, _lib.DTYPE_U8, ptr(self.rows),
                                                                INFO_STATE_SIZE, stream))
            adv = self._adv
            adv[:rows_now] = self.advantages(self.rows[:rows_now], (lvl["words"][:rows_now] >> 18) & 1)
            check(lib.coup_cfr_level(ptr(adv), ptr(lvl["words"]), ptr(cnt), w, player, int(self.method == "external"), self.factor,
                                     e_outcome, self.expl, self._seed, self._counter, ptr(lvl["strategy"]), ptr(lvl["expand"]),
                                     ptr(lvl["offset"]), ptr(self.parent), ptr(self.action), ptr(self.counts[l + 1:l + 2]),
                                     ptr(self.overflow), stream))
            self._counter += 1
            self.counts_host[l + 1:l + 2].copy_(self.counts[l + 1:l + 2], non_blocking=True)
            self._count_ready[l + 1].record()
            check(lib.coup_vec_fork_counted(nxt._h, cur._h, ptr(self.parent), ptr(self.action), ptr(self.counts[l + 1:l + 2]),
                                            min(w, bound * self._growth), stream))
            cur, nxt = nxt, cur
            depth = l + 1
        else:
            torch.cuda.synchronize(dev)
            if int(self.counts[self.max_levels]) != 0:
                raise RuntimeError("traversal deeper than max_levels")
        # ---- backward sweep (:468-480): values of level l from the values of level l + 1 ----
        self._level(depth)["value"].zero_()
        for l in range(depth - 1, -1, -1):
            lvl, child = self._level(l), self._level(l + 1)
            check(lib.coup_cfr_backward(ptr(lvl["words"]), ptr(self.counts[l:l + 1]), w, player, ptr(lvl["strategy"]), ptr(lvl["expand"]),
                                        ptr(lvl["offset"]), ptr(child["value"]), ptr(lvl["value"]), ptr(lvl["regret"]), stream))
        sizes = self.counts[:depth + 1].cpu().tolist()
        if int(self.overflow.item()):
            raise RuntimeError(f"a level grew past the capacity of {w} nodes: use fewer roots per batch, a smaller "
                               "outcome_factor or a larger capacity")
        while sizes and sizes[-1] == 0:
            sizes.pop()
        return {"sizes": sizes, "levels": self._levels[:len(sizes)], "nodes": sum(sizes),
                "root_values": self._levels[0]["value"][:num_roots]}

    @torch.no_grad()
    def memory_records(self, result, player):
        """The records the reference appends during the walk, for the whole batch at once: for the traverser's nodes
        `AdvantageMemory(info_state, ., sampled regrets, action)` (:476-480; `action` is the largest legal id, the loop
        variable the reference leaves behind) and for the opponent's `StrategyMemory(info_state, ., strategy)` (:488-491).
        Returns two dicts of device tensors; info states are uint8 rows decoded from the packed records."""
        sizes, levels = result["sizes"], result["levels"]
        words = torch.cat([lv["words"][:m] for lv, m in zip(levels, sizes)])
        records = torch.cat([lv["records"][:m] for lv, m in zip(levels, sizes)])
        strategy = torch.cat([lv["strategy"][:m] for lv, m in zip(levels, sizes)])
        regret = torch.cat([lv["regret"][:m] for lv, m in zip(levels, sizes)])
        live = ((words >> 19) & 1) == 0
        is_trav = ((words >> 18) & 1) == player
        trav = (live & is_trav).nonzero(as_tuple=True)[0]
        opp = (live & ~is_trav).nonzero(as_tuple=True)[0]
        env = self.slabs[0]
        decode = lambda idx: env.records_information_state_tensor(records, idx, PLAYER_CURRENT, dtype=torch.uint8)
        legal = words[trav] & 0x3FFFF
        last_legal = torch.floor(torch.log2(legal.double())).to(torch.uint8)          # highest legal action id
        adv = {"info_state": decode(trav), "advantage": regret[trav], "action": last_legal}
        strat = {"info_state": decode(opp), "strategy_action_probs": strategy[opp]}
        return adv, strat

    def close(self):
        for s in self.slabs:
            s.close()
