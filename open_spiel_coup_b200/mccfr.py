"""Outcome-sampling MCCFR on Coup with many episodes per iteration on the device.

Mirror of `open_spiel/python/algorithms/outcome_sampling_mccfr.py` (`OutcomeSamplingSolver`) and
`open_spiel/python/algorithms/mccfr.py` (`MCCFRSolverBase`, `AveragePolicy`), as driven by
`coup_experiments/scripts/mccfr.py:46-60`. The reference plays ONE episode per update player per iteration,
recursing through Python and keying its regret table by `information_state_string`. Here one iteration plays
`num_envs` episodes per update player at once:

  forward   every env follows the sampling policy (regret matching on the table; the update player mixes in
            `expl` uniform exploration, outcome_sampling_mccfr.py:88-98); the info-state key is the 64-bit hash of
            the player's info-state tensor row (coup_tensor_row_hash), which identifies the same information sets
            as the string does; the table is a sorted key array searched with `searchsorted`
  backward  one sweep over the recorded steps computes the value estimates, the sampled counterfactual regrets and
            the average-strategy increments of outcome_sampling_mccfr.py:100-145
  merge     the increments of all episodes are added into the table at once (new information sets start from
            1e-6 on their legal actions, mccfr.py:93-98)

All episodes of one iteration read the table as it was when the iteration started (batch-synchronous updates);
`num_envs` = 1 is the reference's sequential order. Chance nodes are resolved inside the env step; their
probabilities, which the reference multiplies into `opp_reach` and `sample_reach` (outcome_sampling_mccfr.py:73-77),
are recomputed from the deal codes the step appended to the history and the deck it left behind.
"""
import torch

from ._lib import NUM_DISTINCT_ACTIONS, PLAYER_CURRENT
from .deep_cfr import _legal_bool
from .vector_env import CoupVectorEnv

REGRET_INDEX = 0          # mccfr.py:20-21
AVG_POLICY_INDEX = 1
_INIT = 1e-6              # mccfr.py:93-98


class InfostateTable:
    """`MCCFRSolverBase._infostates` (mccfr.py:70-107) on the device: sorted 64-bit keys with, per key, the legal
    action bits and two float64 rows over the 18 action ids (cumulative regrets, cumulative average strategy)."""

    def __init__(self, device):
        self.device = device
        self.keys = torch.empty(0, dtype=torch.int64, device=device)
        self.legal_bits = torch.empty(0, dtype=torch.int32, device=device)
        self.values = torch.empty((0, 2, NUM_DISTINCT_ACTIONS), dtype=torch.float64, device=device)

    def __len__(self):
        return int(self.keys.numel())

    def lookup(self, keys, legal):
        """Rows [k, 2, 18] for `keys`; information sets not in the table read as their initial value."""
        init = (legal.double() * _INIT).unsqueeze(1).expand(-1, 2, -1)
        if not len(self):
            return init.clone(), torch.zeros(keys.shape, dtype=torch.bool, device=self.device)
        idx = torch.searchsorted(self.keys, keys).clamp_max(len(self) - 1)
        found = self.keys[idx] == keys
        return torch.where(found.view(-1, 1, 1), self.values[idx], init), found

    def add(self, keys, legal_bits, increments):
        """Inserts the keys that are new (initial value on their legal actions) and adds `increments` [k, 2, 18]."""
        if keys.numel() == 0:
            return
        uniq, inverse = torch.unique(keys, return_inverse=True)
        summed = torch.zeros((uniq.numel(), 2, NUM_DISTINCT_ACTIONS), dtype=torch.float64, device=self.device)
        summed.index_add_(0, inverse, increments)
        bits = torch.zeros(uniq.numel(), dtype=torch.int32, device=self.device)
        bits[inverse] = legal_bits.to(torch.int32)
        if len(self):
            pos = torch.searchsorted(self.keys, uniq).clamp_max(len(self) - 1)
            known = self.keys[pos] == uniq
            self.values.index_add_(0, pos[known], summed[known])
            uniq, summed, bits = uniq[~known], summed[~known], bits[~known]
        if uniq.numel():
            fresh = (_legal_bool(bits).double() * _INIT).unsqueeze(1) + summed
            keys_all = torch.cat([self.keys, uniq])
            order = torch.argsort(keys_all)
            self.keys = keys_all[order]
            self.legal_bits = torch.cat([self.legal_bits, bits])[order]
            self.values = torch.cat([self.values, fresh])[order]


def regret_matching(regrets, legal):
    """`MCCFRSolverBase._regret_matching` (mccfr.py:117-131) over the legal actions of each row."""
    pos = regrets.clamp_min(0.0) * legal
    total = pos.sum(-1, keepdim=True)
    uniform = legal.double() / legal.sum(-1, keepdim=True).clamp_min(1)
    return torch.where(total > 0, pos / total.clamp_min(1e-300), uniform)


class AveragePolicy:
    """`mccfr.AveragePolicy` (mccfr.py:24-60): the normalised cumulative average strategy; uniform over the legal
    actions where the information set was never visited."""

    def __init__(self, solver):
        self._solver = solver

    @torch.no_grad()
    def action_probs(self, info_state, legal_bits):
        """Batched form: info-state rows [k, 2492] (any of uint8 / bf16 / f32) and legal bits -> probs [k, 18]."""
        s = self._solver
        keys = s._env.tensor_row_hash(info_state)
        legal = _legal_bool(legal_bits)
        rows, found = s._table.lookup(keys, legal)
        av = rows[:, AVG_POLICY_INDEX] * legal
        uniform = legal.double() / legal.sum(-1, keepdim=True).clamp_min(1)
        return torch.where(found.view(-1, 1), av / av.sum(-1, keepdim=True).clamp_min(1e-300), uniform).float()

    def action_probabilities(self, state, player_id=None):
        if player_id is None:
            player_id = state.current_player()
        legal_actions = state.legal_actions()
        dev = self._solver._env.device
        info = torch.tensor(state.information_state_tensor(player_id), dtype=torch.float32, device=dev).view(1, -1)
        bits = torch.tensor([sum(1 << a for a in legal_actions)], dtype=torch.int32, device=dev)
        p = self.action_probs(info, bits)[0].tolist()
        return {a: p[a] for a in legal_actions}


def chance_reach(env, ids, first_deal_pos):
    """Product of the chance probabilities `deck_[c] / sum(deck_)` (coup.cc:1062-1077) of the deals that the last
    reset / step appended to the history of envs `ids` at positions >= first_deal_pos (at most 4): walking the deal
    codes backwards from the end of the history puts each dealt card back into the deck it was drawn from."""
    state = env.state[ids].long()
    g = state[:, 2]
    deck = torch.stack([(g >> (4 * c)) & 15 for c in range(5)], dim=1)
    new_moves = state[:, 3] & 127
    hist = env.history[ids].long() & 0xFFFFFFFF
    rows = torch.arange(ids.numel(), device=ids.device)
    prob = torch.ones(ids.numel(), dtype=torch.float64, device=ids.device)
    for t in range(4):
        pos = new_moves - 1 - t
        valid = pos >= first_deal_pos
        pos = pos.clamp_min(0)
        code = (hist[rows, pos // 6] >> (5 * (pos % 6))) & 31
        is_deal = valid & (code >= 18)
        card = ((code - 18) % 5).clamp(0, 4)
        deck = deck + torch.nn.functional.one_hot(card, 5) * is_deal.view(-1, 1)
        p = deck.gather(1, card.view(-1, 1)).view(-1).double() / deck.sum(1).double()
        prob = torch.where(is_deal, prob * p, prob)
    return prob


class OutcomeSamplingSolver:
    """`outcome_sampling_mccfr.OutcomeSamplingSolver` with `num_envs` episodes per update player per iteration."""

    def __init__(self, game=None, num_envs=4096, device=0, seed=0, record=False):
        self._game = game
        self._num_players = 2
        self._expl = 0.6                                              # outcome_sampling_mccfr.py:31
        self._env = CoupVectorEnv(num_envs, seed=seed + 1, device=device, auto_reset=False)
        self.device = self._env.device
        self._table = InfostateTable(self.device)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed)
        self._record = record
        self.last_episodes = None

    @property
    def infostates(self):
        return self._table

    def average_policy(self):
        return AveragePolicy(self)

    def iteration(self):
        """outcome_sampling_mccfr.py:38-48: one batch of episodes for each player as the update player."""
        self.last_episodes = []
        for update_player in range(self._num_players):
            self._episodes(update_player)

    @torch.no_grad()
    def _episodes(self, update_player):
        env, dev = self._env, self.device
        n = env.num_envs
        env.reset()
        idx = torch.arange(n, device=dev)
        my_reach = torch.ones(n, dtype=torch.float64, device=dev)
        opp_reach = chance_reach(env, idx, torch.zeros(n, dtype=torch.long, device=dev))     # the four initial deals
        sample_reach = opp_reach.clone()
        steps = []
        alive = torch.ones(n, dtype=torch.bool, device=dev)
        while bool(alive.any()):
            ids = idx[alive]
            rows = env.information_state_tensor_gather(ids, PLAYER_CURRENT, dtype=torch.uint8)
            keys = env.tensor_row_hash(rows)
            legal_bits = env.legal_mask[ids]
            legal = _legal_bool(legal_bits)
            updating = env.current_player[ids].long() == update_player
            entry, _ = self._table.lookup(keys, legal)
            policy = regret_matching(entry[:, REGRET_INDEX], legal)
            uniform = legal.double() / legal.sum(-1, keepdim=True)
            sample_policy = torch.where(updating.view(-1, 1), self._expl * uniform + (1.0 - self._expl) * policy, policy)
            action = torch.multinomial(sample_policy, 1, generator=self._gen).view(-1)
            p_a = policy.gather(1, action.view(-1, 1)).view(-1)
            s_a = sample_policy.gather(1, action.view(-1, 1)).view(-1)
            steps.append({"ids": ids, "keys": keys, "legal_bits": legal_bits, "legal": legal, "policy": policy,
                          "sample_policy": sample_policy, "action": action, "updating": updating,
                          "my_reach": my_reach[ids], "opp_reach": opp_reach[ids], "sample_reach": sample_reach[ids]})
            my_reach[ids] = torch.where(updating, my_reach[ids] * p_a, my_reach[ids])
            opp_reach[ids] = torch.where(updating, opp_reach[ids], opp_reach[ids] * p_a)
            sample_reach[ids] *= s_a
            moves = torch.full((n,), 0xFF, dtype=torch.uint8, device=dev)
            moves[ids] = action.to(torch.uint8)
            action_pos = env.move_numbers()[ids]                       # the action lands here, deals follow it
            env.step(moves)
            dealt = chance_reach(env, ids, action_pos + 1)
            opp_reach[ids] *= dealt
            sample_reach[ids] *= dealt
            alive &= ~env.done.bool()
        value = env.returns[:, update_player].double()                 # state.player_return(update_player)
        all_keys, all_bits, all_inc = [], [], []
        for st in reversed(steps):
            ids = st["ids"]
            child_value = value[ids]
            # baseline-corrected child values with baseline 0 (outcome_sampling_mccfr.py:50-60): only the sampled
            # action has a non-zero estimate
            sampled = torch.zeros_like(st["policy"]).scatter_(1, st["action"].view(-1, 1), 1.0)
            s_a = st["sample_policy"].gather(1, st["action"].view(-1, 1)).view(-1)
            child_values = sampled * (child_value / s_a).view(-1, 1)
            value_estimate = (st["policy"] * child_values).sum(-1)
            ratio = st["opp_reach"] / st["sample_reach"]
            cf_value = value_estimate * ratio
            regret_inc = (child_values * ratio.view(-1, 1) - cf_value.view(-1, 1)) * st["legal"]
            avstrat_inc = (st["my_reach"] / st["sample_reach"]).view(-1, 1) * st["policy"] * st["legal"]
            inc = torch.stack([regret_inc, avstrat_inc], dim=1) * st["updating"].view(-1, 1, 1)
            all_keys.append(st["keys"]); all_bits.append(st["legal_bits"]); all_inc.append(inc)
            value[ids] = value_estimate
        # every visited information set gets a table entry (mccfr.py:70-98), the update player's also its increments
        self._table.add(torch.cat(all_keys), torch.cat(all_bits), torch.cat(all_inc))
        if self._record:
            self.last_episodes.append({"update_player": update_player, "steps": steps, "root_value": value.clone(),
                                       "histories": env.trajectories()})
