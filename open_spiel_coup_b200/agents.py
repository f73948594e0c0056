"""The reference's learning agents for Coup, batched over many concurrent games on the device.

Mirrors, with the same hyper-parameters and record layouts:
  * `open_spiel/python/algorithms/dqn.py`  -> `DQN`  (epsilon-greedy Q-learning agent of one seat)
  * `open_spiel/python/algorithms/nfsp.py` -> `NFSP` (average-policy network + inner DQN, one mode per episode)
  * `coup_experiments/scripts/nfsp.py:134-144` and `coup_experiments/algorithms/rl_response.py:230-268`
    -> `run_episodes` (the loop "current player's agent steps, env steps, at the end every agent steps")
  * `coup_experiments/algorithms/rl_response.py` -> `rl_resp`, `eval_against_fixed_bots`, `FirstActionPolicy`
  * `coup_experiments/utils/nfsp_policies.py` -> `NFSPPolicies`

The reference steps ONE game and calls its networks with a batch of one. Here an agent is stepped with every env in
which its seat is to move (or whose episode just ended) at once: the info-state rows come from the gather encoder,
the per-(env, seat) "previous time step / previous action" of the reference agents lives in device arrays indexed
by env, and the replay / reservoir buffers are device tensors with uint8 info states.

Learning cadence. The reference counts agent steps and learns every `learn_every`-th one. A batched step advances
an agent's counter by the number of envs it was stepped with, and runs one gradient step for every multiple of
`learn_every` the counter passed, so the ratio of gradient steps to environment steps is the reference's. (Inside
one batched call the gradient steps see all of the call's new transitions, which a sequential run would feed in
one by one; `num_envs` = 1 reproduces the sequential order exactly.)
"""
import math
import os

import torch
from torch import nn

from ._lib import INFO_STATE_SIZE, NUM_DISTINCT_ACTIONS, PLAYER_BOTH, PLAYER_CURRENT
from .deep_cfr import MLP, _legal_bool
from .selfplay import ReplayBuffer, UniformRandomPolicy, masked_action_probs
from .vector_env import CoupVectorEnv

ILLEGAL_ACTION_LOGITS_PENALTY = -1e9   # dqn.py:34


class StepBatch:
    """The part of a batch of `rl_environment.TimeStep`s that one agent reads (rl_environment.py:57-97), for the envs
    it is stepped with: its own info-state rows (uint8), the legal-action bits (0 at a last step), `Rewards()` of its
    seat for the env step just made, and the LAST flag."""

    def __init__(self, env_ids, info_state, legal_bits, rewards, last):
        self.env_ids, self.info_state, self.legal_bits, self.rewards, self.last = env_ids, info_state, legal_bits, rewards, last

    def __len__(self):
        return int(self.env_ids.numel())

    def select(self, mask):
        return StepBatch(self.env_ids[mask], self.info_state[mask], self.legal_bits[mask], self.rewards[mask], self.last[mask])


def _crossings(old, new, every):
    return new // every - old // every


def _make_optimizer(params, optimizer_str, lr):
    if optimizer_str == "adam":
        return torch.optim.Adam(params, lr=lr)
    if optimizer_str == "sgd":
        return torch.optim.SGD(params, lr=lr)
    raise ValueError("Not implemented, choose from 'adam' and 'sgd'.")          # dqn.py:166, nfsp.py:133


def _uniform_legal(legal, gen):
    return torch.multinomial(legal.float(), 1, generator=gen).view(-1)


class DQN:
    """`dqn.DQN` (dqn.py:37-420) for seat `player_id`, stepped with batches of envs. Defaults as the reference."""

    def __init__(self, player_id, num_envs, hidden_layers_sizes=128, replay_buffer_capacity=10000, batch_size=128,
                 learning_rate=0.01, update_target_network_every=1000, learn_every=10, discount_factor=1.0,
                 min_buffer_size_to_learn=1000, epsilon_start=1.0, epsilon_end=0.1, epsilon_decay_duration=int(1e6),
                 optimizer_str="sgd", loss_str="mse", device="cuda", seed=0):
        if isinstance(hidden_layers_sizes, int):
            hidden_layers_sizes = [hidden_layers_sizes]
        if loss_str not in ("mse", "huber"):
            raise ValueError("Not implemented, choose from 'mse', 'huber'.")     # dqn.py:157
        if not isinstance(replay_buffer_capacity, int):
            raise ValueError("Replay buffer capacity not an integer.")           # dqn.py:85-86
        self.player_id = player_id
        self.device = torch.device(device)
        self._batch_size = batch_size
        self._update_target_network_every = update_target_network_every
        self._learn_every = learn_every
        self._min_buffer_size_to_learn = min_buffer_size_to_learn
        self._discount_factor = discount_factor
        self._epsilon_start, self._epsilon_end, self._epsilon_decay_duration = epsilon_start, epsilon_end, epsilon_decay_duration
        self._loss_str = loss_str
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed * 7919 + player_id)
        torch.manual_seed(seed * 7919 + player_id)
        self._q_network = MLP(INFO_STATE_SIZE, list(hidden_layers_sizes), NUM_DISTINCT_ACTIONS).to(self.device)
        self._target_q_network = MLP(INFO_STATE_SIZE, list(hidden_layers_sizes), NUM_DISTINCT_ACTIONS).to(self.device)
        self._optimizer = _make_optimizer(self._q_network.parameters(), optimizer_str, learning_rate)
        self._replay_buffer = ReplayBuffer(replay_buffer_capacity, self.device)
        # _prev_timestep / _prev_action of the reference (dqn.py:88-89), one slot per env
        self._prev_info = torch.zeros((num_envs, INFO_STATE_SIZE), dtype=torch.uint8, device=self.device)
        self._prev_action = torch.zeros(num_envs, dtype=torch.uint8, device=self.device)
        self._prev_valid = torch.zeros(num_envs, dtype=torch.bool, device=self.device)
        self._step_counter = 0
        self._last_loss_value = None

    # ---- accessors of the reference ------------------------------------------------------------------
    @property
    def replay_buffer(self):
        return self._replay_buffer

    @property
    def loss(self):
        return self._last_loss_value

    @property
    def step_counter(self):
        return self._step_counter

    def get_step_counter(self):
        return self._step_counter

    @property
    def q_network(self):
        return self._q_network

    def _get_epsilon(self, is_evaluation, power=1.0):
        """dqn.py:296-304."""
        if is_evaluation:
            return 0.0
        decay_steps = min(self._step_counter, self._epsilon_decay_duration)
        return self._epsilon_end + (self._epsilon_start - self._epsilon_end) * (1 - decay_steps / self._epsilon_decay_duration) ** power

    @torch.no_grad()
    def _epsilon_greedy(self, info_state, legal, epsilon):
        """dqn.py:270-294 for a batch: with probability epsilon a uniform legal action (probs uniform over the legal
        ones), else the legal action with the largest Q-value (probs one-hot)."""
        k = info_state.shape[0]
        q = self._q_network(info_state.float())
        greedy = torch.where(legal, q, torch.full_like(q, -math.inf)).argmax(-1)
        explore = torch.rand(k, device=self.device, generator=self._gen) < epsilon
        action = torch.where(explore, _uniform_legal(legal, self._gen), greedy)
        uniform = legal / legal.sum(-1, keepdim=True).clamp_min(1)
        onehot = torch.zeros_like(uniform).scatter_(1, greedy.view(-1, 1), 1.0)
        return action, torch.where(explore.view(-1, 1), uniform, onehot)

    def add_transitions(self, batch):
        """`add_transition(self._prev_timestep, self._prev_action, time_step)` (dqn.py:225-248) for every env of the
        batch that has a previous decision of this seat pending."""
        had = self._prev_valid[batch.env_ids]
        if bool(had.any()):
            ids = batch.env_ids[had]
            self._replay_buffer.add(self._prev_info[ids], self._prev_action[ids], batch.rewards[had], batch.info_state[had],
                                    batch.last[had], batch.legal_bits[had])

    def remember(self, batch, actions):
        """End of `step` (dqn.py:214-221): forget at a last step, else this decision becomes the pending one."""
        ids = batch.env_ids
        self._prev_valid[ids] = ~batch.last
        keep = ~batch.last
        self._prev_info[ids[keep]] = batch.info_state[keep]
        self._prev_action[ids[keep]] = actions[keep].to(torch.uint8)

    def step(self, batch, is_evaluation=False, add_transition_record=True):
        """dqn.py:175-223. Returns (actions int64 [k], probs float32 [k, 18]); rows at a last step get action 0 and
        all-zero probs (the reference returns nothing there)."""
        k = len(batch)
        actions = torch.zeros(k, dtype=torch.int64, device=self.device)
        probs = torch.zeros((k, NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=self.device)
        act = ~batch.last
        if bool(act.any()):
            a, p = self._epsilon_greedy(batch.info_state[act], _legal_bool(batch.legal_bits[act]), self._get_epsilon(is_evaluation))
            actions[act], probs[act] = a, p
        if not is_evaluation and k:
            old = self._step_counter
            self._step_counter += k
            for _ in range(_crossings(old, self._step_counter, self._learn_every)):
                self._last_loss_value = self.learn()
            if _crossings(old, self._step_counter, self._update_target_network_every):
                self._target_q_network.load_state_dict(self._q_network.state_dict())
            if add_transition_record:
                self.add_transitions(batch)
            self.remember(batch, actions)
        return actions, probs

    def learn(self):
        """dqn.py:306-337: one gradient step on `batch_size` transitions drawn without replacement; None while the
        buffer is smaller than the batch or than `min_buffer_size_to_learn`."""
        buf = self._replay_buffer
        if buf.size < self._batch_size or buf.size < self._min_buffer_size_to_learn:
            return None
        j = torch.randperm(buf.size, device=self.device, generator=self._gen)[: self._batch_size]
        with torch.no_grad():
            target_q = self._target_q_network(buf.next_info_state[j].float())
            illegal = (~_legal_bool(buf.legal_actions_mask[j])).float() * ILLEGAL_ACTION_LOGITS_PENALTY
            max_next_q = (target_q + illegal).max(-1).values
            target = buf.reward[j].float() + (1.0 - buf.is_final_step[j].float()) * self._discount_factor * max_next_q
        q = self._q_network(buf.info_state[j].float())
        predictions = q.gather(1, buf.action[j].long().view(-1, 1)).view(-1)
        loss = (nn.functional.mse_loss(predictions, target) if self._loss_str == "mse"
                else nn.functional.huber_loss(predictions, target, delta=1.0))
        self._optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self._optimizer.step()
        return float(loss.detach())

    def save(self, checkpoint_dir):
        os.makedirs(checkpoint_dir, exist_ok=True)
        for name, net in (("q_network", self._q_network), ("target_q_network", self._target_q_network)):
            torch.save(net.state_dict(), os.path.join(checkpoint_dir, f"{name}_pid{self.player_id}.pt"))

    def has_checkpoint(self, checkpoint_dir):
        return all(os.path.exists(os.path.join(checkpoint_dir, f"{n}_pid{self.player_id}.pt")) for n in ("q_network", "target_q_network"))

    def restore(self, checkpoint_dir):
        for name, net in (("q_network", self._q_network), ("target_q_network", self._target_q_network)):
            net.load_state_dict(torch.load(os.path.join(checkpoint_dir, f"{name}_pid{self.player_id}.pt"), map_location=self.device))


class _Reservoir:
    """`nfsp.ReservoirBuffer` (nfsp.py:322-371) holding `Transition(info_state, action_probs, legal_actions_mask)`
    (nfsp.py:36-37), on the device."""

    def __init__(self, capacity, device, seed):
        from .deep_cfr import ReservoirBuffer
        self._buf = ReservoirBuffer(capacity, device, {"info_state": ((INFO_STATE_SIZE,), torch.uint8),
                                                       "action_probs": ((NUM_DISTINCT_ACTIONS,), torch.float32),
                                                       "legal_actions_mask": ((), torch.int32)}, seed=seed)

    def add(self, info_state, action_probs, legal_bits):
        self._buf.add(info_state=info_state, action_probs=action_probs, legal_actions_mask=legal_bits)

    def sample(self, n):
        return self._buf.sample(n)

    def __len__(self):
        return len(self._buf)


class MODE:
    best_response = "best_response"       # nfsp.py:39
    average_policy = "average_policy"


class NFSP:
    """`nfsp.NFSP` (nfsp.py:40-320) for seat `player_id`, stepped with batches of envs. Every env carries its own
    episode mode (`_sample_episode_policy`, nfsp.py:146-150: best response with probability `anticipatory_param`)."""

    def __init__(self, player_id, num_envs, hidden_layers_sizes, reservoir_buffer_capacity, anticipatory_param,
                 batch_size=128, rl_learning_rate=0.01, sl_learning_rate=0.01, min_buffer_size_to_learn=1000,
                 learn_every=64, optimizer_str="sgd", device="cuda", seed=0, **kwargs):
        self.player_id = player_id
        self.device = torch.device(device)
        self._batch_size = batch_size
        self._learn_every = learn_every
        self._anticipatory_param = anticipatory_param
        self._min_buffer_size_to_learn = min_buffer_size_to_learn
        self._reservoir_buffer = _Reservoir(reservoir_buffer_capacity, self.device, seed * 31 + player_id)
        self._step_counter = 0
        kwargs.update(batch_size=batch_size, learning_rate=rl_learning_rate, learn_every=learn_every,
                      min_buffer_size_to_learn=min_buffer_size_to_learn, optimizer_str=optimizer_str)
        self._rl_agent = DQN(player_id, num_envs, hidden_layers_sizes, device=device, seed=seed, **kwargs)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed * 104729 + player_id)
        self._avg_network = MLP(INFO_STATE_SIZE, list(hidden_layers_sizes), NUM_DISTINCT_ACTIONS).to(self.device)
        self._optimizer = _make_optimizer(self._avg_network.parameters(), optimizer_str, sl_learning_rate)
        self._last_sl_loss_value = None
        self._forced_mode = None
        self._best_response = torch.zeros(num_envs, dtype=torch.bool, device=self.device)
        self._sample_episode_policy(torch.arange(num_envs, device=self.device))

    def _sample_episode_policy(self, env_ids):
        self._best_response[env_ids] = torch.rand(env_ids.numel(), device=self.device, generator=self._gen) < self._anticipatory_param

    class _TempMode:
        def __init__(self, agent, mode):
            self.agent, self.mode = agent, mode

        def __enter__(self):
            self.prev, self.agent._forced_mode = self.agent._forced_mode, self.mode

        def __exit__(self, *exc):
            self.agent._forced_mode = self.prev

    def temp_mode_as(self, mode):
        """nfsp.py:135-141: every env acts in `mode` inside the context."""
        return NFSP._TempMode(self, mode)

    def get_step_counter(self):
        return self._step_counter

    @property
    def loss(self):
        return (self._last_sl_loss_value, self._rl_agent.loss)

    @property
    def avg_network(self):
        return self._avg_network

    @property
    def rl_agent(self):
        return self._rl_agent

    @property
    def reservoir_buffer(self):
        return self._reservoir_buffer

    @torch.no_grad()
    def _act(self, info_state, legal_bits):
        """nfsp.py:152-167: softmax of the average network, illegal actions removed, renormalised, sampled."""
        probs = masked_action_probs(self._avg_network(info_state.float()), legal_bits)
        return torch.multinomial(probs, 1, generator=self._gen).view(-1), probs

    def step(self, batch, is_evaluation=False):
        """nfsp.py:177-227."""
        k = len(batch)
        actions = torch.zeros(k, dtype=torch.int64, device=self.device)
        probs = torch.zeros((k, NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=self.device)
        if k == 0:
            return actions, probs
        if self._forced_mode is None:
            br = self._best_response[batch.env_ids]
        else:
            br = torch.full((k,), self._forced_mode == MODE.best_response, dtype=torch.bool, device=self.device)
        n_br = int(br.sum())
        if n_br:
            sub = batch.select(br)
            a, p = self._rl_agent.step(sub, is_evaluation)
            actions[br], probs[br] = a, p
            if not is_evaluation:                                   # nfsp.py:191-192, 233-247
                rec = ~sub.last
                self._reservoir_buffer.add(sub.info_state[rec], p[rec], sub.legal_bits[rec])
        if n_br < k:
            avg = ~br
            sub = batch.select(avg)
            act = ~sub.last
            a = torch.zeros(len(sub), dtype=torch.int64, device=self.device)
            p = torch.zeros((len(sub), NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=self.device)
            if bool(act.any()):
                a[act], p[act] = self._act(sub.info_state[act], sub.legal_bits[act])
            actions[avg], probs[avg] = a, p
            if not is_evaluation:                                   # nfsp.py:202-204: the inner agent still learns from it
                self._rl_agent.add_transitions(sub)
                self._rl_agent.remember(sub, a)
        if not is_evaluation:
            old = self._step_counter
            self._step_counter += k
            crossed = _crossings(old, self._step_counter, self._learn_every)
            for _ in range(crossed):
                self._last_sl_loss_value = self._learn()
            # "If learn step not triggered by rl policy, learn" (nfsp.py:213-215): the share of this call's steps
            # that were made in average-policy mode
            for _ in range(round(crossed * (k - n_br) / k)):
                self._rl_agent._last_loss_value = self._rl_agent.learn()
            ended = batch.env_ids[batch.last]
            if ended.numel():
                self._sample_episode_policy(ended)                  # nfsp.py:218-222
        return actions, probs

    def _learn(self):
        """nfsp.py:249-277: cross entropy of the average network against the stored behaviour probabilities."""
        if len(self._reservoir_buffer) < self._batch_size or len(self._reservoir_buffer) < self._min_buffer_size_to_learn:
            return None
        t = self._reservoir_buffer.sample(self._batch_size)
        logits = self._avg_network(t["info_state"].float())
        loss = -(t["action_probs"] * torch.log_softmax(logits, dim=-1)).sum(-1).mean()
        self._optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self._optimizer.step()
        return float(loss.detach())

    def save(self, checkpoint_dir, checkpoint_id=""):
        os.makedirs(checkpoint_dir, exist_ok=True)
        for name, net in (("q_network", self._rl_agent.q_network), ("avg_network", self._avg_network)):
            torch.save(net.state_dict(), os.path.join(checkpoint_dir, f"{name}{checkpoint_id}_pid{self.player_id}.pt"))

    def has_checkpoint(self, checkpoint_dir, checkpoint_id=""):
        return all(os.path.exists(os.path.join(checkpoint_dir, f"{n}{checkpoint_id}_pid{self.player_id}.pt")) for n in ("q_network", "avg_network"))

    def restore(self, checkpoint_dir, checkpoint_id=""):
        for name, net in (("q_network", self._rl_agent.q_network), ("avg_network", self._avg_network)):
            net.load_state_dict(torch.load(os.path.join(checkpoint_dir, f"{name}{checkpoint_id}_pid{self.player_id}.pt"), map_location=self.device))


# ---- fixed agents ------------------------------------------------------------------------------------
class FirstActionPolicy:
    """`rl_response.FirstActionAgent` (rl_response.py:132-149): always the first (lowest) legal action."""


class PolicyAgent:
    """A fixed agent acting from logits of an `nn.Module` (masked softmax, as `NFSP._act`), from the built-in
    `UniformRandomPolicy` / `FirstActionPolicy`, or from any object with `action_probs(info_state, legal_bits)` such
    as `deep_cfr.DeepCFRSolver` (rl_response.py:112-129 `PolicyAgent`, `random_agent.RandomAgent`)."""

    def __init__(self, player_id, policy, device="cuda", seed=0):
        self.player_id, self.policy, self.device = player_id, policy, torch.device(device)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed * 15485863 + player_id)

    @torch.no_grad()
    def step(self, batch, is_evaluation=False):
        k = len(batch)
        actions = torch.zeros(k, dtype=torch.int64, device=self.device)
        probs = torch.zeros((k, NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=self.device)
        act = ~batch.last
        if not bool(act.any()):
            return actions, probs
        legal_bits = batch.legal_bits[act]
        legal = _legal_bool(legal_bits)
        if isinstance(self.policy, UniformRandomPolicy):
            p = legal / legal.sum(-1, keepdim=True)
        elif isinstance(self.policy, FirstActionPolicy):
            p = torch.zeros_like(legal, dtype=torch.float32).scatter_(1, legal.to(torch.int32).argmax(-1, keepdim=True), 1.0)
        elif hasattr(self.policy, "action_probs"):
            p = self.policy.action_probs(batch.info_state[act], legal_bits).float()
        else:
            p = masked_action_probs(self.policy(batch.info_state[act].float()), legal_bits)
        actions[act] = torch.multinomial(p, 1, generator=self._gen).view(-1)
        probs[act] = p
        return actions, probs


class NFSPPolicies:
    """`coup_experiments/utils/nfsp_policies.NFSPPolicies`: the joint policy of two NFSP agents in a fixed mode, in
    the batched form the evaluation loops here consume (`action_probs`) and the single-state form of the reference
    (`action_probabilities(state)`)."""

    def __init__(self, nfsp_policies, mode=MODE.average_policy):
        self._policies, self._mode = nfsp_policies, mode

    def agents(self):
        return [_ModeAgent(a, self._mode) for a in self._policies]

    def action_probabilities(self, state, player_id=None):
        cur = state.current_player()
        legal_actions = state.legal_actions(cur)
        agent = self._policies[cur]
        info = torch.tensor(state.information_state_tensor(cur), dtype=torch.float32, device=agent.device).to(torch.uint8).view(1, -1)
        bits = torch.tensor([sum(1 << a for a in legal_actions)], dtype=torch.int32, device=agent.device)
        batch = StepBatch(torch.zeros(1, dtype=torch.long, device=agent.device), info, bits,
                          torch.zeros(1, device=agent.device), torch.zeros(1, dtype=torch.bool, device=agent.device))
        with agent.temp_mode_as(self._mode):
            _, p = agent.step(batch, is_evaluation=True)
        return {a: float(p[0, a]) for a in legal_actions}


class _ModeAgent:
    def __init__(self, agent, mode):
        self.agent, self.mode, self.player_id = agent, mode, agent.player_id

    def step(self, batch, is_evaluation=True):
        with self.agent.temp_mode_as(self.mode):
            return self.agent.step(batch, is_evaluation=True)


# ---- the episode loop --------------------------------------------------------------------------------
def run_episodes(env, agents, num_episodes, is_evaluation=False, on_episode_end=None):
    """Plays `num_episodes` complete games on the envs of `env` (a CoupVectorEnv WITHOUT auto-reset; at most
    `env.num_envs` at a time), agents[p] acting for seat p: the loop of `coup_experiments/scripts/nfsp.py:134-144`.
    At each step the agent of the seat to move is stepped with the envs where that is the case; when an env's
    episode ends, BOTH agents are stepped with its final time step (info state of their own seat, their own reward,
    no legal actions); `on_episode_end(env, env_ids)` is then called, before those envs are touched again. Returns
    the sum over episodes of `Returns()` per seat and the number of decision steps."""
    dev = env.device
    n = env.num_envs
    idx = torch.arange(n, device=dev)
    totals = torch.zeros(2, dtype=torch.float64, device=dev)
    played = steps = 0
    while played < num_episodes:
        active_n = min(n, num_episodes - played)
        mask = idx < active_n
        env.reset(envs_to_reset=mask.to(torch.uint8))
        alive = mask.clone()
        while bool(alive.any()):
            ids = idx[alive]
            rows = env.information_state_tensor_gather(ids, PLAYER_CURRENT, dtype=torch.uint8)
            seat = env.current_player[ids].long()
            legal = env.legal_mask[ids]
            actions = torch.full((n,), 0xFF, dtype=torch.uint8, device=dev)     # 0xFF: envs not being played sit out
            for p in (0, 1):
                sel = seat == p
                if bool(sel.any()):
                    batch = StepBatch(ids[sel], rows[sel], legal[sel], env.rewards[ids[sel], p].float(),
                                      torch.zeros(int(sel.sum()), dtype=torch.bool, device=dev))
                    a, _ = agents[p].step(batch, is_evaluation)
                    actions[ids[sel]] = a.to(torch.uint8)
            env.step(actions)
            steps += int(ids.numel())
            ended = alive & env.done.bool()
            if bool(ended.any()):
                eids = idx[ended]
                final = env.information_state_tensor_gather(eids, PLAYER_BOTH, dtype=torch.uint8).view(-1, 2, INFO_STATE_SIZE)
                for p in (0, 1):
                    batch = StepBatch(eids, final[:, p], torch.zeros(eids.numel(), dtype=torch.int32, device=dev),
                                      env.rewards[eids, p].float(), torch.ones(eids.numel(), dtype=torch.bool, device=dev))
                    agents[p].step(batch, is_evaluation)
                totals += env.returns[eids].double().sum(0)
                if on_episode_end is not None:
                    on_episode_end(env, eids)
                alive &= ~ended
        played += active_n
    return totals.cpu(), steps


def eval_against_fixed_bots(env, trained_agents, fixed_agents, num_episodes):
    """rl_response.py:65-91: for each seat, the trained agent of that seat against the fixed agent in the other one,
    `num_episodes` games, no learning; returns the mean episode reward of the trained seat."""
    out = []
    for player_pos in range(2):
        cur_agents = list(fixed_agents)
        cur_agents[player_pos] = trained_agents[player_pos]
        totals, _ = run_episodes(env, cur_agents, num_episodes, is_evaluation=True)
        out.append(float(totals[player_pos]) / num_episodes)
    return out


class RollingAverage:
    """rl_response.py:152-171."""

    def __init__(self, size=100):
        self._size, self._values, self._index, self._total_additions = size, [0.0] * size, 0, 0

    def add(self, value):
        self._values[self._index] = value
        self._total_additions += 1
        self._index = (self._index + 1) % self._size

    def mean(self):
        n = min(self._size, self._total_additions)
        return 0 if n == 0 else sum(self._values) / n


def rl_resp(exploitee="random", seed=0, window_size=30, num_train_episodes=1000000, eval_every=1000, eval_episodes=1000,
            replay_buffer_capacity=100000, batch_size=32, hidden_layers_sizes=None, num_envs=1024, device=0, log=None):
    """`rl_response.rl_resp` (rl_response.py:174-268): trains one DQN per seat (discount 0.99, epsilon 0.5 -> 0.1,
    rl_response.py:94-110) against a fixed exploitee and reports how much it wins: every `eval_every` episodes the
    mean reward of each trained seat over `eval_episodes` games and their sum, the exploitability estimate the
    reference logs as `value`. `exploitee`: "random", "first", an `nn.Module` producing logits, an object with
    `action_probs` (a `DeepCFRSolver`), or an `NFSPPolicies`. `num_envs` games are played at a time. Returns the list
    of evaluation records."""
    hidden_layers_sizes = [int(x) for x in (hidden_layers_sizes or [64, 64, 64])]
    dev = torch.device("cuda", device)
    env = CoupVectorEnv(num_envs, seed=seed + 1, device=device, auto_reset=False)
    if isinstance(exploitee, NFSPPolicies):
        exploitee_agents = exploitee.agents()
    elif exploitee == "random":
        exploitee_agents = [PolicyAgent(p, UniformRandomPolicy(), dev, seed) for p in range(2)]
    elif exploitee == "first":
        exploitee_agents = [PolicyAgent(p, FirstActionPolicy(), dev, seed) for p in range(2)]
    elif isinstance(exploitee, str):
        raise RuntimeError("Unknown exploitee")                                   # rl_response.py:206
    else:
        exploitee_agents = [PolicyAgent(p, exploitee, dev, seed) for p in range(2)]
    learning_agents = [DQN(p, num_envs, hidden_layers_sizes, replay_buffer_capacity=replay_buffer_capacity,
                           batch_size=batch_size, discount_factor=0.99, epsilon_start=0.5, epsilon_end=0.1,
                           device=dev, seed=seed) for p in range(2)]
    rolling, rolling_p0, rolling_p1 = RollingAverage(window_size), RollingAverage(window_size), RollingAverage(window_size)
    total_value = total_value_n = 0
    records = []
    ep = 0
    while ep < num_train_episodes:
        chunk = min(num_envs, eval_every - ep % eval_every, num_train_episodes - ep)
        for agents in ([learning_agents[0], exploitee_agents[1]], [exploitee_agents[0], learning_agents[1]]):
            run_episodes(env, agents, chunk)
        ep += chunk
        if ep % eval_every == 0:
            r_mean = eval_against_fixed_bots(env, learning_agents, exploitee_agents, eval_episodes)
            value = r_mean[0] + r_mean[1]
            for avg, v in ((rolling, value), (rolling_p0, r_mean[0]), (rolling_p1, r_mean[1])):
                avg.add(v)
            total_value += value
            total_value_n += 1
            rec = {"episode": ep, "r_mean": r_mean, "value": value, "rval": rolling.mean(), "rval_p0": rolling_p0.mean(),
                   "rval_p1": rolling_p1.mean(), "aval": total_value / total_value_n}
            records.append(rec)
            if log is not None:
                log("[{episode}] Mean episode rewards {r_mean}, value: {value}, rval: {rval} (p0/p1: {rval_p0} / {rval_p1}), "
                    "aval: {aval}".format(**rec))
    env.close()
    return records


def train_nfsp(num_train_episodes, hidden_layers_sizes, num_envs=1024, eval_every=None, eval_func=None, device=0, seed=0,
               **kwargs):
    """The main loop of `coup_experiments/scripts/nfsp.py:120-160`: two NFSP agents in self-play, `num_envs` games at
    a time; every `eval_every` episodes `eval_func(joint_average_policy, episode, losses)` is called (the reference
    runs `rl_resp` on the joint average policy there). Returns the agents."""
    dev = torch.device("cuda", device)
    env = CoupVectorEnv(num_envs, seed=seed + 1, device=device, auto_reset=False)
    agents = [NFSP(p, num_envs, hidden_layers_sizes, device=dev, seed=seed, **kwargs) for p in range(2)]
    joint_avg_policy = NFSPPolicies(agents, MODE.average_policy)
    ep = 0
    while ep < num_train_episodes:
        chunk = min(num_envs, num_train_episodes - ep, (eval_every - ep % eval_every) if eval_every else num_envs)
        run_episodes(env, agents, chunk)
        ep += chunk
        if eval_every and eval_func is not None and ep % eval_every == 0:
            eval_func(joint_avg_policy, ep, [a.loss for a in agents])
    env.close()
    return agents


def agent_cmp(policy_a, policy_b, cmp_test_eps, num_envs=4096, device=0, seed=0):
    """`coup_experiments/scripts/agent_cmp.py:123-149`: two fixed policies against each other, `cmp_test_eps` games
    with `policy_a` in seat 0 and as many with the seats swapped. Policies: anything `PolicyAgent` accepts (an
    `nn.Module` giving logits, `UniformRandomPolicy`, `FirstActionPolicy`, an object with `action_probs` such as a
    `DeepCFRSolver` or an MCCFR `AveragePolicy`) or agents with a batched `step` (`NFSPPolicies(...).agents()[seat]`).
    Returns `policy_a`'s mean reward (the other's is its negative) and the mean episode length in moves, chance
    nodes included, as the reference counts them."""
    dev = torch.device("cuda", device)
    env = CoupVectorEnv(min(num_envs, cmp_test_eps), seed=seed + 1, device=device, auto_reset=False)

    def agent(policy, seat):
        return policy if hasattr(policy, "step") else PolicyAgent(seat, policy, dev, seed)

    total_reward, total_moves = 0.0, 0
    for a_seat in (0, 1):
        players = [None, None]
        players[a_seat], players[1 - a_seat] = agent(policy_a, a_seat), agent(policy_b, 1 - a_seat)
        env.clear_stats()
        totals, _ = run_episodes(env, players, cmp_test_eps, is_evaluation=True)
        stats = env.stats()
        total_reward += float(totals[a_seat])
        total_moves += stats["decision_steps"] + stats["chance_moves"]
    env.close()
    return total_reward / (2 * cmp_test_eps), total_moves / (2 * cmp_test_eps)
