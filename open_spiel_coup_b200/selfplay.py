"""Self-play data generation on the device (BASELINE.json configs[3]).

Replaces the batch-1 loops of the reference's experiments
(`coup_experiments/scripts/nfsp.py:134-144`, `coup_experiments/algorithms/rl_response.py:65-91`):
a PyTorch MLP policy reads the info-state tensor of the player to move straight from HBM, actions are
drawn on the device with the reference agents' acting rule (`open_spiel/python/algorithms/nfsp.py:154-167`:
softmax, zero the illegal actions, renormalise, sample), and the supervised-learning records of NFSP
(`nfsp.py:36-37,226-242`: `Transition(info_state, action_probs, legal_actions_mask)`) go into a device
reservoir buffer (`nfsp.py:322-371`). PyTorch is used for the network and buffers only; environment
rules, encoding and sampling are the CUDA kernels behind the C ABI.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib
from ._lib import INFO_STATE_SIZE, NUM_DISTINCT_ACTIONS, PLAYER_CURRENT, PLAYER_FROM_RECORD, RECORD_WORDS

# Row stride of the policy input: 2492 padded to a multiple of 32 elements. K = 2492 is not a multiple of 8, which
# takes a bf16 cuBLAS GEMM off its fast path (5.2 ms vs 0.93 ms for [2^18, K] x [K, 1024] on B200); the encoder
# writes zeros into the four pad columns (coup_vec_information_state_tensor_strided).
PADDED_INFO_STATE_SIZE = 2496
# History rows 91..134 of the info-state tensor can never be set (a game has at most 91 moves, chance nodes included:
# MaxGameLength 90, coup.h:219, termination at move_number_ > 90, coup.cc:990; the tensor reserves 135 rows, coup.cc:1104-1116),
# so elements >= 62 + 18 * 91 = 1700 are always zero and the first layer's product over them is exactly zero. The first
# GEMM therefore runs over the first 1728 columns only (1700 rounded up to 64 elements), and the self-play loop asks the
# encoder for exactly those columns (COUP_LIVE_INFO_STATE_SIZE rows: 31 % fewer bytes written and read back).
from ._lib import LIVE_INFO_STATE_SIZE  # noqa: E402
from .vector_env import CoupVectorEnv


class MLPPolicy(nn.Module):
    """Linear+ReLU stack with a linear head: the shape of the reference's `MLP`
    (`open_spiel/python/pytorch/dqn.py:36-107`) at the thesis sizes 2492 -> 1024 -> 1024 -> 18
    (`coup_experiments/scripts/flags/thesis_runs/nfsp-final1.cfg:2`)."""

    def __init__(self, input_size=INFO_STATE_SIZE, hidden_sizes=(1024, 1024), output_size=NUM_DISTINCT_ACTIONS,
                 padded_input_size=None):
        super().__init__()
        self.input_size = input_size
        # weights of the pad columns only ever see zeros: they do not change the function or its gradients
        self.padded_input_size = padded_input_size or input_size
        layers, prev = [], self.padded_input_size
        for h in hidden_sizes:
            layers += [nn.Linear(prev, h), nn.ReLU()]
            prev = h
        layers.append(nn.Linear(prev, output_size))
        self.net = nn.Sequential(*layers)

    def _first_layer(self, x):
        """(input, weight) of the first layer. An info-state input narrower than the weight is a live prefix (or the
        unpadded row): the weight is cut to its width -- the columns dropped only ever multiply zeros. On the inference
        path a full-width input is cut to the live prefix as well."""
        w = self.net[0].weight
        if self.input_size != INFO_STATE_SIZE:
            return x, w
        k = x.shape[-1]
        if k > LIVE_INFO_STATE_SIZE and not torch.is_grad_enabled():
            k = LIVE_INFO_STATE_SIZE
        if k < LIVE_INFO_STATE_SIZE:
            raise ValueError(f"info-state rows must have at least {LIVE_INFO_STATE_SIZE} columns")
        return x[..., :k], w[:, :min(k, w.shape[1])]          # views: lda / ldb stay the row strides

    def forward(self, info_state):
        x, w = self._first_layer(info_state)
        layers = list(self.net)
        if torch.is_grad_enabled() or not info_state.is_cuda or info_state.dim() != 2:
            x = torch.nn.functional.linear(x, w, layers[0].bias)
            for layer in layers[1:]:
                x = layer(x)
            return x
        # inference on the device: bias + ReLU run in the GEMM epilogue (cuBLASLt) instead of as two more
        # passes over the [num_envs, hidden] activations
        i = 0
        while i < len(layers):
            lin = layers[i]
            if i > 0:
                w = lin.weight
            if i + 1 < len(layers) and isinstance(layers[i + 1], nn.ReLU):
                x = torch._addmm_activation(lin.bias, x, w.t())
                i += 2
            else:
                x = torch.addmm(lin.bias, x, w.t())
                i += 1
        return x


def masked_action_probs(logits, legal_mask_bits):
    """Plain PyTorch statement of `NFSP._act` (nfsp.py:154-167), used to check the fused kernel:
    probs = softmax(logits); probs[illegal] = 0; probs /= probs.sum()."""
    bits = torch.arange(NUM_DISTINCT_ACTIONS, device=logits.device, dtype=torch.int32)
    legal = ((legal_mask_bits.view(-1, 1).to(torch.int32) >> bits) & 1).to(torch.bool)
    probs = torch.softmax(logits.float(), dim=-1) * legal
    return probs / probs.sum(-1, keepdim=True).clamp_min(1e-30)


class ReservoirBuffer:
    """Uniform sample of everything ever added (`nfsp.py:322-371`), batched and on the device: the element
    with running index t replaces slot randint(0, t) when that is < capacity; within one batch the later
    element wins a slot, as it would sequentially."""

    def __init__(self, capacity, device, info_dtype=torch.uint8, seed=0):
        self.capacity = int(capacity)
        self.device = device
        self.info_state = torch.zeros((self.capacity, INFO_STATE_SIZE), dtype=info_dtype, device=device)
        self.action_probs = torch.zeros((self.capacity, NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=device)
        self.legal_actions_mask = torch.zeros(self.capacity, dtype=torch.int32, device=device)
        self.add_calls = 0
        self.size = 0
        self._gen = torch.Generator(device=device)
        self._gen.manual_seed(seed)

    def add(self, info_state, action_probs, legal_mask_bits, valid=None):
        idx = torch.arange(info_state.shape[0], device=self.device)
        if valid is not None:
            idx = idx[valid]
        b = int(idx.numel())
        if b == 0:
            return
        t = self.add_calls + torch.arange(b, device=self.device)           # running index of each element
        fill = t < self.capacity
        draw = (torch.rand(b, device=self.device, generator=self._gen, dtype=torch.float64) * (t + 1).double()).long()
        slot = torch.where(fill, t, draw)
        keep = slot < self.capacity
        slot, src = slot[keep], idx[keep]
        order = torch.arange(src.numel(), device=self.device)
        winner = torch.full((self.capacity,), -1, dtype=torch.long, device=self.device)
        winner.scatter_reduce_(0, slot, order, reduce="amax", include_self=True)   # later element wins
        chosen = winner[winner >= 0]
        dst, src = slot[chosen], src[chosen]
        self.info_state[dst] = info_state[src].to(self.info_state.dtype)
        self.action_probs[dst] = action_probs[src]
        self.legal_actions_mask[dst] = legal_mask_bits[src].to(torch.int32)
        self.add_calls += b
        self.size = min(self.capacity, self.add_calls)

    def sample(self, batch_size):
        j = torch.randint(0, self.size, (batch_size,), device=self.device, generator=self._gen)
        return self.info_state[j], self.action_probs[j], self.legal_actions_mask[j]


class DeviceRecorder:
    """The reference agents' two data stores, filled by the step kernel itself (coup_vec_step_record,
    include/coup_b200.h): the NFSP reservoir of `Transition(info_state, action_probs, legal_actions_mask)`
    (nfsp.py:36-37,226-242,322-371) and the DQN replay buffer of `Transition(info_state, action, reward,
    next_info_state, is_final_step, legal_actions_mask)` (dqn.py:30-80,223-246), both as PACKED 96-byte observation
    records. `sample_*` decodes only the sampled records into network inputs. No per-step torch op, no host read-back;
    the only host state is the count of offered reservoir elements, which is `num_envs` per step."""

    def __init__(self, env, reservoir_capacity=0, replay_capacity=0, seed=0):
        self.env, dev, n = env, env.device, env.num_envs
        self.reservoir_capacity, self.replay_capacity = int(reservoir_capacity), int(replay_capacity)
        self.offered = 0
        z = lambda *shape, dtype=torch.int32: torch.zeros(shape, dtype=dtype, device=dev)
        self.res_records = z(max(self.reservoir_capacity, 1), RECORD_WORDS)
        self.res_probs = z(max(self.reservoir_capacity, 1), NUM_DISTINCT_ACTIONS, dtype=torch.float32)
        self.res_winner = z(max(self.reservoir_capacity, 1), dtype=torch.int64)
        self.transitions = z(max(self.replay_capacity, 1), 2, RECORD_WORDS)
        self.replay_total = z(1, dtype=torch.int64)
        self.pending = z(n, 2, RECORD_WORDS)
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(seed)

    def _buffers(self):
        p = lambda t: C.c_void_p(t.data_ptr())
        b = _lib.RecorderBuffers()
        if self.reservoir_capacity:
            b.d_reservoir_records, b.d_reservoir_probs, b.d_reservoir_winner = p(self.res_records), p(self.res_probs), p(self.res_winner)
            b.reservoir_capacity, b.reservoir_offered = self.reservoir_capacity, self.offered
        if self.replay_capacity:
            b.d_transitions, b.replay_capacity = p(self.transitions), self.replay_capacity
            b.d_replay_total, b.d_pending = p(self.replay_total), p(self.pending)
        return b

    def step(self, actions, action_probs=None):
        """env.step(actions) + recording. `action_probs` float32 [num_envs, 18] is needed for the reservoir."""
        env = self.env
        if self.reservoir_capacity and (action_probs is None or action_probs.dtype != torch.float32 or not action_probs.is_contiguous()):
            raise ValueError("the reservoir needs contiguous float32 action probabilities")
        b = self._buffers()
        _lib.check(env._lib.coup_vec_step_record(env._h, C.c_void_p(actions.data_ptr()),
                                                 C.c_void_p(action_probs.data_ptr()) if action_probs is not None else None,
                                                 C.byref(b), C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)))
        self.offered += env.num_envs

    # ---- learner side --------------------------------------------------------------------------------
    @property
    def reservoir_size(self):
        return min(self.reservoir_capacity, self.offered)

    @property
    def replay_size(self):
        return min(self.replay_capacity, int(self.replay_total.item()))

    def sample_reservoir(self, batch_size, dtype=torch.float32):
        """(info_state [b, 2492], action_probs [b, 18], legal mask bits int32 [b]) of uniformly drawn records."""
        j = torch.randint(0, self.reservoir_size, (batch_size,), device=self.env.device, generator=self._gen)
        info = self.env.records_information_state_tensor(self.res_records, j, PLAYER_FROM_RECORD, dtype=dtype)
        return info, self.res_probs[j], self.res_records[j, 21] & 0x3FFFF

    def decode_transitions(self, idx, dtype=torch.float32):
        """The transitions at buffer positions `idx`: (info_state, action, reward, next_info_state, is_final_step,
        legal mask bits of the next state)."""
        recs = self.transitions.view(-1, RECORD_WORDS)
        idx = idx.to(torch.int64)
        info = self.env.records_information_state_tensor(recs, 2 * idx, PLAYER_FROM_RECORD, dtype=dtype)
        nxt = self.env.records_information_state_tensor(recs, 2 * idx + 1, PLAYER_FROM_RECORD, dtype=dtype)
        meta = self.transitions[idx, 0, 21]
        return (info, (meta & 31).to(torch.uint8), (((meta >> 5) & 7) - 2).to(torch.int8), nxt, ((meta >> 8) & 1).to(torch.uint8),
                self.transitions[idx, 1, 21] & 0x3FFFF)

    def sample_replay(self, batch_size, dtype=torch.float32):
        j = torch.randint(0, self.replay_size, (batch_size,), device=self.env.device, generator=self._gen)
        return self.decode_transitions(j, dtype=dtype)


# Upper bounds of the move-number buckets of `bucketed_policy_forward` and the first-layer K each needs:
# history row i is move i of the episode (coup.cc:230-256), so a state with move number m has no non-zero beyond
# column 62 + 18 m; K is that bound rounded up to a multiple of 64 (the last bucket takes the whole row).
_BUCKET_MOVES = (7, 10, 14, 17, 21, 28, 39, 60, 127)
_BUCKET_K = tuple(min(PADDED_INFO_STATE_SIZE, -(-(62 + 18 * m) // 64) * 64) for m in _BUCKET_MOVES[:-1]) + (PADDED_INFO_STATE_SIZE,)


class _Tail:
    """The layers after the first Linear+ReLU of an MLPPolicy, for MLPPolicy.forward's fused path."""

    input_size = None                      # not an info-state input: the first layer takes it as it is
    _first_layer = MLPPolicy._first_layer

    def __init__(self, net):
        self.net = net


@torch.no_grad()
def bucketed_policy_forward(env, policy, info_buf, logits_out=None):
    """`policy(info_state)` for every env of `env`, with the first Linear layer's contraction cut to the columns
    that can be non-zero: envs are ordered by move number, their rows are encoded in that order
    (coup_vec_information_state_tensor_gather), and each bucket multiplies only its first K columns. A state
    half-way through an average game uses K = 320 of 2496, so the dominant GEMM shrinks ~5x; the result differs
    from the dense forward only by the summation order of exact zeros. Returns (logits [num_envs, 18] in env
    order, perm, rows) where rows = info_buf holds the encoded rows in `perm` order.

    Measured on one B200 at 2^18 envs (scripts/selfplay_breakdown.py): the layer itself drops from 0.71 to ~0.47 ms
    (the small-K GEMMs are bound by writing the [n, 1024] activations, not by flops), but the bucket sizes need one
    device->host read per step, which exposes the launch latency of everything queued behind it; the whole
    self-play step is slower with it, so SelfPlayDataGen keeps the dense forward by default."""
    first = policy.net[0]
    n = env.num_envs
    moves = env.move_numbers()
    bounds = torch.tensor(_BUCKET_MOVES[:-1], device=moves.device)
    bucket = torch.bucketize(moves, bounds)                       # moves <= bounds[b] -> b
    perm = torch.argsort(bucket, stable=True)
    counts = torch.bincount(bucket, minlength=len(_BUCKET_MOVES)).cpu().tolist()
    env.information_state_tensor_gather(perm, PLAYER_CURRENT, out=info_buf)
    width = info_buf.shape[1]
    hidden = torch.empty((n, first.out_features), dtype=info_buf.dtype, device=info_buf.device)
    start = 0
    for k, c in zip(_BUCKET_K, counts):
        if c:
            k = min(k, width)
            torch._addmm_activation(first.bias, info_buf[start:start + c, :k], first.weight[:, :k].t(), out=hidden[start:start + c])
            start += c
    sorted_logits = MLPPolicy.forward(_Tail(policy.net[2:]), hidden)   # net[1] (ReLU) ran in the epilogue above
    if logits_out is None:
        logits_out = torch.empty_like(sorted_logits)
    logits_out.index_copy_(0, perm, sorted_logits)
    return logits_out, perm, info_buf


class SelfPlayDataGen:
    """`num_envs` concurrent self-play games driven by one policy network.

    One `step()`: encode the info-state of the player to move (bf16 by default: the network's input
    dtype, values are exact), run the policy, draw masked actions on the device, record, apply the
    actions (chance nodes and auto-reset are resolved inside the step kernel)."""

    def __init__(self, num_envs=1 << 18, policy=None, seed=1234, device=0, tensor_dtype=torch.bfloat16,
                 reservoir_capacity=0, global_env_offset=0, bucketed_first_layer=False, replay_capacity=0,
                 torch_reservoir=False):
        self.env = CoupVectorEnv(num_envs, seed=seed, device=device, global_env_offset=global_env_offset,
                                 auto_reset=True)
        self.bucketed_first_layer = bucketed_first_layer and isinstance(policy, (MLPPolicy, type(None)))
        dev = self.env.device
        if policy is None:
            policy = MLPPolicy(padded_input_size=PADDED_INFO_STATE_SIZE)
        self.policy = policy.to(device=dev, dtype=tensor_dtype).eval()
        width = getattr(policy, "padded_input_size", INFO_STATE_SIZE)
        if (isinstance(policy, MLPPolicy) and policy.input_size == INFO_STATE_SIZE and not self.bucketed_first_layer
                and not (reservoir_capacity and torch_reservoir)):
            width = LIVE_INFO_STATE_SIZE       # the encoder writes the live prefix of every row, the first GEMM reads it
        self.info_state = torch.zeros((num_envs, width), dtype=tensor_dtype, device=dev)
        self.action_probs = torch.empty((num_envs, NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=dev)
        self.actions = torch.empty(num_envs, dtype=torch.uint8, device=dev)
        self.acting_player = torch.empty(num_envs, dtype=torch.int8, device=dev)
        self.legal_before = torch.empty(num_envs, dtype=torch.int32, device=dev)
        # Records are kept by the step kernel (DeviceRecorder). torch_reservoir=True keeps the older torch-op reservoir
        # of dense uint8 rows instead (used by tests as a cross-check of the reservoir rule).
        self.reservoir = (ReservoirBuffer(reservoir_capacity, dev, seed=seed) if reservoir_capacity and torch_reservoir else None)
        self.recorder = None
        if (reservoir_capacity and not torch_reservoir) or replay_capacity:
            self.recorder = DeviceRecorder(self.env, 0 if torch_reservoir else reservoir_capacity, replay_capacity, seed=seed)
        self.steps = 0

    @torch.no_grad()
    def step(self):
        env = self.env
        perm = None
        if self.bucketed_first_layer:
            logits, perm, _ = bucketed_policy_forward(env, self.policy, self.info_state)
        else:
            env.information_state_tensor(PLAYER_CURRENT, out=self.info_state)
            logits = self.policy(self.info_state)
        self.acting_player.copy_(env.current_player)
        self.legal_before.copy_(env.legal_mask)
        env.sample_policy(logits, probs_out=self.action_probs, actions_out=self.actions)
        if self.reservoir is not None:
            if perm is None:
                self.reservoir.add(self.info_state[:, :INFO_STATE_SIZE], self.action_probs, self.legal_before)
            else:           # rows are in move-number order: line the other columns up with them
                self.reservoir.add(self.info_state[:, :INFO_STATE_SIZE], self.action_probs[perm], self.legal_before[perm])
        if self.recorder is not None:
            self.recorder.step(self.actions, self.action_probs)      # step + reservoir + replay in the step kernel
        else:
            env.step(self.actions)
        self.steps += 1
        # After the call: env.rewards / env.returns / env.done describe the transition just made
        # (done envs already hold a freshly dealt episode; env.returns is the finished episode's return).
        return self.actions

    def run(self, n_steps):
        for _ in range(n_steps):
            self.step()


class UniformRandomPolicy:
    """The `random` exploitee / baseline bot of the reference's evaluation scripts
    (`coup_experiments/algorithms/rl_response.py:60`): uniform over the legal actions."""


@torch.no_grad()
def evaluate_policies(policies, num_episodes, num_envs=1 << 16, seed=1234, device=0, tensor_dtype=torch.bfloat16):
    """Batched agent-vs-agent evaluation (`coup_experiments/scripts/agent_cmp.py`, and the inner loop of
    `rl_response.eval_against_fixed_bots`, rl_response.py:65-91): `policies[p]` acts whenever player p is
    to move; exactly `num_episodes` complete games are played (batches without auto-reset, so the sample of
    episodes is unbiased). Returns the mean episode reward of each seat and the number of decision steps."""
    totals = torch.zeros(2, dtype=torch.float64)
    played = steps = 0
    batch_seed = seed
    while played < num_episodes:
        n = min(num_envs, num_episodes - played)
        env = CoupVectorEnv(n, seed=batch_seed, device=device, auto_reset=False)
        dev = env.device
        nets = [None if isinstance(p, UniformRandomPolicy) else p.to(device=dev, dtype=tensor_dtype).eval() for p in policies]
        width = max([getattr(p, "padded_input_size", INFO_STATE_SIZE) for p in nets if p is not None], default=INFO_STATE_SIZE)
        info = torch.zeros((n, width), dtype=tensor_dtype, device=dev) if any(p is not None for p in nets) else None
        for _ in range(100):                      # a game has at most 91 moves
            if bool(env.done.all()):
                break
            if info is not None:
                env.information_state_tensor(PLAYER_CURRENT, out=info)
            acts = []
            for p in nets:
                if p is None:
                    acts.append(env.sample_uniform())
                else:
                    acts.append(env.sample_policy(p(info[:, :getattr(p, "padded_input_size", INFO_STATE_SIZE)])))
            a = torch.where(env.current_player == 1, acts[1], acts[0])
            a = torch.where(env.done.bool(), torch.zeros_like(a), a)
            env.step(a)
        assert bool(env.done.all())
        totals += env.returns.double().sum(0).cpu()   # sum of per-step rewards == Returns() (coup.cc:1016-1032)
        steps += env.stats()["decision_steps"]
        played += n
        batch_seed += 1
        env.close()
    return (totals / played).tolist(), steps


def eval_against_fixed_bots(trained_policies, fixed_policies, num_episodes, **kw):
    """`rl_response.eval_against_fixed_bots` (rl_response.py:65-91): for each seat, the trained policy of that
    seat against the fixed policy in the other seat; returns the mean episode reward of the trained seat."""
    out = []
    for seat in range(2):
        cur = list(fixed_policies)
        cur[seat] = trained_policies[seat]
        mean, _ = evaluate_policies(cur, num_episodes, **kw)
        out.append(mean[seat])
    return out


class ReplayBuffer:
    """Circular replay buffer of DQN transitions on the device
    (`open_spiel/python/algorithms/dqn.py:30-32`: `Transition(info_state, action, reward, next_info_state,
    is_final_step, legal_actions_mask)`; `ReplayBuffer`, dqn.py:35-80). Info states are stored as uint8
    (values are 0, 1 and coin counts, exact)."""

    def __init__(self, capacity, device):
        self.capacity, self.device = int(capacity), device
        self.info_state = torch.zeros((self.capacity, INFO_STATE_SIZE), dtype=torch.uint8, device=device)
        self.next_info_state = torch.zeros((self.capacity, INFO_STATE_SIZE), dtype=torch.uint8, device=device)
        self.action = torch.zeros(self.capacity, dtype=torch.uint8, device=device)
        self.reward = torch.zeros(self.capacity, dtype=torch.int8, device=device)
        self.is_final_step = torch.zeros(self.capacity, dtype=torch.uint8, device=device)
        self.legal_actions_mask = torch.zeros(self.capacity, dtype=torch.int32, device=device)   # of the NEXT state
        self.total = 0

    @property
    def size(self):
        return min(self.capacity, self.total)

    def add(self, info_state, action, reward, next_info_state, is_final_step, legal_mask_bits):
        k = int(action.numel())
        if k == 0:
            return
        pos = (self.total + torch.arange(k, device=self.device)) % self.capacity
        self.info_state[pos] = info_state
        self.action[pos] = action
        self.reward[pos] = reward.to(torch.int8)
        self.next_info_state[pos] = next_info_state
        self.is_final_step[pos] = is_final_step.to(torch.uint8) if torch.is_tensor(is_final_step) else int(is_final_step)
        self.legal_actions_mask[pos] = legal_mask_bits.to(torch.int32) if torch.is_tensor(legal_mask_bits) else int(legal_mask_bits)
        self.total += k


class ReplayRecorder:
    """Self-play with the reference agents' transition bookkeeping, batched: every env plays both seats with
    the same acting policy, and for each (env, seat) the recorder keeps the seat's previous decision exactly as
    `DQN.step` keeps `_prev_timestep` / `_prev_action` (dqn.py:175-222): a transition is emitted when that
    seat acts again (reward = `Rewards()[seat]` at that moment, next state = its info state then), and, when
    the episode ends, for BOTH seats with the terminal info state, `is_final_step = 1` and an empty legal
    mask -- every agent is stepped with the final time step (`coup_experiments/scripts/nfsp.py:141-143`).
    Rewards that fall between a seat's own turns are dropped, as in the reference."""

    def __init__(self, num_envs, policy=None, seed=1234, device=0, tensor_dtype=torch.bfloat16, replay_capacity=1 << 20,
                 on_episode_end=None):
        self.env = CoupVectorEnv(num_envs, seed=seed, device=device, auto_reset=False)
        dev = self.env.device
        self.policy = None
        if not isinstance(policy, UniformRandomPolicy):
            policy = policy if policy is not None else MLPPolicy(padded_input_size=PADDED_INFO_STATE_SIZE)
            self.policy = policy.to(device=dev, dtype=tensor_dtype).eval()
            width = getattr(policy, "padded_input_size", INFO_STATE_SIZE)
            self.policy_input = torch.zeros((num_envs, width), dtype=tensor_dtype, device=dev)
        self.info_u8 = torch.empty((num_envs, INFO_STATE_SIZE), dtype=torch.uint8, device=dev)
        self.pend_info = torch.zeros((num_envs, 2, INFO_STATE_SIZE), dtype=torch.uint8, device=dev)
        self.pend_action = torch.zeros((num_envs, 2), dtype=torch.uint8, device=dev)
        self.pend_valid = torch.zeros((num_envs, 2), dtype=torch.bool, device=dev)
        self.replay = ReplayBuffer(replay_capacity, dev)
        self.on_episode_end = on_episode_end
        self._idx = torch.arange(num_envs, device=dev)

    @torch.no_grad()
    def step(self):
        env, idx = self.env, self._idx
        seat = env.current_player.long()                  # every env is at a decision node here
        env.information_state_tensor(PLAYER_CURRENT, out=self.info_u8)
        if self.policy is None:
            actions = env.sample_uniform()
        else:
            env.information_state_tensor(PLAYER_CURRENT, out=self.policy_input)
            actions = env.sample_policy(self.policy(self.policy_input))
        # the acting seat's previous decision becomes a transition (dqn.py:223-246)
        had = self.pend_valid[idx, seat]
        rows = idx[had]
        s = seat[rows]
        self.replay.add(self.pend_info[rows, s], self.pend_action[rows, s], env.rewards[rows, s], self.info_u8[rows],
                        torch.zeros_like(rows), env.legal_mask[rows])
        self.pend_info[idx, seat] = self.info_u8
        self.pend_action[idx, seat] = actions
        self.pend_valid[idx, seat] = True
        env.step(actions)
        done = env.done.bool()
        ids = idx[done]
        if ids.numel():
            final = env.information_state_tensor_gather(ids, _lib_player_both(), dtype=torch.uint8).view(-1, 2, INFO_STATE_SIZE)
            for p in (0, 1):
                sel = self.pend_valid[ids, p]
                rows = ids[sel]
                self.replay.add(self.pend_info[rows, p], self.pend_action[rows, p], env.rewards[rows, p], final[sel, p],
                                torch.ones_like(rows), torch.zeros_like(rows))
            if self.on_episode_end is not None:
                self.on_episode_end(self, ids)
            self.pend_valid[ids] = False
            env.reset(envs_to_reset=done)
        return actions


def _lib_player_both():
    from ._lib import PLAYER_BOTH
    return PLAYER_BOTH
