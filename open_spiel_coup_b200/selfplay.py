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
import torch
from torch import nn

from ._lib import INFO_STATE_SIZE, NUM_DISTINCT_ACTIONS, PLAYER_CURRENT

# Row stride of the policy input: 2492 padded to a multiple of 32 elements. K = 2492 is not a multiple of 8, which
# takes a bf16 cuBLAS GEMM off its fast path (5.2 ms vs 0.93 ms for [2^18, K] x [K, 1024] on B200); the encoder
# writes zeros into the four pad columns (coup_vec_information_state_tensor_strided).
PADDED_INFO_STATE_SIZE = 2496
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

    def forward(self, info_state):
        return self.net(info_state)


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


class SelfPlayDataGen:
    """`num_envs` concurrent self-play games driven by one policy network.

    One `step()`: encode the info-state of the player to move (bf16 by default: the network's input
    dtype, values are exact), run the policy, draw masked actions on the device, record, apply the
    actions (chance nodes and auto-reset are resolved inside the step kernel)."""

    def __init__(self, num_envs=1 << 18, policy=None, seed=1234, device=0, tensor_dtype=torch.bfloat16,
                 reservoir_capacity=0, global_env_offset=0):
        self.env = CoupVectorEnv(num_envs, seed=seed, device=device, global_env_offset=global_env_offset,
                                 auto_reset=True)
        dev = self.env.device
        if policy is None:
            policy = MLPPolicy(padded_input_size=PADDED_INFO_STATE_SIZE)
        self.policy = policy.to(device=dev, dtype=tensor_dtype).eval()
        width = getattr(policy, "padded_input_size", INFO_STATE_SIZE)
        self.info_state = torch.zeros((num_envs, width), dtype=tensor_dtype, device=dev)
        self.action_probs = torch.empty((num_envs, NUM_DISTINCT_ACTIONS), dtype=torch.float32, device=dev)
        self.actions = torch.empty(num_envs, dtype=torch.uint8, device=dev)
        self.acting_player = torch.empty(num_envs, dtype=torch.int8, device=dev)
        self.legal_before = torch.empty(num_envs, dtype=torch.int32, device=dev)
        self.reservoir = (ReservoirBuffer(reservoir_capacity, dev, seed=seed) if reservoir_capacity else None)
        self.steps = 0

    @torch.no_grad()
    def step(self):
        env = self.env
        env.information_state_tensor(PLAYER_CURRENT, out=self.info_state)
        logits = self.policy(self.info_state)
        self.acting_player.copy_(env.current_player)
        self.legal_before.copy_(env.legal_mask)
        env.sample_policy(logits, probs_out=self.action_probs, actions_out=self.actions)
        if self.reservoir is not None:
            self.reservoir.add(self.info_state[:, :INFO_STATE_SIZE], self.action_probs, self.legal_before)
        env.step(self.actions)
        self.steps += 1
        # After the call: env.rewards / env.returns / env.done describe the transition just made
        # (done envs already hold a freshly dealt episode; env.returns is the finished episode's return).
        return self.actions

    def run(self, n_steps):
        for _ in range(n_steps):
            self.step()
