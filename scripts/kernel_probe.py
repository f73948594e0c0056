"""Launches one of the secondary kernels a few times on 2^20 envs (for ncu): python scripts/kernel_probe.py obs|policy|mask|step|record|fork"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.vector_env import CoupVectorEnv

what = sys.argv[1]
n = 1 << 20
env = CoupVectorEnv(n, seed=1234, auto_reset=True)
env.rollout(100)
acts = torch.empty(n, dtype=torch.uint8, device=env.device)
for _ in range(5):
    if what == "obs":
        out = env.observation_tensor(_lib.PLAYER_BOTH)
    elif what == "policy":
        logits = torch.randn((n, 18), device=env.device)
        probs = torch.empty((n, 18), device=env.device)
        env.sample_policy(logits, probs_out=probs, actions_out=acts)
    elif what == "mask":
        env.legal_actions_mask()
    elif what == "step":
        env.sample_uniform(out=acts)
        env.step(acts)
    elif what == "record":
        from open_spiel_coup_b200.selfplay import DeviceRecorder
        if "rec" not in globals():
            rec = DeviceRecorder(env, reservoir_capacity=1 << 22, replay_capacity=1 << 22)
            probs = torch.full((n, 18), 1.0 / 18, device=env.device)
        env.sample_uniform(out=acts)
        rec.step(acts, probs)
    elif what == "fork":
        if "child" not in globals():
            child = CoupVectorEnv(n, seed=5, auto_reset=False)
            parents = torch.arange(n, device=env.device, dtype=torch.int32) // 2
        env.sample_uniform(out=acts)
        child.fork_from(env, parents, acts[parents.long()].contiguous())
torch.cuda.synchronize()
print("ok", what)
