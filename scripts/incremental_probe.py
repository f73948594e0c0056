"""Runs the incremental contract (k_rollout_incremental<float>, 2^20 envs, both views) for timing / ncu:
python scripts/incremental_probe.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.vector_env import CoupVectorEnv

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
n = 1 << 20
env = CoupVectorEnv(n, seed=1, auto_reset=True)
env.rollout(100)
buf = torch.empty((2 * n, 2492), dtype=torch.float32, device=env.device)
env.information_state_tensor(_lib.PLAYER_BOTH, out=buf)
env.rollout_incremental(5, buf)
torch.cuda.synchronize()
t0 = time.perf_counter()
env.rollout_incremental(steps, buf)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
print("incremental f32 both views: %.3f ms/step, %.3e steps/s, %.0f GB/s algorithmic (945 B/step)" % (dt * 1e3, n / dt, 945 * n / dt / 1e9))
