"""Runs the incremental contract (k_rollout_incremental<float>, 2^20 envs, both views) for timing / ncu:
python scripts/incremental_probe.py [steps] [row stride: 2492 (default, reference layout) or 2496 (32-byte aligned rows)]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.vector_env import CoupVectorEnv

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
strides = [int(x) for x in sys.argv[2:]] or [2492]
n = 1 << 20
for stride in strides:
    env = CoupVectorEnv(n, seed=1, auto_reset=True)
    env.rollout(100)
    buf = torch.empty((2 * n, stride), dtype=torch.float32, device=env.device)
    env.information_state_tensor(_lib.PLAYER_BOTH, out=buf)
    env.rollout_incremental(5, buf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.rollout_incremental(steps, buf); e1.record(); torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / steps
    ref = torch.empty_like(buf)
    env.information_state_tensor(_lib.PLAYER_BOTH, out=ref)
    same = bool((env.tensor_row_hash(buf[:, :2492].contiguous()) == env.tensor_row_hash(ref[:, :2492].contiguous())).all()) if stride != 2492 else bool((env.tensor_row_hash(buf) == env.tensor_row_hash(ref)).all())
    print("incremental f32 both views, stride %d: %.3f ms/step, %.3e steps/s, %.0f GB/s algorithmic (945 B/step), buffer == dense encoder: %s"
          % (stride, dt * 1e3, n / dt, 945 * n / dt / 1e9, same))
    env.close(); del buf, ref
    torch.cuda.empty_cache()
