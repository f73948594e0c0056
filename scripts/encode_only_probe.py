"""Stand-alone encoder vs fused rollout step at 2^20 envs (how much of the fused step is the rules phase)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.vector_env import CoupVectorEnv
n = 1 << 20
env = CoupVectorEnv(n, auto_reset=True); env.rollout(100)
def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16"), (torch.uint8, "u8")):
    out = torch.empty((n, 2492), dtype=dt, device="cuda")
    enc = t(lambda: env.information_state_tensor(_lib.PLAYER_CURRENT, out=out))
    fused = t(lambda: env.rollout(1, _lib.PLAYER_CURRENT, out=out))
    b = out.numel() * out.element_size()
    print(f"{name}: encode-only {enc:.3f} ms ({b/enc/1e6:.0f} GB/s)   fused step {fused:.3f} ms ({b/fused/1e6:.0f} GB/s)")
    del out
print(f"rules only: {t(lambda: env.rollout(1)):.4f} ms")
