"""Throughput of coup_vec_fork (k_fork): 2^20 children of random parents per launch. Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200.vector_env import CoupVectorEnv

n = 1 << 20
src = CoupVectorEnv(n, seed=1)
src.rollout(6)                                    # mid-game parents, (almost) all alive
dst = CoupVectorEnv(n, seed=2)
g = torch.Generator(device="cuda").manual_seed(0)
alive = (src.done == 0).nonzero(as_tuple=True)[0]
parents = alive[torch.randint(0, alive.numel(), (n,), device="cuda", generator=g)]
legal = ((src.legal_mask[parents].view(-1, 1) >> torch.arange(18, device="cuda", dtype=torch.int32)) & 1).float()
actions = torch.multinomial(legal, 1, generator=g).view(-1).to(torch.uint8)
for order, par, act in (("random parents", parents, actions),):
    for _ in range(3):
        dst.fork_from(src, par, act)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    ev0.record()
    for _ in range(reps):
        dst.fork_from(src, par, act)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    bytes_per_child = 80 + 80 + 5 + 14
    print("k_fork, %s: %.1f us per 2^20 children, %.2e children/s, %.0f GB/s algorithmic (%d B per child)" % (
        order, ms * 1e3, n / (ms * 1e-3), bytes_per_child * n / (ms * 1e-3) / 1e9, bytes_per_child))
dst.check_errors()
