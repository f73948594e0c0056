"""Where the time of a small Deep CFR traversal batch goes (1500 roots, the thesis setting): wall time per batch and
the CPU-side op table of torch.profiler. Run on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200.deep_cfr import DeepCFRSolver

solver = DeepCFRSolver(policy_network_layers=(1024, 1024), advantage_network_layers=(512, 512), num_traversals=1500,
                       sampling_method="outcome", memory_capacity=100000, max_nodes=1 << 16, seed=1)
for _ in range(3):
    solver.traverse(0, 1500); solver.traverse(1, 1500)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 10
for _ in range(reps):
    solver.traverse(0, 1500); solver.traverse(1, 1500)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / (2 * reps)
print("1500 outcome-sampling traversals: %.1f ms per batch, %d levels, %.2f ms per level" % (dt * 1e3, len(solver.last_level_widths), dt * 1e3 / len(solver.last_level_widths)))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    solver.traverse(0, 1500)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=18, max_name_column_width=40))
