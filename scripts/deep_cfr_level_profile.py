"""Where the time of a small Deep CFR traversal batch goes (1500 roots, the thesis setting: e-outcome, factor 2, e 0.2,
512-512 advantage networks): wall time per batch and the op table of torch.profiler. Run on the GPU box.
    python scripts/deep_cfr_level_profile.py [outcome|e-outcome]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open_spiel_coup_b200.deep_cfr import DeepCFRSolver

method = sys.argv[1] if len(sys.argv) > 1 else "e-outcome"
kw = dict(sampling_method="e-outcome", outcome_factor=2, e_outcome=0.2) if method == "e-outcome" else dict(sampling_method="outcome")
def make(device_levels):
    return DeepCFRSolver(policy_network_layers=(1024, 1024), advantage_network_layers=(512, 512), num_traversals=1500,
                         memory_capacity=1 << 22, max_nodes=1 << 15, roots_per_batch=1500, seed=1, max_tree_nodes=1 << 25,
                         device_levels=device_levels, **kw)


for device_levels in (False, True):            # the host-driven level loop of round 1, then the device-count engine
    solver = make(device_levels)
    for _ in range(8):                         # warm-up: also grows the reservoir buffers to their working size
        solver.traverse(0, 1500); solver.traverse(1, 1500)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps, nodes, levels = 10, 0, 0
    for _ in range(reps):
        for p in (0, 1):
            nodes += solver.traverse(p, 1500)[1]
            levels += len(solver.last_level_widths)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / (2 * reps)
    print("%s, %s: 1500 traversals: %.1f ms per batch, %.0f nodes, %.1f levels, %.2f ms per level" % (
        method, "device-side level sizes" if device_levels else "host-driven levels", dt * 1e3, nodes / (2 * reps),
        levels / (2 * reps), dt * 1e3 * 2 * reps / levels), flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    solver.traverse(0, 1500)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=14, max_name_column_width=50))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=50))
