"""Tree sizes and traversal throughput of the batched Deep CFR traversals (open_spiel_coup_b200/deep_cfr.py) on one
GPU: python scripts/deep_cfr_probe.py [--layers 128,128]. Prints one JSON line per configuration."""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from open_spiel_coup_b200.deep_cfr import DeepCFRSolver  # noqa: E402


def run(method, factor, roots, layers, max_nodes, reps=3, device_levels=True, e_outcome=0.0):
    solver = DeepCFRSolver(policy_network_layers=layers, advantage_network_layers=layers, sampling_method=method,
                           outcome_factor=factor, e_outcome=e_outcome, memory_capacity=1 << 22, max_nodes=max_nodes,
                           roots_per_batch=roots, seed=1, max_tree_nodes=1 << 25, device_levels=device_levels)
    solver.traverse(0, roots)                                  # warm-up + tree shape
    widths = solver.last_level_widths
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nodes = 0
    for r in range(reps):
        nodes += solver.traverse(r & 1, roots)[1]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"method": method, "outcome_factor": factor, "roots": roots, "levels": len(widths),
                      "engine": "device-side level sizes (cfr_traversal)" if solver._device_levels else "host-driven levels",
                      "widest_level": max(widths), "nodes_per_traversal_batch": sum(widths),
                      "nodes_per_root": sum(widths) / roots, "seconds_per_batch": dt / reps,
                      "nodes_per_s": nodes / dt, "traversals_per_s": roots * reps / dt,
                      "advantage_records": len(solver.advantage_buffers[0]) + len(solver.advantage_buffers[1]),
                      "strategy_records": len(solver.strategy_buffer),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    for s in solver._slabs:
        s.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", default="128,128")
    ap.add_argument("--external-roots", type=int, default=0, help="also time full-width external-sampling traversals of this many deals")
    ap.add_argument("--thesis-batch", action="store_true",
                    help="the thesis configuration (deep_cfr-final4.cfg): 1500 roots, e-outcome, outcome_factor 2, e 0.2, "
                         "advantage networks 512-512, with both engines")
    args = ap.parse_args()
    layers = tuple(int(x) for x in args.layers.split(","))
    if args.thesis_batch:
        for dl in (False, True):
            run("e-outcome", 2, 1500, (512, 512), 1 << 15, reps=5, device_levels=dl, e_outcome=0.2)
            run("outcome", 1, 1500, (512, 512), 1 << 15, reps=5, device_levels=dl)
    elif args.external_roots:
        run("external", 1, args.external_roots, layers, 1 << 21, reps=1)
    else:
        run("outcome", 1, 1 << 14, layers, 1 << 18)
        run("outcome", 1, 1 << 17, layers, 1 << 18)
        try:
            run("outcome", 2, 4, layers, 1 << 20, reps=1)
        except RuntimeError as e:                              # 2^(traverser decisions) nodes per root
            print(json.dumps({"method": "outcome", "outcome_factor": 2, "error": str(e)}), flush=True)
