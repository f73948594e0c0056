"""Times the tensor-free kernels on one GPU: env-only rollout (k_rollout_env_multi), k_step + k_sample_uniform, k_fork.
    python scripts/env_only_probe.py [--envs N] [--steps K]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_spiel_coup_b200.vector_env import CoupVectorEnv  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=1280)
    ap.add_argument("--ring", type=int, default=1 << 18)
    args = ap.parse_args()
    n, k = args.envs, args.steps
    out = {}
    for ring in sorted({0, args.ring}):
        env = CoupVectorEnv(n, seed=1234, auto_reset=True, finished_ring=ring)
        env.rollout(128)
        env.clear_stats()
        t = timed(lambda: env.rollout(k))
        st = env.stats()
        assert st["decision_steps"] == n * k and st["illegal"] == 0
        out[f"env_only_ring{ring}"] = {"steps_per_s": n * k / t, "ms_per_step": 1e3 * t / k, "episodes": st["episodes"]}
        if ring == 0:
            acts = torch.empty(n, dtype=torch.uint8, device=env.device)

            def loop():
                for _ in range(200):
                    env.sample_uniform(out=acts)
                    env.step(acts)
            loop()
            t = timed(loop)
            out["sample_uniform_plus_step"] = {"steps_per_s": n * 200 / t, "ms_per_step": 1e3 * t / 200}
            child = CoupVectorEnv(n, seed=5, auto_reset=False)
            env.sample_uniform(out=acts)
            for tag, parents in (("random_parents", torch.randint(0, n, (n,), device=env.device, dtype=torch.int32)),
                                 ("two_children_per_parent", (torch.arange(n, device=env.device, dtype=torch.int32) // 2))):
                a = acts[parents.long()].contiguous()
                child.fork_from(env, parents, a)
                t = timed(lambda: [child.fork_from(env, parents, a) for _ in range(50)])
                out["fork_" + tag] = {"children_per_s": n * 50 / t, "ms_per_call": 1e3 * t / 50}
            child.close()
        env.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
