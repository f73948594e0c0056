"""Bit-exactness replay (BASELINE.json configs[4]).

GPU trajectories (action + chance logs) are replayed through the reference C++ CoupState
(oracle/_ref/libcoup_ref.so, i.e. the unmodified /root/reference sources) -- or through the C oracle when
that build is absent -- and every reported state must agree on: legal-action set, current player, terminal
flag, rewards, returns, and the info-state / observation tensors of both players. Tensors are compared
through a 64-bit position-keyed hash of what the encoder kernels actually wrote to HBM
(coup_tensor_row_hash); all fields are folded into one 64-bit digest per trajectory on both sides.

    python scripts/replay_check.py --envs 1048576 --steered 4096 --out profiles/replay_check.json

This script is verification tooling: it may use oracle/. The product path never does.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from open_spiel_coup_b200 import _lib  # noqa: E402
from open_spiel_coup_b200.vector_env import CoupVectorEnv  # noqa: E402

K = 0x9E3779B97F4A7C15 - (1 << 64)   # the fold multiplier as a signed int64


def fold(d, v):
    return d * K + v.to(torch.int64) + 1          # int64 arithmetic wraps mod 2^64, as on the CPU side


def fold_state(env, d, active):
    n = env.num_envs
    new = d
    new = fold(new, env.legal_mask.to(torch.int64) & 0xFFFFFFFF)
    new = fold(new, env.current_player.to(torch.int64) & 0xFF)
    new = fold(new, env.done)
    rew, ret = env.rewards.to(torch.int64), env.returns.to(torch.int64)
    for t in (rew[:, 0], rew[:, 1], ret[:, 0], ret[:, 1]):
        new = fold(new, t + 2)
    info = env.information_state_tensor(_lib.PLAYER_BOTH)
    h = env.tensor_row_hash(info).view(n, 2)
    del info
    new = fold(fold(new, h[:, 0]), h[:, 1])
    obs = env.observation_tensor(_lib.PLAYER_BOTH)
    h = env.tensor_row_hash(obs).view(n, 2)
    new = fold(fold(new, h[:, 0]), h[:, 1])
    return torch.where(active, new, d)


def steering_policy(env):
    """Prefers Pass, Exchange and the ExchangeReturn actions: produces the 91-move truncated games that
    uniform-random play reaches once in ~1.6 M episodes (SURVEY.md section 8a verification recipe)."""
    legal = env.legal_mask
    a = env.sample_uniform().to(torch.int64)
    for act in (15, 16, 17, 5, 9):                 # later entries win
        a = torch.where((legal >> act) & 1 == 1, torch.full_like(a, act), a)
    a = torch.where(env.done.bool(), torch.zeros_like(a), a)
    return a.to(torch.uint8)


def run(env, max_steps, steered):
    n = env.num_envs
    d = torch.zeros(n, dtype=torch.int64, device=env.device)
    all_on = torch.ones(n, dtype=torch.bool, device=env.device)
    d = fold_state(env, d, all_on)                # the first decision node after the deals
    reports = torch.ones(n, dtype=torch.int64, device=env.device)
    for _ in range(max_steps):
        active = ~env.done.bool()
        if not bool(active.any()):
            break
        if steered:
            env.step(steering_policy(env))
        else:
            env.rollout(1)
        d = fold_state(env, d, active)
        reports += active
    assert bool(env.done.all()), "some episode did not finish"
    return d, reports


def check(envs, steered, seed, device, threads, use_reference=True):
    from oracle.bindings import Oracle, Reference
    checker_kind = "reference (oracle/_ref/libcoup_ref.so)" if (use_reference and Reference.available()) else "oracle C port"
    checker = Reference() if (use_reference and Reference.available()) else Oracle()
    out = {"checker": checker_kind, "batches": []}
    total_mismatch = 0
    for name, n, st in (("uniform-random (fused rollout kernel)", envs, False), ("steered long games (step kernel)", steered, True)):
        if n <= 0:
            continue
        t0 = time.time()
        env = CoupVectorEnv(n, seed=seed + (1 if st else 0), device=device, auto_reset=False)
        d, reports = run(env, 100, st)
        torch.cuda.synchronize()
        t_gpu = time.time() - t0
        trajs = env.trajectories()
        stats = env.stats()
        flat = np.concatenate([a for a, _ in trajs])
        off = np.concatenate([[0], np.cumsum([len(a) for a, _ in trajs])]).astype(np.int64)
        t0 = time.time()
        dig, rep, bad = checker.replay_digest_batch(flat, off, threads)
        t_cpu = time.time() - t0
        gd = d.cpu().numpy().view(np.uint64)
        gr = reports.cpu().numpy()
        mism = int((gd != dig).sum() + (gr != rep).sum()) + int(bad)
        total_mismatch += mism
        lens = np.diff(off)
        # a chance move that follows an ExchangeReturn somewhere earlier in the same game (deck quirk exercised)
        post_er = 0
        for a, tgt in trajs:
            er = np.nonzero((tgt < 0) & (a >= 12))[0]
            if len(er) and (tgt[er[0]:] >= 0).any():
                post_er += 1
        out["batches"].append({
            "policy": name, "trajectories": int(n), "moves": int(lens.sum()), "reported_states": int(gr.sum()),
            "mismatching_trajectories": mism, "rejected_by_checker": int(bad), "truncated_91_move_games": int((lens > 90).sum()),
            "games_with_chance_after_exchange_return": int(post_er), "episodes_stat": stats["episodes"],
            "gpu_seconds": round(t_gpu, 2), "checker_seconds": round(t_cpu, 2), "checker_threads": threads,
        })
        del env
        torch.cuda.empty_cache()
    out["total_mismatches"] = total_mismatch
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steered", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    res = check(args.envs, args.steered, args.seed, args.device, args.threads)
    text = json.dumps(res, indent=1)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")
    sys.exit(0 if res["total_mismatches"] == 0 else 1)


if __name__ == "__main__":
    main()
