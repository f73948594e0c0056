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


def check_autoreset(envs, episodes, seed, device, threads, use_reference=True, dtype=torch.float32):
    """The same bit-exactness replay in the mode bench.py times: COUP_FLAG_AUTO_RESET, fused rollout kernel, finished
    episodes harvested from the device ring (coup_vec_finished_*). Every state of every finished episode is folded into a
    per-episode digest on the GPU side -- decision states from the slab, the terminal state (which the env no longer
    holds: it was re-dealt inside the step kernel) from the ring record -- and compared with the checker's digest of the
    ring's action + chance log."""
    from oracle.bindings import Oracle, Reference
    from open_spiel_coup_b200.vector_env import decode_finished_records
    ref = use_reference and Reference.available()
    checker = Reference() if ref else Oracle()
    n = envs
    cap = 1
    while cap < episodes + 2 * n:
        cap *= 2
    env = CoupVectorEnv(n, seed=seed, device=device, auto_reset=True, finished_ring=cap)
    dev = env.device
    t0 = time.time()
    buf = torch.empty((n, 2492), dtype=dtype, device=dev)
    d = fold_state(env, torch.zeros(n, dtype=torch.int64, device=dev), torch.ones(n, dtype=torch.bool, device=dev))
    reports = torch.ones(n, dtype=torch.int64, device=dev)
    fin_digest, fin_reports, fin_env = [], [], []
    term_info = torch.empty((2 * n, 2492), dtype=torch.float32, device=dev)
    term_obs = torch.empty((2 * n, 98), dtype=torch.float32, device=dev)
    fused_rows_checked = 0
    steps = 0
    while int(env.finished_ctrl[0]) < episodes:
        env.rollout(1, _lib.PLAYER_CURRENT, out=buf)             # the benchmarked kernel, ring on
        steps += 1
        done = env.done.bool()
        # the row the fused kernel wrote for the player to move == that player's row of the both-views encoder
        info = env.information_state_tensor(_lib.PLAYER_BOTH)
        hb = env.tensor_row_hash(info).view(n, 2)
        hf = env.tensor_row_hash(buf)
        assert bool((hf == hb.gather(1, env.current_player.long().clamp(min=0).view(n, 1)).view(n)).all())
        fused_rows_checked += n
        ho = env.tensor_row_hash(env.observation_tensor(_lib.PLAYER_BOTH)).view(n, 2)
        del info
        # terminal states of this step's finished episodes, from the ring
        _, ids, cnt = env.finished_information_state_tensor(_lib.PLAYER_BOTH, out=term_info)
        _, ids2, cnt2 = env.finished_observation_tensor(_lib.PLAYER_BOTH, out=term_obs)
        k = int(cnt.item())
        assert k == int(cnt2.item()) == int(done.sum().item()) and bool((ids[:k] == ids2[:k]).all())
        ids = ids[:k].long()
        assert bool(done[ids].all()) and ids.unique().numel() == k
        hti = env.tensor_row_hash(term_info[: 2 * k]).view(k, 2)
        hto = env.tensor_row_hash(term_obs[: 2 * k]).view(k, 2)
        rew, ret = env.rewards.to(torch.int64)[ids], env.returns.to(torch.int64)[ids]
        f = d[ids]
        for v in (torch.zeros(k, dtype=torch.int64, device=dev), torch.full((k,), 252, dtype=torch.int64, device=dev),
                  torch.ones(k, dtype=torch.int64, device=dev), rew[:, 0] + 2, rew[:, 1] + 2, ret[:, 0] + 2, ret[:, 1] + 2,
                  hti[:, 0], hti[:, 1], hto[:, 0], hto[:, 1]):
            f = fold(f, v)
        fin_digest.append(f)
        fin_reports.append(reports[ids] + 1)
        fin_env.append(ids)
        # every env now sits at a decision node: fold it (finished envs start a new digest at their re-dealt state)
        base = torch.where(done, torch.zeros_like(d), d)
        zero = torch.zeros(n, dtype=torch.int64, device=dev)
        new = base
        new = fold(new, env.legal_mask.to(torch.int64) & 0xFFFFFFFF)
        new = fold(new, env.current_player.to(torch.int64) & 0xFF)
        new = fold(new, zero)
        rew, ret = env.rewards.to(torch.int64), env.returns.to(torch.int64)
        for t in (rew[:, 0], rew[:, 1], ret[:, 0], ret[:, 1]):
            new = fold(new, torch.where(done, zero, t) + 2)
        new = fold(fold(new, hb[:, 0]), hb[:, 1])
        d = fold(fold(new, ho[:, 0]), ho[:, 1])
        reports = torch.where(done, torch.ones_like(reports), reports + 1)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    records, dropped = env.finished_drain()
    stats = env.stats()
    rec = decode_finished_records(records)
    gd = torch.cat(fin_digest).cpu().numpy().view(np.uint64)
    gr = torch.cat(fin_reports).cpu().numpy()
    ge = torch.cat(fin_env).cpu().numpy()
    assert dropped == 0 and len(records) == len(gd) == stats["episodes"], (dropped, len(records), len(gd), stats["episodes"])
    assert (rec["env"] == ge).all()
    trajs = rec["trajectories"]
    flat = np.concatenate([a for a, _ in trajs])
    off = np.concatenate([[0], np.cumsum([len(a) for a, _ in trajs])]).astype(np.int64)
    t0 = time.time()
    dig, rep, bad = checker.replay_digest_batch(flat, off, threads)
    t_cpu = time.time() - t0
    mism = int((gd != dig).sum() + (gr != rep).sum()) + int(bad)
    lens = np.diff(off)
    meta_bad = int((rec["moves"] != lens).sum())
    return {"checker": "reference (oracle/_ref/libcoup_ref.so)" if ref else "oracle C port", "mode": "auto-reset, fused rollout kernel, finished-episode ring",
            "envs": int(n), "steps": steps, "episodes": int(len(gd)), "moves": int(lens.sum()), "reported_states": int(gr.sum()),
            "fused_rows_checked": int(fused_rows_checked), "mismatching_trajectories": mism, "rejected_by_checker": int(bad),
            "ring_meta_mismatches": meta_bad, "truncated_91_move_games": int((lens > 90).sum()),
            "ring_truncated_flags": int(rec["truncated"].sum()), "gpu_seconds": round(t_gpu, 2),
            "checker_seconds": round(t_cpu, 2), "checker_threads": threads,
            "total_mismatches": mism + meta_bad}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steered", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--autoreset-envs", type=int, default=1 << 17)
    ap.add_argument("--autoreset-episodes", type=int, default=1 << 20)
    args = ap.parse_args()
    res = check(args.envs, args.steered, args.seed, args.device, args.threads)
    if args.autoreset_episodes > 0:
        res["autoreset"] = check_autoreset(args.autoreset_envs, args.autoreset_episodes, args.seed + 17, args.device, args.threads)
        res["total_mismatches"] += res["autoreset"]["total_mismatches"]
    text = json.dumps(res, indent=1)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")
    sys.exit(0 if res["total_mismatches"] == 0 else 1)


if __name__ == "__main__":
    main()
