#!/usr/bin/env python
"""bench.py -- Coup env steps/sec (BASELINE.json metric) on N B200s, one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W     # the reference's CPU implementation
  torchrun --nproc-per-node N ... bench.py --gpus N ...     # envs sharded, NCCL only for the stats reduce

Workload (BASELINE.json configs[1]): 2-player Coup, uniform-random rollouts, 2^20 envs per GPU, every
step = sample a legal action, apply it, resolve the following chance nodes, auto-reset finished
episodes, emit legal mask / current player / rewards / done and the dense fp32 2492-float
info-state tensor of the player to move (the reference's benchmark_game.cc:53-60 protocol).
One "step" of this script = one such pass over all envs of a rank (one kernel launch).

`value`  : decision steps/s, tensors and state resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : the same metric through the host-buffer C-ABI call (coup_vec_step_host_packed) on 8-16 sub-slabs
           with a stream each: actions come from pinned HOST memory every step, one step word per env (legal
           mask, current player, done, reward, return) goes back to pinned HOST memory every step, the host
           policy (coup_host_sample_uniform) turns the words into the next actions, and the info-state tensor
           is encoded every step and stays device-resident for the on-device consumer (the policy network of
           north_star config 4).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "coup_env_steps_per_sec"
UNIT = "env steps/s"
# Algorithmic bytes per decision step (restated in DESIGN.md):
#   env-only 42 B = state 16 R + 16 W, legal mask 4, action 1, rewards+done+cur_player 4, log append 1.41
#   + dense info-state row (2492 elements) + ~11 B of history read by the encoder
BYTES_PER_STEP = {"d32": 42 + 9968 + 11, "bf16": 42 + 4984 + 11, "d8": 42 + 2492 + 11, "env": 42}
# incremental contract (persistent [N,2,2492] fp32 buffer, only changed elements rewritten): SURVEY.md section 8(d)
BYTES_PER_STEP_INCREMENTAL_F32 = 945
CONTRACT_DTYPE = {"d32": "f32", "bf16": "bf16", "d8": "u8", "env": "u32"}


def workload_text(contract):
    return ("2-player Coup uniform-random rollouts, 2^20 envs per GPU, legal mask + dense info-state tensor "
            "(2492 x %s, player to move) per decision step, auto-reset, Philox4x32-10 chance (BASELINE.json configs[1])"
            % CONTRACT_DTYPE[contract])


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--contract", default="d32", choices=list(BYTES_PER_STEP))
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-contracts", action="store_true")
    ap.add_argument("--no-selfplay", action="store_true")
    ap.add_argument("--no-host-tensor", action="store_true")
    ap.add_argument("--e2e-slabs", type=int, default=0,
                    help="sub-slabs (handles/streams) of the host-buffer leg; 0 = 16 with >= 12 host-policy threads per rank, else 8")
    ap.add_argument("--host-threads", type=int, default=0, help="threads of the host policy per rank (0 = host cores / ranks)")
    ap.add_argument("--blocking-sync", action="store_true", help="host-buffer calls sleep instead of spinning while they wait")
    ap.add_argument("--fused", default="ws", choices=["ws", "cta"],
                    help="fused rollout kernel: persistent warp-specialised (rules warps + encoder warps), or one CTA per 256 envs")
    ap.add_argument("--ring", type=int, default=1 << 18,
                    help="capacity (records, power of two) of the finished-episode ring the timed kernels append to; 0 = off")
    ap.add_argument("--encoder", default="staged", choices=["staged", "plain"],
                    help="info-state encoder: shared-memory staging + bulk (TMA) stores, or per-lane vector stores")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(contract):
    """Per-launch dram bytes of the dominant kernel from the committed ncu capture, or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return d.get(contract)
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed regions run."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows = []
        self.proc = None
        self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) == 9:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        loaded = [r for r in self.rows if r[4].isdigit() and int(r[4]) >= 50] or self.rows
        sm = [int(r[1]) for r in loaded if r[1].isdigit()]
        smax = [int(r[2]) for r in self.rows if r[2].isdigit()]
        reasons = []
        for i, name in ((5, "hw_slowdown"), (6, "hw_thermal_slowdown"), (7, "sw_thermal_slowdown"), (8, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in loaded):
                reasons.append(name)
        power = [float(r[3]) for r in loaded if r[3].replace(".", "", 1).isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": reasons, "samples": len(self.rows), "samples_under_load": len(loaded),
                "power_w_max": max(power) if power else None}


def cpu_reference_run(mode, threads, seconds, seed=1234):
    """Times the reference's CPU implementation (oracle/_ref if it was built, else the C port) on a
    bounded sample sized to ~`seconds`. Returns (dict for the JSON line, raw bench dict)."""
    from oracle.bindings import Oracle, Reference
    impl = Reference() if Reference.available() else Oracle()
    kind = "reference" if Reference.available() else "port"
    probe = impl.bench(mode, threads, 2000 * threads, seed)
    eps = max(2000 * threads, int(probe["episodes_per_s"] * seconds))
    res = impl.bench(mode, threads, eps, seed + 1)
    sample = (f"{res['episodes']} uniform-random episodes ({res['decisions']} decision steps, {res['seconds']:.1f} s), "
              f"LegalActions + InformationStateTensor(current player) at every decision node "
              f"(open_spiel/examples/benchmark_game.cc protocol), one State + std::mt19937 per thread, -O3 -DNDEBUG"
              + ("" if kind == "reference" else "; C restatement of the reference (oracle/_ref not built)"))
    out = {"value": res["decisions_per_s"], "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
           "moves_per_s": res["moves_per_s"], "episodes_per_s": res["episodes_per_s"]}
    # The other protocols of SURVEY.md section 8(d) / BASELINE.md section 3, on short samples (decision steps/s):
    # mode 0 = step + LegalActions only, mode 2 = info-state of BOTH players (rl_environment semantics), and
    # BASELINE.json configs[0]: 10^4 episodes on ONE thread with the benchmark_game.cc protocol.
    extra = {}
    for name, m, th, eps in (("step_and_legal_actions_only", 0, threads, int(probe["episodes_per_s"] * 6)),
                             ("info_state_both_players", 2, threads, int(probe["episodes_per_s"] * 1.5)),
                             ("config0_single_thread_10k_episodes", 1, 1, 10000)):
        r = impl.bench(m, th, max(eps, 1000 * th), seed + 7)
        extra[name] = {"steps_per_s": r["decisions_per_s"], "moves_per_s": r["moves_per_s"], "threads": th,
                       "episodes": r["episodes"], "seconds": round(r["seconds"], 2)}
    out["other_protocols"] = extra
    return out, res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle.bindings import Oracle, Reference
    impl = Reference() if Reference.available() else Oracle()
    kind = "reference" if Reference.available() else "port"
    probe = impl.bench(1, threads, 2000 * threads, args.seed)
    # each "step" is a bounded sample of the workload: ~0.5 s of all host threads
    eps_per_step = max(1000 * threads, int(probe["episodes_per_s"] * 0.5))
    steps = max(1, min(args.steps, 40))
    for i in range(min(args.warmup, 3)):
        impl.bench(1, threads, eps_per_step, args.seed + 10 + i)
    dec = secs = 0.0
    for i in range(steps):
        r = impl.bench(1, threads, eps_per_step, args.seed + 100 + i)
        dec += r["decisions"]
        secs += r["seconds"]
    value = dec / secs
    sample = (f"{steps} samples x {eps_per_step} uniform-random episodes, LegalActions + InformationStateTensor(current player) "
              f"per decision node (benchmark_game.cc protocol), {threads} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 3), "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text("d32"), "envs_per_gpu": args.envs, "contract": "d32",
                   "bytes_per_step": BYTES_PER_STEP["d32"],
                   "reference_arm": "the reference's CoupState on %d host threads: uniform-random legal actions, SampleAction on "
                                    "ChanceOutcomes, LegalActions + InformationStateTensor(current player) at every decision node "
                                    "(examples/benchmark_game.cc protocol); each step = a bounded sample of %d episodes" % (threads, eps_per_step)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from open_spiel_coup_b200 import _lib
    from open_spiel_coup_b200.vector_env import CoupVectorEnv

    from open_spiel_coup_b200.distributed import reduce_stats, shard_envs, world_from_env
    rank, local, world = world_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the Coup environment has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, K, W = args.envs, args.steps, max(args.warmup, 3)
    slab_offset, slab_count = shard_envs(n * world, world, rank)   # weak scaling: n envs per GPU
    assert slab_count == n
    contract = args.contract
    torch_dtype = {"d32": torch.float32, "bf16": torch.bfloat16, "d8": torch.uint8, "env": None}[contract]
    env = CoupVectorEnv(n, seed=args.seed, device=local, global_env_offset=slab_offset, auto_reset=True,
                        plain_store_encoder=(args.encoder == "plain"), warp_specialised=(args.fused == "ws"),
                        finished_ring=args.ring)
    out = None if torch_dtype is None else torch.empty((n, 2492), dtype=torch_dtype, device=dev)
    sel = None if out is None else _lib.PLAYER_CURRENT

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """CUDA-event time of `steps` calls of fn on the current stream, max over ranks (ms)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # Desynchronise episodes (100 untimed env-only steps), then W warm-up steps of the measured kind.
    env.rollout(100)
    env.rollout(W, sel, out=out)
    env.clear_stats()
    ms = timed(lambda k: env.rollout(k, sel, out=out), K)
    stats = reduce_stats(env.stats_device)               # the one collective: NCCL sum of the stats vector
    stats = [int(x) for x in stats.cpu()]
    steps_done = stats[_lib.STAT_DECISION_STEPS]
    assert steps_done == K * n * world and stats[_lib.STAT_ILLEGAL] == 0, (steps_done, K * n * world, stats[_lib.STAT_ILLEGAL])
    value = steps_done / (ms * 1e-3)
    per_launch_s = ms * 1e-3 / K

    # Context for the roofline: what a pure zero-fill of the same output buffer reaches on this GPU right now (the
    # driver's peak is a COPY, half reads; this path only writes). Not a denominator, just reported beside `frac`.
    def zero_fill_gbs(t):
        # the same BYTES zero-filled as 32-bit words (torch's fill of 1- and 2-byte elements is itself far from the
        # write ceiling: 3.9 TB/s on this buffer shape)
        flat = t.view(-1).view(torch.int32) if (t.numel() * t.element_size()) % 4 == 0 else t
        flat.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            flat.zero_()
        e1.record()
        torch.cuda.synchronize()
        return t.numel() * t.element_size() * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9

    write_only_gbs = None
    if out is not None and rank == 0:
        write_only_gbs = zero_fill_gbs(out)

    # Other contracts on the same slab (short runs), reported beside the headline.
    extra = {}
    if not args.no_extra_contracts:
        ke = max(20, min(K, 200))
        for name, dt in (("env", None), ("d8", torch.uint8), ("bf16", torch.bfloat16), ("d32", torch.float32)):
            if name == contract:
                extra[name] = {"steps_per_s": value, "hbm_gbs": BYTES_PER_STEP[name] * value / world / 1e9}
                continue
            buf = None if dt is None else torch.empty((n, 2492), dtype=dt, device=dev)
            s2 = None if dt is None else _lib.PLAYER_CURRENT
            env.rollout(3, s2, out=buf)
            kk = ke * (10 if dt is None else 1)
            m2 = timed(lambda k: env.rollout(k, s2, out=buf), kk)
            v2 = kk * n * world / (m2 * 1e-3)
            extra[name] = {"steps_per_s": v2, "hbm_gbs": BYTES_PER_STEP[name] * v2 / world / 1e9}
            if buf is not None and rank == 0:      # the write ceiling for THIS buffer size (smaller buffers fill a little slower)
                fill = zero_fill_gbs(buf)
                extra[name].update(write_only_fill_gbs=fill, frac_of_write_only_fill=extra[name]["hbm_gbs"] / fill)
            del buf
            if dt is None and args.ring:
                # the same env-only rollout without the finished-episode ring (episodes' logs are lost at the re-deal)
                env.enable_finished_ring(0)
                env.rollout(3)
                m2 = timed(lambda k: env.rollout(k), kk)
                v2 = kk * n * world / (m2 * 1e-3)
                extra["env_no_ring"] = {"steps_per_s": v2, "hbm_gbs": BYTES_PER_STEP["env"] * v2 / world / 1e9}
                env.enable_finished_ring(args.ring)
        # live-prefix contract: fp32 rows of 1728 elements = elements [0, 1728) of the reference row; everything they drop
        # (history rows 91..134) is zero in every reachable state (COUP_LIVE_INFO_STATE_SIZE, include/coup_b200.h)
        for name, dt in (("d32_live_prefix", torch.float32), ("bf16_live_prefix", torch.bfloat16), ("d8_live_prefix", torch.uint8)):
            buf = torch.empty((n, _lib.LIVE_INFO_STATE_SIZE), dtype=dt, device=dev)
            live_bytes = 42 + buf.element_size() * _lib.LIVE_INFO_STATE_SIZE + 11
            env.rollout(3, _lib.PLAYER_CURRENT, out=buf)
            m2 = timed(lambda k: env.rollout(k, _lib.PLAYER_CURRENT, out=buf), ke)
            v2 = ke * n * world / (m2 * 1e-3)
            extra[name] = {"steps_per_s": v2, "hbm_gbs": live_bytes * v2 / world / 1e9, "bytes_per_step": live_bytes,
                           "note": "rows cut after element 1727: the dropped columns are always zero"}
            if rank == 0:
                fill = zero_fill_gbs(buf)
                extra[name].update(write_only_fill_gbs=fill, frac_of_write_only_fill=extra[name]["hbm_gbs"] / fill)
            del buf
        # contract I: persistent fp32 buffer with both views of every env, updated in place
        if out is not None:
            del out
            torch.cuda.empty_cache()
        inc = torch.empty((2 * n, 2492), dtype=torch.float32, device=dev)
        env.information_state_tensor(_lib.PLAYER_BOTH, out=inc)
        env.rollout_incremental(5, inc)
        m2 = timed(lambda k: env.rollout_incremental(k, inc), ke)
        v2 = ke * n * world / (m2 * 1e-3)
        extra["incremental_f32_both_views"] = {"steps_per_s": v2, "hbm_gbs": BYTES_PER_STEP_INCREMENTAL_F32 * v2 / world / 1e9,
                                               "bytes_per_step": BYTES_PER_STEP_INCREMENTAL_F32}
        del inc
        torch.cuda.empty_cache()

        # ---- every other kernel of SURVEY section 8(a), one at a time: algorithmic bytes per env and the fraction of
        # the HBM peak it reaches. Rows = envs (x2 for both views). Kernels whose working set fits the 126 MB L2 say so.
        peak_gbs, _ = measured_peak()
        kern = {}

        def add(name, kernel, fn, bytes_per_env, reps, note=None):
            # warm-up of ~30 ms: the allocations and frees between the legs leave the GPU idle for milliseconds, and a short
            # kernel timed right after such a gap can catch the clocks still ramping (80 vs 99 us for the first leg)
            t_w = time.perf_counter()
            while time.perf_counter() - t_w < 0.03:
                for _ in range(10):
                    fn()
                torch.cuda.synchronize()
            msk = timed(lambda k: [fn() for _ in range(k)], reps)
            per_s = reps * n * world / (msk * 1e-3)
            gbs = bytes_per_env * per_s / world / 1e9
            kern[name] = {"kernel": kernel, "envs_per_s": per_s, "us_per_launch": 1e3 * msk / reps, "bytes_per_env": bytes_per_env,
                          "hbm_gbs": gbs, "frac": gbs / peak_gbs}
            if note:
                kern[name]["note"] = note

        for dt, tag, es in ((torch.float32, "f32", 4), (torch.uint8, "u8", 1)):
            for sel2, views, vt in ((_lib.PLAYER_CURRENT, 1, "one_view"), (_lib.PLAYER_BOTH, 2, "both_views")):
                o = torch.empty((views * n, 98), dtype=dt, device=dev)
                add(f"obs98_{tag}_{vt}", "k_encode_obs", lambda: env.observation_tensor(sel2, out=o), 16 + views * 98 * es, 50,
                    None if views * n * 98 * es > (1 << 28) else "output %d MB: partly L2-resident" % (views * n * 98 * es >> 20))
                del o
        for dt, tag, es in ((torch.float32, "f32", 4), (torch.uint8, "u8", 1)):
            o = torch.empty((2 * n, 2492), dtype=dt, device=dev)
            add(f"info_state_{tag}_both_views", "k_encode_info_tma", lambda: env.information_state_tensor(_lib.PLAYER_BOTH, out=o),
                80 + 2 * 2492 * es, 20)
            del o
        o = torch.empty((n, 2492), dtype=torch.float32, device=dev)
        acts = torch.empty(n, dtype=torch.uint8, device=dev)

        def step_pair():
            env.sample_uniform(out=acts)
            env.step(acts)
            env.information_state_tensor(_lib.PLAYER_CURRENT, out=o)
        add("sample_uniform+step+encode_f32", "k_sample_uniform, k_step, k_encode_info_tma (the unfused form of the headline step, "
            "what coup_vec_step_host_packed launches)", step_pair, 17 + 42 + 80 + 9968, 50)
        del o

        def step_only():
            env.sample_uniform(out=acts)
            env.step(acts)
        add("sample_uniform+step", "k_sample_uniform, k_step", step_only, 17 + 42, 200,
            "state + history + outputs = 100 MB: L2-resident; instruction bound")
        logits = torch.randn((n, 18), device=dev)
        probs = torch.empty((n, 18), device=dev)
        add("sample_policy_f32", "k_sample_policy", lambda: env.sample_policy(logits, probs_out=probs, actions_out=acts), 72 + 4 + 1 + 72, 200,
            "logits + probs = 151 MB: about L2 size")
        add("sample_policy_f32_no_probs", "k_sample_policy", lambda: env.sample_policy(logits, actions_out=acts), 72 + 4 + 1, 200,
            "L2-resident")
        dense = torch.empty((n, 18), dtype=torch.uint8, device=dev)
        add("legal_actions_mask", "k_legal_actions_mask", lambda: env.legal_actions_mask(out=dense), 4 + 18, 200, "23 MB: L2-resident")
        del logits, probs, dense, acts
        extra["kernels"] = kern
        torch.cuda.empty_cache()
        out = None if torch_dtype is None else torch.empty((n, 2492), dtype=torch_dtype, device=dev)

    # ---- end to end through the host-buffer C-ABI call -------------------------------------------
    # The slab is driven as `args.e2e_slabs` sub-slabs, each with its own handle, stream and pinned host
    # buffers, so that the host-side policy and the PCIe copies of one sub-slab overlap the encoder
    # of the other (what a host-driven caller does to keep the GPU busy). Same total number of envs.
    lib = _lib.load()
    threads = args.host_threads or max(1, (os.cpu_count() or 1) // world)
    # Measured (profiles/README.md): finer sub-slabs hide the host policy and the PCIe round trip better (1 GPU, 16 threads:
    # 6.55e8 / 6.67e8 / 6.80e8 steps/s with 4 / 8 / 16), until the per-call cost of waking a small thread pool shows
    # (8 GPUs, 4 threads per rank: 4.77e9 / 5.28e9 / 4.44e9).
    S_ = args.e2e_slabs if args.e2e_slabs > 0 else (16 if threads >= 12 else 8)
    ns = n // S_
    del env
    torch.cuda.empty_cache()
    slabs = []
    for i in range(S_):
        ev = CoupVectorEnv(ns, seed=args.seed, device=local, global_env_offset=slab_offset + i * ns, auto_reset=True,
                           plain_store_encoder=(args.encoder == "plain"), blocking_sync=args.blocking_sync,
                           finished_ring=max(0, args.ring // S_))
        ev.rollout(100)
        h_act = torch.empty(ns, dtype=torch.uint8).pin_memory()
        h_words = torch.empty(ns, dtype=torch.int32).pin_memory()
        h_words.copy_(ev.step_word)
        t_out = None if out is None else out[i * ns:(i + 1) * ns]
        slabs.append((ev, h_act, h_words, t_out, torch.cuda.Stream(device=dev), slab_offset + i * ns))
    torch.cuda.synchronize()

    def e2e_steps(k):
        # One host thread drives all sub-slabs: a sub-slab's step is queued without waiting (step_host_packed_async), and
        # its step words are awaited only when its next actions are about to be chosen, one round later.
        for _ in range(k):
            for ev, h_act, h_words, t_out, stream, offset in slabs:
                ev.host_outputs_wait()
                rc = lib.coup_host_sample_uniform(C.c_void_p(h_words.data_ptr()), ns, args.seed, offset,
                                                  ev.step_counter, C.c_void_p(h_act.data_ptr()), threads)
                assert rc == 0
                ev.step_host_packed_async(h_act, h_words, tensor_out=t_out, stream=stream)
        for sl in slabs:
            sl[0].host_outputs_wait()

    Ke = max(3, min(K, 200))
    e2e_steps(3)
    for sl in slabs:
        sl[0].clear_stats()
    barrier()
    t0 = time.perf_counter()
    e2e_steps(Ke)
    torch.cuda.synchronize(dev)     # the encoders of the last step may still be running on the sub-slab streams
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    st2 = reduce_stats(sum(sl[0].stats_device.clone() for sl in slabs))
    st2 = [int(x) for x in st2.cpu()]
    assert st2[_lib.STAT_DECISION_STEPS] == Ke * ns * S_ * world and st2[_lib.STAT_ILLEGAL] == 0
    e2e_value = st2[_lib.STAT_DECISION_STEPS] / float(e2e_s.item())

    # Secondary end-to-end number for a HOST-side consumer: additionally copy the whole (uint8, exact) info-state
    # tensor to pinned host memory every step. This is PCIe-bound by construction (2492 B per step over ~55 GB/s).
    e2e_host_tensor = None
    if not args.no_host_tensor and rank == 0 and out is not None:
        try:
            h_tensor = torch.empty((n, 2492), dtype=torch.uint8).pin_memory()
            d_u8 = [torch.empty((ns, 2492), dtype=torch.uint8, device=dev) for _ in slabs]

            def host_tensor_steps(k):
                for _ in range(k):
                    for i, (ev, h_act, h_words, t_out, stream, offset) in enumerate(slabs):
                        lib.coup_host_sample_uniform(C.c_void_p(h_words.data_ptr()), ns, args.seed, offset, ev.step_counter,
                                                     C.c_void_p(h_act.data_ptr()), threads)
                        ev.step_host_packed(h_act, h_words, tensor_out=d_u8[i], stream=stream)
                        with torch.cuda.stream(stream):
                            h_tensor[i * ns:(i + 1) * ns].copy_(d_u8[i], non_blocking=True)
                    torch.cuda.synchronize()

            host_tensor_steps(1)
            t0 = time.perf_counter()
            host_tensor_steps(3)
            dt = time.perf_counter() - t0
            e2e_host_tensor = {"value": 3 * ns * S_ / dt, "unit": UNIT, "dtype": "u8", "d2h_bytes_per_step": (2492 + 4) * ns * S_,
                               "d2h_gb_per_s": 3 * (2492 + 4) * ns * S_ / dt / 1e9, "steps": 3,
                               "note": "same loop, plus the full uint8 info-state tensor copied to pinned host memory every step (rank 0 only)"}
            del h_tensor, d_u8
            # the same with the LIVE PREFIX of every row (1728 of the 2492 columns; the rest is zero in every reachable
            # state, COUP_LIVE_INFO_STATE_SIZE): 31 % fewer bytes over PCIe
            live = _lib.LIVE_INFO_STATE_SIZE
            h_tensor = torch.empty((n, live), dtype=torch.uint8).pin_memory()
            d_u8 = [torch.empty((ns, live), dtype=torch.uint8, device=dev) for _ in slabs]

            def host_live_steps(k):
                for _ in range(k):
                    for i, (ev, h_act, h_words, t_out, stream, offset) in enumerate(slabs):
                        lib.coup_host_sample_uniform(C.c_void_p(h_words.data_ptr()), ns, args.seed, offset, ev.step_counter,
                                                     C.c_void_p(h_act.data_ptr()), threads)
                        ev.step_host_packed(h_act, h_words, tensor_out=None, stream=stream)
                        with torch.cuda.stream(stream):
                            ev.information_state_tensor(_lib.PLAYER_CURRENT, out=d_u8[i])
                            h_tensor[i * ns:(i + 1) * ns].copy_(d_u8[i], non_blocking=True)
                    torch.cuda.synchronize()

            host_live_steps(1)
            t0 = time.perf_counter()
            host_live_steps(3)
            dt = time.perf_counter() - t0
            e2e_host_tensor["live_prefix"] = {"value": 3 * ns * S_ / dt, "unit": UNIT, "dtype": "u8", "columns": live,
                                              "d2h_bytes_per_step": (live + 4) * ns * S_,
                                              "d2h_gb_per_s": 3 * (live + 4) * ns * S_ / dt / 1e9, "steps": 3}
            del h_tensor, d_u8
        except Exception as exc:       # pinned allocation can fail on small hosts; the primary e2e number stands
            e2e_host_tensor = {"unavailable": str(exc)[:200]}

    # ---- BASELINE configs[3]: self-play data generation, MLP policy on the info-state tensor, 2^18 envs ----
    selfplay = None
    if not args.no_selfplay:
        from open_spiel_coup_b200.selfplay import MLPPolicy, SelfPlayDataGen
        for sl in slabs:
            sl[0].close()
        del slabs, out
        torch.cuda.empty_cache()
        torch.manual_seed(args.seed)
        n_sp = 1 << 18
        # Data generation, i.e. WITH the records the reference's agents keep (NFSP reservoir + DQN replay, thesis capacities
        # scaled to the batch: a step offers 2^18 decisions): kept by the step kernel as packed 96-byte records.
        cap_res, cap_rb = 1 << 22, 1 << 22
        gen = SelfPlayDataGen(num_envs=n_sp, policy=None, seed=args.seed, device=local, global_env_offset=rank * n_sp,
                              reservoir_capacity=cap_res, replay_capacity=cap_rb)
        gen.env.rollout(100)
        gen.run(20)                      # fills the reservoir (16 steps) and gets past its start-up phase
        k_sp = max(10, min(K, 60))
        ms_sp = timed(lambda k: gen.run(k), k_sp)
        # the recording step alone (k_reservoir_claim + k_step_record) against a plain k_step on the same slab
        acts, probs = gen.actions, gen.action_probs
        ms_rec = timed(lambda k: [gen.recorder.step(acts, probs) for _ in range(k)], 50)
        ms_plain = timed(lambda k: [gen.env.step(acts) for _ in range(k)], 50)
        # algorithmic bytes per env of the recording step: step itself 42; replay: history row 64 R, pending record 96 R + 96 W,
        # one transition per decision 192 W; reservoir (full, t >> capacity): ~0; claims 8
        rec_bytes = 42 + 64 + 96 + 96 + 192 + 8
        peak_gbs, _ = measured_peak()
        selfplay = {"steps_per_s": k_sp * n_sp * world / (ms_sp * 1e-3), "envs_per_gpu": n_sp, "ms_per_step": ms_sp / k_sp,
                    "policy": "MLP 2492(+4 zero pad)-1024-1024-18 (bf16, torch; the encoder writes and the first layer reads the 1728 live columns), masked softmax sampling fused on device (k_sample_policy)",
                    "recording": "NFSP reservoir (%d) + DQN replay (%d) kept by the step kernel as packed records (coup_vec_step_record)" % (cap_res, cap_rb),
                    "recording_step": {"kernel": "k_reservoir_claim + k_step_record", "us_per_launch": 1e3 * ms_rec / 50,
                                       "plain_k_step_us": 1e3 * ms_plain / 50, "bytes_per_env": rec_bytes,
                                       "hbm_gbs": rec_bytes * n_sp / (ms_rec / 50 * 1e-3) / 1e9,
                                       "frac": rec_bytes * n_sp / (ms_rec / 50 * 1e-3) / 1e9 / peak_gbs,
                                       "note": "slab + records of 2^18 envs (~90 MB) are L2-resident"},
                    "note": "per step: encode bf16 info-state of the player to move, policy forward, masked sampling, "
                            "env step + reservoir + replay records"}
        del gen
        torch.cuda.empty_cache()

    clocks = sampler.stop() if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference_run(1, os.cpu_count() or 1, args.cpu_seconds, args.seed)

    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_per_launch = BYTES_PER_STEP[contract] * n
        achieved = bytes_per_launch / per_launch_s / 1e9
        kname = ("k_rollout_ws<%s>" if args.fused == "ws" else "k_rollout_tma<%s>") if args.encoder == "staged" else "k_rollout<%s,true>"
        kernel = {"d32": kname % "float", "bf16": kname % "__nv_bfloat16", "d8": kname % "uint8_t",
                  "env": "k_rollout_env_multi"}[contract]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": CONTRACT_DTYPE[contract], "data": "synthetic",
            "config": {
                "workload": workload_text(contract),
                "envs_per_gpu": n, "contract": contract,
                "finished_ring": ("every finished episode's action+chance log and terminal state appended to a %d-record device ring "
                                  "inside the timed kernels (auto-reset re-deals in place)" % args.ring) if args.ring else "off", "bytes_per_step": BYTES_PER_STEP[contract], "encoder": args.encoder, "fused_kernel": args.fused,
                "parallelism": f"env-slab x{world} (no data-path collective; NCCL all-reduce of the stats vector only)",
                "l2": "per-step working set (%.2f GB written + 80 MB state/history) >> 126 MB L2, no flush needed" % (bytes_per_launch / 1e9)
                      if contract != "env" else "state+history 80 MiB/GPU is L2-resident by design (env-only contract)",
            },
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(contract), "peak_source": peak_src,
                         "bytes_per_launch": bytes_per_launch, "launch_ms": per_launch_s * 1e3,
                         "write_only_fill_gbs": write_only_gbs,
                         "frac_of_write_only_fill": None if not write_only_gbs else achieved / write_only_gbs},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ns * S_, "d2h_bytes_per_step": 4 * ns * S_,
                    "steps": Ke, "host_policy_threads": threads, "sub_slabs": S_,
                    "note": "coup_vec_step_host_packed on %d sub-slabs (own stream each): uint8 actions from pinned host memory in, "
                            "one uint32 step word per env (legal mask, current player, done, reward, return) to pinned host memory out, "
                            "every step; the host policy (coup_host_sample_uniform) picks the next actions from those words; the "
                            "info-state tensor is encoded every step and left in HBM for the on-device consumer" % S_},
            "e2e_tensor_to_host": e2e_host_tensor,
            # one launch per step with a tensor contract; the env-only contract runs up to 64 steps per launch
            "gpu_launches": K if contract != "env" else -(-K // 64),
            "clocks": clocks,
            "contracts": extra,
            "selfplay": selfplay,
            "episodes": stats[_lib.STAT_EPISODES], "moves_per_episode": stats[_lib.STAT_EPISODE_MOVES] / max(1, stats[_lib.STAT_EPISODES]),
            "moves_per_s": value * (stats[_lib.STAT_DECISION_STEPS] + stats[_lib.STAT_CHANCE_MOVES]) / max(1, stats[_lib.STAT_DECISION_STEPS]),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
