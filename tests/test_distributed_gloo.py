"""CPU tests of the N>1 path (world_size 2, gloo): env sharding and the one collective of this
path, the sum all-reduce of the statistics vector. The per-rank statistics are produced by the CPU
oracle rolling out the rank's own slab, so the reduced vector can be checked against a single-process run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from open_spiel_coup_b200 import _lib
from open_spiel_coup_b200.distributed import (barrier, init_process_group, max_over_ranks, reduce_stats, shard_envs,
                                              world_from_env)


def test_shards_partition_exactly():
    for total in (1, 2, 7, 1 << 20, (1 << 23) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_envs(total, world, r) for r in range(world)]
            assert spans[0][0] == 0
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1                      # contiguous, no overlap, no gap
            assert spans[-1][0] + spans[-1][1] == total
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    assert shard_envs(1 << 23, 8, 3) == (3 << 20, 1 << 20)   # BASELINE config 3: 2^23 envs over 8 GPUs
    with pytest.raises(ValueError):
        shard_envs(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _stats_from_oracle(count, seed):
    """Per-slab statistics vector in the library's layout, produced by the CPU oracle."""
    from oracle.bindings import Oracle
    r = Oracle().bench(0, 1, count, seed)
    v = np.zeros(_lib.STATS_LEN, np.int64)
    v[_lib.STAT_DECISION_STEPS] = r["decisions"]
    v[_lib.STAT_CHANCE_MOVES] = r["chance"]
    v[_lib.STAT_EPISODES] = r["episodes"]
    v[_lib.STAT_TRUNCATED] = r["truncated"]
    v[_lib.STAT_EPISODE_MOVES] = r["moves"]
    v[_lib.STAT_RETURN_HIST:_lib.STAT_RETURN_HIST + 5] = r["returns_hist_p0"]
    v[_lib.STAT_LEGAL_HIST:_lib.STAT_LEGAL_HIST + 8] = r["legal_count_hist"]
    return v


def _worker(rank, world, port, total, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, local, w = init_process_group(backend="gloo")
    assert (r, w) == (rank, world) and world_from_env() == (rank, rank, world)
    offset, count = shard_envs(total, world, rank)
    mine = torch.from_numpy(_stats_from_oracle(count, seed=1000 + offset))
    barrier()
    summed = reduce_stats(mine)
    assert torch.equal(mine, torch.from_numpy(_stats_from_oracle(count, seed=1000 + offset)))  # input untouched
    slowest = max_over_ranks(10.0 + rank)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.concatenate([summed.numpy(), [int(slowest)], mine.numpy()]))
    dist.destroy_process_group()


def test_stats_allreduce_world2_gloo(tmp_path):
    total, world = 3001, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    L = _lib.STATS_LEN
    assert (got[0][:L] == got[1][:L]).all(), "every rank must end with the same reduced vector"
    assert got[0][L] == 11 and got[1][L] == 11            # max over ranks of (10, 11)
    expected = sum(g[L + 1:] for g in got)
    assert (got[0][:L] == expected).all()
    assert got[0][_lib.STAT_EPISODES] == total             # slabs covered every env exactly once
    assert got[0][_lib.STAT_RETURN_HIST:_lib.STAT_RETURN_HIST + 5].sum() == total
    assert got[0][_lib.STAT_DECISION_STEPS] + got[0][_lib.STAT_CHANCE_MOVES] == got[0][_lib.STAT_EPISODE_MOVES]


def test_single_process_is_a_noop():
    t = torch.arange(_lib.STATS_LEN, dtype=torch.int64)
    assert torch.equal(reduce_stats(t), t)
    assert max_over_ranks(3.5) == 3.5
    barrier()
