"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the row-range sharding and the
collectives bench.py / a multi-GPU deployment use (the kernels themselves need a GPU; here the
per-rank partial results come from numpy so only the plumbing is under test)."""
import os
import socket

import numpy as np
import pytest

from arrow_go_b200 import sharding


def test_shard_ranges_cover_and_align():
    for n in (0, 1, 63, 64, 65, 1000, 100_000_000, 1_000_000_007):
        for world in (1, 2, 4, 8):
            r = sharding.all_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            for (a, b), (c, d) in zip(r, r[1:]):
                assert b == c and a <= b
            for a, b in r[:-1]:
                assert (a % 64 == 0 or a == n) and (b % 64 == 0 or b == n)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1_000_003
    rng = np.random.default_rng(0x94378165)
    xi = rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, n, dtype=np.int64, endpoint=True)
    xf = rng.integers(-(1 << 20), 1 << 20, n).astype(np.float64)
    mask = rng.random(n) < 0.1
    a, b = sharding.shard_range(n, rank, world)
    with np.errstate(over="ignore"):
        local_i = int(xi[a:b].sum(dtype=np.int64))
    gi = sharding.global_sum_int(local_i, dist)
    gf = sharding.global_sum_float(float(xf[a:b].sum()), dist)
    off, total = sharding.filter_output_offsets(int(mask[a:b].sum()), dist)
    with np.errstate(over="ignore"):
        want_i = int(xi.sum(dtype=np.int64))
    ok = (gi == want_i) and (gf == float(xf.sum())) and (off == int(mask[:a].sum())) and (total == int(mask.sum()))
    # min/max and the cumulative-sum carry: per-rank partials from numpy, plumbing under test
    lo, hi = sharding.global_min_max(int(xi[a:b].min()), int(xi[a:b].max()), dist)
    ok = ok and (lo, hi) == (int(xi.min()), int(xi.max()))
    valid = np.ones(n, dtype=bool)
    valid[700_001] = False                      # one null, in the last shard for world=2
    for skip in (False, True):
        v = valid[a:b]
        first = int(np.argmin(v)) if not v.all() else len(v)
        contrib = xi[a:b][v] if skip else xi[a:b][:first]
        with np.errstate(over="ignore"):
            local_total = int(contrib.sum(dtype=np.int64))
        start, dead = sharding.cumulative_sum_carry(local_total, not v.all(), dist, skip_nulls=skip)
        with np.errstate(over="ignore"):
            want_start = int(xi[:a][valid[:a]].sum(dtype=np.int64)) if skip else int(xi[:min(a, 700_001)].sum(dtype=np.int64))
        want_dead = (not skip) and (not valid[:a].all())
        ok = ok and start == want_start and dead == want_dead
    q.put((rank, ok, gi, want_i))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] for r in res), res
