"""Multi-GPU surface of the C ABI: single-process multi-device runtime, row-range sharding, and the
global Sum whose cross-GPU fold is fused into the reduction kernel (HBM mailboxes over NVLink).

The mailbox protocol does not need two GPUs to be exercised: ag_comm_create_local accepts the same device twice, which
gives two ranks whose kernels meet through the same mailboxes on one GPU (different streams).  The tests that really
cross NVLink skip on a one-GPU box; the driver's 8-GPU scaling run goes through the same entry points (bench.py)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from arrow_go_b200 import sharding
from helpers import Dev

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_matches_python_rule():
    """CPU: the C ABI's sharding rule is the one sharding.py (and the gloo tests) use."""
    for n in (0, 1, 63, 64, 65, 1000, 100_000_000, 1_000_000_007):
        for world in (1, 2, 3, 8):
            for r in range(world):
                a, b = C.c_int64(), C.c_int64()
                N.call("ag_shard_range", n, r, world, C.byref(a), C.byref(b))
                assert (a.value, b.value) == sharding.shard_range(n, r, world)
    assert N.call_status("ag_shard_range", 10, 2, 2, C.byref(a), C.byref(b))[0] == N.AG_ERR_INVALID


def _device_count():
    c = C.c_int()
    N.call("ag_device_count", C.byref(c))
    return c.value


def _local_comms(devices):
    comms = (C.c_void_p * len(devices))()
    N.call("ag_comm_create_local", comms, len(devices), (C.c_int * len(devices))(*devices))
    return [C.c_void_p(c) for c in comms]


def _global_sums(ag, devices, shards_i64, shards_f64):
    """Every rank launches its global Sum (async), then all are synchronised; returns per-rank results."""
    world = len(devices)
    comms = _local_comms(devices)
    streams, bufs, outs = [], [], []
    for k, d in enumerate(devices):
        ag.call("ag_set_device", d)
        st = C.c_void_p(); ag.call("ag_stream_create", C.byref(st))
        streams.append(st)
        bufs.append((Dev(shards_i64[k]), Dev(shards_f64[k])))
        outs.append((Dev(np.zeros(1, dtype=np.int64)), Dev(np.zeros(1))))
    res = []
    for rep in range(3):   # several epochs back to back: exercises the two-buffer mailbox
        for k, d in enumerate(devices):
            ag.call("ag_sum_i64_global_dev", comms[k], bufs[k][0].ptr, shards_i64[k].size, outs[k][0].ptr, streams[k])
        for k, d in enumerate(devices):
            ag.call("ag_sum_f64_global_dev", comms[k], bufs[k][1].ptr, shards_f64[k].size, outs[k][1].ptr, streams[k])
        for st in streams:
            ag.call("ag_stream_sync", st)
        res.append([(int(o[0].get()[0]), float(o[1].get()[0])) for o in outs])
    for k, d in enumerate(devices):
        ag.call("ag_set_device", d)
        ag.call("ag_stream_destroy", streams[k])
        ag.call("ag_comm_destroy", comms[k])
    ag.call("ag_set_device", devices[0])
    assert all(r == res[0] for r in res)
    return res[0]


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_global_sum_mailboxes_one_gpu(ag, cpu, world):
    """`world` ranks on device 0: wrapping int64 sum bit-exact, float64 fold in rank order — the value every rank gets
    equals the oracle's sum of the whole column (int) / math.fsum within 1 ULP (float), and all ranks agree."""
    import math
    ag.call("ag_init_all", None)
    rng = np.random.default_rng(world)
    n = 3_000_011
    xi = rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, n, dtype=np.int64, endpoint=True)
    xf = rng.standard_normal(n)
    cuts = [sharding.shard_range(n, r, world) for r in range(world)]
    got = _global_sums(ag, [0] * world, [xi[a:b] for a, b in cuts], [xf[a:b] for a, b in cuts])
    want_i = cpu.ref_sum_i64(xi.ctypes.data, n)
    exact = math.fsum(xf)
    for gi, gf in got:
        assert gi == want_i
        assert gf == got[0][1]
        assert abs(np.float64(gf).view(np.int64) - np.float64(exact).view(np.int64)) <= 1
    # empty shard on one rank still takes part
    got = _global_sums(ag, [0] * world, [xi[:0]] + [xi[a:b] for a, b in cuts[1:]], [xf[:0]] + [xf[a:b] for a, b in cuts[1:]])
    assert got[0][0] == cpu.ref_sum_i64(xi[cuts[0][1]:].ctypes.data, n - cuts[0][1])


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_single_process_all_devices(ag, cpu):
    """One process drives every GPU of the box (ag_init_all / ag_set_device): per-device shards of Add, Take and the
    global Sum over NVLink peer memory.  Needs >= 2 GPUs."""
    if _device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    nd = C.c_int()
    ag.call("ag_init_all", C.byref(nd))
    world = nd.value
    devices = list(range(world))
    rng = np.random.default_rng(5)
    n = 4_000_037
    xi = rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, n, dtype=np.int64, endpoint=True)
    xf = rng.integers(-(1 << 20), 1 << 20, n).astype(np.float64)
    cuts = [sharding.shard_range(n, r, world) for r in range(world)]
    got = _global_sums(ag, devices, [xi[a:b] for a, b in cuts], [xf[a:b] for a, b in cuts])
    for gi, gf in got:
        assert gi == cpu.ref_sum_i64(xi.ctypes.data, n) and gf == float(xf.sum())
    # per-shard Add and Take on each device's own stream; values replicated, indices sharded
    table = rng.integers(0, 1 << 62, 1 << 20, dtype=np.int64).view(np.uint64)
    idx = rng.integers(0, table.size, n).astype(np.int32)
    outs = []
    for k, d in enumerate(devices):
        ag.call("ag_set_device", d)
        a, b = cuts[k]
        dl, dr, do = Dev(xf[a:b]), Dev(xf[a:b][::-1].copy()), Dev(np.zeros(b - a))
        ag.call("ag_arith_binary_dev", N.FLOAT64, N.OP_ADD, N.SHAPE_AA, dl.ptr, dr.ptr, do.ptr, b - a, None)
        dt, di, dto = Dev(table), Dev(idx[a:b]), Dev(np.zeros(b - a, dtype=np.uint64))
        bad = Dev(np.zeros(1, dtype=np.int64))
        ag.call("ag_error_word_reset_dev", bad.ptr, None)
        ag.call("ag_take_primitive_dev", 64, dt.ptr, None, 0, table.size, 32, 1, di.ptr, None, 0, b - a, 1, dto.ptr, None, bad.ptr, None)
        outs.append((do, dto, bad, a, b))
    for k, d in enumerate(devices):
        ag.call("ag_set_device", d)
        ag.call("ag_stream_sync", None)
        do, dto, bad, a, b = outs[k]
        assert np.array_equal(do.get(), xf[a:b] + xf[a:b][::-1])
        assert np.array_equal(dto.get(), table[idx[a:b]]) and bad.get()[0] == N.NO_ERROR_POS
    ag.call("ag_set_device", 0)


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_multi_process_ranks_on_real_kernels():
    """One process per GPU under torch.distributed.run (the launch the driver uses): sharding.py on real kernels — IPC
    mailboxes, the fused global Sum, the NCCL form, sharded filter offsets and a sharded take with a planted bad
    index.  Needs >= 2 GPUs."""
    if _device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(_device_count(), 8)
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29633", os.path.join(ROOT, "tests", "mp_ranks.py")], capture_output=True, text=True, env=env, timeout=800)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "ALL RANKS OK" in out.stdout
