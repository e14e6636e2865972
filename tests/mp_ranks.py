"""Worker of tests/test_multi_gpu.py::test_multi_process_ranks_on_real_kernels — one process per GPU (torchrun).
Every rank owns a row range of the same seeded columns; results are checked against the oracle on rank-local data and
through values every rank can compute (the generator is counter based)."""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from arrow_go_b200 import _native as N  # noqa: E402
from arrow_go_b200 import sharding  # noqa: E402
from helpers import Dev  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N.call("ag_init", local)
    cpu = oracle.cpu()
    comm = sharding.create_comm(dist, device="cuda")
    sharding.attach_nccl(comm, dist, device="cuda")
    n = 20_000_003
    rng = np.random.default_rng(0x94378165)     # same seed on every rank -> same global column
    xi_all = rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, n, dtype=np.int64, endpoint=True)
    xf_all = rng.standard_normal(n)
    a, b = sharding.shard_range(n, rank, world)
    di, df = Dev(xi_all[a:b]), Dev(xf_all[a:b])
    oi, of, on = Dev(np.zeros(1, dtype=np.int64)), Dev(np.zeros(1)), Dev(np.zeros(1, dtype=np.int64))
    for _ in range(4):
        N.call("ag_sum_i64_global_dev", comm, di.ptr, b - a, oi.ptr, None)
        N.call("ag_sum_f64_global_dev", comm, df.ptr, b - a, of.ptr, None)
        N.call("ag_sum_i64_global_nccl_dev", comm, di.ptr, b - a, on.ptr, None)
    N.call("ag_stream_sync", None)
    want_i = cpu.ref_sum_i64(xi_all.ctypes.data, n)
    assert int(oi.get()[0]) == want_i, (rank, int(oi.get()[0]), want_i)
    assert int(on.get()[0]) == want_i
    gf = float(of.get()[0])
    exact = math.fsum(xf_all)
    assert abs(int(np.float64(gf).view(np.int64)) - int(np.float64(exact).view(np.int64))) <= 1, (gf, exact)
    allf = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(allf, torch.tensor([gf], dtype=torch.float64, device="cuda"))
    assert all(float(t.item()) == gf for t in allf), "ranks disagree on the float64 global sum"

    # sharded filter: local compaction + exclusive scan of the counts (sharding.filter_output_offsets)
    vals = rng.integers(0, 100, n).astype(np.int64)
    mask_bits = np.packbits(vals[a:b] > 89, bitorder="little")
    dv, dm = Dev(vals[a:b]), Dev(np.concatenate([mask_bits, np.zeros(8, dtype=np.uint8)]))
    cnt = int((vals[a:b] > 89).sum())
    dout, dlen = Dev(np.zeros(max(cnt, 1), dtype=np.int64)), Dev(np.zeros(2, dtype=np.int64))
    N.call("ag_filter_primitive_dev", 64, dv.ptr, None, 0, dm.ptr, None, 0, b - a, 0, dout.ptr, None, cnt, dlen.ptr, None)
    N.call("ag_stream_sync", None)
    assert int(dlen.get()[0]) == cnt
    off, total = sharding.filter_output_offsets(cnt, dist, device="cuda")
    glob = vals[vals > 89]
    assert total == glob.size and np.array_equal(dout.get()[:cnt], glob[off:off + cnt])

    # sharded take: table replicated, indices by row range; a planted bad index is found at its GLOBAL row
    table = rng.integers(0, 1 << 62, 1 << 22, dtype=np.int64)
    idx = rng.integers(0, table.size, n).astype(np.int32)
    planted = n // 2 + 12345
    idx[planted] = -7
    dt, dix, dto, bad = Dev(table), Dev(idx[a:b]), Dev(np.zeros(b - a, dtype=np.int64)), Dev(np.zeros(1, dtype=np.int64))
    N.call("ag_error_word_reset_dev", bad.ptr, None)
    N.call("ag_take_primitive_dev", 64, dt.ptr, None, 0, table.size, 32, 1, dix.ptr, None, 0, b - a, 1, dto.ptr, None, bad.ptr, None)
    N.call("ag_stream_sync", None)
    local_bad = int(bad.get()[0])
    g = torch.tensor([local_bad + a if local_bad != N.NO_ERROR_POS else N.NO_ERROR_POS], dtype=torch.int64, device="cuda")
    dist.all_reduce(g, op=dist.ReduceOp.MIN)
    assert int(g.item()) == planted
    ok = np.ones(b - a, dtype=bool)
    if a <= planted < b:
        ok[planted - a] = False
    assert np.array_equal(dto.get()[ok], table[idx[a:b][ok]])

    N.call("ag_comm_destroy", comm)
    dist.barrier()
    if rank == 0:
        print("ALL RANKS OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
