#!/usr/bin/env python
"""CUDA-event timings of the Parquet decode primitives at 100M values (device resident)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from arrow_go_b200 import _native as N  # noqa: E402
from arrow_go_b200.device import DeviceBuffer, Event  # noqa: E402

N.call("ag_init", 0)
n = 100_000_000
a, o = DeviceBuffer(n * 8), DeviceBuffer(n * 8)
N.call("ag_generate_dev", 2, 0xDEF, 0, 1, a.ptr, 2 * n, None)
cnt = DeviceBuffer(64)
bm = DeviceBuffer(n // 8 + 64)
unp = C.c_int64()


def timed(name, fn, nbytes):
    fn(); fn()
    best = 1e9
    for _ in range(3):
        e0, e1 = Event(), Event()
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); e1.sync()
        best = min(best, e0.elapsed_ms(e1) / 10)
    print(f"{name:36s} {best * 1e3:8.1f} us  frac {nbytes / (best * 1e-3) / 1e9 / 6586.4:.3f}", flush=True)


timed("unpack32 13-bit", lambda: N.call("ag_parquet_unpack32_dev", a.ptr, o.ptr, n, 13, C.byref(unp), None), n * 13 / 8 + n * 4)
timed("bytes_to_bools", lambda: N.call("ag_parquet_bytes_to_bools_dev", a.ptr, n // 8, o.ptr, n, None), n * 1.125)
timed("def_levels_to_bitmap (flat)", lambda: N.call("ag_parquet_def_levels_to_bitmap_dev", a.ptr, n, 1, -1, bm.ptr, 0, n, cnt.ptr, None), n * 2.125)
