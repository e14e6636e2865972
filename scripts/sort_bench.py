#!/usr/bin/env python
"""sort_indices / is_in / unique timings at 100M rows (CUDA events, device resident)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_go_b200 import _native as N
from arrow_go_b200.device import DeviceBuffer, Event

N.call("ag_init", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
v, out = DeviceBuffer(n * 8), DeviceBuffer(n * 8)
nn, na = C.c_int64(), C.c_int64()


def timed(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = Event(), Event()
        e0.record(); fn(); e1.record(); e1.sync()
        best = min(best, e0.elapsed_ms(e1))
    return best


for name, lo, hi in (("int64 uniform in [-2^31, 2^31) (4 digits vary)", -(1 << 31), (1 << 31) - 1), ("int64 in [0, 100) (1 digit)", 0, 99), ("int64 in [0, 65535] (2 digits)", 0, 65535)):
    N.call("ag_generate_dev", 1, 0x5027, lo, hi, v.ptr, n, None)
    ms = timed(lambda: N.call("ag_sort_indices_dev", N.INT64, v.ptr, None, 0, n, 0, 0, out.ptr, C.byref(nn), C.byref(na), None))
    print(f"sort_indices {name:48s} {ms:8.3f} ms  {n / ms / 1e6:7.2f} G rows/s", flush=True)
N.call("ag_generate_dev", 0, 0x5027, 0, 0, v.ptr, n, None)
ms = timed(lambda: N.call("ag_sort_indices_dev", N.UINT64, v.ptr, None, 0, n, 0, 0, out.ptr, C.byref(nn), C.byref(na), None))
print(f"sort_indices {'uint64 full range (8 digits)':48s} {ms:8.3f} ms  {n / ms / 1e6:7.2f} G rows/s", flush=True)
N.call("ag_generate_dev", 3, 0x5027, -(1 << 20), 1 << 20, v.ptr, n, None)
ms = timed(lambda: N.call("ag_sort_indices_dev", N.FLOAT64, v.ptr, None, 0, n, 1, 1, out.ptr, C.byref(nn), C.byref(na), None))
print(f"sort_indices {'float64 integers in +-2^20, descending':48s} {ms:8.3f} ms  {n / ms / 1e6:7.2f} G rows/s", flush=True)
