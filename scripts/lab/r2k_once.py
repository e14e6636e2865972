#!/usr/bin/env python
"""One call each of sort_indices / unique / is_in at 100M rows (for an ncu launch list: per-kernel durations)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from arrow_go_b200 import _native as N  # noqa: E402
from arrow_go_b200.device import DeviceBuffer  # noqa: E402

N.call("ag_init", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
v, out = DeviceBuffer(n * 8), DeviceBuffer(n * 8)
scal = DeviceBuffer(64)
nn, na = C.c_int64(), C.c_int64()
N.call("ag_generate_dev", 1, 0x5027, -(1 << 31), (1 << 31) - 1, v.ptr, n, None)
N.call("ag_sort_indices_dev", N.INT64, v.ptr, None, 0, n, 0, 0, out.ptr, C.byref(nn), C.byref(na), None)
N.call("ag_stream_sync", None)
for distinct in (100, 1_000_000, (1 << 32) - 1):
    N.call("ag_generate_dev", 1, 0x16, 0, distinct - 1, v.ptr, n, None)
    N.call("ag_unique_dev", 64, v.ptr, None, 0, n, out.ptr, None, n, scal.ptr, None)
    N.call("ag_stream_sync", None)
    print("unique", distinct, "->", int(scal.to_numpy(np.int64, 1)[0]), flush=True)
sset = DeviceBuffer(8000)
hs = np.arange(0, 100_000, 100, dtype=np.int64)
N.call("ag_upload", sset.ptr, hs.ctypes.data, 8000, None)
N.call("ag_generate_dev", 1, 0x15, 0, 99_999, v.ptr, n, None)
bm1, bm2 = DeviceBuffer(n // 8 + 64), DeviceBuffer(n // 8 + 64)
N.call("ag_is_in_dev", 64, v.ptr, None, 0, n, sset.ptr, None, 0, 1000, 0, bm1.ptr, bm2.ptr, scal.ptr, None)
N.call("ag_stream_sync", None)
