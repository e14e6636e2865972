#!/usr/bin/env python
"""One fused Greater+Filter at 100M int64 rows, 10 % selected (for ncu captures)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_go_b200 import _native as N
from arrow_go_b200.device import DeviceBuffer

rows = 100_000_000
N.call("ag_init", 0)
v = DeviceBuffer(rows * 8); o = DeviceBuffer(rows * 8 // 4); scal = DeviceBuffer(64)
N.call("ag_generate_dev", 1, 0x0FF1CE, 0, 99, v.ptr, rows, None)
sc = np.array([89], dtype=np.int64)
for _ in range(3):
    N.call("ag_filter_compare_scalar_dev", N.INT64, N.CMP_GT, v.ptr, sc.ctypes.data, rows, o.ptr, rows // 4, scal.ptr, None)
N.call("ag_stream_sync", None)
