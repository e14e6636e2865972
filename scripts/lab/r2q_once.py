#!/usr/bin/env python
"""One sort_indices call (for ncu)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from arrow_go_b200 import _native as N  # noqa: E402
from arrow_go_b200.device import DeviceBuffer  # noqa: E402
N.call("ag_init", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
v, out = DeviceBuffer(n * 8), DeviceBuffer(n * 8)
nn, na = C.c_int64(), C.c_int64()
N.call("ag_generate_dev", 1, 0x5027, -(1 << 31), (1 << 31) - 1, v.ptr, n, None)
N.call("ag_sort_indices_dev", N.INT64, v.ptr, None, 0, n, 0, 0, out.ptr, C.byref(nn), C.byref(na), None)
N.call("ag_stream_sync", None)
