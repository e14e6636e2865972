#!/usr/bin/env python
"""One windowed take (int64 values <- int32 indices) for ncu captures: take_once.py [table_rows] [rows] [mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_go_b200 import _native as N
from arrow_go_b200.device import DeviceBuffer

vlen = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N.call("ag_init", 0)
N.call("ag_take_set_policy", mode, 0, 0, 0)
v = DeviceBuffer(vlen * 8); idx = DeviceBuffer(n * 4); o = DeviceBuffer(n * 8); bad = DeviceBuffer(64)
N.call("ag_generate_dev", 0, 1, 0, 0, v.ptr, vlen, None)
N.call("ag_generate_dev", 2, 0x0FF1CE, 0, vlen - 1, idx.ptr, n, None)
N.call("ag_error_word_reset_dev", bad.ptr, None)
for _ in range(2):
    N.call("ag_take_primitive_dev", 64, v.ptr, None, 0, vlen, 32, 1, idx.ptr, None, 0, n, 1, o.ptr, None, bad.ptr, None)
N.call("ag_stream_sync", None)
