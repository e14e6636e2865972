#!/usr/bin/env python
"""Only the cumulative-sum kernel on a 100M-row int64 column (for an ncu capture)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_go_b200 import _native as N  # noqa: E402
from arrow_go_b200.device import DeviceBuffer  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
t = {"i64": N.INT64, "f64": N.FLOAT64, "i32": N.INT32}[sys.argv[2] if len(sys.argv) > 2 else "i64"]
N.call("ag_init", 0)
a, o, state, bad = DeviceBuffer(rows * 8), DeviceBuffer(rows * 8), DeviceBuffer(64), DeviceBuffer(64)
N.call("ag_generate_dev", 1, 0x94378165, -1000, 1000, a.ptr, rows, None)
N.call("ag_error_word_reset_dev", bad.ptr, None)
for _ in range(3):
    N.call("ag_cumulative_sum_state_init_dev", state.ptr, t, None, None)
    N.call("ag_cumulative_sum_dev", t, a.ptr, None, 0, rows, 1, 0, o.ptr, None, 0, state.ptr, bad.ptr, None)
N.call("ag_stream_sync", None)
