#!/usr/bin/env python
"""Round-2 lab: CUDA-event timings of the scan / filter / fused filter / take kernels at 100M rows, each under the
experiment switches of ag_lab_set (0 = shipped).  Not a bench: numbers feed decisions, bench.py is the record."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from arrow_go_b200 import _native as N  # noqa: E402
from arrow_go_b200.device import DeviceBuffer, Event  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N.call("ag_init", 0)
lab = getattr(N.raw(), "ag_lab_set", lambda k, v: None)   # experiment switches existed only in lab builds
a, b, o = DeviceBuffer(rows * 8), DeviceBuffer(rows * 8), DeviceBuffer(rows * 8)
N.call("ag_generate_dev", 1, 0x94378165, -1000, 1000, a.ptr, rows, None)
N.call("ag_generate_dev", 1, 0x0FF1CE, 0, 99, b.ptr, rows, None)
scal = DeviceBuffer(64)
sc = np.array([89], dtype=np.int64)
mask = DeviceBuffer(rows // 8 + 64)
idx = DeviceBuffer(rows * 4)
N.call("ag_generate_dev", 2, 0x0FF1CE + 7, 0, rows - 1, idx.ptr, rows, None)
bad = DeviceBuffer(64)
N.call("ag_error_word_reset_dev", bad.ptr, None)
N.call("ag_compare_dev", N.INT64, N.CMP_GT, N.SHAPE_AS, b.ptr, sc.ctypes.data, mask.ptr, rows, 0, None)
N.call("ag_filter_output_size_dev", mask.ptr, None, 0, rows, 0, scal.ptr, None)
N.call("ag_stream_sync", None)
cnt = int(scal.to_numpy(np.int64, 1)[0])
PEAK = 6586.4


def timed(name, fn, nbytes):
    fn(); fn()
    N.call("ag_stream_sync", None)
    best = 1e9
    tot = 0.0
    for _ in range(3):
        e0, e1 = Event(), Event()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.sync()
        t = e0.elapsed_ms(e1) / reps
        best = min(best, t)
        tot += t
    print(f"{name:44s} best {best * 1e3:9.1f} us  mean {tot / 3 * 1e3:9.1f} us  frac {nbytes / (best * 1e-3) / 1e9 / PEAK:6.3f}", flush=True)


state = DeviceBuffer(64)
ovalid = DeviceBuffer(rows // 8 + 64)
valid = DeviceBuffer(rows // 8 + 64)
N.call("ag_generate_dev", 4, 0x1234, 9, 10, valid.ptr, rows, None)


def cumsum(t, src, v):
    N.call("ag_cumulative_sum_state_init_dev", state.ptr, t, None, None)
    N.call("ag_cumulative_sum_dev", t, src, v, 3, rows, 1, 0, o.ptr, ovalid.ptr if v else None, 0, state.ptr, bad.ptr, None)


timed("cumsum_i64 (stream kernel)", lambda: cumsum(N.INT64, a.ptr, None), rows * 16)
timed("cumsum_f64 (stream kernel)", lambda: cumsum(N.FLOAT64, b.ptr, None), rows * 16)
timed("cumsum_i32 (stream kernel)", lambda: cumsum(N.INT32, idx.ptr, None), rows * 8)
timed("cumsum_i64_nulls_skip (stream kernel + validity)", lambda: cumsum(N.INT64, a.ptr, valid.ptr), rows * 16)
timed("cumsum_i64 input 8 B off 16 (stream kernel)", lambda: (N.call("ag_cumulative_sum_state_init_dev", state.ptr, N.INT64, None, None),
      N.call("ag_cumulative_sum_dev", N.INT64, a.ptr + 8, None, 0, rows - 1, 1, 0, o.ptr, None, 0, state.ptr, bad.ptr, None)), rows * 16)
fb = rows * 8 + rows // 8 + cnt * 8
timed("filter_i64 (2-level look-back)", lambda: N.call("ag_filter_primitive_dev", 64, b.ptr, None, 0, mask.ptr, None, 0, rows, 0, o.ptr, None, cnt, scal.ptr + 8, None), fb)
for k in (0, 1):
    lab(1, k)
    timed(f"fused_greater_filter_i64 knob1={k}", lambda: N.call("ag_filter_compare_scalar_dev", N.INT64, N.CMP_GT, b.ptr, sc.ctypes.data, rows, o.ptr, cnt, scal.ptr + 8, None), rows * 8 + cnt * 8)
lab(1, 0)
for k in (0, 1):
    lab(0, k)
    timed(f"take_i64_i32 random knob0={k}", lambda: N.call("ag_take_primitive_dev", 64, b.ptr, None, 0, rows, 32, 1, idx.ptr, None, 0, rows, 1, o.ptr, None, bad.ptr, None), rows * 20)
lab(0, 0)
timed("greater_i64_scalar", lambda: N.call("ag_compare_dev", N.INT64, N.CMP_GT, N.SHAPE_AS, b.ptr, sc.ctypes.data, mask.ptr, rows, 0, None), rows * 8 + rows // 8)
# set lookup / unique
N.call("ag_generate_dev", 1, 0x15, 0, 99_999, a.ptr, rows, None)
sset = DeviceBuffer(8000)
hs = np.arange(0, 100_000, 100, dtype=np.int64)
N.call("ag_upload", sset.ptr, hs.ctypes.data, 8000, None)
bm1, bm2 = DeviceBuffer(rows // 8 + 64), DeviceBuffer(rows // 8 + 64)
timed("is_in_i64 (1000-value set)", lambda: N.call("ag_is_in_dev", 64, a.ptr, None, 0, rows, sset.ptr, None, 0, 1000, 0, bm1.ptr, bm2.ptr, scal.ptr, None), rows * 8.25)
N.call("ag_generate_dev", 1, 0x16, 0, 99, a.ptr, rows, None)
timed("unique_i64 (100 distinct)", lambda: N.call("ag_unique_dev", 64, a.ptr, None, 0, rows, o.ptr, None, rows, scal.ptr, None), rows * 8.0)
print("selected rows:", cnt)
