#!/usr/bin/env python
"""SURVEY §8(d) row counts: 8192 (BASELINE config 1, the reference's own README case), 1M and 100M rows on one
GPU — device-resident kernel time per call (CUDA events, mean of 200 / 50 / 10 calls) for Sum / Add / Greater /
Filter / Take, the host-pointer call latency at 8192 rows (wall clock: what a per-span cgo binding pays), and the
reference's own loops on this host for the same sizes (hot cache at 8192, like `go test -bench`)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_go_b200 import _native as N  # noqa: E402
from arrow_go_b200.device import DeviceBuffer, Event  # noqa: E402
from oracle import oracle  # noqa: E402

N.call("ag_init", 0)
ref = oracle.ref()
out = {}
for rows, reps in ((8192, 200), (1_000_000, 50), (100_000_000, 10)):
    a, b, o = DeviceBuffer(rows * 8 + 64), DeviceBuffer(rows * 8 + 64), DeviceBuffer(rows * 8 + 64)
    idx, mask, scal, bad = DeviceBuffer(rows * 4 + 64), DeviceBuffer(rows // 8 + 64), DeviceBuffer(64), DeviceBuffer(64)
    N.call("ag_generate_dev", 3, 1, -(1 << 20), 1 << 20, a.ptr, rows, None)
    N.call("ag_generate_dev", 1, 2, 0, 99, b.ptr, rows, None)
    N.call("ag_generate_dev", 2, 3, 0, rows - 1, idx.ptr, rows, None)
    N.call("ag_error_word_reset_dev", bad.ptr, None)
    sc = np.array([89], dtype=np.int64)
    N.call("ag_compare_dev", N.INT64, N.CMP_GT, N.SHAPE_AS, b.ptr, sc.ctypes.data, mask.ptr, rows, 0, None)
    N.call("ag_filter_output_size_dev", mask.ptr, None, 0, rows, 0, scal.ptr, None)
    N.call("ag_stream_sync", None)
    cnt = int(scal.to_numpy(np.int64, 1)[0])

    def timed(fn):
        for _ in range(3):
            fn()
        N.call("ag_stream_sync", None)
        e0, e1 = Event(), Event()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.sync()
        return e0.elapsed_ms(e1) / reps * 1e3

    r = {
        "sum_f64_us": timed(lambda: N.call("ag_sum_f64_dev", a.ptr, rows, scal.ptr, None)),
        "add_f64_us": timed(lambda: N.call("ag_arith_binary_dev", N.FLOAT64, N.OP_ADD, N.SHAPE_AA, a.ptr, a.ptr, o.ptr, rows, None)),
        "greater_i64_us": timed(lambda: N.call("ag_compare_dev", N.INT64, N.CMP_GT, N.SHAPE_AS, b.ptr, sc.ctypes.data, mask.ptr, rows, 0, None)),
        "filter_i64_us": timed(lambda: N.call("ag_filter_primitive_dev", 64, b.ptr, None, 0, mask.ptr, None, 0, rows, 0, o.ptr, None, max(cnt, 1), scal.ptr + 8, None)),
        "take_i64_us": timed(lambda: N.call("ag_take_primitive_dev", 64, b.ptr, None, 0, rows, 32, 1, idx.ptr, None, 0, rows, 1, o.ptr, None, bad.ptr, None)),
    }
    if rows <= 1_000_000:  # host-pointer flavour: one synchronous call, copies inside
        x = np.arange(rows, dtype=np.float64)
        res = C.c_double()
        for _ in range(5):
            N.call("ag_sum_f64", x.ctypes.data, rows, C.byref(res))
        t0 = time.perf_counter()
        for _ in range(50):
            N.call("ag_sum_f64", x.ctypes.data, rows, C.byref(res))
        r["sum_f64_host_call_us"] = (time.perf_counter() - t0) / 50 * 1e6
        assert res.value == float(x.sum())
    if ref is not None:  # the reference's own loops here (1 thread)
        x = np.arange(rows, dtype=np.float64)
        y = np.empty(rows)
        res = C.c_double()
        k = max(3, min(2000, 200_000_000 // rows))
        ref.sum_float64_avx2(x.ctypes.data, rows, C.addressof(res))
        t0 = time.perf_counter()
        for _ in range(k):
            ref.sum_float64_avx2(x.ctypes.data, rows, C.addressof(res))
        r["reference_sum_f64_avx2_us"] = (time.perf_counter() - t0) / k * 1e6
        ref.arithmetic_binary_avx2(12, 0, x.ctypes.data, x.ctypes.data, y.ctypes.data, rows)
        t0 = time.perf_counter()
        for _ in range(k):
            ref.arithmetic_binary_avx2(12, 0, x.ctypes.data, x.ctypes.data, y.ctypes.data, rows)
        r["reference_add_f64_avx2_us"] = (time.perf_counter() - t0) / k * 1e6
    out[str(rows)] = {k2: round(v, 2) for k2, v in r.items()}
    for d in (a, b, o, idx, mask, scal, bad):
        d.free()
print(json.dumps(out, indent=1))
