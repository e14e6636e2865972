#!/usr/bin/env python
"""Runs each hot-path kernel a few times on device-resident 100M-row columns so that ONE
`ncu --set full -k regex:...` call can capture them (bench.py is the timing source; numbers
printed here under a profiler are not bench values).

  python scripts/prof_kernels.py [rows] [reps]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_go_b200 import _native as N  # noqa: E402
from arrow_go_b200.device import DeviceBuffer, Event  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N.call("ag_init", 0)
a, b, o = DeviceBuffer(rows * 8), DeviceBuffer(rows * 8), DeviceBuffer(rows * 8)
N.call("ag_generate_dev", 3, 0x94378165, -(1 << 20), 1 << 20, a.ptr, rows, None)
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


def timed(name, fn):
    fn()
    N.call("ag_stream_sync", None)
    e0, e1 = Event(), Event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.sync()
    print(f"{name:32s} {e0.elapsed_ms(e1) / reps * 1e3:10.1f} us")


timed("sum_f64", lambda: N.call("ag_sum_f64_dev", a.ptr, rows, scal.ptr, None))
timed("add_f64_contig", lambda: N.call("ag_arith_binary_dev", N.FLOAT64, N.OP_ADD, N.SHAPE_AA, a.ptr, a.ptr, o.ptr, rows, None))
timed("greater_i64_scalar", lambda: N.call("ag_compare_dev", N.INT64, N.CMP_GT, N.SHAPE_AS, b.ptr, sc.ctypes.data, mask.ptr, rows, 0, None))
timed("filter_i64", lambda: N.call("ag_filter_primitive_dev", 64, b.ptr, None, 0, mask.ptr, None, 0, rows, 0, o.ptr, None, cnt, scal.ptr + 8, None))
timed("fused_greater_filter_i64", lambda: N.call("ag_filter_compare_scalar_dev", N.INT64, N.CMP_GT, b.ptr, sc.ctypes.data, rows, o.ptr, cnt, scal.ptr + 8, None))
timed("take_i64_i32", lambda: N.call("ag_take_primitive_dev", 64, b.ptr, None, 0, rows, 32, 1, idx.ptr, None, 0, rows, 1, o.ptr, None, bad.ptr, None))
# the default compute.Add on integers is the CHECKED kernel (ScalarBinaryNotNull semantics)
N.call("ag_generate_dev", 1, 0x94378165, -1000, 1000, a.ptr, rows, None)
valid = DeviceBuffer(rows // 8 + 64)
N.call("ag_generate_dev", 4, 0x1234, 9, 10, valid.ptr, rows, None)  # 90 % valid
timed("add_checked_i64_nonull", lambda: N.call("ag_arith_checked_dev", N.INT64, N.OP_ADD_CHECKED, N.SHAPE_AA, a.ptr, None, 0, b.ptr, None, 0, o.ptr, rows, bad.ptr, None))
timed("add_checked_i64_nulls", lambda: N.call("ag_arith_checked_dev", N.INT64, N.OP_ADD_CHECKED, N.SHAPE_AA, a.ptr, valid.ptr, 0, b.ptr, valid.ptr, 3, o.ptr, rows, bad.ptr, None))
timed("add_unchecked_i64", lambda: N.call("ag_arith_binary_dev", N.INT64, N.OP_ADD, N.SHAPE_AA, a.ptr, b.ptr, o.ptr, rows, None))
timed("bitmap_and_100m_bits", lambda: N.call("ag_bitmap_op_dev", N.BITOP_AND, valid.ptr, 0, mask.ptr, 5, o.ptr, 3, rows, None))
timed("bitmap_popcount_100m_bits", lambda: N.call("ag_bitmap_popcount_dev", valid.ptr, 3, rows - 3, scal.ptr, None))
timed("abs_f64", lambda: N.call("ag_arith_unary_same_dev", N.FLOAT64, N.OP_ABS, a.ptr, o.ptr, rows, None))
# numeric casts (implicit promotion): a holds int64 in [-1000, 1000]; idx holds int32
timed("cast_i32_to_i64", lambda: N.call("ag_cast_numeric_dev", N.INT32, N.INT64, idx.ptr, o.ptr, rows, None))
timed("cast_i64_to_f64_checked", lambda: N.call("ag_cast_numeric_checked_dev", N.INT64, N.FLOAT64, a.ptr, None, 0, o.ptr, rows, 0, 0, bad.ptr, None))
timed("cast_i64_to_f64_checked_nulls", lambda: N.call("ag_cast_numeric_checked_dev", N.INT64, N.FLOAT64, a.ptr, valid.ptr, 3, o.ptr, rows, 0, 0, bad.ptr, None))
timed("cast_i64_to_i32_checked", lambda: N.call("ag_cast_numeric_checked_dev", N.INT64, N.INT32, a.ptr, None, 0, o.ptr, rows, 0, 0, bad.ptr, None))
timed("cast_i32_to_i8_unsafe", lambda: N.call("ag_cast_numeric_dev", N.INT32, N.INT8, idx.ptr, o.ptr, rows, None))
timed("cast_f64_to_f32", lambda: N.call("ag_cast_numeric_dev", N.FLOAT64, N.FLOAT32, b.ptr, o.ptr, rows, None))
timed("min_max_i64", lambda: N.call("ag_min_max_dev", N.INT64, b.ptr, rows, scal.ptr, None))
timed("min_max_i32", lambda: N.call("ag_min_max_dev", N.INT32, idx.ptr, rows, scal.ptr, None))
state = DeviceBuffer(64)
ovalid = DeviceBuffer(rows // 8 + 64)
def cumsum(t, src, v):
    N.call("ag_cumulative_sum_state_init_dev", state.ptr, t, None, None)
    N.call("ag_cumulative_sum_dev", t, src, v, 3, rows, 1, 0, o.ptr, ovalid.ptr if v else None, 0, state.ptr, bad.ptr, None)
timed("cumsum_i64", lambda: cumsum(N.INT64, a.ptr, None))
timed("cumsum_i64_nulls_skip", lambda: cumsum(N.INT64, a.ptr, valid.ptr))
timed("cumsum_f64", lambda: cumsum(N.FLOAT64, b.ptr, None))
timed("cumsum_i32", lambda: cumsum(N.INT32, idx.ptr, None))
print("selected rows:", cnt)
