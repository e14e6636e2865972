#!/usr/bin/env python
"""Where does the host-pointer Add lose time against the link?  Compares ag_arith_binary (contiguous), the span form,
and hand-driven upload/kernel/download pipelines with different chunk sizes / stream counts on the same pinned buffers."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_go_b200 import _native as N
from arrow_go_b200.device import DeviceBuffer, PinnedArray, Stream

N.call("ag_init", 0)
n = 100_000_000
ha, hb, ho = PinnedArray(n, np.float64), PinnedArray(n, np.float64), PinnedArray(n, np.float64)
ha.array[:] = 1.0; hb.array[:] = 2.0


def wall(fn, reps=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


ms = wall(lambda: N.call("ag_arith_binary", N.FLOAT64, N.OP_ADD, ha.ptr, hb.ptr, ho.ptr, n))
print(f"ag_arith_binary contiguous      {ms:7.2f} ms  {2.4 * n / 1e8 / ms * 1e2:6.1f} GB/s link")
spans, pos = [], 0
while pos < n:
    ln = min(1_000_000 - pos % 1_000_000, 999_983 - pos % 999_983, n - pos)
    spans.append((ha.ptr + 8 * pos, hb.ptr + 8 * pos, ho.ptr + 8 * pos, ln)); pos += ln
tab = N.span_table(spans)
ms = wall(lambda: N.call("ag_arith_binary_spans", N.FLOAT64, N.OP_ADD, N.SHAPE_AA, tab, len(spans)))
print(f"ag_arith_binary_spans (200)     {ms:7.2f} ms  {2.4 * n / 1e8 / ms * 1e2:6.1f} GB/s link")
for chunk_mb, nstreams in ((8, 3), (32, 3), (32, 4), (64, 3), (16, 6), (128, 3)):
    rows = chunk_mb * (1 << 20) // 8
    streams = [Stream() for _ in range(nstreams)]
    bufs = [(DeviceBuffer(rows * 8), DeviceBuffer(rows * 8), DeviceBuffer(rows * 8)) for _ in range(nstreams)]

    def run():
        k = 0
        for r0 in range(0, n, rows):
            ln = min(rows, n - r0)
            st = streams[k % nstreams]; dl, dr, do = bufs[k % nstreams]
            N.call("ag_upload", dl.ptr, ha.ptr + 8 * r0, ln * 8, st.handle)
            N.call("ag_upload", dr.ptr, hb.ptr + 8 * r0, ln * 8, st.handle)
            N.call("ag_arith_binary_dev", N.FLOAT64, N.OP_ADD, N.SHAPE_AA, dl.ptr, dr.ptr, do.ptr, ln, st.handle)
            N.call("ag_download", ho.ptr + 8 * r0, do.ptr, ln * 8, st.handle)
            k += 1
        for st in streams:
            st.sync()
    ms = wall(run)
    print(f"manual pipeline {chunk_mb:3d} MB x {nstreams} streams {ms:7.2f} ms  {2.4 * n / 1e8 / ms * 1e2:6.1f} GB/s link")
    for st in streams:
        st.close()
    for t in bufs:
        for b in t:
            b.free()
# separate engines: all H2D on one stream, all D2H on another, events between
assert np.all(ho.array[:1000] == 3.0)
