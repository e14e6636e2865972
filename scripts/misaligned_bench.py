#!/usr/bin/env python
"""Add f64 on element-aligned (sliced) operands: l.slice(1, n) + r.slice(0, n) and friends, 100M rows, device resident."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_go_b200 import _native as N
from arrow_go_b200.device import DeviceBuffer, Event

N.call("ag_init", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
PEAK = 6586.4


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = Event(), Event()
        e0.record(); fn(); e1.record(); e1.sync()
        best = min(best, e0.elapsed_ms(e1))
    return best


res = {}
for tname, tid, w in (("f64", N.FLOAT64, 8), ("i32", N.INT32, 4), ("u8", N.UINT8, 1)):
    l, r, o = DeviceBuffer(n * w + 256), DeviceBuffer(n * w + 256), DeviceBuffer(n * w + 256)
    N.call("ag_dev_memset", l.ptr, 1, n * w + 256, None); N.call("ag_dev_memset", r.ptr, 2, n * w + 256, None)
    for name, (lo, ro, oo) in {"aligned": (0, 0, 0), "l+1": (1, 0, 0), "r+1": (0, 1, 0), "l+1,r+1": (1, 1, 0), "out+1": (0, 0, 1),
                               "all+1": (1, 1, 1), "l+1,r+3,out+2" if w < 8 else "l+1,out+1": (1, 3 if w < 8 else 0, 2 if w < 8 else 1)}.items():
        ms = timed(lambda: N.call("ag_arith_binary_dev", tid, N.OP_ADD, N.SHAPE_AA, l.ptr + lo * w, r.ptr + ro * w, o.ptr + oo * w, n, None))
        res[f"add_{tname}_{name}"] = {"ms": round(ms, 4), "frac": round(3.0 * w * n / ms / 1e6 / PEAK, 4)}
        print(f"add_{tname}_{name:16s} {ms:8.4f} ms  frac {3.0 * w * n / ms / 1e6 / PEAK:.3f}", flush=True)
    for b in (l, r, o):
        b.free()
# unary Abs on a sliced input
n8 = n
x, o = DeviceBuffer(n8 * 8 + 256), DeviceBuffer(n8 * 8 + 256)
N.call("ag_dev_memset", x.ptr, 1, n8 * 8 + 256, None)
for name, (io, oo) in {"aligned": (0, 0), "in+1": (1, 0), "out+1": (0, 1), "both+1": (1, 1)}.items():
    ms = timed(lambda: N.call("ag_arith_unary_same_dev", N.FLOAT64, N.OP_ABS, x.ptr + io * 8, o.ptr + oo * 8, n8, None))
    res[f"abs_f64_{name}"] = {"ms": round(ms, 4), "frac": round(16.0 * n8 / ms / 1e6 / PEAK, 4)}
    print(f"abs_f64_{name:20s} {ms:8.4f} ms  frac {16.0 * n8 / ms / 1e6 / PEAK:.3f}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/misaligned_bench.json", "w"), indent=1)
