#!/usr/bin/env python
"""Take: direct vs windowed path over (table rows, gathered rows) — where the automatic policy should switch.
Device-resident int64 values, int32 indices, CUDA-event timing, L2 flushed between iterations."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_go_b200 import _native as N
from arrow_go_b200.device import DeviceBuffer, Event

N.call("ag_init", 0)


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps + 1):
        N.call("ag_flush_l2", None)
        e0, e1 = Event(), Event()
        e0.record(); fn(); e1.record(); e1.sync()
        best = min(best, e0.elapsed_ms(e1))
    return best


def main():
    out = []
    grid = [(100_000_000, 100_000_000), (100_000_000, 50_000_000), (100_000_000, 25_000_000), (100_000_000, 12_500_000),
            (25_000_000, 100_000_000), (50_000_000, 50_000_000), (400_000_000, 100_000_000), (1_000_000_000, 125_000_000),
            (1_000_000_000, 62_500_000), (30_000_000, 30_000_000)]
    if len(sys.argv) > 1:
        grid = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
    for vlen, n in grid:
        v = DeviceBuffer(vlen * 8); idx = DeviceBuffer(n * 4); o = DeviceBuffer(n * 8); bad = DeviceBuffer(64)
        N.call("ag_generate_dev", 0, 1, 0, 0, v.ptr, vlen, None)
        N.call("ag_generate_dev", 2, 0x0FF1CE, 0, vlen - 1, idx.ptr, n, None)
        N.call("ag_error_word_reset_dev", bad.ptr, None)
        take = lambda: N.call("ag_take_primitive_dev", 64, v.ptr, None, 0, vlen, 32, 1, idx.ptr, None, 0, n, 1, o.ptr, None, bad.ptr, None)
        row = {"table_rows": vlen, "rows": n}
        for name, mode in (("direct", 1), ("windowed", 2), ("auto", 0)):
            N.call("ag_take_set_policy", mode, 0, 0, 0)
            ms = timed(take)
            row[name + "_ms"] = round(ms, 4)
            row[name + "_frac"] = round(20.0 * n / ms / 1e6 / 6586.4, 4)
        if os.environ.get("WINDOWS"):
            for wmb in (4, 8, 16, 32):
                N.call("ag_take_set_policy", 2, 0, 0, wmb << 20)
                row[f"windowed_{wmb}mb_ms"] = round(timed(take), 4)
            N.call("ag_take_set_policy", 0, 0, 0, 16 << 20)
        print(json.dumps(row), flush=True)
        out.append(row)
        for b in (v, idx, o, bad):
            b.free()
    N.call("ag_take_set_policy", 0, 0, 0, 0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/take_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
