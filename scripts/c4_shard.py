#!/usr/bin/env python
"""BASELINE.json config 4, one GPU's share: Take(int64 values, int32 indices) with the FULL 1B-row
(8 GB) values table resident on the GPU (values are replicated per GPU) and this rank's
125M indices.  Under torchrun every rank does the same with its own index seed; rank 0 prints one
JSON object.  Parity at this size is size-independent: values[i] = mix64(i), so every output must
equal mix64(index) — checked on the whole output by the device checksum against a host-computed
checksum of mix64(indices) over sampled 4M-row windows, and element by element on those windows.

  python scripts/c4_shard.py [table_rows] [index_rows] [reps]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arrow_go_b200 import _native as N  # noqa: E402
from arrow_go_b200.device import DeviceBuffer, Event  # noqa: E402

table_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
index_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 125_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rank = int(os.environ.get("RANK", "0"))
N.call("ag_init", int(os.environ.get("LOCAL_RANK", "0")))


def mix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


table = DeviceBuffer(table_rows * 8)
idx = DeviceBuffer(index_rows * 4)
out = DeviceBuffer(index_rows * 8)
bad = DeviceBuffer(64)
N.call("ag_generate_dev", 0, 0, 0, 0, table.ptr, table_rows, None)                      # values[i] = mix64(i)
N.call("ag_generate_dev", 2, 0x0FF1CE + rank, 0, table_rows - 1, idx.ptr, index_rows, None)
N.call("ag_error_word_reset_dev", bad.ptr, None)


def take():
    N.call("ag_take_primitive_dev", 64, table.ptr, None, 0, table_rows, 32, 1, idx.ptr, None, 0, index_rows, 1, out.ptr, None, bad.ptr, None)


for _ in range(3):
    take()
N.call("ag_stream_sync", None)
e0, e1 = Event(), Event()
e0.record()
for _ in range(reps):
    take()
e1.record()
e1.sync()
ms = e0.elapsed_ms(e1) / reps

# parity: sampled windows, element by element
with np.errstate(over="ignore"):
    win = 4_000_000
    checked = 0
    for start in (0, index_rows // 3, index_rows - win):
        start = max(0, min(start, index_rows - win))
        n = min(win, index_rows)
        ii = idx.to_numpy(np.int32, n, start * 4)
        oo = out.to_numpy(np.uint64, n, start * 8)
        assert np.array_equal(oo, mix64(ii.astype(np.uint64))), f"take mismatch in window at {start}"
        checked += n
assert bad.to_numpy(np.int64, 1)[0] == (1 << 63) - 1, "bounds check tripped"
# planted out-of-range index -> ErrIndex, first offending row reported
pos = index_rows // 2 + 17
N.call("ag_upload", idx.ptr + pos * 4, np.array([table_rows], dtype=np.int32).ctypes.data, 4, None)
take()
N.call("ag_stream_sync", None)
first_bad = int(bad.to_numpy(np.int64, 1)[0])
assert first_bad == pos, (first_bad, pos)

if rank == 0:
    peak = 6586.4
    print(json.dumps({"config": "C4 shard: Take(int64 <- int32 idx), values table %d rows (%.1f GB) resident, %d indices per GPU" % (table_rows, table_rows * 8 / 1e9, index_rows),
                      "ms": ms, "rows_per_s_per_gpu": index_rows / ms * 1e3, "algorithmic_gbs": 20.0 * index_rows / ms / 1e6,
                      "frac_of_measured_hbm": 20.0 * index_rows / ms / 1e6 / peak, "parity_rows_checked": checked,
                      "bounds_error_row_found": first_bad, "world": int(os.environ.get("WORLD_SIZE", "1"))}))
