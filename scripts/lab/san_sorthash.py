#!/usr/bin/env python
"""Small sort_indices / is_in / unique calls that reach every kernel variant of sort.cu and hash.cu (for compute-sanitizer:
memcheck and racecheck).  Results are compared with numpy so a sanitizer-clean run is also a correct one."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from arrow_go_b200 import _native as N  # noqa: E402

N.call("ag_init", 0)
rng = np.random.default_rng(11)


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bits(valid):
    return np.packbits(valid, bitorder="little")


n = int(sys.argv[1]) if len(sys.argv) > 1 else 40_003
# ---- sort: source pass only / pair passes / class pass / every equal / floats with NaN and nulls
for t, dt, lo, hi in ((N.INT64, np.int64, -(1 << 40), 1 << 40), (N.INT64, np.int64, -50, 50), (N.UINT8, np.uint8, 0, 255), (N.INT32, np.int32, 7, 7)):
    v = rng.integers(lo, hi, n, endpoint=True).astype(dt)
    for valid in (None, rng.random(n) > 0.2):
        for order in (0, 1):
            out = np.zeros(n, dtype=np.uint64); nn, na = C.c_int64(), C.c_int64()
            N.call("ag_sort_indices", t, ptr(v), ptr(bits(valid)) if valid is not None else None, 0, n, order, order, ptr(out), C.byref(nn), C.byref(na))
            keep = np.ones(n, bool) if valid is None else valid
            fin = out[nn.value:] if order else out[:n - nn.value]
            k = v[fin.astype(np.int64)].astype(np.int64)
            assert keep[fin.astype(np.int64)].all() and (np.all(k[1:] <= k[:-1]) if order else np.all(k[1:] >= k[:-1]))
f = rng.integers(-300, 300, n).astype(np.float64) * 0.25
f[rng.integers(0, n, 20)] = np.nan
out = np.zeros(n, dtype=np.uint64); nn, na = C.c_int64(), C.c_int64()
N.call("ag_sort_indices", N.FLOAT64, ptr(f), None, 0, n, 0, 0, ptr(out), C.byref(nn), C.byref(na))
assert np.array_equal(out[:n - na.value], np.argsort(f, kind="stable")[:n - na.value].astype(np.uint64))
# ---- is_in: bitmap / shared-memory table (two load factors) / HBM table
for bw, dt in ((64, np.int64), (32, np.uint32), (16, np.uint16), (8, np.int8)):
    info = np.iinfo(dt)
    v = rng.integers(max(info.min, -30_000), min(info.max, 30_000), n, endpoint=True).astype(dt)
    for sn in (200, 3000, 20_000):
        s = rng.integers(max(info.min, -30_000), min(info.max, 30_000), sn, endpoint=True).astype(dt)
        for valid in (None, rng.random(n) > 0.1):
            d = np.zeros(n // 8 + 8, np.uint8); ov = np.zeros(n // 8 + 8, np.uint8); cnt = C.c_int64()
            N.call("ag_is_in", bw, ptr(v), ptr(bits(valid)) if valid is not None else None, 0, n, ptr(s), None, 0, sn, 1, ptr(d), ptr(ov), C.byref(cnt))
            want = np.isin(v, s) & (valid if valid is not None else True)
            assert np.array_equal(np.unpackbits(d, bitorder="little")[:n].astype(bool), want)
# ---- unique: one table / small table only / small table overflowing into the full-size one
for policy in ((0, 0), (1000, 1024)):
    N.call("ag_unique_set_policy", *policy)
    for card in (7, 400, 5000):
        v = rng.integers(0, card, n).astype(np.int64)
        for valid in (None, rng.random(n) > 0.3):
            o = np.zeros(n, np.int64); ov = np.zeros(n // 8 + 8, np.uint8) if valid is not None else None
            ln, nn = C.c_int64(), C.c_int64()
            N.call("ag_unique", 64, ptr(v), ptr(bits(valid)) if valid is not None else None, 0, n, ptr(o), ptr(ov), C.byref(ln), C.byref(nn))
            vv = v if valid is None else v[valid]
            _, first = np.unique(vv, return_index=True)
            got = o[:ln.value] if valid is None else o[:ln.value][np.unpackbits(ov, bitorder="little")[:ln.value].astype(bool)]
            assert np.array_equal(got, vv[np.sort(first)]), (policy, card)
N.call("ag_unique_set_policy", 0, 0)
print("sanitizer driver: all checks passed", flush=True)
