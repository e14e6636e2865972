"""Thread-safety of the C ABI: cgo calls arrive on arbitrary OS threads and several CallFunctions may
run concurrently (arrow/compute/exec.go:164-170, selection.go:127-150), so the host-pointer entry
points must be re-entrant.  ctypes releases the GIL during the calls, so these really overlap."""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import pack_bits, ptr

pytestmark = pytest.mark.gpu


def test_concurrent_host_calls(ag, cpu):
    rng = np.random.default_rng(77)
    n = 300_001
    a = rng.integers(-(1 << 20), 1 << 20, n).astype(np.float64)
    b = rng.integers(-(1 << 20), 1 << 20, n).astype(np.float64)
    vi = rng.integers(0, 100, n).astype(np.int64)
    mask = pack_bits(rng.random(n) < 0.3, 3)
    idx = rng.integers(0, n, n).astype(np.int32)
    want_add = a + b
    want_sum = float(a.sum())
    want_f = np.zeros(n, dtype=np.int64); wl = C.c_int64()
    assert cpu.ref_filter_primitive(64, ptr(vi), None, 0, ptr(mask), None, 3, n, 0, ptr(want_f), None, C.byref(wl), None) == 0
    want_t = vi[idx]

    def job(k):
        for _ in range(5):
            kind = k % 4
            if kind == 0:
                out = np.empty(n)
                ag.call("ag_arith_binary", N.FLOAT64, N.OP_ADD, ptr(a), ptr(b), ptr(out), n)
                assert np.array_equal(out, want_add)
            elif kind == 1:
                r = C.c_double()
                ag.call("ag_sum_f64", ptr(a), n, C.byref(r))
                assert r.value == want_sum
            elif kind == 2:
                out = np.zeros(n, dtype=np.int64); gl = C.c_int64()
                ag.call("ag_filter_primitive", 64, ptr(vi), None, 0, ptr(mask), None, 3, n, 0, ptr(out), None, C.byref(gl), None)
                assert gl.value == wl.value and np.array_equal(out[: gl.value], want_f[: wl.value])
            else:
                out = np.zeros(n, dtype=np.int64)
                ag.call("ag_take_primitive", 64, ptr(vi), None, 0, n, 32, 1, ptr(idx), None, 0, n, 1, ptr(out), None, None, None, None)
                assert np.array_equal(out, want_t)
        return True

    with ThreadPoolExecutor(8) as pool:
        assert all(pool.map(job, range(16)))
