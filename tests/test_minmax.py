"""Integer min/max (internal/utils/min_max.go; Parquet statistics reduction).

Shape of the reference's own test (internal/utils/min_max_test.go:27-118): sizes
{0,1,2,3,4,7,8,9,15,16,31,63,64,100,1024}, random values with the type's extremes planted,
native loop compared with the pure-Go loop.  CPU part: restatement vs the reference's
AVX2/SSE4 loops; GPU part: the CUDA kernel vs both."""
import ctypes as C

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import INT_TYPES, NP_OF, TYPE_NAME, Dev, ptr

SIZES = [0, 1, 2, 3, 4, 7, 8, 9, 15, 16, 31, 63, 64, 100, 1024]


def make(rng, t, n, plant=True):
    info = np.iinfo(NP_OF[t])
    x = rng.integers(info.min // 2, info.max // 2, n, dtype=NP_OF[t], endpoint=True)
    if n and plant:
        x[rng.integers(0, n)] = info.min
        x[rng.integers(0, n)] = info.max
    return x


def oracle_minmax(cpu, t, x):
    lo, hi = np.zeros(1, dtype=NP_OF[t]), np.zeros(1, dtype=NP_OF[t])
    assert cpu.ref_min_max(t, ptr(x), x.size, ptr(lo), ptr(hi)) == 0
    return lo[0], hi[0]


@pytest.mark.parametrize("t", INT_TYPES, ids=lambda t: TYPE_NAME[t])
def test_oracle_matches_reference_loops(cpu, ref, t):
    rng = np.random.default_rng(t)
    info = np.iinfo(NP_OF[t])
    for n in SIZES + [4099, 100_003]:
        for plant in (True, False):
            x = make(rng, t, n, plant)
            lo, hi = oracle_minmax(cpu, t, x)
            if n == 0:
                assert (lo, hi) == (info.max, info.min)      # min_max.c:24-25 initial values
            else:
                assert (lo, hi) == (x.min(), x.max())
            for isa in ("avx2", "sse4"):
                rlo, rhi = np.zeros(1, dtype=NP_OF[t]), np.zeros(1, dtype=NP_OF[t])
                getattr(ref, f"{TYPE_NAME[t]}_max_min_{isa}")(ptr(x), n, ptr(rlo), ptr(rhi))
                assert (rlo[0], rhi[0]) == (lo, hi), (TYPE_NAME[t], n, isa)


def test_oracle_rejects_floats(cpu):
    x = np.zeros(4)
    assert cpu.ref_min_max(N.FLOAT64, ptr(x), 4, ptr(x), ptr(x)) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("t", INT_TYPES, ids=lambda t: TYPE_NAME[t])
def test_gpu_min_max(ag, cpu, t):
    rng = np.random.default_rng(100 + t)
    isz = np.dtype(NP_OF[t]).itemsize
    for n in SIZES + [4099, (1 << 20) + 7, 5_000_011]:
        for plant in (True, False):
            x = make(rng, t, n, plant)
            want = oracle_minmax(cpu, t, x)
            lo, hi = np.zeros(1, dtype=NP_OF[t]), np.zeros(1, dtype=NP_OF[t])
            ag.call("ag_min_max", t, ptr(x), n, ptr(lo), ptr(hi))
            assert (lo[0], hi[0]) == want, (TYPE_NAME[t], n)
            for mis in (0, 1, 3):  # Arrow slices: element-aligned starts
                dx = Dev(x, byte_offset=mis * isz)
                do = Dev(np.zeros(2, dtype=NP_OF[t]))
                ag.call("ag_min_max_dev", t, dx.ptr, n, do.ptr, None)
                ag.call("ag_stream_sync", None)
                got = do.get()
                assert (got[0], got[1]) == want, (TYPE_NAME[t], n, mis)
    # extremes far from the planted positions: single outlier at the very end / start
    x = np.full(3_000_001, 5, dtype=NP_OF[t])
    x[-1], x[0] = 1, 9
    lo, hi = np.zeros(1, dtype=NP_OF[t]), np.zeros(1, dtype=NP_OF[t])
    ag.call("ag_min_max", t, ptr(x), x.size, ptr(lo), ptr(hi))
    assert (lo[0], hi[0]) == (1, 9)


@pytest.mark.gpu
def test_gpu_min_max_errors(ag):
    x = np.zeros(4)
    st, msg = ag.call_status("ag_min_max", N.FLOAT64, ptr(x), 4, ptr(x), ptr(x))
    assert st == N.AG_ERR_TYPE
    st, msg = ag.call_status("ag_min_max", N.INT32, None, 4, ptr(x), ptr(x))
    assert st == N.AG_ERR_INVALID
