"""The oracle (CPU) and the GPU library against tests/golden/ref_simd_vectors.npz — outputs of
the reference's own AVX2 loops recorded by tests/golden/make_golden.py.  These keep both pinned
to the reference's bits on machines where /root/reference and oracle/_ref do not exist."""
import ctypes as C
import os

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import ALL_TYPES, NP_OF, TYPE_NAME, ptr, same_bits, same_float_class

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_simd_vectors.npz"))
SUM_NS = (0, 1, 31, 32, 33, 100, 8192)


def eq_for(t):
    return same_float_class if t in (N.FLOAT32, N.FLOAT64) else same_bits


def test_oracle_sum(cpu):
    for n in SUM_NS:
        for kind in ("exact", "normal"):
            x = G[f"sum_f64/{kind}/{n}/x"]
            assert cpu.ref_sum_f64_avx2_order(ptr(x), n) == G[f"sum_f64/{kind}/{n}/res"][0]
        x = G[f"sum_i64/{n}/x"]
        assert cpu.ref_sum_i64(ptr(x), n) == G[f"sum_i64/{n}/res"][0]


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_oracle_arith_and_compare(cpu, t):
    nm = TYPE_NAME[t]
    l, r = G[f"arith/{nm}/l"], G[f"arith/{nm}/r"]
    n = l.size
    for op in (0, 1, 2):
        for shape in (0, 1, 2):
            o = np.empty(n, dtype=NP_OF[t])
            assert cpu.ref_arith_binary(t, op, shape, ptr(l), ptr(r), ptr(o), n) == 0
            assert eq_for(t)(o, G[f"arith/{nm}/{op}/{shape}"])
    for op in (4, 5, 20, 26):
        o = np.empty(n, dtype=NP_OF[t])
        assert cpu.ref_arith_unary_same(t, op, ptr(l), ptr(o), n) == 0
        assert eq_for(t)(o, G[f"unary/{nm}/{op}"])
    l, r = G[f"cmp/{nm}/l"], G[f"cmp/{nm}/r"]
    for ci in range(4):
        for shape in (0, 1, 2):
            o = np.full(12, 0xA5, dtype=np.uint8)
            assert cpu.ref_compare(t, ci, shape, ptr(l), ptr(r), ptr(o), n, 3) == 0
            assert o.tobytes() == G[f"cmp/{nm}/{ci}/{shape}"].tobytes()


@pytest.mark.gpu
def test_gpu_sum(ag):
    for n in SUM_NS:
        x = G[f"sum_f64/exact/{n}/x"]
        r = C.c_double(-1)
        ag.call("ag_sum_f64", ptr(x), n, C.byref(r))
        assert r.value == G[f"sum_f64/exact/{n}/res"][0]
        for kind in ("exact", "normal"):
            x = G[f"sum_f64/{kind}/{n}/x"]
            ag.call("ag_sum_f64_reforder", ptr(x), n, C.byref(r))
            assert r.value == G[f"sum_f64/{kind}/{n}/res"][0]
        xi = G[f"sum_i64/{n}/x"]
        ri = C.c_int64(-1)
        ag.call("ag_sum_i64", ptr(xi), n, C.byref(ri))
        assert ri.value == G[f"sum_i64/{n}/res"][0]


@pytest.mark.gpu
@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_gpu_arith_and_compare(ag, t):
    nm = TYPE_NAME[t]
    l, r = G[f"arith/{nm}/l"], G[f"arith/{nm}/r"]
    n = l.size
    fns = {0: "ag_arith_binary", 1: "ag_arith_arr_scalar", 2: "ag_arith_scalar_arr"}
    for op in (0, 1, 2):
        for shape in (0, 1, 2):
            o = np.empty(n, dtype=NP_OF[t])
            ag.call(fns[shape], t, op, ptr(l), ptr(r), ptr(o), n)
            assert eq_for(t)(o, G[f"arith/{nm}/{op}/{shape}"]), (nm, op, shape)
    for op in (4, 5, 20, 26):
        o = np.empty(n, dtype=NP_OF[t])
        ag.call("ag_arith_unary_same", t, op, ptr(l), ptr(o), n)
        assert eq_for(t)(o, G[f"unary/{nm}/{op}"]), (nm, op)
    l, r = G[f"cmp/{nm}/l"], G[f"cmp/{nm}/r"]
    for ci in range(4):
        for shape in (0, 1, 2):
            o = np.full(12, 0xA5, dtype=np.uint8)
            ag.call("ag_compare", t, ci, shape, ptr(l), ptr(r), ptr(o), n, 3)
            assert o.tobytes() == G[f"cmp/{nm}/{ci}/{shape}"].tobytes(), (nm, ci, shape)
