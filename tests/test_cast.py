"""Numeric casts (the data movement under implicit promotion, compute/exec.go:101-121).

CPU part (-m "not gpu"): pins the restatement oracle/cpu_ref.c:ref_cast_numeric against
  * the reference's own cast_type_numeric_{avx2,sse4} loops (oracle/_ref) on representable inputs,
  * the recorded outputs of those loops (tests/golden/ref_simd_vectors.npz, cast/*),
  * the literal vectors of arrow/compute/cast_test.go:483-629 (tests/golden/cast_numeric.json),
  * the safe-bound table of helpers.go:496-543 / numeric_cast.go:698-729.
GPU part (-m gpu): the CUDA kernels through the C ABI against all of the above."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import ALL_TYPES, INT_TYPES, NP_OF, TYPE_NAME, Dev, cast_inputs, pack_bits, ptr, same_bits, same_float_class

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_simd_vectors.npz"))
CASES = json.load(open(os.path.join(HERE, "golden", "cast_numeric.json")))["cases"]
ID_OF = {v: k for k, v in TYPE_NAME.items()}
PAIRS = [(a, b) for a in ALL_TYPES for b in ALL_TYPES if a != b]
NO_POS = (1 << 63) - 1


def eq_for(t):
    return same_float_class if t in (N.FLOAT32, N.FLOAT64) else same_bits


def oracle_cast(cpu, ti, to, x, valid=None, voff=0, aio=True, aft=True):
    x = np.ascontiguousarray(x)
    o = np.zeros(x.size, dtype=NP_OF[to])
    bad = C.c_int64(-1)
    st = cpu.ref_cast_numeric(ti, to, ptr(x), ptr(valid) if valid is not None else None, voff, ptr(o), x.size, int(aio), int(aft), C.byref(bad))
    return st, o, bad.value


def case_arrays(case):
    ti, to = ID_OF[case["from"]], ID_OF[case["to"]]
    vals = case["in"]
    validity = np.array([v is not None for v in vals])
    for i in case.get("null_at", []):
        validity[i] = False
    x = np.array([0 if v is None else v for v in vals], dtype=NP_OF[ti])
    lo, hi = case.get("slice", [0, len(vals)])
    return ti, to, x, validity, lo, hi


def check_case(case, run):
    """run(ti, to, x_slice, validity_bitmap_or_None, bit_offset, aio, aft) -> (status, out, first_bad)"""
    ti, to, x, validity, lo, hi = case_arrays(case)
    bitmap = None if validity.all() else pack_bits(validity)
    st, out, bad = run(ti, to, x[lo:hi], bitmap, lo, bool(case.get("allow_int_overflow")), bool(case.get("allow_float_truncate")))
    if case.get("fails"):
        assert st == N.AG_ERR_INVALID, case
        assert bad == case["first_bad"] - lo, case
    else:
        assert st == 0 and bad == NO_POS, case
        for i, want in enumerate(case["out"]):
            if want is not None:
                assert out[i] == np.array(want, dtype=NP_OF[to]), (case, i)


# ------------------------------------------------------------------ CPU: the oracle ------
@pytest.mark.parametrize("pair", PAIRS, ids=lambda p: f"{TYPE_NAME[p[0]]}-{TYPE_NAME[p[1]]}")
def test_oracle_matches_recorded_reference_loops(cpu, pair):
    ti, to = pair
    x, want = G[f"cast/{TYPE_NAME[ti]}/{TYPE_NAME[to]}/x"], G[f"cast/{TYPE_NAME[ti]}/{TYPE_NAME[to]}/o"]
    st, o, bad = oracle_cast(cpu, ti, to, x)
    assert st == 0 and bad == NO_POS
    assert eq_for(to)(o, want)


@pytest.mark.parametrize("pair", PAIRS, ids=lambda p: f"{TYPE_NAME[p[0]]}-{TYPE_NAME[p[1]]}")
def test_oracle_matches_reference_loops_live(cpu, ref, pair):
    ti, to = pair
    rng = np.random.default_rng(ti * 31 + to)
    for n in (0, 1, 7, 33, 1000, 4099):
        x = cast_inputs(rng, ti, to, n)
        _, mine, _ = oracle_cast(cpu, ti, to, x)
        for isa in ("avx2", "sse4"):
            want = np.zeros(n, dtype=NP_OF[to])
            getattr(ref, f"cast_type_numeric_{isa}")(ti, to, ptr(x), ptr(want), n)
            assert eq_for(to)(mine, want), (TYPE_NAME[ti], TYPE_NAME[to], n, isa)


def test_oracle_reference_known_answers(cpu):
    def run(ti, to, x, bitmap, off, aio, aft):
        return oracle_cast(cpu, ti, to, x, bitmap, off, aio, aft)
    for case in CASES:
        check_case(case, run)


def safe_bounds(ti, to):
    """Expected [lo, hi] of getSafeMinMax* / checkIntToFloatTrunc, derived independently:
    the intersection of the two types' value ranges (ints) or +-2^mantissa (floats)."""
    ii = np.iinfo(NP_OF[ti])
    if to in (N.FLOAT32, N.FLOAT64):
        m = 24 if to == N.FLOAT32 else 53
        if ii.bits <= m or (ii.bits == 32 and to == N.FLOAT64):
            return None
        return (-(1 << m) if ii.min < 0 else 0), 1 << m
    oi = np.iinfo(NP_OF[to])
    lo, hi = max(ii.min, oi.min), min(ii.max, oi.max)
    return None if (lo == ii.min and hi == ii.max) else (lo, hi)


def boundary_values(ti, to):
    b = safe_bounds(ti, to)
    ii = np.iinfo(NP_OF[ti])
    if b is None:
        return None, [ii.min, ii.max, 0, 1]
    lo, hi = b
    cand = [lo, hi, lo - 1, hi + 1, ii.min, ii.max, 0]
    return b, [v for v in cand if ii.min <= v <= ii.max]


@pytest.mark.parametrize("ti", INT_TYPES, ids=lambda t: TYPE_NAME[t])
def test_oracle_safe_bounds(cpu, ti):
    for to in ALL_TYPES:
        if to == ti:
            continue
        b, vals = boundary_values(ti, to)
        for v in vals:
            x = np.array([0, 0, v, 0], dtype=NP_OF[ti])
            st, _, bad = oracle_cast(cpu, ti, to, x, aio=False, aft=False)
            inside = b is None or b[0] <= v <= b[1]
            assert (st == 0 and bad == NO_POS) if inside else (st != 0 and bad == 2), (TYPE_NAME[ti], TYPE_NAME[to], v)
            # a null slot is never checked
            st, _, bad = oracle_cast(cpu, ti, to, x, pack_bits(np.array([1, 1, 0, 1], dtype=bool)), 0, False, False)
            assert st == 0 and bad == NO_POS


# ------------------------------------------------------------------ GPU -----------------
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("pair", PAIRS, ids=lambda p: f"{TYPE_NAME[p[0]]}-{TYPE_NAME[p[1]]}")
def test_gpu_cast_matches_reference(ag, cpu, ref, isa, pair):
    ti, to = pair
    rng = np.random.default_rng(0xCA57 + ti * 31 + to)
    x, want = G[f"cast/{TYPE_NAME[ti]}/{TYPE_NAME[to]}/x"], G[f"cast/{TYPE_NAME[ti]}/{TYPE_NAME[to]}/o"]
    got = np.zeros(x.size, dtype=NP_OF[to])
    ag.call("ag_cast_numeric", ti, to, ptr(x), ptr(got), x.size)
    assert eq_for(to)(got, want)
    isz, osz = np.dtype(NP_OF[ti]).itemsize, np.dtype(NP_OF[to]).itemsize
    for n in (0, 1, 7, 33, 1000, 4099, (1 << 18) + 5):
        x = cast_inputs(rng, ti, to, n)
        want = np.zeros(n, dtype=NP_OF[to])
        getattr(ref, f"cast_type_numeric_{isa}")(ti, to, ptr(x), ptr(want), n)
        _, mine, _ = oracle_cast(cpu, ti, to, x)
        assert eq_for(to)(mine, want)
        got = np.zeros(n, dtype=NP_OF[to])
        ag.call("ag_cast_numeric", ti, to, ptr(x), ptr(got), n)
        assert eq_for(to)(got, want), (TYPE_NAME[ti], TYPE_NAME[to], n)
        if n:
            for mis in (0, 1, 3):  # element-offset operands: vector and scalar kernels
                dx = Dev(x, byte_offset=mis * isz)
                do = Dev(np.zeros(n, dtype=NP_OF[to]), byte_offset=mis * osz)
                ag.call("ag_cast_numeric_dev", ti, to, dx.ptr, do.ptr, n, None)
                ag.call("ag_stream_sync", None)
                assert eq_for(to)(do.get(), want), (TYPE_NAME[ti], TYPE_NAME[to], n, mis)


def gpu_checked(ag, ti, to, x, bitmap, off, aio, aft):
    x = np.ascontiguousarray(x)
    out = np.zeros(x.size, dtype=NP_OF[to])
    bad = C.c_int64(-1)
    st, _ = ag.call_status("ag_cast_numeric_checked", ti, to, ptr(x), ptr(bitmap) if bitmap is not None else None, off,
                           ptr(out), x.size, int(aio), int(aft), C.byref(bad))
    return st, out, bad.value


@gpu
def test_gpu_reference_known_answers(ag):
    for case in CASES:
        check_case(case, lambda *a: gpu_checked(ag, *a))


@gpu
@pytest.mark.parametrize("ti", INT_TYPES, ids=lambda t: TYPE_NAME[t])
def test_gpu_safe_bounds(ag, ti):
    for to in ALL_TYPES:
        if to == ti:
            continue
        b, vals = boundary_values(ti, to)
        for v in vals:
            x = np.array([0, 0, v, 0], dtype=NP_OF[ti])
            st, _, bad = gpu_checked(ag, ti, to, x, None, 0, False, False)
            inside = b is None or b[0] <= v <= b[1]
            assert (st == 0 and bad == NO_POS) if inside else (st == N.AG_ERR_INVALID and bad == 2), (TYPE_NAME[ti], TYPE_NAME[to], v)
            st, _, bad = gpu_checked(ag, ti, to, x, pack_bits(np.array([1, 1, 0, 1], dtype=bool)), 0, False, False)
            assert st == 0 and bad == NO_POS


@gpu
@pytest.mark.parametrize("pair", [(N.INT64, N.INT32), (N.INT32, N.UINT8), (N.UINT64, N.FLOAT64), (N.INT64, N.FLOAT32),
                                  (N.FLOAT64, N.INT32), (N.FLOAT32, N.INT64), (N.FLOAT64, N.UINT8), (N.UINT32, N.INT16)],
                         ids=lambda p: f"{TYPE_NAME[p[0]]}-{TYPE_NAME[p[1]]}")
def test_gpu_checked_random_vs_oracle(ag, cpu, pair):
    """Random data with planted offenders, nulls at random, bit offsets: status, first failing
    row and (when the cast succeeds) every output slot equal the oracle's."""
    ti, to = pair
    rng = np.random.default_rng(ti * 100 + to)
    for n in (1, 33, 1000, 70001):
        for trial in range(4):
            x = cast_inputs(rng, ti, to, n)
            if NP_OF[ti] in (np.float32, np.float64):
                x = np.trunc(x).astype(NP_OF[ti])      # whole numbers pass; then plant fractions
                bad_val = NP_OF[ti](0.5)
            else:
                b = safe_bounds(ti, to)
                x = np.clip(x, b[0], b[1]).astype(NP_OF[ti]) if b else x
                ii = np.iinfo(NP_OF[ti])
                bad_val = NP_OF[ti](b[1] + 1 if b and b[1] + 1 <= ii.max else (b[0] - 1 if b else 0))
            n_bad = int(rng.integers(0, 3)) if trial else 0
            for p in rng.integers(0, n, n_bad):
                x[p] = bad_val
            off = int(rng.integers(0, 13))
            validity = rng.random(n) > 0.3
            bitmap = pack_bits(validity, offset=off) if trial % 2 else None
            wst, wout, wbad = oracle_cast(cpu, ti, to, x, bitmap, off, False, False)
            st, out, bad = gpu_checked(ag, ti, to, x, bitmap, off, False, False)
            assert (st == 0) == (wst == 0) and bad == wbad, (TYPE_NAME[ti], TYPE_NAME[to], n, trial)
            if st == 0:
                assert eq_for(to)(out, wout)
