"""sort_indices: the oracle restatement pinned against the reference's literal test vectors
(arrow/compute/vector_sort_test.go:40-325, tests/golden/sort_indices.json), against numpy's stable argsort and against
pyarrow (Arrow C++ is what the reference's kernel is ported from); then the GPU radix sort against the oracle."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import ALL_TYPES, NP_OF, TYPE_NAME, Dev, pack_bits, ptr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sort_indices.json")
TYPE_ID = {v: k for k, v in TYPE_NAME.items()}
gpu = pytest.mark.gpu


def oracle_sort(cpu, t, vals, valid_bits, voff, n, order, placement):
    out = np.zeros(max(n, 1), dtype=np.uint64)
    nn, na = C.c_int64(), C.c_int64()
    assert cpu.ref_sort_indices(t, ptr(vals), ptr(valid_bits), voff, n, order, placement, ptr(out), C.byref(nn), C.byref(na)) == 0
    return out[:n], nn.value, na.value


def literal(case):
    dt = NP_OF[TYPE_ID[case["type"]]]
    vals = np.array([0 if v is None else (np.nan if v == "NaN" else v) for v in case["values"]], dtype=dt)
    valid = np.array([v is not None for v in case["values"]], dtype=bool)
    return vals, (pack_bits(valid) if not valid.all() else None)


def test_oracle_matches_reference_literals(cpu):
    for case in json.load(open(GOLDEN))["cases"]:
        vals, bits = literal(case)
        got, _, _ = oracle_sort(cpu, TYPE_ID[case["type"]], vals, bits, 0, vals.size, case["order"], case["null_placement"])
        assert got.tolist() == case["expected"], case["name"]


def model_sort(vals, valid, order, placement):
    """Independent numpy model: stable argsort of the finite rows, NaNs and nulls appended / prepended in row order."""
    idx = np.arange(vals.size)
    isnan = np.isnan(vals) if vals.dtype.kind == "f" else np.zeros(vals.size, dtype=bool)
    fin = idx[valid & ~isnan]
    key = vals[fin]
    if vals.dtype.kind == "f":
        key = key + 0.0          # -0.0 == +0.0 either way for a comparison sort
    o = np.argsort(-key.astype(np.float64) if (order and vals.dtype.kind == "f") else key, kind="stable")
    if order and vals.dtype.kind != "f":
        # descending with ties in row order: stable sort of the negated rank
        ranks = np.unique(key, return_inverse=True)[1]
        o = np.argsort(-ranks, kind="stable")
    fin = fin[o]
    nan, nul = idx[valid & isnan], idx[~valid]
    return np.concatenate([nul, nan, fin] if placement else [fin, nan, nul]).astype(np.uint64)


def random_column(rng, t, n, p_null, small):
    dt = np.dtype(NP_OF[t])
    if dt.kind == "f":
        v = rng.standard_normal(n).astype(dt)
        if small:
            v = np.round(v * 2) / 2                      # many ties
        if n > 8:
            v[rng.integers(0, n, max(1, n // 50))] = np.nan
            v[rng.integers(0, n, 3)] = [0.0, -0.0, np.inf]
    elif small:
        v = rng.integers(0, 7, n).astype(dt)
    else:
        info = np.iinfo(dt)
        v = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    valid = rng.random(n) >= p_null if p_null else np.ones(n, dtype=bool)
    return v, valid


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_oracle_matches_numpy_model_and_pyarrow(cpu, t):
    import pyarrow as pa
    import pyarrow.compute as pc
    rng = np.random.default_rng(t)
    for n in (1, 2, 33, 1000, 20_011):
        for small in (True, False):
            for p_null in (0.0, 0.2):
                v, valid = random_column(rng, t, n, p_null, small)
                bits = pack_bits(valid, 3) if p_null else None
                vv = np.concatenate([np.zeros(3, dtype=v.dtype), v]) if p_null else v
                for order in (0, 1):
                    for placement in (0, 1):
                        got, nn, _ = oracle_sort(cpu, t, vv, bits, 3 if p_null else 0, n, order, placement)
                        assert np.array_equal(got, model_sort(v, valid, order, placement)), (TYPE_NAME[t], n, small, p_null, order, placement)
                        assert nn == int((~valid).sum())
                        arr = pa.array(v, mask=~valid)
                        want = pc.array_sort_indices(arr, order="descending" if order else "ascending",
                                                     null_placement="at_start" if placement else "at_end")
                        assert np.array_equal(got, want.to_numpy().astype(np.uint64)), ("pyarrow", TYPE_NAME[t], n, order, placement)


# ------------------------------------------------------------------------------------------------- GPU
def gpu_sort(ag, t, vals, bits, voff, n, order, placement, dev=False):
    nn, na = C.c_int64(), C.c_int64()
    if not dev:
        out = np.full(max(n, 1), 0xEE, dtype=np.uint64)
        ag.call("ag_sort_indices", t, ptr(vals), ptr(bits), voff, n, order, placement, ptr(out), C.byref(nn), C.byref(na))
        return out[:n], nn.value, na.value
    dv = Dev(vals)
    db = Dev(bits) if bits is not None else None
    do = Dev(np.zeros(max(n, 1), dtype=np.uint64))
    ag.call("ag_sort_indices_dev", t, dv.ptr, db.ptr if db else None, voff, n, order, placement, do.ptr, C.byref(nn), C.byref(na), None)
    ag.call("ag_stream_sync", None)
    return do.get()[:n], nn.value, na.value


@gpu
def test_gpu_reference_literals(ag):
    for case in json.load(open(GOLDEN))["cases"]:
        vals, bits = literal(case)
        if vals.size == 0:
            vals = np.zeros(1, dtype=vals.dtype)
            n = 0
        else:
            n = vals.size
        for dev in (False, True):
            got, _, _ = gpu_sort(ag, TYPE_ID[case["type"]], vals, bits, 0, n, case["order"], case["null_placement"], dev)
            assert got.tolist() == case["expected"], (case["name"], dev)


@gpu
@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_gpu_matches_oracle(ag, cpu, t):
    rng = np.random.default_rng(100 + t)
    for n in (1, 2, 31, 32, 33, 4095, 4096, 4097, 70_001, 300_007):
        for small in (True, False):
            for p_null in (0.0, 0.3, 1.0):
                v, valid = random_column(rng, t, n, p_null, small)
                voff = int(rng.integers(0, 9)) if p_null else 0
                bits = pack_bits(valid, voff) if p_null else None
                vv = np.concatenate([np.zeros(voff, dtype=v.dtype), v])
                for order in (0, 1):
                    for placement in (0, 1):
                        want, wn, wa = oracle_sort(cpu, t, vv, bits, voff, n, order, placement)
                        for dev in (False, True):
                            got, gn, ga = gpu_sort(ag, t, vv, bits, voff, n, order, placement, dev)
                            assert np.array_equal(got, want), (TYPE_NAME[t], n, small, p_null, order, placement, dev)
                            assert (gn, ga) == (wn, wa)


@gpu
def test_gpu_100m_rows_properties(ag):
    """BASELINE-size column through size-independent properties: the output is a permutation of 0..n-1, the gathered
    keys are non-decreasing, equal keys keep row order (checked on the device-generated int64 column by bringing the
    sorted keys of three 4M-row windows back), and a small-range column (one radix pass) agrees with numpy on a
    2M-row prefix view of its own."""
    n = 100_000_000
    v = Dev(nbytes=n * 8)
    ag.call("ag_generate_dev", 1, 0x5027, -(1 << 31), (1 << 31) - 1, v.ptr, n, None)
    out = Dev(nbytes=n * 8)
    nn, na = C.c_int64(), C.c_int64()
    ag.call("ag_sort_indices_dev", N.INT64, v.ptr, None, 0, n, 0, 0, out.ptr, C.byref(nn), C.byref(na), None)
    ag.call("ag_stream_sync", None)
    # permutation: the wrapping sum and the xor-free checksum of the indices equal those of 0..n-1
    s = Dev(np.zeros(1, dtype=np.int64))
    ag.call("ag_sum_i64_dev", out.ptr, n, s.ptr, None)
    ag.call("ag_stream_sync", None)
    assert int(s.get()[0]) == n * (n - 1) // 2
    hv = v.buf.to_numpy(np.int64, n)
    for start in (0, 48_000_000, n - 4_000_000):
        idx = out.buf.to_numpy(np.uint64, 4_000_000, start * 8)
        keys = hv[idx]
        assert np.all(keys[1:] >= keys[:-1])
        ties = keys[1:] == keys[:-1]
        assert np.all(idx[1:][ties] > idx[:-1][ties]), "equal keys must keep row order"
    m = 2_000_000
    want = np.argsort(hv[:m], kind="stable").astype(np.uint64)
    ag.call("ag_sort_indices_dev", N.INT64, v.ptr, None, 0, m, 0, 0, out.ptr, C.byref(nn), C.byref(na), None)
    ag.call("ag_stream_sync", None)
    assert np.array_equal(out.buf.to_numpy(np.uint64, m), want)


@gpu
def test_gpu_key_range_reduction(ag, cpu):
    """The sorted key is (key' - min) >> (trailing bits every key shares): columns built to stress that transform —
    values straddling zero (the sign-biased image straddles 2^63), values packed against the type's limits, doubles
    holding integers (32 constant low mantissa bits), multiples of 2^k, all rows equal, two distinct values, one finite row
    among nulls — with and without NaN / null rows (class pass on / off), every order and placement."""
    rng = np.random.default_rng(4242)
    n = 50_021
    cols = []
    for dt in (np.int64, np.int32, np.int8):
        info = np.iinfo(dt)
        cols += [(N.INT64 if dt is np.int64 else N.INT32 if dt is np.int32 else N.INT8, c) for c in (
            rng.integers(-100, 100, n).astype(dt),
            (info.max - rng.integers(0, 100, n)).astype(dt),
            (info.min + rng.integers(0, 100, n)).astype(dt),
            np.where(rng.random(n) < 0.5, info.min, info.max).astype(dt),
            np.full(n, -7, dtype=dt),
        )]
    cols += [(N.INT64, (rng.integers(-3000, 3000, n) << 20).astype(np.int64)),
             (N.UINT64, (np.uint64(2**64 - 1) - rng.integers(0, 70_000, n).astype(np.uint64))),
             (N.UINT64, rng.integers(0, 2**63, n).astype(np.uint64) * np.uint64(2)),
             (N.UINT16, rng.integers(65_000, 65_536, n).astype(np.uint16)),
             (N.FLOAT64, rng.integers(-(1 << 20), 1 << 20, n).astype(np.float64)),
             (N.FLOAT64, rng.integers(-5, 5, n).astype(np.float64) * 0.0),          # -0.0 and +0.0 only: one key
             (N.FLOAT32, rng.integers(-300, 300, n).astype(np.float32) * np.float32(0.5)),
             (N.FLOAT64, np.where(rng.random(n) < 0.5, -np.inf, np.inf))]
    for t, v in cols:
        for p_null in (0.0, 0.3, 0.9999):
            valid = rng.random(n) >= p_null
            vv = v.copy()
            if vv.dtype.kind == "f" and p_null:
                vv[rng.integers(0, n, 50)] = np.nan
            bits = pack_bits(valid, 1) if p_null else None
            col = np.concatenate([np.zeros(1 if p_null else 0, dtype=vv.dtype), vv])
            for order in (0, 1):
                for placement in (0, 1):
                    want, wn, wa = oracle_sort(cpu, t, col, bits, 1 if p_null else 0, n, order, placement)
                    got, gn, ga = gpu_sort(ag, t, col, bits, 1 if p_null else 0, n, order, placement, True)
                    assert np.array_equal(got, want) and (gn, ga) == (wn, wa), (TYPE_NAME[t], str(vv[:3]), p_null, order, placement)
