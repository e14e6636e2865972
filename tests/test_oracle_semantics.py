"""Pins the Go-only part of the oracle (filter, take, take-indices, Kleene logic, checked integer
arithmetic) — logic with no native counterpart in the reference — against
  * the literal known-answer cases of the reference's Go tests (tests/golden/*.json, restated
    from arrow/compute/{arithmetic,scalar_bool,vector_selection}_test.go), and
  * pyarrow (Arrow C++), an independent implementation of the same Arrow semantics.
"""
import ctypes as C
import json
import os

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from arrow_go_b200 import _native as N
from helpers import INT_TYPES, NP_OF, TYPE_NAME, pack_bits, ptr, random_values, unpack_bits

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WDT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def to_np(lst, dtype):
    """JSON list with nulls -> (values, validity-bools)."""
    valid = np.array([x is not None for x in lst], dtype=bool)
    vals = np.array([0 if x is None else x for x in lst], dtype=dtype)
    return vals, valid


# ------------------------------------------------------------------ filter ---------
def oracle_filter(cpu, values, vvalid, mask, mvalid, sel, voff=0, moff=0, bw=32):
    n = len(mask)
    vals = np.concatenate([np.zeros(voff, dtype=WDT[bw]), values.astype(WDT[bw])])
    vv = pack_bits(vvalid, voff) if vvalid is not None else None
    m = pack_bits(mask, moff)
    mv = pack_bits(mvalid, moff) if mvalid is not None else None
    out = np.zeros(n + 1, dtype=WDT[bw]); ov = np.zeros(n // 8 + 2, dtype=np.uint8)
    ln, nulls = C.c_int64(), C.c_int64()
    assert cpu.ref_filter_primitive(bw, ptr(vals), ptr(vv), voff, ptr(m), ptr(mv), moff, n, sel, ptr(out), ptr(ov), C.byref(ln), C.byref(nulls)) == 0
    assert ln.value == cpu.ref_filter_output_size(ptr(m), ptr(mv), moff, n, sel)
    return out[: ln.value], unpack_bits(ov, 0, ln.value), nulls.value


def test_filter_golden(cpu):
    for case in load("filter_numeric.json")["cases"]:
        vals, vvalid = to_np(case["values"], np.int32)
        mask, mvalid = to_np(case["filter"], bool)
        for sel_name, sel in (("drop", 0), ("emit_null", 1)):
            exp_vals, exp_valid = to_np(case[sel_name], np.int32)
            for voff, moff in ((0, 0), (3, 2)):  # "sliced" re-run: values padded by 3, mask by 2 (vector_selection_test.go:114-145)
                got, gvalid, nulls = oracle_filter(cpu, vals.view(np.uint32), None if vvalid.all() else vvalid, mask,
                                                   None if mvalid.all() else mvalid, sel, voff, moff)
                assert len(got) == len(exp_vals), case
                if not (vvalid.all() and mvalid.all()):
                    assert gvalid.tolist() == exp_valid.tolist(), (case, sel_name)
                    assert nulls == int((~exp_valid).sum())
                assert got.view(np.int32)[exp_valid].tolist() == exp_vals[exp_valid].tolist(), (case, sel_name)


def test_filter_vs_pyarrow(cpu):
    rng = np.random.default_rng(0x0FF1CE)  # vector_selection_test.go:41
    for n in (0, 1, 8, 63, 64, 65, 200, 512, 5000):
        for p_mask in (0.0, 0.2, 0.5, 1.0):
            for p_vnull, p_mnull in ((0, 0), (0.2, 0), (0, 0.2), (0.3, 0.3)):
                values = rng.integers(-1000, 1000, n).astype(np.int64)
                vvalid = rng.random(n) >= p_vnull if p_vnull else None
                mask = rng.random(n) < p_mask
                mvalid = rng.random(n) >= p_mnull if p_mnull else None
                pv = pa.array(values, mask=None if vvalid is None else ~vvalid)
                pm = pa.array(mask, mask=None if mvalid is None else ~mvalid)
                for sel, name in ((0, "drop"), (1, "emit_null")):
                    want = pc.filter(pv, pm, null_selection_behavior=name)
                    got, gvalid, nulls = oracle_filter(cpu, values.view(np.uint64), vvalid, mask, mvalid, sel, 5, 3, bw=64)
                    assert len(got) == len(want)
                    wvalid = np.array([x.is_valid for x in want], dtype=bool)
                    if vvalid is not None or mvalid is not None:
                        assert gvalid.tolist() == wvalid.tolist()
                        assert nulls == want.null_count
                    wvals = np.array([x.as_py() if x.is_valid else 0 for x in want], dtype=np.int64)
                    assert got.view(np.int64)[wvalid].tolist() == wvals[wvalid].tolist()


def test_take_indices_vs_numpy(cpu):
    rng = np.random.default_rng(2)
    for n in (1, 64, 65, 1000):
        for p_mnull in (0, 0.3):
            mask = rng.random(n) < 0.4
            mvalid = rng.random(n) >= p_mnull if p_mnull else None
            m, mv = pack_bits(mask, 3), (pack_bits(mvalid, 3) if mvalid is not None else None)
            for sel in (0, 1):
                out = np.zeros(n, dtype=np.uint16); ov = np.zeros(n // 8 + 2, dtype=np.uint8); ln = C.c_int64()
                assert cpu.ref_take_indices(16, ptr(m), ptr(mv), 3, n, sel, ptr(out), ptr(ov), C.byref(ln)) == 0
                v = np.ones(n, bool) if mvalid is None else mvalid
                exp = [(i, True) for i in range(n) if mask[i] and v[i]] if sel == 0 else \
                      [(i if v[i] else 0, bool(v[i])) for i in range(n) if (mask[i] and v[i]) or not v[i]]
                assert ln.value == len(exp)
                assert out[: ln.value].tolist() == [e[0] for e in exp]
                assert unpack_bits(ov, 0, ln.value).tolist() == [e[1] for e in exp]


# ------------------------------------------------------------------ take -----------
def oracle_take(cpu, values, vvalid, idx, ivalid, bw=32, voff=0, ioff=0, signed=1):
    n, vlen = len(idx), len(values)
    vals = np.concatenate([np.zeros(voff, dtype=WDT[bw]), values.astype(WDT[bw])])
    vv = pack_bits(vvalid, voff) if vvalid is not None else None
    iv = pack_bits(ivalid, ioff) if ivalid is not None else None
    out = np.zeros(max(n, 1), dtype=WDT[bw]); ov = np.zeros(n // 8 + 2, dtype=np.uint8)
    nulls, bp, bi = C.c_int64(), C.c_int64(), C.c_int64()
    st = cpu.ref_take_primitive(bw, ptr(vals), ptr(vv), voff, vlen, idx.dtype.itemsize * 8, signed, ptr(idx), ptr(iv), ioff, n, 1,
                                ptr(out), ptr(ov), C.byref(nulls), C.byref(bp), C.byref(bi))
    return st, out[:n], unpack_bits(ov, 0, n), nulls.value, bp.value, bi.value


def test_take_golden(cpu):
    g = load("take_numeric.json")
    for case in g["cases"]:
        vals, vvalid = to_np(case["values"], np.int32)
        for idt in (np.int32, np.int8, np.uint32):  # checkTake re-runs with int8 and uint32 indices (:213-253)
            if idt == np.uint32 and any((x is not None and x < 0) for x in case["indices"]):
                continue
            idx, ivalid = to_np(case["indices"], idt)
            st, got, gvalid, nulls, bp, bi = oracle_take(cpu, vals.view(np.uint32), None if vvalid.all() else vvalid, idx,
                                                         None if ivalid.all() else ivalid, voff=2, ioff=1, signed=int(idt != np.uint32))
            if "error_index" in case:
                assert st == 2 and bi == case["error_index"], case
                continue
            assert st == 0
            exp_vals, exp_valid = to_np(case["expected"], np.int32)
            if not (vvalid.all() and ivalid.all()):
                assert gvalid.tolist() == exp_valid.tolist(), case
            assert got.view(np.int32)[exp_valid].tolist() == exp_vals[exp_valid].tolist(), case
            assert (got[~exp_valid] == 0).all()  # untouched slots keep the allocator's zero


def test_take_vs_pyarrow(cpu):
    rng = np.random.default_rng(3)
    for vlen in (1, 10, 1000):
        for n in (0, 1, 64, 65, 500):
            for p_vnull, p_inull in ((0, 0), (0.2, 0), (0, 0.2), (0.3, 0.3)):
                values = rng.integers(-1000, 1000, vlen).astype(np.int64)
                vvalid = rng.random(vlen) >= p_vnull if p_vnull else None
                idx = rng.integers(0, vlen, n).astype(np.int32)
                ivalid = rng.random(n) >= p_inull if p_inull else None
                pv = pa.array(values, mask=None if vvalid is None else ~vvalid)
                pi = pa.array(idx, mask=None if ivalid is None else ~ivalid)
                want = pc.take(pv, pi)
                st, got, gvalid, nulls, _, _ = oracle_take(cpu, values.view(np.uint64), vvalid, idx, ivalid, bw=64, voff=3, ioff=5)
                assert st == 0
                wvalid = np.array([x.is_valid for x in want], dtype=bool)
                if vvalid is not None or ivalid is not None:
                    assert gvalid.tolist() == wvalid.tolist()
                    assert nulls == want.null_count
                wvals = np.array([x.as_py() if x.is_valid else 0 for x in want], dtype=np.int64)
                assert got.view(np.int64)[wvalid].tolist() == wvals[wvalid].tolist()


# ------------------------------------------------------------------ Kleene ----------
def test_kleene_vs_pyarrow(cpu):
    rng = np.random.default_rng(4)
    fns = {N.KLEENE_AND: pc.and_kleene, N.KLEENE_OR: pc.or_kleene, N.KLEENE_ANDNOT: pc.and_not_kleene}
    for n in (1, 9, 64, 65, 1000):
        ld, rd = rng.random(n) < 0.5, rng.random(n) < 0.5
        lv, rv = rng.random(n) < 0.7, rng.random(n) < 0.7
        for kop, fn in fns.items():
            want = fn(pa.array(ld, mask=~lv), pa.array(rd, mask=~rv))
            ov, od = np.zeros(n // 8 + 2, dtype=np.uint8), np.zeros(n // 8 + 2, dtype=np.uint8)
            blv, bld, brv, brd = pack_bits(lv, 3), pack_bits(ld, 3), pack_bits(rv, 6), pack_bits(rd, 6)
            assert cpu.ref_kleene(kop, ptr(blv), ptr(bld), 3, ptr(brv), ptr(brd), 6, ptr(ov), ptr(od), 1, n) == 0
            gv, gd = unpack_bits(ov, 1, n), unpack_bits(od, 1, n)
            wv = np.array([x.is_valid for x in want])
            assert gv.tolist() == wv.tolist()
            wd = np.array([bool(x.as_py()) if x.is_valid else False for x in want])
            assert gd[wv].tolist() == wd[wv].tolist()


# ------------------------------------------------------------------ arithmetic ------
def test_arithmetic_golden(cpu):
    """arithmetic_test.go TestAdd/TestSub/TestMultiply literal cases (:325-425) at the kernel
    level: values in valid slots must match; for the checked integer kernels null slots are 0."""
    g = load("arithmetic.json")
    for type_id in (N.INT8, N.UINT8, N.INT16, N.INT32, N.UINT32, N.INT64, N.UINT64, N.FLOAT32, N.FLOAT64):
        dt = NP_OF[type_id]
        isf = type_id in (N.FLOAT32, N.FLOAT64)
        for case in g["cases"]:
            op = {"add": (N.OP_ADD, N.OP_ADD_CHECKED), "subtract": (N.OP_SUB, N.OP_SUB_CHECKED), "multiply": (N.OP_MUL, N.OP_MUL_CHECKED)}[case["op"]]
            left, right = case["left"], case["right"]
            shape = N.SHAPE_SA if not isinstance(left, list) else (N.SHAPE_AS if not isinstance(right, list) else N.SHAPE_AA)
            exp, evalid = to_np(case["expected"], dt)
            n = len(exp)
            l, lvalid = to_np(left if isinstance(left, list) else [left], dt)
            r, rvalid = to_np(right if isinstance(right, list) else [right], dt)
            if (shape == N.SHAPE_SA and not lvalid[0]) or (shape == N.SHAPE_AS and not rvalid[0]):
                continue  # null scalar: the executor answers without calling the kernel
            out = np.zeros(n, dtype=dt)
            assert cpu.ref_arith_binary(type_id, op[0], shape, ptr(l), ptr(r), ptr(out), n) == 0
            assert out[evalid].tolist() == exp[evalid].tolist(), (TYPE_NAME[type_id], case)
            if not isf:
                out = np.full(n, 77, dtype=dt); bad = C.c_int64()
                lv = pack_bits(lvalid) if shape != N.SHAPE_SA and not lvalid.all() else None
                rv = pack_bits(rvalid) if shape != N.SHAPE_AS and not rvalid.all() else None
                assert cpu.ref_arith_checked(type_id, op[1], shape, ptr(l), ptr(lv), 0, ptr(r), ptr(rv), 0, ptr(out), n, C.byref(bad)) == 0
                assert out[evalid].tolist() == exp[evalid].tolist()
                if case["op"] != "multiply":
                    assert (out[~evalid] == 0).all()


@pytest.mark.parametrize("type_id", INT_TYPES, ids=lambda t: TYPE_NAME[t])
def test_checked_overflow_cases(cpu, type_id):
    """arithmetic_test.go:352-355 (max+max -> "overflow"), :388-390 (min-max -> "overflow"),
    plus mulWithOverflow against exact Python integers on random operands."""
    dt = NP_OF[type_id]
    info = np.iinfo(dt)
    mx, mn = np.array([info.max], dtype=dt), np.array([info.min], dtype=dt)
    out = np.zeros(1, dtype=dt); bad = C.c_int64()
    assert cpu.ref_arith_checked(type_id, N.OP_ADD_CHECKED, 0, ptr(mx), None, 0, ptr(mx), None, 0, ptr(out), 1, C.byref(bad)) == 1 and bad.value == 0
    assert cpu.ref_arith_checked(type_id, N.OP_SUB_CHECKED, 0, ptr(mn), None, 0, ptr(mx), None, 0, ptr(out), 1, C.byref(bad)) == 1
    rng = np.random.default_rng(type_id)
    n = 2000
    a, b = random_values(rng, type_id, n), random_values(rng, type_id, n)
    a[: n // 2] = random_values(rng, type_id, n // 2, small=True)
    b[: n // 2] = random_values(rng, type_id, n // 2, small=True)
    out = np.zeros(n, dtype=dt)
    st = cpu.ref_arith_checked(type_id, N.OP_MUL_CHECKED, 0, ptr(a), None, 0, ptr(b), None, 0, ptr(out), n, C.byref(bad))
    exact = [int(x) * int(y) for x, y in zip(a, b)]
    ovf = [not (info.min <= e <= info.max) for e in exact]
    assert (st == 1) == any(ovf)
    if any(ovf):
        assert bad.value == ovf.index(True)
    for i in range(n):
        assert int(out[i]) == (0 if ovf[i] else exact[i])
    # unsigned add / sub: carry == exact overflow
    if info.min == 0:
        for op, f in ((N.OP_ADD_CHECKED, lambda x, y: x + y), (N.OP_SUB_CHECKED, lambda x, y: x - y)):
            st = cpu.ref_arith_checked(type_id, op, 0, ptr(a), None, 0, ptr(b), None, 0, ptr(out), n, C.byref(bad))
            ovf = [not (0 <= f(int(x), int(y)) <= info.max) for x, y in zip(a, b)]
            assert (st == 1) == any(ovf) and (not any(ovf) or bad.value == ovf.index(True))
    # divide: zero divisor is the only error; MinInt / -1 wraps like Go
    b2 = b.copy(); b2[b2 == 0] = 1
    assert cpu.ref_arith_checked(type_id, N.OP_DIV, 0, ptr(a), None, 0, ptr(b2), None, 0, ptr(out), n, C.byref(bad)) == 0
    for i in range(0, n, 37):
        x, y = int(a[i]), int(b2[i])
        q = abs(x) // abs(y) * (1 if (x >= 0) == (y >= 0) else -1)  # truncation toward zero
        q = (q - info.min) % (info.max - info.min + 1) + info.min
        assert int(out[i]) == q
    b2[5] = 0
    assert cpu.ref_arith_checked(type_id, N.OP_DIV_CHECKED, 0, ptr(a), None, 0, ptr(b2), None, 0, ptr(out), n, C.byref(bad)) == 1 and bad.value == 5


def test_boolean_values_filter_take_vs_pyarrow(cpu):
    """bit-width-1 values: boolFilterWriter / booleanTakeImpl (vector_selection.go:423-447, 990-1074)
    restated with the documented row semantics; pyarrow is the independent check."""
    rng = np.random.default_rng(9)
    for n in (1, 9, 64, 65, 1000):
        vals = rng.random(n) < 0.5
        vvalid = rng.random(n) < 0.8
        mask = rng.random(n) < 0.4
        mvalid = rng.random(n) < 0.85
        voff, moff = 5, 3
        bv, bvv, bm, bmv = pack_bits(vals, voff), pack_bits(vvalid, voff), pack_bits(mask, moff), pack_bits(mvalid, moff)
        for sel, name in ((0, "drop"), (1, "emit_null")):
            want = pc.filter(pa.array(vals, mask=~vvalid), pa.array(mask, mask=~mvalid), null_selection_behavior=name)
            out = np.zeros(n // 8 + 2, dtype=np.uint8); ov = np.zeros(n // 8 + 2, dtype=np.uint8); ln, nulls = C.c_int64(), C.c_int64()
            assert cpu.ref_filter_primitive(1, ptr(bv), ptr(bvv), voff, ptr(bm), ptr(bmv), moff, n, sel, ptr(out), ptr(ov), C.byref(ln), C.byref(nulls)) == 0
            assert ln.value == len(want) and nulls.value == want.null_count
            wv = np.array([x.is_valid for x in want], dtype=bool)
            assert unpack_bits(ov, 0, ln.value).tolist() == wv.tolist()
            wd = np.array([bool(x.as_py()) if x.is_valid else False for x in want], dtype=bool)
            assert unpack_bits(out, 0, ln.value)[wv].tolist() == wd[wv].tolist()
        idx = rng.integers(0, n, 2 * n).astype(np.int32)
        ivalid = rng.random(2 * n) < 0.9
        want = pc.take(pa.array(vals, mask=~vvalid), pa.array(idx, mask=~ivalid))
        out = np.zeros(2 * n // 8 + 2, dtype=np.uint8); ov = np.zeros(2 * n // 8 + 2, dtype=np.uint8)
        nulls, bp, bi = C.c_int64(), C.c_int64(), C.c_int64()
        biv = pack_bits(ivalid, 2)
        assert cpu.ref_take_primitive(1, ptr(bv), ptr(bvv), voff, n, 32, 1, ptr(idx), ptr(biv), 2, 2 * n, 1, ptr(out), ptr(ov),
                                      C.byref(nulls), C.byref(bp), C.byref(bi)) == 0
        wv = np.array([x.is_valid for x in want], dtype=bool)
        assert unpack_bits(ov, 0, 2 * n).tolist() == wv.tolist() and nulls.value == want.null_count
        wd = np.array([bool(x.as_py()) if x.is_valid else False for x in want], dtype=bool)
        assert unpack_bits(out, 0, 2 * n)[wv].tolist() == wd[wv].tolist()


# ---------------------------------------------------------------- casts / cumulative sum vs pyarrow ----
PA_OF = {N.UINT8: pa.uint8(), N.INT8: pa.int8(), N.UINT16: pa.uint16(), N.INT16: pa.int16(), N.UINT32: pa.uint32(), N.INT32: pa.int32(),
         N.UINT64: pa.uint64(), N.INT64: pa.int64(), N.FLOAT32: pa.float32(), N.FLOAT64: pa.float64()}


def _oracle_cast(cpu, ti, to, x, valid, aio, aft):
    out = np.zeros(x.size, dtype=NP_OF[to])
    bad = C.c_int64(-1)
    bm = pack_bits(valid, offset=3) if valid is not None else None
    st = cpu.ref_cast_numeric(ti, to, ptr(x), ptr(bm) if bm is not None else None, 3, ptr(out), x.size, int(aio), int(aft), C.byref(bad))
    return st, out, bad.value


def test_safe_cast_accept_reject_vs_pyarrow(cpu):
    """Arrow C++ (pyarrow) is an independent implementation of the same safe-cast rules the Go code ports
    (int range, |int| <= 2^24 / 2^53 for floats, float -> int truncation): for boundary and random inputs the
    restatement must accept / reject exactly the same calls, and agree on the values of the accepted ones."""
    rng = np.random.default_rng(2024)
    all_t = list(PA_OF)
    for ti in all_t:
        for to in all_t:
            if ti == to:
                continue
            idt, odt = np.dtype(NP_OF[ti]), np.dtype(NP_OF[to])
            cands = []
            if idt.kind != "f":
                ii = np.iinfo(idt)
                pts = {ii.min, ii.max, 0, 1}
                if odt.kind != "f":
                    oi = np.iinfo(odt)
                    pts |= {v for v in (oi.min, oi.max, oi.min - 1, oi.max + 1) if ii.min <= v <= ii.max}
                else:
                    m = 24 if odt.itemsize == 4 else 53
                    pts |= {v for v in ((1 << m), (1 << m) + 1, -(1 << m), -(1 << m) - 1) if ii.min <= v <= ii.max}
                cands = [np.array([0, p, 0], dtype=idt) for p in sorted(pts)]
            else:
                oi = np.iinfo(odt) if odt.kind != "f" else None
                vals = [0.0, 1.0, -1.0, 1.5, -0.5, 255.0, 256.0, -129.0, 65535.0, 3e9, -3e9, 1e19, float("nan")] if oi else [0.0, 1.5, 1e39, -1e39, float("inf"), float("nan")]
                with np.errstate(over="ignore"):  # 1e39 -> float32 is inf on purpose
                    cands = [np.array([0, v, 0], dtype=idt) for v in vals]
            for x in cands:
                for valid in (None, np.array([True, False, True])):
                    arr = pa.array(x, type=PA_OF[ti], mask=None if valid is None else ~valid)
                    try:
                        want = pc.cast(arr, PA_OF[to], safe=True)
                        pa_ok = True
                    except pa.ArrowInvalid:
                        pa_ok = False
                    st, out, bad = _oracle_cast(cpu, ti, to, x, valid, False, False)
                    if idt.kind == "f" and odt.kind == "f":
                        assert st == 0   # float -> float never fails in either implementation
                        continue
                    assert (st == 0) == pa_ok, (TYPE_NAME[ti], TYPE_NAME[to], x.tolist(), None if valid is None else valid.tolist(), st, pa_ok)
                    if pa_ok:
                        ok = np.ones(3, dtype=bool) if valid is None else valid
                        wv = want.to_numpy(zero_copy_only=False)
                        got, exp = out[ok], np.asarray(wv[ok], dtype=NP_OF[to])
                        assert np.array_equal(got, exp, equal_nan=(odt.kind == "f")), (TYPE_NAME[ti], TYPE_NAME[to], x.tolist())
                    else:
                        assert bad == 1
    # unsafe integer casts wrap identically
    for ti, to in ((N.INT32, N.UINT8), (N.INT64, N.INT16), (N.UINT32, N.INT8), (N.UINT64, N.INT32)):
        x = random_values(rng, ti, 1000)
        st, out, _ = _oracle_cast(cpu, ti, to, x, None, True, True)
        want = pc.cast(pa.array(x), PA_OF[to], safe=False).to_numpy()
        assert st == 0 and np.array_equal(out, want)


def test_cumulative_sum_vs_pyarrow(cpu):
    """pyarrow's cumulative_sum[_checked] (start, skip_nulls) against the restatement of vector_cumulative.go."""
    rng = np.random.default_rng(77)

    class St(C.Structure):
        _fields_ = [("cur", C.c_uint8 * 8), ("enc", C.c_int64)]

    for t in (N.INT8, N.UINT16, N.INT32, N.INT64, N.UINT64, N.FLOAT64):
        dt = np.dtype(NP_OF[t])
        for n in (0, 1, 17, 500):
            for skip in (False, True):
                for checked in (False, True):
                    x = rng.integers(-3, 4, n).astype(dt) if dt.kind != "u" else rng.integers(0, 4, n).astype(dt)
                    valid = rng.random(n) > 0.2
                    start = 2
                    st_ = St()
                    for i, b in enumerate(np.array([start], dtype=dt).tobytes()):
                        st_.cur[i] = b
                    out = np.zeros(n, dtype=dt)
                    ov = np.zeros(n // 8 + 2, dtype=np.uint8)
                    nulls, bad = C.c_int64(), C.c_int64()
                    bm = pack_bits(valid, offset=0)
                    rc = cpu.ref_cumulative_sum(t, ptr(x), ptr(bm), 0, n, int(skip), int(checked), ptr(out), ptr(ov), 0, C.byref(st_), C.byref(nulls), C.byref(bad))
                    arr = pa.array(x, type=PA_OF[t], mask=~valid)
                    fn = pc.cumulative_sum_checked if checked else pc.cumulative_sum
                    want = fn(arr, start=pa.scalar(start, type=PA_OF[t]), skip_nulls=skip)
                    assert rc == 0
                    wv = np.array([v is not None for v in want.to_pylist()], dtype=bool)
                    assert np.array_equal(unpack_bits(ov, 0, n).astype(bool), wv), (TYPE_NAME[t], n, skip)
                    wvals = np.array([0 if v is None else v for v in want.to_pylist()], dtype=dt)
                    assert np.array_equal(out[wv], wvals[wv]) and nulls.value == int((~wv).sum())
    # overflow: checked fails in both, unchecked wraps in both
    x = np.array([127, 1], dtype=np.int8)
    with pytest.raises(pa.ArrowInvalid):
        pc.cumulative_sum_checked(pa.array(x))
    st_ = St()
    out = np.zeros(2, dtype=np.int8)
    bad = C.c_int64()
    assert cpu.ref_cumulative_sum(N.INT8, ptr(x), None, 0, 2, 0, 1, ptr(out), None, 0, C.byref(st_), None, C.byref(bad)) != 0 and bad.value == 1
    assert pc.cumulative_sum(pa.array(x)).to_pylist() == [127, -128]
