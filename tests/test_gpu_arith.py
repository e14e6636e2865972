"""Arithmetic kernels on the GPU vs the reference's own SIMD loops (oracle/_ref) and the
restatement.  Test shapes follow arrow/compute/arithmetic_test.go: BinaryArithmeticSuite over
all 10 numeric types (:699-713), arr⊕arr / arr⊕scalar / scalar⊕arr (:115-132), overflow cases
(:352-355,388-390)."""
import ctypes as C

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import ALL_TYPES, INT_TYPES, NP_OF, TYPE_NAME, Dev, misaligned, pack_bits, ptr, random_values, same_bits, same_float_class

pytestmark = pytest.mark.gpu

BIN_OPS = [N.OP_ADD, N.OP_SUB, N.OP_MUL, N.OP_ADD_CHECKED, N.OP_SUB_CHECKED, N.OP_MUL_CHECKED]
SIZES = [0, 1, 7, 33, 255, 1000, 4099, 1 << 16, (1 << 20) + 3]


def ref_binary(ref, isa, type_id, op, shape, l, r, n):
    out = np.empty(n, dtype=NP_OF[type_id])
    fn = getattr(ref, {N.SHAPE_AA: "arithmetic_binary_", N.SHAPE_AS: "arithmetic_arr_scalar_", N.SHAPE_SA: "arithmetic_scalar_arr_"}[shape] + isa)
    fn(type_id, op, ptr(l), ptr(r), ptr(out), n)
    return out


@pytest.mark.parametrize("type_id", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
@pytest.mark.parametrize("shape", [N.SHAPE_AA, N.SHAPE_AS, N.SHAPE_SA], ids=["aa", "as", "sa"])
def test_binary_matches_reference_simd(ag, ref, cpu, isa, type_id, shape):
    rng = np.random.default_rng(0x0FF1CE + type_id * 7 + shape)
    isf = type_id in (N.FLOAT32, N.FLOAT64)
    for n in SIZES:
        for op in BIN_OPS:
            l = random_values(rng, type_id, 1 if shape == N.SHAPE_SA else n)
            r = random_values(rng, type_id, 1 if shape == N.SHAPE_AS else n)
            want = ref_binary(ref, isa, type_id, op, shape, l, r, n)
            # the restatement agrees with the reference's instruction stream
            mine = np.empty(n, dtype=NP_OF[type_id])
            assert cpu.ref_arith_binary(type_id, op, shape, ptr(l), ptr(r), ptr(mine), n) == 0
            assert (same_float_class if isf else same_bits)(mine, want)
            # host-pointer flavour
            got = np.empty(n, dtype=NP_OF[type_id])
            fn = {N.SHAPE_AA: "ag_arith_binary", N.SHAPE_AS: "ag_arith_arr_scalar", N.SHAPE_SA: "ag_arith_scalar_arr"}[shape]
            ag.call(fn, type_id, op, ptr(l), ptr(r), ptr(got), n)
            assert (same_float_class if isf else same_bits)(got, want), (TYPE_NAME[type_id], op, shape, n)
            # device flavour, aligned and element-misaligned operands (Arrow slices)
            if n:
                for mis in (0, 1):
                    isz = np.dtype(NP_OF[type_id]).itemsize
                    dl = l if shape == N.SHAPE_SA else Dev(l, byte_offset=mis * isz)
                    dr = r if shape == N.SHAPE_AS else Dev(r, byte_offset=mis * isz)
                    do = Dev(np.zeros(n, dtype=NP_OF[type_id]), byte_offset=mis * isz)
                    ag.call("ag_arith_binary_dev", type_id, op, shape,
                            ptr(dl) if shape == N.SHAPE_SA else dl.ptr, ptr(dr) if shape == N.SHAPE_AS else dr.ptr, do.ptr, n, None)
                    ag.call("ag_stream_sync", None)
                    assert (same_float_class if isf else same_bits)(do.get(), want), (TYPE_NAME[type_id], op, shape, n, mis)


@pytest.mark.parametrize("type_id", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_unary_matches_reference_simd(ag, ref, cpu, isa, type_id):
    rng = np.random.default_rng(0xABC + type_id)
    isf = type_id in (N.FLOAT32, N.FLOAT64)
    for n in SIZES:
        x = random_values(rng, type_id, n)
        if n > 4 and not isf:
            info = np.iinfo(NP_OF[type_id])
            x[:3] = [info.min, info.max, 0]
        for op in (N.OP_ABS, N.OP_ABS_CHECKED, N.OP_NEGATE, N.OP_NEGATE_CHECKED, N.OP_SIGN):
            want = np.empty(n, dtype=NP_OF[type_id])
            getattr(ref, "arithmetic_unary_same_types_" + isa)(type_id, op, ptr(x), ptr(want), n)
            mine = np.empty(n, dtype=NP_OF[type_id])
            assert cpu.ref_arith_unary_same(type_id, op, ptr(x), ptr(mine), n) == 0
            assert (same_float_class if isf else same_bits)(mine, want), ("oracle", TYPE_NAME[type_id], op, n)
            got = np.empty(n, dtype=NP_OF[type_id])
            ag.call("ag_arith_unary_same", type_id, op, ptr(x), ptr(got), n)
            assert (same_float_class if isf else same_bits)(got, want), (TYPE_NAME[type_id], op, n)
            if n:
                dx = Dev(x, byte_offset=np.dtype(NP_OF[type_id]).itemsize)
                do = Dev(np.zeros(n, dtype=NP_OF[type_id]))
                ag.call("ag_arith_unary_same_dev", type_id, op, dx.ptr, do.ptr, n, None)
                ag.call("ag_stream_sync", None)
                assert (same_float_class if isf else same_bits)(do.get(), want)


@pytest.mark.parametrize("itype", INT_TYPES, ids=lambda t: TYPE_NAME[t])
def test_sign_diff_type_matches_reference_simd(ag, ref, isa, itype):
    # the Go side asks for Sign int -> int of another width (base_arithmetic_amd64.go:80-86)
    rng = np.random.default_rng(itype)
    n = 4099
    x = random_values(rng, itype, n, small=True)
    for otype in (N.INT8, N.INT32, N.INT64, N.UINT8):
        want = np.empty(n, dtype=NP_OF[otype])
        getattr(ref, "arithmetic_unary_diff_type_" + isa)(itype, otype, N.OP_SIGN, ptr(x), ptr(want), n)
        got = np.empty(n, dtype=NP_OF[otype])
        ag.call("ag_arith_unary_diff", itype, otype, N.OP_SIGN, ptr(x), ptr(got), n)
        assert same_bits(got, want), (TYPE_NAME[itype], TYPE_NAME[otype])


def test_float_special_values(ag, ref, isa):
    # arithmetic_test.go compares with NaNsEqual (:210-213); everything else must be bit exact
    a = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e308, -1e308, 5e-324, 1.0, 2.5], dtype=np.float64)
    b = np.array([-0.0, -0.0, -np.inf, -np.inf, 1.0, 1e308, 1e308, 5e-324, np.nan, -2.5], dtype=np.float64)
    for op in (N.OP_ADD, N.OP_SUB, N.OP_MUL):
        want = ref_binary(ref, isa, N.FLOAT64, op, N.SHAPE_AA, a, b, a.size)
        got = np.empty_like(a)
        ag.call("ag_arith_binary", N.FLOAT64, op, ptr(a), ptr(b), ptr(got), a.size)
        assert same_float_class(got, want)


@pytest.mark.parametrize("type_id", INT_TYPES, ids=lambda t: TYPE_NAME[t])
@pytest.mark.parametrize("op", [N.OP_ADD_CHECKED, N.OP_SUB_CHECKED, N.OP_MUL_CHECKED, N.OP_DIV, N.OP_DIV_CHECKED])
def test_checked_integer_ops_match_oracle(ag, cpu, type_id, op):
    rng = np.random.default_rng(type_id * 31 + op)
    dt = NP_OF[type_id]
    for shape in (N.SHAPE_AA, N.SHAPE_AS, N.SHAPE_SA):
        for n in (1, 65, 1000, 70001):
            for small in (True, False):  # small operands: no overflow; full range: overflow almost surely
                for nullp in (0.0, 0.3):
                    l = random_values(rng, type_id, 1 if shape == N.SHAPE_SA else n, small)
                    r = random_values(rng, type_id, 1 if shape == N.SHAPE_AS else n, small)
                    if op in (N.OP_DIV, N.OP_DIV_CHECKED) and small and shape != N.SHAPE_SA:
                        r[r == 0] = 1  # make the no-error case reachable
                    loff, roff = 3, 5
                    lv = pack_bits(rng.random(n) >= nullp, loff) if (nullp and shape != N.SHAPE_SA) else None
                    rv = pack_bits(rng.random(n) >= nullp, roff) if (nullp and shape != N.SHAPE_AS) else None
                    want = np.full(n, 7, dtype=dt)
                    wbad = C.c_int64()
                    wst = cpu.ref_arith_checked(type_id, op, shape, ptr(l), ptr(lv), loff, ptr(r), ptr(rv), roff, ptr(want), n, C.byref(wbad))
                    got = np.full(n, 9, dtype=dt)
                    gbad = C.c_int64()
                    gst, msg = ag.call_status("ag_arith_checked", type_id, op, shape, ptr(l), ptr(lv), loff, ptr(r), ptr(rv), roff, ptr(got), n, C.byref(gbad))
                    assert gst == wst, (TYPE_NAME[type_id], op, shape, n, small, nullp, msg)
                    assert gbad.value == wbad.value
                    if wst == 0:
                        assert same_bits(got, want)
                    else:
                        assert msg in ("overflow", "divide by zero")


def test_checked_add_reference_cases(ag):
    # arithmetic_test.go:352-355: max + max overflows when checked; wraps when unchecked
    for type_id in INT_TYPES:
        info = np.iinfo(NP_OF[type_id])
        a = np.array([info.max], dtype=NP_OF[type_id])
        out = np.zeros(1, dtype=NP_OF[type_id])
        bad = C.c_int64()
        st, msg = ag.call_status("ag_arith_checked", type_id, N.OP_ADD_CHECKED, N.SHAPE_AA, ptr(a), None, 0, ptr(a), None, 0, ptr(out), 1, C.byref(bad))
        assert st == N.AG_ERR_INVALID and msg == "overflow" and bad.value == 0
        ag.call("ag_arith_binary", type_id, N.OP_ADD, ptr(a), ptr(a), ptr(out), 1)
        with np.errstate(over="ignore"):
            assert out[0] == (a + a)[0]
    # null slots are not computed: max + max under a null is fine and yields 0 (helpers.go:296-306)
    a = np.array([np.iinfo(np.int32).max, 1], dtype=np.int32)
    lv = np.array([0b10], dtype=np.uint8)
    out = np.full(2, -1, dtype=np.int32)
    ag.call("ag_arith_checked", N.INT32, N.OP_ADD_CHECKED, N.SHAPE_AA, ptr(a), ptr(lv), 0, ptr(a), None, 0, ptr(out), 2, None)
    assert out.tolist() == [0, 2]


def test_add_f64_100m_rows_properties(ag):
    """BASELINE config 2 size: 100M-row float64 add, checked through (a) a 64-bit order-sensitive
    checksum against the oracle on a sampled window, (b) commutativity a+b == b+a bit for bit,
    (c) (a+b)-b == a exactly on integer-valued data (every step exact)."""
    n = 100_000_000
    a = Dev(nbytes=n * 8)
    b = Dev(nbytes=n * 8)
    o1 = Dev(nbytes=n * 8)
    o2 = Dev(nbytes=n * 8)
    ag.call("ag_generate_dev", 3, 0x94378165, -(1 << 20), 1 << 20, a.ptr, n, None)
    ag.call("ag_generate_dev", 3, 0x94378166, -(1 << 20), 1 << 20, b.ptr, n, None)
    ag.call("ag_arith_binary_dev", N.FLOAT64, N.OP_ADD, N.SHAPE_AA, a.ptr, b.ptr, o1.ptr, n, None)
    ag.call("ag_arith_binary_dev", N.FLOAT64, N.OP_ADD, N.SHAPE_AA, b.ptr, a.ptr, o2.ptr, n, None)
    cs = Dev(np.zeros(4, dtype=np.uint64))
    ag.call("ag_checksum64_dev", o1.ptr, n, cs.ptr, None)
    ag.call("ag_checksum64_dev", o2.ptr, n, cs.ptr + 8, None)
    ag.call("ag_arith_binary_dev", N.FLOAT64, N.OP_SUB, N.SHAPE_AA, o1.ptr, b.ptr, o2.ptr, n, None)
    ag.call("ag_checksum64_dev", o2.ptr, n, cs.ptr + 16, None)
    ag.call("ag_checksum64_dev", a.ptr, n, cs.ptr + 24, None)
    ag.call("ag_stream_sync", None)
    c = cs.get()
    assert c[0] == c[1]
    assert c[2] == c[3]
    # window vs oracle
    from oracle import oracle
    w = 1 << 20
    start = 77_777_776
    ha = a.buf.to_numpy(np.float64, w, start * 8)
    hb = b.buf.to_numpy(np.float64, w, start * 8)
    want = np.empty(w)
    assert oracle.cpu().ref_arith_binary(N.FLOAT64, N.OP_ADD, N.SHAPE_AA, ptr(ha), ptr(hb), ptr(want), w) == 0
    assert o1.buf.to_numpy(np.float64, w, start * 8).tobytes() == want.tobytes()


@pytest.mark.parametrize("type_id", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_unary_checked_matches_oracle(ag, cpu, type_id):
    """abs / negate (checked): MinInt anywhere -> "overflow" (base_arithmetic.go:295-340)."""
    rng = np.random.default_rng(type_id + 900)
    dt = NP_OF[type_id]
    isf = type_id in (N.FLOAT32, N.FLOAT64)
    for n in (1, 100, 4099, 70_001):
        for plant in (False, True):
            x = random_values(rng, type_id, n, small=True)
            if plant and not isf:
                x[n // 2] = np.iinfo(dt).min
                x[n - 1] = np.iinfo(dt).min
            for op in (N.OP_ABS_CHECKED, N.OP_NEGATE_CHECKED):
                want = np.empty(n, dtype=dt); wb = C.c_int64()
                wst = cpu.ref_arith_unary_checked(type_id, op, ptr(x), ptr(want), n, C.byref(wb))
                got = np.empty(n, dtype=dt); gb = C.c_int64()
                gst, msg = ag.call_status("ag_arith_unary_checked", type_id, op, ptr(x), ptr(got), n, C.byref(gb))
                assert gst == wst and gb.value == wb.value, (TYPE_NAME[type_id], op, n, plant, msg)
                if wst == 0:
                    assert (same_float_class if isf else same_bits)(got, want)
                else:
                    assert msg == "overflow" and gb.value == n // 2


# ---------------------------------------------------------------------------------------------------------
# The batched entry point the headline benchmark times: ag_arith_binary_spans[_dev] (one launch for every aligned span
# of a chunked call, executor.go:598-623,757-863), checked directly against the reference's SIMD loop run span by span.
def _spans_case(ag, ref, isa, type_id, op, shape, lens, offs, rng, host=False):
    """lens: rows per span; offs: per span (l, r, out) element offsets into oversized buffers -> every operand of every
    span gets its own 16-byte phase."""
    dt = NP_OF[type_id]
    isz = np.dtype(dt).itemsize
    isf = type_id in (N.FLOAT32, N.FLOAT64)
    total = int(sum(lens)) + 64 * len(lens) + 64
    L = random_values(rng, type_id, total)
    R = random_values(rng, type_id, total)
    scal = random_values(rng, type_id, 1)
    want = np.zeros(total, dtype=dt)
    got_h = np.zeros(total, dtype=dt)
    spans, pos = [], 0
    for ln, (lo, ro, oo) in zip(lens, offs):
        spans.append((pos + lo, pos + ro, pos + oo, ln))
        pos += ln + 64
    fn = getattr(ref, {N.SHAPE_AA: "arithmetic_binary_", N.SHAPE_AS: "arithmetic_arr_scalar_", N.SHAPE_SA: "arithmetic_scalar_arr_"}[shape] + isa)
    for lp, rp, op_, ln in spans:
        if ln:
            fn(type_id, op, scal.ctypes.data if shape == N.SHAPE_SA else L.ctypes.data + lp * isz,
               scal.ctypes.data if shape == N.SHAPE_AS else R.ctypes.data + rp * isz, want.ctypes.data + op_ * isz, ln)
    cmp_ = same_float_class if isf else same_bits
    if host:
        table = N.span_table([(scal.ctypes.data if shape == N.SHAPE_SA else L.ctypes.data + lp * isz,
                               scal.ctypes.data if shape == N.SHAPE_AS else R.ctypes.data + rp * isz, got_h.ctypes.data + op_ * isz, ln)
                              for lp, rp, op_, ln in spans])
        ag.call("ag_arith_binary_spans", type_id, op, shape, table, len(spans))
        assert cmp_(got_h, want), (TYPE_NAME[type_id], op, shape, "host")
        return
    dL, dR, dO = Dev(L), Dev(R), Dev(np.zeros(total, dtype=dt))
    table = N.span_table([(scal.ctypes.data if shape == N.SHAPE_SA else dL.ptr + lp * isz,
                           scal.ctypes.data if shape == N.SHAPE_AS else dR.ptr + rp * isz, dO.ptr + op_ * isz, ln) for lp, rp, op_, ln in spans])
    ag.call("ag_arith_binary_spans_dev", type_id, op, shape, table, len(spans), None)
    ag.call("ag_stream_sync", None)
    assert cmp_(dO.get(), want), (TYPE_NAME[type_id], op, shape, lens[:4], offs[:4])  # rows between spans stay 0: no stray write


@pytest.mark.parametrize("type_id", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
@pytest.mark.parametrize("shape", [N.SHAPE_AA, N.SHAPE_AS, N.SHAPE_SA], ids=["aa", "as", "sa"])
def test_spans_entry_point_vs_reference_simd(ag, ref, isa, type_id, shape):
    rng = np.random.default_rng(0x5BA2 + type_id * 13 + shape)
    isz = np.dtype(NP_OF[type_id]).itemsize
    nvec = 16 // isz
    # spans far longer than a tile (32 KB per operand), exactly a tile, one row short / long of a tile, tiny, empty
    tile = 32768 // isz
    lens = [5 * tile + 17, tile, tile - 1, tile + 1, 1, 0, 3, 2 * tile + nvec - 1, 70_001]
    for trial in range(3):
        if trial == 0:     # every operand on a 16-byte boundary
            offs = [(0, 0, 0)] * len(lens)
        elif trial == 1:   # one shared misalignment per span (what executeSpans produces: out, l, r all at pos*width)
            offs = [(k % nvec,) * 3 for k in range(1, len(lens) + 1)]
        else:              # independent phases per operand
            offs = [tuple(int(x) for x in rng.integers(0, 2 * nvec, 3)) for _ in lens]
        for op in (N.OP_ADD, N.OP_SUB_CHECKED, N.OP_MUL):
            _spans_case(ag, ref, isa, type_id, op, shape, lens, offs, rng)
    _spans_case(ag, ref, isa, type_id, N.OP_ADD, shape, lens, [tuple(int(x) for x in rng.integers(0, 2 * nvec, 3)) for _ in lens], rng, host=True)


def test_spans_config2_layout_100m_rows_checksum(ag, cpu):
    """BASELINE config 2 at full size: 100M float64 rows, left chunks of 1M rows and right chunks of 999,983 rows -> 200
    spans into one contiguous output (the layout bench.py times).  Dataset E (integers stored as double: every sum is
    exact) generated on the device; the output is checked by the order-sensitive checksum of SURVEY §8d against the
    oracle's, and three 1M-row windows are compared bit for bit."""
    n, lc, rc = 100_000_000, 1_000_000, 999_983
    dl, dr, do = Dev(nbytes=n * 8), Dev(nbytes=n * 8), Dev(nbytes=n * 8)
    ag.call("ag_generate_dev", 3, 0x94378165, -(1 << 20), 1 << 20, dl.ptr, n, None)
    ag.call("ag_generate_dev", 3, 0x94378166, -(1 << 20), 1 << 20, dr.ptr, n, None)
    spans, pos = [], 0
    while pos < n:
        ln = min(lc - pos % lc, rc - pos % rc, n - pos)
        spans.append((dl.ptr + 8 * pos, dr.ptr + 8 * pos, do.ptr + 8 * pos, ln))
        pos += ln
    assert len(spans) == 200
    ag.call("ag_arith_binary_spans_dev", N.FLOAT64, N.OP_ADD_CHECKED, N.SHAPE_AA, N.span_table(spans), len(spans), None)
    ck = Dev(np.zeros(1, dtype=np.uint64))
    ag.call("ag_checksum64_dev", do.ptr, n, ck.ptr, None)
    ag.call("ag_stream_sync", None)
    # oracle: regenerate both columns on the CPU (the generator's twin), add with the restatement, checksum
    a, b, want = np.empty(n), np.empty(n), np.empty(n)
    cpu.ref_generate(3, 0x94378165, -(1 << 20), 1 << 20, a.ctypes.data, n)
    cpu.ref_generate(3, 0x94378166, -(1 << 20), 1 << 20, b.ctypes.data, n)
    assert cpu.ref_arith_binary(N.FLOAT64, N.OP_ADD_CHECKED, 0, a.ctypes.data, b.ctypes.data, want.ctypes.data, n) == 0
    assert int(ck.get()[0]) == cpu.ref_checksum64(want.ctypes.data, n)
    for start in (0, 49_999_000, n - 1_000_000):
        assert do.buf.to_numpy(np.float64, 1_000_000, start * 8).tobytes() == want[start:start + 1_000_000].tobytes()
