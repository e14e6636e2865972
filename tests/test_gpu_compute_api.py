"""The host-side mirror of the reference's compute API (CallFunction / Add / Filter / Take /
math.Sum) running on the GPU.  These read like the reference's own Go tests and use its literal
cases (tests/golden/*.json restated from arrow/compute/{arithmetic,vector_selection}_test.go and
arrow/math/float64_test.go)."""
import json
import os

import numpy as np
import pytest

from arrow_go_b200 import compute as pc

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NUMERIC = [pc.UINT8, pc.INT8, pc.UINT16, pc.INT16, pc.UINT32, pc.INT32, pc.UINT64, pc.INT64, pc.FLOAT32, pc.FLOAT64]
INTS = NUMERIC[:8]


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def datum(x, t):
    return pc.Array.from_pylist(x, t) if isinstance(x, list) else pc.Scalar(x, t)


@pytest.fixture(scope="module", autouse=True)
def _init(ag):
    pc.lib()


# ---------------------------------------------------------------- arithmetic -------------
@pytest.mark.parametrize("t", NUMERIC)
@pytest.mark.parametrize("no_check", [False, True])
def test_binary_arithmetic_suite(t, no_check):
    """BinaryArithmeticSuite[T].TestAdd/TestSub/TestMultiply (arithmetic_test.go:325-425) with both
    settings of ArithmeticOptions.NoCheckOverflow."""
    fns = {"add": pc.Add, "subtract": pc.Subtract, "multiply": pc.Multiply}
    for case in load("arithmetic.json")["cases"]:
        out = fns[case["op"]](datum(case["left"], t), datum(case["right"], t), no_check_overflow=no_check)
        assert out.type == t
        assert out.to_pylist() == case["expected"], (case, no_check)


@pytest.mark.parametrize("t", INTS)
def test_overflow_errors(t):
    # arithmetic_test.go:352-355, 388-390: checked [max]+[max] and [min]-[max] -> ErrInvalid "overflow"
    info = np.iinfo(pc.NP_OF[t])
    mx, mn = pc.Array.from_pylist([int(info.max)], t), pc.Array.from_pylist([int(info.min)], t)
    with pytest.raises(pc.ArrowError) as e:
        pc.Add(mx, mx)
    assert e.value.sentinel == "ErrInvalid" and "overflow" in e.value.msg
    with pytest.raises(pc.ArrowError) as e:
        pc.Subtract(mn, mx)
    assert e.value.sentinel == "ErrInvalid" and "overflow" in e.value.msg
    # unchecked wraps
    w = pc.Add(mx, mx, no_check_overflow=True).to_pylist()[0]
    assert w == (2 * int(info.max) - int(info.min)) % (int(info.max) - int(info.min) + 1) + int(info.min)
    # a null slot hides the overflow (ScalarBinaryNotNull visits valid slots only)
    a = pc.Array.from_pylist([None, 1], t)
    b = pc.Array.from_pylist([int(info.max), int(info.max) - 1], t)
    assert pc.Add(a, b).to_pylist() == [None, int(info.max)]


def test_add_chunked_and_sliced():
    """Chunked x chunked with misaligned chunk boundaries exercises iterateExecSpans +
    contiguous preallocation + per-span null propagation at non-zero output offsets
    (executor.go:598-623, 237-349)."""
    rng = np.random.default_rng(1)
    n = 10_000
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    av, bv = rng.random(n) > 0.1, rng.random(n) > 0.2
    cuts_a, cuts_b = [0, 1000, 1001, 4099, 7777, n], [0, 37, 5000, 5003, n]
    ca = pc.Chunked([pc.Array.from_numpy(a[s:e], av[s:e]) for s, e in zip(cuts_a, cuts_a[1:])], pc.FLOAT64)
    cb = pc.Chunked([pc.Array.from_numpy(b[s:e], bv[s:e]) for s, e in zip(cuts_b, cuts_b[1:])], pc.FLOAT64)
    out = pc.Add(ca, cb)
    assert out.kind == pc.KIND_CHUNKED and len(out) == n
    vals, valid, nulls = out.to_numpy()
    assert np.array_equal(valid, av & bv) and nulls == int((~(av & bv)).sum())
    assert np.array_equal(vals[valid], (a + b)[valid])
    # sliced arrays: offsets that are not multiples of 8
    full_a = pc.Array.from_numpy(a, av)
    full_b = pc.Array.from_numpy(b, bv)
    out = pc.Subtract(full_a.slice(3, 5000), full_b.slice(13, 5000))
    vals, valid, _ = out.to_numpy()
    ev = av[3:5003] & bv[13:5013]
    assert np.array_equal(valid, ev)
    assert np.array_equal(vals[ev], (a[3:5003] - b[13:5013])[ev])
    # int32 checked add over chunks, with nulls
    ia, ib = rng.integers(-1000, 1000, n).astype(np.int32), rng.integers(-1000, 1000, n).astype(np.int32)
    cia = pc.Chunked([pc.Array.from_numpy(ia[s:e], av[s:e]) for s, e in zip(cuts_a, cuts_a[1:])], pc.INT32)
    out = pc.Add(cia, pc.Array.from_numpy(ib, bv))
    vals, valid, _ = out.to_numpy()
    assert np.array_equal(valid, av & bv)
    assert np.array_equal(vals[valid], (ia + ib)[valid])
    assert (vals[~valid] == 0).all()  # null slots of the NotNull kernels are zero


def test_length_mismatch_is_invalid():
    with pytest.raises(pc.ArrowError) as e:
        pc.Add(pc.Array.from_pylist([1, 2, 3], pc.INT32), pc.Array.from_pylist([1, 2], pc.INT32))
    assert e.value.sentinel == "ErrInvalid"


# ---------------------------------------------------------------- comparisons ------------
@pytest.mark.parametrize("t", NUMERIC)
def test_numeric_compare_suite(t):
    """NumericCompareSuite (scalar_compare_test.go:299-483): array ⊕ scalar, scalar ⊕ array,
    null scalar, array ⊕ array."""
    arr = [0, 0, 1, 1, 2, 2, None]
    a = pc.Array.from_pylist(arr, t)
    one = pc.Scalar(1, t)
    ops = {"equal": lambda x, y: x == y, "not_equal": lambda x, y: x != y, "greater": lambda x, y: x > y,
           "greater_equal": lambda x, y: x >= y, "less": lambda x, y: x < y, "less_equal": lambda x, y: x <= y}
    for name, f in ops.items():
        assert pc.CallFunction(name, [a, one]).to_pylist() == [None if x is None else f(x, 1) for x in arr], name
        assert pc.CallFunction(name, [one, a]).to_pylist() == [None if x is None else f(1, x) for x in arr], name
        assert pc.CallFunction(name, [a, pc.Scalar(None, t)]).to_pylist() == [None] * len(arr)
        other = [1, 0, 2, None, 2, 5, 5]
        b = pc.Array.from_pylist(other, t)
        assert pc.CallFunction(name, [a, b]).to_pylist() == [None if (x is None or y is None) else f(x, y) for x, y in zip(arr, other)]
        assert pc.CallFunction(name, [a, one]).type == pc.BOOL
    empty = pc.Array.from_pylist([], t)
    assert pc.CallFunction("equal", [empty, one]).to_pylist() == []


def test_compare_sliced_output_offsets():
    # chunked input => contiguous boolean output written at bit offsets that are not byte aligned
    rng = np.random.default_rng(2)
    n = 5001
    v = rng.integers(0, 100, n).astype(np.int64)
    cuts = [0, 3, 1000, 1003, 2048, n]
    c = pc.Chunked([pc.Array.from_numpy(v[s:e]) for s, e in zip(cuts, cuts[1:])], pc.INT64)
    out = pc.CallFunction("greater", [c, pc.Scalar(89, pc.INT64)])
    vals, valid, _ = out.to_numpy()
    assert valid.all() and np.array_equal(vals, v > 89)


# ---------------------------------------------------------------- boolean ----------------
def test_boolean_kernels():
    """scalar_bool_test.go TestBooleanKernels truth tables (null-intersecting and Kleene)."""
    T, F, U = True, False, None
    left = [T, T, T, F, F, F, U, U, U]
    right = [T, F, U, T, F, U, T, F, U]
    l, r = pc.Array.from_pylist(left, pc.BOOL), pc.Array.from_pylist(right, pc.BOOL)
    assert pc.CallFunction("and", [l, r]).to_pylist() == [T, F, U, F, F, U, U, U, U]
    assert pc.CallFunction("or", [l, r]).to_pylist() == [T, T, U, T, F, U, U, U, U]
    assert pc.CallFunction("xor", [l, r]).to_pylist() == [F, T, U, T, F, U, U, U, U]
    assert pc.CallFunction("and_not", [l, r]).to_pylist() == [F, T, U, F, F, U, U, U, U]
    assert pc.CallFunction("and_kleene", [l, r]).to_pylist() == [T, F, U, F, F, F, U, F, U]
    assert pc.CallFunction("or_kleene", [l, r]).to_pylist() == [T, T, T, T, F, U, T, U, U]
    assert pc.CallFunction("and_not_kleene", [l, r]).to_pylist() == [F, T, U, F, F, F, F, U, U]
    assert pc.CallFunction("not", [l]).to_pylist() == [F, F, F, T, T, T, U, U, U]
    # array ⊕ scalar equivalence helper (scalar_bool_test.go:40-58)
    for name in ("and", "or", "xor", "and_not"):
        for sv in (T, F):
            want = pc.CallFunction(name, [l, pc.Array.from_pylist([sv] * 9, pc.BOOL)]).to_pylist()
            assert pc.CallFunction(name, [l, pc.Scalar(sv, pc.BOOL)]).to_pylist() == want, (name, sv)
            want = pc.CallFunction(name, [pc.Array.from_pylist([sv] * 9, pc.BOOL), r]).to_pylist()
            assert pc.CallFunction(name, [pc.Scalar(sv, pc.BOOL), r]).to_pylist() == want, (name, sv)


# ---------------------------------------------------------------- filter / take ----------
@pytest.mark.parametrize("t", [pc.INT8, pc.UINT16, pc.INT32, pc.INT64, pc.FLOAT32, pc.FLOAT64])
def test_filter_numeric(t):
    for case in load("filter_numeric.json")["cases"]:
        for sel, key in ((pc.EMIT_NULLS, "emit_null"), (pc.DROP_NULLS, "drop")):
            v, m = pc.Array.from_pylist(case["values"], t), pc.Array.from_pylist(case["filter"], pc.BOOL)
            assert pc.Filter(v, m, sel).to_pylist() == case[key], (case, key)
            # sliced re-run: 3 filler values / 2 filler filter slots either side (vector_selection_test.go:114-145)
            vs = pc.Array.from_pylist([None] * 3 + case["values"] + [None] * 3, t).slice(3, len(case["values"]))
            ms = pc.Array.from_pylist([True, False] + case["filter"] + [True, False], pc.BOOL).slice(2, len(case["filter"]))
            assert pc.Filter(vs, ms, sel).to_pylist() == case[key], ("sliced", case, key)
    # length mismatch -> ErrInvalid (:480-483)
    with pytest.raises(pc.ArrowError) as e:
        pc.Filter(pc.Array.from_pylist([7, 8, 9], t), pc.Array.from_pylist([], pc.BOOL))
    assert e.value.sentinel == "ErrInvalid"


def test_filter_random_vs_compare():
    """vector_selection_test.go:554-613: filter(values, values <op> scalar) against a naive loop."""
    rng = np.random.default_rng(0x0FF1CE)
    for n in (8, 64, 100, 512, 10_000):
        v = rng.integers(0, 100, n).astype(np.int64)
        valid = rng.random(n) > 0.1
        arr = pc.Array.from_numpy(v, valid)
        for name, f in (("equal", v == 50), ("not_equal", v != 50), ("greater", v > 50), ("less_equal", v <= 50)):
            mask = pc.CallFunction(name, [arr, pc.Scalar(50, pc.INT64)])
            got = pc.Filter(arr, mask, pc.DROP_NULLS).to_pylist()
            assert got == [int(x) for x, ok, keep in zip(v, valid, f) if ok and keep]
            got = pc.Filter(arr, mask, pc.EMIT_NULLS).to_pylist()
            assert got == [(int(x) if ok else None) for x, ok, keep in zip(v, valid, f) if (not ok) or keep]


def test_filter_chunked():
    # TestFilterChunkedArray-style (:1006-1018): chunked values x chunked filter -> chunked result
    v = pc.Chunked([pc.Array.from_pylist([7, 8], pc.INT32), pc.Array.from_pylist([9, 10, 11], pc.INT32)], pc.INT32)
    m = pc.Chunked([pc.Array.from_pylist([True, False, True], pc.BOOL), pc.Array.from_pylist([None, True], pc.BOOL)], pc.BOOL)
    out = pc.Filter(v, m, pc.EMIT_NULLS)
    assert out.kind == pc.KIND_CHUNKED and out.to_pylist() == [7, 9, None, 11]
    assert pc.Filter(v, m, pc.DROP_NULLS).to_pylist() == [7, 9, 11]


@pytest.mark.parametrize("t", [pc.INT8, pc.UINT16, pc.INT32, pc.INT64, pc.FLOAT64])
def test_take_numeric(t):
    for case in load("take_numeric.json")["cases"]:
        for it in (pc.INT32, pc.INT8, pc.UINT32):
            if it == pc.UINT32 and any(x is not None and x < 0 for x in case["indices"]):
                continue
            v, i = pc.Array.from_pylist(case["values"], t), pc.Array.from_pylist(case["indices"], it)
            if "error_index" in case:
                with pytest.raises(pc.ArrowError) as e:
                    pc.Take(v, i)
                assert e.value.sentinel == "ErrIndex" and f"{case['error_index']} out of bounds" in e.value.msg
                continue
            assert pc.Take(v, i).to_pylist() == case["expected"], case
            # sliced values / sliced indices (checkTake, vector_selection_test.go:213-253)
            vs = pc.Array.from_pylist([None, None] + case["values"] + [None], t).slice(2, len(case["values"]))
            is_ = pc.Array.from_pylist([0] + case["indices"] + [0, 0], it).slice(1, len(case["indices"]))
            assert pc.Take(vs, is_).to_pylist() == case["expected"], ("sliced", case)


def test_take_chunked():
    # :1609-1642 chunked values / chunked indices
    v = pc.Chunked([pc.Array.from_pylist([7], pc.INT32), pc.Array.from_pylist([8, 9], pc.INT32)], pc.INT32)
    assert pc.Take(v, pc.Array.from_pylist([0, 1, 0, 2], pc.INT32)).to_pylist() == [7, 8, 7, 9]
    i = pc.Chunked([pc.Array.from_pylist([0, 1, 0], pc.INT32), pc.Array.from_pylist([], pc.INT32), pc.Array.from_pylist([2], pc.INT32)], pc.INT32)
    out = pc.Take(v, i)
    assert out.kind == pc.KIND_CHUNKED and out.to_pylist() == [7, 8, 7, 9]
    with pytest.raises(pc.ArrowError) as e:
        pc.Take(v, pc.Array.from_pylist([0, 5], pc.INT32))
    assert e.value.sentinel == "ErrIndex"


# ---------------------------------------------------------------- arrow/math --------------
def test_math_sum():
    # arrow/math/float64_test.go:30-48
    f = pc.Array.from_numpy(np.arange(10000, dtype=np.float64))
    assert pc.math.sum_float64(f) == 49995000.0
    assert pc.math.sum_float64(f, reference_order=True) == 49995000.0
    assert pc.math.sum_float64(pc.Array.from_numpy(np.zeros(0))) == 0.0
    assert pc.math.sum_int64(pc.Array.from_numpy(np.arange(10000, dtype=np.int64))) == 49995000
    assert pc.math.sum_uint64(pc.Array.from_numpy(np.arange(10000, dtype=np.uint64))) == 49995000
    # validity is ignored and slices work (Float64Values() = values[offset:offset+len])
    x = np.arange(100, dtype=np.float64)
    a = pc.Array.from_numpy(x, np.arange(100) % 3 != 0)
    assert pc.math.sum_float64(a.slice(7, 50)) == x[7:57].sum()


def test_abs_negate_checked():
    # "abs" / "negate" fail on MinInt (arithmetic.go:822-846); the _unchecked forms wrap
    for t in (pc.INT8, pc.INT16, pc.INT32, pc.INT64):
        mn = int(np.iinfo(pc.NP_OF[t]).min)
        ok = pc.Array.from_pylist([3, -4, None, 0], t)
        assert pc.CallFunction("abs", [ok]).to_pylist() == [3, 4, None, 0]
        assert pc.CallFunction("negate", [ok]).to_pylist() == [-3, 4, None, 0]
        bad = pc.Array.from_pylist([1, mn], t)
        for fn in ("abs", "negate"):
            with pytest.raises(pc.ArrowError) as e:
                pc.CallFunction(fn, [bad])
            assert e.value.sentinel == "ErrInvalid" and "overflow" in e.value.msg
        assert pc.CallFunction("abs_unchecked", [bad]).to_pylist() == [1, mn]
        assert pc.CallFunction("negate_unchecked", [bad]).to_pylist() == [-1, mn]
    assert pc.CallFunction("abs", [pc.Array.from_pylist([1.5, -2.5, None], pc.FLOAT64)]).to_pylist() == [1.5, 2.5, None]
    assert pc.CallFunction("abs", [pc.Array.from_pylist([7, 0], pc.UINT16)]).to_pylist() == [7, 0]
    assert pc.CallFunction("sign", [pc.Array.from_pylist([-5, 0, 9, None], pc.INT32)]).to_pylist() == [-1, 0, 1, None]
    with pytest.raises(pc.ArrowError):
        pc.dispatch("negate", [pc.UINT32])  # negate (checked) has signed / floating kernels only


def test_kleene_scalar_operands():
    """Kleene kernels with a scalar on either side must equal the array ⊕ array result with the
    scalar broadcast (scalar_bool_test.go:40-58 helper), for true / false / null scalars and arrays
    with and without nulls."""
    T, F, U = True, False, None
    for arr in ([T, F, U, T, F, U, T, T, F], [T, F, T, T, F, F, T, T, F]):
        a = pc.Array.from_pylist(arr, pc.BOOL)
        for name in ("and_kleene", "or_kleene", "and_not_kleene"):
            for sv in (T, F, U):
                bcast = pc.Array.from_pylist([sv] * len(arr), pc.BOOL)
                assert pc.CallFunction(name, [a, pc.Scalar(sv, pc.BOOL)]).to_pylist() == pc.CallFunction(name, [a, bcast]).to_pylist(), (name, sv, "right")
                assert pc.CallFunction(name, [pc.Scalar(sv, pc.BOOL), a]).to_pylist() == pc.CallFunction(name, [bcast, a]).to_pylist(), (name, sv, "left")


def test_boolean_values_filter_take_api():
    # TestFilterBoolean / TestTakeBoolean-style literals (vector_selection_test.go)
    T, F, U = True, False, None
    v = pc.Array.from_pylist([T, F, U, T, F], pc.BOOL)
    m = pc.Array.from_pylist([T, T, T, U, F], pc.BOOL)
    assert pc.Filter(v, m, pc.DROP_NULLS).to_pylist() == [T, F, U]
    assert pc.Filter(v, m, pc.EMIT_NULLS).to_pylist() == [T, F, U, U]
    assert pc.Take(v, pc.Array.from_pylist([4, 0, 2, None, 1], pc.INT32)).to_pylist() == [F, T, U, U, F]
    assert pc.Take(pc.Array.from_pylist([T, F, T], pc.BOOL), pc.Array.from_pylist([0, 1, 0], pc.INT8)).to_pylist() == [T, F, T]
    with pytest.raises(pc.ArrowError) as e:
        pc.Take(v, pc.Array.from_pylist([0, 5], pc.INT32))
    assert e.value.sentinel == "ErrIndex"


def test_is_null_is_not_null_is_nan():
    # scalar_compare_test.go TestIsNull / TestIsNaN shapes (kernels scalar_comparisons.go:718-813)
    a = pc.Array.from_pylist([1.5, None, float("nan"), 0.0, None], pc.FLOAT64)
    assert pc.CallFunction("is_null", [a]).to_pylist() == [False, True, False, False, True]
    assert pc.CallFunction("is_not_null", [a]).to_pylist() == [True, False, True, True, False]
    no_nulls = pc.Array.from_pylist([1, 2, 3], pc.INT32)
    assert pc.CallFunction("is_null", [no_nulls]).to_pylist() == [False, False, False]
    assert pc.CallFunction("is_not_null", [no_nulls]).to_pylist() == [True, True, True]
    assert pc.CallFunction("is_nan", [no_nulls]).to_pylist() == [False, False, False]
    nan = pc.CallFunction("is_nan", [pc.Array.from_pylist([1.5, float("nan"), float("inf"), -0.0], pc.FLOAT64)])
    assert nan.to_pylist() == [False, True, False, False] and nan.type == pc.BOOL
    s = pc.Array.from_pylist([None, 1, None, 4, 5, None, 7, 8, 9], pc.INT64).slice(1, 7)
    assert pc.CallFunction("is_null", [s]).to_pylist() == [False, True, False, False, True, False, False]


# ---------------------------------------------------------------- cast + implicit promotion ----
def test_cast_reference_vectors():
    """arrow/compute/cast_test.go:483-629 through compute.CastDatum (tests/golden/cast_numeric.json)."""
    name_to_id = {"int8": pc.INT8, "int16": pc.INT16, "int32": pc.INT32, "int64": pc.INT64, "uint8": pc.UINT8, "uint16": pc.UINT16,
                  "uint32": pc.UINT32, "uint64": pc.UINT64, "float32": pc.FLOAT32, "float64": pc.FLOAT64}
    for case in load("cast_numeric.json")["cases"]:
        ti, to = name_to_id[case["from"]], name_to_id[case["to"]]
        vals = list(case["in"])
        for i in case.get("null_at", []):
            vals[i] = None
        arr = pc.Array.from_pylist(vals, ti)
        if "slice" in case:
            arr = arr.slice(case["slice"][0], case["slice"][1] - case["slice"][0])
        kw = dict(allow_int_overflow=bool(case.get("allow_int_overflow")), allow_float_truncate=bool(case.get("allow_float_truncate")))
        if case.get("fails"):
            with pytest.raises(pc.ArrowError) as e:
                pc.Cast(arr, to, **kw)
            assert e.value.sentinel == "ErrInvalid", case
            assert ("not in range" in e.value.msg) or ("was truncated" in e.value.msg), e.value.msg
        else:
            out = pc.Cast(arr, to, **kw)
            assert out.type == to and out.to_pylist() == case["out"], case
    # error wording (helpers.go:591-594, numeric_cast.go:614-617)
    with pytest.raises(pc.ArrowError) as e:
        pc.Cast(pc.Array.from_pylist([0, None, 2000, 70000, 2], pc.INT32), pc.INT16)
    assert "integer value 70000 not in range: -32768 to 32767" in e.value.msg
    with pytest.raises(pc.ArrowError) as e:
        pc.Cast(pc.Array.from_pylist([1.0, 2.5], pc.FLOAT64), pc.INT32)
    assert "float value 2.500000 was truncated converting to int32" in e.value.msg


def test_cast_same_type_chunked_scalar():
    a = pc.Array.from_pylist([1, None, 3], pc.INT32)
    assert pc.Cast(a, pc.INT32).to_pylist() == [1, None, 3]                      # castMetaFunc: same type returns the input
    rng = np.random.default_rng(5)
    x = rng.integers(-1000, 1000, 5000).astype(np.int16)
    v = rng.random(5000) > 0.2
    cuts = [0, 1, 999, 1000, 4099, 5000]
    c = pc.Chunked([pc.Array.from_numpy(x[s:e], v[s:e]) for s, e in zip(cuts, cuts[1:])], pc.INT16)
    out = pc.Cast(c, pc.FLOAT64)
    assert out.kind == pc.KIND_CHUNKED and out.type == pc.FLOAT64
    vals, valid, _ = out.to_numpy()
    assert np.array_equal(valid, v) and np.array_equal(vals[v], x.astype(np.float64)[v])
    # sliced input (offset not a multiple of 8) keeps its validity
    full = pc.Array.from_numpy(x, v)
    vals, valid, _ = pc.Cast(full.slice(13, 777), pc.INT64).to_numpy()
    assert np.array_equal(valid, v[13:790]) and np.array_equal(vals[valid], x[13:790].astype(np.int64)[valid])
    # scalars
    assert pc.scalar_value(pc.Cast(pc.Scalar(7, pc.INT8), pc.FLOAT32)) == 7.0
    assert pc.scalar_value(pc.Cast(pc.Scalar(None, pc.INT8), pc.INT64)) is None
    with pytest.raises(pc.ArrowError):
        pc.Cast(pc.Scalar(300, pc.INT32), pc.UINT8)


@pytest.mark.parametrize("fname", ["add", "sub", "multiply", "add_unchecked"])
def test_arithmetic_implicit_promotion(fname):
    """exec.go:101-121: DispatchBest picks the common numeric type and the arguments are cast with
    SafeCastOptions before the kernel runs (arithmetic_test.go:715-753 names the expected types)."""
    rng = np.random.default_rng(11)
    n = 3000
    npf = {"add": np.add, "sub": np.subtract, "multiply": np.multiply, "add_unchecked": np.add}[fname]
    combos = [(pc.INT32, pc.INT64, pc.INT64), (pc.INT8, pc.UINT8, pc.INT16), (pc.UINT16, pc.INT32, pc.INT32),
              (pc.INT32, pc.FLOAT64, pc.FLOAT64), (pc.FLOAT32, pc.INT16, pc.FLOAT32), (pc.UINT8, pc.UINT32, pc.UINT32)]
    for lt, rt, want_t in combos:
        # left operand >= right operand so that checked "sub" on unsigned types stays in range
        l = rng.integers(20, 31, n).astype(pc.NP_OF[lt]) if lt not in (pc.FLOAT32, pc.FLOAT64) else rng.standard_normal(n).astype(pc.NP_OF[lt])
        r = rng.integers(0, 11, n).astype(pc.NP_OF[rt]) if rt not in (pc.FLOAT32, pc.FLOAT64) else rng.standard_normal(n).astype(pc.NP_OF[rt])
        lv, rv = rng.random(n) > 0.1, rng.random(n) > 0.1
        out = pc.CallFunction(fname, [pc.Array.from_numpy(l, lv), pc.Array.from_numpy(r, rv)])
        assert out.type == want_t, (lt, rt)
        vals, valid, _ = out.to_numpy()
        assert np.array_equal(valid, lv & rv)
        want = npf(l.astype(pc.NP_OF[want_t]), r.astype(pc.NP_OF[want_t]))
        assert np.array_equal(vals[valid], want[valid]), (fname, lt, rt)
        # array (op) scalar of another type: the scalar is cast too
        out = pc.CallFunction(fname, [pc.Array.from_numpy(l, lv), pc.Scalar(r[0].item(), rt)])
        assert out.type == want_t
        vals, valid, _ = out.to_numpy()
        want = npf(l.astype(pc.NP_OF[want_t]), r[:1].astype(pc.NP_OF[want_t]))
        assert np.array_equal(valid, lv) and np.array_equal(vals[valid], want[valid])
    # int64 -> float64 promotion is a SAFE cast: values beyond 2^53 make the call fail
    big = pc.Array.from_pylist([1, (1 << 53) + 1], pc.INT64)
    with pytest.raises(pc.ArrowError) as e:
        pc.CallFunction("add", [big, pc.Array.from_pylist([0.5, 0.5], pc.FLOAT64)])
    assert "not in range" in e.value.msg
    # compare promotes the same way
    out = pc.CallFunction("greater", [pc.Array.from_pylist([1, 200, None], pc.UINT8), pc.Array.from_pylist([-1, 300, 5], pc.INT16)])
    assert out.to_pylist() == [True, False, None]
    # chunked (op) array of another type
    l = rng.integers(-100, 100, n).astype(np.int32)
    r = rng.standard_normal(n)
    cuts = [0, 7, 1500, n]
    out = pc.CallFunction(fname, [pc.Chunked([pc.Array.from_numpy(l[s:e]) for s, e in zip(cuts, cuts[1:])], pc.INT32), pc.Array.from_numpy(r)])
    vals, _, _ = out.to_numpy()
    assert out.type == pc.FLOAT64 and np.array_equal(vals, npf(l.astype(np.float64), r))


# ---------------------------------------------------------------- cumulative_sum -------------
def test_cumulative_sum_reference_vectors():
    """arrow/compute/vector_cumulative_test.go through compute.CumulativeSum[Checked] (tests/golden/cumulative_sum.json)."""
    name_to_id = {"int8": pc.INT8, "int16": pc.INT16, "int32": pc.INT32, "int64": pc.INT64, "uint8": pc.UINT8, "uint16": pc.UINT16,
                  "uint32": pc.UINT32, "uint64": pc.UINT64, "float32": pc.FLOAT32, "float64": pc.FLOAT64}
    for case in load("cumulative_sum.json")["cases"]:
        t = name_to_id[case["type"]]
        chunks = [pc.Array.from_pylist(c, t) for c in case["chunks"]]
        values = chunks[0] if len(chunks) == 1 else pc.Chunked(chunks, t)
        start = pc.Scalar(case["start"], pc.INT64) if "start" in case else None   # :268-270: an int64 start for int32 input
        kw = dict(start=start, skip_nulls=bool(case.get("skip_nulls")), checked=bool(case.get("checked")))
        if case.get("fails"):
            with pytest.raises(pc.ArrowError) as e:
                pc.CumulativeSum(values, **kw)
            assert e.value.sentinel == "ErrInvalid" and "overflow" in e.value.msg, case
            continue
        out = pc.CumulativeSum(values, **kw)
        assert out.type == t and out.to_pylist() == case["out"], case
        if len(chunks) > 1:   # TestCumulativeSumChunked: a chunked result with ONE chunk
            assert out.kind == pc.KIND_CHUNKED and len(out.chunks()) == 1
    # scalar input (TestCumulativeSumAdditionalInputs :141-147) and a sliced input (:149-160)
    assert pc.CumulativeSum(pc.Scalar(3, pc.INT32)).to_pylist() == [3]
    assert pc.CumulativeSum(pc.Array.from_pylist([0, 1, 2, 3], pc.INT32).slice(1, 2)).to_pylist() == [1, 3]
    # start value: null scalar rejected (:277-301), unsafe start rejected by the safe cast (:303-344)
    with pytest.raises(pc.ArrowError) as e:
        pc.CumulativeSum(pc.Array.from_pylist([1], pc.INT32), start=pc.Scalar(None, pc.INT32))
    assert "must be valid" in e.value.msg
    with pytest.raises(pc.ArrowError) as e:
        pc.CumulativeSum(pc.Array.from_pylist([1], pc.INT8), start=pc.Scalar(300, pc.INT32))
    assert "cannot cast cumulative sum start value" in e.value.msg
    with pytest.raises(pc.ArrowError) as e:
        pc.CumulativeSum(pc.Array.from_pylist([True], pc.BOOL))
    assert e.value.sentinel == "ErrType"


def test_cumulative_sum_chunked_random():
    rng = np.random.default_rng(21)
    n = 300_000
    x = rng.integers(-1000, 1000, n).astype(np.int64)
    valid = rng.random(n) > 0.001
    cuts = [0, 1, 1, 40_000, 40_001, 250_000, n]
    c = pc.Chunked([pc.Array.from_numpy(x[a:b], valid[a:b]) for a, b in zip(cuts, cuts[1:])], pc.INT64)
    out = pc.CumulativeSum(c, skip_nulls=True, start=pc.Scalar(5, pc.INT8))
    vals, v, nulls = out.to_numpy()
    want = np.cumsum(np.where(valid, x, 0)) + 5
    assert np.array_equal(v, valid) and np.array_equal(vals[v], want[v]) and nulls == int((~valid).sum())
    out = pc.CumulativeSum(c)      # propagate: everything from the first null on is null
    vals, v, nulls = out.to_numpy()
    first = int(np.argmin(valid))
    assert v[:first].all() and not v[first:].any() and np.array_equal(vals[:first], np.cumsum(x[:first])) and nulls == n - first


# ---------------------------------------------------------------- sort_indices / unique / is_in ---
TYPE_BY_NAME = {"int32": pc.INT32, "uint64": pc.UINT64, "float64": pc.FLOAT64}


def test_sort_indices_reference_cases():
    """TestSortIndices (vector_sort_test.go:40-325), the fixed-width numeric cases, through compute.SortIndices."""
    for case in load("sort_indices.json")["cases"]:
        t = TYPE_BY_NAME[case["type"]]
        vals = [None if v is None else (float("nan") if v == "NaN" else v) for v in case["values"]]
        if not vals:
            continue
        out = pc.SortIndices(pc.Array.from_pylist(vals, t), order=case["order"], null_placement=case["null_placement"])
        assert out.type == pc.UINT64 and out.to_pylist() == case["expected"], case["name"]
    # registry name, default options (Ascending, NullsAtEnd)
    assert pc.CallFunction("sort_indices", [pc.Array.from_pylist([3, None, 1], pc.INT64)]).to_pylist() == [2, 0, 1]
    # sliced input: offsets apply to values and validity alike
    a = pc.Array.from_pylist([9, 3, None, 1, 2, 8], pc.INT16).slice(1, 4)
    assert pc.SortIndices(a, order=pc.DESCENDING, null_placement=pc.NULLS_AT_START).to_pylist() == [1, 0, 3, 2]


@pytest.mark.parametrize("t", NUMERIC)
def test_is_in_and_unique_reference_cases(t):
    """TestIsInPrimitive (scalar_set_lookup_test.go:104-167) and PrimitiveHashKernelSuite.TestUnique
    (vector_hash_test.go:236-255) for every numeric type."""
    g = load("set_lookup.json")
    for case in g["is_in"]:
        if not case["input"]:
            continue
        for matching, expected in case["cases"]:
            out = pc.IsIn(pc.Array.from_pylist(case["input"], t), pc.Array.from_pylist(case["set"], t), null_behavior=matching)
            assert out.type == pc.BOOL and out.to_pylist() == expected, (case["name"], matching)
    for case in g["unique"]:
        out = pc.Unique(pc.Array.from_pylist(case["input"], t))
        assert out.type == t and out.to_pylist() == case["expected"]
    a = pc.Array.from_pylist([1, 2, None, 3, 2, None], t).slice(1, 4)   # vector_hash_test.go:244-254
    assert pc.Unique(a).to_pylist() == [2, None, 3]
    with pytest.raises(pc.ArrowError):
        pc.IsIn(pc.Array.from_pylist([1], t), pc.Array.from_pylist([1], pc.BOOL))
