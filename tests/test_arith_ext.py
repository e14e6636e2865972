"""divide / bit_wise_* / shift_* (SURVEY Appendix A names that round 1 left out; VERDICT "missing" item 5).

Reference: kernels base_arithmetic.go:154-161,287-294,386-397 (Div), scalar_arithmetic.go:191-259 (bitwise),
:293-412 (shifts); registration arithmetic.go:782-785,944-996.  The Go-only logic has no native counterpart to
assemble, so the oracle is the C restatement (oracle/cpu_ref.c), pinned here by the literal cases of
arithmetic_test.go:427-480,571-680 and by pyarrow's kernels of the same names (CPU part); the CUDA path is then
compared with the oracle (GPU part) and driven through the host mirror like the Go suites do."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pyarrow.compute as pac
import pytest

from arrow_go_b200 import _native as N
from helpers import INT_TYPES, NP_OF, TYPE_NAME, pack_bits, ptr, random_values, unpack_bits

gpu = pytest.mark.gpu
NO_POS = (1 << 63) - 1
SHIFT_MSG = "shift amount must be >= 0 and less than precision of type"
BITOPS = {N.OP_BIT_AND: ("bit_wise_and", np.bitwise_and), N.OP_BIT_OR: ("bit_wise_or", np.bitwise_or), N.OP_BIT_XOR: ("bit_wise_xor", np.bitwise_xor)}


def oracle_checked(cpu, t, op, shape, l, lv, loff, r, rv, roff, n):
    out = np.full(n, 7, dtype=NP_OF[t])
    bad = C.c_int64()
    st = cpu.ref_arith_checked(t, op, shape, ptr(l), ptr(lv), loff, ptr(r), ptr(rv), roff, ptr(out), n, C.byref(bad))
    return st, out, bad.value


# ------------------------------------------------------------------ CPU: the oracle is pinned ----
@pytest.mark.parametrize("t", INT_TYPES, ids=lambda t: TYPE_NAME[t])
def test_oracle_shift_literals_of_the_reference(cpu, t):
    """TestShiftLeft / TestShiftRight / Test*OverflowError (arithmetic_test.go:571-680)."""
    dt = NP_OF[t]
    info = np.iinfo(dt)
    signed = info.min < 0
    bits = info.bits - (1 if signed else 0)          # the test's bitWidth

    def run(op, l, r):
        l, r = np.array(l, dtype=dt), np.array(r, dtype=dt)
        return oracle_checked(cpu, t, op, N.SHAPE_AA, l, None, 0, r, None, 0, len(l))

    for op in (N.OP_SHIFT_LEFT, N.OP_SHIFT_LEFT_CHECKED):
        st, out, _ = run(op, [0, 1, 2, 3], [2, 3, 4, 5])
        assert st == 0 and out.tolist() == [0, 8, 32, 96]
    for op in (N.OP_SHIFT_RIGHT, N.OP_SHIFT_RIGHT_CHECKED):
        st, out, _ = run(op, [0, 1, 4, 8], [1, 1, 1, 4])
        assert st == 0 and out.tolist() == [0, 0, 2, 0]
    st, out, _ = run(N.OP_SHIFT_LEFT_CHECKED, [1], [bits - 1])
    assert st == 0 and out[0] == dt(1 << (bits - 1))
    st, out, _ = run(N.OP_SHIFT_LEFT_CHECKED, [4], [bits - 1])
    assert st == 0 and out[0] == 0                    # shifted past the top
    st, _, bad = run(N.OP_SHIFT_LEFT_CHECKED, [1], [bits])
    assert st == N.AG_ERR_INVALID and bad == 0        # == precision of the type: refused
    st, out, _ = run(N.OP_SHIFT_RIGHT_CHECKED, [info.max], [bits - 1])
    assert st == 0 and out[0] == 1
    st, _, _ = run(N.OP_SHIFT_RIGHT_CHECKED, [1], [bits])
    assert st == N.AG_ERR_INVALID
    if signed:
        st, out, _ = run(N.OP_SHIFT_LEFT_CHECKED, [2], [bits - 1])
        assert st == 0 and out[0] == info.min         # a bit into the sign bit
        st, out, _ = run(N.OP_SHIFT_LEFT_CHECKED, [info.min], [1])
        assert st == 0 and out[0] == 0
        st, _, bad = run(N.OP_SHIFT_LEFT_CHECKED, [1, 2], [1, -1])
        assert st == N.AG_ERR_INVALID and bad == 1
        st, out, _ = run(N.OP_SHIFT_LEFT, [1, 1], [-1, bits])
        assert st == 0 and out.tolist() == [1, 1]     # unchecked: lhs comes back
        st, out, _ = run(N.OP_SHIFT_RIGHT_CHECKED, [-1, -1], [1, 5])
        assert st == 0 and out.tolist() == [-1, -1]   # arithmetic shift
        st, out, _ = run(N.OP_SHIFT_RIGHT_CHECKED, [info.min], [1])
        assert st == 0 and out[0] == info.min // 2
        st, out, _ = run(N.OP_SHIFT_RIGHT, [1, 1], [-1, bits])
        assert st == 0 and out.tolist() == [1, 1]


@pytest.mark.parametrize("t", INT_TYPES, ids=lambda t: TYPE_NAME[t])
def test_oracle_bitwise_and_shifts_vs_pyarrow(cpu, t):
    rng = np.random.default_rng(1000 + t)
    dt = NP_OF[t]
    info = np.iinfo(dt)
    n = 5000
    a, b = random_values(rng, t, n), random_values(rng, t, n)
    for op, (name, _) in BITOPS.items():
        out = np.zeros(n, dtype=dt)
        assert cpu.ref_arith_binary(t, op, N.SHAPE_AA, ptr(a), ptr(b), ptr(out), n) == 0
        assert np.array_equal(out, getattr(pac, name)(pa.array(a), pa.array(b)).to_numpy())
    out = np.zeros(n, dtype=dt)
    assert cpu.ref_arith_unary_same(t, N.OP_BIT_NOT, ptr(a), ptr(out), n) == 0
    assert np.array_equal(out, pac.bit_wise_not(pa.array(a)).to_numpy())
    # shifts: amounts around the limits, both signs
    sh = rng.integers(-2 if info.min < 0 else 0, info.bits + 2, n).astype(dt)
    for op, name in ((N.OP_SHIFT_LEFT, "shift_left"), (N.OP_SHIFT_RIGHT, "shift_right")):
        st, out, _ = oracle_checked(cpu, t, op, N.SHAPE_AA, a, None, 0, sh, None, 0, n)
        assert st == 0 and np.array_equal(out, getattr(pac, name)(pa.array(a), pa.array(sh)).to_numpy()), name
    valid_max = info.bits - (1 if info.min < 0 else 0)
    ok = rng.integers(0, valid_max, n).astype(dt)
    for op, name in ((N.OP_SHIFT_LEFT_CHECKED, "shift_left_checked"), (N.OP_SHIFT_RIGHT_CHECKED, "shift_right_checked")):
        st, out, _ = oracle_checked(cpu, t, op, N.SHAPE_AA, a, None, 0, ok, None, 0, n)
        assert st == 0 and np.array_equal(out, getattr(pac, name)(pa.array(a), pa.array(ok)).to_numpy()), name
        bad_sh = ok.copy()
        bad_sh[1234] = valid_max
        st, _, bad = oracle_checked(cpu, t, op, N.SHAPE_AA, a, None, 0, bad_sh, None, 0, n)
        assert st == N.AG_ERR_INVALID and bad == 1234
        with pytest.raises(pa.ArrowInvalid):
            getattr(pac, name)(pa.array(a), pa.array(bad_sh))


@pytest.mark.parametrize("t", [N.FLOAT32, N.FLOAT64], ids=lambda t: TYPE_NAME[t])
def test_oracle_float_divide_literals_and_pyarrow(cpu, t):
    """TestDiv / TestDivideByZero (arithmetic_test.go:427-480) for the floating point kernels."""
    dt = NP_OF[t]

    def run(op, l, r, lv=None):
        l, r = np.array(l, dtype=dt), np.array(r, dtype=dt)
        return oracle_checked(cpu, t, op, N.SHAPE_AA, l, lv, 0, r, None, 0, len(l))

    for op in (N.OP_DIV, N.OP_DIV_CHECKED):
        st, out, _ = run(op, [3.4, 0.64, 1.28], [1, 2, 4])
        assert st == 0 and np.array_equal(out, np.array([3.4, 0.32, 0.32], dtype=dt))
        st, out, _ = run(op, [3.4, np.inf, -np.inf], [1, 2, 3])
        assert st == 0 and out.tolist()[1:] == [np.inf, -np.inf]
    for num in (6, 0, -6):
        st, _, bad = run(N.OP_DIV_CHECKED, [3, 2, num], [1, 1, 0])
        assert st == N.AG_ERR_INVALID and bad == 2
    st, out, _ = run(N.OP_DIV, [3, 2, 6, 0, -6], [1, 1, 0, 0, 0])
    assert st == 0 and out[2] == np.inf and np.isnan(out[3]) and out[4] == -np.inf
    # a zero divisor under a null slot is not looked at (ScalarBinaryNotNull)
    st, out, _ = run(N.OP_DIV_CHECKED, [3, 2, 6], [1, 1, 0], lv=np.array([0b011], dtype=np.uint8))
    assert st == 0 and out.tolist() == [3, 2, 0]
    rng = np.random.default_rng(5)
    a, b = rng.standard_normal(4000).astype(dt), rng.standard_normal(4000).astype(dt)
    st, out, _ = oracle_checked(cpu, t, N.OP_DIV, N.SHAPE_AA, a, None, 0, b, None, 0, 4000)
    assert st == 0 and np.array_equal(out, pac.divide(pa.array(a), pa.array(b)).to_numpy())


# ------------------------------------------------------------------ GPU: CUDA path vs the oracle ----
@gpu
@pytest.mark.parametrize("t", INT_TYPES, ids=lambda t: TYPE_NAME[t])
def test_gpu_bitwise_matches_oracle(ag, cpu, t):
    rng = np.random.default_rng(2000 + t)
    dt = NP_OF[t]
    for n in (1, 33, 4097, 300_001):
        a, b = random_values(rng, t, n + 3), random_values(rng, t, n + 3)
        for op in BITOPS:
            for shape, (l, r) in ((N.SHAPE_AA, (a[1:], b[3:])), (N.SHAPE_AS, (a[2:], b[:1])), (N.SHAPE_SA, (a[:1], b[1:]))):
                want, got = np.zeros(n, dtype=dt), np.zeros(n, dtype=dt)
                assert cpu.ref_arith_binary(t, op, shape, ptr(l), ptr(r), ptr(want), n) == 0
                fn = {N.SHAPE_AA: "ag_arith_binary", N.SHAPE_AS: "ag_arith_arr_scalar", N.SHAPE_SA: "ag_arith_scalar_arr"}[shape]
                ag.call(fn, t, op, ptr(l), ptr(r), ptr(got), n)
                assert got.tobytes() == want.tobytes(), (TYPE_NAME[t], op, shape, n)
        want, got = np.zeros(n, dtype=dt), np.zeros(n, dtype=dt)
        assert cpu.ref_arith_unary_same(t, N.OP_BIT_NOT, ptr(a[1:]), ptr(want), n) == 0
        ag.call("ag_arith_unary_same", t, N.OP_BIT_NOT, ptr(a[1:]), ptr(got), n)
        assert got.tobytes() == want.tobytes()
    st, msg = ag.call_status("ag_arith_binary", N.FLOAT64, N.OP_BIT_AND, ptr(np.zeros(4)), ptr(np.zeros(4)), ptr(np.zeros(4)), 4)
    assert st == N.AG_ERR_TYPE and "integer" in msg


@gpu
@pytest.mark.parametrize("t", INT_TYPES, ids=lambda t: TYPE_NAME[t])
@pytest.mark.parametrize("op", [N.OP_SHIFT_LEFT, N.OP_SHIFT_RIGHT, N.OP_SHIFT_LEFT_CHECKED, N.OP_SHIFT_RIGHT_CHECKED])
def test_gpu_shifts_match_oracle(ag, cpu, t, op):
    rng = np.random.default_rng(t * 131 + op)
    dt = NP_OF[t]
    info = np.iinfo(dt)
    valid_max = info.bits - (1 if info.min < 0 else 0)
    for shape in (N.SHAPE_AA, N.SHAPE_AS, N.SHAPE_SA):
        for n in (1, 65, 1000, 70_001):
            for legal in (True, False):
                for nullp in (0.0, 0.3):
                    l = random_values(rng, t, 1 if shape == N.SHAPE_SA else n)
                    lo = 0 if (legal or info.min == 0) else -2
                    hi = valid_max if legal else info.bits + 2
                    r = rng.integers(lo, hi, 1 if shape == N.SHAPE_AS else n).astype(dt)
                    loff, roff = 3, 5
                    lv = pack_bits(rng.random(n) >= nullp, loff) if (nullp and shape != N.SHAPE_SA) else None
                    rv = pack_bits(rng.random(n) >= nullp, roff) if (nullp and shape != N.SHAPE_AS) else None
                    wst, want, wbad = oracle_checked(cpu, t, op, shape, l, lv, loff, r, rv, roff, n)
                    got = np.full(n, 9, dtype=dt)
                    gbad = C.c_int64()
                    gst, msg = ag.call_status("ag_arith_checked", t, op, shape, ptr(l), ptr(lv), loff, ptr(r), ptr(rv), roff, ptr(got), n, C.byref(gbad))
                    assert gst == wst and gbad.value == wbad, (TYPE_NAME[t], op, shape, n, legal, nullp, msg)
                    if wst == 0:
                        assert got.tobytes() == want.tobytes()


@gpu
@pytest.mark.parametrize("t", [N.FLOAT32, N.FLOAT64], ids=lambda t: TYPE_NAME[t])
@pytest.mark.parametrize("op", [N.OP_DIV, N.OP_DIV_CHECKED])
def test_gpu_float_divide_matches_oracle(ag, cpu, t, op):
    rng = np.random.default_rng(t * 17 + op)
    dt = NP_OF[t]
    for shape in (N.SHAPE_AA, N.SHAPE_AS, N.SHAPE_SA):
        for n in (1, 65, 4096, 100_003):
            for zeros in (False, True):
                for nullp in (0.0, 0.2):
                    l = random_values(rng, t, 1 if shape == N.SHAPE_SA else n)
                    r = random_values(rng, t, 1 if shape == N.SHAPE_AS else n, small=True)
                    if zeros and r.size > 1:
                        r[rng.integers(0, r.size, 3)] = 0
                    loff, roff = 1, 6
                    lv = pack_bits(rng.random(n) >= nullp, loff) if (nullp and shape != N.SHAPE_SA) else None
                    rv = pack_bits(rng.random(n) >= nullp, roff) if (nullp and shape != N.SHAPE_AS) else None
                    wst, want, wbad = oracle_checked(cpu, t, op, shape, l, lv, loff, r, rv, roff, n)
                    got = np.full(n, 9, dtype=dt)
                    gbad = C.c_int64()
                    gst, msg = ag.call_status("ag_arith_checked", t, op, shape, ptr(l), ptr(lv), loff, ptr(r), ptr(rv), roff, ptr(got), n, C.byref(gbad))
                    assert gst == wst and gbad.value == wbad, (TYPE_NAME[t], op, shape, n, zeros, nullp, msg)
                    if wst == 0:
                        nan = np.isnan(want)                         # IEEE division: bit-exact; NaNs compared by class (x86's 0/0 is
                        assert np.array_equal(np.isnan(got), nan)    # the negative default NaN, CUDA's the canonical positive one)
                        assert got[~nan].tobytes() == want[~nan].tobytes()
                    else:
                        assert msg == "divide by zero"


# ------------------------------------------------------------------ GPU: the host mirror, like the Go suites ----
@gpu
def test_host_mirror_divide_shift_bitwise(ag):
    from arrow_go_b200 import compute as pc
    A = pc.Array.from_pylist
    for t in (pc.INT8, pc.UINT16, pc.INT32, pc.INT64, pc.UINT64):
        for fn in ("divide", "divide_unchecked"):       # TestDiv, integers
            assert pc.CallFunction(fn, [A([3, 2, 6], t), A([1, 1, 2], t)]).to_pylist() == [3, 2, 3]
            assert pc.CallFunction(fn, [A([None, 10, 30, None, 20], t), A([1, 5, 2, 5, 10], t)]).to_pylist() == [None, 2, 15, None, 2]
            assert pc.CallFunction(fn, [pc.Scalar(33, t), A([None, 1, 3, None, 2], t)]).to_pylist() == [None, 33, 11, None, 16]
            assert pc.CallFunction(fn, [A([None, 10, 30, None, 2], t), pc.Scalar(3, t)]).to_pylist() == [None, 3, 10, None, 0]
            with pytest.raises(pc.ArrowError) as e:     # TestDivideByZero: integers fail in both flavours
                pc.CallFunction(fn, [A([3, 2, 6], t), A([1, 1, 0], t)])
            assert e.value.sentinel == "ErrInvalid" and "divide by zero" in e.value.msg
        for fn in ("shift_left", "shift_left_unchecked"):   # TestShiftLeft
            assert pc.CallFunction(fn, [A([0, 1, 2, 3], t), A([2, 3, 4, 5], t)]).to_pylist() == [0, 8, 32, 96]
            assert pc.CallFunction(fn, [A([0, None, 2, 3], t), A([2, 3, None, 5], t)]).to_pylist() == [0, None, None, 96]
            assert pc.CallFunction(fn, [pc.Scalar(2, t), A([None, 5], t)]).to_pylist() == [None, 64]
            assert pc.CallFunction(fn, [A([None, 5], t), pc.Scalar(3, t)]).to_pylist() == [None, 40]
        for fn in ("shift_right", "shift_right_unchecked"):  # TestShiftRight
            assert pc.CallFunction(fn, [A([0, 1, 4, 8], t), A([1, 1, 1, 4], t)]).to_pylist() == [0, 0, 2, 0]
            assert pc.CallFunction(fn, [pc.Scalar(64, t), A([None, 2, 6], t)]).to_pylist() == [None, 16, 1]
        bits = np.iinfo(pc.NP_OF[t]).bits - (1 if np.iinfo(pc.NP_OF[t]).min < 0 else 0)
        with pytest.raises(pc.ArrowError) as e:
            pc.CallFunction("shift_left", [A([1], t), A([bits], t)])
        assert SHIFT_MSG in e.value.msg
        assert pc.CallFunction("shift_left_unchecked", [A([1], t), A([bits], t)]).to_pylist() == [1]
        assert pc.CallFunction("bit_wise_and", [A([0b1100, None, 7], t), A([0b1010, 1, None], t)]).to_pylist() == [0b1000, None, None]
        assert pc.CallFunction("bit_wise_or", [A([0b1100, 1], t), pc.Scalar(0b0011, t)]).to_pylist() == [0b1111, 3]
        assert pc.CallFunction("bit_wise_xor", [pc.Scalar(0b0110, t), A([0b1100, None], t)]).to_pylist() == [0b1010, None]
        allones = int(np.iinfo(pc.NP_OF[t]).max) if np.iinfo(pc.NP_OF[t]).min == 0 else -1
        assert pc.CallFunction("bit_wise_not", [A([0, None, allones], t)]).to_pylist() == [allones, None, 0]
    for t in (pc.FLOAT32, pc.FLOAT64):                   # TestDiv / TestDivideByZero, floats
        got = pc.CallFunction("divide", [A([None, 1.0, 3.5, None, 2.0], t), A([1.0, 4.0, 2.0, 5.0, 0.5], t)]).to_pylist()
        assert got == [None, 0.25, 1.75, None, 4.0]
        with pytest.raises(pc.ArrowError) as e:
            pc.CallFunction("divide", [A([3.0, 2.0, 6.0], t), A([1.0, 1.0, 0.0], t)])
        assert "divide by zero" in e.value.msg
        got = pc.CallFunction("divide_unchecked", [A([3.0, 2.0, 6.0, -6.0], t), A([1.0, 1.0, 0.0, 0.0], t)]).to_pylist()
        assert got == [3.0, 2.0, float("inf"), float("-inf")]
    # implicit promotion like every arithmetic function (arithmeticFunction.DispatchBest)
    assert pc.CallFunction("divide", [A([7, 9], pc.INT8), A([2, 3], pc.INT32)]).type == pc.INT32
    names = pc.function_names()
    for nm in ("divide", "divide_unchecked", "bit_wise_and", "bit_wise_or", "bit_wise_xor", "bit_wise_not",
               "shift_left", "shift_left_unchecked", "shift_right", "shift_right_unchecked"):
        assert nm in names
