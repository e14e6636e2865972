"""Pins the oracle (oracle/cpu_ref.c) against the reference's OWN instruction stream:
oracle/_ref/libarrowgo_ref.so is assembled from the clang output the reference checks in under
_lib/*.s (the code c2goasm turned into the Plan9 assembly Go links), see oracle/Makefile.
Covers every function that has a native counterpart in the reference: Sum, arithmetic
(binary / unary / sign), comparisons (incl. the reference's own offset x length sweep),
byte-aligned bitmap ops."""
import ctypes as C

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import ALL_TYPES, INT_TYPES, NP_OF, TYPE_NAME, pack_bits, ptr, random_values, same_bits, same_float_class

ISAS = ["avx2", "sse4"]


def test_sum_known_answers(cpu, ref):
    # arrow/math/float64_test.go:30-48, int64_test.go, uint64_test.go
    x = np.arange(10000, dtype=np.float64)
    assert cpu.ref_sum_f64_avx2_order(ptr(x), x.size) == 49995000.0
    assert cpu.ref_sum_f64_sequential(ptr(x), x.size) == 49995000.0
    xi, xu = x.astype(np.int64), x.astype(np.uint64)  # keep the buffers alive across the calls
    assert cpu.ref_sum_i64(ptr(xi), x.size) == 49995000
    assert cpu.ref_sum_u64(ptr(xu), x.size) == 49995000
    assert cpu.ref_sum_f64_avx2_order(None, 0) == 0.0
    for isa in ISAS:
        r = C.c_double()
        getattr(ref, f"sum_float64_{isa}")(ptr(x), x.size, C.addressof(r))
        assert r.value == 49995000.0


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 63, 64, 65, 255, 256, 1000, 8192, 100003])
def test_sum_f64_association_order_is_the_avx2_one(cpu, ref, n):
    """General (non exactly-summable) data: only the reference's exact association reproduces
    its bits.  Also documents that the reference has three different answers (AVX2 / SSE4 /
    pure Go), SURVEY.md §3.4."""
    rng = np.random.default_rng(0x0FF1CE + n)
    for dist in ("normal", "lognormal"):
        x = rng.standard_normal(n) if dist == "normal" else rng.lognormal(0, 4, n)
        r = C.c_double()
        ref.sum_float64_avx2(ptr(x), n, C.addressof(r))
        assert cpu.ref_sum_f64_avx2_order(ptr(x), n) == r.value


@pytest.mark.parametrize("n", [0, 1, 33, 8192, 100003])
def test_sum_ints_wrap(cpu, ref, n):
    rng = np.random.default_rng(n)
    x = rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, n, dtype=np.int64, endpoint=True)
    for isa in ISAS:
        r = C.c_int64()
        getattr(ref, f"sum_int64_{isa}")(ptr(x), n, C.addressof(r))
        assert cpu.ref_sum_i64(ptr(x), n) == r.value
        ru = C.c_uint64()
        xu = x.view(np.uint64)
        getattr(ref, f"sum_uint64_{isa}")(ptr(xu), n, C.addressof(ru))
        assert cpu.ref_sum_u64(ptr(x), n) == ru.value


@pytest.mark.parametrize("type_id", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_arithmetic_binary(cpu, ref, type_id):
    rng = np.random.default_rng(type_id)
    isf = type_id in (N.FLOAT32, N.FLOAT64)
    eq = same_float_class if isf else same_bits
    for n in (0, 1, 7, 33, 1000, 4099):
        for shape, nm in ((0, "binary"), (1, "arr_scalar"), (2, "scalar_arr")):
            l = random_values(rng, type_id, 1 if shape == 2 else n)
            r = random_values(rng, type_id, 1 if shape == 1 else n)
            for op in (N.OP_ADD, N.OP_SUB, N.OP_MUL, N.OP_ADD_CHECKED, N.OP_SUB_CHECKED, N.OP_MUL_CHECKED):
                mine = np.empty(n, dtype=NP_OF[type_id])
                assert cpu.ref_arith_binary(type_id, op, shape, ptr(l), ptr(r), ptr(mine), n) == 0
                for isa in ISAS:
                    want = np.empty(n, dtype=NP_OF[type_id])
                    getattr(ref, f"arithmetic_{nm}_{isa}")(type_id, op, ptr(l), ptr(r), ptr(want), n)
                    assert eq(mine, want), (TYPE_NAME[type_id], nm, op, n, isa)


@pytest.mark.parametrize("type_id", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_arithmetic_unary(cpu, ref, type_id):
    rng = np.random.default_rng(type_id + 100)
    isf = type_id in (N.FLOAT32, N.FLOAT64)
    eq = same_float_class if isf else same_bits
    for n in (0, 1, 33, 4099):
        x = random_values(rng, type_id, n)
        if n > 4 and not isf:
            info = np.iinfo(NP_OF[type_id])
            x[:3] = [info.min, info.max, 0]
        for op in (N.OP_ABS, N.OP_ABS_CHECKED, N.OP_NEGATE, N.OP_NEGATE_CHECKED, N.OP_SIGN):
            mine = np.empty(n, dtype=NP_OF[type_id])
            assert cpu.ref_arith_unary_same(type_id, op, ptr(x), ptr(mine), n) == 0
            for isa in ISAS:
                want = np.empty(n, dtype=NP_OF[type_id])
                getattr(ref, f"arithmetic_unary_same_types_{isa}")(type_id, op, ptr(x), ptr(want), n)
                assert eq(mine, want), (TYPE_NAME[type_id], op, n, isa)
    if not isf:
        x = random_values(rng, type_id, 1000, small=True)
        for otype in INT_TYPES:
            mine = np.empty(1000, dtype=NP_OF[otype])
            assert cpu.ref_arith_unary_diff(type_id, otype, N.OP_SIGN, ptr(x), ptr(mine), 1000) == 0
            want = np.empty(1000, dtype=NP_OF[otype])
            ref.arithmetic_unary_diff_type_avx2(type_id, otype, N.OP_SIGN, ptr(x), ptr(want), 1000)
            assert same_bits(mine, want), (TYPE_NAME[type_id], TYPE_NAME[otype])


CMP_NAMES = {N.CMP_EQ: "equal", N.CMP_NE: "not_equal", N.CMP_GT: "greater", N.CMP_GE: "greater_equal"}
SHAPE_NAMES = {0: "arr_arr", 1: "arr_scalar", 2: "scalar_arr"}


def test_compare_reference_sweep(cpu, ref):
    """kernels/scalar_comparisons_test.go:31-119: offsets 0..7 x lengths 0..65 x 3 shapes x 4 ops,
    output pre-filled 0xa5, left[i]=(7i+1)%11, right[i]=(5i+3)%11, scalars 6 and 4."""
    i = np.arange(65)
    for type_id in (N.INT32, N.INT64, N.FLOAT64):
        dt = NP_OF[type_id]
        left, right = ((7 * i + 1) % 11).astype(dt), ((5 * i + 3) % 11).astype(dt)
        six, four = np.array([6], dtype=dt), np.array([4], dtype=dt)
        for offset in range(8):
            for n in range(66):
                for cmp, cname in CMP_NAMES.items():
                    for shape, (l, r) in {0: (left, right), 1: (left, six), 2: (four, right)}.items():
                        nbytes = (offset + n + 7) // 8 + 2
                        mine = np.full(nbytes, 0xA5, dtype=np.uint8)
                        assert cpu.ref_compare(type_id, cmp, shape, ptr(l), ptr(r), ptr(mine), n, offset) == 0
                        # naive expectation, like the Go test builds it
                        a = l if shape != 2 else np.repeat(l, n)
                        b = r if shape != 1 else np.repeat(r, n)
                        exp = {N.CMP_EQ: a[:n] == b[:n], N.CMP_NE: a[:n] != b[:n], N.CMP_GT: a[:n] > b[:n], N.CMP_GE: a[:n] >= b[:n]}[cmp]
                        assert mine.tobytes() == pack_bits(exp, offset, 0xA5, pad_bytes=nbytes - (offset + n + 7) // 8).tobytes()
                        for isa in ISAS:
                            want = np.full(nbytes, 0xA5, dtype=np.uint8)
                            getattr(ref, f"comparison_{cname}_{SHAPE_NAMES[shape]}_{isa}")(type_id, ptr(l), ptr(r), ptr(want), n, offset)
                            assert mine.tobytes() == want.tobytes(), (TYPE_NAME[type_id], offset, n, cname, shape, isa)


@pytest.mark.parametrize("type_id", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_compare_random(cpu, ref, type_id):
    rng = np.random.default_rng(type_id + 7)
    for n in (1, 33, 1000, 4099):
        for shape in range(3):
            l = random_values(rng, type_id, 1 if shape == 2 else n, small=True)
            r = random_values(rng, type_id, 1 if shape == 1 else n, small=True)
            if type_id in (N.FLOAT32, N.FLOAT64) and n > 8 and shape != 2:
                l[[1, 5]] = [np.nan, -0.0]
            for cmp, cname in CMP_NAMES.items():
                off = int(rng.integers(0, 8))
                mine = np.full((off + n + 7) // 8 + 1, 0x5A, dtype=np.uint8)
                want = mine.copy()
                assert cpu.ref_compare(type_id, cmp, shape, ptr(l), ptr(r), ptr(mine), n, off) == 0
                getattr(ref, f"comparison_{cname}_{SHAPE_NAMES[shape]}_avx2")(type_id, ptr(l), ptr(r), ptr(want), n, off)
                assert mine.tobytes() == want.tobytes()


def test_bitmap_aligned_ops(cpu, ref):
    rng = np.random.default_rng(3)
    for nbytes in (1, 7, 8, 31, 32, 33, 1000, 10001):
        l = rng.integers(0, 256, nbytes, dtype=np.uint8)
        r = rng.integers(0, 256, nbytes, dtype=np.uint8)
        for op, name in ((N.BITOP_AND, "and"), (N.BITOP_OR, "or"), (N.BITOP_XOR, "xor"), (N.BITOP_ANDNOT, "and_not")):
            mine = np.zeros(nbytes, dtype=np.uint8)
            assert cpu.ref_bitmap_op(op, ptr(l), 0, ptr(r), 0, ptr(mine), 0, nbytes * 8) == 0
            for isa in ISAS:
                want = np.zeros(nbytes, dtype=np.uint8)
                getattr(ref, f"bitmap_aligned_{name}_{isa}")(ptr(l), ptr(r), ptr(want), nbytes)
                assert mine.tobytes() == want.tobytes()


def test_bitmap_unaligned_ops_naive(cpu):
    # arrow/bitutil/bitmaps_test.go style: every offset combination against a naive bit loop
    rng = np.random.default_rng(4)
    for n in (0, 1, 9, 64, 65, 300):
        for lo in (0, 1, 7, 13):
            for ro in (0, 3, 8):
                for oo in (0, 5, 16):
                    a, b = rng.random(n) < 0.5, rng.random(n) < 0.5
                    l, r = pack_bits(a, lo), pack_bits(b, ro)
                    for op, exp in ((0, a & b), (1, a | b), (2, a ^ b), (3, a & ~b), (4, ~(a ^ b))):
                        out = pack_bits(np.zeros(n, bool), oo, 0xA5)
                        assert cpu.ref_bitmap_op(op, ptr(l), lo, ptr(r), ro, ptr(out), oo, n) == 0
                        assert out.tobytes() == pack_bits(exp, oo, 0xA5).tobytes()
                    assert cpu.ref_bitmap_popcount(ptr(l), lo, n) == int(a.sum())
