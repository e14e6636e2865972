"""Bitmap kernels (validity propagation, boolean kernels) on the GPU vs the oracle; aligned
cases also against the reference's bitmap_aligned_* SIMD loops.  Shapes follow
arrow/bitutil/bitmaps_test.go (offset sweeps) and arrow/compute/scalar_bool_test.go."""
import ctypes as C

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import Dev, pack_bits, ptr, unpack_bits

pytestmark = pytest.mark.gpu

NAMES = {N.BITOP_AND: "and", N.BITOP_OR: "or", N.BITOP_XOR: "xor", N.BITOP_ANDNOT: "and_not"}


def test_aligned_ops_match_reference_simd(ag, ref, isa):
    rng = np.random.default_rng(1)
    for nbytes in (1, 7, 8, 31, 32, 33, 1000, 100001):
        l = rng.integers(0, 256, nbytes, dtype=np.uint8)
        r = rng.integers(0, 256, nbytes, dtype=np.uint8)
        for op, name in NAMES.items():
            want = np.zeros(nbytes, dtype=np.uint8)
            getattr(ref, f"bitmap_aligned_{name}_{isa}")(ptr(l), ptr(r), ptr(want), nbytes)
            got = np.zeros(nbytes, dtype=np.uint8)
            ag.call("ag_bitmap_op", op, ptr(l), 0, ptr(r), 0, ptr(got), 0, nbytes * 8)
            assert got.tobytes() == want.tobytes(), (name, nbytes)


def test_binary_ops_all_offsets(ag, cpu):
    rng = np.random.default_rng(2)
    for n in (0, 1, 5, 8, 31, 32, 33, 63, 64, 65, 200, 1025, 70001):
        for trial in range(6):
            lo, ro, oo = (int(x) for x in rng.integers(0, 40, 3))
            lb, rb = rng.random(n) < 0.5, rng.random(n) < 0.5
            l, r = pack_bits(lb, lo, 0x3C), pack_bits(rb, ro, 0xC3)
            for op in range(5):
                want = pack_bits(np.zeros(n, bool), oo, 0xA5)
                assert cpu.ref_bitmap_op(op, ptr(l), lo, ptr(r), ro, ptr(want), oo, n) == 0
                got = pack_bits(np.zeros(n, bool), oo, 0xA5)
                ag.call("ag_bitmap_op", op, ptr(l), lo, ptr(r), ro, ptr(got), oo, n)
                assert got.tobytes() == want.tobytes(), (n, lo, ro, oo, op)
                # device flavour with odd byte phases
                dl, dr = Dev(l, byte_offset=1), Dev(r, byte_offset=2)
                do = Dev(pack_bits(np.zeros(n, bool), oo, 0xA5), byte_offset=3)
                ag.call("ag_bitmap_op_dev", op, dl.ptr, lo, dr.ptr, ro, do.ptr, oo, n, None)
                ag.call("ag_stream_sync", None)
                assert do.get().tobytes() == want.tobytes(), ("dev", n, lo, ro, oo, op)


def test_in_place_accumulate(ag, cpu):
    # propagateNulls ANDs further validity bitmaps into the output in place (executor.go:340-347)
    rng = np.random.default_rng(3)
    n, oo, ro = 5000, 5, 11
    a, b = rng.random(n) < 0.7, rng.random(n) < 0.7
    out = Dev(pack_bits(a, oo, 0xA5))
    r = Dev(pack_bits(b, ro, 0x11))
    ag.call("ag_bitmap_op_dev", N.BITOP_AND, out.ptr, oo, r.ptr, ro, out.ptr, oo, n, None)
    ag.call("ag_stream_sync", None)
    assert out.get().tobytes() == pack_bits(a & b, oo, 0xA5).tobytes()


def test_copy_invert_set_popcount(ag, cpu):
    rng = np.random.default_rng(4)
    for n in (0, 1, 7, 8, 9, 64, 65, 1000, 65537):
        for _ in range(5):
            so, do_ = (int(x) for x in rng.integers(0, 50, 2))
            bits = rng.random(n) < 0.4
            src = pack_bits(bits, so, 0x0F)
            for name, want_bits in (("ag_bitmap_copy", bits), ("ag_bitmap_invert", ~bits)):
                got = pack_bits(np.zeros(n, bool), do_, 0xA5)
                ag.call(name, ptr(src), so, n, ptr(got), do_)
                assert got.tobytes() == pack_bits(want_bits, do_, 0xA5).tobytes(), (name, n, so, do_)
                dsrc = Dev(src, byte_offset=1)
                ddst = Dev(pack_bits(np.zeros(n, bool), do_, 0xA5), byte_offset=2)
                ag.call(name + "_dev", dsrc.ptr, so, n, ddst.ptr, do_, None)
                ag.call("ag_stream_sync", None)
                assert ddst.get().tobytes() == pack_bits(want_bits, do_, 0xA5).tobytes()
            for val in (0, 1):
                got = pack_bits(bits, do_, 0xA5)
                ag.call("ag_bitmap_set", ptr(got), do_, n, val)
                assert got.tobytes() == pack_bits(np.full(n, bool(val)), do_, 0xA5).tobytes()
            c = C.c_int64(-1)
            ag.call("ag_bitmap_popcount", ptr(src), so, n, C.byref(c))
            assert c.value == int(bits.sum()) == cpu.ref_bitmap_popcount(ptr(src), so, n)


def test_kleene_truth_tables(ag, cpu):
    """scalar_bool_test.go TestBooleanKernels: Kleene and / or / and_not over {true,false,null}^2."""
    T, F, NUL = (1, 1), (1, 0), (0, 0)  # (valid, data)
    combos = [(a, b) for a in (T, F, NUL) for b in (T, F, NUL)]
    lv = np.array([a[0] for a, _ in combos], bool); ld = np.array([a[1] for a, _ in combos], bool)
    rv = np.array([b[0] for _, b in combos], bool); rd = np.array([b[1] for _, b in combos], bool)
    expect = {
        # and_kleene: false dominates
        N.KLEENE_AND: [T, F, NUL, F, F, F, NUL, F, NUL],
        N.KLEENE_OR: [T, T, T, T, F, NUL, T, NUL, NUL],
        N.KLEENE_ANDNOT: [F, T, NUL, F, F, F, F, NUL, NUL],
    }
    blv, bld, brv, brd = pack_bits(lv), pack_bits(ld), pack_bits(rv), pack_bits(rd)  # keep the buffers alive across the call
    for kop, exp in expect.items():
        ov = np.zeros(2, dtype=np.uint8); od = np.zeros(2, dtype=np.uint8)
        ag.call("ag_kleene", kop, ptr(blv), ptr(bld), 0, ptr(brv), ptr(brd), 0, ptr(ov), ptr(od), 0, 9)
        gv, gd = unpack_bits(ov, 0, 9), unpack_bits(od, 0, 9)
        for i, (v, d) in enumerate(exp):
            assert bool(gv[i]) == bool(v), (kop, i)
            if v:
                assert bool(gd[i]) == bool(d), (kop, i)


def test_kleene_random_vs_oracle(ag, cpu):
    rng = np.random.default_rng(5)
    for n in (1, 33, 64, 1000, 40001):
        for kop in range(3):
            for lnull, rnull in ((True, True), (False, True), (True, False)):
                lo, ro, oo = (int(x) for x in rng.integers(0, 30, 3))
                lv = pack_bits(rng.random(n) < 0.8, lo) if lnull else None
                rv = pack_bits(rng.random(n) < 0.8, ro) if rnull else None
                ld, rd = pack_bits(rng.random(n) < 0.5, lo), pack_bits(rng.random(n) < 0.5, ro)
                wv, wd = pack_bits(np.zeros(n, bool), oo, 0xA5), pack_bits(np.zeros(n, bool), oo, 0x5A)
                assert cpu.ref_kleene(kop, ptr(lv), ptr(ld), lo, ptr(rv), ptr(rd), ro, ptr(wv), ptr(wd), oo, n) == 0
                gv, gd = pack_bits(np.zeros(n, bool), oo, 0xA5), pack_bits(np.zeros(n, bool), oo, 0x5A)
                ag.call("ag_kleene", kop, ptr(lv), ptr(ld), lo, ptr(rv), ptr(rd), ro, ptr(gv), ptr(gd), oo, n)
                assert gv.tobytes() == wv.tobytes() and gd.tobytes() == wd.tobytes(), (n, kop, lnull, rnull)
