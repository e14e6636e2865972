"""Take on the GPU vs the oracle.  Shapes follow arrow/compute/vector_selection_test.go:
TestTakeNumeric (:1127-1141, incl. ErrIndex for 9 and -1), checkTake re-run with index types
int8 / uint32 and sliced inputs (:213-253)."""
import ctypes as C

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import Dev, pack_bits, ptr, unpack_bits

pytestmark = pytest.mark.gpu

VDT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}
IDT = {(8, 1): np.int8, (8, 0): np.uint8, (16, 1): np.int16, (16, 0): np.uint16, (32, 1): np.int32, (32, 0): np.uint32,
       (64, 1): np.int64, (64, 0): np.uint64}


def both(ag, cpu, bw, vals, vvalid, voff, vlen, iw, signed, idx, ivalid, ioff, n, bounds=1):
    dt = VDT[bw]
    want = np.full(n, 0xEE, dtype=dt); wv = np.zeros(n // 8 + 8, dtype=np.uint8)
    wn, wp, wi = C.c_int64(), C.c_int64(), C.c_int64()
    wst = cpu.ref_take_primitive(bw, ptr(vals), ptr(vvalid), voff, vlen, iw, signed, ptr(idx), ptr(ivalid), ioff, n, bounds,
                                 ptr(want), ptr(wv), C.byref(wn), C.byref(wp), C.byref(wi))
    need_valid = vvalid is not None or ivalid is not None
    got = np.full(n, 0xDD, dtype=dt); gv = np.zeros(n // 8 + 8, dtype=np.uint8) if need_valid else None
    gn, gp, gi = C.c_int64(), C.c_int64(), C.c_int64()
    gst, msg = ag.call_status("ag_take_primitive", bw, ptr(vals), ptr(vvalid), voff, vlen, iw, signed, ptr(idx), ptr(ivalid), ioff, n, bounds,
                              ptr(got), ptr(gv), C.byref(gn), C.byref(gp), C.byref(gi))
    assert gst == wst, msg
    if wst != 0:
        assert gst == N.AG_ERR_INDEX and gp.value == wp.value and gi.value == wi.value
        assert msg == f"{wi.value} out of bounds"
        return
    assert got.tobytes() == want.tobytes()
    if need_valid:
        assert np.array_equal(unpack_bits(gv, 0, n), unpack_bits(wv, 0, n))
        assert gn.value == wn.value
    # device flavour
    dv = Dev(vals); di = Dev(idx)
    dvv = Dev(vvalid, byte_offset=1) if vvalid is not None else None
    div = Dev(ivalid, byte_offset=3) if ivalid is not None else None
    dout = Dev(np.full(n, 0xCC, dtype=dt))
    dov = Dev(np.full((n + 31) // 32 * 4, 0xFF, dtype=np.uint8)) if need_valid else None
    dbad = Dev(np.zeros(1, dtype=np.int64))
    ag.call("ag_error_word_reset_dev", dbad.ptr, None)
    ag.call("ag_take_primitive_dev", bw, dv.ptr, dvv.ptr if dvv else None, voff, vlen, iw, signed, di.ptr, div.ptr if div else None, ioff, n,
            bounds, dout.ptr, dov.ptr if dov else None, dbad.ptr, None)
    ag.call("ag_stream_sync", None)
    assert dbad.get()[0] == N.NO_ERROR_POS
    assert dout.get().tobytes() == want.tobytes()
    if need_valid:
        assert np.array_equal(unpack_bits(dov.get(), 0, n), unpack_bits(wv, 0, n))


def test_reference_literal_cases(ag, cpu):
    # TestTakeNumeric: values [7,8,9]; indices [], [0,1,0], [null,1,0], [3,0,...] errors
    vals = np.array([7, 8, 9], dtype=np.int32).view(np.uint32)
    for idx in ([0, 1, 0], [2, 1, 0], [0, 0, 0, 0, 1, 2, 2]):
        i = np.array(idx, dtype=np.int32)
        both(ag, cpu, 32, vals, None, 0, 3, 32, 1, i, None, 0, i.size)
    i = np.array([0, 1, 0], dtype=np.int32)
    both(ag, cpu, 32, vals, None, 0, 3, 32, 1, i, pack_bits([0, 1, 1]), 0, 3)          # [null, 1, 0]
    both(ag, cpu, 32, vals, pack_bits([0, 1, 1]), 0, 3, 32, 1, i, None, 0, 3)          # values [null, 8, 9]
    # out of bounds: index 9 -> ErrIndex; index -1 -> ErrIndex (:1139-1140)
    both(ag, cpu, 32, vals, None, 0, 3, 32, 1, np.array([0, 9, 0], dtype=np.int32), None, 0, 3)
    both(ag, cpu, 32, vals, None, 0, 3, 32, 1, np.array([0, -1, 0], dtype=np.int32), None, 0, 3)
    # a null slot holding an out-of-range index is not an error (helpers.go:942)
    both(ag, cpu, 32, vals, None, 0, 3, 32, 1, np.array([0, 9, 0], dtype=np.int32), pack_bits([1, 0, 1]), 0, 3)
    # empty
    out = np.zeros(1, dtype=np.uint32)
    no_idx = np.zeros(1, dtype=np.int32)
    ag.call("ag_take_primitive", 32, ptr(vals), None, 0, 3, 32, 1, ptr(no_idx), None, 0, 0, 1, ptr(out), None, None, None, None)


@pytest.mark.parametrize("bw", [8, 16, 32, 64])
@pytest.mark.parametrize("iw,signed", [(8, 1), (8, 0), (16, 1), (32, 1), (32, 0), (64, 1), (64, 0)])
def test_random_differential(ag, cpu, bw, iw, signed):
    rng = np.random.default_rng(bw * 100 + iw * 2 + signed)
    idt = IDT[(iw, signed)]
    for vlen in (1, 100, 127 if iw == 8 else 50_000):
        vlen = min(vlen, int(np.iinfo(idt).max))
        for n in (1, 31, 32, 33, 1000, 4097, 100_001):
            for p_inull, p_vnull in ((0, 0), (0.2, 0), (0, 0.2), (0.2, 0.2)):
                voff, ioff = int(rng.integers(0, 9)), int(rng.integers(0, 11))
                vals = rng.integers(0, np.iinfo(VDT[bw]).max, vlen + voff, dtype=VDT[bw], endpoint=True)
                idx = rng.integers(0, vlen, n).astype(idt)
                vvalid = pack_bits(rng.random(vlen) >= p_vnull, voff) if p_vnull else None
                ivalid = pack_bits(rng.random(n) >= p_inull, ioff) if p_inull else None
                both(ag, cpu, bw, vals, vvalid, voff, vlen, iw, signed, idx, ivalid, ioff, n)


def test_first_offender_is_lowest_row(ag, cpu):
    rng = np.random.default_rng(5)
    n, vlen = 500_000, 1000
    vals = rng.integers(0, 1 << 60, vlen, dtype=np.uint64)
    idx = rng.integers(0, vlen, n).astype(np.int32)
    idx[[400_000, 123_457, 499_999]] = [5000, -3, 1000]
    both(ag, cpu, 64, vals, None, 0, vlen, 32, 1, idx, None, 0, n)
    # bounds_check off: out-of-range slots must not fault (we skip the load)
    out = np.zeros(n, dtype=np.uint64)
    ag.call("ag_take_primitive", 64, ptr(vals), None, 0, vlen, 32, 1, ptr(idx), None, 0, n, 0, ptr(out), None, None, None, None)
    ok = (idx >= 0) & (idx < vlen)
    assert np.array_equal(out[ok], vals[idx[ok]])


def test_take_125m_rows_config4_shard(ag):
    """One GPU's shard of BASELINE config 4 (125M int32 indices per GPU).  The values table is
    256M rows (2 GB, well past the 126 MB L2) of splitmix64(seed + i), generated on the device, so
    out[i] must equal splitmix64(seed + idx[i]): checked on three 256K-row windows against the
    generator's CPU twin.  A planted out-of-range index must be reported at its row."""
    nv, n = 1 << 28, 125_000_000
    v = Dev(nbytes=nv * 8)
    ag.call("ag_generate_dev", 0, 0x94378165, 0, 0, v.ptr, nv, None)
    idx = Dev(nbytes=n * 4)
    ag.call("ag_generate_dev", 2, 0x0FF1CE, 0, nv - 1, idx.ptr, n, None)
    out = Dev(nbytes=n * 8)
    bad = Dev(np.zeros(1, dtype=np.int64))
    ag.call("ag_error_word_reset_dev", bad.ptr, None)
    ag.call("ag_take_primitive_dev", 64, v.ptr, None, 0, nv, 32, 1, idx.ptr, None, 0, n, 1, out.ptr, None, bad.ptr, None)
    ag.call("ag_stream_sync", None)
    assert bad.get()[0] == N.NO_ERROR_POS
    from oracle import oracle
    cpu = oracle.cpu()
    for start in (0, 64_000_000, n - (1 << 18)):
        w = 1 << 18
        hi = idx.buf.to_numpy(np.int32, w, start * 4)
        want = np.empty(w, dtype=np.uint64)
        # values[i] = mix64(seed + i): evaluate the generator's CPU twin at the gathered positions
        tmp = np.empty(1, dtype=np.uint64)
        pos = hi.astype(np.uint64)
        z = (np.uint64(0x94378165) + pos)
        with np.errstate(over="ignore"):
            z = z + np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            want = z ^ (z >> np.uint64(31))
        cpu.ref_generate(0, 0x94378165 + int(pos[0]), 0, 0, tmp.ctypes.data, 1)
        assert tmp[0] == want[0]
        assert out.buf.to_numpy(np.uint64, w, start * 8).tobytes() == want.tobytes()
    # planted out-of-range index
    h = np.array([nv], dtype=np.int32)
    ag.call("ag_upload", idx.ptr + 4 * 77_777_777, h.ctypes.data, 4, None)
    ag.call("ag_take_primitive_dev", 64, v.ptr, None, 0, nv, 32, 1, idx.ptr, None, 0, n, 1, out.ptr, None, bad.ptr, None)
    ag.call("ag_stream_sync", None)
    assert bad.get()[0] == 77_777_777


def test_boolean_values_take(ag, cpu):
    rng = np.random.default_rng(22)
    for vlen in (1, 77, 10_000):
        for n in (1, 31, 32, 33, 1000, 100_001):
            for p_inull, p_vnull in ((0, 0), (0.2, 0.2)):
                voff, ioff = int(rng.integers(0, 13)), int(rng.integers(0, 9))
                vals = pack_bits(rng.random(vlen) < 0.5, voff)
                vvalid = pack_bits(rng.random(vlen) >= p_vnull, voff) if p_vnull else None
                idx = rng.integers(0, vlen, n).astype(np.int32)
                ivalid = pack_bits(rng.random(n) >= p_inull, ioff) if p_inull else None
                want = np.zeros(n // 8 + 8, dtype=np.uint8); wv = np.zeros(n // 8 + 8, dtype=np.uint8)
                wn, wp, wi = C.c_int64(), C.c_int64(), C.c_int64()
                assert cpu.ref_take_primitive(1, ptr(vals), ptr(vvalid), voff, vlen, 32, 1, ptr(idx), ptr(ivalid), ioff, n, 1, ptr(want), ptr(wv),
                                              C.byref(wn), C.byref(wp), C.byref(wi)) == 0
                got = np.zeros(n // 8 + 8, dtype=np.uint8); gv = np.zeros(n // 8 + 8, dtype=np.uint8) if (p_inull or p_vnull) else None
                gn, gp, gi = C.c_int64(), C.c_int64(), C.c_int64()
                ag.call("ag_take_primitive", 1, ptr(vals), ptr(vvalid), voff, vlen, 32, 1, ptr(idx), ptr(ivalid), ioff, n, 1, ptr(got), ptr(gv),
                        C.byref(gn), C.byref(gp), C.byref(gi))
                assert np.array_equal(unpack_bits(got, 0, n), unpack_bits(want, 0, n)), (vlen, n, p_inull)
                if gv is not None:
                    assert np.array_equal(unpack_bits(gv, 0, n), unpack_bits(wv, 0, n)) and gn.value == wn.value


# ---------------------------------------------------------------------------------------------------------
# Windowed path (partition by table window -> gather out of L2 -> un-permute).  ag_take_set_policy(2, ...) forces it
# on inputs the oracle handles in milliseconds; a 4 KB window turns a 50K-row table into ~100 buckets.
@pytest.fixture
def windowed(ag):
    ag.call("ag_take_set_policy", 2, 0, 0, 4096)
    yield ag
    ag.call("ag_take_set_policy", 0, 32 << 20, 160 << 20, 16 << 20)


@pytest.mark.parametrize("bw", [8, 16, 32, 64])
@pytest.mark.parametrize("iw,signed", [(32, 1), (32, 0), (64, 1), (64, 0)])
def test_windowed_random_differential(windowed, cpu, bw, iw, signed):
    ag = windowed
    rng = np.random.default_rng(bw * 1000 + iw * 2 + signed)
    idt = IDT[(iw, signed)]
    for vlen in (1, 2, 511, 50_000, 700_000):     # 700K rows of int64 over 4 KB windows would be 1367 windows (> 1022): the window grows
        for n in (2, 33, 8191, 8192, 8193, 30_000, 100_001):
            for p_inull in (0, 0.3):
                voff, ioff = int(rng.integers(0, 9)), int(rng.integers(0, 11))
                vals = rng.integers(0, np.iinfo(VDT[bw]).max, vlen + voff, dtype=VDT[bw], endpoint=True)
                idx = rng.integers(0, vlen, n).astype(idt)
                ivalid = pack_bits(rng.random(n) >= p_inull, ioff) if p_inull else None
                both(ag, cpu, bw, vals, None, voff, vlen, iw, signed, idx, ivalid, ioff, n)


def test_windowed_bounds_and_nulls(windowed, cpu):
    ag = windowed
    rng = np.random.default_rng(77)
    n, vlen = 200_000, 60_000
    vals = rng.integers(0, 1 << 60, vlen, dtype=np.uint64)
    idx = rng.integers(0, vlen, n).astype(np.int32)
    idx[[150_000, 123_457, 199_999]] = [vlen, -3, 2_000_000_000]
    both(ag, cpu, 64, vals, None, 0, vlen, 32, 1, idx, None, 0, n)                      # first offender = lowest row
    iv = np.ones(n, dtype=bool); iv[[150_000, 123_457, 199_999]] = False
    both(ag, cpu, 64, vals, None, 0, vlen, 32, 1, idx, pack_bits(iv, 5), 5, n)          # offenders under null slots: no error
    iv[123_457] = True
    both(ag, cpu, 64, vals, None, 0, vlen, 32, 1, idx, pack_bits(iv, 5), 5, n)
    # bounds_check off: out-of-range slots gather nothing and read as 0
    out = np.full(n, 0xEE, dtype=np.uint64)
    ag.call("ag_take_primitive", 64, ptr(vals), None, 0, vlen, 32, 1, ptr(idx), None, 0, n, 0, ptr(out), None, None, None, None)
    ok = (idx >= 0) & (idx < vlen)
    assert np.array_equal(out[ok], vals[idx[ok]]) and not out[~ok].any()
    # all-null indices
    both(ag, cpu, 64, vals, None, 0, vlen, 32, 1, idx, pack_bits(np.zeros(n, dtype=bool)), 0, n)
    # every index the same row / all in the last window
    both(ag, cpu, 64, vals, None, 0, vlen, 32, 1, np.full(n, vlen - 1, dtype=np.int32), None, 0, n)


def _mix64(z):
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


@pytest.mark.parametrize("order", ["random", "sorted", "reversed", "planted_bad"])
def test_auto_routing_40m_rows(ag, order):
    """SURVEY §8d C4 datasets at a size the automatic policy sends to the windowed path (40M int32 indices into a
    64M-row, 512 MB table): random -> windowed, sorted / reversed -> the probe keeps the direct kernel, and one planted
    out-of-range index is reported at its row.  Every output row is checked: values[i] = splitmix64(seed + i)."""
    nv, n = 1 << 26, 40_000_000
    v = Dev(nbytes=nv * 8)
    ag.call("ag_generate_dev", 0, 0x94378165, 0, 0, v.ptr, nv, None)
    idx = Dev(nbytes=n * 4)
    ag.call("ag_generate_dev", 2, 0x0FF1CE, 0, nv - 1, idx.ptr, n, None)
    ag.call("ag_stream_sync", None)
    hidx = idx.buf.to_numpy(np.int32, n)
    if order in ("sorted", "reversed"):
        hidx = np.sort(hidx)
        if order == "reversed":
            hidx = hidx[::-1].copy()
        ag.call("ag_upload", idx.ptr, hidx.ctypes.data, n * 4, None)
    if order == "planted_bad":
        hidx[31_234_567] = -1
        ag.call("ag_upload", idx.ptr, hidx.ctypes.data, n * 4, None)
    out = Dev(nbytes=n * 8)
    bad = Dev(np.zeros(1, dtype=np.int64))
    ag.call("ag_error_word_reset_dev", bad.ptr, None)
    ag.call("ag_take_primitive_dev", 64, v.ptr, None, 0, nv, 32, 1, idx.ptr, None, 0, n, 1, out.ptr, None, bad.ptr, None)
    ag.call("ag_stream_sync", None)
    if order == "planted_bad":
        assert bad.get()[0] == 31_234_567
        return
    assert bad.get()[0] == N.NO_ERROR_POS
    want = _mix64(np.uint64(0x94378165) + hidx.astype(np.uint64))
    assert out.buf.to_numpy(np.uint64, n).tobytes() == want.tobytes()
