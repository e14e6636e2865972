"""Filter / GetTakeIndices on the GPU vs the oracle (which restates primitiveFilterImpl and is
itself pinned against pyarrow and the reference's literal cases in the CPU suite).

Shapes follow arrow/compute/vector_selection_test.go: TestFilterNumeric literal cases (:449-484),
every case re-run on sliced inputs (values padded by 3, mask by 2, :114-145), random
differential over lengths 8..512 with seed 0x0ff1ce (:554-613), both null-selection modes."""
import ctypes as C

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import Dev, NP_OF, pack_bits, ptr, unpack_bits

pytestmark = pytest.mark.gpu

WIDTH_DT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}


def oracle_filter(cpu, bw, vals, vvalid, voff, mask, mvalid, moff, n, sel):
    out = np.zeros(n + 1, dtype=WIDTH_DT[bw])
    ov = np.zeros((n + 7) // 8 + 8, dtype=np.uint8)
    ln, nulls = C.c_int64(), C.c_int64()
    assert cpu.ref_filter_primitive(bw, ptr(vals), ptr(vvalid), voff, ptr(mask), ptr(mvalid), moff, n, sel, ptr(out), ptr(ov), C.byref(ln), C.byref(nulls)) == 0
    return out[: ln.value], ov, ln.value, nulls.value


def run_case(ag, cpu, rng, bw, n, sel, p_mask, p_mnull, p_vnull, voff, moff):
    dt = WIDTH_DT[bw]
    vals = rng.integers(0, np.iinfo(dt).max, n + voff, dtype=dt, endpoint=True)
    vvalid = pack_bits(rng.random(n) >= p_vnull, voff) if p_vnull else None
    mask = pack_bits(rng.random(n) < p_mask, moff, 0x5A)
    mvalid = pack_bits(rng.random(n) >= p_mnull, moff, 0xC3) if p_mnull else None
    want, wv, wlen, wnulls = oracle_filter(cpu, bw, vals, vvalid, voff, mask, mvalid, moff, n, sel)
    # host flavour
    cnt = C.c_int64(-1)
    ag.call("ag_filter_output_size", ptr(mask), ptr(mvalid), moff, n, sel, C.byref(cnt))
    assert cnt.value == wlen == cpu.ref_filter_output_size(ptr(mask), ptr(mvalid), moff, n, sel)
    got = np.zeros(wlen + 1, dtype=dt)
    need_valid = vvalid is not None or mvalid is not None
    gv = np.zeros((wlen + 7) // 8 + 8, dtype=np.uint8) if need_valid else None
    glen, gnulls = C.c_int64(), C.c_int64()
    ag.call("ag_filter_primitive", bw, ptr(vals), ptr(vvalid), voff, ptr(mask), ptr(mvalid), moff, n, sel, ptr(got), ptr(gv), C.byref(glen), C.byref(gnulls))
    assert glen.value == wlen
    assert got[:wlen].tobytes() == want.tobytes(), (bw, n, sel, p_mask, p_mnull, p_vnull, voff, moff)
    if need_valid:
        assert np.array_equal(unpack_bits(gv, 0, wlen), unpack_bits(wv, 0, wlen))
        assert gnulls.value == wnulls
    # device flavour with an upper-bound capacity (n) instead of the exact size
    dv = Dev(vals, byte_offset=0)
    dvv = Dev(vvalid, byte_offset=1) if vvalid is not None else None
    dm = Dev(mask, byte_offset=3)
    dmv = Dev(mvalid, byte_offset=2) if mvalid is not None else None
    dout = Dev(np.zeros(n + 1, dtype=dt))
    dov = Dev(np.full((n + 31) // 32 * 4 + 8, 0xFF, dtype=np.uint8)) if need_valid else None
    dlen = Dev(np.array([-1], dtype=np.int64))
    ag.call("ag_filter_primitive_dev", bw, dv.ptr, dvv.ptr if dvv else None, voff, dm.ptr, dmv.ptr if dmv else None, moff, n, sel,
            dout.ptr, dov.ptr if dov else None, n, dlen.ptr, None)
    ag.call("ag_stream_sync", None)
    assert dlen.get()[0] == wlen
    assert dout.get()[:wlen].tobytes() == want.tobytes()
    if need_valid:
        assert np.array_equal(unpack_bits(dov.get(), 0, wlen), unpack_bits(wv, 0, wlen))


def test_reference_literal_cases(ag, cpu):
    """TestFilterNumeric, vector_selection_test.go:449-484 (values [7,8,9] etc.)."""
    def check(values, vnull, mask, mnull, sel, exp_vals, exp_valid):
        vals = np.array(values, dtype=np.int32).view(np.uint32)
        n = len(values)
        vvalid = pack_bits(~np.array(vnull, bool)) if any(vnull) else None
        m = pack_bits(np.array(mask, bool))
        mvalid = pack_bits(~np.array(mnull, bool)) if any(mnull) else None
        out = np.zeros(n + 1, dtype=np.uint32)
        ov = np.zeros(8, dtype=np.uint8)
        ln, nulls = C.c_int64(), C.c_int64()
        ag.call("ag_filter_primitive", 32, ptr(vals), ptr(vvalid), 0, ptr(m), ptr(mvalid), 0, n, sel, ptr(out), ptr(ov), C.byref(ln), C.byref(nulls))
        assert ln.value == len(exp_vals)
        valid = unpack_bits(ov, 0, ln.value) if (vvalid is not None or mvalid is not None) else np.ones(ln.value, bool)
        assert valid.tolist() == [bool(x) for x in exp_valid]
        for i, (v, ok) in enumerate(zip(exp_vals, exp_valid)):
            if ok:
                assert int(out[i:i + 1].view(np.int32)[0]) == v
    z3 = [0, 0, 0]
    # ValidateFilter("[7, 8, 9]", "[0, 1, 0]", "[8]") ...
    check([7, 8, 9], z3, [0, 1, 0], z3, N.DROP_NULLS, [8], [1])
    check([7, 8, 9], z3, [0, 0, 0], z3, N.DROP_NULLS, [], [])
    check([7, 8, 9], z3, [1, 0, 1], z3, N.DROP_NULLS, [7, 9], [1, 1])
    check([7, 8, 9], z3, [1, 1, 1], z3, N.DROP_NULLS, [7, 8, 9], [1, 1, 1])
    # values with a null: "[null, 8, 9]" mask [0,1,0] / [1,0,1]
    check([0, 8, 9], [1, 0, 0], [0, 1, 0], z3, N.DROP_NULLS, [8], [1])
    check([0, 8, 9], [1, 0, 0], [1, 0, 1], z3, N.DROP_NULLS, [0, 9], [0, 1])
    # mask with a null: "[7, 8, 9]" x "[null, 1, 0]" -> EmitNulls "[null, 8]", DropNulls "[8]"
    check([7, 8, 9], z3, [0, 1, 0], [1, 0, 0], N.EMIT_NULLS, [0, 8], [0, 1])
    check([7, 8, 9], z3, [0, 1, 0], [1, 0, 0], N.DROP_NULLS, [8], [1])
    check([7, 8, 9], z3, [1, 1, 0], [0, 1, 0], N.EMIT_NULLS, [7, 0], [1, 0])
    check([7, 8, 9], z3, [1, 1, 0], [0, 1, 0], N.DROP_NULLS, [7], [1])


@pytest.mark.parametrize("bw", [8, 16, 32, 64])
def test_random_differential(ag, cpu, bw):
    rng = np.random.default_rng(0x0FF1CE + bw)
    for n in (1, 8, 31, 32, 33, 64, 100, 512, 1000, 8191, 8192, 8193, 40000, 300001):
        for sel in (N.DROP_NULLS, N.EMIT_NULLS):
            for p_mask in (0.0, 0.1, 0.5, 1.0):
                for p_mnull, p_vnull in ((0, 0), (0.1, 0), (0, 0.2), (0.3, 0.3)):
                    if n > 50000 and (p_mask in (0.0, 1.0)) and p_mnull:
                        continue
                    run_case(ag, cpu, rng, bw, n, sel, p_mask, p_mnull, p_vnull, int(rng.integers(0, 9)), int(rng.integers(0, 19)))


def test_long_runs_and_block_edges(ag, cpu):
    """Clustered masks (sorted values) exercise the all-set / none-set block paths of
    primitiveFilterImpl (vector_selection.go:303-320)."""
    rng = np.random.default_rng(11)
    n = 100_000
    mask_bits = np.zeros(n, bool)
    mask_bits[1000:5000] = True
    mask_bits[64 * 300:64 * 301] = True
    mask_bits[-70:] = True
    vals = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    for moff in (0, 5):
        mask = pack_bits(mask_bits, moff)
        want, _, wlen, _ = oracle_filter(cpu, 64, vals, None, 0, mask, None, moff, n, 0)
        got = np.zeros(wlen, dtype=np.uint64)
        ln = C.c_int64()
        ag.call("ag_filter_primitive", 64, ptr(vals), None, 0, ptr(mask), None, moff, n, 0, ptr(got), None, C.byref(ln), None)
        assert ln.value == wlen and got.tobytes() == want.tobytes()


def test_take_indices(ag, cpu):
    rng = np.random.default_rng(12)
    for n in (1, 33, 1000, 65534, 65535, 200_001):
        for sel in (0, 1):
            for p_mnull in (0, 0.2):
                moff = int(rng.integers(0, 13))
                mask = pack_bits(rng.random(n) < 0.3, moff)
                mvalid = pack_bits(rng.random(n) >= p_mnull, moff) if p_mnull else None
                iw = 16 if n < 65535 else 32
                want = np.zeros(n + 1, dtype=WIDTH_DT[iw]); wv = np.zeros(n // 8 + 8, dtype=np.uint8); wl = C.c_int64()
                assert cpu.ref_take_indices(iw, ptr(mask), ptr(mvalid), moff, n, sel, ptr(want), ptr(wv), C.byref(wl)) == 0
                got = np.zeros(n + 1, dtype=WIDTH_DT[iw]); gv = np.zeros(n // 8 + 8, dtype=np.uint8); gl = C.c_int64()
                ag.call("ag_take_indices", iw, ptr(mask), ptr(mvalid), moff, n, sel, ptr(got), ptr(gv) if p_mnull else None, C.byref(gl))
                assert gl.value == wl.value
                assert got[: wl.value].tobytes() == want[: wl.value].tobytes()
                if p_mnull:
                    assert np.array_equal(unpack_bits(gv, 0, wl.value), unpack_bits(wv, 0, wl.value))


@pytest.mark.parametrize("type_id,cmp", [(N.INT64, N.CMP_GT), (N.INT64, N.CMP_LE), (N.FLOAT64, N.CMP_GE), (N.INT32, N.CMP_EQ), (N.UINT64, N.CMP_NE), (N.FLOAT32, N.CMP_LT)])
def test_fused_compare_filter_equals_two_step(ag, cpu, type_id, cmp):
    rng = np.random.default_rng(type_id * 10 + cmp)
    dt = NP_OF[type_id]
    for n in (1, 31, 1024, 8192, 8193, 100_003, 1_000_000):
        vals = rng.integers(0, 100, n).astype(dt)
        sc = np.array([89], dtype=dt)
        mask = np.zeros((n + 7) // 8 + 1, dtype=np.uint8)
        assert cpu.ref_compare(type_id, cmp, N.SHAPE_AS, ptr(vals), ptr(sc), ptr(mask), n, 0) == 0
        bw = np.dtype(dt).itemsize * 8
        want, _, wlen, _ = oracle_filter(cpu, bw, vals.view(WIDTH_DT[bw]), None, 0, mask, None, 0, n, 0)
        dv = Dev(vals)
        dout = Dev(np.zeros(max(wlen, 1), dtype=dt))
        dlen = Dev(np.array([-1], dtype=np.int64))
        ag.call("ag_filter_compare_scalar_dev", type_id, cmp, dv.ptr, ptr(sc), n, dout.ptr, wlen, dlen.ptr, None)
        ag.call("ag_stream_sync", None)
        assert dlen.get()[0] == wlen
        assert dout.get()[:wlen].view(WIDTH_DT[bw]).tobytes() == want.tobytes()


def test_filter_100m_rows_config3(ag):
    """BASELINE config 3 at full size: Greater(int64[100M] in [0,100), 89) + Filter, ~10 % selected.
    Properties: out_len == popcount(mask); every output value > 89 (checked with a second compare
    + popcount on the device); stable order checked through the order-sensitive checksum against
    the fused kernel's output; a 1M-row prefix window equals the oracle bit for bit."""
    n = 100_000_000
    v = Dev(nbytes=n * 8)
    ag.call("ag_generate_dev", 1, 0x0FF1CE, 0, 99, v.ptr, n, None)
    sc = np.array([89], dtype=np.int64)
    mask = Dev(nbytes=(n + 7) // 8)
    ag.call("ag_compare_dev", N.INT64, N.CMP_GT, N.SHAPE_AS, v.ptr, ptr(sc), mask.ptr, n, 0, None)
    sc_ = Dev(np.zeros(8, dtype=np.int64))
    ag.call("ag_filter_output_size_dev", mask.ptr, None, 0, n, 0, sc_.ptr, None)
    ag.call("ag_stream_sync", None)
    cnt = int(sc_.get()[0])
    assert abs(cnt / n - 0.1) < 0.001
    out = Dev(nbytes=cnt * 8)
    ag.call("ag_filter_primitive_dev", 64, v.ptr, None, 0, mask.ptr, None, 0, n, 0, out.ptr, None, cnt, sc_.ptr + 8, None)
    out2 = Dev(nbytes=cnt * 8)
    ag.call("ag_filter_compare_scalar_dev", N.INT64, N.CMP_GT, v.ptr, ptr(sc), n, out2.ptr, cnt, sc_.ptr + 16, None)
    cs = Dev(np.zeros(2, dtype=np.uint64))
    ag.call("ag_checksum64_dev", out.ptr, cnt, cs.ptr, None)
    ag.call("ag_checksum64_dev", out2.ptr, cnt, cs.ptr + 8, None)
    omask = Dev(nbytes=(cnt + 7) // 8 + 8)
    ag.call("ag_compare_dev", N.INT64, N.CMP_GT, N.SHAPE_AS, out.ptr, ptr(sc), omask.ptr, cnt, 0, None)
    ag.call("ag_bitmap_popcount_dev", omask.ptr, 0, cnt, sc_.ptr + 24, None)
    ag.call("ag_stream_sync", None)
    s = sc_.get()
    assert s[1] == cnt and s[2] == cnt and s[3] == cnt
    c = cs.get()
    assert c[0] == c[1]
    from oracle import oracle
    w = 1 << 20
    hv = v.buf.to_numpy(np.int64, w)
    hm = mask.buf.to_numpy(np.uint8, w // 8)
    want = np.zeros(w, dtype=np.int64); wl = C.c_int64()
    assert oracle.cpu().ref_filter_primitive(64, ptr(hv), None, 0, ptr(hm), None, 0, w, 0, ptr(want), None, C.byref(wl), None) == 0
    assert out.buf.to_numpy(np.int64, wl.value).tobytes() == want[: wl.value].tobytes()


def test_boolean_values_filter(ag, cpu):
    rng = np.random.default_rng(21)
    for n in (1, 31, 32, 33, 1000, 32768, 32769, 200_001):
        for sel in (0, 1):
            for p_mask in (0.05, 0.5, 1.0):
                for p_mnull, p_vnull in ((0, 0), (0.2, 0.2)):
                    voff, moff = int(rng.integers(0, 13)), int(rng.integers(0, 9))
                    vals = pack_bits(rng.random(n) < 0.5, voff)
                    vvalid = pack_bits(rng.random(n) >= p_vnull, voff) if p_vnull else None
                    mask = pack_bits(rng.random(n) < p_mask, moff)
                    mvalid = pack_bits(rng.random(n) >= p_mnull, moff) if p_mnull else None
                    want = np.zeros(n // 8 + 8, dtype=np.uint8); wv = np.zeros(n // 8 + 8, dtype=np.uint8); wl, wn = C.c_int64(), C.c_int64()
                    assert cpu.ref_filter_primitive(1, ptr(vals), ptr(vvalid), voff, ptr(mask), ptr(mvalid), moff, n, sel, ptr(want), ptr(wv), C.byref(wl), C.byref(wn)) == 0
                    got = np.zeros(n // 8 + 8, dtype=np.uint8); gv = np.zeros(n // 8 + 8, dtype=np.uint8) if (p_mnull or p_vnull) else None
                    gl, gn = C.c_int64(), C.c_int64()
                    ag.call("ag_filter_primitive", 1, ptr(vals), ptr(vvalid), voff, ptr(mask), ptr(mvalid), moff, n, sel, ptr(got), ptr(gv), C.byref(gl), C.byref(gn))
                    assert gl.value == wl.value
                    valid = unpack_bits(wv, 0, wl.value) if gv is not None else np.ones(wl.value, bool)
                    if gv is not None:
                        assert np.array_equal(unpack_bits(gv, 0, wl.value), valid) and gn.value == wn.value
                    assert np.array_equal(unpack_bits(got, 0, wl.value)[valid], unpack_bits(want, 0, wl.value)[valid]), (n, sel, p_mask, p_mnull)
