"""is_in / unique: the oracle pinned against the reference's literal vectors
(scalar_set_lookup_test.go:104-167, vector_hash_test.go:236-255; tests/golden/set_lookup.json) and against pyarrow,
then the GPU hash-table kernels against the oracle."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import ALL_TYPES, NP_OF, TYPE_NAME, Dev, pack_bits, ptr, unpack_bits

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "set_lookup.json")))
gpu = pytest.mark.gpu
WIDTHS = {1: 8, 2: 16, 4: 32, 8: 64}


def col(lst, dt):
    vals = np.array([0 if v is None else v for v in lst], dtype=dt)
    valid = np.array([v is not None for v in lst], dtype=bool)
    return vals, (pack_bits(valid) if not valid.all() else None)


def oracle_is_in(cpu, bw, v, vb, voff, n, s, sb, soff, sn, matching):
    d = np.zeros(n // 8 + 8, dtype=np.uint8); val = np.zeros(n // 8 + 8, dtype=np.uint8)
    nn = C.c_int64()
    assert cpu.ref_is_in(bw, ptr(v), ptr(vb), voff, n, ptr(s), ptr(sb), soff, sn, matching, ptr(d), ptr(val), C.byref(nn)) == 0
    return unpack_bits(d, 0, n), unpack_bits(val, 0, n), nn.value


def oracle_unique(cpu, bw, v, vb, voff, n):
    out = np.zeros(max(n, 1), dtype=v.dtype); ov = np.zeros(n // 8 + 8, dtype=np.uint8)
    ln, nn = C.c_int64(), C.c_int64()
    assert cpu.ref_unique(bw, ptr(v), ptr(vb), voff, n, ptr(out), ptr(ov), C.byref(ln), C.byref(nn)) == 0
    return out[:ln.value], unpack_bits(ov, 0, ln.value), nn.value


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_oracle_matches_reference_literals(cpu, t):
    dt = np.dtype(NP_OF[t]); bw = WIDTHS[dt.itemsize]
    for case in GOLDEN["is_in"]:
        v, vb = col(case["input"], dt); s, sb = col(case["set"], dt)
        if v.size == 0:
            continue
        for matching, expected in case["cases"]:
            d, val, _ = oracle_is_in(cpu, bw, v, vb, 0, v.size, s, sb, 0, s.size, matching)
            got = [None if not ok else bool(b) for b, ok in zip(d, val)]
            assert got == expected, (case["name"], matching)
    for case in GOLDEN["unique"]:
        v, vb = col(case["input"], dt)
        out, ov, _ = oracle_unique(cpu, bw, v, vb, 0, v.size)
        assert [None if not ok else x for x, ok in zip(out.tolist(), ov)] == [None if e is None else dt.type(e).item() for e in case["expected"]]


def test_oracle_matches_pyarrow(cpu):
    import pyarrow as pa
    import pyarrow.compute as pc
    rng = np.random.default_rng(3)
    for dt in (np.int64, np.int32, np.uint8):
        for n in (1, 100, 5000):
            v = rng.integers(0, 50, n).astype(dt); valid = rng.random(n) > 0.2
            s = rng.integers(0, 50, 17).astype(dt); svalid = rng.random(17) > 0.3
            bw = WIDTHS[np.dtype(dt).itemsize]
            d, val, _ = oracle_is_in(cpu, bw, v, pack_bits(valid), 0, n, s, pack_bits(svalid), 0, 17, 0)
            want = pc.is_in(pa.array(v, mask=~valid), value_set=pa.array(s, mask=~svalid), skip_nulls=False)
            assert np.array_equal(d, np.array(want.to_pylist(), dtype=bool)) and val.all()
            d, val, _ = oracle_is_in(cpu, bw, v, pack_bits(valid), 0, n, s, pack_bits(svalid), 0, 17, 1)
            want = pc.is_in(pa.array(v, mask=~valid), value_set=pa.array(s, mask=~svalid), skip_nulls=True)
            assert np.array_equal(d, np.array(want.to_pylist(), dtype=bool))
            out, ov, _ = oracle_unique(cpu, bw, v, pack_bits(valid), 0, n)
            want = pc.unique(pa.array(v, mask=~valid)).to_pylist()
            assert [None if not ok else x for x, ok in zip(out.tolist(), ov)] == want


# ------------------------------------------------------------------------------------------------- GPU
@gpu
@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_gpu_is_in_and_unique_match_oracle(ag, cpu, t):
    dt = np.dtype(NP_OF[t]); bw = WIDTHS[dt.itemsize]
    rng = np.random.default_rng(50 + t)
    for n in (1, 31, 33, 1000, 70_001):
        for card in (5, 3000):
            for p_null in (0.0, 0.25):
                hi = min(card, np.iinfo(dt).max if dt.kind != "f" else card)
                v = rng.integers(0, hi, n).astype(dt)
                if dt.kind == "f" and n > 8:
                    v[rng.integers(0, n, 4)] = [np.nan, -0.0, 0.0, np.nan]
                valid = rng.random(n) >= p_null
                voff = int(rng.integers(0, 9))
                vv = np.concatenate([np.zeros(voff, dtype=dt), v]); vb = pack_bits(valid, voff) if p_null else None
                sn = int(rng.integers(0, 200))
                s = rng.integers(0, hi, max(sn, 1)).astype(dt); svalid = rng.random(max(sn, 1)) >= p_null
                soff = int(rng.integers(0, 5))
                ss = np.concatenate([np.zeros(soff, dtype=dt), s]); sb = pack_bits(svalid, soff) if p_null else None
                for matching in range(4):
                    wd, wv, wn = oracle_is_in(cpu, bw, vv, vb, voff, n, ss, sb, soff, sn, matching)
                    gd = np.zeros(n // 8 + 8, dtype=np.uint8); gv = np.zeros(n // 8 + 8, dtype=np.uint8); gn = C.c_int64()
                    ag.call("ag_is_in", bw, ptr(vv), ptr(vb), voff, n, ptr(ss), ptr(sb), soff, sn, matching, ptr(gd), ptr(gv), C.byref(gn))
                    assert np.array_equal(unpack_bits(gd, 0, n), wd) and np.array_equal(unpack_bits(gv, 0, n), wv) and gn.value == wn, (TYPE_NAME[t], n, card, p_null, matching)
                wo, wov, wnn = oracle_unique(cpu, bw, vv, vb, voff, n)
                go = np.zeros(n, dtype=dt); gov = np.zeros(n // 8 + 8, dtype=np.uint8) if p_null else None
                gl, gnn = C.c_int64(), C.c_int64()
                ag.call("ag_unique", bw, ptr(vv), ptr(vb), voff, n, ptr(go), ptr(gov), C.byref(gl), C.byref(gnn))
                assert gl.value == wo.size, (TYPE_NAME[t], n, card, p_null)
                if p_null:
                    assert np.array_equal(unpack_bits(gov, 0, gl.value), wov) and gnn.value == wnn
                # the value under the (single) null slot is unspecified by the format: compare the valid slots
                assert go[:gl.value][wov].tobytes() == wo[wov].tobytes(), (TYPE_NAME[t], n, card, p_null)


@gpu
def test_gpu_set_lookup_100m_rows(ag):
    """100M int64 rows: is_in against a 1000-value set (every output bit checked on the host with numpy.isin) and
    unique of a 100-value column (exactly the 100 values, in order of first appearance)."""
    n = 100_000_000
    v = Dev(nbytes=n * 8)
    ag.call("ag_generate_dev", 1, 0x15, 0, 99_999, v.ptr, n, None)
    st = np.arange(0, 100_000, 100, dtype=np.int64)
    ds = Dev(st)
    od, ov, nn = Dev(nbytes=n // 8 + 64), Dev(nbytes=n // 8 + 64), Dev(np.zeros(1, dtype=np.int64))
    ag.call("ag_is_in_dev", 64, v.ptr, None, 0, n, ds.ptr, None, 0, st.size, 0, od.ptr, ov.ptr, nn.ptr, None)
    ag.call("ag_stream_sync", None)
    hv = v.buf.to_numpy(np.int64, n)
    got = np.unpackbits(od.buf.to_numpy(np.uint8, n // 8), bitorder="little").astype(bool)
    assert np.array_equal(got, hv % 100 == 0) and nn.get()[0] == 0
    ag.call("ag_generate_dev", 1, 0x16, 0, 99, v.ptr, n, None)
    out, ln = Dev(nbytes=n * 8), Dev(np.zeros(2, dtype=np.int64))
    ag.call("ag_unique_dev", 64, v.ptr, None, 0, n, out.ptr, None, n, ln.ptr, None)
    ag.call("ag_stream_sync", None)
    k = int(ln.get()[0])
    hv = v.buf.to_numpy(np.int64, 1_000_000)
    _, first = np.unique(hv, return_index=True)
    assert k == 100 and np.array_equal(out.buf.to_numpy(np.int64, k), hv[np.sort(first)])


@gpu
@pytest.mark.parametrize("dt", [np.int64, np.uint32, np.float64, np.uint16, np.int8], ids=lambda d: np.dtype(d).name)
def test_gpu_is_in_every_probe_path(ag, cpu, dt):
    """The three membership structures of is_in_kernel — the shared-memory table at load factor <= 1/8 (few hundred
    values) and <= 1/2 (a few thousand), the HBM table (tens of thousands), the bitmap of 1-/2-byte values — against the
    oracle, with nulls on both sides, the all-ones key in the set, and every NullMatchingBehavior."""
    dt = np.dtype(dt); bw = WIDTHS[dt.itemsize]
    rng = np.random.default_rng(7 + dt.itemsize)
    n = 200_003
    for sn in (300, 3000, 20_000):
        if dt.kind == "f":
            v = rng.integers(-40_000, 40_000, n).astype(dt); s = rng.integers(-40_000, 40_000, sn).astype(dt)
            v[:8] = [np.nan, -0.0, 0.0, np.inf, -np.inf, np.nan, 1.5, -1.5]; s[:3] = [np.nan, -0.0, np.inf]
        else:
            info = np.iinfo(dt)
            lo, hi = max(info.min, -40_000), min(info.max, 40_000)
            v = rng.integers(lo, hi, n, endpoint=True).astype(dt); s = rng.integers(lo, hi, sn, endpoint=True).astype(dt)
            v[:4] = [info.max, info.min, info.max, 0]; s[:2] = [info.max, 0]
            if dt.kind == "u" or dt.itemsize == 8:
                v[4] = dt.type(-1) if dt.kind == "i" else info.max     # the all-ones byte pattern
                s[2] = v[4]
        valid = rng.random(n) >= 0.1; svalid = rng.random(sn) >= 0.01
        vb, sb = pack_bits(valid, 5), pack_bits(svalid, 2)
        vv = np.concatenate([np.zeros(5, dtype=dt), v]); ss = np.concatenate([np.zeros(2, dtype=dt), s])
        for matching in range(4):
            wd, wv, wn = oracle_is_in(cpu, bw, vv, vb, 5, n, ss, sb, 2, sn, matching)
            gd = np.zeros(n // 8 + 8, dtype=np.uint8); gv = np.zeros(n // 8 + 8, dtype=np.uint8); gn = C.c_int64()
            ag.call("ag_is_in", bw, ptr(vv), ptr(vb), 5, n, ptr(ss), ptr(sb), 2, sn, matching, ptr(gd), ptr(gv), C.byref(gn))
            assert np.array_equal(unpack_bits(gd, 0, n), wd) and np.array_equal(unpack_bits(gv, 0, n), wv) and gn.value == wn, (dt.name, sn, matching)


@gpu
def test_gpu_unique_small_table_then_overflow(ag, cpu):
    """unique's two-table scheme: with the policy forced down to a 1024-slot first table, columns below its capacity
    (512 distinct values) finish there and columns above it raise the overflow word and are redone on the full-size
    table — same answers as the oracle either way; then the shipped policy at 6M rows (10 / 3M distinct values)."""
    rng = np.random.default_rng(99)
    ag.call("ag_unique_set_policy", 1000, 1024)
    try:
        for dt in (np.int64, np.uint32, np.float64, np.uint8):
            dt = np.dtype(dt); bw = WIDTHS[dt.itemsize]
            for n in (1001, 70_001):
                for card in (5, 400, 520, 3000):
                    for p_null in (0.0, 0.2):
                        hi = min(card, 255) if dt.itemsize == 1 else card
                        v = rng.integers(0, hi, n).astype(dt)
                        if dt.kind == "u":
                            v[rng.integers(0, n, 3)] = np.iinfo(dt).max      # the all-ones key lives outside the table
                        valid = rng.random(n) >= p_null
                        vb = pack_bits(valid, 3) if p_null else None
                        vv = np.concatenate([np.zeros(3, dtype=dt), v])
                        wo, wov, wnn = oracle_unique(cpu, bw, vv, vb, 3, n)
                        go = np.zeros(n, dtype=dt); gov = np.zeros(n // 8 + 8, dtype=np.uint8) if p_null else None
                        gl, gnn = C.c_int64(), C.c_int64()
                        ag.call("ag_unique", bw, ptr(vv), ptr(vb), 3, n, ptr(go), ptr(gov), C.byref(gl), C.byref(gnn))
                        assert gl.value == wo.size, (dt.name, n, card, p_null)
                        if p_null:
                            assert np.array_equal(unpack_bits(gov, 0, gl.value), wov) and gnn.value == wnn
                        assert go[:gl.value][wov].tobytes() == wo[wov].tobytes(), (dt.name, n, card, p_null)
    finally:
        ag.call("ag_unique_set_policy", 0, 0)
    n = 6_000_000
    for card in (10, 3_000_000):
        v = rng.integers(-card // 2, card // 2, n).astype(np.int64)
        dv, out, ln = Dev(v), Dev(nbytes=n * 8), Dev(np.zeros(2, dtype=np.int64))
        ag.call("ag_unique_dev", 64, dv.ptr, None, 0, n, out.ptr, None, n, ln.ptr, None)
        ag.call("ag_stream_sync", None)
        _, first = np.unique(v, return_index=True)
        k = int(ln.get()[0])
        assert k == first.size and np.array_equal(out.buf.to_numpy(np.int64, k), v[np.sort(first)]), card
