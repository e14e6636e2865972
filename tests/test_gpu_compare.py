"""Comparison kernels on the GPU vs the reference's SIMD loops.  The first test restates
arrow/compute/internal/kernels/scalar_comparisons_test.go:31-119 one to one: offsets 0..7 x
lengths 0..65 x 3 shapes x 4 ops, output pre-filled with 0xa5, inputs left[i]=(7i+1)%11,
right[i]=(5i+3)%11, scalars 6 and 4."""
import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import ALL_TYPES, NP_OF, TYPE_NAME, Dev, ptr, random_values, unpack_bits

pytestmark = pytest.mark.gpu

OPS = {N.CMP_EQ: "equal", N.CMP_NE: "not_equal", N.CMP_GT: "greater", N.CMP_GE: "greater_equal"}
SHAPES = {N.SHAPE_AA: "arr_arr", N.SHAPE_AS: "arr_scalar", N.SHAPE_SA: "scalar_arr"}


def ref_cmp(ref, isa, type_id, cmp, shape, l, r, n, offset, fill):
    out = np.full((offset % 8 + n + 7) // 8 + 3, fill, dtype=np.uint8)
    getattr(ref, f"comparison_{OPS[cmp]}_{SHAPES[shape]}_{isa}")(type_id, ptr(l), ptr(r), ptr(out), n, offset)
    return out


def test_reference_kernel_sweep(ag, ref, isa):
    for type_id in (N.INT32, N.INT64, N.FLOAT64, N.UINT8):
        dt = NP_OF[type_id]
        i = np.arange(65)
        left = ((7 * i + 1) % 11).astype(dt)
        right = ((5 * i + 3) % 11).astype(dt)
        six, four = np.array([6], dtype=dt), np.array([4], dtype=dt)
        for offset in range(8):
            for n in range(66):
                for cmp in OPS:
                    for shape, (l, r) in {N.SHAPE_AA: (left, right), N.SHAPE_AS: (left, six), N.SHAPE_SA: (four, right)}.items():
                        want = ref_cmp(ref, isa, type_id, cmp, shape, l, r, n, offset, 0xA5)
                        got = np.full_like(want, 0xA5)
                        ag.call("ag_compare", type_id, cmp, shape, ptr(l), ptr(r), ptr(got), n, offset)
                        assert got.tobytes() == want.tobytes(), (TYPE_NAME[type_id], offset, n, cmp, shape)


@pytest.mark.parametrize("type_id", ALL_TYPES, ids=lambda t: TYPE_NAME[t])
def test_compare_random_all_types(ag, ref, cpu, isa, type_id):
    rng = np.random.default_rng(0x0FF1CE + type_id)
    dt = NP_OF[type_id]
    isz = np.dtype(dt).itemsize
    for n in (1, 31, 32, 33, 1023, 1024, 1025, 50000, (1 << 20) + 7):
        for shape in SHAPES:
            l = random_values(rng, type_id, 1 if shape == N.SHAPE_SA else n, small=True)
            r = random_values(rng, type_id, 1 if shape == N.SHAPE_AS else n, small=True)
            if dt in (np.float32, np.float64) and n > 8:
                (l if shape != N.SHAPE_SA else r)[[1, 5]] = [np.nan, -0.0]
            for cmp in (N.CMP_EQ, N.CMP_NE, N.CMP_GT, N.CMP_GE, N.CMP_LT, N.CMP_LE):
                offset = int(rng.integers(0, 8))
                if cmp in OPS:
                    want = ref_cmp(ref, isa, type_id, cmp, shape, l, r, n, offset, 0x5A)
                    mine = np.full_like(want, 0x5A)
                    assert cpu.ref_compare(type_id, cmp, shape, ptr(l), ptr(r), ptr(mine), n, offset) == 0
                    assert mine.tobytes() == want.tobytes()
                else:  # LT / LE exist only as flipped GT / GE (scalar_compare.go:73-99)
                    flipped = {N.SHAPE_AA: N.SHAPE_AA, N.SHAPE_AS: N.SHAPE_SA, N.SHAPE_SA: N.SHAPE_AS}[shape]
                    want = ref_cmp(ref, isa, type_id, N.CMP_GT if cmp == N.CMP_LT else N.CMP_GE, flipped, r, l, n, offset, 0x5A)
                got = np.full_like(want, 0x5A)
                ag.call("ag_compare", type_id, cmp, shape, ptr(l), ptr(r), ptr(got), n, offset)
                assert got.tobytes() == want.tobytes(), (TYPE_NAME[type_id], n, shape, cmp, offset)
                # device flavour: odd byte phase of the output pointer + element-misaligned inputs
                for obyte, mis in ((0, 0), (3, 1)):
                    dl = l if shape == N.SHAPE_SA else Dev(l, byte_offset=mis * isz)
                    dr = r if shape == N.SHAPE_AS else Dev(r, byte_offset=mis * isz)
                    dout = Dev(np.full(want.size, 0x5A, dtype=np.uint8), byte_offset=obyte)
                    ag.call("ag_compare_dev", type_id, cmp, shape, ptr(dl) if shape == N.SHAPE_SA else dl.ptr,
                            ptr(dr) if shape == N.SHAPE_AS else dr.ptr, dout.ptr, n, offset, None)
                    ag.call("ag_stream_sync", None)
                    assert dout.get().tobytes() == want.tobytes(), ("dev", TYPE_NAME[type_id], n, shape, cmp, offset, obyte, mis)


def test_named_entry_points(ag, ref, isa):
    # the 12 reference-signature entry points (scalar_comparison.cc:210-256)
    l = np.arange(100, dtype=np.int64)
    r = np.full(100, 50, dtype=np.int64)
    sc = np.array([50], dtype=np.int64)
    for op, cmp in (("eq", N.CMP_EQ), ("ne", N.CMP_NE), ("gt", N.CMP_GT), ("ge", N.CMP_GE)):
        for sh, shape, (a, b) in (("aa", N.SHAPE_AA, (l, r)), ("as", N.SHAPE_AS, (l, sc)), ("sa", N.SHAPE_SA, (sc, l))):
            want = ref_cmp(ref, isa, N.INT64, cmp, shape, a, b, 100, 3, 0)
            got = np.zeros_like(want)
            ag.call(f"ag_cmp_{op}_{sh}", N.INT64, ptr(a), ptr(b), ptr(got), 100, 3)
            assert got.tobytes() == want.tobytes()


def test_greater_100m_rows_properties(ag):
    """Config 3a size: Greater(int64[100M] uniform in [0,100), 89): popcount must equal the count
    of the generator's CPU twin on windows scaled, complement identity popcount(>89)+popcount(<=89)==n,
    and a 1M-row window must match the oracle bit for bit."""
    n = 100_000_000
    v = Dev(nbytes=n * 8)
    ag.call("ag_generate_dev", 1, 0x0FF1CE, 0, 99, v.ptr, n, None)
    nb = (n + 7) // 8
    gt = Dev(nbytes=nb)
    le = Dev(nbytes=nb)
    sc = np.array([89], dtype=np.int64)
    ag.call("ag_compare_dev", N.INT64, N.CMP_GT, N.SHAPE_AS, v.ptr, ptr(sc), gt.ptr, n, 0, None)
    ag.call("ag_compare_dev", N.INT64, N.CMP_LE, N.SHAPE_AS, v.ptr, ptr(sc), le.ptr, n, 0, None)
    cnt = Dev(np.zeros(2, dtype=np.int64))
    ag.call("ag_bitmap_popcount_dev", gt.ptr, 0, n, cnt.ptr, None)
    ag.call("ag_bitmap_popcount_dev", le.ptr, 0, n, cnt.ptr + 8, None)
    ag.call("ag_stream_sync", None)
    a, b = cnt.get()
    assert a + b == n
    assert abs(a / n - 0.10) < 0.001
    from oracle import oracle
    w = 1 << 20
    start = 64 * 1_000_003
    hv = v.buf.to_numpy(np.int64, w, start * 8)
    want = np.zeros(w // 8, dtype=np.uint8)
    assert oracle.cpu().ref_compare(N.INT64, N.CMP_GT, N.SHAPE_AS, ptr(hv), ptr(sc), ptr(want), w, 0) == 0
    assert gt.buf.to_numpy(np.uint8, w // 8, start // 8).tobytes() == want.tobytes()
