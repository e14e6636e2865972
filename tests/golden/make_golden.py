"""Generates tests/golden/ref_simd_vectors.npz by running the REFERENCE's own SIMD loops
(oracle/_ref/libarrowgo_ref.so, assembled from /root/reference by oracle/Makefile) on small
seeded inputs.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The .npz holds inputs and the reference's outputs, so the oracle and the GPU library stay pinned
to the reference's bits wherever /root/reference (and oracle/_ref) is absent."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import ALL_TYPES, NP_OF, TYPE_NAME, cast_inputs, random_values  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    ref = oracle.ref()
    assert ref is not None, "oracle/_ref not built (needs /root/reference)"
    rng = np.random.default_rng(0x0FF1CE)
    out = {}
    # Sum: exactly-summable and general data, AVX2 order (the baseline of record)
    for n in (0, 1, 31, 32, 33, 100, 8192):
        for kind in ("exact", "normal"):
            x = rng.integers(-(1 << 20), 1 << 20, n).astype(np.float64) if kind == "exact" else rng.standard_normal(n)
            r = C.c_double()
            ref.sum_float64_avx2(x.ctypes.data, n, C.addressof(r))
            out[f"sum_f64/{kind}/{n}/x"] = x
            out[f"sum_f64/{kind}/{n}/res"] = np.array([r.value])
        xi = rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, n, dtype=np.int64, endpoint=True)
        ri = C.c_int64()
        ref.sum_int64_avx2(xi.ctypes.data, n, C.addressof(ri))
        out[f"sum_i64/{n}/x"] = xi
        out[f"sum_i64/{n}/res"] = np.array([ri.value], dtype=np.int64)
    # arithmetic: every type x {add, sub, mul} x 3 shapes, n = 67
    n = 67
    for t in ALL_TYPES:
        l, r = random_values(rng, t, n), random_values(rng, t, n)
        out[f"arith/{TYPE_NAME[t]}/l"] = l
        out[f"arith/{TYPE_NAME[t]}/r"] = r
        for op in (0, 1, 2):
            for shape, nm in ((0, "binary"), (1, "arr_scalar"), (2, "scalar_arr")):
                o = np.empty(n, dtype=NP_OF[t])
                getattr(ref, f"arithmetic_{nm}_avx2")(t, op, l.ctypes.data, r.ctypes.data, o.ctypes.data, n)
                out[f"arith/{TYPE_NAME[t]}/{op}/{shape}"] = o
        for op in (4, 5, 20, 26):
            o = np.empty(n, dtype=NP_OF[t])
            ref.arithmetic_unary_same_types_avx2(t, op, l.ctypes.data, o.ctypes.data, n)
            out[f"unary/{TYPE_NAME[t]}/{op}"] = o
    # comparisons: every type x 4 ops x 3 shapes, n = 67, bit offset 3, filler 0xa5
    for t in ALL_TYPES:
        l, r = random_values(rng, t, n, small=True), random_values(rng, t, n, small=True)
        out[f"cmp/{TYPE_NAME[t]}/l"] = l
        out[f"cmp/{TYPE_NAME[t]}/r"] = r
        for ci, cn in enumerate(("equal", "not_equal", "greater", "greater_equal")):
            for shape, sn in ((0, "arr_arr"), (1, "arr_scalar"), (2, "scalar_arr")):
                o = np.full(12, 0xA5, dtype=np.uint8)
                getattr(ref, f"comparison_{cn}_{sn}_avx2")(t, l.ctypes.data, r.ctypes.data, o.ctypes.data, n, 3)
                out[f"cmp/{TYPE_NAME[t]}/{ci}/{shape}"] = o
    # numeric casts: every ordered pair of distinct types, n = 67 (vector body + scalar tail),
    # inputs whose converted value is representable (helpers.cast_inputs)
    crng = np.random.default_rng(0xCA57)
    for ti in ALL_TYPES:
        for to in ALL_TYPES:
            if ti == to:
                continue
            x = cast_inputs(crng, ti, to, n)
            o = np.empty(n, dtype=NP_OF[to])
            ref.cast_type_numeric_avx2(ti, to, x.ctypes.data, o.ctypes.data, n)
            o2 = np.empty(n, dtype=NP_OF[to])
            ref.cast_type_numeric_sse4(ti, to, x.ctypes.data, o2.ctypes.data, n)
            assert o.tobytes() == o2.tobytes() or np.dtype(NP_OF[to]).kind == "f", (ti, to)
            out[f"cast/{TYPE_NAME[ti]}/{TYPE_NAME[to]}/x"] = x
            out[f"cast/{TYPE_NAME[ti]}/{TYPE_NAME[to]}/o"] = o
    path = os.path.join(HERE, "ref_simd_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
