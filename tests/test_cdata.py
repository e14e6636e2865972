"""Arrow C Device Data Interface hand-off (include/arrowgpu_cdata.h).

CPU part: the struct layout agrees with the C compiler and with another producer of the same ABI
(pyarrow's _export_to_c_device), format strings map to the reference's arrow.Type ids.
GPU part: export -> import round trips by pointer, sync events, release-callback discipline,
pinned-host (ARROW_DEVICE_CUDA_HOST) arrays consumed in place."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from arrow_go_b200 import _native as N
from helpers import NP_OF, TYPE_NAME, Dev, pack_bits, ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FORMATS = {N.BOOL: b"b", N.INT8: b"c", N.UINT8: b"C", N.INT16: b"s", N.UINT16: b"S", N.INT32: b"i", N.UINT32: b"I",
           N.INT64: b"l", N.UINT64: b"L", N.FLOAT32: b"f", N.FLOAT64: b"g"}


def test_struct_layout_matches_the_c_compiler():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "arrowgpu_cdata.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(struct ArrowSchema), sizeof(struct ArrowArray), sizeof(struct ArrowDeviceArray), sizeof(ag_array_view));
  printf("%zu %zu %zu %zu\n", offsetof(struct ArrowDeviceArray, device_id), offsetof(struct ArrowDeviceArray, device_type),
         offsetof(struct ArrowDeviceArray, sync_event), offsetof(struct ArrowDeviceArray, reserved));
  printf("%zu %zu %zu\n", offsetof(struct ArrowArray, buffers), offsetof(struct ArrowArray, release), offsetof(struct ArrowSchema, release));
  printf("%zu %zu %zu\n", offsetof(ag_array_view, validity), offsetof(ag_array_view, device_type), offsetof(ag_array_view, device_id));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    A, D, S, V = N.ArrowArray, N.ArrowDeviceArray, N.ArrowSchema, N.ArrayView
    want = [C.sizeof(S), C.sizeof(A), C.sizeof(D), C.sizeof(V),
            D.device_id.offset, D.device_type.offset, D.sync_event.offset, D.reserved.offset,
            A.buffers.offset, A.release.offset, S.release.offset,
            V.validity.offset, V.device_type.offset, V.device_id.offset]
    assert got == want
    assert (C.sizeof(S), C.sizeof(A), C.sizeof(D)) == (72, 80, 128)  # the published ABI on LP64


def test_format_strings():
    for t, f in FORMATS.items():
        assert N.raw().ag_type_to_schema_format(t) == f
        out = C.c_int(-1)
        N.call("ag_schema_format_to_type", f, C.byref(out))
        assert out.value == t
    assert N.raw().ag_type_to_schema_format(13) is None           # arrow.STRING
    st, msg = N.call_status("ag_schema_format_to_type", b"u", C.byref(C.c_int()))
    assert st == N.AG_ERR_NOT_IMPLEMENTED and "'u'" in msg


def pyarrow_device_export(arr):
    da, sc = N.ArrowDeviceArray(), N.ArrowSchema()
    arr._export_to_c_device(C.addressof(da), C.addressof(sc))
    return da, sc


def test_layout_matches_pyarrow_producer():
    """Another implementation of the ABI fills the structs; ours reads them back."""
    pa = pytest.importorskip("pyarrow")
    cases = [(pa.array([1, None, 3, 4, None, 6, 7], type=pa.int32()), N.INT32), (pa.array([1.5, 2.5, None], type=pa.float64()), N.FLOAT64),
             (pa.array([True, None, False, True], type=pa.bool_()), N.BOOL), (pa.array(list(range(100)), type=pa.uint16()).slice(13, 50), N.UINT16)]
    for arr, t in cases:
        da, sc = pyarrow_device_export(arr)
        try:
            v = N.ArrayView()
            N.call("ag_device_array_describe", C.byref(da), C.byref(sc), C.byref(v))
            assert (v.type, v.length, v.offset, v.null_count) == (t, len(arr), arr.offset, arr.null_count)
            assert v.device_type == N.DEVICE_CPU and da.array.n_buffers == 2
            if t != N.BOOL:  # CPU memory: read the values through the pointer the view holds
                raw = np.ctypeslib.as_array(C.cast(v.values, C.POINTER(C.c_uint8)), shape=((v.offset + v.length) * np.dtype(NP_OF[t]).itemsize,))
                vals = raw.view(NP_OF[t])[v.offset:]
                want = arr.to_numpy(zero_copy_only=False)
                ok = ~np.isnan(want.astype(np.float64)) if arr.null_count else np.ones(len(arr), dtype=bool)
                assert np.array_equal(vals[ok], want[ok].astype(NP_OF[t]))
            assert bool(v.validity) == (arr.null_count > 0)
        finally:
            da.array.release(C.byref(da.array))
            sc.release(C.byref(sc))
        assert not da.array.release  # released marker
    # layouts outside this path are refused, not misread
    da, sc = pyarrow_device_export(pa.array(["a", "bc"]))
    st, msg = N.call_status("ag_device_array_describe", C.byref(da), C.byref(sc), C.byref(N.ArrayView()))
    assert st == N.AG_ERR_NOT_IMPLEMENTED
    da.array.release(C.byref(da.array))
    sc.release(C.byref(sc))


# ------------------------------------------------------------------ GPU -----------------
gpu = pytest.mark.gpu


@gpu
def test_export_import_roundtrip(ag):
    rng = np.random.default_rng(3)
    n = 100_003
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    valid = pack_bits(rng.random(n) > 0.2, offset=0)
    dx, dy, dv = Dev(x), Dev(y), Dev(valid)
    released = []
    cb = N.RELEASE_BUFFERS_FN(lambda opaque: released.append(opaque))
    da, sc = N.ArrowDeviceArray(), N.ArrowSchema()
    ag.call("ag_export_device_array", N.FLOAT64, n - 5, -1, 5, dv.ptr, dx.ptr, cb, 1234, None, C.byref(da), C.byref(sc))
    assert (da.device_type, da.array.length, da.array.offset, da.array.null_count, da.array.n_buffers) == (N.DEVICE_CUDA, n - 5, 5, -1, 2)
    assert da.sync_event and sc.format == b"g" and sc.flags == 2
    assert (da.array.buffers[0], da.array.buffers[1]) == (dv.ptr, dx.ptr)
    v = N.ArrayView()
    ag.call("ag_import_device_array", C.byref(da), C.byref(sc), None, C.byref(v))
    assert (v.type, v.length, v.offset, v.values, v.validity) == (N.FLOAT64, n - 5, 5, dx.ptr, dv.ptr)
    # the imported pointers go straight into the *_dev entry points
    do = Dev(np.zeros(n - 5))
    ag.call("ag_arith_binary_dev", N.FLOAT64, N.OP_ADD, N.SHAPE_AA, v.values + 8 * v.offset, dy.ptr + 8 * 5, do.ptr, v.length, None)
    ag.call("ag_stream_sync", None)
    assert np.array_equal(do.get(), x[5:] + y[5:])
    cnt = Dev(np.zeros(1, dtype=np.int64))
    ag.call("ag_bitmap_popcount_dev", v.validity, v.offset, v.length, cnt.ptr, None)
    ag.call("ag_stream_sync", None)
    assert cnt.get()[0] == int(np.unpackbits(valid, bitorder="little")[5:n].sum())
    assert released == []
    da.array.release(C.byref(da.array))
    sc.release(C.byref(sc))
    assert released == [1234] and not da.array.release
    # a released array is refused
    st, msg = ag.call_status("ag_import_device_array", C.byref(da), None, None, C.byref(v))
    assert st == N.AG_ERR_INVALID and "released" in msg


@gpu
def test_import_orders_the_consumer_after_the_producer(ag):
    """sync_event: a consumer stream that imports right after the producer queued its kernel sees the
    finished output (the wait is inserted by ag_import_device_array)."""
    n = 20_000_000
    a = Dev(np.zeros(1, dtype=np.float64), byte_offset=0)
    from arrow_go_b200.device import DeviceBuffer, Stream
    x, o, o2 = DeviceBuffer(n * 8), DeviceBuffer(n * 8), DeviceBuffer(n * 8)
    ag.call("ag_generate_dev", 3, 7, -1000, 1000, x.ptr, n, None)
    ag.call("ag_stream_sync", None)
    prod, cons = Stream(), Stream()
    ag.call("ag_arith_binary_dev", N.FLOAT64, N.OP_ADD, N.SHAPE_AA, x.ptr, x.ptr, o.ptr, n, prod.handle)
    da = N.ArrowDeviceArray()
    ag.call("ag_export_device_array", N.FLOAT64, n, 0, 0, None, o.ptr, N.RELEASE_BUFFERS_FN(0), None, prod.handle, C.byref(da), None)
    v = N.ArrayView()
    ag.call("ag_import_device_array", C.byref(da), None, cons.handle, C.byref(v))
    ag.call("ag_arith_binary_dev", N.FLOAT64, N.OP_SUB, N.SHAPE_AA, v.values, x.ptr, o2.ptr, n, cons.handle)  # (x+x)-x
    ag.call("ag_stream_sync", cons.handle)
    s1, s2 = Dev(np.zeros(1, dtype=np.uint64)), Dev(np.zeros(1, dtype=np.uint64))
    ag.call("ag_checksum64_dev", o2.ptr, n, s1.ptr, None)
    ag.call("ag_checksum64_dev", x.ptr, n, s2.ptr, None)
    ag.call("ag_stream_sync", None)
    assert s1.get()[0] == s2.get()[0]
    da.array.release(C.byref(da.array))
    del a


@gpu
def test_pinned_host_array_is_consumed_in_place(ag):
    """ARROW_DEVICE_CUDA_HOST: memory from ag_host_alloc (the pinned memory.Allocator) is handed over
    by pointer and read by the kernel across PCIe — no staging copy."""
    n = 1_000_003
    p = C.c_void_p()
    ag.call("ag_host_alloc", C.byref(p), n * 8)
    host = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int64)), shape=(n,))
    host[:] = np.random.default_rng(9).integers(-1 << 40, 1 << 40, n)
    bufs = (C.c_void_p * 2)(None, p.value)
    released = []
    rel = C.CFUNCTYPE(None, C.POINTER(N.ArrowArray))(lambda a: released.append(1))
    da = N.ArrowDeviceArray()
    da.array.length, da.array.null_count, da.array.offset, da.array.n_buffers = n, 0, 0, 2
    da.array.buffers = C.cast(bufs, C.POINTER(C.c_void_p))
    da.array.release = rel
    da.device_type, da.device_id = N.DEVICE_CUDA_HOST, -1
    sc = N.ArrowSchema()
    sc.format, sc.release = b"l", C.CFUNCTYPE(None, C.POINTER(N.ArrowSchema))(lambda s: None)
    v = N.ArrayView()
    ag.call("ag_import_device_array", C.byref(da), C.byref(sc), None, C.byref(v))
    assert v.type == N.INT64 and v.values == p.value
    res = Dev(np.zeros(1, dtype=np.int64))
    ag.call("ag_sum_i64_dev", v.values, n, res.ptr, None)
    ag.call("ag_stream_sync", None)
    assert res.get()[0] == int(host.sum())
    da.array.release(C.byref(da.array))
    assert released == [1]
    ag.call("ag_host_free", p)


@gpu
def test_cpu_and_foreign_devices_are_rejected(ag):
    pa = pytest.importorskip("pyarrow")
    da, sc = pyarrow_device_export(pa.array([1, 2, 3], type=pa.int64()))
    st, msg = ag.call_status("ag_import_device_array", C.byref(da), C.byref(sc), None, C.byref(N.ArrayView()))
    assert st == N.AG_ERR_INVALID and "ARROW_DEVICE_CPU" in msg
    da.device_type = 10  # ROCm
    st, msg = ag.call_status("ag_import_device_array", C.byref(da), C.byref(sc), None, C.byref(N.ArrayView()))
    assert st == N.AG_ERR_NOT_IMPLEMENTED
    da.device_type, da.device_id = N.DEVICE_CUDA, 63
    st, msg = ag.call_status("ag_import_device_array", C.byref(da), C.byref(sc), None, C.byref(N.ArrayView()))
    assert st == N.AG_ERR_INVALID and "device 63" in msg
    da.device_type, da.device_id = N.DEVICE_CPU, -1
    da.array.release(C.byref(da.array))
    sc.release(C.byref(sc))


@gpu
def test_host_api_export_import(ag):
    from arrow_go_b200 import compute as pc
    rng = np.random.default_rng(4)
    n = 5000
    x, valid = rng.integers(-100, 100, n).astype(np.int32), rng.random(n) > 0.3
    arr = pc.Array.from_numpy(x, valid)
    da, sc = pc.export_device(arr.slice(7, 4000))
    assert da.device_type == N.DEVICE_CUDA and da.array.offset == 7 and da.array.length == 4000 and sc.format == b"i"
    back = pc.import_device(da, sc)          # takes ownership: the caller's struct is marked moved
    assert not da.array.release
    sc.release(C.byref(sc))
    del arr                                  # the exported struct keeps the buffers alive
    out = pc.Add(back, back)
    vals, v, _ = out.to_numpy()
    assert np.array_equal(v, valid[7:4007]) and np.array_equal(vals[v], (2 * x[7:4007])[v])
