"""Device residency helpers over the C ABI (ag_dev_alloc / ag_upload / ag_download / streams /
events).  Used by the host-side compute mirror, the tests and bench.py.  numpy is only the host
container; nothing here computes."""
import ctypes as C

import numpy as np

from . import _native as N


class Stream:
    def __init__(self):
        h = C.c_void_p()
        N.call("ag_stream_create", C.byref(h))
        self.handle = h

    def sync(self):
        N.call("ag_stream_sync", self.handle)

    def close(self):
        if self.handle:
            N.call("ag_stream_destroy", self.handle)
            self.handle = None


class Event:
    def __init__(self):
        h = C.c_void_p()
        N.call("ag_event_create", C.byref(h))
        self.handle = h

    def record(self, stream=None):
        N.call("ag_event_record", self.handle, stream.handle if stream else None)

    def sync(self):
        N.call("ag_event_sync", self.handle)

    def elapsed_ms(self, end):
        ms = C.c_float()
        N.call("ag_event_elapsed_ms", self.handle, end.handle, C.byref(ms))
        return ms.value


class DeviceBuffer:
    """An owned device allocation (64-byte padded, zero-filled like Arrow allocators)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        N.call("ag_dev_alloc", C.byref(p), self.nbytes)
        self.ptr = p.value

    @classmethod
    def from_numpy(cls, arr, stream=None):
        arr = np.ascontiguousarray(arr)
        buf = cls(arr.nbytes)
        if arr.nbytes:
            N.call("ag_upload", buf.ptr, arr.ctypes.data, arr.nbytes, stream.handle if stream else None)
            N.call("ag_stream_sync", stream.handle if stream else None)
        return buf

    def to_numpy(self, dtype, count=None, offset_bytes=0, stream=None):
        dtype = np.dtype(dtype)
        if count is None:
            count = (self.nbytes - offset_bytes) // dtype.itemsize
        out = np.empty(count, dtype=dtype)
        if out.nbytes:
            N.call("ag_download", out.ctypes.data, self.ptr + offset_bytes, out.nbytes, stream.handle if stream else None)
            N.call("ag_stream_sync", stream.handle if stream else None)
        return out

    def free(self):
        if self.ptr:
            N.call("ag_dev_free", self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """numpy view over ag_host_alloc memory (a pinned memory.Allocator buffer)."""

    def __init__(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if not np.isscalar(shape) else int(shape)
        self.nbytes = n * dtype.itemsize
        p = C.c_void_p()
        N.call("ag_host_alloc", C.byref(p), max(self.nbytes, 1))
        self.ptr = p.value
        raw = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=dtype, count=n)

    def free(self):
        if self.ptr:
            self.array = None
            N.call("ag_host_free", self.ptr)
            self.ptr = None


def device_info():
    dev, sms, cc_major, cc_minor = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    hbm = C.c_size_t()
    N.call("ag_device_info", C.byref(dev), C.byref(sms), C.byref(hbm), C.byref(cc_major), C.byref(cc_minor))
    return {"device": dev.value, "sm_count": sms.value, "hbm_bytes": hbm.value, "cc": (cc_major.value, cc_minor.value)}


def gpu_available():
    """True iff the library loads and sees at least one CUDA device."""
    try:
        cnt = C.c_int()
        st = N.raw().ag_device_count(C.byref(cnt))
        return st == N.AG_OK and cnt.value > 0
    except Exception:
        return False
