"""Arrow IPC files as device-resident record batches, and whole batches over the Arrow C Device Stream interface.

Thin handles over arrow_go_b200/host/ipc.cc (the restatement of arrow/ipc/file_reader.go for numeric / boolean
columns): nothing is parsed or computed in Python.

    r = ipc.FileReader(buf)            # buf: bytes / numpy uint8 (kept alive by the reader, like a memory map)
    r.schema                           # [(name, type_id, nullable), ...]
    r.layout(i)                        # metadata only — no device work
    rows, cols = r.record_batch(i)     # ONE host-to-device copy; cols are compute.Datum arrays in HBM
"""
import ctypes as C

import numpy as np

from . import _native as N
from . import compute as pc


class ArrowDeviceArrayStream(C.Structure):
    """arrow/cdata/abi.h:170-200"""
    _fields_ = [("device_type", C.c_int32), ("get_schema", C.c_void_p), ("get_next", C.c_void_p),
                ("get_last_error", C.c_void_p), ("release", C.c_void_p), ("private_data", C.c_void_p)]


_GET_SCHEMA = C.CFUNCTYPE(C.c_int, C.POINTER(ArrowDeviceArrayStream), C.POINTER(N.ArrowSchema))
_GET_NEXT = C.CFUNCTYPE(C.c_int, C.POINTER(ArrowDeviceArrayStream), C.POINTER(N.ArrowDeviceArray))
_RELEASE = C.CFUNCTYPE(None, C.POINTER(ArrowDeviceArrayStream))


def _lib():
    h = pc.lib()
    if not getattr(h, "_ipc_ready", False):
        p, i, i64 = C.c_void_p, C.c_int, C.c_int64
        h.agx_ipc_open.argtypes = [p, i64, C.POINTER(p)]
        h.agx_ipc_close.argtypes = [p]
        h.agx_ipc_close.restype = None
        for nm in ("agx_ipc_num_fields", "agx_ipc_num_records", "agx_ipc_version"):
            getattr(h, nm).argtypes = [p]
        h.agx_ipc_field.argtypes = [p, i, C.c_char_p, i64, C.POINTER(i), C.POINTER(i)]
        h.agx_ipc_layout.argtypes = [p, i, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
        h.agx_ipc_read_batch.argtypes = [p, i, C.POINTER(i64), C.POINTER(p)]
        h.agx_ipc_export_stream.argtypes = [p, C.POINTER(ArrowDeviceArrayStream)]
        h.agx_ipc_import_stream.argtypes = [C.POINTER(ArrowDeviceArrayStream), C.POINTER(i), C.POINTER(i), i, C.POINTER(i), C.POINTER(i64), C.POINTER(p)]
        h._ipc_ready = True
    return h


class FileReader:
    """ipc.NewMappedFileReader (arrow/ipc/file_reader.go:231-250)."""

    def __init__(self, data):
        self._buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        self._h = C.c_void_p()
        pc._check(_lib().agx_ipc_open(self._buf.ctypes.data, self._buf.size, C.byref(self._h)))

    def close(self):
        if self._h:
            _lib().agx_ipc_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def schema(self):
        out = []
        for k in range(_lib().agx_ipc_num_fields(self._h)):
            name = C.create_string_buffer(512)
            t, nullable = C.c_int(), C.c_int()
            pc._check(_lib().agx_ipc_field(self._h, k, name, 512, C.byref(t), C.byref(nullable)))
            out.append((name.value.decode(), t.value, bool(nullable.value)))
        return out

    @property
    def num_records(self):
        return _lib().agx_ipc_num_records(self._h)

    @property
    def version(self):
        return _lib().agx_ipc_version(self._h)

    def layout(self, i):
        """(rows, body_offset, body_length, [dict per column]) — metadata only, no device work."""
        nf = _lib().agx_ipc_num_fields(self._h)
        rows, bo, bl = C.c_int64(), C.c_int64(), C.c_int64()
        cols = (C.c_int64 * (6 * max(nf, 1)))()
        pc._check(_lib().agx_ipc_layout(self._h, i, C.byref(rows), C.byref(bo), C.byref(bl), cols))
        keys = ("length", "null_count", "validity_offset", "validity_length", "data_offset", "data_length")
        return rows.value, bo.value, bl.value, [dict(zip(keys, cols[6 * c:6 * c + 6])) for c in range(nf)]

    def record_batch(self, i):
        """RecordBatchAt(i): (rows, [Datum per column]); the batch body crosses the link once."""
        nf = _lib().agx_ipc_num_fields(self._h)
        rows = C.c_int64()
        cols = (C.c_void_p * max(nf, 1))()
        pc._check(_lib().agx_ipc_read_batch(self._h, i, C.byref(rows), cols))
        return rows.value, [pc.Datum(C.c_void_p(cols[c])) for c in range(nf)]

    def export_stream(self):
        """A producer ArrowDeviceArrayStream over the file's record batches (the stream shares the reader)."""
        s = ArrowDeviceArrayStream()
        pc._check(_lib().agx_ipc_export_stream(self._h, C.byref(s)))
        return s


def import_stream(stream, max_fields=64, max_batches=1024):
    """Drain a device stream: ([type ids], [(rows, [Datum per column]) per batch]).  Releases the stream."""
    nf, nb = C.c_int(), C.c_int(max_batches)
    types = (C.c_int * max_fields)()
    rows = (C.c_int64 * max_batches)()
    cols = (C.c_void_p * (max_fields * max_batches))()
    pc._check(_lib().agx_ipc_import_stream(C.byref(stream), C.byref(nf), types, max_fields, C.byref(nb), rows, cols))
    batches = []
    for b in range(nb.value):
        batches.append((rows[b], [pc.Datum(C.c_void_p(cols[b * nf.value + c])) for c in range(nf.value)]))
    return list(types[:nf.value]), batches


def stream_schema(stream):
    """get_schema of a stream as an ArrowSchema struct (caller releases it)."""
    sc = N.ArrowSchema()
    rc = C.cast(stream.get_schema, _GET_SCHEMA)(C.byref(stream), C.byref(sc))
    if rc != 0:
        raise pc.ArrowError(1, f"get_schema failed: {rc}")
    return sc


def stream_release(stream):
    if stream.release:
        C.cast(stream.release, _RELEASE)(C.byref(stream))
