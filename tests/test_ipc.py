"""Arrow IPC file reader and ArrowDeviceArrayStream hand-off (arrow_go_b200/host/ipc.cc).

Oracle: the reference's reader cannot run here (Go), so the files come from another implementation of the same
format — pyarrow's writer — and every column is checked against pyarrow's own view of it.  The error cases restate
the reference's (file_reader.go:68-99,353-381; metadata.go:78-109; file_block_test.go) with its messages.

CPU part: footer / schema / record-batch metadata only (no device work).  GPU part: batches in HBM, streams."""
import ctypes as C
import io
import struct

import numpy as np
import pyarrow as pa
import pytest

from arrow_go_b200 import _native as N
from arrow_go_b200 import compute as pc
from arrow_go_b200 import ipc

gpu = pytest.mark.gpu
PA_TYPES = [(pa.bool_(), pc.BOOL), (pa.int8(), pc.INT8), (pa.uint8(), pc.UINT8), (pa.int16(), pc.INT16), (pa.uint16(), pc.UINT16),
            (pa.int32(), pc.INT32), (pa.uint32(), pc.UINT32), (pa.int64(), pc.INT64), (pa.uint64(), pc.UINT64),
            (pa.float32(), pc.FLOAT32), (pa.float64(), pc.FLOAT64)]


def random_column(rng, t, n, null_p):
    if pa.types.is_boolean(t):
        v = rng.random(n) > 0.5
    elif pa.types.is_floating(t):
        v = rng.standard_normal(n).astype(t.to_pandas_dtype())
    else:
        info = np.iinfo(t.to_pandas_dtype())
        v = rng.integers(info.min, info.max, n, dtype=t.to_pandas_dtype(), endpoint=True)
    mask = (rng.random(n) < null_p) if null_p else None
    return pa.array(v, type=t, mask=mask)


def make_file(batches, schema, **opts):
    sink = io.BytesIO()
    with pa.ipc.new_file(sink, schema, options=pa.ipc.IpcWriteOptions(**opts)) as w:
        for b in batches:
            w.write_batch(b)
    return sink.getvalue()


def sample_table(seed=7, sizes=(1000, 0, 37, 70_001), null_p=(0.0, 0.1, 0.0, 0.5)):
    rng = np.random.default_rng(seed)
    schema = pa.schema([pa.field(f"c{i}_{t}", t, nullable=(i % 3 != 0)) for i, (t, _) in enumerate(PA_TYPES)])
    batches = []
    for n, p in zip(sizes, null_p):
        cols = [random_column(rng, t, n, p if schema.field(i).nullable else 0.0) for i, (t, _) in enumerate(PA_TYPES)]
        batches.append(pa.record_batch(cols, schema=schema))
    return schema, batches


def test_schema_and_layout_match_pyarrow():
    schema, batches = sample_table()
    data = make_file(batches, schema)
    r = ipc.FileReader(data)
    assert r.num_records == len(batches) and r.version == 4          # MetadataVersion V5 has enum value 4
    assert r.schema == [(f.name, tid, f.nullable) for f, (_, tid) in zip(schema, PA_TYPES)]
    raw = np.frombuffer(data, dtype=np.uint8)
    for i, b in enumerate(batches):
        rows, body_off, body_len, cols = r.layout(i)
        assert rows == b.num_rows and body_len % 8 == 0 and body_off % 8 == 0
        for c, L in enumerate(cols):
            arr = b.column(c)
            assert L["length"] == len(arr) and L["null_count"] == arr.null_count
            vbuf, dbuf = arr.buffers()
            bits = 1 if pa.types.is_boolean(arr.type) else arr.type.bit_width
            nbytes = (len(arr) + 7) // 8 if bits == 1 else len(arr) * bits // 8
            if len(arr) == 0:
                assert L["data_offset"] == -1            # loadPrimitive: an empty array has no data buffer
                continue
            got = raw[body_off + L["data_offset"]: body_off + L["data_offset"] + nbytes]
            want = np.frombuffer(dbuf, dtype=np.uint8)[:nbytes]
            if bits == 1:   # compare logical bits only (padding bits are unspecified)
                assert np.array_equal(np.unpackbits(got, bitorder="little")[:len(arr)], np.unpackbits(want, bitorder="little")[:len(arr)])
            else:
                assert got.tobytes() == want.tobytes()
            if arr.null_count == 0:
                assert L["validity_offset"] == -1        # loadCommon: NullCount() == 0 skips the bitmap
            else:
                vb = raw[body_off + L["validity_offset"]: body_off + L["validity_offset"] + (len(arr) + 7) // 8]
                want_valid = np.array([x is not None for x in arr.to_pylist()])
                assert np.array_equal(np.unpackbits(vb, bitorder="little")[:len(arr)].astype(bool), want_valid)


def test_sliced_columns_and_legacy_framing():
    """The writer re-bases sliced arrays (offsets are 0 in the file), and files written before 0.15 frame their
    messages without the continuation token (validateFileBlockMetadata's 4-byte prefix branch)."""
    t = pa.table({"a": pa.array(np.arange(100, dtype=np.int32)), "b": pa.array([None if i % 7 == 0 else i * 0.5 for i in range(100)])})
    sl = t.slice(13, 50).to_batches()[0]
    for legacy in (False, True):
        r = ipc.FileReader(make_file([sl], sl.schema, use_legacy_format=legacy))
        rows, body_off, _, cols = r.layout(0)
        assert rows == 50 and cols[0]["length"] == 50 and cols[1]["null_count"] == sl.column(1).null_count


def test_error_cases_use_the_reference_wording():
    schema, batches = sample_table(sizes=(100,), null_p=(0.1,))
    data = bytearray(make_file(batches, schema))

    def fails(buf, needle, code=N.AG_ERR_INVALID):
        with pytest.raises(pc.ArrowError) as e:
            r = ipc.FileReader(bytes(buf))
            r.layout(0)
        assert e.value.code == code and needle in e.value.msg, e.value.msg

    fails(data[:10], "file too small (size=10)")
    bad = bytearray(data); bad[-3] = ord("X")
    fails(bad, "not an Arrow file")
    bad = bytearray(data); bad[-10:-6] = struct.pack("<I", len(data))
    fails(bad, "file is smaller than indicated metadata size")
    # corrupt the first record-batch block of the footer: find it through our own reader's layout
    r = ipc.FileReader(bytes(data))
    _, body_off, body_len, _ = r.layout(0)
    footer_size = struct.unpack("<I", data[-10:-6])[0]
    footer = bytes(data[-10 - footer_size:-10])
    # Block {offset:long, metaDataLength:int, pad, bodyLength:long}: locate by value
    blk_body = struct.pack("<q", body_len)
    pos = footer.rfind(blk_body)
    assert pos >= 16
    at = len(data) - 10 - footer_size + pos
    bad = bytearray(data); bad[at:at + 8] = struct.pack("<q", body_len + 4)
    fails(bad, "is not a multiple of 8")
    bad = bytearray(data); bad[at:at + 8] = struct.pack("<q", body_len + 8)
    fails(bad, "does not match message body length")           # mappedFileBlock.NewMessage :993-997
    bad = bytearray(data); bad[at:at + 8] = struct.pack("<q", body_len + (1 << 30))
    fails(bad, "exceeds file size")
    bad = bytearray(data); bad[at - 8:at - 4] = struct.pack("<i", 2)
    fails(bad, "invalid file block metadata length 2")
    bad = bytearray(data); bad[at - 16:at - 8] = struct.pack("<q", -8)
    fails(bad, "invalid file block offset -8")
    # message framing: length prefix disagrees with the footer
    meta_off = struct.unpack("<q", bytes(data[at - 16:at - 8]))[0]
    bad = bytearray(data); bad[meta_off + 4:meta_off + 8] = struct.pack("<I", 12345)
    fails(bad, "does not match footer length")
    bad = bytearray(data); bad[meta_off:meta_off + 4] = struct.pack("<I", 0)
    fails(bad, "unexpected end-of-stream marker in file block")
    with pytest.raises(pc.ArrowError) as e:
        ipc.FileReader(bytes(data)).layout(5)
    assert "record index out of bounds" in e.value.msg


def test_unsupported_content_is_refused_not_misread():
    s = pa.schema([("s", pa.string())])
    with pytest.raises(pc.ArrowError) as e:
        ipc.FileReader(make_file([pa.record_batch([pa.array(["a"])], schema=s)], s))
    assert e.value.code == N.AG_ERR_NOT_IMPLEMENTED and "utf8" in e.value.msg
    d = pa.array(["x", "y", "x"]).dictionary_encode()
    s = pa.schema([("d", d.type)])
    with pytest.raises(pc.ArrowError) as e:
        ipc.FileReader(make_file([pa.record_batch([d], schema=s)], s))
    assert e.value.code == N.AG_ERR_NOT_IMPLEMENTED and "dictionary" in e.value.msg
    s = pa.schema([("x", pa.int32())])
    data = make_file([pa.record_batch([pa.array(np.arange(10000, dtype=np.int32))], schema=s)], s, compression="lz4")
    with pytest.raises(pc.ArrowError) as e:
        ipc.FileReader(data).layout(0)
    assert e.value.code == N.AG_ERR_NOT_IMPLEMENTED and "compressed" in e.value.msg
    s = pa.schema([("l", pa.list_(pa.int32()))])
    with pytest.raises(pc.ArrowError) as e:
        ipc.FileReader(make_file([pa.record_batch([pa.array([[1], [2, 3]], type=pa.list_(pa.int32()))], schema=s)], s))
    assert e.value.code == N.AG_ERR_NOT_IMPLEMENTED


def test_stream_schema_is_importable_by_pyarrow():
    """get_schema needs no device: the struct it fills must be a valid ArrowSchema for another consumer."""
    schema, batches = sample_table(sizes=(10,), null_p=(0.0,))
    r = ipc.FileReader(make_file(batches, schema))
    st = r.export_stream()
    assert st.device_type == N.DEVICE_CUDA
    sc = ipc.stream_schema(st)
    got = pa.Schema._import_from_c(C.addressof(sc))      # consumes (releases) the struct
    assert got.names == schema.names and [f.type for f in got] == [f.type for f in schema]
    assert [f.nullable for f in got] == [f.nullable for f in schema]
    ipc.stream_release(st)
    assert not st.release


def column_equals(datum, arr):
    vals, valid, nulls = datum.to_numpy()
    want_valid = np.array([x is not None for x in arr.to_pylist()], dtype=bool) if len(arr) else np.zeros(0, dtype=bool)
    assert nulls == arr.null_count and np.array_equal(valid, want_valid)
    want = arr.fill_null(False if pa.types.is_boolean(arr.type) else 0).to_numpy(zero_copy_only=False)
    assert np.array_equal(vals[valid], want[want_valid])


@gpu
def test_record_batches_land_in_hbm_and_compute(ag):
    schema, batches = sample_table()
    r = ipc.FileReader(make_file(batches, schema))
    for i, b in enumerate(batches):
        rows, cols = r.record_batch(i)
        assert rows == b.num_rows
        for c, d in enumerate(cols):
            assert d.type == PA_TYPES[c][1] and len(d) == len(b.column(c))
            column_equals(d, b.column(c))
    # the columns are ordinary device arrays: run the hot path on them
    rows, cols = r.record_batch(3)
    i64, want = cols[7], batches[3].column(7)
    assert want.null_count > 0
    out = pc.Add(i64, i64, no_check_overflow=True)
    vals, valid, _ = out.to_numpy()
    ref = want.fill_null(0).to_numpy(zero_copy_only=False)
    assert np.array_equal(valid, np.array([x is not None for x in want.to_pylist()]))
    with np.errstate(over="ignore"):
        assert np.array_equal(vals[valid], (ref + ref)[valid])
    # arrow/math Sum ignores validity: it adds the raw value buffer, nulls included
    f64col = batches[3].column(10)
    raw = np.frombuffer(f64col.buffers()[1], dtype=np.float64)[:len(f64col)]
    assert pc.math.sum_float64(cols[10]) == pytest.approx(float(np.sum(raw)), rel=1e-12, abs=1e-9)


@gpu
def test_device_stream_round_trip(ag):
    """IPC file -> producer stream (one H2D copy per batch) -> consumer: same columns, release discipline."""
    schema, batches = sample_table(seed=11, sizes=(5000, 1, 0, 123))
    r = ipc.FileReader(make_file(batches, schema))
    st = r.export_stream()
    types, got = ipc.import_stream(st)
    assert not st.release                                   # the consumer released the stream
    assert types == [tid for _, tid in PA_TYPES] and len(got) == len(batches)
    for (rows, cols), b in zip(got, batches):
        assert rows == b.num_rows
        for c, d in enumerate(cols):
            column_equals(d, b.column(c))
    del got


@gpu
def test_large_batch_100m_rows_one_copy(ag):
    """A BASELINE-sized column through the feeder: 100M int64 rows = one 800 MB body, Sum on the device."""
    n = 100_000_000
    x = np.arange(n, dtype=np.int64)
    s = pa.schema([("x", pa.int64())])
    r = ipc.FileReader(make_file([pa.record_batch([pa.array(x)], schema=s)], s))
    rows, cols = r.record_batch(0)
    assert rows == n
    assert pc.math.sum_int64(cols[0]) == n * (n - 1) // 2
