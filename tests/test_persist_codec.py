"""The reference's on-disk record format (csrc/persist.cu) against hand-built byte fixtures -- pure host code, no GPU.

PersistedEmbedding {1: binary id, 2: embedding.Embedding} in TBinaryProtocol (serialization.thrift:7-10;
ThriftIteratorIO.scala:14-56): big-endian, field header = type byte + i16 id, STOP = 0x00, binary = i32 length + bytes,
list = element type byte + i32 count.  The inner struct's assumed layout is stated in persist.cu."""
import ctypes
import struct

import numpy as np
import pytest

from the_algorithm_b200 import _capi

STOP = b"\x00"


def fh(ttype, fid):
    return struct.pack(">bh", ttype, fid)


def hand_record(idv, floats, id_bytes=8, tensor_field=5, shape=False):
    """Built field by field with struct.pack, independently of the library."""
    out = fh(11, 1) + struct.pack(">i", id_bytes) + (struct.pack(">q", idv) if id_bytes == 8 else struct.pack(">i", idv))
    tensor = fh(15, 1) + struct.pack(">bi", 4, len(floats)) + b"".join(struct.pack(">d", float(np.float32(x))) for x in floats)
    if shape:
        tensor += fh(15, 2) + struct.pack(">bi", 10, 1) + struct.pack(">q", len(floats))
    tensor += STOP
    out += fh(12, 2) + fh(12, 1) + fh(12, tensor_field) + tensor + STOP + STOP + STOP
    return out


def encode(idv, row, id_format=_capi.ANN_ID_INT64_BE, layout=_capi.ANN_LAYOUT_FLOAT_TENSOR):
    row = np.ascontiguousarray(row, dtype=np.float32)
    L = _capi.lib()
    n = L.ann_persisted_embedding_encode(idv, id_format, row.ctypes.data, row.shape[0], layout, None, 0)
    buf = (ctypes.c_ubyte * n)()
    assert L.ann_persisted_embedding_encode(idv, id_format, row.ctypes.data, row.shape[0], layout, buf, n) == n
    return bytes(buf)


def decode(data, id_format=_capi.ANN_ID_AUTO, cap=64):
    L = _capi.lib()
    idv, dim, used = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int64()
    row = np.zeros(cap, dtype=np.float32)
    buf = (ctypes.c_ubyte * max(len(data), 1)).from_buffer_copy(data or b"\0")
    rc = L.ann_persisted_embedding_decode(buf, len(data), id_format, ctypes.byref(idv), row.ctypes.data, cap, ctypes.byref(dim),
                                          ctypes.byref(used))
    return rc, idv.value, row[: dim.value].copy(), used.value


def test_writer_matches_hand_built_bytes(built_lib):
    row = [0.5, -1.25, 3.0]
    assert encode(0x0102030405060708, row) == hand_record(0x0102030405060708, row)
    assert encode(-2, row, layout=_capi.ANN_LAYOUT_DOUBLE_TENSOR) == hand_record(-2, row, tensor_field=6)
    assert encode(7, row, id_format=_capi.ANN_ID_INT32_BE) == hand_record(7, row, id_bytes=4)
    # first bytes spelled out: field 1 (STRING=11) id=1, length 8, big-endian long
    assert encode(258, [1.0])[:15] == bytes([11, 0, 1, 0, 0, 0, 8, 0, 0, 0, 0, 0, 0, 1, 2])


@pytest.mark.parametrize("layout", [0, 1, 2])
def test_round_trip_every_layout(built_lib, layout):
    rng = np.random.default_rng(layout)
    row = rng.standard_normal(37).astype(np.float32)
    data = encode(-(2 ** 62) + 5, row, layout=layout)
    rc, idv, got, used = decode(data)
    assert rc == 0 and idv == -(2 ** 62) + 5 and used == len(data)
    assert (got.view(np.uint32) == row.view(np.uint32)).all()       # float -> double -> float is exact


def test_reader_accepts_fixture_variants_and_unknown_fields(built_lib):
    row = [1.5, 2.5]
    # a shape list after the floats, an unknown i64 field in the outer struct, an Int (4-byte) id
    data = hand_record(9, row, shape=True)
    rc, idv, got, used = decode(data)
    assert rc == 0 and idv == 9 and got.tolist() == row and used == len(data)
    extra = fh(10, 7) + struct.pack(">q", 123)
    data2 = data[:-1] + extra + STOP
    rc, idv, got, used = decode(data2)
    assert rc == 0 and idv == 9 and got.tolist() == row and used == len(data2)
    rc, idv, got, _ = decode(hand_record(-5, row, id_bytes=4))
    assert rc == 0 and idv == -5 and got.tolist() == row
    # a stream: the second record starts where the first ended
    stream = hand_record(1, [1.0]) + hand_record(2, [2.0, 3.0])
    rc, idv, got, used = decode(stream)
    assert (idv, got.tolist()) == (1, [1.0])
    rc, idv, got, used2 = decode(stream[used:])
    assert (idv, got.tolist()) == (2, [2.0, 3.0]) and used + used2 == len(stream)


def test_end_of_stream_and_errors(built_lib):
    data = hand_record(1, [1.0, 2.0])
    assert decode(b"")[3] == 0                                       # clean end of file
    assert decode(data[:-3])[0] == 0 and decode(data[:-3])[3] == 0   # truncated trailing record ends the stream (ThriftIteratorIO.scala:42-49)
    rc, *_ = decode(fh(11, 1) + struct.pack(">i", 3) + b"abc" + fh(12, 2) + STOP + STOP)     # 3-byte id: String injection, not native
    assert rc == _capi.ANN_ERR_INVALID_ARGUMENT
    rc, *_ = decode(fh(11, 1) + struct.pack(">i", 8) + b"\0" * 8 + STOP)                   # no embedding
    assert rc == _capi.ANN_ERR_INVALID_ARGUMENT
    rc, *_ = decode(hand_record(1, [1.0]), id_format=_capi.ANN_ID_INT32_BE)                  # 8-byte id refused when Int was asked for
    assert rc == _capi.ANN_ERR_INVALID_ARGUMENT
