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


# ---- robustness: no read past the end of the input, malformed records are errors (not a silent end of stream) ----------
class GuardedBuffer:
    """`data` placed so that its last byte is the last byte of a page, with a PROT_NONE page right behind it: a decoder
    that reads one byte too far dies with SIGSEGV instead of passing by luck."""

    def __init__(self, data: bytes):
        import mmap
        self.page = mmap.PAGESIZE
        self.libc = ctypes.CDLL(None, use_errno=True)
        self.libc.mmap.restype = ctypes.c_void_p
        self.libc.mmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long]
        self.libc.mprotect.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        self.libc.munmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        n_pages = (len(data) + self.page - 1) // self.page + 1
        self.size = (n_pages + 1) * self.page
        base = self.libc.mmap(None, self.size, mmap.PROT_READ | mmap.PROT_WRITE, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS, -1, 0)
        assert base not in (None, ctypes.c_void_p(-1).value)
        self.base = base
        guard = base + n_pages * self.page
        assert self.libc.mprotect(guard, self.page, 0) == 0      # prot 0 = PROT_NONE
        self.addr = guard - len(data)
        ctypes.memmove(self.addr, data, len(data)) if data else None
        self.len = len(data)

    def close(self):
        self.libc.munmap(self.base, self.size)


def decode_guarded(data, id_format=_capi.ANN_ID_AUTO, cap=64):
    L = _capi.lib()
    g = GuardedBuffer(data)
    try:
        idv, dim, used = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int64()
        row = np.zeros(cap, dtype=np.float32)
        rc = L.ann_persisted_embedding_decode(ctypes.cast(g.addr, ctypes.POINTER(ctypes.c_ubyte)), g.len, id_format, ctypes.byref(idv),
                                              row.ctypes.data, cap, ctypes.byref(dim), ctypes.byref(used))
        return rc, idv.value, row[: max(0, min(dim.value, cap))].copy(), used.value, dim.value
    finally:
        g.close()


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("id_format", [_capi.ANN_ID_INT64_BE, _capi.ANN_ID_INT32_BE])
def test_every_truncation_ends_the_stream_without_reading_past_the_input(built_lib, layout, id_format):
    row = np.arange(1, 6, dtype=np.float32) * 0.5
    data = encode(77, row, id_format=id_format, layout=layout)
    rc, idv, got, used, dim = decode_guarded(data)
    assert rc == 0 and idv == 77 and used == len(data) and got.tolist() == row.tolist()
    for cut in range(len(data)):      # every proper prefix: a partial trailing record is the end of the stream
        rc, _, _, used, dim = decode_guarded(data[:cut])
        assert rc == 0 and used == 0 and dim == 0, cut


def test_mutated_records_never_crash_and_never_end_the_stream_silently_mid_record(built_lib):
    rng = np.random.default_rng(20261019)
    base = [encode(5, np.linspace(-1, 1, 9, dtype=np.float32), layout=l) for l in (0, 1, 2)]
    base.append(hand_record(9, [1.5, 2.5], shape=True))
    outcomes = {"ok": 0, "eos": 0, "error": 0}
    for it in range(3000):
        data = bytearray(base[it % len(base)])
        for _ in range(int(rng.integers(1, 4))):
            data[int(rng.integers(0, len(data)))] = int(rng.integers(0, 256))
        rc, idv, got, used, dim = decode_guarded(bytes(data))      # a read past the input would be a SIGSEGV here
        assert rc in (0, _capi.ANN_ERR_INVALID_ARGUMENT), rc
        if rc == 0:
            assert 0 <= used <= len(data) and dim >= 0
            outcomes["ok" if used else "eos"] += 1
        else:
            outcomes["error"] += 1
    # all three happen: mutations of padding-like bytes still decode, longer declared sizes run off the end, broken type
    # bytes / negative sizes are errors
    assert min(outcomes.values()) > 0, outcomes


def test_malformed_record_in_mid_stream_is_an_error_not_a_short_index(built_lib):
    """Only END_OF_FILE ends the reference's iterator (ThriftIteratorIO.scala:42-49); a TProtocolException propagates.  A
    record whose embedding holds an unknown thrift type, with a good record BEHIND it, must not load as a shorter index."""
    good = hand_record(1, [1.0, 2.0])
    bad = bytearray(hand_record(2, [3.0, 4.0]))
    pos = bad.index(fh(15, 1)) + 3      # element-type byte of list<double>: 4 -> 99 (no such TType)
    assert bad[pos] == 4
    bad[pos] = 99
    stream = bytes(bad) + good
    rc, *_ = decode_guarded(stream)
    assert rc == _capi.ANN_ERR_INVALID_ARGUMENT
    neg = bytearray(hand_record(2, [3.0, 4.0]))
    neg[pos + 1: pos + 5] = struct.pack(">i", -7)      # negative list length
    rc, *_ = decode_guarded(bytes(neg) + good)
    assert rc == _capi.ANN_ERR_INVALID_ARGUMENT
