"""TFRecord files and tf.train.Example messages without TensorFlow.

The reference reads and writes its patch sets through TensorFlow (importer/TFRecordImporter.py:16-72,
utilities/tfrecord_writer.py:45-81): ``training.tfrecord`` / ``test.tfrecord`` / ``validation.tfrecord`` hold one
``Example{label: int64, image: float[P*P*C]}`` per patch (optionally GZIP-compressed as a whole file), and
``metadata.tfrecord`` one Example with the three ``*_data_shape`` int64 vectors.  This module speaks both published
formats directly:

* TFRecord framing — per record: uint64 length, uint32 masked CRC-32C of those 8 bytes, the payload, uint32 masked
  CRC-32C of the payload; little endian; mask(crc) = rotr(crc, 15) + 0xa282ead8.  The CRC runs in the native library
  (``hyp_crc32c``, host code).
* tf.train.Example — protobuf wire format of
      Example{Features features = 1}   Features{map<string, Feature> feature = 1}
      Feature{oneof: BytesList bytes_list = 1 | FloatList float_list = 2 | Int64List int64_list = 3}
      BytesList{repeated bytes value = 1}  FloatList{repeated float value = 1 [packed]}  Int64List{repeated int64 value = 1 [packed]}
  written with packed repeated fields and sorted map keys; the reader accepts packed and unpacked fields, any key order.
"""
import ctypes
import gzip
import struct

import numpy

from hypelcnn_b200 import _native as N

_MASK_DELTA = 0xA282EAD8


def crc32c(data, crc=0):
    c = ctypes.c_uint32(crc)
    buf = (ctypes.c_char * len(data)).from_buffer_copy(data) if len(data) else None
    N.check(N.lib().hyp_crc32c(buf, len(data), ctypes.byref(c)))
    return c.value


def masked_crc32c(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------ framing
class TFRecordWriter:
    """Same role as tf.io.TFRecordWriter(path, options=GZIP or None): write(bytes), close(); context manager."""

    def __init__(self, path, compressed=False):
        self._f = gzip.open(path, "wb") if compressed else open(path, "wb")

    def write(self, record):
        header = struct.pack("<Q", len(record))
        self._f.write(header + struct.pack("<I", masked_crc32c(header)) + record + struct.pack("<I", masked_crc32c(record)))

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def iter_records(path, compressed=None, verify=True):
    """Yield the payloads of a TFRecord file.  compressed=None sniffs the GZIP magic.  A bad checksum or a truncated
    record raises ValueError (tf_record_iterator raises DataLossError)."""
    if compressed is None:
        with open(path, "rb") as probe:
            compressed = probe.read(2) == b"\x1f\x8b"
    with (gzip.open(path, "rb") if compressed else open(path, "rb")) as f:
        index = 0
        while True:
            header = f.read(12)
            if not header:
                return
            if len(header) < 12:
                raise ValueError(f"{path}: truncated header of record {index}")
            (length,), (hcrc,) = struct.unpack("<Q", header[:8]), struct.unpack("<I", header[8:])
            if verify and hcrc != masked_crc32c(header[:8]):
                raise ValueError(f"{path}: corrupted length of record {index}")
            body = f.read(length + 4)
            if len(body) < length + 4:
                raise ValueError(f"{path}: truncated record {index}")
            if verify and struct.unpack("<I", body[length:])[0] != masked_crc32c(body[:length]):
                raise ValueError(f"{path}: corrupted record {index}")
            yield body[:length]
            index += 1


# ------------------------------------------------------------------------------------------ protobuf wire format
def _varint(value):
    value &= 0xFFFFFFFFFFFFFFFF          # int64: negative numbers take ten bytes
    out = bytearray()
    while True:
        byte = value & 0x7F
        value >>= 7
        if value:
            out.append(byte | 0x80)
        else:
            out.append(byte)
            return bytes(out)


def _read_varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint longer than ten bytes")


def _delimited(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _fields(buf):
    """Yield (field number, wire type, value) of one message; value is an int (varint / fixed) or a memoryview."""
    buf = memoryview(buf)
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            value, pos = _read_varint(buf, pos)
        elif wire == 1:
            value, pos = buf[pos:pos + 8], pos + 8
        elif wire == 2:
            n, pos = _read_varint(buf, pos)
            value, pos = buf[pos:pos + n], pos + n
            if len(value) < n:
                raise ValueError("truncated protobuf message")
        elif wire == 5:
            value, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wire}")
        yield field, wire, value


def encode_example(features):
    """{name: values} -> serialized tf.train.Example.  float arrays -> float_list, integer arrays -> int64_list,
    bytes / str / lists of them -> bytes_list."""
    entries = []
    for name in sorted(features):
        value = features[name]
        if isinstance(value, (bytes, str)):
            value = [value]
        if isinstance(value, (list, tuple)) and value and isinstance(value[0], (bytes, str)):
            items = b"".join(_delimited(1, v.encode() if isinstance(v, str) else v) for v in value)
            feature = _delimited(1, items)
        else:
            arr = numpy.asarray(value)
            if arr.dtype.kind == "f":
                feature = _delimited(2, _delimited(1, arr.astype("<f4").reshape(-1).tobytes()))
            elif arr.dtype.kind in "iub":
                feature = _delimited(3, _delimited(1, b"".join(_varint(int(v)) for v in arr.reshape(-1))))
            else:
                raise TypeError(f"feature {name!r}: unsupported dtype {arr.dtype}")
        entries.append(_delimited(1, _delimited(1, name.encode()) + _delimited(2, feature)))
    return _delimited(1, b"".join(entries))


def _decode_feature(buf):
    for field, wire, value in _fields(buf):
        if wire != 2:
            continue
        if field == 1:      # BytesList
            return [bytes(v) for f, w, v in _fields(value) if f == 1 and w == 2]
        if field == 2:      # FloatList: packed (one blob) and / or unpacked fixed32 entries
            parts = [numpy.frombuffer(v, dtype="<f4") for f, w, v in _fields(value) if f == 1 and w in (2, 5)]
            return numpy.concatenate(parts) if parts else numpy.zeros(0, numpy.float32)
        if field == 3:      # Int64List
            out = []
            for f, w, v in _fields(value):
                if f != 1:
                    continue
                if w == 0:
                    out.append(v)
                elif w == 2:
                    pos = 0
                    while pos < len(v):
                        x, pos = _read_varint(v, pos)
                        out.append(x)
            arr = numpy.array(out, dtype=numpy.uint64).astype(numpy.int64) if out else numpy.zeros(0, numpy.int64)
            return arr
    return None                # Feature with no kind set


def decode_example(record):
    """serialized tf.train.Example -> {name: float32 array | int64 array | list of bytes}."""
    out = {}
    for field, wire, features in _fields(record):
        if field != 1 or wire != 2:
            continue
        for f, w, entry in _fields(features):
            if f != 1 or w != 2:
                continue
            key, feature = None, None
            for ef, ew, ev in _fields(entry):
                if ef == 1 and ew == 2:
                    key = bytes(ev).decode()
                elif ef == 2 and ew == 2:
                    feature = ev
            if key is not None:
                out[key] = _decode_feature(feature) if feature is not None else None
    return out
