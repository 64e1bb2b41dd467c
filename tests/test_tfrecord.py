"""CPU tests of the TFRecord / tf.train.Example path (importer/TFRecordImporter.py:16-72, utilities/tfrecord_writer.py:
45-81 in the reference, which go through TensorFlow).  TensorFlow is absent here, so the formats are pinned against
independent implementations that ARE installed: the CRC-32C vectors of RFC 3720, TensorBoard's TFRecord writer / reader
(tensorboard.summary.writer.record_writer, tensorboard.compat.tensorflow_stub.pywrap_tensorflow) and the official
protobuf runtime with the Example schema declared at run time."""
import gzip
import os
import struct

import numpy
import pytest

from hypelcnn_b200.utilities import tfrecord_io as T
from hypelcnn_b200.utilities.tfrecord_writer import write_metadata_record, write_to_tfrecord


def test_crc32c_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(bytes(32)) == 0x8A9136AA                       # RFC 3720 B.4
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    data = os.urandom(100003)
    assert T.crc32c(data[40001:], T.crc32c(data[:40001])) == T.crc32c(data)   # incremental update
    from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import masked_crc32c
    for n in (0, 1, 7, 8, 9, 4097):
        assert T.masked_crc32c(data[:n]) == masked_crc32c(data[:n])


def test_framing_interoperates_with_tensorboards_record_io(tmp_path):
    from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import PyRecordReader_New
    from tensorboard.summary.writer.record_writer import RecordWriter
    payloads = [b"", b"x", os.urandom(1000), os.urandom(70000)]
    mine = str(tmp_path / "mine.tfrecord")
    with T.TFRecordWriter(mine) as w:
        for p in payloads:
            w.write(p)
    reader = PyRecordReader_New(mine)
    got = []
    while True:
        try:
            reader.GetNext()
        except Exception:
            break
        got.append(bytes(reader.record()))
    assert got == payloads
    theirs = str(tmp_path / "theirs.tfrecord")
    rw = RecordWriter(open(theirs, "wb"))
    for p in payloads:
        rw.write(p)
    rw.close()
    assert list(T.iter_records(theirs)) == payloads
    assert open(mine, "rb").read() == open(theirs, "rb").read()             # byte-identical files


def _example_classes():
    """tf.train.Example declared through descriptor_pb2 (tensorflow/core/example/{example,feature}.proto)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="hyp_test_example.proto", package="hyp_test", syntax="proto3")

    def msg(name, *fields):
        m = fd.message_type.add(name=name)
        for fname, number, ftype, label, type_name, extra in fields:
            f = m.field.add(name=fname, number=number, type=ftype, label=label)
            if type_name:
                f.type_name = ".hyp_test." + type_name
            if extra == "packed":
                f.options.packed = True
            if extra == "oneof":
                f.oneof_index = 0
        return m
    msg("BytesList", ("value", 1, F.TYPE_BYTES, F.LABEL_REPEATED, None, None))
    msg("FloatList", ("value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, None, "packed"))
    msg("Int64List", ("value", 1, F.TYPE_INT64, F.LABEL_REPEATED, None, "packed"))
    feat = msg("Feature", ("bytes_list", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "BytesList", "oneof"),
               ("float_list", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "FloatList", "oneof"),
               ("int64_list", 3, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "Int64List", "oneof"))
    feat.oneof_decl.add(name="kind")
    feats = msg("Features", ("feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, "Features.FeatureEntry", None))
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".hyp_test.Feature")
    entry.options.map_entry = True
    msg("Example", ("features", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "Features", None))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("hyp_test.Example"))


def test_example_encoding_matches_the_protobuf_runtime():
    Example = _example_classes()
    rng = numpy.random.default_rng(0)
    image = rng.random(7 * 7 * 145, dtype=numpy.float32)
    mine = T.encode_example({"label": numpy.array([11]), "image": image})
    parsed = Example.FromString(mine)                                        # the official runtime reads our bytes
    assert list(parsed.features.feature["label"].int64_list.value) == [11]
    assert numpy.array_equal(numpy.array(parsed.features.feature["image"].float_list.value, dtype=numpy.float32), image)
    ref = Example()                                                          # ... and we read the runtime's bytes
    ref.features.feature["label"].int64_list.value.append(3)
    ref.features.feature["image"].float_list.value.extend(image.tolist())
    ref.features.feature["shape"].int64_list.value.extend([-5, 0, 2 ** 40])
    ref.features.feature["names"].bytes_list.value.extend([b"casi", b"lidar"])
    back = T.decode_example(ref.SerializeToString())
    assert back["label"].tolist() == [3] and numpy.array_equal(back["image"], image)
    assert back["shape"].tolist() == [-5, 0, 2 ** 40] and back["names"] == [b"casi", b"lidar"]
    again = Example.FromString(T.encode_example({k: back[k] for k in back}))  # round trip through both
    assert again == ref
    # unpacked repeated scalars (what some writers emit) are accepted too
    unpacked = T._delimited(1, T._delimited(1, T._delimited(1, b"v") + T._delimited(2, T._delimited(
        2, b"".join(T._varint((1 << 3) | 5) + struct.pack("<f", x) for x in (1.5, -2.0))))))
    assert T.decode_example(unpacked)["v"].tolist() == [1.5, -2.0]


@pytest.mark.parametrize("compressed", [False, True])
def test_writer_and_importer_round_trip(tmp_path, compressed):
    from hypelcnn_b200.importer.TFRecordImporter import load_split, read_metadata
    rng = numpy.random.default_rng(5)
    shape = (3, 3, 10)
    train = rng.random((17,) + shape, dtype=numpy.float32)
    test = rng.random((5,) + shape, dtype=numpy.float32)
    val = rng.random((0,) + shape, dtype=numpy.float32)                      # an empty split is legal
    labels = rng.integers(0, 4, 17)
    base = str(tmp_path) + os.sep
    write_metadata_record(base + "metadata.tfrecord", train, test, val)
    write_to_tfrecord(base + "training.tfrecord", train, labels, compressed)
    write_to_tfrecord(base + "validation.tfrecord", val, numpy.zeros(0, numpy.int64), compressed)
    shapes = read_metadata(base + "metadata.tfrecord")
    assert [s.tolist() for s in shapes] == [list(train.shape), list(test.shape), list(val.shape)]
    raw = open(base + "training.tfrecord", "rb").read()
    assert (raw[:2] == b"\x1f\x8b") == compressed
    images, got_labels = load_split(base + "training.tfrecord", shapes[0][1:4], 4)
    assert numpy.array_equal(images.numpy(), train) and got_labels.tolist() == labels.tolist()   # bit-exact floats
    images, got_labels = load_split(base + "validation.tfrecord", shape, 4)
    assert tuple(images.shape) == (0,) + shape and got_labels.numel() == 0
    assert load_split(base + "training.tfrecord", shape, 4, max_records=3)[0].shape[0] == 3
    with pytest.raises(ValueError):
        load_split(base + "training.tfrecord", (3, 3, 9), 4)                 # wrong image size
    with pytest.raises(ValueError):
        load_split(base + "training.tfrecord", shape, 2)                     # label outside the class range


def test_corruption_is_detected(tmp_path):
    path = str(tmp_path / "x.tfrecord")
    with T.TFRecordWriter(path) as w:
        w.write(b"hello world" * 10)
        w.write(b"second")
    good = open(path, "rb").read()
    for pos in (3, 9, 40, len(good) - 2):                                    # length, length crc, payload, payload crc
        bad = bytearray(good)
        bad[pos] ^= 0x10
        open(path, "wb").write(bytes(bad))
        with pytest.raises(ValueError):
            list(T.iter_records(path))
    open(path, "wb").write(good[:-3])
    with pytest.raises(ValueError):
        list(T.iter_records(path))
    open(path, "wb").write(good)
    assert [len(r) for r in T.iter_records(path)] == [110, 6]
    gz = str(tmp_path / "y.tfrecord")
    with gzip.open(gz, "wb") as f:
        f.write(good)
    assert [len(r) for r in T.iter_records(gz)] == [110, 6]                  # GZIP sniffed from the magic number
