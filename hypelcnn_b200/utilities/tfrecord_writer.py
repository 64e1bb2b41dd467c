"""Writes the patch sets the TFRecordImporter reads — same function names, arguments and files as the reference's
utilities/tfrecord_writer.py:45-81, on hypelcnn_b200.utilities.tfrecord_io instead of TensorFlow."""
import os

import numpy

from hypelcnn_b200.utilities.tfrecord_io import TFRecordWriter, encode_example


def _host(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else numpy.asarray(a)


def write_to_tfrecord(train_filename, data, labels, compressed):
    """One Example{label, image} per patch; data [N, P, P, C] float32 (host or device), labels [N] integer."""
    data, labels = _host(data), _host(labels)
    with TFRecordWriter(train_filename, compressed) as writer:
        for i in range(len(data)):
            writer.write(encode_example({"label": numpy.asarray([labels[i]], dtype=numpy.int64),
                                         "image": numpy.asarray(data[i], dtype=numpy.float32).reshape(-1)}))


def write_metadata_record(metadata_filename, training_data, testing_data, validation_data):
    """metadata.tfrecord: the full shapes [N, P, P, C] of the three splits as int64 vectors."""
    with TFRecordWriter(metadata_filename) as writer:
        writer.write(encode_example({"training_data_shape": numpy.asarray(training_data.shape, dtype=numpy.int64),
                                     "testing_data_shape": numpy.asarray(testing_data.shape, dtype=numpy.int64),
                                     "validation_data_shape": numpy.asarray(validation_data.shape, dtype=numpy.int64)}))


def write_data_set(target_path, training, test, validation, compressed=False):
    """What the reference's main() does after InMemoryImporter.read_data_set (:17-38): metadata + the three splits.
    training / test / validation: objects with .data and .labels (the importer's Target tuples)."""
    os.makedirs(target_path, exist_ok=True)
    write_metadata_record(os.path.join(target_path, "metadata.tfrecord"), training.data, test.data, validation.data)
    for name, split in (("training", training), ("test", test), ("validation", validation)):
        write_to_tfrecord(os.path.join(target_path, name + ".tfrecord"), split.data, split.labels, compressed)
