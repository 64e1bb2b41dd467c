"""GPU leg of the TFRecord path: patches cut by the in-memory importer are written with tfrecord_writer, read back by
the TFRecordImporter (importer/TFRecordImporter.py:16-72) and must arrive in HBM bit for bit, ready for create_graph."""
import os

import numpy
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tfrecord_importer_round_trip_and_training(tmp_path):
    from hypelcnn_b200.common import common_nn_ops as ops
    from hypelcnn_b200.utilities.tfrecord_writer import write_data_set
    from tests.util import ALG
    mem = ops.get_importer_from_name("InMemoryImporter")
    train, test, val, _, class_range, _, _ = mem.read_data_set(
        "SyntheticGRSS2013DataLoader", "synthetic:H=20,W=24,samples=96", 1.0, 0.25, 1, True)
    base = str(tmp_path) + os.sep
    write_data_set(base, train, test, val, compressed=True)
    assert sorted(os.listdir(base)) == ["metadata.tfrecord", "test.tfrecord", "training.tfrecord", "validation.tfrecord"]
    imp = ops.get_importer_from_name("TFRecordImporter")
    tr_info, te_info, va_info, shadow, classes, scene_shape, colors = imp.read_data_set(
        "SyntheticGRSS2013DataLoader", base, 1.0, 0.25, 1, True)
    assert shadow is None and scene_shape is None and classes == class_range and colors.shape == (15, 3)
    assert tr_info.data.shape.tolist() == list(train.data.shape) and tr_info.path == base + "training.tfrecord"
    assert te_info.data.shape.tolist() == list(test.data.shape) and va_info.data.shape.tolist() == list(val.data.shape)
    assert not imp.requires_separate_validation_branch()
    testing_tensor, training_tensor, validation_tensor = imp.convert_data_to_tensor(te_info, tr_info, va_info, classes)
    images, one_hot = training_tensor.dataset
    assert images.is_cuda and torch.equal(images, train.data)                       # bit-exact through the file
    assert torch.equal(one_hot.argmax(dim=1).to(torch.uint8), train.labels) and int(one_hot.sum()) == train.labels.numel()
    assert torch.equal(testing_tensor.dataset[0], test.data) and validation_tensor.dataset is testing_tensor.dataset
    model = ops.get_model_from_name("HYPELCNNModel")
    alg = {**ALG, "filter_count": 32, "batch_size": 24}
    ce, lr, testing_nn, train_nn, validation_nn, train_step = ops.create_graph(
        training_tensor.dataset, testing_tensor.dataset, validation_tensor.dataset, classes, 24, 1000, "/gpu:0",
        None, alg, model, None, imp.requires_separate_validation_branch)
    imp.init_tensors(None, training_tensor, train_nn)
    for _ in range(3):
        train_step.run()
    assert train_step.global_step == 3 and numpy.isfinite(float(ce()[0]))
    # the validation branch is built over the TESTING tensor; init_tensors points it at validation.tfrecord
    # (reference importer/TFRecordImporter.py:70-72 feeds nn_params.data_with_labels.path into the path placeholder)
    validation_nn.data_with_labels, testing_nn.data_with_labels = va_info, te_info
    imp.init_tensors(None, validation_tensor, validation_nn)
    assert torch.equal(validation_nn.input_iterator.images, val.data)
    assert torch.equal(validation_nn.input_iterator.labels.argmax(dim=1).to(torch.uint8), val.labels)
    imp.init_tensors(None, testing_tensor, testing_nn)
    assert torch.equal(testing_nn.input_iterator.images, test.data)
    # data parallel: the training file is strided over the ranks when it is parsed
    from hypelcnn_b200 import parallel
    shard = parallel.shard_training_data(tr_info, 1, 2)
    rows = (train.data.shape[0] // 2) * 2
    assert int(shard.data.shape[0]) == rows // 2 and shard.shard == (1, 2)
    images_1, _ = imp._device_split(shard, classes.stop, images.device)
    assert torch.equal(images_1, train.data[1:rows:2])
