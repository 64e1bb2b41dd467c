"""GeneratorImporter on the GPU.

Reference: importer/GeneratorImporter.py:14-97 — a lazy importer: patches are cut from the
scene when a batch is requested (tf.data.from_generator, one Python call per pixel).  Here
each batch is one hyp_gather_patches launch over a slice of the target list, so a full
scene (664 845 pixels for GRSS2013) never needs the 18.9 GB materialised array.
"""
import time
from collections import namedtuple

import numpy
import torch

from hypelcnn_b200.common.common_nn_ops import get_loader_from_name
from hypelcnn_b200.importer.DataImporter import DataImporter

GeneratorDataTensor = namedtuple("GeneratorDataTensor", ["dataset"])
GeneratorDataInfo = namedtuple("GeneratorDataInfo", ["data", "targets", "loader", "dataset"])
GeneratorSpecialData = namedtuple("GeneratorSpecialData", ["shape", "size"])


class LazyPatchDataset:
    """(images, labels) pair whose images are gathered per slice.  Supports the slicing /
    index_select the DeviceBatchIterator performs."""

    class _Images:
        def __init__(self, data_set, targets):
            self.data_set, self.targets = data_set, targets
            self.shape = tuple([targets.shape[0]] + list(data_set.get_data_shape()))
            self.device = data_set.device

        def __getitem__(self, sl):
            return self.data_set.get_data_points(self.targets[sl])

        def index_select(self, dim, idx):
            return self.data_set.get_data_points(self.targets.index_select(0, idx))

    def __init__(self, data_set, targets, class_count):
        t = torch.from_numpy(numpy.ascontiguousarray(numpy.asarray(targets).astype(numpy.int32))).to(data_set.device)
        self.images = self._Images(data_set, t)
        labels = t[:, 2].clamp(0, class_count - 1).long() if t.shape[0] else torch.zeros(0, dtype=torch.long, device=t.device)
        self.labels = torch.zeros((t.shape[0], class_count), dtype=torch.uint8, device=t.device)
        if t.shape[0]:
            self.labels.scatter_(1, labels.unsqueeze(1), 1)

    def __iter__(self):
        return iter((self.images, self.labels))

    def __getitem__(self, i):
        return (self.images, self.labels)[i]


class GeneratorImporter(DataImporter):

    def read_data_set(self, loader_name, path, train_data_ratio, test_data_ratio, neighborhood, normalize):
        start_time = time.time()
        loader = get_loader_from_name(loader_name, path)
        data_set = loader.load_data(neighborhood, normalize)
        sample_set = loader.load_samples(train_data_ratio, test_data_ratio)

        def info(targets):
            shape = numpy.concatenate(([targets.shape[0]], data_set.get_data_shape()))
            return GeneratorDataInfo(data=GeneratorSpecialData(shape=shape, size=numpy.prod(shape)), targets=targets,
                                     loader=loader, dataset=data_set)
        print(f"Loaded dataset({time.time() - start_time:.3f} sec)")
        return info(sample_set.training_targets), info(sample_set.test_targets), info(sample_set.validation_targets), \
            data_set.shadow_creator_dict, loader.get_class_count(), data_set.get_scene_shape(), \
            loader.get_samples_color_list()

    def convert_data_to_tensor(self, test_data_with_labels, training_data_with_labels, validation_data_with_labels,
                               class_range):
        n = class_range.stop
        mk = lambda d: GeneratorDataTensor(dataset=LazyPatchDataset(d.dataset, d.targets, n))  # noqa: E731
        return mk(test_data_with_labels), mk(training_data_with_labels), mk(validation_data_with_labels)

    def init_tensors(self, session, tensor, nn_params):
        nn_params.input_iterator.reset()

    def requires_separate_validation_branch(self):
        return True
