"""InMemoryImporter on the GPU.

Reference: importer/InMemoryImporter.py:14-86.  The per-pixel Python loop of
_get_data_with_labels (:27-38) is one hyp_gather_patches launch; the materialised
[N,P,P,C] fp32 array lives in HBM instead of host memory, and "placeholders + feed_dict"
(:56-83) become device-resident (images, one-hot labels) pairs.
"""
import time
from collections import namedtuple

import numpy
import torch

from hypelcnn_b200.common.common_nn_ops import get_loader_from_name
from hypelcnn_b200.importer.DataImporter import DataImporter

Target = namedtuple("Target", ["data", "labels"])
InMemoryDataTensor = namedtuple("InMemoryDataTensor", ["dataset", "x", "y_"])


class InMemoryImporter(DataImporter):

    @staticmethod
    def _get_data_with_labels(targets, loader, data_set):
        targets = numpy.asarray(targets)
        if targets.shape[0] == 0:
            shape = [0] + list(data_set.get_data_shape())
            return Target(data=torch.zeros(shape, dtype=torch.float32, device=data_set.device),
                          labels=torch.zeros(0, dtype=torch.uint8, device=data_set.device))
        data = data_set.get_data_points(targets)
        labels = torch.from_numpy(targets[:, 2].astype(numpy.uint8)).to(data.device)
        return Target(data=data, labels=labels)

    def read_data_set(self, loader_name, path, train_data_ratio, test_data_ratio, neighborhood, normalize):
        t0 = time.perf_counter()
        loader = get_loader_from_name(loader_name, path)
        scene = loader.load_data(neighborhood, normalize)
        samples = loader.load_samples(train_data_ratio, test_data_ratio)
        # one gather launch per split; the patches stay in HBM
        split = {name: self._get_data_with_labels(getattr(samples, name + "_targets"), loader, scene)
                 for name in ("training", "test", "validation")}
        torch.cuda.synchronize()
        counts = ", ".join(f"{name} {t.labels.shape[0]}" for name, t in split.items())
        print(f"Gathered {counts} patches in {time.perf_counter() - t0:.3f} s")
        # the 7-tuple the reference's callers unpack (importer/InMemoryImporter.py:51-52)
        return (split["training"], split["test"], split["validation"], scene.shadow_creator_dict,
                loader.get_class_count(), scene.get_scene_shape(), loader.get_samples_color_list())

    @staticmethod
    def _one_hot(labels, class_range):
        # tf.one_hot(y_, class_range.stop, dtype=tf.uint8) (:24) — layout only, no arithmetic
        out = torch.zeros((labels.shape[0], class_range.stop), dtype=torch.uint8, device=labels.device)
        if labels.shape[0]:
            out.scatter_(1, labels.long().unsqueeze(1), 1)
        return out

    def convert_data_to_tensor(self, test_data_with_labels, training_data_with_labels, validation_data_with_labels,
                               class_range):
        training = (training_data_with_labels.data, self._one_hot(training_data_with_labels.labels, class_range))
        testing = (test_data_with_labels.data, self._one_hot(test_data_with_labels.labels, class_range))
        # the reference returns the TESTING tensors for validation too (:76-78); the separate
        # validation branch is fed through init_tensors with the validation arrays
        return InMemoryDataTensor(dataset=testing, x=testing[0], y_=test_data_with_labels.labels), \
            InMemoryDataTensor(dataset=training, x=training[0], y_=training_data_with_labels.labels), \
            InMemoryDataTensor(dataset=testing, x=testing[0], y_=test_data_with_labels.labels)

    def init_tensors(self, session, tensor, nn_params):
        """Reference :80-83: (re)initialise the iterator with nn_params.data_with_labels."""
        it = nn_params.input_iterator
        it = getattr(it, "inner", it)       # an AugmentingIterator wraps the batch iterator that holds the data
        d = nn_params.data_with_labels
        if d is not None:
            it.images = d.data
            it.labels = self._one_hot(d.labels, range(0, tensor.dataset[1].shape[1])) if d.labels.dim() == 1 else d.labels
        it.reset()

    def requires_separate_validation_branch(self):
        return True
