"""The feed-strategy contract: how labelled patches reach the training / evaluation loops.

Names and signatures are the reference's (importer/DataImporter.py:4-20) — the drop-in boundary.  In this engine the
"tensors" are CUDA tensors ([B, P, P, C] float32 patches, [B, classes] uint8 one-hot labels) or iterators over them; the
``session`` argument of the reference's TF1 API is accepted and ignored.
"""
import abc


class DataImporter(abc.ABC):

    @abc.abstractmethod
    def read_data_set(self, loader_name, path, train_data_ratio, test_data_ratio, neighborhood, normalize):
        """Instantiate the named loader, load the scene and the sample lists; -> the importer's data bundle
        (training / test / validation patches with labels, the loader, shadow information)."""
        raise NotImplementedError

    @abc.abstractmethod
    def convert_data_to_tensor(self, test_data_with_labels, training_data_with_labels, validation_data_with_labels,
                               class_range):
        """Move the three splits to the device and wrap them as the iterables the loops consume; labels become
        one-hot uint8 over ``class_range`` (reference: tf.one_hot(..., dtype=uint8))."""
        raise NotImplementedError

    @abc.abstractmethod
    def init_tensors(self, session, tensor, nn_params):
        """(Re-)initialise the iterators before a run; nothing to do for in-memory feeds."""
        raise NotImplementedError

    @abc.abstractmethod
    def requires_separate_validation_branch(self):
        """-> bool.  The reference's caller passes this METHOD un-called, i.e. always truthy
        (classify/train_for_classification.py:67) — callers here do the same."""
        raise NotImplementedError
