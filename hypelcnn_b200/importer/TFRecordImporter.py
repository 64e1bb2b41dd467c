"""TFRecordImporter: patches that were cut once and stored as TFRecord files (reference:
importer/TFRecordImporter.py:16-72).  read_data_set returns, like the reference, only the shapes (from
metadata.tfrecord) and the file paths; the records are parsed when the tensors are built — straight into one pinned
host array per split and from there to HBM in a single copy, where the same DeviceBatchIterator as for the in-memory
importer batches them (the reference streams the file through tf.data for every epoch; 180 GB of HBM hold the
reference's largest split, 18.9 GB, many times over)."""
from collections import namedtuple

import numpy
import torch

from hypelcnn_b200.common.common_nn_ops import get_loader_from_name
from hypelcnn_b200.importer.DataImporter import DataImporter
from hypelcnn_b200.utilities.tfrecord_io import decode_example, iter_records

# shard = (rank, world) of a data-parallel run (hypelcnn_b200/parallel.shard_training_data): rows rank::world of the file
TFRecordDataInfo = namedtuple("TFRecordDataInfo", ["data", "path", "shard"], defaults=[None])
TFRecordDataTensor = namedtuple("TFRecordDataTensor", ["dataset", "path_placeholder"])
TFRecordSpecialData = namedtuple("TFRecordSpecialData", ["shape"])


def read_metadata(path):
    """metadata.tfrecord -> (training, testing, validation) shape vectors (the last record wins, as in :24-31)."""
    shapes = None
    for record in iter_records(path):
        ex = decode_example(record)
        shapes = tuple(numpy.asarray(ex[k + "_data_shape"], dtype=numpy.int64) for k in ("training", "testing", "validation"))
    if shapes is None:
        raise ValueError(f"{path}: no metadata record")
    return shapes


def load_split(path, sample_shape, class_count, max_records=None, pin=False):
    """Parse a split file into (images [N, *sample_shape] float32, labels [N] int64) host arrays.  Every record must
    carry exactly prod(sample_shape) floats and one label in [0, class_count) (FixedLenFeature semantics, :43-47)."""
    want = int(numpy.prod(sample_shape))
    images, labels = [], []
    for i, record in enumerate(iter_records(path)):
        if max_records is not None and i >= max_records:
            break
        ex = decode_example(record)
        image, label = ex.get("image"), ex.get("label")
        if image is None or label is None or image.size != want or label.size != 1:
            raise ValueError(f"{path}: record {i} does not hold label[1] and image[{want}]")
        if not 0 <= int(label[0]) < class_count:
            raise ValueError(f"{path}: record {i} has label {int(label[0])} outside [0, {class_count})")
        images.append(image)
        labels.append(int(label[0]))
    n = len(images)
    out = torch.empty((n, *[int(s) for s in sample_shape]), dtype=torch.float32, pin_memory=pin and n > 0)
    flat = out.view(n, want).numpy() if n else None
    for i, image in enumerate(images):
        flat[i] = image
    return out, torch.tensor(labels, dtype=torch.int64)


class TFRecordImporter(DataImporter):

    def read_data_set(self, loader_name, path, train_data_ratio, test_data_ratio, neighborhood, normalize):
        loader = get_loader_from_name(loader_name, path)
        base = loader.get_model_base_dir()
        training_shape, testing_shape, validation_shape = read_metadata(base + "metadata.tfrecord")
        return (TFRecordDataInfo(data=TFRecordSpecialData(training_shape), path=base + "training.tfrecord"),
                TFRecordDataInfo(data=TFRecordSpecialData(testing_shape), path=base + "test.tfrecord"),
                TFRecordDataInfo(data=TFRecordSpecialData(validation_shape), path=base + "validation.tfrecord"),
                None, loader.get_class_count(), None, loader.get_samples_color_list())

    def _device_split(self, info, class_count, device):
        """The split file of ``info`` in HBM as (images, one-hot labels); parsed once per (file, shard)."""
        cache = self.__dict__.setdefault("_splits", {})
        key = (info.path, info.shard, class_count)
        if key not in cache:
            images, labels = load_split(info.path, info.data.shape[1:4], class_count, pin=True)
            if info.shard is not None:
                rank, world = info.shard
                rows = (labels.shape[0] // world) * world
                images, labels = images[rank:rows:world].contiguous(), labels[rank:rows:world].contiguous()
            images = images.to(device, non_blocking=True)
            one_hot = torch.zeros((labels.shape[0], class_count), dtype=torch.uint8, device=device)
            if labels.shape[0]:
                one_hot.scatter_(1, labels.to(device).unsqueeze(1), 1)
            cache[key] = (images, one_hot)
        return cache[key]

    def convert_data_to_tensor(self, test_data_with_labels, training_data_with_labels, validation_data_with_labels,
                               class_range):
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: hypelcnn_b200 has no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
        testing = self._device_split(test_data_with_labels, class_range.stop, device)
        training = self._device_split(training_data_with_labels, class_range.stop, device)
        # like the reference (:66-68) the third tensor is built over the TESTING file; which file a branch really reads
        # is decided by init_tensors from the branch's own data_with_labels.path (the reference's path placeholder)
        return (TFRecordDataTensor(dataset=testing, path_placeholder=test_data_with_labels.path),
                TFRecordDataTensor(dataset=training, path_placeholder=training_data_with_labels.path),
                TFRecordDataTensor(dataset=testing, path_placeholder=test_data_with_labels.path))

    def init_tensors(self, session, tensor, nn_params):
        """Reference :70-72 feeds ``nn_params.data_with_labels.path`` into the path placeholder and re-initialises the
        iterator: the validation branch — built over the testing tensor — thereby reads validation.tfrecord.  Here the
        branch's iterator is pointed at the (cached) HBM copy of that file, then rewound."""
        iterator = nn_params.input_iterator
        info = getattr(nn_params, "data_with_labels", None)
        if info is not None and getattr(info, "path", None) is not None:
            batches = iterator
            while hasattr(batches, "inner"):                     # AugmentingIterator and friends wrap the batcher
                batches = batches.inner
            class_count = batches.labels.shape[1]
            batches.images, batches.labels = self._device_split(info, class_count, batches.images.device)
        iterator.reset() if hasattr(iterator, "reset") else iterator.initializer()

    def requires_separate_validation_branch(self):
        return False
