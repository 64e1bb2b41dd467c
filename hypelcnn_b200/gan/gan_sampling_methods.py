"""Sample pairing for GAN training (reference: gan/gan_sampling_methods.py): which scene pixels become the rows of the
(normal, shadowed) matrices the GAN wrappers train on.  Same class names, constructor arguments and return order
``(normal_data_as_matrix, shadow_data_as_matrix)`` as the reference.

The reference walks the scene pixel by pixel and calls ``data_set.get_data_point`` once per pixel (:37-46, :69-78,
:150-160).  Here every sampler first builds the two target lists ``[N,2] = (x = column, y = row)`` with array operations
(``*_pair_targets``, host side, pinned against the reference's own module in tests/test_gan_samplers.py) and then fetches
all rows of a list with ONE ``data_set.get_data_points`` call — for ``BasicDataSet`` one launch of the patch-gather kernel,
the matrices stay in HBM as CUDA tensors.  A ``DataSet`` plug-in that only implements the per-point ``get_data_point``
of the ABC is read point by point and yields numpy matrices like the reference.
"""
from abc import ABC, abstractmethod

import numpy
from scipy import ndimage

from hypelcnn_b200.loader.DataLoader import SampleSet


class Sampler(ABC):
    @abstractmethod
    def get_sample_pairs(self, data_set, loader, shadow_map):
        pass


def _rows_to_targets(rows, cols):
    return numpy.stack([cols, rows], axis=1).astype(numpy.int32).reshape(-1, 2)


def _fetch(data_set, targets_xy, zero_rows=0):
    """All patches of a target list (+ ``zero_rows`` all-zero rows, see NeighborhoodBasedSampler) as one matrix."""
    shape = [int(s) for s in data_set.get_data_shape()]
    count = int(targets_xy.shape[0])
    batched = getattr(data_set, "get_data_points", None)
    if batched is not None and count > 0:
        matrix = batched(targets_xy)
        if zero_rows:
            pad = matrix.new_zeros([zero_rows] + shape) if hasattr(matrix, "new_zeros") else numpy.zeros(
                [zero_rows] + shape, numpy.float32)
            matrix = (numpy.concatenate if isinstance(matrix, numpy.ndarray) else _torch_cat)([matrix, pad])
        return matrix
    matrix = numpy.zeros([count + zero_rows] + shape, dtype=numpy.float32)
    for row, (x, y) in enumerate(targets_xy):
        matrix[row] = _as_numpy(data_set.get_data_point(int(x), int(y)))
    return matrix


def _torch_cat(parts):
    import torch
    return torch.cat(parts)


def _as_numpy(value):
    return value.detach().cpu().numpy() if hasattr(value, "detach") else value


def neighbourhood_pair_targets(shadow_map, neighborhood_size, margin):
    """Reference :21-51.  Shadowed pixels in scan order; "normal" pixels are the ring between ``margin`` and
    ``neighborhood_size`` city-block dilations of the shadow mask, trimmed to the shadow count.

    Returns (normal_targets, shadow_targets, normal_zero_rows).  The reference allocates ``sum(ring map)`` zero rows
    and fills those whose ring value == 1; with a mask dtype where the subtraction of the two dilations wraps (``margin``
    0 makes scipy dilate until nothing changes, so ring = dil(size) - ones) the sum exceeds the filled rows and all-zero
    rows reach the trimmed matrix — ``normal_zero_rows`` counts them so that quirk is reproduced, not hidden."""
    shadow_map = numpy.asarray(shadow_map)
    ring = ndimage.binary_dilation(shadow_map, iterations=neighborhood_size).astype(shadow_map.dtype) - \
        ndimage.binary_dilation(shadow_map, iterations=margin).astype(shadow_map.dtype)
    is_shadow = shadow_map == 1
    shadow_targets = _rows_to_targets(*numpy.nonzero(is_shadow))
    normal_targets = _rows_to_targets(*numpy.nonzero(~is_shadow & (ring == 1)))
    shadow_rows = int(numpy.sum(shadow_map, dtype=int))
    if shadow_rows != shadow_targets.shape[0]:
        raise ValueError("shadow_map must hold only 0 and 1")
    allocated = int(numpy.sum(ring, dtype=int))
    kept = min(normal_targets.shape[0], shadow_rows)
    zero_rows = max(0, min(allocated, shadow_rows) - kept)
    return normal_targets[:kept], shadow_targets, zero_rows


class NeighborhoodBasedSampler(Sampler):

    def __init__(self, neighborhood_size, margin) -> None:
        self._margin = margin
        self._neighborhood_size = neighborhood_size

    def get_sample_pairs(self, data_set, loader, shadow_map):
        normal_targets, shadow_targets, zero_rows = neighbourhood_pair_targets(shadow_map, self._neighborhood_size,
                                                                               self._margin)
        return _fetch(data_set, normal_targets, zero_rows), _fetch(data_set, shadow_targets)


def random_pair_targets(shadow_map, multiply_shadowed_data):
    """Reference :59-88.  Every pixel is either shadowed (== 1) or normal; the shadow list is repeated element-wise
    ``normal // shadow`` times (numpy.repeat order: a a b b ...) and the normal list trimmed to its length."""
    shadow_map = numpy.asarray(shadow_map)
    is_shadow = shadow_map == 1
    shadow_targets = _rows_to_targets(*numpy.nonzero(is_shadow))
    normal_targets = _rows_to_targets(*numpy.nonzero(~is_shadow))
    shadow_count = int(numpy.sum(shadow_map, dtype=int))
    if shadow_count != shadow_targets.shape[0]:
        raise ValueError("shadow_map must hold only 0 and 1")
    if multiply_shadowed_data:
        shadow_targets = numpy.repeat(shadow_targets, repeats=normal_targets.shape[0] // shadow_count, axis=0)
    return normal_targets[:shadow_targets.shape[0]], shadow_targets


class RandomBasedSampler(Sampler):

    def __init__(self, multiply_shadowed_data) -> None:
        self._multiply_shadowed_data = multiply_shadowed_data

    def get_sample_pairs(self, data_set, loader, shadow_map):
        normal_targets, shadow_targets = random_pair_targets(shadow_map, self._multiply_shadowed_data)
        return _fetch(data_set, normal_targets), _fetch(data_set, shadow_targets)


def target_pair_targets(all_targets, shadow_map, class_count, report=print):
    """Reference :116-187 (_get_targetbased_shadowed_normal_data).  Labelled targets ``[N,3] = (x, y, class)`` with
    class >= 0 are split per class into shadowed / normal points; for every class that has both, all its normal
    points are paired with its shadowed points cycled to the same count (each repeated ``normal // shadow`` times,
    then the first ``normal % shadow`` once more).  Classes ascend; (None, None) when no class has both kinds."""
    all_targets = numpy.asarray(all_targets)
    shadow_map = numpy.asarray(shadow_map)
    valid = all_targets[:, 2] >= 0
    in_shadow = numpy.zeros(all_targets.shape[0], bool)
    in_shadow[valid] = shadow_map[all_targets[valid, 1], all_targets[valid, 0]] == 1
    normal_parts, shadow_parts = [], []
    for target_key in range(class_count):
        of_class = valid & (all_targets[:, 2] == target_key)
        shadow_points = all_targets[of_class & in_shadow, :2]
        normal_points = all_targets[of_class & ~in_shadow, :2]
        if shadow_points.shape[0] == 0:
            continue
        if normal_points.shape[0] == 0:
            report(f"Target key is not found in read target image during "
                   f"target based sampling:{target_key}")
            continue
        multiplier, reminder = divmod(normal_points.shape[0], shadow_points.shape[0])
        normal_parts.append(normal_points)
        shadow_parts.append(numpy.concatenate([numpy.repeat(shadow_points, repeats=multiplier, axis=0),
                                               shadow_points[:reminder]]))
    if not normal_parts:
        return None, None
    return (numpy.concatenate(normal_parts).astype(numpy.int32), numpy.concatenate(shadow_parts).astype(numpy.int32))


def merge_sample_set_targets(samples):
    """Reference :118-123: the training targets, the test targets, or test stacked over training."""
    if samples.test_targets is None and samples.training_targets is not None:
        return samples.training_targets
    if samples.training_targets is None and samples.test_targets is not None:
        return samples.test_targets
    return numpy.vstack([samples.test_targets, samples.training_targets])


class TargetBasedSampler(Sampler):
    def __init__(self, margin):
        self._margin = margin

    def get_sample_pairs(self, data_set, loader, shadow_map):
        """Reference :95-113: targets come from the loader's ``shadow_gen_model/class_result.tif``; those within
        ``margin`` of the scene border (strict comparisons on both sides) are disabled by class -1."""
        samples = SampleSet(training_targets=numpy.array(loader.read_targets("shadow_gen_model/class_result.tif")),
                            test_targets=None, validation_targets=None)
        targets = samples.training_targets
        scene_shape = data_set.get_scene_shape()
        inside = (self._margin < targets[:, 1]) & (targets[:, 1] < scene_shape[0] - self._margin) & \
                 (self._margin < targets[:, 0]) & (targets[:, 0] < scene_shape[1] - self._margin)
        targets[~inside, 2] = -1
        return self._get_targetbased_shadowed_normal_data(data_set, loader, shadow_map, samples)

    @staticmethod
    def _get_targetbased_shadowed_normal_data(data_set, loader, shadow_map, samples):
        normal_targets, shadow_targets = target_pair_targets(merge_sample_set_targets(samples), shadow_map,
                                                             loader.get_class_count().stop)
        if normal_targets is None:
            return None, None
        return _fetch(data_set, normal_targets), _fetch(data_set, shadow_targets)


class DummySampler(Sampler):
    """Reference :191-201, the known-answer fixture: ``element_count`` constant patches, normal = coefficient x shadow."""

    def __init__(self, element_count, fill_value, coefficient):
        self._element_count = element_count
        self._fill_value = fill_value
        self._coefficient = coefficient

    def get_sample_pairs(self, data_set, loader, shadow_map):
        data_shape_info = data_set.get_data_shape()
        shadow_data_as_matrix = numpy.full(numpy.concatenate([[self._element_count], data_shape_info]),
                                           fill_value=self._fill_value, dtype=numpy.float32)
        return shadow_data_as_matrix * self._coefficient, shadow_data_as_matrix
