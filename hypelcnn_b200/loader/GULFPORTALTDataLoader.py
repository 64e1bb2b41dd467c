"""MUUFL Gulfport with the shadow experiment's splits (reference: loader/GULFPORTALTDataLoader.py): training targets come
from the lit area only, every labelled pixel under the shadow map joins the validation list; the data set carries
the shadow augmenters; the scene can be served in its GAN-shadowed / de-shadowed renditions or as a per-sample mix."""
import numpy

from hypelcnn_b200.common.common_nn_ops import INVALID_TARGET_VALUE, DataSet
from hypelcnn_b200.loader.DataLoader import LoadingMode, SampleSet
from hypelcnn_b200.loader.GULFPORTDataLoader import GULFPORTDataLoader


class MultiDataSet(DataSet):
    """Several renditions of one scene; every requested point is served from a randomly chosen one (reference :17-42,
    one ``random.randint`` per point there, one vectorised draw per batch here).  Shape queries go to the first."""

    def __init__(self, *data_sets) -> None:
        super().__init__()
        self._data_sets = data_sets
        self._primary_data_set = data_sets[0]
        self.lidar, self.casi = self._primary_data_set.lidar, self._primary_data_set.casi
        self.neighborhood = self._primary_data_set.neighborhood
        self.shadow_creator_dict = self._primary_data_set.shadow_creator_dict
        self.device = getattr(self._primary_data_set, "device", None)
        self._cmin = getattr(self._primary_data_set, "_cmin", None)      # read by load_shadow_map_common
        self._rng = numpy.random.default_rng(1234)

    def get_data_shape(self):
        return self._primary_data_set.get_data_shape()

    def get_casi_band_count(self):
        return self._primary_data_set.get_casi_band_count()

    def get_scene_shape(self):
        return self._primary_data_set.get_scene_shape()

    def get_unnormalized_casi_dtype(self):
        return self._primary_data_set.get_unnormalized_casi_dtype()

    def get_data_point(self, point_x, point_y):
        chosen = int(self._rng.integers(0, len(self._data_sets)))
        return self._data_sets[chosen].get_data_point(point_x=point_x, point_y=point_y)

    def get_data_points(self, targets_xy, out=None):
        """Batched: one gather launch per rendition over the points that drew it."""
        import torch
        targets = torch.as_tensor(numpy.asarray(targets_xy)[:, :2].astype(numpy.int32)) \
            if not isinstance(targets_xy, torch.Tensor) else targets_xy[:, :2]
        chosen = torch.as_tensor(self._rng.integers(0, len(self._data_sets), targets.shape[0]))
        result = None
        for index, data_set in enumerate(self._data_sets):
            rows = torch.nonzero(chosen == index).reshape(-1)
            if rows.numel() == 0:
                continue
            part = data_set.get_data_points(targets[rows.to(targets.device)])
            if result is None:
                result = part.new_empty((targets.shape[0],) + tuple(part.shape[1:])) if out is None else out
            result[rows.to(part.device)] = part
        return result


class GULFPORTALTDataLoader(GULFPORTDataLoader):
    SHADOW_MAP_FILE = "muulf_shadow_map.tif"
    GAN_CHECKPOINTS = {"cycle_gan": "shadow_gen_model/cycle_gan/model.ckpt-3000",
                       "dcl_gan": "shadow_gen_model/dcl_gan/model.ckpt-3000",
                       "dcl_cycle_gan": "shadow_gen_model/dcl_cycle_gan/v1/model.ckpt-3000"}

    def __init__(self, base_dir):
        super().__init__(base_dir)
        self._load_mode = LoadingMode.ORIGINAL

    def _rendition(self, mode, neighborhood, normalize, original=None):
        """muulf_hsi.tif, or muulf_hsi_shadowed.tif / muulf_hsi_deshadowed.tif normalised with the ORIGINAL range."""
        if mode is LoadingMode.ORIGINAL:
            return self._load_data_utility(self.HSI_FILE + self.FILE_EXT, self.LIDAR_FILE + self.FILE_EXT, neighborhood,
                                           normalize)
        return self._load_data_utility(self.HSI_FILE + "_" + mode.value + self.FILE_EXT, self.LIDAR_FILE + self.FILE_EXT,
                                       neighborhood, normalize, casi_min=original.casi_min, casi_max=original.casi_max)

    def load_data(self, neighborhood, normalize):
        original = self._rendition(LoadingMode.ORIGINAL, neighborhood, normalize)
        mode = self._load_mode
        if mode in (LoadingMode.SHADOWED, LoadingMode.DESHADOWED):
            data_set = self._rendition(mode, neighborhood, normalize, original)
        elif mode is LoadingMode.MIXED:
            shadowed = self._rendition(LoadingMode.SHADOWED, neighborhood, normalize, original)
            self._rendition(LoadingMode.DESHADOWED, neighborhood, normalize, original)   # read (and checked) like the reference
            data_set = MultiDataSet(original, shadowed, shadowed, shadowed)   # the reference mixes 1 : 3, de-shadowed unused
        else:
            data_set = original
        return self.attach_shadow_creators(data_set, neighborhood)

    def load_samples(self, train_data_ratio, test_data_ratio):
        shadow_map, _ = self.load_shadow_map(0, None)
        targets = self.read_raster('muulf_gt_shadow_corrected.tif')
        in_shadow = shadow_map.astype(bool)
        under_shadow = numpy.where(in_shadow, targets, INVALID_TARGET_VALUE).astype(targets.dtype)
        in_clear_area = numpy.where(in_shadow, INVALID_TARGET_VALUE, targets).astype(targets.dtype)
        train_set, validation_set = self.split_training_and_validation(self._convert_targets_aux(in_clear_area),
                                                                       train_data_ratio)
        test_set = numpy.empty([0, train_set.shape[1]])                      # no test list in this experiment
        validation_set = numpy.vstack([validation_set, self._convert_targets_aux(under_shadow)])
        return SampleSet(training_targets=train_set, test_targets=test_set, validation_targets=validation_set)
