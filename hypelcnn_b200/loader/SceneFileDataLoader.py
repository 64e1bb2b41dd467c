"""Shared machinery of the loaders that read the reference's scene files (GeoTIFF rasters read with
hypelcnn_b200.utilities.tiff_io instead of tifffile).  A concrete loader is a table — directory, file names, class
count, colours, band range, generator checkpoints — plus whatever its scene needs beyond that; the cube goes to the
device once (``BasicDataSet``), everything after it (normalisation, padding, window slicing) happens in the gather kernel.
"""
import numpy

from hypelcnn_b200.common.common_nn_ops import (BasicDataSet, load_shadow_map_common, read_targets_from_image,
                                                shuffle_test_data_using_ratio, shuffle_training_data_using_ratio,
                                                shuffle_training_data_using_size)
from hypelcnn_b200.loader.DataLoader import DataLoader, SampleSet
from hypelcnn_b200.utilities.tiff_io import imread


class SceneFileDataLoader(DataLoader):
    DIRECTORY = None            # appended to base_dir, e.g. "/2013_DFTC/"
    CLASSES = 0
    COLORS = ()                 # one RGB row per class (the reference's get_samples_color_list)
    BAND_RANGE = (0, 0, 0)      # numpy.linspace(first nm, last nm, bands)
    SHADOW_MAP_FILE = None
    GAN_CHECKPOINTS = None      # {"cycle_gan": "shadow_gen_model/.../model.ckpt-N", ...}

    def __init__(self, base_dir):
        self.base_dir = base_dir

    # ------------------------------------------------------------------------------------------ constant tables
    def get_model_base_dir(self):
        return self.base_dir + self.DIRECTORY

    def get_class_count(self):
        return range(0, self.CLASSES)

    def get_samples_color_list(self):
        return numpy.array(self.COLORS, dtype=numpy.uint8).reshape(-1, 3)

    def get_band_measurements(self):
        first, last, count = self.BAND_RANGE
        return numpy.linspace(first, last, num=count)

    # ------------------------------------------------------------------------------------------------- file access
    def read_raster(self, name):
        return imread(self.get_model_base_dir() + name)

    def read_targets(self, target_image_path):
        """Label image -> [N,3] rows of (x, y, class) for the classes of this loader."""
        return read_targets_from_image(self.read_raster(target_image_path), self.get_class_count())

    def load_shadow_map(self, neighborhood, data_set):
        """-> (map padded by neighborhood, per-band lit / shadow ratio); None for scenes without a shadow map, like
        the reference's loaders whose method body is ``pass``."""
        if self.SHADOW_MAP_FILE is None:
            return None
        return load_shadow_map_common(data_set, neighborhood, self.read_raster(self.SHADOW_MAP_FILE))

    # ---------------------------------------------------------------------------------------------------- helpers
    def attach_shadow_creators(self, data_set, neighborhood):
        """data_set.shadow_creator_dict as the reference's loaders fill it: one frozen generator pair per GAN family,
        restored from this scene's ``shadow_gen_model`` checkpoints when training starts, plus the ratio augmenter."""
        from hypelcnn_b200.gan.gan_utilities import create_gan_struct, create_simple_shadow_struct
        from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANInferenceWrapper
        _, shadow_ratio = self.load_shadow_map(neighborhood, data_set)
        bands = data_set.get_casi_band_count()
        creators = {name: create_gan_struct(CycleGANInferenceWrapper(bands=bands), self.get_model_base_dir(), path)
                    for name, path in self.GAN_CHECKPOINTS.items()}
        creators["simple"] = create_simple_shadow_struct(shadow_ratio)
        data_set.shadow_creator_dict = creators
        return data_set

    def split_training_and_validation(self, targets, train_data_ratio):
        """A ratio below 1 is a stratified fraction; from 1 upwards it is a sample count per class."""
        if train_data_ratio < 1.0:
            return shuffle_training_data_using_ratio(targets, train_data_ratio)
        return shuffle_training_data_using_size(self.get_class_count(), targets, int(train_data_ratio), None)

    def split_samples(self, targets, train_data_ratio, test_data_ratio):
        train_set, validation_set = self.split_training_and_validation(targets, train_data_ratio)
        test_set, train_set = shuffle_test_data_using_ratio(train_set, test_data_ratio)
        return SampleSet(training_targets=train_set, test_targets=test_set, validation_targets=validation_set)

    @staticmethod
    def basic_data_set(casi, lidar, neighborhood, normalize, **given_range):
        return BasicDataSet(shadow_creator_dict=None, casi=casi, lidar=lidar, neighborhood=neighborhood,
                            normalize=normalize, **given_range)
