"""MUUFL Gulfport: 325 x 220, 64-band HSI + LiDAR, 11 classes (reference: loader/GULFPORTDataLoader.py)."""
import numpy

from hypelcnn_b200.common.common_nn_ops import read_targets_from_image
from hypelcnn_b200.loader.SceneFileDataLoader import SceneFileDataLoader


class GULFPORTDataLoader(SceneFileDataLoader):
    DIRECTORY = "/GULFPORT/"
    CLASSES = 11
    # trees, mostly grass, mixed ground, dirt and sand, road, water, building shadow, building, sidewalk,
    # yellow curb, cloth panels
    COLORS = ((0, 128, 0), (25, 255, 25), (0, 255, 255), (255, 204, 0), (255, 20, 67), (0, 0, 204), (102, 0, 204),
              (255, 132, 156), (204, 102, 0), (255, 255, 207), (208, 45, 115))
    BAND_RANGE = (405, 1005, 64)
    HSI_FILE, LIDAR_FILE, FILE_EXT = "muulf_hsi", "muulf_lidar", ".tif"

    def __init__(self, base_dir):
        super().__init__(base_dir)
        self._base_dir = base_dir

    def load_data(self, neighborhood, normalize):
        return self._load_data_utility(self.HSI_FILE + self.FILE_EXT, self.LIDAR_FILE + self.FILE_EXT, neighborhood,
                                       normalize)

    def _load_data_utility(self, hsi_file, lidar_file, neighborhood, normalize, casi_min=None, casi_max=None):
        lidar = numpy.expand_dims(self.read_raster(lidar_file), axis=2)
        return self.basic_data_set(self.read_raster(hsi_file), lidar, neighborhood, normalize, casi_min=casi_min,
                                   casi_max=casi_max)

    @staticmethod
    def _convert_targets_aux(targets):
        """Labels 1..11 -> classes 0..10; anything else (0 = unlabelled, 255 = masked out) is not a target."""
        return read_targets_from_image(targets, range(1, 12)) - [0, 0, 1]

    def read_targets(self, target_image_path):
        return self._convert_targets_aux(self.read_raster(target_image_path))

    def load_samples(self, train_data_ratio, test_data_ratio):
        return self.split_samples(self.read_targets("muulf_gt.tif"), train_data_ratio, test_data_ratio)
