"""IEEE GRSS DFC 2018 (Houston): 48-band CASI at 1 m (the file carries 50, the last two are dropped) + LiDAR DSM at
0.5 m, 20 classes (reference: loader/GRSS2018DataLoader.py).  The mixed resolution is resolved inside the gather kernel
(``HYP_GATHER_GRSS2018``), the cubes stay at their native sizes in HBM."""
import numpy

from hypelcnn_b200.loader.SceneFileDataLoader import SceneFileDataLoader
from hypelcnn_b200.loader.SyntheticDataLoader import GRSS2018DataSet

TRAINING_TILE_OFFSET = (1194, 1202)     # (x, y) of the labelled tile inside the LiDAR-resolution scene (:69-70)


class GRSS2018DataLoader(SceneFileDataLoader):
    DIRECTORY = "/2018_DFTC/"
    CLASSES = 20
    COLORS = ((0, 180, 0), (0, 124, 0), (0, 137, 69), (0, 69, 0), (255, 0, 0), (172, 125, 11), (0, 190, 194),
              (120, 0, 0), (216, 217, 247), (121, 121, 121), (255, 255, 0), (0, 155, 50), (0, 55, 55),
              (205, 172, 127), (220, 175, 120), (100, 100, 100), (185, 175, 94), (0, 237, 0), (207, 18, 56),
              (0, 0, 255), (0, 0, 0))   # 21 rows like the reference's table (the last one is unused)
    BAND_RANGE = (380, 1050, 48)

    def load_data(self, neighborhood, normalize):
        casi = numpy.ascontiguousarray(self.read_raster("20170218_UH_CASI_S4_NAD83.tiff")[:, :, 0:-2])
        lidar = numpy.array(self.read_raster("UH17c_GEF051.tif"), dtype=numpy.float32)[:, :, None]
        lidar[lidar > 300] = 0          # no-data spikes of the DSM
        return GRSS2018DataSet(shadow_creator_dict=None, casi=casi, lidar=lidar, neighborhood=neighborhood,
                               normalize=normalize)

    @staticmethod
    def print_stats(data):
        for band_index in range(1, data.shape[2]):
            band = data[:, :, band_index]
            print('Band mean:%.5f, band std:%.5f, min:%.5f, max:%.5f' % (band.mean(), band.std(), band.min(), band.max()))

    def read_targets(self, target_image_path):
        """Labels 1..20 of the ground-truth tile -> classes 0..19 at scene coordinates (tile offset added)."""
        labels = self.read_raster(target_image_path)
        rows = []
        for label in range(1, self.CLASSES + 1):
            ys, xs = numpy.where(labels == label)
            rows.append(numpy.stack([xs.astype(int) + TRAINING_TILE_OFFSET[0], ys.astype(int) + TRAINING_TILE_OFFSET[1],
                                     numpy.full(xs.shape, label - 1, dtype=int)], axis=1))
        return numpy.concatenate(rows).reshape(-1, 3)

    def load_samples(self, train_data_ratio, test_data_ratio):
        return self.split_samples(self.read_targets("2018_IEEE_GRSS_DFC_GT_TR.tif"), train_data_ratio, test_data_ratio)
