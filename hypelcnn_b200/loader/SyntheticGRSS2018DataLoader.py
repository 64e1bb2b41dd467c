"""GRSS2018-shaped synthetic scene: 48-band float HSI at 1 m (601x2086 here) under a 0.5 m
LiDAR raster (1202x4172), 20 classes (reference: loader/GRSS2018DataLoader.py:47-57)."""
import numpy

from hypelcnn_b200.loader.SyntheticDataLoader import SyntheticDataLoader


class SyntheticGRSS2018DataLoader(SyntheticDataLoader):
    H, W, BANDS, CLASSES = 1202, 4172, 48, 20
    CASI_DTYPE = numpy.float32
    HALF_RES_HSI = True
    SAMPLES = (20000, 20000)
