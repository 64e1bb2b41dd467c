"""GULFPORT-shaped synthetic scene: 325x220, 64-band HSI + 1 LiDAR, 11 classes."""
import numpy

from hypelcnn_b200.loader.SyntheticDataLoader import SyntheticDataLoader


class SyntheticGULFPORTDataLoader(SyntheticDataLoader):
    H, W, BANDS, CLASSES = 325, 220, 64, 11
    CASI_DTYPE = numpy.float32
    SAMPLES = (2000, 2000)
    SHADOWED = True      # the GAN configurations (BASELINE configs 4 / 5) need shadowed and lit spectra
