"""Synthetic scene loaders with the shapes of the reference's data sets (the GeoTIFF files
are not available; SURVEY §8d S-gather).  They implement the reference's DataLoader
interface (loader/DataLoader.py:5-47) and resolve through the same name registry:
``--loader_name SyntheticGRSS2013DataLoader``.

`base_dir` may carry overrides: "synthetic:H=64,W=80,samples=500".
"""
import numpy

from hypelcnn_b200 import _native as N
from hypelcnn_b200.common.common_nn_ops import BasicDataSet
from hypelcnn_b200.loader.DataLoader import DataLoader, SampleSet


class GRSS2018DataSet(BasicDataSet):
    """Mixed-resolution scene (reference: loader/GRSS2018DataLoader.py:10-44): HSI at half the
    LiDAR resolution; the gather kernel upsamples nearest-neighbour inside the window."""
    gather_mode = N.HYP_GATHER_GRSS2018

    def get_scene_shape(self):
        return [self.lidar.shape[0], self.lidar.shape[1]]


class SyntheticDataLoader(DataLoader):
    H, W, BANDS, CLASSES = 349, 1905, 144, 15
    CASI_DTYPE = numpy.uint16
    HALF_RES_HSI = False
    SAMPLES = (2832, 12197)  # training / validation counts of GRSS2013 (dataset facts)

    def __init__(self, base_dir):
        self.base_dir = base_dir
        self.h, self.w, self.samples = self.H, self.W, self.SAMPLES
        if isinstance(base_dir, str) and base_dir.startswith("synthetic:"):
            for kv in base_dir[len("synthetic:"):].split(","):
                if not kv:
                    continue
                k, v = kv.split("=")
                if k == "H":
                    self.h = int(v)
                elif k == "W":
                    self.w = int(v)
                elif k == "samples":
                    self.samples = (int(v), int(v))
        self.rng = numpy.random.default_rng(1234)

    def load_data(self, neighborhood, normalize):
        hc, wc = (self.h // 2, self.w // 2) if self.HALF_RES_HSI else (self.h, self.w)
        if self.CASI_DTYPE == numpy.uint16:
            casi = self.rng.integers(0, 16384, (hc, wc, self.BANDS)).astype(numpy.uint16)
        else:
            casi = self.rng.random((hc, wc, self.BANDS), dtype=numpy.float32)
        lidar = (self.rng.random((self.h, self.w, 1), dtype=numpy.float32) * 50).astype(numpy.float32)
        cls = GRSS2018DataSet if self.HALF_RES_HSI else BasicDataSet
        return cls(shadow_creator_dict=None, casi=casi, lidar=lidar, neighborhood=neighborhood, normalize=normalize)

    def _targets(self, n):
        return numpy.stack([self.rng.integers(0, self.w, n), self.rng.integers(0, self.h, n),
                            self.rng.integers(0, self.CLASSES, n)], axis=1).astype(numpy.int64)

    def load_samples(self, train_data_ratio, test_data_ratio):
        train = self._targets(self.samples[0])
        validation = self._targets(self.samples[1])
        n_test = int(round(train.shape[0] * test_data_ratio)) if test_data_ratio > 0 else 0
        test, train = train[:n_test], train[n_test:]
        return SampleSet(training_targets=train, test_targets=test, validation_targets=validation)

    def load_shadow_map(self, neighborhood, data_set):
        return None, None

    def get_class_count(self):
        return range(0, self.CLASSES)

    def get_model_base_dir(self):
        return self.base_dir

    def get_samples_color_list(self):
        return (numpy.arange(self.CLASSES * 3).reshape(self.CLASSES, 3) * 5 % 256).astype(numpy.uint8)

    def get_band_measurements(self):
        return numpy.linspace(380, 1050, num=self.BANDS)
