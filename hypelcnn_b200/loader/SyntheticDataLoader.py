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
    SHADOWED = False         # True: the pixels under the synthetic shadow map are darkened by SHADOW_RATIO per band

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
        self._shadow_map = None

    def load_data(self, neighborhood, normalize):
        hc, wc = (self.h // 2, self.w // 2) if self.HALF_RES_HSI else (self.h, self.w)
        if self.CASI_DTYPE == numpy.uint16:
            casi = self.rng.integers(0, 16384, (hc, wc, self.BANDS)).astype(numpy.uint16)
        else:
            casi = self.rng.random((hc, wc, self.BANDS), dtype=numpy.float32)
        lidar = (self.rng.random((self.h, self.w, 1), dtype=numpy.float32) * 50).astype(numpy.float32)
        if self.SHADOWED and not self.HALF_RES_HSI:
            in_shadow = self.synthetic_shadow_map() == 1
            casi[in_shadow] = (casi[in_shadow] / self.shadow_band_ratio()).astype(casi.dtype)
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

    def synthetic_shadow_map(self):
        """[H,W] uint8 blobs (about a fifth of the scene) from a generator of its own, so asking for the shadow map
        does not shift the scene / sample draws."""
        if self._shadow_map is None:
            rng = numpy.random.default_rng(4321)
            shadow = numpy.zeros([self.h, self.w], numpy.uint8)
            for _ in range(max(3, self.h * self.w // 250)):
                r, c = rng.integers(0, self.h), rng.integers(0, self.w)
                shadow[r:r + rng.integers(3, 14), c:c + rng.integers(3, 14)] = 1
            self._shadow_map = shadow
        return self._shadow_map

    def shadow_band_ratio(self):
        """lit / shadowed level per band used when SHADOWED (SURVEY S-C4: linspace(1.5, 4, bands))."""
        return numpy.linspace(1.5, 4.0, self.BANDS).astype(numpy.float32)

    def load_shadow_map(self, neighborhood, data_set):
        """-> (shadow map padded by neighborhood, per-band lit / shadow ratio measured on the data set) like the
        reference loaders (load_shadow_map_common, common/common_nn_ops.py:567-571); the map is synthetic."""
        from hypelcnn_b200.common.common_nn_ops import load_shadow_map_common
        if self.HALF_RES_HSI:
            return numpy.pad(self.synthetic_shadow_map(), neighborhood, mode="symmetric"), None
        return load_shadow_map_common(data_set, neighborhood, self.synthetic_shadow_map())

    def read_targets(self, target_image_path):
        """Labelled pixels of a target image as [N,3] = (x, y, class) (reference loaders: read_targets); synthetic:
        drawn from a generator seeded by the name, independent of the scene draws."""
        rng = numpy.random.default_rng(sum(target_image_path.encode()) + 99)
        n = self.samples[0]
        return numpy.stack([rng.integers(0, self.w, n), rng.integers(0, self.h, n),
                            rng.integers(0, self.CLASSES, n)], axis=1).astype(numpy.int64)

    def get_class_count(self):
        return range(0, self.CLASSES)

    def get_model_base_dir(self):
        return self.base_dir

    def get_samples_color_list(self):
        return (numpy.arange(self.CLASSES * 3).reshape(self.CLASSES, 3) * 5 % 256).astype(numpy.uint8)

    def get_band_measurements(self):
        return numpy.linspace(380, 1050, num=self.BANDS)
