"""CPU ORACLE (test infrastructure, NOT product code) — scene preparation, patch gather,
metrics.  numpy restatement, pinned bit-exactly against fixtures produced by running the
reference's own code (tests/golden/make_golden.py -> tests/golden/dataset_golden.npz).

Reference files followed:
  common/common_nn_ops.py:45-106     BasicDataSet (symmetric pad, min-max normalise)
  common/common_nn_ops.py:169-185    get_data_point_func / _hsi (window slice)
  loader/GRSS2018DataLoader.py:10-44 GRSS2018DataSet.get_data_point (mixed resolution)
  importer/InMemoryImporter.py:27-38 _get_data_with_labels
  common/common_nn_ops.py:243-292    argmax / confusion / class accuracies
  utilities/stat_extractor.py:24-62  kappa
"""
import numpy


class SceneRef:
    """BasicDataSet.__init__ (common/common_nn_ops.py:46-82)."""

    def __init__(self, casi, lidar, neighborhood, normalize):
        self.neighborhood = n = neighborhood
        pad = ((n, n), (n, n), (0, 0))
        self.lidar = None if lidar is None else numpy.pad(lidar, pad, mode="symmetric")
        self.casi = numpy.pad(casi, pad, mode="symmetric")
        self.casi_min, self.casi_max, self.lidar_min, self.lidar_max = 0, 1, 0, 1
        if normalize:
            if self.lidar is not None:
                self.lidar_min = numpy.min(self.lidar)
                self.lidar = self.lidar - self.lidar_min
                self.lidar_max = numpy.max(self.lidar)
                self.lidar = self.lidar / self.lidar_max
            self.casi_min = numpy.min(self.casi, axis=(0, 1))
            self.casi = self.casi - self.casi_min
            self.casi_max = numpy.max(self.casi, axis=(0, 1))
            self.casi = self.casi / self.casi_max.astype(numpy.float32)

    def data_shape(self):
        d = 2 * self.neighborhood + 1
        return [d, d, self.casi.shape[2] + (0 if self.lidar is None else 1)]

    def scene_shape(self):
        p = self.lidar if self.lidar is not None else self.casi
        return [p.shape[0] - 2 * self.neighborhood, p.shape[1] - 2 * self.neighborhood]

    def get_data_point(self, x, y):
        """common/common_nn_ops.py:169-185 — x = column, y = row of the UNPADDED scene."""
        s = 2 * self.neighborhood + 1
        if self.lidar is None:
            return self.casi[y:y + s, x:x + s, :]
        return numpy.concatenate((self.casi[y:y + s, x:x + s, :], self.lidar[y:y + s, x:x + s, :]), axis=2)


class SceneRef2018(SceneRef):
    """GRSS2018DataSet (loader/GRSS2018DataLoader.py:10-44): HSI at half the LiDAR resolution,
    nearest-neighbour 2x upsample inside the window; int() truncation throughout."""

    @staticmethod
    def _position(neighborhood, px, py, scale):
        actual = int(neighborhood * scale)
        return int(px * scale) + neighborhood - actual, int(py * scale) + neighborhood - actual

    def get_data_point(self, x, y):
        n = self.neighborhood
        s = 2 * n + 1
        cb = self.casi.shape[2]
        out = numpy.empty([s, s, cb + 1], dtype=self.casi.dtype)
        sx, sy = self._position(n, x, y, 0.5)
        lx, ly = self._position(n, x, y, 1)
        for xi in range(s):
            for yi in range(s):
                out[yi, xi, 0:cb] = self.casi[sy + int(yi * 0.5), sx + int(xi * 0.5), :]
                out[yi, xi, cb] = self.lidar[ly + yi, lx + xi, 0]
        return out


def gather_patches(scene, targets):
    """InMemoryImporter._get_data_with_labels (importer/InMemoryImporter.py:27-38)."""
    data = numpy.zeros([targets.shape[0]] + scene.data_shape(), dtype=numpy.float32)
    labels = numpy.zeros(targets.shape[0], dtype=numpy.uint8)
    for i, p in enumerate(targets):
        data[i] = scene.get_data_point(int(p[0]), int(p[1]))
        labels[i] = p[2]
    return data, labels


# --------------------------------------------------------------------------- #
def argmax_lowest(logits):
    """tf.argmax: lowest index on ties (App. A.11).  numpy.argmax has the same rule."""
    return numpy.argmax(logits, axis=1)


def confusion_matrix(labels, predictions, num_classes):
    """tf.math.confusion_matrix(labels, predictions): rows = labels, cols = predictions, int32."""
    c = numpy.zeros((num_classes, num_classes), dtype=numpy.int32)
    numpy.add.at(c, (labels.astype(numpy.int64), predictions.astype(numpy.int64)), 1)
    return c


def class_accuracies(conf, class_range):
    """calculate_class_accuracies_using_confusion (common/common_nn_ops.py:280-292)."""
    n = class_range.stop
    precision, recall = numpy.zeros(n), numpy.zeros(n)
    for i in class_range:
        gt = numpy.sum(conf[i, :])
        if gt != 0:
            recall[i] = conf[i, i] / gt
        pr = numpy.sum(conf[:, i])
        if pr != 0:
            precision[i] = conf[i, i] / pr
    return recall[class_range], precision[class_range]


def overall_accuracy(conf):
    return float(numpy.trace(conf)) / float(max(conf.sum(), 1))


def mean_per_class_accuracy(conf):
    """tf.metrics.mean_per_class_accuracy: mean over classes of diag/row-sum; classes with an
    empty row contribute 0 to the sum but are still counted [TF-lib: div_no_nan, then mean]."""
    rows = conf.sum(axis=1).astype(numpy.float64)
    d = numpy.diag(conf).astype(numpy.float64)
    per = numpy.where(rows > 0, d / numpy.maximum(rows, 1), 0.0)
    return float(per.mean())


def kappa(conf):
    """Cohen's kappa (po - pe)/(1 - pe) — same value as utilities/stat_extractor.py:24-62."""
    c = conf.astype(numpy.float64)
    n = c.sum()
    po = numpy.trace(c) / n
    pe = float((c.sum(axis=1) * c.sum(axis=0)).sum()) / (n * n)
    return (po - pe) / (1 - pe)


def scatter_class_map(scene_shape, targets, predictions, fill=255):
    """perform_prediction's scatter (common/common_nn_ops.py:313-327): img[y, x] = class."""
    img = numpy.full([scene_shape[0], scene_shape[1]], fill, dtype=numpy.uint8)
    for t, p in zip(targets, predictions):
        img[t[1], t[0]] = p
    return img


def augment_patches(x, choices, deltas=None):
    """Replay of the reference's per-sample augmentation maps (common/common_nn_ops.py:397-440) for a given draw:
    choices[b] = (k, flip_lr, flip_ud): tf.image.rot90(img, k) (counter-clockwise == numpy.rot90 on the H, W axes),
    then tf.image.flip_left_right / flip_up_down, then img + delta[b] (broadcast over pixels, :428-431)."""
    out = numpy.empty_like(x)
    for b in range(x.shape[0]):
        k, flr, fud = int(choices[b][0]), int(choices[b][1]), int(choices[b][2])
        img = numpy.rot90(x[b], k, axes=(0, 1))
        if flr:
            img = img[:, ::-1, :]
        if fud:
            img = img[::-1, :, :]
        if deltas is not None:
            img = img + deltas[b][None, None, :]
        out[b] = img
    return out
