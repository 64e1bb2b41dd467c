"""The data-set plug-in contract of this engine.

It keeps the names and call signatures of the reference's loader interface (loader/DataLoader.py:5-47: ``DataLoader``,
``SampleSet``, ``LoadingMode``) because that is the drop-in boundary: a loader written against the reference is usable
here once it returns numpy / torch data instead of tf tensors.  What each hook must hand to THIS engine is spelled out
per method below.
"""
import abc
import enum


class LoadingMode(enum.Enum):
    """Which rendition of the scene a loader serves (reference values, loader/DataLoader.py:13-17): the original cube,
    the GAN-shadowed / de-shadowed one, or a per-sample mix."""
    ORIGINAL = ""
    SHADOWED = "shadowed"
    DESHADOWED = "deshadowed"
    MIXED = "mixed"


class SampleSet:
    """The three disjoint target lists a loader produces: rows of (x, y, label) scene coordinates, int."""

    __slots__ = ("training_targets", "test_targets", "validation_targets")

    def __init__(self, validation_targets, training_targets, test_targets) -> None:
        self.training_targets, self.test_targets, self.validation_targets = (training_targets, test_targets,
                                                                             validation_targets)


class DataLoader(abc.ABC):
    """Scene access.  The engine's gather kernel (``hyp_gather_patches``) cuts the patches itself, so a loader only
    has to expose the scene cube and the labelled pixel lists."""

    @abc.abstractmethod
    def load_data(self, neighborhood, normalize):
        """-> data-set object holding the HSI cube [H, W, C] (uint16 or float32), the LiDAR raster(s), the
        neighborhood n (patch edge 2n+1) and, when ``normalize``, the per-band min/max used for scaling."""
        raise NotImplementedError

    @abc.abstractmethod
    def load_samples(self, train_data_ratio, test_data_ratio):
        """-> ``SampleSet``; ratios follow the reference: train ratio of the labelled pixels, test ratio of the rest."""
        raise NotImplementedError

    @abc.abstractmethod
    def load_shadow_map(self, neighborhood, data_set):
        """-> (shadow_map [H, W] bool, per-band shadow_ratio [C]) for the shadow augmenters."""
        raise NotImplementedError

    @abc.abstractmethod
    def get_class_count(self):
        """-> number of classes (width of the logits)."""
        raise NotImplementedError

    @abc.abstractmethod
    def get_model_base_dir(self):
        """-> directory the checkpoints of this data set live under."""
        raise NotImplementedError

    @abc.abstractmethod
    def get_samples_color_list(self):
        """-> per-class RGB rows for rendering class maps (reporting only)."""
        raise NotImplementedError

    @abc.abstractmethod
    def get_band_measurements(self):
        """-> wavelength per band (reporting only)."""
        raise NotImplementedError
