"""SHARE 2012 Avon: a 360-band HSI flight line without LiDAR, two target classes marked in bitmap overlays
(reference: loader/AVONDataLoader.py)."""
import numpy

from hypelcnn_b200.common.common_nn_ops import read_targets_from_image, shuffle_test_data_using_ratio
from hypelcnn_b200.loader.DataLoader import SampleSet
from hypelcnn_b200.loader.SceneFileDataLoader import SceneFileDataLoader

BLANK_OFFSET = 55       # empty margin of the georeferenced strip, cut from both ends


class AVONDataLoader(SceneFileDataLoader):
    DIRECTORY = "/AVON/"
    CLASSES = 2
    COLORS = ((0, 0, 255), (255, 0, 0))
    BAND_RANGE = (400, 2500, 360)
    SHADOW_MAP_FILE = "0920-1857.georef_cropped_shadow.tif"
    GAN_CHECKPOINTS = {"cycle_gan": "shadow_gen_model/cycle_gan/model.ckpt-7000",
                       "dcl_gan": "shadow_gen_model/dcl_gan/model.ckpt-6000",
                       "dcl_cycle_gan": "shadow_gen_model/dcl_cycle_gan/model.ckpt-3000"}

    def __init__(self, base_dir):
        super().__init__(base_dir)
        self.load_shadow_corrected = False

    def load_data(self, neighborhood, normalize):
        if self.load_shadow_corrected:
            casi = self.read_raster("0920-1857.georef_cropped_shcorrected.tif")
        else:       # stored band-major with the blank margin: [bands, W, H-with-margin] -> [H, W, bands]
            casi = numpy.swapaxes(self.read_raster("0920-1857.georef_cropped.tif")[:, :, BLANK_OFFSET:-BLANK_OFFSET], 0, 2)
        casi = numpy.ascontiguousarray(casi.astype(numpy.uint16))
        ceiling = numpy.percentile(casi, 95, axis=(0, 1)).astype(casi.dtype)         # clip the top 5 % per band
        numpy.clip(casi, None, ceiling, out=casi)
        data_set = self.basic_data_set(casi, None, neighborhood, normalize, casi_min=0)
        return self.attach_shadow_creators(data_set, neighborhood)

    def read_each_target(self, target_image_path, target_no):
        """A black / white overlay marks one class: white pixels become class ``target_no - 1``."""
        from PIL import Image
        image = numpy.array(Image.open(self.get_model_base_dir() + target_image_path))[BLANK_OFFSET:-BLANK_OFFSET, :]
        if image.dtype == bool:
            image = image.astype(numpy.uint8) * 255
        targets = ((image / 255).astype(int) * target_no) - 1
        return read_targets_from_image(targets, self.get_class_count())

    def load_samples(self, train_data_ratio, test_data_ratio):
        prefix = "0920-1857.georef_cropped_rgb_with_targets_"
        lit = [self.read_each_target(f"{prefix}{no}_nsh.bmp", target_no=no) for no in (1, 2)]
        shadowed = [self.read_each_target(f"{prefix}{no}_sh.bmp", target_no=no) for no in (1, 2)]
        if train_data_ratio < 1.0:      # the reference splits with shuffle_test_data_using_ratio here (sic)
            splits = [shuffle_test_data_using_ratio(targets, train_data_ratio) for targets in lit]
        else:
            splits = [self.split_training_and_validation(targets, train_data_ratio) for targets in lit]
        train_set = numpy.vstack([s[0] for s in splits])
        validation_set = numpy.vstack(shadowed + [s[1] for s in splits])
        test_set, train_set = shuffle_test_data_using_ratio(train_set, test_data_ratio)
        return SampleSet(training_targets=train_set, test_targets=test_set, validation_targets=validation_set)
