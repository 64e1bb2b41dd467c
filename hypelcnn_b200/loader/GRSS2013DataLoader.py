"""IEEE GRSS DFC 2013 (Houston): 349 x 1905, 144-band CASI + LiDAR DSM, 15 classes
(reference: loader/GRSS2013DataLoader.py)."""
from hypelcnn_b200.common.common_nn_ops import shuffle_test_data_using_ratio
from hypelcnn_b200.loader.DataLoader import SampleSet
from hypelcnn_b200.loader.SceneFileDataLoader import SceneFileDataLoader


class GRSS2013DataLoader(SceneFileDataLoader):
    DIRECTORY = "/2013_DFTC/"
    CLASSES = 15
    # healthy / stressed / synthetic grass, tree, soil, water, residential, commercial, road, highway, railway,
    # parking lot 1 / 2, tennis court, running track
    COLORS = ((0, 180, 0), (0, 124, 0), (0, 137, 69), (0, 69, 0), (172, 125, 11), (0, 190, 194), (120, 0, 0),
              (216, 217, 247), (121, 121, 121), (205, 172, 127), (220, 175, 120), (100, 100, 100), (185, 175, 94),
              (0, 237, 0), (207, 18, 56))
    BAND_RANGE = (380, 1050, 144)
    SHADOW_MAP_FILE = "shadow_map.tif"
    GAN_CHECKPOINTS = {"cycle_gan": "shadow_gen_model/cycle_gan/model.ckpt-5000",
                       "dcl_gan": "shadow_gen_model/dcl_gan/model.ckpt-3000",
                       "dcl_cycle_gan": "shadow_gen_model/dcl_cycle_gan/model.ckpt-5000"}

    def load_data(self, neighborhood, normalize):
        casi = self.read_raster("2013_IEEE_GRSS_DF_Contest_CASI.tif")
        lidar = self.read_raster("2013_IEEE_GRSS_DF_Contest_LiDAR.tif")[:, :, None]
        return self.attach_shadow_creators(self.basic_data_set(casi, lidar, neighborhood, normalize), neighborhood)

    def load_samples(self, train_data_ratio, test_data_ratio):
        """The contest's own training / validation label images; the test list is cut from the training one."""
        train_set = self.read_targets("2013_IEEE_GRSS_DF_Contest_Samples_TR.tif")
        validation_set = self.read_targets("2013_IEEE_GRSS_DF_Contest_Samples_VA.tif")
        test_set, train_set = shuffle_test_data_using_ratio(train_set, test_data_ratio)
        return SampleSet(training_targets=train_set, test_targets=test_set, validation_targets=validation_set)
