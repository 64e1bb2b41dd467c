"""The scene-file loaders (hypelcnn_b200/loader/{GRSS2013,GRSS2018,GULFPORT,GULFPORTALT,AVON}DataLoader.py) against values
produced by running the reference's own loader classes over the same synthetic scene directories
(tests/golden/make_golden_gan_host.py: loader_goldens): constant tables, read_targets row for row, the sizes and unions
of the sample splits.  What load_data hands to the device data set is checked with a recording stand-in (the data set
itself needs CUDA and is covered by the -m gpu tests)."""
import json
import os

import numpy
import pytest

from hypelcnn_b200.utilities.tiff_io import imwrite

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "gan_host_golden.json")))["loaders"]
GOLD = numpy.load(os.path.join(HERE, "golden", "gan_host_golden.npz"))
NAMES = ["GRSS2013DataLoader", "GRSS2018DataLoader", "GULFPORTDataLoader", "GULFPORTALTDataLoader", "AVONDataLoader"]


def _loader(name, base):
    from hypelcnn_b200.common.common_nn_ops import get_loader_from_name
    return get_loader_from_name(name, str(base))


@pytest.fixture()
def scene_dir(tmp_path):
    for key in GOLD.files:
        if key.startswith("file_"):
            path = tmp_path / key[len("file_"):].replace("__", os.sep)
            os.makedirs(path.parent, exist_ok=True)
            imwrite(str(path), GOLD[key])
    return tmp_path


def _sorted_rows(a):
    a = numpy.asarray(a).astype(int).reshape(-1, 3)
    return a[numpy.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


@pytest.mark.parametrize("name", NAMES)
def test_constant_tables_equal_the_reference(name):
    loader, g = _loader(name, "/base"), META[name]
    assert loader.get_samples_color_list().tolist() == g["colors"] and loader.get_samples_color_list().dtype == numpy.uint8
    assert [loader.get_class_count().start, loader.get_class_count().stop] == g["classes"]
    bands = loader.get_band_measurements()
    assert [float(bands[0]), float(bands[-1]), int(bands.shape[0])] == g["bands"]
    assert loader.get_model_base_dir() == "/base" + g["base_suffix"]


def test_grss2013_targets_and_splits(scene_dir):
    loader = _loader("GRSS2013DataLoader", scene_dir)
    assert numpy.array_equal(loader.read_targets("2013_IEEE_GRSS_DF_Contest_Samples_TR.tif"), GOLD["targets_2013_tr"])
    s = loader.load_samples(0.1, 0.25)
    assert [len(s.training_targets), len(s.test_targets), len(s.validation_targets)] == META["GRSS2013DataLoader"]["split_sizes"]
    assert numpy.array_equal(s.validation_targets, GOLD["targets_2013_va"])           # the VA image, untouched
    assert numpy.array_equal(_sorted_rows(numpy.vstack([s.training_targets, s.test_targets])),
                             _sorted_rows(GOLD["targets_2013_tr"]))


def test_grss2018_targets_and_splits(scene_dir):
    loader = _loader("GRSS2018DataLoader", scene_dir)
    everything = loader.read_targets("2018_IEEE_GRSS_DFC_GT_TR.tif")
    assert numpy.array_equal(_sorted_rows(everything), _sorted_rows(GOLD["targets_2018_all"]))
    assert everything[:, 0].min() >= 1194 and everything[:, 1].min() >= 1202 and set(everything[:, 2]) <= set(range(20))
    s = loader.load_samples(0.5, 0.2)
    assert [len(s.training_targets), len(s.test_targets), len(s.validation_targets)] == META["GRSS2018DataLoader"]["split_sizes"]
    assert numpy.array_equal(_sorted_rows(numpy.vstack([s.training_targets, s.test_targets, s.validation_targets])),
                             _sorted_rows(everything))
    assert loader.load_shadow_map(0, None) is None          # `pass` in the reference


def test_gulfport_targets_and_splits(scene_dir):
    loader = _loader("GULFPORTDataLoader", scene_dir)
    targets = loader.read_targets("muulf_gt.tif")
    assert numpy.array_equal(targets, GOLD["targets_gulfport"])
    s = loader.load_samples(3, 0.0)                          # >= 1: samples per class
    assert [len(s.training_targets), len(s.test_targets), len(s.validation_targets)] == META["GULFPORTDataLoader"]["split_sizes"]
    assert numpy.bincount(s.training_targets[:, 2].astype(int), minlength=11).tolist() == [3] * 11


def test_gulfportalt_splits_by_shadow(scene_dir):
    loader = _loader("GULFPORTALTDataLoader", scene_dir)
    g = META["GULFPORTALTDataLoader"]
    s = loader.load_samples(0.5, 0.3)
    assert [len(s.training_targets), len(s.test_targets), len(s.validation_targets)] == g["split_sizes"]
    assert list(s.test_targets.shape) == g["test_shape"]
    assert numpy.array_equal(_sorted_rows(numpy.vstack([s.training_targets, s.validation_targets])),
                             _sorted_rows(GOLD["targets_alt_train_val"]))
    shadow = GOLD["file_GULFPORT__muulf_shadow_map.tif"]
    train = s.training_targets.astype(int)
    assert not shadow[train[:, 1], train[:, 0]].any()        # nothing under the shadow map is trained on
    padded, ratio = loader.load_shadow_map(2, None)
    assert padded.shape == (shadow.shape[0] + 4, shadow.shape[1] + 4) and ratio is None


def test_load_data_hands_the_scene_over_like_the_reference(tmp_path, monkeypatch):
    """File names, band selection, LiDAR clean-up, given normalisation ranges — recorded instead of uploaded."""
    from hypelcnn_b200.loader import GRSS2018DataLoader as M18
    from hypelcnn_b200.loader.SceneFileDataLoader import SceneFileDataLoader
    rng = numpy.random.default_rng(5)
    os.makedirs(tmp_path / "2018_DFTC")
    os.makedirs(tmp_path / "GULFPORT")
    os.makedirs(tmp_path / "AVON")
    casi18 = rng.integers(0, 4000, (6, 8, 50)).astype(numpy.uint16)
    lidar18 = (rng.random((12, 16)) * 400).astype(numpy.float32)
    imwrite(str(tmp_path / "2018_DFTC" / "20170218_UH_CASI_S4_NAD83.tiff"), casi18)
    imwrite(str(tmp_path / "2018_DFTC" / "UH17c_GEF051.tif"), lidar18)
    seen = {}

    class Recorder:
        def __init__(self, **kwargs):
            seen.update(kwargs)
            self.casi_min, self.casi_max = "MIN", "MAX"

    monkeypatch.setattr(M18, "GRSS2018DataSet", Recorder)
    _loader("GRSS2018DataLoader", tmp_path).load_data(5, True)
    assert seen["casi"].shape == (6, 8, 48) and numpy.array_equal(seen["casi"], casi18[:, :, :48])
    assert seen["lidar"].shape == (12, 16, 1) and seen["lidar"].max() <= 300 and (seen["lidar"] == 0).sum() == (lidar18 > 300).sum()
    assert (seen["neighborhood"], seen["normalize"], seen["shadow_creator_dict"]) == (5, True, None)

    hsi = rng.random((7, 9, 64)).astype(numpy.float32)
    for suffix in ("", "_shadowed", "_deshadowed"):
        imwrite(str(tmp_path / "GULFPORT" / f"muulf_hsi{suffix}.tif"), hsi * (1 + len(suffix)))
    imwrite(str(tmp_path / "GULFPORT" / "muulf_lidar.tif"), rng.random((7, 9)).astype(numpy.float32))
    calls = []
    monkeypatch.setattr(SceneFileDataLoader, "basic_data_set",
                        staticmethod(lambda casi, lidar, neighborhood, normalize, **given: calls.append(
                            (casi, None if lidar is None else lidar.shape, neighborhood, normalize, given)) or Recorder()))
    monkeypatch.setattr(SceneFileDataLoader, "attach_shadow_creators", lambda self, data_set, neighborhood: data_set)
    from hypelcnn_b200.loader.DataLoader import LoadingMode
    alt = _loader("GULFPORTALTDataLoader", tmp_path)
    alt.load_data(1, True)
    assert len(calls) == 1 and numpy.array_equal(calls[0][0], hsi) and calls[0][1:] == ((7, 9, 1), 1, True, {"casi_min": None, "casi_max": None})
    alt._load_mode = LoadingMode.SHADOWED
    calls.clear()
    alt.load_data(1, True)
    assert len(calls) == 2 and calls[1][4] == {"casi_min": "MIN", "casi_max": "MAX"}   # the ORIGINAL scene's range
    assert numpy.array_equal(calls[1][0], hsi * (1 + len("_shadowed")))

    avon = rng.integers(0, 3000, (12, 5, 4 + 110)).astype(numpy.int32)                 # [bands, W, H + margins]
    imwrite(str(tmp_path / "AVON" / "0920-1857.georef_cropped.tif"), avon)
    calls.clear()
    _loader("AVONDataLoader", tmp_path).load_data(0, True)
    casi = calls[0][0]
    assert calls[0][1] is None                                      # no LiDAR for this scene
    assert casi.shape == (4, 5, 12) and casi.dtype == numpy.uint16 and calls[0][4] == {"casi_min": 0}
    want = numpy.swapaxes(avon[:, :, 55:-55], 0, 2).astype(numpy.uint16)
    ceiling = numpy.percentile(want, 95, axis=(0, 1)).astype(numpy.uint16)
    assert numpy.array_equal(casi, numpy.minimum(want, ceiling))


def test_avon_overlays(tmp_path):
    from PIL import Image
    os.makedirs(tmp_path / "AVON")
    rng = numpy.random.default_rng(9)
    marks = {}
    for no in (1, 2):
        for kind in ("nsh", "sh"):
            image = (rng.random((130, 20)) < 0.1)
            marks[(no, kind)] = image[55:-55]
            Image.fromarray((image * 255).astype(numpy.uint8)).save(
                str(tmp_path / "AVON" / f"0920-1857.georef_cropped_rgb_with_targets_{no}_{kind}.bmp"))
    loader = _loader("AVONDataLoader", tmp_path)
    t2 = loader.read_each_target("0920-1857.georef_cropped_rgb_with_targets_2_nsh.bmp", target_no=2)
    assert len(t2) == marks[(2, "nsh")].sum() and set(t2[:, 2]) == {1}
    s = loader.load_samples(0.5, 0.0)
    total = sum(m.sum() for m in marks.values())
    assert len(s.training_targets) + len(s.test_targets) + len(s.validation_targets) == total
    shadowed = int(marks[(1, "sh")].sum() + marks[(2, "sh")].sum())
    assert len(s.validation_targets) >= shadowed              # every shadowed target validates


def test_multi_data_set_mixes_renditions_per_point():
    """GULFPORTALT's MIXED mode: every requested point comes from one randomly chosen rendition, the batched form
    issues one fetch per rendition; shape queries go to the first data set."""
    import torch
    from hypelcnn_b200.loader.GULFPORTALTDataLoader import MultiDataSet

    class Rendition:
        neighborhood, shadow_creator_dict, lidar, casi, device = 1, {"simple": 1}, "lidar", "casi", torch.device("cpu")

        def __init__(self, value):
            self.value, self.batched = value, 0

        def get_data_shape(self):
            return [3, 3, 5]

        def get_casi_band_count(self):
            return 4

        def get_scene_shape(self):
            return [10, 12]

        def get_unnormalized_casi_dtype(self):
            return numpy.dtype(numpy.uint16)

        def get_data_point(self, point_x, point_y):
            return torch.full((3, 3, 5), float(self.value))

        def get_data_points(self, targets_xy):
            self.batched += 1
            out = torch.full((targets_xy.shape[0], 3, 3, 5), float(self.value))
            out[:, 0, 0, 0] = targets_xy[:, 0].float()              # which point landed where
            return out

    original, shadowed = Rendition(1), Rendition(2)
    mixed = MultiDataSet(original, shadowed, shadowed, shadowed)
    assert mixed.get_data_shape() == [3, 3, 5] and mixed.get_casi_band_count() == 4 and mixed.get_scene_shape() == [10, 12]
    assert (mixed.casi, mixed.lidar, mixed.neighborhood, mixed.shadow_creator_dict) == ("casi", "lidar", 1, {"simple": 1})
    targets = numpy.stack([numpy.arange(400), numpy.zeros(400, int)], axis=1)
    got = mixed.get_data_points(targets)
    assert got.shape == (400, 3, 3, 5) and torch.equal(got[:, 0, 0, 0], torch.arange(400).float())   # order kept
    share_shadowed = float((got[:, 1, 1, 1] == 2).float().mean())
    assert 0.65 < share_shadowed < 0.85 and set(got[:, 1, 1, 1].tolist()) == {1.0, 2.0}             # 1 : 3 mix
    assert original.batched == 1 and shadowed.batched == 3
    singles = {float(mixed.get_data_point(0, 0)[0, 0, 0]) for _ in range(40)}
    assert singles == {1.0, 2.0}
