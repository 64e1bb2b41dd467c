"""The GAN pair samplers (hypelcnn_b200/gan/gan_sampling_methods.py) against golden matrices produced by running the
reference's own gan/gan_sampling_methods.py over a probe data set (tests/golden/make_golden_gan_host.py): which scene
pixel lands in which row of the (normal, shadow) matrices — neighbourhood ring, random split with element-wise repeat,
per-class target pairing with the margin filter — bit-exact, for the per-point DataSet contract and for the batched
``get_data_points`` path the device data set takes."""
import json
import os

import numpy
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "gan_host_golden.json")))
GOLD = numpy.load(os.path.join(HERE, "golden", "gan_host_golden.npz"))


class ProbeDataSet:
    """Same probe as the golden script: get_data_point(x, y) = (x, y, 7x + 13y) everywhere in the patch."""

    def __init__(self, scene_shape, patch=1):
        self.scene_shape, self.patch, self.calls = list(scene_shape), patch, 0

    def get_data_shape(self):
        return [self.patch, self.patch, 3]

    def get_casi_band_count(self):
        return 2

    def get_scene_shape(self):
        return self.scene_shape

    def get_data_point(self, x, y):
        self.calls += 1
        out = numpy.empty([self.patch, self.patch, 3], numpy.float32)
        out[..., 0], out[..., 1], out[..., 2] = x, y, 7 * x + 13 * y
        return out


class BatchedProbeDataSet(ProbeDataSet):
    """Adds the batched fetch BasicDataSet offers (one gather launch there); counts the calls."""

    def __init__(self, scene_shape, patch=1):
        super().__init__(scene_shape, patch)
        self.batched_calls = 0

    def get_data_points(self, targets_xy):
        self.batched_calls += 1
        t = numpy.asarray(targets_xy).astype(numpy.float32)
        vals = numpy.stack([t[:, 0], t[:, 1], 7 * t[:, 0] + 13 * t[:, 1]], axis=1)
        return numpy.broadcast_to(vals[:, None, None, :], [t.shape[0], self.patch, self.patch, 3]).copy()


class ProbeLoader:
    def __init__(self, targets, class_count):
        self.targets, self.class_count = targets, class_count

    def read_targets(self, name):
        assert name == "shadow_gen_model/class_result.tif"
        return self.targets.copy()

    def get_class_count(self):
        return range(self.class_count)


def _sampler(case):
    from hypelcnn_b200.gan import gan_sampling_methods as S
    if case["kind"] == "neighbour":
        return S.NeighborhoodBasedSampler(neighborhood_size=case["size"], margin=case["margin"])
    if case["kind"] == "random":
        return S.RandomBasedSampler(multiply_shadowed_data=case["multiply"])
    return S.TargetBasedSampler(margin=case["margin"])


@pytest.mark.parametrize("data_set_cls", [ProbeDataSet, BatchedProbeDataSet])
@pytest.mark.parametrize("case", META["sampler_cases"], ids=lambda c: c["key"])
def test_sampler_rows_equal_the_reference(case, data_set_cls, capsys):
    shadow_map = GOLD["map_" + case["map"]]
    data_set = data_set_cls(shadow_map.shape, case.get("patch", 1))
    loader = ProbeLoader(GOLD[case["key"] + "_targets"], case["classes"]) if case["kind"] == "target" else None
    normal, shadow = _sampler(case).get_sample_pairs(data_set, loader, shadow_map)
    if case.get("none"):
        assert normal is None and shadow is None
        return
    want_normal, want_shadow = GOLD[case["key"] + "_normal"], GOLD[case["key"] + "_shadow"]
    assert normal.dtype == numpy.float32 and shadow.dtype == numpy.float32
    assert normal.shape == want_normal.shape and shadow.shape == want_shadow.shape
    assert numpy.array_equal(normal, want_normal) and numpy.array_equal(shadow, want_shadow)
    if case["kind"] == "target":
        assert capsys.readouterr().out == case["printed"]
    if data_set_cls is BatchedProbeDataSet:
        assert data_set.calls == 0 and data_set.batched_calls <= 2      # one fetch per matrix, no per-pixel calls


def test_goldens_cover_the_quirks():
    """The fixtures really contain the cases the mirror has to reproduce: a margin-0 ring whose uint8 subtraction wraps
    (all-zero normal rows), a map touching the scene border, a target set with no class on both sides (None, None)."""
    zero_ring = GOLD["neigh_a_2_0_normal"]
    assert zero_ring.shape[0] == GOLD["neigh_a_2_0_shadow"].shape[0] and not zero_ring.any()
    assert GOLD["map_edge"][0, 0] == 1 and GOLD["map_edge"][-1, -1] == 1
    assert any(c.get("none") for c in META["sampler_cases"]) and any(c["kind"] == "target" and not c["none"]
                                                                     for c in META["sampler_cases"])


def test_target_pairing_counts():
    """Property: per class, every normal point appears once and the shadowed points are cycled to the same count."""
    from hypelcnn_b200.gan.gan_sampling_methods import target_pair_targets
    rng = numpy.random.default_rng(3)
    shadow_map = (rng.random((30, 40)) < 0.3).astype(numpy.uint8)
    targets = numpy.stack([rng.integers(0, 40, 500), rng.integers(0, 30, 500), rng.integers(-1, 6, 500)], axis=1)
    messages = []
    normal, shadow = target_pair_targets(targets, shadow_map, 6, report=messages.append)
    assert normal.shape == shadow.shape and not messages
    assert not (shadow_map[normal[:, 1], normal[:, 0]] == 1).any() and (shadow_map[shadow[:, 1], shadow[:, 0]] == 1).all()
    labelled = targets[targets[:, 2] >= 0]
    assert normal.shape[0] == int((shadow_map[labelled[:, 1], labelled[:, 0]] != 1).sum())
    # a class with shadowed points only is reported and skipped
    only_shadow = numpy.array([[x, y, 0] for y, x in zip(*numpy.nonzero(shadow_map))][:5])
    assert target_pair_targets(only_shadow, shadow_map, 1, report=messages.append) == (None, None)
    assert messages == ["Target key is not found in read target image during target based sampling:0"]


def test_registry_lists_the_reference_pairing_methods():
    """gan/wrapper_registry.py:13-18: target (margin 5), random (multiply), neighbour (20, 2), dummy (2000, 0.5, 2)."""
    from hypelcnn_b200.gan import gan_sampling_methods as S
    from hypelcnn_b200.gan.wrapper_registry import get_sampling_map
    m = get_sampling_map()
    assert list(m) == ["target", "random", "neighbour", "dummy"]
    assert isinstance(m["target"], S.TargetBasedSampler) and m["target"]._margin == 5
    assert isinstance(m["random"], S.RandomBasedSampler) and m["random"]._multiply_shadowed_data is True
    assert isinstance(m["neighbour"], S.NeighborhoodBasedSampler)
    assert (m["neighbour"]._neighborhood_size, m["neighbour"]._margin) == (20, 2)
    assert isinstance(m["dummy"], S.DummySampler)
    assert (m["dummy"]._element_count, m["dummy"]._fill_value, m["dummy"]._coefficient) == (2000, 0.5, 2)
