"""Validation side of GAN training (hypelcnn_b200/gan/wrappers/gan_common.py) against golden values produced by running
the reference's own gan/wrappers/gan_common.py (tests/golden/make_golden_gan_host.py): best-ratio bookkeeping, the
validation-iteration rule, the random validation sample picks, read_hsi_data, the band-ratio statistics and the
printed report.  CPU tensors throughout — the same code runs on CUDA tensors in training."""
import json
import os
import random

import numpy
import pytest
import torch

from tests.test_gan_samplers import BatchedProbeDataSet, ProbeDataSet

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "gan_host_golden.json")))
GOLD = numpy.load(os.path.join(HERE, "golden", "gan_host_golden.npz"))


def _pairs(holder):
    return [list(p) for p in holder.data_holder]


def test_best_ratio_holder_follows_the_reference():
    from hypelcnn_b200.gan.wrappers.gan_common import BestRatioHolder
    for seq in META["best_ratio"]:
        h = BestRatioHolder(seq["max_size"])
        for (it, v), state in zip(seq["points"], seq["states"]):
            h.add_point(numpy.int64(it), numpy.float64(v))
            assert _pairs(h) == state
        assert h.get_best_diver() == seq["best"] and str(h) == seq["str"]
        assert [list(h.get_point_with_itr(it)) for it, _ in seq["points"]] == seq["lookup"]
    assert BestRatioHolder(3).get_best_diver() is META["best_ratio_empty_best"]
    g = META["best_ratio_common"]
    h1, h2 = BestRatioHolder(10), BestRatioHolder(10)
    h1.data_holder, h2.data_holder = [tuple(p) for p in g["h1"]], [tuple(p) for p in g["h2"]]
    assert _pairs(BestRatioHolder.create_common_iterations(h1, h2)) == g["common"]


def test_best_ratio_holder_file_round_trip(tmp_path, capsys):
    from hypelcnn_b200.gan.wrappers.gan_common import BestRatioHolder
    h = BestRatioHolder(4)
    for it, v in [(101, .5), (201, .25), (301, .75)]:
        h.add_point(it, v)
    path = str(tmp_path / "best_ratio_shadowed.json")
    h.save(path)
    assert json.load(open(path)) == [[201, .25], [101, .5], [301, .75]]
    g = BestRatioHolder(4)
    g.load(path)
    assert _pairs(g) == [[201, .25], [101, .5], [301, .75]] and g.get_best_diver() == .25
    g.load(str(tmp_path / "missing.json"))                       # reported, holder unchanged
    (tmp_path / "broken.json").write_text("{not json")
    g.load(str(tmp_path / "broken.json"))
    out = capsys.readouterr().out
    assert "file not found" in out and "can not be decoded" in out and _pairs(g)[0] == [201, .25]


def test_validation_iteration_rule():
    from hypelcnn_b200.gan.wrappers.gan_common import BaseValidationHook
    for freq, true_at in META["is_validation_itr_true_at"].items():
        hook = BaseValidationHook(int(freq), "/tmp", 1.0)
        assert [i for i in range(2005) if hook._is_validation_itr(i)] == true_at
    assert all(BaseValidationHook(0, "/tmp", 1.0)._is_validation_itr(i) for i in range(50)) == \
        META["is_validation_itr_all_true_freq0"]
    hook = BaseValidationHook(1, "/tmp", 1.0)
    assert [i for i in range(20) if hook._is_validation_itr(i)] == META["is_validation_itr_freq1_true_at_first20"]


@pytest.mark.parametrize("data_set_cls", [ProbeDataSet, BatchedProbeDataSet])
def test_validation_sample_picks_equal_the_reference(data_set_cls):
    from hypelcnn_b200.gan.wrappers.gan_common import load_samples_for_testing
    smap = GOLD["lsft_map"]
    for case in META["load_samples_for_testing"]:
        random.seed(case["seed"])
        ds = data_set_cls(smap.shape, patch=2 * case["neighborhood"] + 1)
        got = load_samples_for_testing(ds, case["count"], case["neighborhood"], smap, fetch_shadows=case["fetch_shadows"])
        want = GOLD[case["key"]]
        assert len(got) == case["count"] and numpy.array_equal(numpy.asarray(got), want)
        assert want.shape[-1] == ds.get_casi_band_count()


def test_read_hsi_data():
    from hypelcnn_b200.gan.wrappers.gan_common import read_hsi_data

    class OneSampler:
        def get_sample_pairs(self, data_set, loader, shadow_map):
            base = numpy.arange(12, dtype=numpy.float32).reshape(4, 1, 1, 3)
            return base * 2, base

    n, s = read_hsi_data(None, ProbeDataSet([4, 4]), None, "one", {"one": OneSampler()})
    assert numpy.array_equal(n, GOLD["read_hsi_normal"]) and numpy.array_equal(s, GOLD["read_hsi_shadow"])
    with pytest.raises(ValueError) as e:
        read_hsi_data(None, ProbeDataSet([4, 4]), None, "nope", {"one": OneSampler()})
    assert str(e.value) == META["read_hsi_error"]


def test_band_ratio_statistics_equal_the_reference(tmp_path, capsys):
    """calculate_stats_from_samples: divergence within fp32 summation-order tolerance of the reference's numpy value,
    the printed mean±std table identical, samples with inf / nan ratios dropped."""
    from hypelcnn_b200.gan.wrappers import gan_common as C
    for case in META["calculate_stats"]:
        key = case["key"]
        samples, generated, ratio = GOLD[key + "_samples"], GOLD[key + "_generated"], GOLD[key + "_ratio"]
        div = C.calculate_stats_from_samples(lambda x: torch.from_numpy(generated), samples, ratio, str(tmp_path), 1,
                                             "plt", numpy.arange(samples.shape[-1]))
        assert div == pytest.approx(case["divergence"], rel=2e-5)
        assert capsys.readouterr().out == case["printed"]
        table = numpy.loadtxt(str(tmp_path / "plt_1.csv"), delimiter=",", skiprows=1)
        kept = numpy.isfinite(generated / samples).all(axis=3)
        assert kept.sum() == samples.shape[0] - 2
        want = (generated / samples)[kept] * ratio
        assert numpy.allclose(table[:, 1], numpy.percentile(want, 50, axis=0), rtol=1e-5)
        assert numpy.allclose(table[:, 2], numpy.percentile(want, 10, axis=0), rtol=1e-5)
        # create_stats_tensor: the same numbers through the training-time entry point, plus the upper divergence
        dm, du, r, mean, std = C.create_stats_tensor(torch.from_numpy(generated), torch.from_numpy(samples), ratio)
        assert float(dm) == pytest.approx(case["divergence"], rel=2e-5)
        assert numpy.allclose(mean.numpy(), want.mean(axis=0), rtol=1e-5) and numpy.allclose(std.numpy(), want.std(axis=0), rtol=1e-4)
        p = numpy.abs(want.mean(axis=0, dtype=numpy.float64) + want.std(axis=0, dtype=numpy.float64) - 1)
        assert float(du) == pytest.approx(0.5 * numpy.sum(p * numpy.log(2.0)), rel=1e-4)   # JS(p, 0) = sum p log 2 / 2


def test_print_overall_info_text(capsys):
    from hypelcnn_b200.gan.wrappers.gan_common import adj_shadow_ratio, print_overall_info
    print_overall_info(numpy.linspace(0.5, 2.0, 13), numpy.linspace(0.01, 0.4, 13))
    assert capsys.readouterr().out == META["print_overall_info_13"]
    assert [adj_shadow_ratio(4.0, True), adj_shadow_ratio(4.0, False)] == META["adj_shadow_ratio"]


class _Loader:
    def get_band_measurements(self):
        return numpy.arange(2)


def test_validation_hooks_end_to_end(tmp_path, capsys):
    """Two peer hooks over a probe scene with an 'ideal' generator (output = input / ratio): divergence ~ 0 at the
    validation iterations only, best-ratio files written, TensorBoard scalar present, peers report common bests."""
    from tensorboard.backend.event_processing.event_file_loader import EventFileLoader
    from hypelcnn_b200.gan.wrappers import gan_common as C
    rng = numpy.random.default_rng(0)
    smap = (rng.random((16, 18)) < 0.3).astype(numpy.uint8)
    ds = BatchedProbeDataSet(smap.shape)
    ds.get_data_point = None                                    # the hook must use the batched fetch
    ratio = numpy.array([2.0, 4.0], numpy.float32)
    random.seed(1)
    scale = {"fwd": 1.0}
    hook = C.create_base_validation_hook(ds, _Loader(), str(tmp_path), 0, smap, ratio, 5, 40,
                                         model_forward=lambda x: x / torch.from_numpy(ratio) * scale["fwd"],
                                         model_backward=lambda x: x * torch.from_numpy(ratio))
    hook.after_create_session(None, None)

    class Ctx:
        global_step = 0

    ctx = Ctx()
    for step in range(1, 13):
        ctx.global_step = step
        scale["fwd"] = 1.0 + 0.01 * step                        # the forward generator drifts away from the ideal
        hook.after_run(ctx, None)
    out = capsys.readouterr().out
    assert out.count("Validation metrics for shadowed #") == 2 and "#6" in out and "#11" in out
    assert out.count("Best common options:") == 2
    best_mean = hook.get_best_mean_div()
    assert len(best_mean) == 2 and best_mean[1] == pytest.approx(0.0, abs=1e-6)
    shadowed = json.load(open(tmp_path / "best_ratio_shadowed.json"))
    assert [p[0] for p in shadowed] == [6, 11] and shadowed[0][1] < shadowed[1][1]
    assert shadowed[0][1] == pytest.approx(0.5 * 2 * 0.06 * numpy.log(2.0), rel=1e-3)
    assert (tmp_path / "band_ratio_shadowed_6.csv").exists() and (tmp_path / "band_ratio_deshadowed_11.csv").exists()
    for h in hook._validation_base_hooks:
        h._writer.close()
    tags = [(e.step, v.tag) for f in sorted(os.listdir(tmp_path)) if "tfevents" in f
            for e in EventFileLoader(str(tmp_path / f)).Load() for v in e.summary.value]
    assert (6, "divergence_shadowed") in tags and (11, "divergence_deshadowed") in tags
    # a second run in the same log dir picks the saved bests up again (ValidationHook.__init__ -> load)
    again = C.ValidationHook(5, 4, str(tmp_path), _Loader(), ds, 0, smap, ratio, None, lambda x: x, "shadowed", False)
    assert [p[0] for p in again.best_mean_div_holder.data_holder] == [6, 11]


def test_shadow_ratio_from_the_resident_scene_equals_the_reference():
    """load_shadow_map_common: the ratio computed from the UNPADDED, un-normalised cube (weighted sums) against the
    reference's BasicDataSet (pad + normalise) + calculate_shadow_ratio run by the golden script."""
    from types import SimpleNamespace
    from hypelcnn_b200.common.common_nn_ops import load_shadow_map_common
    for name in ("r0", "r2"):
        casi, smap, n = GOLD[f"sr_{name}_casi"], GOLD[f"sr_{name}_map"], int(GOLD[f"sr_{name}_n"])
        cube = torch.from_numpy(casi)
        data_set = SimpleNamespace(casi=cube, _cmin=cube.amin(dim=(0, 1)))
        padded, ratio = load_shadow_map_common(data_set, n, smap)
        assert padded.shape == (smap.shape[0] + 2 * n, smap.shape[1] + 2 * n)
        assert numpy.array_equal(padded, numpy.pad(smap, n, mode="symmetric"))
        assert ratio.dtype == numpy.float32 and numpy.allclose(ratio, GOLD[f"sr_{name}_ratio"], rtol=2e-5)
    assert load_shadow_map_common(None, 1, smap)[1] is None
    with pytest.raises(ValueError):
        load_shadow_map_common(data_set, n + 1, smap[1:])


def test_shadow_tooling():
    """measure_targets_shadow_ratio.ratio_statistics and remove_test_targets_from_shadow.clear_targets against direct
    numpy (the reference scripts' arithmetic: divide, keep finite rows, mean / std; clear the shadowed targets)."""
    from hypelcnn_b200.utilities.measure_targets_shadow_ratio import ratio_statistics
    from hypelcnn_b200.utilities.remove_test_targets_from_shadow import clear_targets
    rng = numpy.random.default_rng(8)
    normal = rng.uniform(0.2, 1.0, (300, 1, 1, 6)).astype(numpy.float32)
    shadow = (normal / numpy.linspace(1.5, 4, 6).astype(numpy.float32) * rng.uniform(0.8, 1.2, normal.shape)).astype(numpy.float32)
    normal[5, 0, 0, 2] = 0.0                                            # inf -> the pair is dropped
    normal[9, 0, 0, 0], shadow[9, 0, 0, 0] = 0.0, 0.0                   # nan -> dropped
    mean, std = ratio_statistics(normal, shadow)
    with numpy.errstate(all="ignore"):
        ratio = numpy.squeeze(shadow) / numpy.squeeze(normal)
    ratio = ratio[numpy.isfinite(ratio).all(axis=1)]
    assert ratio.shape[0] == 298 and numpy.allclose(mean, ratio.mean(axis=0), rtol=1e-5) and numpy.allclose(std, ratio.std(axis=0), rtol=1e-4)
    shadow_map = (rng.random((12, 15)) < 0.4).astype(numpy.uint8)
    targets = numpy.stack([rng.integers(0, 15, 40), rng.integers(0, 12, 40), rng.integers(0, 5, 40)], axis=1)
    cleared, lit = clear_targets(shadow_map, targets)
    want = shadow_map.copy()
    count = 0
    for x, y, _ in targets:                                             # the reference's loop
        if want[y, x] == 1:
            want[y, x] = 0
        else:
            count += 1
    assert numpy.array_equal(cleared, want) and shadow_map.sum() > cleared.sum()
    # the loop counts a target twice-listed under shadow as "not shadowed" the second time; the array form does not
    assert lit <= count


def test_remaining_gan_common_names():
    from hypelcnn_b200.gan.wrappers import gan_common as C
    from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import get_lr
    for step in (0, 1, 499, 500, 501, 750, 999, 1000, 1500):
        assert C._get_lr(2e-4, 1000, step) == get_lr(2e-4, 1000, step)
    assert C._get_lr(1.0, 7, 3) == 1.0 and C._get_lr(1.0, 7, 5) == pytest.approx(0.5) and C._get_lr(1.0, 7, 7) == 0.0
    ds = ProbeDataSet([4, 4], patch=3)
    assert C.create_input_tensor(ds, True) == {"name": "x", "shape": [None, 3, 3, 2], "dtype": "float32"}
    assert C.create_input_tensor(ds, False)["name"] == "y"
    hook = C.InitializerHook("ITER", None, None, "N", "S")
    hook.after_create_session(None, None)
    assert (hook.input_itr, hook.normal_data, hook.shadow_data) == ("ITER", "N", "S")
    reference_names = ["InitializerHook", "BestRatioHolder", "BaseValidationHook", "PeerValidationHook", "ValidationHook",
                       "adj_shadow_ratio", "_get_lr", "define_standard_train_ops", "create_inference_for_matrix_input",
                       "create_input_tensor", "create_stats_tensor", "calculate_stats_from_samples",
                       "load_samples_for_testing", "read_hsi_data", "plot_overall_info", "print_overall_info",
                       "model_generator_name", "model_base_name", "input_x_tensor_name", "input_y_tensor_name"]
    assert [n for n in reference_names if not hasattr(C, n)] == []
