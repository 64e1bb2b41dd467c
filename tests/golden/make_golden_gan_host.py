#!/usr/bin/env python
"""Golden vectors for the GAN host logic, produced by EXECUTING the reference's own modules (build container only):

* gan/gan_sampling_methods.py — NeighborhoodBasedSampler, RandomBasedSampler, TargetBasedSampler, DummySampler run
  over a probe data set whose get_data_point(x, y) returns (x, y, 7x+13y), so the returned matrices spell out which
  scene pixel landed in which row of the (normal, shadow) pair lists;
* gan/wrappers/gan_common.py (imported against the recording stubs of make_golden.py; none of the functions used here
  touches TensorFlow) — BestRatioHolder insert / trim / common-iteration logic, BaseValidationHook._is_validation_itr,
  load_samples_for_testing under a fixed ``random.seed``, read_hsi_data, the numpy statistics of
  calculate_stats_from_samples and the text of print_overall_info.

``gan_host_golden.npz`` / ``gan_host_golden.json`` are committed and read by tests/test_gan_samplers.py and
tests/test_gan_validation.py.

usage: python tests/golden/make_golden_gan_host.py"""
import contextlib
import io
import json
import os
import random
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


class ProbeDataSet:
    """get_data_point(x, y) -> [P,P,3] float32 filled with (x, y, 7x + 13y)."""

    def __init__(self, scene_shape, patch=1):
        self.scene_shape, self.patch = list(scene_shape), patch
        self.calls = 0

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


class ProbeLoader:
    def __init__(self, targets, class_count):
        self.targets, self.class_count = targets, class_count

    def read_targets(self, name):
        assert name == "shadow_gen_model/class_result.tif"
        return self.targets.copy()

    def get_class_count(self):
        return range(self.class_count)


def blob_map(rng, h, w, blobs, dtype=numpy.uint8):
    m = numpy.zeros([h, w], dtype)
    for _ in range(blobs):
        r, c = rng.integers(0, h), rng.integers(0, w)
        hh, ww = rng.integers(1, 5), rng.integers(1, 6)
        m[r:r + hh, c:c + ww] = 1
    return m


def sampler_goldens(S):
    rng = numpy.random.default_rng(1234)
    arrays, meta = {}, {}
    maps = {"a": blob_map(rng, 23, 31, 5), "b": blob_map(rng, 40, 37, 9), "edge": blob_map(rng, 12, 9, 2)}
    maps["edge"][0, :3] = 1
    maps["edge"][-1, -2:] = 1
    for name, m in maps.items():
        arrays[f"map_{name}"] = m
    cases = []
    for name, m in maps.items():
        for (size, margin) in [(3, 1), (20, 2), (2, 0)]:
            normal, shadow = S.NeighborhoodBasedSampler(neighborhood_size=size, margin=margin).get_sample_pairs(
                ProbeDataSet(m.shape), None, m)
            key = f"neigh_{name}_{size}_{margin}"
            arrays[key + "_normal"], arrays[key + "_shadow"] = normal, shadow
            cases.append({"kind": "neighbour", "map": name, "size": size, "margin": margin, "key": key})
        for mult in (True, False):
            normal, shadow = S.RandomBasedSampler(multiply_shadowed_data=mult).get_sample_pairs(
                ProbeDataSet(m.shape), None, m)
            key = f"random_{name}_{int(mult)}"
            arrays[key + "_normal"], arrays[key + "_shadow"] = normal, shadow
            cases.append({"kind": "random", "map": name, "multiply": mult, "key": key})
    # patch > 1: the samplers only forward get_data_shape
    normal, shadow = S.RandomBasedSampler(True).get_sample_pairs(ProbeDataSet(maps["edge"].shape, patch=3), None, maps["edge"])
    arrays["random_edge_p3_normal"], arrays["random_edge_p3_shadow"] = normal, shadow
    cases.append({"kind": "random", "map": "edge", "multiply": True, "patch": 3, "key": "random_edge_p3"})

    # target based: [N,3] = (x, y, class) targets, margin filter, per-class expansion of the shadowed points
    for name, class_count, n, margin in [("a", 5, 120, 5), ("b", 7, 400, 5), ("b", 3, 60, 2), ("a", 4, 30, 11)]:
        m = maps[name]
        t = numpy.stack([rng.integers(0, m.shape[1], n), rng.integers(0, m.shape[0], n),
                         rng.integers(0, class_count, n)], axis=1).astype(numpy.int32)
        key = f"target_{name}_{class_count}_{n}_{margin}"
        ds = ProbeDataSet(m.shape)
        with contextlib.redirect_stdout(io.StringIO()) as out:
            normal, shadow = S.TargetBasedSampler(margin=margin).get_sample_pairs(ds, ProbeLoader(t, class_count), m)
        arrays[key + "_targets"] = t
        none = normal is None
        if not none:
            arrays[key + "_normal"], arrays[key + "_shadow"] = normal, shadow
        cases.append({"kind": "target", "map": name, "classes": class_count, "margin": margin, "key": key,
                      "none": none, "printed": out.getvalue()})
    meta["sampler_cases"] = cases
    return arrays, meta


def common_goldens(C):
    arrays, meta = {}, {}
    # BestRatioHolder
    seqs = []
    rng = numpy.random.default_rng(99)
    for max_size, n in [(10, 25), (3, 7), (10, 4), (1, 5)]:
        pts = [(int(i * 100 + 1), float(numpy.round(rng.random(), 3))) for i in range(n)]
        pts[n // 2] = (pts[n // 2][0], pts[0][1])  # a tie
        h = C.BestRatioHolder(max_size)
        states = []
        for it, v in pts:
            h.add_point(numpy.int64(it), numpy.float64(v))
            states.append([list(p) for p in h.data_holder])
        seqs.append({"max_size": max_size, "points": pts, "states": states, "best": h.get_best_diver(),
                     "lookup": [list(h.get_point_with_itr(it)) for it, _ in pts], "str": str(h)})
    meta["best_ratio"] = seqs
    h1, h2 = C.BestRatioHolder(10), C.BestRatioHolder(10)
    for it, v in [(101, .5), (201, .2), (301, .9), (401, .1)]:
        h1.add_point(it, v)
    for it, v in [(201, .7), (401, .3), (501, .05), (101, .6)]:
        h2.add_point(it, v)
    meta["best_ratio_common"] = {"h1": [list(p) for p in h1.data_holder], "h2": [list(p) for p in h2.data_holder],
                                 "common": [list(p) for p in C.BestRatioHolder.create_common_iterations(h1, h2).data_holder]}
    meta["best_ratio_empty_best"] = C.BestRatioHolder(3).get_best_diver()

    # _is_validation_itr
    table = {}
    for freq in (0, 1, 2, 100, 1000):
        hook = C.BaseValidationHook(freq, "/tmp", 1.0)
        table[str(freq)] = [bool(hook._is_validation_itr(i)) for i in range(0, 2005)]
    meta["is_validation_itr_true_at"] = {k: [i for i, v in enumerate(t) if v] for k, t in table.items() if k != "0" and k != "1"}
    meta["is_validation_itr_all_true_freq0"] = all(table["0"])
    meta["is_validation_itr_freq1_true_at_first20"] = [i for i, v in enumerate(table["1"][:20]) if v]

    # load_samples_for_testing (python `random`)
    rng = numpy.random.default_rng(5)
    smap = blob_map(rng, 20, 17, 6)
    arrays["lsft_map"] = smap
    lsft = []
    for neighborhood, fetch, count, seed in [(0, True, 12, 7), (0, False, 12, 7), (2, True, 9, 11), (2, False, 20, 3)]:
        random.seed(seed)
        ds = ProbeDataSet(smap.shape, patch=2 * neighborhood + 1)
        out = C.load_samples_for_testing(ds, count, neighborhood, smap, fetch_shadows=fetch)
        key = f"lsft_{neighborhood}_{int(fetch)}_{count}_{seed}"
        arrays[key] = numpy.asarray(out)
        lsft.append({"neighborhood": neighborhood, "fetch_shadows": fetch, "count": count, "seed": seed, "key": key})
    meta["load_samples_for_testing"] = lsft

    # read_hsi_data
    class OneSampler:
        def get_sample_pairs(self, data_set, loader, shadow_map):
            base = numpy.arange(4 * 1 * 1 * 3, dtype=numpy.float32).reshape(4, 1, 1, 3)
            return base * 2, base

    n, s = C.read_hsi_data(None, ProbeDataSet([4, 4]), None, "one", {"one": OneSampler()})
    arrays["read_hsi_normal"], arrays["read_hsi_shadow"] = n, s
    try:
        C.read_hsi_data(None, ProbeDataSet([4, 4]), None, "nope", {"one": OneSampler()})
    except ValueError as e:
        meta["read_hsi_error"] = str(e)

    # calculate_stats_from_samples: numpy statistics with a session that returns a prepared "generated" array
    rng = numpy.random.default_rng(42)
    stats = []
    for bands, count in [(8, 50), (64, 200)]:
        samples = rng.uniform(0.02, 0.5, (count, 1, 1, bands)).astype(numpy.float32)
        samples[3, 0, 0, 2] = 0.0  # -> inf ratio: the sample is eliminated
        samples[7, 0, 0, 0] = 0.0
        generated = (samples * rng.uniform(1.5, 4.0, bands).astype(numpy.float32)
                     * rng.uniform(0.9, 1.1, samples.shape).astype(numpy.float32)).astype(numpy.float32)
        generated[7, 0, 0, 0] = 0.0  # 0/0 -> nan
        shadow_ratio = rng.uniform(0.25, 0.7, bands).astype(numpy.float32)

        class Sess:
            def run(self, tensor, feed_dict):
                return generated

        text = io.StringIO()
        with contextlib.redirect_stdout(text), numpy.errstate(all="ignore"):
            div = C.calculate_stats_from_samples(Sess(), samples, "in", "out", shadow_ratio, "/tmp", 1, "plt",
                                                 numpy.arange(bands))
        key = f"stats_{bands}_{count}"
        arrays[key + "_samples"], arrays[key + "_generated"], arrays[key + "_ratio"] = samples, generated, shadow_ratio
        stats.append({"key": key, "divergence": float(div), "printed": text.getvalue()})
    meta["calculate_stats"] = stats

    text = io.StringIO()
    with contextlib.redirect_stdout(text):
        C.print_overall_info(numpy.linspace(0.5, 2.0, 13), numpy.linspace(0.01, 0.4, 13))
    meta["print_overall_info_13"] = text.getvalue()
    meta["adj_shadow_ratio"] = [float(C.adj_shadow_ratio(4.0, True)), float(C.adj_shadow_ratio(4.0, False))]
    return arrays, meta


def shadow_ratio_goldens():
    """load_shadow_map_common (common/common_nn_ops.py:567-571) minus the tif read: the reference's BasicDataSet pads and
    normalises the cube, the shadow map is padded the same way, calculate_shadow_ratio (:473-483) gives the ratio."""
    from common.common_nn_ops import BasicDataSet, calculate_shadow_ratio
    rng = numpy.random.default_rng(77)
    arrays = {}
    for name, (h, w, c, n) in {"r0": (12, 9, 5, 0), "r2": (14, 11, 6, 2)}.items():
        casi = rng.integers(100, 16384, (h, w, c)).astype(numpy.float32)
        smap = blob_map(rng, h, w, 4, numpy.uint8)
        casi[smap == 1] /= rng.uniform(1.5, 4.0, c).astype(numpy.float32)
        ds = BasicDataSet(None, casi.copy(), None, n, True)
        padded = numpy.pad(smap, n, mode="symmetric")
        ratio = calculate_shadow_ratio(ds.casi, padded, numpy.logical_not(padded).astype(int))
        arrays[f"sr_{name}_casi"], arrays[f"sr_{name}_map"], arrays[f"sr_{name}_ratio"] = casi, smap, ratio
        arrays[f"sr_{name}_n"] = numpy.array(n)
    return arrays


def train_app_goldens():
    """gan/gan_train_for_shadow.py cannot be imported even against the stubs (tfgan namedtuples are subclassed at import
    time), so the two pure functions needed — add_parse_cmds_for_app (:28-77) and get_log_suffix (:187-200) — are
    compiled from the reference file's own AST and executed here; the other flag groups come from common/cmd_parser.py,
    which imports as it is."""
    import argparse
    import ast
    from types import SimpleNamespace
    from common import cmd_parser
    from common.common_ops import replace_abbrs
    path = os.path.join(G.REF, "gan", "gan_train_for_shadow.py")
    tree = ast.parse(open(path).read())
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("add_parse_cmds_for_app", "get_log_suffix")]
    space = {"type_ensure_strtobool": cmd_parser.type_ensure_strtobool, "replace_abbrs": replace_abbrs}
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), space)
    parser = argparse.ArgumentParser()
    cmd_parser.add_parse_cmds_for_loaders(parser)
    cmd_parser.add_parse_cmds_for_loggers(parser)
    cmd_parser.add_parse_cmds_for_trainers(parser)
    space["add_parse_cmds_for_app"](parser)
    flags, _ = parser.parse_known_args([])
    defaults = {k: v for k, v in vars(flags).items() if k not in ("base_log_path", "output_path")}
    suffixes = []
    for over in [{}, {"use_identity_loss": False}, {"loader_name": "GULFPORTALTDataLoader", "gan_type": "DCL_GAN",
                                                    "neighborhood": 2, "regularization_support_rate": 0.25,
                                                    "batch_size": 128}]:
        f = SimpleNamespace(**{**vars(flags), **over})
        suffixes.append({"overrides": over, "suffix": space["get_log_suffix"](f)})
    out = {"train_flag_defaults": defaults, "log_suffixes": suffixes}

    # classify/train_for_classification.py: add_parse_cmds_for_app (:123-157), get_log_suffix (:160-180) the same way
    from common.common_ops import path_leaf
    path = os.path.join(G.REF, "classify", "train_for_classification.py")
    tree = ast.parse(open(path).read())
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("add_parse_cmds_for_app", "get_log_suffix")]
    space = {"type_ensure_strtobool": cmd_parser.type_ensure_strtobool, "replace_abbrs": replace_abbrs,
             "path_leaf": path_leaf, "os": os}
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), space)
    parser = argparse.ArgumentParser()
    for add in (cmd_parser.add_parse_cmds_for_loaders, cmd_parser.add_parse_cmds_for_loggers,
                cmd_parser.add_parse_cmds_for_trainers, cmd_parser.add_parse_cmds_for_models,
                cmd_parser.add_parse_cmds_for_importers, space["add_parse_cmds_for_app"]):
        add(parser)
    flags, _ = parser.parse_known_args([])
    out["classify_flag_defaults"] = {k: v for k, v in vars(flags).items() if k not in ("base_log_path", "output_path")}
    suffixes = []
    for over in [{"algorithm_param_path": "/some/dir/alg_param_hypelcnn.json"},
                 {"algorithm_param_path": "alg_param_dualcnn.json", "model_name": "DUALCNNModel", "neighborhood": 3,
                  "train_ratio": 200.0, "augment_data_with_shadow": "cycle_gan", "augmentation_random_threshold": 0.25},
                 {"algorithm_param_path": "C:\\x\\alg_param_concnn.json", "model_name": "CONCNNModel",
                  "loader_name": "GULFPORTALTDataLoader", "neighborhood": 1, "train_ratio": 0.5,
                  "augment_data_with_spectral": 0.0125, "augment_data_with_shadow": "simple"}]:
        f = SimpleNamespace(**{**vars(flags), **over})
        suffixes.append({"overrides": over, "suffix": space["get_log_suffix"](f)})
    out["classify_log_suffixes"] = suffixes

    # classify/infer_for_classification.py: create_all_scene_data (:24-35), create_sample_data (:38-47)
    from collections import namedtuple
    info = namedtuple("GeneratorDataInfo", ["data", "targets", "loader", "dataset"])
    path = os.path.join(G.REF, "classify", "infer_for_classification.py")
    tree = ast.parse(open(path).read())
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("create_all_scene_data", "create_sample_data")]
    space = {"numpy": numpy, "GeneratorDataInfo": info}
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), space)
    scene = space["create_all_scene_data"]([3, 5], info(None, None, "the loader", "the data set"))
    parts = [info(None, numpy.array(t, dtype=numpy.int64), f"loader{i}", f"set{i}")
             for i, t in enumerate([[[1, 2, 3]], [[4, 5, 6], [7, 8, 9]], [[10, 11, 12]]])]
    sample = space["create_sample_data"](*parts)
    out["infer_targets"] = {"all_3x5": scene.targets.tolist(), "all_loader": scene.loader, "all_dataset": scene.dataset,
                            "sample": sample.targets.tolist(), "sample_dtype": str(sample.targets.dtype),
                            "sample_loader": sample.loader, "sample_dataset": sample.dataset}
    return out


def loader_goldens():
    """The reference's scene loaders run over small synthetic scene directories (written with this repo's TIFF
    writer, read back through ``tifffile.imread`` := this repo's reader): constant tables, read_targets, the union and
    sizes of load_samples' splits (the stratified shuffles themselves are unseeded in the reference)."""
    import importlib.machinery
    import tempfile
    from collections import namedtuple
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from hypelcnn_b200.utilities.tiff_io import imread, imwrite

    class Finder:           # every tensorflow* / tf_slim* / tensorflow_gan* submodule resolves to a stub
        def find_spec(self, name, path, target=None):
            if name.split(".")[0] in ("tensorflow", "tensorflow_gan", "tf_slim", "tensorflow_probability", "imageio"):
                return importlib.machinery.ModuleSpec(name, self)

        def create_module(self, spec):
            return sys.modules.get(spec.name) or G._Stub(spec.name)

        def exec_module(self, module):
            pass

    sys.meta_path.append(Finder())
    import tensorflow_gan as tfgan
    tfgan.CycleGANModel = namedtuple("CycleGANModel", ["model_x2y", "model_y2x", "reconstructed_x", "reconstructed_y"])
    rng = numpy.random.default_rng(2013)
    base = tempfile.mkdtemp()
    arrays, meta = {}, {}

    def labels(h, w, classes, density=0.4, first=1):
        image = rng.integers(first, first + classes, (h, w)).astype(numpy.uint8)
        image[rng.random((h, w)) > density] = 0
        return image

    files = {"2013_DFTC/2013_IEEE_GRSS_DF_Contest_Samples_TR.tif": labels(20, 30, 15, first=0),
             "2013_DFTC/2013_IEEE_GRSS_DF_Contest_Samples_VA.tif": labels(20, 30, 15, first=0),
             "2018_DFTC/2018_IEEE_GRSS_DFC_GT_TR.tif": labels(24, 40, 20),
             "GULFPORT/muulf_gt.tif": labels(25, 22, 11),
             "GULFPORT/muulf_gt_shadow_corrected.tif": labels(25, 22, 11),
             "GULFPORT/muulf_shadow_map.tif": blob_map(rng, 25, 22, 6)}
    for name, image in files.items():
        os.makedirs(os.path.dirname(os.path.join(base, name)), exist_ok=True)
        imwrite(os.path.join(base, name), image)
        arrays["file_" + name.replace("/", "__")] = image
    tables = {}
    for name in ["GRSS2013DataLoader", "GRSS2018DataLoader", "GULFPORTDataLoader", "GULFPORTALTDataLoader", "AVONDataLoader"]:
        module = __import__("loader." + name, fromlist=[name])
        if hasattr(module, "imread"):
            module.imread = imread
        loader = getattr(module, name)(base)
        bands = loader.get_band_measurements()
        tables[name] = {"colors": loader.get_samples_color_list().tolist(),
                        "classes": [loader.get_class_count().start, loader.get_class_count().stop],
                        "bands": [float(bands[0]), float(bands[-1]), int(bands.shape[0])],
                        "base_suffix": loader.get_model_base_dir()[len(base):]}
        if name == "GRSS2013DataLoader":
            arrays["targets_2013_tr"] = loader.read_targets("2013_IEEE_GRSS_DF_Contest_Samples_TR.tif")
            s = loader.load_samples(0.1, 0.25)
            tables[name]["split_sizes"] = [len(s.training_targets), len(s.test_targets), len(s.validation_targets)]
            arrays["targets_2013_va"] = s.validation_targets
        if name == "GRSS2018DataLoader":
            s = loader.load_samples(0.5, 0.2)
            tables[name]["split_sizes"] = [len(s.training_targets), len(s.test_targets), len(s.validation_targets)]
            arrays["targets_2018_all"] = numpy.vstack([s.training_targets, s.test_targets, s.validation_targets])
        if name == "GULFPORTDataLoader":
            arrays["targets_gulfport"] = loader.read_targets("muulf_gt.tif")
            s = loader.load_samples(3, 0.0)        # 3 samples per class, no test list
            tables[name]["split_sizes"] = [len(s.training_targets), len(s.test_targets), len(s.validation_targets)]
        if name == "GULFPORTALTDataLoader":
            sys.modules["common.common_nn_ops"].imread = imread
            s = loader.load_samples(0.5, 0.3)
            tables[name]["split_sizes"] = [len(s.training_targets), len(s.test_targets), len(s.validation_targets)]
            tables[name]["test_shape"] = list(s.test_targets.shape)
            arrays["targets_alt_train_val"] = numpy.vstack([s.training_targets, s.validation_targets]).astype(int)
    meta["loaders"] = tables
    return arrays, meta


def main():
    G.install_stubs()
    for sub in ["tensorflow.python.ops.math_ops", "tensorflow.python.summary", "tensorflow.python.summary.summary",
                "tensorflow.python.training.adam", "tensorflow.python.training.learning_rate_decay",
                "tensorflow.python.training.training_util"]:
        sys.modules[sub] = G._Stub(sub)
    import gan.gan_sampling_methods as S
    import gan.wrappers.gan_common as C
    C.plt.rcParams = {}  # the plotting stub only has to accept the font settings
    a1, m1 = sampler_goldens(S)
    a3 = shadow_ratio_goldens()
    m3 = train_app_goldens()
    a4, m4 = loader_goldens()
    a2, m2 = common_goldens(C)
    numpy.savez_compressed(os.path.join(HERE, "gan_host_golden.npz"), **a1, **a2, **a3, **a4)
    with open(os.path.join(HERE, "gan_host_golden.json"), "w") as f:
        json.dump({**m1, **m2, **m3, **m4}, f, indent=1)
    print("wrote", len(a1) + len(a2), "arrays")


if __name__ == "__main__":
    main()
