"""Host logic of the GAN training run (hypelcnn_b200/gan/gan_train_for_shadow.py) on CPU tensors: flag defaults and
log-directory suffix against values produced by executing the reference's own functions
(tests/golden/make_golden_gan_host.py), the pair iterator's shuffle-repeat-batch contract, the regularisation-support
augmentation, and the training loop's op / hook / checkpoint order with a recording wrapper."""
import json
import os
from types import SimpleNamespace

import numpy
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "gan_host_golden.json")))


def test_flag_defaults_and_log_suffix_equal_the_reference():
    from hypelcnn_b200.gan.gan_train_for_shadow import default_flags, get_log_suffix
    ours = vars(default_flags())
    for name, value in META["train_flag_defaults"].items():
        if name in ("master", "ps_tasks", "task"):      # parameter-server flags: not part of this engine
            continue
        assert ours[name] == value, name
    for case in META["log_suffixes"]:
        assert get_log_suffix(default_flags(**case["overrides"])) == case["suffix"]
    with pytest.raises(KeyError):
        default_flags(no_such_flag=1)


def test_pair_iterator_contract():
    """epoch passes, each a permutation; batches cut from the concatenated passes (they straddle pass boundaries);
    the tail that does not fill a batch is dropped; rows stay paired."""
    from hypelcnn_b200.gan.gan_train_for_shadow import PairIterator
    rows, batch, epoch = 10, 4, 3
    normal = torch.arange(rows, dtype=torch.float32).reshape(rows, 1, 1, 1).expand(rows, 1, 1, 5).contiguous()
    shadow = normal * 0.5
    it = PairIterator(normal, shadow, batch, epoch, shadow_ratio=numpy.full(5, 2.0, numpy.float32), reg_support_rate=0.0)
    served = []
    for x, y in it:
        assert x.shape == (batch, 1, 1, 5) and torch.equal(y, x * 0.5)
        served.append(x[:, 0, 0, 0].to(torch.int64))
    assert len(served) == epoch * rows // batch == it.batches_served == 7
    stream = torch.cat(served)
    for p in range(2):                                                      # complete passes are permutations
        assert sorted(stream[p * rows:(p + 1) * rows].tolist()) == list(range(rows))
    assert len(set(stream[2 * rows:].tolist())) == stream.numel() - 2 * rows   # the cut third pass: no repeats
    assert not torch.equal(stream[:rows], stream[rows:2 * rows])             # reshuffled every pass
    with pytest.raises(StopIteration):
        it.get_next()
    again = torch.cat([x[:, 0, 0, 0] for x, _ in PairIterator(normal, shadow, batch, epoch, 1.0, 0.0)]).to(torch.int64)
    assert torch.equal(again, stream)                                        # seeded: reproducible
    assert list(PairIterator(normal, shadow, batch, 0, 1.0, 0.0)) == []      # iteration_count * batch < pairs
    with pytest.raises(ValueError):
        PairIterator(normal, shadow[:-1], batch, 1, 1.0, 0.0)


def test_load_op_epoch_count_and_band_selection():
    from hypelcnn_b200.gan.gan_train_for_shadow import load_op

    class DS:
        def get_data_shape(self):
            return [1, 1, 7]            # 6 bands + LiDAR

        def get_casi_band_count(self):
            return 6

    it = load_op(batch_size=32, iteration_count=500, loader=None, data_set=DS(), shadow_map=None,
                 shadow_ratio=numpy.full(6, 2.0, numpy.float32), reg_support_rate=0.0, pairing_method="dummy",
                 device="cpu")
    assert it.normal_data.shape == (2000, 1, 1, 6) and it.epoch == 500 * 32 // 2000 == 8
    x, y = it.get_next()
    assert torch.all(x == 1.0) and torch.all(y == 0.5)                       # DummySampler(2000, 0.5, 2)
    with pytest.raises(ValueError):
        load_op(32, 500, None, DS(), None, None, 0.0, "no_such_pairing", device="cpu")


def test_regularisation_support_augmentation():
    from hypelcnn_b200.gan.gan_train_for_shadow import perform_shadow_augmentation_random as aug
    g = torch.Generator().manual_seed(3)
    ratio = torch.linspace(1.5, 4.0, 8)
    normal, shadow = torch.rand(4000, 1, 1, 8) + 1.0, torch.rand(4000, 1, 1, 8) * 0.3 + 0.1
    n0, s0 = aug(normal, shadow, ratio, 0.0)
    assert n0 is normal and s0 is shadow                                     # u >= 0.01: a rate of 0 never fires
    n1, s1 = aug(normal, shadow, ratio, 1.0, g)                              # u < 0.99 < 1: always fires
    assert torch.equal(n1, shadow * ratio) and torch.allclose(s1, shadow, rtol=1e-6)
    n5, s5 = aug(normal, shadow, ratio, 0.5, g)
    n_swapped = (n5 != normal).flatten(1).any(dim=1)
    s_swapped = (s5 != shadow).flatten(1).any(dim=1)
    assert 0.45 < n_swapped.float().mean() < 0.55 and torch.equal(n5[n_swapped], (shadow * ratio)[n_swapped])
    assert torch.equal(n5[~n_swapped], normal[~n_swapped])                   # whole samples, never single bands
    # the second draw is independent of the first and divides the ALREADY replaced normal spectrum
    both = n_swapped & s_swapped
    only_s = ~n_swapped & s_swapped
    assert both.any() and only_s.any() and 0.2 < s_swapped.float().mean() < 0.55
    assert torch.allclose(s5[only_s], (normal / ratio)[only_s]) and torch.allclose(s5[both], shadow[both], rtol=1e-6)
    flat_n, flat_s = aug(normal.reshape(4000, 8), shadow.reshape(4000, 8), ratio, 0.5, g)      # [B,C] works too
    assert flat_n.shape == (4000, 8) and flat_s.shape == (4000, 8)


class _RecordingOps:
    def __init__(self, log):
        self.log, self.global_step = log, 0

    def global_step_inc_op(self):
        self.global_step += 1
        self.log.append(("inc", self.global_step))
        return self.global_step

    def generator_train_op(self, x, y):
        self.log.append(("gen", self.global_step, tuple(x.shape)))
        return torch.tensor(1.0)

    def discriminator_train_op(self, x, y):
        self.log.append(("dis", self.global_step, tuple(x.shape)))
        return torch.tensor(2.0)


class _RecordingHook:
    def __init__(self, log):
        self.log = log

    def after_create_session(self, session, coord):
        self.log.append(("session",))

    def after_run(self, run_context, run_values):
        self.log.append(("hook", run_context.global_step, [float(r) for r in run_context.results]))


def test_training_loop_order_and_stops():
    from hypelcnn_b200.gan.gan_train_for_shadow import PairIterator, gan_train
    data = torch.ones(6, 1, 1, 4)
    hooks_fn = lambda ops: [ops.generator_train_op, ops.discriminator_train_op]      # noqa: E731
    log = []
    last = gan_train(_RecordingOps(log), PairIterator(data, data * 0.5, 2, 100, 1.0, 0.0), "unused", hooks_fn,
                     hooks=[_RecordingHook(log), None], num_steps=5, save_checkpoint_steps=2,
                     saver=lambda step: log.append(("save", step)))
    assert last == 5
    assert log[0] == ("session",)
    assert log[1:6] == [("inc", 1), ("gen", 1, (2, 1, 1, 4)), ("dis", 1, (2, 1, 1, 4)), ("hook", 1, [1.0, 2.0]), ("inc", 2)]
    assert [e[1] for e in log if e[0] == "save"] == [2, 4, 5]            # every 2 steps + the final state (CheckpointSaverHook.end)
    assert [e[1] for e in log if e[0] == "hook"] == [1, 2, 3, 4, 5]
    # exhausted input ends the run before num_steps (tf OutOfRangeError in the reference)
    log2 = []
    last = gan_train(_RecordingOps(log2), PairIterator(data, data * 0.5, 4, 2, 1.0, 0.0), "unused", hooks_fn, num_steps=50)
    assert last == 3 and sum(1 for e in log2 if e[0] == "inc") == 3            # 2 passes x 6 rows // 4
    assert gan_train(_RecordingOps([]), PairIterator(data, data, 4, 0, 1.0, 0.0), "unused", hooks_fn) is None


def test_synthetic_loader_offers_what_the_gan_run_reads():
    """Shadow map, targets and band measurements of the synthetic GULFPORT loader (host side only)."""
    from hypelcnn_b200.loader.SyntheticGULFPORTDataLoader import SyntheticGULFPORTDataLoader
    loader = SyntheticGULFPORTDataLoader("synthetic:H=60,W=50,samples=300")
    shadow = loader.synthetic_shadow_map()
    assert shadow.shape == (60, 50) and shadow.dtype == numpy.uint8 and set(numpy.unique(shadow)) == {0, 1}
    assert 0.03 < shadow.mean() < 0.6 and loader.synthetic_shadow_map() is shadow
    targets = loader.read_targets("shadow_gen_model/class_result.tif")
    assert targets.shape == (300, 3) and targets[:, 0].max() < 50 and targets[:, 1].max() < 60 and targets[:, 2].max() < 11
    assert numpy.array_equal(targets, loader.read_targets("shadow_gen_model/class_result.tif"))
    assert loader.get_band_measurements().shape == (64,) and loader.shadow_band_ratio().shape == (64,)
    padded, ratio = loader.load_shadow_map(2, None)
    assert padded.shape == (64, 54) and ratio is None
    # the samplers work on it
    from hypelcnn_b200.gan.gan_sampling_methods import neighbourhood_pair_targets, random_pair_targets, target_pair_targets
    normal, shadowed = random_pair_targets(shadow, True)
    assert normal.shape == shadowed.shape and normal.shape[0] > 0
    normal, shadowed, zero_rows = neighbourhood_pair_targets(shadow, 20, 2)
    assert zero_rows == 0 and shadowed.shape[0] == int(shadow.sum()) and normal.shape[0] <= shadowed.shape[0]
    normal, shadowed = target_pair_targets(targets, shadow, 11, report=lambda m: None)
    assert normal is None or normal.shape == shadowed.shape


def test_run_session_glue_with_stand_in_wrappers(tmp_path, monkeypatch, capsys):
    """run_session end to end on CPU tensors: the real loader-facing / sampler / iterator / hook / checkpoint code with
    the device pieces (data set gather, wrapper kernels) replaced by torch stand-ins."""
    from hypelcnn_b200.gan import gan_train_for_shadow as T
    from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANInferenceWrapper
    rng = numpy.random.default_rng(0)
    H, W, C = 30, 26, 6
    smap = (rng.random((H, W)) < 0.25).astype(numpy.uint8)
    true_ratio = numpy.linspace(1.5, 3.0, C).astype(numpy.float32)
    scene = rng.random((H, W, C + 1)).astype(numpy.float32) + 0.5
    scene[smap == 1, :C] /= true_ratio

    class DataSet:
        def get_data_shape(self):
            return [1, 1, C + 1]

        def get_casi_band_count(self):
            return C

        def get_scene_shape(self):
            return [H, W]

        def get_data_points(self, targets_xy):
            t = numpy.asarray(targets_xy)
            return torch.from_numpy(scene[t[:, 1], t[:, 0]]).reshape(-1, 1, 1, C + 1)

    class Loader:
        def load_data(self, neighborhood, normalize):
            assert neighborhood == 0 and normalize is True
            return DataSet()

        def load_shadow_map(self, neighborhood, data_set):
            return smap, true_ratio

        def get_band_measurements(self):
            return numpy.arange(C)

    class Variables:
        def __init__(self):
            self.scale = torch.ones(C)

        def export(self):
            return {"net1/weights": self.scale.numpy().copy()}

    class Trainer:
        def __init__(self):
            self.gen_x2y, self.gen_y2x, self.global_step, self.calls = Variables(), Variables(), 0, []

    class TrainOps:
        def __init__(self, trainer):
            self.trainer = trainer

        def global_step_inc_op(self):
            self.trainer.global_step += 1
            return self.trainer.global_step

        def generator_train_op(self, x, y):                # "learn" the ratio: move the scales towards y / x
            t = self.trainer
            t.gen_x2y.scale += 0.5 * ((y / x).mean(dim=(0, 1, 2)) - t.gen_x2y.scale)
            t.gen_y2x.scale += 0.5 * ((x / y).mean(dim=(0, 1, 2)) - t.gen_y2x.scale)
            t.calls.append(tuple(x.shape))
            return torch.tensor(0.0)

    class Wrapper:
        trainer = None

        def define_model(self, images_x, images_y):
            assert images_x.shape == (16, 1, 1, C)
            self.trainer = Trainer()
            return self.trainer

        def define_loss(self, model):
            return model

        def define_train_ops(self, model, loss, max_number_of_steps, **kwargs):
            assert max_number_of_steps == 200 and set(kwargs) == {"generator_lr", "discriminator_lr", "gen_discriminator_lr"}
            return TrainOps(model)

        def get_train_hooks_fn(self):
            return lambda ops: [ops.generator_train_op]

    class Inference(CycleGANInferenceWrapper):
        def construct_inference_graph(self, input_tensor, is_shadow_graph, clip_invalid_values, copy_extra=0):
            gen = self.forward_generator if is_shadow_graph else self.backward_generator
            return input_tensor * gen.scale

    monkeypatch.setattr(T, "get_wrapper_dict", lambda flags: {"cycle_gan": Wrapper()})
    monkeypatch.setattr(T, "get_infer_wrapper", lambda gan_type, trainer=None: Inference(trainer=trainer))
    real_load_op = T.load_op
    monkeypatch.setattr(T, "load_op", lambda *args, **kwargs: real_load_op(*args, device="cpu", **kwargs))
    flags = T.default_flags(batch_size=16, step=200, validation_steps=20, validation_sample_count=40,
                            pairing_method="random", loader_name="StandInLoader")
    base = str(tmp_path / "gan")
    result = T.run_session(vars(flags), base, loader=Loader())
    log_dir = f"{base}_{T.get_log_suffix(flags)}"
    files = os.listdir(log_dir)
    pairs = int((smap == 0).sum() // smap.sum() * smap.sum())                       # random pairing, multiplied shadows
    batches = (200 * 16 // pairs) * pairs // 16
    assert 0 < batches <= 200
    checkpoints = sorted(int(f[len("model.ckpt-"):-4]) for f in files if f.startswith("model.ckpt-"))
    assert checkpoints == sorted(set(range(20, batches + 1, 20)) | {batches})       # every 20 steps + the final state
    ckpt = numpy.load(os.path.join(log_dir, f"model.ckpt-{checkpoints[-1]}.npz"))
    assert set(ckpt.files) == {"global_step", "ModelX2Y/Generator/net1/weights", "ModelY2X/Generator/net1/weights",
                               "train_state/global_step"}
    assert numpy.allclose(ckpt["ModelX2Y/Generator/net1/weights"], 1 / true_ratio, rtol=0.25)
    shadowed = json.load(open(os.path.join(log_dir, "best_ratio_shadowed.json")))
    assert sorted(p[0] for p in shadowed)[:2] == [21, 41] and len(shadowed) == min(10, (batches - 1) // 20)
    assert len(result) == 2 and all(numpy.isfinite(result)) and result[1] < 0.5    # mean divergence small: ratio learnt
    assert "Best common options:" in capsys.readouterr().out
    assert any("tfevents" in f for f in files) and "band_ratio_deshadowed_21.csv" in files


def test_gan_struct_restores_generators_from_a_run_session_checkpoint(tmp_path):
    """create_gan_struct's shadow_op_initializer (gan/gan_utilities.py:36-37): model_base_dir + ckpt_relative_path names
    a checkpoint of the training run; variables are split by generator scope and handed to the restorer."""
    from hypelcnn_b200.gan.gan_utilities import create_gan_struct, read_generator_checkpoint

    class Wrapper:
        restored = None

        def construct_inference_graph(self, input_tensor, is_shadow_graph, clip_invalid_values, copy_extra=0):
            return ("shadow" if is_shadow_graph else "deshadow", copy_extra, clip_invalid_values)

        def create_generator_restorer(self):
            return self

        def restore(self, forward_values=None, backward_values=None):
            self.restored = (forward_values, backward_values)

    os.makedirs(tmp_path / "shadow_gen_model" / "dcl_gan")
    numpy.savez(tmp_path / "shadow_gen_model" / "dcl_gan" / "model.ckpt-3000.npz", global_step=3000,
                **{"ModelX2Y/Generator/net1/weights": numpy.full((4, 1, 1), 1.0), "ModelX2Y/Generator/net1/biases": numpy.zeros(1),
                   "ModelY2X/Generator/net1/weights": numpy.full((4, 1, 1), 2.0)})
    wrapper = Wrapper()
    struct = create_gan_struct(wrapper, str(tmp_path) + os.sep, "shadow_gen_model/dcl_gan/model.ckpt-3000")
    assert struct.shadow_op("x") == ("shadow", 1, False) and struct.deshadow_op("x") == ("deshadow", 1, False)
    restorer = struct.shadow_op_creater()
    struct.shadow_op_initializer(restorer, None)                    # what InitHook.after_create_session calls
    forward, backward = wrapper.restored
    assert sorted(forward) == ["net1/biases", "net1/weights"] and float(forward["net1/weights"].mean()) == 1.0
    assert list(backward) == ["net1/weights"] and float(backward["net1/weights"].mean()) == 2.0
    numpy.savez(tmp_path / "single.npz", **{"Model/Generator/net7/weights": numpy.ones((2, 1, 1))})
    forward, backward = read_generator_checkpoint(str(tmp_path / "single"))
    assert list(forward) == ["net7/weights"] and backward is None
    with pytest.raises(IOError):
        create_gan_struct(wrapper, str(tmp_path) + os.sep, "nope/model.ckpt-1").shadow_op_initializer(wrapper, None)
    # without a path the values may be handed over directly (an in-process trainer), or nothing happens
    create_gan_struct(wrapper).shadow_op_initializer(wrapper, ({"a": 1}, None))
    assert wrapper.restored == ({"a": 1}, None)
    create_gan_struct(wrapper).shadow_op_initializer(wrapper, None)
    from hypelcnn_b200.loader.SyntheticGULFPORTALTDataLoader import SyntheticGULFPORTALTDataLoader
    loader = SyntheticGULFPORTALTDataLoader(f"synthetic:H=8,W=9,models={tmp_path}")
    assert loader.get_model_base_dir() == str(tmp_path) + os.sep and (loader.h, loader.w) == (8, 9)
    assert sorted(loader.GAN_CHECKPOINTS) == ["cycle_gan", "dcl_cycle_gan", "dcl_gan"]


def test_command_lines_parse_like_the_reference():
    """The flag tables of common/cmd_parser.py and the apps: defaults equal the reference's parsers (golden), values
    parse with the reference's types, unknown arguments are tolerated (parse_known_args)."""
    from hypelcnn_b200.classify import infer_for_classification as I
    from hypelcnn_b200.classify import train_for_classification as T
    from hypelcnn_b200.common.cmd_parser import type_ensure_strtobool
    from hypelcnn_b200.gan import gan_train_for_shadow as G
    gan_flags, extra = G.build_parser().parse_known_args(
        ["--gan_type", "dcl_gan", "--use_identity_loss", "false", "--step", "300", "--tau", "0.1", "--unknown", "x"])
    assert (gan_flags.gan_type, gan_flags.use_identity_loss, gan_flags.step, gan_flags.tau) == ("dcl_gan", False, 300, 0.1)
    assert extra == ["--unknown", "x"] and gan_flags.pairing_method == "random"
    flags, _ = T.build_parser().parse_known_args(["--perform_validation", "True", "--augment_data_with_spectral", "0.01",
                                                  "--augment_data_with_shadow", "dcl_gan", "--neighborhood", "3"])
    assert flags.perform_validation is True and flags.augment_data_with_spectral == 0.01
    assert flags.augment_data_with_shadow == "dcl_gan" and flags.neighborhood == 3 and flags.epoch is None
    assert I.build_parser().parse_known_args([])[0].domain == "all"
    for text, value in [("y", True), ("YES", True), ("t", True), ("on", True), ("1", True), (True, True),
                        ("n", False), ("No", False), ("f", False), ("off", False), ("0", False), (False, False)]:
        assert type_ensure_strtobool(text) is value
    with pytest.raises(ValueError):
        type_ensure_strtobool("maybe")


def test_update_flags_from_json(tmp_path):
    from hypelcnn_b200.gan.gan_train_for_shadow import default_flags, update_flags_from_json
    (tmp_path / "flags.json").write_text(json.dumps({"batch_size": 128, "gan_type": "cut_x2y"}))
    flags = update_flags_from_json(default_flags(step=7), str(tmp_path / "flags.json"))
    assert (flags.batch_size, flags.gan_type, flags.step) == (128, "cut_x2y", 7)


def test_scene_conversion_with_a_stand_in_generator(tmp_path):
    """gan_infer_image_for_shadow.convert_scene: which pixels go through the generator, chunking, de-normalisation
    and dtype — against a direct numpy computation (the generator stand-in halves every band)."""
    from hypelcnn_b200.gan.gan_infer_image_for_shadow import conversion_plan, convert_scene
    assert conversion_plan("shadow") == (True, 0, "shadow") and conversion_plan("deshadow") == (False, 1, "deshadow")
    assert conversion_plan("") == (True, -1, "none") and conversion_plan("anything") == (True, -1, "none")
    rng = numpy.random.default_rng(4)
    H, W, C = 9, 11, 5
    raw = rng.integers(100, 4000, (H, W, C)).astype(numpy.uint16)
    lidar = rng.random((H, W, 1)).astype(numpy.float32)
    casi_min = raw.min(axis=(0, 1))
    casi_max = (raw - casi_min).max(axis=(0, 1))
    normalised = ((raw - casi_min) / casi_max.astype(numpy.float32)).astype(numpy.float32)
    shadow_map = (rng.random((H, W)) < 0.3).astype(numpy.uint8)

    class DataSet:
        def __init__(self):
            self.casi_min, self.casi_max, self.calls = casi_min, casi_max, 0

        def get_scene_shape(self):
            return [H, W]

        def get_casi_band_count(self):
            return C

        def get_unnormalized_casi_dtype(self):
            return numpy.dtype(numpy.uint16)

        def get_data_points(self, targets):
            self.calls += 1
            t = numpy.asarray(targets)
            patch = numpy.concatenate([normalised[t[:, 1], t[:, 0]], lidar[t[:, 1], t[:, 0]]], axis=1)
            return torch.from_numpy(patch).reshape(-1, 1, 1, C + 1)

    halve = lambda x: x * 0.5                                                            # noqa: E731
    for sign, convert_all in [(0, False), (1, False), (-1, False), (-1, True)]:
        data_set = DataSet()
        got = convert_scene(data_set, shadow_map, halve, sign, convert_all, chunk=40)
        selected = numpy.ones((H, W), bool) if convert_all else shadow_map == sign
        want_norm = numpy.where(selected[:, :, None], normalised * numpy.float32(0.5), normalised)
        want = (want_norm * casi_max.astype(numpy.float32) + casi_min.astype(numpy.float32)).astype(numpy.uint16)
        assert got.dtype == numpy.uint16 and got.shape == (H, W, C) and numpy.array_equal(got, want)
        assert data_set.calls == -(-H * W // 40)                                           # one gather per chunk
    untouched = convert_scene(DataSet(), shadow_map, halve, -1, False)
    assert numpy.abs(untouched.astype(int) - raw.astype(int)).max() <= 1                   # round trip of the normalisation


def test_generators_restore_into_either_wrapper_kind(tmp_path):
    from hypelcnn_b200.gan.gan_infer_for_shadow import restore_generators
    numpy.savez(tmp_path / "model.ckpt-7.npz", **{"ModelX2Y/Generator/net1/weights": numpy.ones(3),
                                                   "ModelY2X/Generator/net1/weights": numpy.zeros(3)})

    class Pair:
        backward_generator = object()

        def create_generator_restorer(self):
            return self

        def restore(self, forward_values=None, backward_values=None):
            self.got = (sorted(forward_values), sorted(backward_values))

    class Single:
        def create_generator_restorer(self):
            return self

        def restore(self, values):
            self.got = sorted(values)

    assert restore_generators(Pair(), str(tmp_path / "model.ckpt-7")).got == (["net1/weights"], ["net1/weights"])
    assert restore_generators(Single(), str(tmp_path / "model.ckpt-7.npz")).got == ["net1/weights"]


def test_pair_rows_are_strided_over_ranks_with_equal_counts():
    from hypelcnn_b200.gan.gan_train_for_shadow import PairIterator, shard_pair_iterator
    normal = torch.arange(11, dtype=torch.float32).reshape(11, 1, 1, 1)
    whole = PairIterator(normal, normal * 0.5, 2, 4, 1.0, 0.0)
    assert shard_pair_iterator(whole, 0, 1) is whole
    shares = [shard_pair_iterator(whole, r, 3) for r in range(3)]
    assert [s.normal_data.reshape(-1).tolist() for s in shares] == [[0, 3, 6], [1, 4, 7], [2, 5, 8]]   # 11 // 3 rows each
    assert all(torch.equal(s.shadow_data, s.normal_data * 0.5) and (s.batch_size, s.epoch) == (2, 4) for s in shares)
    assert len({len(list(s)) for s in shares}) == 1                      # same number of iterations on every rank


def test_gan_inference_apps_glue(tmp_path, monkeypatch, capsys):
    """gan_infer_for_shadow.run and gan_infer_image_for_shadow.run with stand-ins for the loader and the device
    wrapper: checkpoint restore, hook with frequency 0, output file naming, which pixels get converted."""
    from hypelcnn_b200.gan import gan_infer_for_shadow as V
    from hypelcnn_b200.gan import gan_infer_image_for_shadow as M
    from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANInferenceWrapper
    from hypelcnn_b200.utilities.tiff_io import imread
    rng = numpy.random.default_rng(2)
    H, W, C = 10, 12, 4
    raw = rng.integers(200, 3000, (H, W, C)).astype(numpy.uint16)
    casi_min = raw.min(axis=(0, 1))
    casi_max = (raw - casi_min).max(axis=(0, 1))
    normalised = ((raw - casi_min) / casi_max.astype(numpy.float32)).astype(numpy.float32)
    shadow_map = (rng.random((H, W)) < 0.3).astype(numpy.uint8)

    class DataSet:
        def __init__(self):
            self.casi_min, self.casi_max = casi_min, casi_max

        def get_data_shape(self):
            return [1, 1, C + 1]

        def get_scene_shape(self):
            return [H, W]

        def get_casi_band_count(self):
            return C

        def get_unnormalized_casi_dtype(self):
            return numpy.dtype(numpy.uint16)

        def get_data_points(self, targets):
            t = numpy.asarray(targets)
            patch = numpy.concatenate([normalised[t[:, 1], t[:, 0]], numpy.zeros((len(t), 1), numpy.float32)], axis=1)
            return torch.from_numpy(patch).reshape(-1, 1, 1, C + 1)

    class Loader:
        def load_data(self, neighborhood, normalize):
            return DataSet()

        def load_shadow_map(self, neighborhood, data_set):
            return shadow_map, numpy.full(C, 2.0, numpy.float32)

        def get_band_measurements(self):
            return numpy.arange(C)

    class Variables:
        scale = None

        def load(self, values):
            self.scale = float(values["net1/weights"].reshape(-1)[0])

    class Wrapper(CycleGANInferenceWrapper):
        def __init__(self):
            self.forward_generator, self.backward_generator = Variables(), Variables()

        def construct_inference_graph(self, input_tensor, is_shadow_graph, clip_invalid_values, copy_extra=0):
            gen = self.forward_generator if is_shadow_graph else self.backward_generator
            return input_tensor * gen.scale

    numpy.savez(tmp_path / "model.ckpt-3000.npz", **{"ModelX2Y/Generator/net1/weights": numpy.full((2, 1, 1), 0.5),
                                                      "ModelY2X/Generator/net1/weights": numpy.full((2, 1, 1), 2.0)})
    for module in (V, M):
        monkeypatch.setattr(module, "get_loader_from_name", lambda name, path: Loader())
        monkeypatch.setattr(module, "get_infer_wrapper", lambda gan_type, bands=None: Wrapper())
    flags = SimpleNamespace(loader_name="L", path="P", neighborhood=0, gan_type="cycle_gan", number_of_samples=30,
                            base_log_path=str(tmp_path / "model.ckpt-3000"), output_path=str(tmp_path),
                            make_them_shadow="shadow", convert_all=False)
    best = V.run(flags)
    out = capsys.readouterr().out
    assert len(best) == 2 and "Validation metrics for shadowed #0" in out and "Validation metrics for deshadowed #0" in out
    assert best[0] == pytest.approx(0.0, abs=1e-6)       # forward generator halves, ratio 2: generated / input * ratio = 1
    path, image = M.run(flags)
    assert os.path.basename(path) == "shadow_image_shadow_3000.tif" and numpy.array_equal(imread(path), image)
    lit = shadow_map == 0
    assert numpy.abs(image[~lit].astype(int) - raw[~lit].astype(int)).max() <= 1           # shadowed pixels untouched
    want = ((normalised * numpy.float32(0.5)) * casi_max.astype(numpy.float32) + casi_min.astype(numpy.float32)).astype(numpy.uint16)
    assert numpy.array_equal(image[lit], want[lit])
    flags.make_them_shadow, flags.convert_all = "deshadow", True
    path, image = M.run(flags)
    assert os.path.basename(path) == "shadow_image_deshadow_3000_all.tif"
    want = ((normalised * numpy.float32(2.0)) * casi_max.astype(numpy.float32) + casi_min.astype(numpy.float32)).astype(numpy.uint16)
    assert numpy.array_equal(image, want)
