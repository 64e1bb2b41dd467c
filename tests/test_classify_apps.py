"""Host logic of the classification apps (hypelcnn_b200/classify/{monitored_session_runner,train_for_classification,
infer_for_classification}.py) on the CPU: flag defaults, log suffixes and inference target lists against values
produced by executing the reference's own functions (tests/golden/make_golden_gan_host.py); the monitored loop's stop
rule, hook cadence, NaN stop, summaries and checkpoint rotation with a recording train step."""
import json
import os
from types import SimpleNamespace

import numpy
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "gan_host_golden.json")))


def test_flag_defaults_and_log_suffix_equal_the_reference():
    from hypelcnn_b200.classify.train_for_classification import default_flags, get_log_suffix
    ours = vars(default_flags())
    assert {k: ours[k] for k in META["classify_flag_defaults"]} == META["classify_flag_defaults"]
    for case in META["classify_log_suffixes"]:
        assert get_log_suffix(default_flags(**case["overrides"])) == case["suffix"]
    with pytest.raises(KeyError):
        default_flags(nope=1)


def test_inference_target_lists_equal_the_reference():
    from hypelcnn_b200.classify.infer_for_classification import create_all_scene_data, create_sample_data
    from hypelcnn_b200.importer.GeneratorImporter import GeneratorDataInfo
    g = META["infer_targets"]
    scene = create_all_scene_data([3, 5], GeneratorDataInfo(None, None, "the loader", "the data set"))
    assert scene.targets.tolist() == g["all_3x5"] and scene.data is None
    assert (scene.loader, scene.dataset) == (g["all_loader"], g["all_dataset"])
    parts = [GeneratorDataInfo(None, numpy.array(t, dtype=numpy.int64), f"loader{i}", f"set{i}")
             for i, t in enumerate([[[1, 2, 3]], [[4, 5, 6], [7, 8, 9]], [[10, 11, 12]]])]
    sample = create_sample_data(*parts)
    assert sample.targets.tolist() == g["sample"] and str(sample.targets.dtype) == g["sample_dtype"]
    assert (sample.loader, sample.dataset) == (g["sample_loader"], g["sample_dataset"])
    big = create_all_scene_data([349, 1905], parts[0]).targets                     # GRSS2013 scene: vectorised
    assert big.shape == (664845, 3) and big[1905].tolist() == [0, 1, 0] and big[-1].tolist() == [1904, 348, 0]


class _Engine:
    def __init__(self):
        self.saved, self.loaded, self.global_step = [], [], 0

    def save_checkpoint(self, path):
        open(path, "w").write(str(self.global_step))
        self.saved.append(os.path.basename(path))

    def load_checkpoint(self, path):
        self.global_step = int(open(path).read())
        self.loaded.append(os.path.basename(path))

    def export_variables(self):
        return {"nn_core/fc_final/weights": numpy.linspace(-1, 1, 11)}


class _TrainStep:
    def __init__(self, engine, nan_at=None, exhaust_at=None):
        self.engine, self.nan_at, self.exhaust_at, self.last_loss = engine, nan_at, exhaust_at, None

    @property
    def global_step(self):
        return self.engine.global_step

    def run(self):
        if self.exhaust_at is not None and self.engine.global_step >= self.exhaust_at:
            raise StopIteration
        self.engine.global_step += 1
        value = float("nan") if self.engine.global_step == self.nan_at else 1.0 / self.engine.global_step
        self.last_loss = torch.tensor([value, value, 0.0])
        return self.last_loss


class _Metrics:
    def __init__(self):
        self.confusion = numpy.eye(3, dtype=numpy.int32)
        self.accuracy, self.mean_per_class_accuracy, self.kappa = 0.5, 0.4, 0.3


class _Importer:
    def __init__(self, log):
        self.log = log

    def init_tensors(self, session, tensor, nn_params):
        self.log.append(("init", tensor))


def _params(rows):
    return SimpleNamespace(metrics=_Metrics(), data_with_labels=SimpleNamespace(data=torch.zeros(rows, 1)))


def _run(tmp_path, monkeypatch, required_steps, engine=None, **step_args):
    from hypelcnn_b200.classify import monitored_session_runner as M
    log = []

    def accuracy(sess, nn_params, class_range):
        log.append(("accuracy", nn_params.name, engine.global_step))
        return 0.75 if nn_params.name == "validation" else 0.5, None, None, 0.25, 0.6

    monkeypatch.setattr(M, "calculate_accuracy", accuracy)
    engine = engine or _Engine()
    train_step = _TrainStep(engine, **step_args)
    testing, validation, training = _params(4), _params(4), _params(4)
    testing.name, validation.name, training.name = "testing", "validation", "training"
    cross_entropy = lambda: train_step.last_loss                                       # noqa: E731
    summaries = M.add_classification_summaries(cross_entropy, lambda: 3e-4, True, testing, validation)
    result = M.run_monitored_session(cross_entropy, str(tmp_path), range(0, 3), 100, 150, train_step, required_steps,
                                     None, training, "train-tensor", testing, "test-tensor", validation,
                                     "validation-tensor", _Importer(log), '{"a": 1}', '{"b": 2}', summaries=summaries,
                                     engine_of=lambda: engine)
    return result, log, engine


def test_monitored_loop_cadence_checkpoints_and_summaries(tmp_path, monkeypatch, capsys):
    from tensorboard.backend.event_processing.event_file_loader import EventFileLoader
    result, log, engine = _run(tmp_path, monkeypatch, required_steps=321)
    assert engine.global_step == 320                                                   # StopAtStepHook(last_step=320)
    assert log[0] == ("init", "train-tensor")
    validation_at = [e[2] for e in log if e[:2] == ("accuracy", "validation")]
    testing_at = [e[2] for e in log if e[:2] == ("accuracy", "testing")]
    assert validation_at == [151, 301, 320]           # 1 + k * validation_steps, and required_steps - 1
    assert testing_at == [1, 101, 201, 301, 320]      # 1 + k * 100, and once more at the end
    assert engine.saved == ["model.ckpt-100.safetensors", "model.ckpt-200.safetensors", "model.ckpt-300.safetensors",
                            "model.ckpt-320.safetensors"]
    assert (result.validation_accuracy, result.test_accuracy) == (0.75, 0.5) and result.loss == pytest.approx(1 / 320)
    out = capsys.readouterr().out
    assert "Validation metrics #151 : Overall accuracy=0.75, Class based average accuracy=0.6, Kappa=0.25" in out
    assert "Training step=320, Testing accuracy=0.5, loss=0.00313" in out
    events = [e for f in sorted(os.listdir(tmp_path)) if "tfevents" in f
              for e in EventFileLoader(str(tmp_path / f)).Load() if e.HasField("summary")]
    tags = {(e.step, v.tag) for e in events for v in e.summary.value}
    assert (0, "flags") in tags and (0, "algorithm_params") in tags
    for step in (100, 200, 300, 151, 301, 320):
        assert (step, "training_cross_entropy") in tags and (step, "validation_kappa") in tags
    assert (100, "nn_core/fc_final/weights") in tags                                   # log_all_model_variables

    # a second run in the same directory resumes from the newest checkpoint and only does the missing steps
    result, log, engine2 = _run(tmp_path, monkeypatch, required_steps=331)
    assert engine2.loaded == ["model.ckpt-320.safetensors"] and engine2.global_step == 330
    assert engine2.saved == ["model.ckpt-330.safetensors"]


def test_monitored_loop_stops_on_nan_and_on_exhausted_input(tmp_path, monkeypatch, capsys):
    result, log, engine = _run(tmp_path / "nan", monkeypatch, required_steps=1000, nan_at=7)
    assert engine.global_step == 8 and "Model diverged with loss = NaN." in capsys.readouterr().out   # seen one step late
    assert engine.saved == ["model.ckpt-8.safetensors"] and result.loss == pytest.approx(1 / 8)
    result, log, engine = _run(tmp_path / "short", monkeypatch, required_steps=1000, exhaust_at=12)
    assert engine.global_step == 12 and engine.saved == ["model.ckpt-12.safetensors"]


def test_checkpoint_rotation_keeps_twenty(tmp_path):
    from hypelcnn_b200.classify.monitored_session_runner import CheckpointSaver
    engine = _Engine()
    saver = CheckpointSaver(str(tmp_path), lambda: engine, 10)
    for step in range(1, 301):
        engine.global_step = step
        saver.after_run(step)
    kept = [s for s, _ in saver.existing()]
    assert kept == list(range(110, 301, 10)) and len(kept) == 20
    assert saver.restore_latest() == 300 and engine.loaded == ["model.ckpt-300.safetensors"]
    from hypelcnn_b200.classify.infer_for_classification import latest_checkpoint
    assert latest_checkpoint(str(tmp_path)).endswith("model.ckpt-300.safetensors")
    assert latest_checkpoint("/some/file.safetensors") == "/some/file.safetensors"
    with pytest.raises(IOError):
        latest_checkpoint(str(tmp_path / ".."  / "empty" if (tmp_path / ".." / "empty").mkdir() is None else ""))


def test_non_chief_ranks_write_nothing(tmp_path, monkeypatch):
    from hypelcnn_b200.classify import monitored_session_runner as M
    monkeypatch.setattr(M, "calculate_accuracy", lambda sess, nn_params, class_range: (0.5, None, None, 0.1, 0.2))
    engine = _Engine()
    train_step = _TrainStep(engine)
    testing, validation, training = _params(4), _params(4), _params(4)
    cross_entropy = lambda: train_step.last_loss                                       # noqa: E731
    summaries = M.add_classification_summaries(cross_entropy, lambda: 3e-4, False, testing, validation)
    result = M.run_monitored_session(cross_entropy, str(tmp_path), range(0, 3), 10, 20, train_step, 41, None, training,
                                     "t", testing, "t", validation, "t", _Importer([]), "{}", "{}", summaries=summaries,
                                     engine_of=lambda: engine, is_chief=False)
    assert engine.global_step == 40 and engine.saved == [] and result.validation_accuracy == 0.5
    assert not [f for f in os.listdir(tmp_path) if f.startswith("model.ckpt")]


def test_training_split_is_strided_over_ranks():
    from collections import namedtuple
    from hypelcnn_b200.parallel import shard_training_data
    Target = namedtuple("Target", ["data", "labels"])
    Lazy = namedtuple("Lazy", ["data", "targets", "loader", "dataset"])
    full = Target(data=torch.arange(10).reshape(10, 1), labels=torch.arange(10) % 3)
    parts = [shard_training_data(full, r, 3) for r in range(3)]
    assert [p.data.reshape(-1).tolist() for p in parts] == [[0, 3, 6], [1, 4, 7], [2, 5, 8]]      # equal counts, 9 dropped
    assert all(torch.equal(p.labels, p.data.reshape(-1) % 3) for p in parts)
    assert shard_training_data(full, 0, 1) is full
    lazy = shard_training_data(Lazy(None, numpy.arange(14).reshape(7, 2), "L", "D"), 1, 2)
    assert lazy.targets.tolist() == [[2, 3], [6, 7], [10, 11]] and (lazy.loader, lazy.dataset) == ("L", "D")
    with pytest.raises(ValueError):
        shard_training_data(full, 3, 3)


def test_perform_an_episode_plumbing(monkeypatch, capsys):
    """perform_an_episode with the device pieces replaced: what reaches the importer, create_graph and the monitored
    loop (steps from --epoch, augmentation switches, validation on / off), and the reported result."""
    from collections import namedtuple
    from hypelcnn_b200.classify import train_for_classification as T
    from hypelcnn_b200.common.common_nn_ops import TrainingResult
    Target = namedtuple("Target", ["data", "labels"])
    seen = {}

    class Importer:
        def read_data_set(self, *args):
            seen["read"] = args
            mk = lambda n: Target(torch.zeros(n, 3, 3, 5), torch.zeros(n, dtype=torch.uint8))     # noqa: E731
            return mk(100), mk(10), mk(50), {"simple": "SIMPLE"}, range(0, 4), [8, 9], numpy.zeros((4, 3), numpy.uint8)

        def convert_data_to_tensor(self, test, train, validation, class_range):
            seen["convert"] = (test.data.shape[0], train.data.shape[0], validation.data.shape[0])
            return SimpleNamespace(dataset="TEST"), SimpleNamespace(dataset="TRAIN"), SimpleNamespace(dataset="VAL")

        def requires_separate_validation_branch(self):
            return True

    def create_graph(train_ds, test_ds, val_ds, class_range, batch_size, prefetch, device_id, epochs, **kw):
        seen["graph"] = (train_ds, test_ds, val_ds, class_range, batch_size, prefetch, device_id, epochs, kw)
        return "CE", "LR", SimpleNamespace(), SimpleNamespace(), SimpleNamespace(), SimpleNamespace(allreduce=None)

    def run_monitored_session(*args, **kw):
        seen["run"] = (args, kw)
        return TrainingResult(validation_accuracy=0.75, test_accuracy=0.5, loss=1.25)

    monkeypatch.setattr(T, "get_importer_from_name", lambda name: Importer())
    monkeypatch.setattr(T, "create_graph", create_graph)
    monkeypatch.setattr(T, "run_monitored_session", run_monitored_session)
    monkeypatch.setattr(T, "add_classification_summaries", lambda *a: ("SUMMARIES",) + a)
    flags = T.default_flags(loader_name="L", path="P", neighborhood=1, epoch=3, augment_data_with_shadow="simple",
                            augment_data_with_rotation=True, perform_validation=True, validation_steps=7,
                            save_checkpoint_steps=9, train_ratio=0.2, test_ratio=0.1)
    result = T.perform_an_episode(flags, {"batch_size": 16}, "MODEL", "/logs/x")
    assert seen["read"] == ("L", "P", 0.2, 0.1, 1, True) and seen["convert"] == (10, 100, 50)
    train_ds, test_ds, val_ds, class_range, batch, prefetch, device, epochs, kw = seen["graph"]
    assert (train_ds, test_ds, val_ds, class_range, batch, prefetch, device, epochs) == ("TRAIN", "TEST", "VAL", range(0, 4), 16, 1000, "/gpu:0", 3)
    info = kw["augmentation_info"]
    assert (info.shadow_struct, info.perform_shadow_augmentation, info.perform_rotation_augmentation,
            info.perform_reflection_augmentation, info.perform_spectral_augmentation,
            info.augmentation_random_threshold) == ("SIMPLE", True, True, False, None, 0.5)
    assert kw["model"] == "MODEL" and kw["algorithm_params"] == {"batch_size": 16}
    args, kwargs = seen["run"]
    assert args[:7] == ("CE", "/logs/x", range(0, 4), 9, 7, args[5], 100 * 3 // 16)         # steps from --epoch
    assert args[12] is not None and kwargs["is_chief"] is True and kwargs["summaries"][0] == "SUMMARIES"
    assert args[8].data_with_labels.data.shape[0] == 100 and args[10].data_with_labels.data.shape[0] == 10
    assert (result.validation_accuracy, result.test_accuracy, result.loss) == (0.75, 0.5, 1.25)
    out = capsys.readouterr().out
    assert "Steps: 18," in out and "Validation accuracy=0.75, Testing accuracy=0.5, loss=1.25" in out
    assert "Mean testing accuracy result: (0.5) +- (0), Loss result: (1.25) +- (0)" in out
    # validation off: the loop gets no validation branch, the result carries none
    flags.perform_validation, flags.epoch = False, None
    result = T.perform_an_episode(flags, {"batch_size": 16}, "MODEL", "/logs/x")
    assert seen["run"][0][12] is None and seen["run"][0][6] == flags.step and result.validation_accuracy is None
    flags.device = "cpu"
    with pytest.raises(RuntimeError):
        T.perform_an_episode(flags, {"batch_size": 16}, "MODEL", "/logs/x")


def test_inference_app_plumbing(tmp_path, monkeypatch):
    """infer_for_classification.run with stand-ins for the device pieces: domain handling, checkpoint choice, the
    decoder excluded from the restore, the class image written as TIFF."""
    from hypelcnn_b200.classify import infer_for_classification as I
    from hypelcnn_b200.importer.GeneratorImporter import GeneratorDataInfo
    from hypelcnn_b200.utilities.tiff_io import imread
    seen = {}

    class DataSet:
        device = "cpu"

        def get_data_points(self, targets):
            return torch.zeros(len(targets), 3, 3, 5)

    data_set = DataSet()

    class Importer:
        def read_data_set(self, *args):
            seen["read"] = args
            mk = lambda t: GeneratorDataInfo(None, numpy.array(t), "LOADER", data_set)            # noqa: E731
            return (mk([[1, 1, 0]]), mk([[2, 2, 1]]), mk([[3, 3, 2], [0, 1, 2]]), None, range(0, 3), [4, 5],
                    numpy.array([[255, 0, 0], [0, 255, 0], [0, 0, 255]], numpy.uint8))

        def convert_data_to_tensor(self, test, train, validation, class_range):
            seen["targets"] = numpy.asarray(validation.targets)
            return SimpleNamespace(dataset="T"), SimpleNamespace(dataset="TR"), SimpleNamespace(dataset="V")

        def init_tensors(self, session, tensor, nn_params):
            seen["init"] = tensor.dataset

    class Engine:
        def load_checkpoint(self, path, exclude_prefixes=()):
            seen["restore"] = (os.path.basename(path), exclude_prefixes)

    class Model:
        engine = None

        def engine_for(self, x, alg):
            seen["engine_for"] = (tuple(x.shape), alg["batch_size"], self._class_count)
            self.engine = Engine()

    def perform_prediction(session, nn_params, class_map):
        targets = numpy.asarray(nn_params.data_with_labels.targets)
        class_map[torch.as_tensor(targets[:, 1]), torch.as_tensor(targets[:, 0])] = 1
        return class_map

    monkeypatch.setattr(I, "GeneratorImporter", Importer)
    monkeypatch.setattr(I, "simple_nn_iterator", lambda dataset, batch: ("ITER", dataset, batch))
    monkeypatch.setattr(I, "perform_prediction", perform_prediction)
    (tmp_path / "alg.json").write_text(json.dumps({"filter_count": 8}))
    for step in (5, 20, 100):
        (tmp_path / f"model.ckpt-{step}.safetensors").write_text("x")
    flags = SimpleNamespace(loader_name="L", path="P", neighborhood=1, algorithm_param_path=str(tmp_path / "alg.json"),
                            batch_size=64, model_name="unused", base_log_path=str(tmp_path), output_path=str(tmp_path),
                            domain="all")
    image, colored = I.run(flags, model=Model())
    assert seen["read"] == ("L", "P", 0.1, 0, 1, True) and seen["targets"].shape == (20, 3) and seen["init"] == "V"
    assert seen["engine_for"] == ((1, 3, 3, 5), 64, 3) and seen["restore"] == ("model.ckpt-100.safetensors", ("image_gen_net_",))
    assert image.shape == (4, 5) and (image == 1).all() and colored.shape == (4, 5, 3) and (colored == [0, 255, 0]).all()
    assert numpy.array_equal(imread(str(tmp_path / "result_raw.tif")), image)
    assert numpy.array_equal(imread(str(tmp_path / "result_colorized.tif")), colored)
    flags.domain = "sample"
    image, _ = I.run(flags, model=Model())
    assert seen["targets"].tolist() == [[1, 1, 0], [2, 2, 1], [3, 3, 2], [0, 1, 2]]     # (train, test, validation) order
    assert (image == 1).sum() == 4 and (image == 255).sum() == 16
    flags.algorithm_param_path = None
    with pytest.raises(IOError):
        I.run(flags, model=Model())


def test_loss_watch_reports_one_step_late():
    from hypelcnn_b200.classify.monitored_session_runner import _LossWatch
    watch = _LossWatch()
    seen = [watch.submit(torch.tensor([v, 0.0, 0.0])) for v in (1.0, float("nan"), 2.0, 3.0)]
    assert seen == [False, False, True, False] and watch.check() is False
    assert watch.submit(torch.tensor(float("nan"))) is False and watch.check() is True and watch.check() is False
    assert watch.submit(None) is False and watch.submit(float("nan")) is False and watch.submit(None) is True
