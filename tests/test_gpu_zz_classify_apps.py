"""The classification apps on the device: perform_an_episode (importer -> create_graph -> monitored loop with
validation, test evaluation, TensorBoard summaries, checkpoints) and infer_for_classification.run restoring that
checkpoint into a fresh model and classifying the whole scene — the class image must equal the trained engine's own
argmax over the same patches (checkpoint restore, target order and scatter are the things under test)."""
import json
import os

import numpy
import pytest
import torch

from tests.util import ALG

pytestmark = pytest.mark.gpu

SPEC = "synthetic:H=20,W=24,samples=160"


def test_train_then_infer_whole_scene(tmp_path, capsys):
    from hypelcnn_b200 import engine as E
    from hypelcnn_b200.classify import infer_for_classification as I
    from hypelcnn_b200.classify import train_for_classification as T
    from hypelcnn_b200.common.common_nn_ops import get_model_from_name
    from hypelcnn_b200.loader.SyntheticGRSS2013DataLoader import SyntheticGRSS2013DataLoader
    alg_path = str(tmp_path / "alg_param_small.json")
    json.dump({**ALG, "filter_count": 64}, open(alg_path, "w"))
    flags = T.default_flags(loader_name="SyntheticGRSS2013DataLoader", path=SPEC, neighborhood=3, train_ratio=1.0,
                            test_ratio=0.1, batch_size=32, step=14, algorithm_param_path=alg_path,
                            perform_validation=True, validation_steps=5, save_checkpoint_steps=5,
                            augment_data_with_rotation=True, base_log_path=str(tmp_path))
    assert T.get_log_suffix(flags) == "syntheticgrss2013ldr_hypelcnnmdl_trn100_psmall_7x7"
    model = get_model_from_name(flags.model_name)
    log_dir = os.path.join(flags.base_log_path, T.get_log_suffix(flags))
    algorithm_params = json.load(open(alg_path))
    algorithm_params["batch_size"] = flags.batch_size
    result = T.perform_an_episode(flags, algorithm_params, model, log_dir)
    assert model.engine.global_step == 13                                   # StopAtStepHook(last_step = step - 1)
    assert 0.0 <= result.validation_accuracy <= 1.0 and 0.0 <= result.test_accuracy <= 1.0 and numpy.isfinite(result.loss)
    files = sorted(os.listdir(log_dir))
    assert [f for f in files if f.startswith("model.ckpt-")] == ["model.ckpt-10.safetensors", "model.ckpt-13.safetensors",
                                                                 "model.ckpt-5.safetensors"]
    assert any("tfevents" in f for f in files)
    out = capsys.readouterr().out
    assert "Validation metrics #6 :" in out and "Validation metrics #13 :" in out and "Training step=1," in out

    # whole-scene inference with a FRESH model restored from the log directory
    infer_flags = T.default_flags(loader_name=flags.loader_name, path=SPEC, neighborhood=3, batch_size=50,
                                  algorithm_param_path=alg_path, base_log_path=log_dir, output_path=str(tmp_path))
    infer_flags.domain = "all"
    class_image, colored = I.run(infer_flags)
    assert class_image.shape == (20, 24) and class_image.dtype == numpy.uint8 and class_image.max() < 15
    assert colored.shape == (20, 24, 3) and os.path.exists(tmp_path / "result_raw.tif")
    data_set = SyntheticGRSS2013DataLoader(SPEC).load_data(3, True)         # the same seeded scene
    targets = numpy.array([[x, y] for y in range(20) for x in range(24)], dtype=numpy.int32)
    want = numpy.concatenate([
        E.argmax_confusion(model.engine.forward(data_set.get_data_points(targets[lo:lo + 32]).contiguous(), False,
                                                False)[0].contiguous()).cpu().numpy()
        for lo in range(0, len(targets), 32)]).reshape(20, 24)     # the TRAINED engine, batches within its capacity
    assert numpy.array_equal(class_image, want)

    infer_flags.domain = "sample"
    sample_image, _ = I.run(infer_flags)
    labelled = sample_image != 255
    assert 0 < labelled.sum() <= 320 and numpy.array_equal(sample_image[labelled], want[labelled])
    infer_flags.domain = "gt"
    gt_image, _ = I.run(infer_flags)
    assert gt_image.shape == (20, 24)
    infer_flags.domain = "nope"
    with pytest.raises(ValueError):
        I.run(infer_flags)
