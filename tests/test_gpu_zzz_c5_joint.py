"""BASELINE configs[4] end to end on one device: train a DCLGAN shadow generator pair on the GULFPORTALT-shaped
synthetic scene (gan_train_for_shadow.run_session), file its checkpoint where the loader's shadow_creator_dict expects
it, then train HYPELCNN through perform_an_episode with ``--augment_data_with_shadow=dcl_gan``: InitHook restores the
generators, every training batch passes the per-pixel generator with probability ``augmentation_random_threshold``."""
import json
import os
import shutil

import numpy
import pytest

from tests.util import ALG

pytestmark = pytest.mark.gpu


def test_dclgan_augmenter_feeds_hypelcnn_training(tmp_path, monkeypatch):
    from hypelcnn_b200.classify import train_for_classification as T
    from hypelcnn_b200.common.common_nn_ops import get_model_from_name
    from hypelcnn_b200.gan import gan_train_for_shadow as G
    from hypelcnn_b200.gan import gan_utilities
    from hypelcnn_b200.loader.SyntheticGULFPORTALTDataLoader import SyntheticGULFPORTALTDataLoader
    models = tmp_path / "models"
    spec = f"synthetic:H=40,W=36,samples=300,models={models}"
    gan_flags = G.default_flags(gan_type="dcl_gan", pairing_method="random", batch_size=32, step=60,
                                validation_steps=25, validation_sample_count=40,
                                loader_name="SyntheticGULFPORTALTDataLoader")
    G.run_session(vars(gan_flags), str(tmp_path / "gan"), loader=SyntheticGULFPORTALTDataLoader(spec))
    gan_log = f"{tmp_path / 'gan'}_{G.get_log_suffix(gan_flags)}"
    os.makedirs(models / "shadow_gen_model" / "dcl_gan")
    shutil.copy(os.path.join(gan_log, "model.ckpt-25.npz"), models / "shadow_gen_model" / "dcl_gan" / "model.ckpt-3000.npz")
    trained = numpy.load(os.path.join(gan_log, "model.ckpt-25.npz"))

    restored = []
    real_read = gan_utilities.read_generator_checkpoint

    def spy(path):
        restored.append(path)
        return real_read(path)

    monkeypatch.setattr(gan_utilities, "read_generator_checkpoint", spy)
    alg_path = str(tmp_path / "alg_param_c5.json")
    json.dump({**ALG, "filter_count": 64}, open(alg_path, "w"))
    flags = T.default_flags(loader_name="SyntheticGULFPORTALTDataLoader", path=spec, neighborhood=3, train_ratio=1.0,
                            test_ratio=0.1, batch_size=32, step=9, algorithm_param_path=alg_path,
                            perform_validation=True, validation_steps=100, save_checkpoint_steps=100,
                            augment_data_with_shadow="dcl_gan", augmentation_random_threshold=0.5,
                            base_log_path=str(tmp_path))
    assert T.get_log_suffix(flags) == "syntheticgulfportaltldr_hypelcnnmdl_trn100_pc5_7x7_dcl_gan_aug050"
    algorithm_params = json.load(open(alg_path))
    algorithm_params["batch_size"] = flags.batch_size
    model = get_model_from_name("HYPELCNNModel")
    result = T.perform_an_episode(flags, algorithm_params, model, str(tmp_path / "classify"))
    assert model.engine.global_step == 8 and (model.engine.channels, model.engine.classes) == (65, 11)
    assert numpy.isfinite(result.loss) and 0.0 <= result.validation_accuracy <= 1.0
    assert len(restored) == 1 and restored[0].endswith(os.path.join("shadow_gen_model", "dcl_gan", "model.ckpt-3000"))
    assert "ModelX2Y/Generator/net1/weights" in trained.files and "ModelY2X/Generator/net7/biases" in trained.files
