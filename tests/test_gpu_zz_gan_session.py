"""A whole GAN training run on the device (hypelcnn_b200/gan/gan_train_for_shadow.run_session) over the synthetic
GULFPORT-shaped scene: sampler -> one batched patch gather -> resident pair matrices -> train ops -> validation hooks
(generator launch + band-ratio statistics on the device) -> best-ratio files, TensorBoard scalars, generator
checkpoints.  Also the device path of the samplers against the per-point DataSet contract."""
import json
import os

import numpy
import pytest
import torch

pytestmark = pytest.mark.gpu


def _loader(spec="synthetic:H=64,W=60,samples=600"):
    from hypelcnn_b200.loader.SyntheticGULFPORTDataLoader import SyntheticGULFPORTDataLoader
    return SyntheticGULFPORTDataLoader(spec)


def test_samplers_gather_on_the_device_like_point_by_point():
    from hypelcnn_b200.gan import gan_sampling_methods as S
    loader = _loader()
    data_set = loader.load_data(0, True)
    shadow_map, shadow_ratio = loader.load_shadow_map(0, data_set)
    assert shadow_ratio.shape == (64,) and numpy.allclose(shadow_ratio, loader.shadow_band_ratio(), rtol=0.3)

    class PointWise:                                    # the same scene through the per-point contract only
        get_data_shape, get_scene_shape = data_set.get_data_shape, data_set.get_scene_shape
        get_casi_band_count = data_set.get_casi_band_count

        @staticmethod
        def get_data_point(x, y):
            return data_set.get_data_point(x, y)

    for sampler in (S.NeighborhoodBasedSampler(4, 1), S.RandomBasedSampler(True), S.TargetBasedSampler(5)):
        normal, shadow = sampler.get_sample_pairs(data_set, loader, shadow_map)
        assert normal.is_cuda and shadow.is_cuda and normal.dtype == torch.float32 and normal.shape[1:] == (1, 1, 65)
        if isinstance(sampler, S.NeighborhoodBasedSampler):
            want_normal, want_shadow = sampler.get_sample_pairs(PointWise, loader, shadow_map)
            assert numpy.array_equal(normal.cpu().numpy(), want_normal)
            assert numpy.array_equal(shadow.cpu().numpy(), want_shadow)
        else:
            assert normal.shape == shadow.shape


@pytest.mark.parametrize("gan_type,pairing", [("cycle_gan", "random"), ("cut_x2y", "neighbour"), ("dcl_gan", "target")])
def test_run_session(tmp_path, gan_type, pairing, capsys):
    from hypelcnn_b200.gan.gan_train_for_shadow import default_flags, get_log_suffix, run_session
    # step * batch_size >= scene pixels >= pair count, so load_op's epoch count is >= 1 and the stream holds >= 60 batches
    flags = default_flags(gan_type=gan_type, pairing_method=pairing, batch_size=32, step=120, validation_steps=10,
                          validation_sample_count=50, loader_name="SyntheticGULFPORTDataLoader",
                          regularization_support_rate=0.3)
    base = str(tmp_path / "run")
    result = run_session(vars(flags), base, loader=_loader())
    log_dir = f"{base}_{get_log_suffix(flags)}"
    assert len(result) == 2 and all(r is not None and numpy.isfinite(r) for r in result)
    files = sorted(os.listdir(log_dir))
    assert "model.ckpt-10.npz" in files and "model.ckpt-20.npz" in files
    # the pair stream (epochs x pairs // batch) may end before --step iterations: everything below follows the last
    # iteration that really ran = the final checkpoint (CheckpointSaverHook.end)
    last_step = max(int(f[len("model.ckpt-"):-len(".npz")]) for f in files if f.startswith("model.ckpt-"))
    assert 30 < last_step <= 120
    validated = list(range(11, last_step + 1, 10))                       # iterations 1 + k * validation_steps
    suffixes = ["shadowed"] if gan_type == "cut_x2y" else ["shadowed", "deshadowed"]
    for suffix in suffixes:
        best = json.load(open(os.path.join(log_dir, f"best_ratio_{suffix}.json")))
        # a best-10 list: WHICH validation is evicted once there are more than ten depends on how the divergence moves
        # during training, so only what every run guarantees is asserted
        kept = [p[0] for p in best]
        assert len(kept) == min(10, len(validated)) and len(set(kept)) == len(kept) and set(kept) <= set(validated)
        assert all(numpy.isfinite(p[1]) for p in best) and [p[1] for p in best] == sorted(p[1] for p in best)
        assert f"band_ratio_{suffix}_11.csv" in files
    assert any("tfevents" in f for f in files)
    ckpt = numpy.load(os.path.join(log_dir, "model.ckpt-20.npz"))
    names = [n for n in ckpt.files if n.endswith("net1/weights")]
    assert names and int(ckpt["global_step"]) == 20
    if gan_type == "cycle_gan":
        assert any(numpy.abs(ckpt[n]).sum() > 0 for n in names)              # the zero-initialised generator moved
    assert "Validation metrics for shadowed #11" in capsys.readouterr().out
    # the final state is always written (the pair stream may end before --step iterations), and a second run over the
    # same log dir continues from it: generators, discriminators, Adam moments and the step clocks
    # (MonitoredTrainingSession(checkpoint_dir=log_dir))
    final = numpy.load(os.path.join(log_dir, f"model.ckpt-{last_step}.npz"))
    state_keys = [n for n in final.files if "train_state/" in n]
    assert any(n.endswith("dis_params") for n in state_keys) and any(n.endswith("_m") for n in state_keys)
    if gan_type != "dcl_gan":
        run_session(vars(flags), base, loader=_loader())                     # same flags: as many iterations again
        assert "Restored" in capsys.readouterr().out
        resumed = max(int(f[len("model.ckpt-"):-len(".npz")]) for f in os.listdir(log_dir) if f.startswith("model.ckpt-"))
        assert resumed > last_step
        again = numpy.load(os.path.join(log_dir, f"model.ckpt-{resumed}.npz"))
        assert int(again["global_step"]) == resumed
        moved = [n for n in state_keys if n.endswith("dis_params") and not numpy.array_equal(final[n], again[n])]
        assert moved                                                         # training went on from the restored weights
        steps = [n for n in state_keys if n.endswith("gen_steps") or n.endswith("clock_gen")]
        assert steps and all(int(again[n]) == int(final[n]) + resumed - last_step for n in steps)
