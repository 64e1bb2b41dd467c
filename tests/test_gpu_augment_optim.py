"""GPU parity: device-side augmentation (common/common_nn_ops.py:397-440) bit-exact against the oracle replaying the
same per-sample draw; MomentumOptimizer step (common/common_nn_ops.py:223-227) against its formula."""
import ctypes

import numpy
import pytest
import torch

from oracle import dataset_ref as D

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    from hypelcnn_b200 import engine
    return engine


@pytest.mark.parametrize("P,C,B", [(7, 145, 64), (11, 49, 33), (3, 65, 257), (1, 64, 5)])
def test_augmentation_replays_bit_exact(E, P, C, B):
    rng = numpy.random.default_rng(P * 100 + C)
    x = rng.random((B, P, P, C), dtype=numpy.float32)
    out, choices, deltas = E.augment_patches(torch.tensor(x).cuda(), True, True, 0.05, seed=99, return_draw=True)
    ch, dl = choices.cpu().numpy(), deltas.cpu().numpy()
    assert ch[:, 0].max() <= 2 and set(numpy.unique(ch[:, 1:3])) <= {0, 1}
    assert (dl <= 0).all() and (dl >= -0.05).all()            # U(-s, 0)   (:428-430)
    ref = D.augment_patches(x, ch, dl)
    assert numpy.array_equal(out.cpu().numpy(), ref)            # index permutation + one fp32 add: bit-exact
    if B >= 64 and P > 1:                                       # every branch of the draw occurs
        assert set(ch[:, 0]) == {0, 1, 2} and set(ch[:, 1]) == {0, 1} and set(ch[:, 2]) == {0, 1}
    # a different seed gives a different draw, the same seed the same one
    out2 = E.augment_patches(torch.tensor(x).cuda(), True, True, 0.05, seed=99)
    assert torch.equal(out, out2)


def test_augmentation_switches(E):
    x = torch.rand((16, 5, 5, 8), device="cuda")
    assert torch.equal(E.augment_patches(x, False, False, 0.0, seed=1), x)
    only_rot, ch, _ = E.augment_patches(x, True, False, 0.0, seed=3, return_draw=True)
    assert ch[:, 1:3].sum().item() == 0
    assert numpy.array_equal(only_rot.cpu().numpy(), D.augment_patches(x.cpu().numpy(), ch.cpu().numpy()))


def test_momentum_step_matches_formula(E):
    from hypelcnn_b200 import _native as N
    rng = numpy.random.default_rng(4)
    n = 100003
    p, g, a = (rng.standard_normal(n).astype(numpy.float32) for _ in range(3))
    pd, gd, ad = (torch.tensor(t).cuda() for t in (p, g, a))
    lr, mom, scale = numpy.float32(0.01), numpy.float32(0.9), numpy.float32(0.5)
    N.check(N.lib().hyp_momentum_step(ctypes.c_void_p(pd.data_ptr()), ctypes.c_void_p(gd.data_ptr()),
                                      ctypes.c_void_p(ad.data_ptr()), n, float(lr), float(mom), float(scale),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    acc = mom * a + g * scale                                  # tf MomentumOptimizer, use_nesterov=False [TF-lib]
    ref = p - lr * acc
    numpy.testing.assert_allclose(ad.cpu().numpy(), acc, rtol=2e-7, atol=1e-7)
    numpy.testing.assert_allclose(pd.cpu().numpy(), ref, rtol=2e-7, atol=1e-7)


def test_engine_selects_momentum_optimizer(E):
    from tests.util import ALG, synthetic_batch
    alg = {**ALG, "filter_count": 64, "batch_size": 16, "optimizer": ["MomentumOptimizer", 0.9], "drop_out_ratio": 0.0}
    eng = E.PatchEngine(5, 21, 6, alg, max_batch=16)
    eng.init_variables(1)
    x, y = synthetic_batch(16, 5, 21, 6)
    xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
    p0 = eng.params.clone()
    eng.train_step(xd, yd)
    g = eng.grads.clone()
    lr = eng.learning_rate(0)
    assert torch.allclose(eng.params, p0 - lr * g, rtol=1e-6, atol=1e-9)  # first step: accum = g
    assert eng.global_step == 1
