"""GPU parity of the shadow GAN generator (gan/shadow_data_models.py:43-90) and the per-pixel augmenter built on it
(gan/wrappers/gan_common.py:282-304, gan/gan_utilities.py:18-43) against oracle/gan_ref.py."""
import numpy
import pytest
import torch

from oracle import gan_ref as R

pytestmark = pytest.mark.gpu


def _vars(bands, encoder_only, seed, scale=0.05):
    rng = numpy.random.default_rng(seed)
    v = {}
    for i, k in enumerate(R.kernel_sizes(bands, encoder_only)):
        v[f"net{i + 1}/weights"] = rng.standard_normal((k, 1, 1)) * scale
        v[f"net{i + 1}/biases"] = rng.standard_normal(1) * scale
    return v


@pytest.mark.parametrize("bands,encoder_only,n", [(64, False, 1000), (64, True, 33), (48, False, 257), (144, False, 65)])
def test_generator_forward_parity(bands, encoder_only, n):
    from hypelcnn_b200.gan.shadow_data_models import GeneratorVariables, shadowdata_generator_model
    v = _vars(bands, encoder_only, bands + n)
    gv = GeneratorVariables(bands, encoder_only)
    gv.load(v)
    assert {k: a.shape for k, a in gv.export().items()} == {k: a.shape for k, a in v.items()}
    assert gv.flat.numel() == (239 if (bands, encoder_only) == (64, False) else gv.flat.numel())
    rng = numpy.random.default_rng(1)
    x = rng.uniform(0.02, 0.5, (n, 1, 1, bands)).astype(numpy.float32)
    got = shadowdata_generator_model(torch.tensor(x).cuda(), encoder_only, False, gv).cpu().numpy()
    ref = R.generator_forward(x.reshape(n, bands).astype(numpy.float64), v, encoder_only).reshape(x.shape)
    numpy.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-6)


def test_untrained_generator_is_the_reference_zero_init():
    """weights_initializer = zeros (shadow_data_models.py:47): net7 = tanh(0 + 0) = 0 for every input."""
    from hypelcnn_b200.gan.shadow_data_models import GeneratorVariables, shadowdata_generator_model
    gv = GeneratorVariables(64)
    x = torch.rand((10, 1, 1, 64), device="cuda")
    assert torch.count_nonzero(shadowdata_generator_model(x, False, True, gv)).item() == 0


@pytest.mark.parametrize("clip,is_shadow", [(False, True), (True, True), (True, False)])
def test_patch_augmenter_per_pixel(clip, is_shadow):
    from hypelcnn_b200.gan.gan_utilities import GeneratorInferenceWrapper
    from hypelcnn_b200.gan.shadow_data_models import GeneratorVariables
    bands, B, P = 64, 20, 3
    vf, vb = _vars(bands, False, 5, 0.08), _vars(bands, False, 6, 0.08)
    fwd, bwd = GeneratorVariables(bands), GeneratorVariables(bands)
    wrapper = GeneratorInferenceWrapper(fwd, bwd)
    wrapper.create_generator_restorer().restore(vf, vb)
    rng = numpy.random.default_rng(2)
    x = rng.uniform(0.0, 0.6, (B, P, P, bands + 1)).astype(numpy.float32)
    got = wrapper.construct_inference_graph(torch.tensor(x).cuda(), is_shadow, clip, copy_extra=1).cpu().numpy()
    ref = R.inference_for_matrix_input(x.astype(numpy.float64), vf if is_shadow else vb, is_shadow, clip, copy_extra=1)
    numpy.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-6)
    assert numpy.array_equal(got[..., -1], x[..., -1])  # LiDAR passes through (gan_utilities.py:35)


def test_shadow_structs_feed_the_training_iterator():
    """C5's data path: the (frozen) augmenter applied with probability augmentation_random_threshold inside the
    training iterator (common/common_nn_ops.py:408-422), and the simple ratio struct (gan_utilities.py:18-28)."""
    from hypelcnn_b200.common.common_nn_ops import AugmentationInfo, training_nn_iterator
    from hypelcnn_b200.gan.gan_utilities import GeneratorInferenceWrapper, create_gan_struct, create_simple_shadow_struct
    from hypelcnn_b200.gan.shadow_data_models import GeneratorVariables
    bands, n, P = 64, 64, 3
    rng = numpy.random.default_rng(3)
    x = torch.tensor(rng.uniform(0.1, 0.6, (n, P, P, bands + 1)).astype(numpy.float32)).cuda()
    y = torch.zeros((n, 11), dtype=torch.uint8, device="cuda")
    ratio = numpy.linspace(1.5, 4, bands).astype(numpy.float32)
    simple = create_simple_shadow_struct(ratio)
    shadowed = simple.shadow_op(x)
    assert torch.allclose(shadowed[..., :-1], x[..., :-1] / torch.tensor(ratio).cuda()) and torch.equal(shadowed[..., -1], x[..., -1])
    assert torch.allclose(simple.deshadow_op(shadowed), x, rtol=1e-6)
    fwd, bwd = GeneratorVariables(bands), GeneratorVariables(bands)
    fwd.load(_vars(bands, False, 9, 0.08))
    struct = create_gan_struct(GeneratorInferenceWrapper(fwd, bwd))
    info = AugmentationInfo(struct, True, False, False, False, 0.5)
    it = training_nn_iterator((x, y), info, 32, None, "/gpu:0", 1000)
    bx, by = it.get_next()
    assert bx.shape == (32, P, P, bands + 1)
    changed = (bx[..., :-1] != bx[..., :-1]).any()  # no NaNs
    assert not bool(changed)
