"""CPU checks that pin the DUALCNN / CONCNN / GAN oracles as far as this container allows (TensorFlow is absent):
structure against the reference's own numbers, two independent restatements against each other, and autograd against
finite differences.  No GPU, no native library."""
import numpy
import torch

from oracle import concnn_ref as RC
from oracle import dualcnn_ref as RD
from oracle import gan_ref as RG

DUAL_ALG = {"batch_size": 48, "drop_out_ratio": 0.70, "lrelu_alpha": 0.18, "filter_count": 480, "hs_lidar_diff": 1}
CON_ALG = {"batch_size": 10, "drop_out_ratio": 0.5, "filter_count": 128}


def test_dualcnn_structure_matches_survey_counts():
    specs = RD.variable_specs(7, 145, 15, DUAL_ALG)
    assert sum(int(numpy.prod(s)) for _, s in specs) == 35652784
    shapes = dict(specs)
    assert shapes["nn_core/level3_conv5x5/weights"] == (5, 5, 720, 480)    # 3 kernels (1,3,5) on the cropped 5x5 window
    assert shapes["nn_core/lidar_level1_conv7x7/weights"] == (7, 7, 1, 2)  # LiDAR branch keeps the 7x7 window
    assert shapes["nn_core/fc1/weights"] == (1125 + 1568, 135)             # flat 5*5*45 + 7*7*32
    alg = {**DUAL_ALG, "filter_count": 64}
    v = RD.init_variables(7, 21, 6, alg, seed=1)
    out = RD.forward(v, torch.rand(3, 7, 7, 21, dtype=torch.float64), 6, alg, False)
    assert out["logits"].shape == (3, 6) and out["tensors"]["connector_conv8"].shape == (3, 5, 5, 6)


def test_dualcnn_dropout_keeps_with_probability_drop_out_ratio():
    """slim dropout(net, drop_out_ratio) takes keep_prob positionally (DUALCNNModel.py:49): kept units scale by 1/0.7."""
    alg = {**DUAL_ALG, "filter_count": 32}
    v = RD.init_variables(5, 9, 4, alg, seed=2)
    x = torch.rand(4, 5, 5, 9, dtype=torch.float64)
    ones = {n: torch.ones(4, 4 * m) for n, m in (("fc1", 9), ("fc2", 6), ("fc3", 3))}
    a = RD.forward(v, x, 4, alg, True, ones)["tensors"]["fc1"]
    b = RD.forward(v, x, 4, alg, False)["tensors"]["fc1"]
    assert torch.allclose(a, b / 0.7)


def test_concnn_lrn_matches_brute_force_and_gradcheck():
    x = torch.rand(2, 3, 3, 20, dtype=torch.float64)
    ref = numpy.zeros((2, 3, 3, 20))
    xn = x.numpy()
    for c in range(20):
        lo, hi = max(0, c - 5), min(19, c + 5)
        ref[..., c] = xn[..., c] / numpy.sqrt(1 + (xn[..., lo:hi + 1] ** 2).sum(-1))   # depth_radius 5, bias 1, alpha 1, beta .5
    assert numpy.abs(RC.lrn(x).numpy() - ref).max() < 1e-14
    assert torch.autograd.gradcheck(RC.lrn, (torch.rand(1, 1, 2, 13, dtype=torch.float64, requires_grad=True),), eps=1e-6, atol=1e-6)


def test_concnn_structure():
    specs = dict(RC.variable_specs(7, 65, 11, CON_ALG))
    assert specs["nn_core/conv0_5x5/weights"] == (5, 5, 65, 128) and specs["nn_core/conv33/weights"] == (1, 1, 384, 384)
    assert specs["nn_core/fc/weights"] == (7 * 7 * 384, 11)
    alg = {**CON_ALG, "filter_count": 8}
    v = RC.init_variables(3, 10, 4, alg, seed=0)
    out = RC.forward(v, torch.rand(2, 3, 3, 10, dtype=torch.float64), 4, alg, False)
    assert out["logits"].shape == (2, 4) and (out["tensors"]["conv12"] >= 0).all()   # ReLU


def test_gan_generator_two_restatements_agree():
    """numpy (conv by shifted adds) vs torch (conv1d + autograd) restatements of shadowdata_generator_model."""
    rng = numpy.random.default_rng(0)
    for bands in (64, 48, 10):
        v, flat = {}, []
        for i, k in enumerate(RG.kernel_sizes(bands)):
            w, b = rng.standard_normal((k, 1, 1)) * 0.1, rng.standard_normal(1) * 0.1
            v[f"net{i + 1}/weights"], v[f"net{i + 1}/biases"] = w, b
            flat += [w.ravel(), b]
        x = rng.uniform(0, 1, (7, bands))
        a = RG.generator_forward(x, v)
        b = RG.t_generator(torch.tensor(x), torch.tensor(numpy.concatenate(flat))).numpy()
        assert numpy.abs(a - b).max() < 1e-12
    assert sum(k + 1 for k in RG.kernel_sizes(64)) == 239
    assert RG.kernel_sizes(64, True) == [64, 32, 16, 8]                              # encoder only stops after net4


def test_gan_generator_known_answers():
    """Zero weights (the reference's initializer): the encoder passes residual sums, the full generator outputs 0."""
    x = numpy.full((3, 16), 0.5)
    zero = {}
    for i, k in enumerate(RG.kernel_sizes(16)):
        zero[f"net{i + 1}/weights"], zero[f"net{i + 1}/biases"] = numpy.zeros((k, 1, 1)), numpy.zeros(1)
    assert numpy.array_equal(RG.generator_forward(x, zero), numpy.zeros((3, 16)))
    # encoder: net1 = x, net2 = net1 + net0 = 2x, net3 = net2 + net1 = 3x, net4 = net3 + net2 = 5x
    assert numpy.allclose(RG.generator_forward(x, zero, encoder_only=True), 5 * x)


def test_cyclegan_objective_composition():
    """aux (cycle + identity) is added to BOTH partial generator losses and tfgan sums them: counted twice."""
    torch.manual_seed(0)
    C = 16
    x, y = torch.rand(5, C, dtype=torch.float64), torch.rand(5, C, dtype=torch.float64)
    ng = sum(k + 1 for k in RG.kernel_sizes(C))
    nd = C * C + C + C * C + C + C * (C // 2) + C // 2
    G, F = torch.randn(ng, dtype=torch.float64) * 0.1, torch.randn(ng, dtype=torch.float64) * 0.1
    DY, DX = torch.randn(nd, dtype=torch.float64) * 0.1, torch.randn(nd, dtype=torch.float64) * 0.1
    total, gan, cyc, ident = RG.t_generator_loss(x, y, G, F, DY, DX, 10.0, 0.5)
    gx, fy = RG.t_generator(x, G), RG.t_generator(y, F)
    cyc_once = ((RG.t_generator(gx, F) - x).abs().mean() + (RG.t_generator(fy, G) - y).abs().mean()) / 2
    id_once = (gx - x).abs().mean() + (fy - y).abs().mean()
    assert torch.allclose(cyc, 2 * 10.0 * cyc_once) and torch.allclose(ident, 2 * 0.5 * id_once)
    assert torch.allclose(total, gan + cyc + ident)


def test_fused_xent_gradient_is_softmax_minus_labels():
    """the [TF-lib] convention the oracle encodes: backprop = softmax - labels although the labels sum to S."""
    torch.manual_seed(0)
    fg = torch.randn(3, 4, 2, dtype=torch.float64, requires_grad=True)
    fr = torch.randn(3, 4, 2, dtype=torch.float64)
    RG.t_patchnce(fg, fr, 0.5, True).backward()
    logits = (fg.detach() @ fr.transpose(1, 2) / 0.5).reshape(3, 16)
    bp = (torch.softmax(logits, dim=1) - torch.eye(4, dtype=torch.float64).reshape(1, 16)).reshape(3, 4, 4) / 3
    assert torch.allclose(fg.grad, bp @ fr / 0.5, atol=1e-12)
    fg2 = fg.detach().clone().requires_grad_(True)
    RG.t_patchnce(fg2, fr, 0.5, False).backward()          # exact derivative: S * softmax - labels
    bp2 = (4 * torch.softmax(logits, dim=1) - torch.eye(4, dtype=torch.float64).reshape(1, 16)).reshape(3, 4, 4) / 3
    assert torch.allclose(fg2.grad, bp2 @ fr / 0.5, atol=1e-12)
