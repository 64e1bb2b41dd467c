"""GPU parity of the DUALCNN engine (nnmodel/DUALCNNModel.py) against oracle/dualcnn_ref.py through the C ABI:
variable table, per-level activations, logits (rtol 1e-4), argmax bit-exact, loss and every variable gradient
with the engine's own dropout masks injected into the oracle, and a short training trajectory."""
import numpy
import pytest
import torch

from oracle import dataset_ref as D
from oracle import dualcnn_ref as R
from tests.util import RTOL, ATOL, assert_close, assert_grad_close, synthetic_batch

pytestmark = pytest.mark.gpu

ALG = {"batch_size": 48, "drop_out_ratio": 0.70, "learning_rate": 0.0003, "learning_rate_decay_factor": 0.96,
       "learning_rate_decay_step": 350, "lrelu_alpha": 0.18, "filter_count": 480, "optimizer": "AdamOptimizer",
       "hs_lidar_diff": 1, "l2regularizer_scale": 0.00001}
CASES = {  # name -> (P, C, classes, B, filter_count)
    "small": (7, 21, 6, 16, 64),          # every level one accumulator group
    "c1": (7, 145, 15, 6, 480),           # GRSS2013 shape, alg_param_dualcnn.json: level3 has 480 filters per kernel
    "p5": (5, 12, 4, 130, 32),            # 3x3 HSI window, batch not a multiple of the tile
}


@pytest.fixture(scope="module")
def E():
    from hypelcnn_b200 import engine
    return engine


def _make(E, case):
    P, C, classes, B, fc = CASES[case]
    alg = {**ALG, "filter_count": fc, "batch_size": B}
    eng = E.PatchEngine(P, C, classes, alg, max_batch=B, model="dualcnn")
    v64 = R.init_variables(P, C, classes, alg, seed=7)
    eng.load_variables({k: t.numpy() for k, t in v64.items()})
    x, y = synthetic_batch(B, P, C, classes, seed=11)
    return eng, alg, (P, C, classes, B), v64, x, y


def _lrelu_gates(eng, ref_tensors, B):
    """The LeakyReLU branch the engine took for every activated layer output, recomputed exactly as its kernel does
    (y = z + bias in fp32, branch y > 0) from the engine's own pre-activation — see oracle/dualcnn_ref._lrelu."""
    gates = {}
    for name, t in ref_tensors.items():
        if name == "fc4":
            continue
        z = eng.debug_tensor(name, 1).cpu().reshape(t.shape)
        if name.startswith(("level", "lidar_level")):
            ks = R.level_kernel_sizes(t.shape[1])
            bias = torch.cat([eng.variable(f"nn_core/{name}_conv{k}x{k}/biases").cpu() for k in ks])
        else:
            bias = eng.variable(f"nn_core/{name}/biases").cpu()
        gates[name] = (z + bias) > 0
    return gates


@pytest.mark.parametrize("case", list(CASES))
def test_variable_table_matches_reference(E, case):
    eng, alg, (P, C, classes, B), v64, x, y = _make(E, case)
    specs = R.variable_specs(P, C, classes, alg)
    assert set(eng.variables) == {n for n, _ in specs} and len(eng.variables) == len(specs)
    for n, shape in specs:
        assert tuple(eng.variables[n][2]) == tuple(shape), n
    if case == "c1":
        assert eng.trainable_count == 35652784


@pytest.mark.parametrize("case", list(CASES))
def test_forward_eval_parity(E, case):
    eng, alg, (P, C, classes, B), v64, x, y = _make(E, case)
    ref = R.forward(v64, torch.tensor(x, dtype=torch.float64), classes, alg, False)
    logits, recon = eng.forward(torch.tensor(x).cuda(), False, False, seed=0)
    assert recon is None
    for name, t in ref["tensors"].items():
        if name.startswith("fc"):
            continue
        got = eng.debug_tensor(name, 0).cpu().numpy().reshape(t.shape)
        assert_close(got, t.numpy(), RTOL, 1e-5, f"activation {name}")
    ref32 = R.forward({k: t.float() for k, t in v64.items()}, torch.tensor(x), classes, alg, False)
    floor32 = float((ref32["logits"].double() - ref["logits"]).abs().max())
    assert_close(logits.cpu().numpy(), ref["logits"].numpy(), RTOL, max(ATOL, 3.0 * floor32), "logits")
    pred = E.argmax_confusion(logits)
    assert numpy.array_equal(pred.cpu().numpy(), D.argmax_lowest(ref["logits"].numpy()).astype(numpy.uint8))


@pytest.mark.parametrize("case", list(CASES))
def test_loss_and_gradients_with_dropout(E, case):
    eng, alg, (P, C, classes, B), v64, x, y = _make(E, case)
    seed = 5
    xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
    logits, _ = eng.forward(xd, True, True, seed=seed)
    loss = eng.loss_backward(xd, yd)
    masks = {n: eng.dropout_mask(n, seed, B).cpu() for n in ("fc1", "fc2", "fc3")}
    keep = [float(m.float().mean()) for m in masks.values()]
    assert all(abs(k - alg["drop_out_ratio"]) < 0.25 for k in keep), keep  # keep_prob = drop_out_ratio (:49)
    yl = torch.tensor(y.astype(numpy.int64))
    gates = _lrelu_gates(eng, R.forward(v64, torch.tensor(x, dtype=torch.float64), classes, alg, False)["tensors"], B)
    l64, g64, o64 = R.loss_and_grads(v64, torch.tensor(x, dtype=torch.float64), yl, classes, alg, masks, gates)
    l32, g32, _ = R.loss_and_grads({k: t.float() for k, t in v64.items()}, torch.tensor(x), yl, classes, alg, masks,
                                   gates)
    assert_close(logits.cpu().numpy(), o64["logits"].numpy(), RTOL, 2e-5, "training logits")
    assert abs(loss[0].item() - l64.item()) < 1e-5 * max(1.0, abs(l64.item()))
    assert abs(loss[2].item()) == 0.0  # no reconstruction term in DUALCNN's loss (:87-89)
    for name in eng.variables:
        assert_grad_close(eng.gradient(name).cpu().numpy(), g64[name].numpy(), g32[name].numpy(), f"grad {name}")


def test_training_reduces_loss(E):
    eng, alg, (P, C, classes, B), v64, x, y = _make(E, "small")
    xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
    losses = [eng.train_step(xd, yd)[0].item() for _ in range(12)]
    assert losses[-1] < losses[0], losses
    assert eng.global_step == 12


def test_model_plugin_surface(E):
    """DUALCNNModel keeps the NNModel surface: no-arg constructor, create_tensor_graph, get_loss_func."""
    from hypelcnn_b200.common.common_nn_ops import ModelInputParams, get_model_from_name
    model = get_model_from_name("DUALCNNModel")
    P, C, classes, B, fc = CASES["small"]
    alg = {**ALG, "filter_count": fc, "batch_size": B}
    x, y = synthetic_batch(B, P, C, classes, seed=3)
    xd = torch.tensor(x).cuda()
    out = model.create_tensor_graph(ModelInputParams(xd, None, "/gpu:0", False), classes, alg)
    assert out.y_conv.shape == (B, classes) and out.image_output is None and out.image_original is None
    onehot = torch.nn.functional.one_hot(torch.tensor(y.astype(numpy.int64)), classes).to(torch.uint8).cuda()
    per_sample = model.get_loss_func(out, onehot)
    ref = torch.nn.functional.cross_entropy(out.y_conv.cpu().double(), torch.tensor(y.astype(numpy.int64)), reduction="none")
    assert_close(per_sample.cpu().numpy(), ref.numpy(), 1e-5, 1e-6, "per-sample CE")
