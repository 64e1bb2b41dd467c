"""GPU parity of the CONCNN engine (nnmodel/CONCNNModel.py) against oracle/concnn_ref.py through the C ABI."""
import numpy
import pytest
import torch

from oracle import concnn_ref as R
from oracle import dataset_ref as D
from tests.util import RTOL, ATOL, assert_close, assert_grad_close, synthetic_batch

pytestmark = pytest.mark.gpu

ALG = {"batch_size": 10, "drop_out_ratio": 0.5, "learning_rate": 0.001, "learning_rate_decay_factor": 0.01,
       "learning_rate_decay_step": 33333, "filter_count": 128, "optimizer": ["MomentumOptimizer", 0.9]}
CASES = {  # name -> (P, C, classes, B, filter_count)
    "small": (5, 12, 4, 16, 16),
    "gulfport": (7, 65, 11, 10, 128),   # SURVEY a23's shape: 384-channel flatten, alg_param_concnn.json
    "p3": (3, 20, 5, 130, 8),           # 5x5 kernel wider than the patch; batch not a multiple of the tile
}


@pytest.fixture(scope="module")
def E():
    from hypelcnn_b200 import engine
    return engine


def _make(E, case):
    P, C, classes, B, fc = CASES[case]
    alg = {**ALG, "filter_count": fc, "batch_size": B}
    eng = E.PatchEngine(P, C, classes, alg, max_batch=B, model="concnn")
    v64 = R.init_variables(P, C, classes, alg, seed=3)
    eng.load_variables({k: t.numpy() for k, t in v64.items()})
    x, y = synthetic_batch(B, P, C, classes, seed=13)
    return eng, alg, (P, C, classes, B), v64, x, y


def _gates(eng, ref_tensors):
    gates = {}
    for name, t in ref_tensors.items():
        if name == "fc":
            continue
        scope = "conv0" if name == "net0_out" else name
        z = eng.debug_tensor(scope, 1).cpu().reshape(t.shape)
        bias = torch.cat([eng.variable(f"nn_core/conv0_{k}x{k}/biases").cpu() for k in (1, 3, 5)]) \
            if name == "net0_out" else eng.variable(f"nn_core/{name}/biases").cpu()
        gates[name] = (z + bias) > 0
    return gates


@pytest.mark.parametrize("case", list(CASES))
def test_variable_table_and_forward_eval(E, case):
    eng, alg, (P, C, classes, B), v64, x, y = _make(E, case)
    specs = R.variable_specs(P, C, classes, alg)
    assert set(eng.variables) == {n for n, _ in specs} and len(eng.variables) == len(specs)
    for n, shape in specs:
        assert tuple(eng.variables[n][2]) == tuple(shape), n
    ref = R.forward(v64, torch.tensor(x, dtype=torch.float64), classes, alg, False)
    logits, recon = eng.forward(torch.tensor(x).cuda(), False, False, seed=0)
    assert recon is None
    for name, t in ref["tensors"].items():
        if name == "fc":
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
    seed = 9
    xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
    logits, _ = eng.forward(xd, True, True, seed=seed)
    loss = eng.loss_backward(xd, yd)
    masks = {n: eng.dropout_mask(n, seed, B, rows_per_sample=P * P).cpu().permute(1, 0, 2).reshape(B, P, P, -1)
             for n in ("conv31", "conv32")}
    assert all(abs(float(m.float().mean()) - 0.5) < 0.1 for m in masks.values())
    yl = torch.tensor(y.astype(numpy.int64))
    x64 = torch.tensor(x, dtype=torch.float64)
    gates = _gates(eng, R.forward(v64, x64, classes, alg, False)["tensors"])
    l64, g64, o64 = R.loss_and_grads(v64, x64, yl, classes, alg, masks, gates)
    l32, g32, _ = R.loss_and_grads({k: t.float() for k, t in v64.items()}, torch.tensor(x), yl, classes, alg, masks, gates)
    assert_close(logits.cpu().numpy(), o64["logits"].numpy(), RTOL, 2e-5, "training logits")
    assert abs(loss[0].item() - l64.item()) < 1e-5 * max(1.0, abs(l64.item()))
    for name in eng.variables:
        assert_grad_close(eng.gradient(name).cpu().numpy(), g64[name].numpy(), g32[name].numpy(), f"grad {name}")


def test_training_with_reference_momentum_config(E):
    """alg_param_concnn.json selects ["MomentumOptimizer", 0.9]."""
    eng, alg, (P, C, classes, B), v64, x, y = _make(E, "small")
    xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
    losses = [eng.train_step(xd, yd)[0].item() for _ in range(30)]
    assert losses[-1] < losses[0], losses


def test_model_plugin_surface(E):
    from hypelcnn_b200.common.common_nn_ops import ModelInputParams, get_model_from_name
    model = get_model_from_name("CONCNNModel")
    P, C, classes, B, fc = CASES["small"]
    alg = {**ALG, "filter_count": fc, "batch_size": B}
    x, y = synthetic_batch(B, P, C, classes, seed=3)
    out = model.create_tensor_graph(ModelInputParams(torch.tensor(x).cuda(), None, "/gpu:0", False), classes, alg)
    assert out.y_conv.shape == (B, classes) and out.image_output is None
    onehot = torch.nn.functional.one_hot(torch.tensor(y.astype(numpy.int64)), classes).to(torch.uint8).cuda()
    assert model.get_loss_func(out, onehot).shape == (B,)
