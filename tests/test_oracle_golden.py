"""Pin the CPU oracle against fixtures produced by executing the reference's own code
(tests/golden/make_golden.py) and against the known answers of SURVEY.md §8a."""
import json
import os

import numpy
import pytest
import torch

from oracle import dataset_ref as D
from oracle import hypelcnn_ref as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ALG = {"batch_size": 48, "drop_out_ratio": 0.70, "filter_count": 480, "learning_rate": 0.0003,
       "learning_rate_decay_factor": 0.96, "learning_rate_decay_step": 350, "lrelu_alpha": 0.18,
       "optimizer": "AdamOptimizer", "bn_decay": 0.95, "l2regularizer_scale": 0.00001,
       "spectral_hierarchy_level": 3, "spatial_hierarchy_level": 3, "degradation_coeff": 3, "use_residual": True}


def test_scale_in_to_out_tables_match_reference():
    gold = json.load(open(os.path.join(GOLD, "scale_in_to_out.json")))
    assert len(gold) >= 20
    for key, g in gold.items():
        cin, cout = map(int, key.split("->"))
        assert R.scale_in_to_out_index(cin, cout) == g["idx"], key


def test_scale_in_to_out_known_answers():
    i = R.scale_in_to_out_index
    assert i(145, 120)[:12] == [0, 1, 2, 4, 5, 6, 7, 8, 10, 11, 12, 13] and i(145, 120)[-4:] == [140, 141, 143, 144]
    assert sum(i(145, 120)) == 8627 and sum(i(145, 480)) == 34726 and i(145, 480)[-4:] == [144] * 4
    assert i(120, 240)[:4] == [0, 0, 1, 1] and i(120, 360)[:6] == [0, 0, 0, 1, 1, 1]
    assert i(480, 240) == [2 * j for j in range(240)] and sum(i(480, 120)) == 28560
    assert sum(i(49, 120)) == 2914 and sum(i(65, 120)) == 3867


def test_fc_stage_sizes_known_answers():
    assert R.fc_stage_sizes(2940, 15, 3) == [980, 326, 108]
    assert R.fc_stage_sizes(10890, 20, 3) == [3630, 1210, 403, 134]
    assert R.fc_stage_sizes(270, 11, 3) == [90]
    assert R.fc_stage_sizes(2940, 15, 6) == [490]
    assert R.fc_stage_sizes(2940, 15, 9) == [326]
    assert R.fc_stage_sizes(2940, 15, 27) == []
    assert R.fc_stage_sizes(2940, 15, 2) == [1470, 735, 367, 183, 91, 45]


@pytest.mark.parametrize("key", ["c2_train", "c2_eval", "c3_train", "c5_train", "tiny_train", "nonres_train"])
def test_plan_matches_reference_graph_trace(key):
    """The oracle's layer plan == the op sequence the reference's HYPELCNNModel emits."""
    g = json.load(open(os.path.join(GOLD, "hypelcnn_graph_trace.json")))[key]
    plan = R.build_plan(g["patch"], g["channels"], g["classes"], g["alg"], g["is_training"])
    ref_layers = [e for e in g["trace"] if e["op"] in ("conv2d", "fully_connected")]
    mine = [l for l in plan if l.kind in ("conv", "fc")]
    assert len(ref_layers) == len(mine)
    for e, l in zip(ref_layers, mine):
        assert e["scope"] == l.scope and e["cin"] == l.cin and e["cout"] == l.cout, (e, l.scope)
        assert e["normalizer"] == "batch_norm"
        if e["op"] == "conv2d":
            assert e["kernel"] == [l.kernel, l.kernel]
        act = {"<lambda>": "lrelu", "sigmoid": "sigmoid", None: None}[e["activation"]]
        assert act == l.act, e["scope"]
    # residual structure: every gather/repeat in the trace corresponds to a non-identity residual
    n_resample = sum(1 for e in g["trace"] if e["op"] in ("gather", "repeat"))
    mine_resample = 0
    for l in plan:
        for _, c in l.residuals:
            if R.scale_in_to_out_index(c, l.cout) != list(range(l.cout)):
                mine_resample += 1
    assert n_resample == mine_resample
    n_add = sum(1 for e in g["trace"] if e["op"] == "add")
    assert n_add == sum(len(l.residuals) for l in plan)
    drops = [e for e in g["trace"] if e["op"] == "dropout"]
    assert len(drops) == sum(1 for l in plan if l.dropout)
    for d in drops:
        assert d["keep_prob"] == 1 - g["alg"]["drop_out_ratio"]


def test_param_count_known_answer():  # 8 160 297 trainable (train graph)
    n = sum(int(numpy.prod(s)) for _, s, k in R.variable_specs(7, 145, 15, ALG) if k in ("weights", "beta"))
    assert n == 8160297
    n3 = sum(int(numpy.prod(s)) for _, s, k in R.variable_specs(11, 49, 20, ALG) if k in ("weights", "beta"))
    assert n3 == 54407798


def test_useful_flops_known_answer():  # 157.16 MFLOP useful fwd / patch
    assert abs(R.useful_flops_per_patch(7, 145, 15, ALG) / 1e6 - 157.16) < 0.05


def test_dataset_matches_reference_bit_exact():
    g = numpy.load(os.path.join(GOLD, "dataset_golden.npz"))
    s = D.SceneRef(g["same_casi"].copy(), g["same_lidar"].copy(), int(g["same_n"]), True)
    assert numpy.array_equal(s.casi.astype(numpy.float64), g["same_padded_casi"])
    assert numpy.array_equal(s.lidar.astype(numpy.float64), g["same_padded_lidar"])
    assert s.data_shape() == list(g["same_shape"]) and s.scene_shape() == list(g["same_scene"])
    pts = numpy.concatenate([g["same_pts"], numpy.zeros((len(g["same_pts"]), 1), numpy.int32)], axis=1)
    data, _ = D.gather_patches(s, pts)
    assert data.dtype == numpy.float32 and numpy.array_equal(data, g["same_patches"])
    s2 = D.SceneRef(g["raw_casi"].copy(), g["same_lidar"].copy(), 1, False)
    assert numpy.array_equal(numpy.stack([s2.get_data_point(int(p[0]), int(p[1])) for p in g["same_pts"]]),
                             g["raw_patches"])
    s3 = D.SceneRef(g["raw_casi"].copy(), None, 1, True)
    assert s3.data_shape() == list(g["hsi_shape"])
    assert numpy.array_equal(numpy.stack([s3.get_data_point(int(p[0]), int(p[1])) for p in g["same_pts"]]),
                             g["hsi_patches"])


@pytest.mark.parametrize("n", [2, 3, 5])
def test_grss2018_gather_matches_reference_bit_exact(n):
    g = numpy.load(os.path.join(GOLD, "dataset_golden.npz"))
    s = D.SceneRef2018(g[f"g18_{n}_casi"].copy(), g[f"g18_{n}_lidar"].copy(), n, True)
    assert numpy.array_equal(s.casi.astype(numpy.float64), g[f"g18_{n}_padded_casi"])
    got = numpy.stack([s.get_data_point(int(p[0]), int(p[1])) for p in g[f"g18_{n}_pts"]]).astype(numpy.float32)
    assert numpy.array_equal(got, g[f"g18_{n}_patches"])


def test_confusion_metrics_match_reference():
    g = numpy.load(os.path.join(GOLD, "dataset_golden.npz"))
    conf = g["conf"]
    rec, prec = D.class_accuracies(conf, range(0, 15))
    assert numpy.array_equal(rec, g["conf_recall"]) and numpy.array_equal(prec, g["conf_precision"])
    assert abs(D.kappa(conf) - float(g["conf_kappa"])) < 1e-12
    rng = numpy.random.default_rng(0)
    lab, pred = rng.integers(0, 15, 1000), rng.integers(0, 15, 1000)
    c = D.confusion_matrix(lab, pred, 15)
    assert c.sum() == 1000 and c[lab[0], pred[0]] >= 1 and c.dtype == numpy.int32
    assert abs(D.overall_accuracy(c) - float((lab == pred).mean())) < 1e-12
    assert D.argmax_lowest(numpy.array([[1.0, 3.0, 3.0], [2.0, 2.0, 1.0]])).tolist() == [1, 0]


def test_conv_layout_restatement():
    """F.conv2d-based SAME conv == literal NHWC / [kh,kw,Cin,Cout] loop form."""
    torch.manual_seed(0)
    for k in (1, 3, 5, 7):
        x = torch.randn(2, 7, 7, 5, dtype=torch.float64)
        w = torch.randn(k, k, 5, 4, dtype=torch.float64)
        assert torch.allclose(R.conv2d_same_nhwc(x, w), R.conv2d_same_nhwc_naive(x, w), atol=1e-12)


def _tiny():
    alg = {**ALG, "filter_count": 32, "drop_out_ratio": 0.0}
    v = R.init_variables(3, 10, 4, alg, seed=1, dtype=torch.float64)
    rng = numpy.random.default_rng(1234)
    x = torch.tensor(rng.random((6, 3, 3, 10)), dtype=torch.float64)
    y = torch.tensor(rng.integers(0, 4, 6))
    return alg, v, x, y


def test_oracle_forward_shapes_and_bn_semantics():
    alg, v, x, y = _tiny()
    out = R.forward(v, x, 4, alg, True)
    assert out["logits"].shape == (6, 4) and out["recon"].shape == (6, 90)
    # training-mode BN output of fc_final (no activation): zero mean, biased var/(var+eps) per column
    lg = out["logits"]
    assert torch.allclose(lg.mean(0), torch.zeros(4, dtype=torch.float64), atol=1e-9)
    z = out["pre"]["fc_final"]
    var = z.var(0, unbiased=False)
    assert torch.allclose(lg.var(0, unbiased=False), var / (var + R.BN_EPS), atol=1e-9)
    # moving stats: decay 0.95, Bessel-corrected variance
    mv = out["new_variables"]["nn_core/fc_final/BatchNorm/moving_variance"]
    assert torch.allclose(mv, 0.95 + 0.05 * z.var(0, unbiased=True), atol=1e-12)
    ev = R.forward(v, x, 4, alg, False)
    assert ev["recon"] is None and "image_gen_net_1" not in ev["pre"]


def test_oracle_gradients_finite_difference():
    alg, v, x, y = _tiny()
    loss, g, _ = R.loss_and_grads(v, x, y, 4, alg)
    rng = numpy.random.default_rng(0)
    for name in ["nn_core/conv_enc_0/weights", "nn_core/connector_1_conv3x3/weights", "nn_core/fc_final/BatchNorm/beta",
                 "nn_core/image_gen_net_4/weights", "nn_core/conv_dec_1/BatchNorm/beta"]:
        flat = v[name].reshape(-1)
        for _ in range(2):
            i = int(rng.integers(0, flat.numel()))
            h = 1e-6
            old = flat[i].item()
            flat[i] = old + h
            lp = R.per_sample_loss(*(lambda o: (o["logits"], o["recon"]))(R.forward(v, x, 4, alg, True)), x, y).mean()
            flat[i] = old - h
            lm = R.per_sample_loss(*(lambda o: (o["logits"], o["recon"]))(R.forward(v, x, 4, alg, True)), x, y).mean()
            flat[i] = old
            fd = (lp - lm).item() / (2 * h)
            assert abs(fd - g[name].reshape(-1)[i].item()) < 1e-6 + 1e-4 * abs(fd), (name, fd)


def test_oracle_fp32_tracks_fp64():
    alg, v, x, y = _tiny()
    v32 = {k: t.float() for k, t in v.items()}
    a = R.forward(v, x, 4, alg, True)["logits"]
    b = R.forward(v32, x.float(), 4, alg, True)["logits"].double()
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)


def test_adam_tf1_known_answer():
    p, g = torch.tensor([1.0], dtype=torch.float64), torch.tensor([0.5], dtype=torch.float64)
    p1, m, v = R.adam_tf1(p, g, torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64), 0.001, 1)
    # t=1: m=0.05 v=0.00025 lr_t=0.001*sqrt(0.001)/0.1 ; step = lr_t*m/(sqrt(v)+eps)
    expect = 1.0 - (0.001 * (0.001 ** 0.5) / 0.1) * 0.05 / (0.00025 ** 0.5 + 1e-8)
    assert abs(p1.item() - expect) < 1e-15
    assert R.learning_rate(ALG, 349) == 0.0003 and abs(R.learning_rate(ALG, 700) - 0.0003 * 0.96 ** 2) < 1e-18
