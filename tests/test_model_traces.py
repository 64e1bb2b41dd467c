"""DUALCNN / CONCNN structure pinned to the reference's own code.  tests/golden/{dualcnn,concnn}_graph_trace.json were
recorded by executing nnmodel/DUALCNNModel.py and nnmodel/CONCNNModel.py against recording stubs
(tests/golden/make_golden_models.py).  Here the trace is (1) compared with the oracles' variable tables and (2) REPLAYED
as a dataflow program over torch tensors with the oracles' variables and primitive ops; the replay must reproduce the
oracles' forward pass, so any difference in wiring — kernel sets per level, concat order, the band split and crop,
residual adds, LRN and dropout positions, which layers have an activation — shows up as a numeric mismatch."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import concnn_ref as RC
from oracle import dualcnn_ref as RD

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DUAL = json.load(open(os.path.join(GOLD, "dualcnn_graph_trace.json")))
CON = json.load(open(os.path.join(GOLD, "concnn_graph_trace.json")))


def trace_variables(trace):
    specs = {}
    for e in trace:
        if e["op"] == "conv2d":
            assert e["normalizer"] is None                       # no BatchNorm in these models -> biases
            specs[f"nn_core/{e['scope']}/weights"] = (e["kernel"][0], e["kernel"][1], e["cin"], e["cout"])
            specs[f"nn_core/{e['scope']}/biases"] = (e["cout"],)
        elif e["op"] == "fully_connected":
            specs[f"nn_core/{e['scope']}/weights"] = (e["cin"], e["cout"])
            specs[f"nn_core/{e['scope']}/biases"] = (e["cout"],)
    return specs


def replay(case, v, x, activation, lrn):
    t = {case["input_id"]: x}
    for e in case["trace"]:
        op = e["op"]
        if op == "split":
            for out, part in zip(e["outs"], torch.split(t[e["in"]], e["sizes"], dim=e["axis"])):
                t[out] = part
        elif op == "slice":
            y = t[e["in"]]
            for axis, start, stop in e["crop"]:
                y = y.narrow(axis, start, stop - start)
            t[e["out"]] = y
        elif op == "conv2d":
            w, b = v[f"nn_core/{e['scope']}/weights"], v[f"nn_core/{e['scope']}/biases"]
            k = e["kernel"][0]
            y = F.conv2d(t[e["in"]].permute(0, 3, 1, 2), w.permute(3, 2, 0, 1).contiguous(), b, padding=k // 2)
            t[e["out"]] = activation(y.permute(0, 2, 3, 1), e["activation"])
        elif op == "fully_connected":
            y = t[e["in"]] @ v[f"nn_core/{e['scope']}/weights"] + v[f"nn_core/{e['scope']}/biases"]
            t[e["out"]] = activation(y, e["activation"])
        elif op == "concat":
            t[e["out"]] = torch.cat([t[i] for i in e["ins"]], dim=e["axis"])
        elif op == "flatten":
            t[e["out"]] = t[e["in"]].reshape(t[e["in"]].shape[0], -1)
        elif op == "add":
            t[e["out"]] = t[e["a"]] + t[e["b"]]
        elif op == "dropout":
            t[e["out"]] = t[e["in"]]                             # evaluated without masks, like the oracle below
        elif op == "lrn":
            assert e["args"] == [] and e["kwargs"] == {}         # TF defaults: depth_radius 5, bias 1, alpha 1, beta 0.5
            t[e["out"]] = lrn(t[e["in"]])
        else:
            raise AssertionError(f"unknown op {op}")
    return t[case["y_conv"]]


@pytest.mark.parametrize("case", DUAL, ids=lambda c: f"P{c['patch']}C{c['channels']}")
def test_dualcnn_oracle_follows_the_reference_graph(case):
    P, C, classes, alg = case["patch"], case["channels"], case["classes"], case["alg"]
    assert trace_variables(case["trace"]) == {n: tuple(s) for n, s in RD.variable_specs(P, C, classes, alg)}
    drops = [e for e in case["trace"] if e["op"] == "dropout"]
    assert [d["keep_prob"] for d in drops] == [alg["drop_out_ratio"]] * 3        # passed positionally AS keep_prob
    first = case["trace"][:2]
    assert first[0]["op"] == "split" and first[0]["sizes"] == [C - 1, 1]
    d = alg["hs_lidar_diff"]
    assert first[1]["op"] == "slice" and first[1]["crop"] == [[1, d, P - d], [2, d, P - d]]


def test_dualcnn_replay_matches_the_oracle_numerically():
    case = DUAL[1]                                                               # filter_count 64: seconds on CPU
    P, C, classes, alg = case["patch"], case["channels"], case["classes"], case["alg"]
    v = RD.init_variables(P, C, classes, alg, seed=3)
    x = torch.rand(3, P, P, C, dtype=torch.float64)
    got = replay(case, v, x, lambda y, name: y if name is None else RD._lrelu(y, alg["lrelu_alpha"]), None)
    ref = RD.forward(v, x, classes, alg, False)["logits"]
    assert got.shape == ref.shape and torch.allclose(got, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("case", CON, ids=lambda c: f"P{c['patch']}C{c['channels']}")
def test_concnn_oracle_follows_the_reference_graph(case):
    P, C, classes, alg = case["patch"], case["channels"], case["classes"], case["alg"]
    assert trace_variables(case["trace"]) == {n: tuple(s) for n, s in RC.variable_specs(P, C, classes, alg)}
    ops = [(e["op"], e.get("scope")) for e in case["trace"]]
    assert ops[3:7] == [("concat", None), ("lrn", None), ("conv2d", "conv11"), ("lrn", None)]
    assert [e["keep_prob"] for e in case["trace"] if e["op"] == "dropout"] == [alg["drop_out_ratio"]] * 2
    assert [e["activation"] for e in case["trace"] if e["op"] in ("conv2d", "fully_connected")][-1] is None
    if alg["filter_count"] > 16:
        return
    v = RC.init_variables(P, C, classes, alg, seed=5)
    x = torch.rand(4, P, P, C, dtype=torch.float64)
    got = replay(case, v, x, lambda y, name: y if name is None else RC._relu(y), RC.lrn)
    ref = RC.forward(v, x, classes, alg, False)["logits"]
    assert got.shape == ref.shape and torch.allclose(got, ref, rtol=1e-10, atol=1e-10)


# ---------------------------------------------------------------------------------------------------------------
# HYPELCNN: the same replay for the headline model (tests/golden/hypelcnn_graph_trace.json, make_golden.py)
HYP = json.load(open(os.path.join(GOLD, "hypelcnn_graph_trace.json")))


def replay_hypelcnn(case, v, x, alg):
    from oracle import hypelcnn_ref as R
    t = {case["input_id"]: x}

    def normalise_activate(z, e):
        assert e["normalizer"] == "batch_norm"
        params = e.get("norm_params", {"decay": alg["bn_decay"], "is_training": case["is_training"]})   # FCs: arg_scope's
        assert params["decay"] == alg["bn_decay"] and params["is_training"] == case["is_training"]
        scope = f"nn_core/{e['scope']}/BatchNorm/"
        y = R.batch_norm(z, v[scope + "beta"], v[scope + "moving_mean"], v[scope + "moving_variance"],
                         params["is_training"], alg["bn_decay"])[0]
        if e["activation"] == "<lambda>":
            return R.leaky_relu(y, alg["lrelu_alpha"])
        return torch.sigmoid(y) if e["activation"] == "sigmoid" else y

    for e in case["trace"]:
        op = e["op"]
        if op == "conv2d":
            t[e["out"]] = normalise_activate(R.conv2d_same_nhwc(t[e["in"]], v[f"nn_core/{e['scope']}/weights"]), e)
        elif op == "fully_connected":
            t[e["out"]] = normalise_activate(t[e["in"]] @ v[f"nn_core/{e['scope']}/weights"], e)
        elif op == "gather":
            t[e["out"]] = t[e["in"]].index_select(e["axis"], torch.tensor(e["indices"]))
        elif op == "repeat":
            t[e["out"]] = t[e["in"]].repeat_interleave(e["repeats"], dim=e["axis"])
        elif op == "add":
            t[e["out"]] = t[e["a"]] + t[e["b"]]
        elif op == "concat":
            t[e["out"]] = torch.cat([t[i] for i in e["ins"]], dim=e["axis"])
        elif op == "flatten":
            t[e["out"]] = t[e["in"]].reshape(t[e["in"]].shape[0], -1)
        elif op == "dropout":
            t[e["out"]] = t[e["in"]]                             # compared against the oracle with dropout off
        else:
            raise AssertionError(f"unknown op {op}")
    recon = None if case["image_output"] is None else t[case["image_output"]]
    return t[case["y_conv"]], recon


@pytest.mark.parametrize("key", sorted(HYP))
def test_hypelcnn_oracle_follows_the_reference_graph(key):
    from oracle import hypelcnn_ref as R
    case = HYP[key]
    P, C, classes = case["patch"], case["channels"], case["classes"]
    alg = {**case["alg"], "drop_out_ratio": 0.0}
    v = R.init_variables(P, C, classes, alg, seed=7, dtype=torch.float64)
    for name in v:                                               # non-trivial BN state on both sides
        if name.endswith("moving_mean") or name.endswith("beta"):
            v[name] = v[name] + 0.05 * torch.randn(v[name].shape, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    x = torch.rand(4, P, P, C, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    logits, recon = replay_hypelcnn(case, v, x, alg)
    ref = R.forward(v, x, classes, alg, case["is_training"], update_moving=False)
    assert torch.allclose(logits, ref["logits"], rtol=1e-9, atol=1e-9)
    if recon is not None:
        assert torch.allclose(recon.reshape(ref["recon"].shape), ref["recon"], rtol=1e-9, atol=1e-9)
    else:
        assert ref.get("recon") is None
