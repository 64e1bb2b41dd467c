"""CPU ORACLE (test infrastructure, NOT product code) — CONCNN forward / loss / gradients.

Restates nnmodel/CONCNNModel.py:23-64 on torch-CPU tensors with the slim / TF semantics of SURVEY.md Appendix A:
conv2d = stride 1, SAME, NHWC, weights [kh,kw,Cin,Cout] + biases then ReLU (slim's default activation_fn);
tf.nn.local_response_normalization defaults (depth_radius 5, bias 1, alpha 1, beta 0.5); dropout(x, keep_prob) with
keep_prob = the JSON's ``drop_out_ratio`` (second positional argument, CONCNNModel.py:53-58); fully_connected without
activation.  PARITY STATUS: unpinned against TensorFlow numerics (TF is not installable here).  Only tests/ imports this."""
import math

import numpy
import torch
import torch.nn.functional as F

CONVS = ["conv11", "conv12", "conv13", "conv21", "conv22", "conv31", "conv32", "conv33"]


def variable_specs(patch, channels, classes, alg):
    f = alg["filter_count"]
    specs = []
    for k in (1, 3, 5):
        specs += [(f"nn_core/conv0_{k}x{k}/weights", (k, k, channels, f)), (f"nn_core/conv0_{k}x{k}/biases", (f,))]
    c = 3 * f
    for name in CONVS:
        specs += [(f"nn_core/{name}/weights", (1, 1, c, c)), (f"nn_core/{name}/biases", (c,))]
    specs += [("nn_core/fc/weights", (patch * patch * c, classes)), ("nn_core/fc/biases", (classes,))]
    return specs


def init_variables(patch, channels, classes, alg, seed=1234, dtype=torch.float64):
    rng = numpy.random.default_rng(seed)
    v = {}
    for name, shape in variable_specs(patch, channels, classes, alg):
        if name.endswith("weights"):
            rf = int(numpy.prod(shape[:-2]))
            limit = math.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))  # slim default xavier_initializer (uniform)
            v[name] = torch.tensor(rng.uniform(-limit, limit, shape), dtype=dtype)
        else:
            v[name] = torch.tensor(rng.uniform(-0.1, 0.1, shape), dtype=dtype)  # reference: zeros; tests randomise
    return v


def lrn(x, radius=5, bias=1.0, alpha=1.0, beta=0.5):
    """tf.nn.local_response_normalization over the last axis: x / (bias + alpha * sum_{|j-c|<=radius} x_j^2)^beta."""
    C = x.shape[-1]
    sq = F.pad(x * x, (radius, radius))
    win = sq.unfold(-1, 2 * radius + 1, 1).sum(-1)
    assert win.shape[-1] == C
    return x / (bias + alpha * win) ** beta


def _relu(y, gate=None):
    return torch.relu(y) if gate is None else torch.where(gate, y, torch.zeros_like(y))


def _conv(x, w, b, gate=None):
    k = w.shape[0]
    y = F.conv2d(x.permute(0, 3, 1, 2).contiguous(), w.permute(3, 2, 0, 1).contiguous(), b,
                 padding=k // 2).permute(0, 2, 3, 1)
    return _relu(y, gate)


def forward(v, x, classes, alg, is_training, dropout_masks=None, gates=None):
    """x [B,P,P,C] -> dict(logits, tensors).  dropout_masks {"conv31": 0/1 [B,P,P,C1], "conv32": ...}; gates: ReLU
    branches per layer output (see oracle/dualcnn_ref._lrelu)."""
    gates = gates or {}
    keep = alg["drop_out_ratio"]
    t = {}
    outs, c0 = [], 0
    for k in (1, 3, 5):
        w = v[f"nn_core/conv0_{k}x{k}/weights"]
        g = gates["net0_out"][..., c0:c0 + w.shape[3]] if "net0_out" in gates else None
        outs.append(_conv(x, w, v[f"nn_core/conv0_{k}x{k}/biases"], g))
        c0 += w.shape[3]
    net = t["net0_out"] = lrn(torch.cat(outs, dim=3))                                    # :36-37

    def conv(name, inp):
        return _conv(inp, v[f"nn_core/{name}/weights"], v[f"nn_core/{name}/biases"], gates.get(name))

    def drop(name, y):
        if is_training and dropout_masks is not None:
            return y * dropout_masks[name].to(y.dtype) / keep
        return y

    net11 = t["conv11"] = lrn(conv("conv11", net))                                       # :40-41
    net12 = t["conv12"] = conv("conv12", net11)
    net13 = t["conv13"] = conv("conv13", net12) + net11                                  # :43-45
    net21 = t["conv21"] = conv("conv21", net13)
    net22 = t["conv22"] = conv("conv22", net21) + net13                                  # :48-50
    net31 = t["conv31"] = drop("conv31", conv("conv31", net22))                          # :52-54
    net32 = t["conv32"] = drop("conv32", conv("conv32", net31))                          # :56-58
    net33 = t["conv33"] = conv("conv33", net32)
    logits = net33.reshape(x.shape[0], -1) @ v["nn_core/fc/weights"] + v["nn_core/fc/biases"]   # :62-63
    t["fc"] = logits
    return {"logits": logits, "tensors": t}


def loss_and_grads(v, x, labels, classes, alg, dropout_masks=None, gates=None):
    vv = {k: a.clone().requires_grad_(True) for k, a in v.items()}
    out = forward(vv, x, classes, alg, True, dropout_masks, gates)
    loss = F.cross_entropy(out["logits"], labels, reduction="mean")
    grads = torch.autograd.grad(loss, list(vv.values()))
    return loss.detach(), {k: g for k, g in zip(vv, grads)}, {"logits": out["logits"].detach()}
