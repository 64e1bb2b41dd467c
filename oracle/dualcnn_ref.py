"""CPU ORACLE (test infrastructure, NOT product code) — DUALCNN forward / loss / gradients.

A restatement on torch-CPU tensors of the graph nnmodel/DUALCNNModel.py:11-104 builds, with the slim
semantics of SURVEY.md Appendix A (conv2d: stride 1, SAME, NHWC, weights [kh,kw,Cin,Cout] + biases, then the
activation; fully_connected alike; dropout(x, keep_prob) = x * mask / keep_prob with keep_prob = the JSON's
``drop_out_ratio`` because it is passed positionally, DUALCNNModel.py:49).  Only ``tests/`` imports this.

PARITY STATUS: unpinned against TensorFlow numerics (TensorFlow / tf_slim are not installable here and the
reference ships no tests); the layer list, scopes, filter counts, kernel sizes and the crop are read off the
reference source cited per function.
"""
import math

import numpy
import torch
import torch.nn.functional as F


def level_kernel_sizes(patch):
    """DUALCNNModel._create_a_level (:92-104): square odd kernels 1, 3, .. <= patch."""
    return [k for k in range(1, patch + 1) if k % 2 == 1]


def variable_specs(patch, channels, classes, alg):
    """[(tf variable name, shape)] in creation order: HSI branch (:58-85), LiDAR branch (:36-43), FCs (:46-55)."""
    diff = alg["hs_lidar_diff"]
    ph = patch - 2 * diff if patch > 1 else patch
    Fc = alg["filter_count"]
    specs = []

    def level(name, cin, f, p):
        for k in level_kernel_sizes(p):
            specs.append((f"nn_core/{name}_conv{k}x{k}/weights", (k, k, cin, f)))
            specs.append((f"nn_core/{name}_conv{k}x{k}/biases", (f,)))
        return f * len(level_kernel_sizes(p))

    def conv1(name, c):
        specs.append((f"nn_core/{name}/weights", (1, 1, c, c)))
        specs.append((f"nn_core/{name}/biases", (c,)))

    c = channels - 1
    for i, f in enumerate([Fc // 4, Fc // 2, Fc, Fc // 2, Fc // 4, Fc // 8, Fc // 16, Fc // 32]):
        c = level(f"level{i + 1}", c, f, ph)
        conv1(f"connector_conv{i + 1}", c)
    hs_flat = ph * ph * c
    c = 1
    for i, f in enumerate([2, 4, 8]):
        c = level(f"lidar_level{i + 1}", c, f, patch)
        conv1(f"lidar_connector_conv{i + 1}", c)
    flat = hs_flat + patch * patch * c
    for name, n in (("fc1", classes * 9), ("fc2", classes * 6), ("fc3", classes * 3), ("fc4", classes)):
        specs.append((f"nn_core/{name}/weights", (flat, n)))
        specs.append((f"nn_core/{name}/biases", (n,)))
        flat = n
    return specs


def init_variables(patch, channels, classes, alg, seed=1234, dtype=torch.float64, random_biases=True):
    """xavier-uniform weights (slim default); biases are zero in the reference — tests randomise them so that
    the bias path is exercised."""
    rng = numpy.random.default_rng(seed)
    v = {}
    for name, shape in variable_specs(patch, channels, classes, alg):
        if name.endswith("weights"):
            rf = int(numpy.prod(shape[:-2]))
            limit = math.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))
            v[name] = torch.tensor(rng.uniform(-limit, limit, shape), dtype=dtype)
        else:
            v[name] = torch.tensor(rng.uniform(-0.1, 0.1, shape) if random_biases else numpy.zeros(shape), dtype=dtype)
    return v


def _lrelu(y, alpha, gate=None):
    """leaky_relu(alpha) = max(y, alpha y) (A.4).  ``gate`` (bool, y's shape) fixes the branch per element: max is
    not differentiable at 0 and an element within fp32 round-off of 0 may legitimately take either branch, so the
    gradient tests hand over the branch the implementation under test took."""
    if gate is None:
        return torch.maximum(y, alpha * y)
    return torch.where(gate, y, alpha * y)


def _conv(x, w, b, alpha, gate=None):
    """slim conv2d, NHWC in / out (App. A.1) followed by leaky_relu."""
    k = w.shape[0]
    y = F.conv2d(x.permute(0, 3, 1, 2).contiguous(), w.permute(3, 2, 0, 1).contiguous(), b,
                 padding=k // 2).permute(0, 2, 3, 1)
    return _lrelu(y, alpha, gate)


def forward(v, x, classes, alg, is_training, dropout_masks=None, gates=None):
    """x [B,P,P,C] -> dict(logits=[B,classes], tensors={scope: activation}).  dropout_masks: {"fc1": 0/1 [B,n], ..}
    (training only; None = no dropout, i.e. keep everything unscaled as in eval).  gates: {scope: bool tensor} LeakyReLU
    branches per layer output (levels: the concatenated tensor), see _lrelu."""
    gates = gates or {}
    alpha, diff = alg["lrelu_alpha"], alg["hs_lidar_diff"]
    P = x.shape[1]
    hs, lidar = x[..., :-1], x[..., -1:]                      # tf.split [band-1, 1]        (:19-21)
    if P > 1:
        hs = hs[:, diff:P - diff, diff:P - diff, :]           # crop by hs_lidar_diff      (:23-26)
    tensors = {}

    def level(net, name):
        outs, c0 = [], 0
        for k in level_kernel_sizes(net.shape[1]):
            w = v[f"nn_core/{name}_conv{k}x{k}/weights"]
            g = gates[name][..., c0:c0 + w.shape[3]] if name in gates else None
            outs.append(_conv(net, w, v[f"nn_core/{name}_conv{k}x{k}/biases"], alpha, g))
            c0 += w.shape[3]
        tensors[name] = torch.cat(outs, dim=3)                # tf.concat axis=3           (:103)
        return tensors[name]

    def conv1(net, name):
        tensors[name] = _conv(net, v[f"nn_core/{name}/weights"], v[f"nn_core/{name}/biases"], alpha, gates.get(name))
        return tensors[name]

    net = hs
    for i in range(8):
        net = conv1(level(net, f"level{i + 1}"), f"connector_conv{i + 1}")
    hs_net = net
    net = lidar
    for i in range(3):
        net = conv1(level(net, f"lidar_level{i + 1}"), f"lidar_connector_conv{i + 1}")
    B = x.shape[0]
    net = torch.cat([hs_net.reshape(B, -1), net.reshape(B, -1)], dim=1)   # NHWC flatten + concat axis=1 (:31)
    keep = alg["drop_out_ratio"]
    for name in ("fc1", "fc2", "fc3"):
        y = net @ v[f"nn_core/{name}/weights"] + v[f"nn_core/{name}/biases"]
        net = _lrelu(y, alpha, gates.get(name))
        if is_training and dropout_masks is not None:
            net = net * dropout_masks[name].to(net.dtype) / keep
        tensors[name] = net
    logits = net @ v["nn_core/fc4/weights"] + v["nn_core/fc4/biases"]      # activation_fn=None (:53)
    tensors["fc4"] = logits
    return {"logits": logits, "tensors": tensors}


def loss_and_grads(v, x, labels, classes, alg, dropout_masks=None, gates=None):
    """mean_B(softmax CE) (get_loss_func :87-89 + common_nn_ops.py:214) and d loss / d variable."""
    vv = {k: t.clone().requires_grad_(True) for k, t in v.items()}
    out = forward(vv, x, classes, alg, True, dropout_masks, gates)
    loss = F.cross_entropy(out["logits"], labels, reduction="mean")
    grads = torch.autograd.grad(loss, list(vv.values()))
    return loss.detach(), {k: g for k, g in zip(vv, grads)}, {"logits": out["logits"].detach()}
