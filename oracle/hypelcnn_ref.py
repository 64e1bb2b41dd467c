"""CPU ORACLE (test infrastructure, NOT product code) — HYPELCNN forward / loss / backward / Adam.

A restatement, on torch-CPU tensors, of what the reference's TF1 graph computes for the
hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.

PARITY STATUS: **unpinned against TensorFlow numerics** — TensorFlow / tf_slim cannot be
installed in the build container and the reference ships no tests.  What IS pinned, by
executing the reference's own Python (``tests/golden/make_golden.py``): the layer
sequence, scopes, channel counts, kernel sizes, FC stage sizes, dropout keep_prob, and every
``scale_in_to_out`` index table.  The arithmetic inside each TF op follows the published
semantics of tensorflow 2.9 / tf-slim 1.1.0, cited per function.

Reference files followed (relative to the reference repo root):
  nnmodel/HYPELCNNModel.py:34-183      graph
  common/common_nn_ops.py:208-240      optimize_nn (loss mean, LR decay, Adam)
  common/common_nn_ops.py:546-564      scale_in_to_out
"""
import math

import numpy
import torch
import torch.nn.functional as F

BN_EPS = 0.001  # slim batch_norm default epsilon [TF-lib]
TRUNC_STD_FIX = 0.87962566103423978  # variance_scaling truncated-normal correction [TF-lib]


# --------------------------------------------------------------------------- #
# pure-Python pieces (pinned exactly by tests/golden/*.json)
# --------------------------------------------------------------------------- #
def scale_in_to_out_index(cin, cout):
    """Index table of scale_in_to_out (common/common_nn_ops.py:546-564).

    out[..., j] = in[..., idx[j]].  Integer inverse ratio -> tf.repeat (each channel
    repeated consecutively); otherwise gather with Python banker's ``round``.
    """
    scale_ratio = cin / cout
    inv_scale_ratio = 1 / scale_ratio
    if float(inv_scale_ratio).is_integer():
        rep = int(inv_scale_ratio)
        return [j // rep for j in range(cin * rep)]
    return [min(round(j * scale_ratio), cin - 1) for j in range(cout)]


def fc_stage_sizes(flat, classes, deg):
    """FC block sizing (nnmodel/HYPELCNNModel.py:115-125)."""
    stages = math.floor(math.log(flat / classes, deg))
    sizes = []
    size = flat
    for _ in range(0, stages - 1):
        size = size // deg
        sizes.append(size)
    return sizes


def level_kernel_sizes(patch):
    """Square odd kernels 1,3,..,<=patch (nnmodel/HYPELCNNModel.py:171-176)."""
    return [k for k in range(1, patch + 1) if k % 2 == 1]


class LayerSpec:
    """One conv2d / fully_connected of the graph with everything attached to it."""

    def __init__(self, scope, kind, cin, cout, kernel, act, src, dst, residuals=(), dropout=False,
                 concat_slot=None):
        self.scope, self.kind, self.cin, self.cout, self.kernel = scope, kind, cin, cout, kernel
        self.act, self.src, self.dst = act, src, dst
        self.residuals = list(residuals)  # [(source tensor name, source channel count)] added after activation
        self.dropout = dropout
        self.concat_slot = concat_slot  # (tensor name, channel offset) for multi-kernel levels


def build_plan(patch, channels, classes, alg, is_training):
    """Layer list in TF variable-creation order (nnmodel/HYPELCNNModel.py:34-99).

    Each entry is a LayerSpec; multi-kernel levels are one entry per conv whose outputs
    concatenate along channels (``:167-183``), the level-wide residuals hang on the
    pseudo-entry kind='level_end'.
    """
    F_ = alg["filter_count"]
    L = alg["spectral_hierarchy_level"]
    S = alg["spatial_hierarchy_level"]
    res = bool(alg["use_residual"])
    plan = []
    cur, cur_c = "x", channels
    # spectral encoder / decoder (:146-164)
    for enc in (True, False):
        block_in, block_in_c = cur, cur_c
        for i in range(L):
            cout = F_ // pow(2, (L - 1) - i) if enc else F_ // pow(2, i)
            name = ("conv_enc_" if enc else "conv_dec_") + str(i)
            r = [(cur, cur_c)] if res else []
            last = i == L - 1
            if last and res:
                r.append((block_in, block_in_c))  # net1 += S(net0) / net2 += S(net1)   (:57-58, :63-64)
            plan.append(LayerSpec(name, "conv", cur_c, cout, 1, "lrelu", cur, name, r))
            cur, cur_c = name, cout
    net2, net2_c = cur, cur_c
    # spatial blocks (:128-143)
    lf = net2_c // 2
    for i in range(S):
        f = lf // pow(2, i)
        ks = level_kernel_sizes(patch)
        lvl = f"connector_{i}"
        for j, k in enumerate(ks):
            plan.append(LayerSpec(f"{lvl}_conv{k}x{k}", "conv", cur_c, f, k, "lrelu", cur, lvl,
                                  concat_slot=(lvl, j * f)))
        lvl_c = f * len(ks)
        plan.append(LayerSpec(lvl, "level_end", cur_c, lvl_c, 0, None, cur, lvl,
                              [(cur, cur_c)] if res else []))
        cname = f"connector_conv_{i}"
        r = [(lvl, lvl_c)] if res else []
        if i == S - 1 and res:
            r.append((net2, net2_c))  # net3 += S(net2)   (:71-72)
        plan.append(LayerSpec(cname, "conv", lvl_c, lvl_c, 1, "lrelu", lvl, cname, r))
        cur, cur_c = cname, lvl_c
    flat = patch * patch * cur_c
    plan.append(LayerSpec("flatten", "flatten", cur_c, flat, 0, None, cur, "flat"))
    cur, cur_c = "flat", flat
    for i, size in enumerate(fc_stage_sizes(flat, classes, alg["degradation_coeff"])):
        plan.append(LayerSpec(f"fc_{i}", "fc", cur_c, size, 0, "lrelu", cur, f"fc_{i}", dropout=True))
        cur, cur_c = f"fc_{i}", size
    plan.append(LayerSpec("fc_final", "fc", cur_c, classes, 0, None, cur, "fc_final"))
    cur, cur_c = "fc_final", classes
    if is_training:  # decoder only in the training graph (:84-94)
        for i, mult in enumerate((3, 9, 27)):
            plan.append(LayerSpec(f"image_gen_net_{i + 1}", "fc", cur_c, classes * mult, 0, "lrelu", cur,
                                  f"image_gen_net_{i + 1}"))
            cur, cur_c = f"image_gen_net_{i + 1}", classes * mult
        plan.append(LayerSpec("image_gen_net_4", "fc", cur_c, patch * patch * channels, 0, "sigmoid", cur,
                              "image_gen_net_4"))
    return plan


def variable_specs(patch, channels, classes, alg):
    """[(tf variable name, shape, kind)] for the TRAINING graph, creation order.

    Names are the checkpoint names of the reference: ``nn_core/<scope>/weights``
    ``[kh,kw,Cin,Cout]`` or ``[Cin,Cout]``; ``nn_core/<scope>/BatchNorm/{beta,moving_mean,
    moving_variance}``.  No biases (normalizer_fn set), no gamma (scale=False).  [TF-lib]
    """
    out = []
    for l in build_plan(patch, channels, classes, alg, True):
        if l.kind not in ("conv", "fc"):
            continue
        shape = (l.kernel, l.kernel, l.cin, l.cout) if l.kind == "conv" else (l.cin, l.cout)
        out.append((f"nn_core/{l.scope}/weights", shape, "weights"))
        out.append((f"nn_core/{l.scope}/BatchNorm/beta", (l.cout,), "beta"))
        out.append((f"nn_core/{l.scope}/BatchNorm/moving_mean", (l.cout,), "moving_mean"))
        out.append((f"nn_core/{l.scope}/BatchNorm/moving_variance", (l.cout,), "moving_variance"))
    return out


def init_variables(patch, channels, classes, alg, seed=1234, dtype=torch.float32):
    """variance_scaling(scale=2.0): fan_in, truncated normal at +-2 sigma,
    std = sqrt(2/fan_in)/0.87962566 (nnmodel/HYPELCNNModel.py:41) [TF-lib]; beta=0,
    moving_mean=0, moving_variance=1.  The RNG stream is numpy's, not TF's."""
    rng = numpy.random.default_rng(seed)
    v = {}
    for name, shape, kind in variable_specs(patch, channels, classes, alg):
        if kind == "weights":
            fan_in = int(numpy.prod(shape[:-1]))
            std = math.sqrt(2.0 / fan_in) / TRUNC_STD_FIX
            w = rng.standard_normal(shape)
            bad = numpy.abs(w) > 2.0
            while bad.any():
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = numpy.abs(w) > 2.0
            v[name] = torch.tensor(w * std, dtype=dtype)
        elif kind == "moving_variance":
            v[name] = torch.ones(shape, dtype=dtype)
        else:
            v[name] = torch.zeros(shape, dtype=dtype)
    return v


# --------------------------------------------------------------------------- #
# TF op restatements
# --------------------------------------------------------------------------- #
def conv2d_same_nhwc(x, w):
    """slim conv2d: stride 1, SAME, NHWC, weights [kh,kw,Cin,Cout], no bias (App. A.1)."""
    k = w.shape[0]
    y = F.conv2d(x.permute(0, 3, 1, 2).contiguous(), w.permute(3, 2, 0, 1).contiguous(), padding=k // 2)
    return y.permute(0, 2, 3, 1)


def conv2d_same_nhwc_naive(x, w):
    """Literal loop form of the same op — used by the tests to pin the layout conventions of
    conv2d_same_nhwc on small cases."""
    B, H, W, C = x.shape
    k = w.shape[0]
    p = k // 2
    out = torch.zeros(B, H, W, w.shape[3], dtype=x.dtype)
    for dy in range(k):
        for dx in range(k):
            for h in range(H):
                for ww in range(W):
                    sh, sw = h + dy - p, ww + dx - p
                    if 0 <= sh < H and 0 <= sw < W:
                        out[:, h, ww, :] += x[:, sh, sw, :] @ w[dy, dx]
    return out


def batch_norm(z, beta, moving_mean, moving_var, is_training, decay):
    """slim batch_norm, center=True scale=False eps=1e-3, fused kernel (App. A.3) [TF-lib].

    Training: normalise with biased batch variance; moving stats updated with the
    Bessel-corrected variance, ``moving = moving*decay + batch*(1-decay)``.
    Returns (y, new_moving_mean, new_moving_var, saved) where saved=(mean, rstd).
    """
    red = tuple(range(z.dim() - 1))
    if is_training:
        n = z.numel() // z.shape[-1]
        mean = z.mean(dim=red)
        var = ((z - mean) ** 2).mean(dim=red)
        rstd = torch.rsqrt(var + BN_EPS)
        y = (z - mean) * rstd + beta
        unbiased = var * (n / max(n - 1, 1))
        with torch.no_grad():
            new_mm = moving_mean * decay + mean * (1 - decay)
            new_mv = moving_var * decay + unbiased * (1 - decay)
        return y, new_mm.detach(), new_mv.detach(), (mean.detach(), rstd.detach())
    rstd = torch.rsqrt(moving_var + BN_EPS)
    return (z - moving_mean) * rstd + beta, moving_mean, moving_var, (moving_mean, rstd)


def leaky_relu(x, alpha):
    """tf.nn.leaky_relu = max(x, alpha*x) (App. A.4)."""
    return torch.maximum(x, alpha * x)


def resample(src, cout):
    """scale_in_to_out applied on the last axis."""
    idx = scale_in_to_out_index(src.shape[-1], cout)
    assert len(idx) == cout, (src.shape[-1], cout, len(idx))
    if idx == list(range(cout)):
        return src
    return src[..., torch.tensor(idx, dtype=torch.long)]


def forward(variables, x, classes, alg, is_training, dropout_masks=None, update_moving=True, lrelu_gates=None):
    """HYPELCNN forward.  x [B,P,P,C].  Returns dict: logits, recon (or None), tensors{name:
    activation}, pre{scope: pre-BN conv output}, new_variables (moving stats updated when
    training), saved BN stats.

    dropout_masks: {scope: 0/1 tensor [B,size]} — kept elements; None -> no dropout applied
    when keep_prob==1, else error in training (TF's RNG stream cannot be reproduced).
    lrelu_gates: optional {scope: bool tensor shaped like the layer output} — which side of
    the LeakyReLU kink each element is on.  max(y, alpha*y) is not differentiable at y == 0
    and, among millions of activations, a few land within fp32 rounding of 0; a gradient
    comparison is only meaningful when both sides take the same branch there, so the parity
    tests pass the branch the implementation under test took (the forward value changes by
    at most (1-alpha)*|y| ~ 1e-7 for those elements).
    """
    P, C = x.shape[1], x.shape[3]
    plan = build_plan(P, C, classes, alg, is_training)
    alpha, decay = alg["lrelu_alpha"], alg["bn_decay"]
    keep = 1 - alg["drop_out_ratio"]
    T = {"x": x}
    pre, saved = {}, {}
    newv = dict(variables)
    level_parts = {}
    for l in plan:
        if l.kind == "flatten":
            T[l.dst] = T[l.src].reshape(T[l.src].shape[0], -1)  # NHWC order (App. A.9)
            continue
        if l.kind == "level_end":
            t = torch.cat(level_parts.pop(l.dst), dim=-1)  # tf.concat(axis=3) (:182)
            for s, _ in l.residuals:
                t = t + resample(T[s], l.cout)
            T[l.dst] = t
            continue
        w = variables[f"nn_core/{l.scope}/weights"]
        z = conv2d_same_nhwc(T[l.src], w) if l.kind == "conv" else T[l.src] @ w
        pre[l.scope] = z
        bn = f"nn_core/{l.scope}/BatchNorm/"
        y, mm, mv, sv = batch_norm(z, variables[bn + "beta"], variables[bn + "moving_mean"],
                                   variables[bn + "moving_variance"], is_training, decay)
        saved[l.scope] = sv
        if is_training and update_moving:
            newv[bn + "moving_mean"], newv[bn + "moving_variance"] = mm, mv
        if l.act == "lrelu":
            if lrelu_gates is not None and l.scope in lrelu_gates:
                y = torch.where(lrelu_gates[l.scope], y, alpha * y)
            else:
                y = leaky_relu(y, alpha)
        elif l.act == "sigmoid":
            y = torch.sigmoid(y)
        if l.dropout and is_training and keep < 1.0:  # slim dropout: x*mask/keep_prob (App. A.5)
            if dropout_masks is None:
                raise ValueError("training with keep_prob<1 needs injected dropout masks")
            y = y * dropout_masks[l.scope].to(y.dtype) / keep
        if l.concat_slot is not None:
            level_parts.setdefault(l.concat_slot[0], []).append(y)
            continue
        for s, _ in l.residuals:
            y = y + resample(T[s], l.cout)
        T[l.dst] = y
    return {"logits": T["fc_final"], "recon": T.get("image_gen_net_4"), "tensors": T, "pre": pre,
            "new_variables": newv, "saved": saved}


def per_sample_loss(logits, recon, x, labels):
    """HYPELCNNModel.get_loss_func (nnmodel/HYPELCNNModel.py:101-112): softmax CE per sample
    (+ scalar reconstruction MSE broadcast, training graph)."""
    ce = -(F.log_softmax(logits, dim=1)[torch.arange(logits.shape[0]), labels.long()])
    if recon is None:
        return ce
    mse = ((recon - x.reshape(x.shape[0], -1)) ** 2).mean()
    return ce + mse


def loss_and_grads(variables, x, labels, classes, alg, dropout_masks=None, lrelu_gates=None):
    """optimize_nn's loss = mean_B(per-sample loss) (common/common_nn_ops.py:214) and its
    gradients w.r.t. every trainable variable (autograd on the restatement).  L2 regulariser
    terms are NOT part of the optimised loss (App. A.7)."""
    leaf = {k: (v.clone().requires_grad_(True) if ("weights" in k or k.endswith("beta")) else v)
            for k, v in variables.items()}
    out = forward(leaf, x, classes, alg, True, dropout_masks, lrelu_gates=lrelu_gates)
    loss = per_sample_loss(out["logits"], out["recon"], x, labels).mean()
    names = [k for k, v in leaf.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)
    g = {k: (gi if gi is not None else torch.zeros_like(leaf[k])) for k, gi in zip(names, grads)}
    return loss.detach(), g, out


def learning_rate(alg, global_step):
    """exponential_decay(staircase=True) (common/common_nn_ops.py:217-221; App. A.14)."""
    return alg["learning_rate"] * alg["learning_rate_decay_factor"] ** (global_step // alg["learning_rate_decay_step"])


def adam_tf1(param, grad, m, v, lr, t, b1=0.9, b2=0.999, eps=1e-8):
    """TF1 AdamOptimizer (App. A.6): epsilon OUTSIDE the bias correction.  t starts at 1."""
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    m = m + (grad - m) * (1 - b1)
    v = v + (grad * grad - v) * (1 - b2)
    return param - lr_t * m / (torch.sqrt(v) + eps), m, v


def train_step(variables, opt_state, x, labels, classes, alg, global_step, dropout_masks=None):
    """One optimize_nn train step: grads at the current variables, Adam, BN moving stats."""
    loss, g, out = loss_and_grads(variables, x, labels, classes, alg, dropout_masks)
    lr = learning_rate(alg, global_step)
    newv = dict(out["new_variables"])
    for k, gk in g.items():
        m, v = opt_state.get(k, (torch.zeros_like(gk), torch.zeros_like(gk)))
        newv[k], m, v = adam_tf1(variables[k], gk, m, v, lr, global_step + 1)
        opt_state[k] = (m, v)
    return loss, newv, opt_state, g


def useful_flops_per_patch(patch, channels, classes, alg, is_training=True):
    """Forward MAC*2 per patch excluding multiply-by-zero SAME-padding taps."""
    total = 0
    for l in build_plan(patch, channels, classes, alg, is_training):
        if l.kind == "conv":
            k = l.kernel
            valid = sum(1 for h in range(patch) for w in range(patch) for dy in range(-(k // 2), k // 2 + 1)
                        for dx in range(-(k // 2), k // 2 + 1) if 0 <= h + dy < patch and 0 <= w + dx < patch)
            total += 2 * valid * l.cin * l.cout
        elif l.kind == "fc":
            total += 2 * l.cin * l.cout
    return total
