"""PatchEngine — owns the flat device buffers (parameters, gradients, BN state, Adam moments,
workspace) as torch tensors and drives a native ``hyp_model`` through the C ABI.

PyTorch is used here for device memory, streams and (in parallel.py) the NCCL process
group only; every arithmetic step of the hot path is a kernel of libhypelcnn_b200.so.
"""
import ctypes
import math

import numpy
import torch

from hypelcnn_b200 import _native as N

TRUNC_STD_FIX = 0.87962566103423978  # variance_scaling truncated-normal correction [TF-lib]
# tcgen05 tensor-core engine: "3xf16" (fp16 hi/lo operand planes, fp32-accurate, the default for HYPELCNN), "3xtf32"
# (TF32 planes, fp32-accurate, half the rate), "bf16" (one bf16 plane: the labelled fast mode); "fp32": FFMA engine
_PRECISIONS = {"fp32": N.HYP_PRECISION_FP32, "3xtf32": N.HYP_PRECISION_3XTF32, "bf16": N.HYP_PRECISION_BF16,
               "3xf16": N.HYP_PRECISION_3XF16}


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA torch.Tensor (this engine has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


class PatchEngine:
    """One HYPELCNN instance on one GPU.  ``algorithm_params`` is the reference's JSON dict
    (nnmodel/modelconfigs/alg_param_hypelcnn.json) — a missing key raises KeyError exactly
    like the reference's dict lookups."""

    def __init__(self, patch, channels, classes, algorithm_params, max_batch, device=None, precision=None,
                 model="hypelcnn"):
        if not torch.cuda.is_available():
            raise N.NativeError(N.HYP_E_CUDA, "no CUDA device: hypelcnn_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.patch, self.channels, self.classes = int(patch), int(channels), int(classes)
        self.alg = dict(algorithm_params)
        # default operand format: fp16 hi/lo planes for HYPELCNN (BatchNorm after every layer keeps activations in the
        # fp16 range); TF32 planes for DUALCNN / CONCNN, whose un-normalised activations have no such bound
        self.precision = precision if precision is not None else ("3xf16" if model == "hypelcnn" else "3xtf32")
        if model not in ("hypelcnn", "dualcnn", "concnn"):
            raise ValueError(f"unknown model kind {model!r}")
        self.model = model
        self._handle = ctypes.c_void_p()
        self.global_step = 0
        self._max_batch = 0
        self.params = self.grads = self.state = self.adam_m = self.adam_v = self.workspace = None
        self._notify_events = self._comm_stream = self._notify_handle = None
        self._create(int(max_batch))

    # ------------------------------------------------------------------ lifecycle
    def _desc(self, max_batch):
        a = self.alg
        if self.model == "concnn":  # nnmodel/modelconfigs/alg_param_concnn.json keys; slim's default activation is ReLU
            return N.ModelDesc(kind=N.HYP_MODEL_CONCNN, patch=self.patch, channels=self.channels, classes=self.classes,
                               filter_count=int(a["filter_count"]), spectral_levels=0, spatial_levels=0, degradation=0,
                               use_residual=1, precision_mode=_PRECISIONS[self.precision], max_batch=max_batch,
                               reserved=0, lrelu_alpha=0.0, bn_decay=0.0, bn_eps=0.0,
                               drop_out_ratio=float(a["drop_out_ratio"]))
        if self.model == "dualcnn":  # nnmodel/modelconfigs/alg_param_dualcnn.json keys
            return N.ModelDesc(kind=N.HYP_MODEL_DUALCNN, patch=self.patch, channels=self.channels, classes=self.classes,
                               filter_count=int(a["filter_count"]), spectral_levels=0, spatial_levels=0, degradation=0,
                               use_residual=0, precision_mode=_PRECISIONS[self.precision], max_batch=max_batch,
                               reserved=int(a["hs_lidar_diff"]), lrelu_alpha=float(a["lrelu_alpha"]), bn_decay=0.0,
                               bn_eps=0.0, drop_out_ratio=float(a["drop_out_ratio"]))
        return N.ModelDesc(kind=0, patch=self.patch, channels=self.channels, classes=self.classes,
                           filter_count=int(a["filter_count"]), spectral_levels=int(a["spectral_hierarchy_level"]),
                           spatial_levels=int(a["spatial_hierarchy_level"]), degradation=int(a["degradation_coeff"]),
                           use_residual=int(bool(a["use_residual"])), precision_mode=_PRECISIONS[self.precision],
                           max_batch=max_batch, reserved=0, lrelu_alpha=float(a["lrelu_alpha"]),
                           bn_decay=float(a["bn_decay"]), bn_eps=0.001, drop_out_ratio=float(a["drop_out_ratio"]))

    def _create(self, max_batch):
        L = N.lib()
        with torch.cuda.device(self.device):
            handle = ctypes.c_void_p()
            desc = self._desc(max_batch)
            N.check(L.hyp_model_create(ctypes.byref(desc), ctypes.byref(handle)))
            n_params, n_state, ws, nvar = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
            N.check(L.hyp_model_sizes(handle, ctypes.byref(n_params), ctypes.byref(n_state), ctypes.byref(ws),
                                      ctypes.byref(nvar)))
            if self.params is None:
                z = dict(dtype=torch.float32, device=self.device)
                self.params = torch.zeros(n_params.value, **z)
                self.grads = torch.zeros(n_params.value, **z)
                self.adam_m = torch.zeros(n_params.value, **z)
                self.adam_v = torch.zeros(n_params.value, **z)
                self.state = torch.zeros(n_state.value, **z)
                self.variables = {}
                name = ctypes.create_string_buffer(128)
                kind, off, rank = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int32()
                shape = (ctypes.c_int32 * 4)()
                for i in range(nvar.value):
                    N.check(L.hyp_model_variable(handle, i, name, ctypes.byref(kind), ctypes.byref(off), shape,
                                                 ctypes.byref(rank)))
                    self.variables[name.value.decode()] = (kind.value, off.value, tuple(shape[:rank.value]))
                for nm, (k, o, s) in self.variables.items():
                    if k == 3:
                        self.variable(nm).fill_(1.0)  # moving_variance initial value
            self.workspace = torch.empty(ws.value + 256, dtype=torch.uint8, device=self.device)
            base = self.workspace.data_ptr()
            self._ws_ptr = (base + 255) // 256 * 256
            N.check(L.hyp_model_bind(handle, _ptr(self.params), _ptr(self.grads), _ptr(self.state),
                                     ctypes.c_void_p(self._ws_ptr), ctypes.c_size_t(ws.value)))
            if self._handle:
                L.hyp_model_destroy(self._handle)
            self._handle, self._max_batch, self.workspace_bytes = handle, max_batch, ws.value

    def ensure_capacity(self, batch):
        """Grow the workspace (parameters and optimizer state are kept)."""
        if batch > self._max_batch:
            torch.cuda.synchronize(self.device)
            self._create(int(batch))

    def close(self):
        if self._handle:
            N.lib().hyp_model_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ variables
    @property
    def trainable_count(self):
        return sum(int(numpy.prod(s)) for k, o, s in self.variables.values() if k in (0, 1))

    def variable(self, name):
        """View of one TF-named variable inside the flat buffers."""
        kind, off, shape = self.variables[name]
        buf = self.params if kind in (0, 1) else self.state
        n = int(numpy.prod(shape))
        return buf[off:off + n].view(shape)

    def gradient(self, name):
        kind, off, shape = self.variables[name]
        if kind not in (0, 1):
            raise KeyError(f"{name} is not trainable")
        return self.grads[off:off + int(numpy.prod(shape))].view(shape)

    def init_variables(self, seed=1234):
        """variance_scaling(scale=2.0) truncated normal (nnmodel/HYPELCNNModel.py:41), beta=0,
        moving_mean=0, moving_variance=1."""
        rng = numpy.random.default_rng(seed)
        self.params.zero_()
        self.state.zero_()
        self.adam_m.zero_()
        self.adam_v.zero_()
        self.global_step = 0
        if self.model in ("dualcnn", "concnn"):
            # slim defaults (nnmodel/DUALCNNModel.py:13-18 sets only the activation): xavier-uniform weights
            # (limit = sqrt(6 / (fan_in + fan_out)), receptive field included), zero biases [TF-lib]
            for name, (kind, off, shape) in self.variables.items():
                if kind == 0:
                    rf = int(numpy.prod(shape[:-2]))
                    limit = math.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))
                    self.variable(name).copy_(torch.from_numpy(rng.uniform(-limit, limit, shape).astype(numpy.float32)))
            return
        for name, (kind, off, shape) in self.variables.items():
            if kind == 0:
                fan_in = int(numpy.prod(shape[:-1]))
                std = math.sqrt(2.0 / fan_in) / TRUNC_STD_FIX
                w = rng.standard_normal(shape)
                bad = numpy.abs(w) > 2.0
                while bad.any():
                    w[bad] = rng.standard_normal(int(bad.sum()))
                    bad = numpy.abs(w) > 2.0
                self.variable(name).copy_(torch.from_numpy((w * std).astype(numpy.float32)))
            elif kind == 3:
                self.variable(name).fill_(1.0)

    def load_variables(self, values):
        """values: {tf variable name: array-like}.  Unknown names raise KeyError."""
        for name, v in values.items():
            t = torch.as_tensor(numpy.asarray(v), dtype=torch.float32)
            dst = self.variable(name)
            if tuple(t.shape) != tuple(dst.shape):
                raise ValueError(f"{name}: shape {tuple(t.shape)} != {tuple(dst.shape)}")
            dst.copy_(t)

    # ------------------------------------------------------------------ checkpoints
    def save_checkpoint(self, path):
        """Variables under their TF checkpoint names plus the optimizer slots and global_step, as safetensors — the
        on-disk role of the reference's Saver(nn_core/*, global_step, training_optimizer/*)
        (classify/monitored_session_runner.py:164-168).  TF's own ckpt format needs TensorFlow and is out of scope."""
        from safetensors.torch import save_file
        t = {}
        for name, (kind, off, shape) in self.variables.items():
            n = int(numpy.prod(shape))
            t[name] = self.variable(name).detach().cpu().contiguous()
            if kind in (0, 1):  # Adam slots are named <variable>/Adam and /Adam_1 in TF checkpoints [TF-lib]
                t[name + "/Adam"] = self.adam_m[off:off + n].view(shape).detach().cpu().contiguous()
                t[name + "/Adam_1"] = self.adam_v[off:off + n].view(shape).detach().cpu().contiguous()
        t["global_step"] = torch.tensor([self.global_step], dtype=torch.int64)
        save_file(t, path, metadata={"model": self.model, "patch": str(self.patch), "channels": str(self.channels),
                                     "classes": str(self.classes)})

    def load_checkpoint(self, path, exclude_prefixes=()):
        """Restore what save_checkpoint wrote.  exclude_prefixes mirrors the inference restore, which skips
        ``image_gen_net_*`` (classify/infer_for_classification.py:121-128)."""
        from safetensors.torch import load_file
        t = load_file(path)
        for name, (kind, off, shape) in self.variables.items():
            if any(name.startswith("nn_core/" + p) for p in exclude_prefixes):
                continue
            n = int(numpy.prod(shape))
            self.variable(name).copy_(t[name])
            if kind in (0, 1) and name + "/Adam" in t:
                self.adam_m[off:off + n].view(shape).copy_(t[name + "/Adam"])
                self.adam_v[off:off + n].view(shape).copy_(t[name + "/Adam_1"])
        self.global_step = int(t["global_step"][0])

    def export_variables(self):
        return {name: self.variable(name).detach().cpu().numpy().copy() for name in self.variables}

    # ------------------------------------------------------------------ compute
    def _check_x(self, x):
        require_cuda(x, "x", torch.float32)
        if x.dim() != 4 or tuple(x.shape[1:]) != (self.patch, self.patch, self.channels):
            raise ValueError(f"x must be [B,{self.patch},{self.patch},{self.channels}], got {tuple(x.shape)}")
        self.ensure_capacity(x.shape[0])
        return x.shape[0]

    def forward(self, x, is_training, update_moving=True, seed=0):
        """-> (logits [B,classes], recon [B,P*P*C] or None).  hyp_model_forward."""
        B = self._check_x(x)
        logits = torch.empty((B, self.classes), dtype=torch.float32, device=x.device)
        recon = torch.empty((B, self.patch * self.patch * self.channels), dtype=torch.float32,
                            device=x.device) if (is_training and self.model == "hypelcnn") else None
        N.check(N.lib().hyp_model_forward(self._handle, _ptr(x), B, int(bool(is_training)), int(bool(update_moving)),
                                          ctypes.c_uint64(seed), _ptr(logits), _ptr(recon), _stream()))
        self._last_x = x  # keep alive until backward
        return logits, recon

    def forward_inplace(self, x, is_training, update_moving=True, seed=0):
        """Forward without copying logits/recon out of the workspace (used by train_step)."""
        B = self._check_x(x)
        N.check(N.lib().hyp_model_forward(self._handle, _ptr(x), B, int(bool(is_training)), int(bool(update_moving)),
                                          ctypes.c_uint64(seed), None, None, _stream()))
        self._last_x = x

    def per_sample_loss(self, logits, recon, x, labels):
        require_cuda(logits, "logits", torch.float32)
        require_cuda(labels, "labels", torch.uint8)
        B = logits.shape[0]
        out = torch.empty(B, dtype=torch.float32, device=logits.device)
        N.check(N.lib().hyp_model_loss(self._handle, _ptr(logits), _ptr(recon), _ptr(x), _ptr(labels), B, _ptr(out),
                                       _stream()))
        return out

    def loss_backward(self, x, labels):
        """-> device tensor [3] = (loss, mean CE, reconstruction MSE); fills self.grads."""
        require_cuda(labels, "labels", torch.uint8)
        out = torch.empty(3, dtype=torch.float32, device=x.device)
        N.check(N.lib().hyp_model_loss_backward(self._handle, _ptr(x), _ptr(labels), x.shape[0], _ptr(out),
                                                _stream()))
        return out

    def learning_rate(self, global_step=None):
        """exponential_decay(staircase=True) (common/common_nn_ops.py:217-221)."""
        s = self.global_step if global_step is None else global_step
        a = self.alg
        return a["learning_rate"] * a["learning_rate_decay_factor"] ** (s // a["learning_rate_decay_step"])

    def adam_step(self, lr=None, grad_scale=1.0, b1=0.9, b2=0.999, eps=1e-8):
        """The optimizer step of optimize_nn (common/common_nn_ops.py:223-232): "AdamOptimizer" or
        ["MomentumOptimizer", momentum] as the JSON's "optimizer" value."""
        opt = self.alg.get("optimizer", "AdamOptimizer")
        lr = self.learning_rate() if lr is None else lr
        if isinstance(opt, (tuple, list)):
            if opt[0] != "MomentumOptimizer":
                raise N.NativeError(N.HYP_E_UNSUPPORTED, f"optimizer {opt[0]!r} is not one the reference selects")
            N.check(N.lib().hyp_momentum_step(_ptr(self.params), _ptr(self.grads), _ptr(self.adam_m),
                                              self.params.numel(), lr, float(opt[1]), grad_scale, _stream()))
            self.global_step += 1
            return
        if opt != "AdamOptimizer":
            raise N.NativeError(N.HYP_E_UNSUPPORTED, f"optimizer {opt!r} is not one the reference selects")
        N.check(N.lib().hyp_adam_step(_ptr(self.params), _ptr(self.grads), _ptr(self.adam_m), _ptr(self.adam_v),
                                      self.params.numel(), lr, b1, b2, eps, self.global_step + 1, grad_scale,
                                      _stream()))
        self.global_step += 1

    def train_step(self, x, labels, seed=None, allreduce=None):
        """One optimize_nn step (common/common_nn_ops.py:208-240): forward (training), loss,
        backward, [gradient all-reduce], Adam.  Returns the device loss tensor [3]."""
        seed = self.global_step if seed is None else seed
        self.forward_inplace(x, True, True, seed)
        overlap = allreduce is not None and getattr(allreduce, "overlap", False)
        splits = self.grad_split_offsets if overlap else []
        if overlap and splits:
            # Parameters are laid out in layer order and backward walks the layers last to first: the piece of the flat
            # gradient buffer behind each split point is final long before backward ends.  Its all-reduce runs on the
            # communication stream meanwhile; only the head (the six spectral 1x1 convolutions, 2 MB) is reduced after
            # the last layer's backward.
            if self._notify_events is None:
                self._notify_events = [torch.cuda.Event() for _ in splits]
                for ev in self._notify_events:
                    ev.record()  # materialises the cudaEvent_t
                self._comm_stream = torch.cuda.Stream(device=self.device)
                self._notify_handle = None
            if self._notify_handle != self._handle.value:   # a re-created native model (capacity growth) forgets them
                for off, ev in zip(splits, self._notify_events):
                    N.check(N.lib().hyp_model_set_grad_notify(self._handle, off, ctypes.c_void_p(ev.cuda_event)))
                self._notify_handle = self._handle.value
        loss = self.loss_backward(x, labels)
        scale = 1.0
        if overlap and splits:
            main = torch.cuda.current_stream()
            end = self.grads.numel()
            with torch.cuda.stream(self._comm_stream):
                for off, ev in zip(splits, self._notify_events):      # tail first: that is the order they become final
                    self._comm_stream.wait_event(ev)
                    allreduce(self.grads[off:end])
                    end = off
            scale = allreduce(self.grads[:end])
            main.wait_stream(self._comm_stream)
        elif allreduce is not None:
            scale = allreduce(self.grads)
        self.adam_step(grad_scale=scale)
        return loss

    @property
    def grad_split_offsets(self):
        """Split points of the overlapped gradient all-reduce, largest first: the first parameter of the FC block (the
        FC / decoder tail, ~70 % of the buffer, is final once fc_0's backward has run) and the first parameter of the
        spatial levels (final once connector_0's backward has run)."""
        offs = [self.grad_split_offset]
        # (a level is ONE layer of the engine: its split point is the level's first variable, the 1x1 kernel)
        level = [off for name, (kind, off, shape) in self.variables.items() if kind == 0 and "/connector_0_conv" in name]
        if level and 0 < min(level) < offs[0]:
            offs.append(min(level))
        return [o for o in offs if o > 0]

    @property
    def grad_split_offset(self):
        """First parameter of the FC block: gradients from here to the end of the flat buffer are complete once
        fc_0's backward has run (hyp_model_set_grad_notify)."""
        for name, (kind, off, shape) in self.variables.items():
            if kind == 0 and len(shape) == 2:
                return off
        return 0

    def debug_tensor(self, name, what=0):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        N.check(N.lib().hyp_model_debug_tensor(self._handle, name.encode(), what, ctypes.byref(p), ctypes.byref(n)))
        off = p.value - self._ws_ptr
        base = self._ws_ptr - self.workspace.data_ptr()
        return self.workspace[base + off: base + off + 4 * n.value].view(torch.float32).clone()

    def dropout_mask(self, layer_scope, seed, batch, rows_per_sample=1):
        """0/1 keep mask of a dropout layer for `seed`.  FC layers: [batch, width]; conv layers (rows_per_sample =
        P*P): [P*P, batch, width] in the engine's position-major row order."""
        width = None
        for nm, (k, o, s) in self.variables.items():
            if nm in (f"nn_core/{layer_scope}/BatchNorm/beta", f"nn_core/{layer_scope}/biases"):
                width = s[0]
        if width is None:
            raise KeyError(layer_scope)
        out = torch.empty((batch, width) if rows_per_sample == 1 else (rows_per_sample, batch, width), dtype=torch.uint8,
                          device=self.device)
        N.check(N.lib().hyp_model_dropout_mask(self._handle, layer_scope.encode(), ctypes.c_uint64(seed), batch,
                                               _ptr(out), _stream()))
        return out


# ---------------------------------------------------------------------- free functions
def adam_step(params, grads, m, v, lr, t, grad_scale=1.0, b1=0.9, b2=0.999, eps=1e-8):
    for name, t_ in (("params", params), ("grads", grads), ("m", m), ("v", v)):
        require_cuda(t_, name, torch.float32)
    N.check(N.lib().hyp_adam_step(_ptr(params), _ptr(grads), _ptr(m), _ptr(v), params.numel(), lr, b1, b2, eps, t,
                                  grad_scale, _stream()))


def augment_patches(x, rotation=False, reflection=False, spectral=0.0, seed=0, return_draw=False):
    """Device-side form of the reference's augmentation maps (common/common_nn_ops.py:397-440); one random draw per
    sample.  -> augmented [B,P,P,C] (and (choices uint8 [B,4], deltas [B,C]) with return_draw)."""
    require_cuda(x, "x", torch.float32)
    B, P, _, C = x.shape
    out = torch.empty_like(x)
    choices = torch.zeros((B, 4), dtype=torch.uint8, device=x.device) if return_draw else None
    deltas = torch.zeros((B, C), dtype=torch.float32, device=x.device) if return_draw else None
    N.check(N.lib().hyp_augment_patches(_ptr(x), _ptr(out), B, P, C, int(bool(rotation)), int(bool(reflection)),
                                        float(spectral), ctypes.c_uint64(seed), _ptr(choices), _ptr(deltas), _stream()))
    return (out, choices, deltas) if return_draw else out


def argmax_confusion(logits, labels=None, confusion=None):
    """tf.argmax + confusion accumulation (common/common_nn_ops.py:246-262).  -> uint8 [B]."""
    require_cuda(logits, "logits", torch.float32)
    B, classes = logits.shape
    pred = torch.empty(B, dtype=torch.uint8, device=logits.device)
    if labels is not None:
        require_cuda(labels, "labels", torch.uint8)
    if confusion is not None:
        require_cuda(confusion, "confusion", torch.int32)
    N.check(N.lib().hyp_argmax_confusion(_ptr(logits), _ptr(labels), B, classes, _ptr(pred), _ptr(confusion),
                                         _stream()))
    return pred


def scatter_class_map(pred, targets_xy, class_map):
    require_cuda(pred, "pred", torch.uint8)
    require_cuda(targets_xy, "targets_xy", torch.int32)
    require_cuda(class_map, "class_map", torch.uint8)
    H, W = class_map.shape
    N.check(N.lib().hyp_scatter_class_map(_ptr(pred), _ptr(targets_xy), pred.shape[0], H, W, _ptr(class_map),
                                          _stream()))
    return class_map


def scene_minmax(cube):
    require_cuda(cube, "cube")
    dt = {torch.float32: N.HYP_DT_F32, torch.uint16: N.HYP_DT_U16}.get(cube.dtype)
    if dt is None:
        raise TypeError(f"scene dtype {cube.dtype} not supported (float32, uint16)")
    H, W, C = cube.shape
    mn = torch.empty(C, dtype=torch.float32, device=cube.device)
    mx = torch.empty(C, dtype=torch.float32, device=cube.device)
    N.check(N.lib().hyp_scene_minmax(_ptr(cube), dt, H, W, C, _ptr(mn), _ptr(mx), _stream()))
    return mn, mx


def gather_patches(casi, lidar, neighborhood, targets_xy, casi_min=None, casi_max=None, lidar_minmax=None,
                   mode=N.HYP_GATHER_SAME_RES, out=None):
    """[N, S, S, C(+1)] fp32 patches from the unpadded scene (hyp_gather_patches)."""
    require_cuda(casi, "casi")
    require_cuda(targets_xy, "targets_xy", torch.int32)
    dt = {torch.float32: N.HYP_DT_F32, torch.uint16: N.HYP_DT_U16}.get(casi.dtype)
    if dt is None:
        raise TypeError(f"casi dtype {casi.dtype} not supported (float32, uint16)")
    Hc, Wc, C = casi.shape
    n = targets_xy.shape[0]
    S = 2 * neighborhood + 1
    ch = C + (0 if lidar is None else 1)
    Hl = Wl = 0
    if lidar is not None:
        require_cuda(lidar, "lidar", torch.float32)
        Hl, Wl = lidar.shape[0], lidar.shape[1]
    if out is None:
        out = torch.empty((n, S, S, ch), dtype=torch.float32, device=casi.device)
    else:
        require_cuda(out, "out", torch.float32)
    if n == 0:  # empty target list (the reference's test split with --test_ratio 0): nothing to launch
        return out
    N.check(N.lib().hyp_gather_patches(_ptr(casi), dt, Hc, Wc, C, _ptr(casi_min), _ptr(casi_max), _ptr(lidar), Hl, Wl,
                                       _ptr(lidar_minmax), neighborhood, mode, _ptr(targets_xy), n, _ptr(out),
                                       out.shape[-1], _stream()))
    return out
