"""CycleGAN (with identity loss) on [B,1,1,C] spectra — the Wrapper of gan/wrappers/cycle_gan_wrapper.py:48-116 with
the model / loss of :189-333 (tfgan cyclegan_model + cyclegan_loss_with_identity) and the train ops of
gan/wrappers/gan_common.py:222-279, executed eagerly by the kernels of libhypelcnn_b200.so.

    generator objective   L_G = 0.5 mean((D_Y(G x) - 1)^2) + 0.5 mean((D_X(F y) - 1)^2) + 2 aux
                          aux = w_cyc (mean|x - F G x| + mean|y - G F y|) / 2 + w_id (mean|x - G x| + mean|y - F y|)
                          (aux is added to BOTH partial generator losses, which tfgan then sums: hence the 2; the
                          "identity" terms compare x with G(x), the generator on its own input domain, :308-311)
    discriminator         L_D = sum over (Y, X) of 0.5 mean((D(real) - 1)^2) + 0.5 mean(D(pool(fake))^2)
                                + discriminator_reg_scale * (|W1|^2 + |W2|^2) / 2      (slim l2_regularizer, net3 has none)
    Adam(beta1 = 0.5) for both, lr constant for the first half of max_number_of_steps, then linear to 0 (_get_lr).

One train iteration = global_step += 1, one generator step, one discriminator step (tfgan sequential hooks,
GANTrainSteps(1, 1)).  The fake batches the discriminators see go through tfgan's tensor_pool (size 50, p = 0.5), a
host-side random choice — only generator-side quantities are deterministic.
"""
import ctypes
import math
from collections import namedtuple

import numpy
import torch

from hypelcnn_b200 import _native as N
from hypelcnn_b200 import engine as E
from hypelcnn_b200.gan.shadow_data_models import GeneratorVariables
from hypelcnn_b200.gan.wrappers.gan_common import create_base_validation_hook, create_inference_for_matrix_input
from hypelcnn_b200.gan.wrappers.wrapper import InferenceWrapper, Wrapper

model_forward_generator_name = "ModelX2Y"
model_backward_generator_name = "ModelY2X"

CycleGANModel = namedtuple("CycleGANModel", ["trainer", "data_x", "data_y"])
CycleGANLoss = namedtuple("CycleGANLoss", ["trainer"])


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def get_lr(base_lr, max_number_of_steps, global_step):
    """_get_lr (gan/wrappers/gan_common.py:222-244): base_lr while global_step < max // 2, then polynomial_decay
    (power 1) to 0 at max_number_of_steps."""
    const_steps = max_number_of_steps // 2
    if global_step < const_steps:
        return base_lr
    decay_steps = max_number_of_steps - const_steps
    s = min(global_step - const_steps, decay_steps)
    return base_lr * (1.0 - s / decay_steps)


def discriminator_weight_count(bands):
    return bands * bands + bands + bands * bands + bands + bands * (bands // 2) + bands // 2


def discriminator_variable_table(bands):
    """(name, offset, shape) of the discriminator's slim variables inside its flat buffer."""
    h, off, out = bands // 2, 0, []
    for name, shape in (("fully_connected/weights", (bands, bands)), ("fully_connected/biases", (bands,)),
                        ("fully_connected_1/weights", (bands, bands)), ("fully_connected_1/biases", (bands,)),
                        ("fully_connected_2/weights", (bands, h)), ("fully_connected_2/biases", (h,))):
        out.append((name, off, shape))
        off += int(numpy.prod(shape))
    return out


class TensorPool:
    """tfgan.features.tensor_pool(pool_size=50, pooling_probability=0.5) [TF-lib]: until the pool is full the input is
    stored and returned; afterwards with probability p a random stored tensor is returned and replaced by the input."""

    def __init__(self, pool_size=50, pooling_probability=0.5, seed=1234):
        self.pool, self.size, self.p = [], pool_size, pooling_probability
        self.rng = numpy.random.default_rng(seed)

    def __call__(self, t):
        if self.size == 0:
            return t
        if len(self.pool) < self.size:
            self.pool.append(t.clone())
            return t
        if self.rng.random() < self.p:
            i = int(self.rng.integers(0, self.size))
            out, self.pool[i] = self.pool[i], t.clone()
            return out
        return t


class DeviceTensorPool:
    """tfgan.features.tensor_pool for the fused discriminator step: the stored tensors live in ONE device buffer
    [pool_size, rows, bands]; the host only draws what TensorPool draws (same generator, same order) and hands the
    kernel a (mode, slot) pair — 0: no pool, 1: store the fresh fakes into ``slot`` and use them, 2: use ``slot`` and
    replace it."""

    def __init__(self, pool_size=50, pooling_probability=0.5, seed=1234):
        self.size, self.p, self.filled, self.buffer = pool_size, pooling_probability, 0, None
        self.rng = numpy.random.default_rng(seed)

    def draw(self, rows, bands, device):
        if self.size == 0:
            return None, 0, 0
        if self.buffer is None or self.buffer.shape[1:] != (rows, bands):     # a new batch shape starts a new pool
            self.buffer = torch.empty((self.size, rows, bands), dtype=torch.float32, device=device)
            self.filled = 0
        if self.filled < self.size:
            self.filled += 1
            return self.buffer, 1, self.filled - 1
        if self.rng.random() < self.p:
            return self.buffer, 2, int(self.rng.integers(0, self.size))
        return self.buffer, 0, 0


# One warp carries a pair through a whole train op: right for launch-bound batches (the reference trains at 32; measured
# 0.35 ms per iteration against 0.85 ms for the per-op chain), wrong for throughput batches where the per-op kernels
# keep every SM full (batch 16 384: 2.8 ms per iteration chained, 7.3 ms fused)
FUSED_MAX_ROWS = 1024


def fused_steps_enabled():
    import os
    return os.environ.get("HYP_GAN_FUSED", "1") != "0"


class GanKernels:
    """The C-ABI calls every GAN trainer chains (needs self.C and self.loss_acc)."""

    def _gen_fwd(self, x, w):
        nets = torch.empty((x.shape[0], 8, self.C), dtype=torch.float32, device=x.device)
        N.check(N.lib().hyp_gan_generator_train_forward(_p(x), x.shape[0], self.C, _p(w), _p(nets), _st()))
        return nets

    def _gen_bwd(self, nets, gout, w, gw, need_gin=True):
        gin = torch.empty((nets.shape[0], self.C), dtype=torch.float32, device=nets.device) if need_gin else None
        N.check(N.lib().hyp_gan_generator_backward(_p(nets), _p(gout), nets.shape[0], self.C, _p(w), _p(gin), _p(gw), _st()))
        return gin

    def _dis_fwd(self, x, w):
        h = torch.empty((x.shape[0], 2, self.C), dtype=torch.float32, device=x.device)
        out = torch.empty((x.shape[0], self.C // 2), dtype=torch.float32, device=x.device)
        N.check(N.lib().hyp_gan_discriminator_forward(_p(x), x.shape[0], self.C, _p(w), _p(h), _p(out), _st()))
        return h, out

    def _dis_bwd(self, x, h, gout, w, gw=None, need_gin=False):
        gin = torch.empty((x.shape[0], self.C), dtype=torch.float32, device=x.device) if need_gin else None
        N.check(N.lib().hyp_gan_discriminator_backward(_p(x), _p(h), _p(gout), x.shape[0], self.C, _p(w), _p(gin), _p(gw), _st()))
        return gin

    def _loss(self, mode, a, b, target, weight, grad, accumulate, slot):
        N.check(N.lib().hyp_gan_loss_grad(mode, _p(a), _p(b), target, weight / a.numel(), a.numel(), _p(grad),
                                          int(accumulate), ctypes.c_void_p(self.loss_acc[slot:].data_ptr()), _st()))

    @staticmethod
    def _rows(t):
        if t.dim() == 4:
            t = t.reshape(t.shape[0], t.shape[3])
        if not t.is_cuda or t.dtype != torch.float32:
            raise TypeError("spectra must be CUDA float32 tensors (no CPU path)")
        return t.contiguous()


def init_discriminator(params, bands, rng):
    """variance_scaling(scale=2.0): fan_in, truncated normal (gan/shadow_data_models.py:95) into a flat buffer."""
    for name, off, shape in discriminator_variable_table(bands):
        if name.endswith("weights"):
            params[off:off + shape[0] * shape[1]].copy_(torch.from_numpy(truncated_normal(rng, shape, math.sqrt(2.0 / shape[0]))))


def truncated_normal(rng, shape, std):
    w = rng.standard_normal(shape)
    bad = numpy.abs(w) > 2.0
    while bad.any():
        w[bad] = rng.standard_normal(int(bad.sum()))
        bad = numpy.abs(w) > 2.0
    return (w * (std / E.TRUNC_STD_FIX)).astype(numpy.float32).ravel()


class CycleGANTrainer(GanKernels):
    """Variables, optimizer slots and the two train ops.  Generators: [G (x -> y) | F (y -> x)] in one flat buffer,
    discriminators: [D_Y | D_X] in another, so each Adam step and (multi-GPU) each all-reduce is one call."""

    def __init__(self, bands, cycle_consistency_loss_weight=10.0, identity_loss_weight=0.5, use_identity_loss=True,
                 discriminator_reg_scale=1e-5, device=None, seed=1234, pool_size=50):
        if not torch.cuda.is_available():
            raise N.NativeError(N.HYP_E_CUDA, "no CUDA device: hypelcnn_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.C = int(bands)
        self.w_cyc, self.w_id = float(cycle_consistency_loss_weight), float(identity_loss_weight) if use_identity_loss else 0.0
        self.reg = float(discriminator_reg_scale)
        self.gen_x2y, self.gen_y2x = GeneratorVariables(bands, False, self.device), GeneratorVariables(bands, False, self.device)
        ng, nd = self.gen_x2y.flat.numel(), discriminator_weight_count(bands)
        z = dict(dtype=torch.float32, device=self.device)
        self.gen_params = torch.zeros(2 * ng, **z)   # zeros: shadow_data_models.py:47
        self.gen_x2y.flat, self.gen_y2x.flat = self.gen_params[:ng], self.gen_params[ng:]
        self.dis_params = torch.zeros(2 * nd, **z)
        rng = numpy.random.default_rng(seed)
        for d in range(2):
            init_discriminator(self.dis_params[d * nd:(d + 1) * nd], bands, rng)
        self.ng, self.nd = ng, nd
        self.gen_grads, self.dis_grads = torch.zeros_like(self.gen_params), torch.zeros_like(self.dis_params)
        self.gen_m, self.gen_v = torch.zeros_like(self.gen_params), torch.zeros_like(self.gen_params)
        self.dis_m, self.dis_v = torch.zeros_like(self.dis_params), torch.zeros_like(self.dis_params)
        self.gen_steps = self.dis_steps = self.global_step = 0
        self.pool_y, self.pool_x = TensorPool(pool_size, seed=seed), TensorPool(pool_size, seed=seed + 1)
        self.loss_acc = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.allreduce = None  # set to a parallel.GradientAllReduce for data-parallel training
        self.last = {}
        # the train ops compute all gradients of a step in ONE kernel each (hyp_gan_cycle_generator_step /
        # hyp_gan_cycle_discriminator_step); the per-op chain below (generator_gradients / discriminator_gradients) is
        # what they are tested against and what other band counts fall back to
        self.use_fused = fused_steps_enabled() and self.C % 8 == 0 and 8 <= self.C <= 64
        self.dev_pool_y, self.dev_pool_x = DeviceTensorPool(pool_size, seed=seed), DeviceTensorPool(pool_size, seed=seed + 1)
        self._outs = None

    # ---- views
    def G(self):
        return self.gen_params[:self.ng]

    def F(self):
        return self.gen_params[self.ng:]

    def DY(self):
        return self.dis_params[:self.nd]

    def DX(self):
        return self.dis_params[self.nd:]

    # ---- generator step (tfgan generator train op over both partial models)
    def generator_gradients(self, images_x, images_y):
        """Fills gen_grads with dL_G / d[G | F]; returns the device loss tensor [total, gan, cycle, identity]."""
        x, y = self._rows(images_x), self._rows(images_y)
        G, Fw, ng = self.G(), self.F(), self.ng
        gG, gF = self.gen_grads[:ng], self.gen_grads[ng:]
        self.gen_grads.zero_()
        self.loss_acc.zero_()
        n_gx, n_fy = self._gen_fwd(x, G), self._gen_fwd(y, Fw)            # G(x), F(y)
        gx, fy = n_gx[:, 7, :].contiguous(), n_fy[:, 7, :].contiguous()
        n_rx, n_ry = self._gen_fwd(gx, Fw), self._gen_fwd(fy, G)          # F(G x), G(F y)
        rx, ry = n_rx[:, 7, :].contiguous(), n_ry[:, 7, :].contiguous()
        h_y, d_gx = self._dis_fwd(gx, self.DY())
        h_x, d_fy = self._dis_fwd(fy, self.DX())
        g_dgx, g_dfy = torch.empty_like(d_gx), torch.empty_like(d_fy)
        self._loss(0, d_gx, None, 1.0, 1.0, g_dgx, False, 1)             # least_squares_generator_loss
        self._loss(0, d_fy, None, 1.0, 1.0, g_dfy, False, 1)
        g_gx = self._dis_bwd(gx, h_y, g_dgx, self.DY(), None, True)       # through the (frozen) discriminators
        g_fy = self._dis_bwd(fy, h_x, g_dfy, self.DX(), None, True)
        g_rx, g_ry = torch.empty_like(rx), torch.empty_like(ry)
        self._loss(1, rx, x, 0.0, 2.0 * self.w_cyc / 2.0, g_rx, False, 2)  # cycle consistency, counted twice
        self._loss(1, ry, y, 0.0, 2.0 * self.w_cyc / 2.0, g_ry, False, 2)
        if self.w_id != 0.0:
            self._loss(1, gx, x, 0.0, 2.0 * self.w_id, g_gx, True, 3)     # "identity": |x - G(x)|
            self._loss(1, fy, y, 0.0, 2.0 * self.w_id, g_fy, True, 3)
        g_gx += self._gen_bwd(n_rx, g_rx, Fw, gF)                         # F applied to G(x)
        g_fy += self._gen_bwd(n_ry, g_ry, G, gG)                          # G applied to F(y)
        self._gen_bwd(n_gx, g_gx, G, gG, need_gin=False)
        self._gen_bwd(n_fy, g_fy, Fw, gF, need_gin=False)
        self.last = {"generated_y": gx, "generated_x": fy, "reconstructed_x": rx, "reconstructed_y": ry}
        loss = self.loss_acc.clone()
        loss[0] = loss[1] + loss[2] + loss[3]
        return loss

    def generator_gradients_fused(self, images_x, images_y):
        """generator_gradients in one launch (hyp_gan_cycle_generator_step)."""
        x, y = self._rows(images_x), self._rows(images_y)
        if self._outs is None or self._outs[0].shape != x.shape:
            self._outs = [torch.empty_like(x) for _ in range(4)]
        gx, fy, rx, ry = self._outs
        self.gen_grads.zero_()
        self.loss_acc.zero_()
        ng = self.ng
        N.check(N.lib().hyp_gan_cycle_generator_step(
            _p(x), _p(y), x.shape[0], self.C, _p(self.G()), _p(self.F()), _p(self.DY()), _p(self.DX()), self.w_cyc,
            self.w_id, _p(self.gen_grads[:ng]), _p(self.gen_grads[ng:]), ctypes.c_void_p(self.loss_acc.data_ptr()),
            _p(gx), _p(fy), _p(rx), _p(ry), _st()))
        self.last = {"generated_y": gx, "generated_x": fy, "reconstructed_x": rx, "reconstructed_y": ry}
        return self.loss_acc.clone()

    def discriminator_gradients_fused(self, images_x, images_y, use_pool=True):
        """discriminator_gradients in one launch (hyp_gan_cycle_discriminator_step): fakes, tensor pool, both
        discriminators on real and fake, regulariser."""
        x, y = self._rows(images_x), self._rows(images_y)
        nd = self.nd
        self.dis_grads.zero_()
        self.loss_acc.zero_()
        pool_y = self.dev_pool_y.draw(x.shape[0], self.C, x.device) if use_pool else (None, 0, 0)
        pool_x = self.dev_pool_x.draw(x.shape[0], self.C, x.device) if use_pool else (None, 0, 0)
        N.check(N.lib().hyp_gan_cycle_discriminator_step(
            _p(x), _p(y), x.shape[0], self.C, _p(self.G()), _p(self.F()), _p(self.DY()), _p(self.DX()), self.reg,
            _p(self.dis_grads[:nd]), _p(self.dis_grads[nd:]), ctypes.c_void_p(self.loss_acc.data_ptr()),
            _p(pool_y[0]), pool_y[1], pool_y[2], _p(pool_x[0]), pool_x[1], pool_x[2], _st()))
        return self.loss_acc.clone()

    def generator_train_op(self, images_x, images_y, lr):
        if self.use_fused and images_x.shape[0] <= FUSED_MAX_ROWS:
            loss = self.generator_gradients_fused(images_x, images_y)
        else:
            loss = self.generator_gradients(images_x, images_y)
        scale = self.allreduce(self.gen_grads) if self.allreduce is not None else 1.0
        self.gen_steps += 1
        E.adam_step(self.gen_params, self.gen_grads, self.gen_m, self.gen_v, lr, self.gen_steps, scale, b1=0.5)
        return loss

    # ---- discriminator step
    def discriminator_gradients(self, images_x, images_y, use_pool=True):
        """Fills dis_grads with dL_D / d[D_Y | D_X]; returns the device loss tensor [total, gan, regularisation, -]."""
        x, y = self._rows(images_x), self._rows(images_y)
        nd = self.nd
        gDY, gDX = self.dis_grads[:nd], self.dis_grads[nd:]
        self.dis_grads.zero_()
        self.loss_acc.zero_()
        gx = self._gen_fwd(x, self.G())[:, 7, :].contiguous()
        fy = self._gen_fwd(y, self.F())[:, 7, :].contiguous()
        if use_pool:
            gx, fy = self.pool_y(gx), self.pool_x(fy)
        for real, fake, w, gw in ((y, gx, self.DY(), gDY), (x, fy, self.DX(), gDX)):
            for data, target in ((real, 1.0), (fake, 0.0)):            # least_squares_discriminator_loss
                h, d = self._dis_fwd(data, w)
                g = torch.empty_like(d)
                self._loss(0, d, None, target, 1.0, g, False, 1)
                self._dis_bwd(data, h, g, w, gw, False)
            C = self.C                                                  # l2_regularizer on net1, net2 weights
            for off in (0, C * C + C):
                N.check(N.lib().hyp_gan_l2_regularizer(_p(w[off:]), _p(gw[off:]), C * C, self.reg,
                                                       ctypes.c_void_p(self.loss_acc[2:].data_ptr()), _st()))
        loss = self.loss_acc.clone()
        loss[0] = loss[1] + loss[2]
        return loss

    def discriminator_train_op(self, images_x, images_y, lr):
        loss = self.discriminator_gradients_fused(images_x, images_y) \
            if self.use_fused and images_x.shape[0] <= FUSED_MAX_ROWS else \
            self.discriminator_gradients(images_x, images_y)
        scale = self.allreduce(self.dis_grads) if self.allreduce is not None else 1.0
        self.dis_steps += 1
        E.adam_step(self.dis_params, self.dis_grads, self.dis_m, self.dis_v, lr, self.dis_steps, scale, b1=0.5)
        return loss


class GANTrainOps:
    """What define_train_ops returns (tfgan GANTrainOps): generator_train_op, discriminator_train_op and the
    global_step_inc_op the training loop runs every iteration (gan/gan_train_for_shadow.py:142)."""

    def __init__(self, trainer, max_number_of_steps, generator_lr, discriminator_lr):
        self.trainer, self.max_steps = trainer, max_number_of_steps
        self.generator_lr, self.discriminator_lr = generator_lr, discriminator_lr

    def global_step_inc_op(self):
        self.trainer.global_step += 1
        return self.trainer.global_step

    def generator_train_op(self, images_x, images_y):
        return self.trainer.generator_train_op(images_x, images_y,
                                               get_lr(self.generator_lr, self.max_steps, self.trainer.global_step))

    def discriminator_train_op(self, images_x, images_y):
        return self.trainer.discriminator_train_op(images_x, images_y,
                                                   get_lr(self.discriminator_lr, self.max_steps, self.trainer.global_step))

    def train_iteration(self, images_x, images_y):
        """global_step += 1, then the sequential hooks: 1 generator step, 1 discriminator step."""
        self.global_step_inc_op()
        lg = self.generator_train_op(images_x, images_y)
        ld = self.discriminator_train_op(images_x, images_y)
        return lg, ld


class CycleGANWrapper(Wrapper):
    """Same constructor / method names as the reference (cycle_gan_wrapper.py:48-116).  generator_fn / discriminator_fn
    are accepted for signature compatibility; the models are the fixed shadow_data_models architectures."""

    def __init__(self, cycle_consistency_loss_weight, identity_loss_weight, use_identity_loss,
                 generator_fn=None, discriminator_fn=None, discriminator_reg_scale=1e-5) -> None:
        super().__init__()
        self._cycle_consistency_loss_weight = cycle_consistency_loss_weight
        self._identity_loss_weight = identity_loss_weight
        self._use_identity_loss = use_identity_loss
        self._discriminator_reg_scale = discriminator_reg_scale
        self.trainer = None

    def define_model(self, images_x, images_y):
        bands = images_x.shape[-1]
        if self.trainer is None:
            self.trainer = CycleGANTrainer(bands, self._cycle_consistency_loss_weight, self._identity_loss_weight,
                                           self._use_identity_loss, self._discriminator_reg_scale, images_x.device)
        return CycleGANModel(self.trainer, images_x, images_y)

    def define_loss(self, model):
        return CycleGANLoss(model.trainer)

    def define_train_ops(self, model, loss, max_number_of_steps, **kwargs):
        return GANTrainOps(model.trainer, max_number_of_steps, kwargs["generator_lr"], kwargs["discriminator_lr"])

    def get_train_hooks_fn(self):
        return lambda train_ops: [train_ops.generator_train_op, train_ops.discriminator_train_op]  # GANTrainSteps(1, 1)


class CycleGANInferenceWrapper(InferenceWrapper):
    """cycle_gan_wrapper.py:119-166: forward (x -> y, "shadow") / backward generator over a [B,H,W,C] matrix."""

    def __init__(self, shadow_generator_fn=None, trainer=None, bands=None):
        if trainer is not None:
            self.forward_generator, self.backward_generator = trainer.gen_x2y, trainer.gen_y2x
        else:
            self.forward_generator, self.backward_generator = GeneratorVariables(bands), GeneratorVariables(bands)

    def construct_inference_graph(self, input_tensor, is_shadow_graph, clip_invalid_values, copy_extra=0):
        gen = self.forward_generator if is_shadow_graph else self.backward_generator
        return create_inference_for_matrix_input(input_tensor, is_shadow_graph, clip_invalid_values, gen, copy_extra)

    def make_inference_graph(self, data_set, is_shadow_graph, clip_invalid_values):
        return None, lambda x: self.construct_inference_graph(x, is_shadow_graph, clip_invalid_values)

    def create_generator_restorer(self):
        return self

    def restore(self, forward_values=None, backward_values=None):
        if forward_values is not None:
            self.forward_generator.load(forward_values)
        if backward_values is not None:
            self.backward_generator.load(backward_values)

    def create_inference_hook(self, data_set, loader, log_dir, neighborhood, shadow_map, shadow_ratio,
                              validation_iteration_count, validation_sample_count):
        """cycle_gan_wrapper.py:149-166: peer validation of both generators (no clipping of invalid values)."""
        return create_base_validation_hook(
            data_set=data_set, loader=loader, log_dir=log_dir, neighborhood=neighborhood, shadow_map=shadow_map,
            shadow_ratio=shadow_ratio, validation_iteration_count=validation_iteration_count,
            validation_sample_count=validation_sample_count,
            model_forward=lambda x: self.construct_inference_graph(x, True, False),
            model_backward=lambda x: self.construct_inference_graph(x, False, False))
