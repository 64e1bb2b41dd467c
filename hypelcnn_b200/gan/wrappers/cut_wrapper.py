"""CUT (contrastive unpaired translation) on [B,1,1,C] spectra — the Wrapper of gan/wrappers/cut_wrapper.py:587-665 with
the model of :256-356, the losses of :90-208 / :360-420 and the three train ops of :467-584, executed eagerly by the
kernels of libhypelcnn_b200.so.

    gen = G(inp)                      idt = G(real)                       Enc = G's first four layers (net4)
    generator      L_G = 0.5 mean((D(gen) - 1)^2) + w_nce NCE(F(Enc gen), F(Enc inp)) + w_id NCE(F(Enc idt), F(Enc real))
    discriminator  L_D = 0.5 mean((D(real) - 1)^2) + 0.5 mean(D(gen)^2) + reg (no tensor pool: tensor_pool_fn=None, :171)
    feature disc.  L_F = NCE(F(Enc gen), F(Enc inp)) + l2 reg over F's weight matrices            (:198-208)
    NCE(a, b)      = mean_b [ -sum_i log_softmax(flatten(a_b b_b^T / tau))[i, i] ]                 (:360-393)

F is the patch feature discriminator (gan/shadow_data_models.py:126-149).  The generator gradient flows through the
frozen D and F and through every use of G (G(inp), Enc(gen), Enc(inp), G(real), Enc(idt), Enc(real)); the encoder of
inp / real is read out of the full forward pass (net4 of the same kernel launch).  swap_inputs=True trains y -> x.
One train iteration = global_step += 1, then generator, discriminator, feature-discriminator steps (:67-87,
CUTTrainSteps(1, 1, 1)), each Adam(beta1 0.5) with the _get_lr schedule.

TensorFlow evaluates tf.nn.softmax_cross_entropy_with_logits with a fused kernel whose gradient is softmax - labels
[TF-lib]; the reference's labels (a flattened identity) sum to `slices`, so this is not the exact derivative of the
loss it reports.  fused_xent_grad=True (default) trains with the reference's gradient, False with the exact one.
"""
import ctypes
import math
from collections import namedtuple

import numpy
import torch

from hypelcnn_b200 import _native as N
from hypelcnn_b200 import engine as E
from hypelcnn_b200.gan.shadow_data_models import GeneratorVariables
from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import (GanKernels, _p, _st, discriminator_variable_table,
                                                          discriminator_weight_count, get_lr, init_discriminator,
                                                          truncated_normal)
from hypelcnn_b200.gan.wrappers.gan_wrapper import GANInferenceWrapper
from hypelcnn_b200.gan.wrappers.wrapper import Wrapper

CUTTrainSteps = namedtuple("CUTTrainSteps", ["generator_train_steps", "discriminator_train_steps",
                                             "gen_discriminator_train_steps"])
CUTModel = namedtuple("CUTModel", ["trainer", "generator_inputs", "real_data"])
CUTLoss = namedtuple("CUTLoss", ["trainer"])


def feature_discriminator_variable_table(bands, patch_count, embedded_feature_size):
    """(name, offset, shape) of the feature discriminator's slim variables in its flat buffer: slice s, layer l is
    fully_connected_{4 s + l} (slim's default scope names in creation order, shadow_data_models.py:141-146)."""
    ps = bands // patch_count
    dims = [ps, ps, ps // 4, ps // 2, embedded_feature_size]
    n_full = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(4))
    out = []
    for s, start in enumerate(range(0, bands, ps)):
        d = [min(ps, bands - start)] + dims[1:]
        off = s * n_full
        for l in range(4):
            k = 4 * s + l
            scope = "fully_connected" if k == 0 else f"fully_connected_{k}"
            out.append((f"{scope}/weights", off, (d[l], d[l + 1])))
            out.append((f"{scope}/biases", off + d[l] * d[l + 1], (d[l + 1],)))
            off += d[l] * d[l + 1] + d[l + 1]
    return out


class CUTTrainer(GanKernels):
    """Variables (one generator, one discriminator, one feature discriminator), Adam slots and the three train ops.
    `clock` holds global_step and the three optimizers' step counts; DCLGAN shares one clock between its two CUT
    models because the reference hands both the same AdamOptimizer objects (dcl_gan_wrapper.py:285-309)."""

    def __init__(self, bands, nce_loss_weight=10.0, identity_loss_weight=0.5, use_identity_loss=True, tau=0.07,
                 patches=6, embedded_feat_size=2, swap_inputs=False, discriminator_reg_scale=1e-5,
                 gen_disc_reg_scale=1e-4, device=None, seed=1234, fused_xent_grad=True, clock=None):
        if not torch.cuda.is_available():
            raise N.NativeError(N.HYP_E_CUDA, "no CUDA device: hypelcnn_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.C, self.patches, self.E = int(bands), int(patches), int(embedded_feat_size)
        self.ps = self.C // self.patches
        self.slices = (self.C + self.ps - 1) // self.ps
        self.w_nce = float(nce_loss_weight)
        self.w_id = float(identity_loss_weight) if use_identity_loss else 0.0   # cut_wrapper.py:593
        self.tau, self.swap, self.fused = float(tau), bool(swap_inputs), bool(fused_xent_grad)
        self.reg, self.feat_reg = float(discriminator_reg_scale), float(gen_disc_reg_scale)
        self.generator = GeneratorVariables(bands, False, self.device)
        self.ng, self.nd = self.generator.flat.numel(), discriminator_weight_count(bands)
        self.nf = int(N.lib().hyp_gan_feature_discriminator_weight_count(self.C, self.patches, self.E))
        if self.nf <= 0:
            raise ValueError("bad patches / embedded_feat_size for this band count")
        z = dict(dtype=torch.float32, device=self.device)
        self.gen_params = self.generator.flat                      # zeros (shadow_data_models.py:47)
        self.dis_params, self.feat_params = torch.zeros(self.nd, **z), torch.zeros(self.nf, **z)
        rng = numpy.random.default_rng(seed)
        init_discriminator(self.dis_params, bands, rng)
        self.feat_table = feature_discriminator_variable_table(self.C, self.patches, self.E)
        for name, off, shape in self.feat_table:                   # variance_scaling(scale=2.0), :127
            if name.endswith("weights"):
                w = truncated_normal(rng, shape, math.sqrt(2.0 / shape[0]))
                self.feat_params[off:off + w.size].copy_(torch.from_numpy(w))
        self.gen_grads, self.dis_grads, self.feat_grads = (torch.zeros_like(t) for t in
                                                           (self.gen_params, self.dis_params, self.feat_params))
        self.slots = {k: (torch.zeros_like(p), torch.zeros_like(p)) for k, p in
                      (("gen", self.gen_params), ("dis", self.dis_params), ("feat", self.feat_params))}
        self.clock = clock if clock is not None else {"global_step": 0, "gen": 0, "dis": 0, "feat": 0}
        self.loss_acc = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.allreduce = None
        self.last = {}

    # ---- kernels beyond GanKernels
    def _gen_bwd_enc(self, nets, gout, gout_enc, need_gin):
        gin = torch.empty((nets.shape[0], self.C), dtype=torch.float32, device=nets.device) if need_gin else None
        N.check(N.lib().hyp_gan_generator_backward_enc(_p(nets), _p(gout), _p(gout_enc), nets.shape[0], self.C,
                                                       _p(self.gen_params), _p(gin), _p(self.gen_grads), _st()))
        return gin

    def _feat_fwd(self, e):
        z = torch.empty((e.shape[0], self.slices, self.E), dtype=torch.float32, device=e.device)
        ss = torch.empty(self.slices, dtype=torch.float32, device=e.device)
        N.check(N.lib().hyp_gan_feature_discriminator_forward(_p(e), e.shape[0], self.C, self.patches, self.E,
                                                              _p(self.feat_params), _p(z), _p(ss), _st()))
        return z, ss

    def _feat_bwd(self, e, z, ss, gf, dot, need_gin, gweights):
        gin = torch.empty_like(e) if need_gin else None
        N.check(N.lib().hyp_gan_feature_discriminator_backward(_p(e), _p(z), _p(ss), _p(gf), _p(dot), e.shape[0], self.C,
                                                               self.patches, self.E, _p(self.feat_params), _p(gin),
                                                               _p(gweights), _st()))
        return gin

    def _nce(self, zg, ssg, zr, ssr, weight, slot, want_grad=True):
        gg = torch.empty_like(zg) if want_grad else None
        gr = torch.empty_like(zr) if want_grad else None
        dg = torch.empty(self.slices, dtype=torch.float32, device=zg.device) if want_grad else None
        dr = torch.empty(self.slices, dtype=torch.float32, device=zg.device) if want_grad else None
        N.check(N.lib().hyp_gan_patchnce(_p(zg), _p(zr), _p(ssg), _p(ssr), zg.shape[0], self.slices, self.E, self.tau,
                                         weight / zg.shape[0], int(self.fused), _p(gg), _p(gr), _p(dg), _p(dr),
                                         ctypes.c_void_p(self.loss_acc[slot:].data_ptr()), _st()))
        return gg, gr, dg, dr

    def _pick(self, images_x, images_y):
        x, y = self._rows(images_x), self._rows(images_y)
        return (y, x) if self.swap else (x, y)   # (generator_inputs, real_data)  cut_wrapper.py:611-616

    def feature_embeddings(self, spectra):
        """F(Enc(spectra)) normalised like the reference's feat_discriminator_* tensors: [B, slices, E]."""
        nets = self._gen_fwd(self._rows(spectra), self.gen_params)
        z, ss = self._feat_fwd(nets[:, 4, :].contiguous())
        return z * torch.rsqrt(torch.clamp(ss, min=1e-12)).view(1, -1, 1)

    # ---- generator step
    def _nce_branch(self, nets_a, nets_b, weight, slot):
        """NCE(F(Enc a), F(Enc b)) where Enc a / Enc b are net4 of two saved forward passes; returns dL/dnet4 of both."""
        ea, eb = nets_a[:, 4, :].contiguous(), nets_b[:, 4, :].contiguous()
        (za, sa), (zb, sb) = self._feat_fwd(ea), self._feat_fwd(eb)
        ga, gb, da, db = self._nce(za, sa, zb, sb, weight, slot)
        return self._feat_bwd(ea, za, sa, ga, da, True, None), self._feat_bwd(eb, zb, sb, gb, db, True, None)

    def generator_gradients(self, images_x, images_y):
        """Fills gen_grads with dL_G/dG; returns the device loss tensor [total, gan, w_nce*nce_x, w_id*nce_identity]."""
        inp, real = self._pick(images_x, images_y)
        G, D = self.gen_params, self.dis_params
        self.gen_grads.zero_()
        self.loss_acc.zero_()
        n_gen = self._gen_fwd(inp, G)
        gen = n_gen[:, 7, :].contiguous()
        h, d = self._dis_fwd(gen, D)
        g_d = torch.empty_like(d)
        self._loss(0, d, None, 1.0, 1.0, g_d, False, 1)                      # least_squares_generator_loss
        g_gen = self._dis_bwd(gen, h, g_d, D, None, True)
        n_eg = self._gen_fwd(gen, G)                                         # Enc(gen) = net4 of G(gen)
        ge_gen, ge_inp = self._nce_branch(n_eg, n_gen, self.w_nce, 2)
        g_gen += self._gen_bwd_enc(n_eg, None, ge_gen, True)
        if self.w_id != 0.0:
            n_idt = self._gen_fwd(real, G)                                   # G(real), Enc(real)
            idt = n_idt[:, 7, :].contiguous()
            n_ei = self._gen_fwd(idt, G)                                     # Enc(idt)
            ge_idt, ge_real = self._nce_branch(n_ei, n_idt, self.w_id, 3)
            g_idt = self._gen_bwd_enc(n_ei, None, ge_idt, True)
            self._gen_bwd_enc(n_idt, g_idt, ge_real, False)
        self._gen_bwd_enc(n_gen, g_gen, ge_inp, False)
        self.last = {"generated": gen}
        loss = self.loss_acc.clone()
        loss[0] = loss[1] + loss[2] + loss[3]
        return loss

    # ---- discriminator step (no tensor pool in CUT)
    def discriminator_gradients(self, images_x, images_y):
        inp, real = self._pick(images_x, images_y)
        D, gD, C = self.dis_params, self.dis_grads, self.C
        self.dis_grads.zero_()
        self.loss_acc.zero_()
        gen = self._gen_fwd(inp, self.gen_params)[:, 7, :].contiguous()
        for data, target in ((real, 1.0), (gen, 0.0)):                       # least_squares_discriminator_loss
            h, d = self._dis_fwd(data, D)
            g = torch.empty_like(d)
            self._loss(0, d, None, target, 1.0, g, False, 1)
            self._dis_bwd(data, h, g, D, gD, False)
        for off in (0, C * C + C):
            N.check(N.lib().hyp_gan_l2_regularizer(_p(D[off:]), _p(gD[off:]), C * C, self.reg,
                                                   ctypes.c_void_p(self.loss_acc[2:].data_ptr()), _st()))
        loss = self.loss_acc.clone()
        loss[0] = loss[1] + loss[2]
        return loss

    # ---- feature discriminator step
    def feat_discriminator_gradients(self, images_x, images_y):
        """dL_F/dF, L_F = NCE(F(Enc gen), F(Enc inp)) + regularisation; returns [total, nce_x, reg, 0]."""
        inp, _ = self._pick(images_x, images_y)
        Fw, gF = self.feat_params, self.feat_grads
        self.feat_grads.zero_()
        self.loss_acc.zero_()
        n_gen = self._gen_fwd(inp, self.gen_params)
        n_eg = self._gen_fwd(n_gen[:, 7, :].contiguous(), self.gen_params)
        ea, eb = n_eg[:, 4, :].contiguous(), n_gen[:, 4, :].contiguous()
        (za, sa), (zb, sb) = self._feat_fwd(ea), self._feat_fwd(eb)
        ga, gb, da, db = self._nce(za, sa, zb, sb, 1.0, 1)
        self._feat_bwd(ea, za, sa, ga, da, False, gF)
        self._feat_bwd(eb, zb, sb, gb, db, False, gF)
        for name, off, shape in self.feat_table:
            if name.endswith("weights"):
                N.check(N.lib().hyp_gan_l2_regularizer(_p(Fw[off:]), _p(gF[off:]), shape[0] * shape[1], self.feat_reg,
                                                       ctypes.c_void_p(self.loss_acc[2:].data_ptr()), _st()))
        loss = self.loss_acc.clone()
        loss[0] = loss[1] + loss[2]
        return loss

    # ---- train ops
    def _apply(self, key, params, grads, lr):
        scale = self.allreduce(grads) if self.allreduce is not None else 1.0
        self.clock[key] += 1
        m, v = self.slots[key]
        E.adam_step(params, grads, m, v, lr, self.clock[key], scale, b1=0.5)

    def generator_train_op(self, images_x, images_y, lr):
        loss = self.generator_gradients(images_x, images_y)
        self._apply("gen", self.gen_params, self.gen_grads, lr)
        return loss

    def discriminator_train_op(self, images_x, images_y, lr):
        loss = self.discriminator_gradients(images_x, images_y)
        self._apply("dis", self.dis_params, self.dis_grads, lr)
        return loss

    def gen_discriminator_train_op(self, images_x, images_y, lr):
        loss = self.feat_discriminator_gradients(images_x, images_y)
        self._apply("feat", self.feat_params, self.feat_grads, lr)
        return loss

    def variables(self):
        """{scope/name: array} under the reference's scope names (cut_wrapper.py:264-266)."""
        out = {f"Generator/{k}": v for k, v in self.generator.export().items()}
        for scope, table, flat in (("Discriminator", discriminator_variable_table(self.C), self.dis_params),
                                   ("FeatDiscriminator", self.feat_table, self.feat_params)):
            for name, off, shape in table:
                out[f"{scope}/{name}"] = flat[off:off + int(numpy.prod(shape))].cpu().numpy().reshape(shape)
        return out


class CUTTrainOps:
    """cut_train_ops' result (cut_wrapper.py:48-64): generator / discriminator / gen_discriminator train ops and the
    global_step_inc_op the training loop runs each iteration."""

    def __init__(self, trainer, max_number_of_steps, generator_lr, discriminator_lr, gen_discriminator_lr):
        self.trainer, self.max_steps = trainer, max_number_of_steps
        self.generator_lr, self.discriminator_lr, self.gen_discriminator_lr = generator_lr, discriminator_lr, gen_discriminator_lr
        self.train_hooks = []

    def _lr(self, base):
        return get_lr(base, self.max_steps, self.trainer.clock["global_step"])

    def global_step_inc_op(self):
        self.trainer.clock["global_step"] += 1
        return self.trainer.clock["global_step"]

    def generator_train_op(self, images_x, images_y):
        return self.trainer.generator_train_op(images_x, images_y, self._lr(self.generator_lr))

    def discriminator_train_op(self, images_x, images_y):
        return self.trainer.discriminator_train_op(images_x, images_y, self._lr(self.discriminator_lr))

    def gen_discriminator_train_op(self, images_x, images_y):
        return self.trainer.gen_discriminator_train_op(images_x, images_y, self._lr(self.gen_discriminator_lr))

    def run_sequential(self, images_x, images_y):
        """get_sequential_train_hooks_cut(CUTTrainSteps(1, 1, 1)) (:67-87): generator, discriminator, feature disc."""
        return (self.generator_train_op(images_x, images_y), self.discriminator_train_op(images_x, images_y),
                self.gen_discriminator_train_op(images_x, images_y))

    def train_iteration(self, images_x, images_y):
        self.global_step_inc_op()
        return self.run_sequential(images_x, images_y)


class CUTWrapper(Wrapper):
    """Same constructor / method names as the reference (cut_wrapper.py:587-665).  The *_fn arguments are accepted for
    signature compatibility; patches / embedded_feat_size / the regularisation scales are what the reference binds
    into feat_discriminator_fn / discriminator_fn with functools.partial (gan/wrapper_registry.py:35-40)."""

    def __init__(self, nce_loss_weight, identity_loss_weight, use_identity_loss, tau, batch_size, swap_inputs,
                 generator_fn=None, discriminator_fn=None, feat_discriminator_fn=None, patches=6, embedded_feat_size=2,
                 discriminator_reg_scale=1e-5, gen_disc_reg_scale=1e-4) -> None:
        super().__init__()
        self._nce_loss_weight = nce_loss_weight
        self._identity_loss_weight = 0.0 if not use_identity_loss else identity_loss_weight
        self._swap_inputs = swap_inputs
        self._tau = tau
        self._batch_size = batch_size
        self._model_args = dict(patches=patches, embedded_feat_size=embedded_feat_size,
                                discriminator_reg_scale=discriminator_reg_scale, gen_disc_reg_scale=gen_disc_reg_scale)
        self.trainer = None

    def define_model(self, images_x, images_y):
        if self.trainer is None:
            self.trainer = CUTTrainer(images_x.shape[-1], self._nce_loss_weight, self._identity_loss_weight, True,
                                      self._tau, swap_inputs=self._swap_inputs, device=images_x.device,
                                      **self._model_args)
        gi, rd = (images_y, images_x) if self._swap_inputs else (images_x, images_y)
        return CUTModel(self.trainer, gi, rd)

    def define_loss(self, model):
        return CUTLoss(model.trainer)

    def define_train_ops(self, model, loss, max_number_of_steps, **kwargs):
        return CUTTrainOps(model.trainer, max_number_of_steps, kwargs["generator_lr"], kwargs["discriminator_lr"],
                           kwargs["gen_discriminator_lr"])

    def get_train_hooks_fn(self):
        return lambda train_ops: [train_ops.generator_train_op, train_ops.discriminator_train_op,
                                  train_ops.gen_discriminator_train_op]


class CUTInferenceWrapper(GANInferenceWrapper):
    """cut_wrapper.py:668-: inference is the single generator, exactly as for the plain GAN wrapper."""

    def __init__(self, fetch_shadows, shadow_generator_fn=None, trainer=None, bands=None):
        if trainer is not None:
            self._fetch_shadows, self.generator = fetch_shadows, trainer.generator
        else:
            super().__init__(fetch_shadows, shadow_generator_fn, None, bands)
