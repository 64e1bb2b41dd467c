"""Single-direction GAN (gan_x2y / gan_y2x) — the Wrapper of gan/wrappers/gan_wrapper.py:14-66: one generator, one
discriminator, tfgan.gan_model + gan_loss with tfgan's DEFAULT losses, i.e. Wasserstein:
    generator  L_G = -mean(D(G(inp)))            discriminator  L_D = mean(D(pool(G(inp)))) - mean(D(real)) + l2 reg
(the constructor's identity_loss_weight / use_identity_loss are stored but never used by define_loss, :46-54 —
reproduced).  swap_inputs=True trains y -> x.  Train ops, Adam(beta1 0.5) and the LR schedule are the standard ones."""
import ctypes

import torch

from hypelcnn_b200 import _native as N
from hypelcnn_b200 import engine as E
from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANTrainer, GANTrainOps, _p, _st
from hypelcnn_b200.gan.wrappers.gan_common import (ValidationHook, adj_shadow_ratio,
                                                   create_inference_for_matrix_input)
from hypelcnn_b200.gan.wrappers.wrapper import InferenceWrapper, Wrapper


class GANTrainer(CycleGANTrainer):
    """Reuses the CycleGAN trainer's buffers and kernels; only G (= gen_x2y) and D_Y are trained."""

    def __init__(self, bands, swap_inputs, discriminator_reg_scale=1e-5, device=None, seed=1234, pool_size=50):
        super().__init__(bands, 0.0, 0.0, False, discriminator_reg_scale, device, seed, pool_size)
        self.use_fused = False       # the fused step kernels are CycleGAN's; the one-directional GAN chains per-op kernels
        self.swap = bool(swap_inputs)

    def _pick(self, images_x, images_y):
        x, y = self._rows(images_x), self._rows(images_y)
        return (y, x) if self.swap else (x, y)  # (generator_inputs, real_data)   (:37-42)

    def _lin(self, a, weight, grad, slot):
        N.check(N.lib().hyp_gan_loss_grad(2, _p(a), None, 0.0, weight / a.numel(), a.numel(), _p(grad), 0,
                                          ctypes.c_void_p(self.loss_acc[slot:].data_ptr()), _st()))

    def generator_gradients(self, images_x, images_y):
        inp, _ = self._pick(images_x, images_y)
        self.gen_grads.zero_()
        self.loss_acc.zero_()
        nets = self._gen_fwd(inp, self.G())
        gen = nets[:, 7, :].contiguous()
        h, d = self._dis_fwd(gen, self.DY())
        g_d = torch.empty_like(d)
        self._lin(d, -1.0, g_d, 1)                                   # wasserstein_generator_loss = -mean(D(G))
        g_gen = self._dis_bwd(gen, h, g_d, self.DY(), None, True)
        self._gen_bwd(nets, g_gen, self.G(), self.gen_grads[:self.ng], need_gin=False)
        self.last = {"generated": gen}
        loss = self.loss_acc.clone()
        loss[0] = loss[1]
        return loss

    def discriminator_gradients(self, images_x, images_y, use_pool=True):
        inp, real = self._pick(images_x, images_y)
        self.dis_grads.zero_()
        self.loss_acc.zero_()
        gen = self._gen_fwd(inp, self.G())[:, 7, :].contiguous()
        if use_pool:
            gen = self.pool_y(gen)
        w, gw = self.DY(), self.dis_grads[:self.nd]
        for data, sign in ((gen, 1.0), (real, -1.0)):                # mean(D(fake)) - mean(D(real))
            h, d = self._dis_fwd(data, w)
            g = torch.empty_like(d)
            self._lin(d, sign, g, 1)
            self._dis_bwd(data, h, g, w, gw, False)
        C = self.C
        for off in (0, C * C + C):
            N.check(N.lib().hyp_gan_l2_regularizer(_p(w[off:]), _p(gw[off:]), C * C, self.reg,
                                                   ctypes.c_void_p(self.loss_acc[2:].data_ptr()), _st()))
        loss = self.loss_acc.clone()
        loss[0] = loss[1] + loss[2]
        return loss


class GANWrapper(Wrapper):

    def __init__(self, identity_loss_weight, use_identity_loss, swap_inputs, generator_fn=None, discriminator_fn=None,
                 discriminator_reg_scale=1e-5) -> None:
        super().__init__()
        self._identity_loss_weight = identity_loss_weight
        self._use_identity_loss = use_identity_loss
        self._swap_inputs = swap_inputs
        self._discriminator_reg_scale = discriminator_reg_scale
        self.trainer = None

    def define_model(self, images_x, images_y):
        if self.trainer is None:
            self.trainer = GANTrainer(images_x.shape[-1], self._swap_inputs, self._discriminator_reg_scale, images_x.device)
        return self.trainer

    def define_loss(self, model):
        return model

    def define_train_ops(self, model, loss, max_number_of_steps, **kwargs):
        return GANTrainOps(model, max_number_of_steps, kwargs["generator_lr"], kwargs["discriminator_lr"])

    def get_train_hooks_fn(self):
        return lambda train_ops: [train_ops.generator_train_op, train_ops.discriminator_train_op]


class GANInferenceWrapper(InferenceWrapper):
    """gan_wrapper.py:69-: one generator serves both inference directions' call signature."""

    def __init__(self, fetch_shadows, shadow_generator_fn=None, trainer=None, bands=None):
        from hypelcnn_b200.gan.shadow_data_models import GeneratorVariables
        self._fetch_shadows = fetch_shadows
        self.generator = trainer.gen_x2y if trainer is not None else GeneratorVariables(bands)

    def construct_inference_graph(self, input_tensor, is_shadow_graph, clip_invalid_values, copy_extra=0):
        return create_inference_for_matrix_input(input_tensor, is_shadow_graph, clip_invalid_values, self.generator, copy_extra)

    def make_inference_graph(self, data_set, is_shadow_graph, clip_invalid_values):
        return None, lambda x: self.construct_inference_graph(x, is_shadow_graph, clip_invalid_values)

    def create_generator_restorer(self):
        return self

    def restore(self, values):
        self.generator.load(values)

    def create_inference_hook(self, data_set, loader, log_dir, neighborhood, shadow_map, shadow_ratio,
                              validation_iteration_count, validation_sample_count):
        """gan_wrapper.py:94-106: one hook; a y2x wrapper validates on shadowed pixels against 1 / shadow_ratio."""
        return ValidationHook(iteration_freq=validation_iteration_count, sample_count=validation_sample_count,
                              log_dir=log_dir, loader=loader, data_set=data_set, neighborhood=neighborhood,
                              shadow_map=shadow_map, shadow_ratio=adj_shadow_ratio(shadow_ratio, self._fetch_shadows),
                              input_tensor=None,
                              infer_model=lambda x: self.construct_inference_graph(x, None, clip_invalid_values=False),
                              fetch_shadows=self._fetch_shadows,
                              name_suffix="deshadowed" if self._fetch_shadows else "shadowed")
