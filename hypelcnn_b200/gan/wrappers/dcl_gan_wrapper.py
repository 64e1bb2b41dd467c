"""DCLGAN (dual contrastive learning GAN) — the Wrapper of gan/wrappers/dcl_gan_wrapper.py:232-319: two independent CUT
models, ModelX2Y (x -> y) and ModelY2X (y -> x), each with its own generator, discriminator and feature discriminator
(dcl_gan_model, :28-68).

Two properties of the reference are reproduced on purpose:
  * dcl_gan_loss's cross-sum of the generator losses is computed into `_replace` results that are discarded
    (:189-190), so each CUT model is trained on its own CUT loss only;
  * both models' train ops are built from the SAME three AdamOptimizer objects (:285-309), whose beta-power
    accumulators therefore advance twice per iteration [TF-lib: non-slot variables are per optimizer and graph] —
    here both CUTTrainers share one `clock`.
One train iteration = global_step += 1, then x2y's generator / discriminator / feature-discriminator steps, then
y2x's (get_sequential_train_hooks_dclgan, :213-229)."""
from collections import namedtuple

from hypelcnn_b200.gan.wrappers.cut_wrapper import CUTTrainer, CUTTrainOps
from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANInferenceWrapper
from hypelcnn_b200.gan.wrappers.wrapper import Wrapper

DCLGANModel = namedtuple("DCLGANModel", ["model_x2y", "model_y2x"])
DCLGANLoss = namedtuple("DCLGANLoss", ["loss_x2y", "loss_y2x"])


class DCLGANTrainer:
    def __init__(self, bands, nce_loss_weight=10.0, identity_loss_weight=0.5, use_identity_loss=True, tau=0.07,
                 patches=6, embedded_feat_size=2, discriminator_reg_scale=1e-5, gen_disc_reg_scale=1e-4, device=None,
                 seed=1234, fused_xent_grad=True):
        self.clock = {"global_step": 0, "gen": 0, "dis": 0, "feat": 0}
        kw = dict(nce_loss_weight=nce_loss_weight, identity_loss_weight=identity_loss_weight,
                  use_identity_loss=use_identity_loss, tau=tau, patches=patches, embedded_feat_size=embedded_feat_size,
                  discriminator_reg_scale=discriminator_reg_scale, gen_disc_reg_scale=gen_disc_reg_scale, device=device,
                  fused_xent_grad=fused_xent_grad, clock=self.clock)
        self.model_x2y = CUTTrainer(bands, swap_inputs=False, seed=seed, **kw)       # generator_inputs = x (:41-47)
        self.model_y2x = CUTTrainer(bands, swap_inputs=True, seed=seed + 1, **kw)    # generator_inputs = y (:56-62)

    # what CycleGANInferenceWrapper reads
    @property
    def gen_x2y(self):
        return self.model_x2y.generator

    @property
    def gen_y2x(self):
        return self.model_y2x.generator

    def variables(self):
        out = {f"ModelX2Y/{k}": v for k, v in self.model_x2y.variables().items()}
        out.update({f"ModelY2X/{k}": v for k, v in self.model_y2x.variables().items()})
        return out


class DCLGANTrainOps:
    """DCLGANTrainOps (:195-210): x2y_ops, y2x_ops, global_step_inc_op (x2y's), train_hooks."""

    def __init__(self, trainer, max_number_of_steps, generator_lr, discriminator_lr, gen_discriminator_lr):
        self.trainer = trainer
        self.x2y_ops = CUTTrainOps(trainer.model_x2y, max_number_of_steps, generator_lr, discriminator_lr, gen_discriminator_lr)
        self.y2x_ops = CUTTrainOps(trainer.model_y2x, max_number_of_steps, generator_lr, discriminator_lr, gen_discriminator_lr)
        self.train_hooks = []

    def global_step_inc_op(self):
        return self.x2y_ops.global_step_inc_op()     # one shared global step (clock)

    def train_iteration(self, images_x, images_y):
        self.global_step_inc_op()
        return self.x2y_ops.run_sequential(images_x, images_y) + self.y2x_ops.run_sequential(images_x, images_y)


class DCLGANWrapper(Wrapper):

    def __init__(self, nce_loss_weight, identity_loss_weight, use_identity_loss, tau, batch_size,
                 generator_fn=None, discriminator_fn=None, feat_discriminator_fn=None, patches=6, embedded_feat_size=2,
                 discriminator_reg_scale=1e-5, gen_disc_reg_scale=1e-4) -> None:
        super().__init__()
        self._nce_loss_weight = nce_loss_weight
        self._identity_loss_weight = 0.0 if not use_identity_loss else identity_loss_weight
        self._tau = tau
        self._batch_size = batch_size
        self._model_args = dict(patches=patches, embedded_feat_size=embedded_feat_size,
                                discriminator_reg_scale=discriminator_reg_scale, gen_disc_reg_scale=gen_disc_reg_scale)
        self.trainer = None

    def define_model(self, images_x, images_y):
        if self.trainer is None:
            self.trainer = DCLGANTrainer(images_x.shape[-1], self._nce_loss_weight, self._identity_loss_weight, True,
                                         self._tau, device=images_x.device, **self._model_args)
        return DCLGANModel(self.trainer.model_x2y, self.trainer.model_y2x)

    def define_loss(self, model):
        return DCLGANLoss(model.model_x2y, model.model_y2x)

    def define_train_ops(self, model, loss, max_number_of_steps, **kwargs):
        return self.base_trainops_method(self.trainer, max_number_of_steps, kwargs)

    @staticmethod
    def base_trainops_method(trainer, max_number_of_steps, kwargs):
        return DCLGANTrainOps(trainer, max_number_of_steps, kwargs["generator_lr"], kwargs["discriminator_lr"],
                              kwargs["gen_discriminator_lr"])

    def get_train_hooks_fn(self):
        def get_hooks(train_ops):
            ops = []
            for o in (train_ops.x2y_ops, train_ops.y2x_ops):
                ops += [o.generator_train_op, o.discriminator_train_op, o.gen_discriminator_train_op]
            return ops
        return get_hooks


class DCLGANInferenceWrapper(CycleGANInferenceWrapper):
    """dcl_gan_wrapper.py:322-324: the two CUT generators serve the forward (shadow) / backward inference graphs."""
