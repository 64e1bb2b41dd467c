"""DCL-CycleGAN — the Wrapper of gan/wrappers/dcl_cycle_gan_wrapper.py:155-211.  The reference builds the DCLGAN model
plus the two cycle reconstructions (dcl_cycle_gan_model, :16-46) and computes a cycle-consistency term, but adds it to
the generator losses through `_replace` calls whose results are discarded (:149-150): the term never reaches a train
op, and training is DCLGAN's exactly (define_train_ops delegates to DCLGANWrapper.base_trainops_method, :199-200).
This wrapper therefore IS the DCLGAN wrapper with the extra constructor argument kept, and `reconstructions()` for the
two tensors the reference's model tuple carries."""
from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANInferenceWrapper
from hypelcnn_b200.gan.wrappers.dcl_gan_wrapper import DCLGANWrapper


class DCLCycleGANWrapper(DCLGANWrapper):

    def __init__(self, nce_loss_weight, identity_loss_weight, cycle_consistency_loss_weight, use_identity_loss, tau,
                 batch_size, generator_fn=None, discriminator_fn=None, feat_discriminator_fn=None, **model_args) -> None:
        super().__init__(nce_loss_weight, identity_loss_weight, use_identity_loss, tau, batch_size, generator_fn,
                         discriminator_fn, feat_discriminator_fn, **model_args)
        self._cycle_consistency_loss_weight = cycle_consistency_loss_weight   # inert in the reference, see above

    def reconstructions(self, images_x, images_y):
        """(reconstructed_x, reconstructed_y) = (G_y2x(G_x2y(x)), G_x2y(G_y2x(y)))  (:33-38), [B,C] each."""
        t = self.trainer
        x, y = t.model_x2y._rows(images_x), t.model_x2y._rows(images_y)
        gx = t.model_x2y._gen_fwd(x, t.model_x2y.gen_params)[:, 7, :].contiguous()
        fy = t.model_y2x._gen_fwd(y, t.model_y2x.gen_params)[:, 7, :].contiguous()
        return (t.model_y2x._gen_fwd(gx, t.model_y2x.gen_params)[:, 7, :].contiguous(),
                t.model_x2y._gen_fwd(fy, t.model_x2y.gen_params)[:, 7, :].contiguous())


class DCLCycleGANInferenceWrapper(CycleGANInferenceWrapper):
    """dcl_cycle_gan_wrapper.py:209-211."""
