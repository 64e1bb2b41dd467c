"""gan_type -> Wrapper registry (reference: gan/wrapper_registry.py:13-94): cycle_gan, gan_x2y, gan_y2x, cut_x2y,
cut_y2x, dcl_gan, dcl_cycle_gan.  What the reference binds into its model functions with functools.partial
(flags.patches, flags.embedded_feat_size, the two regularisation scales, :31-40) is passed to the wrappers here."""
from hypelcnn_b200.gan.gan_sampling_methods import (DummySampler, NeighborhoodBasedSampler, RandomBasedSampler,
                                                    TargetBasedSampler)
from hypelcnn_b200.gan.wrappers.cut_wrapper import CUTInferenceWrapper, CUTWrapper
from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANInferenceWrapper, CycleGANWrapper
from hypelcnn_b200.gan.wrappers.dcl_cycle_gan_wrapper import DCLCycleGANInferenceWrapper, DCLCycleGANWrapper
from hypelcnn_b200.gan.wrappers.dcl_gan_wrapper import DCLGANInferenceWrapper, DCLGANWrapper
from hypelcnn_b200.gan.wrappers.gan_wrapper import GANInferenceWrapper, GANWrapper


def get_sampling_map():
    """--pairing_method -> sampler, with the reference's constants (gan/wrapper_registry.py:13-18)."""
    return {"target": TargetBasedSampler(margin=5),
            "random": RandomBasedSampler(multiply_shadowed_data=True),
            "neighbour": NeighborhoodBasedSampler(neighborhood_size=20, margin=2),
            "dummy": DummySampler(element_count=2000, fill_value=0.5, coefficient=2)}


_INFER_WRAPPERS = {"cycle_gan": (CycleGANInferenceWrapper, {}),
                   "gan_x2y": (GANInferenceWrapper, {"fetch_shadows": False}),
                   "gan_y2x": (GANInferenceWrapper, {"fetch_shadows": True}),
                   "cut_x2y": (CUTInferenceWrapper, {"fetch_shadows": False}),
                   "cut_y2x": (CUTInferenceWrapper, {"fetch_shadows": True}),
                   "dcl_gan": (DCLGANInferenceWrapper, {}),
                   "dcl_cycle_gan": (DCLCycleGANInferenceWrapper, {})}


def get_infer_wrapper(gan_type, bands=None, trainer=None):
    """One inference wrapper: over fresh (zero) generator variables of ``bands`` bands to be restored from a
    checkpoint, or bound to a live ``trainer``'s generators — what sharing the "Generator" variable scope between the
    train and the validation graph does in the reference (gan/gan_train_for_shadow.py:268-275)."""
    cls, kwargs = _INFER_WRAPPERS[gan_type]
    return cls(trainer=trainer, bands=bands, **kwargs)


def get_infer_wrapper_dict(bands=64):
    """Reference: gan/wrapper_registry.py:21-32."""
    return {gan_type: get_infer_wrapper(gan_type, bands=bands) for gan_type in _INFER_WRAPPERS}


def get_wrapper_dict(flags):
    reg = getattr(flags, "discriminator_reg_scale", 1e-5)
    contrastive = dict(nce_loss_weight=getattr(flags, "nce_loss_weight", 10.0),
                       identity_loss_weight=flags.identity_loss_weight, use_identity_loss=flags.use_identity_loss,
                       tau=getattr(flags, "tau", 0.07), batch_size=getattr(flags, "batch_size", None),
                       patches=getattr(flags, "patches", 6), embedded_feat_size=getattr(flags, "embedded_feat_size", 2),
                       discriminator_reg_scale=reg, gen_disc_reg_scale=getattr(flags, "gen_disc_reg_scale", 1e-4))
    return {"cycle_gan": CycleGANWrapper(cycle_consistency_loss_weight=flags.cycle_consistency_loss_weight,
                                         identity_loss_weight=flags.identity_loss_weight,
                                         use_identity_loss=flags.use_identity_loss, discriminator_reg_scale=reg),
            "gan_x2y": GANWrapper(identity_loss_weight=flags.identity_loss_weight,
                                  use_identity_loss=flags.use_identity_loss, swap_inputs=False,
                                  discriminator_reg_scale=reg),
            "gan_y2x": GANWrapper(identity_loss_weight=flags.identity_loss_weight,
                                  use_identity_loss=flags.use_identity_loss, swap_inputs=True,
                                  discriminator_reg_scale=reg),
            "cut_x2y": CUTWrapper(swap_inputs=False, **contrastive),
            "cut_y2x": CUTWrapper(swap_inputs=True, **contrastive),
            "dcl_gan": DCLGANWrapper(**contrastive),
            "dcl_cycle_gan": DCLCycleGANWrapper(
                cycle_consistency_loss_weight=flags.cycle_consistency_loss_weight, **contrastive)}


def get_wrapper(gan_type, flags):
    return get_wrapper_dict(flags)[gan_type]
