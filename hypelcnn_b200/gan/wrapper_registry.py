"""gan_type -> Wrapper registry (reference: gan/wrapper_registry.py:13-94).  Built here: cycle_gan, gan_x2y, gan_y2x.
cut_x2y / cut_y2x / dcl_gan / dcl_cycle_gan (patch feature discriminator + PatchNCE) are not built yet and raise."""
from hypelcnn_b200.gan.gan_sampling_methods import DummySampler
from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANInferenceWrapper, CycleGANWrapper
from hypelcnn_b200.gan.wrappers.gan_wrapper import GANInferenceWrapper, GANWrapper

_NOT_BUILT = ("cut_x2y", "cut_y2x", "dcl_gan", "dcl_cycle_gan")


def get_sampling_map():
    return {"dummy": DummySampler(element_count=2000, fill_value=0.5, coefficient=2)}


def get_infer_wrapper_dict(bands=64):
    return {"cycle_gan": CycleGANInferenceWrapper(bands=bands),
            "gan_x2y": GANInferenceWrapper(fetch_shadows=False, bands=bands),
            "gan_y2x": GANInferenceWrapper(fetch_shadows=True, bands=bands)}


def get_wrapper_dict(flags):
    reg = getattr(flags, "discriminator_reg_scale", 1e-5)
    return {"cycle_gan": CycleGANWrapper(cycle_consistency_loss_weight=flags.cycle_consistency_loss_weight,
                                         identity_loss_weight=flags.identity_loss_weight,
                                         use_identity_loss=flags.use_identity_loss, discriminator_reg_scale=reg),
            "gan_x2y": GANWrapper(identity_loss_weight=flags.identity_loss_weight,
                                  use_identity_loss=flags.use_identity_loss, swap_inputs=False,
                                  discriminator_reg_scale=reg),
            "gan_y2x": GANWrapper(identity_loss_weight=flags.identity_loss_weight,
                                  use_identity_loss=flags.use_identity_loss, swap_inputs=True,
                                  discriminator_reg_scale=reg)}


def get_wrapper(gan_type, flags):
    if gan_type in _NOT_BUILT:
        raise NotImplementedError(f"gan_type {gan_type!r}: the CUT / DCL wrappers are not built in this round")
    return get_wrapper_dict(flags)[gan_type]
