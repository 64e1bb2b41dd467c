"""GULFPORTALT-shaped synthetic scene (BASELINE configs[4]): the GULFPORT scene plus the shadow augmenters the
reference's GULFPORTALTDataLoader attaches to its data set (loader/GULFPORTALTDataLoader.py:85-95): ``cycle_gan``,
``dcl_gan``, ``dcl_cycle_gan`` (generators restored from ``<base>/shadow_gen_model/<type>/model.ckpt-3000``) and
``simple`` (per-band shadow ratio)."""
import os

from hypelcnn_b200.loader.SyntheticGULFPORTDataLoader import SyntheticGULFPORTDataLoader


class SyntheticGULFPORTALTDataLoader(SyntheticGULFPORTDataLoader):
    GAN_CHECKPOINTS = {"cycle_gan": "shadow_gen_model/cycle_gan/model.ckpt-3000",
                       "dcl_gan": "shadow_gen_model/dcl_gan/model.ckpt-3000",
                       "dcl_cycle_gan": "shadow_gen_model/dcl_cycle_gan/v1/model.ckpt-3000"}

    def __init__(self, base_dir):
        """``base_dir``: ``synthetic:H=..,W=..,samples=..[,models=<directory>]`` — ``models`` is where the generator
        checkpoints live (default: the current directory)."""
        super().__init__(base_dir)
        self.models_dir = os.getcwd()
        if isinstance(base_dir, str) and base_dir.startswith("synthetic:"):
            for kv in base_dir[len("synthetic:"):].split(","):
                if kv.startswith("models="):
                    self.models_dir = kv[len("models="):]

    def get_model_base_dir(self):
        return os.path.join(self.models_dir, "")

    def load_data(self, neighborhood, normalize):
        from hypelcnn_b200.gan.gan_utilities import create_gan_struct, create_simple_shadow_struct
        from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANInferenceWrapper
        data_set = super().load_data(neighborhood, normalize)
        _, shadow_ratio = self.load_shadow_map(neighborhood, data_set)
        data_set.shadow_creator_dict = {
            name: create_gan_struct(CycleGANInferenceWrapper(bands=self.BANDS), self.get_model_base_dir(), path)
            for name, path in self.GAN_CHECKPOINTS.items()}
        data_set.shadow_creator_dict["simple"] = create_simple_shadow_struct(shadow_ratio)
        return data_set
