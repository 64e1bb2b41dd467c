"""Validation of trained shadow generators on scene samples (reference: gan/gan_infer_for_shadow.py): restore the
generator(s) from a checkpoint, push ``number_of_samples`` random scene spectra through them, print and log the
band-ratio statistics (the inference wrappers' validation hooks with frequency 0 = "run now")."""
import argparse
from types import SimpleNamespace

import numpy

from hypelcnn_b200.common.cmd_parser import add_flags, add_parse_cmds_for_loaders, add_parse_cmds_for_loggers
from hypelcnn_b200.common.common_nn_ops import get_loader_from_name
from hypelcnn_b200.gan.gan_utilities import read_generator_checkpoint
from hypelcnn_b200.gan.wrapper_registry import get_infer_wrapper

APP_FLAGS = (("number_of_samples", int, 6000, "Number of samples."),
             ("gan_type", str, "cycle_gan", "cycle_gan, gan_x2y, gan_y2x, cut_x2y, cut_y2x, dcl_gan, dcl_cycle_gan"))


def add_parse_cmds_for_app(parser):
    add_flags(parser, APP_FLAGS)


def restore_generators(inference_wrapper, checkpoint_path):
    """The reference's ``create_generator_restorer().restore(sess, path)``; path = a ``model.ckpt-<step>[.npz]`` written
    by gan_train_for_shadow.run_session."""
    forward, backward = read_generator_checkpoint(checkpoint_path)
    restorer = inference_wrapper.create_generator_restorer()
    if hasattr(inference_wrapper, "backward_generator"):
        restorer.restore(forward, backward)
    else:
        restorer.restore(forward)
    return inference_wrapper


def run(flags):
    numpy.set_printoptions(precision=5, suppress=True)
    loader = get_loader_from_name(flags.loader_name, flags.path)
    data_set = loader.load_data(flags.neighborhood, True)
    shadow_map, shadow_ratio = loader.load_shadow_map(flags.neighborhood, data_set)
    wrapper = restore_generators(get_infer_wrapper(flags.gan_type, bands=data_set.get_casi_band_count()),
                                 flags.base_log_path)
    hook = wrapper.create_inference_hook(data_set=data_set, loader=loader, log_dir=flags.output_path,
                                         neighborhood=flags.neighborhood, shadow_map=shadow_map,
                                         shadow_ratio=shadow_ratio, validation_iteration_count=0,
                                         validation_sample_count=flags.number_of_samples)
    hook.after_create_session(None, None)
    hook.after_run(SimpleNamespace(global_step=0), None)
    return hook.get_best_mean_div()


def main(argv=None):
    parser = argparse.ArgumentParser()
    add_parse_cmds_for_loaders(parser)
    add_parse_cmds_for_loggers(parser)
    add_parse_cmds_for_app(parser)
    flags, _ = parser.parse_known_args(argv)
    print("Mean divergence:", run(flags))


if __name__ == "__main__":
    main()
