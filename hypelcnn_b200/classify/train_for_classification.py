"""Classifier training run (reference: classify/train_for_classification.py): importer -> create_graph -> monitored
training loop -> TrainingResult.  ``perform_an_episode`` and ``get_log_suffix`` keep the reference's names and
arguments; ``default_flags`` carries the reference's command-line defaults (hyper-parameter search via optuna is not
part of this engine)."""
import argparse
import json
import os
import time
from types import SimpleNamespace

from numpy import mean, std

from hypelcnn_b200 import parallel
from hypelcnn_b200.classify.monitored_session_runner import (add_classification_summaries, run_monitored_session,
                                                             set_run_seed)
from hypelcnn_b200.common.common_nn_ops import (AugmentationInfo, TrainingResult, create_graph, get_importer_from_name,
                                                get_model_from_name)
from hypelcnn_b200.common.cmd_parser import (add_flags, add_parse_cmds_for_importers, add_parse_cmds_for_loaders,
                                             add_parse_cmds_for_loggers, add_parse_cmds_for_models,
                                             add_parse_cmds_for_trainers, type_ensure_strtobool)
from hypelcnn_b200.common.common_ops import path_leaf, replace_abbrs


APP_FLAGS = (("perform_validation", type_ensure_strtobool, False, "Validate during / after training."),
             ("augment_data_with_rotation", type_ensure_strtobool, False, "Augment with 90-degree rotations."),
             ("augment_data_with_spectral", float, None, "Augment with a random spectral offset of this size."),
             ("augment_data_with_shadow", str, None, "Shadow augmenter: cycle_gan, dcl_gan, dcl_cycle_gan or simple."),
             ("augment_data_with_reflection", type_ensure_strtobool, False, "Augment with reflections."),
             ("augmentation_random_threshold", float, 0.5, "Probability of the shadow augmentation per sample."),
             ("device", str, "gpu", "gpu (this engine has no CPU path)"),
             ("save_checkpoint_steps", int, 2000, "Checkpoint frequency"),
             ("validation_steps", int, 40000, "Validation frequency"),
             ("all_data_shuffle_ratio", float, None, "Unused (kept for the reference's command lines)."),
             ("log_model_params", type_ensure_strtobool, False, "Log variable histograms to TensorBoard."))


def add_parse_cmds_for_app(parser):
    """Reference :123-157."""
    add_flags(parser, APP_FLAGS)


def build_parser():
    parser = argparse.ArgumentParser()
    for add in (add_parse_cmds_for_loaders, add_parse_cmds_for_loggers, add_parse_cmds_for_trainers,
                add_parse_cmds_for_models, add_parse_cmds_for_importers, add_parse_cmds_for_app):
        add(parser)
    return parser


def default_flags(**overrides):
    """The parser's defaults as a flags object (reference: common/cmd_parser.py:15-67 and :123-157 here)."""
    flags = vars(build_parser().parse_known_args([])[0])
    unknown = set(overrides) - set(flags)
    if unknown:
        raise KeyError(f"unknown flags: {sorted(unknown)}")
    flags.update(overrides)
    return SimpleNamespace(**flags)


def _augmentation_info(flags, shadow_dict):
    """:32-41 — the shadow augmenter named by --augment_data_with_shadow (a key of the data set's shadow_creator_dict)
    plus the rotation / reflection / spectral switches."""
    wants_shadow = flags.augment_data_with_shadow is not None
    struct = shadow_dict[flags.augment_data_with_shadow] if (wants_shadow and shadow_dict is not None) else None
    return AugmentationInfo(shadow_struct=struct, perform_shadow_augmentation=wants_shadow,
                            perform_rotation_augmentation=flags.augment_data_with_rotation,
                            perform_reflection_augmentation=flags.augment_data_with_reflection,
                            perform_spectral_augmentation=flags.augment_data_with_spectral,
                            augmentation_random_threshold=flags.augmentation_random_threshold)


def _report(result, with_validation):
    """:98-120.  The reference keeps one-element lists (a left-over of multi-run episodes) and prints mean +- std of
    them; the printed lines are kept, the std of one value is 0."""
    head = f"Validation accuracy={result.validation_accuracy:g}, " if with_validation else ""
    print(f"{head}Testing accuracy={result.test_accuracy:g}, loss={result.loss:.2f}")
    validation = None
    if with_validation:
        validation = mean([result.validation_accuracy])
        print(f"Validation result: ({validation:g}) +- ({std([result.validation_accuracy]):g})")
    print(f"Mean testing accuracy result: ({mean([result.test_accuracy]):g}) +- ({std([result.test_accuracy]):g}), "
          f"Loss result: ({mean([result.loss]):g}) +- ({std([result.loss]):g})")
    return TrainingResult(validation_accuracy=validation, test_accuracy=mean([result.test_accuracy]),
                          loss=mean([result.loss]))


def perform_an_episode(flags, algorithm_params, model, base_log_path):
    """Reference :20-120: read the data set through the importer named by the flags, build the train / test /
    validation branches over one shared model, run the monitored loop, report."""
    print("Args:", json.dumps(vars(flags), indent=3))
    if flags.device == "cpu":
        raise RuntimeError("--device=cpu: this engine has no CPU path (the kernels are sm_100a only)")
    importer = get_importer_from_name(flags.importer_name)
    # under torchrun: one process per GPU, the training split strided over the ranks, one gradient all-reduce per step
    # (hypelcnn_b200/parallel.py); every rank evaluates the whole validation / test lists, rank 0 writes the files.
    # The loaders' random splits must come out identical on every rank before they are strided: one broadcast seed.
    rank, _, world = parallel.init_from_env()
    parallel.sync_split_seed()
    train, test, validation, shadow_dict, class_range, scene_shape, color_list = importer.read_data_set(
        flags.loader_name, flags.path, flags.train_ratio, flags.test_ratio, flags.neighborhood, True)
    augmentation_info = _augmentation_info(flags, shadow_dict)
    train = parallel.shard_training_data(train, rank, world)

    batch_size = algorithm_params["batch_size"]
    required_steps = flags.step if flags.epoch is None else (train.data.shape[0] * flags.epoch) // batch_size
    print(f"Steps: {required_steps:d}, Algorithm Params: {algorithm_params}")

    set_run_seed()
    testing_tensor, training_tensor, validation_tensor = importer.convert_data_to_tensor(test, train, validation,
                                                                                         class_range)
    cross_entropy, learning_rate, testing_nn, training_nn, validation_nn, train_step = create_graph(
        training_tensor.dataset, testing_tensor.dataset, validation_tensor.dataset, class_range, batch_size,
        1000, "/gpu:0", flags.epoch, augmentation_info=augmentation_info, algorithm_params=algorithm_params,
        model=model, create_separate_validation_branch=importer.requires_separate_validation_branch)
    training_nn.data_with_labels, testing_nn.data_with_labels, validation_nn.data_with_labels = train, test, validation
    if world > 1:
        train_step.allreduce = parallel.GradientAllReduce(overlap=True)
    if not flags.perform_validation:
        validation_nn = None

    summaries = add_classification_summaries(cross_entropy, learning_rate, flags.log_model_params, testing_nn,
                                             validation_nn)
    started = time.time()
    result = run_monitored_session(cross_entropy, base_log_path, class_range, flags.save_checkpoint_steps,
                                   flags.validation_steps, train_step, required_steps, augmentation_info,
                                   training_nn, training_tensor, testing_nn, testing_tensor, validation_nn,
                                   validation_tensor, importer, json.dumps(vars(flags), indent=3),
                                   json.dumps(algorithm_params, indent=3), summaries=summaries, is_chief=rank == 0)
    print(f"Done training for {time.time() - started:.3f} sec")
    return _report(result, flags.perform_validation)


def get_log_suffix(flags):
    """Reference :160-180: ``<loader>_<model>_trn<ratio>_<algorithm file stem>_<P>x<P>[_<shadow>_aug<thr>][_spectral<s>]``,
    dots dropped from the numbers, then the abbreviations model -> mdl, dataloader -> ldr, alg_param_ -> p."""
    ratio = flags.train_ratio
    parts = [flags.loader_name.lower(), flags.model_name.lower(),
             "trn" + (f"{int(ratio):d}" if ratio > 1.0 else f"{ratio:.2f}".replace(".", "")),
             os.path.splitext(path_leaf(flags.algorithm_param_path))[0].lower(),
             "{0:d}x{0:d}".format(flags.neighborhood * 2 + 1)]
    if flags.augment_data_with_shadow is not None:
        parts += [str(flags.augment_data_with_shadow), f"aug{flags.augmentation_random_threshold:.2f}".replace(".", "")]
    if flags.augment_data_with_spectral is not None:
        parts.append(f"spectral{flags.augment_data_with_spectral:.3f}".replace(".", ""))
    return replace_abbrs("_".join(parts), {"model": "mdl", "dataloader": "ldr", "alg_param_": "p"})


def run(flags):
    """The non-optuna branch of the reference's main (:213-220)."""
    nn_model = get_model_from_name(flags.model_name)
    if flags.algorithm_param_path is None:
        raise IOError("Algorithm parameter file is not given")
    algorithm_params = json.load(open(flags.algorithm_param_path, "r"))
    algorithm_params["batch_size"] = flags.batch_size
    return perform_an_episode(flags, algorithm_params, nn_model,
                              os.path.join(flags.base_log_path, get_log_suffix(flags)))


def main(argv=None):
    flags, _ = build_parser().parse_known_args(argv)
    print("Running on training mode")
    result = run(flags)
    parallel.finish()
    return result


if __name__ == "__main__":
    main()
