"""Scene classification with a trained model (reference: classify/infer_for_classification.py): every pixel of the
scene (``domain="all"``), the labelled samples (``"sample"``) or the ground truth itself (``"gt"``) as a uint8 class
image + its coloured rendering.

The per-pixel Python generator of the reference (GeneratorImporter -> tf.data.from_generator) becomes one patch-gather
launch per batch followed by the model's eval forward, argmax and a scatter into the class image, all on the device
(common_nn_ops.perform_prediction).  The weights come from a checkpoint written by the training loop
(``model.ckpt-<step>.safetensors``); like the reference's Saver the ``image_gen_net_*`` decoder is not restored.
The class images are returned and written as ``result_raw.tif`` / ``result_colorized.tif`` like the reference's
(utilities/tiff_io.py; no GeoTIFF georeferencing tags — the reference writes none either).
"""
import json
import os
import time

import numpy
import torch

from hypelcnn_b200 import parallel
from hypelcnn_b200.common.common_nn_ops import (ModelInputParams, NNParams, create_colored_image,
                                                create_target_image_via_samples, get_loader_from_name,
                                                get_model_from_name, perform_prediction, simple_nn_iterator)
from hypelcnn_b200.importer.GeneratorImporter import GeneratorDataInfo, GeneratorImporter
from hypelcnn_b200.utilities.tiff_io import imwrite


def create_all_scene_data(scene_shape, data_with_labels_to_copy):
    """Reference :24-35: one target per scene pixel, row by row: (x = column, y = row, class 0)."""
    rows, cols = numpy.mgrid[0:scene_shape[0], 0:scene_shape[1]]
    targets = numpy.stack([cols.reshape(-1), rows.reshape(-1), numpy.zeros(rows.size, dtype=int)], axis=1).astype(int)
    return GeneratorDataInfo(data=None, targets=targets, loader=data_with_labels_to_copy.loader,
                             dataset=data_with_labels_to_copy.dataset)


def create_sample_data(test_data_with_labels, training_data_with_labels, validation_data_with_labels):
    """Reference :38-47: the three labelled target lists stacked (int32), in the order the arguments are given."""
    targets = numpy.vstack([numpy.asarray(d.targets).astype(numpy.int32) for d in
                            (test_data_with_labels, training_data_with_labels, validation_data_with_labels)])
    return GeneratorDataInfo(data=None, targets=targets, loader=test_data_with_labels.loader,
                             dataset=test_data_with_labels.dataset)


def gt_process(flags):
    """Reference :78-85."""
    loader = get_loader_from_name(flags.loader_name, flags.path)
    sample_set = loader.load_samples(0.1, 0)
    data_set = loader.load_data(0, False)
    scene_as_image = create_target_image_via_samples(sample_set, data_set.get_scene_shape())
    return scene_as_image, loader.get_samples_color_list()


def latest_checkpoint(path):
    """``flags.base_log_path`` names a checkpoint file, or a log directory whose newest model.ckpt-<step> is taken."""
    if os.path.isdir(path):
        from hypelcnn_b200.classify.monitored_session_runner import CheckpointSaver
        found = CheckpointSaver(path, None, None).existing()
        if not found:
            raise IOError(f"no model.ckpt-*.safetensors under {path}")
        return found[-1][1]
    return path


def prediction_process(flags, model=None):
    """Reference :88-134.  Note the reference's argument mix-up is kept: read_data_set returns (train, test,
    validation, ...) and ``create_sample_data(training, test, validation)`` receives them in that order although its
    parameters are named (test, training, validation) — only the stacking order of the targets depends on it."""
    data_importer = GeneratorImporter()
    # under torchrun every rank classifies a contiguous slice of the pixel list; the slices meet in merge_class_map.
    # The sample lists are random splits: same seed on every rank first, or the slices would not tile one list.
    rank, _, world = parallel.init_from_env()
    parallel.sync_split_seed()
    training_data_with_labels, test_data_with_labels, validation_data_with_labels, shadow_dict, class_range, \
        scene_shape, color_list = data_importer.read_data_set(flags.loader_name, flags.path, 0.1, 0,
                                                               flags.neighborhood, True)
    if flags.domain == "all":
        validation_data_with_labels = create_all_scene_data(scene_shape, validation_data_with_labels)
    elif flags.domain == "sample":
        validation_data_with_labels = create_sample_data(training_data_with_labels, test_data_with_labels,
                                                         validation_data_with_labels)

    if world > 1:
        validation_data_with_labels = validation_data_with_labels._replace(
            targets=parallel.shard_targets(validation_data_with_labels.targets, rank, world))

    if flags.algorithm_param_path is None:
        raise IOError("Algorithm parameter file is not given")
    algorithm_params = json.load(open(flags.algorithm_param_path, "r"))
    algorithm_params["batch_size"] = flags.batch_size
    nn_model = model if model is not None else get_model_from_name(flags.model_name)

    testing_tensor, training_tensor, validation_tensor = data_importer.convert_data_to_tensor(
        test_data_with_labels, training_data_with_labels, validation_data_with_labels, class_range)
    validation_input_iter = simple_nn_iterator(validation_tensor.dataset, flags.batch_size)

    def predict(images):
        return nn_model.create_tensor_graph(ModelInputParams(x=images, y=None, device_id="/gpu:0", is_training=False),
                                            class_range.stop, algorithm_params).y_conv

    validation_nn_params = NNParams(input_iterator=validation_input_iter,
                                    data_with_labels=validation_data_with_labels, metrics=None, predict_tensor=predict)
    data_set = validation_data_with_labels.dataset
    if nn_model.engine is None:          # build the engine from the patch shape, then restore into it
        nn_model._class_count = class_range.stop
        nn_model.engine_for(data_set.get_data_points(numpy.zeros([1, 2], numpy.int32)), algorithm_params)
    if flags.base_log_path is not None:
        nn_model.engine.load_checkpoint(latest_checkpoint(flags.base_log_path), exclude_prefixes=("image_gen_net_",))

    scene_as_image = torch.full(tuple(scene_shape), 255, dtype=torch.uint8, device=data_set.device)
    data_importer.init_tensors(None, validation_tensor, validation_nn_params)
    perform_prediction(None, validation_nn_params, scene_as_image)
    parallel.merge_class_map(scene_as_image)
    return scene_as_image.cpu().numpy(), color_list


def run(flags, model=None):
    """The reference's main (:50-75) without the argument parsing; returns (class image, coloured image)."""
    start_time = time.time()
    if flags.domain == "all" or flags.domain == "sample":
        scene_as_image, color_list = prediction_process(flags, model)
    elif flags.domain == "gt":
        scene_as_image, color_list = gt_process(flags)
    else:
        raise ValueError(f"Domain flags does not support value:{flags.domain}")
    colored = create_colored_image(scene_as_image, color_list)
    if flags.output_path is not None and parallel.init_from_env()[0] == 0:       # rank 0 writes the merged images
        imwrite(os.path.join(flags.output_path, "result_raw.tif"), scene_as_image)
        imwrite(os.path.join(flags.output_path, "result_colorized.tif"), colored)
    print(f"Done evaluation({time.time() - start_time:.3f} sec)")
    return scene_as_image, colored


def build_parser():
    import argparse
    from hypelcnn_b200.common import cmd_parser as C
    parser = argparse.ArgumentParser()
    for add in (C.add_parse_cmds_for_loaders, C.add_parse_cmds_for_loggers, C.add_parse_cmds_for_trainers,
                C.add_parse_cmds_for_models, C.add_parse_cmds_for_importers):
        add(parser)
    C.add_flags(parser, (("domain", str, "all", "all (whole scene), sample (labelled samples) or gt (ground truth)"),))
    return parser


def main(argv=None):
    flags, _ = build_parser().parse_known_args(argv)
    result = run(flags)
    parallel.finish()
    return result


if __name__ == "__main__":
    main()
