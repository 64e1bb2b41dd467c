"""GAN training run (reference: gan/gan_train_for_shadow.py): pair the scene's normal / shadowed spectra, feed them
batch by batch to the train ops of the chosen wrapper, validate the generators every ``validation_steps`` iterations.

Same function names and argument meaning as the reference where a function exists there (``load_op``,
``perform_shadow_augmentation_random``, ``get_log_suffix``, ``gan_train``, ``run_session``).  What differs is where the
work happens: the paired matrices are gathered once into HBM (gan_sampling_methods), ``load_op`` returns an iterator
whose batches are index selections of those resident matrices (the reference re-feeds them through a tf.data pipeline
pinned to the CPU, :147-169), the regularisation-support augmentation is applied to the whole batch at once with one
uniform draw per sample, and ``gan_train`` is a plain loop instead of a MonitoredTrainingSession.
Hyper-parameter search (optuna, :215-231) and parameter-server flags are not part of this engine.
"""
import argparse
import json
import os
from collections.abc import Sequence
from types import SimpleNamespace

import numpy
import torch

from hypelcnn_b200.common.cmd_parser import (add_flags, add_parse_cmds_for_json_loader, add_parse_cmds_for_loaders,
                                             add_parse_cmds_for_loggers, add_parse_cmds_for_trainers,
                                             type_ensure_strtobool)
from hypelcnn_b200 import parallel
from hypelcnn_b200.common.common_ops import replace_abbrs
from hypelcnn_b200.gan.wrapper_registry import get_infer_wrapper, get_sampling_map, get_wrapper_dict
from hypelcnn_b200.gan.wrappers.gan_common import read_hsi_data


APP_FLAGS = (("gan_type", str, "cycle_gan", "cycle_gan, gan_x2y, gan_y2x, cut_x2y, cut_y2x, dcl_gan, dcl_cycle_gan"),
             ("use_identity_loss", type_ensure_strtobool, True, "Add the identity loss to the generator objective."),
             ("identity_loss_weight", float, 0.5, "Weight of the identity loss."),
             ("regularization_support_rate", float, 0.0, "Rate of ratio-made pairs mixed into the batches."),
             ("cycle_consistency_loss_weight", float, 10.0, "Weight of the cycle-consistency loss."),
             ("nce_loss_weight", float, 10.0, "Weight of the PatchNCE loss."),
             ("tau", float, 0.07, "PatchNCE temperature."),
             ("patches", int, 6, "Band slices of the feature discriminator."),
             ("embedded_feat_size", int, 2, "Embedding size of the feature discriminator."),
             ("validation_steps", int, 1000, "Validation / checkpoint frequency."),
             ("validation_sample_count", int, 300, "Validation samples."),
             ("generator_lr", float, 0.0002, "Generator learning rate."),
             ("discriminator_lr", float, 0.0001, "Discriminator learning rate."),
             ("gen_discriminator_lr", float, 0.0001, "Feature-discriminator learning rate."),
             ("discriminator_reg_scale", float, 0.00001, "L2 scale of the discriminator."),
             ("gen_disc_reg_scale", float, 0.0001, "L2 scale of the feature discriminator."),
             ("pairing_method", str, "random", "target, random, neighbour or dummy."))


def add_parse_cmds_for_app(parser):
    """Reference :28-77 (without the parameter-server flags master / ps_tasks / task)."""
    add_flags(parser, APP_FLAGS)


def build_parser():
    parser = argparse.ArgumentParser()
    for add in (add_parse_cmds_for_loaders, add_parse_cmds_for_loggers, add_parse_cmds_for_trainers,
                add_parse_cmds_for_json_loader, add_parse_cmds_for_app):
        add(parser)
    return parser


def default_flags(**overrides):
    """The parser's defaults as the flags object ``run_session`` / ``get_wrapper_dict`` read."""
    flags = vars(build_parser().parse_known_args([])[0])
    unknown = set(overrides) - set(flags)
    if unknown:
        raise KeyError(f"unknown flags: {sorted(unknown)}")
    flags.update(overrides)
    return SimpleNamespace(**flags)


def update_flags_from_json(flags, flag_config_file):
    """Reference :308-314: a JSON file overrides the command line."""
    print("Updating flags from json file,", flag_config_file)
    merged = dict(vars(flags))
    merged.update(json.load(open(flag_config_file, "r")))
    return SimpleNamespace(**merged)


def perform_shadow_augmentation_random(normal_images, shadow_images, shadow_ratio, reg_support_rate, generator=None):
    """Reference :172-184, for a batch ``[B,1,1,C]`` (or ``[B,C]``): per sample, with probability governed by
    ``u < reg_support_rate``, u ~ U(0.01, 0.99), the normal spectrum is replaced by ``shadow * shadow_ratio``, and —
    with a second, independent draw — the shadowed spectrum by ``normal_rand / shadow_ratio`` (the already
    replaced normal spectrum, as in the reference).  Rates <= 0.01 never fire: the inputs are returned as they are."""
    if reg_support_rate <= 0.01:
        return normal_images, shadow_images
    batch = normal_images.shape[0]
    ratio = torch.as_tensor(shadow_ratio, dtype=normal_images.dtype, device=normal_images.device)
    draws = torch.empty(2, batch, device=normal_images.device).uniform_(0.01, 0.99, generator=generator)
    pick = (draws < reg_support_rate).reshape(2, batch, *([1] * (normal_images.dim() - 1)))
    normal_images_rand = torch.where(pick[0], shadow_images * ratio, normal_images)
    shadow_images_rand = torch.where(pick[1], normal_images_rand / ratio, shadow_images)
    return normal_images_rand, shadow_images_rand


class PairIterator:
    """What the reference's ``load_op`` builds as tf.data (:147-169): ``shuffle_and_repeat(count=epoch)`` over the
    paired rows, the random augmentation, ``batch(batch_size, drop_remainder=True)``.  As there, batching happens after
    the repeat, so a batch may straddle two passes and the stream ends after ``epoch * N // batch_size`` batches
    (``StopIteration`` = tf's OutOfRangeError, which ends the training loop).  Each pass is a full permutation drawn on
    the device (tf shuffles within a 10 000-element window; for the reference's pair counts that is the whole set).
    The matrices stay where they are (HBM); a batch is one index_select per side."""

    def __init__(self, normal_data, shadow_data, batch_size, epoch, shadow_ratio, reg_support_rate, seed=1234):
        self.normal_data = torch.as_tensor(normal_data)
        self.shadow_data = torch.as_tensor(shadow_data).to(self.normal_data.device)
        if self.normal_data.shape[0] != self.shadow_data.shape[0]:
            raise ValueError("normal and shadow matrices must pair row by row")
        self.batch_size, self.epoch = int(batch_size), int(epoch)
        self.shadow_ratio, self.reg_support_rate = shadow_ratio, reg_support_rate
        self._generator = torch.Generator(device=self.normal_data.device)
        self._generator.manual_seed(seed)
        self._pending = torch.empty(0, dtype=torch.int64, device=self.normal_data.device)
        self._passes_started = 0
        self.batches_served = 0

    @property
    def initializer(self):
        return None     # nothing to feed: the data is resident (the reference's InitializerHook feeds placeholders)

    def __iter__(self):
        return self

    def __next__(self):
        return self.get_next()

    def get_next(self):
        rows = self.normal_data.shape[0]
        while self._pending.numel() < self.batch_size:
            if self._passes_started >= self.epoch or rows == 0:
                raise StopIteration
            self._passes_started += 1
            order = torch.randperm(rows, generator=self._generator, device=self.normal_data.device)
            self._pending = torch.cat([self._pending, order])
        take, self._pending = self._pending[:self.batch_size], self._pending[self.batch_size:]
        self.batches_served += 1
        return perform_shadow_augmentation_random(self.normal_data.index_select(0, take),
                                                  self.shadow_data.index_select(0, take),
                                                  self.shadow_ratio, self.reg_support_rate, self._generator)


def shard_pair_iterator(iterator, rank, world):
    """Data-parallel GAN training: rank's strided share of the paired rows, every rank the same number of rows (the
    remainder is dropped) so that all ranks run the same number of iterations and meet in every all-reduce; each
    rank shuffles with its own seed."""
    if world <= 1:
        return iterator
    rows = (iterator.normal_data.shape[0] // world) * world
    share = slice(rank, rows, world)
    return PairIterator(iterator.normal_data[share], iterator.shadow_data[share], iterator.batch_size, iterator.epoch,
                        iterator.shadow_ratio, iterator.reg_support_rate, seed=1234 + rank)


def _trainers_of(wrapper):
    trainer = wrapper.trainer
    return [trainer.model_x2y, trainer.model_y2x] if hasattr(trainer, "model_x2y") else [trainer]


def load_op(batch_size, iteration_count, loader, data_set, shadow_map, shadow_ratio, reg_support_rate, pairing_method,
            device=None):
    """Reference :147-169.  Pairs via the registry's sampler, HSI bands only; ``epoch = iteration_count * batch_size
    // pairs`` passes over them."""
    normal_data_as_matrix, shadow_data_as_matrix = read_hsi_data(loader, data_set, shadow_map, pairing_method,
                                                                 get_sampling_map())
    normal = torch.as_tensor(normal_data_as_matrix)
    if device is None:
        device = normal.device if normal.is_cuda else torch.device("cuda", torch.cuda.current_device())
    normal = normal.to(device).contiguous()
    shadow = torch.as_tensor(shadow_data_as_matrix).to(device).contiguous()
    epoch = (iteration_count * batch_size) // normal.shape[0]
    return PairIterator(normal, shadow, batch_size, epoch, shadow_ratio, reg_support_rate)


def get_log_suffix(flags):
    """Reference :187-200 (including its formatting of the boolean ``use_identity_loss`` as ``1.00`` -> ``idnty100``)."""
    abbreviations = {"dataloader": "ldr"}
    patch_size = (flags.neighborhood * 2) + 1
    suffix = f"{flags.loader_name.lower():s}_{flags.gan_type.lower():s}_" \
             f"{patch_size:d}x{patch_size:d}_" \
             f"regsup{flags.regularization_support_rate:.2f}_" \
             f"batch{flags.batch_size:d}".replace(".", "")
    if flags.use_identity_loss is True:
        suffix = suffix + f"_idnty{flags.use_identity_loss:.2f}".replace(".", "")
    return replace_abbrs(suffix, abbreviations)


class RunContext:
    """What hooks see after every iteration: the global step just completed and the last losses."""

    def __init__(self):
        self.global_step, self.results = 0, None


def gan_train(train_ops, input_iterator, logdir, get_hooks_fn, hooks=None, num_steps=None, save_checkpoint_steps=None,
              saver=None):
    """Reference :80-144.  One iteration = ``global_step_inc_op`` + the wrapper's sequential train ops on the next
    batch (gan/wrappers: generator step, discriminator step[, feature-discriminator step]), then every hook's
    ``after_run``.  Stops after ``num_steps`` iterations (StopAtStepHook) or when the input is exhausted.
    ``saver(global_step)`` is called every ``save_checkpoint_steps`` iterations.  Returns the last global step."""
    step_ops = get_hooks_fn(train_ops)
    hooks = [h for h in (hooks or []) if h is not None]
    for hook in hooks:
        hook.after_create_session(None, None)
    context, gstep, done = RunContext(), None, 0
    while num_steps is None or done < num_steps:
        try:
            images_x, images_y = input_iterator.get_next()
        except StopIteration:
            break
        gstep = train_ops.global_step_inc_op()
        context.results = [op(images_x, images_y) for op in step_ops]
        context.global_step = gstep
        done += 1
        for hook in hooks:
            hook.after_run(context, None)
        if saver is not None and save_checkpoint_steps and gstep % save_checkpoint_steps == 0:
            saver(gstep)
    # CheckpointSaverHook.end: MonitoredTrainingSession writes the final state whatever the step count
    if saver is not None and gstep is not None and not (save_checkpoint_steps and gstep % save_checkpoint_steps == 0):
        saver(gstep)
    return gstep


_STATE_TENSORS = ("gen_params", "dis_params", "feat_params", "gen_m", "gen_v", "dis_m", "dis_v")
_STATE_COUNTERS = ("gen_steps", "dis_steps", "global_step")


def _named_trainers(trainer):
    if hasattr(trainer, "model_x2y"):
        return [("ModelX2Y/", trainer.model_x2y), ("ModelY2X/", trainer.model_y2x)]
    return [("", trainer)]


def trainer_state(trainer):
    """Everything a restarted run needs beyond the generator weights (MonitoredTrainingSession(checkpoint_dir=log_dir)
    saves every global variable, reference :123-141): all flat parameter buffers — discriminators and feature
    discriminators included —, the Adam moments, the optimizers' step counts and the global step."""
    out = {}
    for scope, t in _named_trainers(trainer):
        for name in _STATE_TENSORS:
            if hasattr(t, name):
                out[f"{scope}train_state/{name}"] = getattr(t, name).detach().cpu().numpy()
        for key, (m, v) in getattr(t, "slots", {}).items():
            out[f"{scope}train_state/{key}_m"] = m.detach().cpu().numpy()
            out[f"{scope}train_state/{key}_v"] = v.detach().cpu().numpy()
        for name in _STATE_COUNTERS:
            if hasattr(t, name):
                out[f"{scope}train_state/{name}"] = numpy.int64(getattr(t, name))
        for key, value in getattr(t, "clock", {}).items():
            out[f"{scope}train_state/clock_{key}"] = numpy.int64(value)
    return out


def restore_trainer_state(trainer, values):
    """Inverse of trainer_state; buffers are overwritten in place (the generator objects hold views of them)."""
    for scope, t in _named_trainers(trainer):
        def get(name):
            return values.get(f"{scope}train_state/{name}")
        for name in _STATE_TENSORS:
            if hasattr(t, name) and get(name) is not None:
                getattr(t, name).copy_(torch.as_tensor(numpy.asarray(get(name))))
        for key, (m, v) in getattr(t, "slots", {}).items():
            if get(f"{key}_m") is not None:
                m.copy_(torch.as_tensor(numpy.asarray(get(f"{key}_m"))))
                v.copy_(torch.as_tensor(numpy.asarray(get(f"{key}_v"))))
        for name in _STATE_COUNTERS:
            if hasattr(t, name) and get(name) is not None:
                setattr(t, name, int(get(name)))
        for key in list(getattr(t, "clock", {})):
            if get(f"clock_{key}") is not None:
                t.clock[key] = int(get(f"clock_{key}"))


def latest_checkpoint(log_dir):
    """(step, path) of the newest model.ckpt-N.npz in log_dir, or (None, None)."""
    best = (None, None)
    if os.path.isdir(log_dir):
        for name in os.listdir(log_dir):
            if name.startswith("model.ckpt-") and name.endswith(".npz"):
                try:
                    step = int(name[len("model.ckpt-"):-len(".npz")])
                except ValueError:
                    continue
                if best[0] is None or step > best[0]:
                    best = (step, os.path.join(log_dir, name))
    return best


def _generator_exports(inference_wrapper):
    gens = {}
    for scope, attr in (("ModelX2Y", "forward_generator"), ("ModelY2X", "backward_generator"), ("Model", "generator")):
        variables = getattr(inference_wrapper, attr, None)
        if variables is not None:
            gens.update({f"{scope}/Generator/{name}": value for name, value in variables.export().items()})
    return gens


def run_session(params, base_log_path, loader=None):
    """Reference :234-305.  ``params``: the flags as a dict (``vars(default_flags(...))``).  The loader is resolved
    by name like the reference does unless one is handed in.  Returns ``[best upper divergence, best mean
    divergence]`` (the maximum over the two validation directions where the wrapper validates both)."""
    from hypelcnn_b200.common.common_nn_ops import get_loader_from_name
    flags = SimpleNamespace(**params)
    print("Args:", json.dumps(vars(flags), indent=3, default=str))
    log_dir = f"{base_log_path}_{get_log_suffix(flags)}"
    os.makedirs(log_dir, exist_ok=True)

    validation_iteration_count = flags.validation_steps
    validation_sample_count = flags.validation_sample_count
    neighborhood = 0

    # under torchrun: the pairs are strided over the ranks, every train op all-reduces its gradient buffer once; the
    # samplers draw from the global generators, so every rank seeds them alike before the pair list is built.  First
    # thing of all: it selects this rank's GPU, and the scene must be loaded onto THAT device.
    rank, _, world = parallel.init_from_env()
    parallel.sync_split_seed()
    if loader is None:
        loader = get_loader_from_name(flags.loader_name, flags.path)
    data_set = loader.load_data(neighborhood, True)
    shadow_map, shadow_ratio = loader.load_shadow_map(neighborhood, data_set)
    input_iterator = load_op(flags.batch_size, flags.step, loader, data_set, shadow_map, shadow_ratio,
                             flags.regularization_support_rate, flags.pairing_method)
    input_iterator = shard_pair_iterator(input_iterator, rank, world)
    wrapper = get_wrapper_dict(flags)[flags.gan_type]
    the_gan_model = wrapper.define_model(input_iterator.normal_data[:flags.batch_size],
                                         input_iterator.shadow_data[:flags.batch_size])
    the_gan_loss = wrapper.define_loss(the_gan_model)
    inference_wrapper = get_infer_wrapper(flags.gan_type, trainer=wrapper.trainer)
    peer_validation_hook = inference_wrapper.create_inference_hook(
        data_set, loader, log_dir, neighborhood, shadow_map, shadow_ratio,
        validation_iteration_count, validation_sample_count)
    train_ops = wrapper.define_train_ops(the_gan_model, the_gan_loss, max_number_of_steps=flags.step,
                                         generator_lr=flags.generator_lr, discriminator_lr=flags.discriminator_lr,
                                         gen_discriminator_lr=flags.gen_discriminator_lr)

    if world > 1:
        for trainer in _trainers_of(wrapper):
            trainer.allreduce = parallel.GradientAllReduce()
    is_chief = rank == 0                        # validation, summaries and checkpoints are the chief's
    if is_chief:
        from hypelcnn_b200.classify.summaries import ClassificationSummaryWriter
        writer = ClassificationSummaryWriter(log_dir)
        writer.add_text("flags", json.dumps(vars(flags), indent=3, default=str), 0)  # TextSummaryAtStartHook
        writer.close()

    def saver(global_step):
        numpy.savez(os.path.join(log_dir, f"model.ckpt-{global_step}.npz"), global_step=global_step,
                    **_generator_exports(inference_wrapper), **trainer_state(wrapper.trainer))

    # a log dir that already holds a checkpoint continues from it (every rank restores the same file)
    restored_step, restored_path = latest_checkpoint(log_dir)
    if restored_path is not None:
        with numpy.load(restored_path) as values:
            restore_trainer_state(wrapper.trainer, values)
        print(f"Restored {restored_path} (global step {restored_step})")

    gan_train(train_ops, input_iterator, log_dir, get_hooks_fn=wrapper.get_train_hooks_fn(),
              hooks=[peer_validation_hook] if is_chief else [], num_steps=flags.step,
              save_checkpoint_steps=validation_iteration_count, saver=saver if is_chief else None)
    best_upper_div = peer_validation_hook.get_best_upper_div()
    best_mean_div = peer_validation_hook.get_best_mean_div()

    def worst(value):       # both validation directions: the larger divergence; none validated (e.g. a non-chief rank): None
        return (max(value) if value else None) if isinstance(value, Sequence) else value

    return [worst(best_upper_div), worst(best_mean_div)]


def main(argv=None):
    flags, _ = build_parser().parse_known_args(argv)
    if flags.flag_config_file:
        flags = update_flags_from_json(flags, flags.flag_config_file)
    print("Running on training mode")
    print("Output divergence values:", run_session(params=dict(vars(flags)), base_log_path=flags.base_log_path))
    parallel.finish()


if __name__ == "__main__":
    main()
