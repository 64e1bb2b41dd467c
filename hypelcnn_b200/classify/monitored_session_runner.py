"""The classification training loop and its hooks (reference: classify/monitored_session_runner.py).

Same function / class names and argument lists; ``run_monitored_session`` is a plain loop around the engine's train step
instead of a tf MonitoredTrainingSession.  What the session machinery did implicitly is spelled out here:

* StopAtStepHook(last_step = required_steps - 1): the loop ends once global_step reaches ``required_steps - 1``;
* NanTensorHook(fail_on_nan_loss=False): a NaN loss ends the loop with a message instead of raising (noticed one step
  late, so that the check never stalls the launch queue);
* the summary saver (``save_summaries_steps = 100``) and ValidationHook write the tags of
  ``add_classification_summaries`` through classify/summaries.py;
* the checkpoint saver: ``model.ckpt-<step>.safetensors`` (TF variable names, Adam slots, global_step) every
  ``save_checkpoint_steps`` and at the end, the latest one restored at start, 20 kept.
"""
import gc
import glob
import os
import re

import torch

from hypelcnn_b200.classify.summaries import ClassificationSummaryWriter
from hypelcnn_b200.common.common_nn_ops import TrainingResult, calculate_accuracy

TEST_ITERATION_COUNT = 100      # monitored_session_runner.py:139: test cadence = summary cadence
CHECKPOINTS_TO_KEEP = 20        # Saver(max_to_keep=20) (:165)


def set_run_seed():
    """Reference :11-13: the graph-level seed 1234.  Here it seeds torch's generators (iterator shuffles); the engine's
    own draws (initialisation, dropout, augmentation) are seeded with the same constant per model."""
    torch.manual_seed(1234)


class ClassificationSummaries:
    """What ``add_classification_summaries`` registers in the reference's "summary_op" collection: evaluated at an
    iteration it logs the training loss / learning rate, both confusion matrices, the accuracies and kappa."""

    def __init__(self, cross_entropy, learning_rate, log_all_model_variables, testing_nn_params, validation_nn_params):
        self.cross_entropy, self.learning_rate = cross_entropy, learning_rate
        self.log_all_model_variables = log_all_model_variables
        self.testing_nn_params, self.validation_nn_params = testing_nn_params, validation_nn_params
        self.model_variables = None     # callable -> {TF variable name: array}; set by run_monitored_session

    def write(self, writer, iteration):
        loss = self.cross_entropy()
        if loss is None:
            return
        loss = loss[0] if getattr(loss, "dim", lambda: 0)() > 0 else loss
        variables = self.model_variables() if (self.log_all_model_variables and self.model_variables) else None
        validation = self.validation_nn_params if self.validation_nn_params is not None else self.testing_nn_params
        writer.add_classification_summaries(iteration, loss, self.learning_rate(), self.testing_nn_params.metrics,
                                            validation.metrics, variables)


def add_classification_summaries(cross_entropy, learning_rate, log_all_model_variables, testing_nn_params,
                                 validation_nn_params):
    """Reference :16-28."""
    return ClassificationSummaries(cross_entropy, learning_rate, log_all_model_variables, testing_nn_params,
                                   validation_nn_params)


class RunContext:
    """What hooks get after every train step: ``global_step`` and the ``session`` slot the importers' ``init_tensors``
    accept (always None: execution is eager)."""

    def __init__(self):
        self.session, self.global_step, self.stop_requested = None, 0, False

    def request_stop(self):
        self.stop_requested = True


class InitHook:
    """Reference :31-46: restore the shadow augmenter's generator weights, then (re)initialise the training input."""

    def __init__(self, training_nn_params, training_tensor, augmentation_info, restorer, importer):
        self.importer = importer
        self.restorer = restorer
        self.augmentation_info = augmentation_info
        self.training_nn_params = training_nn_params
        self.training_tensor = training_tensor

    def after_create_session(self, session, coord):
        info = self.augmentation_info
        if info is not None and info.perform_shadow_augmentation:
            if info.shadow_struct is not None and info.shadow_struct.shadow_op_initializer is not None:
                info.shadow_struct.shadow_op_initializer(self.restorer, session)
        self.importer.init_tensors(session, self.training_tensor, self.training_nn_params)

    def after_run(self, run_context, run_values):
        pass


class ValidationHook:
    """Reference :49-89: at ``required_steps - 1`` and at every ``1 + k * iteration`` (k >= 1) evaluate the validation
    set and log all classification summaries."""

    def __init__(self, validation_nn_params, validation_tensor, class_range, required_steps, iteration, summary_dir,
                 importer):
        self.importer = importer
        self.required_steps = required_steps
        self.validation_nn_params = validation_nn_params
        self.validation_tensor = validation_tensor
        self.validation_accuracy = 0
        self.class_range = class_range
        self.summary_dir = summary_dir
        self._iteration_count = iteration
        self._writer = None
        self.summaries = None           # ClassificationSummaries; set by run_monitored_session

    def after_create_session(self, session, coord):
        self._writer = ClassificationSummaryWriter(self.summary_dir)

    def after_run(self, run_context, run_values):
        iteration = run_context.global_step
        if self.validation_nn_params is not None:
            if (iteration == self.required_steps - 1) or (iteration % self._iteration_count == 1 and iteration != 1):
                self.importer.init_tensors(run_context.session, self.validation_tensor, self.validation_nn_params)
                self.validation_accuracy, class_recall, class_precisions, kappa, mean_per_class_accuracy = \
                    calculate_accuracy(run_context.session, self.validation_nn_params, self.class_range)
                print('Validation metrics #%d : Overall accuracy=%g, Class based average accuracy=%g, Kappa=%g' % (
                    iteration, self.validation_accuracy, mean_per_class_accuracy, kappa))
                if self.summaries is not None and self._writer is not None:
                    self.summaries.write(self._writer, iteration)
                    self._writer.flush()
                gc.collect()

    def end(self, session):
        if self._writer is not None:
            self._writer.close()


class TestHook:
    """Reference :92-124: every 100 iterations (at 1, 101, ...) and once at the end, read the loss and evaluate the
    test set if there is one."""

    def __init__(self, testing_nn_params, testing_tenser, cross_entropy, test_iteration_count, class_range, importer):
        self.importer = importer
        self.testing_nn_params = testing_nn_params
        self.testing_tensor = testing_tenser
        self.testing_accuracy = 0
        self.loss = 0
        self.cross_entropy = cross_entropy
        self._test_iteration_count = test_iteration_count
        self.class_range = class_range
        self._last_iteration = 0

    def after_create_session(self, session, coord):
        pass

    def after_run(self, run_context, run_values):
        self._last_iteration = run_context.global_step
        if self._last_iteration % self._test_iteration_count == 1:
            self.__perform_action(run_context.session, self._last_iteration)

    def end(self, session):
        self.__perform_action(session, self._last_iteration)

    def __perform_action(self, session, iteration):
        loss = self.cross_entropy()
        if loss is not None:
            self.loss = float(loss[0] if getattr(loss, "dim", lambda: 0)() > 0 else loss)
        if _element_count(self.testing_nn_params.data_with_labels.data) != 0:
            self.importer.init_tensors(session, self.testing_tensor, self.testing_nn_params)
            self.testing_accuracy, class_recall, class_precisions, kappa, mean_per_class_accuracy = calculate_accuracy(
                session, self.testing_nn_params, self.class_range)
        print('Training step=%d, Testing accuracy=%g, loss=%.5f' % (iteration, self.testing_accuracy, self.loss))


def _element_count(data):
    """``data.size`` of the reference's numpy array, for a device tensor as well."""
    if data is None:
        return 0
    return data.numel() if hasattr(data, "numel") else data.size


class CheckpointSaver:
    """The Saver(nn_core/*, global_step, training_optimizer/*) + CheckpointSaverHook pair (:163-180) over
    PatchEngine.save_checkpoint / load_checkpoint."""

    PATTERN = re.compile(r"model\.ckpt-(\d+)\.safetensors$")

    def __init__(self, log_dir, engine_of, save_checkpoint_steps, max_to_keep=CHECKPOINTS_TO_KEEP):
        self.log_dir, self.engine_of = log_dir, engine_of
        self.save_checkpoint_steps, self.max_to_keep = save_checkpoint_steps, max_to_keep
        self.last_saved = None

    def existing(self):
        found = []
        for path in glob.glob(os.path.join(self.log_dir, "model.ckpt-*.safetensors")):
            m = self.PATTERN.search(path)
            if m:
                found.append((int(m.group(1)), path))
        return sorted(found)

    def restore_latest(self):
        found = self.existing()
        if not found:
            return None
        step, path = found[-1]
        self.engine_of().load_checkpoint(path)
        print(f"Restored {path}")
        return step

    def save(self, step):
        if self.last_saved == step:
            return
        os.makedirs(self.log_dir, exist_ok=True)
        self.engine_of().save_checkpoint(os.path.join(self.log_dir, f"model.ckpt-{step}.safetensors"))
        self.last_saved = step
        for _, path in self.existing()[:-self.max_to_keep]:
            os.remove(path)

    def after_run(self, step):
        if self.save_checkpoint_steps and step % self.save_checkpoint_steps == 0:
            self.save(step)


def _engine_of(train_step):
    """The PatchEngine behind a TrainOp; built from the training input's patch shape if no step has run yet."""
    model = train_step.model
    if model.engine is None:
        iterator = train_step.iterator
        images = getattr(iterator, "images", None)
        if images is None:
            images = iterator.inner.images
        model.engine_for(images[:1], train_step.alg)
    return model.engine


class _LossWatch:
    """NaN watch that never drains the launch queue: the loss of step n is copied into pinned host memory right behind
    step n (asynchronously, two alternating slots) and an event marks the copy; it is looked at after step n + 1 has
    been launched, waiting at most for that event — i.e. for step n, not for the work queued behind it."""

    def __init__(self):
        self._slots, self._events, self._turn, self._pending = None, None, 0, None

    def submit(self, loss):
        """Register this step's loss; returns True if the PREVIOUS step's loss was NaN."""
        previous_was_nan = self.check()
        if loss is None:
            return previous_was_nan
        loss = loss.reshape(-1)[:1] if hasattr(loss, "reshape") else torch.as_tensor([float(loss)])
        if loss.is_cuda:
            if self._slots is None:
                self._slots = [torch.empty(1, dtype=loss.dtype, pin_memory=True) for _ in range(2)]
                self._events = [torch.cuda.Event() for _ in range(2)]
            slot = self._turn
            self._turn ^= 1
            self._slots[slot].copy_(loss, non_blocking=True)
            self._events[slot].record()
            self._pending = slot
        else:
            self._pending = loss.clone()
        return previous_was_nan

    def check(self):
        """Is the registered (not yet examined) loss NaN?"""
        pending, self._pending = self._pending, None
        if pending is None:
            return False
        if isinstance(pending, int):
            self._events[pending].synchronize()
            return bool(torch.isnan(self._slots[pending]).item())
        return bool(torch.isnan(pending).item())


def _loss_over_ranks(loss):
    import torch.distributed as dist
    if loss is None or not (dist.is_initialized() and dist.get_world_size() > 1):
        return loss
    total = torch.as_tensor(loss, dtype=torch.float32).detach().reshape(-1)[:1].clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total


def run_monitored_session(cross_entropy, log_dir, class_range,
                          save_checkpoint_steps, validation_steps,
                          train_step, required_steps,
                          augmentation_info, training_nn_params, training_tensor,
                          testing_nn_params, testing_tensor,
                          validation_nn_params, validation_tensor,
                          importer, flags_as_json_str, alg_params_as_json_str,
                          summaries=None, engine_of=None, is_chief=True):
    """Reference :127-188.  ``summaries`` is what add_classification_summaries returned (the reference finds it
    through the graph's "summary_op" collection); ``engine_of`` overrides how the checkpoint saver reaches the engine;
    ``is_chief`` (MonitoredTrainingSession's flag, always True in the reference): only the chief writes checkpoints and
    summaries — every rank of a data-parallel run restores the same checkpoint and runs the same hooks."""
    augmentation_restorer = None
    if augmentation_info is not None and augmentation_info.perform_shadow_augmentation:
        if augmentation_info.shadow_struct is not None and \
                augmentation_info.shadow_struct.shadow_op_initializer is not None:
            augmentation_restorer = augmentation_info.shadow_struct.shadow_op_creater()

    if not is_chief:
        summaries = None
    validation_hook = ValidationHook(validation_nn_params, validation_tensor, class_range, required_steps,
                                     validation_steps, log_dir, importer)
    validation_hook.summaries = summaries
    test_hook = TestHook(testing_nn_params, testing_tensor, cross_entropy, TEST_ITERATION_COUNT, class_range, importer)
    initializer_hook = InitHook(training_nn_params, training_tensor, augmentation_info, augmentation_restorer, importer)
    hooks = [initializer_hook, validation_hook, test_hook]

    os.makedirs(log_dir, exist_ok=True)
    writer = ClassificationSummaryWriter(log_dir) if is_chief else None
    if writer is not None:
        writer.add_text("flags", flags_as_json_str, 0)                   # TextSummaryAtStartHook x 2 (:144-145)
        writer.add_text("algorithm_params", alg_params_as_json_str, 0)
    saver = CheckpointSaver(log_dir, engine_of or (lambda: _engine_of(train_step)),
                            save_checkpoint_steps if is_chief else None)
    if summaries is not None and summaries.model_variables is None:
        summaries.model_variables = lambda: saver.engine_of().export_variables()

    for hook in hooks:
        hook.after_create_session(None, None)
    context = RunContext()
    restored = saver.restore_latest()
    context.global_step = train_step.global_step if restored is not None else 0
    last_step = required_steps - 1                                       # StopAtStepHook(last_step=...)
    loss_watch = _LossWatch()
    while context.global_step < last_step and not context.stop_requested:
        try:
            train_step.run()
        except StopIteration:                                            # --epoch given: the input ran out
            break
        context.global_step = train_step.global_step
        # NanTensorHook(fail_on_nan_loss=False).  The loss of step n is looked at after step n + 1 has been launched, so
        # reading it does not drain the device queue every step (at the reference's batch of 48 a step is ~170 short
        # launches and the host has to run ahead); a divergence is noticed one step late.
        # Data parallel: the watched value is the SUM over ranks of the local losses (NaN on one rank -> NaN everywhere,
        # one 4-byte all-reduce queued behind the step, no host wait), so every rank takes the stop decision at the same
        # global step and nobody is left alone in the next gradient all-reduce.
        if loss_watch.submit(_loss_over_ranks(cross_entropy())):
            print("Model diverged with loss = NaN.")
            context.request_stop()
        for hook in hooks:
            hook.after_run(context, None)
        if summaries is not None and context.global_step % TEST_ITERATION_COUNT == 0:
            summaries.write(writer, context.global_step)
        saver.after_run(context.global_step)
    if not context.stop_requested and loss_watch.check():
        print("Model diverged with loss = NaN.")
    for hook in hooks:
        if hasattr(hook, "end"):
            hook.end(None)
    if context.global_step > 0 and is_chief:
        saver.save(context.global_step)
    if writer is not None:
        writer.close()
    return TrainingResult(validation_accuracy=validation_hook.validation_accuracy,
                          test_accuracy=test_hook.testing_accuracy, loss=test_hook.loss)
