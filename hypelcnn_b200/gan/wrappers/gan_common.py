"""gan/wrappers/gan_common.py of the reference on the device: the per-pixel generator inference over a matrix, and the
validation side of GAN training (best-ratio bookkeeping, band-ratio statistics, validation hooks, sample loading)."""
import bisect
import json
import os
import random
from json import JSONDecodeError

import numpy
import torch

from hypelcnn_b200.gan.shadow_data_models import _generator_rows

model_generator_name = "Generator"
model_base_name = "Model"


def adj_shadow_ratio(shadow_ratio, is_shadow):
    return 1. / shadow_ratio if is_shadow else shadow_ratio


def create_inference_for_matrix_input(input_tensor, is_shadow_graph, clip_invalid_values, generator_variables,
                                      copy_extra=0):
    """Reference: gan/wrappers/gan_common.py:282-304 — the generator applied to every pixel of [B,H,W,C(+extra)]
    separately (the reference builds H*W sub-graphs sharing weights; here every pixel is one row of ONE launch).
    With clip_invalid_values a pixel keeps its input spectrum unless the generated mean moved the expected way.
    copy_extra trailing channels (the LiDAR band, gan/gan_utilities.py:31-35) pass through unchanged."""
    if input_tensor.dim() != 4:
        raise ValueError("input must be [B,H,W,C]")
    x = input_tensor.contiguous()
    C = x.shape[3] - copy_extra
    rows = x.reshape(-1, x.shape[3])
    out = torch.empty_like(rows)
    _generator_rows(rows, out, C, copy_extra, generator_variables, clip_invalid_values, is_shadow_graph)
    return out.reshape(x.shape)


# --------------------------------------------------------------------------------------------------------------- #
# Validation during GAN training (reference: gan/wrappers/gan_common.py:32-219, 307-429).  The reference evaluates a
# second TF sub-graph through session.run inside SessionRunHooks; here a hook holds the validation spectra in HBM,
# calls the inference callable (one generator launch) and reduces the band ratios with a handful of tensor ops.
# --------------------------------------------------------------------------------------------------------------- #
input_x_tensor_name = "x"
input_y_tensor_name = "y"


class BestRatioHolder:
    """The ``max_size`` lowest divergences seen so far, kept as an ascending list of (iteration, divergence) in
    ``data_holder`` (the attribute callers and the JSON file use; reference gan_common.py:47-104).  A new value lands in
    front of the held values that are >= it — ``bisect_left`` over the divergences — and the largest drops off the end
    once the list is over ``max_size``."""

    def __init__(self, max_size) -> None:
        self.max_size = max_size
        self.data_holder = []

    def add_point(self, iteration, diver_val):
        point = (int(iteration), float(diver_val))                 # plain Python numbers: the list goes to JSON
        at = bisect.bisect_left([held[1] for held in self.data_holder], point[1])
        self.data_holder[at:at] = [point]
        del self.data_holder[self.max_size:]

    def get_best_diver(self):
        return self.data_holder[0][1] if self.data_holder else None

    def get_point_with_itr(self, iteration):
        return next(((it, div) for it, div in self.data_holder if it == iteration), (None, None))

    def load(self, file_address):
        """A missing or undecodable file is reported and leaves the holder as it is."""
        try:
            with open(file_address, "rb") as stream:
                self.data_holder = json.load(stream)
        except IOError:
            print(f"File {file_address} file not found. No best ratio is loaded.")
        except JSONDecodeError:
            print(f"File {file_address} file can not be decoded. No best ratio is loaded.")
        else:
            print(f"Best ratio file {file_address} is loaded.", self.data_holder)

    def save(self, file_address):
        with open(file_address, "w") as stream:
            json.dump(self.data_holder, stream)

    @staticmethod
    def create_common_iterations(ratio_holder_1, ratio_holder_2):
        """Iterations held by both, with the second holder's divergences."""
        common = BestRatioHolder(ratio_holder_1.max_size)
        second = {}
        for it, div in ratio_holder_2.data_holder:
            second.setdefault(it, div)
        for it, _ in ratio_holder_1.data_holder:
            if it in second:
                common.add_point(it, second[it])
        return common

    def __str__(self) -> str:
        return str(self.data_holder)


def _iteration_of(run_context):
    """Hooks are called as ``after_run(run_context, run_values)`` like tf SessionRunHooks; the training loop of
    gan_train_for_shadow.gan_train passes an object with ``global_step`` (an int is accepted too)."""
    return int(getattr(run_context, "global_step", run_context))


class BaseValidationHook:
    """State shared by the validation hooks (reference :107-138): two best-10 lists and the validation cadence."""

    def __init__(self, iteration_freq, log_dir, shadow_ratio):
        self._iteration_frequency, self._log_dir, self._shadow_ratio = iteration_freq, log_dir, shadow_ratio
        self.best_mean_div_holder, self.best_upper_div_holder = BestRatioHolder(10), BestRatioHolder(10)
        self.validation_itr_mark = False

    def after_create_session(self, session=None, coord=None):
        pass

    def _is_validation_itr(self, current_iteration):
        """Frequency 0: every iteration.  Otherwise iterations 1 + k * frequency with k >= 1 (never iteration 1)."""
        freq = self._iteration_frequency
        return freq == 0 or (current_iteration != 1 and current_iteration % freq == 1)

    def get_best_mean_div(self):
        return self.best_mean_div_holder.get_best_diver()

    def get_best_upper_div(self):
        return self.best_upper_div_holder.get_best_diver()


class PeerValidationHook:
    """Several validation hooks run side by side (reference :141-166: the shadowed and the de-shadowed direction); on
    a validation iteration the iterations that are among the best of the first two are reported."""

    def __init__(self, *validation_base_hooks):
        self._validation_base_hooks = validation_base_hooks

    def after_create_session(self, session=None, coord=None):
        for hook in self._validation_base_hooks:
            hook.after_create_session(session, coord)

    def after_run(self, run_context, run_values=None):
        for hook in self._validation_base_hooks:
            hook.after_run(run_context, run_values)
        first, second = self._validation_base_hooks[0], self._validation_base_hooks[1]
        if first.validation_itr_mark:
            print("Best common options:",
                  BestRatioHolder.create_common_iterations(first.best_mean_div_holder, second.best_mean_div_holder))

    def _bests(self, getter):
        values = [getter(hook) for hook in self._validation_base_hooks]
        return [v for v in values if v is not None]

    def get_best_mean_div(self):
        return self._bests(lambda hook: hook.get_best_mean_div())

    def get_best_upper_div(self):
        return self._bests(lambda hook: hook.get_best_upper_div())


def _kl_divergence(p, q):
    return torch.sum(torch.where(p != 0., p * torch.log(p / q), torch.zeros_like(p)))


def _js_divergence(p, q):
    m = 0.5 * (p + q)
    return 0.5 * _kl_divergence(p, m) + 0.5 * _kl_divergence(q, m)


def create_stats_tensor(generate_y_tensor, images_x_input_tensor, shadow_ratio):
    """Reference :314-330, evaluated eagerly on whatever device the spectra live on (fp32 like the TF graph).
    generated / input * shadow_ratio per band -> samples with a non-finite ratio dropped -> band-wise mean and
    population std -> |JS(|mean - 1|, 0)| and the same for mean + std.
    Returns (div_mean, div_upper, ratio [kept, bands], mean [bands], std [bands])."""
    generated = torch.as_tensor(generate_y_tensor, dtype=torch.float32)
    inputs = torch.as_tensor(images_x_input_tensor, dtype=torch.float32).to(generated.device)
    ratio = torch.as_tensor(numpy.asarray(shadow_ratio, dtype=numpy.float32) if not isinstance(shadow_ratio, torch.Tensor)
                            else shadow_ratio, dtype=torch.float32).to(generated.device)
    ratio_tensor = (generated / inputs * ratio).squeeze(2).squeeze(1)
    finite_map = torch.isfinite(ratio_tensor).all(dim=1)
    ratio_inf_eliminated = ratio_tensor[finite_map]
    mean = ratio_inf_eliminated.mean(dim=0)
    std = ratio_inf_eliminated.std(dim=0, unbiased=False)
    div_mean = torch.abs(_js_divergence(torch.abs(mean - 1), torch.zeros_like(mean)))
    div_upper = torch.abs(_js_divergence(torch.abs(mean + std - 1), torch.zeros_like(mean)))
    return div_mean, div_upper, ratio_inf_eliminated, mean, std


def calculate_stats_from_samples(infer_model, data_sample_list, shadow_ratio, log_dir, current_iteration, plt_name,
                                 bands):
    """Reference :333-361 (the offline variant used by gan_infer_for_shadow): same statistics, returns the mean
    divergence.  ``infer_model`` is the callable of InferenceWrapper.make_inference_graph; the reference's
    (sess, input placeholder, output tensor) triple collapses into it."""
    samples = torch.as_tensor(data_sample_list, dtype=torch.float32)
    generated = infer_model(samples)
    div_mean, _, final_ratio, mean, std = create_stats_tensor(generated, samples.to(generated.device), shadow_ratio)
    final_ratio, mean, std = final_ratio.cpu().numpy(), mean.cpu().numpy(), std.cpu().numpy()
    print_overall_info(mean, std)
    plot_overall_info(bands, numpy.percentile(final_ratio, 50, axis=0), numpy.percentile(final_ratio, 10, axis=0),
                      numpy.percentile(final_ratio, 90, axis=0), current_iteration, plt_name, log_dir)
    return float(div_mean)


def load_samples_for_testing(data_set, sample_count, neighborhood, shadow_map, fetch_shadows):
    """Reference :364-385: ``sample_count`` random shadowed (or shadow-free) pixels, HSI bands only.  The draws use
    Python's ``random.randint`` in the reference's order, so ``random.seed`` reproduces the reference's picks; the
    patches come from one batched gather when the data set offers it.  Note the reference crops the map by
    ``neighborhood`` but uses the cropped indices as scene coordinates unchanged — kept."""
    band_size = data_set.get_casi_band_count()
    if neighborhood > 0:
        shadow_map = shadow_map[neighborhood:-neighborhood, neighborhood:-neighborhood]
    indices = numpy.where(shadow_map > 0) if fetch_shadows else numpy.where(shadow_map == 0)
    picks = [random.randint(0, indices[0].size - 1) for _ in range(sample_count)]
    targets = numpy.stack([indices[1][picks], indices[0][picks]], axis=1).astype(numpy.int32).reshape(-1, 2)
    batched = getattr(data_set, "get_data_points", None)
    if batched is not None:
        points = batched(targets)
        return [points[i, :, :, 0:band_size] for i in range(sample_count)]
    return [data_set.get_data_point(int(x), int(y))[:, :, 0:band_size] for x, y in targets]


def read_hsi_data(loader, data_set, shadow_map, pairing_method, sampling_method_map):
    """The chosen sampler's (normal, shadowed) pair matrices without the trailing LiDAR channel (reference :388-395)."""
    sampler = sampling_method_map.get(pairing_method)
    if sampler is None:
        raise ValueError(f"Wrong sampling parameter value ({pairing_method}).")
    bands = data_set.get_casi_band_count()
    return tuple(matrix[..., :bands] for matrix in sampler.get_sample_pairs(data_set, loader, shadow_map))


def plot_overall_info(bands, mean, lower_bound, upper_bound, iteration, plt_name, log_dir):
    """The reference (:398-419) draws the band-ratio curve and its 10-90 % envelope with matplotlib, which is not part of
    this image; the curve goes to ``{plt_name}_{iteration}.csv`` (band, median, p10, p90) for any plotting tool."""
    table = numpy.column_stack([numpy.asarray(bands, dtype=numpy.float64), mean, lower_bound, upper_bound])
    numpy.savetxt(os.path.join(log_dir, f"{plt_name}_{iteration}.csv"), table, delimiter=",",
                  header="band,ratio_p50,ratio_p10,ratio_p90", comments="")


def print_overall_info(mean, std):
    """``mean±std`` per band in the reference's text layout (:422-429): brackets around the list, a line break after
    the bands whose index is 1 mod 5, a space after the others."""
    cells = [f"{m:2.4f}\u00B1{s:2.2f}" for m, s in zip(mean, std)]
    if cells:
        cells[0] = "[ " + cells[0]
        if len(cells) > 1:
            cells[-1] += " ]"
    text = "".join(cell + ("\n" if i % 5 == 1 else " ") for i, cell in enumerate(cells))
    print("Mean&std Generated vs Original Ratio: ")
    print(text, end="")


class ValidationHook(BaseValidationHook):
    """Reference :169-219.  ``infer_model`` is a callable [N,1,1,bands] -> [N,1,1,bands] (the inference wrapper's
    generator over a matrix); ``input_tensor`` — a placeholder in the reference — is accepted and unused.  The
    validation spectra are picked once (load_samples_for_testing) and kept on the device of the data set."""

    def __init__(self, iteration_freq, sample_count, log_dir, loader, data_set, neighborhood, shadow_map, shadow_ratio,
                 input_tensor, infer_model, name_suffix, fetch_shadows):
        super().__init__(iteration_freq, log_dir, shadow_ratio)
        self._writer = None
        self._infer_model = infer_model
        self._name_suffix = name_suffix
        self._plt_name = f"band_ratio_{name_suffix}"
        self._best_mean_div_addr = os.path.join(self._log_dir, f"best_ratio_{name_suffix}.json")
        self.best_mean_div_holder.load(self._best_mean_div_addr)
        self._bands = loader.get_band_measurements()
        samples = load_samples_for_testing(data_set, sample_count, neighborhood, shadow_map, fetch_shadows=fetch_shadows)
        if samples and isinstance(samples[0], torch.Tensor):
            self._data_sample_list = torch.stack(samples).contiguous()
        else:
            self._data_sample_list = torch.as_tensor(numpy.asarray(samples, dtype=numpy.float32))

    def after_create_session(self, session=None, coord=None):
        from hypelcnn_b200.classify.summaries import ClassificationSummaryWriter
        self._writer = ClassificationSummaryWriter(self._log_dir)

    def after_run(self, run_context, run_values=None):
        current_iter = _iteration_of(run_context)
        self.validation_itr_mark = self._is_validation_itr(current_iter)
        if self.validation_itr_mark:
            generated = self._infer_model(self._data_sample_list)
            div_mean, div_upper, ratio, mean, std = create_stats_tensor(
                generated, self._data_sample_list.to(generated.device), self._shadow_ratio)
            div_mean, div_upper = float(div_mean), float(div_upper)
            self.best_mean_div_holder.add_point(current_iter, div_mean)
            self.best_mean_div_holder.save(self._best_mean_div_addr)
            self.best_upper_div_holder.add_point(current_iter, div_upper)
            if self._writer is not None:
                self._writer.add_scalar(f"divergence_{self._name_suffix}", div_mean, current_iter)
            self.print_stats(current_iter, div_mean, div_upper, mean.cpu().numpy(), ratio.cpu().numpy(),
                             std.cpu().numpy())

    def print_stats(self, current_iteration, div_mean, div_upper, mean, ratio, std):
        print(f"Validation metrics for {self._name_suffix} #{current_iteration}")
        print_overall_info(mean, std)
        plot_overall_info(self._bands, numpy.percentile(ratio, 50, axis=0), numpy.percentile(ratio, 10, axis=0),
                          numpy.percentile(ratio, 90, axis=0), current_iteration, self._plt_name, self._log_dir)
        print(f"Divergence for {self._name_suffix}; mean:{div_mean}, upper:{div_upper}")
        print(f"Best {self._name_suffix} options:{self.best_mean_div_holder}")


def create_base_validation_hook(data_set, loader, log_dir, neighborhood, shadow_map, shadow_ratio,
                                validation_iteration_count, validation_sample_count, model_forward, model_backward,
                                x_input_tensor=None, y_input_tensor=None):
    """Reference: gan/wrappers/cycle_gan_wrapper.py:22-45 — normal spectra through the forward generator judged
    against shadow_ratio ("shadowed"), shadowed spectra through the backward generator against 1 / shadow_ratio
    ("deshadowed"), as peers."""
    shadowed = ValidationHook(iteration_freq=validation_iteration_count, sample_count=validation_sample_count,
                              log_dir=log_dir, loader=loader, data_set=data_set, neighborhood=neighborhood,
                              shadow_map=shadow_map, shadow_ratio=shadow_ratio, input_tensor=x_input_tensor,
                              infer_model=model_forward, fetch_shadows=False, name_suffix="shadowed")
    de_shadowed = ValidationHook(iteration_freq=validation_iteration_count, sample_count=validation_sample_count,
                                 log_dir=log_dir, loader=loader, data_set=data_set, neighborhood=neighborhood,
                                 shadow_map=shadow_map, shadow_ratio=1. / shadow_ratio, input_tensor=y_input_tensor,
                                 infer_model=model_backward, fetch_shadows=True, name_suffix="deshadowed")
    return PeerValidationHook(shadowed, de_shadowed)


# --------------------------------------------------------------------------------------------------------------- #
# The remaining names of the reference module, for callers written against it.
# --------------------------------------------------------------------------------------------------------------- #
def _get_lr(base_lr, max_number_of_steps, global_step=0):
    """Reference :222-244: ``base_lr`` for the first half of the run, then linear to zero at ``max_number_of_steps``
    (tf polynomial_decay, power 1).  The global step is an argument here (a graph variable in the reference)."""
    lr_constant_steps = max_number_of_steps // 2
    if global_step < lr_constant_steps:
        return base_lr
    decay_steps = max_number_of_steps - lr_constant_steps
    return base_lr * (1.0 - min(global_step - lr_constant_steps, decay_steps) / decay_steps)


def define_standard_train_ops(gan_model, gan_loss, max_number_of_steps, generator_lr, discriminator_lr):
    """Reference :247-279: Adam(beta1 0.5) generator / discriminator train ops with the schedule above.  ``gan_model``
    is what a wrapper's define_model returned (it carries the trainer)."""
    from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import GANTrainOps
    return GANTrainOps(getattr(gan_model, "trainer", gan_model), max_number_of_steps, generator_lr, discriminator_lr)


def create_input_tensor(data_set, is_shadow_graph):
    """Reference :307-312 creates the placeholder ``x`` / ``y`` of shape [None, P, P, bands]; eager code feeds tensors
    directly, so only the description of that input is returned."""
    shape = data_set.get_data_shape()
    return {"name": input_x_tensor_name if is_shadow_graph else input_y_tensor_name,
            "shape": [None, shape[0], shape[1], data_set.get_casi_band_count()], "dtype": "float32"}


class InitializerHook:
    """Reference :32-44 feeds the pair matrices into the tf.data placeholders when the session starts.  The pair
    iterator of gan_train_for_shadow.load_op already owns its (device-resident) data: nothing to do, kept so that hook
    lists written for the reference still construct."""

    def __init__(self, input_itr, normal_placeholder=None, shadow_placeholder=None, normal_data=None, shadow_data=None):
        self.input_itr = input_itr
        self.normal_data, self.shadow_data = normal_data, shadow_data
        self.normal_placeholder, self.shadow_placeholder = normal_placeholder, shadow_placeholder

    def after_create_session(self, session=None, coord=None):
        pass

    def after_run(self, run_context, run_values=None):
        pass
