"""The GAN plug-in contract: one ``Wrapper`` per gan_type for training, one ``InferenceWrapper`` for applying the
trained generators.  Method names and signatures are the reference's (gan/wrappers/wrapper.py:4-38); what flows through
them here are CUDA float32 tensors and eager callables instead of TF graph nodes:

  define_model(images_x, images_y)        [B,1,1,C] normal / shadowed spectra -> a model tuple owning the trainer
  define_loss(model)                      -> a loss tuple (the losses are evaluated inside the train ops)
  define_train_ops(model, loss, max_number_of_steps, generator_lr=, discriminator_lr=[, gen_discriminator_lr=])
                                          -> train-ops object with *_train_op callables and global_step_inc_op
  get_train_hooks_fn()                    -> f(train_ops) listing the ops in the order one iteration runs them
"""
import abc


class Wrapper(abc.ABC):

    @abc.abstractmethod
    def define_model(self, images_x, images_y):
        """Create (once) the trainer with its variables and optimizer slots for spectra of this band count."""
        raise NotImplementedError

    @abc.abstractmethod
    def define_loss(self, model):
        """Bind the loss configuration (weights, tau, regularisation) to the model; no arithmetic happens here."""
        raise NotImplementedError

    @abc.abstractmethod
    def define_train_ops(self, model, loss, max_number_of_steps, **kwargs):
        """Adam(beta1 0.5) per variable group, learning rates constant for the first half of max_number_of_steps and
        then linear to zero."""
        raise NotImplementedError

    @abc.abstractmethod
    def get_train_hooks_fn(self):
        """The sequential-hook order of one iteration (tfgan get_sequential_train_hooks equivalents)."""
        raise NotImplementedError


class InferenceWrapper:
    """construct_inference_graph maps a [B,H,W,C] matrix through the forward ("shadow") or backward generator;
    clip_invalid_values keeps the input spectrum wherever the generated mean moves the wrong way.
    Like the reference's (gan/wrappers/wrapper.py:21), this class does not derive from ABC: the abstractmethod markers
    document the contract but a partial implementation still instantiates."""

    @abc.abstractmethod
    def construct_inference_graph(self, input_tensor, is_shadow_graph, clip_invalid_values):
        raise NotImplementedError

    @abc.abstractmethod
    def make_inference_graph(self, data_set, is_shadow_graph, clip_invalid_values):
        """-> (input placeholder or None, callable applying the generator to a matrix of that data set's shape)."""
        raise NotImplementedError

    @abc.abstractmethod
    def create_generator_restorer(self):
        """-> object with restore(...) that loads generator variables by their reference checkpoint names."""
        raise NotImplementedError

    @abc.abstractmethod
    def create_inference_hook(self, data_set, loader, log_dir, neighborhood, shadow_map, shadow_ratio,
                              validation_iteration_count, validation_sample_count):
        raise NotImplementedError
