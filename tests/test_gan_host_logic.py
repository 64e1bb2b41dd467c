"""CPU checks of the GAN host logic that needs no device: the gan_type registry (gan/wrapper_registry.py:21-94), the
identity-weight rule of the contrastive wrappers (cut_wrapper.py:593, dcl_gan_wrapper.py:238), the _get_lr schedule
(gan/wrappers/gan_common.py:222-244), tfgan's tensor_pool semantics and the sequential-hook order."""
from types import SimpleNamespace

import torch

FLAGS = SimpleNamespace(cycle_consistency_loss_weight=10.0, identity_loss_weight=0.5, use_identity_loss=True,
                        nce_loss_weight=10.0, tau=0.07, patches=6, embedded_feat_size=2, batch_size=32,
                        discriminator_reg_scale=1e-5, gen_disc_reg_scale=1e-4)


def test_registry_has_every_reference_gan_type():
    from hypelcnn_b200.gan.wrapper_registry import get_sampling_map, get_wrapper, get_wrapper_dict
    from hypelcnn_b200.gan.wrappers.wrapper import Wrapper
    d = get_wrapper_dict(FLAGS)
    assert set(d) == {"cycle_gan", "gan_x2y", "gan_y2x", "cut_x2y", "cut_y2x", "dcl_gan", "dcl_cycle_gan"}
    assert all(isinstance(w, Wrapper) for w in d.values())
    assert d["cut_x2y"]._swap_inputs is False and d["cut_y2x"]._swap_inputs is True
    assert d["dcl_cycle_gan"]._cycle_consistency_loss_weight == 10.0
    assert type(get_wrapper("dcl_gan", FLAGS)).__name__ == "DCLGANWrapper"
    assert "dummy" in get_sampling_map()


def test_identity_weight_is_zeroed_without_identity_loss():
    from hypelcnn_b200.gan.wrapper_registry import get_wrapper_dict
    off = SimpleNamespace(**{**vars(FLAGS), "use_identity_loss": False})
    for key in ("cut_x2y", "cut_y2x", "dcl_gan", "dcl_cycle_gan"):
        assert get_wrapper_dict(FLAGS)[key]._identity_loss_weight == 0.5
        assert get_wrapper_dict(off)[key]._identity_loss_weight == 0.0


def test_lr_schedule_constant_then_linear_to_zero():
    from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import get_lr
    assert get_lr(2e-4, 1000, 0) == 2e-4 and get_lr(2e-4, 1000, 499) == 2e-4
    assert abs(get_lr(2e-4, 1000, 750) - 1e-4) < 1e-12 and get_lr(2e-4, 1000, 1000) == 0.0
    assert get_lr(2e-4, 1000, 5000) == 0.0                      # polynomial_decay clamps at decay_steps
    assert get_lr(1e-3, 7, 3) == 1e-3 * (1 - 0 / 4)            # odd step counts: const_steps = 3, decay over 4


def test_tensor_pool_fills_then_swaps_with_probability_half():
    from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import TensorPool
    pool = TensorPool(4, 0.5, seed=0)
    first = [pool(torch.full((2,), float(i)))[0].item() for i in range(4)]
    assert first == [0.0, 1.0, 2.0, 3.0] and len(pool.pool) == 4          # until full: stored and returned
    later = [pool(torch.full((2,), float(i)))[0].item() for i in range(4, 400)]
    swapped = sum(1 for i, v in zip(range(4, 400), later) if v != float(i))
    assert 140 < swapped < 260                                             # p = 0.5
    assert all(v <= float(i) for i, v in zip(range(4, 400), later))        # only older tensors come back
    assert TensorPool(0)(torch.ones(1)).item() == 1.0                      # pool_size 0: identity


def test_sequential_hook_order_of_the_contrastive_wrappers():
    from hypelcnn_b200.gan.wrappers.cut_wrapper import CUTTrainOps, CUTWrapper
    from hypelcnn_b200.gan.wrappers.dcl_gan_wrapper import DCLGANTrainOps, DCLGANWrapper
    clock = {"global_step": 0, "gen": 0, "dis": 0, "feat": 0}
    calls = []

    class FakeTrainer:
        def __init__(self, tag):
            self.tag, self.clock = tag, clock

        def generator_train_op(self, x, y, lr):
            calls.append((self.tag, "gen", lr))

        def discriminator_train_op(self, x, y, lr):
            calls.append((self.tag, "dis", lr))

        def gen_discriminator_train_op(self, x, y, lr):
            calls.append((self.tag, "feat", lr))

    ops = CUTTrainOps(FakeTrainer("cut"), 10, 2e-4, 1e-4, 5e-5)
    ops.train_iteration(None, None)
    assert clock["global_step"] == 1 and calls == [("cut", "gen", 2e-4), ("cut", "dis", 1e-4), ("cut", "feat", 5e-5)]
    hooks = CUTWrapper(10.0, 0.5, True, 0.07, 32, False).get_train_hooks_fn()(ops)
    assert [h.__name__ for h in hooks] == ["generator_train_op", "discriminator_train_op", "gen_discriminator_train_op"]
    calls.clear()
    both = SimpleNamespace(model_x2y=FakeTrainer("x2y"), model_y2x=FakeTrainer("y2x"))
    dops = DCLGANTrainOps(both, 10, 2e-4, 1e-4, 5e-5)
    for _ in range(6):
        dops.train_iteration(None, None)                                   # steps 2..7: the decay half starts at 5
    assert clock["global_step"] == 7
    assert [c[:2] for c in calls[:6]] == [("x2y", "gen"), ("x2y", "dis"), ("x2y", "feat"),
                                          ("y2x", "gen"), ("y2x", "dis"), ("y2x", "feat")]
    assert calls[-1][2] == 5e-5 * (1 - 2 / 5)                              # global_step 7 of 10: linear decay
    assert len(DCLGANWrapper(10.0, 0.5, True, 0.07, 32).get_train_hooks_fn()(dops)) == 6
