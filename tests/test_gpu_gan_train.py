"""GPU parity of the CycleGAN training path (gan/wrappers/cycle_gan_wrapper.py, gan/wrappers/gan_common.py:222-279)
against the torch-autograd restatement in oracle/gan_ref.py: generator / discriminator backward kernels, the
generator-step and discriminator-step gradients, the LR schedule, and a short training run on the reference's
DummySampler data (gan/gan_sampling_methods.py:191-201: shadow = 0.5, normal = 1.0)."""
import numpy
import pytest
import torch

from oracle import gan_ref as R

pytestmark = pytest.mark.gpu


def _trainer(bands=64, seed=3, **kw):
    from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANTrainer
    t = CycleGANTrainer(bands, seed=seed, **kw)
    rng = numpy.random.default_rng(seed)
    t.gen_params.copy_(torch.tensor(rng.standard_normal(t.gen_params.numel()).astype(numpy.float32) * 0.05))
    t.dis_params.add_(torch.tensor(rng.standard_normal(t.dis_params.numel()).astype(numpy.float32) * 0.01).cuda())
    return t


def _data(n, bands, seed=1):
    rng = numpy.random.default_rng(seed)
    y = rng.uniform(0.02, 0.5, (n, bands)).astype(numpy.float32)          # shadowed spectra
    x = (y * numpy.linspace(1.5, 4, bands)).astype(numpy.float32)         # lit spectra (SURVEY S-C4)
    return torch.tensor(x).cuda(), torch.tensor(y).cuda()


def _rel(got, ref):
    return float((got.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("bands,n", [(64, 300), (32, 37)])
def test_generator_backward_matches_autograd(bands, n):
    t = _trainer(bands)
    x, _ = _data(n, bands)
    w = t.G()
    nets = t._gen_fwd(x, w)
    gout = torch.randn((n, bands), device="cuda")
    gw = torch.zeros_like(w)
    gin = t._gen_bwd(nets, gout, w, gw)
    xr = x.double().cpu().requires_grad_(True)
    wr = w.double().cpu().requires_grad_(True)
    out = R.t_generator(xr, wr)
    assert _rel(nets[:, 7, :], out.detach()) < 1e-5
    (out * gout.double().cpu()).sum().backward()
    assert _rel(gin, xr.grad) < 1e-4 and _rel(gw, wr.grad) < 1e-4


@pytest.mark.parametrize("bands,n", [(64, 300), (16, 9)])
def test_discriminator_forward_backward_match_autograd(bands, n):
    t = _trainer(bands)
    x, _ = _data(n, bands)
    w = t.DY()
    h, d = t._dis_fwd(x, w)
    gout = torch.randn_like(d)
    gw = torch.zeros_like(w)
    gin = t._dis_bwd(x, h, gout, w, gw, True)
    xr = x.double().cpu().requires_grad_(True)
    wr = w.double().cpu().requires_grad_(True)
    out = R.t_discriminator(xr, wr)
    assert _rel(d, out.detach()) < 1e-5
    (out * gout.double().cpu()).sum().backward()
    assert _rel(gin, xr.grad) < 1e-4 and _rel(gw, wr.grad) < 1e-4


def test_generator_step_gradients_match_the_cyclegan_objective():
    t = _trainer(64, cycle_consistency_loss_weight=10.0, identity_loss_weight=0.5)
    x, y = _data(64, 64)
    loss = t.generator_gradients(x, y).cpu()
    P = [p.double().cpu().requires_grad_(True) for p in (t.G(), t.F())]
    D = [p.double().cpu() for p in (t.DY(), t.DX())]
    total, gan, cyc, ident = R.t_generator_loss(x.double().cpu(), y.double().cpu(), P[0], P[1], D[0], D[1], 10.0, 0.5)
    total.backward()
    for got, ref in zip(loss.tolist(), (total.item(), gan.item(), cyc.item(), ident.item())):
        assert abs(got - ref) < 1e-5 * max(1.0, abs(ref))
    assert _rel(t.gen_grads[:t.ng], P[0].grad) < 2e-4 and _rel(t.gen_grads[t.ng:], P[1].grad) < 2e-4


def test_discriminator_step_gradients_match():
    t = _trainer(64, discriminator_reg_scale=1e-3)
    x, y = _data(48, 64)
    loss = t.discriminator_gradients(x, y, use_pool=False).cpu()
    G, Fw = t.G().double().cpu(), t.F().double().cpu()
    gx, fy = R.t_generator(x.double().cpu(), G), R.t_generator(y.double().cpu(), Fw)
    D = [p.double().cpu().requires_grad_(True) for p in (t.DY(), t.DX())]
    ref = R.t_discriminator_loss(x.double().cpu(), y.double().cpu(), gx, fy, D[0], D[1], 1e-3)
    ref.backward()
    assert abs(loss[0].item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert _rel(t.dis_grads[:t.nd], D[0].grad) < 2e-4 and _rel(t.dis_grads[t.nd:], D[1].grad) < 2e-4


def test_lr_schedule_and_tensor_pool():
    from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import TensorPool, get_lr
    assert get_lr(2e-4, 1000, 0) == 2e-4 and get_lr(2e-4, 1000, 499) == 2e-4        # constant first half
    assert abs(get_lr(2e-4, 1000, 750) - 1e-4) < 1e-12 and get_lr(2e-4, 1000, 1000) == 0.0
    pool = TensorPool(3, 0.5, seed=0)
    seen = [pool(torch.full((1,), float(i))) .item() for i in range(40)]
    assert seen[:3] == [0.0, 1.0, 2.0] and any(s != float(i) for i, s in enumerate(seen[3:], 3))


def test_wrapper_trains_on_dummy_pairs():
    """--pairing_method=dummy: shadow = 0.5, normal = 1.0 — the only known-answer fixture of the reference."""
    from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANInferenceWrapper, CycleGANWrapper
    wrapper = CycleGANWrapper(10.0, 0.5, True)
    shadow = torch.full((32, 1, 1, 64), 0.5, device="cuda")
    normal = shadow * 2
    model = wrapper.define_model(normal, shadow)
    loss = wrapper.define_loss(model)
    ops = wrapper.define_train_ops(model, loss, 400, generator_lr=2e-3, discriminator_lr=1e-4)
    first = last = None
    for it in range(200):
        lg, ld = ops.train_iteration(normal, shadow)
        if it == 0:
            first = lg.cpu()
        last = lg.cpu()
    assert ops.trainer.global_step == 200 and torch.isfinite(last).all()
    assert last[2] < first[2]                                   # the cycle-consistency term goes down
    infer = CycleGANInferenceWrapper(trainer=ops.trainer)
    out = infer.construct_inference_graph(normal, True, False)  # x -> y: towards the shadowed level
    assert abs(out.mean().item() - 0.5) < abs(1.0 - 0.5)


@pytest.mark.parametrize("swap", [False, True])
def test_single_direction_gan_gradients_and_registry(swap):
    """gan_x2y / gan_y2x (gan/wrappers/gan_wrapper.py): tfgan's default Wasserstein losses."""
    from types import SimpleNamespace
    from hypelcnn_b200.gan.wrapper_registry import get_wrapper
    flags = SimpleNamespace(cycle_consistency_loss_weight=10.0, identity_loss_weight=0.5, use_identity_loss=True,
                            discriminator_reg_scale=1e-3)
    wrapper = get_wrapper("gan_y2x" if swap else "gan_x2y", flags)
    x, y = _data(40, 64)
    t = wrapper.define_model(x.view(40, 1, 1, 64), y.view(40, 1, 1, 64))
    rng = numpy.random.default_rng(5)
    t.gen_params.copy_(torch.tensor(rng.standard_normal(t.gen_params.numel()).astype(numpy.float32) * 0.05))
    inp, real = (y, x) if swap else (x, y)
    lg = t.generator_gradients(x, y).cpu()
    G = t.G().double().cpu().requires_grad_(True)
    D = t.DY().double().cpu().requires_grad_(True)
    ref_g = -R.t_discriminator(R.t_generator(inp.double().cpu(), G), D.detach()).mean()
    ref_g.backward()
    assert abs(lg[0].item() - ref_g.item()) < 1e-5 * max(1.0, abs(ref_g.item()))
    assert _rel(t.gen_grads[:t.ng], G.grad) < 2e-4
    ld = t.discriminator_gradients(x, y, use_pool=False).cpu()
    gen = R.t_generator(inp.double().cpu(), G.detach())
    C = 64
    ref_d = R.t_discriminator(gen, D).mean() - R.t_discriminator(real.double().cpu(), D).mean() + \
        1e-3 * 0.5 * ((D[:C * C] ** 2).sum() + (D[C * C + C:2 * C * C + C] ** 2).sum())
    ref_d.backward()
    assert abs(ld[0].item() - ref_d.item()) < 1e-5 * max(1.0, abs(ref_d.item()))
    assert _rel(t.dis_grads[:t.nd], D.grad) < 2e-4
    ops = wrapper.define_train_ops(t, wrapper.define_loss(t), 100, generator_lr=2e-4, discriminator_lr=1e-4)
    ops.train_iteration(x, y)
    assert t.global_step == 1 and t.gen_steps == 1 and t.dis_steps == 1
    with pytest.raises(KeyError):
        get_wrapper("no_such_gan", flags)


@pytest.mark.parametrize("bands,n,identity", [(64, 32, True), (64, 301, True), (32, 37, False), (16, 5, True)])
def test_fused_step_kernels_equal_the_per_op_chain(bands, n, identity):
    """hyp_gan_cycle_generator_step / hyp_gan_cycle_discriminator_step (one kernel per train op) against the chain of
    per-op kernels that the oracle tests pin: same losses, same gradients, same generated spectra."""
    x, y = _data(n, bands, seed=4)
    t = _trainer(bands, pool_size=0, use_identity_loss=identity)
    assert t.use_fused
    want_loss = t.generator_gradients(x, y).clone()
    want_grads, want_last = t.gen_grads.clone(), {k: v.clone() for k, v in t.last.items()}
    got_loss = t.generator_gradients_fused(x, y)
    assert torch.allclose(got_loss, want_loss, rtol=2e-5, atol=1e-7), (got_loss, want_loss)
    assert _rel(t.gen_grads, want_grads.double().cpu()) < 2e-5
    for key, value in want_last.items():
        assert torch.allclose(t.last[key], value, rtol=1e-6, atol=1e-7), key
    want_loss = t.discriminator_gradients(x, y, use_pool=False).clone()
    want_grads = t.dis_grads.clone()
    got_loss = t.discriminator_gradients_fused(x, y, use_pool=False)
    assert torch.allclose(got_loss, want_loss, rtol=2e-5, atol=1e-7), (got_loss, want_loss)
    assert _rel(t.dis_grads, want_grads.double().cpu()) < 2e-5


def test_device_tensor_pool_follows_the_host_pool():
    """The fused discriminator step keeps tfgan's tensor pool on the device; the host draws (mode, slot) exactly like
    TensorPool draws: over a run with changing inputs both paths see the same fakes, i.e. produce the same gradients."""
    eager, fused = _trainer(pool_size=3), _trainer(pool_size=3)
    assert fused.use_fused
    used_pool = False
    for step in range(12):
        x, y = _data(32, 64, seed=10 + step)
        want = eager.discriminator_gradients(x, y).clone()
        got = fused.discriminator_gradients_fused(x, y)
        assert torch.allclose(got, want, rtol=2e-5, atol=1e-7), step
        assert _rel(fused.dis_grads, eager.dis_grads.double().cpu()) < 2e-5, step
        used_pool = used_pool or fused.dev_pool_y.filled == 3
    assert used_pool


def test_fused_train_ops_train_like_the_chain():
    x, y = _data(32, 64, seed=4)
    chain, fused = _trainer(pool_size=0), _trainer(pool_size=0)
    chain.use_fused = False
    for step in range(1, 5):
        x2, y2 = x * (1.0 + 0.01 * step), y * (1.0 - 0.01 * step)
        assert torch.allclose(chain.generator_train_op(x2, y2, 2e-4), fused.generator_train_op(x2, y2, 2e-4), rtol=1e-4, atol=1e-7)
        assert torch.allclose(chain.discriminator_train_op(x2, y2, 1e-4), fused.discriminator_train_op(x2, y2, 1e-4), rtol=1e-4, atol=1e-7)
    assert _rel(fused.gen_params, chain.gen_params.double().cpu()) < 1e-4
    assert _rel(fused.dis_params, chain.dis_params.double().cpu()) < 1e-4
