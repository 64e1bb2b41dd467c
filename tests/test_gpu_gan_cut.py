"""GPU parity of the CUT / DCLGAN / DCL-CycleGAN training path (gan/wrappers/cut_wrapper.py:90-208,256-420,467-584,
gan/wrappers/dcl_gan_wrapper.py, gan/wrappers/dcl_cycle_gan_wrapper.py, gan/shadow_data_models.py:126-149) against the
torch-autograd restatement in oracle/gan_ref.py: patch feature discriminator forward / backward, PatchNCE loss and
both of its gradient conventions, the encoder-gradient entry of the generator backward, the three CUT train-op
gradients, and the train-op sequencing / shared optimizer clock of DCLGAN."""
import numpy
import pytest
import torch

from oracle import gan_ref as R

pytestmark = pytest.mark.gpu


def _trainer(bands=64, seed=3, **kw):
    from hypelcnn_b200.gan.wrappers.cut_wrapper import CUTTrainer
    t = CUTTrainer(bands, seed=seed, **kw)
    rng = numpy.random.default_rng(seed)
    t.gen_params.copy_(torch.tensor(rng.standard_normal(t.gen_params.numel()).astype(numpy.float32) * 0.05))
    t.dis_params.add_(torch.tensor(rng.standard_normal(t.dis_params.numel()).astype(numpy.float32) * 0.01).cuda())
    bias = torch.tensor(rng.standard_normal(t.feat_params.numel()).astype(numpy.float32) * 0.05).cuda()
    mask = torch.zeros_like(t.feat_params)
    for name, off, shape in t.feat_table:        # biases are zero-initialised: give them values, keep padding zero
        if name.endswith("biases"):
            mask[off:off + shape[0]] = 1
    t.feat_params.add_(bias * mask)
    return t


def _data(n, bands, seed=1):
    rng = numpy.random.default_rng(seed)
    y = rng.uniform(0.02, 0.5, (n, bands)).astype(numpy.float32)
    x = (y * numpy.linspace(1.5, 4, bands)).astype(numpy.float32)
    return torch.tensor(x).cuda(), torch.tensor(y).cuda()


def _rel(got, ref):
    return float((got.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("bands,patches,E,n", [(64, 6, 2, 300), (144, 6, 2, 129), (48, 4, 5, 7), (32, 8, 8, 128)])
def test_feature_discriminator_forward_backward_match_autograd(bands, patches, E, n):
    t = _trainer(bands, patches=patches, embedded_feat_size=E)
    x, _ = _data(n, bands)
    z, ss = t._feat_fwd(x)
    xr = x.double().cpu().requires_grad_(True)
    wr = t.feat_params.double().cpu().requires_grad_(True)
    f_ref = R.t_feature_discriminator(xr, wr, patches, E)
    f = z * torch.rsqrt(torch.clamp(ss, min=1e-12)).view(1, -1, 1)
    assert f.shape == f_ref.shape and _rel(f, f_ref.detach()) < 2e-5
    assert _rel(t.feature_embeddings(x), R.t_feature_discriminator(
        R.t_generator(xr.detach(), t.gen_params.double().cpu(), True), wr.detach(), patches, E)) < 5e-5
    gf = torch.randn_like(z)
    dot = (gf * z).sum(dim=(0, 2)).contiguous()
    gw = torch.zeros_like(t.feat_params)
    gin = t._feat_bwd(x, z, ss, gf, dot, True, gw)
    (f_ref * gf.double().cpu()).sum().backward()
    assert _rel(gin, xr.grad) < 2e-4 and _rel(gw, wr.grad) < 2e-4


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("S,E,n", [(7, 2, 200), (6, 2, 33), (16, 8, 5), (2, 1, 4)])
def test_patchnce_loss_and_gradients(S, E, n, fused):
    t = _trainer(64, fused_xent_grad=fused)
    t.slices, t.E = S, E
    rng = numpy.random.default_rng(7)
    zg = torch.tensor(rng.standard_normal((n, S, E)).astype(numpy.float32)).cuda()
    zr = torch.tensor(rng.standard_normal((n, S, E)).astype(numpy.float32)).cuda()
    ssg, ssr = (zg * zg).sum(dim=(0, 2)).contiguous(), (zr * zr).sum(dim=(0, 2)).contiguous()
    t.loss_acc.zero_()
    gg, gr, dg, dr = t._nce(zg, ssg, zr, ssr, 10.0, 2)
    fg = (zg.double().cpu() / ssg.double().cpu().sqrt().view(1, -1, 1)).requires_grad_(True)
    fr = (zr.double().cpu() / ssr.double().cpu().sqrt().view(1, -1, 1)).requires_grad_(True)
    ref = 10.0 * R.t_patchnce(fg, fr, t.tau, fused)
    ref.backward()
    assert abs(t.loss_acc[2].item() - ref.item()) < 2e-5 * max(1.0, abs(ref.item()))
    assert _rel(gg, fg.grad) < 2e-4 and _rel(gr, fr.grad) < 2e-4
    assert _rel(dg, (fg.grad * zg.double().cpu()).sum(dim=(0, 2))) < 2e-3
    assert _rel(dr, (fr.grad * zr.double().cpu()).sum(dim=(0, 2))) < 2e-3


@pytest.mark.parametrize("with_out", [True, False])
def test_generator_backward_with_encoder_gradient(with_out):
    t = _trainer(64)
    x, _ = _data(150, 64)
    nets = t._gen_fwd(x, t.gen_params)
    gout = torch.randn((150, 64), device="cuda") if with_out else None
    genc = torch.randn((150, 64), device="cuda")
    t.gen_grads.zero_()
    gin = t._gen_bwd_enc(nets, gout, genc, True)
    xr = x.double().cpu().requires_grad_(True)
    wr = t.gen_params.double().cpu().requires_grad_(True)
    enc = R.t_generator(xr, wr, True)
    assert _rel(nets[:, 4, :], enc.detach()) < 1e-5
    obj = (enc * genc.double().cpu()).sum()
    if with_out:
        obj = obj + (R.t_generator(xr, wr) * gout.double().cpu()).sum()
    obj.backward()
    assert _rel(gin, xr.grad) < 1e-4 and _rel(t.gen_grads, wr.grad) < 1e-4


@pytest.mark.parametrize("bands,swap,fused,use_id", [(64, False, True, True), (64, True, False, True),
                                                     (96, False, True, True), (64, False, True, False)])
def test_cut_train_op_gradients_match_the_cut_objective(bands, swap, fused, use_id):
    t = _trainer(bands, swap_inputs=swap, fused_xent_grad=fused, use_identity_loss=use_id,
                 discriminator_reg_scale=1e-3, gen_disc_reg_scale=1e-2)
    x, y = _data(96, bands)
    inp, real = (y, x) if swap else (x, y)
    lg = t.generator_gradients(x, y).cpu()
    ld = t.discriminator_gradients(x, y).cpu()
    lf = t.feat_discriminator_gradients(x, y).cpu()
    G, D, Fd = (p.double().cpu().requires_grad_(True) for p in (t.gen_params, t.dis_params, t.feat_params))
    w_id = 0.5 if use_id else 0.0
    gen_loss, dis_loss, feat_loss, parts = R.t_cut_losses(inp.double().cpu(), real.double().cpu(), G, D, Fd, t.patches,
                                                          t.E, t.tau, 10.0, w_id, 1e-3, 1e-2, fused)
    gG, = torch.autograd.grad(gen_loss, G, retain_graph=True)
    gD, = torch.autograd.grad(dis_loss, D, retain_graph=True)
    gF, = torch.autograd.grad(feat_loss, Fd)
    for got, ref in zip(lg.tolist(), (gen_loss.item(), parts["gan"].item(), 10.0 * parts["nce_x"].item(),
                                      w_id * parts["nce_identity"].item())):
        assert abs(got - ref) < 2e-5 * max(1.0, abs(ref))
    assert abs(ld[0].item() - dis_loss.item()) < 2e-5 * max(1.0, abs(dis_loss.item()))
    assert abs(lf[0].item() - feat_loss.item()) < 2e-5 * max(1.0, abs(feat_loss.item()))
    assert _rel(t.gen_grads, gG) < 5e-4
    assert _rel(t.dis_grads, gD) < 2e-4
    assert _rel(t.feat_grads, gF) < 5e-4


def test_cut_wrapper_trains_and_infers():
    from types import SimpleNamespace
    from hypelcnn_b200.gan.wrapper_registry import get_infer_wrapper_dict, get_wrapper
    from hypelcnn_b200.gan.wrappers.cut_wrapper import CUTInferenceWrapper
    flags = SimpleNamespace(cycle_consistency_loss_weight=10.0, identity_loss_weight=0.5, use_identity_loss=True,
                            nce_loss_weight=10.0, tau=0.07, patches=6, embedded_feat_size=2, batch_size=64)
    wrapper = get_wrapper("cut_x2y", flags)
    x, y = _data(64, 64)
    model = wrapper.define_model(x.view(64, 1, 1, 64), y.view(64, 1, 1, 64))
    ops = wrapper.define_train_ops(model, wrapper.define_loss(model), 100, generator_lr=2e-4, discriminator_lr=1e-4,
                                   gen_discriminator_lr=1e-4)
    before = {k: v.copy() for k, v in model.trainer.variables().items()}
    assert "FeatDiscriminator/fully_connected_27/weights" in before and "Generator/net7/weights" in before
    for _ in range(5):
        lg, ld, lf = ops.train_iteration(x, y)
    t = model.trainer
    assert t.clock == {"global_step": 5, "gen": 5, "dis": 5, "feat": 5}
    assert all(torch.isfinite(l).all() for l in (lg, ld, lf))
    after = t.variables()
    for key in ("Generator/net1/weights", "Discriminator/fully_connected/weights", "FeatDiscriminator/fully_connected/weights"):
        assert numpy.abs(after[key] - before[key]).max() > 0
    infer = CUTInferenceWrapper(False, trainer=t)
    out = infer.construct_inference_graph(x.view(64, 1, 1, 64), True, False)
    assert out.shape == (64, 1, 1, 64)
    assert _rel(out.view(64, 64), R.t_generator(x.double().cpu(), t.gen_params.double().cpu())) < 1e-5
    assert set(get_infer_wrapper_dict(64)) == {"cycle_gan", "gan_x2y", "gan_y2x", "cut_x2y", "cut_y2x", "dcl_gan",
                                               "dcl_cycle_gan"}


@pytest.mark.parametrize("gan_type", ["dcl_gan", "dcl_cycle_gan"])
def test_dcl_wrappers_are_two_cut_models_on_one_optimizer_clock(gan_type):
    from types import SimpleNamespace
    from hypelcnn_b200.gan.wrapper_registry import get_wrapper
    from hypelcnn_b200.gan.wrappers.cut_wrapper import CUTTrainer
    from hypelcnn_b200.gan.wrappers.cycle_gan_wrapper import CycleGANInferenceWrapper
    flags = SimpleNamespace(cycle_consistency_loss_weight=10.0, identity_loss_weight=0.5, use_identity_loss=True,
                            nce_loss_weight=10.0, tau=0.07, patches=6, embedded_feat_size=2, batch_size=48)
    wrapper = get_wrapper(gan_type, flags)
    x, y = _data(48, 64)
    model = wrapper.define_model(x.view(48, 1, 1, 64), y.view(48, 1, 1, 64))
    ops = wrapper.define_train_ops(model, wrapper.define_loss(model), 100, generator_lr=2e-4, discriminator_lr=1e-4,
                                   gen_discriminator_lr=1e-4)
    tr = wrapper.trainer
    # an independent CUT model with the same initial variables takes the same first x2y generator step
    solo = CUTTrainer(64, swap_inputs=False, seed=1234)
    for a, b in ((solo.dis_params, tr.model_x2y.dis_params), (solo.feat_params, tr.model_x2y.feat_params)):
        assert torch.equal(a, b)
    losses = ops.train_iteration(x, y)
    assert len(losses) == 6 and all(torch.isfinite(l).all() for l in losses)
    assert tr.clock == {"global_step": 1, "gen": 2, "dis": 2, "feat": 2}      # shared AdamOptimizer objects
    solo.generator_train_op(x, y, 2e-4)
    assert torch.allclose(solo.gen_params, tr.model_x2y.gen_params, atol=1e-7)
    assert not torch.equal(tr.model_x2y.gen_params, tr.model_y2x.gen_params)
    assert len(wrapper.get_train_hooks_fn()(ops)) == 6
    names = tr.variables()
    assert "ModelX2Y/Generator/net1/weights" in names and "ModelY2X/FeatDiscriminator/fully_connected_3/biases" in names
    infer = CycleGANInferenceWrapper(trainer=tr)
    assert infer.forward_generator is tr.model_x2y.generator and infer.backward_generator is tr.model_y2x.generator
    if gan_type == "dcl_cycle_gan":
        rx, ry = wrapper.reconstructions(x, y)
        Gx, Gy = tr.model_x2y.gen_params.double().cpu(), tr.model_y2x.gen_params.double().cpu()
        assert _rel(rx, R.t_generator(R.t_generator(x.double().cpu(), Gx), Gy)) < 1e-5
        assert _rel(ry, R.t_generator(R.t_generator(y.double().cpu(), Gy), Gx)) < 1e-5


def test_dclgan_generators_feed_the_classifier_augmenter():
    """BASELINE configs[4]'s data path: --augment_data_with_shadow=dcl_gan hands the DCLGAN generators to the
    classifier's input pipeline through create_gan_struct (gan/gan_utilities.py:30-43)."""
    from hypelcnn_b200.gan.gan_utilities import create_gan_struct
    from hypelcnn_b200.gan.wrappers.dcl_gan_wrapper import DCLGANInferenceWrapper, DCLGANTrainer
    bands, n, P = 64, 24, 3
    tr = DCLGANTrainer(bands)
    x, y = _data(64, bands)
    from hypelcnn_b200.gan.wrappers.dcl_gan_wrapper import DCLGANTrainOps
    ops = DCLGANTrainOps(tr, 100, 2e-3, 1e-4, 1e-4)
    for _ in range(3):
        ops.train_iteration(x, y)
    struct = create_gan_struct(DCLGANInferenceWrapper(trainer=tr))
    rng = numpy.random.default_rng(5)
    patches = torch.tensor(rng.uniform(0.1, 0.6, (n, P, P, bands + 1)).astype(numpy.float32)).cuda()
    for op, gen in ((struct.shadow_op, tr.model_x2y.generator), (struct.deshadow_op, tr.model_y2x.generator)):
        out = op(patches)
        ref = R.inference_for_matrix_input(patches.double().cpu().numpy(), {k: v.astype(numpy.float64) for k, v in
                                                                            gen.export().items()}, True, False, 1)
        assert out.shape == patches.shape and torch.equal(out[..., -1], patches[..., -1])
        assert numpy.abs(out.cpu().numpy() - ref).max() < 1e-5
        assert float((out[..., :-1] - patches[..., :-1]).abs().max()) > 0      # the trained generator does something
