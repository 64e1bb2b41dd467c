"""CPU ORACLE (test infrastructure, NOT product code) — shadow GAN models and objectives.

numpy: the generator forward (gan/shadow_data_models.py:43-90; float64 or float32): seven slim convolution1d layers
with one filter (SAME padding: total pad K-1, left = (K-1)//2; bias; leaky_relu 0.1), the dense
residual pattern net_i = conv(net_{i-1}) + net_{i-1} + net_{i-2} (net1: + net0 only), tanh and no residual on net7;
encoder-only stops after net4 (:75); create_inference_for_matrix_input (gan/wrappers/gan_common.py:282-304).
torch (autograd, for the gradient parity tests): the generator, the discriminator (:93-123), the patch feature
discriminator (:126-149), the CycleGAN-with-identity objective (gan/wrappers/cycle_gan_wrapper.py:189-333), PatchNCE
and the CUT objective (gan/wrappers/cut_wrapper.py:90-208,256-420) including TensorFlow's fused softmax-xent gradient
convention.  Parity with TensorFlow is UNPINNED (no TF in this image): the [TF-lib] semantics follow SURVEY App. A;
the reference's DummySampler fixture (gan/gan_sampling_methods.py:191-201) gives the only known answers
(tests/test_gpu_gan.py).  Only tests/ imports this module."""
import numpy


def kernel_sizes(bands, encoder_only=False):
    k = bands
    s = [k, k // 2, k // 4, k // 8]
    return s if encoder_only else s + [k // 4, k // 2, k]


def conv1d_same(x, w, b):
    """x [N,C], one filter w [K], SAME zero padding, stride 1."""
    K, C = len(w), x.shape[1]
    left = (K - 1) // 2
    xp = numpy.zeros((x.shape[0], C + K - 1), dtype=x.dtype)
    xp[:, left:left + C] = x
    out = numpy.full(x.shape, b, dtype=x.dtype)
    for t in range(K):
        out += w[t] * xp[:, t:t + C]
    return out


def generator_forward(x, variables, encoder_only=False):
    """x [N,C]; variables {"net1/weights": [K,1,1], "net1/biases": [1], ...} -> [N,C]."""
    dt = x.dtype
    nets = [x]
    sizes = kernel_sizes(x.shape[1], encoder_only)
    for i, k in enumerate(sizes):
        w = numpy.asarray(variables[f"net{i + 1}/weights"], dtype=dt).reshape(-1)
        b = dt.type(numpy.asarray(variables[f"net{i + 1}/biases"]).reshape(-1)[0])
        assert len(w) == k
        y = conv1d_same(nets[-1], w, b)
        if i == 6:
            y = numpy.tanh(y)
        else:
            y = numpy.maximum(y, dt.type(0.1) * y) + nets[-1]
            if i > 0:
                y = y + nets[-2]
        nets.append(y)
    return nets[-1]


def inference_for_matrix_input(x, variables, is_shadow, clip, copy_extra=0):
    """create_inference_for_matrix_input (gan/wrappers/gan_common.py:282-304) + LiDAR pass-through
    (gan/gan_utilities.py:31-35) on [B,H,W,C(+extra)]."""
    C = x.shape[3] - copy_extra
    rows = x.reshape(-1, x.shape[3])
    gen = generator_forward(rows[:, :C].copy(), variables)
    if clip:
        gm, im = gen.mean(axis=1), rows[:, :C].mean(axis=1)
        keep = (gm < im) if is_shadow else (gm > im)
        gen = numpy.where(keep[:, None], gen, rows[:, :C])
    out = rows.copy()
    out[:, :C] = gen
    return out.reshape(x.shape)


# ---------------------------------------------------------------------------------------------------------------
# torch (autograd) restatement of the CycleGAN-with-identity objective for the gradient parity tests
def t_generator(x, w, encoder_only=False):
    """x [N,C] torch, w flat [net1 w, b, net2 w, b, ...] -> net7 (net4 when encoder_only, shadow_data_models.py:75)."""
    import torch
    import torch.nn.functional as F
    C = x.shape[1]
    nets, off = [x], 0
    for i, k in enumerate(kernel_sizes(C, encoder_only)):
        wk, b = w[off:off + k], w[off + k]
        off += k + 1
        left = (k - 1) // 2
        xp = F.pad(nets[-1], (left, k - 1 - left))
        y = F.conv1d(xp.unsqueeze(1), wk.view(1, 1, k)).squeeze(1) + b
        if i == 6:
            y = torch.tanh(y)
        else:
            y = torch.maximum(y, 0.1 * y) + nets[-1]
            if i > 0:
                y = y + nets[-2]
        nets.append(y)
    return nets[-1]


def t_discriminator(x, w):
    import torch
    C = x.shape[1]
    H = C // 2
    o = 0
    W1 = w[o:o + C * C].view(C, C); o += C * C
    b1 = w[o:o + C]; o += C
    W2 = w[o:o + C * C].view(C, C); o += C * C
    b2 = w[o:o + C]; o += C
    W3 = w[o:o + C * H].view(C, H); o += C * H
    b3 = w[o:o + H]
    h = x @ W1 + b1
    h = torch.maximum(h, 0.1 * h)
    h = h @ W2 + b2
    h = torch.maximum(h, 0.1 * h)
    return h @ W3 + b3


def t_generator_loss(x, y, G, Fw, DY, DX, w_cyc, w_id):
    """L_G of cyclegan_loss_with_identity summed over both partial models (aux counted twice)."""
    gx, fy = t_generator(x, G), t_generator(y, Fw)
    rx, ry = t_generator(gx, Fw), t_generator(fy, G)
    gan = 0.5 * ((t_discriminator(gx, DY) - 1) ** 2).mean() + 0.5 * ((t_discriminator(fy, DX) - 1) ** 2).mean()
    cyc = ((rx - x).abs().mean() + (ry - y).abs().mean()) / 2
    ident = (gx - x).abs().mean() + (fy - y).abs().mean()
    return gan + 2 * (w_cyc * cyc + w_id * ident), gan, 2 * w_cyc * cyc, 2 * w_id * ident


def t_discriminator_loss(x, y, gx, fy, DY, DX, reg):
    C = x.shape[1]
    total = 0
    for real, fake, w in ((y, gx, DY), (x, fy, DX)):
        total = total + 0.5 * ((t_discriminator(real, w) - 1) ** 2).mean() + 0.5 * (t_discriminator(fake, w) ** 2).mean()
        total = total + reg * 0.5 * ((w[:C * C] ** 2).sum() + (w[C * C + C:2 * C * C + C] ** 2).sum())
    return total


# ---------------------------------------------------------------------------------------------------------------
# CUT / DCLGAN (gan/shadow_data_models.py:126-149, gan/wrappers/cut_wrapper.py:90-208,256-420)
def feature_discriminator_layout(bands, patch_count, E):
    """[(slice start, in width, [(w offset, (n_in, n_out), b offset), ...])] inside the flat weight buffer; every slice
    occupies the size of a full slice (ps wide) so that slice s starts at s * n_full."""
    ps = bands // patch_count
    dims = [ps, ps, ps // 4, ps // 2, E]
    n_full = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(4))
    out = []
    for s, start in enumerate(range(0, bands, ps)):
        width = min(ps, bands - start)
        d = [width] + dims[1:]
        off, layers = s * n_full, []
        for i in range(4):
            layers.append((off, (d[i], d[i + 1]), off + d[i] * d[i + 1]))
            off += d[i] * d[i + 1] + d[i + 1]
        out.append((start, width, layers))
    return out, n_full


def t_feature_discriminator(x, w, patch_count, E):
    """x [N,C] -> [N, slices, E]: per slice 4 FC + leaky_relu(0.1) and tf.math.l2_normalize over the WHOLE [N,E]
    output of the slice (axis=None, epsilon 1e-12)."""
    import torch
    layout, _ = feature_discriminator_layout(x.shape[1], patch_count, E)
    outs = []
    for start, width, layers in layout:
        h = x[:, start:start + width]
        for wo, (ni, no), bo in layers:
            h = h @ w[wo:wo + ni * no].view(ni, no) + w[bo:bo + no]
            h = torch.maximum(h, 0.1 * h)
        outs.append((h * torch.rsqrt(torch.clamp((h * h).sum(), min=1e-12))).unsqueeze(1))
    return torch.cat(outs, dim=1)


def t_feature_discriminator_reg(w, bands, patch_count, E, scale):
    """slim l2_regularizer(scale) on every FC weight matrix of the feature discriminator (:128-129)."""
    layout, _ = feature_discriminator_layout(bands, patch_count, E)
    total = 0
    for _, _, layers in layout:
        for wo, (ni, no), _ in layers:
            total = total + scale * 0.5 * (w[wo:wo + ni * no] ** 2).sum()
    return total


def _fused_xent():
    import torch

    class FusedXent(torch.autograd.Function):
        """tf.nn.softmax_cross_entropy_with_logits as TensorFlow's fused kernel evaluates it [TF-lib]:
        loss = -sum(labels * log_softmax(logits)), backprop = softmax(logits) - labels (exact only when the labels
        sum to one; the reference's eye() labels sum to `slices`)."""

        @staticmethod
        def forward(ctx, logits, labels):
            ls = torch.log_softmax(logits, dim=1)
            ctx.save_for_backward(ls.exp() - labels)
            return -(labels * ls).sum(dim=1)

        @staticmethod
        def backward(ctx, g):
            (bp,) = ctx.saved_tensors
            return g.unsqueeze(1) * bp, None

    return FusedXent


def t_patchnce(f_gen, f_real, tau, fused_grad=True):
    """_calc_cross_feats + softmax CE + SUM_OVER_BATCH_SIZE (cut_wrapper.py:360-393)."""
    import torch
    B, S, _ = f_gen.shape
    logits = (f_gen @ f_real.transpose(1, 2) / tau).reshape(B, S * S)
    labels = torch.eye(S, dtype=f_gen.dtype).reshape(1, S * S).expand(B, S * S)
    if fused_grad:
        return _fused_xent().apply(logits, labels).mean()
    return -(labels * torch.log_softmax(logits, dim=1)).sum(dim=1).mean()


def t_cut_losses(inp, real, G, D, Fd, patch_count, E, tau, nce_w, id_w, dis_reg, feat_reg, fused_grad=True):
    """cut_model + cut_loss with the LSGAN losses of CUTWrapper.define_loss (cut_wrapper.py:626-636).  Returns
    (generator_loss, discriminator_loss, gen_discriminator_loss, parts)."""
    C = inp.shape[1]
    gen = t_generator(inp, G)
    d_gen, d_real = t_discriminator(gen, D), t_discriminator(real, D)
    fd = lambda t: t_feature_discriminator(t_generator(t, G, True), Fd, patch_count, E)
    nce_x = t_patchnce(fd(gen), fd(inp), tau, fused_grad)
    idt = t_generator(real, G)
    nce_id = t_patchnce(fd(idt), fd(real), tau, fused_grad)
    gan_g = 0.5 * ((d_gen - 1) ** 2).mean()
    gen_loss = gan_g + nce_w * nce_x + id_w * nce_id
    dis_loss = 0.5 * ((d_real - 1) ** 2).mean() + 0.5 * (d_gen ** 2).mean() + \
        dis_reg * 0.5 * ((D[:C * C] ** 2).sum() + (D[C * C + C:2 * C * C + C] ** 2).sum())
    feat_loss = nce_x + t_feature_discriminator_reg(Fd, C, patch_count, E, feat_reg)
    return gen_loss, dis_loss, feat_loss, {"gan": gan_g, "nce_x": nce_x, "nce_identity": nce_id}
