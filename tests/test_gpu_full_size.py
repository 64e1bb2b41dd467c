"""Parity at BASELINE.json's full size (C2: 4096 patches of 7x7x145, 15 classes, the bench workload), where the fp64
oracle with autograd no longer fits a few seconds: the forward pass against the fp32/fp64 CPU oracle without autograd,
and size-independent properties for everything else —
  * rows are independent in eval mode: the 4096-patch call equals eight 512-patch calls bit for bit;
  * a batch made of the same 2048 patches twice has the batch statistics of the half batch, so its logits repeat, and
    its loss and every gradient equal those of the 2048-patch step (the mean loss halves each sample's weight, every
    sample appears twice) — this walks the full-size tilings (32 batch tiles, wgrad K slices, CTA pairs) against the
    half-size ones."""
import numpy
import pytest
import torch

from oracle import dataset_ref as D
from oracle import hypelcnn_ref as R
from tests.util import ALG, ATOL, RTOL, assert_close, oracle_variables, synthetic_batch

pytestmark = pytest.mark.gpu

P, C, CLASSES, B = 7, 145, 15, 4096
ALG0 = {**ALG, "drop_out_ratio": 0.0}


def _engine(max_batch):
    from hypelcnn_b200 import engine as E
    eng = E.PatchEngine(P, C, CLASSES, ALG0, max_batch=max_batch, precision="3xf16")
    eng.init_variables(seed=1234)
    return E, eng


def test_full_size_forward_matches_the_oracle():
    E, eng = _engine(B)
    x, y = synthetic_batch(B, P, C, CLASSES)
    xd = torch.as_tensor(x).cuda()
    logits, recon = eng.forward(xd, True, True, seed=0)
    with torch.no_grad():
        ref32 = R.forward(oracle_variables(eng, torch.float32), torch.tensor(x), CLASSES, ALG0, True)
        sub = slice(0, 4096, 16)    # fp64 needs the full batch for the BN statistics; it fits without autograd
        ref64 = R.forward(oracle_variables(eng), torch.tensor(x, dtype=torch.float64), CLASSES, ALG0, True)
    floor32 = float((ref32["logits"].double() - ref64["logits"]).abs().max())
    assert_close(logits.cpu().numpy(), ref64["logits"].numpy(), RTOL, max(ATOL, 3.0 * floor32), "logits")
    assert_close(recon.cpu().numpy()[sub], ref64["recon"].numpy()[sub], RTOL, ATOL, "recon")
    pred = E.argmax_confusion(logits).cpu().numpy()
    ref_pred = D.argmax_lowest(ref64["logits"].numpy()).astype(numpy.uint8)
    # bit-exact class map wherever the fp64 top-2 margin exceeds the fp32 oracle's own error
    top2 = numpy.sort(ref64["logits"].numpy(), axis=1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 4.0 * max(floor32, ATOL)
    assert decided.mean() > 0.99 and numpy.array_equal(pred[decided], ref_pred[decided])
    per = eng.per_sample_loss(logits, recon, xd, torch.as_tensor(y).cuda()).cpu().numpy()
    ref_per = R.per_sample_loss(ref64["logits"], ref64["recon"], torch.tensor(x, dtype=torch.float64),
                                torch.tensor(y.astype(numpy.int64))).numpy()
    assert_close(per, ref_per, RTOL, ATOL, "per-sample loss")


def test_eval_rows_are_independent_of_the_batch_they_travel_in():
    E, eng = _engine(B)
    x, _ = synthetic_batch(B, P, C, CLASSES, seed=5)
    xd = torch.as_tensor(x).cuda()
    for _ in range(2):                                  # non-trivial moving statistics
        eng.forward(xd, True, True, seed=0)
    full, _ = eng.forward(xd, False)
    full = full.clone()
    parts = torch.cat([eng.forward(xd[i:i + 512].contiguous(), False)[0].clone() for i in range(0, B, 512)])
    assert torch.equal(full, parts)
    ragged = torch.cat([eng.forward(xd[:1000].contiguous(), False)[0].clone(),
                        eng.forward(xd[1000:].contiguous(), False)[0].clone()])
    assert torch.equal(full, ragged)


def test_doubled_batch_equals_the_half_batch_step():
    E, big = _engine(B)
    _, half = _engine(B // 2)
    h, yh = synthetic_batch(B // 2, P, C, CLASSES, seed=11)
    x2, y2 = numpy.concatenate([h, h]), numpy.concatenate([yh, yh])
    xd2, yd2 = torch.as_tensor(x2).cuda(), torch.as_tensor(y2).cuda()
    xdh, ydh = torch.as_tensor(h).cuda(), torch.as_tensor(yh).cuda()
    l2, _ = big.forward(xd2, True, True, seed=0)
    lh, _ = half.forward(xdh, True, True, seed=0)
    l2 = l2.cpu()
    assert torch.equal(l2[:B // 2], l2[B // 2:])                       # duplicate rows, same statistics
    scale = float(lh.abs().max())
    assert float((l2[:B // 2] - lh.cpu()).abs().max()) <= 2e-5 * scale   # statistics summed in another order
    loss2, lossh = big.loss_backward(xd2, yd2).cpu().numpy(), half.loss_backward(xdh, ydh).cpu().numpy()
    assert numpy.allclose(loss2, lossh, rtol=2e-6, atol=1e-7)
    worst = 0.0
    for name in big.variables:
        if "moving_" in name:
            continue
        g2, gh = big.gradient(name).cpu().double(), half.gradient(name).cpu().double()
        rel = float((g2 - gh).abs().max() / gh.abs().max().clamp_min(1e-30))
        worst = max(worst, rel)
        assert rel < 2e-3, (name, rel)
    assert worst > 0.0   # different tilings did run (bit-identical gradients would mean the same code path)


def test_large_batch_gradients_match_the_oracle():
    """Every variable gradient of a 1024-patch C2 step against autograd of the CPU oracle (fp64, with the fp32 oracle as
    the yardstick of what fp32 arithmetic can deliver): the multi-tile plans of the bench size — 8 batch tiles per
    position, K-split wgrads, CTA pairs — checked against something other than the engine itself."""
    from tests.test_gpu_parity import CASES, engine_lrelu_gates
    from tests.util import assert_grad_close
    Bg = 1024
    E, eng = _engine(Bg)
    x, y = synthetic_batch(Bg, P, C, CLASSES, seed=21)
    xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
    eng.forward(xd, True, True, seed=0)
    case = dict(CASES["c2"], B=Bg)
    gates = engine_lrelu_gates(eng, case, ALG0)
    yt = torch.tensor(y.astype(numpy.int64))
    loss_ref, g_ref, _ = R.loss_and_grads(oracle_variables(eng), torch.tensor(x, dtype=torch.float64), yt, CLASSES, ALG0,
                                          lrelu_gates=gates)
    _, g_ref32, _ = R.loss_and_grads(oracle_variables(eng, torch.float32), torch.tensor(x), yt, CLASSES, ALG0,
                                     lrelu_gates=gates)
    loss = eng.loss_backward(xd, yd).cpu().numpy()
    assert abs(loss[0] - loss_ref.item()) <= RTOL * abs(loss_ref.item()) + ATOL
    for name in reversed([n for n in eng.variables if "moving_" not in n]):
        assert_grad_close(eng.gradient(name).cpu().numpy(), g_ref[name].numpy(), g_ref32[name].numpy(), f"grad {name}")
