"""CPU test of the pair-tile plan for level forward launches (hyp_tc_engine.cuh plan_level_pairs — groundwork for the
next kernel step): the plan is replayed in numpy exactly as the GEMM kernel would execute it (one
matrix product per segment into a column range of a 2W-wide accumulator, B rows taken from a mirrored and a normal
packed weight copy) and must reproduce the SAME-padded k x k convolutions of the level for every output position."""
import ctypes

import numpy
import pytest

from hypelcnn_b200 import _native as N


@pytest.fixture(scope="module")
def lib():
    N.build_native()
    return N.lib()


def plan(lib, P, R, fpad):
    tiles = numpy.zeros((P * P, 4), numpy.int32)
    segs = numpy.zeros((P * P * P * P, 6), numpy.int32)
    counts = numpy.zeros(2, numpy.int32)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    N.check(lib.hyp_debug_plan_level_pairs(P, R, fpad, ptr(tiles), len(tiles), ptr(segs), len(segs), ptr(counts)))
    return tiles[:counts[0]], segs[:counts[1]]


def packed_weights(weights, R, f, fpad, h):
    """normal copy: row (tap * R + slot) * fpad + n, slot 0 = the largest kernel; mirrored copy: slots reversed.
    A tap outside a kernel's support contributes zero rows (as tc_pack_weights_kernel packs them)."""
    TW, W, cin = 2 * h + 1, R * fpad, weights[0].shape[2]
    normal = numpy.zeros((TW * TW * W, cin))
    mirrored = numpy.zeros_like(normal)
    for dy in range(-h, h + 1):
        for dx in range(-h, h + 1):
            tap = (dy + h) * TW + (dx + h)
            for slot in range(R):
                q = R - 1 - slot                                  # kernel of size 2q + 1
                if max(abs(dy), abs(dx)) > q:
                    continue
                rows = weights[q][dy + q, dx + q].T               # [f, cin]
                normal[tap * W + slot * fpad: tap * W + slot * fpad + f] = rows
                mirrored[tap * W + (R - 1 - slot) * fpad: tap * W + (R - 1 - slot) * fpad + f] = rows
    return normal, mirrored


@pytest.mark.parametrize("P,R,f,fpad", [(7, 4, 30, 32), (7, 4, 15, 16), (3, 2, 20, 32), (5, 4, 8, 16), (4, 3, 32, 32),
                                        (1, 1, 16, 16), (2, 4, 16, 32)])
def test_pair_plan_reproduces_the_level_convolutions(lib, P, R, f, fpad):
    rng = numpy.random.default_rng(P * 100 + R)
    cin, B = 12, 5
    h, W = min(R - 1, P - 1), R * fpad
    a = rng.standard_normal((P * P, B, cin))                      # position-major activations
    weights = [rng.standard_normal((2 * q + 1, 2 * q + 1, cin, f)) for q in range(R)]
    normal, mirrored = packed_weights(weights, R, f, fpad, h)
    tiles, segs = plan(lib, P, R, fpad)
    got = numpy.full((P * P, B, R * f), numpy.nan)
    covered = numpy.zeros(P * P, int)
    for p1, p2, s0, ns in tiles:
        acc = numpy.zeros((B, 2 * W))
        widths = []
        for q, n1, n2, brow1, brow2, dcol in segs[s0:s0 + ns]:
            assert dcol == W - n1 and n1 % fpad == 0 and n2 % fpad == 0 and 0 < n1 + n2 <= 2 * W
            assert n2 == 0 or p2 >= 0
            b = numpy.concatenate([mirrored[brow1:brow1 + n1], normal[brow2:brow2 + n2]])
            acc[:, dcol:dcol + n1 + n2] += a[q] @ b.T             # one MMA chain: N = n1 + n2 contiguous columns
            widths.append((n1, n2))
        assert len({int(q) for q in segs[s0:s0 + ns, 0]}) == ns    # every input position at most once per tile
        for slot in range(R):                                     # the epilogue's column blocks
            kq = R - 1 - slot
            got[p1, :, kq * f:(kq + 1) * f] = acc[:, (R - 1 - slot) * fpad:(R - 1 - slot) * fpad + f]
            if p2 >= 0:
                got[p2, :, kq * f:(kq + 1) * f] = acc[:, W + slot * fpad: W + slot * fpad + f]
        covered[p1] += 1
        if p2 >= 0:
            assert p2 == p1 + 1 and p1 // P == p2 // P
            covered[p2] += 1
    assert numpy.all(covered == 1)
    ref = numpy.zeros((P * P, B, R * f))                          # direct SAME convolution per kernel size
    for p in range(P * P):
        for q in range(R):
            for dy in range(-q, q + 1):
                for dx in range(-q, q + 1):
                    y, x = p // P + dy, p % P + dx
                    if 0 <= y < P and 0 <= x < P:
                        ref[p, :, q * f:(q + 1) * f] += a[y * P + x] @ weights[q][dy + q, dx + q]
    assert numpy.allclose(got, ref, rtol=1e-12, atol=1e-12)


def test_pairing_halves_the_mma_count_of_the_7x7_level(lib):
    tiles, segs = plan(lib, 7, 4, 32)
    single = sum(min(y + 3, 6) - max(y - 3, 0) + 1 for y in range(7)) ** 2          # (sum of per-axis tap counts)^2 = 1369
    assert single == 1369 and len(segs) < 0.62 * single                              # 28 tiles instead of 49, 60 % of the MMAs
    assert len(tiles) == 7 * 4 and int((tiles[:, 1] < 0).sum()) == 7                 # one unpaired position per row
    assert int(numpy.minimum(segs[:, 1] + segs[:, 2], 256).max()) <= 256


def test_bad_arguments(lib):
    z = numpy.zeros(8, numpy.int32)
    ptr = z.ctypes.data_as(ctypes.c_void_p)
    assert lib.hyp_debug_plan_level_pairs(7, 5, 32, ptr, 2, ptr, 1, ptr) == N.HYP_E_INVALID     # R * fpad > 128
    assert lib.hyp_debug_plan_level_pairs(7, 4, 32, ptr, 2, ptr, 1, ptr) == N.HYP_E_INVALID     # tables too small
