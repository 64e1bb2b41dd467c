"""The index walk of gather_kernel_v2 (hypelcnn_b200/csrc/hyp_kernels.cuh) replayed in Python: a thread's
(row, column, channel) position advanced incrementally by the block stride must equal the division-based position
gather_kernel computes for the same flat element index, for every thread and every patch / channel geometry the
engine meets (and a few it does not)."""
import pytest

BLOCK = 256


def _walk(S, per_pixel, thread):
    """The (i, py, px, c) sequence of one thread, transcribed from the kernel's loop."""
    total = S * S * per_pixel
    dpix = BLOCK // per_pixel
    dc = BLOCK - dpix * per_pixel
    pix = thread // per_pixel
    c = thread - pix * per_pixel
    py, px = divmod(pix, S)
    i = thread
    while i < total:
        yield i, py, px, c
        c += dc
        advance = dpix
        if c >= per_pixel:
            c -= per_pixel
            advance += 1
        px += advance
        if px >= S:
            rows = px // S
            py += rows
            px -= rows * S
        i += BLOCK


@pytest.mark.parametrize("S,per_pixel", [(7, 145), (11, 49), (11, 51), (3, 65), (7, 65), (1, 65), (1, 145), (5, 21),
                                         (1, 1), (3, 2), (13, 4), (5, 256), (5, 257), (3, 361), (9, 1000), (7, 255)])
def test_incremental_position_equals_division(S, per_pixel):
    total = S * S * per_pixel
    seen = 0
    for thread in range(BLOCK):
        for i, py, px, c in _walk(S, per_pixel, thread):
            pix, want_c = divmod(i, per_pixel)
            assert (py, px, c) == (pix // S, pix % S, want_c), (thread, i)
            seen += 1
    assert seen == total                      # every element of the patch is written exactly once
