"""GPU tests of the tcgen05/TMA segment-GEMM building block (3xTF32 split) through the C ABI probe
``hyp_debug_tc_gemm``, against a float64 matmul of the same inputs."""
import ctypes

import numpy
import pytest
import torch

pytestmark = pytest.mark.gpu


def tc_gemm(mn, A, B, raw_hi=0, bn=0, ksplit=1, stats=False, chunk_kb=0):
    from hypelcnn_b200 import _native as N
    A, B = A.cuda().contiguous(), B.cuda().contiguous()
    if mn & 1:
        K, M = A.shape
        Nn = B.shape[1]
    else:
        M, K = A.shape
        Nn = B.shape[0]
    D = torch.zeros((M, Nn), dtype=torch.float32, device="cuda")
    S = torch.zeros(((M + 127) // 128, 2, Nn), dtype=torch.float32, device="cuda") if stats else None
    N.check(N.lib().hyp_debug_tc_gemm(mn, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), M, Nn, K,
                                      ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(S.data_ptr() if stats else 0),
                                      raw_hi, bn, ksplit, chunk_kb, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return (D.cpu(), S.cpu()) if stats else D.cpu()


def rel_err(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (128, 256, 64), (300, 120, 145), (256, 240, 480), (1000, 480, 240),
                                   (77, 16, 8), (130, 980, 2940)])
def test_kmajor_3xtf32_matches_fp64(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn((M, K), generator=g)
    B = torch.randn((N, K), generator=g)
    ref = A.double() @ B.double().T
    got = tc_gemm(0, A, B)
    fp32 = rel_err(A @ B.T, ref)
    err = rel_err(got, ref)
    assert err < max(3e-6, 4 * fp32), (err, fp32)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (128, 64, 32), (300, 120, 145), (1000, 480, 240), (77, 16, 8),
                                   (130, 980, 2940), (40000, 240, 480)])
def test_kmajor_cta_pair_matches_fp64(M, N, K):
    """cta_group::2: two CTAs share the B tile (each loads half of it); odd tile counts get a phantom tile.
    The last shape has more tile pairs than SM pairs, so the persistent tile walk wraps."""
    g = torch.Generator().manual_seed(M * 7 + N + 1)
    A = torch.randn((M, K), generator=g)
    B = torch.randn((N, K), generator=g)
    ref = A.double() @ B.double().T
    got = tc_gemm(2, A, B)
    fp32 = rel_err(A @ B.T, ref)
    err = rel_err(got, ref)
    assert err < max(3e-6, 4 * fp32), (err, fp32)


@pytest.mark.parametrize("bn", [16, 32, 64])
def test_kmajor_cta_pair_multiple_b_boxes(bn):
    g = torch.Generator().manual_seed(bn + 100)
    A = torch.randn((512, 120), generator=g)
    B = torch.randn((240, 120), generator=g)
    ref = A.double() @ B.double().T
    assert rel_err(tc_gemm(2, A, B, bn=bn), ref) < 4e-6


def test_kmajor_persistent_wraps_tiles():
    """more tiles than SMs: every CTA walks several tiles, smem / TMEM rings keep running across them"""
    g = torch.Generator().manual_seed(77)
    A = torch.randn((128 * 400, 96), generator=g)
    B = torch.randn((120, 96), generator=g)
    ref = A.double() @ B.double().T
    D, S = tc_gemm(0, A, B, stats=True)
    assert rel_err(D, ref) < 4e-6
    assert float((S[:, 0, :].double().sum(0) - ref.sum(0)).abs().max()) < 2e-2
    D2, S2 = tc_gemm(2, A, B, stats=True)
    assert rel_err(D2, ref) < 4e-6
    assert torch.equal(S, S2) or float((S - S2).abs().max()) < 1e-3


@pytest.mark.parametrize("bn", [16, 32, 64])
def test_kmajor_multiple_b_boxes(bn):
    g = torch.Generator().manual_seed(bn)
    A = torch.randn((256, 120), generator=g)
    B = torch.randn((240, 120), generator=g)
    ref = A.double() @ B.double().T
    assert rel_err(tc_gemm(0, A, B, bn=bn), ref) < 4e-6


@pytest.mark.parametrize("M,N,K,ksplit", [(128, 64, 64, 1), (120, 240, 1000, 1), (480, 480, 4096, 4), (145, 120, 777, 3),
                                          (60, 16, 200, 1)])
def test_mnmajor_3xtf32_matches_fp64(M, N, K, ksplit):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M), generator=g)
    B = torch.randn((K, N), generator=g)
    ref = A.double().T @ B.double()
    err = rel_err(tc_gemm(1, A, B, ksplit=ksplit), ref)
    fp32 = rel_err(A.T @ B, ref)
    assert err < max(3e-6, 4 * fp32), (err, fp32)


@pytest.mark.parametrize("M,N,K,ksplit", [(256, 64, 64, 1), (480, 240, 1000, 1), (480, 480, 4096, 4), (145, 120, 777, 3),
                                          (240, 128, 300, 2)])
def test_mnmajor_cta_pair_matches_fp64(M, N, K, ksplit):
    """wgrad layout as CTA pairs: two 128-row tiles of the M side share the B columns (each CTA loads half of them)."""
    g = torch.Generator().manual_seed(M + N + K + 7)
    A = torch.randn((K, M), generator=g)
    B = torch.randn((K, N), generator=g)
    ref = A.double().T @ B.double()
    err = rel_err(tc_gemm(3, A, B, ksplit=ksplit), ref)
    fp32 = rel_err(A.T @ B, ref)
    assert err < max(3e-6, 4 * fp32), (err, fp32)


def test_epilogue_column_statistics():
    g = torch.Generator().manual_seed(5)
    A = torch.randn((300, 96), generator=g)
    B = torch.randn((120, 96), generator=g)
    D, S = tc_gemm(0, A, B, stats=True)
    ref = (A.double() @ B.double().T)
    s1 = S[:, 0, :].double().sum(0)
    s2 = S[:, 1, :].double().sum(0)
    assert float((s1 - ref.sum(0)).abs().max()) < 1e-3
    assert float(((s2 - (ref * ref).sum(0)).abs() / (ref * ref).sum(0)).max()) < 1e-5


def test_tensor_core_input_truncation_probe():
    """Informational: does kind::tf32 ignore the 13 low mantissa bits of raw fp32 operands?
    (The engine never relies on it: plane 0 is always pre-rounded.)"""
    g = torch.Generator().manual_seed(9)
    A = torch.randn((128, 256), generator=g)
    B = torch.randn((64, 256), generator=g)
    ref = A.double() @ B.double().T
    err_raw = rel_err(tc_gemm(0, A, B, raw_hi=1), ref)
    err_rna = rel_err(tc_gemm(0, A, B, raw_hi=0), ref)
    print(f"raw-hi rel err {err_raw:.3e}; rna-hi rel err {err_rna:.3e}")
    assert err_rna < 4e-6


# ---- 16-bit operand formats: mn flag bits 2..3 select the format (1: fp16 hi/lo planes x 3 MMAs, 2: one bf16 plane) ----
OP_F16X3, OP_BF16 = 1 << 2, 2 << 2


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 256, 128), (300, 120, 145), (256, 240, 480), (1000, 480, 240),
                                   (77, 16, 8), (130, 980, 2940)])
@pytest.mark.parametrize("pair", [0, 2])
def test_kmajor_f16x3_matches_fp64(M, N, K, pair):
    """fp16 (hi, lo) planes, hi*hi + lo*hi + hi*lo on kind::f16 at twice the TF32 rate: the same 3e-6 bound."""
    g = torch.Generator().manual_seed(M * 7 + N + pair)
    A = torch.randn((M, K), generator=g)
    B = torch.randn((N, K), generator=g)
    ref = A.double() @ B.double().T
    err = rel_err(tc_gemm(pair | OP_F16X3, A, B), ref)
    fp32 = rel_err(A @ B.T, ref)
    assert err < max(3e-6, 4 * fp32), (err, fp32)


@pytest.mark.parametrize("M,N,K,ksplit", [(128, 64, 64, 1), (120, 240, 1000, 1), (480, 480, 4096, 4), (145, 120, 777, 3),
                                          (60, 16, 200, 1), (256, 64, 128, 1), (240, 128, 300, 2)])
@pytest.mark.parametrize("pair", [1, 3])
def test_mnmajor_f16x3_matches_fp64(M, N, K, ksplit, pair):
    """wgrad layout with 16-bit operands: 64 x 64 SWIZZLE_128B boxes, MN-major descriptors (LBO = one box, SBO 1024)."""
    if pair == 3 and (M + 127) // 128 % 2:
        pytest.skip("CTA pairs need an even number of 128-row tiles")
    g = torch.Generator().manual_seed(M + N + K + pair)
    A = torch.randn((K, M), generator=g)
    B = torch.randn((K, N), generator=g)
    ref = A.double().T @ B.double()
    err = rel_err(tc_gemm(pair | OP_F16X3, A, B, ksplit=ksplit), ref)
    fp32 = rel_err(A.T @ B, ref)
    assert err < max(3e-6, 4 * fp32), (err, fp32)


@pytest.mark.parametrize("mn", [0, 2, 1, 3])
def test_bf16_single_plane(mn):
    """The fast mode: one bf16 plane, one MMA per K step.  Error is bf16 input rounding (2^-9 per operand)."""
    g = torch.Generator().manual_seed(40 + mn)
    M, N, K = 256, 240, 512
    if mn & 1:
        A, B = torch.randn((K, M), generator=g), torch.randn((K, N), generator=g)
        ref = A.double().T @ B.double()
        bf = (A.bfloat16().double().T @ B.bfloat16().double())
    else:
        A, B = torch.randn((M, K), generator=g), torch.randn((N, K), generator=g)
        ref = A.double() @ B.double().T
        bf = A.bfloat16().double() @ B.bfloat16().double().T
    got = tc_gemm(mn | OP_BF16, A, B)
    assert rel_err(got, bf) < 3e-6          # exact products of the rounded operands, fp32 accumulation
    assert rel_err(got, ref) < 2e-2


def test_f16x3_small_magnitudes_keep_relative_accuracy():
    """Operands around 1e-3 (weights) and 1e-4: the probe scales A by 4 and B by 64 before the split, as the engine does
    with its per-tensor power-of-two scales; the error stays relative to the result's own scale."""
    g = torch.Generator().manual_seed(3)
    A = torch.randn((256, 480), generator=g) * 1e-2
    B = torch.randn((240, 480), generator=g) * 3e-2
    ref = A.double() @ B.double().T
    assert rel_err(tc_gemm(OP_F16X3, A, B), ref) < 2e-5
