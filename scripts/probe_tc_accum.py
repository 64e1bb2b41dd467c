"""Probe: how does tcgen05 kind::tf32 round its fp32 accumulation?  (run on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_tc import tc_gemm


def rna_tf32(x):
    b = x.view(torch.int32)
    b = (b + 0x1000) & ~0x1fff
    return b.view(torch.float32)


def split(x):
    hi = rna_tf32(x)
    lo = rna_tf32(x - hi)
    return hi.double(), lo.double()


for positive in (True, False):
    for K in (512, 4096, 16384):
        g = torch.Generator().manual_seed(K)
        if positive:
            A = torch.rand((128, K), generator=g) + 0.5
            B = torch.rand((64, K), generator=g) + 0.5
        else:
            A = torch.randn((128, K), generator=g)
            B = torch.randn((64, K), generator=g)
        ah, al = split(A)
        bh, bl = split(B)
        exact3 = ah @ bh.T + al @ bh.T + ah @ bl.T
        full = A.double() @ B.double().T
        scale = full.abs().max() if not positive else full.abs()
        for ch in (1 << 20, 16, 8, 4):
            got = tc_gemm(0, A, B, chunk_kb=ch).double()
            e_model = ((got - exact3) / scale)
            print(f"  chunk_kb={ch:8d}: vs-split-exact mean {e_model.mean():+.3e} std {e_model.std():.3e} max {e_model.abs().max():.3e}")
        got = tc_gemm(0, A, B).double()
        e_model = ((got - exact3) / scale)
        e_full = ((got - full) / scale)
        fp32 = (((A @ B.T).double() - full) / scale)
        print(f"positive={positive} K={K:6d} adds={K // 8:5d}  vs-split-exact: mean {e_model.mean():+.3e} std {e_model.std():.3e} "
              f"max {e_model.abs().max():.3e} | vs-fp64: max {e_full.abs().max():.3e} | torch fp32 max {fp32.abs().max():.3e}")
