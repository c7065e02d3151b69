"""tcgen05 conventions (umma.cuh) pinned on hardware: split-precision GEMM against fp64."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(rows, n, k, alias, terms, scale=1.0):
    from elg_b200 import _lib
    g = torch.Generator().manual_seed(rows * 1000 + n + k)
    a = (torch.randn(rows, k, generator=g) * scale).cuda()
    b = (torch.randn(n, k, generator=g) * scale).cuda()
    d = torch.zeros(128, n, device="cuda")
    _lib.check(_lib.lib.elg_selftest_umma(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(d.data_ptr()),
                                          rows, n, k, alias, terms, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    ref = a.double() @ b.double().T
    return d[:rows].double(), ref


@pytest.mark.parametrize("rows,n,k,alias", [(128, 112, 128, 0), (100, 64, 32, 0), (52, 112, 128, 1), (64, 16, 16, 1)])
def test_split_precision_gemm(rows, n, k, alias):
    got, ref = _run(rows, n, k, alias, 3)
    err = (got - ref).abs().max() / ref.abs().max()
    assert err < 3e-6, err
    got5, _ = _run(rows, n, k, alias, 5)          # cross terms in their own accumulator: fp32-matmul-grade
    err5 = (got5 - ref).abs().mean() / ref.abs().mean()
    fp32 = ((ref.float().double() - ref).abs().mean() / ref.abs().mean())
    assert err5 < 6e-7 and err5 < 0.6 * ((got - ref).abs().mean() / ref.abs().mean()) + 1e-7, (err5, fp32)
    got1, _ = _run(rows, n, k, alias, 1)          # hi*hi only: fp16-grade, proves the lo terms matter
    err1 = (got1 - ref).abs().max() / ref.abs().max()
    assert 1e-5 < err1 < 5e-3, err1


@pytest.mark.parametrize("rows,n,k", [(128, 112, 128), (100, 16, 112), (77, 64, 32)])
def test_a_operand_from_tensor_memory(rows, n, k):
    """terms=6: A hi/lo written to TMEM with tcgen05.st (lane = row, two fp16 per column) and consumed by the
    TS form of tcgen05.mma -- the operand path of the tensor-core attention in rollout_tc.cu."""
    got, ref = _run(rows, n, k, 0, 6)
    err = (got - ref).abs().mean() / ref.abs().mean()
    assert err < 6e-7, err


@pytest.mark.parametrize("rows,n,k,alias", [(128, 112, 128, 0), (52, 112, 128, 1)])
def test_cross_terms_first_single_accumulator(rows, n, k, alias):
    """terms=7: the two small cross terms issued BEFORE hi*hi into one accumulator lose nothing (the reverse order
    truncates them against the large partial sums), so no second TMEM accumulator is needed."""
    got7, ref = _run(rows, n, k, alias, 7)
    got5, _ = _run(rows, n, k, alias, 5)
    e7 = (got7 - ref).abs().mean() / ref.abs().mean()
    e5 = (got5 - ref).abs().mean() / ref.abs().mean()
    assert e7 < 6e-7 and e7 < 1.3 * e5 + 1e-8, (e7, e5)
