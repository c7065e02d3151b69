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
