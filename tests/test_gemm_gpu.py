"""Stand-alone checks of the two tcgen05 GEMM kernels against torch matmul (fp16 in, fp32 accumulate).

Both GEMMs are exact up to fp32 summation order, so the tolerance is tight (1e-3 relative to the
largest magnitude is far above the observed ~1e-6)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (128, 256, 512), (300, 512, 1024), (16384, 512, 512), (1000, 256, 1472)])
def test_kmajor_gemm(native, m, n, k):
    torch.manual_seed(0)
    a = torch.randn(m, k, device="cuda").half()
    b = torch.randn(n, k, device="cuda").half()
    c = torch.full((m, n), float("nan"), device="cuda")
    native.check(native.lib().npp_debug_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, native.current_stream()))
    ref = a.float() @ b.float().t()
    assert torch.isfinite(c).all()
    assert _rel(c, ref) < 1e-4


@pytest.mark.parametrize("rows,m,n,splits", [(64, 256, 256, 1), (512, 256, 256, 1), (1000, 256, 512, 3), (16384, 512, 1024, 6), (4100, 512, 512, 7)])
def test_wgrad_gemm(native, rows, m, n, splits):
    torch.manual_seed(0)
    a = torch.randn(rows, m, device="cuda").half()
    b = torch.randn(rows, n, device="cuda").half()
    c = torch.full((splits, m, n), float("nan"), device="cuda")
    native.check(native.lib().npp_debug_wgrad(a.data_ptr(), b.data_ptr(), c.data_ptr(), rows, m, n, splits, native.current_stream()))
    ref = a.float().t() @ b.float()
    got = c.sum(0)
    assert torch.isfinite(got).all()
    assert _rel(got, ref) < 1e-4
