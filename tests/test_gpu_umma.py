"""GPU: tcgen05 building blocks (descriptors, operand layouts, TMEM, BF16x3 split) against a plain fp64 matmul."""
import pytest
import torch

from pointnerf_b200 import lib as _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(A, W, layout):
    l = _lib.load_selftest()
    K, N = A.shape[1], W.shape[0]
    D = torch.full((128, N), float("nan"), device=DEV)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    _lib.check_selftest(l.pnb_umma_selftest(A.data_ptr(), W.data_ptr(), D.data_ptr(), K, N, layout, err.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream), "pnb_umma_selftest")
    torch.cuda.synchronize()
    return D, int(err.item())


@pytest.mark.parametrize("layout", [0, 4])
@pytest.mark.parametrize("K,N", [(16, 16), (32, 256), (64, 128), (288, 256), (272, 256), (256, 256), (100, 64)])
def test_umma_bf16x3_matches_matmul(layout, K, N):
    g = torch.Generator(device=DEV).manual_seed(K * 1000 + N)
    A = torch.randn(128, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    D, err = _run(A.contiguous(), W.contiguous(), layout)
    assert err == 0, "pipeline timeout code %d" % err
    ref = (A.double() @ W.double().t())
    d = (D.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert d <= 2e-5 * max(scale, 1.0), "layout %d K=%d N=%d: max abs err %.3e (scale %.2f)" % (layout, K, N, d, scale)
    # the split must beat a single bf16 pass by orders of magnitude
    single = (A.bfloat16().double() @ W.bfloat16().double().t() - ref).abs().max().item()
    assert d < single / 50


@pytest.mark.parametrize("layout,K,N", [(100, 16, 16), (100, 64, 256), (100, 256, 256), (101, 32, 256), (101, 272, 256)])
def test_umma_a_operand_in_tensor_memory(layout, K, N):
    """TS form (A in TMEM, written with tcgen05.st), optionally mixed with one SS k-step on the same accumulator."""
    g = torch.Generator(device=DEV).manual_seed(K * 7 + N)
    A = torch.randn(128, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    D, err = _run(A.contiguous(), W.contiguous(), layout)
    assert err == 0, "pipeline timeout code %d" % err
    ref = (A.double() @ W.double().t())
    d = (D.double() - ref).abs().max().item()
    assert d <= 2e-5 * max(ref.abs().max().item(), 1.0), "mode %d K=%d N=%d: max abs err %.3e" % (layout, K, N, d)


@pytest.mark.parametrize("mode,K,N", [(0, 32, 256), (0, 288, 256), (0, 64, 64), (1, 16, 32), (1, 256, 256), (1, 64, 128)])
def test_umma_cta_pair(mode, K, N):
    """cta_group::2 (cluster of 2): M=256, each CTA stages half of B; SS and TS forms, multicast commit."""
    l = _lib.load_selftest()
    g = torch.Generator(device=DEV).manual_seed(K * 13 + N + mode)
    A = torch.randn(256, K, device=DEV, generator=g).contiguous()
    W = (torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).contiguous()
    D = torch.full((256, N), float("nan"), device=DEV)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = torch.zeros(2, dtype=torch.int64, device=DEV)
    _lib.check_selftest(l.pnb_umma_selftest2(A.data_ptr(), W.data_ptr(), D.data_ptr(), K, N, mode, 0, 0, out.data_ptr(), err.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream), "pnb_umma_selftest2")
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "pipeline timeout code %d" % int(err.item())
    ref = A.double() @ W.double().t()
    d = (D.double() - ref).abs().max().item()
    assert d <= 2e-5 * max(ref.abs().max().item(), 1.0), "mode %d K=%d N=%d: max abs err %.3e" % (mode, K, N, d)
