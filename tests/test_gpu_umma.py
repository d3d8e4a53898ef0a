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


# ---- the tcgen05 GEMM engine of the backward pass (csrc/gemm_tc.cu) against fp64 matmuls, every operand form it is used in
def _gemm_tc(A, a_rs, a_ks, B, b_rs, b_ks, M, N, K, ldc, bias=None, act=0, dact=None, ldd=0, dact_n=0, splits=1, C0=None, precise=0):
    l = _lib.load_selftest()
    C = torch.full((M, ldc), 7.0, device="cuda") if C0 is None else C0.clone()
    err = torch.zeros(4, dtype=torch.int32, device="cuda")
    part = torch.empty(max(max(splits, 1) * M * N, 1 << 19), device="cuda")       # >= 2 MB: also holds the weight images of the k_gemm_tcw path
    ptr = lambda t: t.data_ptr() if t is not None else None
    _lib.check_selftest(l.pnb_gemm_tc_test(A.data_ptr(), a_rs, a_ks, B.data_ptr(), b_rs, b_ks, C.data_ptr(), ldc, M, N, K, ptr(bias), act,
                                           ptr(dact), ldd, dact_n, splits, part.data_ptr(), part.numel() * 4, 1 if C0 is not None else 0,
                                           precise, err.data_ptr(), torch.cuda.current_stream().cuda_stream), "pnb_gemm_tc_test")
    torch.cuda.synchronize()
    assert int(err[0]) == 0
    return C


def _rel(a, ref):
    return ((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,N,K", [(300, 256, 288), (128, 128, 128), (1000, 256, 272), (77, 288, 256), (4096, 128, 128)])
def test_gemm_tc_nn_nt(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    X = torch.randn(M, K + 16, device="cuda", generator=g)[:, :K]                 # leading dimension != K
    lda = X.stride(0)
    Wt = (0.1 * torch.randn(K, N, device="cuda", generator=g)).contiguous()       # W^T [K][N] as the backward holds the weights
    bias = torch.randn(N, device="cuda", generator=g)
    # NN (forward recompute): H = LeakyReLU(X Wt + b);  A(m,k) = X[m*lda + k], B(n,k) = Wt[k*N + n]
    H = _gemm_tc(X, lda, 1, Wt, 1, N, M, N, K, N + 16, bias=bias, act=1)
    ref = torch.nn.functional.leaky_relu(X.double() @ Wt.double() + bias.double(), 0.01)
    assert _rel(H[:, :N], ref) < 2e-5 and torch.all(H[:, N:] == 7.0)
    # the 3-part split (6 products): stops at ~1e-6 - the tensor core does not accumulate with full fp32 precision
    Hp = _gemm_tc(X, lda, 1, Wt, 1, N, M, N, K, N + 16, bias=bias, act=1, precise=1)
    assert _rel(Hp[:, :N], ref) < 4e-6, _rel(Hp[:, :N], ref)      # ~1e-6: the accumulation precision of the tensor core, not fp32
    # NT (dX = dZ W * LeakyReLU'(Y)): A = dZ [M x N], B(n=k_in, k=n_out) = Wt[k_in*N + n_out]; result [M x K]
    if K % 16 == 0:
        dZ = torch.randn(M, N, device="cuda", generator=g)
        Y = torch.randn(M, K, device="cuda", generator=g)
        dact_n = (K // 16 - 1) * 16
        dX = _gemm_tc(dZ, N, 1, Wt, N, 1, M, K, N, K, dact=Y, ldd=K, dact_n=dact_n)
        refx = dZ.double() @ Wt.double().t()
        mask = torch.where(Y > 0, 1.0, 0.01).double()
        mask[:, dact_n:] = 1.0
        assert _rel(dX, refx * mask) < 2e-5


@pytest.mark.parametrize("rows,Kin,Nout,splits", [(5000, 288, 256, 3), (130000, 256, 256, 64), (999, 128, 128, 2), (40000, 272, 256, 20)])
def test_gemm_tc_tn_splitk_accumulate(rows, Kin, Nout, splits):
    """dWt[Kin x Nout] += X^T dZ: both operands transposed on the way into shared memory, deterministic split-K, accumulation."""
    g = torch.Generator(device="cuda").manual_seed(rows)
    X = torch.randn(rows, Kin, device="cuda", generator=g)
    dZ = torch.randn(rows, Nout + 16, device="cuda", generator=g)[:, :Nout]
    C0 = torch.randn(Kin, Nout, device="cuda", generator=g)
    out = _gemm_tc(X, 1, Kin, dZ, 1, dZ.stride(0), Kin, Nout, rows, Nout, splits=splits, C0=C0)
    ref = C0.double() + X.double().t() @ dZ.double()
    assert _rel(out, ref) < 2e-5
    out2 = _gemm_tc(X, 1, Kin, dZ, 1, dZ.stride(0), Kin, Nout, rows, Nout, splits=splits, C0=C0)
    assert torch.equal(out, out2)                                                   # deterministic (no atomics)
