"""Diagnostic (run on the GPU box): which operand layouts of pnb_umma_selftest reproduce a matmul, with fingerprints."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnerf_b200 import lib as _lib

l = _lib.load_selftest()
dev = "cuda:0"
for layout in (0, 4):
    for (K, N) in ((16, 16), (32, 32), (32, 256), (64, 256), (288, 256)):
        torch.manual_seed(0)
        A = torch.randn(128, K, device=dev); W = torch.randn(N, K, device=dev)
        D = torch.full((128, N), float("nan"), device=dev); err = torch.zeros(1, dtype=torch.int32, device=dev)
        rc = l.pnb_umma_selftest(A.data_ptr(), W.data_ptr(), D.data_ptr(), K, N, layout, err.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = A.double() @ W.double().t()
        d = (D.double() - ref).abs()
        print("layout %d K=%3d N=%3d rc=%d err=%d max|d|=%.3e nan=%d" % (layout, K, N, rc, int(err), float(d.nan_to_num(1e9).max()), int(torch.isnan(D).sum())))
        if float(d.nan_to_num(1e9).max()) > 1e-3 and K <= 32 and N <= 32:
            # fingerprint: one-hot A rows / W rows to see which (r,k) lands where
            for (r, k) in ((0, 0), (1, 0), (0, 1), (0, 8), (8, 0), (3, 17)):
                if k >= K: continue
                A1 = torch.zeros(128, K, device=dev); A1[r, k] = 1.0
                W1 = torch.arange(N * K, device=dev, dtype=torch.float32).reshape(N, K) * 0 + torch.arange(K, device=dev)[None, :] + 100 * torch.arange(N, device=dev)[:, None]
                D1 = torch.zeros(128, N, device=dev)
                l.pnb_umma_selftest(A1.data_ptr(), W1.contiguous().data_ptr(), D1.data_ptr(), K, N, layout, err.data_ptr(), torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                nz = torch.nonzero(D1)
                print("   A[%d,%d]=1 -> nonzero rows %s ; D[row,0:4]=%s" % (r, k, sorted(set(nz[:, 0].tolist()))[:6], D1[nz[0, 0] if len(nz) else 0, :4].tolist()))
