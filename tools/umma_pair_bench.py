"""Run on the GPU box: cycles per tcgen05.mma.cta_group::2 (M=256 over a CTA pair, N=256, K=16) for SS / TS forms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnerf_b200 import lib as _lib
l = _lib.load()
dev = "cuda:0"
A = torch.randn(256, 32, device=dev)
W = torch.randn(256, 32, device=dev)
D = torch.zeros(256, 256, device=dev)
out = torch.zeros(2, dtype=torch.int64, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for mode in (0, 1):
    for iters in (300, 3000):
        _lib.check(l.pnb_umma_selftest2(A.data_ptr(), W.data_ptr(), D.data_ptr(), 32, 256, mode, iters, out.data_ptr(), err.data_ptr(), st), "selftest2")
        torch.cuda.synchronize()
        o = out.tolist()
        print("cta_group::2 %s iters=%4d: issue %.1f cyc/mma, complete %.1f cyc/mma (err %d)" % (
            "TS" if mode else "SS", iters, o[0] / iters, o[1] / iters, int(err[0])))
