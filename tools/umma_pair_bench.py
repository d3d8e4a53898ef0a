"""Run on the GPU box: cycles per tcgen05.mma.cta_group::2 (M=256 over a CTA pair, N=256, K=16) for SS / TS forms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnerf_b200 import lib as _lib
l = _lib.load_selftest()
dev = "cuda:0"
A = torch.randn(256, 32, device=dev)
W = torch.randn(256, 32, device=dev)
D = torch.zeros(256, 256, device=dev)
out = torch.zeros(2, dtype=torch.int64, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
KIND = ["multicast(3)", "multicast(1)", "cta_group::2 local", "cta_group::1"]
for mode in (0, 1):
    for flags in (0, 1 << 16, (1 << 16) | (1 << 13), (1 << 16) | (1 << 12) | (1 << 13), (1 << 16) | 2, (1 << 16) | 2 | (1 << 12) | (1 << 13)):
        iters = 3000
        _lib.check_selftest(l.pnb_umma_selftest2(A.data_ptr(), W.data_ptr(), D.data_ptr(), 32, 256, mode, iters, flags, out.data_ptr(), err.data_ptr(), st), "selftest2")
        torch.cuda.synchronize()
        o = out.tolist()
        print("cta_group::2 %s rotate=%d stress[tmem=%d bulk=%d test_spin=%d try_spin=%d] commit every %d (%s): issue %.1f cyc/mma, complete %.1f cyc/mma (err %d)" % (
            "TS" if mode else "SS", (flags >> 16) & 1, (flags >> 12) & 1, (flags >> 13) & 1, (flags >> 14) & 1, (flags >> 15) & 1, flags & 255, KIND[(flags >> 8) & 3], o[0] / iters, o[1] / iters, int(err[0])))
