"""Run on the GPU box: cycles per tcgen05.mma for SS / TS operand forms, two operand layouts, with and without a
concurrent cp.async.bulk stream into shared memory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnerf_b200 import lib as _lib
l = _lib.load_selftest()
dev = "cuda:0"
src = torch.zeros(2 << 20, dtype=torch.uint8, device=dev)
out = torch.zeros(4, dtype=torch.int64, device=dev)
err = torch.zeros(4, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for layout in (0,):
    for mode in (0, 1):
        for bulk in (0, 2 << 8, 4 << 8, 1 << 16, (1 << 16) | 1, (1 << 16) | (2 << 8)):   # bits 8+: commit every n MMAs   # bit 0: concurrent cp.async.bulk stream, bit 1: concurrent tcgen05.ld/st traffic from 3 warps
            for iters in (3000,):
                _lib.check_selftest(l.pnb_umma_bench(layout, mode, iters, bulk, src.data_ptr(), out.data_ptr(), err.data_ptr(), st), "bench")
                torch.cuda.synchronize()
                o = out.tolist()
                print("cta_group::1 layout %d %s flags=0x%x iters=%4d: issue %.1f cyc/mma, complete %.1f cyc/mma (err %d)" % (
                    layout, "TS" if mode else "SS", bulk, iters, o[0] / iters, o[1] / iters, int(err[0])))
