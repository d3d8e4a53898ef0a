"""Run on the GPU box (round 2): the REFERENCE's own CUDA query extension (oracle/_ref/query_worldcoords_cuda.so, built
un-modified from /root/reference by oracle/build_ref.py in the container) next to libpnb200 on the same inputs.

The reference kernel is nondeterministic by construction: which occupied voxel receives occupancy slot 0 - and
therefore silently loses its points (query_worldcoords.cu:147) - depends on the order of atomicAdd in claim_occ, and the
order of the points inside a voxel on the atomics of fill_occ2pnts.  So the comparison is on neighbour SETS per sample, and
samples whose neighbourhood can see the reference's or our slot-0 voxel are reported separately.  This script only prints a
report; it is not a pytest test (it was written when the round's GPU budget was spent and has not run yet).

    python tests/ref_kernel_check.py [config] [patch_side]      (lives under tests/: only tests may import oracle/)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import build_ref
from pointnerf_b200 import harness, scene


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "lego_render"
    side = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    ext = build_ref.load_prebuilt()
    if ext is None:
        print("oracle/_ref/query_worldcoords_cuda.so is missing: run `python -m oracle.build_ref` in the build container")
        return 2
    dev = torch.device("cuda:0")
    cfg = scene.CONFIGS[name]
    net, pts, opt = harness.build_model(cfg, dev, alpha_bias=3.0)
    rays = scene.make_rays(cfg, scene.centre_patch(cfg, side))
    raydir = rays["raydir"].to(dev)
    R = raydir.shape[1]
    querier = net.neural_points.querier
    xyz = net.neural_points.xyz.detach().contiguous()
    r = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in rays.items()}

    # ---- ours, through the drop-in query_points (the reference's 7-tuple, point_query.py:72-98)
    ours = querier.query_points(r["pixel_idx"], None, xyz[None], torch.tensor([xyz.shape[0]], dtype=torch.int32, device=dev),
                                r["h"], r["w"], r["intrinsic"], cfg.near, cfg.far, raydir, r["campos"], r["camrotc2w"])
    ours_pidx = ours[0][0].cpu().numpy()                        # [R', SR, K]
    ours_mask = ours[4][0].cpu().numpy() > 0                    # [R]

    # ---- the reference kernel with the 18 arguments of point_query.py:85-92 (raypos = campos + raydir * t, fp32 mul then add)
    ranges_tensor, ranges_np, vsize_np, scaled_vdim_np = querier._hyper
    t = querier._t_for(cfg.near, cfg.far, R, dev)               # [D] eval table (host evaluation of the reference formula)
    raypos = (r["campos"][:, None, None, :] + raydir[:, :, None, :] * t[None, None, :, None]).contiguous()       # [1,R,D,3]
    D = raypos.shape[2]
    max_o = int(opt.max_o) if opt.max_o is not None else int(xyz.shape[0])
    out = ext.woord_query_grid_point_index(
        r["pixel_idx"], raypos, xyz[None], torch.tensor([xyz.shape[0]], dtype=torch.int32, device=dev),
        querier.kernel_size_tensor, querier.query_size_tensor, int(opt.SR), int(opt.K), R, D,
        torch.as_tensor(scaled_vdim_np, device=dev), max_o, int(opt.P), float(querier.radius_limit_np), ranges_tensor.to(dev),
        querier.scaled_vsize_tensor, 1024, 2)
    ref_pidx = out[0][0].cpu().numpy()                          # [R', SR, K]
    ref_mask = out[2][0].cpu().numpy().reshape(-1) > 0

    our_mask = ours_mask.reshape(-1)
    print("rays %d | hit: ours %d, reference %d, both %d" % (R, our_mask.sum(), ref_mask.sum(), (our_mask & ref_mask).sum()))
    rows_o = np.cumsum(our_mask) - 1
    rows_r = np.cumsum(ref_mask) - 1
    both = np.nonzero(our_mask & ref_mask)[0]
    same = diff = 0
    for r in both:
        a, b = ours_pidx[rows_o[r]], ref_pidx[rows_r[r]]
        for j in range(a.shape[0]):
            sa, sb = set(a[j][a[j] >= 0].tolist()), set(b[j][b[j] >= 0].tolist())
            if sa == sb:
                same += 1
            else:
                diff += 1
    print("samples compared %d | identical neighbour sets %d | different %d (expected: only around a slot-0 voxel)" % (same + diff, same, diff))
    return 0


if __name__ == "__main__":
    sys.exit(main())
