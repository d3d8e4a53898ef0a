"""diagnostic: deferred vs non-deferred last epilogue of k_shade_tc8 on a patch; where do the outputs differ?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnerf_b200 import harness, scene
DEV = "cuda:0"
name, side = sys.argv[1], int(sys.argv[2])
cfg = scene.CONFIGS[name]
a, _, _ = harness.build_model(cfg, DEV, alpha_bias=3.0)
b, _, _ = harness.build_model(cfg, DEV, alpha_bias=3.0, pnb_dbg_flags=8)
rays = scene.make_rays(cfg, scene.centre_patch(cfg, side))
rd = rays["raydir"].to(DEV)
def R(net):
    with torch.no_grad():
        o = net.render_full(list(cfg.campos), rd, torch.eye(3), cfg.near, cfg.far, [1., 1., 1.])
    torch.cuda.synchronize(); net.check_errors()
    return {k: o[k].clone() for k in ("coarse_raycolor", "coarse_point_opacity", "coarse_is_background")}
outs = {"a1": R(a), "b1": R(b), "a2": R(a), "b2": R(b), "a3": R(a)}
cnt = a.last.counters if a.last is not None and getattr(a.last, "counters", None) else None
print("counters", cnt)
for x, y in (("a1", "a2"), ("a2", "a3"), ("b1", "b2"), ("a1", "b1"), ("a2", "b2")):
    for k in outs[x]:
        d = (outs[x][k] - outs[y][k]).abs()
        nz = (d > 0)
        if k == "coarse_point_opacity":
            rays_bad = nz.reshape(nz.shape[1], -1).any(-1).nonzero().flatten()
        else:
            rays_bad = nz.reshape(-1, nz.shape[-1]).any(-1).nonzero().flatten()
        print("%s vs %s  %-22s differing elements %7d  max %.3e  rays %d  first %s" % (x, y, k, int(nz.sum()), float(d.max()), rays_bad.numel(), rays_bad[:12].tolist()))
