"""GPU box: per-kernel time table of the per-scene optimisation step (config 3, ship_optimise: N=600k, 3600 rays) with torch.profiler
(CUPTI): where the forward / backward / Adam milliseconds go.  python tools/train_profile.py [config] [bwd_flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
from pointnerf_b200 import harness, parallel, scene

name = sys.argv[1] if len(sys.argv) > 1 else "ship_optimise"
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
cfg = scene.CONFIGS[name]
net, pts, opt = harness.build_model(cfg, dev, alpha_bias=3.0, is_train=True, pnb_bwd_fp32=flags)
ts = parallel.TrainStep(net)
rng = np.random.RandomState(0)
g = torch.Generator().manual_seed(1)


def batch():
    px = rng.randint(0, cfg.W, size=(3600,)).astype(np.float32)
    py = rng.randint(0, cfg.H, size=(3600,)).astype(np.float32)
    rays = {k: (v.to(dev) if k in ("raydir", "pixel_idx") else v) for k, v in scene.make_rays(cfg, np.stack([px, py], -1)).items()}
    kw = dict(campos=rays["campos"], raydir=rays["raydir"], bg_color=rays["bg_color"], camrotc2w=rays["camrotc2w"], pixel_idx=rays["pixel_idx"],
              near=rays["near"], far=rays["far"], h=rays["h"], w=rays["w"], intrinsic=rays["intrinsic"])
    return kw, torch.rand(3600, 3, generator=g).to(dev)


for _ in range(4):
    ts.step(*batch())
torch.cuda.synchronize()
print("valid samples of the last step:", net.last.counters["n_valid"], "pairs", net.last.counters["n_pairs"])
N_STEPS = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N_STEPS):
        ts.step(*batch())
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print("total device time per step %.3f ms (%d kernels per step)" % (tot / N_STEPS / 1e3, sum(e.count for e in rows) // N_STEPS))
for e in rows[:40]:
    print("%8.3f ms/step  x%-4d  %s" % (e.device_time_total / N_STEPS / 1e3, e.count // N_STEPS, e.key[:110]))
