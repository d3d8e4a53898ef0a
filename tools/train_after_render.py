"""GPU box diagnostic: does a big render section earlier in the same process slow the optimisation step down?  Renders N frames of
config 4 (truck, 2 M points) first (PNB_PRE=truck | sr80 | none), then times the config-3 optimisation step exactly as bench.py does and
prints the per-kernel device times (torch.profiler) next to the wall time per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
from pointnerf_b200 import harness, parallel, scene

dev = torch.device("cuda:0")
pre = os.environ.get("PNB_PRE", "truck")
if pre != "none":
    tcfg = scene.CONFIGS["truck_8gpu" if pre == "truck" else "lego_render"]
    over = dict(SR=80) if pre == "sr80" else {}
    tnet, _, _ = harness.build_model(tcfg, dev, seed=0, alpha_bias=3.0, **over)
    rd = scene.make_rays(tcfg)["raydir"][0].to(dev)
    for _ in range(8):
        with torch.no_grad():
            tnet.render_full(list(tcfg.campos), rd, torch.eye(3), tcfg.near, tcfg.far, [1., 1., 1.])
    tnet.check_errors()
    torch.cuda.synchronize()
    print("pre-section %s done, allocated %.1f GB reserved %.1f GB" % (pre, torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
    if os.environ.get("PNB_KEEP") != "1":
        del tnet, rd
        torch.cuda.empty_cache()
cfg = scene.CONFIGS["ship_optimise"]
net, pts, opt = harness.build_model(cfg, dev, alpha_bias=3.0, is_train=True)
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
t0 = time.perf_counter()
for _ in range(20):
    ts.step(*batch())
torch.cuda.synchronize()
print("wall per step %.3f ms (reserved %.1f GB)" % ((time.perf_counter() - t0) / 20 * 1e3, torch.cuda.memory_reserved() / 1e9))
N_STEPS = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N_STEPS):
        ts.step(*batch())
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print("total device time per step %.3f ms (%d kernels per step)" % (tot / N_STEPS / 1e3, sum(e.count for e in rows) // N_STEPS))
for e in rows[:14]:
    print("%8.3f ms/step  x%-4d  %s" % (e.device_time_total / N_STEPS / 1e3, e.count // N_STEPS, e.key[:100]))
