"""BASELINE.json configs[2] ("ship optimise"): per-scene optimisation steps (forward + backward + 2x Adam, fp32) on the
synthetic N=600k scene, 3600 random rays of one 800x800 view per step (run/train_ft.py, lego_cuda.sh:109), jitter on.
Prints one JSON line (steps/s, Mrays/s, ms per phase).  Run on the GPU box:  python tools/bench_train.py --steps 20"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pointnerf_b200 import harness, scene

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--precision", default="bf16x3")
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = scene.CONFIGS["ship_optimise"]
net, pts, opt = harness.build_model(cfg, dev, alpha_bias=3.0, is_train=True, pnb_precision=args.precision)
mlp_params = list(net.aggregator.parameters())
pt_params = [p for p in net.neural_points.parameters() if p.requires_grad]
opt_mlp = torch.optim.Adam(mlp_params, lr=5e-4, betas=(0.9, 0.999))       # lego_cuda.sh: lr
opt_pts = torch.optim.Adam(pt_params, lr=2e-3, betas=(0.9, 0.999))        # plr
rng = np.random.RandomState(0)
gt = torch.rand(1, 3600, 3, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tf = tb = to = 0.0
hit = 0
for it in range(args.warmup + args.steps):
    px = rng.randint(0, cfg.W, size=(3600,)).astype(np.float32)
    py = rng.randint(0, cfg.H, size=(3600,)).astype(np.float32)
    rays = {k: v.to(dev) for k, v in scene.make_rays(cfg, np.stack([px, py], -1)).items()}
    ev[0].record()
    out = net(rays["campos"], rays["raydir"], bg_color=rays["bg_color"], camrotc2w=rays["camrotc2w"], pixel_idx=rays["pixel_idx"],
              near=rays["near"], far=rays["far"], h=rays["h"], w=rays["w"], intrinsic=rays["intrinsic"])
    mask = out["ray_mask"][0] > 0
    loss = ((out["coarse_raycolor"][0] - gt[0][mask]) ** 2).mean() + 1e-4 * (-torch.log(out["conf_coefficient"] + 1e-3)).mean()
    ev[1].record()
    opt_mlp.zero_grad(); opt_pts.zero_grad()
    loss.backward()
    ev[2].record()
    opt_mlp.step(); opt_pts.step()
    ev[3].record()
    torch.cuda.synchronize()
    if it >= args.warmup:
        tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2]); to += ev[2].elapsed_time(ev[3])
        hit += int(mask.sum())
n = args.steps
tot = (tf + tb + to) / n
print(json.dumps(dict(metric="optimisation steps/s (3600 rays/step, N=600k, fwd+bwd+Adam, fp32)", value=1e3 / tot, unit="steps/s",
                      mrays_per_s=3600 / tot / 1e3, ms_forward=tf / n, ms_backward=tb / n, ms_adam=to / n, hit_rays_per_step=hit / n,
                      precision=args.precision, loss=float(loss.detach()), config="ship_optimise (BASELINE configs[2])")))
