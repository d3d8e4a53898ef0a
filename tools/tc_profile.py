"""Run on the GPU box, lego_render frame: in-kernel cycle accounting of the pair kernel (block 0; v3 / v5 / v6) and, with
PNB_DBG_FLAGS=4, the cycles of every CTA (v5 / v6 / v7).  PNB_TC_VERSION selects the variant, PNB_NO_PROF=1 switches the
accounting off (it slows block 0 down), PNB_NO_WEIGHTS=1 removes the weight traffic (garbage results)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnerf_b200 import harness, scene
dev = torch.device("cuda:0")
cfg = scene.CONFIGS["lego_render"]
net, pts, opt = harness.build_model(cfg, dev, alpha_bias=3.0, pnb_tc_version=int(os.environ.get("PNB_TC_VERSION", "7")), pnb_color_version=int(os.environ.get("PNB_COLOR_VERSION", "2")))
rays = scene.make_rays(cfg)
rd = rays["raydir"].to(dev)
for i in range(3):
    with torch.no_grad():
        net.render_full(list(cfg.campos), rd, torch.eye(3), cfg.near, cfg.far, [1., 1., 1.])
if os.environ.get("PNB_NO_WEIGHTS"):
    net.tc_mask |= 64
net.tc_mask |= (int(os.environ.get("PNB_DBG_FLAGS", "0")) | (0 if os.environ.get("PNB_NO_PROF") else 1)) << 8
torch.cuda.synchronize()
net._err.zero_()
with torch.no_grad():
    net.render_full(list(cfg.campos), rd, torch.eye(3), cfg.near, cfg.far, [1., 1., 1.])
torch.cuda.synchronize()
c = net._err.cpu().view(torch.int64)[1:17].tolist()
names = ["loader wait empty", "issuer wait a1_ready", "issuer wait at_ready", "issuer wait full(weights)", "builder wait a1_free",
         "builder busy", "epilogue wait acc_full", "epilogue busy (l<3)", "epilogue busy (l==3)", "kernel total (thread 0)", "issuer: in ring commits (v6)", "issuer: K-block issue incl. commits (v6)", "issuer: wait kblk (probe)", "issuer: wait weights (probe)", "issuer: #weight waits", "peer loader wait empty"]
ntiles = (net.last.counters["n_valid"] if net.last.counters else 3472901) if False else None
tot = c[9]
print("status", int(net._err[0]), "version", os.environ.get("PNB_TC_VERSION", "7"), "no_weights", bool(os.environ.get("PNB_NO_WEIGHTS")))
for n, v in zip(names, c):
    print("%-28s %12d cycles  %5.1f%% of kernel" % (n, v, 100.0 * v / max(tot, 1)))

if int(os.environ.get("PNB_DBG_FLAGS", "0")) & 4:
    per = net._err.cpu().view(torch.int64)[32:32 + 148].tolist()
    print("n_quads (v7):", int(net._err.cpu().view(torch.int64)[32 + 192]), "n_valid:", net.last.counters.get("n_valid") if net.last and net.last.counters else None)
    print("per-CTA kernel cycles (M) @smid:", " ".join("%.1f@%d" % ((v & 0xffffffffffff) / 1e6, v >> 48) for v in per))
