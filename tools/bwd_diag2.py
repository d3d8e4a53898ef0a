"""Diagnostic (GPU box): workspace buffers of the backward pass, tensor-core GEMMs vs fp32 CUDA-core GEMMs (same inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pointnerf_b200 import harness, scene

DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "tiny_thin_sr8"
fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name + ".npz"))
cfg = scene.CONFIGS["tiny"]


def align(x, a=256):
    return (x + a - 1) // a * a


def layout(S, cap, P):
    items = [("X1", P * 288), ("H1", P * 256), ("X3", P * 272), ("H3", P * 256), ("H4", P * 256), ("wc", P), ("wn", P), ("sp", P), ("sg", P), ("dwc", P),
             ("pidx", P), ("CX", S * 288), ("C1", S * 128), ("C2", S * 128), ("C3", S * 128), ("O3", S * 4), ("G1", P * 288), ("G2", P * 272), ("G3", P * 256),
             ("GS1", S * 288), ("GS2", S * 128), ("GS3", S * 128), ("dO3", S * 4), ("dsig", S)]
    off, out = 0, {}
    for k, n in items:
        off = align(off)
        out[k] = (off, n)
        off += n * 4
    return out


ws = {}
for mode in (1, 0):
    net, pts, opt = harness.build_model(cfg, DEV, SR=int(fx["SR"]), max_o=100000, pnb_precision="fp32", pnb_bwd_fp32=mode)
    net.aggregator.load_state_dict({k[4:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("mlp.")})
    r = {k: v.to(DEV) for k, v in scene.make_rays(cfg, fx["pixels"]).items()}
    out = net(r["campos"], r["raydir"], bg_color=r["bg_color"], camrotc2w=r["camrotc2w"], pixel_idx=r["pixel_idx"],
              near=r["near"], far=r["far"], h=r["h"], w=r["w"], intrinsic=r["intrinsic"])
    ((out["coarse_raycolor"] ** 2).sum() + 1e-3 * out["conf_coefficient"].sum()).backward()
    torch.cuda.synchronize()
    S = net.last.counters["n_valid"]
    NP = net.last.counters["n_pairs"]
    ws[mode] = net._bwd_ws.clone()
    cap = net.last.desc.cap_samples
L = layout(S, cap, NP)
print("S", S, "P", NP)
ld = dict(X1=288, H1=256, X3=272, H3=256, H4=256, CX=288, C1=128, C2=128, C3=128, O3=4, G1=288, G2=272, G3=256, GS1=288, GS2=128, GS3=128, dO3=4)
for k, (off, n) in L.items():
    if k == "pidx":
        continue
    a = ws[0][off:off + n * 4].view(torch.float32).double()
    b = ws[1][off:off + n * 4].view(torch.float32).double()
    d = (a - b).abs()
    sc = b.abs().max().clamp_min(1e-30)
    i = int(d.argmax())
    w = ld.get(k, 1)
    print("%-5s scale %.3e  max|tc-fp32|/scale %.3e  at row %d col %d (tc %.6e fp32 %.6e)  rows with rel diff > 1e-3: %d" % (
        k, float(sc), float(d.max() / sc), i // w, i % w, float(a[i]), float(b[i]),
        int(((d.view(-1, w).max(1)[0]) > 1e-3 * sc).sum()) if n % w == 0 else -1))

# sign disagreements of the recomputed activations (LeakyReLU masks of the backward)
pid_off, pid_n = L["pidx"]
pidx = ws[0][pid_off:pid_off + pid_n * 4].view(torch.int32)
for k in ("H1", "X3", "H3", "H4", "C1", "C2", "C3"):
    off, n = L[k]
    w = ld[k]
    a = ws[0][off:off + n * 4].view(torch.float32).view(-1, w)
    b = ws[1][off:off + n * 4].view(torch.float32).view(-1, w)
    ncol = 256 if k == "X3" else w
    flip = ((a[:, :ncol] > 0) != (b[:, :ncol] > 0))
    idx = flip.nonzero()
    print("%s: %d sign flips of %d units" % (k, idx.shape[0], a.shape[0] * ncol))
    for r_, c_ in idx[:8].tolist():
        print("   row %d (point %s) col %d: tc %.3e fp32 %.3e" % (r_, int(pidx[r_]) if k in ("H1", "X3", "H3", "H4") else "-", c_, float(a[r_, c_]), float(b[r_, c_])))
off, n = L["G1"]
a = ws[0][off:off + n * 4].view(torch.float32).view(-1, 288); b = ws[1][off:off + n * 4].view(torch.float32).view(-1, 288)
d = (a - b).abs().max(1)[0]
top = torch.argsort(-d)[:8]
print("rows of G1 (dX1) with the largest tc-vs-fp32 difference:", [(int(r_), int(pidx[r_]), float(d[r_])) for r_ in top])
