"""Diagnostic (GPU box): gradients of the tensor-core backward (tcgen05 GEMMs) vs the fp32 CUDA-core backward vs the reference fixture,
per tensor: max abs diff / scale, and where the largest difference sits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pointnerf_b200 import harness, scene

DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "tiny_opaque"
fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name + ".npz"))
cfg = scene.CONFIGS["tiny"]
grads = {}
for mode in (0, 1, 2):
    net, pts, opt = harness.build_model(cfg, DEV, SR=int(fx["SR"]), max_o=100000, pnb_precision="fp32", pnb_bwd_fp32=mode)
    net.aggregator.load_state_dict({k[4:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("mlp.")})
    r = {k: v.to(DEV) for k, v in scene.make_rays(cfg, fx["pixels"]).items()}
    out = net(r["campos"], r["raydir"], bg_color=r["bg_color"], camrotc2w=r["camrotc2w"], pixel_idx=r["pixel_idx"],
              near=r["near"], far=r["far"], h=r["h"], w=r["w"], intrinsic=r["intrinsic"])
    loss = (out["coarse_raycolor"] ** 2).sum() + 1e-3 * out["conf_coefficient"].sum()
    loss.backward()
    net.check_errors()
    g = {"embedding": net.neural_points.points_embeding.grad, "color": net.neural_points.points_color.grad,
         "dir": net.neural_points.points_dir.grad, "conf": net.neural_points.points_conf.grad}
    for k, p in net.aggregator.named_parameters():
        g["mlp." + k] = p.grad
    grads[mode] = {k: v.detach().cpu().double().numpy() for k, v in g.items()}
ref = {"embedding": fx["grad_embedding"], "color": fx["grad_color"], "dir": fx["grad_dir"], "conf": fx["grad_conf"]}
for k in fx.files:
    if k.startswith("gradmlp."):
        ref["mlp." + k[8:]] = fx[k]
print("%-28s %10s | %12s %12s %12s %12s" % ("tensor", "scale", "tc-vs-ref", "fp32-vs-ref", "tc-vs-fp32", "tc3part-vs-ref"))
for k in grads[0]:
    rf = np.asarray(ref[k], np.float64).reshape(grads[0][k].shape)
    sc = max(np.abs(rf).max(), 1e-30)
    d0, d1, d01 = np.abs(grads[0][k] - rf), np.abs(grads[1][k] - rf), np.abs(grads[0][k] - grads[1][k])
    i = np.unravel_index(np.argmax(d0), d0.shape)
    d2 = np.abs(grads[2][k] - rf)
    print("%-28s %10.3e | %12.3e %12.3e %12.3e %12.3e   worst at %s: tc %.6e fp32 %.6e ref %.6e" % (k, sc, d0.max() / sc, d1.max() / sc, d01.max() / sc, d2.max() / sc, i,
                                                                                           grads[0][k][i], grads[1][k][i], rf[i]))
e = grads[0]["embedding"][0] - grads[1]["embedding"][0]
rows = np.argsort(-np.abs(e).max(1))[:5]
print("rows of points_embeding.grad with the largest tc-vs-fp32 difference:", rows.tolist())
for r_ in rows:
    print(" row", r_, "max|diff| %.3e" % np.abs(e[r_]).max(), "max|grad| %.3e" % np.abs(grads[1]["embedding"][0][r_]).max(), "cols", np.argsort(-np.abs(e[r_]))[:6].tolist())
