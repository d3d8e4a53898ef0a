"""GPU box, ONE GPU: what a rank of an N-way frame split has to do.  Renders the shard `rays[0::N]` of the lego frame (and of the
truck frame) for N = 1, 2, 4, 8 and prints device ms per frame (CUDA events, L2 flushed), host issue ms per frame (perf_counter
around the calls, no synchronisation) and the implied strong-scaling efficiency t(1) / (N * t(N)) before any collective.
The fixed per-call latency (launch chain, kernel tails, host issue time) is what bounds the 8-GPU strong-scaled frame; this
separates it from the NCCL part without spending 8 GPUs.     python tools/shard_latency.py [lego_render|truck_8gpu]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnerf_b200 import harness, scene

name = sys.argv[1] if len(sys.argv) > 1 else "lego_render"
dev = torch.device("cuda:0")
cfg = scene.CONFIGS[name]
net, _, opt = harness.build_model(cfg, dev, seed=0, alpha_bias=3.0)
dirs = scene.make_rays(cfg)["raydir"][0]
cam = (list(cfg.campos), torch.eye(3), cfg.near, cfg.far, [1.0, 1.0, 1.0])
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
res = {}
for n in (1, 2, 4, 8):
    rd = dirs[0::n].contiguous().to(dev)

    def frame():
        with torch.no_grad():
            return net.render_full(cam[0], rd, cam[1], cam[2], cam[3], cam[4])
    for _ in range(4):
        frame()
    torch.cuda.synchronize()
    net.check_errors()
    K = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = 0.0
    e0.record()
    for _ in range(K):
        flush.fill_(1)
        t0 = time.perf_counter()
        frame()
        host += time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    net.check_errors()
    res[n] = dict(rays=int(rd.shape[0]), device_ms=e0.elapsed_time(e1) / K, host_issue_ms=host / K * 1e3)
for n, r in res.items():
    r["efficiency_before_collective"] = res[1]["device_ms"] / (n * r["device_ms"])
    print("%s  shard 1/%d: %7d rays  device %.3f ms/frame  host issue %.3f ms/frame  efficiency %.3f" % (
        name, n, r["rays"], r["device_ms"], r["host_issue_ms"], r["efficiency_before_collective"]))
print(json.dumps({name: res}))
