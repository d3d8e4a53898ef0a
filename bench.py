#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, measured on B200.

Metric: Mrays/s (whole job) for full-image renders of the synthetic "lego_render" scene
(800x800 = 640,000 rays per image, K=8, N=400,000 neural points, SR=24 shading samples per ray, D=400 march
steps, fp32).  A "step" = one full image through the hot path (voxel query -> fused shading -> composite).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--sr 24]

N>1 is launched by torchrun (one rank per GPU): the global batch is N images, rays interleave-sharded
(ray i -> rank i % N), the point cloud / grid / MLP replicated, and ONE NCCL all-gather of the rendered colours per
step inside the timed region ("weak" scaling: per-GPU work fixed).

`--impl reference` times the reference's own CPU path (the oracle port: oracle/query_oracle.c +
oracle/shade_oracle.py, i.e. the reference's algorithm on host cores) on a bounded sample of the same workload.
The oracle is used here ONLY as the measured CPU baseline, never inside the GPU arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pointnerf_b200 import harness, scene  # noqa: E402

METRIC = "Mrays/s (800x800 render, K=8, 400k pts, SR=24), whole job"
FLOPS_PER_PAIR = 542720.0    # 2*(284*256 + 256*256 + 263*256 + 256*256 + 256)
FLOPS_PER_SAMPLE = 137984.0  # 2*(280*128 + 128*128 + 128*128 + 128*3)
# kernels of libpnb200.so launched per step (memsets are not kernels): march, 2 scans x 3, expand, knn,
# valid_list, count_rays, shade, composite
LAUNCHES_PER_STEP = 1 + 6 + 1 + 1 + 1 + 1 + 1 + 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi polled every 25 ms from before the warm-up (its start-up latency would otherwise eat a 0.2-s timed region);
    stop(t0, t1) keeps the samples whose timestamps fall inside the timed region [t0, t1] (host wall clock)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "25"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)

        def collect(lo, hi):
            sm, mx, reasons, pw = [], [], set(), []
            for r in rows:
                try:
                    ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if lo is not None and not (lo <= ts <= hi):
                        continue
                    sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                except Exception:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons, pw

        window = "timed region"
        sm, mx, reasons, pw = collect(t0, t1)
        if len(sm) < 2 and t0 is not None:          # very short region: fall back to everything since the warm-up started (also under load)
            sm, mx, reasons, pw = collect(None, None)
            window = "warm-up + timed region"
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)), samples=len(sm),
                    window=window, reasons=sorted(reasons))


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own path on the host cores = oracle port (C query + torch-CPU shading).  A pool of worker
# processes (each with a few intra-op threads: 128-thread eager torch on 1e5-row tensors is slower than 8 threads)
# renders reference-sized chunks of the frame concurrently, so that every host core is used.
_W = {}


def _cpu_worker_init(sr, threads):
    import torch as _t
    from oracle import query_oracle
    from pointnerf_b200.ray_marching import PointAggregator
    _t.set_num_threads(threads)
    query_oracle.build()
    cfg = scene.CONFIGS["lego_render"]
    cfg.SR = sr
    opt = harness.make_opt(cfg)
    agg = PointAggregator(opt, seed=0)
    with _t.no_grad():
        agg.alpha_branch[0].bias += 3.0
    _W.update(cfg=cfg, opt=opt, pts=scene.make_points(cfg), mlp=harness.mlp_cpu(agg))


def _cpu_worker_chunk(i):
    """Render one 2304-ray chunk (48x48 block i of the central 384x384 region: every ray hits the shell)."""
    from oracle import pipeline
    cfg, opt, pts, mlp = _W["cfg"], _W["opt"], _W["pts"], _W["mlp"]
    bx, by = i % 8, (i // 8) % 8
    x0, y0 = cfg.W // 2 - 192 + 48 * bx, cfg.H // 2 - 192 + 48 * by
    px, py = np.meshgrid(np.arange(x0, x0 + 48), np.arange(y0, y0 + 48))
    rays = scene.make_rays(cfg, np.stack((px, py), -1).reshape(-1, 2).astype(np.float32))
    t0 = time.perf_counter()
    pipeline.render(pts, mlp, rays["raydir"][0], cfg.campos, np.eye(3, dtype=np.float32), cfg.near, cfg.far, opt.vsize,
                    opt.vscale, opt.kernel_size, opt.query_size, opt.ranges, opt.SR, opt.K, opt.P, pts["xyz"].shape[0], D=cfg.D)
    return 2304, time.perf_counter() - t0


class CpuArm:
    def __init__(self, sr):
        import multiprocessing as mp
        self.cores = os.cpu_count() or 1
        self.threads = 8 if self.cores >= 16 else self.cores
        self.workers = max(1, self.cores // self.threads)
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_cpu_worker_init, initargs=(sr, self.threads))

    def step(self, k=0):
        """One step = every worker renders one chunk concurrently.  Returns (rays, seconds)."""
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker_chunk, [k * self.workers + j for j in range(self.workers)])
        return sum(r for r, _ in res), time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self, steps, secs):
        return ("%d steps x %d concurrent 2304-ray chunks (48x48 blocks of the central 384x384 region of the 800x800 frame, all rays "
                "hit) = %d worker processes x %d torch threads; voxel grid rebuilt per chunk as the reference does; %.1f s"
                % (steps, self.workers, self.workers, self.threads, secs))


WORKLOAD = "lego_render: 800x800 image per GPU, K=8, N=400000 points, SR=%d, D=400, P=16, vsize 0.004 x vscale 2"      # both arms


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    arm = CpuArm(args.sr)
    for i in range(args.warmup):
        arm.step(i)
    rays = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        r, _ = arm.step(args.warmup + i)
        rays += r
    dt = time.perf_counter() - t0
    arm.close()
    val = rays / dt / 1e6
    line = dict(impl="reference", metric=METRIC, value=val, unit="Mrays/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=WORKLOAD % args.sr, rays_per_step_sampled=rays // max(args.steps, 1),
                            note="each step = a bounded sample of the frame (see cpu_baseline.sample); value = rays of the sample / time"),
                cpu_baseline=dict(value=val, unit="Mrays/s", cores=arm.workers * arm.threads, kind="port", sample=arm.describe(args.steps, dt)),
                e2e=dict(value=val, unit="Mrays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


def cpu_baseline_leg(sr, budget_s=20.0):
    """Bounded sample of the same workload on the host cores (rank 0, N=1 only)."""
    arm = CpuArm(sr)
    arm.step(0)
    rays, n, t0 = 0, 0, time.perf_counter()
    while True:
        r, _ = arm.step(1 + n)
        rays += r
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 6:
            break
    dt = time.perf_counter() - t0
    arm.close()
    return dict(value=rays / dt / 1e6, unit="Mrays/s", cores=arm.workers * arm.threads, kind="port", sample=arm.describe(n, dt))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="pnb200")
    ap.add_argument("--sr", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", type=str, default="bf16x3", help="bf16x3 (tcgen05, default) | fp32 (CUDA cores)")
    ap.add_argument("--frozen", type=int, default=1, help="1 (default): frozen-cloud pair kernel k_shade_tc8 (point-only layer-1 inputs hoisted per point) | 0: general kernel k_shade_tc7")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus, "WORLD_SIZE %d != --gpus %d (launch N>1 with torchrun)" % (world, args.gpus)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    cfg = scene.CONFIGS["lego_render"]
    cfg.SR = args.sr
    net, pts, opt = harness.build_model(cfg, dev, seed=0, alpha_bias=3.0, pnb_precision=args.precision, pnb_frozen=args.frozen)
    full = scene.make_rays(cfg)
    R_img = full["raydir"].shape[1]
    # global batch = `world` images; ray i of the global batch belongs to rank i % world (interleaved)
    glob = full["raydir"][0].repeat(world, 1)
    mine_host = glob[rank::world].contiguous().pin_memory()
    R = mine_host.shape[0]
    raydir_dev = mine_host.to(dev)
    cam = (list(cfg.campos), torch.eye(3), cfg.near, cfg.far, [1., 1., 1.])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gathered = torch.empty((world, R, 3), dtype=torch.float32, device=dev) if world > 1 else None
    out_host = torch.empty((R, 3), dtype=torch.float32).pin_memory()

    def step_resident():
        with torch.no_grad():
            out = net.render_full(cam[0], raydir_dev, cam[1], cam[2], cam[3], cam[4])
        col = out["coarse_raycolor"][0]
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1, 3), col)
        return col

    def step_e2e():
        rd = mine_host.to(dev, non_blocking=True)
        with torch.no_grad():
            out = net.render_full(cam[0], rd, cam[1], cam[2], cam[3], cam[4])
        col = out["coarse_raycolor"][0]
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1, 3), col)
        out_host.copy_(col, non_blocking=True)
        return col

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps (outside the
        events).  Returns total ms (max over ranks)."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in evs:
            flush.fill_(1)
            a.record()
            fn()
            b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    # workload counters (oracle-independent: the library's own device counters)
    qc = net.neural_points.querier.run_query(net.neural_points.xyz.detach(), raydir_dev, cam[0], cam[2], cam[3], want_counters=True).counters
    gc = net.neural_points.querier.last_grid_counters

    t_w0 = time.time()
    ms_res = timed(step_resident, args.steps)
    t_w1 = time.time()
    clocks = sampler.stop(t_w0, t_w1) if sampler else None

    # dominant kernel alone (shade): CUDA events around the shade launch on the launching stream, same inputs
    from pointnerf_b200 import lib as _lib
    l = _lib.load()
    q = net.last
    mlp = net._mlp.get(net.aggregator)
    ptsd = net.neural_points.points_desc()
    from pointnerf_b200.point_query import make_cam_opts
    o = make_cam_opts(cam[0], cam[1], Rw2c=None, vsize_z=float(opt.vsize[2]), bg_color=cam[4],
                      raydist_mode_unit=opt.raydist_mode_unit)
    stream = torch.cuda.current_stream(dev).cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_kernel(launch):
        ms = []
        for i in range(args.steps + 1):
            flush.fill_(1)
            e0.record()
            launch()
            e1.record()
            torch.cuda.synchronize(dev)
            if i > 0:
                ms.append(e0.elapsed_time(e1))
        return float(np.mean(ms))

    if args.precision == "fp32":
        shade_avg = time_kernel(lambda: _lib.check(l.pnb_shade_forward(
            _lib.C.byref(q.desc), _lib.C.byref(ptsd), _lib.C.byref(mlp), _lib.C.byref(o), net._sigma_rgb.data_ptr(), None, 0, stream),
            "pnb_shade_forward"))
        color_avg = None
        kname = "k_shade_fwd (fp32 CUDA-core kernel: pair MLPs + colour branch)"
        kflops = FLOPS_PER_PAIR * qc["n_pairs"] + FLOPS_PER_SAMPLE * qc["n_valid"]
    else:
        frozen = bool(net.frozen_ok)
        pre_ptr = net._point_pre(mlp, ptsd, stream).data_ptr() if frozen else None

        def tc(flags):
            _lib.check(l.pnb_shade_forward_tc(_lib.C.byref(q.desc), _lib.C.byref(ptsd), _lib.C.byref(mlp), net._mlp.packed.data_ptr(), pre_ptr,
                                              _lib.C.byref(o), net._sigma_rgb.data_ptr(), net._tc_ws.data_ptr(), net._tc_ws.numel(),
                                              net._max_valid, flags | (_lib.TC_FROZEN if frozen else 0), net._err.data_ptr(), stream),
                       "pnb_shade_forward_tc")
        shade_avg = time_kernel(lambda: tc(_lib.TC_PAIRS))
        color_avg = time_kernel(lambda: tc(_lib.TC_COLOR))
        kname = ("k_shade_tc8 (frozen cloud: layer-1 point inputs hoisted)" if frozen else "k_shade_tc7") + \
            " + k_pack_* (tcgen05 BF16x3 pair MLPs 284-256-256 | 263-256-256 + alpha + K-reduction)"
        kflops = FLOPS_PER_PAIR * qc["n_pairs"]
        net.check_errors()

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    total_rays = R * world * args.steps
    value = total_rays / (ms_res * 1e-3) / 1e6
    e2e_val = total_rays / (ms_e2e * 1e-3) / 1e6
    pk = peaks()
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of this same command
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "r02_ncu_dram_traffic.json")
    if os.path.exists(tfile) and args.precision != "fp32" and world == 1 and args.sr == 24:
        traffic = json.load(open(tfile)).get("k_shade_tc8" if net.frozen_ok else "k_shade_tc7", None)
    flops = kflops
    achieved = flops / (shade_avg * 1e-3) / 1e12
    peak = pk["bf16_tflops"]
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(args.sr)
    if rank == 0:
        line = dict(
            metric=METRIC, value=value, unit="Mrays/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=ms_res / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=("f32 (bf16x3 split on tcgen05, fp32 accumulate)" if args.precision != "fp32" else "f32"),
            data="synthetic",
            config=dict(workload=WORKLOAD % args.sr,
                        rays_per_step_per_gpu=R, parallelism="rays interleave-sharded x%d, points replicated%s" % (world, ", all-gather of colours" if world > 1 else ""),
                        l2="flushed between timed steps (256 MiB write, outside the CUDA events)",
                        workload_counters=dict(hit_rays=qc["R2"], valid_samples=qc["n_valid"], valid_pairs=qc["n_pairs"],
                                               candidate_samples=qc["n_cand"], occupied_voxels=gc["n_occ"], max_pts_per_voxel=gc["max_pts"])),
            e2e=dict(value=e2e_val, unit="Mrays/s", h2d_bytes_per_step=int(mine_host.numel() * 4 * world),
                     d2h_bytes_per_step=int(out_host.numel() * 4 * world), ms_per_step=ms_e2e / args.steps),
            # + colour kernel + the 3 row-packing kernels of the tcgen05 path; cf. profiles/r02_ncu_launches_summary.txt
            gpu_launches=(LAUNCHES_PER_STEP + (4 if args.precision != "fp32" else 0)) * args.steps,
            clocks=clocks,
            roofline=dict(bound="tensor", kernel=kname, achieved=achieved, peak=peak, unit="TFLOP/s",
                          frac=achieved / peak, traffic=traffic, peak_source="%s bf16 cuBLAS burst (MEASURED_PEAKS.json)" % pk["source"],
                          algorithmic_flops_per_launch=flops, kernel_ms=shade_avg,
                          kernel_share_of_step=shade_avg / (ms_res / args.steps),
                          issued_mma_flops_per_launch=(3 * flops if args.precision != "fp32" else None),
                          tensor_pipe_frac_issued=(3 * achieved / peak if args.precision != "fp32" else None),
                          colour_branch_kernel_ms=color_avg),
            cpu_baseline=cpu,
        )
        if _RETRY_NOTE:
            line["retried_after"] = _RETRY_NOTE        # the first attempt tripped the in-kernel watchdog (see __main__)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


_RETRY_NOTE = None

if __name__ == "__main__":
    try:
        rc = main()
    except Exception as e:  # noqa: BLE001
        from pointnerf_b200.lib import PnbError
        single = int(os.environ.get("WORLD_SIZE", "1")) == 1       # with several ranks a one-sided retry would hang the collectives
        if isinstance(e, PnbError) and "time-out" in str(e) and single:
            # every in-kernel mbarrier wait is bounded (2 s): a protocol stall surfaces as this error instead of a hung GPU.
            # Never seen on the final pipeline; if it ever happens the run is re-measured once and the JSON line says so.
            print("bench.py: %s -- re-measuring once" % e, file=sys.stderr)
            _RETRY_NOTE = str(e)
            rc = main()
        else:
            raise
    sys.exit(rc)
