"""CPU: host-side mirror of the reference interface (hyper-parameters, t table, option gate)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import pipeline
from pointnerf_b200 import harness, scene
from pointnerf_b200.point_query import host_t_table, lighting_fast_querier
from pointnerf_b200.ray_marching import check_opt


def test_t_table_matches_reference_fixture(golden_dir):
    fx = np.load(os.path.join(golden_dir, "hyper.npz"))
    for (near, far, D) in ((2.0, 6.0, 400), (0.0, 3.5, 400), (0.1, 8.0, 400), (2.0, 6.0, 37)):
        ref = fx["t_%g_%g_%d" % (near, far, D)]
        assert np.array_equal(host_t_table(near, far, D).numpy(), ref)
        assert np.array_equal(pipeline.t_table(near, far, D), ref)


@pytest.mark.parametrize("name", ["tiny", "chair_plumbing", "lego_render"])
def test_hyperparameters_match_reference_fixture(name, golden_dir):
    fx = np.load(os.path.join(golden_dir, "hyper.npz"))
    cfg = scene.CONFIGS[name]
    pts = scene.make_points(cfg)
    opt = harness.make_opt(cfg)
    q = lighting_fast_querier(torch.device("cpu"), opt)   # constructor only loads the library
    rng_t, vsz, sdim = q.get_hyperparameters(opt.vsize, pts["xyz"][None], ranges=opt.ranges)
    assert np.array_equal(rng_t.numpy(), fx[name + ".ranges6"])
    assert np.array_equal(sdim, fx[name + ".scaled_vdim"])
    assert np.array_equal(q.scaled_vsize_np, fx[name + ".scaled_vsize"])
    assert np.array_equal(q.radius_limit_np, fx[name + ".radius_limit"])
    o_rng, o_svs, o_dim = pipeline.hyperparameters(pts["xyz"], opt.vsize, opt.vscale, opt.kernel_size, opt.ranges)
    assert np.array_equal(o_rng, fx[name + ".ranges6"]) and np.array_equal(o_dim, fx[name + ".scaled_vdim"])


def test_option_gate_rejects_unimplemented_values():
    cfg = scene.CONFIGS["tiny"]
    check_opt(harness.make_opt(cfg))
    check_opt(harness.make_opt(cfg, agg_intrp_order=1))          # alpha_branch on the aggregated feature
    for k, v in (("agg_intrp_order", 0), ("agg_dist_pers", 10), ("act_type", "ReLU"), ("num_feat_freqs", 0),
                 ("shading_color_mlp_layer", 2), ("prob", 2), ("agg_axis_weight", [1.0, 2.0, 1.0])):
        with pytest.raises(NotImplementedError):
            check_opt(harness.make_opt(cfg, **{k: v}))
    with pytest.raises(NotImplementedError):
        lighting_fast_querier(torch.device("cpu"), harness.make_opt(cfg, K=16))
    with pytest.raises(NotImplementedError):
        lighting_fast_querier(torch.device("cpu"), harness.make_opt(cfg, inverse=1))


def test_scene_generator_is_seeded():
    cfg = scene.CONFIGS["tiny"]
    a, b = scene.make_points(cfg), scene.make_points(cfg)
    for k in a:
        assert torch.equal(a[k], b[k])
    r = scene.make_rays(cfg)
    assert r["raydir"].shape == (1, cfg.H * cfg.W, 3) and torch.all(r["raydir"][..., 2] == 1)


def test_prune_and_grow_points_mirror_reference_semantics():
    """neural_points.py:347-399: prune keeps conf >= thresh, grow appends; parameter names / shapes / grad flags kept."""
    from pointnerf_b200.ray_marching import NeuralPoints
    cfg = scene.CONFIGS["tiny"]
    opt = harness.make_opt(cfg)
    pts = scene.make_points(cfg)
    npn = NeuralPoints(opt, torch.device("cpu"))
    npn.set_points(pts["xyz"], pts["embedding"], points_color=pts["color"], points_dir=pts["dir"], points_conf=pts["conf"])
    N = pts["xyz"].shape[0]
    keep = pts["conf"][0, :, 0] >= 0.5
    removed = npn.prune(0.5)
    assert removed == int((~keep).sum()) and npn.xyz.shape == (int(keep.sum()), 3)
    assert torch.equal(npn.points_embeding[0], pts["embedding"][0][keep]) and npn.points_embeding.requires_grad
    assert not npn.xyz.requires_grad and npn.points_conf.shape == (1, int(keep.sum()), 1)
    n0 = npn.xyz.shape[0]
    npn.grow_points(torch.zeros(5, 3), torch.ones(5, 32), torch.ones(5, 3), torch.ones(5, 3), torch.full((5, 1), 0.3))
    assert npn.xyz.shape[0] == n0 + 5 and npn.points_color.shape == (1, n0 + 5, 3)
    assert sorted(k for k, _ in npn.named_parameters()) == ["points_color", "points_conf", "points_dir", "points_embeding", "xyz"]


def test_segmented_scan_order_depends_on_count_only():
    """Design invariant behind the packed-row K-reduction (csrc/shade_tc.cu: seg_scan8): a sample's rows are 1..8 consecutive
    lanes anywhere in a 32-lane quadrant; the 3-step segmented shuffle scan must add them in an order that depends on the
    neighbour count only, so that a ray's colour is bit-identical whichever samples share its quadrant.  Restated in numpy
    (fp32, same dataflow as the warp shuffles) and checked for every (offset, count) against the offset-0 result."""
    rng = np.random.default_rng(0)

    def seg_scan(vals, st_of_lane):                     # vals[32] fp32, st_of_lane[32] = first lane of each lane's segment
        v = vals.astype(np.float32).copy()
        for d in (1, 2, 4):
            up = np.concatenate([np.zeros(d, np.float32), v[:-d]])           # __shfl_up_sync(v, d)
            take = (np.arange(32) - d) >= st_of_lane
            v = np.where(take, (v + up).astype(np.float32), v)
        return v

    for cnt in range(1, 9):
        x = (rng.standard_normal(cnt) * 10 ** rng.uniform(-3, 3, cnt)).astype(np.float32)
        ref = None
        for st in range(0, 32 - cnt + 1):
            vals = rng.standard_normal(32).astype(np.float32)               # neighbours' rows hold other samples' values
            vals[st:st + cnt] = x
            st_of = np.arange(32)                                            # every other lane: its own 1-row segment
            st_of[st:st + cnt] = st
            if st > 0:
                st_of[:st] = 0                                               # a longer foreign segment right before ours
            out = seg_scan(vals, st_of)[st + cnt - 1]
            if ref is None:
                ref = out
            assert out.tobytes() == ref.tobytes(), (cnt, st)
        assert abs(float(ref) - float(np.sum(x.astype(np.float64)))) <= 1e-5 * float(np.sum(np.abs(x)) + 1e-30)


def test_first_fit_row_packing_properties():
    """Design invariants of the row packing (csrc/shade_tc.cu: k_pack_quads, restated in Python line by line): every valid sample
    is placed exactly once, a quadrant never exceeds 32 rows, quadrants are contiguous ranges of the permutation, the result
    is deterministic, and a lego-like neighbour-count distribution fills > 97 % of the rows (in-order packing: ~92 %)."""
    PACK_S, PACK_WIN = 512, 64

    def pack(counts, win):
        n_valid = len(counts)
        vorder, vcntp, quad_first = np.zeros(n_valid, np.int64), np.zeros(n_valid, np.int64), []
        for i0 in range(0, n_valid, PACK_S):
            c = counts[i0:i0 + PACK_S].copy()
            n, pos, emitted = len(c), 0, 0
            while pos < n:
                quad_first.append(i0 + emitted)
                rows = 0
                for i in range(pos, min(n, pos + win)):
                    if rows >= 32:
                        break
                    if c[i] != 0 and rows + c[i] <= 32:
                        vorder[i0 + emitted] = i0 + i
                        vcntp[i0 + emitted] = c[i]
                        rows += c[i]
                        c[i] = 0
                        emitted += 1
                while pos < n and c[pos] == 0:
                    pos += 1
        quad_first.append(n_valid)
        return vorder, vcntp, np.asarray(quad_first)

    rng = np.random.default_rng(1)
    counts = np.minimum(8, np.maximum(1, rng.poisson(6.5, size=5000))).astype(np.int64)      # mean ~6, many saturated at K=8
    counts[:600] = rng.integers(1, 9, size=600)
    vorder, vcntp, qf = pack(counts, PACK_WIN)
    assert sorted(vorder.tolist()) == list(range(len(counts)))                   # a permutation: every sample exactly once
    assert np.array_equal(vcntp, counts[vorder])
    rows = np.array([vcntp[a:b].sum() for a, b in zip(qf[:-1], qf[1:])])
    assert rows.max() <= 32 and rows.min() >= 1 and np.all(np.diff(qf) >= 1)
    for a in range(0, len(counts), PACK_S):                                      # super-chunks are packed independently
        assert a in set(qf.tolist()) and set(vorder[a:a + PACK_S].tolist()) == set(range(a, min(a + PACK_S, len(counts))))
    v2, c2, q2 = pack(counts, PACK_WIN)
    assert np.array_equal(v2, vorder) and np.array_equal(q2, qf)
    fill = counts.sum() / (32.0 * (len(qf) - 1))
    assert fill > 0.97, fill
    # in-order greedy packing (what a look-ahead of 0 would do) for comparison
    nq, r = 0, 0
    for ci in counts:
        if r + ci > 32:
            nq, r = nq + 1, 0
        r += ci
    assert counts.sum() / (32.0 * (nq + 1)) < fill
