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
    for k, v in (("agg_intrp_order", 1), ("agg_dist_pers", 10), ("act_type", "ReLU"), ("num_feat_freqs", 0),
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
