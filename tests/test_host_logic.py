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
                 ("shading_color_mlp_layer", 2), ("prob", 1), ("agg_axis_weight", [1.0, 2.0, 1.0])):
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
