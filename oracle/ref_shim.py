"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference Python hot path.

Only usable where /root/reference exists (this build container, never the GPU box).  It is
used by oracle/make_golden.py to (a) pin oracle/shade_oracle.py and oracle/query_oracle.c
against the reference's own code and (b) generate the committed fixtures in tests/golden/.

What it does:
  * stubs import-time-only dependencies that are not installed here (matplotlib, imageio,
    scipy.special.{sph_harm,lpmn,lpmv});
  * replaces torch.utils.cpp_extension.load so that importing the reference's
    models/neural_points/point_query.py (which JIT-builds query_worldcoords.{cpp,cu} at
    import, point_query.py:15-22) yields a module object whose only function,
    woord_query_grid_point_index (query_worldcoords.cpp:34-82), is served by the serial C
    restatement in oracle/query_oracle.c.  Everything else (lighting_fast_querier,
    NeuralPoints, PointAggregator, ray_march, NeuralPointsRayMarching) is the reference's
    own code, executed as-is on CPU.
"""
import os
import sys
import types
import argparse

REF = os.environ.get("PNB_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "models", "neural_points"))


_installed = False


def install(query_op=None):
    """Make `import models...` resolve to the reference tree.  query_op: callable with the
    18-argument signature of woord_query_grid_point_index (or None -> oracle C port)."""
    global _installed
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    if _installed:
        return
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "imageio"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    import scipy.special as sp
    for n in ("sph_harm", "lpmn", "lpmv"):
        if not hasattr(sp, n):
            setattr(sp, n, None)

    import torch.utils.cpp_extension as cpp_ext

    fake_ext = types.ModuleType("query_worldcoords_cuda")
    if query_op is None:
        from oracle import query_oracle
        query_op = query_oracle.woord_query_grid_point_index
    fake_ext.woord_query_grid_point_index = query_op

    real_load = cpp_ext.load

    def fake_load(name, sources, **kw):
        if name == "query_worldcoords_cuda":
            return fake_ext
        return real_load(name, sources, **kw)

    cpp_ext.load = fake_load
    _installed = True


def make_opt(extra_flags=(), is_train=False):
    """argparse namespace built by the reference's own option registration
    (neural_points_volumetric_model.py:10-70) with the shipped hot-path flags."""
    install()
    from models.neural_points_volumetric_model import NeuralPointsVolumetricModel
    parser = argparse.ArgumentParser()
    NeuralPointsVolumetricModel.modify_commandline_options(parser, is_train)
    flags = ("--K 8 --NN 2 --SR 24 --P 16 --max_o 100000 --vscale 2 2 2 --kernel_size 3 3 3 "
             "--query_size 3 3 3 --vsize 0.004 0.004 0.004 --radius_limit_scale 4 --z_depth_dim 400 "
             "--point_features_dim 32 --agg_dist_pers 20 --agg_distance_kernel linear "
             "--agg_intrp_order 2 --apply_pnt_mask 1 --num_feat_freqs 3 --dist_xyz_freq 5 "
             "--dist_xyz_deno 0 --num_viewdir_freqs 4 --num_pos_freqs 10 "
             "--shading_feature_mlp_layer1 2 --shading_feature_mlp_layer2 0 --shading_feature_mlp_layer3 2 "
             "--shading_alpha_mlp_layer 1 --shading_color_mlp_layer 4 --shading_feature_num 256 "
             "--act_type LeakyReLU --point_conf_mode 1 --point_dir_mode 1 --point_color_mode 1 "
             "--agg_feat_xyz_mode None --agg_alpha_xyz_mode None --agg_color_xyz_mode None "
             "--raydist_mode_unit 1 --which_render_func radiance --which_blend_func alpha "
             "--which_tonemap_func off --bg_color white --wcoord_query -1 --load_points 0 --num_point 0 "
             "--ranges -1.1 -1.1 -1.1 1.1 1.1 1.1 --which_agg_model viewmlp "
             "--zero_one_loss_items conf_coefficient").split()
    opt, _ = parser.parse_known_args(flags + list(extra_flags))
    opt.is_train = is_train
    return opt
