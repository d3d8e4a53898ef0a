"""Checkpoint wire format: our module's state_dict IS the reference's `{epoch}_net_ray_marching.pth`
layout (tests/golden/checkpoint_layout.json was written from the reference module's own state_dict by oracle/make_golden.py),
plus the loader's validation and a save/load round trip.  CPU only."""
import json
import os

import pytest
import torch

from pointnerf_b200 import checkpoint, harness, scene


@pytest.fixture()
def layout(golden_dir):
    return json.load(open(os.path.join(golden_dir, "checkpoint_layout.json")))


def _net():
    cfg = scene.CONFIGS["tiny"]
    net, pts, opt = harness.build_model(cfg, "cpu", seed=3)
    return cfg, net, pts, opt


def test_state_dict_is_the_reference_layout(layout):
    cfg, net, pts, opt = _net()
    ours = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in net.state_dict().items()}
    assert layout["n_points"] == pts["xyz"].shape[0]
    assert ours == layout["layout"]


def test_save_load_round_trip(tmp_path):
    cfg, net, pts, opt = _net()
    net_path, states_path = checkpoint.save_checkpoint(net, str(tmp_path), "200000", other_states=dict(epoch_count=7, total_steps=200000))
    assert os.path.basename(net_path) == "200000_net_ray_marching.pth" and os.path.basename(states_path) == "200000_states.pth"
    assert torch.load(states_path)["total_steps"] == 200000
    net2 = checkpoint.load_checkpoint(net_path, opt, "cpu")
    a, b = net.state_dict(), net2.state_dict()
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert net2.neural_points.points_embeding.requires_grad and not net2.neural_points.xyz.requires_grad


def test_loader_accepts_dataparallel_prefix_and_rejects_bad_files(tmp_path):
    cfg, net, pts, opt = _net()
    sd = {("module." + k): v.detach().clone() for k, v in net.state_dict().items()}
    net2 = checkpoint.load_checkpoint(sd, opt, "cpu")
    assert torch.equal(net2.aggregator.block1[0].weight, net.aggregator.block1[0].weight)
    good = {k: v.detach().clone() for k, v in net.state_dict().items()}
    bad = dict(good); bad.pop("neural_points.points_conf")
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.load_checkpoint(bad, opt, "cpu")
    bad = dict(good); bad["neural_points.points_dir"] = bad["neural_points.points_dir"][:, :-1]
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.load_checkpoint(bad, opt, "cpu")
    bad = dict(good); bad["aggregator.block1.0.wieght"] = bad.pop("aggregator.block1.0.weight")     # a strict=False load would ignore this
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.load_checkpoint(bad, opt, "cpu")
    bad = dict(good); bad["neural_points.eulers"] = torch.zeros(3)
    with pytest.raises(NotImplementedError):
        checkpoint.load_checkpoint(bad, opt, "cpu")


def test_rw2c_round_trip(tmp_path):
    """A non-identity Rw2c is part of the reference state dict (neural_points.py:463-467: nn.Parameter): load -> save -> load keeps it."""
    cfg, net, pts, opt = _net()
    a = 0.3
    R = torch.tensor([[1.0, 0.0, 0.0], [0.0, float(torch.cos(torch.tensor(a))), -float(torch.sin(torch.tensor(a)))],
                      [0.0, float(torch.sin(torch.tensor(a))), float(torch.cos(torch.tensor(a)))]])
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    assert "neural_points.Rw2c" not in sd                      # default identity: a plain tensor, not in the state dict (as the reference)
    sd["neural_points.Rw2c"] = R
    net2 = checkpoint.load_checkpoint(sd, opt, "cpu")
    assert isinstance(net2.neural_points.Rw2c, torch.nn.Parameter) and not net2.neural_points.Rw2c.requires_grad
    net_path, _ = checkpoint.save_checkpoint(net2, str(tmp_path), "latest")
    saved = torch.load(net_path)
    assert torch.equal(saved["neural_points.Rw2c"], R)
    net3 = checkpoint.load_checkpoint(net_path, opt, "cpu")
    assert torch.equal(net3.neural_points.Rw2c.detach(), R)
    # back to the default: a later set_points without Rw2c drops the parameter again
    p = net3.neural_points
    p.set_points(p.xyz.detach(), p.points_embeding.detach(), points_color=p.points_color.detach(), points_dir=p.points_dir.detach(),
                 points_conf=p.points_conf.detach())
    assert "neural_points.Rw2c" not in net3.state_dict() and torch.equal(p.Rw2c, torch.eye(3))
