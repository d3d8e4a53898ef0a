"""CPU: the host-side assembly logic of pointnerf_b200.runner with a stub network (the GPU parity of the two entry points is in
tests/test_gpu_runner.py)."""
import numpy as np
import pytest
import torch

from pointnerf_b200 import runner


class _StubNet:
    """Colours a ray by its direction; rays with x < 0 "miss" (no neighbour) exactly like the drop-in forward()/render_full()."""

    def __init__(self, sr=4):
        self.sr = sr
        self.calls = 0

    def _full(self, raydir):
        d = raydir[0]
        mask = (d[:, 0] >= 0).to(torch.int8)
        col = torch.where(mask[:, None] > 0, d * 0.25 + 0.5, torch.ones_like(d))          # background = white
        op = mask[:, None].float().expand(-1, self.sr) * 0.5
        return mask, col, op

    def check_errors(self):
        self.checked = getattr(self, "checked", 0) + 1

    def render_full(self, campos, raydir, camrotc2w, near, far, bg_color, t=None):
        mask, col, op = self._full(raydir)
        return dict(coarse_raycolor=col[None], coarse_point_opacity=op[None], coarse_is_background=(1 - mask.float())[None, :, None], ray_mask=mask[None])

    def __call__(self, campos, raydir, bg_color=None, camrotc2w=None, near=None, far=None, **kw):
        self.calls += 1
        mask, col, op = self._full(raydir)
        sel = mask > 0
        return dict(coarse_raycolor=col[sel][None], coarse_point_opacity=op[sel][None], ray_mask=mask[None])


def _data(h, w):
    g = torch.Generator().manual_seed(0)
    return dict(campos=torch.zeros(1, 3), raydir=torch.randn(1, h * w, 3, generator=g), camrotc2w=torch.eye(3)[None],
                near=torch.tensor([[2.0]]), far=torch.tensor([[6.0]]), bg_color=torch.ones(1, 3))


def test_whole_image_and_chunk_loop_assemble_the_same_pixels():
    h, w = 7, 9
    net, data = _StubNet(), _data(h, w)
    whole = runner.render_image(net, data, h, w)
    chunked = runner.render_image_chunked(net, data, h, w, chunk_size=16)         # ragged last chunk (63 = 3*16 + 15)
    assert net.calls == 4
    assert whole["coarse_raycolor"].shape == (h, w, 3) and whole["coarse_point_opacity"].shape == (h, w, 4)
    assert whole["ray_mask"].shape == (h, w) and whole["coarse_is_background"].shape == (h, w, 1)
    assert np.array_equal(whole["coarse_raycolor"].numpy(), chunked["coarse_raycolor"])
    miss = whole["ray_mask"].numpy() == 0
    assert miss.any() and np.all(chunked["coarse_raycolor"][miss] == 1.0)


def test_render_image_checks_the_ray_count():
    with pytest.raises(ValueError):
        runner.render_image(_StubNet(), _data(4, 4), 4, 5)
