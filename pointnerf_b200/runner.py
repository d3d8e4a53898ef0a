"""Whole-image entry points either side of the seam.

The reference renders an image as a host loop over `random_sample_size**2`-ray chunks
(`/root/reference/run/render_vid.py:45-71`, `/root/reference/run/train_ft.py:283-320`): per chunk `model.set_input`,
`model.test()`, `.cpu().numpy()` of every `*color` visual into a `[H*W,3]` host array, then `fill_invalid`
(`neural_points_volumetric_model.py:87-123`) — 278 launches + host syncs for an 800x800 frame at the shipped 2304-ray
chunk.  `render_image` is the same computation as ONE call (rays of the whole frame resident on the device, fill_invalid
applied in-kernel, no host sync until the caller reads the result); `render_image_chunked` reproduces the reference loop
through the drop-in `forward()` and is what the parity test compares it with (a ray's colour is bit-identical whichever
rays share its call).
"""
import numpy as np
import torch


def _near_far(near, far):
    n = float(torch.min(near)) if isinstance(near, torch.Tensor) else float(near)
    f = float(torch.max(far)) if isinstance(far, torch.Tensor) else float(far)
    return n, f


def render_image(net, data, height, width, check=True):
    """data: the dict a reference dataset item carries (`campos [1,3]`, `raydir [1,H*W,3]`, `camrotc2w [1,3,3]`, `near`, `far`,
    `bg_color`).  Returns device tensors: `coarse_raycolor [H,W,3]`, `coarse_point_opacity [H,W,SR]`,
    `coarse_is_background [H,W,1]`, `ray_mask [H,W]` (all rays, background filled as fill_invalid does).
    check=True (default): the device status of the frame is read before returning (one stream synchronisation; the caller is
    about to read the image anyway) and a frame whose shading workspace was too small is rendered again with the exact size;
    check=False: fully asynchronous, the status surfaces at a later call or through `net.check_errors()`."""
    near, far = _near_far(data["near"], data["far"])
    raydir = data["raydir"]
    if raydir.shape[1] != height * width:
        raise ValueError("render_image: %d rays for a %dx%d image" % (raydir.shape[1], height, width))
    bg = data.get("bg_color", None)
    from .lib import PnbOverflow
    with torch.no_grad():
        for attempt in range(2):
            out = net.render_full(data["campos"], raydir, data["camrotc2w"], near, far, bg if bg is not None else torch.zeros(3))
            if not check:
                break
            try:
                net.check_errors()
                break
            except PnbOverflow:
                if attempt == 1:
                    raise
    sr = out["coarse_point_opacity"].shape[-1]
    return dict(coarse_raycolor=out["coarse_raycolor"][0].view(height, width, 3),
                coarse_point_opacity=out["coarse_point_opacity"][0].view(height, width, sr),
                coarse_is_background=out["coarse_is_background"][0].view(height, width, 1),
                ray_mask=out["ray_mask"][0].view(height, width))


def render_image_chunked(net, data, height, width, chunk_size):
    """The reference chunk loop (render_vid.py:45-71) through the drop-in `forward()` + fill_invalid on the host:
    returns `{"coarse_raycolor": np.ndarray [H,W,3]}` exactly as the reference assembles `visuals`."""
    near, far = _near_far(data["near"], data["far"])
    raydir = data["raydir"]
    total = height * width
    bg = data.get("bg_color", None)
    bgv = (bg if bg is not None else torch.zeros(3)).reshape(-1)[:3].cpu().numpy().astype(np.float32)
    img = np.zeros((total, 3), dtype=np.float32)
    with torch.no_grad():
        for start in range(0, total, chunk_size):
            end = min(start + chunk_size, total)
            out = net(data["campos"], raydir[:, start:end, :], bg_color=bg, camrotc2w=data["camrotc2w"],
                      near=near, far=far)
            mask = out["ray_mask"][0].bool().cpu().numpy()
            chunk = np.tile(bgv[None, :], (end - start, 1))                      # fill_invalid: background where no neighbour
            chunk[mask] = out["coarse_raycolor"][0].cpu().numpy()
            img[start:end] = chunk
    return dict(coarse_raycolor=img.reshape(height, width, 3))
