"""Multi-GPU plumbing of the hot path: one process per GPU, neural point cloud / voxel grid / MLP
replicated, rays interleave-sharded (ray i -> rank i % world).  Two collectives only:
  * render:        one all_gather of the [R/world, 3] colours per image (7.7 MB @800x800), re-interleaved;
  * optimisation:  one all_reduce(sum) over the point-feature and MLP gradients per step -- every rank then takes
                   the same dense Adam step (the reference's optimisers, mvs_points_volumetric_model.py:87-91), so
                   replicas stay bit-identical without a broadcast.
Backend: NCCL over NVLink on the GPU box; the same code runs on gloo/CPU tensors for the host-logic tests.
"""
import torch
import torch.distributed as dist


def shard_indices(R, rank, world, device="cpu"):
    """Ray ids of this rank: rank, rank+world, ...  (hit density is spatially clustered -> interleave, not blocks)."""
    return torch.arange(rank, R, world, device=device)


def padded_shard_len(R, world):
    return (R + world - 1) // world


def gather_interleaved(local, R, world, group=None):
    """local: [ceil(R/world), C] rows of this rank (rows beyond its share are padding).  Returns the full [R, C]
    tensor in original ray order on every rank (one all_gather)."""
    n = padded_shard_len(R, world)
    assert local.shape[0] == n, "pad the local shard to ceil(R/world) rows"
    out = torch.empty((world, n) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if world > 1:
        dist.all_gather_into_tensor(out.view(-1, *local.shape[1:]), local.contiguous(), group=group)
    else:
        out[0] = local
    # out[g, j] is ray j*world + g
    full = out.transpose(0, 1).reshape(n * world, *local.shape[1:])
    return full[:R]


def pad_rows(t, n):
    if t.shape[0] == n:
        return t
    pad = torch.zeros((n - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    return torch.cat([t, pad], dim=0)


def allreduce_gradients(params, world, group=None, average=False):
    """One flat all_reduce(sum) over the gradients of `params` (point features + MLP), written back in place."""
    grads = [p.grad for p in params if p.grad is not None]
    if world <= 1 or not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel()
