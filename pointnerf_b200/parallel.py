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


# ---- point growing across ranks ("Grow / prune"): probe_hole (run/train_ft.py:417-530) renders whole
# training frames -> shard FRAMES across ranks, then one variable-length all-gather (counts, then payload) of the new
# points; every rank appends them in rank order, so the replicated point clouds stay identical without a broadcast.
# Prune needs no communication: it is a deterministic function of the replicated points_conf.

def shard_frames(frame_ids, rank, world):
    """Frames this rank probes: round-robin over the (already shuffled / ranked) id list, as rays are."""
    return list(frame_ids)[rank::world]


def allgather_varlen(t, world, group=None):
    """t: [n_rank, C] (n differs per rank).  Returns the [sum n, C] concatenation in rank order on every rank:
    one all_gather of the counts, one of the payload padded to the largest count."""
    if world <= 1:
        return t
    cnt = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = torch.empty(world, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts = counts.tolist()
    m = max(counts)
    if m == 0:
        return t
    buf = torch.empty((world, m) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(buf.view(world * m, *t.shape[1:]), pad_rows(t.contiguous(), m), group=group)
    return torch.cat([buf[g, :counts[g]] for g in range(world)], dim=0)


def allgather_new_points(add_xyz, add_embedding, add_color, add_dir, add_conf, world, group=None):
    """The five tensors probe_hole returns (train_ft.py:530), gathered as ONE payload [n, 3+C+3+3+1] so that a rank
    cannot interleave them differently; returns them split again, identical on every rank."""
    c = add_embedding.shape[1]
    packed = torch.cat([add_xyz, add_embedding, add_color, add_dir, add_conf], dim=1)
    allp = allgather_varlen(packed, world, group=group)
    return allp[:, 0:3], allp[:, 3:3 + c], allp[:, 3 + c:6 + c], allp[:, 6 + c:9 + c], allp[:, 9 + c:10 + c]


# ---- sparse gradient exchange ("sparse alternative"): a step touches at most 3600*SR*K point rows, so for large
# clouds (N = 5 M: 780 MB of dense gradients per all-reduce) exchanging only the touched rows is far cheaper.

def allreduce_rows_sparse(grad, world, group=None):
    """grad: [N, C] or [1, N, C] dense gradient of a per-point tensor, non-zero only on the rows this rank touched.  In place:
    grad <- sum over ranks, identical bit for bit on every rank (each rank's rows are unique, and the ranks' contributions are
    added in rank order - no atomics on duplicate indices).  One variable-length all-gather of [index | row] pairs.
    Returns the number of rows exchanged."""
    if world <= 1:
        return 0
    g2 = grad.view(-1, grad.shape[-1])
    idx = torch.nonzero(g2.abs().sum(dim=1) > 0)[:, 0]
    payload = torch.cat([idx.to(g2.dtype)[:, None], g2[idx]], dim=1) if idx.numel() else g2.new_zeros((0, g2.shape[1] + 1))
    assert g2.shape[0] < (1 << 24) or g2.dtype == torch.float64, "row indices travel as fp32: N must stay below 2^24"
    cnt = torch.tensor([payload.shape[0]], dtype=torch.int64, device=g2.device)
    counts = torch.empty(world, dtype=torch.int64, device=g2.device)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts = counts.tolist()
    allp = allgather_varlen(payload, world, group=group)
    g2.zero_()
    off = 0
    for n in counts:                                       # rank order; indices are unique within a rank's block
        if n:
            blk = allp[off:off + n]
            g2.index_add_(0, blk[:, 0].to(torch.int64), blk[:, 1:])
        off += n
    return int(sum(counts))
