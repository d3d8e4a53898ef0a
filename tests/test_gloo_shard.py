"""CPU, world_size 2 over gloo: the N>1 host logic (interleaved ray sharding, colour all-gather, gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointnerf_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, R, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = torch.rand(R, 3, generator=g)                      # what a single process would render
        ids = parallel.shard_indices(R, rank, world)
        local = parallel.pad_rows(full[ids], parallel.padded_shard_len(R, world))
        out = parallel.gather_interleaved(local, R, world)
        ok_gather = torch.equal(out, full)
        # gradient all-reduce: each rank holds the gradient of its own rays; the sum must equal the unsharded one
        w = torch.nn.Parameter(torch.ones(5, 3))
        feats = torch.rand(R, 5, generator=g)
        (feats[ids] @ w).sum().backward()
        n = parallel.allreduce_gradients([w], world)
        ref = feats.sum(0)[:, None].expand(5, 3)
        ok_grad = torch.allclose(w.grad, ref, rtol=1e-6, atol=1e-6) and n == 15
        # point growing: each rank probed its own frames and found a different number of new points
        assert parallel.shard_frames(list(range(7)), rank, world) == ([0, 2, 4, 6] if rank == 0 else [1, 3, 5])
        n_new = [3, 0] if R == 10 else [2, 5]
        gg = torch.Generator().manual_seed(100 + rank)
        mine = [torch.rand(n_new[rank], c, generator=gg) for c in (3, 32, 3, 3, 1)]
        got = parallel.allgather_new_points(*mine, world)
        exp = [[], [], [], [], []]
        for g_ in range(world):
            g2 = torch.Generator().manual_seed(100 + g_)
            for i, c in enumerate((3, 32, 3, 3, 1)):
                exp[i].append(torch.rand(n_new[g_], c, generator=g2))
        ok_grow = all(torch.equal(a, torch.cat(e, 0)) for a, e in zip(got, exp)) and got[0].shape[0] == sum(n_new)
        # sparse gradient exchange: each rank touched a few (overlapping) point rows
        N, C = 50, 4
        gs = torch.Generator().manual_seed(7)
        dense = [torch.zeros(N, C) for _ in range(world)]
        for g_ in range(world):
            rows = torch.randperm(N, generator=gs)[:9 + 3 * g_]
            dense[g_][rows] = torch.randn(len(rows), C, generator=gs)
        mine = dense[rank].clone()[None]                              # the [1, N, C] shape of points_embeding.grad
        n_rows = parallel.allreduce_rows_sparse(mine, world)
        expect = dense[0].clone()
        for g_ in range(1, world):
            expect += dense[g_]                                       # same rank order -> bit-identical
        ok_sparse = torch.equal(mine[0], expect) and n_rows == sum(int((d.abs().sum(1) > 0).sum()) for d in dense)
        ret[rank] = (ok_gather, ok_grad and ok_grow and ok_sparse)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_shard_gather_allreduce():
    for R in (10, 11):                                            # even and ragged split
        mgr = mp.Manager()
        ret = mgr.dict()
        port = _free_port()
        mp.spawn(_worker, args=(2, port, R, ret), nprocs=2, join=True)
        assert dict(ret) == {0: (True, True), 1: (True, True)}, (R, dict(ret))


def test_single_process_paths():
    full = torch.arange(21.0).view(7, 3)
    assert torch.equal(parallel.gather_interleaved(full, 7, 1), full)
    assert parallel.shard_indices(7, 1, 3).tolist() == [1, 4]
    assert parallel.padded_shard_len(7, 3) == 3
    t = torch.rand(4, 5)
    assert parallel.allgather_varlen(t, 1) is t
