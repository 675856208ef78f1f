"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row blocks, the edge all-gather that
follows the user-sharded rebuild, and the X-block all-gather of the row-partitioned propagation.
The per-rank compute is stood in by the numpy oracle (the CUDA kernels need a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from diffmm_b200 import dist as ddist
from oracle import diffmm_oracle as O


def test_row_blocks_cover_and_balance():
    for n, w in [(10, 1), (10, 3), (19445, 8), (5, 8)]:
        b = ddist.row_blocks(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
    ptr = torch.tensor([0, 100, 100, 101, 102, 103, 104, 105, 106])
    b = ddist.row_blocks(8, 2, ptr)
    assert b == [(0, 1), (1, 8)] or b[0][1] <= 2            # the heavy first row gets its own block


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, U, I, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                       # same data on every rank
        scores = rng.standard_normal((U, I)).astype(np.float32)
        deg = rng.integers(0, 9, U)
        indptr = np.zeros(U + 1, dtype=np.int64)
        np.cumsum(deg, out=indptr[1:])
        r0, r1 = ddist.shard_rows(U, world, rank)
        items = torch.full((int(indptr[-1]),), -1, dtype=torch.int32)
        local = O.topk_edges(scores[r0:r1], deg[r0:r1])      # stand-in for denoise chain + dmm_topk_edges
        items[indptr[r0]:indptr[r1]] = torch.from_numpy(np.concatenate(local).astype(np.int32))
        full = ddist.allgather_edges(items, torch.from_numpy(indptr), U)
        want = np.concatenate(O.topk_edges(scores, deg)).astype(np.int32)
        ok_edges = bool((full.numpy() == want).all())

        # row-partitioned propagation: local SpMM on the local row block after an all-gather of X blocks
        users = np.repeat(np.arange(U), deg)
        adj = O.normalized_adj_csr(users, want, U, I)
        N = U + I
        x = rng.standard_normal((N, 8)).astype(np.float32)
        blocks = ddist.row_blocks(N, world)
        a, b = blocks[rank]
        xg = ddist.allgather_rows(torch.from_numpy(x[a:b].copy()), blocks)
        y_local = O.spmm_csr(*adj, xg.numpy())[a:b]
        yg = ddist.allgather_rows(torch.from_numpy(y_local.copy()), blocks)
        ok_spmm = bool(np.allclose(yg.numpy(), O.spmm_csr(*adj, x), rtol=1e-6, atol=1e-6))
        ret[rank] = (ok_edges, ok_spmm)
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_rebuild_and_propagation_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 101, 37, ret), nprocs=world, join=True)
    assert dict(ret) == {r: (True, True) for r in range(world)}
