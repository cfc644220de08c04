"""N > 1 host logic on CPU: two gloo ranks shard a sample by chunk, each computes its own chunk
range (with the CPU oracle standing in for the GPU library), results are assembled by row range and
the column sums are combined in rank order.  Sharded == unsharded, bitwise."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as O
    from scasa_b200.shard import shard_chunks, local_problem, combine_colsum
    from scasa_b200.xmatrix import XMatrix
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    d = np.load(os.path.join(tmp, "problem.npz"))
    xm = XMatrix.load_npz(os.path.join(tmp, "xm.npz"))
    chunks, y = d["chunks"], d["y"]
    rng_ = shard_chunks(chunks, world)[rank]
    cell0, cell1, loc, _, _ = local_problem(chunks, np.arange(len(y) + 1), rng_)
    iso, iters = O.estim_chunks(xm, y[cell0:cell1], loc, None)
    colsum = combine_colsum(iso.sum(axis=0))
    np.savez(os.path.join(tmp, f"out{rank}.npz"), iso=iso, iters=iters, cell0=cell0, colsum=colsum)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path, oracle):
    from tests.test_oracle import _xm_from_blocks
    rng = np.random.default_rng(11)
    blocks = []
    for q, p in [(3, 2), (2, 1), (5, 3), (1, 1), (4, 2), (1, 1)]:
        X = rng.random((q, p)) + 0.05
        blocks.append(X / X.sum(0) if q > 1 else np.ones((1, 1)))
    xm = _xm_from_blocks(blocks)
    n = 430                                                       # 4 chunks of 107-108 cells
    y = np.concatenate([rng.poisson(rng.gamma(0.5, 8, (n, b.shape[1])) @ b.T) for b in blocks], axis=1).astype(float)
    chunks = oracle.chunk_split(n, 4)
    assert len(chunks) == 5
    xm.save_npz(str(tmp_path / "xm.npz"))
    np.savez(tmp_path / "problem.npz", chunks=chunks, y=y)
    import socket
    with socket.socket() as sk:                                   # a port the kernel says is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    whole, whole_it = oracle.estim_chunks(xm, y, chunks, None)
    parts = [np.load(tmp_path / f"out{r}.npz") for r in range(2)]
    assert int(parts[0]["cell0"]) == 0 and int(parts[1]["cell0"]) == chunks[2]
    assert np.array_equal(np.concatenate([p["iso"] for p in parts]), whole)          # bitwise
    assert np.array_equal(np.concatenate([p["iters"] for p in parts]), whole_it)
    want = parts[0]["iso"].sum(axis=0) + parts[1]["iso"].sum(axis=0)                 # rank order
    assert np.array_equal(parts[0]["colsum"], want) and np.array_equal(parts[1]["colsum"], want)
