"""CPU, gloo, world_size 2: the row-partition + all-gather plumbing of the multi-GPU path.
Each rank multiplies its row block with the ORACLE (there is no GPU here) -- what is under test is the host logic:
partition bounds, CSR slicing, padded all-gather, equal-flops split for mxm."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from oracle import bigref as R


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from graphblas_b200 import distributed as D

    r, c, n = H.rmat_edges(9, seed=3)
    rng = np.random.default_rng(0)
    w = rng.integers(1, 50, r.size).astype(np.int64)
    A = R.BigMat.from_coo(r, c, w, n, n)
    # ---- iterative min_plus mxv with the vector all-gathered every iteration
    bounds, per = D.row_blocks_equal(n, world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    p, ci, vv = D.slice_csr(A.indptr, A.indices, A.values, r0, r1)
    Ag = R.BigMat(p, ci, vv, r1 - r0, n)
    x = rng.integers(0, 100, n).astype(np.int64)
    xfull = x.copy()
    for it in range(3):
        y_local = R.mxv_T("min_plus", Ag, R.BigVec(x, np.ones(n, np.uint8)))
        vals = np.where(y_local.present.astype(bool), y_local.vals, np.iinfo(np.int64).max)
        x = D.all_gather_padded(torch.from_numpy(vals), per, n).numpy().copy()
        yf = R.mxv_T("min_plus", A, R.BigVec(xfull, np.ones(n, np.uint8)))
        xfull = np.where(yf.present.astype(bool), yf.vals, np.iinfo(np.int64).max)
        assert np.array_equal(x, xfull), (rank, it)
    # ---- mxm: rows split by equal flops, B replicated; concatenated blocks == full product
    deg = np.diff(A.indptr)
    rowflops = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowflops[1:], np.repeat(np.arange(n), deg), deg[A.indices])
    fb = D.row_blocks_by_prefix(np.cumsum(rowflops), world)
    assert fb[0] == 0 and fb[-1] == n and all(a <= b for a, b in zip(fb, fb[1:]))
    f0, f1 = fb[rank], fb[rank + 1]
    p, ci, vv = D.slice_csr(A.indptr, A.indices, A.values, f0, f1)
    Cg = R.mxm_T("plus_times", R.BigMat(p, ci, vv, f1 - f0, n), A)
    Cfull = R.mxm_T("plus_times", A, A)
    k0, k1 = Cfull.indptr[f0], Cfull.indptr[f1]
    assert np.array_equal(Cg.indices, Cfull.indices[k0:k1]) and np.array_equal(Cg.values, Cfull.values[k0:k1])
    nn = D.reduce_scalar(float(Cg.nvals), "sum")
    assert nn == Cfull.nvals
    share = rowflops[f0 + 1:f1 + 1].sum() / rowflops.sum()
    assert 0.2 < share < 0.8, share     # equal-flops split is balanced where an equal-row split would not be
    ret[rank] = 1
    dist.destroy_process_group()


def test_row_partition_gloo_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, ret)) for rk in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {0: 1, 1: 1}


def test_partition_bounds():
    from graphblas_b200 import distributed as D

    b, per = D.row_blocks_equal(10, 4)
    assert b == [0, 3, 6, 9, 10] and per == 3
    pref = np.array([0, 100, 100, 100, 101, 102, 200])
    assert D.row_blocks_by_prefix(pref, 2) == [0, 1, 6]
