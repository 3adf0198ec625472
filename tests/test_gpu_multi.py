"""GPU, NCCL, 2 ranks (skipped on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
the CUDA path under the row partition of SURVEY.md section 8e must reproduce the 1-GPU results --
  * mxm: the row blocks of the partitioned product, concatenated, equal the 1-GPU product bit-exactly (int64 plus_times);
  * SSSP (BASELINE config 4): min_plus sweeps with the distance vector all-gathered every sweep, bit-exact (int64);
  * PageRank (config 5): plus_second fp64 with the per-iteration all-gather, rel <= 1e-10 (summation order differs per block);
  * the fused SpMV + peer-write exchange equals the NCCL all-gather.
Every rank also runs the 1-GPU formulation on the full matrix as its own check."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import bench
    import graphblas_b200 as gb
    from graphblas_b200 import distributed as D

    gb.init(device=rank)
    torch.cuda.set_stream(torch.cuda.Stream())
    gb.cuda.use_torch_stream()
    dev = torch.device("cuda", rank)
    scale = 13
    ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2B, 42, device=dev)   # same seed on every rank: identical graph
    nnz = c.numel()
    g = torch.Generator(device=dev); g.manual_seed(43)
    w = torch.randint(1, 256, (nnz,), device=dev, generator=g, dtype=torch.int64)

    # ---- mxm: A split by equal flops, B replicated
    B = gb.cuda.matrix_from_device_csr(ip, c, w, n, n)
    deg = ip[1:] - ip[:-1]
    rowflops = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowflops[1:].index_add_(0, torch.repeat_interleave(torch.arange(n, device=dev), deg), deg[c.long()])
    fb = D.row_blocks_by_prefix(torch.cumsum(rowflops, 0).cpu().numpy(), world)
    A_blk = D.local_block(gb, ip, c, w, n, fb[rank], fb[rank + 1])
    C_blk = A_blk.mxm(B, gb.semiring.plus_times).new()
    C_full = B.mxm(B, gb.semiring.plus_times).new()
    gb.cuda.matrix_sort(C_blk); gb.cuda.matrix_sort(C_full)
    bp, bj, bx = gb.cuda.matrix_as_torch(C_blk)
    fp_, fj, fx = gb.cuda.matrix_as_torch(C_full)
    k0, k1 = int(fp_[fb[rank]]), int(fp_[fb[rank + 1]])
    assert torch.equal(bp, fp_[fb[rank]:fb[rank + 1] + 1] - k0), "mxm row pointers"
    assert torch.equal(bj, fj[k0:k1]) and torch.equal(bx, fx[k0:k1]), "mxm block differs from the 1-GPU product"
    tot = torch.tensor([C_blk.nvals], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    assert int(tot[0]) == C_full.nvals

    # ---- SSSP: rows of W' partitioned by equal nnz
    src = int(torch.nonzero(deg > 0)[0])
    W = gb.cuda.matrix_from_device_csr(ip, c, w, n, n)
    d1 = gb.Vector.from_coo([src], [0], size=n, dtype=gb.dtypes.INT64)
    sweeps1 = 0
    for _ in range(64):
        old = d1.dup()
        d1(gb.binary.min) << d1.vxm(W, gb.semiring.min_plus)
        sweeps1 += 1
        if d1.isequal(old):
            break
    tp, tc, tv = D.transpose_csr_torch(ip, c, w, n)
    nb = D.row_blocks_by_nnz(tp, world)
    Wt_blk = D.local_block(gb, tp, tc, tv, n, nb[rank], nb[rank + 1])
    d_loc, full, sweeps = D.sssp_partitioned(gb, Wt_blk, nb, rank, n, src)
    v1, p1 = gb.cuda.vector_as_torch(d1)
    assert sweeps == sweeps1, (sweeps, sweeps1)
    assert torch.equal(full.present, p1) and torch.equal(full.vals[p1.bool()], v1[p1.bool()]), "SSSP differs from the 1-GPU loop"

    # ---- PageRank on the row partition of A'
    ones = torch.ones(nnz, dtype=torch.float64, device=dev)
    Af = gb.cuda.matrix_from_device_csr(ip, c, ones, n, n)
    dvec_t = torch.clamp(deg, min=1).to(torch.float64)
    dvec = gb.cuda.vector_from_torch(dvec_t)
    damping, teleport = 0.85, 0.15 / n
    t1 = gb.cuda.vector_from_torch(torch.full((n,), 1.0 / n, dtype=torch.float64, device=dev))
    for _ in range(10):
        wv = t1.ewise_mult(dvec, gb.binary.truediv).new()
        wv = wv.apply(gb.binary.times, right=damping).new()
        r = gb.Vector(gb.dtypes.FP64, n)
        r[:] = teleport
        r(gb.binary.plus) << Af.T.mxv(wv, gb.semiring.plus_second)
        t1 = r
    atp, atc, atv = D.transpose_csr_torch(ip, c, ones, n)
    pb = D.row_blocks_by_nnz(atp, world)
    At_blk = D.local_block(gb, atp, atc, atv, n, pb[rank], pb[rank + 1])
    dloc = gb.cuda.vector_from_torch(dvec_t[pb[rank]:pb[rank + 1]].contiguous())
    for exchange in ("nccl", "peer"):
        t_loc = D.pagerank_partitioned(gb, At_blk, dloc, pb, rank, n, iters=10, exchange=exchange)
        tl, _ = gb.cuda.vector_as_torch(t_loc)
        tf, _ = gb.cuda.vector_as_torch(t1)
        ref = tf[pb[rank]:pb[rank + 1]]
        rel = float(((tl - ref).abs() / ref.abs()).max())
        assert rel <= 1e-10, (exchange, rel)
    ret[rank] = 1
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_path_matches_single_gpu():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    port = 29700 + (os.getpid() % 1000)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, ret)) for rk in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {0: 1, 1: 1}
