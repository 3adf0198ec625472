"""1-D row partition of the hot path across GPUs: one process per GPU,
``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU tests) only for plumbing.

  * mxv-type iterations: rank g owns a contiguous row block of A and the matching slice of every output vector;
    the dense input vector is re-assembled each iteration with ONE all-gather of equal (padded) slices.
  * mxm: C_g = A_g (+).(x) B is independent per rank once B is replicated (shipped once by broadcast); rows are
    split by equal *flops* prefix, not equal row count (R-MAT rows are heavily skewed).

The reference has no distributed layer at all (single process, OpenMP inside its C library); nothing here
mirrors a reference file.  Pure index arithmetic + collectives, so it is testable on CPU with gloo.
"""
import numpy as np


def row_blocks_equal(n, world):
    """Equal-count contiguous blocks (padded all-gather needs equal slice lengths): bounds[g]..bounds[g+1]."""
    per = (n + world - 1) // world
    return [min(n, g * per) for g in range(world + 1)], per


def row_blocks_by_prefix(prefix, world):
    """Contiguous blocks with (nearly) equal weight; `prefix` is the inclusive-exclusive cumulative weight, len n+1."""
    prefix = np.asarray(prefix)
    n = prefix.shape[0] - 1
    total = int(prefix[-1])
    targets = [(total * g) // world for g in range(world + 1)]
    bounds = np.searchsorted(prefix, targets, side="left")
    bounds[0], bounds[-1] = 0, n
    return [int(b) for b in np.maximum.accumulate(bounds)]


def mxm_row_costs(rowflops, *, split_above=10440, part_keys=16320, reread=1.0):
    """Per-row cost estimate of the hash SpGEMM for the row partition of A: the flop bound, except that a row too big for one
    shared-memory table is hashed by ceil(1.125 flops / part_keys) CTAs which ALL stream the row's products (csrc/spgemm.cu,
    split rows) -- its cost grows by `reread` per extra part.  Balancing this instead of the plain flops keeps the rank that
    owns the heavy low-index rows of an R-MAT matrix from finishing last: slowest / mean block time of the 8-way split at scale 22 is
    1.13 with plain flops, 1.04 at reread = 0.5, 1.03 at 1.0, 1.08 at 2.0 (scripts/part_balance.py, one B200).
    `rowflops`: torch tensor or numpy array; returns the same kind (float64)."""
    f = rowflops.double() if hasattr(rowflops, "double") else np.asarray(rowflops, dtype=np.float64)
    parts = (f * 1.125 / part_keys).ceil() if hasattr(f, "ceil") else np.ceil(f * 1.125 / part_keys)
    extra = (parts - 1).clamp(min=0) if hasattr(parts, "clamp") else np.maximum(parts - 1, 0)
    extra = extra * (f > split_above)   # rows that fit one table are not split at all
    return f * (1.0 + reread * extra)


def slice_csr(indptr, cols, vals, r0, r1):
    """Rows [r0, r1) of a CSR given as array-likes supporting slicing (numpy arrays or torch tensors)."""
    k0, k1 = int(indptr[r0]), int(indptr[r1])
    return indptr[r0:r1 + 1] - k0, cols[k0:k1], vals[k0:k1]


def all_gather_padded(local, rows_per, n, group=None):
    """Concatenate every rank's slice (len <= rows_per) into the full length-n vector; torch tensors, any backend."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:n]
    pad = torch.zeros(rows_per, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    out = torch.empty(rows_per * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n]


def reduce_scalar(x, op="max", device=None, group=None):
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op], group=group)
    return float(t[0])


# ------------------------------------------------------------------ iterative mxv / vxm on a row partition (BFS pull / SSSP / PageRank)
def row_blocks_by_nnz(indptr, world):
    """Equal-nnz prefix split of a CSR (R-MAT rows are heavily skewed, equal-row blocks put most of the
    entries on one GPU).  `indptr` may be a numpy array or a torch tensor (any device); returns world + 1 row bounds."""
    try:
        import torch

        if isinstance(indptr, torch.Tensor):
            n = indptr.numel() - 1
            total = int(indptr[-1])
            targets = torch.tensor([(total * g) // world for g in range(world + 1)], dtype=indptr.dtype, device=indptr.device)
            b = torch.searchsorted(indptr, targets, right=False).tolist()
            b[0], b[-1] = 0, n
            return [int(x) for x in np.maximum.accumulate(np.asarray(b))]
    except ImportError:   # pragma: no cover
        pass
    return row_blocks_by_prefix(indptr, world)


def all_gather_uneven(views, local, group=None):
    """Every rank's `local` (lengths may differ) into `views[g]` on every rank -- the ncclAllGatherV of SURVEY.md section 8e.
    `views` are slices of ONE replicated buffer, so the gather lands in place.  Equal lengths: one all_gather_into_tensor on the
    flat buffer; unequal: torch's all_gather (grouped ncclBroadcasts under NCCL), else one broadcast per rank (gloo)."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        if views[0].data_ptr() != local.data_ptr():
            views[0].copy_(local)
        return
    sizes = [v.numel() for v in views]
    if len(set(sizes)) == 1 and all(views[g].data_ptr() == views[0].data_ptr() + g * sizes[0] * views[0].element_size() for g in range(world)):
        import torch

        flat = torch.as_strided(views[0], (sizes[0] * world,), (1,))
        dist.all_gather_into_tensor(flat, local, group=group)
        return
    if dist.get_backend(group) == "nccl":
        dist.all_gather(list(views), local, group=group)
        return
    views[rank].copy_(local)
    for g in range(world):
        dist.broadcast(views[g], src=dist.get_global_rank(group, g) if group is not None else g, group=group)


class GatheredVector:
    """The replicated input vector of a row-partitioned pull multiply: a full-length library Vector on every rank whose device
    arrays are refilled, in place, from the ranks' output slices (values and -- for sparse vectors -- presence bytes)."""

    def __init__(self, gb, dtype, n, bounds, *, sparse=True):
        import torch

        self.gb, self.n, self.bounds, self.sparse = gb, n, bounds, sparse
        tdt = gb.cuda._np_to_torch_dtype(gb.dtypes.lookup_dtype(dtype).np_type)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.vector = gb.cuda.vector_from_torch(torch.zeros(n, dtype=tdt, device=dev),
                                                torch.zeros(n, dtype=torch.uint8, device=dev) if sparse else None)
        self.vals, self.present = gb.cuda.vector_as_torch(self.vector, sync=False)
        world = len(bounds) - 1
        self._vviews = [self.vals[bounds[g]:bounds[g + 1]] for g in range(world)]
        self._pviews = [self.present[bounds[g]:bounds[g + 1]] for g in range(world)]

    def gather(self, local, group=None):
        """local: this rank's slice (a library Vector of length bounds[rank + 1] - bounds[rank])"""
        lv, lp = self.gb.cuda.vector_as_torch(local, sync=False)   # same stream as torch / NCCL: no host synchronisation
        all_gather_uneven(self._vviews, lv, group)
        if self.sparse:
            all_gather_uneven(self._pviews, lp, group)
            self.gb.cuda.vector_touch(self.vector)
        else:
            self.gb.cuda.vector_assume_full(self.vector)
        return self.vector


class PeerExchange:
    """The replicated input vector of a row-partitioned pull multiply, refilled WITHOUT a collective: two buffer sets (values +
    presence bytes) per rank in IPC-exportable memory, every rank maps every other rank's sets once, and the SpMV epilogue of
    iteration k stores each finished output position into set (k + 1) & 1 of ALL ranks over NVLink (GrB_cuda_set_peer_targets).
    The only cross-rank operation per iteration is a tiny all-reduce (which SSSP needs anyway for its convergence flag): after
    it, every writer's kernel has completed, so set (k + 1) & 1 is complete everywhere and set k & 1 may be overwritten."""

    def __init__(self, gb, dtype, n, bounds, rank, group=None, *, dense=False):
        import torch
        import torch.distributed as dist

        self.gb, self.n, self.bounds, self.rank = gb, n, bounds, rank
        self.dense = dense   # every position always holds an entry (PageRank, dense mxv): no presence traffic, no recounts
        self.dtype = gb.dtypes.lookup_dtype(dtype)
        world = len(bounds) - 1
        es = self.dtype.np_type.itemsize
        self._own = [(gb.cuda.peer_alloc(n * es), gb.cuda.peer_alloc(n)) for _ in range(2)]
        gb.cuda.sync()
        handles = [(gb.cuda.ipc_handle(v), gb.cuda.ipc_handle(p)) for v, p in self._own]
        allh = [None] * world
        if world > 1:
            dist.all_gather_object(allh, handles, group=group)
        else:
            allh[0] = handles
        self._opened = []
        self.targets = []   # per buffer set: ([vals pointer of every rank], [presence pointer of every rank])
        for k in range(2):
            vs, ps = [], []
            for g in range(world):
                if g == rank:
                    v, p = self._own[k]
                else:
                    v, p = gb.cuda.ipc_open(allh[g][k][0]), gb.cuda.ipc_open(allh[g][k][1])
                    self._opened += [v, p]
                vs.append(v)
                ps.append(p)
            self.targets.append((vs, ps))
        self.vectors = [gb.cuda.vector_wrap(self.dtype, n, v, p) for v, p in self._own]
        self.step = 0
        self._flag = torch.zeros(1, device=torch.device("cuda", torch.cuda.current_device()))
        self._group = group

    @property
    def current(self):
        return self.vectors[self.step & 1]

    def writing(self, scale_ptr=None, *, with_presence=True):
        """context manager: multiplies inside it also store their result slice into the NEXT buffer set of every rank"""
        vs, ps = self.targets[(self.step + 1) & 1]
        return self.gb.cuda.peer_targets(vs, ps if (with_presence and not self.dense) else None, self.bounds[self.rank], scale_ptr)

    def fill_current(self, local, group=None):
        """first fill of the current set from the ranks' slices (one NCCL all-gather, before the loop)"""
        import torch

        es = self.dtype.np_type.itemsize
        v, p = self._own[self.step & 1]
        tdt = self.gb.cuda._np_to_torch_dtype(self.dtype.np_type)
        vals = torch.as_tensor(self.gb.cuda._CudaArray(v, (self.n,), self.dtype.np_type.str if self.dtype.name != "BOOL" else "|u1"), device="cuda")
        pres = torch.as_tensor(self.gb.cuda._CudaArray(p, (self.n,), "|u1"), device="cuda")
        lv, lp = self.gb.cuda.vector_as_torch(local, sync=False)
        world = len(self.bounds) - 1
        all_gather_uneven([vals[self.bounds[g]:self.bounds[g + 1]] for g in range(world)], lv.view(vals.dtype), group)
        all_gather_uneven([pres[self.bounds[g]:self.bounds[g + 1]] for g in range(world)], lp, group)
        self._refilled()

    def _refilled(self):
        (self.gb.cuda.vector_assume_full if self.dense else self.gb.cuda.vector_touch)(self.current)

    def advance(self, flag=0.0):
        """cross-rank ordering point of the iteration (max all-reduce of `flag`, returned); then the next set becomes current"""
        import torch.distributed as dist

        self._flag.fill_(float(flag))
        if dist.is_initialized() and dist.get_world_size(self._group) > 1:
            dist.all_reduce(self._flag, op=dist.ReduceOp.MAX, group=self._group)
        self.step += 1
        self._refilled()
        return self._flag

    def close(self):
        self.vectors = []
        for ptr in self._opened:
            self.gb.cuda.ipc_close(ptr)
        for v, p in self._own:
            self.gb.cuda.peer_free(v)
            self.gb.cuda.peer_free(p)
        self._opened, self._own = [], []


def transpose_csr_torch(indptr, cols, vals, n):
    """CSR of the transpose (== CSC of the matrix) with torch ops on the device: stable sort of the entries by column."""
    import torch

    nnz = cols.numel()
    rows = torch.repeat_interleave(torch.arange(n, device=cols.device, dtype=torch.int32), (indptr[1:] - indptr[:-1]))
    key = cols.to(torch.int64) * n + rows.to(torch.int64)
    order = torch.argsort(key)
    del key
    tcols = rows[order].contiguous()
    tvals = vals[order].contiguous() if vals is not None else None
    counts = torch.bincount(cols.to(torch.int64), minlength=n)
    tptr = torch.zeros(n + 1, dtype=torch.int64, device=cols.device)
    tptr[1:] = torch.cumsum(counts, 0)
    assert int(tptr[-1]) == nnz
    return tptr, tcols, tvals


def local_block(gb, indptr, cols, vals, ncols, r0, r1):
    """rows [r0, r1) of a device CSR as a library Matrix of shape (r1 - r0) x ncols"""
    k0, k1 = int(indptr[r0]), int(indptr[r1])
    return gb.cuda.matrix_from_device_csr((indptr[r0:r1 + 1] - k0).contiguous(), cols[k0:k1].contiguous(), vals[k0:k1].contiguous(),
                                          r1 - r0, ncols)


def sssp_partitioned(gb, Wt_block, bounds, rank, n, src, *, max_iters=64, group=None, exchange="nccl", px=None):
    """Bellman-Ford sweeps d(min) << d.vxm(W, min_plus) on a row partition of W' (BASELINE config 4): rank g owns rows
    bounds[g]..bounds[g + 1] of W' (= columns of W) and the matching slice of d; the full d is all-gathered every sweep.
    Returns (local slice of d, full gathered d, sweeps).  Bit-identical to the single-GPU loop: min over int64 is exact.
    `px`: a PeerExchange built beforehand (mapping the peers' buffers is a one-off set-up, not part of an iteration)."""
    import torch
    import torch.distributed as dist

    r0, r1 = bounds[rank], bounds[rank + 1]
    d_loc = gb.Vector(gb.dtypes.INT64, r1 - r0)
    if r0 <= src < r1:
        d_loc[src - r0] = 0
    if exchange == "peer":
        # fused: the epilogue of each sweep's multiply stores the new distances into every rank's next input vector; the
        # all-reduce of the convergence flag is the only collective (and the ordering point) of a sweep
        own_px = px is None
        if own_px:
            px = PeerExchange(gb, gb.dtypes.INT64, n, bounds, rank, group)
        try:
            px.fill_current(d_loc, group)
            sweeps = 0
            for _ in range(max_iters):
                old = d_loc.dup()
                with px.writing():
                    d_loc(gb.binary.min) << Wt_block.mxv(px.current, gb.semiring.min_plus)
                sweeps += 1
                changed = px.advance(0.0 if d_loc.isequal(old) else 1.0)
                if float(changed[0]) == 0.0:
                    break
            full = GatheredVector(gb, gb.dtypes.INT64, n, bounds, sparse=True)
            full.gather(d_loc, group)
        finally:
            if own_px:
                px.close()
        return d_loc, full, sweeps
    full = GatheredVector(gb, gb.dtypes.INT64, n, bounds, sparse=True)
    full.gather(d_loc, group)
    sweeps = 0
    for _ in range(max_iters):
        old = d_loc.dup()
        d_loc(gb.binary.min) << Wt_block.mxv(full.vector, gb.semiring.min_plus)
        sweeps += 1
        changed = 0.0 if d_loc.isequal(old) else 1.0
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            t = torch.tensor([changed], device=full.vals.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            changed = float(t[0])
        full.gather(d_loc, group)
        if not changed:
            break
    return d_loc, full, sweeps


def pagerank_partitioned(gb, At_block, outdeg_loc, bounds, rank, n, *, iters=20, damping=0.85, group=None, exchange="nccl", px=None):
    """The notebook recurrence (reference notebooks/Pagerank Demo.ipynb cell 9) on a row partition of A' (BASELINE config 5):
    w = damping * t / d ; r = teleport ; r(plus) << A'.mxv(w, plus_second) ; t = r -- here t, d, r are local slices, w is
    all-gathered (fp64) once per iteration.  Returns the local slice of t."""
    import torch

    r0, r1 = bounds[rank], bounds[rank + 1]
    m = r1 - r0
    dev = torch.device("cuda", torch.cuda.current_device())
    teleport = (1 - damping) / n
    t = gb.cuda.vector_from_torch(torch.full((m,), 1.0 / n, dtype=torch.float64, device=dev))
    if exchange == "peer":
        # fused: r(plus) << A'.mxv(w) stores damping * r / d -- the NEXT w -- into every rank's next input vector from the
        # multiply's epilogue (scale = damping / d per local row); per iteration: one assign, one multiply, one tiny all-reduce
        dv, _ = gb.cuda.vector_as_torch(outdeg_loc, sync=False)
        scale = (damping / dv).contiguous()
        own_px = px is None
        if own_px:
            px = PeerExchange(gb, gb.dtypes.FP64, n, bounds, rank, group, dense=True)
        try:
            w0 = gb.cuda.vector_from_torch((scale / n).contiguous())   # w of the first iteration: damping * (1 / n) / d
            px.fill_current(w0, group)
            for _ in range(iters):
                r = gb.Vector(gb.dtypes.FP64, m)
                r[:] = teleport
                with px.writing(scale.data_ptr()):
                    r(gb.binary.plus) << At_block.mxv(px.current, gb.semiring.plus_second)
                px.advance()
                t = r
            gb.cuda.sync()
        finally:
            if own_px:
                px.close()
        return t
    full = GatheredVector(gb, gb.dtypes.FP64, n, bounds, sparse=False)
    for _ in range(iters):
        w = t.ewise_mult(outdeg_loc, gb.binary.truediv).new()
        w = w.apply(gb.binary.times, right=damping).new()
        full.gather(w, group)
        r = gb.Vector(gb.dtypes.FP64, m)
        r[:] = teleport
        r(gb.binary.plus) << At_block.mxv(full.vector, gb.semiring.plus_second)
        t = r
    return t
