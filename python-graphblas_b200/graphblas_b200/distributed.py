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
