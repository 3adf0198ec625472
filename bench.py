#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's metric: mxm nnz-out/s (+ mxv GB/s) on synthetic R-MAT CSR.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

A "step" is one A.mxm(A, plus_times) fp32 on the R-MAT matrix of BASELINE.json configs[1]
("R-MAT scale-22 avg-deg-16 A.mxm(A) plus_times fp32 on 1 B200"), in the feasible variant SURVEY.md
section 8(d) names "2a": R-MAT (a,b,c,d)=(0.45,0.15,0.15,0.25), edge factor 16, seed 42, deduplicated
(Graph500 skew at scale 22 would emit ~6e10 entries = 0.5 TB, which no single GPU holds).
`value` = nnz(C) / device time with A resident in HBM (symbolic + numeric; the result is left "jumbled",
as the reference's C library also leaves it -- sort time is reported separately).
`e2e`   = the same through the C-ABI with HOST buffers: H2D of A's CSR + mxm + D2H of C's CSR per step
          (`e2e_reference_interface`: through GrB_Matrix_import_FP32 / export_FP32 with uint64 pageable numpy arrays).
`roofline` = SURVEY.md 8(d): B_min / t_step / measured HBM peak for the whole mxm step.
`mxv`   = plus_times fp32 A.mxv(x) on the Graph500-skew matrix "2b" (the north_star's >=50%-of-HBM target).
`workloads` = BFS / SSSP / PageRank loops of BASELINE configs 3-5 (N > 1: SSSP and PageRank on the 1-D row partition with a
          per-iteration exchange, NCCL all-gather and the fused SpMV + NVLink peer-store variant); `scale25` = the scale-25 runs.
`cpu_baseline` = the in-repo OpenMP port on all host cores (mxm sample; mxv / min_plus vxm / plus_second mxv on the whole matrix).
One process per GPU (torchrun); every rank generates the inputs from the seed itself; N>1 row-partitions A by equal flops
(B = A replicated) and the vector workloads by equal nnz; time = max over ranks.
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

RMAT_2A = (0.45, 0.15, 0.15)   # mild skew: nnz(C) fits
RMAT_2B = (0.57, 0.19, 0.19)   # Graph500


def peaks():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text()), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------ synthetic input (on the GPU, torch as a carrier)
def rmat_csr_torch(scale, params, seed, edge_factor=16, device="cuda"):
    import torch

    a, b, c = params
    n = 1 << scale
    m = edge_factor * n
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rows = torch.zeros(m, dtype=torch.int64, device=device)
    cols = torch.zeros(m, dtype=torch.int64, device=device)
    ab, abc = a + b, a + b + c
    for bit in range(scale):
        r = torch.rand(m, device=device, generator=g)
        rows += (r >= ab).to(torch.int64) << (scale - 1 - bit)
        cols += (((r >= a) & (r < ab)) | (r >= abc)).to(torch.int64) << (scale - 1 - bit)
        del r
    key = torch.unique(rows * n + cols)   # sorted + deduplicated
    del rows, cols
    rows = key // n
    cols = (key % n).to(torch.int32)
    del key
    counts = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    return indptr, cols, n


def values_torch(nnz, seed, dtype, device="cuda"):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.rand(nnz, device=device, generator=g, dtype=torch.float32).to(dtype)


# ------------------------------------------------------------------ clocks sampling during the timed region
class ClockSampler:
    """SM clock + throttle reasons sampled every 100 ms DURING the timed region.  NVML is queried in-process (pynvml): spawning
    `nvidia-smi` every 200 ms initialises NVML and enumerates all GPUs of the box each time, which on a busy 8-GPU host takes
    hundreds of ms under the driver lock and stalled this process's own CUDA calls (one run measured 168 ms per step for 56 ms of
    kernels).  nvidia-smi remains the fallback when pynvml is unavailable."""

    def __init__(self, gpu_index=0):
        self.samples, self.reasons = [], set()
        self._stop = threading.Event()
        self.gpu = gpu_index
        self.sm_max = None
        self.source = "nvml"
        self._h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = gpu_index
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                ids = [int(t) for t in vis.split(",")]
                phys = ids[gpu_index] if gpu_index < len(ids) else gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None
            self.source = "nvidia-smi"
        self._t = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        nv = self._nv
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        for nm, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                        ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(nm)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0]))
        self.sm_max = float(out[1])
        for nm, v in zip(names, out[2:]):
            if v.strip().lower().startswith("active"):
                self.reasons.add(nm)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._h is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.1 if self._h is not None else 0.5)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}


# ------------------------------------------------------------------ CPU legs (reference arm and cpu_baseline)
L2_NOTE = "no flush: every step streams inputs and outputs far larger than the 126 MB L2 (scale 22: 0.5 GB in, 20 GB out)"
WORKLOAD = ("R-MAT scale-{scale} (0.45,0.15,0.15,0.25) ef16 seed42 A.mxm(A) plus_times fp32 [BASELINE configs[1], variant 2a]; "
            "CPU arm: random row sample of A times the full A per step, rate in nnz-out/s")


def _cpu_threads():
    """All host cores, set explicitly: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which silently made the
    round-1 reference arm single-threaded at N > 1."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return max(1, n)


def cpu_mxm_sample(indptr, indices, values, n, budget_s=15.0, seed=0, rows_hint=None):
    """Times the CPU baseline (oracle/grb_oracle.c `oracle_mxm_baseline_f32`: OpenMP, one numeric pass, unsorted rows, 32-bit
    indices, hash accumulator for short rows / dense Gustavson workspace for long ones -- the way SuiteSparse's saxpy3 organises
    it) on a random row sample sized for ~budget_s of CPU work.  The timed region covers the per-row flop bound, its prefix and
    the numeric pass, not the 32-bit index conversion (input preparation).  Returns (nnz_out_per_s, description, threads,
    rows_used, seconds).  `rows_hint` skips the sizing search."""
    from oracle import bigref as R

    R.set_num_threads(_cpu_threads())
    A = R.BigMat(indptr, indices, values, n, n)
    Aj32 = A.indices.astype(np.int32)
    rng = np.random.default_rng(seed)
    deg = np.diff(A.indptr)

    def sub(rows):
        rows = np.sort(rows)
        lens = deg[rows]
        ptr = np.zeros(rows.size + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        idx = np.repeat(A.indptr[rows] - ptr[:-1], lens) + np.arange(ptr[-1])
        return R.BigMat(ptr, A.indices[idx], A.values[idx], rows.size, n)

    # grow the sample until it costs a meaningful fraction of the budget (per-call set-up of the per-thread workspaces --
    # O(threads x ncols) -- would otherwise dominate a small sample and understate the CPU)
    cap = n if n < (1 << 21) else n // 2   # bounded sample: at most half the rows of a big matrix (host memory of the staging arrays)
    m = min(cap, rows_hint if rows_hint else 1 << 15)
    while True:
        rows = rng.choice(n, size=m, replace=False)
        As = sub(rows)
        nvals, dt, out = R.mxm_baseline_f32(As, A, indices32=(As.indices.astype(np.int32), Aj32))
        del out
        if rows_hint or dt >= budget_s / 3 or m >= cap:
            break
        m = int(min(cap, max(m * 2, m * (budget_s / max(dt, 1e-3)) * 0.7)))
    return nvals / dt, f"{m} random rows of A (of {n}) times full A, {nvals} output entries in {dt:.2f} s", R.num_threads(), m, dt


def cpu_mxv_legs(hp, hc, n, scale, reps=5):
    """CPU baselines of the vector multiplies on the Graph500-skew matrix (north_star: min_plus vxm >= 10x CPU): the oracle's
    row-parallel OpenMP pull (oracle/grb_oracle.c oracle_mxv, 64-bit indices) on all host cores, whole matrix, median of `reps`."""
    from oracle import bigref as R

    R.set_num_threads(_cpu_threads())
    rng = np.random.default_rng(7)
    nnz = hc.size
    out = {}
    for name, sr, dt in (("mxv_plus_times_fp32", "plus_times", np.float32), ("vxm_min_plus_int64", "min_plus", np.int64),
                         ("mxv_plus_second_fp64", "plus_second", np.float64)):
        vals = rng.integers(1, 256, nnz).astype(dt) if dt == np.int64 else rng.random(nnz).astype(dt)
        A = R.BigMat(hp, hc, vals, n, n)
        x = R.BigVec((rng.integers(0, 1000, n) if dt == np.int64 else rng.random(n)).astype(dt), np.ones(n, np.uint8))
        R.mxv_T(sr, A, x)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            R.mxv_T(sr, A, x)
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts))
        out[name] = {"ms": t * 1e3, "nnz_per_s": nnz / t, "cores": R.num_threads(), "kind": "port",
                     "sample": f"whole R-MAT scale-{scale} Graph500-skew matrix ({nnz} entries), dense input vector, median of {reps}"}
        del A
    return out


def rmat_host_csr(scale, params, seed):
    """The bench matrix on the host for the CPU legs (torch CPU ops: the numpy generator needs minutes at scale 22)."""
    import torch

    torch.set_num_threads(_cpu_threads())
    indptr, cols, n = rmat_csr_torch(scale, params, seed, device="cpu")
    vals = values_torch(cols.numel(), seed + 1, torch.float32, device="cpu")
    return indptr.numpy(), cols.numpy().astype(np.int64), vals.numpy(), n


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  SuiteSparse:GraphBLAS is not installable here
    (no network, not vendored), so this arm times the in-repo OpenMP port on ALL host cores, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.pop("OMP_NUM_THREADS", None)
    scale = args.scale
    indptr, c, vals, n = rmat_host_csr(scale, RMAT_2A, 42)
    rates, secs, hint = [], [], None
    for s in range(args.warmup + args.steps):
        # the first (warm-up) step sizes the sample for the per-step budget; the timed steps reuse that size with fresh rows
        rate, desc, threads, hint, dt = cpu_mxm_sample(indptr, c, vals, n, budget_s=max(2.0, 60.0 / max(1, args.steps + args.warmup)),
                                                       seed=s, rows_hint=hint)
        if s >= args.warmup:
            rates.append(rate)
            secs.append(dt)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "mxm nnz-out/s (R-MAT plus_times fp32)", "value": value, "unit": "nnz-out/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(scale=scale), "l2": L2_NOTE},
        "cpu_baseline": {"value": value, "unit": "nnz-out/s", "cores": threads, "kind": "port", "sample": desc,
                         "note": "SuiteSparse unavailable -- baseline is the in-repo OpenMP SpGEMM (one pass, unsorted rows, hash / dense Gustavson accumulators)"},
        "e2e": {"value": value, "unit": "nnz-out/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ the iterative workloads (BASELINE configs 3-5)
def _time_cuda(torch, fn):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    out = fn()
    ev1.record()
    torch.cuda.synchronize()
    return out, ev0.elapsed_time(ev1)


def run_workloads(gb, torch, dev, scale):
    """Level BFS (any_pair, complemented structural mask + replace), SSSP (min_plus, min accum, int64) and PageRank
    (plus_second on A.T, fp64) on the Graph500-skew R-MAT, written exactly like the reference notebooks.
    Device-resident; times are CUDA-event times of the whole loop, including the O(n) vector ops around the multiply."""
    ip, c, n = rmat_csr_torch(scale, RMAT_2B, 42, device=dev)
    nnz = c.numel()
    deg = ip[1:] - ip[:-1]
    src = int(torch.nonzero(deg > 0)[0])
    out = {"graph": f"R-MAT scale-{scale} (0.57,0.19,0.19,0.05) ef16 seed42", "n": n, "nnz": nnz}

    # ---- BFS
    A = gb.cuda.matrix_from_device_csr(ip, c, torch.ones(nnz, dtype=torch.bool, device=dev), n, n)
    gb.cuda.matrix_build_transpose(A)

    def bfs():
        q = gb.Vector.from_coo([src], [True], size=n)
        v = gb.Vector(gb.dtypes.INT64, n)
        levels = 0
        for level in range(1, n):
            v(mask=q.V)[:] = level
            q(~v.S, replace=True) << q.vxm(A, gb.semiring.any_pair)
            levels = level
            if q.nvals == 0:
                break
        return v, levels

    bfs()
    (v, levels), ms = _time_cuda(torch, bfs)
    vv, vp = gb.cuda.vector_as_torch(v, sync=False)
    edges = int(deg[vp.bool()].sum())
    out["bfs"] = {"ms": ms, "levels": levels, "reached": int(vp.sum()), "edges_traversed": edges, "GTEPS": edges / ms / 1e6}
    del A

    # ---- SSSP (Bellman-Ford sweeps to a fixed point, capped)
    g = torch.Generator(device=dev); g.manual_seed(43)
    w = torch.randint(1, 256, (nnz,), device=dev, generator=g, dtype=torch.int64)
    W = gb.cuda.matrix_from_device_csr(ip, c, w, n, n)
    gb.cuda.matrix_build_transpose(W)

    def sssp():
        d = gb.Vector.from_coo([src], [0], size=n, dtype=gb.dtypes.INT64)
        its = 0
        for it in range(64):
            old = d.dup()
            d(gb.binary.min) << d.vxm(W, gb.semiring.min_plus)
            its += 1
            if d.isequal(old):
                break
        return d, its

    sssp()
    (d, its), ms = _time_cuda(torch, sssp)
    out["sssp"] = {"ms": ms, "iterations": its, "ms_per_iteration": ms / its, "reached": d.nvals,
                   "mxv_algorithmic_GB_per_iteration": (nnz * 12 + (n + 1) * 8 + n * 8 * 3) / 1e9}
    del W

    # ---- PageRank: 20 iterations of the notebook recurrence, fp64
    Af = gb.cuda.matrix_from_device_csr(ip, c, torch.ones(nnz, dtype=torch.float64, device=dev), n, n)
    gb.cuda.matrix_build_transpose(Af)
    dvec = gb.cuda.vector_from_torch(torch.clamp(deg, min=1).to(torch.float64))
    damping, teleport = 0.85, (1 - 0.85) / n

    def pagerank(iters):
        t = gb.cuda.vector_from_torch(torch.full((n,), 1.0 / n, dtype=torch.float64, device=dev))
        for _ in range(iters):
            wv = t.ewise_mult(dvec, gb.binary.truediv).new()
            wv = wv.apply(gb.binary.times, right=damping).new()
            r = gb.Vector(gb.dtypes.FP64, n)
            r[:] = teleport
            r(gb.binary.plus) << Af.T.mxv(wv, gb.semiring.plus_second)
            t = r
        return t

    pagerank(10)   # warm-up covers the library's one-off kernel-selection trial for this CSR (6-8 multiplies, format builds)
    t, ms = _time_cuda(torch, lambda: pagerank(20))
    out["pagerank"] = {"ms": ms, "iterations": 20, "ms_per_iteration": ms / 20, "sum": t.reduce(gb.monoid.plus).new().value,
                       "mxv_algorithmic_GB_per_iteration": (nnz * 4 + (n + 1) * 8 + n * 8 + n * 9 * 2) / 1e9}
    return out


def run_workloads_partitioned(gb, torch, dist, dev, scale, rank, world, max_over_ranks, which=("sssp", "pagerank")):
    """Configs 4 and 5 on the 1-D row partition (graphblas_b200/distributed.py): every rank generates the graph itself from the
    seed (no rank-0 materialisation, no input traffic), keeps its equal-nnz row block of the transposed CSR, and all-gathers
    the frontier / rank vector once per iteration.  Times are max over ranks of the CUDA-event time of the whole loop."""
    from graphblas_b200 import distributed as D

    ip, c, n = rmat_csr_torch(scale, RMAT_2B, 42 if scale <= 22 else 44, device=dev)
    nnz = c.numel()
    deg = ip[1:] - ip[:-1]
    src = int(torch.nonzero(deg > 0)[0])
    out = {"graph": f"R-MAT scale-{scale} (0.57,0.19,0.19,0.05) ef16", "n": n, "nnz": nnz, "partition": f"equal-nnz row blocks of the transposed CSR x{world}",
           "exchange": "one all-gather of the vector slice per iteration (NCCL; values + presence bytes for sparse vectors)"}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if "sssp" in which:
        g = torch.Generator(device=dev); g.manual_seed(43)
        w = torch.randint(1, 256, (nnz,), device=dev, generator=g, dtype=torch.int64)
        tp, tc, tv = D.transpose_csr_torch(ip, c, w, n)
        del w
        nb = D.row_blocks_by_nnz(tp, world)
        Wt = D.local_block(gb, tp, tc, tv, n, nb[rank], nb[rank + 1])
        local_nnz = int(tp[nb[rank + 1]] - tp[nb[rank]])
        del tp, tc, tv
        res = {}
        for exchange in (("nccl", "peer") if world > 1 else ("nccl",)):
            px = None
            try:
                px = D.PeerExchange(gb, gb.dtypes.INT64, n, nb, rank) if exchange == "peer" else None
                D.sssp_partitioned(gb, Wt, nb, rank, n, src, exchange=exchange, px=px)
                sync_all()
                (d_loc, full, sweeps), ms = _time_cuda(torch, lambda: D.sssp_partitioned(gb, Wt, nb, rank, n, src, exchange=exchange, px=px))
                ms = max_over_ranks(ms)
                res[exchange] = {"ms": ms, "iterations": sweeps, "ms_per_iteration": ms / sweeps, "reached": int(full.present.sum()),
                                 "checksum": int(full.vals[full.present.bool()].sum())}
                del d_loc, full
            except Exception as exc:
                res[exchange] = {"error": repr(exc)}
            finally:
                if px is not None:
                    px.close()
        best = min((v for v in res.values() if "ms" in v), key=lambda v: v["ms"], default=None)
        out["sssp"] = {"exchange_variants": res, "local_nnz_rank0": local_nnz,
                       "mxv_algorithmic_GB_per_iteration_all_ranks": (nnz * 12 + (n + 1) * 8 + world * n * 8 + n * 8 * 2) / 1e9}
        if best:
            out["sssp"].update(best)
        del Wt
    if "pagerank" in which:
        atp, atc, _ = D.transpose_csr_torch(ip, c, None, n)
        pb = D.row_blocks_by_nnz(atp, world)
        k0, k1 = int(atp[pb[rank]]), int(atp[pb[rank + 1]])
        At = gb.cuda.matrix_from_device_csr((atp[pb[rank]:pb[rank + 1] + 1] - k0).contiguous(), atc[k0:k1].contiguous(),
                                            torch.ones(k1 - k0, dtype=torch.float64, device=dev), pb[rank + 1] - pb[rank], n)
        del atp, atc
        dloc = gb.cuda.vector_from_torch(torch.clamp(deg[pb[rank]:pb[rank + 1]], min=1).to(torch.float64))
        res = {}
        for exchange in (("nccl", "peer") if world > 1 else ("nccl",)):
            px = None
            try:
                # the peers' buffers are mapped once (CUDA IPC), outside the timed iterations; the warm-up also covers the
                # library's one-off kernel-selection trial for this CSR block (6-8 multiplies, format builds)
                px = D.PeerExchange(gb, gb.dtypes.FP64, n, pb, rank, dense=True) if exchange == "peer" else None
                D.pagerank_partitioned(gb, At, dloc, pb, rank, n, iters=10, exchange=exchange, px=px)
                sync_all()
                t_loc, ms = _time_cuda(torch, lambda: D.pagerank_partitioned(gb, At, dloc, pb, rank, n, iters=20, exchange=exchange, px=px))
                ms = max_over_ranks(ms)
                tl, _ = gb.cuda.vector_as_torch(t_loc, sync=False)
                ssum = torch.tensor([float(tl.sum())], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(ssum)
                res[exchange] = {"ms": ms, "ms_per_iteration": ms / 20, "sum": float(ssum[0])}
            except Exception as exc:   # the peer path needs CUDA IPC between the ranks; report rather than lose the run
                res[exchange] = {"error": repr(exc)}
            finally:
                if px is not None:
                    px.close()
        best = min((v for v in res.values() if "ms" in v), key=lambda v: v["ms"], default=None)
        out["pagerank"] = {"iterations": 20, "exchange_variants": res, "local_nnz_rank0": k1 - k0,
                           "mxv_algorithmic_GB_per_iteration_per_rank": ((k1 - k0) * 4 + (pb[rank + 1] - pb[rank] + 1) * 8 + n * 8 + (pb[rank + 1] - pb[rank]) * 9 * 2) / 1e9}
        if best:
            out["pagerank"].update(ms=best["ms"], ms_per_iteration=best["ms_per_iteration"], sum=best["sum"])
    return out


def mxv_fused_exchange(gb, D, M, x, sr, dtype, n, bounds, rank, iters, barrier, max_over_ranks, ev0, ev1):
    """ms per iteration of y = M.mxv(x) with the fused multiply + exchange: the SpMV epilogue stores every finished y(row) into the
    next x buffer of ALL ranks over NVLink peer memory (no collective on the data path; one tiny all-reduce orders the iterations).
    Returns {"peer": ms} or {"peer_error": text}; never raises (the NCCL number must survive a failure here)."""
    px = None
    try:
        px = D.PeerExchange(gb, dtype, n, bounds, rank, dense=True)
        px.fill_current(M.mxv(x, sr).new())

        def step():
            with px.writing(None, with_presence=True):
                yy = M.mxv(px.current, sr).new()
            px.advance()
            return yy

        for _ in range(5):
            step()
        barrier()
        ev0.record()
        for _ in range(iters):
            step()
        ev1.record()
        barrier()
        return {"peer": max_over_ranks(ev0.elapsed_time(ev1)) / iters}
    except Exception as exc:
        return {"peer_error": repr(exc)}
    finally:
        if px is not None:
            try:
                px.close()
            except Exception:
                pass


# ------------------------------------------------------------------ our arm
def ncu_traffic(mxv_kernel="spmv_merge_kernel"):
    """DRAM bytes per launch measured by ncu (profiles/traffic_r02.json if present, else the round-1 capture; produced by
    scripts/summarize_profiles.py from the `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` pass over the same
    workload).  mxm: ALL kernels of one A.mxm(A) step (hash kernels, compaction, row flops, binning); None when absent."""
    t = None
    for name in ("traffic_r02.json", "traffic_r01.json"):
        try:
            t = json.loads((ROOT / "profiles" / name).read_text())
            break
        except Exception:
            continue
    if t is None:
        return None, None
    mxm = mxv = None
    try:
        calls = t.get("mxm22_calls", 2)
        mxm = sum(r["dram_MB"] for r in t["mxm22"].values()) * 1e6 / calls
    except Exception:
        pass
    try:
        r = [v for k, v in t["mxv22"].items() if k.startswith(mxv_kernel)][0]
        mxv = r["dram_MB"] * 1e6 / r["launches"]
    except Exception:
        pass
    return mxm, mxv


def run_ours(args):
    import torch
    import torch.distributed as dist

    import graphblas_b200 as gb
    from graphblas_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gb.init(device=local)
    torch.cuda.set_stream(torch.cuda.Stream())   # a real (non-NULL) stream shared by torch events, NCCL ordering and our kernels
    gb.cuda.use_torch_stream()
    dev = torch.device("cuda", local)
    pk, pk_kind = peaks()
    hbm = float(pk.get("hbm_gbs", 6650.0))
    scale = args.scale
    sr = gb.semiring.plus_times

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    def mxm_partition(indptr, cols, n):
        """row bounds of this rank: equal prefix of the per-row flop bounds"""
        if world == 1:
            return 0, n
        deg = (indptr[1:] - indptr[:-1])
        rowflops = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        rowflops[1:].index_add_(0, torch.repeat_interleave(torch.arange(n, device=dev), deg), deg[cols.long()])
        cost = torch.zeros(n + 1, dtype=torch.float64, device=dev)
        cost[1:] = D.mxm_row_costs(rowflops[1:])
        b = D.row_blocks_by_prefix(torch.cumsum(cost, 0).cpu().numpy(), world)
        return b[rank], b[rank + 1]

    # ---------------- inputs: every rank generates the matrix from the seed itself (B = A replicated, no input traffic)
    indptr, cols, n = rmat_csr_torch(scale, RMAT_2A, 42, device=dev)
    vals = values_torch(cols.numel(), 43, torch.float32, device=dev)
    nnz = cols.numel()
    B = gb.cuda.matrix_from_device_csr(indptr, cols, vals, n, n)
    r0, r1 = mxm_partition(indptr, cols, n)
    k0, k1 = int(indptr[r0]), int(indptr[r1])
    a_ptr = (indptr[r0:r1 + 1] - k0).contiguous()
    a_cols, a_vals = cols[k0:k1].contiguous(), vals[k0:k1].contiguous()
    A = gb.cuda.matrix_from_device_csr(a_ptr, a_cols, a_vals, r1 - r0, n) if world > 1 else B
    torch.cuda.synchronize()

    # ---------------- device-resident timing (value)
    C = None
    for _ in range(args.warmup):
        C = None
        C = A.mxm(B, sr).new()
    nnz_c_local = C.nvals if C is not None else A.mxm(B, sr).new().nvals
    C = None
    barrier()
    launches1 = gb.cuda.launch_count()
    with ClockSampler(local) as clk:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            C = None   # the previous result is released (stream-ordered) before the next is built
            C = A.mxm(B, sr).new()
        ev1.record()
        barrier()
    ms_dev = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    launches_per_step = (gb.cuda.launch_count() - launches1) / args.steps
    nnz_c = sum_over_ranks(float(nnz_c_local))
    value = nnz_c / (ms_dev * 1e-3)
    flops_local, _ = gb.cuda.mxm_symbolic(A, B)
    flops = sum_over_ranks(float(flops_local))

    # The product is left as the hash kernels wrote it: rows in row order but with a few unused slots between them (row-end CSR)
    # and unsorted inside a row.  Later multiplies read that form directly; export / sort / SpMV squeeze it once.  Both
    # on-demand costs are measured here and reported; `value_compact` is the rate with the compaction forced into every step.
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
    gb.cuda.matrix_compact(C); gb.cuda.matrix_sort(C)   # once untimed: the first compaction grows the memory pool by the exact-size arrays
    C = None
    C = A.mxm(B, sr).new()
    t0.record(); gb.cuda.matrix_compact(C); t1.record(); gb.cuda.matrix_sort(C); t2.record(); torch.cuda.synchronize()
    compact_ms = max_over_ranks(t0.elapsed_time(t1))
    sort_ms = max_over_ranks(t1.elapsed_time(t2))
    C = None
    barrier()
    ev0.record()
    for _ in range(args.steps):
        C = None
        C = A.mxm(B, sr).new()
        gb.cuda.matrix_compact(C)
    ev1.record()
    barrier()
    ms_compact = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps

    # ---------------- per-kernel breakdown of one step (profile mode serialises; explains the roofline entry)
    gb.cuda.set_option("profile", "1")
    gb.cuda.kernel_times(reset=True)
    C = None
    C = A.mxm(B, sr).new()
    kt = gb.cuda.kernel_times(reset=True)
    gb.cuda.set_option("profile", "0")
    numeric_ms = sum(ms for k, (ms, cnt) in kt.items() if k.startswith("spgemm_numeric"))
    symbolic_ms = sum(ms for k, (ms, cnt) in kt.items() if k.startswith("spgemm_symbolic"))
    # roofline.achieved for mxm := B_min / t / peak, B_min = (nnzA + nnzB + nnzC) (s_idx + s_val) + 3 (n + 1) s_ptr,
    # t = the whole step (row flops + binning + numeric + compaction).  The no-reuse gather bound of the numeric phase is kept
    # beside it as an explanation, not as the fraction.
    bmin_bytes = ((k1 - k0) + nnz + nnz_c_local) * 8 + 3 * (n + 1) * 8
    gather_bytes = (k1 - k0) * 8 + flops_local * 8 + nnz_c_local * 8 + (r1 - r0 + 1) * 8 * 2
    bmin_gbs = bmin_bytes / (ms_dev * 1e-3) / 1e9
    C = None

    # ---------------- end to end through the C-ABI with host buffers
    e2e = None
    e2e_std = None
    if not args.no_e2e:
        reps = max(1, args.steps // 2)
        # (1) headline: the 32-bit-index import / export entry points with pinned host buffers
        h_ptr = torch.empty(a_ptr.numel(), dtype=torch.int64).pin_memory(); h_ptr.copy_(a_ptr)
        h_col = torch.empty(a_cols.numel(), dtype=torch.int32).pin_memory(); h_col.copy_(a_cols)
        h_val = torch.empty(a_vals.numel(), dtype=torch.float32).pin_memory(); h_val.copy_(a_vals)
        o_ptr = torch.empty(r1 - r0 + 1, dtype=torch.int64).pin_memory()
        o_col = torch.empty(int(nnz_c_local), dtype=torch.int32).pin_memory()
        o_val = torch.empty(int(nnz_c_local), dtype=torch.float32).pin_memory()
        h2d = h_ptr.numel() * 8 + h_col.numel() * 4 + h_val.numel() * 4
        d2h = o_ptr.numel() * 8 + o_col.numel() * 4 + o_val.numel() * 4

        def e2e_step(blocks):
            Ah = gb.cuda.matrix_from_host_csr32(h_ptr.numpy(), h_col.numpy(), h_val.numpy(), r1 - r0, n)
            if blocks <= 1:   # one GrB_mxm, then one export
                Ch = Ah.mxm(B if world > 1 else Ah, sr).new()
                gb.cuda.matrix_export_host_csr32(Ch, o_ptr.numpy(), o_col.numpy(), o_val.numpy(), sort=False)
            else:             # the product leaves the device in row blocks while later blocks are still being multiplied
                gb.cuda.mxm_to_host_csr32(Ah, B if world > 1 else Ah, sr, o_ptr.numpy(), o_col.numpy(), o_val.numpy(), blocks=blocks)

        def time_e2e(blocks):
            e2e_step(blocks)
            barrier()
            t_e = time.perf_counter()
            ev0.record()
            for _ in range(reps):
                e2e_step(blocks)
            ev1.record()
            barrier()
            return max_over_ranks(ev0.elapsed_time(ev1)) / reps, (time.perf_counter() - t_e) * 1e3 / reps

        def checksum():
            # order-independent inside a row (rows come out unsorted): row pointers, and the column sum over whole leading rows
            cut = int(o_ptr[min(o_ptr.numel() - 1, 4096)])
            return (int(o_ptr[-1]), int(o_ptr[: 1 << 16].sum()), int(o_col[:cut].long().sum()))

        ms_single, _ = time_e2e(1)
        check = checksum()
        ms_e2e, wall_e2e = time_e2e(args.e2e_blocks)
        check2 = checksum()
        api_single = "GrB_cuda_Matrix_import_csr32 / one GrB_mxm / GrB_cuda_Matrix_export_csr32 (int32 column indices, pinned host buffers)"
        api_blocks = (f"GrB_cuda_Matrix_import_csr32, then graphblas_b200.cuda.mxm_to_host_csr32: GrB_mxm on {args.e2e_blocks} row blocks, each exported "
                      "with GrB_cuda_Matrix_export_csr32_async while the next is multiplied (int32 column indices, pinned host buffers)")
        # both are calls a user can make; the headline is the faster one (the D2H of the 20 GB result at PCIe speed bounds either)
        best_ms, best_api = (ms_single, api_single) if ms_single <= ms_e2e else (ms_e2e, api_blocks)
        e2e = {"value": nnz_c / (best_ms * 1e-3), "unit": "nnz-out/s", "ms_per_step": best_ms,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "api": best_api,
               "variants": {"single_call": {"ms_per_step": ms_single, "value": nnz_c / (ms_single * 1e-3), "api": api_single},
                            "row_blocks_overlapped": {"ms_per_step": ms_e2e, "wall_ms_per_step": wall_e2e, "value": nnz_c / (ms_e2e * 1e-3), "api": api_blocks}},
               "same_result": check == check2 if world == 1 else None}
        del h_ptr, h_col, h_val, o_ptr, o_col, o_val
        # (2) what the reference's Matrix.from_csr / to_csr hand over (graphblas/core/matrix.py:992-1068, 1601-1645): uint64 index
        # arrays in pageable numpy memory through GrB_Matrix_import_FP32 / GrB_Matrix_export_FP32 -- the host mirror's from_csr / to_csr
        if world == 1 and not args.no_e2e_std:
            try:
                hp = a_ptr.cpu().numpy().astype(np.uint64)
                hc = a_cols.cpu().numpy().astype(np.uint64)
                hv = a_vals.cpu().numpy()

                def std_step():
                    Ah = gb.Matrix.from_csr(hp, hc, hv, ncols=n)
                    Ch = Ah.mxm(Ah, sr).new()
                    return Ch.to_csr(sort=False)

                std_step()
                torch.cuda.synchronize()
                t_s = time.perf_counter()
                out = std_step()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t_s
                e2e_std = {"value": nnz_c / dt, "unit": "nnz-out/s", "ms_per_step": dt * 1e3,
                           "h2d_bytes_per_step": int(hp.nbytes + hc.nbytes + hv.nbytes), "d2h_bytes_per_step": int(sum(o.nbytes for o in out)),
                           "api": "Matrix.from_csr -> GrB_Matrix_import_FP32, GrB_mxm, Matrix.to_csr -> GrB_Matrix_export_FP32 (uint64 indices, pageable numpy)",
                           "timing": "host wall clock, 1 repetition (allocations of the 30 GB result arrays included)"}
                del out, hp, hc, hv
            except Exception as exc:
                e2e_std = {"error": repr(exc)}

    # ---------------- mxv: plus_times fp32 on the Graph500-skew matrix, rows partitioned by equal nnz, x all-gathered per iteration
    mxv = None
    host_2b = None
    if not args.no_mxv:
        del B, A
        B = A = None
        gb.cuda.set_option("trim", "1")   # hand the cached 10-20 GB mxm blocks back before torch builds the next input
        ip2, c2, n2 = rmat_csr_torch(scale, RMAT_2B, 42, device=dev)
        v2 = values_torch(c2.numel(), 45, torch.float32, device=dev)
        nnz2 = c2.numel()
        if rank == 0 and world == 1 and not args.no_cpu:
            host_2b = (ip2.cpu().numpy(), c2.cpu().numpy().astype(np.int64))
        nb = D.row_blocks_by_nnz(ip2, world)
        q0, q1 = nb[rank], nb[rank + 1]
        p0, p1 = int(ip2[q0]), int(ip2[q1])
        M = gb.cuda.matrix_from_device_csr((ip2[q0:q1 + 1] - p0).contiguous(), c2[p0:p1].contiguous(), v2[p0:p1].contiguous(), q1 - q0, n2)
        x = gb.cuda.vector_from_torch(values_torch(n2, 46, torch.float32, device=dev))
        # the per-iteration exchange of the row-partitioned path: every rank's output slice is all-gathered straight into
        # the device buffer of the next input vector (zero-copy torch views of the library's arrays; NCCL over NVLink)
        x_next = D.GatheredVector(gb, gb.dtypes.FP32, n2, nb, sparse=False) if world > 1 else None
        iters = max(20, args.steps * 4)

        def mxv_iter(xv):
            y = M.mxv(xv, sr).new()
            if world > 1:
                x_next.gather(y)
            return y

        y = None
        for _ in range(10):   # covers the library's one-off kernel-selection trial (6 multiplies) as well
            y = mxv_iter(x)
        barrier()
        ev0.record()
        for _ in range(iters):
            y = mxv_iter(x)
        ev1.record()
        barrier()
        ms_mxv = max_over_ranks(ev0.elapsed_time(ev1)) / iters
        mxv_variants = {"nccl": ms_mxv}
        if world > 1:
            mxv_variants.update(mxv_fused_exchange(gb, D, M, x, sr, gb.dtypes.FP32, n2, nb, rank, iters, barrier, max_over_ranks, ev0, ev1))
            ms_mxv = min(v for k, v in mxv_variants.items() if isinstance(v, float))
        # the dominant kernel alone (library profile mode: CUDA events around each launch on the library stream)
        gb.cuda.set_option("profile", "1")
        gb.cuda.kernel_times(reset=True)
        for _ in range(5):
            y = M.mxv(x, sr).new()
        ktm = gb.cuda.kernel_times(reset=True)
        gb.cuda.set_option("profile", "0")
        kname = max((k for k in ktm if k.startswith("spmv_") and not k.endswith("fixup")), key=lambda k: ktm[k][0], default=None)
        kernel_ms = max_over_ranks(ktm[kname][0] / ktm[kname][1]) if kname else ms_mxv
        # algorithmic bytes: nnz*(s_idx+s_val) + (nrows+1)*s_ptr + ncols*s_x + nrows*(s_y + 1 presence byte)
        bytes_local = (p1 - p0) * 8 + (q1 - q0 + 1) * 8 + n2 * 4 + (q1 - q0) * 5
        bytes_total = sum_over_ranks(float(bytes_local))
        bytes_max = max_over_ranks(float(bytes_local))
        gbs = bytes_total / (ms_mxv * 1e-3) / 1e9
        gbs_kernel = bytes_max / (kernel_ms * 1e-3) / 1e9
        kfull = {"spmv_merge": "spmv_merge_kernel", "spmv_seg": "spmv_seg_kernel", "spmv_seg_hot": "spmv_seg_kernel", "spmv_band": "spmv_band_kernel"}.get(kname, kname or "?")
        mxv = {"workload": f"R-MAT scale-{scale} (0.57,0.19,0.19,0.05) plus_times fp32 A.mxv(x), x dense", "nnz": nnz2,
               "partition": f"equal-nnz row blocks x{world}" + ("; one all-gather of the fp32 slice per iteration" if world > 1 else ""),
               "ms_per_iter": ms_mxv, "ms_per_iter_by_exchange": mxv_variants, "GB_per_s": gbs, "nnz_per_s": nnz2 / (ms_mxv * 1e-3),
               "roofline": {"bound": "hbm", "kernel": f"{kfull} (chosen by the library's timed trial)", "achieved": gbs_kernel, "peak": hbm,
                            "unit": "GB/s", "frac": gbs_kernel / hbm, "peak_source": pk_kind, "kernel_us": kernel_ms * 1e3,
                            "algorithmic_bytes": int(bytes_max), "frac_whole_call": gbs / world / hbm,
                            "traffic": (ncu_traffic(kfull)[1] if (scale == 22 and world == 1) else None)}}
        del M, x, y, x_next, ip2, c2, v2
        gb.cuda.set_option("trim", "1")

    # ---------------- the iterative workloads of BASELINE.json configs 3-5
    workloads = None
    if not args.no_workloads:
        try:
            if world == 1:
                workloads = run_workloads(gb, torch, dev, scale)
            else:
                workloads = run_workloads_partitioned(gb, torch, dist, dev, scale, rank, world, max_over_ranks)
        except Exception as exc:
            workloads = {"error": repr(exc)}
        gb.cuda.set_option("trim", "1")

    # ---------------- scale 25 (north_star: ">= 6x the 1-GPU nnz-out/s at 8 GPUs on scale-25"; config 5: PageRank scale-25)
    big = None
    if not args.no_scale25 and scale == 22:
        try:
            big = run_scale25(gb, torch, dist, D, dev, rank, world, barrier, max_over_ranks, sum_over_ranks)
        except Exception as exc:
            big = {"error": repr(exc)}
        gb.cuda.set_option("trim", "1")

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- CPU baselines (rank 0, N == 1 only): the in-repo OpenMP port on all host cores, bounded samples
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            hp, hc, hv = indptr.cpu().numpy(), cols.cpu().numpy().astype(np.int64), vals.cpu().numpy()
            rate, desc, threads, _, _ = cpu_mxm_sample(hp, hc, hv, n, budget_s=args.cpu_budget)
            cpu = {"value": rate, "unit": "nnz-out/s", "cores": threads, "kind": "port", "sample": desc,
                   "note": "SuiteSparse:GraphBLAS unavailable on this box -- baseline is the in-repo OpenMP SpGEMM (one pass, unsorted rows, hash / dense Gustavson accumulators)"}
            del hp, hc, hv
            if host_2b is not None:
                cpu["vector_multiplies"] = cpu_mxv_legs(host_2b[0], host_2b[1], n, scale)
                if mxv:
                    cpu["vector_multiplies"]["gpu_over_cpu_mxv_plus_times_fp32"] = mxv["nnz_per_s"] / cpu["vector_multiplies"]["mxv_plus_times_fp32"]["nnz_per_s"]
                if workloads and "sssp" in workloads:
                    gpu_rate = workloads["nnz"] / (workloads["sssp"]["ms_per_iteration"] * 1e-3)
                    cpu["vector_multiplies"]["gpu_over_cpu_vxm_min_plus_int64"] = gpu_rate / cpu["vector_multiplies"]["vxm_min_plus_int64"]["nnz_per_s"]
        except Exception as exc:   # never lose the GPU numbers because the CPU leg failed
            cpu = cpu or {"value": None}
            cpu["error"] = repr(exc)

    line = {
        "metric": "mxm nnz-out/s (R-MAT plus_times fp32)", "value": value, "unit": "nnz-out/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(scale=scale), "l2": L2_NOTE},
        "problem": {"n": n, "nnz_A": nnz, "nnz_C": int(nnz_c), "flops": int(flops), "parallelism": f"row-partition x{world} (equal estimated cost: flops, split rows weighted), B replicated",
                    "l2": "inputs (0.5 GB) and outputs (>=GBs) exceed the 126 MB L2",
                    "result_form": "row-end CSR (rows in order, unused slots between rows where products merged), columns unsorted inside a row; "
                                   "compaction and sort happen on demand and are reported in phases_ms"},
        "value_compact": {"value": nnz_c / (ms_compact * 1e-3), "ms_per_step": ms_compact,
                          "note": "the same steps with the result squeezed to a compact CSR inside every step (what export / SpMV / sort trigger once)"},
        "clocks": clk.summary(), "gpu_launches": launches_per_step,
        "phases_ms": {"symbolic": symbolic_ms, "numeric": numeric_ms, "compact_on_demand": compact_ms, "sort_on_demand": sort_ms, "kernels": {k: v[0] for k, v in kt.items()}},
        "roofline": {"bound": "hbm", "kernel": "whole A.mxm(A) step (row flops, binning, hash numeric kernels, count scan)",
                     "achieved": bmin_gbs, "peak": hbm, "unit": "GB/s", "frac": bmin_gbs / hbm, "peak_source": pk_kind,
                     "definition": "SURVEY.md 8(d): B_min / t_step / peak, B_min = (nnzA + nnzB + nnzC)(s_idx + s_val) + 3(n + 1) s_ptr",
                     "algorithmic_bytes": int(bmin_bytes),
                     "traffic": (ncu_traffic()[0] if (scale == 22 and world == 1) else None),
                     "traffic_note": "ncu dram bytes of ALL kernels of one A.mxm(A) (profiles/ncu_mxm22_r0*.txt)",
                     "numeric_gather_bound": {"bytes": int(gather_bytes), "GB_per_s": gather_bytes / (numeric_ms * 1e-3) / 1e9 if numeric_ms > 0 else None,
                                              "frac": gather_bytes / (numeric_ms * 1e-3) / 1e9 / hbm if numeric_ms > 0 else None,
                                              "note": "no-reuse gather bytes of the numeric hash kernels over their own time (explanatory)"}},
        "e2e": e2e, "e2e_reference_interface": e2e_std, "mxv": mxv, "workloads": workloads, "scale25": big, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_scale25(gb, torch, dist, D, dev, rank, world, barrier, max_over_ranks, sum_over_ranks):
    """R-MAT scale 25.  (a) A.mxm(A) plus_times fp32 (2a parameters): rows split by equal flops over the ranks; a rank whose share
    of the result does not fit its memory budget walks its rows in blocks and drops each block's result after counting it (at
    N = 1 the 2.6e10-entry product is 8 blocks), so nnz-out/s is measured with the same kernels at every N.  (b) PageRank on
    the Graph500-skew graph, rows of A' partitioned by equal nnz (BASELINE config 5)."""
    scale = 25
    out = {}
    sr = gb.semiring.plus_times
    indptr, cols, n = rmat_csr_torch(scale, RMAT_2A, 42, device=dev)
    vals = values_torch(cols.numel(), 43, torch.float32, device=dev)
    nnz = cols.numel()
    B = gb.cuda.matrix_from_device_csr(indptr, cols, vals, n, n)
    deg = indptr[1:] - indptr[:-1]
    rowflops = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowflops[1:].index_add_(0, torch.repeat_interleave(torch.arange(n, device=dev), deg), deg[cols.long()])
    cum = torch.cumsum(rowflops, 0)
    cost = torch.zeros(n + 1, dtype=torch.float64, device=dev)
    cost[1:] = D.mxm_row_costs(rowflops[1:])
    del rowflops
    cum_h = cum.cpu().numpy()
    total_flops = int(cum_h[-1])
    b = D.row_blocks_by_prefix(torch.cumsum(cost, 0).cpu().numpy(), world)   # ranks: equal estimated cost (heavy rows weigh more)
    del cost
    r0, r1 = b[rank], b[rank + 1]
    # blocks of at most ~3e9 products: result (8 B / entry) + staging stay below ~50 GB
    my_flops = int(cum_h[r1] - cum_h[r0])
    nblk = max(1, -(-my_flops // 3_000_000_000))
    sub = D.row_blocks_by_prefix(cum_h[r0:r1 + 1] - cum_h[r0], nblk)
    blocks = []
    for s0, s1 in zip(sub[:-1], sub[1:]):
        q0, q1 = r0 + s0, r0 + s1
        k0, k1 = int(indptr[q0]), int(indptr[q1])
        blocks.append(gb.cuda.matrix_from_device_csr((indptr[q0:q1 + 1] - k0).contiguous(), cols[k0:k1].contiguous(), vals[k0:k1].contiguous(), q1 - q0, n))
    del cum

    def one_pass():
        tot = 0
        for Ab in blocks:
            C = Ab.mxm(B, sr).new()
            tot += C.nvals
            C = None
        return tot

    nnz_local = one_pass()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    reps = 2
    for _ in range(reps):
        one_pass()
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1)) / reps
    nnz_c = sum_over_ranks(float(nnz_local))
    out["mxm"] = {"workload": "R-MAT scale-25 (0.45,0.15,0.15,0.25) ef16 seed42 A.mxm(A) plus_times fp32", "n": n, "nnz_A": nnz,
                  "flops": total_flops, "nnz_C": int(nnz_c), "ms_per_step": ms, "value": nnz_c / (ms * 1e-3), "unit": "nnz-out/s",
                  "row_blocks_per_rank": nblk, "note": "each block's result is counted and released before the next block is formed"}
    del blocks, B, indptr, cols, vals
    gb.cuda.set_option("trim", "1")
    # (b) plus_times fp32 mxv on the Graph500-skew scale-25 matrix, rows by equal nnz, x all-gathered every iteration
    ip2, c2, n2 = rmat_csr_torch(scale, RMAT_2B, 44, device=dev)
    nb = D.row_blocks_by_nnz(ip2, world)
    q0, q1 = nb[rank], nb[rank + 1]
    p0, p1 = int(ip2[q0]), int(ip2[q1])
    M = gb.cuda.matrix_from_device_csr((ip2[q0:q1 + 1] - p0).contiguous(), c2[p0:p1].contiguous(),
                                       values_torch(p1 - p0, 45 + rank, torch.float32, device=dev), q1 - q0, n2)
    nnz2 = c2.numel()
    del ip2, c2
    x = gb.cuda.vector_from_torch(values_torch(n2, 46, torch.float32, device=dev))
    x_next = D.GatheredVector(gb, gb.dtypes.FP32, n2, nb, sparse=False) if world > 1 else None

    def mxv_iter():
        y = M.mxv(x, sr).new()
        if world > 1:
            x_next.gather(y)
        return y

    for _ in range(10):
        mxv_iter()
    barrier()
    ev0.record()
    for _ in range(20):
        mxv_iter()
    ev1.record()
    barrier()
    ms_mxv = max_over_ranks(ev0.elapsed_time(ev1)) / 20
    variants25 = {"nccl": ms_mxv}
    if world > 1:
        variants25.update(mxv_fused_exchange(gb, D, M, x, sr, gb.dtypes.FP32, n2, nb, rank, 20, barrier, max_over_ranks, ev0, ev1))
        ms_mxv = min(v for v in variants25.values() if isinstance(v, float))
    bytes_local = (p1 - p0) * 8 + (q1 - q0 + 1) * 8 + n2 * 4 + (q1 - q0) * 5
    out["mxv"] = {"workload": "R-MAT scale-25 (0.57,0.19,0.19,0.05) plus_times fp32 A.mxv(x), x dense, equal-nnz row blocks, x exchanged per iteration "
                              "(NCCL all-gather, or the fused SpMV + peer-store exchange)",
                  "nnz": nnz2, "ms_per_iter": ms_mxv, "ms_per_iter_by_exchange": variants25, "nnz_per_s": nnz2 / (ms_mxv * 1e-3),
                  "GB_per_s_all_ranks": sum_over_ranks(float(bytes_local)) / (ms_mxv * 1e-3) / 1e9}
    del M, x, x_next
    gb.cuda.set_option("trim", "1")
    w = run_workloads_partitioned(gb, torch, dist, dev, 25, rank, world, max_over_ranks, which=("pagerank",))
    out["pagerank"] = w.get("pagerank")
    out["pagerank_graph"] = {k: w[k] for k in ("graph", "n", "nnz", "partition", "exchange")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=int(os.environ.get("GRB_BENCH_SCALE", "22")))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-std", action="store_true")
    ap.add_argument("--no-mxv", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-workloads", action="store_true")
    ap.add_argument("--no-scale25", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--e2e-blocks", type=int, default=12)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
