#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's metric: mxm nnz-out/s (+ mxv GB/s) on synthetic R-MAT CSR.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

A "step" is one A.mxm(A, plus_times) fp32 on the R-MAT matrix of BASELINE.json configs[1]
("R-MAT scale-22 avg-deg-16 A.mxm(A) plus_times fp32 on 1 B200"), in the feasible variant SURVEY.md
section 8(d) names "2a": R-MAT (a,b,c,d)=(0.45,0.15,0.15,0.25), edge factor 16, seed 42, deduplicated
(Graph500 skew at scale 22 would emit ~6e10 entries = 0.5 TB, which no single GPU holds).
`value` = nnz(C) / device time with A resident in HBM (symbolic + numeric; the result is left "jumbled",
as the reference's C library also leaves it -- sort time is reported separately).
`e2e`   = the same through the C-ABI with HOST buffers: H2D of A's CSR + mxm + D2H of C's CSR per step.
`mxv`   = plus_times fp32 A.mxv(x) on the Graph500-skew matrix "2b" (the north_star's >=50%-of-HBM target).
One process per GPU (torchrun); N>1 row-partitions A by equal flops, B = A replicated (shipped once with NCCL
broadcast, outside the timed region); time = max over ranks.
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


def ncu_traffic(mxv_kernel="spmv_merge_kernel"):
    """DRAM bytes per launch measured by ncu (profiles/traffic_r01.json, produced by scripts/summarize_profiles.py from the
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` capture of the same workload); None when absent."""
    try:
        t = json.loads((ROOT / "profiles" / "traffic_r01.json").read_text())
    except Exception:
        return None, None
    mxm = mxv = None
    try:
        rows = [v for k, v in t["mxm22"].items() if k.startswith(("spgemm_block_kernel", "spgemm_warp_kernel", "spgemm_block_elect", "spgemm_warp_elect", "spgemm_group"))]
        calls = 2   # the capture ran A.mxm(A) twice
        mxm = sum(r["dram_MB"] for r in rows) * 1e6 / calls
    except Exception:
        pass
    try:
        r = [v for k, v in t["mxv22"].items() if k.startswith(mxv_kernel)][0]
        mxv = r["dram_MB"] * 1e6 / r["launches"]
    except Exception:
        pass
    return mxm, mxv


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


# ------------------------------------------------------------------ the reference arm (CPU)
def cpu_mxm_sample(indptr, indices, values, n, budget_s=15.0, seed=0, rows_hint=None):
    """Times the CPU baseline (oracle/grb_oracle.c `oracle_mxm_baseline_f32`: OpenMP, one numeric pass, unsorted rows, 32-bit
    indices, hash accumulator for short rows / dense Gustavson workspace for long ones -- the way SuiteSparse's saxpy3 organises
    it) on a random row sample sized for ~budget_s of CPU work.  The timed region covers the per-row flop bound, its prefix and
    the numeric pass, not the 32-bit index conversion (input preparation).  Returns (nnz_out_per_s, description, threads,
    rows_used, seconds).  `rows_hint` skips the sizing search."""
    from oracle import bigref as R

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


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  SuiteSparse:GraphBLAS is not installable here
    (no network, not vendored), so this arm times the oracle port on all host cores, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import bigref as R

    scale = args.scale
    import helpers

    r, c, n = helpers.rmat_edges(scale, a=RMAT_2A[0], b=RMAT_2A[1], c=RMAT_2A[2], seed=42)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, r + 1, 1)
    np.cumsum(indptr, out=indptr)
    vals = np.random.default_rng(43).random(r.size).astype(np.float32)
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
        "config": {"workload": f"R-MAT scale-{scale} (0.45,0.15,0.15,0.25) ef16 seed42 A.mxm(A) plus_times fp32", "sample": desc},
        "cpu_baseline": {"value": value, "unit": "nnz-out/s", "cores": threads, "kind": "port", "sample": desc,
                         "note": "SuiteSparse unavailable -- baseline is the in-repo OpenMP SpGEMM (one pass, unsorted rows, hash / dense Gustavson accumulators)"},
        "e2e": {"value": value, "unit": "nnz-out/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))



def run_workloads(gb, torch, dev, scale):
    """Level BFS (any_pair, complemented structural mask + replace), SSSP (min_plus, min accum, int64) and PageRank
    (plus_second on A.T, fp64) on the Graph500-skew R-MAT, written exactly like the reference notebooks.
    Device-resident; times are CUDA-event times of the whole loop, including the O(n) vector ops around the multiply."""
    ip, c, n = rmat_csr_torch(scale, RMAT_2B, 42, device=dev)
    nnz = c.numel()
    deg = ip[1:] - ip[:-1]
    src = int(torch.nonzero(deg > 0)[0])
    out = {"graph": f"R-MAT scale-{scale} (0.57,0.19,0.19,0.05) ef16 seed42", "n": n, "nnz": nnz}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

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

    v, levels = bfs()
    torch.cuda.synchronize()
    ev0.record()
    v, levels = bfs()
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
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

    d, its = sssp()
    torch.cuda.synchronize()
    ev0.record()
    d, its = sssp()
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
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

    t = pagerank(2)
    torch.cuda.synchronize()
    ev0.record()
    t = pagerank(20)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    out["pagerank"] = {"ms": ms, "iterations": 20, "ms_per_iteration": ms / 20, "sum": t.reduce(gb.monoid.plus).value,
                       "mxv_algorithmic_GB_per_iteration": (nnz * 4 + (n + 1) * 8 + n * 8 + n * 9 * 2) / 1e9}
    return out


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import graphblas_b200 as gb

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

    # ---------------- inputs: generated once on the device (rank 0 ships B = A to the others with NCCL broadcast)
    if rank == 0:
        indptr, cols, n = rmat_csr_torch(scale, RMAT_2A, 42, device=dev)
        vals = values_torch(cols.numel(), 43, torch.float32, device=dev)
        meta = torch.tensor([n, cols.numel()], dtype=torch.int64, device=dev)
    else:
        meta = torch.zeros(2, dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(meta, 0)
        n, nnz = int(meta[0]), int(meta[1])
        if rank != 0:
            indptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
            cols = torch.empty(nnz, dtype=torch.int32, device=dev)
            vals = torch.empty(nnz, dtype=torch.float32, device=dev)
        for t in (indptr, cols, vals):
            dist.broadcast(t, 0)
    n, nnz = int(meta[0]), int(meta[1])
    B = gb.cuda.matrix_from_device_csr(indptr, cols, vals, n, n)
    # row partition by equal flops prefix
    if world > 1:
        deg = (indptr[1:] - indptr[:-1])
        rowflops = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        seg = torch.repeat_interleave(torch.arange(n, device=dev), deg)
        rowflops[1:].index_add_(0, seg, deg[cols.long()])
        cum = torch.cumsum(rowflops, 0)
        total = int(cum[-1])
        bounds = torch.searchsorted(cum, torch.tensor([total * k // world for k in range(world + 1)], device=dev, dtype=torch.int64))
        bounds[0], bounds[-1] = 0, n
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        del seg, rowflops, cum
    else:
        r0, r1 = 0, n
    k0, k1 = int(indptr[r0]), int(indptr[r1])
    a_ptr = (indptr[r0:r1 + 1] - k0).contiguous()
    a_cols, a_vals = cols[k0:k1].contiguous(), vals[k0:k1].contiguous()
    A = gb.cuda.matrix_from_device_csr(a_ptr, a_cols, a_vals, r1 - r0, n) if world > 1 else B
    sr = gb.semiring.plus_times
    torch.cuda.synchronize()

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

    # ---------------- device-resident timing (value)
    launches0 = gb.cuda.launch_count()
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

    # sort-on-demand cost (not part of `value`; the reference's library is lazy here too)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); gb.cuda.matrix_sort(C); t1.record(); torch.cuda.synchronize()
    sort_ms = max_over_ranks(t0.elapsed_time(t1))

    # ---------------- per-kernel breakdown of one step (profile mode serialises; used for the roofline entry only)
    gb.cuda.set_option("profile", "1")
    gb.cuda.kernel_times(reset=True)
    C = None
    C = A.mxm(B, sr).new()
    kt = gb.cuda.kernel_times(reset=True)
    gb.cuda.set_option("profile", "0")
    numeric_ms = sum(ms for k, (ms, cnt) in kt.items() if k.startswith("spgemm_numeric"))
    symbolic_ms = sum(ms for k, (ms, cnt) in kt.items() if k.startswith("spgemm_symbolic"))
    # algorithmic bytes of the numeric phase: read A once, read the B rows every product touches
    # (flops * (4 + 4) bytes, the no-reuse gather bound of SURVEY.md 8d), write C once
    numeric_bytes = (k1 - k0) * 8 + flops_local * 8 + nnz_c_local * 8 + (r1 - r0 + 1) * 8 * 2
    bmin_bytes = ((k1 - k0) + nnz + nnz_c_local) * 8 + 3 * (n + 1) * 8
    roof_ach = numeric_bytes / (numeric_ms * 1e-3) / 1e9 if numeric_ms > 0 else None
    C = None

    # ---------------- end to end through the C-ABI with host buffers
    e2e = None
    if not args.no_e2e:
        h_ptr = torch.empty(a_ptr.numel(), dtype=torch.int64).pin_memory(); h_ptr.copy_(a_ptr)
        h_col = torch.empty(a_cols.numel(), dtype=torch.int32).pin_memory(); h_col.copy_(a_cols)
        h_val = torch.empty(a_vals.numel(), dtype=torch.float32).pin_memory(); h_val.copy_(a_vals)
        o_ptr = torch.empty(r1 - r0 + 1, dtype=torch.int64).pin_memory()
        o_col = torch.empty(int(nnz_c_local), dtype=torch.int32).pin_memory()
        o_val = torch.empty(int(nnz_c_local), dtype=torch.float32).pin_memory()
        h2d = h_ptr.numel() * 8 + h_col.numel() * 4 + h_val.numel() * 4
        d2h = o_ptr.numel() * 8 + o_col.numel() * 4 + o_val.numel() * 4

        def e2e_step():
            Ah = gb.cuda.matrix_from_host_csr32(h_ptr.numpy(), h_col.numpy(), h_val.numpy(), r1 - r0, n)
            Ch = Ah.mxm(B if world > 1 else Ah, sr).new()
            gb.cuda.matrix_export_host_csr32(Ch, o_ptr.numpy(), o_col.numpy(), o_val.numpy(), sort=False)

        e2e_step()
        barrier()
        t_e = time.perf_counter()
        ev0.record()
        for _ in range(max(1, args.steps // 2)):
            e2e_step()
        ev1.record()
        barrier()
        ms_e2e = max_over_ranks(ev0.elapsed_time(ev1)) / max(1, args.steps // 2)
        e2e = {"value": nnz_c / (ms_e2e * 1e-3), "unit": "nnz-out/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "wall_ms_per_step": (time.perf_counter() - t_e) * 1e3 / max(1, args.steps // 2)}
        del h_ptr, h_col, h_val, o_ptr, o_col, o_val

    # ---------------- mxv: plus_times fp32 on the Graph500-skew matrix, rows partitioned by nnz, x all-gathered per iteration
    mxv = None
    if not args.no_mxv:
        del B, A
        B = A = None
        gb.cuda.set_option("trim", "1")   # hand the cached 10-20 GB mxm blocks back before torch builds the next input
        if rank == 0:
            ip2, c2, n2 = rmat_csr_torch(scale, RMAT_2B, 42, device=dev)
            v2 = values_torch(c2.numel(), 45, torch.float32, device=dev)
            meta2 = torch.tensor([n2, c2.numel()], dtype=torch.int64, device=dev)
        else:
            meta2 = torch.zeros(2, dtype=torch.int64, device=dev)
        if world > 1:
            dist.broadcast(meta2, 0)
            if rank != 0:
                ip2 = torch.empty(int(meta2[0]) + 1, dtype=torch.int64, device=dev)
                c2 = torch.empty(int(meta2[1]), dtype=torch.int32, device=dev)
                v2 = torch.empty(int(meta2[1]), dtype=torch.float32, device=dev)
            for t in (ip2, c2, v2):
                dist.broadcast(t, 0)
        n2, nnz2 = int(meta2[0]), int(meta2[1])
        # equal-nnz prefix split, padded to equal row counts for all_gather_into_tensor
        rows_per = (n2 + world - 1) // world
        q0, q1 = min(n2, rank * rows_per), min(n2, (rank + 1) * rows_per)
        p0, p1 = int(ip2[q0]), int(ip2[q1])
        M = gb.cuda.matrix_from_device_csr((ip2[q0:q1 + 1] - p0).contiguous(), c2[p0:p1].contiguous(), v2[p0:p1].contiguous(), q1 - q0, n2)
        x_t = values_torch(n2, 46, torch.float32, device=dev)
        x = gb.cuda.vector_from_torch(x_t)
        # the per-iteration exchange of the row-partitioned path: every rank's output slice is all-gathered straight into
        # the device buffer of the next input vector (zero-copy torch views of the library's arrays; NCCL over NVLink)
        x_next = gb.cuda.vector_from_torch(torch.zeros(n2, dtype=torch.float32, device=dev)) if world > 1 else None
        xn_vals = gb.cuda.vector_as_torch(x_next, sync=False)[0] if world > 1 else None
        zero_copy = world > 1 and n2 % world == 0
        gathered = torch.empty(rows_per * world, dtype=torch.float32, device=dev) if (world > 1 and not zero_copy) else None
        iters = max(20, args.steps * 4)

        def mxv_iter(xv):
            y = M.mxv(xv, sr).new()
            if world == 1:
                return y
            yv, _ = gb.cuda.vector_as_torch(y, sync=False)   # same stream as torch: no synchronisation needed
            if zero_copy:
                dist.all_gather_into_tensor(xn_vals, yv)
            else:
                pad = torch.zeros(rows_per, dtype=torch.float32, device=dev)
                pad[: q1 - q0] = yv
                dist.all_gather_into_tensor(gathered, pad)
                xn_vals.copy_(gathered[:n2])
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
        # the dominant kernel alone (library profile mode: CUDA events around each launch on the library stream)
        gb.cuda.set_option("profile", "1")
        gb.cuda.kernel_times(reset=True)
        for _ in range(5):
            y = M.mxv(x, sr).new()
        ktm = gb.cuda.kernel_times(reset=True)
        gb.cuda.set_option("profile", "0")
        kname = max((k for k in ktm if k in ("spmv_merge", "spmv_seg", "spmv_seg_hot")), key=lambda k: ktm[k][0], default=None)
        kernel_ms = max_over_ranks(ktm[kname][0] / ktm[kname][1]) if kname else ms_mxv
        # algorithmic bytes: nnz*(s_idx+s_val) + (nrows+1)*s_ptr + ncols*s_x + nrows*(s_y + 1 presence byte)
        bytes_local = (p1 - p0) * 8 + (q1 - q0 + 1) * 8 + n2 * 4 + (q1 - q0) * 5
        bytes_total = sum_over_ranks(float(bytes_local))
        gbs = bytes_total / (ms_mxv * 1e-3) / 1e9
        gbs_kernel = bytes_local / (kernel_ms * 1e-3) / 1e9
        kfull = {"spmv_merge": "spmv_merge_kernel", "spmv_seg": "spmv_seg_kernel", "spmv_seg_hot": "spmv_seg_kernel"}.get(kname, "?")
        mxv = {"workload": f"R-MAT scale-{scale} (0.57,0.19,0.19,0.05) plus_times fp32 A.mxv(x), x dense", "nnz": nnz2,
               "ms_per_iter": ms_mxv, "GB_per_s": gbs, "nnz_per_s": nnz2 / (ms_mxv * 1e-3),
               "roofline": {"bound": "hbm", "kernel": f"{kfull} (chosen by the library's timed trial)", "achieved": gbs_kernel, "peak": hbm,
                            "unit": "GB/s", "frac": gbs_kernel / hbm, "peak_source": pk_kind, "kernel_us": kernel_ms * 1e3,
                            "algorithmic_bytes": int(bytes_local), "frac_whole_call": gbs / world / hbm,
                            "traffic": (ncu_traffic(kfull)[1] if (scale == 22 and world == 1) else None),
                            "note": "gather-bound, not HBM-bound: one L1 wavefront per distinct 32 B sector a warp gathers (DESIGN.md section 3)"}}
        del M

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- the iterative workloads of BASELINE.json configs 3-5 as the reference notebooks write them (N == 1)
    workloads = None
    if world == 1 and not args.no_workloads:
        try:
            workloads = run_workloads(gb, torch, dev, scale)
        except Exception as exc:
            workloads = {"error": repr(exc)}

    # ---------------- CPU baseline (rank 0, N == 1 only): oracle port on the host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            hp, hc, hv = indptr.cpu().numpy(), cols.cpu().numpy().astype(np.int64), vals.cpu().numpy()
            rate, desc, threads, _, _ = cpu_mxm_sample(hp, hc, hv, n, budget_s=args.cpu_budget)
            cpu = {"value": rate, "unit": "nnz-out/s", "cores": threads, "kind": "port", "sample": desc,
                   "note": "SuiteSparse:GraphBLAS unavailable on this box -- baseline is the in-repo OpenMP SpGEMM (one pass, unsorted rows, hash / dense Gustavson accumulators)"}
        except Exception as exc:   # never lose the GPU numbers because the CPU leg failed
            cpu = {"value": None, "error": repr(exc)}

    line = {
        "metric": "mxm nnz-out/s (R-MAT plus_times fp32)", "value": value, "unit": "nnz-out/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"R-MAT scale-{scale} (0.45,0.15,0.15,0.25) ef16 seed42 A.mxm(A) plus_times fp32 [BASELINE configs[1], variant 2a]",
                   "n": n, "nnz_A": nnz, "nnz_C": int(nnz_c), "flops": int(flops), "parallelism": f"row-partition x{world}",
                   "l2": "inputs (0.5 GB) and outputs (>=GBs) exceed the 126 MB L2", "result_order": "jumbled (lazy sort); sort_ms reported"},
        "clocks": clk.summary(), "gpu_launches": launches_per_step,
        "phases_ms": {"symbolic": symbolic_ms, "numeric": numeric_ms, "sort_on_demand": sort_ms, "kernels": {k: v[0] for k, v in kt.items()}},
        "roofline": {"bound": "hbm", "kernel": "spgemm_numeric_*", "achieved": roof_ach, "peak": hbm, "unit": "GB/s",
                     "frac": (roof_ach / hbm) if roof_ach else None, "peak_source": pk_kind,
                     "traffic": (ncu_traffic()[0] if (scale == 22 and world == 1) else None),
                     "traffic_note": "ncu dram bytes of all numeric hash kernels of one A.mxm(A) (profiles/ncu_mxm22_r01.txt)",
                     "algorithmic_bytes": int(numeric_bytes), "bmin_frac_whole_step": bmin_bytes / (ms_dev * 1e-3) / 1e9 / hbm},
        "e2e": e2e, "mxv": mxv, "workloads": workloads, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=int(os.environ.get("GRB_BENCH_SCALE", "22")))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mxv", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-workloads", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
