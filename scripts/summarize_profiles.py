"""Turn the raw ncu CSVs brought back in gpurun_out/ into the committed summaries under profiles/."""
import collections, csv, json, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
P.mkdir(exist_ok=True)
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r01"


def read(path):
    rows = list(csv.reader(open(path, errors="replace")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            return r, rows[i + 1:]
    return None, []


def short(name):
    n = name.split("(")[0].replace("void ", "")
    return n[:110]


def launch_list(src, dst, title):
    hdr, rows = read(src)
    if not hdr:
        return
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    ours = {k: v for k, v in agg.items() if not k.startswith(("at::", "native::", "at_cuda", "cuda::", "CUB"))}
    with open(dst, "w") as f:
        f.write(f"# {title}\n# source: gpurun_out/{src.name} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n")
        f.write(f"# total captured device time: {tot/1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")
        f.write(f"{'share':>7} {'total_ms':>10} {'launches':>8}  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{t/tot*100:6.2f}% {t/1e6:10.3f} {n:8d}  {k}\n")
        ot = sum(v[1] for v in ours.values())
        f.write(f"# library kernels only: {ot/1e6:.2f} ms ({ot/tot*100:.1f}% of captured time; the rest is torch generating the synthetic input)\n")


def per_kernel_metrics(src, dst, title, out_json_key, traffic):
    hdr, rows = read(src)
    if not hdr:
        return
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = collections.OrderedDict()
    for r in rows:
        if len(r) <= vi:
            continue
        d = launches.setdefault(r[ii], {"name": short(r[ki])})
        try:
            d[r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            pass
    unit = {}
    ui = hdr.index("Metric Unit")
    for r in rows:
        if len(r) > ui:
            unit[r[mi]] = r[ui]
    with open(dst, "w") as f:
        f.write(f"# {title}\n# source: gpurun_out/{src.name}\n")
        f.write(f"{'id':>4} {'time_us':>10} {'dram_rd_MB':>11} {'dram_wr_MB':>11} {'TB/s':>8} {'l1tex_wf%':>9} {'warps%':>7}  kernel\n")
        agg = collections.OrderedDict()
        for i, d in launches.items():
            def val(k, scale_map):
                v = d.get(k)
                if v is None:
                    return 0.0
                return v * scale_map.get(unit.get(k, ""), 1.0)
            t = val("gpu__time_duration.sum", {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3})
            rd = val("dram__bytes_read.sum", {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3})
            wr = val("dram__bytes_write.sum", {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3})
            wf = d.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 0.0)
            wa = d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
            f.write(f"{i:>4} {t:10.1f} {rd:11.1f} {wr:11.1f} {(rd+wr)/max(t,1e-9):8.2f} {wf:9.1f} {wa:7.1f}  {d['name']}\n")
            a = agg.setdefault(d["name"], [0, 0.0, 0.0])
            a[0] += 1; a[1] += t; a[2] += rd + wr
        f.write("# per kernel name: launches, total time (us), total DRAM traffic (MB)\n")
        for k, (n, t, b) in agg.items():
            f.write(f"#   x{n:<3d} {t:12.1f} us {b:12.1f} MB  {k}\n")
    traffic[out_json_key] = {k: {"launches": n, "time_us": t, "dram_MB": b} for k, (n, t, b) in agg.items()}


traffic = {}
if (G / f"launches_bench_{ROUND}.csv").exists():
    launch_list(G / f"launches_bench_{ROUND}.csv", P / f"launches_bench_{ROUND}.txt", "launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu` (scale 22)")
if (G / f"ncu_mxm22_{ROUND}.csv").exists():
    per_kernel_metrics(G / f"ncu_mxm22_{ROUND}.csv", P / f"ncu_mxm22_{ROUND}.txt", "A.mxm(A) plus_times fp32, R-MAT scale 22 variant 2a: per-launch time and DRAM traffic (2 calls)", "mxm22", traffic)
if (G / f"ncu_mxv22_{ROUND}.csv").exists():
    per_kernel_metrics(G / f"ncu_mxv22_{ROUND}.csv", P / f"ncu_mxv22_{ROUND}.txt", "A.mxv(x) plus_times fp32, R-MAT scale 22 Graph500 skew: per-launch time and DRAM traffic (4 calls)", "mxv22", traffic)
(P / f"traffic_{ROUND}.json").write_text(json.dumps(traffic, indent=1))
print("wrote", sorted(p.name for p in P.iterdir()))
