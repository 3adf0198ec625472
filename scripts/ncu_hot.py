"""Top stalled SASS instructions of an `ncu --page source --csv` dump (gz): python scripts/ncu_hot.py file.csv.gz [N]"""
import csv, gzip, io, sys
rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and (r[ix['# Samples']] or '0').isdigit()]   # a multi-kernel dump repeats the header
tot = sum(int(r[ix['# Samples']] or 0) for r in body)
execd = sum(int(r[ix['Instructions Executed']] or 0) for r in body)
print("total samples", tot, "warp instructions executed", execd)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']] or 0))[:N]
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for i in sorted(top):
    r = body[i]
    s = int(r[ix['# Samples']] or 0)
    reasons = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {s:7d} {100*s/tot:5.1f}%  ex={r[ix['Instructions Executed']]:>9s} thr={r[ix['Avg. Threads Executed']]:>5s}  {r[ix['Source']].strip()[:70]:70s} {reasons}")
