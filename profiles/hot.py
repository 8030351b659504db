#!/usr/bin/env python
"""Hottest source lines / SASS of an ncu source-page CSV (gpurun_out/src_*.csv.gz): python profiles/hot.py FILE [N] [sass]"""
import csv, gzip, io, sys
f = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = gzip.open(f, "rt").read() if f.endswith(".gz") else open(f).read()
rows = list(csv.reader(io.StringIO(raw)))
if rows[0][0] == "Kernel Name": print(rows[0][1]); rows = rows[1:]
hdr = rows[0]
def col(name):
    for i, h in enumerate(hdr):
        if h.strip() == name: return i
    return None
ci = {k: col(k) for k in ["#", "Address", "Source", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (Not-issued Samples)", "# Samples", "Instructions Executed", "Thread Instructions Executed", "Predicated-On Thread Instructions Executed"]}
samp = ci["Warp Stall Sampling (All Samples)"] if ci["Warp Stall Sampling (All Samples)"] is not None else ci["# Samples"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
data = rows[1:]
tot = sum(num(r[samp]) for r in data if len(r) > samp)
print("columns:", [h for h in hdr][:12], "... total samples", tot, "rows", len(data))
ie = ci["Instructions Executed"]
print("total warp instructions", sum(num(r[ie]) for r in data if ie is not None and len(r) > ie))
order = sorted(range(len(data)), key=lambda k: -num(data[k][samp]) if len(data[k]) > samp else 0)
for k in order[:n]:
    r = data[k]
    st = sorted(((num(r[i]), hdr[i][6:]) for i in stall_cols if num(r[i]) > 0), reverse=True)[:3]
    print("%5d %6.2f%% exec=%-9s %-70s %s" % (k, 100 * num(r[samp]) / max(tot, 1), r[ie] if ie is not None else "", r[ci["Source"]][:70], " ".join("%s:%d" % (b, a) for a, b in st)))
