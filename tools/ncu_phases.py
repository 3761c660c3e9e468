"""Aggregate an ncu source-page CSV by line ranges of phd_kernels.cuh (phases). usage: ncu_phases.py csv name:lo-hi ..."""
import csv, sys, collections
path = sys.argv[1]
ranges = []
for a in sys.argv[2:]:
    n, r = a.split(":"); lo, hi = r.split("-"); ranges.append((n, int(lo), int(hi)))
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None
agg = collections.OrderedDict((n, [0, 0]) for n, _, _ in ranges); agg["other"] = [0, 0]
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    try:
        inst = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
    except ValueError: continue
    ln = int(r[0]); key = "other"
    if cur_file == "phd_kernels.cuh":
        for n, lo, hi in ranges:
            if lo <= ln <= hi: key = n; break
    else:
        key = "other:" + cur_file
        agg.setdefault(key, [0, 0])
    agg[key][0] += inst; agg[key][1] += samp
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
for k, v in agg.items():
    print("%-28s inst %5.1f%%  samples %5.1f%%" % (k, 100.0 * v[0] / ti, 100.0 * v[1] / ts))
