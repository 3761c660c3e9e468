"""Per-source-line table (in line order) from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`:
instructions executed per particle, share of instructions / stall samples, active threads per instruction.
usage: python tools/ncu_bylines.py dump.csv [n_particles] [file-substring] [lo] [hi]"""
import csv, sys, collections
path = sys.argv[1]
NP = float(sys.argv[2]) if len(sys.argv) > 2 else 8000.0
sub = sys.argv[3] if len(sys.argv) > 3 else ""
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi = int(sys.argv[5]) if len(sys.argv) > 5 else 10**9
rows = list(csv.reader(open(path)))
cur = None; hdr = None
per = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    try:
        ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); it = hdr.index("Thread Instructions Executed")
        inst = int(r[ii] or 0); s = int(r[isamp] or 0); t = int(r[it] or 0)
    except (ValueError, IndexError):
        continue
    d = per.setdefault((cur, int(r[0])), [0, 0, 0, r[1]])
    d[0] += inst; d[1] += s; d[2] += t
tot = sum(v[0] for v in per.values()); ts = sum(v[1] for v in per.values())
print("total inst %d (%.0f per particle)  samples %d" % (tot, tot / NP, ts))
acc_i = acc_s = 0
for (f, ln), v in sorted(per.items()):
    if sub not in f or ln < lo or ln > hi or (v[0] == 0 and v[1] == 0): continue
    acc_i += v[0]; acc_s += v[1]
    print("%-22s %5d %7.1f i/p %5.2f%%i %5.2f%%s thr %4.1f | %s" % (f[:22], ln, v[0] / NP, 100.0 * v[0] / tot, 100.0 * v[1] / max(ts, 1), v[2] / max(v[0], 1), v[3].strip()[:100]))
print("selected: %.1f i/p  %.2f%% inst  %.2f%% samples" % (acc_i / NP, 100.0 * acc_i / tot, 100.0 * acc_s / max(ts, 1)))
