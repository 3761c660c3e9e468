"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump: instructions executed and
stall samples per CUDA source line (top N)."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None
per = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("Function Name", "Kernel Name"): continue
    if r[0] != "" and r[0].isdigit():
        i_inst = hdr.index("Instructions Executed"); i_samp = hdr.index("# Samples"); i_thr = hdr.index("Thread Instructions Executed")
        key = (cur_file, int(r[0]))
        try:
            inst = int(r[i_inst]); samp = int(r[i_samp]); thr = int(r[i_thr])
        except ValueError:
            continue
        d = per.setdefault(key, [0, 0, 0, r[1]])
        d[0] += inst; d[1] += samp; d[2] += thr
tot_i = sum(v[0] for v in per.values()); tot_s = sum(v[1] for v in per.values())
print("total inst %d samples %d" % (tot_i, tot_s))
for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% samp %5.1f%% inst  thr/inst %4.1f  %s:%d  %s" % (100.0 * v[1] / max(tot_s, 1), 100.0 * v[0] / max(tot_i, 1), v[2] / max(v[0], 1), f, ln, v[3].strip()[:90]))
