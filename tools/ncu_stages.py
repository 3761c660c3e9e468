"""Stage shares of phd_update_kernel from an `ncu --page source --csv --print-source cuda,sass` dump; the stage
boundaries are found by their marker comments in phd_kernels.cuh.  usage: python tools/ncu_stages.py dump.csv [n_particles]"""
import csv, sys, os, re, collections
path = sys.argv[1]; NP = float(sys.argv[2]) if len(sys.argv) > 2 else 8000.0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "rfs-slam_b200", "csrc", "phd_kernels.cuh")).read().split("\n")
marks = [("helpers", "^namespace rfsb200"), ("M1", r"int merge_clustered\(T\* cur"), ("M2a", "---- M2: candidate"), ("M2b", r"^  int npass = 0;"),
         ("M3", "---- M3: clusters"), ("M4", "---- M4: one lane"), ("M5", "---- M5: commit"), ("mf_helpers", "^// S5 helpers"),
         ("epilogue", "^// End of a step"), ("tables", r"^phd_update_kernel\("), ("S0", r"^  while \(pi < p.N\)"), ("S1a", "// S1a: every component"),
         ("S1b", "// S1b: the queued"), ("S2-4", "-- S2: per-measurement"), ("S5", "-- S5: multi-feature"), ("S6call", "-- S6: merge"),
         ("S7", "-- S7: prune"), ("end", r"^  step_epilogue<T>\(p, lane")]
bounds = []
for name, pat in marks:
    for i, l in enumerate(src):
        if re.search(pat, l):
            bounds.append((i + 1, name)); break
bounds.sort()
def stage_of(f, ln):
    if "phd_kernels" not in f: return "inl:" + f[:24]
    cur = "pre"
    for b, name in bounds:
        if ln >= b: cur = name
    return cur
rows = list(csv.reader(open(path)))
cur = None; hdr = None
inst = collections.Counter(); samp = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    try:
        ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        a = int(r[ii] or 0); s = int(r[isamp] or 0)
    except (ValueError, IndexError):
        continue
    st = stage_of(cur or "", int(r[0]))
    inst[st] += a; samp[st] += s
ti = sum(inst.values()); ts = sum(samp.values())
print("total %.0f inst/particle, %d samples" % (ti / NP, ts))
for k in sorted(inst, key=lambda k: -inst[k]):
    print("%-28s %7.1f i/p %5.1f%% inst %5.1f%% samples" % (k, inst[k] / NP, 100.0 * inst[k] / ti, 100.0 * samp[k] / max(ts, 1)))
