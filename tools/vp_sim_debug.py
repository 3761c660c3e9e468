"""Run the unchanged Victoria Park driver on the reference header and on the drop-in, report the first divergence."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
def run(binary, wd, env=None):
    os.makedirs(wd, exist_ok=True)
    if not os.path.exists(wd + "/vpdata"): os.symlink(REFDIR + "/vpdata", wd + "/vpdata")
    e = dict(os.environ); e.update(env or {})
    r = subprocess.run([REFDIR + "/" + binary, "-c", REFDIR + "/rbphdslam_VictoriaPark.xml", "-s", "1"], cwd=wd, env=e, capture_output=True, text=True)
    print(binary, "rc", r.returncode, r.stdout[-300:], r.stderr[-500:])
    return np.loadtxt(wd + "/vpout/particlePose.dat"), np.loadtxt(wd + "/vpout/landmarkEst.dat")
prec = sys.argv[1] if len(sys.argv) > 1 else "64"
a, la = run("rbphdslam_VictoriaPark_ref", "/tmp/vp_ref")
b, lb = run("rbphdslam_VictoriaPark_b200", "/tmp/vp_b", {"RFSB200_PRECISION": prec})
print(a.shape, b.shape, la.shape, lb.shape)
n = min(len(a), len(b))
d = np.abs(a[:n, 2:6] - b[:n, 2:6])
bad = np.nonzero((d[:, :3] > 2e-3).any(1) | (d[:, 3] > 2e-3 + 1e-2 * np.abs(a[:n, 5])))[0]
print("rows differing:", len(bad), "of", n)
if len(bad):
    k = bad[0]
    print("first differing row", k, "t", a[k, 0], "ref", a[k], "b200", b[k])
    t0 = a[k, 0]
    ts = np.unique(a[:, 0]); i0 = np.searchsorted(ts, t0)
    for t in ts[max(0, i0 - 2):i0 + 1]:
        ra, rb = a[a[:, 0] == t], b[b[:, 0] == t]
        print(" t", t, "max dpose", np.abs(ra[:, 2:5] - rb[:, 2:5]).max(), "max dw", np.abs(ra[:, 5] - rb[:, 5]).max(), "w ref", ra[:3, 5], "w b200", rb[:3, 5])
        ma, mb = la[la[:, 0] == t], lb[lb[:, 0] == t]
        print("   landmarkEst rows ref", len(ma), "b200", len(mb), "best particle", ma[0, 1] if len(ma) else None, mb[0, 1] if len(mb) else None)
        if len(ma) and len(mb) and len(ma) == len(mb):
            print("   max d lm", np.abs(np.sort(ma[:, 2:], 0) - np.sort(mb[:, 2:], 0)).max())
# landmark count over time
ta = np.unique(la[:, 0])
for t in ta[:12]:
    print(" t", t, "n lm ref", (la[:, 0] == t).sum(), "b200", (lb[:, 0] == t).sum())
