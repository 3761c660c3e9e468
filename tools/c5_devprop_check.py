"""The unchanged Victoria Park driver (100 particles) with device-side propagation vs the reference: deviation of the
weighted-mean trajectory (different random streams, so this is a statistical comparison)."""
import os, subprocess, sys
import numpy as np
R = "/root/repo/oracle/_ref"
def run(binary, wd, env):
    os.makedirs(wd, exist_ok=True)
    if not os.path.exists(wd + "/vpdata"): os.symlink(R + "/vpdata", wd + "/vpdata")
    e = dict(os.environ); e.update(env)
    r = subprocess.run([R + "/" + binary, "-c", R + "/rbphdslam_VictoriaPark.xml", "-s", "1"], cwd=wd, env=e, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    return np.loadtxt(wd + "/vpout/particlePose.dat")
def track(pp):
    out = []
    for tk in np.unique(pp[:, 0])[::10]:
        r = pp[pp[:, 0] == tk]; w = r[:, 5] / r[:, 5].sum(); out.append((r[:, 2:4] * w[:, None]).sum(0))
    return np.array(out)
ref = track(run("rbphdslam_VictoriaPark_ref", "/tmp/dp_ref", {}))
for seed in (1, 2, 3):
    dev = track(run("rbphdslam_VictoriaPark_b200", "/tmp/dp_dev%d" % seed, {"RFSB200_DEVICE_PROPAGATE": "1", "RFSB200_SEED": str(seed)}))
    d = np.linalg.norm(dev - ref, axis=1)
    print("seed", seed, "max deviation of the weighted-mean track %.3f m, mean %.3f m, path length %.1f m" % (d.max(), d.mean(), np.linalg.norm(np.diff(ref, axis=0), axis=1).sum()))
