"""Golden fixtures of the candidate-list births, written by the REFERENCE ITSELF: its own
RBPHDFilter::addBirthGaussians() (include/RBPHDFilter.hpp:1000-1080) driven on injected state through
oracle/_ref/libphd_ref.so (oracle/ref_births.hpp).  Run in the build container only (`make -C oracle ref` first):
    python tests/golden/make_golden_births.py
-> tests/golden/births_rngbrg.npz, births_rngbrg_posecov.npz, births_vp.npz (inputs and outputs of every step of a
sequence; the candidate lists are carried from step to step by the reference)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import rfs_slam_b200  # noqa: E402,F401
from oracle import binding as ob  # noqa: E402
import helpers  # noqa: E402

CASES = {"births_rngbrg": dict(dim=2, N=24, steps=9, seed=201), "births_rngbrg_posecov": dict(dim=2, N=16, steps=7, seed=202, pose_cov=True),
         "births_vp": dict(dim=3, N=24, steps=9, seed=203)}
CAND_CAP, ADD_CAP = 32, 48

if __name__ == "__main__":
    for name, kw in CASES.items():
        model, steps = helpers.birth_scenario(**kw)
        st = ob.BirthState(kw["N"], kw["dim"], cap=CAND_CAP)
        out = dict(meta=np.array([kw["dim"], kw["N"], kw["steps"], kw["seed"], int(kw.get("pose_cov", False)), CAND_CAP, ADD_CAP]))
        tot_add = tot_sup = 0
        for t, s in enumerate(steps):
            add_n, add_mean, add_cov = ob.birth_candidates(model, helpers.BIRTH_CFG, st, s["pose"], s["Z"], s["mask"], s["nfov"],
                                                           parent=s["parent"], pose_cov=s["pose_cov"], which="ref", add_cap=ADD_CAP)
            assert add_n.max() <= ADD_CAP and st.n.max() <= CAND_CAP
            out.update({f"s{t}_cand_n": st.n.copy(), f"s{t}_cand_mean": st.mean.copy(), f"s{t}_cand_cov": st.cov.copy(),
                        f"s{t}_cand_support": st.support.copy(), f"s{t}_cand_checks": st.checks.copy(),
                        f"s{t}_add_n": add_n, f"s{t}_add_mean": add_mean, f"s{t}_add_cov": add_cov})
            tot_add += int(add_n.sum())
            tot_sup += int((st.support > 1).sum())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "real Gaussians", tot_add, "supported candidate-steps", tot_sup, "final list lengths", st.n.sum())
