"""Randomised sequence sweep: the drop-in header (include/rfs_b200/RBPHDFilter.hpp over the C ABI) against the reference
class it replaces, both driven through their public API by oracle/seq_harness.cpp on random scenarios — particle counts,
sequence lengths, landmark densities, measurement counts, both weightings, with and without resampling; births, landmark
process noise, the same drand48 stream and an empty measurement set are part of every sequence.  fp64 device build:
identical resampling decisions, identical map structure, weights to 1e-8.
usage: fuzz_sequences.py [n_cases] [seed]
On a GPU box this runs the sm_100a build.  Without a GPU, bind the ABI symbols to the host interpreter of tests/simt
(test infrastructure):  LD_PRELOAD=tests/simt/_build/librfsb200_simt.so SIMT_SM_COUNT=2 python tools/fuzz_sequences.py 200 1"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rfs_slam_b200  # noqa
from oracle import binding as ob
import helpers
import test_gpu_dropin as td

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
if not ob.have_seq():
    print("oracle/_ref/libseq_{ref,b200}.so not built (needs /root/reference at build time)")
    sys.exit(2)
bad = 0
for case in range(n_cases):
    kw = dict(N=int(rng.integers(4, 48)), K=int(rng.integers(3, 14)), n_lmk=int(rng.integers(10, 90)),
              nZ_max=int(rng.integers(4, 24)), seed=int(rng.integers(1, 1 << 30)), sc=int(rng.random() < 0.5))
    resample = bool(rng.random() < 0.5)
    md, fc, poses, Z, nZ = td._scenario(**kw)
    run = dict(pose_cov=[3e-5, 0, 0, 3e-5, 0, 3e-5], Q_lmk=[1e-5, 0, 0, 1e-5],
               # "resample (almost) always": 0.98 N rather than N itself — with N the gate N_eff > N is decided by the
               # last bit of 1 / sum w^2 whenever all weights are equal (e.g. every map still empty), which a 1e-13
               # relative difference in the common weight flips
               neff_threshold=(0.98 * float(poses.shape[1]) if resample else 0.0), seed48=int(rng.integers(1, 1000)))
    if rng.random() < 0.5:   # the candidate-list form of addBirthGaussians (on the device in the drop-in) with random thresholds
        run["births"] = dict(count_thr=int(rng.integers(2, 5)), check_thr=int(rng.integers(0, 6)), cur_count_thr=int(rng.integers(0, 4)),
                             support_dist=float(rng.uniform(0.5, 3.0)))
    ref, nres_ref, trace_ref = ob.run_sequence("ref", poses, Z, nZ, md, fc, **run)
    got, nres, trace = ob.run_sequence("b200", poses, Z, nZ, md, fc, precision=64, **run)
    ok = nres == nres_ref and np.array_equal(trace, trace_ref) and np.array_equal(got.count, ref.count)
    if ok:
        r = helpers.compare_maps(got.count, got.mean, got.cov, got.w, ref.count, ref.mean, ref.cov, ref.w, helpers.TOL64, ordered=False)
        ok = (not r["bad"]) and np.allclose(got.weight, ref.weight, rtol=1e-8, atol=0)
    if not ok:
        bad += 1
        print(f"CASE {case} FAILED: kw={kw} resample={resample} seed48={run['seed48']} births={run.get('births')} nres {nres} / {nres_ref}")
print(f"{n_cases} random sequences, {bad} with differences")
sys.exit(1 if bad else 0)
