"""Pins the restatement of the candidate-list births (oracle/phd_oracle_births.cpp) to the reference's own
RBPHDFilter::addBirthGaussians() (include/RBPHDFilter.hpp:1000-1080): against the golden sequences the compiled
reference wrote (tests/golden/make_golden_births.py) and, where oracle/_ref is present, against the reference live on
further random sequences.  Both are fp64 and evaluate the same closed forms, so lists, counters and the Gaussians that
become real agree to rounding."""
import os

import numpy as np
import pytest

import helpers
from oracle import binding

GOLDEN = ["births_rngbrg", "births_rngbrg_posecov", "births_vp"]


def _same(st, ref_n, ref_mean, ref_cov, ref_sup, ref_chk, tag):
    assert np.array_equal(st.n, ref_n), (tag, np.nonzero(st.n != ref_n)[0][:8])
    for i in range(st.N):
        k = int(ref_n[i])
        assert np.array_equal(st.support[i, :k], ref_sup[i, :k]) and np.array_equal(st.checks[i, :k], ref_chk[i, :k]), (tag, i)
        assert np.allclose(st.mean[i, :k], ref_mean[i, :k], rtol=1e-11, atol=1e-11), (tag, i)
        assert np.allclose(st.cov[i, :k], ref_cov[i, :k], rtol=1e-9, atol=1e-14), (tag, i)


def _same_adds(a, b, tag):
    assert np.array_equal(a[0], b[0]), tag
    for i in range(len(a[0])):
        k = int(a[0][i])
        assert np.allclose(a[1][i, :k], b[1][i, :k], rtol=1e-11, atol=1e-11), (tag, i)
        assert np.allclose(a[2][i, :k], b[2][i, :k], rtol=1e-9, atol=1e-14), (tag, i)


@pytest.mark.parametrize("case", GOLDEN)
def test_oracle_births_against_the_reference_golden_sequences(case):
    g = np.load(os.path.join(helpers.GOLDEN_DIR, case + ".npz"))
    dim, N, steps, seed, pose_cov, cand_cap, add_cap = (int(v) for v in g["meta"])
    model, seq = helpers.birth_scenario(dim, N, steps, seed, pose_cov=bool(pose_cov))
    st = binding.BirthState(N, dim, cap=cand_cap)
    n_real = 0
    for t, s in enumerate(seq):
        adds = binding.birth_candidates(model, helpers.BIRTH_CFG, st, s["pose"], s["Z"], s["mask"], s["nfov"], parent=s["parent"],
                                        pose_cov=s["pose_cov"], add_cap=add_cap)
        _same(st, g[f"s{t}_cand_n"], g[f"s{t}_cand_mean"], g[f"s{t}_cand_cov"], g[f"s{t}_cand_support"], g[f"s{t}_cand_checks"], (case, t))
        _same_adds(adds, (g[f"s{t}_add_n"], g[f"s{t}_add_mean"], g[f"s{t}_add_cov"]), (case, t))
        n_real += int(adds[0].sum())
    assert n_real > 100


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("bcfg", [helpers.BIRTH_CFG, dict(count_thr=1, check_thr=2, cur_count_thr=0, support_dist=1.0),
                                  dict(count_thr=5, check_thr=10, cur_count_thr=2, support_dist=2.0),     # the Victoria Park cfg
                                  dict(count_thr=2, check_thr=0, cur_count_thr=6, support_dist=3.0)],
                         ids=["default", "direct", "vpcfg", "always_few_in_view"])
def test_oracle_births_against_the_reference_live(dim, bcfg):
    if not binding.have_ref():
        pytest.skip("oracle/_ref/libphd_ref.so not built (needs /root/reference at build time)")
    for seed in (1, 2, 3):
        model, seq = helpers.birth_scenario(dim, 20, 8, 1000 * dim + seed, pose_cov=(dim == 2 and seed == 2))
        a = binding.BirthState(20, dim, cap=48)
        b = binding.BirthState(20, dim, cap=48)
        for t, s in enumerate(seq):
            kw = dict(parent=s["parent"], pose_cov=s["pose_cov"], add_cap=64)
            ra = binding.birth_candidates(model, bcfg, a, s["pose"], s["Z"], s["mask"], s["nfov"], which="oracle", **kw)
            rb = binding.birth_candidates(model, bcfg, b, s["pose"], s["Z"], s["mask"], s["nfov"], which="ref", **kw)
            _same(a, b.n, b.mean, b.cov, b.support, b.checks, (seed, t))
            _same_adds(ra, rb, (seed, t))
