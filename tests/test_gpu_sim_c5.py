"""BASELINE config C5: the reference's Victoria Park driver, src/rbphdslam_VictoriaPark.cpp UNCHANGED
(MotionModel_Ackerman2d + MeasurementModel_VictoriaPark + KalmanFilter_VictoriaPark), compiled once against
the reference's RBPHDFilter.hpp (CPU) and once against the drop-in header + librfsb200 (GPU) by
`make -C oracle simvp`; both run the shipped artificial-clutter configuration on the head of the Victoria
Park dataset (2 500 sensor messages, 100 particles, -s 1).  Particle propagation and the resampling draw are host code in both
builds and the candidate-list births run on the device in fp64, so with the fp64 device build the two runs log the
same particle poses, weights and best-particle maps until an exact tie in Gaussian weights is ordered differently
by the reference's unstable std::sort (Q9), and the same trajectory estimate afterwards."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
NEEDED = ("rbphdslam_VictoriaPark_ref", "rbphdslam_VictoriaPark_b200", "rbphdslam_VictoriaPark.xml", "vpdata/LASER.txt")


def _run(binary, workdir, env=None):
    os.makedirs(workdir, exist_ok=True)
    link = os.path.join(workdir, "vpdata")
    if not os.path.exists(link):
        os.symlink(os.path.join(REFDIR, "vpdata"), link)
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([os.path.join(REFDIR, binary), "-c", os.path.join(REFDIR, "rbphdslam_VictoriaPark.xml"), "-s", "1"],
                       cwd=workdir, env=e, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    pp = np.loadtxt(os.path.join(workdir, "vpout", "particlePose.dat"))      # t id x y theta w
    lm = np.loadtxt(os.path.join(workdir, "vpout", "landmarkEst.dat"))       # t pid x y d ... w
    return pp, lm


def test_unchanged_victoria_park_driver_runs_on_the_dropin(cuda_required, tmp_path):
    if not all(os.path.exists(os.path.join(REFDIR, f)) for f in NEEDED):
        pytest.skip("oracle/_ref/rbphdslam_VictoriaPark_{ref,b200} not built (needs /root/reference at build time)")
    pp_ref, lm_ref = _run("rbphdslam_VictoriaPark_ref", str(tmp_path / "ref"))
    pp64, lm64 = _run("rbphdslam_VictoriaPark_b200", str(tmp_path / "b64"), {"RFSB200_PRECISION": "64"})
    assert pp64.shape == pp_ref.shape and pp_ref.shape[0] > 10000
    # Same RNG stream, same resampling decisions, same candidate-list births: identical poses / weights / best-particle
    # maps to print precision for as long as no two Gaussians of a particle have EXACTLY the same weight.  The
    # sensing-limit heuristic clips weights to exactly 1.0 (include/RBPHDFilter.hpp:699-701), and the order the reference's
    # std::sort leaves such ties in is implementation-defined (Q9; the device uses weight desc, position asc), which
    # changes the eval-point choice of importanceWeighting.  On this dataset the first such tie that matters comes after
    # 60 updates; up to there the two runs must agree exactly, afterwards as estimates of the same trajectory.
    early = pp_ref[:, 0] < 13.9
    assert early.sum() >= 60 * 100
    assert np.allclose(pp64[early, :5], pp_ref[early, :5], atol=2e-3)
    assert np.allclose(pp64[early, 5], pp_ref[early, 5], rtol=1e-2, atol=2e-3)

    def canon(a):
        a = a[a[:, 0] < 13.9]
        return a[np.lexsort((np.round(a[:, 3], 2), np.round(a[:, 2], 2), a[:, 0]))]
    a, b = canon(lm64), canon(lm_ref)
    assert a.shape == b.shape
    close = np.isclose(a, b, atol=2e-3).all(axis=1)
    assert close.mean() > 0.995, f"{(~close).sum()} of {len(close)} logged Gaussians differ"

    def mean_track(pp):
        out = []
        for tk in np.unique(pp[:, 0])[::10]:
            r = pp[pp[:, 0] == tk]
            w = r[:, 5] / r[:, 5].sum()
            out.append((r[:, 2:4] * w[:, None]).sum(0))
        return np.array(out)
    ref_track = mean_track(pp_ref)
    assert np.linalg.norm(ref_track[-1] - ref_track[0]) > 10.0      # the vehicle has driven off
    d64 = np.linalg.norm(mean_track(pp64) - ref_track, axis=1)
    assert d64.max() < 1.0, d64.max()
    # map size of the best particle over time stays with the reference's
    n_ref = np.array([(lm_ref[:, 0] == t).sum() for t in np.unique(lm_ref[:, 0])])
    n_64 = np.array([(lm64[:, 0] == t).sum() for t in np.unique(lm64[:, 0])])
    assert len(n_ref) == len(n_64) and np.abs(n_ref - n_64).max() <= 6
    # the candidate lists of addBirthGaussians live on the device (rfsb200_birth_candidates); kept on the host and
    # evaluated with the live plugin objects instead (RFSB200_HOST_BIRTHS=1), the run logs the same numbers
    pph, lmh = _run("rbphdslam_VictoriaPark_b200", str(tmp_path / "b64h"), {"RFSB200_PRECISION": "64", "RFSB200_HOST_BIRTHS": "1"})
    assert pph.shape == pp64.shape and lmh.shape == lm64.shape
    assert np.allclose(pph[early], pp64[early], rtol=1e-9, atol=1e-9)
    assert np.allclose(pph[:, :5], pp64[:, :5], atol=2e-3) and np.allclose(lmh, lm64, atol=2e-3)
    # fp32 product build: same trajectory estimate
    pp32, lm32 = _run("rbphdslam_VictoriaPark_b200", str(tmp_path / "b32"), {"RFSB200_PRECISION": "32"})
    assert pp32.shape == pp_ref.shape
    d32 = np.linalg.norm(mean_track(pp32) - ref_track, axis=1)
    assert d32.max() < 1.0, d32.max()
