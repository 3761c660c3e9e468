"""Comparison helpers shared by the parity tests (fp64 oracle vs device results)."""
from __future__ import annotations

import numpy as np

# SURVEY.md §8(d) tolerances, fp64 reference -> fp32 device
TOL32 = dict(w_abs=1e-5, w_rel=1e-4, mean_abs=1e-4, cov_rel=1e-3, logw_abs=2e-3, wnorm_rel=1e-3)
# fp64 device build vs fp64 oracle
TOL64 = dict(w_abs=1e-12, w_rel=1e-10, mean_abs=1e-10, cov_rel=1e-10, logw_abs=1e-9, wnorm_rel=1e-9)


def offsets(count):
    o = np.zeros(len(count) + 1, dtype=np.int64)
    np.cumsum(count, out=o[1:])
    return o


def compare_maps(count_a, mean_a, cov_a, w_a, count_b, mean_b, cov_b, w_b, tol, ordered=True):
    """Per-particle comparison. Returns dict with the particles whose STRUCTURE differs
    (different number of Gaussians or no one-to-one match) and the max errors over the rest."""
    N = len(count_a)
    oa, ob = offsets(count_a), offsets(count_b)
    bad = []
    max_w = max_mean = max_cov = 0.0
    for i in range(N):
        na, nb = int(count_a[i]), int(count_b[i])
        if na != nb:
            bad.append(i)
            continue
        if na == 0:
            continue
        ma, mb = mean_a[oa[i]:oa[i + 1]], mean_b[ob[i]:ob[i + 1]]
        ca, cb = cov_a[oa[i]:oa[i + 1]], cov_b[ob[i]:ob[i + 1]]
        wa, wb = w_a[oa[i]:oa[i + 1]], w_b[ob[i]:ob[i + 1]]
        perm = np.arange(na)
        if not ordered or np.abs(ma - mb).max() > tol["mean_abs"]:
            # match as sets: nearest (mean, weight) neighbour, must be one-to-one
            d = ((ma[:, None, :] - mb[None, :, :]) ** 2).sum(-1) + (wa[:, None] - wb[None, :]) ** 2
            perm = d.argmin(axis=1)
            if len(set(perm.tolist())) != na:
                bad.append(i)
                continue
        mb, cb, wb = mb[perm], cb[perm], wb[perm]
        ew = np.abs(wa - wb) - (tol["w_abs"] + tol["w_rel"] * np.abs(wb))
        em = np.abs(ma - mb).max()
        full = lambda c: np.stack([c[:, 0], c[:, 1], c[:, 1], c[:, 2]], 1)
        ec = np.linalg.norm(full(ca) - full(cb), axis=1) / np.maximum(np.linalg.norm(full(cb), axis=1), 1e-300)
        if ew.max() > 0 or em > tol["mean_abs"] or ec.max() > tol["cov_rel"]:
            bad.append(i)
            continue
        max_w = max(max_w, float(np.abs(wa - wb).max()))
        max_mean = max(max_mean, float(em))
        max_cov = max(max_cov, float(ec.max()))
    return dict(bad=bad, max_w=max_w, max_mean=max_mean, max_cov=max_cov)


def compare_weights(w_dev, w_ref, tol, mask=None):
    """Unnormalised particle weights: compare log w (abs) and normalised weights (rel)."""
    w_dev = np.asarray(w_dev, dtype=np.float64)
    w_ref = np.asarray(w_ref, dtype=np.float64)
    if mask is None:
        mask = np.ones(len(w_ref), dtype=bool)
    ok = mask & (w_ref > 0) & (w_dev > 0)
    dl = np.abs(np.log(w_dev[ok]) - np.log(w_ref[ok]))
    return dict(max_dlog=float(dl.max()) if dl.size else 0.0,
                n_bad=int((dl > tol["logw_abs"]).sum()) + int((mask & ((w_ref > 0) != (w_dev > 0))).sum()),
                idx_bad=np.nonzero(ok)[0][dl > tol["logw_abs"]] if dl.size else np.array([], dtype=int))
