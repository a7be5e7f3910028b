"""The Monte Carlo agreement gate of SURVEY 8d / BASELINE.json north_star: two independent runs of the same scene agree
"per pixel within 3 sigma of the combined photon noise".

What sigma is.  The reference's own per-pixel error (src/ARTES.f90:3490-3493: sqrt(sum W^2 / n - (sum W / n)^2) * sqrt(n))
is the scatter of the deposits about their mean.  It leaves out the Poisson noise of the NUMBER of deposits and, above
all, the correlation between the many peel-off deposits one packet makes into the same pixel (one per scattering; ~40 per
packet in the template atmosphere).  Two runs of the ORACLE ITSELF with different seeds disagree by 3.1 of those sigmas rms
in Stokes I on the template atmosphere (34 % of the pixels beyond "3 sigma"; test_reference_sigma_underestimates_the_noise),
so that estimate cannot gate anything.  The photon noise used here is measured: every run is made of K statistically
independent batches and the variance of a pixel is K times the variance of its batch sums.  The thresholds are SURVEY 8d's:
pixels with at least 30 deposits per batch, at most 1 % of them beyond 3 sigma, none beyond 5 sigma -- for I, Q, U and for
the degree of polarisation P = sqrt(Q^2 + U^2) / I, whose sigma follows the reference's propagation (:995-1002, :3506-3516)
fed with the measured sigmas.  Two statistical footnotes: (i) z is a ratio with a variance estimated from K batches, i.e.
Student-t with ~2(K-1) degrees of freedom (P(|t| > 3) = 0.4 % at K = 32), and on a detector with n valid pixels the count
beyond 3 sigma is binomial -- "at most 1 %" of n = 150 pixels is one pixel, which two correct runs exceed 12 % of the time;
the gate therefore allows max(1 % of n, the 99.9 % quantile of that binomial).  (ii) P is a ratio of noisy numbers and only
Gaussian where the polarised flux is well measured, so P is gated on pixels with sqrt(Q^2+U^2) above 3 of its sigmas.
"""
import math

import numpy as np

MIN_PER_BATCH = 30


def _stack(batches):
    """batches: K arrays det[l, stokes, ...pixels] -> sums [K, 4, ...], counts [4, ...] (totals)."""
    s = np.stack([b[0] for b in batches])
    n = np.sum([b[2] for b in batches], axis=0)
    return s, n


def ref_sigma(det):
    """:3490-3493 on det[l, stokes, ...]."""
    n = det[2]
    with np.errstate(all="ignore"):
        v = np.where(n > 0, det[1] / n - (det[0] / n) ** 2, 0.0)
    return np.where((n > 0) & (v > 0), np.sqrt(np.abs(v)) * np.sqrt(n), 0.0)


def pol_sigma(i_, q, u, si, sq, su, reference_formula=True):
    """sigma of P = sqrt(q^2 + u^2) / i.  reference_formula: dpol as in :995-1002 / :3506-3516 (it carries a factor 1/2
    under the root); otherwise plain first-order propagation."""
    pol2 = q * q + u * u
    with np.errstate(all="ignore"):
        pol = np.sqrt(pol2)
        dpol = np.sqrt(((q * sq) ** 2 + (u * su) ** 2) / ((2.0 if reference_formula else 1.0) * pol2))
        s = (pol / i_) * np.sqrt((dpol / pol) ** 2 + (si / i_) ** 2)
    return np.where((pol2 > 0) & (i_ > 0), s, 0.0)


def z_report(batches_a, batches_b, min_per_batch=MIN_PER_BATCH):
    """z scores of two runs made of K independent batches each.  Returns {name: (n_valid, frac > 3, max, rms)} for
    I, Q, U, P (P with the reference's propagation) and P1 (first-order propagation)."""
    A_, na = _stack(batches_a)
    B_, nb = _stack(batches_b)
    Ka, Kb = A_.shape[0], B_.shape[0]
    va, vb = Ka * A_.var(axis=0, ddof=1), Kb * B_.var(axis=0, ddof=1)      # variance of the run totals
    ta, tb = A_.sum(axis=0), B_.sum(axis=0)
    rep = {}
    valid = {}
    for k, nm in enumerate("IQU"):
        m = (na[k] >= min_per_batch * Ka) & (nb[k] >= min_per_batch * Kb) & (va[k] + vb[k] > 0)
        valid[nm] = m
        z = np.abs(ta[k] - tb[k])[m] / np.sqrt((va[k] + vb[k])[m])
        rep[nm] = _summ(z)
    m = valid["I"] & valid["Q"] & valid["U"]
    with np.errstate(all="ignore"):
        pa = np.sqrt(ta[1] ** 2 + ta[2] ** 2) / ta[0]
        pb = np.sqrt(tb[1] ** 2 + tb[2] ** 2) / tb[0]
    with np.errstate(all="ignore"):      # polarised flux measured to better than 3 sigma on both sides
        snr_a = np.sqrt(ta[1] ** 2 + ta[2] ** 2) / np.sqrt((ta[1] ** 2 * va[1] + ta[2] ** 2 * va[2]) / (ta[1] ** 2 + ta[2] ** 2))
        snr_b = np.sqrt(tb[1] ** 2 + tb[2] ** 2) / np.sqrt((tb[1] ** 2 * vb[1] + tb[2] ** 2 * vb[2]) / (tb[1] ** 2 + tb[2] ** 2))
    well = np.nan_to_num(snr_a) > 3.0
    well &= np.nan_to_num(snr_b) > 3.0
    for nm, ref in (("P", True), ("P1", False)):
        sa = pol_sigma(ta[0], ta[1], ta[2], np.sqrt(va[0]), np.sqrt(va[1]), np.sqrt(va[2]), ref)
        sb = pol_sigma(tb[0], tb[1], tb[2], np.sqrt(vb[0]), np.sqrt(vb[1]), np.sqrt(vb[2]), ref)
        mm = m & well & (sa > 0) & (sb > 0)
        rep[nm] = _summ((np.abs(pa - pb) / np.sqrt(sa ** 2 + sb ** 2))[mm])
    rep["dof"] = (Ka - 1) + (Kb - 1)
    return rep


def z_report_ref_sigma(batches_a, batches_b, min_count=30):
    """The same comparison with the reference's own sigma (run totals, :3490-3493) -- reported, not gated (module docstring)."""
    ta, tb = np.sum(batches_a, axis=0), np.sum(batches_b, axis=0)
    sa, sb = ref_sigma(ta), ref_sigma(tb)
    rep = {}
    for k, nm in enumerate("IQU"):
        m = (ta[2][k] >= min_count) & (tb[2][k] >= min_count) & (sa[k] > 0) & (sb[k] > 0)
        rep[nm] = _summ(np.abs(ta[0][k] - tb[0][k])[m] / np.sqrt(sa[k][m] ** 2 + sb[k][m] ** 2))
    return rep


def _summ(z):
    z = np.asarray(z, dtype=float).ravel()
    if z.size == 0:
        return (0, 0.0, 0.0, 0.0)
    return (int(z.size), float((z > 3).mean()), float(z.max()), float(math.sqrt((z ** 2).mean())))


def allowed_beyond_3(n, dof):
    """max(1 % of n, 99.9 % quantile of Binomial(n, P(|t_dof| > 3)))."""
    from scipy import stats
    p3 = 2.0 * stats.t.sf(3.0, dof)
    return int(max(math.floor(0.01 * n), stats.binom.ppf(0.999, n, p3)))


def assert_gate(rep, names=("I", "Q", "U", "P"), min_valid=1, what="", min_valid_p=None):
    """at most 1 % of the valid pixels beyond 3 sigma, none beyond 5 sigma (and an rms near 1)."""
    dof = rep.get("dof", 62)
    for nm in names:
        n, f3, zmax, rms = rep[nm]
        need = min_valid if nm in "IQU" else (min_valid_p if min_valid_p is not None else max(1, min_valid // 10))
        assert n >= need, (what, nm, rep)
        assert round(f3 * n) <= allowed_beyond_3(n, dof), (what, nm, allowed_beyond_3(n, dof), rep)
        assert zmax < 5.0, (what, nm, rep)
        # rms of n z scores: 1 within ~1 / sqrt(2 n) (more between neighbouring pixels, which share packets).  The reference's
        # propagation for P is not an exact sigma (factor 1/2 under its root, I-Q-U correlations ignored): wider band.
        assert rms < 1.05 + 4.0 / math.sqrt(2.0 * n) + (0.25 if nm == "P" else 0.0), (what, nm, rep)


def totals_z(batches_a, batches_b):
    """Disk-integrated I, Q, U and P: z of the run totals with the batch variance."""
    out = {}
    fa = [np.array([b[0, k].sum() for b in batches_a]) for k in range(3)]
    fb = [np.array([b[0, k].sum() for b in batches_b]) for k in range(3)]
    for k, nm in enumerate("IQU"):
        out[nm] = abs(fa[k].mean() - fb[k].mean()) / math.sqrt(fa[k].var(ddof=1) / len(fa[k]) + fb[k].var(ddof=1) / len(fb[k]))
    pa = np.hypot(fa[1], fa[2]) / fa[0]
    pb = np.hypot(fb[1], fb[2]) / fb[0]
    out["P"] = abs(pa.mean() - pb.mean()) / math.sqrt(pa.var(ddof=1) / len(pa) + pb.var(ddof=1) / len(pb))
    return out
