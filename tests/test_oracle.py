"""CPU tests of the oracle itself (oracle/artes_oracle.cc).

The reference ships no golden vectors (PARITY UNPINNED), so the oracle is pinned by
 * known-answer vectors of the generators it restates,
 * analytic anchors of the physics (Lambert sphere phase law, Rayleigh single scattering),
 * geometric invariants of cell_face,
 * the committed fixtures under tests/golden (regression of the oracle, and the vectors the GPU
   tests are compared against).
"""
import math
import os

import numpy as np
import pytest

import oracle_lib as OL
from artes_b200.abi import make_launch
from tools import atmospheres as A

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- generators ---------------------------------------------------------------------------------
def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert OL.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert OL.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert OL.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_uniform_stream_layout():
    u = OL.philox_uniforms(seed=5, pid=(7 << 32) | 3, n=8)
    w = OL.philox([3, 7, 0, 0], [5, 0]) + OL.philox([3, 7, 1, 0], [5, 0])
    np.testing.assert_array_equal(u, (np.array(w, dtype=np.float64) + 0.5) / 4294967296.0)
    assert u.min() > 0.0 and u.max() < 1.0


def test_marsaglia_zaman_matches_python_restatement():
    """src/ARTES.f90:4205-4216 restated with Python integers (32-bit wrap emulated)."""
    def wrap(v):
        v &= 0xffffffff
        return v - (1 << 32) if v & 0x80000000 else v
    s = [123456, 362436069, 16163801, 1131199299]
    exp = []
    for _ in range(1000):
        imz = s[0] - s[2]
        if imz < 0:
            imz += 2147483579
        s[0], s[1], s[2] = s[1], s[2], imz
        s[3] = wrap(69069 * s[3] + 1013904243)
        imz = wrap(imz + s[3])
        exp.append(0.5 + 0.23283064e-9 * imz)
    got = OL.mz_uniforms(123456, 1000)
    np.testing.assert_array_equal(got, np.array(exp))
    assert 0.0 < got.min() and got.max() < 1.0 and abs(got.mean() - 0.5) < 0.03


# ---- analytic anchors ------------------------------------------------------------------------------
@pytest.mark.parametrize("alpha_deg", [20.0, 75.0, 130.0])
def test_lambert_sphere_phase_law(oracle_factory, alpha_deg):
    """Empty atmosphere over a Lambertian surface: A_g = 2/3 and the Lambert phase law."""
    atm = A.lambert_sphere()
    o, depth = oracle_factory(atm)
    assert depth == 0
    n = 150000
    a = math.radians(alpha_deg)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(n_photons=n, x_max=xm, y_max=xm, seed=3, surface_albedo=1.0, det_phi=a, nx=1, ny=1)
    r = o.run(L)
    hit = (atm.rfront[0] / atm.rfront[-1]) ** 2  # photons are launched over the disk of the top face
    expect = hit * (2.0 / (3.0 * math.pi)) * (math.sin(a) + (math.pi - a) * math.cos(a)) / math.pi
    got = r["det"][0, 0].sum() / n
    assert abs(got / expect - 1.0) < 0.012
    assert r["det"][0, 1].sum() == 0.0 and r["det"][0, 2].sum() == 0.0  # Lambert surface depolarises
    assert int(r["err"].sum()) == 0


def test_thin_rayleigh_single_scattering_polarisation(oracle_factory):
    """Optically thin Rayleigh gas seen at 90 deg phase angle: P = sin^2/(1+cos^2) -> ~1, U -> 0,
    and the stored Q is negative (sign flip of src/ARTES.f90:4956)."""
    rfront = A.R_JUP + np.array([0.0, 100e3, 200e3])
    b = A._Builder(rfront, [0.0, 180.0], [0.0], [0.7])
    b.add_region(A.rayleigh([0.7]), 1e-7, (0, 2), (0, 1), (0, 1))
    atm = b.finish("thin")
    assert atm.radial_tau() < 1e-3
    o, _ = oracle_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(n_photons=100000, x_max=xm, y_max=xm, seed=5, nx=1, ny=1)
    r = o.run(L)
    I, Q, U, V = (r["det"][0, k].sum() for k in range(4))
    assert I > 0 and Q < 0
    assert -Q / I > 0.985
    assert abs(U) / I < 0.01 and V == 0.0


def rayleigh_deep_observables(runner, n, seed=3):
    """(pi * I / n at full phase, -Q/I and U/I at 90 deg) of A.rayleigh_deep through `runner(launch) -> result`."""
    atm = A.rayleigh_deep()
    xm = 1.3 * atm.rfront[-1]
    out = []
    for adeg in (0.0573, 90.0):     # det_phi is kept 1e-3 rad off zero by the reference (:492)
        L = make_launch(n_photons=n, x_max=xm, y_max=xm, seed=seed, surface_albedo=1.0, det_phi=math.radians(adeg), nx=1, ny=1, fstop=1e-7)
        r = runner(atm, L)
        assert int(r["err"].sum()) == 0
        out.append([r["det"][0, k].sum() / n for k in range(4)])
    (i0, _, _, _), (i90, q90, u90, _) = out
    return math.pi * i0, -q90 / i90, u90 / i90


def test_rayleigh_semi_infinite_literature_anchor():
    """Multiple scattering WITH polarisation against the literature: the conservative semi-infinite Rayleigh planet has
    geometric albedo 0.7975 when polarisation is carried through every scattering (Prather 1974; Buenzli & Schmid 2009,
    A&A 504, 259) but 0.75 in the scalar approximation, and a disk-integrated polarisation of ~0.325 near quadrature,
    perpendicular to the scattering plane (stored Q < 0, :4956).  A wrong Mueller algebra, rotation sign or phase-function
    sampling in the oracle's scattering loop moves these numbers by several per cent (the scalar value is 6 % away)."""
    from oracle_lib import Oracle

    def runner(atm, L):
        o = Oracle()
        o.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
        o.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], 0)    # surface = the planet (no tau > 30 cut)
        return o.run(L)
    ag, p90, u90 = rayleigh_deep_observables(runner, 40000)
    assert abs(ag / 0.7975 - 1.0) < 0.015, ag          # 40 000 packets: sigma ~ 0.5 %
    assert 0.310 < p90 < 0.340, p90
    assert abs(u90) < 0.01


def chandrasekhar_h(omega=1.0, n=64):
    """H-function of isotropic scattering with single-scattering albedo omega on a Gauss-Legendre grid of (0, 1):
    1/H(mu) = sqrt(1 - omega) + omega/2 int mu' H(mu') / (mu + mu') dmu'  (Chandrasekhar 1960, ch. V).  Returns (mu, weights, H);
    conservative case: H(1) = 2.9078, int H = 2, int H mu = 2/sqrt(3)."""
    x, w = np.polynomial.legendre.leggauss(n)
    mu, w = 0.5 * (x + 1.0), 0.5 * w
    H = np.ones_like(mu)
    for _ in range(4000):
        Hn = 1.0 / (math.sqrt(1.0 - omega) + 0.5 * omega * np.sum(w * mu * H / (mu[:, None] + mu[None, :]), axis=1))
        done = np.max(np.abs(Hn - H)) < 1e-13
        H = 0.5 * (H + Hn)
        if done:
            break
    return mu, w, H


def h_at(m, omega, mu, w, H):
    """H(m) for arbitrary m in [0, 1] from the grid solution (the integral equation itself is the interpolation formula)"""
    m = np.asarray(m, dtype=np.float64)
    return 1.0 / (math.sqrt(1.0 - omega) + 0.5 * omega * np.sum(w * mu * H / (m[..., None] + mu), axis=-1))


def isotropic_phase_law(alpha, omega=1.0):
    """pi x (reflected flux / incident flux on the disk) of the semi-infinite isotropic atmosphere at phase angle alpha, i.e. geometric
    albedo x phase function: (1/pi) int int mu0 R(mu, mu0) mu cos(psi) dpsi dlambda over the lit and visible part of the sphere,
    R = omega H(mu) H(mu0) / (4 (mu + mu0)), mu = cos(psi) cos(lambda), mu0 = cos(psi) cos(alpha - lambda)."""
    mu_g, w_g, H = chandrasekhar_h(omega)
    x, wq = np.polynomial.legendre.leggauss(96)
    lam = 0.5 * (x + 1.0) * (math.pi - alpha) + (alpha - 0.5 * math.pi)       # lambda in [alpha - pi/2, pi/2]
    wl = 0.5 * (math.pi - alpha) * wq
    psi = 0.5 * math.pi * x                                                    # psi in [-pi/2, pi/2]
    wp = 0.5 * math.pi * wq
    cp = np.cos(psi)[:, None]
    m, m0 = cp * np.cos(lam)[None, :], cp * np.cos(alpha - lam)[None, :]
    R = 0.25 * omega * h_at(m, omega, mu_g, w_g, H) * h_at(m0, omega, mu_g, w_g, H) / (m + m0)
    return float(np.sum(wp[:, None] * wl[None, :] * m0 * R * m * cp) / math.pi)


def isotropic_deep_observables(runner, n, npix=31, seed=5, omega=1.0, ntheta=1, nphi=1):
    """Full-phase image of A.isotropic_deep through `runner(atm, launch) -> result` against Chandrasekhar's semi-infinite atmosphere:
    returns (geometric albedo, its theoretical value, measured / expected intensity in five rings of equal projected area)."""
    atm = A.isotropic_deep(omega=omega, ntheta=ntheta, nphi=nphi)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(n_photons=n, x_max=xm, y_max=xm, seed=seed, surface_albedo=1.0, det_phi=math.radians(0.0573), nx=npix, ny=npix, fstop=1e-7)
    r = runner(atm, L)
    assert int(r["err"].sum()) == 0
    img = r["det"][0, 0] / n                      # sum = A_g / pi; a pixel holds (1/pi) (I/F) dA / (pi R^2)
    assert r["det"][0, 1].sum() == 0.0 and r["det"][0, 2].sum() == 0.0      # F11 only: unpolarised
    mu, w, H = chandrasekhar_h(omega)
    ag_theory = 0.25 * omega * np.sum(w * H * H * mu)
    # expected image: I/F = omega H(mu)^2 / 8 at mu = sqrt(1 - rho^2), integrated over every pixel on an 8 x 8 sub-grid
    sub = 8
    edges = np.linspace(-xm, xm, npix * sub + 1)
    c = 0.5 * (edges[1:] + edges[:-1]) / atm.rfront[-1]
    rho2 = c[None, :] ** 2 + c[:, None] ** 2
    Hm = h_at(np.sqrt(np.clip(1.0 - rho2, 0.0, None)), omega, mu, w, H)
    fine = np.where(rho2 < 1.0, omega * Hm * Hm / 8.0, 0.0) * (edges[1] - edges[0]) ** 2 / (math.pi * atm.rfront[-1] ** 2) / math.pi
    expect = fine.reshape(npix, sub, npix, sub).sum(axis=(1, 3))
    pc = 0.5 * (np.linspace(-xm, xm, npix + 1)[1:] + np.linspace(-xm, xm, npix + 1)[:-1]) / atm.rfront[-1]
    ring = np.minimum((5.0 * (pc[None, :] ** 2 + pc[:, None] ** 2)).astype(int), 5)      # equal-area rings by the pixel centre; 5 = off the disk
    ratios = [img[ring == k].sum() / expect[ring == k].sum() for k in range(5)]
    return math.pi * img.sum(), ag_theory, ratios


def isotropic_deep_phase_points(runner, n, alphas_deg=(60.0, 90.0, 120.0), omega=1.0, seed=6):
    """measured / expected disk-integrated brightness of A.isotropic_deep at the given phase angles (1 x 1 detector)"""
    atm = A.isotropic_deep(omega=omega)
    xm = 1.3 * atm.rfront[-1]
    out = []
    for adeg in alphas_deg:
        L = make_launch(n_photons=n, x_max=xm, y_max=xm, seed=seed, surface_albedo=1.0, det_phi=math.radians(adeg), nx=1, ny=1, fstop=1e-7)
        r = runner(atm, L)
        assert int(r["err"].sum()) == 0
        out.append(math.pi * r["det"][0, 0].sum() / n / isotropic_phase_law(math.radians(adeg), omega))
    return out


def _oracle_runner(atm, L):
    from oracle_lib import Oracle
    o = Oracle()
    o.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
    o.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], 0)    # surface = the planet (no tau > 30 cut)
    return o.run(L)


def test_isotropic_semi_infinite_h_function_anchor():
    """Multiple scattering against an exact solution: the semi-infinite isotropically scattering atmosphere reflects
    I(mu, mu0) = F mu0 omega H(mu) H(mu0) / (4 (mu + mu0)) (Chandrasekhar's H-function, computed here from its integral equation):
    conservative case, full phase: geometric albedo 0.6897 and a definite limb darkening; with absorption (omega = 0.8) the same
    with the H-function of that albedo (survival weighting, :791-813); away from full phase the disk-integrated phase law.  The
    oracle's optical-depth sampling, survival / scattering loop, peel-off weights, e^-tau walks and pixel mapping all enter; none
    of the expected numbers comes from the code under test."""
    mu, w, H = chandrasekhar_h()
    assert abs(h_at(1.0, 1.0, mu, w, H) - 2.9078) < 2e-4 and abs(np.sum(w * H) - 2.0) < 1e-9
    assert abs(np.sum(w * H * mu) - 2.0 / math.sqrt(3.0)) < 1e-9
    assert abs(isotropic_phase_law(1e-9) - 0.25 * np.sum(w * H * H * mu)) < 2e-4      # the phase law ends in the geometric albedo
    ag, ag_theory, ratios = isotropic_deep_observables(_oracle_runner, 40000)
    assert abs(ag_theory - 0.68967) < 1e-4
    assert abs(ag / ag_theory - 1.0) < 0.015, ag
    for k, q in enumerate(ratios):
        assert abs(q - 1.0) < (0.03 if k < 4 else 0.05), (k, q)      # the outermost ring holds the limb pixels (mu < 0.45)
    ag, ag_theory, ratios = isotropic_deep_observables(_oracle_runner, 40000, omega=0.8)
    assert abs(ag / ag_theory - 1.0) < 0.015, (ag, ag_theory)
    for k, q in enumerate(ratios):
        assert abs(q - 1.0) < (0.03 if k < 4 else 0.05), (k, q)
    # the same medium on a 3-D grid (6 polar x 8 azimuthal cells): cones and half-planes are crossed, the answer must not change
    ag, ag_theory, ratios = isotropic_deep_observables(_oracle_runner, 40000, ntheta=6, nphi=8)
    assert abs(ag / ag_theory - 1.0) < 0.015, ag
    for k, q in enumerate(ratios):
        assert abs(q - 1.0) < (0.03 if k < 4 else 0.05), (k, q)
    # (40 000 packets: sigma 0.4 / 0.7 / 1.6 % at 60 / 90 / 120 deg; the crescent at 120 deg also sits ~1 % low: sphericity at the limb)
    for adeg, tol, q in zip((60.0, 90.0, 120.0), (0.02, 0.025, 0.045), isotropic_deep_phase_points(_oracle_runner, 40000)):
        assert abs(q - 1.0) < tol, (adeg, q)


def reflection_semi_infinite(phase, omega, n_mu=32, n_phi=256, m_max=24, tol=1e-12):
    """Scalar reflection function S(mu, mu0, dphi) of a homogeneous semi-infinite atmosphere from Ambartsumian's invariance equation
    (Chandrasekhar 1960, section 29), one azimuthal Fourier component at a time:
      (1/mu + 1/mu0) S^m(mu, mu0) = p^m(mu, -mu0) + 1/2 int S^m(mu, mu'') p^m(-mu'', -mu0) dmu''/mu'' + 1/2 int p^m(mu, mu') S^m(mu', mu0) dmu'/mu'
                                   + 1/4 int int S^m(mu, mu'') p^m(-mu'', mu') S^m(mu', mu0) dmu'' dmu' / (mu'' mu')
    with p = omega x phase function of cos(Theta) = mu mu' + sqrt(1 - mu^2) sqrt(1 - mu'^2) cos(dphi) (signed direction cosines of the
    PROPAGATION directions) and p^m = (1/2pi) int p cos(m dphi) ddphi.  I(0, mu, phi) = F S / (4 mu) for a beam of flux pi F.
    Returns (mu, weights, S^m[m, i, i0])."""
    x, w = np.polynomial.legendre.leggauss(n_mu)
    mu, w = 0.5 * (x + 1.0), 0.5 * w
    dphi = 2.0 * math.pi * np.arange(n_phi) / n_phi
    sn = np.sqrt(1.0 - mu * mu)

    def pm(sign_a, sign_b):          # p^m(sign_a mu_i, sign_b mu_j), m = 0 .. m_max
        ct = (sign_a * sign_b) * mu[:, None, None] * mu[None, :, None] + sn[:, None, None] * sn[None, :, None] * np.cos(dphi)[None, None, :]
        p = omega * phase(np.clip(ct, -1.0, 1.0))
        return np.stack([np.mean(p * np.cos(m * dphi)[None, None, :], axis=2) for m in range(m_max + 1)])
    p_ud, p_dd, p_uu, p_du = pm(+1, -1), pm(-1, -1), pm(+1, +1), pm(-1, +1)      # (out <- in): up <- down, down <- down, up <- up, down <- up
    inv = 1.0 / (1.0 / mu[:, None] + 1.0 / mu[None, :])
    wm = w / mu
    S = np.zeros_like(p_ud)
    for m in range(m_max + 1):
        Sm = p_ud[m] * inv
        for _ in range(5000):
            a = Sm * wm[None, :]
            new = (p_ud[m] + 0.5 * a @ p_dd[m] + 0.5 * (p_uu[m] * wm[None, :]) @ Sm + 0.25 * a @ (p_du[m] * wm[None, :]) @ Sm) * inv
            done = np.max(np.abs(new - Sm)) < tol
            Sm = new
            if done:
                break
        S[m] = Sm
    return mu, w, S


def backscatter_brightness(mu, S):
    """I/F of the reflected light in the exact backscattering direction (a planet at full phase): S(mu, mu, dphi = pi) / (4 mu)"""
    d = np.zeros(len(mu))
    for m in range(S.shape[0]):
        d += (1.0 if m == 0 else 2.0) * ((-1.0) ** m) * np.diag(S[m])
    return d / (4.0 * mu)


def hg_deep_observables(runner, n, g=0.5, omega=0.9, npix=31, seed=5):
    """Full-phase image of A.hg_deep through `runner(atm, launch) -> result` against the invariance-equation solution: returns (geometric
    albedo, its expected value, measured / expected intensity in five rings of equal projected area)."""
    atm = A.hg_deep(g=g, omega=omega)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(n_photons=n, x_max=xm, y_max=xm, seed=seed, surface_albedo=1.0, det_phi=math.radians(0.0573), nx=npix, ny=npix, fstop=1e-7)
    r = runner(atm, L)
    assert int(r["err"].sum()) == 0
    img = r["det"][0, 0] / n
    assert r["det"][0, 1].sum() == 0.0 and r["det"][0, 2].sum() == 0.0      # F12 = 0: unpolarised light stays unpolarised
    mu, w, S = reflection_semi_infinite(lambda c: (1.0 - g * g) / (1.0 + g * g - 2.0 * g * c) ** 1.5, omega)
    f = backscatter_brightness(mu, S)
    ag_expected = float(np.sum(w * f * 2.0 * mu))
    sub = 8
    edges = np.linspace(-xm, xm, npix * sub + 1)
    c = 0.5 * (edges[1:] + edges[:-1]) / atm.rfront[-1]
    rho2 = c[None, :] ** 2 + c[:, None] ** 2
    fm = np.interp(np.sqrt(np.clip(1.0 - rho2, 0.0, None)), mu, f)
    fine = np.where(rho2 < 1.0, fm, 0.0) * (edges[1] - edges[0]) ** 2 / (math.pi * atm.rfront[-1] ** 2) / math.pi
    expect = fine.reshape(npix, sub, npix, sub).sum(axis=(1, 3))
    pc = 0.5 * (np.linspace(-xm, xm, npix + 1)[1:] + np.linspace(-xm, xm, npix + 1)[:-1]) / atm.rfront[-1]
    ring = np.minimum((5.0 * (pc[None, :] ** 2 + pc[:, None] ** 2)).astype(int), 5)
    ratios = [img[ring == k].sum() / expect[ring == k].sum() for k in range(5)]
    return math.pi * img.sum(), ag_expected, ratios


def test_henyey_greenstein_semi_infinite_invariance_anchor():
    """ANISOTROPIC multiple scattering against an independent numerical solution: the reflection function of a semi-infinite atmosphere
    obeys Ambartsumian's invariance equation, solved here by fixed-point iteration per azimuthal Fourier component.  The solver is
    checked first against Chandrasekhar's closed form omega H(mu) H(mu0) / (1/mu + 1/mu0) for isotropic scattering (1e-10); then a
    Henyey-Greenstein phase function (g = 0.5, omega = 0.9, unpolarising) gives the full-phase brightness S(mu, mu, pi) / 4 mu that the
    oracle's image must show.  Tabulated-phase-function sampling (the 180-bin polar CDF, :1534-1661), the interpolated matrix of the
    peel-off (:4763-4951) and the survival weighting all enter beyond first order."""
    for omega in (0.8, 0.95):
        mu, w, S = reflection_semi_infinite(lambda c: np.ones_like(c), omega, n_mu=24, m_max=2)
        _, _, H = chandrasekhar_h(omega, 24)
        exact = omega * H[:, None] * H[None, :] / (1.0 / mu[:, None] + 1.0 / mu[None, :])
        assert np.max(np.abs(S[0] / exact - 1.0)) < 1e-10 and np.max(np.abs(S[1:])) < 1e-14
    ag, ag_expected, ratios = hg_deep_observables(_oracle_runner, 150000)
    assert abs(ag_expected - 0.186924) < 2e-6          # converged: 16, 24, 32 Gauss points agree to 1e-8
    assert abs(ag / ag_expected - 1.0) < 0.015, (ag, ag_expected)
    for k, q in enumerate(ratios):
        assert abs(q - 1.0) < (0.03 if k < 4 else 0.05), (k, q)


def rayleigh_phase_blocks(mu, phi, s_out, s_in, unpolarising=None):
    """3 x 3 phase-matrix blocks P[a, b] in Chandrasekhar's (I_l, I_r, U) representation for the propagation directions
    out_a = (s_out mu[a], phi[a]) and in_b = (s_in mu[b], phi[b]).  Rayleigh scattering projects the incident field on the plane transverse
    to the new direction: E_l' = (l'.l) E_l + (l'.r) E_r, E_r' = (r'.l) E_l + (r'.r) E_r with the unit vectors l = d/dtheta, r = d/dphi of
    each direction (the meridian-plane frame); I_l = |E_l|^2, I_r = |E_r|^2, U = 2 Re E_l E_r*; the factor 3/2 normalises
    (1/4pi) int P dOmega to 1 for unpolarised light.  unpolarising: a test scatterer with that phase function and unpolarised output."""
    def frame(sgn):
        ct, st = sgn * mu, np.sqrt(1.0 - mu * mu)
        l = np.stack([ct * np.cos(phi), ct * np.sin(phi), -st], axis=-1)
        r = np.stack([-np.sin(phi), np.cos(phi), np.zeros_like(phi)], axis=-1)
        k = np.stack([st * np.cos(phi), st * np.sin(phi), ct], axis=-1)
        return l, r, k
    lo, ro, ko = frame(s_out)
    li, ri, ki = frame(s_in)
    ll, lr, rl, rr = lo @ li.T, lo @ ri.T, ro @ li.T, ro @ ri.T
    P = np.zeros(ll.shape + (3, 3))
    if unpolarising is not None:
        P[..., 0, 0] = P[..., 0, 1] = P[..., 1, 0] = P[..., 1, 1] = 0.5 * unpolarising(np.clip(ko @ ki.T, -1.0, 1.0))
        return P
    P[..., 0, 0], P[..., 0, 1], P[..., 0, 2] = ll * ll, lr * lr, ll * lr
    P[..., 1, 0], P[..., 1, 1], P[..., 1, 2] = rl * rl, rr * rr, rl * rr
    P[..., 2, 0], P[..., 2, 1], P[..., 2, 2] = 2.0 * ll * rl, 2.0 * lr * rr, ll * rr + lr * rl
    return 1.5 * P


def vector_reflection_semi_infinite(omega, n_mu=24, n_phi=8, tol=1e-11, unpolarising=None):
    """The invariance equation of reflection_semi_infinite for POLARISED light, Rayleigh scattering: S and P are 3 x 3 matrices per pair of
    directions ((1/mu + 1/mu0) S = P_ud + 1/4pi int S P_dd dOmega''/mu'' + 1/4pi int P_uu S dOmega'/mu' + 1/16pi^2 int int S P_du S ...;
    the rightmost factor acts first), discretised on n_mu Gauss points x n_phi azimuths (Rayleigh scattering has azimuthal orders <= 2, so 8
    equidistant azimuths integrate every product exactly).  Returns (mu, weights, S[a, k, b, k'], n_phi) with direction index a = i n_phi + j."""
    x, w = np.polynomial.legendre.leggauss(n_mu)
    mu, w = 0.5 * (x + 1.0), 0.5 * w
    MU, PH = np.repeat(mu, n_phi), np.tile(2.0 * math.pi * np.arange(n_phi) / n_phi, n_mu)
    nd = n_mu * n_phi

    def big(s_out, s_in):
        return (omega * rayleigh_phase_blocks(MU, PH, s_out, s_in, unpolarising)).transpose(0, 2, 1, 3).reshape(3 * nd, 3 * nd)
    P_ud, P_dd, P_uu, P_du = big(+1, -1), big(-1, -1), big(+1, +1), big(-1, +1)
    W3 = np.repeat(np.repeat(w / mu, n_phi) * (2.0 * math.pi / n_phi), 3)      # dOmega / mu per (direction, Stokes) column
    inv = np.repeat(np.repeat(1.0 / (1.0 / MU[:, None] + 1.0 / MU[None, :]), 3, axis=0), 3, axis=1)
    c1, c2 = 1.0 / (4.0 * math.pi), 1.0 / (16.0 * math.pi ** 2)
    S = P_ud * inv
    for _ in range(20000):
        SW = S * W3[None, :]
        new = (P_ud + c1 * SW @ P_dd + c1 * (P_uu * W3[None, :]) @ S + c2 * SW @ (P_du * W3[None, :]) @ S) * inv
        done = np.max(np.abs(new - S)) < tol
        S = new
        if done:
            break
    return mu, w, S.reshape(nd, 3, nd, 3), n_phi


def vector_backscatter(mu, S, n_phi):
    """(I_l, I_r, U) / F of the light reflected straight back (mu = mu0, phi = phi0 + pi) for unpolarised incident light"""
    out = np.zeros((len(mu), 3))
    for i in range(len(mu)):
        out[i] = S[i * n_phi + n_phi // 2, :, i * n_phi, :] @ np.array([0.5, 0.5, 0.0]) / (4.0 * mu[i])
    return out


def rayleigh_absorbing_observables(runner, n, omega=0.9, npix=31, seed=5):
    """Full-phase image of A.rayleigh_deep(omega) against the vector invariance-equation solution: returns (geometric albedo, its expected
    value, the scalar-theory value, measured / expected intensity in five rings, measured and expected radial polarisation Q_r / I per ring)."""
    atm = A.rayleigh_deep(omega=omega)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(n_photons=n, x_max=xm, y_max=xm, seed=seed, surface_albedo=1.0, det_phi=math.radians(0.0573), nx=npix, ny=npix, fstop=1e-7)
    r = runner(atm, L)
    assert int(r["err"].sum()) == 0
    I, Q, U = (r["det"][0, k] / n for k in range(3))
    mu, w, S, n_phi = vector_reflection_semi_infinite(omega)
    f = vector_backscatter(mu, S, n_phi)
    fI, fQ = f[:, 0] + f[:, 1], f[:, 0] - f[:, 1]              # Q > 0: along the meridian plane, i.e. radial on the disk
    assert np.max(np.abs(f[:, 2])) < 1e-9                        # no U in the meridian frame at exact backscattering (symmetry)
    mus, ws, Ss = reflection_semi_infinite(lambda c: 0.75 * (1.0 + c * c), omega, n_mu=24, m_max=4)
    ag_scalar = float(np.sum(ws * backscatter_brightness(mus, Ss) * 2.0 * mus))
    sub = 8
    edges = np.linspace(-xm, xm, npix * sub + 1)
    c = 0.5 * (edges[1:] + edges[:-1]) / atm.rfront[-1]
    rho2 = c[None, :] ** 2 + c[:, None] ** 2
    m = np.sqrt(np.clip(1.0 - rho2, 0.0, None))
    area = (edges[1] - edges[0]) ** 2 / (math.pi * atm.rfront[-1] ** 2) / math.pi
    eI = (np.where(rho2 < 1.0, np.interp(m, mu, fI), 0.0) * area).reshape(npix, sub, npix, sub).sum(axis=(1, 3))
    eQ = (np.where(rho2 < 1.0, np.interp(m, mu, fQ), 0.0) * area).reshape(npix, sub, npix, sub).sum(axis=(1, 3))
    pc = 0.5 * (np.linspace(-xm, xm, npix + 1)[1:] + np.linspace(-xm, xm, npix + 1)[:-1]) / atm.rfront[-1]
    X, Y = np.meshgrid(pc, pc)                                   # det[.., iy, ix]
    ring = np.minimum((5.0 * (X ** 2 + Y ** 2)).astype(int), 5)
    chi = np.arctan2(Y, X)
    Qr = Q * np.cos(2.0 * chi) + U * np.sin(2.0 * chi)           # radial Stokes parameter from the detector-frame Q, U as the reference stores them
    rI = [I[ring == k].sum() / eI[ring == k].sum() for k in range(5)]
    pol = [Qr[ring == k].sum() / I[ring == k].sum() for k in range(5)]
    pol_expected = [eQ[ring == k].sum() / eI[ring == k].sum() for k in range(5)]
    return math.pi * I.sum(), float(np.sum(w * fI * 2.0 * mu)), ag_scalar, rI, pol, pol_expected


def test_rayleigh_vector_invariance_anchor():
    """POLARISED multiple scattering against an independent numerical solution: the 3 x 3 reflection matrix of the semi-infinite Rayleigh
    atmosphere (omega = 0.9) from the invariance equation in Chandrasekhar's (I_l, I_r, U) representation.  The matrix machinery is checked
    first with an unpolarising scatterer (it must reproduce the scalar solver).  Polarisation raises the geometric albedo from 0.3657
    (scalar theory) to 0.3999 and polarises the limb RADIALLY at full phase (up to 7 % -- single scattering gives exactly zero there): the
    oracle must show both.  A wrong sign in polarization_rotation (:1663-2052) or in the Mueller product lands on the scalar value or
    beyond and flips or kills the limb polarisation."""
    ray = lambda c: 0.75 * (1.0 + c * c)
    mu, w, S, n_phi = vector_reflection_semi_infinite(0.9, n_mu=12, unpolarising=ray)
    f = vector_backscatter(mu, S, n_phi)
    mus, ws, Ss = reflection_semi_infinite(ray, 0.9, n_mu=12, m_max=4)
    assert np.max(np.abs((f[:, 0] + f[:, 1]) / backscatter_brightness(mus, Ss) - 1.0)) < 1e-9
    ag, ag_vector, ag_scalar, rI, pol, pol_expected = rayleigh_absorbing_observables(_oracle_runner, 150000)
    assert abs(ag_vector - 0.39990) < 2e-5 and abs(ag_scalar - 0.36571) < 2e-5
    assert abs(ag / ag_vector - 1.0) < 0.01, (ag, ag_vector)
    for k in range(5):
        assert abs(rI[k] - 1.0) < (0.03 if k < 4 else 0.05), (k, rI[k])
        assert abs(pol[k] - pol_expected[k]) < 0.008, (k, pol[k], pol_expected[k])
    assert pol_expected[4] > 0.06 and pol[4] > 0.05                # radial, and strongest in the limb ring


def rayleigh_disk_theory(alphas, omega, n_mu=24, nq=64):
    """Disk-integrated (geometric albedo x phase function, Q / I, U / I) of the semi-infinite Rayleigh planet at the phase angles `alphas` [rad]
    from the 3 x 3 invariance-equation solution: at every point of the lit and visible crescent the reflected (I_l, I_r, U) for unpolarised
    sunlight (the azimuth dependence of S is a trigonometric polynomial of order 2: exact from the 8 solved azimuths; cubic splines in mu, mu0),
    rotated from the local meridian frame to the frame of the planet's scattering plane (l' = c l + s r with c = l.l', s = r.l':
    Q' = (c^2 - s^2) Q + 2 c s U), integrated with the projected area.  Q < 0: polarised perpendicular to the scattering plane."""
    from scipy.interpolate import RectBivariateSpline
    mu, w, S, M = vector_reflection_semi_infinite(omega, n_mu=n_mu)
    n = len(mu)
    v = np.einsum('ijkab,b->iajk', S.reshape(n, M, 3, n, M, 3)[:, :, :, :, 0, :], np.array([0.5, 0.5, 0.0]))     # [i, i0, j, k]: incident azimuth 0
    v = v * (1.0 / mu[:, None, None, None] + 1.0 / mu[None, :, None, None])                                          # smooth in (mu, mu0)
    F = np.fft.rfft(v, axis=2) / M
    spl = {(m, k, c): RectBivariateSpline(mu, mu, (F[:, :, m, k].real, F[:, :, m, k].imag)[c]) for m in range(3) for k in range(3) for c in (0, 1)}
    x, wq = np.polynomial.legendre.leggauss(nq)
    out = []
    for alpha in alphas:
        o = np.array([1.0, 0.0, 0.0])                                  # towards the observer
        st_ = np.array([math.cos(alpha), math.sin(alpha), 0.0])        # towards the star
        lam = 0.5 * (x + 1.0) * (math.pi - alpha) + (alpha - 0.5 * math.pi)
        LAM, PSI = np.meshgrid(lam, 0.5 * math.pi * x)
        WW = (0.5 * math.pi * wq)[:, None] * (0.5 * (math.pi - alpha) * wq)[None, :]
        nrm = np.stack([np.cos(PSI) * np.cos(LAM), np.cos(PSI) * np.sin(LAM), np.sin(PSI)], axis=-1)
        m_, m0 = nrm @ o, nrm @ st_
        e1 = o[None, None, :] - m_[..., None] * nrm
        e1 /= np.linalg.norm(e1, axis=-1, keepdims=True)               # local azimuth 0 = the outgoing direction
        e2 = np.cross(nrm, e1)
        dphi = -np.arctan2(e2 @ (-st_), e1 @ (-st_))                   # phi_out - phi_in of the propagation directions
        vec = np.zeros(m_.shape + (3,))
        for k in range(3):
            acc = spl[(0, k, 0)].ev(m_, m0)
            for m in (1, 2):
                acc = acc + 2.0 * (spl[(m, k, 0)].ev(m_, m0) * np.cos(m * dphi) - spl[(m, k, 1)].ev(m_, m0) * np.sin(m * dphi))
            vec[..., k] = acc / (1.0 / m_ + 1.0 / m0) / (4.0 * m_)
        l = m_[..., None] * e1 - np.sqrt(np.clip(1.0 - m_ * m_, 0.0, None))[..., None] * nrm      # d/dtheta at o, theta from the local normal
        lp = st_ - (st_ @ o) * o
        lp /= np.linalg.norm(lp)                                       # l' in the scattering plane, perpendicular to o
        c, sn = l @ lp, e2 @ lp
        I, Q, U = vec[..., 0] + vec[..., 1], vec[..., 0] - vec[..., 1], vec[..., 2]
        area = m_ * np.cos(PSI) * WW / math.pi
        tot = float(np.sum(I * area))
        out.append((tot, float(np.sum(((c * c - sn * sn) * Q + 2.0 * c * sn * U) * area)) / tot,
                    float(np.sum((-2.0 * c * sn * Q + (c * c - sn * sn) * U) * area)) / tot))
    return out


def rayleigh_phase_points(runner, n, alphas_deg=(60.0, 90.0, 120.0), omega=0.9, seed=3):
    """[(pi I / n, -Q / I, U / I)] of A.rayleigh_deep(omega) at the given phase angles (1 x 1 detector; the reference stores -Q, :4956)"""
    atm = A.rayleigh_deep(omega=omega)
    xm = 1.3 * atm.rfront[-1]
    out = []
    for adeg in alphas_deg:
        L = make_launch(n_photons=n, x_max=xm, y_max=xm, seed=seed, surface_albedo=1.0, det_phi=math.radians(adeg), nx=1, ny=1, fstop=1e-7)
        r = runner(atm, L)
        assert int(r["err"].sum()) == 0
        i, q, u = (r["det"][0, k].sum() / n for k in range(3))
        out.append((math.pi * i, -q / i, u / i))
    return out


def test_rayleigh_polarised_phase_curve_anchor():
    """The disk-integrated brightness and degree of polarisation of the semi-infinite Rayleigh planet (omega = 0.9) at 60 / 90 / 120 deg phase
    angle against the 3 x 3 invariance-equation solution integrated over the crescent: 0.1772 / 0.0876 / 0.0414 and P = 37.2 / 57.3 / 37.4 %,
    perpendicular to the scattering plane, U = 0.  (At alpha -> 0 the same integral returns the geometric albedo 0.3999 and P = 0.)"""
    th = rayleigh_disk_theory([1e-6] + [math.radians(a) for a in (60.0, 90.0, 120.0)], 0.9)
    assert abs(th[0][0] - 0.39990) < 1e-4 and abs(th[0][1]) < 1e-6
    assert abs(-th[2][1] - 0.5732) < 5e-4 and all(abs(t[2]) < 1e-9 for t in th)
    for (i, p, u), (ti, tq, _) in zip(rayleigh_phase_points(_oracle_runner, 100000), th[1:]):
        assert abs(i / ti - 1.0) < 0.02, (i, ti)
        assert abs(p - (-tq)) < 0.012, (p, -tq)
        assert abs(u) < 0.01


def general_phase_blocks(mu_o, phi_o, s_out, mu_i, phi_i, s_in, Fmat):
    """4 x 4 phase-matrix blocks Z[a, b] in the (I, Q, U, V) representation w.r.t. the meridian frames (l = d/dtheta, r = d/dphi, l x r = k) for
    the propagation directions out_a = (s_out mu_o[a], phi_o[a]), in_b = (s_in mu_i[b], phi_i[b]): rotation of the incident Stokes vector into the
    scattering plane (r_s = k_in x k_out / |..|, l_s = r_s x k), the scattering matrix F(cos Theta) of a macroscopically isotropic, mirror-symmetric
    medium [[F11 F12 0 0] [F12 F22 0 0] [0 0 F33 F34] [0 0 -F34 F44]], rotation into the meridian frame of the new direction.  Every rotation
    angle comes from dot products of unit vectors (for l' = c l + s r: Q' = (c^2 - s^2) Q + 2 c s U, U' = -2 c s Q + (c^2 - s^2) U)."""
    def frame(mu, phi, sgn):
        ct, st = sgn * mu, np.sqrt(1.0 - mu * mu)
        return (np.stack([ct * np.cos(phi), ct * np.sin(phi), -st], axis=-1), np.stack([-np.sin(phi), np.cos(phi), np.zeros_like(phi)], axis=-1),
                np.stack([st * np.cos(phi), st * np.sin(phi), ct], axis=-1))
    lo, ro, ko = frame(mu_o, phi_o, s_out)
    li, ri, ki = frame(mu_i, phi_i, s_in)
    shape = (len(mu_o), len(mu_i), 3)
    KI, LI, RI = (np.broadcast_to(v[None, :, :], shape) for v in (ki, li, ri))
    KO, LO = (np.broadcast_to(v[:, None, :], shape) for v in (ko, lo))
    ct = np.clip(np.sum(KI * KO, axis=-1), -1.0, 1.0)
    rs = np.cross(KI, KO)
    nrm = np.linalg.norm(rs, axis=-1)
    deg = nrm < 1e-12                                                                # forward / backward: any plane through k_in will do
    rs = np.where(deg[..., None], RI, rs / np.where(deg, 1.0, nrm)[..., None])
    ls_i = np.cross(rs, KI)
    ls_o = np.where(deg[..., None], np.where((ct < 0.0)[..., None], -ls_i, ls_i), np.cross(rs, KO))
    c1, s1 = np.sum(ls_i * LI, axis=-1), np.sum(ls_i * RI, axis=-1)                   # l_s = c1 l + s1 r at k_in
    c2, s2 = np.sum(LO * ls_o, axis=-1), np.sum(LO * rs, axis=-1)                    # l_out = c2 l_s + s2 r_s at k_out

    def rot(c, s_):
        R = np.zeros(c.shape + (4, 4))
        R[..., 0, 0] = R[..., 3, 3] = 1.0
        R[..., 1, 1] = R[..., 2, 2] = c * c - s_ * s_
        R[..., 1, 2] = 2.0 * c * s_
        R[..., 2, 1] = -2.0 * c * s_
        return R
    return rot(c2, s2) @ Fmat(ct) @ rot(c1, s1)


def fourier_reflection(Fmat, omega, n_mu=24, n_phi=64, m_max=24, tol=1e-11):
    """Reflection matrix of the semi-infinite atmosphere for a general scattering matrix: the invariance equation of
    vector_reflection_semi_infinite per complex azimuthal Fourier component, S(mu, mu0, dphi) = sum_m S^m e^{i m dphi} (S^-m = conj S^m).
    Returns (mu, weights, S^m[m, i, k, i0, k'] for m = 0 .. m_max)."""
    x, w = np.polynomial.legendre.leggauss(n_mu)
    mu, w = 0.5 * (x + 1.0), 0.5 * w
    dphi = 2.0 * math.pi * np.arange(n_phi) / n_phi

    def pm(s_out, s_in):          # (1/2pi) int Z(mu_a, dphi; mu_b, 0) e^{-i m dphi} ddphi
        Z = omega * general_phase_blocks(np.repeat(mu, n_phi), np.tile(dphi, n_mu), s_out, mu, np.zeros(n_mu), s_in, Fmat)
        Zm = np.fft.fft(Z.reshape(n_mu, n_phi, n_mu, 4, 4), axis=1) / n_phi
        return np.stack([Zm[:, m].transpose(0, 2, 1, 3).reshape(4 * n_mu, 4 * n_mu) for m in range(m_max + 1)])
    P_ud, P_dd, P_uu, P_du = pm(+1, -1), pm(-1, -1), pm(+1, +1), pm(-1, +1)
    W4 = np.repeat(w / mu, 4)
    inv = np.repeat(np.repeat(1.0 / (1.0 / mu[:, None] + 1.0 / mu[None, :]), 4, axis=0), 4, axis=1)
    out = np.zeros((m_max + 1, n_mu, 4, n_mu, 4), dtype=complex)
    for m in range(m_max + 1):
        S = P_ud[m] * inv
        for _ in range(20000):
            SW = S * W4[None, :]
            new = (P_ud[m] + 0.5 * SW @ P_dd[m] + 0.5 * (P_uu[m] * W4[None, :]) @ S + 0.25 * SW @ (P_du[m] * W4[None, :]) @ S) * inv
            done = np.max(np.abs(new - S)) < tol
            S = new
            if done:
                break
        out[m] = S.reshape(n_mu, 4, n_mu, 4)
    return mu, w, out


def backscatter_stokes(mu, Sm):
    """(I, Q, U, V) / F of the light reflected straight back (mu = mu0, dphi = pi) for unpolarised incident light; Q > 0: radial on the disk"""
    res = np.zeros((len(mu), 4))
    for i in range(len(mu)):
        acc = Sm[0, i, :, i, 0].real.copy()
        for m in range(1, Sm.shape[0]):
            acc += 2.0 * (Sm[m, i, :, i, 0] * np.exp(1j * m * math.pi)).real
        res[i] = acc / (4.0 * mu[i])
    return res


def tabulated_matrix(atm):
    """F(cos Theta) of a homogeneous test atmosphere as the transport code sees it: the 180 one-degree bins of its matrix table, interpolated
    linearly between the bin centres (matrix_at_deg, :4763-4810), normalised to (1/4pi) int F11 dOmega = 1"""
    tab = np.asarray(atm.uniq[0]).reshape(-1, 180, 16)[np.asarray(atm.cell_to_uniq[0])[0]] * 4.0 * math.pi
    centres = np.arange(180) + 0.5

    def F(ct):
        deg = np.degrees(np.arccos(np.clip(ct, -1.0, 1.0)))
        out = np.zeros(ct.shape + (4, 4))
        for (r, c), e in {(0, 0): 0, (0, 1): 1, (1, 0): 4, (1, 1): 5, (2, 2): 10, (2, 3): 11, (3, 2): 14, (3, 3): 15}.items():
            out[..., r, c] = np.interp(deg, centres, tab[:, e])
        return out
    return F


def polarising_hg_observables(runner, n, g=0.5, omega=0.9, p_linear=0.5, npix=31, seed=5):
    """Full-phase image of A.hg_deep with the polarising Henyey-Greenstein matrix against the 4 x 4 invariance-equation solution for the
    tabulated matrix: (geometric albedo, expected, measured / expected intensity in five rings, measured and expected radial polarisation)."""
    atm = A.hg_deep(g=g, omega=omega, p_linear=p_linear)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(n_photons=n, x_max=xm, y_max=xm, seed=seed, surface_albedo=1.0, det_phi=math.radians(0.0573), nx=npix, ny=npix, fstop=1e-7)
    r = runner(atm, L)
    assert int(r["err"].sum()) == 0
    I, Q, U = (r["det"][0, k] / n for k in range(3))
    mu, w, Sm = fourier_reflection(tabulated_matrix(atm), omega)
    bs = backscatter_stokes(mu, Sm)
    assert np.max(np.abs(bs[:, 2:])) < 1e-9
    sub = 8
    edges = np.linspace(-xm, xm, npix * sub + 1)
    c = 0.5 * (edges[1:] + edges[:-1]) / atm.rfront[-1]
    rho2 = c[None, :] ** 2 + c[:, None] ** 2
    m = np.sqrt(np.clip(1.0 - rho2, 0.0, None))
    area = (edges[1] - edges[0]) ** 2 / (math.pi * atm.rfront[-1] ** 2) / math.pi
    eI = (np.where(rho2 < 1.0, np.interp(m, mu, bs[:, 0]), 0.0) * area).reshape(npix, sub, npix, sub).sum(axis=(1, 3))
    eQ = (np.where(rho2 < 1.0, np.interp(m, mu, bs[:, 1]), 0.0) * area).reshape(npix, sub, npix, sub).sum(axis=(1, 3))
    pc = 0.5 * (np.linspace(-xm, xm, npix + 1)[1:] + np.linspace(-xm, xm, npix + 1)[:-1]) / atm.rfront[-1]
    X, Y = np.meshgrid(pc, pc)
    ring = np.minimum((5.0 * (X ** 2 + Y ** 2)).astype(int), 5)
    chi = np.arctan2(Y, X)
    Qr = Q * np.cos(2.0 * chi) + U * np.sin(2.0 * chi)
    return (math.pi * I.sum(), float(np.sum(w * bs[:, 0] * 2.0 * mu)), [I[ring == k].sum() / eI[ring == k].sum() for k in range(5)],
            [Qr[ring == k].sum() / I[ring == k].sum() for k in range(5)], [eQ[ring == k].sum() / eI[ring == k].sum() for k in range(5)])


def test_polarising_henyey_greenstein_invariance_anchor():
    """Polarised AND anisotropic: the cloud species of C2 (Henyey-Greenstein F11 with F12 = -p_linear F11 sin^2 / (1 + cos^2), F33 = F11 2 cos /
    (1 + cos^2); g = 0.5, omega = 0.9) against the 4 x 4 invariance-equation solution for a general scattering matrix -- the Stokes vector is
    rotated into the scattering plane, multiplied with F, rotated into the new meridian frame, all angles from dot products of unit vectors.
    That machinery is checked first with the Rayleigh matrix against the field-projection solver (two formulations of the same physics:
    1e-10).  Expected: geometric albedo 0.1932, radial limb polarisation 0.5 % (centre) to 7.5 % (limb) at full phase."""
    def rayleigh_f(ct):
        F = np.zeros(ct.shape + (4, 4))
        F[..., 0, 0] = F[..., 1, 1] = 0.75 * (1.0 + ct * ct)
        F[..., 0, 1] = F[..., 1, 0] = 0.75 * (ct * ct - 1.0)
        F[..., 2, 2] = F[..., 3, 3] = 1.5 * ct
        return F
    mu, w, Sm = fourier_reflection(rayleigh_f, 0.9, n_mu=12, n_phi=16, m_max=4)
    b = backscatter_stokes(mu, Sm)
    mu2, _, S2, n_phi = vector_reflection_semi_infinite(0.9, n_mu=12)
    f2 = vector_backscatter(mu2, S2, n_phi)
    assert np.max(np.abs(b[:, 0] / (f2[:, 0] + f2[:, 1]) - 1.0)) < 1e-9 and np.max(np.abs(b[:, 1] - (f2[:, 0] - f2[:, 1]))) < 1e-10
    assert np.max(np.abs(Sm[3:])) < 1e-12                            # Rayleigh scattering has azimuthal orders 0, 1, 2 only
    ag, ag_expected, rI, pol, pol_expected = polarising_hg_observables(_oracle_runner, 150000)
    assert abs(ag_expected - 0.19319) < 5e-5
    assert abs(ag / ag_expected - 1.0) < 0.012, (ag, ag_expected)
    for k in range(5):
        assert abs(rI[k] - 1.0) < (0.03 if k < 4 else 0.05), (k, rI[k])
        assert abs(pol[k] - pol_expected[k]) < (0.008 if k < 4 else 0.012), (k, pol[k], pol_expected[k])
    assert pol_expected[4] > 0.07 and pol[4] > 0.06


def test_limb_biased_emission_against_the_crescent_solution():
    """Phase angles >= 170 deg are run with limb-biased emission (packets start on the outer 19 % of the disk only, :1041-1055) and a package
    energy scaled by that area fraction (:2526-2530).  At 172.5 deg the thin crescent of the semi-infinite Rayleigh planet (omega = 0.9) has, from
    the invariance-equation solution, a brightness of 5.77e-4 and a NEGATIVE polarisation (parallel to the scattering plane: multiple scattering
    wins over the vanishing single-scattering polarisation near forward scattering), P = -1.7 %."""
    adeg = 172.5
    (ti, tq, _), = rayleigh_disk_theory([math.radians(adeg)], 0.9)
    assert abs(ti - 5.7708e-4) < 2e-7 and abs(-tq - (-0.0171)) < 3e-4
    atm = A.rayleigh_deep(omega=0.9)
    xm = 1.3 * atm.rfront[-1]
    n = 200000
    L = make_launch(n_photons=n, x_max=xm, y_max=xm, seed=3, surface_albedo=1.0, det_phi=math.radians(adeg), nx=1, ny=1, fstop=1e-7, limb_emission=1)
    r = _oracle_runner(atm, L)
    i, q, u = (r["det"][0, k].sum() / n for k in range(3))
    assert abs(0.19 * math.pi * i / ti - 1.0) < 0.04, (0.19 * math.pi * i, ti)        # sigma ~ 1.5 % at 2e5 packets
    assert -0.035 < -q / i < -0.005 and abs(u / i) < 0.015, (-q / i, u / i)


def test_vector_invariance_solver_reaches_the_literature_value():
    """The reference solution itself against the literature: towards the conservative limit the geometric albedo of the semi-infinite
    Rayleigh atmosphere behaves as A(1) - b sqrt(1 - omega) + c (1 - omega); the 3 x 3 solver at omega = 0.99, 0.999, 0.9999 extrapolates to
    0.7976 -- Prather (1974) / Buenzli & Schmid (2009): 0.7975 (scalar theory: 0.7506)."""
    s, a = [], []
    for om in (0.99, 0.999, 0.9999):
        mu, w, S, n_phi = vector_reflection_semi_infinite(om, n_mu=12, tol=1e-10)
        f = vector_backscatter(mu, S, n_phi)
        s.append(math.sqrt(1.0 - om))
        a.append(float(np.sum(w * (f[:, 0] + f[:, 1]) * 2.0 * mu)))
    a1 = np.linalg.solve(np.array([[1.0, -x, x * x] for x in s]), np.array(a))[0]
    assert abs(a1 - 0.7975) < 1e-3, a1


def test_host_photometry_and_error_planes_match_oracle_restatement():
    """The tail of radiative_transfer (:957-1004) and the error planes of write_output (:3481-3519): the Python host
    mirror (artes_b200/host.py, which the driver tests compare bin/ARTES with) against the oracle's restatement, on
    detectors with empty pixels, single-deposit pixels and unpolarised pixels."""
    import oracle_lib
    from artes_b200 import host
    rs = np.random.RandomState(1)
    for nx, ny in ((1, 1), (7, 5), (25, 25)):
        cnt = rs.poisson(2.0, (4, ny, nx)).astype(float)
        cnt[1:] = cnt[1]                                        # Q, U, V share one count plane (:4969-4972)
        w = rs.random_sample((4, ny, nx)) * cnt
        w[1:3] -= 0.5 * cnt[1:3]
        w2 = w ** 2 / np.maximum(cnt, 1.0) + rs.random_sample((4, ny, nx)) * (cnt > 1)      # single deposits: variance exactly 0
        if nx > 1:
            w[1:3, 0, 0] = 0.0                                  # an unpolarised pixel: sigma_P stays 0
        det_sum = np.stack([w, w2, cnt])
        energy = 3.3e-20
        d1, p1 = oracle_lib.finish_detector(det_sum, energy)
        d2 = host.detector_from_sums(det_sum, energy)
        p2 = host.photometry(d2)
        np.testing.assert_array_equal(d1, d2)
        np.testing.assert_allclose(p2, p1, rtol=1e-12, atol=0.0)
        e1, e2 = oracle_lib.stokes_error(d1), host.stokes_error(d2)
        np.testing.assert_allclose(e2, e1, rtol=1e-12, atol=1e-14 * np.abs(e1).max())
        if nx > 1:
            assert e1[4, 0, 0] == 0.0
    p = host.Params(phase_curve=True, det_phi=math.radians(172.5))
    for src in (1, 2):
        p.photon_source = src
        a = oracle_lib.package_energy(p, [1.0, 7.0e7], 0.7e-6, 1e6, 4.2e11)
        b = host.package_energy(p, [1.0, 7.0e7], 0.7e-6, 1e6, 4.2e11)
        assert abs(a / b - 1.0) < 1e-14
    assert abs(oracle_lib.lib().artes_ref_planck(5800.0, 0.7e-6, 1) / host.planck_function(5800.0, 0.7e-6, 1) - 1.0) < 1e-14


def test_reference_sigma_underestimates_the_noise(oracle_factory, atmospheres):
    """Why the statistical gate (tests/stat_gate.py) measures the photon noise from independent batches instead of using
    the reference's per-pixel sigma (:3490-3493): two runs of the oracle ITSELF with different seeds are consistent under
    the batch variance (rms z ~ 1) but differ by ~3 of the reference's sigmas in Stokes I on the template atmosphere --
    that estimate ignores the correlation of the ~40 peel-off deposits a packet makes into neighbouring pixels."""
    import oracle_lib
    import stat_gate
    atm = atmospheres("c1_template_rayleigh")
    o, _ = oracle_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    K, n = 16, 8000
    runs = [[o.run(make_launch(n_photons=n, x_max=xm, y_max=xm, seed=s0 + i, nx=25, ny=25), rng=oracle_lib.RNG_MZ)["det"] for i in range(K)]
            for s0 in (100, 5000)]
    rep = stat_gate.z_report(*runs)
    for nm in "IQU":
        assert 0.7 < rep[nm][3] < 1.3 and rep[nm][2] < 5.0, rep
    ref = stat_gate.z_report_ref_sigma(*runs)
    assert ref["I"][3] > 2.0 and ref["I"][1] > 0.1, ref          # rms z > 2, > 10 % of the pixels beyond "3 sigma"


def test_mirror_symmetry_of_detector_azimuth(oracle_factory, atmospheres):
    """phi_det -> -phi_det mirrors the scene: same I and Q, opposite U (within noise)."""
    atm = atmospheres("c2_hg_deck")
    o, _ = oracle_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    out = []
    for phi in (math.radians(60.0), math.radians(300.0)):
        L = make_launch(n_photons=60000, x_max=xm, y_max=xm, seed=9, nx=1, ny=1, det_phi=phi)
        d = o.run(L)["det"]
        out.append([d[0, k].sum() for k in range(3)])
    (i1, q1, u1), (i2, q2, u2) = out
    assert abs(i1 / i2 - 1.0) < 0.03 and abs(q1 / q2 - 1.0) < 0.06
    assert abs(u1 + u2) < 0.1 * abs(q1)


def test_thread_count_does_not_change_philox_result(oracle_factory, atmospheres):
    atm = atmospheres("c1_template_rayleigh")
    o, _ = oracle_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(n_photons=3000, x_max=xm, y_max=xm, seed=2)
    a = o.run(L, nthreads=1)
    b = o.run(L, nthreads=4)
    np.testing.assert_allclose(a["det"], b["det"], rtol=1e-11, atol=0)
    assert a["stats"]["n_cell_face"] == b["stats"]["n_cell_face"]


def test_thermal_emission_bookkeeping(oracle_factory, atmospheres):
    from artes_b200 import host
    atm = atmospheres("c3_molecular")
    o, depth = oracle_factory(atm, l=0, photon_source=2)
    vol = host.cell_volume(atm.rfront, atm.thetafront(), atm.phifront())
    cw, lum, cdf = host.thermal_tables(depth, atm.k_abs[0], atm.temperature, vol, atm.wavelengths[0] * 1e-6,
                                       atm.nr, atm.ntheta, atm.nphi)
    o.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], depth, cw, cdf)
    xm = 1.3 * atm.rfront[-1]
    for emission in (1, 2):
        L = make_launch(n_photons=20000, x_max=xm, y_max=xm, seed=4, photon_source=2, photon_emission=emission, nx=1, ny=1)
        r = o.run(L)
        assert r["flux"][0] > 0 and 0 < r["flux"][1] <= r["flux"][0] * 1.0000001
        assert r["det"][0, 0].sum() > 0
        assert int(r["err"].sum()) == 0


# ---- geometry ------------------------------------------------------------------------------------------
def random_interior_points(atm, n, seed):
    rs = np.random.RandomState(seed)
    ir = rs.randint(0, atm.nr, n); it = rs.randint(0, atm.ntheta, n); ip = rs.randint(0, atm.nphi, n)
    r = atm.rfront[ir] + rs.random_sample(n) * (atm.rfront[ir + 1] - atm.rfront[ir])
    th = np.radians(atm.theta_deg[it] + rs.random_sample(n) * (atm.theta_deg[it + 1] - atm.theta_deg[it]))
    phf = np.append(atm.phi_deg, 360.0)
    ph = np.radians(phf[ip] + rs.random_sample(n) * (phf[ip + 1] - phf[ip]))
    pos = np.stack([r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), r * np.cos(th)], 1)
    d = rs.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    return pos, d, np.zeros((n, 2), np.int32), np.stack([ir, it, ip], 1).astype(np.int32)


def test_cell_face_lands_on_the_reported_face(oracle_factory, atmospheres):
    atm = atmospheres("c4_mie_patches")
    o, _ = oracle_factory(atm)
    pos, d, face, cell = random_interior_points(atm, 20000, 1)
    oi, od = o.cell_face(pos, d, face, cell)
    assert (oi[:, 6] == 0).all() and (od > 0).all()
    hit = pos + od[:, None] * d
    r = np.linalg.norm(hit, axis=1)
    rad = oi[:, 0] == 1
    np.testing.assert_allclose(r[rad], atm.rfront[oi[rad, 1]], rtol=1e-12)
    pol = oi[:, 0] == 2
    np.testing.assert_allclose(np.degrees(np.arccos(hit[pol, 2] / r[pol])), atm.theta_deg[oi[pol, 1]], atol=1e-8)
    azi = oi[:, 0] == 3
    ph = np.degrees(np.arctan2(hit[azi, 1], hit[azi, 0])) % 180.0   # full planes: modulo 180
    want = atm.phi_deg[oi[azi, 1]] % 180.0
    dphi = np.abs(ph - want); dphi = np.minimum(dphi, 180.0 - dphi)
    assert dphi.max() < 1e-8
    # the next cell differs from the current one in exactly the crossed coordinate
    changed = (oi[:, 2:5] != cell).sum(axis=1)
    assert (changed == 1).all()
    assert rad.sum() > 0 and pol.sum() > 0 and azi.sum() > 0


# ---- committed fixtures -------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "c2", "c4"])
def test_oracle_reproduces_golden(oracle_factory, atmospheres, name):
    from golden.make_golden import CASES, build_case
    g = np.load(os.path.join(GOLDEN, f"golden_{name}.npz"))
    atm, launch_trace, launch_run, xi = build_case(name, atmospheres)
    o, depth = oracle_factory(atm)
    assert depth == int(g["cell_depth"])
    t = o.trace(launch_trace, xi, max_rec=0)
    np.testing.assert_array_equal(t["len"], g["seq_len"])
    np.testing.assert_array_equal(t["hash"], g["seq_hash"])
    r = o.run(launch_run)
    np.testing.assert_allclose(r["det"], g["det"], rtol=1e-9, atol=1e-300)
    assert CASES[name]["trace_n"] == len(g["seq_len"])


def test_statistical_gate_on_cpu_mz_against_philox(oracle_factory, atmospheres):
    """The statistical gate of the GPU tests (tests/stat_gate.py, SURVEY 8d thresholds), exercised CPU-only: the reference's
    Marsaglia-Zaman stream against the product's Philox stream through the same oracle -- the CPU proxy of
    tests/test_gpu_statistical.py (the GPU replays the Philox side to rounding)."""
    import stat_gate
    atm = atmospheres("c4_mie_patches")
    o, _ = oracle_factory(atm)
    xm = 1.3 * atm.rfront[-1]
    kw = dict(x_max=xm, y_max=xm, nx=12, ny=12, det_phi=math.radians(60.0))
    K, n = 16, 12000
    ba = [o.run(make_launch(n_photons=n, seed=100 + i, **kw), rng=OL.RNG_MZ)["det"] for i in range(K)]
    bb = [o.run(make_launch(n_photons=n, seed=7, photon_id_base=i * n, **kw), rng=OL.RNG_PHILOX)["det"] for i in range(K)]
    rep = stat_gate.z_report(ba, bb)
    stat_gate.assert_gate(rep, names=("I", "Q", "U", "P"), min_valid=30, what="c4 12x12 MZ vs Philox")
