"""Synthetic ARTES atmospheres C1..C5 in Python 3 / numpy.

Setup tooling, NOT part of the hot path: the reference builds `atmosphere.fits` with the
Python-2 scripts `python/opacity*.py` + `python/atmosphere.py`, which cannot run here
(Python 2 syntax, astropy missing, ComputePart binary missing).  This module restates the
formulas those scripts use and the file layout they write, so that the same inputs can be fed to
the CPU oracle, to libartes_gpu and to the `bin/ARTES` driver:

  * Rayleigh species            python/opacityRayleigh.py:54-122
  * Henyey-Greenstein species   python/opacityHenyeyGreenstein.py:61-105
  * isotropic species           python/opacityIsotropic.py
  * Mie species                 own Lorenz-Mie integration (bin/ComputePartLinux is a missing
                                blob); 6 -> 16 element expansion of python/opacityMie.py:118-129
  * P11 normalisation           python/atmosphere.py:29-65 (Simpson on the bin centres)
  * hydrostatic radial grid     python/atmosphere.py:127-167
  * per-cell mixing             python/atmosphere.py:330-372 (incl. the fact that `density` only
                                holds the gas density while species are mixed)
  * atmosphere.fits layout      python/atmosphere.py:449-459, read at src/ARTES.f90:2067-2201

The per-cell scattering matrices are kept DE-DUPLICATED (unique 180x16 blocks + a cell->block
map); `dense_matrix()` expands them to the reference's dense HDU for small grids.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

R_JUP = 69911e3  # python/atmosphere.py:116
ANGLE_C = (np.arange(180) + 0.5) * math.pi / 180.0  # bin centres, python/atmosphere.py:25-27


# ----------------------------------------------------------------------------------------------
# species: opacity [cm2 g-1] (ext, abs, sca) and 180x16 matrix per wavelength
# ----------------------------------------------------------------------------------------------
@dataclass
class Species:
    name: str
    wavelengths: np.ndarray  # [nl] micron
    k_abs: np.ndarray        # [nl] cm2/g
    k_sca: np.ndarray        # [nl] cm2/g
    matrix: np.ndarray       # [nl,180,16]


def _simps_avg(y, x):
    """scipy.integrate.simps(y, x, even='avg') of the scipy the reference was written for."""
    def simp(yy, xx):  # odd number of samples, uniform spacing assumed
        h = xx[1] - xx[0]
        return h / 3.0 * (yy[0] + yy[-1] + 4.0 * yy[1:-1:2].sum() + 2.0 * yy[2:-1:2].sum())
    n = len(y)
    if n % 2 == 1:
        return simp(y, x)
    first = simp(y[:-1], x[:-1]) + 0.5 * (x[-1] - x[-2]) * (y[-1] + y[-2])
    last = simp(y[1:], x[1:]) + 0.5 * (x[1] - x[0]) * (y[1] + y[0])
    return 0.5 * (first + last)


def normalise(matrix):
    """python/atmosphere.py:60-65: 2 pi int P11 sin(T) dT = 1 on the bin centres."""
    out = np.array(matrix, dtype=np.float64, copy=True)
    for l in range(out.shape[0]):
        norm = 2.0 * math.pi * _simps_avg(out[l, :, 0] * np.sin(ANGLE_C), ANGLE_C)
        out[l] /= norm
    return out


def _bin_average(fn):
    """Average of the matrix at the two bin edges j deg and j+1 deg (opacityRayleigh.py:110-122)."""
    m = np.zeros((180, 16))
    for j in range(180):
        lo = fn(math.cos(j * math.pi / 180.0))
        up = fn(math.cos((j + 1) * math.pi / 180.0))
        m[j] = (lo + up) / 2.0
    return m


def rayleigh(wavelengths, depol=0.0, ssa=1.0, mmw=2.02):
    wavelengths = np.atleast_1d(np.asarray(wavelengths, dtype=np.float64))
    avogadro, loschmidt = 6.02214129e23, 2.6867805e19
    mass = mmw / avogadro
    ksca = np.zeros(len(wavelengths))
    for i, wl in enumerate(wavelengths):  # opacityRayleigh.py:54-72
        a, b = 13.58e-5, 7.52e-3
        ri = 1.0 + a + a * b / (wl * wl)
        rindex = (ri * ri - 1.0) ** 2 / (ri * ri + 2.0) ** 2
        dep = (6.0 + 3.0 * depol) / (6.0 - 7.0 * depol)
        cross = 24.0 * math.pi ** 3 * rindex * dep / (((wl * 1e-4) ** 4) * loschmidt ** 2)
        ksca[i] = cross / mass
    delta = (1.0 - depol) / (1.0 + depol / 2.0)
    delta_p = (1.0 - 2.0 * depol) / (1.0 - depol)

    def F(alpha):  # opacityRayleigh.py:86-104
        m = np.zeros(16)
        m[0] = alpha * alpha + 1.0
        m[1] = alpha * alpha - 1.0
        m[4] = m[1]
        m[5] = m[0]
        m[10] = 2.0 * alpha
        m[15] = delta_p * m[10]
        m = delta * m
        m[0] += 1.0 - delta
        return m

    mat = _bin_average(F)
    mats = normalise(np.repeat(mat[None], len(wavelengths), axis=0))
    return Species("rayleigh", wavelengths, ksca / ssa - ksca, ksca, mats)


def henyey_greenstein(wavelengths, g=0.9, p_linear=0.5, k_sca=1.0, k_abs=0.0):
    wavelengths = np.atleast_1d(np.asarray(wavelengths, dtype=np.float64))

    def F(alpha):  # opacityHenyeyGreenstein.py:75-93 (single term, pCircular = skew = 0)
        m = np.zeros(16)
        m[0] = (1.0 - g * g) / ((1.0 + g * g - 2.0 * g * alpha) ** 1.5)
        m[1] = -p_linear * m[0] * (1.0 - alpha * alpha) / (1.0 + alpha * alpha)
        m[4] = m[1]
        m[5] = m[0]
        m[10] = m[0] * (2.0 * alpha) / (1.0 + alpha * alpha)
        m[15] = m[10]
        return m

    mats = normalise(np.repeat(_bin_average(F)[None], len(wavelengths), axis=0))
    n = len(wavelengths)
    return Species("hg", wavelengths, np.full(n, k_abs), np.full(n, k_sca), mats)


def isotropic(wavelengths, k_sca=1.0, k_abs=0.0):
    wavelengths = np.atleast_1d(np.asarray(wavelengths, dtype=np.float64))
    m = np.zeros((180, 16))
    m[:, 0] = 1.0 / (4.0 * math.pi)
    mats = normalise(np.repeat(m[None], len(wavelengths), axis=0))
    n = len(wavelengths)
    return Species("isotropic", wavelengths, np.full(n, k_abs), np.full(n, k_sca), mats)


def _mie_amplitudes(x, m, mu):
    """Lorenz-Mie S1, S2 at cos(angle)=mu plus Qext, Qsca (Bohren & Huffman recurrences)."""
    nmax = int(x + 4.0 * x ** (1.0 / 3.0) + 2.0)
    mx = m * x
    nmx = int(max(nmax, abs(mx)) + 16)
    D = np.zeros(nmx + 1, dtype=complex)
    for n in range(nmx, 0, -1):
        D[n - 1] = n / mx - 1.0 / (D[n] + n / mx)
    psi0, psi1 = math.cos(x), math.sin(x)
    chi0, chi1 = -math.sin(x), math.cos(x)
    xi1 = complex(psi1, -chi1)
    pi0 = np.zeros_like(mu)
    pi1 = np.ones_like(mu)
    S1 = np.zeros_like(mu, dtype=complex)
    S2 = np.zeros_like(mu, dtype=complex)
    qext = qsca = 0.0
    for n in range(1, nmax + 1):
        psi = (2 * n - 1) / x * psi1 - psi0
        chi = (2 * n - 1) / x * chi1 - chi0
        xi = complex(psi, -chi)
        an = ((D[n] / m + n / x) * psi - psi1) / ((D[n] / m + n / x) * xi - xi1)
        bn = ((m * D[n] + n / x) * psi - psi1) / ((m * D[n] + n / x) * xi - xi1)
        pi = pi1
        tau = n * mu * pi - (n + 1) * pi0
        fn = (2 * n + 1) / (n * (n + 1.0))
        S1 += fn * (an * pi + bn * tau)
        S2 += fn * (an * tau + bn * pi)
        qext += (2 * n + 1) * (an + bn).real
        qsca += (2 * n + 1) * (abs(an) ** 2 + abs(bn) ** 2)
        psi0, psi1 = psi1, psi
        chi0, chi1 = chi1, chi
        xi1 = complex(psi1, -chi1)
        pi1 = ((2 * n + 1) * mu * pi - (n + 1) * pi0) / n
        pi0 = pi
    return S1, S2, 2.0 / (x * x) * qext, 2.0 / (x * x) * qsca


def mie(wavelengths, r_eff=1.4, v_eff=0.05, n_re=1.42, n_im=1e-6, rho_p=1.0, n_radii=60):
    """Gamma (Hansen) size distribution of homogeneous spheres; [micron], rho_p [g cm-3]."""
    wavelengths = np.atleast_1d(np.asarray(wavelengths, dtype=np.float64))
    mu = np.cos(ANGLE_C)
    radii = np.linspace(max(0.05 * r_eff, r_eff * (1 - 6 * math.sqrt(v_eff))), r_eff * (1 + 8 * math.sqrt(v_eff)), n_radii)
    w = radii ** ((1.0 - 3.0 * v_eff) / v_eff) * np.exp(-radii / (r_eff * v_eff))
    w /= w.sum()
    mats = np.zeros((len(wavelengths), 180, 16))
    kext = np.zeros(len(wavelengths))
    ksca = np.zeros(len(wavelengths))
    for l, wl in enumerate(wavelengths):
        f = np.zeros((180, 6))
        cext = csca = vol = 0.0
        for r, wr in zip(radii, w):
            x = 2.0 * math.pi * r / wl
            S1, S2, qe, qs = _mie_amplitudes(x, complex(n_re, n_im), mu)
            f[:, 0] += wr * 0.5 * (abs(S1) ** 2 + abs(S2) ** 2)
            f[:, 1] += wr * 0.5 * (abs(S2) ** 2 - abs(S1) ** 2)
            f[:, 2] += wr * 0.5 * (abs(S1) ** 2 + abs(S2) ** 2)
            f[:, 3] += wr * (S1 * np.conj(S2)).real
            f[:, 4] += wr * (S2 * np.conj(S1)).imag
            f[:, 5] += wr * (S1 * np.conj(S2)).real
            cext += wr * qe * math.pi * r * r
            csca += wr * qs * math.pi * r * r
            vol += wr * 4.0 / 3.0 * math.pi * r ** 3
        m = mats[l]  # opacityMie.py:118-129
        m[:, 0] = f[:, 0]; m[:, 1] = f[:, 1]; m[:, 4] = f[:, 1]; m[:, 5] = f[:, 2]
        m[:, 10] = f[:, 3]; m[:, 11] = f[:, 4]; m[:, 14] = -f[:, 4]; m[:, 15] = f[:, 5]
        mass = vol * 1e-12 * rho_p  # micron^3 -> cm^3 * g/cm^3
        kext[l] = cext * 1e-8 / mass  # micron^2 -> cm^2
        ksca[l] = csca * 1e-8 / mass
    return Species("mie", wavelengths, kext - ksca, ksca, normalise(mats))


# ----------------------------------------------------------------------------------------------
# atmosphere = grid + per-wavelength opacities + de-duplicated matrices
# ----------------------------------------------------------------------------------------------
@dataclass
class Atmosphere:
    name: str
    rfront: np.ndarray        # [nr+1] m
    theta_deg: np.ndarray     # [ntheta+1]
    phi_deg: np.ndarray       # [nphi]
    wavelengths: np.ndarray   # [nl] micron
    k_sca: np.ndarray         # [nl, cells]  (cell index: r fastest, then theta, then phi)
    k_abs: np.ndarray         # [nl, cells]
    uniq: list                # nl arrays [n_uniq_l, 180, 16]
    cell_to_uniq: list        # nl int32 arrays [cells]
    density: np.ndarray       # [cells] kg m-3 (HDU 4; read and discarded by ARTES)
    temperature: np.ndarray   # [cells] K
    artes_in: dict = field(default_factory=dict)  # keyword -> value for artes.in
    photons: float = 1e6
    seed: int = 1

    @property
    def nr(self): return len(self.rfront) - 1
    @property
    def ntheta(self): return len(self.theta_deg) - 1
    @property
    def nphi(self): return len(self.phi_deg)
    @property
    def cells(self): return self.nr * self.ntheta * self.nphi

    # grid in the units / derived forms src/ARTES.f90:2085-2122 produces
    def thetafront(self):
        return self.theta_deg * math.pi / 180.0
    def thetaplane(self):
        t = self.theta_deg
        return np.where((t < 90.0 - 1e-6) | (t > 90.0 + 1e-6), 1, 2).astype(np.int32)
    def phifront(self):
        return self.phi_deg * math.pi / 180.0

    def dense_matrix(self, l):
        """HDU-8 block of wavelength l in numpy order (180,16,nphi,ntheta,nr)."""
        u = self.uniq[l][self.cell_to_uniq[l]]  # [cells,180,16]
        return np.ascontiguousarray(u.reshape(self.nphi, self.ntheta, self.nr, 180, 16).transpose(3, 4, 0, 1, 2))

    def radial_tau(self, l=0, j=0, k=0):
        kap = (self.k_sca[l] + self.k_abs[l]).reshape(self.nphi, self.ntheta, self.nr)[k, j]
        return float((kap * np.diff(self.rfront)).sum())


class _Builder:
    """Restates the cell loop of python/atmosphere.py:330-372 with de-duplicated matrices."""

    def __init__(self, rfront, theta_deg, phi_deg, wavelengths):
        self.rfront = np.asarray(rfront, dtype=np.float64)
        self.theta_deg = np.asarray(theta_deg, dtype=np.float64)
        self.phi_deg = np.asarray(phi_deg, dtype=np.float64)
        self.wl = np.atleast_1d(np.asarray(wavelengths, dtype=np.float64))
        self.nr, self.nt, self.np_ = len(self.rfront) - 1, len(self.theta_deg) - 1, len(self.phi_deg)
        nl = len(self.wl)
        shp = (self.np_, self.nt, self.nr)
        self.ksca = np.zeros((nl,) + shp)
        self.kabs = np.zeros((nl,) + shp)
        self.gas_density = np.zeros(shp)   # what atmosphere.py's `density` holds during mixing
        self.density = np.zeros(shp)
        self.temperature = np.zeros(shp)
        self.uniq = [[np.zeros((180, 16))] for _ in range(nl)]   # block 0 = empty cell
        self.key = np.zeros((nl,) + shp, dtype=np.int64)

    def add_gas(self, density_layers, species_layers, temperature_layers=None):
        """gas: on (atmosphere.py:335-347): per radial layer density [kg m-3] and species."""
        for i in range(self.nr):
            sp = species_layers[i] if isinstance(species_layers, (list, tuple)) else species_layers
            for l in range(len(self.wl)):
                self.kabs[l, :, :, i] = density_layers[i] * sp.k_abs[l] / 10.0
                self.ksca[l, :, :, i] = density_layers[i] * sp.k_sca[l] / 10.0
                self.uniq[l].append(sp.matrix[l].copy())
                self.key[l, :, :, i] = len(self.uniq[l]) - 1
            self.gas_density[:, :, i] = density_layers[i]
            self.density[:, :, i] = density_layers[i]
            if temperature_layers is not None:
                self.temperature[:, :, i] = temperature_layers[i]

    def add_region(self, sp, density_gcm3, r, t, p):
        """opacityNN line (atmosphere.py:349-372): r/t/p are (in, out) index ranges."""
        dens = density_gcm3 * 1e3
        sl = (slice(p[0], p[1]), slice(t[0], t[1]), slice(r[0], r[1]))
        for l in range(len(self.wl)):
            o_sca = dens * sp.k_sca[l] / 10.0
            o_abs = dens * sp.k_abs[l] / 10.0
            ks, ka, key = self.ksca[l][sl], self.kabs[l][sl], self.key[l][sl]
            gd = self.gas_density[sl]
            weight = np.where(gd > 0.0, (o_sca + o_abs) / (o_sca + o_abs + ks + ka + (gd <= 0.0)), 1.0)
            pairs = np.stack([key.ravel(), weight.ravel().view(np.int64)], axis=1)
            uq, inv = np.unique(pairs, axis=0, return_inverse=True)
            new_ids = np.zeros(len(uq), dtype=np.int64)
            for n, (k_old, wbits) in enumerate(uq):
                w = np.array([wbits], dtype=np.int64).view(np.float64)[0]
                if w == 1.0:
                    m = sp.matrix[l].copy()
                else:
                    m = self.uniq[l][k_old] * (1.0 - w)
                    m += w * sp.matrix[l]
                self.uniq[l].append(m)
                new_ids[n] = len(self.uniq[l]) - 1
            self.key[l][sl] = new_ids[inv.ravel()].reshape(key.shape)
            self.ksca[l][sl] = ks + o_sca
            self.kabs[l][sl] = ka + o_abs
        self.density[sl] += dens

    def finish(self, name, **kw):
        nl = len(self.wl)
        uniq, c2u = [], []
        for l in range(nl):
            used, inv = np.unique(self.key[l].ravel(), return_inverse=True)
            uniq.append(np.stack([self.uniq[l][k] for k in used]))
            c2u.append(inv.astype(np.int32))
        return Atmosphere(name, self.rfront, self.theta_deg, self.phi_deg, self.wl,
                          self.ksca.reshape(nl, -1).copy(), self.kabs.reshape(nl, -1).copy(), uniq, c2u,
                          self.density.ravel().copy(), self.temperature.ravel().copy(), **kw)


TEMPLATE_ARTES_IN = {  # template/artes.in
    "general:log": "off", "general:email": "", "photon:source": "star", "photon:fstop": "1d-5",
    "photon:minimum": "1d-20", "photon:weight": "on", "photon:scattering": "on",
    "photon:emission": "isotropic", "photon:bias": "0.8", "star:temperature": "5800", "star:radius": "1",
    "star:direction": "off", "planet:surface_albedo": "0", "planet:oblateness": "0", "planet:orbit": "5",
    "planet:ring": "off", "detector:type": "imaging_mono", "detector:theta": "90", "detector:phi": "90",
    "detector:pixel": "25", "detector:distance": "10", "output:flow_global": "off",
    "output:flow_latitudinal": "off",
}


def _artes_in(**over):
    d = dict(TEMPLATE_ARTES_IN)
    d.update(over)
    return d


def c1_template_rayleigh():
    """C1: template/atmosphere.in grid, homogeneous Rayleigh gas, 0.7 micron."""
    rfront = R_JUP + np.array([0.0, 100e3, 200e3])
    theta = np.array([0.0, 30, 75, 89.9, 90.1, 95, 150, 180.0])
    b = _Builder(rfront, theta, [0.0], [0.7])
    b.add_region(rayleigh([0.7]), 1e-2, (0, 2), (0, 7), (0, 1))
    return b.finish("c1_template_rayleigh", artes_in=_artes_in(), photons=1e6, seed=1)


def _exp_gas(nr, dz, rho0_gcm3, H):
    z = (np.arange(nr) + 0.5) * dz
    return rho0_gcm3 * 1e3 * np.exp(-z / H)


def c2_hg_deck(nr=20, ntheta=18):
    """C2: 2-D r-theta grid, Rayleigh gas + Henyey-Greenstein polar caps, phase curve."""
    rfront = R_JUP + np.arange(nr + 1) * 10e3
    theta = np.linspace(0.0, 180.0, ntheta + 1)
    b = _Builder(rfront, theta, [0.0], [0.7])
    b.add_gas(_exp_gas(nr, 10e3, 1.16e-2, 40e3), rayleigh([0.7]))
    hg = henyey_greenstein([0.7])
    b.add_region(hg, 2e-6, (5, 10), (0, 3), (0, 1))
    b.add_region(hg, 2e-6, (5, 10), (ntheta - 3, ntheta), (0, 1))
    return b.finish("c2_hg_deck", artes_in=_artes_in(**{"detector:type": "phase"}), photons=1e6, seed=2)


def c3_molecular(nr=100, nl=32):
    """C3: isothermal 800 K hydrostatic column, 100 layers, Rayleigh + synthetic absorption band."""
    T_iso, mmw, log_g = 800.0, 2.02e-3, 3.4
    P = np.logspace(-3, 2, nr + 1)[::-1] * 1e5  # pressureTemperatureIsothermal.py:16, reversed as atmosphere.py:141
    g = 1e-2 * 10.0 ** log_g
    H = 8.3144621 * T_iso / (mmw * g)
    radial = np.zeros(nr + 1)
    for i in range(1, nr + 1):  # atmosphere.py:151-155
        radial[i] = radial[i - 1] - H * np.log(P[i] / P[i - 1])
    dens = P[:-1] / (g * H)
    wl = np.linspace(0.5, 1.0, nl)
    ray = rayleigh(wl)
    kabs = 5e-5 * (1.0 + np.cos(2.0 * math.pi * wl / 0.1))  # SURVEY 8d fallback band [cm2 g-1]
    gas = Species("gas", wl, kabs, ray.k_sca, ray.matrix)
    b = _Builder(R_JUP + radial, [0.0, 180.0], [0.0], wl)
    b.add_gas(dens, gas, temperature_layers=np.full(nr, T_iso))
    return b.finish("c3_molecular", artes_in=_artes_in(**{"detector:type": "spectrum"}), photons=1e6, seed=3)


def _patchy(name, nr, ntheta, nphi, wl, dz, layers, phi_blocks, seed, photons, pixels, cloud_tau=10.0):
    rfront = R_JUP + np.arange(nr + 1) * dz
    theta = np.linspace(0.0, 180.0, ntheta + 1)
    phi = np.arange(nphi) * (360.0 / nphi)
    b = _Builder(rfront, theta, phi, wl)
    b.add_gas(_exp_gas(nr, dz, 1.16e-2, 40e3), rayleigh(wl))
    cloud = mie(wl)
    # mass density giving a radial cloud optical depth `cloud_tau` at the first wavelength
    kext = (cloud.k_sca[0] + cloud.k_abs[0]) / 10.0
    dens_gcm3 = cloud_tau / ((layers[1] - layers[0]) * dz) / kext * 1e-3
    for p0, p1 in phi_blocks:
        b.add_region(cloud, dens_gcm3, layers, (0, ntheta), (p0, p1))
    return b.finish(name, artes_in=_artes_in(**{"detector:phi": "60", "detector:pixel": str(pixels)}),
                    photons=photons, seed=seed)


def c4_mie_patches(nr=20, ntheta=18, nphi=36):
    """C4: 3-D grid, Rayleigh gas + Mie cloud patches in longitude, Stokes images."""
    return _patchy("c4_mie_patches", nr, ntheta, nphi, [0.7], 10e3, (8, 13), [(0, 9), (18, 27)], 4, 1e7, 64)


def c5_scale(nr=100, ntheta=60, nphi=120, nl=4):
    """C5: scale grid 100 x 60 x 120, 4 wavelengths."""
    wl = np.linspace(0.55, 0.85, nl)
    return _patchy("c5_scale", nr, ntheta, nphi, wl, 2e3, (40, 65), [(0, 30), (60, 90)], 5, 1e10, 64)


def lambert_sphere():
    """Analytic anchor: empty atmosphere over a Lambertian surface (tau=0 everywhere)."""
    rfront = R_JUP + np.array([0.0, 100e3])
    b = _Builder(rfront, [0.0, 180.0], [0.0], [0.7])
    b.add_region(isotropic([0.7]), 0.0, (0, 1), (0, 1), (0, 1))
    return b.finish("lambert_sphere", artes_in=_artes_in(**{"planet:surface_albedo": "1"}))


def rayleigh_deep(tau=32.0, nr=8, thickness=8e3, omega=1.0):
    """Literature anchor: a homogeneous, conservative Rayleigh atmosphere of radial optical depth `tau` over a white
    Lambert surface (planet:surface_albedo=1, cell_depth 0) -- for tau >~ 30 the conservative semi-infinite Rayleigh
    planet of Prather (1974) / Buenzli & Schmid (2009): geometric albedo 0.7975 with polarisation (0.75 without),
    disk-integrated polarisation ~ 0.325 near 90 deg phase angle."""
    rfront = R_JUP + np.linspace(0.0, thickness, nr + 1)
    b = _Builder(rfront, [0.0, 180.0], [0.0], [0.7])
    b.add_region(rayleigh([0.7]), 1.0, (0, nr), (0, 1), (0, 1))
    atm = b.finish("rayleigh_deep", artes_in=_artes_in(**{"planet:surface_albedo": "1"}))
    atm.k_sca = atm.k_sca * (tau / atm.radial_tau())
    atm.k_abs = atm.k_abs * 0.0
    if omega < 1.0:      # the same with absorption: single-scattering albedo omega at the same total optical depth
        atm.k_abs = atm.k_sca * (1.0 - omega)
        atm.k_sca = atm.k_sca * omega
    return atm


def isotropic_deep(tau=32.0, nr=8, thickness=8e3, omega=1.0, ntheta=1, nphi=1):
    """Analytic anchor for multiple scattering: a homogeneous, conservative, isotropically scattering atmosphere of radial optical
    depth `tau` over a white Lambert surface -- for tau >~ 30 Chandrasekhar's conservative semi-infinite atmosphere, whose emergent
    intensity at full phase is F H(mu)^2 / 8 (H = the H-function of isotropic scattering; geometric albedo 1/4 int H^2 mu dmu = 0.6897).
    omega < 1: the same with absorption (single-scattering albedo omega): I = omega F H^2 / 8 with the H-function of that albedo.
    ntheta, nphi > 1: the same homogeneous medium cut into polar / azimuthal cells (the walk then crosses cones and half-planes too)."""
    rfront = R_JUP + np.linspace(0.0, thickness, nr + 1)
    b = _Builder(rfront, np.linspace(0.0, 180.0, ntheta + 1), np.arange(nphi) * (360.0 / nphi), [0.7])
    b.add_region(isotropic([0.7]), 1.0, (0, nr), (0, ntheta), (0, nphi))
    atm = b.finish("isotropic_deep", artes_in=_artes_in(**{"planet:surface_albedo": "1"}))
    atm.k_abs = atm.k_abs * 0.0
    atm.k_sca = atm.k_sca * (tau / atm.radial_tau())          # radial optical depth `tau` in total,
    atm.k_abs = atm.k_sca * (1.0 - omega)                       # of which the single-scattering albedo `omega` scatters
    atm.k_sca = atm.k_sca * omega
    return atm


def hg_deep(g=0.5, omega=0.9, tau=32.0, nr=8, thickness=8e3, p_linear=0.0):
    """Numerical-solution anchor for ANISOTROPIC multiple scattering: a homogeneous atmosphere of radial optical depth `tau` with an
    unpolarising Henyey-Greenstein phase function (p_linear = 0: F12 = 0, so unpolarised light stays unpolarised and scalar transfer
    theory is exact) and single-scattering albedo `omega`; semi-infinite for tau >~ 30 (tests/test_oracle.py solves Ambartsumian's
    invariance equation for its reflection function).  p_linear > 0: the polarising Henyey-Greenstein matrix of
    python/opacityHenyeyGreenstein.py:75-93 (the cloud species of C2): the 4 x 4 solver of tests/test_oracle.py applies."""
    rfront = R_JUP + np.linspace(0.0, thickness, nr + 1)
    b = _Builder(rfront, [0.0, 180.0], [0.0], [0.7])
    b.add_region(henyey_greenstein([0.7], g=g, p_linear=p_linear), 1.0, (0, nr), (0, 1), (0, 1))
    atm = b.finish("hg_deep", artes_in=_artes_in(**{"planet:surface_albedo": "1"}))
    atm.k_abs = atm.k_abs * 0.0
    atm.k_sca = atm.k_sca * (tau / atm.radial_tau())
    atm.k_abs = atm.k_sca * (1.0 - omega)
    atm.k_sca = atm.k_sca * omega
    return atm


CONFIGS = {"c1": c1_template_rayleigh, "c2": c2_hg_deck, "c3": c3_molecular, "c4": c4_mie_patches, "c5": c5_scale}


def write_atmosphere_fits(atm: Atmosphere, path):
    """atmosphere.fits with the 9 HDUs of python/atmosphere.py:449-459 (dense matrices)."""
    from artes_b200 import fitsio
    nl = len(atm.wavelengths)
    shp3 = (atm.nphi, atm.ntheta, atm.nr)
    mats = np.stack([atm.dense_matrix(l) for l in range(nl)], axis=2)  # (180,16,nl,nphi,ntheta,nr)
    hdus = [
        ("radial", atm.rfront), ("polar", atm.theta_deg), ("azimuthal", atm.phi_deg),
        ("wavelength", atm.wavelengths), ("density", atm.density.reshape(shp3)),
        ("temperature", atm.temperature.reshape(shp3)),
        ("scattering", atm.k_sca.reshape((nl,) + shp3)), ("absorption", atm.k_abs.reshape((nl,) + shp3)),
        ("scattermatrix", mats),
    ]
    fitsio.write_hdus(path, hdus)
