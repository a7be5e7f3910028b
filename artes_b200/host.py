"""Host-side mirror of the reference's interface around `radiative_transfer`.

The reference keeps everything in program-scope variables; what feeds the transport call and what
is done with its result is restated here with the reference's names so that parity tests and
bench.py read like `run` (src/ARTES.f90:121-267):

    grid_initialize(2)  -> cell_depth(), thermal_tables()       (:2329-2453)
    photon_package      -> package_energy()                      (:2509-2539)
    radiative_transfer  -> Transport.radiative_transfer()        (:518-1006; the photon loop runs on
                           the GPU through libartes_gpu, the scaling/photometry of :957-1004 here)
    write_output        -> stokes_error()                        (:3481-3519)

Python stands in for the Fortran host only in tests and bench.py; the shipped driver is the C++
`bin/ARTES` (src/host), which calls the same C-ABI.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import abi
from .lib import GpuTransport

# constants src/ARTES.f90:9-16
PI = 4.0 * math.atan(1.0)
K_B = 1.3806488e-23
HH = 6.62606957e-34
CC = 2.99792458e8
R_SUN = 6.95500e8
PC = 3.08572e16
AU = 1.49598e11


def planck_function(temperature, wavelength, photon_source):
    """:1350-1367 [W m-2 m-1] (star) or [W m-2 m-1 sr-1] (planet)."""
    pref = 2.0 * PI if photon_source == 1 else 2.0
    return (pref * HH * CC * CC / (wavelength ** 5.0)) / (math.exp(HH * CC / (wavelength * K_B * temperature)) - 1.0)


def cell_depth(rfront, k_sca, k_abs, nr, ntheta, nphi, photon_source=1, ring=False):
    """:2329-2393 -- deepest radial layer a photon may reach (tau > 30 extinction / > 5 absorption)."""
    kap = (np.asarray(k_sca) + np.asarray(k_abs)) if photon_source == 1 else np.asarray(k_abs)
    kap = kap.reshape(nphi, ntheta, nr)
    limit = 30.0 if photon_source == 1 else 5.0
    grid_out = 2 if (photon_source == 2 and ring) else 0
    rf = np.asarray(rfront, dtype=np.float64)
    # top-down running optical depth of every (theta, phi) column, summed in the reference's order (:2345-2375);
    # the depth of a column is the layer in which it first exceeds the limit, else the bottom layer
    layers = np.arange(nr - 1 - grid_out, -1, -1)                       # nr-i-1 for i = grid_out .. nr-1
    dtau = kap[:, :, layers] * (rf[layers + 1] - rf[layers])
    tot = np.cumsum(dtau, axis=2)
    over = tot > limit
    first = np.where(over.any(axis=2), over.argmax(axis=2), len(layers) - 1)
    return int(min(1000000, layers[first].min())) if len(layers) else 1000000


def cell_volume(rfront, thetafront, phifront, oblate=(1.0, 1.0, 1.0)):
    """:2272-2307, returned flat in cell order (r fastest)."""
    nr, nt, npn = len(rfront) - 1, len(thetafront) - 1, len(phifront)
    tcos = np.cos(thetafront)
    vol = np.zeros((npn, nt, nr))
    for k in range(npn):
        dphi = 2.0 * PI if npn == 1 else ((phifront[k + 1] - phifront[k]) if k < npn - 1 else (2.0 * PI - phifront[k]))
        for j in range(nt):
            for i in range(nr):
                vol[k, j, i] = oblate[0] * oblate[1] * oblate[2] * (1.0 / 3.0) * (rfront[i + 1] ** 3 - rfront[i] ** 3) * \
                    (tcos[j] - tcos[j + 1]) * dphi
    return vol.ravel()


def thermal_tables(depth, k_abs, temperature, volume, wavelength, nr, ntheta, nphi, thermal_weight=True):
    """:2395-2453 -> cell_weight, cell_luminosity, emissivity_cumulative (flat, r fastest)."""
    ka = np.asarray(k_abs).reshape(nphi, ntheta, nr)
    tt = np.asarray(temperature).reshape(nphi, ntheta, nr)
    vv = np.asarray(volume).reshape(nphi, ntheta, nr)
    weight_norm = 0.0
    for i in range(depth, nr):
        for j in range(ntheta):
            for k in range(nphi):
                if tt[k, j, i] > 0.0:
                    weight_norm = weight_norm + ka[k, j, i] * planck_function(tt[k, j, i], wavelength, 2) * vv[k, j, i]
    cw = np.zeros((nphi, ntheta, nr))
    lum = np.zeros((nphi, ntheta, nr))
    cdf = np.zeros((nphi, ntheta, nr))
    total = 0.0
    for i in range(depth, nr):
        for j in range(ntheta):
            for k in range(nphi):
                if tt[k, j, i] > 0.0 and ka[k, j, i] > 0.0:
                    pf = planck_function(tt[k, j, i], wavelength, 2)
                    cw[k, j, i] = weight_norm / (vv[k, j, i] * ka[k, j, i] * pf) if thermal_weight else 1.0
                    lum[k, j, i] = 4.0 * PI * vv[k, j, i] * ka[k, j, i] * pf
                    cdf[k, j, i] = total + lum[k, j, i] * cw[k, j, i]
                    total = cdf[k, j, i]
                else:
                    cdf[k, j, i] = total
    return cw.ravel(), lum.ravel(), cdf.ravel()


@dataclass
class Params:
    """The artes.in keywords that reach the transport call, with the defaults of :283-314."""
    photon_source: int = 1
    fstop: float = 1e-5
    photon_minimum: float = 1e-20
    thermal_weight: bool = True
    photon_scattering: bool = True
    photon_emission: int = 1
    photon_bias: float = 0.8
    t_star: float = 5800.0
    r_star: float = R_SUN
    stellar_direction: bool = False
    theta_star: float = PI / 2.0
    phi_star: float = 0.0
    surface_albedo: float = 0.0
    oblateness: float = 0.0
    orbit: float = 5.0 * AU
    ring: bool = False
    phase_curve: bool = False
    det_theta: float = PI / 2.0
    det_phi: float = PI / 2.0
    nx: int = 25
    ny: int = 25
    distance_planet: float = 10.0 * PC
    flow_global: bool = False
    flow_theta: bool = False


def package_energy(p: Params, rfront, wavelength, packages, emis_total=0.0):
    """photon_package :2509-2539 (wavelength in metres)."""
    if p.photon_source == 1:
        e = PI * planck_function(p.t_star, wavelength, 1) * rfront[-1] * rfront[-1] * p.r_star * p.r_star / \
            (p.orbit * p.orbit * p.distance_planet * p.distance_planet * float(packages))
        if p.phase_curve and p.det_phi * 180.0 / PI >= 170.0:
            e = e * (PI * p.r_star * p.r_star - 0.9 * 0.9 * PI * p.r_star * p.r_star) / (PI * p.r_star * p.r_star)
        return e
    return emis_total / (p.distance_planet * p.distance_planet * float(packages))


def detector_from_sums(det_sum, energy):
    """:959-975: detector(:,:,:,1) = sum*E, (:,:,:,2) = sum*E^2, (:,:,:,3) = counts."""
    det = np.array(det_sum, copy=True)
    det[0] *= energy
    det[1] = det[1] * energy * energy      # the reference's order: (sum * E) * E
    return det


def photometry(det):
    """:977-1004 on det[l, stokes, iy, ix]."""
    ph = np.zeros(11)
    ph[0], ph[2], ph[4], ph[6] = det[0, 0].sum(), det[0, 1].sum(), det[0, 2].sum(), det[0, 3].sum()
    ph[8] = math.sqrt(ph[2] ** 2 + ph[4] ** 2)
    ph[9] = ph[8] / ph[0] if ph[0] != 0 else 0.0
    for i in range(4):
        n = det[2, i].sum()
        if n > 0.0:
            dummy = det[1, i].sum() / n - (det[0, i].sum() / n) ** 2
            if dummy > 0.0:
                ph[2 * i + 1] = math.sqrt(dummy) * math.sqrt(n)
    if ph[2] ** 2 + ph[4] ** 2 > 0.0:
        dpi = math.sqrt(((ph[2] * ph[3]) ** 2 + (ph[4] * ph[5]) ** 2) / (2.0 * (ph[2] ** 2 + ph[4] ** 2)))
        ph[10] = ph[9] * math.sqrt((dpi / ph[8]) ** 2 + (ph[1] / ph[0]) ** 2)
    return ph


def stokes_error(det):
    """write_output :3481-3519 -> error[5, iy, ix] (sigma I,Q,U,V and sigma P)."""
    err = np.zeros((5,) + det.shape[2:])
    n = det[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        dummy = np.where(n > 0, det[1] / n - (det[0] / n) ** 2, 0.0)
        err[:4] = np.where((n > 0) & (dummy > 0), np.sqrt(np.abs(dummy)) * np.sqrt(n), 0.0)
        q, u, i_ = det[0, 1], det[0, 2], det[0, 0]
        pol2 = q * q + u * u
        pol = np.sqrt(pol2)
        dpol = np.sqrt(((q * err[1]) ** 2 + (u * err[2]) ** 2) / (2.0 * pol2))
        e5 = (pol / i_) * np.sqrt((dpol / pol) ** 2 + (err[0] / i_) ** 2)
        err[4] = np.where((i_ > 0) & (pol2 > 0), e5, 0.0)
    return np.nan_to_num(err)


class Transport:
    """One atmosphere on the GPU(s): get_atmosphere + grid_initialize + radiative_transfer."""

    def __init__(self, atm, params: Params | None = None, devices=(0,), mode=abi.MODE_FAST):
        self.atm = atm
        self.p = params or Params()
        self.mode = mode
        self.gpu = GpuTransport(devices)
        ox = 1.0 / (1.0 - self.p.oblateness)  # :469-471
        self.oblate = (ox, ox, 1.0)
        self.gpu.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront(), self.oblate)
        self.x_max = (self.p.oblateness + 1.0) * 1.3 * atm.rfront[-1]  # :475-479
        self.wl_index = None
        self.depth = 0
        self.emis_total = 0.0
        self._depth_cache = {}       # grid_initialize(2) is done once per wavelength of a run (:209, :141), not once per launch

    def close(self):
        self.gpu.close()

    def set_wavelength(self, l=0):
        """grid_initialize(2) for wavelength index l, then the table upload."""
        a, p = self.atm, self.p
        self.depth = self._cell_depth(l)
        cw = cdf = None
        if p.photon_source == 2:
            vol = cell_volume(a.rfront, a.thetafront(), a.phifront(), self.oblate)
            cw, _, cdf = thermal_tables(self.depth, a.k_abs[l], a.temperature, vol, a.wavelengths[l] * 1e-6,
                                        a.nr, a.ntheta, a.nphi, p.thermal_weight)
            self.emis_total = float(cdf.reshape(a.nphi, a.ntheta, a.nr)[-1, -1, -1])
        self.gpu.set_wavelength(a.k_sca[l], a.k_abs[l], a.uniq[l], a.cell_to_uniq[l], self.depth, cw, cdf)
        self.wl_index = l

    def _cell_depth(self, l):
        if l not in self._depth_cache:
            a, p = self.atm, self.p
            self._depth_cache[l] = cell_depth(a.rfront, a.k_sca[l], a.k_abs[l], a.nr, a.ntheta, a.nphi, p.photon_source, p.ring)
        return self._depth_cache[l]

    def set_all_wavelengths(self, wls=None):
        """grid_initialize(2) for every wavelength of the atmosphere, uploaded as ONE stacked table set
        (artes_gpu_set_wavelengths): launches then pick their wavelength with `wl_index` and the wavelength loop of
        `run` (:132-204) can be one batched launch.  Star source only (the thermal tables stay per wavelength here)."""
        a, p = self.atm, self.p
        wls = list(range(len(a.wavelengths))) if wls is None else list(wls)
        uniq, c2u, off = [], [], 0
        for l in wls:
            uniq.append(a.uniq[l]); c2u.append(np.asarray(a.cell_to_uniq[l]) + off); off += a.uniq[l].shape[0]
        self.depths = [self._cell_depth(l) for l in wls]
        whole = wls == list(range(len(a.wavelengths)))      # the atmosphere's own (n_wl, cells) arrays: no copy
        self.gpu.set_wavelengths(a.k_sca if whole else np.stack([a.k_sca[l] for l in wls]),
                                 a.k_abs if whole else np.stack([a.k_abs[l] for l in wls]), np.concatenate(uniq), np.stack(c2u), self.depths)
        self.wl_index = wls[0]
        self.depth = self.depths[0]
        return wls

    def launch_struct(self, packages, seed=1, photon_id_base=0, det_phi=None, wl_index=0):
        p = self.p
        det_phi = p.det_phi if det_phi is None else det_phi
        return abi.make_launch(
            mode=self.mode, n_photons=int(packages), photon_id_base=int(photon_id_base), seed=int(seed),
            photon_source=p.photon_source, photon_scattering=int(p.photon_scattering), photon_emission=p.photon_emission,
            stellar_direction=int(p.stellar_direction),
            limb_emission=int(p.phase_curve and det_phi * 180.0 / PI >= 170.0),
            flow_global=int(p.flow_global), flow_theta=int(p.flow_theta), nx=p.nx, ny=p.ny, fstop=p.fstop,
            photon_minimum=p.photon_minimum, photon_bias=p.photon_bias, surface_albedo=p.surface_albedo,
            theta_star=p.theta_star, phi_star=p.phi_star, det_theta=p.det_theta, det_phi=det_phi,
            x_max=self.x_max, y_max=self.x_max, wl_index=int(wl_index))

    def radiative_transfer(self, packages, seed=1, total_packages=None, photon_id_base=0, det_phi=None):
        """`call radiative_transfer`: returns detector(l, stokes, iy, ix), photometry(11), raw result."""
        if self.wl_index is None:
            self.set_wavelength(0)
        L = self.launch_struct(packages, seed, photon_id_base, det_phi)
        res = self.gpu.run(L, flows=self.p.flow_global or self.p.flow_theta)
        wl = self.atm.wavelengths[self.wl_index] * 1e-6
        energy = package_energy(self.p, self.atm.rfront, wl, total_packages or packages, self.emis_total)
        det = detector_from_sums(res["det"], energy)
        return det, photometry(det), res
