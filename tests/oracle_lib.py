"""ctypes binding of the CPU oracle (oracle/libartes_oracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from artes_b200.abi import ERR_SLOTS, Launch, Stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libartes_oracle.so")

RNG_MZ, RNG_PHILOX = 0, 1

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build_oracle():
    src = os.path.join(ORACLE_DIR, "artes_oracle.cc")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def _load():
    lib = C.CDLL(build_oracle())
    lib.artes_ref_create.restype = C.c_void_p
    lib.artes_ref_destroy.argtypes = [C.c_void_p]
    lib.artes_ref_set_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _ip, _dp, C.c_double, C.c_double, C.c_double]
    lib.artes_ref_set_wavelength.argtypes = [C.c_void_p, _dp, _dp, C.c_int, _dp, _ip, C.c_int, C.c_void_p, C.c_void_p]
    lib.artes_ref_run.argtypes = [C.c_void_p, C.POINTER(Launch), C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_void_p, C.c_void_p, _up, C.POINTER(Stats)]
    lib.artes_ref_trace.argtypes = [C.c_void_p, C.POINTER(Launch), _dp, C.c_uint64, C.c_int, _ip, _up, C.c_void_p, C.c_int, C.c_void_p]
    lib.artes_ref_cell_face.argtypes = [C.c_void_p, C.c_uint64, _dp, _dp, _ip, _ip, _ip, _dp]
    lib.artes_ref_scatter.argtypes = [C.c_void_p, C.c_uint64, _dp, _dp, _ip, _dp, _dp]
    lib.artes_ref_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.artes_ref_philox_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_int, _dp]
    lib.artes_ref_mz_uniforms.argtypes = [C.c_int32, C.c_int, _dp]
    lib.artes_ref_cell_depth.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.artes_ref_planck.restype = C.c_double
    lib.artes_ref_planck.argtypes = [C.c_double, C.c_double, C.c_int]
    lib.artes_ref_package_energy.restype = C.c_double
    lib.artes_ref_package_energy.argtypes = [C.c_int] + [C.c_double] * 7 + [C.c_int, C.c_double, C.c_double]
    lib.artes_ref_finish_detector.argtypes = [C.c_int, C.c_int, _dp, C.c_double, _dp, _dp]
    lib.artes_ref_stokes_error.argtypes = [C.c_int, C.c_int, _dp, _dp]
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _load()
    return _LIB


class Oracle:
    """Same method names as artes_b200.lib.GpuTransport so parity tests read symmetrically."""

    def __init__(self):
        self.lib = lib()
        self.h = C.c_void_p(self.lib.artes_ref_create())
        self.cells = 0
        self.nr = self.ntheta = self.nphi = 0

    def close(self):
        if self.h:
            self.lib.artes_ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_grid(self, rfront, thetafront, thetaplane, phifront, oblate=(1.0, 1.0, 1.0)):
        rfront = np.ascontiguousarray(rfront, dtype=np.float64)
        thetafront = np.ascontiguousarray(thetafront, dtype=np.float64)
        thetaplane = np.ascontiguousarray(thetaplane, dtype=np.int32)
        phifront = np.ascontiguousarray(phifront, dtype=np.float64)
        self.nr, self.ntheta, self.nphi = len(rfront) - 1, len(thetafront) - 1, len(phifront)
        self.cells = self.nr * self.ntheta * self.nphi
        rc = self.lib.artes_ref_set_grid(self.h, self.nr, self.ntheta, self.nphi, rfront, thetafront, thetaplane,
                                         phifront, *[float(o) for o in oblate])
        assert rc == 0, rc

    def set_wavelength(self, k_sca, k_abs, uniq, cell_to_uniq, cell_depth, cell_weight=None, emis_cdf=None):
        k_sca = np.ascontiguousarray(k_sca, dtype=np.float64)
        k_abs = np.ascontiguousarray(k_abs, dtype=np.float64)
        uniq = np.ascontiguousarray(uniq, dtype=np.float64)
        c2u = np.ascontiguousarray(cell_to_uniq, dtype=np.int32)
        assert k_sca.size == self.cells and k_abs.size == self.cells and c2u.size == self.cells
        cw = ce = None
        if cell_weight is not None:
            self._cw = np.ascontiguousarray(cell_weight, dtype=np.float64)
            self._ce = np.ascontiguousarray(emis_cdf, dtype=np.float64)
            cw, ce = self._cw.ctypes.data, self._ce.ctypes.data
        rc = self.lib.artes_ref_set_wavelength(self.h, k_sca, k_abs, uniq.shape[0], uniq, c2u, int(cell_depth), cw, ce)
        assert rc == 0, rc

    def set_atmosphere(self, atm, l=0, photon_source=1, oblateness=0.0):
        ox = 1.0 / (1.0 - oblateness)
        self.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront(), (ox, ox, 1.0))
        self.set_wavelength(atm.k_sca[l], atm.k_abs[l], atm.uniq[l], atm.cell_to_uniq[l], 0)
        depth = self.cell_depth(photon_source)
        self.set_wavelength(atm.k_sca[l], atm.k_abs[l], atm.uniq[l], atm.cell_to_uniq[l], depth)
        return depth

    def cell_depth(self, photon_source=1, ring=False):
        return self.lib.artes_ref_cell_depth(self.h, photon_source, int(ring))

    def run(self, launch, rng=RNG_PHILOX, nthreads=0, emulate_stat=False, flows=False):
        det = np.zeros(launch.nx * launch.ny * 12)
        flux = np.zeros(2)
        err = np.zeros(ERR_SLOTS, dtype=np.uint64)
        st = Stats()
        f4 = f3 = None
        p4 = p3 = None
        if flows:
            f4 = np.zeros(4 * self.cells)
            f3 = np.zeros(3 * self.cells)
            p4, p3 = f4.ctypes.data, f3.ctypes.data
        rc = self.lib.artes_ref_run(self.h, C.byref(launch), rng, nthreads, int(emulate_stat), det, flux, p4, p3, err, C.byref(st))
        assert rc == 0, rc
        out = dict(det=det.reshape(3, 4, launch.ny, launch.nx), flux=flux, err=err, stats=st.as_dict())
        if flows:
            out["flow4"], out["flow3"] = f4, f3
        return out

    def trace(self, launch, xi, max_rec=0):
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        n, max_draws = xi.shape
        seq_len = np.zeros(n, dtype=np.int32)
        seq_hash = np.zeros(n, dtype=np.uint64)
        head = np.full((n, max_rec, 5), -1, dtype=np.int32) if max_rec else None
        fstate = np.zeros((n, 8))
        rc = self.lib.artes_ref_trace(self.h, C.byref(launch), xi, n, max_draws, seq_len, seq_hash,
                                      head.ctypes.data if max_rec else None, max_rec, fstate.ctypes.data)
        assert rc == 0, rc
        return dict(len=seq_len, hash=seq_hash, head=head, fstate=fstate)

    def cell_face(self, pos, dirs, face, cell):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        dirs = np.ascontiguousarray(dirs, dtype=np.float64)
        face = np.ascontiguousarray(face, dtype=np.int32)
        cell = np.ascontiguousarray(cell, dtype=np.int32)
        n = pos.shape[0]
        oi = np.zeros((n, 7), dtype=np.int32)
        od = np.zeros(n)
        rc = self.lib.artes_ref_cell_face(self.h, n, pos, dirs, face, cell, oi, od)
        assert rc == 0, rc
        return oi, od

    def scatter(self, stokes, dirs, cell_idx, xi):
        stokes = np.ascontiguousarray(stokes, dtype=np.float64)
        dirs = np.ascontiguousarray(dirs, dtype=np.float64)
        cell_idx = np.ascontiguousarray(cell_idx, dtype=np.int32)
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        out = np.zeros((stokes.shape[0], 9))
        rc = self.lib.artes_ref_scatter(self.h, stokes.shape[0], stokes, dirs, cell_idx, xi, out)
        assert rc == 0, rc
        return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().artes_ref_philox(c, k, o)
    return list(o)


def philox_uniforms(seed, pid, n):
    out = np.zeros(n)
    lib().artes_ref_philox_uniforms(seed, pid, n, out)
    return out


def mz_uniforms(s1, n):
    out = np.zeros(n)
    lib().artes_ref_mz_uniforms(s1, n, out)
    return out


def package_energy(p, rfront, wavelength, packages, emis_total=0.0):
    """photon_package :2509-2539 for an artes_b200.host.Params-like object (wavelength in metres)."""
    return lib().artes_ref_package_energy(int(p.photon_source), p.t_star, p.r_star, p.orbit, p.distance_planet, float(rfront[-1]),
                                          float(wavelength), float(packages), int(p.phase_curve), float(p.det_phi), float(emis_total))


def finish_detector(det_sum, energy):
    """:957-1004 on det_sum[l, stokes, iy, ix] -> detector (same layout), photometry(11)."""
    d = np.ascontiguousarray(det_sum, dtype=np.float64)
    ny, nx = d.shape[-2:]
    out = np.zeros_like(d)
    phot = np.zeros(11)
    lib().artes_ref_finish_detector(nx, ny, d.ravel(), float(energy), out.reshape(-1), phot)
    return out, phot


def stokes_error(detector):
    """write_output :3481-3519 -> error[5, iy, ix]."""
    d = np.ascontiguousarray(detector, dtype=np.float64)
    ny, nx = d.shape[-2:]
    err = np.zeros((5, ny, nx))
    lib().artes_ref_stokes_error(nx, ny, d.ravel(), err.reshape(-1))
    return err
