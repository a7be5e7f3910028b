"""ctypes binding of libartes_gpu.so (include/artes_gpu.h).

This is the product path: it fails loudly when the CUDA library is missing or no device is usable.
There is no CPU fallback and nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .abi import ABI_VERSION, ERR_SLOTS, NCCL_ID_BYTES, Launch, Stats

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARTES_GPU_LIB") or os.path.join(_HERE, "libartes_gpu.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")

# every symbol include/artes_gpu.h declares
SYMBOLS = [
    "artes_gpu_create", "artes_gpu_destroy", "artes_gpu_last_error", "artes_gpu_abi_version",
    "artes_gpu_set_grid", "artes_gpu_set_wavelength", "artes_gpu_set_wavelength_dense", "artes_gpu_set_wavelength_dense_wl",
    "artes_gpu_set_wavelengths", "artes_gpu_run", "artes_gpu_run_batch", "artes_gpu_run_multi", "artes_gpu_run_async", "artes_gpu_wait", "artes_gpu_nccl_unique_id",
    "artes_gpu_nccl_init_rank", "artes_gpu_trace", "artes_gpu_cell_face", "artes_gpu_device_info",
    "artes_gpu_fma_peak", "artes_gpu_last_engine", "artes_gpu_test_ingest_chunk", "artes_gpu_test_full_matrix",
]


class ArtesGpuError(RuntimeError):
    pass


_LIB = None


def load():
    """Load libartes_gpu.so (built by `make -C artes_b200/csrc` / __graft_entry__.build())."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ArtesGpuError(f"{LIB_PATH} is missing: build it with `make -C artes_b200/csrc` "
                            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.artes_gpu_last_error.restype = C.c_char_p
    lib.artes_gpu_last_error.argtypes = [C.c_void_p]
    lib.artes_gpu_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    lib.artes_gpu_destroy.argtypes = [C.c_void_p]
    lib.artes_gpu_set_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _ip, _dp,
                                       C.c_double, C.c_double, C.c_double]
    lib.artes_gpu_set_wavelength.argtypes = [C.c_void_p, _dp, _dp, C.c_int, _dp, _ip, C.c_int, C.c_void_p, C.c_void_p]
    lib.artes_gpu_set_wavelengths.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp, _ip, _ip, C.c_void_p, C.c_void_p]
    lib.artes_gpu_set_wavelength_dense.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int, C.c_void_p, C.c_void_p]
    lib.artes_gpu_set_wavelength_dense_wl.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_void_p, C.c_void_p]
    lib.artes_gpu_run.argtypes = [C.c_void_p, C.POINTER(Launch), _dp, _dp, C.c_void_p, C.c_void_p, _up, C.POINTER(Stats)]
    lib.artes_gpu_run_async.argtypes = [C.c_void_p, C.POINTER(Launch)]
    lib.artes_gpu_run_batch.argtypes = [C.c_void_p, C.POINTER(Launch), C.c_int, _dp, _dp, _up, C.POINTER(Stats)]
    lib.artes_gpu_run_multi.argtypes = [C.c_void_p, C.POINTER(Launch), C.c_int, _dp, _dp, _up, C.POINTER(Stats)]
    lib.artes_gpu_wait.argtypes = [C.c_void_p, _dp, _dp, C.c_void_p, C.c_void_p, _up, C.POINTER(Stats)]
    lib.artes_gpu_nccl_unique_id.argtypes = [C.c_void_p]
    lib.artes_gpu_nccl_init_rank.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.artes_gpu_trace.argtypes = [C.c_void_p, C.POINTER(Launch), _dp, C.c_uint64, C.c_int, _ip, _up,
                                    C.c_void_p, C.c_int, C.c_void_p]
    lib.artes_gpu_cell_face.argtypes = [C.c_void_p, C.c_int, C.c_uint64, _dp, _dp, _ip, _ip, _ip, _dp]
    lib.artes_gpu_device_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.c_char_p, C.c_int]
    lib.artes_gpu_fma_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.artes_gpu_last_engine.argtypes = [C.c_void_p]
    lib.artes_gpu_test_ingest_chunk.argtypes = [C.c_int]
    lib.artes_gpu_test_full_matrix.argtypes = [C.c_int]
    if lib.artes_gpu_abi_version() != ABI_VERSION:
        raise ArtesGpuError("libartes_gpu.so ABI version mismatch")
    _LIB = lib
    return lib


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    rc = load().artes_gpu_nccl_unique_id(buf)
    if rc:
        raise ArtesGpuError(f"artes_gpu_nccl_unique_id failed ({rc}): {load().artes_gpu_last_error(None).decode()}")
    return buf.raw


class GpuTransport:
    """Thin object wrapper of the C-ABI handle: one context = the devices of this process."""

    def __init__(self, devices=(0,)):
        self.lib = load()
        self.h = C.c_void_p()
        devs = list(devices)
        arr = (C.c_int * len(devs))(*devs)
        rc = self.lib.artes_gpu_create(C.byref(self.h), len(devs), arr)
        if rc:
            raise ArtesGpuError(f"artes_gpu_create failed ({rc}): {self.lib.artes_gpu_last_error(None).decode()}")
        self.cells = 0
        self.nr = self.ntheta = self.nphi = 0
        self._keep = []

    def _check(self, rc, what):
        if rc:
            raise ArtesGpuError(f"{what} failed ({rc}): {self.lib.artes_gpu_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.artes_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs ---------------------------------------------------------------------------
    def set_grid(self, rfront, thetafront, thetaplane, phifront, oblate=(1.0, 1.0, 1.0)):
        rfront = np.ascontiguousarray(rfront, dtype=np.float64)
        thetafront = np.ascontiguousarray(thetafront, dtype=np.float64)
        thetaplane = np.ascontiguousarray(thetaplane, dtype=np.int32)
        phifront = np.ascontiguousarray(phifront, dtype=np.float64)
        self.nr, self.ntheta, self.nphi = len(rfront) - 1, len(thetafront) - 1, len(phifront)
        self.cells = self.nr * self.ntheta * self.nphi
        self._check(self.lib.artes_gpu_set_grid(self.h, self.nr, self.ntheta, self.nphi, rfront, thetafront,
                                                thetaplane, phifront, *[float(o) for o in oblate]), "artes_gpu_set_grid")

    def set_wavelength(self, k_sca, k_abs, uniq, cell_to_uniq, cell_depth, cell_weight=None, emis_cdf=None):
        k_sca = np.ascontiguousarray(k_sca, dtype=np.float64)
        k_abs = np.ascontiguousarray(k_abs, dtype=np.float64)
        uniq = np.ascontiguousarray(uniq, dtype=np.float64)
        c2u = np.ascontiguousarray(cell_to_uniq, dtype=np.int32)
        if not (k_sca.size == self.cells and k_abs.size == self.cells and c2u.size == self.cells):
            raise ValueError("per-cell arrays do not match the grid")
        cw = ce = None
        if cell_weight is not None:
            a = np.ascontiguousarray(cell_weight, dtype=np.float64)
            b = np.ascontiguousarray(emis_cdf, dtype=np.float64)
            self._keep = [a, b]
            cw, ce = a.ctypes.data, b.ctypes.data
        self._check(self.lib.artes_gpu_set_wavelength(self.h, k_sca, k_abs, uniq.shape[0], uniq, c2u, int(cell_depth), cw, ce),
                    "artes_gpu_set_wavelength")

    def set_wavelengths(self, k_sca, k_abs, uniq, cell_to_uniq, cell_depths, cell_weight=None, emis_cdf=None):
        """Tables of several wavelengths at once: k_sca, k_abs, cell_to_uniq are (n_wl, cells); uniq is ONE common list."""
        k_sca = np.ascontiguousarray(k_sca, dtype=np.float64)
        k_abs = np.ascontiguousarray(k_abs, dtype=np.float64)
        uniq = np.ascontiguousarray(uniq, dtype=np.float64)
        c2u = np.ascontiguousarray(cell_to_uniq, dtype=np.int32)
        depths = np.ascontiguousarray(cell_depths, dtype=np.int32)
        n_wl = depths.size
        if not (k_sca.size == n_wl * self.cells and k_abs.size == n_wl * self.cells and c2u.size == n_wl * self.cells):
            raise ValueError("per-cell arrays do not match the grid x wavelengths")
        cw = ce = None
        if cell_weight is not None:      # thermal source: (n_wl, cells) each
            a = np.ascontiguousarray(cell_weight, dtype=np.float64)
            b = np.ascontiguousarray(emis_cdf, dtype=np.float64)
            if a.size != n_wl * self.cells or b.size != n_wl * self.cells:
                raise ValueError("thermal tables do not match the grid x wavelengths")
            self._keep = [a, b]
            cw, ce = a.ctypes.data, b.ctypes.data
        self._check(self.lib.artes_gpu_set_wavelengths(self.h, n_wl, k_sca, k_abs, uniq.shape[0], uniq, c2u, depths, cw, ce),
                    "artes_gpu_set_wavelengths")

    def set_wavelength_dense(self, k_sca, k_abs, dense, cell_depth, cell_weight=None, emis_cdf=None):
        """dense: numpy array (180, 16, nphi, ntheta, nr) = HDU 8 of atmosphere.fits for one wavelength."""
        k_sca = np.ascontiguousarray(k_sca, dtype=np.float64)
        k_abs = np.ascontiguousarray(k_abs, dtype=np.float64)
        dense = np.ascontiguousarray(dense, dtype=np.float64)
        if dense.size != self.cells * 2880:
            raise ValueError("dense matrix does not match the grid")
        cw = ce = None
        if cell_weight is not None:
            a = np.ascontiguousarray(cell_weight, dtype=np.float64)
            b = np.ascontiguousarray(emis_cdf, dtype=np.float64)
            self._keep = [a, b]
            cw, ce = a.ctypes.data, b.ctypes.data
        self._check(self.lib.artes_gpu_set_wavelength_dense(self.h, k_sca, k_abs, dense, int(cell_depth), cw, ce),
                    "artes_gpu_set_wavelength_dense")

    def set_wavelength_dense_wl(self, k_sca_all, k_abs_all, dense_all, wl_index, cell_depth, cell_weight=None, emis_cdf=None):
        """The reference's whole arrays: k_sca_all / k_abs_all numpy (n_wl, nphi, ntheta, nr), dense_all numpy
        (180, 16, n_wl, nphi, ntheta, nr) = HDUs 6-8 of atmosphere.fits; wavelength wl_index is picked with strides and
        de-duplicated on the device."""
        k_sca_all = np.ascontiguousarray(k_sca_all, dtype=np.float64)
        k_abs_all = np.ascontiguousarray(k_abs_all, dtype=np.float64)
        dense_all = np.ascontiguousarray(dense_all, dtype=np.float64)
        n_wl = k_sca_all.size // self.cells
        if dense_all.size != self.cells * 2880 * n_wl or k_abs_all.size != k_sca_all.size:
            raise ValueError("dense arrays do not match the grid x wavelengths")
        cw = ce = None
        if cell_weight is not None:
            a = np.ascontiguousarray(cell_weight, dtype=np.float64)
            b = np.ascontiguousarray(emis_cdf, dtype=np.float64)
            self._keep = [a, b]
            cw, ce = a.ctypes.data, b.ctypes.data
        self._check(self.lib.artes_gpu_set_wavelength_dense_wl(self.h, n_wl, int(wl_index), k_sca_all, k_abs_all, dense_all,
                                                               int(cell_depth), cw, ce), "artes_gpu_set_wavelength_dense_wl")

    # ---- the hot path ----------------------------------------------------------------------
    def _outputs(self, launch, flows):
        det = np.zeros(launch.nx * launch.ny * 12)
        flux = np.zeros(2)
        err = np.zeros(ERR_SLOTS, dtype=np.uint64)
        f4 = np.zeros(4 * self.cells) if flows else None
        f3 = np.zeros(3 * self.cells) if flows else None
        return det, flux, err, f4, f3

    def _result(self, launch, det, flux, err, f4, f3, st):
        out = dict(det=det.reshape(3, 4, launch.ny, launch.nx), flux=flux, err=err, stats=st.as_dict())
        if f4 is not None:
            out["flow4"], out["flow3"] = f4, f3
        return out

    def run(self, launch, flows=False):
        det, flux, err, f4, f3 = self._outputs(launch, flows)
        st = Stats()
        self._check(self.lib.artes_gpu_run(self.h, C.byref(launch), det, flux,
                                           f4.ctypes.data if flows else None, f3.ctypes.data if flows else None,
                                           err, C.byref(st)), "artes_gpu_run")
        return self._result(launch, det, flux, err, f4, f3, st)

    def run_multi(self, launches):
        """ONE walk observed by all detectors of `launches` (include/artes_gpu.h: artes_gpu_run_multi)."""
        return self.run_batch(launches, multi=True)

    def run_batch(self, launches, multi=False):
        """Launches that differ only in the detector direction, as one kernel launch (include/artes_gpu.h)."""
        n = len(launches)
        arr = (Launch * n)(*launches)
        L0 = launches[0]
        det = np.zeros(n * L0.nx * L0.ny * 12)
        flux = np.zeros(2 * n)
        err = np.zeros(ERR_SLOTS, dtype=np.uint64)
        st = Stats()
        fn = self.lib.artes_gpu_run_multi if multi else self.lib.artes_gpu_run_batch
        self._check(fn(self.h, arr, n, det, flux, err, C.byref(st)), "artes_gpu_run_multi" if multi else "artes_gpu_run_batch")
        return dict(det=det.reshape(n, 3, 4, L0.ny, L0.nx), flux=flux.reshape(n, 2), err=err, stats=st.as_dict())

    def run_async(self, launch):
        self._pending = launch
        self._check(self.lib.artes_gpu_run_async(self.h, C.byref(launch)), "artes_gpu_run_async")

    def wait(self, flows=False):
        launch = self._pending
        det, flux, err, f4, f3 = self._outputs(launch, flows)
        st = Stats()
        self._check(self.lib.artes_gpu_wait(self.h, det, flux, f4.ctypes.data if flows else None,
                                            f3.ctypes.data if flows else None, err, C.byref(st)), "artes_gpu_wait")
        return self._result(launch, det, flux, err, f4, f3, st)

    def nccl_init_rank(self, nranks, rank, uid: bytes):
        buf = C.create_string_buffer(uid, NCCL_ID_BYTES)
        self._check(self.lib.artes_gpu_nccl_init_rank(self.h, nranks, rank, buf), "artes_gpu_nccl_init_rank")

    # ---- test hooks -------------------------------------------------------------------------
    def trace(self, launch, xi, max_rec=0):
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        n, max_draws = xi.shape
        seq_len = np.zeros(n, dtype=np.int32)
        seq_hash = np.zeros(n, dtype=np.uint64)
        head = np.full((n, max_rec, 5), -1, dtype=np.int32) if max_rec else None
        fstate = np.zeros((n, 8))
        self._check(self.lib.artes_gpu_trace(self.h, C.byref(launch), xi, n, max_draws, seq_len, seq_hash,
                                             head.ctypes.data if max_rec else None, max_rec, fstate.ctypes.data),
                    "artes_gpu_trace")
        return dict(len=seq_len, hash=seq_hash, head=head, fstate=fstate)

    def cell_face(self, pos, dirs, face, cell, mode=0):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        dirs = np.ascontiguousarray(dirs, dtype=np.float64)
        face = np.ascontiguousarray(face, dtype=np.int32)
        cell = np.ascontiguousarray(cell, dtype=np.int32)
        n = pos.shape[0]
        oi = np.zeros((n, 7), dtype=np.int32)
        od = np.zeros(n)
        self._check(self.lib.artes_gpu_cell_face(self.h, mode, n, pos, dirs, face, cell, oi, od), "artes_gpu_cell_face")
        return oi, od

    def device_info(self):
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        name = C.create_string_buffer(256)
        self._check(self.lib.artes_gpu_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), name, 256),
                    "artes_gpu_device_info")
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), name=name.value.decode())

    def last_engine(self):
        """1 = persistent-lane engine, 2 = ray/event engine (the last run / trace of this context)."""
        return int(self.lib.artes_gpu_last_engine(self.h))

    def fma_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.artes_gpu_fma_peak(self.h, C.byref(a), C.byref(b)), "artes_gpu_fma_peak")
        return dict(fp64_tflops=a.value, fp32_tflops=b.value)
