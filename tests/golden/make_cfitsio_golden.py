"""Golden FITS files written by the REFERENCE's own output calls: write_fits_3D / write_fits_4D (src/ARTES.f90:3774-3841) are
ftinit / ftphpr(simple=T, bitpix=-64, naxis, naxes, pcount=0, gcount=1, extend=T) / ftpprd / ftclos of the vendored CFITSIO
3.34 (lib/libcfitsio.so.3).  This script makes the same calls through the library's C entry points (ffinit,
ffphpr, ffpprd, ffclos) in THIS container, where /root/reference exists, and commits the resulting bytes as fixtures, so that
tests/test_output_stage.py can compare the driver's writer with them anywhere.

    python tests/golden/make_cfitsio_golden.py
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = "/root/reference/lib/libcfitsio.so.3"
CASES = {"stokes_5x4x4": (5, 4, 4), "error_3x3x5": (3, 3, 5), "flow_3x2x3x2": (3, 2, 3, 2), "lum_361": (361,)}


def case_data(name, naxes):
    rs = np.random.RandomState(sum(map(ord, name)))
    d = rs.standard_normal(int(np.prod(naxes))) * 10.0 ** rs.randint(-30, 30, int(np.prod(naxes)))
    d[:6] = [0.0, -0.0, np.inf, -np.inf, 5e-324, 1.7976931348623157e308]
    return d


def cfitsio_write(lib, path, naxes, data):
    fptr, st = C.c_void_p(), C.c_int(0)
    lib.ffinit(C.byref(fptr), ("!" + path).encode(), C.byref(st))
    ax = (C.c_long * len(naxes))(*naxes)
    lib.ffphpr(fptr, 1, -64, len(naxes), ax, C.c_longlong(0), C.c_longlong(1), 1, C.byref(st))
    d = np.ascontiguousarray(data, dtype=np.float64)
    lib.ffpprd(fptr, C.c_long(1), C.c_longlong(1), C.c_longlong(d.size), d.ctypes.data_as(C.c_void_p), C.byref(st))
    lib.ffclos(fptr, C.byref(st))
    assert st.value == 0, st.value


if __name__ == "__main__":
    lib = C.CDLL(LIB)
    for name, naxes in CASES.items():
        out = os.path.join(HERE, f"cfitsio_{name}.fits")
        cfitsio_write(lib, out, naxes, case_data(name, naxes))
        print(out, os.path.getsize(out))
