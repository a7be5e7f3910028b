"""Output stage: the files bin/ARTES writes are byte-compatible with what the reference's write_output
(src/ARTES.f90:3472-3772) produces.

  * FITS images (stokes.fits, error.fits, cell_luminosity.fits, flow_*.fits): write_fits_3D/4D (:3774-3841) are four calls
    of the vendored CFITSIO 3.34.  The driver's writer (src/host/fits_min.cc, through bin/fits_tool) and the Python mirror
    must produce the same BYTES as those calls -- against golden files made with the reference's own library
    (tests/golden/make_cfitsio_golden.py) and, where /root/reference is present, against the library live; and the library
    must read the driver's files back (keywords, axis order, pixel values).
  * Text tables (phase.dat, photometry.dat, spectrum.dat, normalization.dat, ...): Fortran list-directed records
    (`write (100,*)`, :3525-3709) in gfortran's layout.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from artes_b200 import fitsio
from golden.make_cfitsio_golden import CASES, LIB, case_data, cfitsio_write

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TOOL = os.path.join(ROOT, "bin", "fits_tool")


@pytest.fixture(scope="module")
def tool():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "src", "host"), os.path.join("..", "..", "bin", "fits_tool")], stdout=subprocess.DEVNULL)
    return TOOL


def tool_write(tool, path, naxes, data):
    subprocess.run([tool, "write", str(path)] + [str(n) for n in naxes], input=np.ascontiguousarray(data, dtype="<f8").tobytes(), check=True)


@pytest.mark.parametrize("name", sorted(CASES))
def test_fits_writers_reproduce_the_cfitsio_bytes(tool, tmp_path, name):
    """Driver writer and Python mirror against the golden file the reference's CFITSIO calls produced: identical bytes
    (header cards with their comments, the two COMMENT cards of ftphpr, big-endian pixels incl. +-0, +-inf, denormals, padding)."""
    naxes = CASES[name]
    data = case_data(name, naxes)
    gold = open(os.path.join(GOLDEN, f"cfitsio_{name}.fits"), "rb").read()
    out = tmp_path / "a.fits"
    tool_write(tool, out, naxes, data)
    assert open(out, "rb").read() == gold
    out2 = tmp_path / "b.fits"
    fitsio.write_image(str(out2), data.reshape(tuple(reversed(naxes))))      # numpy shape = reversed FITS axes
    assert open(out2, "rb").read() == gold
    # and the readers give the array back in the reference's index order
    name0, arr = fitsio.read_hdus(str(out))[0]
    assert arr.shape == tuple(reversed(naxes))
    np.testing.assert_array_equal(arr.ravel(), data)


@pytest.mark.skipif(not os.path.exists(LIB), reason="the vendored CFITSIO of /root/reference is not on this machine")
def test_vendored_cfitsio_live_roundtrip(tool, tmp_path):
    """Live against lib/libcfitsio.so.3: (a) its output for the image shapes of write_output equals the driver's bytes,
    (b) it opens the driver's files and returns BITPIX -64, the axes in (nx, ny, plane) order and the pixel values."""
    lib = C.CDLL(LIB)
    rs = np.random.RandomState(3)
    for naxes in ((25, 25, 4), (25, 25, 5), (64, 64, 4), (20, 18, 36), (3, 20, 18, 36), (4, 2, 7, 1), (1, 1, 4)):
        data = rs.standard_normal(int(np.prod(naxes)))
        ours, theirs = tmp_path / "ours.fits", tmp_path / "theirs.fits"
        tool_write(tool, ours, naxes, data)
        cfitsio_write(lib, str(theirs), naxes, data)
        assert open(ours, "rb").read() == open(theirs, "rb").read(), naxes
        fptr, st = C.c_void_p(), C.c_int(0)
        lib.ffopen(C.byref(fptr), str(ours).encode(), 0, C.byref(st))
        assert st.value == 0
        bitpix, naxis = C.c_int(), C.c_int()
        ax = (C.c_long * 8)()
        lib.ffgipr(fptr, 8, C.byref(bitpix), C.byref(naxis), ax, C.byref(st))
        assert (bitpix.value, naxis.value, tuple(ax[:naxis.value])) == (-64, len(naxes), tuple(naxes))
        buf = np.zeros(data.size)
        anynul = C.c_int()
        lib.ffgpvd(fptr, C.c_long(1), C.c_longlong(1), C.c_longlong(data.size), C.c_double(0.0), buf.ctypes.data_as(C.c_void_p), C.byref(anynul), C.byref(st))
        lib.ffclos(fptr, C.byref(st))
        assert st.value == 0
        np.testing.assert_array_equal(buf, data)


LD_CASES = [
    # value -> the record gfortran writes for `write (100,*) value` (REAL(8): one blank + G25.17E3); well-known outputs
    (1.0, "   1.0000000000000000     "),
    (0.7, "  0.69999999999999996     "),
    (1.0e-5, "   1.0000000000000001E-005"),
    (123456.789, "   123456.78900000000     "),
    (0.0, "   0.0000000000000000     "),
    (-2.5, "  -2.5000000000000000     "),
    (180.0, "   180.00000000000000     "),
    (0.1, "  0.10000000000000001     "),
    (1.0e16, "   10000000000000000.     "),
    (1.0e17, "   1.0000000000000000E+017"),
]


def test_list_directed_records_have_gfortran_layout(tool):
    for v, want in LD_CASES:
        got = subprocess.run([tool, "ld", repr(v)], capture_output=True, text=True, check=True).stdout.rstrip("\n")
        assert got == want, (v, got, want)
        assert len(got) == 26 and float(got) == v
    # any value: 26 characters, 17 significant digits (round trip), F layout inside [0.1, 1e17) with the five trailing blanks,
    # otherwise a three-digit exponent
    rs = np.random.RandomState(0)
    for v in np.concatenate([rs.standard_normal(200) * 10.0 ** rs.randint(-40, 40, 200), [5e-324, 1.7976931348623157e308, -0.1, 0.09999999]]):
        got = subprocess.run([tool, "ld", repr(float(v))], capture_output=True, text=True, check=True).stdout.rstrip("\n")
        assert len(got) == 26 and got[0] == " " and float(got) == v, (v, got)
        if 0.1 <= abs(v) < 1e17:
            assert got.endswith("     ") and "E" not in got and len(got.strip().lstrip("-").replace(".", "").lstrip("0")) <= 17, (v, got)
        else:
            assert got[-5] == "E" and got[-4] in "+-" and got[-3:].isdigit() and got.strip().lstrip("-")[1] == ".", (v, got)
    # a photometry.dat record (:3576-3587): nine reals; a cell_depth.dat record (:3689-3709): real + default integer (I12 with the separator)
    rec = subprocess.run([tool, "ld", "0.7", "1e-14", "2e-16", "-3e-15", "1e-16", "0", "0", "0", "0"], capture_output=True, text=True, check=True).stdout.rstrip("\n")
    assert len(rec) == 9 * 26 and rec.split() == ["0.69999999999999996", "1.0000000000000000E-014", "2.0000000000000000E-016", "-2.9999999999999998E-015",
                                                   "9.9999999999999998E-017", "0.0000000000000000", "0.0000000000000000", "0.0000000000000000", "0.0000000000000000"]
    rec = subprocess.run([tool, "ld", "0.7", "i:3"], capture_output=True, text=True, check=True).stdout.rstrip("\n")
    assert rec == "  0.69999999999999996                3"
    hdr = subprocess.run([tool, "ld", "s:# Wavelength [micron] - Stokes I, Q, U, V [W m-2 micron-1]"], capture_output=True, text=True, check=True).stdout.rstrip("\n")
    assert hdr == " # Wavelength [micron] - Stokes I, Q, U, V [W m-2 micron-1]"      # character item: one blank, then the text (:3535)
