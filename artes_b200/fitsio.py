"""Minimal FITS image reader/writer (numpy only) for the two file contracts of the path:

  * input  `atmosphere.fits`: 9 float64 image HDUs (python/atmosphere.py:449-459, read with
    CFITSIO at src/ARTES.f90:2067-2201);
  * output `stokes.fits` / `error.fits` / `flow_*.fits`: one primary HDU, BITPIX -64
    (src/ARTES.f90:3774-3841).

The reference links a vendored CFITSIO binary without headers; this module (and its C++ twin in
src/host/fits_min.cc) only implements what those call sites need: 2880-byte blocks, 80-char cards,
big-endian data, BITPIX 8/16/32/64/-32/-64, no scaling, no tables.
"""
from __future__ import annotations

import numpy as np

BLOCK = 2880
_DTYPES = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_BITPIX = {"u1": 8, "i2": 16, "i4": 32, "i8": 64, "f4": -32, "f8": -64}


def _card(key, value=None, comment=""):
    if value is None:
        s = key
    elif isinstance(value, bool):
        s = f"{key:<8}= {('T' if value else 'F'):>20}"
    elif isinstance(value, int):
        s = f"{key:<8}= {value:>20d}"
    elif isinstance(value, float):
        s = f"{key:<8}= {value:>20.13E}"
    else:
        s = f"{key:<8}= '{str(value):<8}'"
    if comment:
        s += f" / {comment}"
    return s[:80].ljust(80)


def _header(arr, primary, name):
    cards = []
    if primary:
        cards.append(_card("SIMPLE", True, "file does conform to FITS standard"))
    else:
        cards.append(_card("XTENSION", "IMAGE", "Image extension"))
    cards.append(_card("BITPIX", _BITPIX[arr.dtype.str[1:]], "number of bits per data pixel"))
    cards.append(_card("NAXIS", arr.ndim, "number of data axes"))
    for i, n in enumerate(reversed(arr.shape)):  # FITS axis order = reversed numpy shape
        cards.append(_card(f"NAXIS{i + 1}", int(n), f"length of data axis {i + 1}"))
    if primary:
        # the primary header ftphpr of the vendored CFITSIO 3.34 writes (write_fits_3D/4D, src/ARTES.f90:3774-3841), byte for byte
        cards.append(_card("EXTEND", True, "FITS dataset may contain extensions"))
        cards.append("COMMENT   FITS (Flexible Image Transport System) format is defined in 'Astronomy".ljust(80))
        cards.append("COMMENT   and Astrophysics', volume 376, page 359; bibcode: 2001A&A...376..359H".ljust(80))
    else:
        cards.append(_card("PCOUNT", 0))
        cards.append(_card("GCOUNT", 1))
    if name:
        cards.append(_card("EXTNAME", name.upper()))
    cards.append(_card("END"))
    h = "".join(cards)
    h += " " * (-len(h) % BLOCK)
    return h.encode("ascii")


def write_hdus(path, hdus):
    """hdus: list of (name, ndarray); the first becomes the primary HDU."""
    with open(path, "wb") as f:
        for i, (name, arr) in enumerate(hdus):
            arr = np.asarray(arr)
            if arr.dtype.kind == "f" and arr.dtype.itemsize != 4:
                arr = arr.astype(np.float64)
            f.write(_header(arr, i == 0, name))
            data = np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder(">"), copy=False).tobytes()
            f.write(data)
            f.write(b"\0" * (-len(data) % BLOCK))


def write_image(path, arr):
    """Single primary HDU, float64: what write_fits_3D/4D produce (src/ARTES.f90:3774-3841).
    `arr` is given in Fortran index order (n1 fastest): pass a numpy array of shape (n3, n2, n1)."""
    write_hdus(path, [("", np.asarray(arr, dtype=np.float64))])


def _parse_header(f):
    cards = {}
    order = []
    while True:
        block = f.read(BLOCK)
        if len(block) < BLOCK:
            return None, None
        done = False
        for i in range(0, BLOCK, 80):
            c = block[i:i + 80].decode("ascii", "replace")
            key = c[:8].strip()
            if key == "END":
                done = True
                break
            if c[8:10] == "= ":
                v = c[10:].split("/")[0].strip()
                if v.startswith("'"):
                    val = v.strip("'").strip()
                elif v in ("T", "F"):
                    val = v == "T"
                else:
                    try:
                        val = int(v)
                    except ValueError:
                        try:
                            val = float(v.replace("D", "E"))
                        except ValueError:
                            val = v
                cards[key] = val
                order.append(key)
        if done:
            return cards, order


def read_hdus(path):
    """Returns a list of (header dict, ndarray) for every image HDU of the file."""
    out = []
    with open(path, "rb") as f:
        while True:
            hdr, _ = _parse_header(f)
            if hdr is None:
                break
            naxis = hdr.get("NAXIS", 0)
            shape = tuple(hdr[f"NAXIS{i}"] for i in range(naxis, 0, -1))
            n = int(np.prod(shape)) if naxis else 0
            dt = np.dtype(_DTYPES[hdr["BITPIX"]])
            nbytes = n * dt.itemsize
            raw = f.read(nbytes)
            f.seek((-nbytes) % BLOCK, 1)
            arr = np.frombuffer(raw, dtype=dt).astype(dt.newbyteorder("=")).reshape(shape) if n else np.zeros(shape)
            out.append((hdr, arr))
    return out
