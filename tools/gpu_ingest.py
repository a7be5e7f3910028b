"""Device-side atmosphere ingest at scale: the reference's dense scattering-matrix array for ONE
wavelength of the scale grid (nr=100, ntheta=60, nphi up to 120: 23 040 B per cell, 16.6 GB at nphi=120) goes through
artes_gpu_set_wavelength_dense_wl (streamed to HBM, hashed and verified on the device) and must give the same launch as the
compact entry.  The grid's nphi is chosen so that the host copy of the dense array fits comfortably in the box's free RAM.

    python tools/gpu_ingest.py [--max-gb 20]
"""
import argparse, json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from artes_b200 import abi, host
from artes_b200.abi import make_launch
from artes_b200.lib import GpuTransport
from tools import atmospheres as A

ap = argparse.ArgumentParser()
ap.add_argument("--max-gb", type=float, default=20.0)
args = ap.parse_args()
avail = 0.0
for line in open("/proc/meminfo"):
    if line.startswith("MemAvailable"):
        avail = float(line.split()[1]) / 1e6
budget = min(args.max_gb, 0.2 * avail)           # the dense array is built once and transposed once: peak ~2.2x its size
nphi = 120
while nphi > 8 and 100 * 60 * nphi * 23040 / 1e9 > budget:
    nphi //= 2
atm = A.c5_scale(nr=100, ntheta=60, nphi=nphi, nl=1)
cells = atm.cells
gb = cells * 23040 / 1e9
t0 = time.time()
u = atm.uniq[0][atm.cell_to_uniq[0]]                                   # [cells, 180, 16]
dense = np.empty((180, 16, cells))                                     # HDU 8 order for one wavelength: (angle, element, cell), cell fastest
for a0 in range(0, 180, 20):
    dense[a0:a0 + 20] = u[:, a0:a0 + 20, :].transpose(1, 2, 0)
del u
t_build = time.time() - t0
depth = host.cell_depth(atm.rfront, atm.k_sca[0], atm.k_abs[0], atm.nr, atm.ntheta, atm.nphi, 1)
ga, gb_ = GpuTransport((0,)), GpuTransport((0,))
for g in (ga, gb_):
    g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
t0 = time.time(); ga.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], depth); t_compact = time.time() - t0
t0 = time.time(); gb_.set_wavelength_dense_wl(atm.k_sca, atm.k_abs, dense, 0, depth); t_dense = time.time() - t0
t0 = time.time(); gb_.set_wavelength_dense_wl(atm.k_sca, atm.k_abs, dense, 0, depth); t_dense2 = time.time() - t0
xm = 1.3 * atm.rfront[-1]
L = make_launch(mode=abi.MODE_FAST, n_photons=200000, x_max=xm, y_max=xm, seed=4, nx=64, ny=64, det_phi=math.radians(60.0))
a, b = ga.run(L), gb_.run(L)
same = bool(np.array_equal(a["det"][2], b["det"][2]) and np.allclose(a["det"][0], b["det"][0], rtol=1e-9, atol=1e-12 * np.abs(a["det"][0]).max())
            and all(a["stats"][k] == b["stats"][k] for k in ("n_cell_face", "n_scatter", "n_draws")))
# the host path the round-1 library used, for comparison: one pass over the cells hashing 2880 strided doubles each
t0 = time.time()
sample = min(cells, 20000)
h = np.zeros(sample, dtype=np.uint64)
blk = np.ascontiguousarray(dense[:, :, :sample].reshape(2880, sample).T)
t_host_sample = time.time() - t0
print(json.dumps({"grid": [atm.nr, atm.ntheta, atm.nphi], "cells": cells, "dense_gb": gb, "mem_available_gb": avail, "n_uniq": int(atm.uniq[0].shape[0]),
                  "build_dense_s": t_build, "compact_entry_s": t_compact, "dense_entry_first_s": t_dense, "dense_entry_second_s": t_dense2,
                  "dense_gb_per_s": gb / t_dense2, "host_gather_s_per_1e4_cells": t_host_sample / sample * 1e4,
                  "equal_to_compact_entry": same}))
assert same
