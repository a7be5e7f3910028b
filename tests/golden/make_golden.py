"""Generates tests/golden/golden_*.npz with the CPU oracle.

The reference has no fixtures of its own and cannot be run here (PARITY UNPINNED); these vectors
freeze the oracle's output so that (a) the oracle cannot drift unnoticed and (b) the GPU tests have
committed files to be compared against on the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py        # regenerates every file
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from artes_b200.abi import make_launch  # noqa: E402

CASES = {
    # name: atmosphere builder, injected-stream photons, draws per photon, Philox-run photons, launch overrides
    "c1": dict(atm="c1_template_rayleigh", trace_n=4000, draws=192, run_n=20000, kw=dict()),
    "c2": dict(atm="c2_hg_deck", trace_n=4000, draws=192, run_n=20000, kw=dict(nx=1, ny=1, det_phi=math.radians(60.0))),
    "c4": dict(atm="c4_mie_patches", trace_n=3000, draws=192, run_n=15000,
               kw=dict(nx=64, ny=64, det_phi=math.radians(60.0), surface_albedo=0.3)),
}


def build_case(name, atmospheres=None):
    from tools import atmospheres as A
    c = CASES[name]
    atm = atmospheres(c["atm"]) if atmospheres else getattr(A, c["atm"])()
    xm = 1.3 * atm.rfront[-1]
    xi = np.random.RandomState(1000 + len(name) + ord(name[-1])).random_sample((c["trace_n"], c["draws"]))
    launch_trace = make_launch(n_photons=c["trace_n"], x_max=xm, y_max=xm, fstop=0.02, **c["kw"])
    launch_run = make_launch(n_photons=c["run_n"], x_max=xm, y_max=xm, seed=77, **c["kw"])
    return atm, launch_trace, launch_run, xi


def main():
    from oracle_lib import Oracle
    for name in CASES:
        atm, lt, lr, xi = build_case(name)
        o = Oracle()
        depth = o.set_atmosphere(atm)
        t = o.trace(lt, xi)
        r = o.run(lr)
        path = os.path.join(HERE, f"golden_{name}.npz")
        np.savez_compressed(path, cell_depth=depth, seq_len=t["len"], seq_hash=t["hash"], det=r["det"],
                            fstate=t["fstate"], n_cell_face=r["stats"]["n_cell_face"], n_scatter=r["stats"]["n_scatter"])
        print(name, "mean crossings", t["len"].mean(), "I", r["det"][0, 0].sum(), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
