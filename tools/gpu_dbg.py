import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from tools import atmospheres as A
from oracle_lib import Oracle
from artes_b200.lib import GpuTransport
from artes_b200.abi import make_launch
atm = A.c4_mie_patches()
obl = 0.06; ox = 1.0 / (1.0 - obl)
o = Oracle(); depth = o.set_atmosphere(atm, oblateness=obl)
g = GpuTransport((0,)); g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront(), (ox, ox, 1.0))
g.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], depth)
xm = 1.06 * 1.3 * atm.rfront[-1]; n = 3000
xi = np.random.RandomState(8).random_sample((n, 160))
L = make_launch(mode=1, n_photons=n, x_max=xm, y_max=xm, fstop=0.03, surface_albedo=0.5)
ro, rg = o.trace(L, xi, max_rec=60), g.trace(L, xi, max_rec=60)
bad = np.where((ro["hash"] != rg["hash"]) | (ro["len"] != rg["len"]))[0]
print("bad", len(bad), "of", n)
for b in bad[:3]:
    ho, hg = ro["head"][b], rg["head"][b]
    k = next((i for i in range(60) if (ho[i] != hg[i]).any()), None)
    print("photon", b, "len", ro["len"][b], rg["len"][b], "first diff at", k)
    if k is not None:
        print(" oracle", ho[max(0, k - 3):k + 3].tolist()); print(" gpu   ", hg[max(0, k - 3):k + 3].tolist())
