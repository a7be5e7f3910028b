"""Small launches of every kernel family for compute-sanitizer (memcheck): plain / general / trace, both engines."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tools import atmospheres as A
from artes_b200 import abi, host
from artes_b200.abi import make_launch
for name, kw in (("c4_mie_patches", dict()), ("c4_mie_patches", dict(surface_albedo=0.5, flow_theta=1)), ("c1_template_rayleigh", dict())):
    atm = getattr(A, name)()
    t = host.Transport(atm, host.Params(nx=16, ny=16, det_phi=math.radians(60.0)), mode=abi.MODE_FAST)
    t.set_wavelength(0)
    xm = t.x_max
    for mode in (abi.MODE_FAST, abi.MODE_FAITHFUL):
        L = make_launch(mode=mode, n_photons=3000, x_max=xm, y_max=xm, seed=3, nx=16, ny=16, det_phi=math.radians(60.0), **kw)
        r = t.gpu.run(L, flows=bool(kw.get("flow_theta")))
        print(name, kw, "mode", mode, "engine", t.gpu.last_engine(), "I", float(r["det"][0, 0].sum()), "err", int(r["stats"]["n_error"]))
        xi = np.random.RandomState(1).random_sample((500, 120))
        Lt = make_launch(mode=mode, n_photons=500, x_max=xm, y_max=xm, fstop=0.05, nx=16, ny=16, **{k: v for k, v in kw.items() if k != "flow_theta"})
        tr = t.gpu.trace(Lt, xi, max_rec=8)
        print("  trace engine", t.gpu.last_engine(), "mean len", float(tr["len"].mean()))
    t.close()
print("SANITIZE_RUN_DONE")
