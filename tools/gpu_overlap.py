"""Developer experiment: do the drain of one launch and the ramp-up of the next overlap when they are issued on
two streams (two contexts)?  usage: python tools/gpu_overlap.py c4 1000000 8"""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import atmospheres as A
from artes_b200 import abi, host
name = sys.argv[1]; n = int(float(sys.argv[2])); reps = int(sys.argv[3])
atm = getattr(A, {"c1": "c1_template_rayleigh", "c2": "c2_hg_deck", "c3": "c3_molecular", "c4": "c4_mie_patches", "c5": "c5_scale"}[name])()
px = {"c1": 25, "c2": 1, "c3": 1, "c4": 64, "c5": 64}[name]
ts = [host.Transport(atm, host.Params(nx=px, ny=px, det_phi=math.radians(60.0)), mode=abi.MODE_FAST) for _ in range(2)]
for t in ts:
    t.set_wavelength(0); t.gpu.run(t.launch_struct(n, seed=1))
t0 = time.perf_counter()
for i in range(reps):
    ts[0].gpu.run(ts[0].launch_struct(n, seed=2, photon_id_base=i * n))
seq = time.perf_counter() - t0
t0 = time.perf_counter()
pend = []
for i in range(reps):
    t = ts[i & 1]
    if len(pend) == 2:
        pend.pop(0).gpu.wait()
    t.gpu.run_async(t.launch_struct(n, seed=2, photon_id_base=i * n)); pend.append(t)
for t in pend:
    t.gpu.wait()
ovl = time.perf_counter() - t0
print(f"{name} n={n} reps={reps}: sequential {seq*1e3:.1f} ms ({reps*n/seq:.3e} pkt/s)  two streams {ovl*1e3:.1f} ms ({reps*n/ovl:.3e} pkt/s)")
