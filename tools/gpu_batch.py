"""Developer tool: a phase curve (n launches of P packets that differ in det_phi) as single launches and as one
batched launch.  usage: python tools/gpu_batch.py c2 1000000 73 [pixels]"""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import atmospheres as A
from artes_b200 import abi, host
name = sys.argv[1]; P = int(float(sys.argv[2])); n = int(sys.argv[3])
atm = getattr(A, {"c1": "c1_template_rayleigh", "c2": "c2_hg_deck", "c3": "c3_molecular", "c4": "c4_mie_patches", "c5": "c5_scale"}[name])()
px = int(sys.argv[4]) if len(sys.argv) > 4 else {"c1": 25, "c2": 1, "c3": 1, "c4": 64, "c5": 64}[name]
t = host.Transport(atm, host.Params(nx=px, ny=px, det_phi=math.radians(60.0), phase_curve=True), mode=abi.MODE_FAST)
t.set_wavelength(0)
phis = [math.radians(180.0 * k / max(n - 1, 1)) for k in range(n)]      # the reference's sweep: 0 .. 180 deg
t.gpu.run(t.launch_struct(P, seed=1))
t0 = time.perf_counter(); ms = 0.0; cf = 0
for k, phi in enumerate(phis):
    r = t.gpu.run(t.launch_struct(P, seed=2, photon_id_base=k * P, det_phi=phi)); ms += r["stats"]["kernel_ms"]; cf += r["stats"]["n_cell_face"]
seq = time.perf_counter() - t0
Ls = [t.launch_struct(P, seed=2, det_phi=phi) for phi in phis]
t.gpu.run_batch(Ls)
t0 = time.perf_counter()
b = t.gpu.run_batch(Ls)
bat = time.perf_counter() - t0
print(f"{name} {n} launches x {P} px {px}: single launches {seq*1e3:.1f} ms wall / {ms:.1f} ms kernels ({n*P/seq:.3e} pkt/s)   "
      f"batched {bat*1e3:.1f} ms wall / {b['stats']['kernel_ms']:.1f} ms kernel ({n*P/bat:.3e} pkt/s)  cf {cf} vs {b['stats']['n_cell_face']}  "
      f"last image I {r['det'][0,0].sum():.6e} vs {b['det'][n-1][0,0].sum():.6e}", flush=True)

if len(atm.wavelengths) > 1:      # the spectrum loop: one launch per wavelength vs one batched launch over wl_index
    import numpy as np
    nl = len(atm.wavelengths)
    t0 = time.perf_counter(); ms = 0.0
    for l in range(nl):
        t.set_wavelength(l)
        r = t.gpu.run(t.launch_struct(P, seed=2, photon_id_base=l * P)); ms += r["stats"]["kernel_ms"]
    seq = time.perf_counter() - t0
    uq, c2u, off, depths = [], [], 0, []
    for l in range(nl):
        uq.append(atm.uniq[l]); c2u.append(np.asarray(atm.cell_to_uniq[l]) + off); off += atm.uniq[l].shape[0]
        depths.append(host.cell_depth(atm.rfront, atm.k_sca[l], atm.k_abs[l], atm.nr, atm.ntheta, atm.nphi, 1))
    t0 = time.perf_counter()
    t.gpu.set_wavelengths(np.stack(atm.k_sca), np.stack(atm.k_abs), np.concatenate(uq), np.stack(c2u), depths)
    Ls = []
    for l in range(nl):
        L = t.launch_struct(P, seed=2); L.wl_index = l; Ls.append(L)
    b = t.gpu.run_batch(Ls)
    bat = time.perf_counter() - t0
    print(f"{name} spectrum {nl} wavelengths x {P}: single launches {seq*1e3:.1f} ms wall / {ms:.1f} ms kernels ({nl*P/seq:.3e} pkt/s)   "
          f"batched {bat*1e3:.1f} ms wall incl. table upload / {b['stats']['kernel_ms']:.1f} ms kernel ({nl*P/bat:.3e} pkt/s)  "
          f"last I {r['det'][0,0].sum():.6e} vs {b['det'][nl-1][0,0].sum():.6e}", flush=True)
