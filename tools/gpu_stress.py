"""Developer tool: many launches of random sizes, seeds and options through the ray/event engine (plain, general, batched and
multi-detector instantiations) and the event-list engine of the faithful mode -- a scheduling dead-lock would show as a
watchdog error or a time-out.  usage: python tools/gpu_stress.py [n_launches]"""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import atmospheres as A
from artes_b200 import abi, host
n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(12345)
ts = {}
for name, builder, px in (("c1", "c1_template_rayleigh", 25), ("c2", "c2_hg_deck", 1), ("c4", "c4_mie_patches", 32)):
    t = host.Transport(getattr(A, builder)(), host.Params(nx=px, ny=px, det_phi=math.radians(60.0)), mode=abi.MODE_FAST)
    t.set_wavelength(0)
    ts[name] = t
t0 = time.perf_counter(); tot = 0
for i in range(n_launch):
    name = ("c1", "c2", "c4")[int(rng.integers(3))]
    t = ts[name]
    n = int(10 ** rng.uniform(0.0, 5.7))
    L = t.launch_struct(n, seed=int(rng.integers(1 << 30)), photon_id_base=int(rng.integers(1 << 40)), det_phi=float(rng.uniform(0, 2 * math.pi)))
    kind = int(rng.integers(6))
    if kind == 1: L.surface_albedo = float(rng.uniform(0.05, 1.0))
    if kind == 2: L.flow_theta = 1
    if kind == 3:
        nb = int(rng.integers(2, 40))
        Ls = []
        for k in range(nb):
            Lk = t.launch_struct(max(n // nb, 1), seed=L.seed, photon_id_base=L.photon_id_base, det_phi=float(rng.uniform(0, 2 * math.pi)))
            Lk.surface_albedo = 0.3 if (i & 1) else 0.0
            Ls.append(Lk)
        r = t.gpu.run_batch(Ls); tot += nb * max(n // nb, 1)
        assert r["stats"]["n_emit"] == nb * max(n // nb, 1)
    elif kind == 4:      # one walk observed by nb detectors
        nb = int(rng.integers(1, 100))
        Ls = [t.launch_struct(n, seed=L.seed, photon_id_base=L.photon_id_base, det_phi=float(rng.uniform(1e-3, math.pi - 1e-3))) for k in range(nb)]
        r = t.gpu.run_multi(Ls); tot += n
        assert r["stats"]["n_emit"] == n, (r["stats"]["n_emit"], n, nb)
    elif kind == 5:      # faithful mode: the event-list engine (engine3.cuh), small launches
        n = min(n, 20000)
        L.n_photons = n
        L.mode = abi.MODE_FAITHFUL
        if i & 1: L.surface_albedo = 0.5
        if i & 2: L.flow_global = 1; L.flow_theta = 1
        r = t.gpu.run(L, flows=bool(L.flow_theta)); tot += n
        assert r["stats"]["n_emit"] == n and t.gpu.last_engine() == 3
    else:
        r = t.gpu.run(L, flows=bool(L.flow_theta)); tot += n
        assert r["stats"]["n_emit"] == n
    assert t.gpu.last_engine() == (3 if kind == 5 else 2)
print(f"stress: {n_launch} launches, {tot} packets, {time.perf_counter() - t0:.1f} s, no dead-lock, all emitted", flush=True)
