"""Scratch GPU check: oracle vs libartes_gpu on a few configurations (developer tool)."""
import math, sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from tools import atmospheres as A
from oracle_lib import Oracle, RNG_PHILOX
from artes_b200 import abi
from artes_b200.lib import GpuTransport
from artes_b200.abi import make_launch


def setup(atm, photon_source=1):
    o = Oracle(); depth = o.set_atmosphere(atm, photon_source=photon_source)
    g = GpuTransport((0,))
    g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
    g.set_wavelength(atm.k_sca[0], atm.k_abs[0], atm.uniq[0], atm.cell_to_uniq[0], depth)
    return o, g, depth


def trace_cmp(name, atm, n=20000, max_draws=256, **kw):
    o, g, depth = setup(atm)
    xm = 1.3 * atm.rfront[-1]
    rs = np.random.RandomState(7)
    xi = rs.random_sample((n, max_draws))
    for mode in (0, 1):
        L = make_launch(mode=mode, n_photons=n, x_max=xm, y_max=xm, fstop=0.02, **kw)
        t = time.time(); ro = o.trace(L, xi, max_rec=8); to = time.time() - t
        t = time.time(); rg = g.trace(L, xi, max_rec=8); tg = time.time() - t
        same = (ro["hash"] == rg["hash"]) & (ro["len"] == rg["len"])
        dpos = np.abs(ro["fstate"] - rg["fstate"])
        print(f"[trace {name} mode={mode}] n={n} identical={same.sum()}/{n} mean_len={ro['len'].mean():.1f} "
              f"max|dfstate|={np.nanmax(dpos[same]) if same.any() else float('nan'):.3e} oracle {to:.2f}s gpu {tg:.2f}s")
        bad = np.where(~same)[0][:3]
        for b in bad:
            print("  mismatch photon", b, "len", ro["len"][b], rg["len"][b]); print(ro["head"][b].tolist()); print(rg["head"][b].tolist())


def run_cmp(name, atm, n=200000, **kw):
    o, g, depth = setup(atm)
    xm = 1.3 * atm.rfront[-1]
    for mode in (0, 1):
        L = make_launch(mode=mode, n_photons=n, x_max=xm, y_max=xm, seed=11, **kw)
        ro = o.run(L, rng=RNG_PHILOX)
        rg = g.run(L)
        do, dg = ro["det"], rg["det"]
        tot = np.abs(do[0, 0]).sum()
        print(f"[run {name} mode={mode}] I oracle {do[0,0].sum():.9e} gpu {dg[0,0].sum():.9e} rel {abs(do[0,0].sum()-dg[0,0].sum())/tot:.2e} "
              f"Q {do[0,1].sum():.6e}/{dg[0,1].sum():.6e} U {do[0,2].sum():.4e}/{dg[0,2].sum():.4e} counts {do[2,0].sum():.0f}/{dg[2,0].sum():.0f} "
              f"max pix rel {np.abs(do[0,0]-dg[0,0]).max()/do[0,0].max():.2e}")
        so, sg = ro["stats"], rg["stats"]
        print("   stats oracle", {k: so[k] for k in ("n_emit","n_cell_face","n_scatter","n_peel","n_surface","n_draws","n_error")})
        print("   stats gpu   ", {k: sg[k] for k in ("n_emit","n_cell_face","n_scatter","n_peel","n_surface","n_draws","n_error")},
              f"kernel {sg['kernel_ms']:.2f} ms -> {n/sg['kernel_ms']*1e3:.3e} pkt/s ; oracle {so['kernel_ms']:.0f} ms -> {n/so['kernel_ms']*1e3:.3e} pkt/s")
        eo = {i: int(v) for i, v in enumerate(ro["err"]) if v}; eg = {i: int(v) for i, v in enumerate(rg["err"]) if v}
        if eo or eg: print("   err oracle", eo, "gpu", eg)


if __name__ == "__main__":
    g = GpuTransport((0,)); print(g.device_info(), g.fma_peak()); g.close()
    c1 = A.c1_template_rayleigh(); c2 = A.c2_hg_deck(); c4 = A.c4_mie_patches()
    # cell_face parity on random interior points of c4
    o, g, depth = setup(c4)
    rs = np.random.RandomState(3); n = 200000
    ir = rs.randint(0, c4.nr, n); it = rs.randint(0, c4.ntheta, n); ip = rs.randint(0, c4.nphi, n)
    r = c4.rfront[ir] + rs.random_sample(n) * (c4.rfront[ir + 1] - c4.rfront[ir])
    th = np.radians(c4.theta_deg[it] + rs.random_sample(n) * (c4.theta_deg[it + 1] - c4.theta_deg[it]))
    phf = np.append(c4.phi_deg, 360.0); ph = np.radians(phf[ip] + rs.random_sample(n) * (phf[ip + 1] - phf[ip]))
    pos = np.stack([r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), r * np.cos(th)], 1)
    d = rs.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    face = np.zeros((n, 2), np.int32); cell = np.stack([ir, it, ip], 1).astype(np.int32)
    oi, od = o.cell_face(pos, d, face, cell)
    for mode in (0, 1):
        gi, gd = g.cell_face(pos, d, face, cell, mode=mode)
        print(f"[cell_face mode={mode}] int equal {np.all(oi == gi, axis=1).sum()}/{n} dist bit-equal {(od == gd).sum()}/{n} max rel {np.max(np.abs(od-gd)/od):.2e}")
    trace_cmp("c1", c1); trace_cmp("c2", c2); trace_cmp("c4", c4, n=10000)
    trace_cmp("c4 albedo surface", c4, n=10000, surface_albedo=0.5)
    run_cmp("c1", c1, n=100000); run_cmp("c2", c2, n=100000, nx=1, ny=1, det_phi=math.radians(60)); run_cmp("c4", c4, n=100000, nx=64, ny=64, det_phi=math.radians(60))
