"""CPU prototype of the engine2 ray geometry (per-axis incremental crossings) checked against the oracle's
cell_face iterated along the same ray.  Developer tool: validates the ALGORITHM of engine2.cuh."""
import math, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from tools import atmospheres as A
from oracle_lib import Oracle
from test_oracle import random_interior_points

NONE = 1e300
PI = math.pi

def quadric_next(qa, hb, qc, hs, lim, z0, n2):
    disc = hb * hb - qa * qc
    if not (disc >= 0.0): return NONE
    q = -(hb + math.copysign(math.sqrt(disc), hb))
    t1 = q / qa if abs(qa) > 1e-100 else NONE
    t2 = qc / q if abs(q) > 1e-100 else NONE
    if not (t1 > lim) or (z0 + t1 * n2) * hs < 0.0: t1 = NONE
    if not (t2 > lim) or (z0 + t2 * n2) * hs < 0.0: t2 = NONE
    return min(t1, t2)

class Ray:
    def __init__(s, atm, x, n, cell, sface=-1, obl=(1, 1, 1)):
        s.atm = atm; s.r = atm.rfront; s.tf = atm.thetafront(); s.tpl = atm.thetaplane(); s.pf = atm.phifront()
        s.nr, s.nt, s.np = atm.nr, atm.ntheta, atm.nphi
        a, b, c = 1 / obl[0], 1 / obl[1], 1 / obl[2]
        s.A1 = a*a*n[0]*n[0] + b*b*n[1]*n[1]; s.A2 = c*c*n[2]*n[2]
        s.B1 = a*a*x[0]*n[0] + b*b*x[1]*n[1]; s.B2 = c*c*x[2]*n[2]
        s.C1 = a*a*x[0]*x[0] + b*b*x[1]*x[1]; s.C2 = c*c*x[2]*x[2]
        qa = s.A1 + s.A2; hb = s.B1 + s.B2; Cs = s.C1 + s.C2
        s.iq = 1 / qa; s.hbn = -hb * s.iq; s.D0 = s.hbn * s.hbn - Cs * s.iq
        s.g0 = n[2] * Cs - x[2] * hb; s.g1 = n[2] * hb - x[2] * qa
        s.Xx, s.Xy, s.Nx, s.Ny, s.z0, s.n2 = a*x[0], b*x[1], a*n[0], b*n[1], x[2], n[2]
        s.c = list(cell); s.t = 0.0
        # radial first
        s.inward = s.hbn > 0
        s.tr = NONE
        done = False
        if s.inward:
            disc = s.r[s.c[0]]**2 * s.iq + s.D0
            if disc >= 0:
                t = s.hbn - math.sqrt(disc)
                if t > 1e-15: s.tr = t; done = True
            if not done: s.inward = False
        if not done:
            disc = s.r[s.c[0] + 1]**2 * s.iq + s.D0
            if disc >= 0:
                t = s.hbn + math.sqrt(disc)
                if t > (1e-3 if sface == s.c[0] + 1 else 1e-15): s.tr = t
        s.tt, s.upper = (s.theta_next(0.0) if s.nt > 1 else (NONE, 0))
        s.tp, s.up = s.phi_next(0.0)
    def cone_root(s, k, t):
        if s.tpl[k] != 1:
            if s.tpl[k] != 2 or s.n2 == 0: return NONE
            r = -s.z0 / s.n2
            return r if r > t else NONE
        T2 = math.tan(s.tf[k])**2
        hs = 1 if s.tf[k] < PI/2 else (-1 if s.tf[k] > PI/2 else 0)
        return quadric_next(s.A1 - s.A2*T2, s.B1 - s.B2*T2, s.C1 - s.C2*T2, hs, t, s.z0, s.n2)
    def theta_next(s, t):
        down = (s.g0 + s.g1 * t) > 0
        for _ in range(2):
            k = s.c[1] if down else s.c[1] + 1
            if k != 0 and k != s.nt:
                tk = s.cone_root(k, t)
                if tk < NONE: return tk, (0 if down else 1)
            down = not down
        return NONE, 0
    def phi_next(s, t):
        up = (s.Xx * s.Ny - s.Xy * s.Nx) > 0
        if s.np <= 1: return NONE, up
        k = ((s.c[2] + 1) % s.np) if up else s.c[2]
        ps, pc = math.sin(s.pf[k]), math.cos(s.pf[k])
        den = s.Ny * pc - s.Nx * ps
        if den == 0: return NONE, up
        r = (s.Xx * ps - s.Xy * pc) / den
        return (r if r > t else NONE), up
    def step(s):
        """returns (nf0, nf1, c0, c1, c2, dist) like cell_face"""
        tn, ax = s.tr, 0
        if s.tt < tn: tn, ax = s.tt, 1
        if s.tp < tn: tn, ax = s.tp, 2
        if not tn < NONE: return None
        d = tn - s.t; s.t = tn
        if ax == 0:
            f = s.c[0] if s.inward else s.c[0] + 1
            s.c[0] += -1 if s.inward else 1
            if f != s.nr and s.c[0] >= 0:
                ok = False
                if s.inward:
                    disc = s.r[s.c[0]]**2 * s.iq + s.D0
                    if disc >= 0: s.tr = s.hbn - math.sqrt(disc); ok = True
                if not ok:
                    s.inward = False
                    disc = s.r[s.c[0] + 1]**2 * s.iq + s.D0
                    s.tr = s.hbn + math.sqrt(disc) if disc >= 0 else NONE
        elif ax == 1:
            f = s.c[1] + 1 if s.upper else s.c[1]
            s.c[1] += 1 if s.upper else -1
            s.tt, s.upper = s.theta_next(s.t)
        else:
            f = ((s.c[2] + 1) % s.np) if s.up else s.c[2]
            s.c[2] = (s.c[2] + 1) % s.np if s.up else (s.c[2] - 1) % s.np
            s.tp, _ = s.phi_next(s.t)
        return (ax + 1, f, s.c[0], s.c[1], s.c[2], d)

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c4_mie_patches"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    atm = getattr(A, name)()
    o = Oracle(); o.set_atmosphere(atm)
    pos, d, face, cell = random_interior_points(atm, n, 5)
    bad = 0; steps = 0
    for i in range(n):
        ray = Ray(atm, pos[i], d[i], cell[i])
        p = pos[i:i+1].copy(); f = face[i:i+1].copy(); c = cell[i:i+1].copy()
        for k in range(400):
            oi, od = o.cell_face(p, d[i:i+1], f, c)
            mine = ray.step()
            steps += 1
            ref = tuple(int(v) for v in oi[0, :5])
            if oi[0, 6] != 0:
                break
            if mine is None or mine[:5] != ref or abs(mine[5] - od[0]) > 1e-6 * max(1.0, od[0]):
                bad += 1
                if bad <= 8:
                    print("ray", i, "step", k, "oracle", ref, od[0], "mine", mine, "pos", pos[i], "dir", d[i], "cell", cell[i])
                break
            if oi[0, 5]: break   # grid exit
            if ref[0] == 1 and ref[1] == 0: break
            p = p + od[0] * d[i:i+1]; f = oi[:, 0:2].copy(); c = oi[:, 2:5].copy()
    print(name, "rays", n, "steps", steps, "bad rays", bad)

if __name__ == "__main__":
    main()
