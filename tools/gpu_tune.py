"""Developer tool: throughput of one workload for the regrouping thresholds given in the environment."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import atmospheres as A
from artes_b200 import abi, host
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fast"]
builder = {"c1": "c1_template_rayleigh", "c2": "c2_hg_deck", "c3": "c3_molecular", "c4": "c4_mie_patches", "c5": "c5_scale"}[name]
atm = getattr(A, builder)()
px = {"c1": 25, "c2": 1, "c3": 1, "c4": 64, "c5": 64}[name]
for m in modes:
    t = host.Transport(atm, host.Params(nx=px, ny=px, det_phi=math.radians(60.0)), mode=abi.MODE_FAST if m == "fast" else abi.MODE_FAITHFUL)
    t.set_wavelength(0)
    t.gpu.run(t.launch_struct(n // 4, seed=1))
    r = t.gpu.run(t.launch_struct(n, seed=2))
    st = r["stats"]
    print(f"{name} {m} ev={os.environ.get('ARTES_DEFER_EVENTS','-')} rf={os.environ.get('ARTES_DEFER_REFILL','-')} "
          f"n={n} kernel {st['kernel_ms']:.1f} ms -> {n/st['kernel_ms']*1e3:.4e} pkt/s  launches {st['reserved']} cf/pkt {st['n_cell_face']/n:.1f} sc/pkt {st['n_scatter']/n:.2f} I={r['det'][0,0].sum():.6e}", flush=True)
    if os.environ.get("E2_STATS"):
        e = [int(x) for x in r["err"][50:59]]
        print(f"  passes {e[0]:.3e} lanes@start {e[1]/max(e[0],1):.1f} rdy_empty {e[2]/max(e[0],1):.2f} step-iters/pass {e[3]/max(e[0],1):.1f} lanes/step {e[4]/max(e[3],1):.1f} "
              f"evlist(H+DEP+RES)/block {e[5]/max(e[0],1):.0f} rdy/block {e[6]/max(e[0],1):.0f} event batches {e[7]:.3e} lanes/batch {e[8]/max(e[7],1):.1f}")
        names = ["EMIT", "PRE", "H", "DEP", "RES", "SURF", "FAN", "SC", "RDY"]
        e = [int(x) for x in r["err"]]
        print("  " + "  ".join(f"{nm}: backlog {e[4+i]/max(e[50],1):.1f} batches {e[13+i]:.2e} lanes {e[22+i]/max(e[13+i],1):.1f}" for i, nm in enumerate(names) if e[4+i] or e[13+i]))
    t.close()
