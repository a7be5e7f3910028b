"""Developer tool: the eight-element matrix table (DevTables::Mc) against the 16-element path on the same launch."""
import math, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from tools import atmospheres as A
    from artes_b200 import abi, host
    from artes_b200 import lib
    lib.load().artes_gpu_test_full_matrix(1 if sys.argv[2] == "full" else 0)
    for name, builder, px in (("c1", "c1_template_rayleigh", 25), ("c2", "c2_hg_deck", 1), ("c4", "c4_mie_patches", 64)):
        t = host.Transport(getattr(A, builder)(), host.Params(nx=px, ny=px, det_phi=math.radians(60.0)), mode=abi.MODE_FAST)
        t.set_wavelength(0)
        r = t.gpu.run(t.launch_struct(400000, seed=3))
        np.save(f"/tmp/compact_{name}_{sys.argv[2]}.npy", r["det"])
        print(name, sys.argv[2], "kernel_ms %.2f" % r["stats"]["kernel_ms"], "I %.12e" % r["det"][0].sum(), flush=True)
        t.close()
else:
    import numpy as np
    for tag, env in (("compact", {}), ("full", {})):
        subprocess.check_call([sys.executable, __file__, "child", tag], env=dict(os.environ, **env))
    for name in ("c1", "c2", "c4"):
        a, b = np.load(f"/tmp/compact_{name}_compact.npy"), np.load(f"/tmp/compact_{name}_full.npy")
        print(name, "identical" if np.array_equal(a, b) else "max rel diff %.3e, count planes equal %s" % (np.abs(a - b).max() / np.abs(b).max(), np.array_equal(a[8:], b[8:])))
