"""Host check of the fast-mode elementary functions (artes_b200/csrc/fastmath.cuh: sin/cos on [0, pi], acos on (-1, 1),
log on (0, 1], exp(-x) on [0, 700]) against glibc: at most 1 ulp over 4e6 random arguments each.  The header compiles for
the host with the same arithmetic (no FMA contraction here; on the device the fused forms only lower the error)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fastmath_within_one_ulp_of_glibc(tmp_path):
    exe = str(tmp_path / "fastmath_check")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", os.path.join(ROOT, "tools", "fastmath_check.cc"), "-o", exe], check=True, cwd=ROOT)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    m = re.search(r"max ulp: sin ([\d.]+) cos ([\d.]+).*acos ([\d.]+) log ([\d.]+) exp ([\d.]+)", out)
    assert m, out
    assert all(float(v) <= 1.0 + 1e-9 for v in m.groups()), out
    assert "log(1)=0" in out and "exp(0)=1" in out
