"""Per-function shares of an .ncu-rep with source correlation: folds the SASS rows of `--page source` onto the device
functions of a source file (by line ranges taken from the file as it was when the profiled library was built).

usage: python tools/ncu_functions.py report.ncu-rep path/to/engine2.cuh [git-rev]
"""
import csv, io, re, subprocess, sys

rep, src = sys.argv[1], sys.argv[2]
rev = sys.argv[3] if len(sys.argv) > 3 else None
import os
SRC_DIR = os.path.dirname(src)
_cache = {}
def starts_of(fname):
    """function start lines of a source file of the profiled build (git revision `rev` if given)"""
    if fname in _cache:
        return _cache[fname]
    path = os.path.join(SRC_DIR, os.path.basename(fname))
    text = subprocess.run(["git", "show", f"{rev}:{path}"], capture_output=True, text=True).stdout if rev else (open(path).read() if os.path.exists(path) else "")
    st = []
    for i, l in enumerate(text.split("\n"), 1):
        m = re.search(r"(?:__device__|__global__)[^;(]*?\b(\w+)\s*\(", l)
        if m and not l.strip().startswith("//"):
            st.append((i, m.group(1) if m.group(1) != "__launch_bounds__" else "kernel main loop"))
    _cache[fname] = st
    return st
def func_of(fname, n):
    name = "?"
    for i, nm in starts_of(fname):
        if i <= n:
            name = nm
        else:
            break
    return name

both = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rdr = list(csv.reader(io.StringIO(both)))
h = None
agg = {}
tot = [0.0, 0.0, 0.0]
cur_file = ""
for r in rdr:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1]
        continue
    if "Instructions Executed" in r:
        h = r
        ix = {k: h.index(k) for k in ("Instructions Executed", "Thread Instructions Executed", "# Samples")}
        continue
    if h is None or len(r) < len(h) or not r[0].isdigit():
        continue          # SASS rows repeat what their source-line row already sums
    try:
        ni, nt, ns = float(r[ix["Instructions Executed"]] or 0), float(r[ix["Thread Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    base = os.path.basename(cur_file)
    key = f"{base}:{func_of(cur_file, int(r[0]))}" if cur_file.startswith("/root/repo") else f"[{base}]"
    a = agg.setdefault(key, [0.0, 0.0, 0.0])
    a[0] += ns; a[1] += ni; a[2] += nt
    tot[0] += ns; tot[1] += ni; tot[2] += nt
print(f"total warp-inst {tot[1]:.3e}  thread-inst {tot[2]:.3e}  avg lanes {tot[2] / max(tot[1], 1):.2f}  samples {tot[0]:.0f}")
print(f"{'function':40s} {'samples%':>9s} {'inst%':>7s} {'lanes':>6s}")
for k, (ns, ni, nt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if ns / max(tot[0], 1) < 0.002:
        continue
    print(f"{k:40s} {ns / tot[0] * 100:9.1f} {ni / tot[1] * 100:7.1f} {nt / max(ni, 1):6.1f}")
