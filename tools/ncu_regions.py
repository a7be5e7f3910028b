"""Instruction / stall-sample share per function of engine2.cuh (+ helper files) from an .ncu-rep taken with the CURRENT sources.
usage: python tools/ncu_regions.py report.ncu-rep"""
import csv, io, os, re, subprocess, sys
rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "artes_b200", "csrc", "engine2.cuh")).read().split("\n")
# function start lines
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:    )?(?:__device__ __forceinline__|__global__)[^(]*?\b([A-Za-z_0-9]+)\s*\(", l)
    if m:
        name = m.group(1)
        if l.startswith("    ") and name in ("init", "load", "step", "finish", "trip"): name = "Marcher::" + name
        elif l.startswith("    "): continue
        if name == "__launch_bounds__":
            name = re.search(r"(transport[234]_kernel)", l).group(1)
        starts.append((i, name))
def func_of(line):
    name = "engine2:other"
    for s, n in starts:
        if s <= line: name = n
        else: break
    return name
both = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; h = None; agg = {}; tot = [0.0, 0.0, 0.0]
for r in csv.reader(io.StringIO(both)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if "Instructions Executed" in r and "Line No" in r:
        h = r; iL, iA = h.index("Line No"), h.index("Address"); iI, iT, iN = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples"); continue
    if h is None or len(r) < len(h) or r[iA] != "-": continue
    try: ni, nt, ns, ln = float(r[iI] or 0), float(r[iT] or 0), float(r[iN] or 0), int(r[iL])
    except ValueError: continue
    key = func_of(ln) if cur == "engine2.cuh" else cur
    a = agg.setdefault(key, [0.0, 0.0, 0.0]); a[0] += ni; a[1] += nt; a[2] += ns
    tot[0] += ni; tot[1] += nt; tot[2] += ns
print(f"total warp-inst {tot[0]:.3e} lanes {tot[1] / tot[0]:.1f}")
for k, (ni, nt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if ni / tot[0] > 0.002: print(f"{ni / tot[0] * 100:5.1f}% inst {ns / tot[2] * 100:5.1f}% smp lanes {nt / max(ni, 1):5.1f} | {k}")
