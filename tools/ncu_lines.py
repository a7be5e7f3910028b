"""Per-source-line view of an .ncu-rep (needs -lineinfo + --import-source on): instructions, lanes, stall samples.

usage: python tools/ncu_lines.py report.ncu-rep [n_lines] [file-substring]
Rows of the ncu "cuda,sass" source page are grouped per file; only the per-line aggregate rows are used."""
import csv, io, subprocess, sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 50
both = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = "?"
h = None
agg = {}
tot = [0.0, 0.0, 0.0]
per_file = {}
for r in csv.reader(io.StringIO(both)):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if "Instructions Executed" in r and "Line No" in r:
        h = r
        iL, iA = h.index("Line No"), h.index("Address")
        iS = [i for i, x in enumerate(h) if x == "Source"][0]
        iI, iT, iN = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
        continue
    if h is None or len(r) < len(h) or r[iA] != "-":
        continue
    try:
        ni, nt, ns = float(r[iI] or 0), float(r[iT] or 0), float(r[iN] or 0)
    except ValueError:
        continue
    key = (cur_file, int(r[iL]))
    a = agg.setdefault(key, [0.0, 0.0, 0.0, r[iS].strip()[:110]])
    a[0] += ns; a[1] += ni; a[2] += nt
    tot[0] += ns; tot[1] += ni; tot[2] += nt
    f = per_file.setdefault(cur_file, [0.0, 0.0, 0.0]); f[0] += ns; f[1] += ni; f[2] += nt
print(f"total warp-inst {tot[1]:.3e} thread-inst {tot[2]:.3e} avg lanes {tot[2] / max(tot[1], 1):.2f} samples {tot[0]:.0f}")
for f, (ns, ni, nt) in sorted(per_file.items(), key=lambda kv: -kv[1][1]):
    print(f"  file {f:28s} {ni / tot[1] * 100:5.1f}% inst {ns / tot[0] * 100:5.1f}% smp lanes {nt / max(ni, 1):5.1f}")
flt = sys.argv[3] if len(sys.argv) > 3 else ""
for (f, ln), (ns, ni, nt, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:nlines]:
    if flt and flt not in f: continue
    print(f"{ni / tot[1] * 100:5.1f}% inst {ns / tot[0] * 100:5.1f}% smp lanes {nt / max(ni, 1):5.1f} | {f}:{ln}: {s}")

# region split for engine2.cuh: marcher loop vs event phase (line ranges read from the source file itself)
import os
src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "artes_b200", "csrc", "engine2.cuh")
if os.path.exists(src):
    L = open(src).read().split("\n")
    def find(s):
        return next((i + 1 for i, x in enumerate(L) if s in x), None)
    m0, m1, e1 = find("// ================= marcher phase"), find("// ================= event phase"), find("// counters")
    if m0 and m1 and e1:
        reg = {"marcher": [0, 0, 0], "event-dispatch": [0, 0, 0], "events": [0, 0, 0], "other files": [0, 0, 0]}
        for (f, ln), (ns, ni, nt, s) in agg.items():
            if f != "engine2.cuh":
                k = "other files"
            elif m0 <= ln < m1: k = "marcher"
            elif m1 <= ln < e1: k = "event-dispatch"
            elif ln < m0 - 40: k = "events"
            else: k = "marcher"
            reg[k][0] += ns; reg[k][1] += ni; reg[k][2] += nt
        for k, (ns, ni, nt) in reg.items():
            print(f"region {k:16s} {ni / tot[1] * 100:5.1f}% inst {ns / tot[0] * 100:5.1f}% smp lanes {nt / max(ni, 1):5.1f}")
