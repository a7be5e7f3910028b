"""Top source lines of an .ncu-rep by one stall reason.  usage: ncu_stalls.py rep [reason=long_sb] [n=25]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; reason = sys.argv[2] if len(sys.argv) > 2 else "long_sb"; n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
both = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; h = None; agg = {}; tot = 0.0; allsmp = 0.0
for r in csv.reader(io.StringIO(both)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if "Instructions Executed" in r and "Line No" in r:
        h = r; iL, iA = h.index("Line No"), h.index("Address"); iS = [i for i, x in enumerate(h) if x == "Source"][0]
        iR = h.index("stall_" + reason); iN = h.index("# Samples"); continue
    if h is None or len(r) < len(h) or r[iA] != "-": continue
    try: v = float(r[iR] or 0); ns = float(r[iN] or 0)
    except ValueError: continue
    k = (cur, int(r[iL])); a = agg.setdefault(k, [0.0, r[iS].strip()[:120]]); a[0] += v; tot += v; allsmp += ns
print(f"stall_{reason}: {tot:.0f} samples = {tot / max(allsmp, 1) * 100:.1f}% of all samples")
for (f, l), (v, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{v / max(tot, 1) * 100:5.1f}% | {f}:{l}: {s}")
