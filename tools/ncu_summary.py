"""Summarise an .ncu-rep: key metrics + hottest source lines (SASS metrics folded onto CUDA lines).

usage: python tools/ncu_summary.py report.ncu-rep [n_lines]
"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("=" * 100)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:75s} {r[i]} {units[i]}")
    st = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"   stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.3f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# line info: ncu's sass page has no line column in csv; use the "cuda,sass" correlated view if available
both = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rdr = list(csv.reader(io.StringIO(both)))
# find header row containing 'Source'
h = None
agg = {}
tot_inst = tot_thr = tot_samp = 0.0
cur_line = None
for r in rdr:
    if "Instructions Executed" in r:
        h = r
        i_src = h.index("Source"); i_inst = h.index("Instructions Executed"); i_thr = h.index("Thread Instructions Executed"); i_samp = h.index("# Samples")
        i_ln = h.index("Line No") if "Line No" in h else None
        continue
    if h is None or len(r) < len(h):
        continue
    try:
        ni, nt, ns = float(r[i_inst] or 0), float(r[i_thr] or 0), float(r[i_samp] or 0)
    except ValueError:
        continue
    key = r[i_ln] if i_ln is not None else r[i_src][:60]
    a = agg.setdefault(key, [0.0, 0.0, 0.0, r[i_src][:120]])
    a[0] += ns; a[1] += ni; a[2] += nt
    tot_inst += ni; tot_thr += nt; tot_samp += ns
print("=" * 100)
print(f"total warp-inst {tot_inst:.3e} thread-inst {tot_thr:.3e} avg lanes {tot_thr / max(tot_inst, 1):.2f} samples {tot_samp:.0f}")
for key, (ns, ni, nt, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:nlines]:
    print(f"{ns / max(tot_samp, 1) * 100:5.1f}% smp {ni / max(tot_inst, 1) * 100:5.1f}% inst lanes {nt / max(ni, 1):5.1f} | {key}: {s}")
