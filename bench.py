#!/usr/bin/env python
"""bench.py -- photon packets / s (peel-off on) of the ARTES transport path on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c4] [--photons P]

One "step" = one pass of the hot path (one `artes_gpu_run`, i.e. one `call radiative_transfer`,
src/ARTES.f90:518) over P photon packets per GPU on the synthetic atmosphere named by --workload
(default C4: 3-D r-theta-phi grid, Rayleigh gas + Mie cloud patches, 64x64 Stokes images -- the
configuration the north-star target is quoted on).  Photons shard over ranks by photon id (weak
scaling: P per GPU), the images are summed with NCCL inside the library.

Numbers on the JSON line:
  value      photon packets/s, all ranks, tables resident in HBM, CUDA-event time of kernel+reduce
  e2e        same through the public call with HOST buffers: table upload + launch + image download
  roofline   algorithmic FP64 flop (SURVEY 8d event accounting x exact event counters) / kernel time
             against the FP64 FMA peak measured in this run (this path is FP64-issue bound; the HBM
             side is reported next to it as roofline_hbm)
  cpu_baseline  the CPU oracle (C++ restatement of ARTES.f90 -- the image has no Fortran compiler)
             on the box's host cores, bounded sample of the same workload
`--impl reference` times that CPU implementation alone on the same workload/metric.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "photon packets/sec (peel-off on)"
UNIT = "packets/s"

WORKLOADS = {
    # name: (builder, launch overrides, default photons per GPU per step)
    "c1": ("c1_template_rayleigh", dict(nx=25, ny=25, det_phi=90.0), 4_000_000),
    "c2": ("c2_hg_deck", dict(nx=1, ny=1, det_phi=60.0), 8_000_000),
    "c3": ("c3_molecular", dict(nx=1, ny=1, det_phi=90.0), 8_000_000),
    "c4": ("c4_mie_patches", dict(nx=64, ny=64, det_phi=60.0), 10_000_000),   # BASELINE.json configs[3]: 1e7 packets
    # BASELINE.json configs[4]: multi-wavelength; one step = ONE batched launch of all wavelengths x --photons packets each
    "c5": ("c5_scale", dict(nx=64, ny=64, det_phi=60.0, multi_wl=True), 1_000_000),
}

# algorithmic FP64 flop per event
F_CF = {3: 104.0, 2: 90.0, 1: 50.0}
F_SC = {"faithful": 2700.0, "fast": 2700.0 - 900.0 - 1620.0 + 2 * 64.0}
F_PEEL, F_EMIT = 200.0, 60.0
B_CF, B_PEEL = 8.0, 256.0 + 12 * 8.0
B_SC = {"faithful": 4 * 180 * 8.0 + 2 * 16 * 8.0, "fast": 2 * 8 * 32.0 + 2 * 16 * 8.0}


def grid_dim(atm):
    return 3 if atm.nphi > 1 else (2 if atm.ntheta > 1 else 1)


def algorithmic(stats, atm, mode):
    d = grid_dim(atm)
    fl = stats["n_emit"] * F_EMIT + stats["n_cell_face"] * F_CF[d] + stats["n_scatter"] * F_SC[mode] + stats["n_peel"] * F_PEEL
    by = stats["n_cell_face"] * B_CF + stats["n_scatter"] * B_SC[mode] + stats["n_peel"] * B_PEEL
    return fl, by


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")] + [time.perf_counter()])
        except Exception:
            pass

    def stop(self, t0=None, t1=None):
        """Samples of the timed region [t0, t1] (host clock at arrival); nvidia-smi needs a few hundred ms to deliver its
        first line, so the sampler is started during the warm-up steps (same load) and a timed region shorter than the
        sampling period falls back to the samples under load around it."""
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        window = "timed region"
        if t0 is not None:
            inside = [r for r in self.rows if t0 <= r[-1] <= t1]
            if inside:
                self.rows = inside
            else:
                window = "warm-up + timed region (timed region shorter than the sampling period)"
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_arm(atm, wl_kw, photons, seed, nthreads=0, emulate_stat=False):
    """The reference's CPU implementation of the path (oracle port; gfortran is not in the image)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, RNG_MZ
    from artes_b200.abi import make_launch
    o = Oracle()
    o.set_atmosphere(atm)
    xm = 1.3 * atm.rfront[-1]
    L = make_launch(n_photons=int(photons), x_max=xm, y_max=xm, seed=seed, nx=wl_kw["nx"], ny=wl_kw["ny"],
                    det_phi=math.radians(wl_kw["det_phi"]))
    if nthreads <= 0:     # all host cores of this process, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
        nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    r = o.run(L, rng=RNG_MZ, nthreads=nthreads, emulate_stat=emulate_stat)
    return r["stats"]["kernel_ms"] * 1e-3, int(r["stats"]["reserved"]), r


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries write banners to file descriptor 1 (NCCL prints its version there when NCCL_DEBUG=VERSION and ignores
    NCCL_DEBUG_FILE at that level): point fd 1 at stderr for the run and keep the real stdout for the one JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--photons", type=float, default=0, help="photon packets per GPU per step")
    ap.add_argument("--mode", default="fast", choices=["fast", "faithful"])
    ap.add_argument("--batch", type=int, default=0, help="N > 1: every step is ONE batched launch (artes_gpu_run_batch) of N launches x --photons "
                                                         "packets whose det_phi sweeps 0..180 deg like the phase-curve loop src/ARTES.f90:215-245")
    ap.add_argument("--multi", type=int, default=0, help="N > 1: every step is ONE walk of --photons packets observed by N detectors "
                                                         "(artes_gpu_run_multi; det_phi = 0 .. 167.5 deg like the non-limb angles of a phase curve); "
                                                         "value counts N x photons detector-packets, i.e. what a per-detector loop would have to walk")
    ap.add_argument("--cpu-sample", type=float, default=300000, help="photons of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from tools import atmospheres as A
    builder, wl_kw, default_p = WORKLOADS[args.workload]
    wl_kw = dict(wl_kw)
    multi_wl = bool(wl_kw.pop("multi_wl", False)) and args.batch <= 1 and args.multi <= 1 and args.mode == "fast"
    if args.multi > 1:
        args.batch = args.multi
    atm = getattr(A, builder)()
    P = int(args.photons) if args.photons else (1_000_000 if args.batch > 1 else default_p)
    NB = args.batch if args.batch > 1 else (len(atm.wavelengths) if multi_wl else 1)
    config = {"workload": f"{args.workload}:{builder} nr={atm.nr} ntheta={atm.ntheta} nphi={atm.nphi} "
                          f"image={wl_kw['nx']}x{wl_kw['ny']} det_phi={wl_kw['det_phi']}deg star source, peel-off on",
              "photons_per_gpu_per_step": P * NB, "mode": args.mode, "parallelism": f"photon-id sharding x{world}",
              "l2": "tables (<= few MB) are L2-resident by design; 256 MiB memset flushes L2 between steps"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.multi > 1:
        config["batch"] = (f"ONE walk of {P} packets per step observed by {NB} detectors (artes_gpu_run_multi), det_phi = 0..167.5 deg; "
                           f"value = {NB} x packets / s (detector-packets: the packets a per-detector loop walks for the same curve)")
    elif multi_wl:
        config["batch"] = f"{NB} wavelengths x {P} packets per step as one kernel (artes_gpu_set_wavelengths + artes_gpu_run_batch)"
    elif NB > 1:
        config["batch"] = f"{NB} launches x {P} packets per step as one kernel (artes_gpu_run_batch), det_phi = 0..180 deg"
    if args.impl == "reference":
        if rank != 0:
            return
        sample = int(min(args.cpu_sample, P))
        # the CPU arm times a bounded sample of the workload (a rate: photons are independent), not the GPU arm's step size
        config["photons_per_step_timed_on_cpu"] = sample
        config["cpu_note"] = "photons_per_gpu_per_step is the GPU arm's step; this arm times photons_per_step_timed_on_cpu packets per step of the same workload" + (" (wavelength 0)" if multi_wl else "")
        for _ in range(min(args.warmup, 1)):
            cpu_arm(atm, wl_kw, max(sample // 10, 1000), 1)
        tot = 0.0
        cores = 0
        for s in range(args.steps):
            dt, cores, _ = cpu_arm(atm, wl_kw, sample, 100 + s)
            tot += dt
        v = sample * args.steps / tot
        # the reference as written also pays a stat() system call per packet and per cell_face call (src/ARTES.f90:549, :2820);
        # the headline CPU number above is the FASTER, stat()-free variant -- this is the slower one, on a smaller sample
        st_sample = max(sample // 4, 1000)
        dt_stat, _, _ = cpu_arm(atm, wl_kw, st_sample, 99, emulate_stat=True)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{sample} packets per step of the same workload; C++ restatement of ARTES.f90 "
                                           f"(g++ -O3 -fopenmp, gfortran unavailable), all host threads"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "cpu_with_stat_calls": {"value": st_sample / dt_stat, "unit": UNIT, "sample": f"{st_sample} packets",
                                        "note": "same port with the reference's stat() call per packet and per cell_face evaluation (:549, :2820) kept"}}
        emit(line)
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    from artes_b200 import abi, host, dist as adist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    mode = abi.MODE_FAST if args.mode == "fast" else abi.MODE_FAITHFUL
    params = host.Params(nx=wl_kw["nx"], ny=wl_kw["ny"], det_phi=math.radians(wl_kw["det_phi"]), phase_curve=(NB > 1 and not multi_wl))
    t = host.Transport(atm, params, devices=(local_rank,), mode=mode)
    if world > 1:  # NCCL communicator of the library itself: the id travels through torch.distributed
        adist.init_library_comm(t.gpu, dist, rank, world)

    def load_tables(tr):
        if multi_wl:
            tr.set_all_wavelengths()
        else:
            tr.set_wavelength(0)

    load_tables(t)
    peaks = t.gpu.fma_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches(tr, packets, base):
        if multi_wl:
            return [tr.launch_struct(packets, seed=4, photon_id_base=base, wl_index=k) for k in range(NB)]
        if args.multi > 1:
            return [tr.launch_struct(packets, seed=4, photon_id_base=base, det_phi=max(math.radians(2.5 * k), 1e-3)) for k in range(NB)]
        return [tr.launch_struct(packets, seed=4, photon_id_base=base, det_phi=math.pi * k / (NB - 1)) for k in range(NB)]

    def run_many(tr, ls):
        return tr.gpu.run_multi(ls) if args.multi > 1 else tr.gpu.run_batch(ls)

    def step(i):
        base = adist.step_base(i, world, rank, P * NB)   # disjoint photon ids for every step and rank
        if NB > 1:
            return run_many(t, launches(t, P, base))
        L = t.launch_struct(P, seed=4, photon_id_base=base)
        return t.gpu.run(L)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
        flush.zero_()
        torch.cuda.synchronize()
    barrier()
    t_timed0 = time.perf_counter()
    dev_ms = 0.0
    kern_ms = 0.0
    red_ms = 0.0
    agg = dict(n_emit=0, n_cell_face=0, n_scatter=0, n_peel=0)
    last = None
    for i in range(args.steps):
        res = step(args.warmup + i)
        dev_ms += res["stats"]["kernel_ms"] + res["stats"]["reduce_ms"]
        kern_ms += res["stats"]["kernel_ms"]
        red_ms += res["stats"]["reduce_ms"]
        for k in agg:
            agg[k] += res["stats"][k]          # already summed over ranks by the NCCL reduce
        last = res
        flush.zero_()                          # L2 flush, outside the event-timed region
        torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop(t_timed0, time.perf_counter()) if rank == 0 else None

    # max over ranks of the device times; min / max of the kernel time and of the reduce wait per rank say where a scaling loss sits:
    # reduce_ms is the time from this rank's kernel end to the end of the all-reduce, i.e. mostly the wait for the slowest rank
    tm = torch.tensor([dev_ms, kern_ms, red_ms], dtype=torch.float64, device="cuda")
    tmin = tm.clone()
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    dev_ms, kern_ms = float(tm[0]), float(tm[1])
    rank_times = {"kernel_ms_per_step_min": float(tmin[1]) / args.steps, "kernel_ms_per_step_max": float(tm[1]) / args.steps,
                  "reduce_ms_per_step_min": float(tmin[2]) / args.steps, "reduce_ms_per_step_max": float(tm[2]) / args.steps,
                  "note": "per rank: kernel = transport kernel; reduce = from the rank's kernel end to the end of the NCCL all-reduce (includes the wait for the slowest rank)"}
    value = world * args.steps * P * NB / (dev_ms * 1e-3)

    # ---- end to end through the public call with host buffers: upload tables, launch, download image
    a = atm
    nwl_up = len(a.wavelengths) if multi_wl else 1
    h2d = sum(a.k_sca[l].nbytes + a.k_abs[l].nbytes + a.uniq[l].nbytes + a.cell_to_uniq[l].nbytes for l in range(nwl_up))
    d2h = ((10 * wl_kw["nx"] * wl_kw["ny"] + 2) * NB + (7 * a.cells if NB == 1 else 0)) * 8 + (64 + 8) * 8
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        load_tables(t)
        step(args.warmup + args.steps + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = world * args.steps * P * NB / float(te[0])

    # ---- shard check: the N-rank result of one small step against the same photon ids walked by rank 0 alone.
    # Philox is keyed by the global photon id, so the two differ only in the order of the floating-point sums:
    # the count planes must be identical, the Stokes sums equal to rounding.
    pc = int(min(P, 200_000))
    shards = world if world > 1 else 2          # one GPU: two half-launches on the same device added on the host
    base0 = 1 << 40                             # ids no timed step used
    if world > 1:
        rs = run_many(t, launches(t, pc, base0 + rank * pc * NB)) if NB > 1 else t.gpu.run(t.launch_struct(pc, seed=4, photon_id_base=base0 + rank * pc))
        det_n = rs["det"]
    else:
        det_n = None
        for r_ in range(shards):
            rs = run_many(t, launches(t, pc, base0 + r_ * pc * NB)) if NB > 1 else t.gpu.run(t.launch_struct(pc, seed=4, photon_id_base=base0 + r_ * pc))
            det_n = rs["det"].copy() if det_n is None else det_n + rs["det"]
    shard_check = None
    if rank == 0:
        t1 = host.Transport(atm, params, devices=(local_rank,), mode=mode)      # a second context WITHOUT the communicator
        load_tables(t1)
        if NB > 1:       # launch k of shard r walks ids base0 + r*pc*NB + k*pc + [0, pc): the same ids, shard by shard
            det_1 = None
            for r_ in range(shards):
                rr = run_many(t1, launches(t1, pc, base0 + r_ * pc * NB))
                det_1 = rr["det"].copy() if det_1 is None else det_1 + rr["det"]
        else:
            det_1 = t1.gpu.run(t1.launch_struct(pc * shards, seed=4, photon_id_base=base0))["det"]
        t1.close()
        cn, c1 = det_n[..., 2, :, :, :], det_1[..., 2, :, :, :]
        sn, s1 = det_n[..., 0, :, :, :], det_1[..., 0, :, :, :]
        scale = float(np.abs(s1).max()) or 1.0
        shard_check = {"photons": pc * shards * NB, "shards": shards,
                       "counts_sum_sharded": float(cn.sum()), "counts_sum_single": float(c1.sum()),
                       "count_planes_identical": bool(np.array_equal(cn, c1)),
                       "sum_I_sharded": float(sn[..., 0, :, :].sum()), "sum_I_single": float(s1[..., 0, :, :].sum()),
                       "sum_I_rel_diff": float(abs(sn[..., 0, :, :].sum() - s1[..., 0, :, :].sum()) / abs(s1[..., 0, :, :].sum())),
                       "max_pixel_diff_over_max": float(np.abs(sn - s1).max() / scale),
                       "ok": bool(np.array_equal(cn, c1) and np.abs(sn - s1).max() <= 1e-9 * scale),
                       "what": "one small step: photon ids sharded over the ranks and NCCL-summed (one GPU: two half-launches) vs the "
                               "same ids in one launch on rank 0, a context without communicator"}

    if rank == 0:
        fl, by = algorithmic(agg, atm, args.mode)          # all ranks, all timed steps
        per_launch_fl = fl / (world * args.steps)
        per_launch_by = by / (world * args.steps)
        k_s = kern_ms * 1e-3 / args.steps                  # average kernel duration (max over ranks)
        achieved = per_launch_fl / k_s / 1e12
        peaks_file = {}
        try:
            peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
        traffic = traffic_at = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
            # as captured (the DRAM traffic of this kernel is mostly launch-constant: tables once, the L2-resident photon records
            # as they are evicted, the image; 32 MB at 4e6 packets on C4 with the final kernel of round 2), not scaled to this launch size
            traffic = tr["dram_bytes_per_launch"] if tr else None
            traffic_at = tr["photons_per_launch"] if tr else None
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": args.steps * world,
                "clocks": clocks,
                "rank_times": rank_times,
                "shard_check": shard_check,
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s",
                             "frac": achieved / peaks["fp64_tflops"], "traffic": traffic, "traffic_captured_at_packets_per_launch": traffic_at,
                             "note": "dominant kernel transport3_kernel (ray/event engine, asynchronous scheduling); this path is FP64-issue / latency bound, "
                                     "not HBM or tensor bound: algorithmic FP64 flop = exact event counters x SURVEY 8d "
                                     "per-event figures; peak = FP64 FMA rate measured in this run by artes_gpu_fma_peak "
                                     "(MEASURED_PEAKS.json holds no FP64 figure); traffic = ncu dram bytes per launch of "
                                     "the same workload (profiles/traffic.json)",
                             "flop_per_packet": fl / (world * args.steps * P * NB), "fp32_peak_tflops": peaks["fp32_tflops"]},
                "roofline_hbm": {"bound": "hbm", "achieved": per_launch_by / k_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": per_launch_by / k_s / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks_file else "fallback",
                                 "note": "algorithmic bytes are table reads served by L1/L2; HBM is not the bound of this path"},
                "events_per_packet": {k: v / (world * args.steps * P * NB) for k, v in agg.items()}}
        if world == 1 and not args.no_cpu_baseline:
            sample = int(min(args.cpu_sample, P))
            dt, cores, _ = cpu_arm(atm, wl_kw, sample, 7)
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} packets of the same workload; C++ restatement of ARTES.f90 "
                                              f"(g++ -O3 -fopenmp; gfortran unavailable), all host threads"}
        emit(line)
    t.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
