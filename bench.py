#!/usr/bin/env python
"""bench.py -- the hot path's headline numbers on B200, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload direct|tree|direct4m|galaxy]
                    [--impl reference] [--no-extras] [--no-cpu-baseline]

BASELINE.json's metric has two halves: pairwise interactions/s (direct summation) and
particle-steps/s (Barnes-Hut tree), both at 1/2/4/8 B200.  The default run measures BOTH at every
--gpus N:

  headline   BASELINE.json configs[2]: Plummer sphere N = 1,048,576, direct summation, fp32 pair
             arithmetic, one full DKD leapfrog step per "step" (force on all N particles from all N
             + kick + drift, state resident in HBM); metric = pairwise interactions/s = N^2 / step.
  "tree"     BASELINE.json configs[3]: Hernquist sphere N = 4,194,304, Barnes-Hut theta = 0.7, fp32
             walk, one DKD step per "step"; particle-steps/s, with its own e2e (pinned and
             pageable host buffers), roofline (walk against the FP32 FMA peak, build against the
             measured HBM peak), accuracy, parity_check and cpu_baseline.
  "fp64"     configs[2]'s fp64 arm: the same Plummer N = 2^20 through the fp64 direct kernel.
  "galaxy"   (--gpus 8 only) BASELINE.json configs[4]: exponential disk + Hernquist halo, N = 10M, tree.

With --gpus N > 1 (launched by torchrun, one rank per GPU) the targets are sharded N/P per rank and
each step all-gathers the half-drifted positions over NCCL (strong scaling: total work fixed).
Before anything is timed every rank checks 256 of its own targets against the oracle
("parity_check": the CPU restatement of the reference, used here only as the checker).

--workload tree|galaxy|direct4m makes that workload the headline instead (no extra blocks).
--impl reference: times the reference's own CPU implementation of the same path
(oracle/_ref = the unmodified /root/reference/gravhopper/_jbgrav.c compiled by oracle/Makefile;
falls back to the oracle port) on the host cores, on a bounded sample of the workload.

Keys follow the driver's contract: value = whole-job throughput with inputs resident in HBM;
e2e = the same metric through the public call with HOST buffers (pinned), H2D and D2H inside the
timed region; roofline = the dominant kernel against the FP32 FMA peak (this path is FMA-pipe
bound, not HBM or tensor bound: SURVEY 8d); cpu_baseline = reference C timed beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_DIRECT = int(os.environ.get("GH_BENCH_DIRECT_N", 1 << 20))  # override only for tests/experiments
N_TREE = int(os.environ.get("GH_BENCH_TREE_N", 1 << 22))
FLOP_PER_INTERACTION = 20  # north_star / GPU-Gems-3 convention
SM_COUNT = 148
FP32_LANES_PER_SM = 128
FP64_LANES_PER_SM = 64
PARITY_TARGETS = 256
C_ACC = 4.398600412921223e-09  # jbgrav.py:48


def workload(kind):
    from gravhopper_b200 import ic_raw
    if kind == "direct4m":
        n = 1 << 22
        x, v, m = ic_raw.Plummer(n, 1e-3, 1e6, seed=42)
        return dict(x=np.ascontiguousarray(x), v=np.ascontiguousarray(v), m=m, eps=5e-5, dt=0.005,
                    theta=0.7, alg="direct", prec="fp32",
                    name="Plummer N=4194304, direct summation fp32, 1 DKD leapfrog step (north_star: 8-GPU scaling at N=4M)")
    if kind == "galaxy":
        n = 10_000_000
        x, v, m = ic_raw.galaxy_model(n)
        return dict(x=np.ascontiguousarray(x), v=np.ascontiguousarray(v), m=m, eps=0.05, dt=1.0, theta=0.7,
                    alg="tree", prec="fp32",
                    name="Exponential disk (2M) + Hernquist halo (8M) N=10000000, Barnes-Hut theta=0.7 fp32 walk, "
                         "1 DKD leapfrog step (BASELINE.json configs[4])")
    if kind in ("direct", "direct64"):
        x, v, m = ic_raw.Plummer(N_DIRECT, 1e-3, 1e6, seed=42)
        prec = "fp32" if kind == "direct" else "fp64"
        return dict(x=np.ascontiguousarray(x), v=np.ascontiguousarray(v), m=m, eps=5e-5, dt=0.005,
                    theta=0.7, alg="direct", prec=prec,
                    name="Plummer N=%d b=1pc M=1e6Msun eps=0.05pc dt=0.005Myr, direct summation %s, "
                         "1 DKD leapfrog step (BASELINE.json configs[2])" % (N_DIRECT, prec))
    x, v, m = ic_raw.Hernquist(N_TREE, 1.0, 1e10, seed=42)
    return dict(x=np.ascontiguousarray(x), v=np.ascontiguousarray(v), m=m, eps=0.05, dt=1.0, theta=0.7,
                alg="tree", prec="fp32",
                name="Hernquist N=%d a=1kpc M=1e10Msun eps=0.05kpc dt=1Myr, Barnes-Hut theta=0.7 fp32 walk, "
                     "1 DKD leapfrog step (BASELINE.json configs[3])" % N_TREE)


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def merge_clocks(a, b):
    """Clock records of two timed regions -> one (the line's `clocks` covers everything timed)."""
    if not a or not b:
        return a or b
    na, nb = a["samples"], b["samples"]
    return {"sm_mhz": (a["sm_mhz"] * na + b["sm_mhz"] * nb) / (na + nb),
            "sm_max_mhz": max(a["sm_max_mhz"], b["sm_max_mhz"]),
            "power_w_max": max(a["power_w_max"], b["power_w_max"]), "samples": na + nb,
            "reasons": sorted(set(a["reasons"]) | set(b["reasons"]))}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except ValueError:
            pass
    return {}


def cpu_baseline_direct(w, seconds_target=12.0):
    """Reference C (oracle/_ref) on ONE core (the reference is single threaded), bounded sample:
    nf targets x all N sources through direct_summation_position (the only reference
    entry point that can run at N = 2^20)."""
    from oracle import oracle as O
    ref = O.ref()
    n = len(w["m"])
    nf = 192  # 32*nf*np bytes of scratch inside the reference = 6.4 GB at N = 2^20
    if ref is not None:
        fn, kind = (lambda t: ref.direct_summation_position(w["x"], w["m"], t, w["eps"])), "reference"
    else:
        fn, kind = (lambda t: O.direct_summation_position(w["x"], w["m"], t, w["eps"], nthreads=1)), "port"
    t0 = time.perf_counter()
    fn(w["x"][:8])
    est = (time.perf_counter() - t0) / 8
    nf = int(max(16, min(nf, seconds_target / max(est, 1e-9))))
    t0 = time.perf_counter()
    fn(w["x"][:nf])
    dt = time.perf_counter() - t0
    return {"value": nf * n / dt, "unit": "interactions/s", "cores": 1, "kind": kind,
            "sample": "%d targets x %d sources, direct_summation_position, %.1f s" % (nf, n, dt)}


def cpu_baseline_tree(w, ntargets=16384):
    from oracle import oracle as O
    ref = O.ref()
    sel = np.random.default_rng(0).choice(len(w["m"]), min(ntargets, len(w["m"])), replace=False)
    t0 = time.perf_counter()
    if ref is not None:
        ref.tree_force_position(w["x"], w["m"], w["x"][sel], w["eps"], w["theta"])
        kind = "reference"
    else:
        O.tree_force_position(w["x"], w["m"], w["x"][sel], w["eps"], w["theta"], nthreads=1)
        kind = "port"
    dt = time.perf_counter() - t0
    # one evaluation = build (all N) + walk (sampled targets); extrapolate the walk to all N targets
    return {"value": len(sel) / dt, "unit": "particle-steps/s", "cores": 1, "kind": kind,
            "sample": "tree build over %d sources + walk of %d targets, %.1f s (build included once)"
                      % (len(w["m"]), len(sel), dt)}


def tree_roofline(n, world, ms_per_step, kernel_ms, st, mode, accuracy, peaks, fp32_peak_tflops, phases=None,
                  peak_how=None):
    """roofline object of the tree workloads: the walk kernel against the FP32 FMA peak (20 flop x
    list entries), the build against the HBM roofline (SURVEY 8d bytes per particle-step)."""
    acc_per = st["accepted"] / float(n)
    vis_per = st["visited"] / float(n)
    achieved = (n / world) * acc_per * FLOP_PER_INTERACTION / (kernel_ms * 1e-3) / 1e12
    if mode == "group":
        kernel = "walk_group_kernel"
        how = ("20 flop x interaction-list entries (%.0f per target: one warp-cooperative traversal per 32 "
               "Morton-consecutive targets with the bounding-box form of the reference's opening test, which "
               "opens every cell the reference opens and some more; the reference's own per-target set is "
               "~1.9x shorter) / mean CUDA-event time of the walk kernel over the timed steps, against the FP32 "
               "FMA peak; visited_per_target = entries tested by the target's group (shared by its 32 targets); "
               "the build (ms_per_step - kernel_ms) is HBM-streaming bound" % acc_per)
        bound = "fp32_fma (list evaluation) + issue (traversal)"
    else:
        kernel = "walk_kernel"
        how = ("20 flop x accepted nodes (the reference's own accepted set: %.0f per target) / mean CUDA-event "
               "time of the walk kernel, against the FP32 FMA peak; ncu (profiles/) shows this walk is entry-load "
               "latency / instruction-issue bound; the build (ms_per_step - kernel_ms) is HBM-streaming bound"
               % acc_per)
        bound = "issue (walk); fp32_fma peak quoted"
    if peak_how:
        how += "; peak: " + peak_how
    # dram__bytes_read.sum + dram__bytes_write.sum of walk_group_kernel at N = 2^22 on one GPU in the
    # engine path (this bench's), from `ncu --set full` (profiles/r02_walk_group_kernel_4194304.txt):
    # 1010.1 MB + 478.5 MB.  Algorithmic: 32 B x 6.2M entries read once (0.20 GB) + 160 B x N of
    # targets and fused kick/drift state (x_half, v, m read; x, v, x_half, float4 source written:
    # 0.67 GB) = 0.87 GB; the excess is entry re-reads that miss L2 (hit rate 70 %)
    traffic = 1488.6e6 if (mode == "group" and n == (1 << 22) and world == 1) else None
    # the build (everything of the step that is not the walk kernel) against the HBM roofline:
    # SURVEY 8d's algorithmic bytes per particle-step, 190 + 24 x radix passes (8) = 382 B
    build_ms = ms_per_step - kernel_ms
    hbm_peak = float(peaks.get("hbm_gbs") or 0.0) or 6650.0
    build_gbs = 382.0 * n / (build_ms * 1e-3) / 1e9
    build_roofline = {"bound": "hbm", "achieved": build_gbs, "peak": hbm_peak, "unit": "GB/s",
                      "frac": build_gbs / hbm_peak, "ms": build_ms,
                      "how": "382 B per particle (SURVEY 8d's accounting: 190 + 24 x 8 radix passes; the running "
                             "simulation's splitter sort moves the pairs through memory 3 times, not 8, so this is "
                             "the algorithmic figure, not the build's traffic) x N / (ms_per_step - walk "
                             "kernel ms; with --gpus N > 1 this includes the NCCL exchanges); peak = %s"
                             % ("MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks.get("hbm_gbs")
                                else "6.65 TB/s (of fallback, B200_PROFILING.md)")}
    r = {"bound": bound, "achieved": achieved, "peak": fp32_peak_tflops,
         "unit": "TFLOP/s", "frac": achieved / fp32_peak_tflops, "traffic": traffic,
         "kernel": kernel, "kernel_ms": kernel_ms, "walk": mode,
         "accepted_per_target": acc_per, "visited_per_target": vis_per,
         "tree_entries": st["entries"], "tree_cells": st["cells"], "deepest_level": st["maxlevel"],
         "build_ms": build_ms, "build_roofline": build_roofline, "accuracy": accuracy, "how": how}
    if phases:
        r["phases_ms"] = phases
    return r


_W = None  # workload shared with forked reference workers (no per-step pickling of the sources)


def _ref_worker(args):
    kind, targets = args
    x, m, eps, theta = _W["x"], _W["m"], _W["eps"], _W["theta"]
    from oracle import oracle as O
    ref = O.ref()
    if kind == "direct":
        if ref is not None:
            ref.direct_summation_position(x, m, targets, eps)
        else:
            O.direct_summation_position(x, m, targets, eps, nthreads=1)
    else:
        if ref is not None:
            ref.tree_force_position(x, m, targets, eps, theta)
        else:
            O.tree_force_position(x, m, targets, eps, theta, nthreads=1)
    return len(targets)


def run_reference(args):
    """--impl reference: the reference C backend on all host cores (P processes, each evaluating a
    slice of targets against all sources -- the only way the single-threaded reference can use
    more than one core)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import oracle as O
    global _W
    w = workload(args.workload)
    _W = w
    n = len(w["m"])
    cores = os.cpu_count() or 1
    P = max(1, min(cores, 32))
    kind = "reference" if O.ref() is not None else "port"
    wl_direct = w["alg"] == "direct"
    if wl_direct:
        per = 24  # targets per process per step: ~1.3 s of reference C, 0.8 GB scratch each
    else:
        per = 2048
    rng = np.random.default_rng(1)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(P) as pool:
        for s in range(args.warmup + args.steps):
            jobs = []
            for p in range(P):
                sel = rng.choice(n, min(per, n), replace=False)
                jobs.append(("direct" if wl_direct else "tree", w["x"][sel]))
            t0 = time.perf_counter()
            pool.map(_ref_worker, jobs)
            if s >= args.warmup:
                times.append(time.perf_counter() - t0)
    tot = sum(times)
    units = (min(per, n) * P * n) if wl_direct else (min(per, n) * P)
    value = units * args.steps / tot
    unit = "interactions/s" if wl_direct else "particle-steps/s"
    sample = ("%d processes x %d targets x %d sources per step (direct_summation_position)" % (P, per, n)
              if wl_direct else
              "%d processes, each: tree build over %d sources + walk of %d targets per step" % (P, n, per))
    # the reference computes in IEEE fp64 whatever the GPU arm's pair arithmetic is: same particles,
    # same N, same step; `precision` names the arithmetic of THIS arm
    name = w["name"].replace(" fp32", "").replace(" fp64", "")
    line = {"impl": "reference", "metric": "pairwise interactions/s" if wl_direct else "particle-steps/s",
            "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "n_particles": n, "precision": "fp64",
                       "parallelism": "%d host processes, targets sharded (the reference is single threaded)" % P},
            "cpu_baseline": {"value": value, "unit": unit, "cores": P, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


class Ctx(object):
    """Process-wide state of one bench run (rank, world, torch handles, peaks)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from gravhopper_b200 import _jbgrav as J, _lib, _pinned
        self.torch, self.dist, self.J, self._lib, self._pinned = torch, dist, J, _lib, _pinned
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        _lib.require_gpu()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        self.peaks = measured_peaks()
        self.fp32_peak = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def allsum(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def fp32_peak(ctx, sm_max):
    """FP32 FMA peak: measured by the library's FFMA-only probe kernel (gh_fp32_fma_probe: 8
    independent FFMA chains per thread, every SM full, no memory traffic) when the library has
    it, else the nominal lane count x clock."""
    nominal = SM_COUNT * FP32_LANES_PER_SM * 2 * sm_max * 1e6 / 1e12
    if ctx.fp32_peak is None:
        ctx.fp32_peak = (nominal, "nominal: 148 SMs x 128 FP32 lanes x 2 flop x clocks.max.sm (%.0f MHz); "
                                  "MEASURED_PEAKS.json holds HBM GB/s and bf16 TF/s only" % sm_max)
        fn = getattr(ctx._lib.lib(), "gh_fp32_fma_probe", None)
        if fn is not None:
            import ctypes as C
            tf = C.c_double()
            fn.restype, fn.argtypes = C.c_int, [C.c_int, C.POINTER(C.c_double)]
            if fn(5, C.byref(tf)) == 0 and tf.value > 0:
                ctx.fp32_peak = (tf.value, "measured in this process: gh_fp32_fma_probe (FFMA-only kernel, 8 "
                                           "independent chains per thread, best of 5) = %.2f TFLOP/s; nominal 148 x "
                                           "128 x 2 x %.0f MHz = %.2f" % (tf.value, sm_max, nominal))
    return ctx.fp32_peak


def parity_check(ctx, w, sim, pos_l, vel_l, mass_l):
    """Before timing: ONE step of the sharded engine from the initial state, then every rank
    recovers the accelerations of 256 of its own targets from the kick (a = (v1 - v0)/dt/C_ACC,
    exact to ~1e-14 here) and compares them with the oracle evaluated at the same half-drifted
    positions (direct_summation_position; for the tree also the oracle's reference tree,
    _jbgrav.c:487-541, on rank 0).  This exercises the path that is timed: all-gather, force
    kernel(s), fused kick/drift epilogue."""
    from oracle import oracle as O
    torch = ctx.torch
    b, c = sim.begin, sim.count
    nthreads = max(1, (os.cpu_count() or 1) // max(1, ctx.world))
    xh_all = O.half_drift(pos_l, vel_l, w["dt"])
    sel = np.sort(np.random.default_rng(100 + ctx.rank).choice(c, min(PARITY_TARGETS, c), replace=False))
    sim.step()
    _, v1 = sim.local_state()
    a_gpu = (v1[sel] - vel_l[b:b + c][sel]) / w["dt"] / C_ACC
    tgt = np.ascontiguousarray(xh_all[b:b + c][sel])
    truth = O.direct_summation_position(xh_all, mass_l, tgt, w["eps"], nthreads=nthreads)
    e = np.linalg.norm(a_gpu - truth, axis=1) / np.linalg.norm(truth, axis=1)
    if w["alg"] == "direct":
        tol = 1e-5 if w["prec"] == "fp32" else 1e-12
        mx, = ctx.allmax([float(e.max())])
        return {"max_rel_err": mx, "tol": tol, "ok": bool(mx <= tol), "targets_per_rank": len(sel),
                "vs": "oracle direct_summation_position (reference _jbgrav.c:299-353 restated) at the same x_half, "
                      "through one step of the sharded engine"}
    # tree: the contract is statistical (north_star: error against direct summation no worse than
    # the reference tree's); every rank contributes its targets' errors, rank 0 evaluates the
    # reference tree on its own targets for the bar
    s_mean, cnt = ctx.allsum([float(e.sum()), float(len(e))])
    mx, = ctx.allmax([float(e.max())])
    ref_mean = ref_max = own_max = 0.0
    if ctx.rank == 0:
        rt = O.tree_force_position(xh_all, mass_l, tgt, w["eps"], w["theta"], nthreads=nthreads)
        er = np.linalg.norm(rt - truth, axis=1) / np.linalg.norm(truth, axis=1)
        ref_mean, ref_max, own_max = float(er.mean()), float(er.max()), float(e.max())
    ref_mean, ref_max, own_max = ctx.allmax([ref_mean, ref_max, own_max])
    mean = s_mean / cnt
    # like for like: the mean over all ranks' targets against the reference tree's mean on rank 0's
    # 256 (the mean is stable across samples); the MAX on the same 256 targets for both
    tol_mean, tol_max = 1.05 * ref_mean, 1.5 * ref_max
    return {"mean_rel_err": mean, "max_rel_err_all_ranks": mx, "max_rel_err": own_max,
            "reference_tree_mean_rel_err": ref_mean, "reference_tree_max_rel_err": ref_max,
            "tol": {"mean": tol_mean, "max": tol_max},
            "ok": bool(mean <= tol_mean and own_max <= tol_max), "targets_per_rank": len(sel),
            "vs": "error against the oracle's direct summation at the same x_half; bar = the oracle's reference tree "
                  "(theta = %.2f): mean over all ranks' targets <= 1.05 x its mean on rank 0's targets, max on rank 0's "
                  "targets <= 1.5 x its max on the same targets (samples of 256: the full-distribution tail checks are "
                  "in tests/test_gpu_scale.py)" % w["theta"]}


def run_block(ctx, kind, steps, warmup, with_cpu_baseline, with_e2e=True, with_parity=True):
    """One workload through the sharded engine: parity check, timed steps, e2e, roofline."""
    torch, dist, J = ctx.torch, ctx.dist, ctx.J
    from gravhopper_b200.sharded import ShardedSimulation
    rank, world, local = ctx.rank, ctx.world, ctx.local
    w = workload(kind)
    n = len(w["m"])
    sim = ShardedSimulation(w["x"], w["v"], w["m"], w["dt"], w["eps"], algorithm=w["alg"], theta=w["theta"],
                            precision=w["prec"], rank=rank, world=world, device=local)
    shard = sim.shard
    if sim.perm is not None:
        pos_l, vel_l, mass_l = w["x"][sim.perm], w["v"][sim.perm], w["m"][sim.perm]
    else:
        pos_l, vel_l, mass_l = w["x"], w["v"], w["m"]
    parity = parity_check(ctx, w, sim, pos_l, vel_l, mass_l) if with_parity else None
    del pos_l, vel_l, mass_l

    def one_step():
        with shard.stream_context():
            ctx.flush.zero_()  # evict L2 between steps (the 16-24 MB source array would otherwise stay hot)
        sim.step()

    for _ in range(warmup):
        one_step()
    ctx.barrier()
    launches0 = shard.launches()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with shard.stream_context():
        ev0.record()
    for _ in range(steps):
        one_step()
    with shard.stream_context():
        ev1.record()
    ctx.barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = shard.launches() - launches0
    clocks = sampler.stop() if sampler else None
    # per-kernel time of the dominant (force) kernel: events the engine records around it, MEAN over
    # the timed steps (the last 64 at most)
    kernel_ms = shard.force_ms_mean(steps)
    phases = sim.phase_ms() if hasattr(sim, "phase_ms") else None
    ms_total, kernel_ms = ctx.allmax([ms_total, kernel_ms])
    ms_per_step = ms_total / steps

    is_direct = w["alg"] == "direct"
    if is_direct:
        units_per_step = float(n) * float(n)
        metric, unit = "pairwise interactions/s", "interactions/s"
    else:
        units_per_step = float(n)
        metric, unit = "particle-steps/s", "particle-steps/s"
    value = units_per_step / (ms_per_step * 1e-3)

    # ---- end to end through the reference-facing call with HOST buffers ----
    # Every rank evaluates its share of the targets against all sources through the public
    # _jbgrav call with host arrays (H2D of the sources + its targets, D2H of its accelerations
    # inside the timed region); the job's rate is all units / the slowest rank's time.
    e2e = None
    if with_e2e:
        b0, cnt = sim.begin, sim.count

        def make_call(hx, hm, ht):
            if world == 1:
                if is_direct:
                    return lambda: J.direct_summation(hx, hm, w["eps"], precision=w["prec"])
                return lambda: J.tree_force(hx, hm, w["eps"], w["theta"], precision=w["prec"])
            if is_direct:
                return lambda: J.direct_summation_position(hx, hm, ht, w["eps"], precision=w["prec"])
            return lambda: J.tree_force_position(hx, hm, ht, w["eps"], w["theta"], precision=w["prec"])

        def time_calls(call, nwarm, reps):
            # untimed calls first: the stateless path's device scratch and the binding's pool of page-locked
            # result blocks (two alternate while `out` is rebound) reach steady state before the clock starts
            out = None
            for _ in range(nwarm):
                out = call()
            ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                out = call()
            te = (time.perf_counter() - t0) / reps
            te, = ctx.allmax([te])
            return te, out

        big = is_direct and n > (1 << 21)
        slow = is_direct and w["prec"] == "fp64"
        nwarm = 1 if (big or slow) else 3
        reps = (1 if (big or slow) else 3) if is_direct else 5
        hx = torch.from_numpy(w["x"]).pin_memory().numpy()
        hm = torch.from_numpy(w["m"]).pin_memory().numpy()
        tgt = np.ascontiguousarray(w["x"][b0:b0 + cnt])
        ht = torch.from_numpy(tgt).pin_memory().numpy()
        te, out = time_calls(make_call(hx, hm, ht), nwarm, reps)
        name = ("direct_summation" if is_direct else "tree_force") + ("" if world == 1 else "_position")
        h2d = int(hx.nbytes + hm.nbytes + (ht.nbytes if world > 1 else 0))
        e2e = {"value": units_per_step / te, "unit": unit, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(out.nbytes), "ms_per_call": te * 1e3, "n_gpus_used": world,
               "host_buffers": "page-locked inputs (torch pin_memory) and page-locked results (the binding's pool)",
               "call": "_jbgrav.%s(host ndarrays) -> host ndarray, one call per rank on its share of the targets"
                       % name}
        if not is_direct:
            # what a user's plain ndarrays pay: pageable inputs, a fresh pageable result per call
            del out
            ctx._pinned.ENABLED = False
            try:
                tp, out = time_calls(make_call(w["x"], w["m"], tgt), 2, 3)
            finally:
                ctx._pinned.ENABLED = True
            e2e["pageable"] = {"value": units_per_step / tp, "unit": unit, "ms_per_call": tp * 1e3,
                               "host_buffers": "plain (pageable) ndarrays in, a fresh pageable ndarray out"}
        del hx, hm, ht, out

    block = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup,
             "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
             "dtype": "f32" if w["prec"] == "fp32" else "f64", "data": "synthetic",
             "config": {"workload": w["name"], "n_particles": n, "precision": w["prec"],
                        "parallelism": sim.describe() if hasattr(sim, "describe") else
                        "targets sharded over %d rank(s), NCCL all-gather of x_half per step" % world,
                        "l2": "256 MB flush write between steps"},
             "gpu_launches": int(launches), "clocks": clocks}
    if e2e:
        block["e2e"] = e2e
    if parity:
        block["parity_check"] = parity

    if rank == 0:
        peaks = ctx.peaks
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        if is_direct and w["prec"] == "fp32":
            peak, peak_how = fp32_peak(ctx, sm_max)
            per_rank_units = units_per_step / world
            achieved = per_rank_units * FLOP_PER_INTERACTION / (kernel_ms * 1e-3) / 1e12
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at N = 2^20 on one GPU from
            # `ncu --set full` (profiles/r01_direct_f32_N1M.txt): 24.4 MB + 76.5 MB
            traffic = 101.0e6 if (n == (1 << 20) and world == 1) else None
            nominal = SM_COUNT * FP32_LANES_PER_SM * 2 * sm_max * 1e6 / 1e12
            roofline = {"bound": "fp32_fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak, "traffic": traffic,
                        "traffic_unit": "bytes per launch (ncu); algorithmic: 16.8 MB sources read + 24 MB x S partial sums",
                        "kernel": "direct_f32_kernel", "kernel_ms": kernel_ms,
                        "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
                        "how": "20 flop/interaction x N_i x N_j per launch / mean CUDA-event time of the force kernel "
                               "over the timed steps; peak: " + peak_how}
            if clocks:
                roofline["frac_at_observed_clock"] = achieved / (peak * clocks["sm_mhz"] / sm_max)
            block["roofline"] = roofline
        elif is_direct:
            dfma_peak = SM_COUNT * FP64_LANES_PER_SM * 2 * sm_max * 1e6 / 1e12
            achieved = (units_per_step / world) * FLOP_PER_INTERACTION / (kernel_ms * 1e-3) / 1e12
            block["roofline"] = {"bound": "fp64_fma", "achieved": achieved, "peak": dfma_peak, "unit": "TFLOP/s",
                                 "frac": achieved / dfma_peak, "traffic": None, "kernel": "direct_f64_kernel",
                                 "kernel_ms": kernel_ms,
                                 "how": "20 flop/interaction convention against 148 SMs x 64 DFMA lanes x 2 x clocks.max.sm "
                                        "(nominal); the kernel executes 16-17 DP operations + 1 MUFU per interaction "
                                        "(DESIGN 4.2), i.e. 32-34 flop-equivalents"}
        else:
            peak, peak_how = fp32_peak(ctx, sm_max)
            # one extra, untimed evaluation with counters on: list entries / tested entries per target
            J.tree_stats(True)
            tx = torch.from_numpy(w["x"]).cuda()
            tm = torch.from_numpy(w["m"]).cuda()
            a32 = J.tree_force(tx, tm, w["eps"], w["theta"], precision=w["prec"])
            torch.cuda.synchronize()
            st = J.tree_stats()
            J.tree_stats(False)
            mode = J.tree_walk()
            # accuracy beside the rate (north_star: "reported with accuracy matching the reference"):
            # per-particle relative acceleration error against fp64 direct summation on sampled
            # targets, for the timed fp32 walk and for the reference's criterion (the fp64 per-target
            # walk accepts exactly the reference's node set: tests/test_gpu_parity.py)
            sel = torch.from_numpy(np.random.default_rng(0).choice(n, min(4096, n), replace=False)).cuda()
            tsel = tx[sel].contiguous()
            d = J.direct_summation_position(tx, tm, tsel, w["eps"])
            r64 = J.tree_force_position(tx, tm, tsel, w["eps"], w["theta"])
            torch.cuda.synchronize()

            def errs(a):
                e = (torch.linalg.norm(a - d, dim=1) / torch.linalg.norm(d, dim=1)).cpu().numpy()
                return {"mean": float(e.mean()), "median": float(np.median(e)), "p99": float(np.percentile(e, 99)),
                        "max": float(e.max())}
            accuracy = {"vs": "fp64 direct summation, 4096 sampled targets, same theta",
                        "timed_fp32_walk": errs(a32[sel]), "reference_criterion_fp64_walk": errs(r64),
                        "hybrid_kappa": J.tree_walk_hybrid(),
                        "hybrid_targets_fraction": st.get("hybrid_targets", 0) / float(n)}
            del tx, tm, a32, d, r64
            block["roofline"] = tree_roofline(n, world, ms_per_step, kernel_ms, st, mode, accuracy, peaks, peak,
                                              phases, peak_how)
        if with_cpu_baseline and world == 1:
            block["cpu_baseline"] = cpu_baseline_direct(w) if is_direct else cpu_baseline_tree(w)
    sim.close() if hasattr(sim, "close") else None
    del sim, shard
    torch.cuda.empty_cache()
    return block


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=["direct", "tree", "direct4m", "galaxy", "direct64"], default="direct")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no tree / fp64 blocks)")
    ap.add_argument("--quick", action="store_true",
                    help="experiments only: no parity check, no e2e, no cpu baseline (NOT a valid bench line)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    ctx = Ctx(args)
    cpu = not args.no_cpu_baseline
    if args.quick:
        line = run_block(ctx, args.workload, args.steps, args.warmup, False, with_e2e=False, with_parity=False)
        line["quick"] = "experiment run: parity check, e2e and cpu baseline skipped -- not a bench line"
    else:
        line = run_block(ctx, args.workload, args.steps, args.warmup, cpu)
    if args.workload == "direct" and not args.no_extras and not args.quick:
        # the other half of BASELINE.json's metric, at the same --gpus N: particle-steps/s of the tree
        tree = run_block(ctx, "tree", max(args.steps, 10), args.warmup, cpu)
        # configs[2]'s fp64 arm: ~1.07 s per step on one B200, so few steps
        f64 = run_block(ctx, "direct64", 2, 3, False, with_e2e=False)
        # BASELINE.json configs[4] (galaxy model, N = 10M) is quoted on 8 GPUs only
        galaxy = run_block(ctx, "galaxy", 10, 3, False, with_e2e=False) if ctx.world == 8 else None
        if ctx.rank == 0 and galaxy is not None:
            line["galaxy"] = galaxy
        if ctx.rank == 0:
            line["clocks"] = merge_clocks(merge_clocks(line.get("clocks"), tree.pop("clocks", None)),
                                          f64.pop("clocks", None))
            line["gpu_launches"] = int(line["gpu_launches"]) + int(tree["gpu_launches"]) + int(f64["gpu_launches"])
            line["tree"] = tree
            r = f64.get("roofline", {})
            line["fp64"] = {"metric": f64["metric"], "value": f64["value"], "unit": f64["unit"],
                            "ms_per_step": f64["ms_per_step"], "steps": f64["steps"], "warmup": f64["warmup"],
                            "frac_of_dfma_peak": r.get("frac"), "roofline": r, "config": f64["config"],
                            "parity_check": f64.get("parity_check")}
    if ctx.rank == 0:
        print(json.dumps(line))
    ctx.close()


if __name__ == "__main__":
    main()
