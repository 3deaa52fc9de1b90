#!/usr/bin/env python
"""bench.py -- the hot path's headline number on B200, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload direct|tree] [--impl reference]

Default workload (BASELINE.json configs[2], the configuration the metric and the north_star
target are quoted on): Plummer sphere N = 1,048,576, direct summation, fp32 pair arithmetic,
one full DKD leapfrog step per "step" (force on all N particles from all N + kick + drift, state
resident in HBM).  metric = pairwise interactions/s = N^2 per step.  With --gpus N > 1 (launched
by torchrun, one rank per GPU) the targets are sharded N/P per rank and each step all-gathers the
half-drifted positions (strong scaling: the total work is fixed).

--workload tree: BASELINE.json configs[3]: Hernquist N = 4,194,304, Barnes-Hut theta = 0.7, fp32
walk; metric = particle-steps/s.

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
    if kind == "direct":
        x, v, m = ic_raw.Plummer(N_DIRECT, 1e-3, 1e6, seed=42)
        return dict(x=np.ascontiguousarray(x), v=np.ascontiguousarray(v), m=m, eps=5e-5, dt=0.005,
                    theta=0.7, alg="direct", prec="fp32",
                    name="Plummer N=%d b=1pc M=1e6Msun eps=0.05pc dt=0.005Myr, direct summation fp32, "
                         "1 DKD leapfrog step (BASELINE.json configs[2])" % N_DIRECT)
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
    sel = np.random.default_rng(0).choice(len(w["m"]), ntargets, replace=False)
    t0 = time.perf_counter()
    if ref is not None:
        ref.tree_force_position(w["x"], w["m"], w["x"][sel], w["eps"], w["theta"])
        kind = "reference"
    else:
        O.tree_force_position(w["x"], w["m"], w["x"][sel], w["eps"], w["theta"], nthreads=1)
        kind = "port"
    dt = time.perf_counter() - t0
    # one evaluation = build (all N) + walk (sampled targets); extrapolate the walk to all N targets
    return {"value": ntargets / dt, "unit": "particle-steps/s", "cores": 1, "kind": kind,
            "sample": "tree build over %d sources + walk of %d targets, %.1f s (build included once)"
                      % (len(w["m"]), ntargets, dt)}


def tree_roofline(n, world, ms_per_step, kernel_ms, st, mode, accuracy, peaks, fp32_peak_tflops):
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
               "~1.9x shorter) / CUDA-event time of the walk kernel, against the FP32 FMA peak; "
               "visited_per_target = entries tested by the target's group (shared by its 32 targets); the build "
               "(ms_per_step - kernel_ms) is HBM-streaming bound" % acc_per)
        bound = "fp32_fma (list evaluation) + issue (traversal)"
    else:
        kernel = "walk_kernel"
        how = ("20 flop x accepted nodes (the reference's own accepted set: %.0f per target) / CUDA-event time "
               "of the walk kernel, against the FP32 FMA peak; ncu (profiles/) shows this walk is entry-load "
               "latency / instruction-issue bound; the build (ms_per_step - kernel_ms) is HBM-streaming bound"
               % acc_per)
        bound = "issue (walk); fp32_fma peak quoted"
    # dram__bytes_read.sum + dram__bytes_write.sum of walk_group_kernel at N = 2^22 on one GPU from
    # `ncu --set full` (profiles/r01_walk_group_f32_N4M_v2.txt): 591.8 MB + 192.3 MB; the
    # algorithmic bytes are 32 B x 6.2M entries read once + 64 B x N targets/epilogue = 0.47 GB
    traffic = 784.2e6 if (mode == "group" and n == (1 << 22) and world == 1) else None
    # the build (everything of the step that is not the walk kernel) against the HBM roofline:
    # SURVEY 8d's algorithmic bytes per particle-step, 190 + 24 x radix passes (8) = 382 B
    build_ms = ms_per_step - kernel_ms
    hbm_peak = float(peaks.get("hbm_gbs") or 0.0) or 6650.0
    build_gbs = 382.0 * n / (build_ms * 1e-3) / 1e9
    build_roofline = {"bound": "hbm", "achieved": build_gbs, "peak": hbm_peak, "unit": "GB/s",
                      "frac": build_gbs / hbm_peak, "ms": build_ms,
                      "how": "382 B per particle (SURVEY 8d: 190 + 24 x 8 radix passes) x N / (ms_per_step - walk "
                             "kernel ms); peak = %s" % ("MEASURED_PEAKS.json hbm_gbs (of measured)"
                                                        if peaks.get("hbm_gbs")
                                                        else "6.65 TB/s (of fallback, B200_PROFILING.md)")}
    return {"bound": bound, "achieved": achieved, "peak": fp32_peak_tflops,
            "unit": "TFLOP/s", "frac": achieved / fp32_peak_tflops, "traffic": traffic,
            "kernel": kernel, "kernel_ms": kernel_ms, "walk": mode,
            "accepted_per_target": acc_per, "visited_per_target": vis_per,
            "tree_entries": st["entries"], "tree_cells": st["cells"], "deepest_level": st["maxlevel"],
            "build_ms": build_ms, "build_roofline": build_roofline, "accuracy": accuracy, "how": how}


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
                sel = rng.choice(n, per, replace=False)
                jobs.append(("direct" if wl_direct else "tree", w["x"][sel]))
            t0 = time.perf_counter()
            pool.map(_ref_worker, jobs)
            if s >= args.warmup:
                times.append(time.perf_counter() - t0)
    tot = sum(times)
    units = (per * P * n) if wl_direct else (per * P)
    value = units * args.steps / tot
    unit = "interactions/s" if wl_direct else "particle-steps/s"
    sample = ("%d processes x %d targets x %d sources per step (direct_summation_position)" % (P, per, n)
              if wl_direct else
              "%d processes, each: tree build over %d sources + walk of %d targets per step" % (P, n, per))
    line = {"impl": "reference", "metric": "pairwise interactions/s" if wl_direct else "particle-steps/s",
            "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"]},
            "cpu_baseline": {"value": value, "unit": unit, "cores": P, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=["direct", "tree", "direct4m", "galaxy"], default="direct")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from gravhopper_b200 import _jbgrav as J, _lib
    from gravhopper_b200.sharded import ShardedSimulation

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.require_gpu()

    w = workload(args.workload)
    n = len(w["m"])
    sim = ShardedSimulation(w["x"], w["v"], w["m"], w["dt"], w["eps"], algorithm=w["alg"], theta=w["theta"],
                            precision=w["prec"], rank=rank, world=world, device=local)
    shard = sim.shard
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        with shard.stream_context():
            flush.zero_()  # evict L2 between steps (the 16-24 MB source array would otherwise stay hot)
        sim.step()

    for _ in range(args.warmup):
        one_step()
    barrier()
    launches0 = shard.launches()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with shard.stream_context():
        ev0.record()
    for _ in range(args.steps):
        one_step()
    with shard.stream_context():
        ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = shard.launches() - launches0
    clocks = sampler.stop() if sampler else None
    # per-kernel time of the dominant (force) kernel: events the engine records around it, last step
    kernel_ms = shard.last_force_ms()
    t = torch.tensor([ms_total, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps

    is_direct = w["alg"] == "direct"
    if is_direct:
        units_per_step = float(n) * float(n)
        metric, unit = "pairwise interactions/s", "interactions/s"
    else:
        units_per_step = float(n)
        metric, unit = "particle-steps/s", "particle-steps/s"
    value = units_per_step / (ms_per_step * 1e-3)

    # ---- end to end through the reference-facing call with HOST (pinned) buffers ----
    # Every rank evaluates its share of the targets against all sources through the public
    # _jbgrav call with host arrays (H2D of the sources + its targets, D2H of its accelerations
    # inside the timed region); the job's rate is all units / the slowest rank's time.
    hx = torch.from_numpy(w["x"]).pin_memory().numpy()
    hm = torch.from_numpy(w["m"]).pin_memory().numpy()
    b0, cnt = sim.begin, sim.count
    ht = torch.from_numpy(np.ascontiguousarray(w["x"][b0:b0 + cnt])).pin_memory().numpy()
    if world == 1:
        if is_direct:
            call = lambda: J.direct_summation(hx, hm, w["eps"], precision=w["prec"])  # noqa: E731
        else:
            call = lambda: J.tree_force(hx, hm, w["eps"], w["theta"], precision=w["prec"])  # noqa: E731
        name = "direct_summation" if is_direct else "tree_force"
        h2d = int(hx.nbytes + hm.nbytes)
    else:
        if is_direct:
            call = lambda: J.direct_summation_position(hx, hm, ht, w["eps"], precision=w["prec"])  # noqa: E731
        else:
            call = lambda: J.tree_force_position(hx, hm, ht, w["eps"], w["theta"], precision=w["prec"])  # noqa: E731
        name = "direct_summation_position" if is_direct else "tree_force_position"
        h2d = int(hx.nbytes + hm.nbytes + ht.nbytes)
    # W >= 3 untimed calls: the stateless path's device scratch and the binding's pool of page-locked
    # result blocks (two alternate while `out` is rebound) reach steady state before the clock starts
    for _ in range(1 if (is_direct and n > (1 << 21)) else 3):
        out = call()
    reps = (1 if n > (1 << 21) else 3) if is_direct else 5
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = call()
    te = (time.perf_counter() - t0) / reps
    tt = torch.tensor([te], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    te = float(tt[0])
    e2e = {"value": units_per_step / te, "unit": unit, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": int(out.nbytes), "ms_per_call": te * 1e3, "n_gpus_used": world,
           "call": "_jbgrav.%s(host ndarrays) -> host ndarray, one call per rank on its share of the targets"
                   % name}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    fp32_peak_tflops = SM_COUNT * FP32_LANES_PER_SM * 2 * sm_max * 1e6 / 1e12
    if is_direct:
        per_rank_units = units_per_step / world
        achieved = per_rank_units * FLOP_PER_INTERACTION / (kernel_ms * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at N = 2^20 on one GPU from
        # `ncu --set full` (profiles/r01_direct_f32_N1M.txt): 24.4 MB + 76.5 MB
        traffic = 101.0e6 if (n == (1 << 20) and world == 1) else None
        roofline = {"bound": "fp32_fma", "achieved": achieved, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                    "frac": achieved / fp32_peak_tflops, "traffic": traffic,
                    "traffic_unit": "bytes per launch (ncu); algorithmic: 16.8 MB sources read + 24 MB x S partial sums",
                    "kernel": "direct_f32_kernel", "kernel_ms": kernel_ms,
                    "how": "20 flop/interaction x N_i x N_j per launch / CUDA-event time of the force kernel; "
                           "peak = 148 SMs x 128 FP32 lanes x 2 flop x clocks.max.sm (%.0f MHz); no measured FP32 "
                           "figure exists in MEASURED_PEAKS.json (it holds HBM GB/s and bf16 TF/s)" % sm_max}
        if clocks:
            roofline["frac_at_observed_clock"] = achieved / (fp32_peak_tflops * clocks["sm_mhz"] / sm_max)
    else:
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
        sel = torch.from_numpy(np.random.default_rng(0).choice(n, 4096, replace=False)).cuda()
        tsel = tx[sel].contiguous()
        d = J.direct_summation_position(tx, tm, tsel, w["eps"])
        r64 = J.tree_force_position(tx, tm, tsel, w["eps"], w["theta"])
        torch.cuda.synchronize()

        def errs(a):
            e = (torch.linalg.norm(a - d, dim=1) / torch.linalg.norm(d, dim=1)).cpu().numpy()
            return {"mean": float(e.mean()), "median": float(np.median(e)), "p99": float(np.percentile(e, 99)),
                    "max": float(e.max())}
        accuracy = {"vs": "fp64 direct summation, 4096 sampled targets, same theta",
                    "timed_fp32_walk": errs(a32[sel]), "reference_criterion_fp64_walk": errs(r64)}
        roofline = tree_roofline(n, world, ms_per_step, kernel_ms, st, mode, accuracy, peaks, fp32_peak_tflops)

    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "n_particles": n, "precision": w["prec"],
                       "parallelism": "targets sharded over %d rank(s), NCCL all-gather of x_half per step" % world,
                       "l2": "256 MB flush write between steps"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_direct(w) if is_direct else cpu_baseline_tree(w)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
