#!/usr/bin/env python
"""bench.py -- headline benchmark of the DQMC sweep hot path.

Metric (BASELINE.json): DetSDW O(2) L=12 beta=10 sweeps/sec.  Workload = config C3: a replica-
exchange ladder of 64 replicas (D = 288, m = 100, s = 10, delayed updates with delaySteps = 16,
weak z-flux, global shift every 10 sweeps, exchange every sweep), synthetic random fields drawn
like setupRandomField from the reference's dSFMT stream (seed 1020304050).  One "step" = one
sweep() of EVERY replica of the ladder (one direction) + one replica-exchange step; `value` is
replica-sweeps per second over the whole job (64 x ladder-sweeps/s).

  python bench.py --gpus N --steps K --warmup W            (this framework, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's CPU path on host cores)

Timed regions
  value : K steps with the random-number stream resident in HBM (dqmc_rng_preload), CUDA events on
          the launching stream, barrier + synchronize on both sides, max over ranks.
  e2e   : K steps through the plain public API with HOST buffers: every step copies the step's
          random-number window from pinned host memory to the device and reads the consumption
          cursors / control data / exchange payload back.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOAD = dict(opdim=2, L=12, m=100, s=10, dtau=0.1, c=3.0, u=1.0, lam=1.0, mu=-0.5, accRatio=0.5,
                weakZflux=True, bc=0, updateMethod=2, delaySteps=16, globalShift=True,
                globalUpdateInterval=10, seed=1020304050)
LADDER_LO, LADDER_HI = -1.9, 0.4            # r ladder, example/simulation.job:21-23 scaled to 64 values


def ladder_values(P):
    return np.linspace(LADDER_LO, LADDER_HI, P)


def workload_name(P):
    return "DetSDW O(2) L=12 beta=10 dtau=0.1 s=10, %d-replica exchange ladder (BASELINE config C3)" % P


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference / cpu baseline arm
# --------------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One replica on one host core through the UNMODIFIED reference (oracle/_ref), or the NumPy port
    when the reference library has not been built.  Returns (kind, sweeps, seconds)."""
    idx, r, sweeps, warm = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    import ref_bindings as rb
    from dqmc_oracle import SdwOracle, SdwParams
    p = SdwParams(r=float(r), rngIndex=int(idx), **WORKLOAD)
    # the first sweep of a replica contains a global-shift move (performedSweeps % interval == 0): it runs
    # untimed, so that the timed window starts one sweep after a global move like the GPU arm's windows
    if rb.available():
        rep = rb.RefSdw(p)
        for _ in range(warm):
            rep.sweep(therm=True)
        t0 = time.perf_counter()
        for _ in range(sweeps):
            rep.sweep(therm=True)
        return "reference", sweeps, time.perf_counter() - t0
    rep = SdwOracle(p)
    for _ in range(warm):
        rep.sweep_thermalization()
    t0 = time.perf_counter()
    for _ in range(sweeps):
        rep.sweep_thermalization()
    return "port", sweeps, time.perf_counter() - t0


def cpu_baseline(n_procs, sweeps_each, P, warm=1):
    """Bounded sample of the same workload on the host cores: n_procs replicas of the ladder, one
    single-threaded process each (the reference's own parallel model: one MPI rank per replica)."""
    import multiprocessing as mp
    try:                                                   # map the reference library in this process as well, so that
        import ref_bindings as rb                          # a loader hook on the parent sees which native code runs
        if rb.available():
            rb.lib()
    except Exception:
        pass
    vals = ladder_values(P)
    jobs = [(i + 1, vals[i * P // n_procs], sweeps_each, warm) for i in range(n_procs)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(n_procs) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    kind = res[0][0]
    sweep_time = max(r[2] for r in res)                     # excludes construction
    total = sum(r[1] for r in res)
    return {"value": total / sweep_time, "unit": "replica-sweeps/s", "cores": n_procs, "kind": kind,
            "per_core": total / sweep_time / n_procs, "timed_s": sweep_time,
            "sample": "%d replicas of the ladder x %d sweeps each, one single-threaded process per replica "
                      "(OPENBLAS_NUM_THREADS=1), construction and %d warm-up sweep(s) (the first one holds the global-shift "
                      "move) excluded; wall %.1f s" % (n_procs, sweeps_each, warm, wall)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_procs = max(1, min(cores, args.replicas))
    # a step of this arm = one sweep of the SAMPLED replicas (one per host core, all cores busy); --steps / --warmup are
    # honoured up to 8 / 2 (7-16 s per sweep and core: the whole run stays within a few minutes) and reported as run
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    base = cpu_baseline(n_procs, steps, args.replicas, warm)
    line = {"impl": "reference", "metric": "DetSDW O(2) L=12 beta=10 sweeps/sec", "value": base["value"],
            "unit": "replica-sweeps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * base["timed_s"] / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": workload_name(args.replicas), "replicas": args.replicas,
                       "replicas_per_step": n_procs,
                       "note": "value = replica-sweeps per second of this host (%d cores, one replica per core); a step "
                               "sweeps %d of the %d replicas, ms_per_step is the time of such a step; per core: see "
                               "cpu_baseline.per_core" % (n_procs, n_procs, args.replicas)},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "replica-sweeps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# this framework
# --------------------------------------------------------------------------------------------------
def measure_fp64_peak(torch):
    """cuBLAS DGEMM throughput (MEASURED_PEAKS.json has no FP64 figure): best of 5, 6144^3."""
    n = 6144
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    torch.cuda.empty_cache()
    return best


def run_b200(args):
    import torch
    import torch.distributed as dist
    from detqmc_b200 import DetSDWBatch, ReplicaExchangeLadder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created: keep stdout for the JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    P = args.replicas
    assert P % world == 0, "replicas must be divisible by the number of GPUs"
    n_local = P // world
    vals = ladder_values(P)
    lo = rank * n_local
    # one explicit CUDA stream for the library's kernels AND torch's copies, collectives and timing events: torch's
    # default stream has the handle 0, which dqmc_set_stream reads as "own stream" -- the exchange payload copy and
    # the all-gather would then not be ordered after the kernel that packs the payload
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)

    # DetQMCPT seeds rank p with RngWrapper(seed, (simindex+1)*(p+1)) (detqmcpt.h:301)
    batch = DetSDWBatch(dict(WORKLOAD), n_replicas=n_local, device=local_rank,
                        rng_indices=[lo + i + 1 for i in range(n_local)], r_values=vals[lo:lo + n_local],
                        stream=stream.cuda_stream)
    lad = ReplicaExchangeLadder(vals, n_local, rank, world)
    payload = torch.zeros(lad.payload_len, dtype=torch.float64, device="cuda")
    gathered = torch.zeros(world * lad.payload_len, dtype=torch.float64, device="cuda")
    host_gathered = torch.zeros(world * lad.payload_len, dtype=torch.float64).pin_memory()

    def exchange():
        batch.exchange_pack(payload.data_ptr(), lad.n_uniforms)
        if world > 1:
            dist.all_gather_into_tensor(gathered, payload)
            host_gathered.copy_(gathered, non_blocking=True)
        else:
            host_gathered.copy_(payload, non_blocking=True)
        stream.synchronize()
        r_new, ctrl_new, used = lad.walk(host_gathered.numpy())
        batch.exchange_apply(r_new, ctrl_new, used)

    def step():
        batch.sweepThermalization()
        exchange()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W, K = args.warmup, args.steps
    gint = WORKLOAD["globalUpdateInterval"]

    def align():
        """Every timed window (value, e2e, and the reference arm's) starts one sweep after a global-shift move
        (DetSDW::globalMove runs before every `globalUpdateInterval`-th sweep and costs a full re-setup), so
        that the windows hold the same number of them: floor((K - 1 + 1) / interval) for K steps."""
        n = 0
        while batch.sweep_state()["performedSweeps"] % gint != 1:
            step()
            n += 1
        return n

    def global_moves_between(s0, s1):
        return sum(1 for x in range(s0, s1) if x % gint == 0)

    # ---------------------------------------------------------------- value: stream resident in HBM
    batch.rng_preload(W + K + 2 + gint)
    for _ in range(W):
        step()
    align()
    sweeps0 = batch.sweep_state()["performedSweeps"]
    launches0 = batch.launch_count
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    gpu_launches = batch.launch_count - launches0
    gm_value = global_moves_between(sweeps0, batch.sweep_state()["performedSweeps"])
    batch.rng_release()

    # ---------------------------------------------------------------- e2e: host buffers every step
    step()                                                   # warm the non-resident path
    align()
    sweeps0 = batch.sweep_state()["performedSweeps"]
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(K):
        step()
    f1.record(stream)
    barrier()
    wall_e2e = max_over_ranks(time.perf_counter() - t0)
    gm_e2e = global_moves_between(sweeps0, batch.sweep_state()["performedSweeps"])
    rng_cap = batch.m * batch.N * (batch.opdim + 1)
    h2d = n_local * rng_cap * 8 + n_local * (8 + 840)
    d2h = n_local * (4 + 840) + 4 + world * lad.payload_len * 8

    # ---------------------------------------------------------------- per-kernel-family device time
    lanes_default = int(os.environ.get("DQMC_LANES", str(min(n_local, 32))))   # the library's default
    batch.set_lanes(1)                                      # per-kernel times are only meaningful without overlap
    batch.profile_enable(True)
    acc0 = batch.accepted_total().astype(np.float64).sum()
    prof_steps = 2
    for _ in range(prof_steps):
        step()
    prof = batch.profile_get()
    acc_per_step = (batch.accepted_total().astype(np.float64).sum() - acc0) / prof_steps
    batch.profile_enable(False)
    batch.set_lanes(lanes_default)

    # ---------------------------------------------------------------- weak-scaling line (N > 1 only): 64 replicas PER GPU.  Last: creating a second
    # context changes the library's process-wide batch-size hints (tile shapes, programmatic launch)
    weak = None
    if world > 1 and not args.no_weak:
        try:
            Pw = args.replicas * world
            valsw = ladder_values(Pw)
            low = rank * args.replicas
            bw = DetSDWBatch(dict(WORKLOAD), n_replicas=args.replicas, device=local_rank,
                             rng_indices=[low + i + 1 for i in range(args.replicas)],
                             r_values=valsw[low:low + args.replicas], stream=stream.cuda_stream)
            ladw = ReplicaExchangeLadder(valsw, args.replicas, rank, world)
            pw = torch.zeros(ladw.payload_len, dtype=torch.float64, device="cuda")
            gw = torch.zeros(world * ladw.payload_len, dtype=torch.float64, device="cuda")
            hw = torch.zeros(world * ladw.payload_len, dtype=torch.float64).pin_memory()

            def stepw():
                bw.sweepThermalization()
                bw.exchange_pack(pw.data_ptr(), ladw.n_uniforms)
                dist.all_gather_into_tensor(gw, pw)
                hw.copy_(gw, non_blocking=True)
                stream.synchronize()
                r_new, ctrl_new, used = ladw.walk(hw.numpy())
                bw.exchange_apply(r_new, ctrl_new, used)

            bw.rng_preload(W + K + 2 + gint)
            for _ in range(W):
                stepw()
            while bw.sweep_state()["performedSweeps"] % gint != 1:
                stepw()
            barrier()
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record(stream)
            for _ in range(K):
                stepw()
            w1.record(stream)
            barrier()
            ms_w = max_over_ranks(w0.elapsed_time(w1))
            weak = {"replicas": Pw, "replicas_per_gpu": args.replicas, "value": Pw * K / (ms_w * 1e-3),
                    "unit": "replica-sweeps/s", "ms_per_step": ms_w / K,
                    "note": "same step on a %d-value ladder, %d replicas per GPU (fixed per-GPU work); reported beside the "
                            "64-replica metric, not instead of it" % (Pw, args.replicas)}
            bw.rng_release()
            bw.close()
            del bw
        except Exception as exc:         # the headline line must survive a failure of the extra measurement
            weak = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    D, N, msf, R = batch.D, batch.N, 2, n_local
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6.65 TB/s (of fallback)"
    fp64_peak = measure_fp64_peak(torch)

    fam = {}
    total_kernel_ms = sum(v[0] for v in prof.values())
    for name, (ms, cnt) in prof.items():
        if cnt == 0:
            continue
        per = ms / cnt
        entry = {"launches_per_step": cnt / prof_steps, "ms_per_step": ms / prof_steps, "avg_launch_ms": per,
                 "share_of_kernel_time": ms / total_kernel_ms}
        if name == "cb_mult":
            # single-slice launches (the wraps): one read + one write of every matrix per launch
            by = 2.0 * D * D * 16 * R
            entry.update(bound="hbm", achieved=by / (per * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s")
        elif name == "cb_chain":
            # s-slice chains of the advance steps: 48 FP64 instructions (96 flop) per complex element and slice
            fl = 96.0 * D * D * WORKLOAD["s"] * R
            entry.update(bound="tensor", achieved=fl / (per * 1e-3) / 1e12, peak=fp64_peak, unit="TFLOP/s",
                         note="FP64 vector pipe / shared-memory bound, not a tensor-core kernel")
        elif name == "gemm_dmma":
            fl = 8.0 * D ** 3 * R
            entry.update(bound="tensor", achieved=fl / (per * 1e-3) / 1e12, peak=fp64_peak, unit="TFLOP/s")
        elif name == "qrcp_factor":
            fl = 16.0 / 3.0 * D ** 3 * R
            entry.update(bound="tensor", achieved=fl / (per * 1e-3) / 1e12, peak=fp64_peak, unit="TFLOP/s")
        elif name == "qr_form_q":
            # explicit Q (16/3 D^3) in the chain steps, Q^H C (8 D^3) in the Green's functions: ~ (16/3 * 15 + 8 * 10.5) / 25.5
            fl = (16.0 / 3.0 * 15.0 + 8.0 * 10.5) / 25.5 * D ** 3 * R
            entry.update(bound="tensor", achieved=fl / (per * 1e-3) / 1e12, peak=fp64_peak, unit="TFLOP/s")
        elif name == "trsm_upper":
            fl = 4.0 * D ** 3 * R
            entry.update(bound="tensor", achieved=fl / (per * 1e-3) / 1e12, peak=fp64_peak, unit="TFLOP/s")
        elif name == "update_slice":
            # window rounds: per accepted update a rank-MSF update of the future part of the (MSF w)^2 window block
            # (on average a third of it) + the K x K coefficient recurrences; a strictly sequential Metropolis chain:
            # latency bound by construction, one CTA per replica
            wp = msf * 2 * WORKLOAD["delaySteps"]
            fl_step = acc_per_step * 8.0 * msf * (wp * wp / 3.0 + 2.0 * msf * WORKLOAD["delaySteps"] * wp)
            fl = fl_step / (cnt / prof_steps)
            entry.update(bound="tensor", achieved=fl / (per * 1e-3) / 1e12, peak=fp64_peak, unit="TFLOP/s",
                         note="sequential Metropolis chain on one SM per replica: latency bound")
        elif name == "update_build_xy":
            # gather: per accepted update MSF columns of G copied (read + write), MSF rows read, MSF rows of Y written
            by = acc_per_step * msf * 4.0 * D * 16 / (cnt / prof_steps)
            entry.update(bound="hbm", achieved=by / (per * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s",
                         note="strided row gathers + K x K x D products on the FP64 pipe; launches of finished rounds are empty")
        elif name == "update_flush":
            # rank-K flush G += X Y: 8 D^2 MSF flop per accepted update (launches of finished rounds are empty)
            fl = acc_per_step * 8.0 * D * D * msf / (cnt / prof_steps)
            entry.update(bound="tensor", achieved=fl / (per * 1e-3) / 1e12, peak=fp64_peak, unit="TFLOP/s")
        if "achieved" in entry:
            entry["frac"] = entry["achieved"] / entry["peak"]
        fam[name] = entry
    dominant = max((n for n in fam if "frac" in fam[n]), key=lambda n: fam[n]["ms_per_step"])
    dom = fam[dominant]
    traffic = None
    for tr_name in ("traffic_r02.json", "traffic_r01.json"):       # dram bytes per launch from the ncu captures
        tr_path = os.path.join(ROOT, "profiles", tr_name)
        if os.path.exists(tr_path):
            try:
                tr = json.load(open(tr_path))
                for fname, entry in fam.items():
                    if fname in tr:
                        entry["traffic"] = tr[fname]
                traffic = tr.get(dominant)
            except Exception:
                traffic = None
            break
    roofline = {"kernel": dominant, "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"],
                "unit": dom["unit"], "frac": dom["frac"], "traffic": traffic,
                "peak_source": hbm_src if dom["bound"] == "hbm" else
                "cuBLAS DGEMM 6144^3 measured live in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "families": fam}

    cores = os.cpu_count() or 1
    if args.no_cpu_baseline:
        base = None
    else:
        base = cpu_baseline(max(1, min(cores, P)), 1, P)

    value = P * K / (ms_total * 1e-3)
    line = {"metric": "DetSDW O(2) L=12 beta=10 sweeps/sec", "value": value, "unit": "replica-sweeps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": workload_name(P), "replicas": P, "replicas_per_gpu": n_local,
                       "ladder_sweeps_per_s": K / (ms_total * 1e-3),
                       "cache": "working set per step (G, UDT storage, fields of %d replicas: %.1f GB) exceeds "
                                "the 126 MB L2; no flush needed" % (n_local, n_local * (2 * 11 + 8) * D * D * 16 / 1e9),
                       "parallelism": "replicas partitioned contiguously, %d per GPU, issued as %d lanes (CUDA streams) per GPU"
                                      % (n_local, lanes_default),
                       "global_shift_moves_in_timed_steps": {"value": gm_value, "e2e": gm_e2e,
                                                             "note": "every timed window starts one sweep after a "
                                                                     "global-shift move (1 per %d sweeps)" % gint}},
            "clocks": clocks, "gpu_launches": int(gpu_launches),
            "e2e": {"value": P * K / wall_e2e, "unit": "replica-sweeps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "device_ms_per_step": f0.elapsed_time(f1) / K},
            "roofline": roofline}
    if base is not None:
        line["cpu_baseline"] = base
    if weak is not None:
        line["weak_scaling"] = weak
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicas", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the additional weak-scaling measurement of multi-GPU runs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
