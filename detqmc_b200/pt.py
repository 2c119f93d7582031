"""DetQMCPT without MPI: the parallel-tempering driver loop of the reference (DetQMCPT<Model, ModelParams>::run,
detqmcpt.h:760-958) over replicas batched per GPU, SURVEY.md section 8(f) row 4.

One process per GPU (torchrun) holds P / world replicas of the ladder in one DetSDWBatch; per exchange step the
actions, the look-ahead uniforms of replica 0 and the control blobs are packed on the device, all-gathered
(torch.distributed: NCCL over NVLink on the GPUs, gloo in the CPU tests of the host logic) and every rank performs the
identical serial ladder walk (ReplicaExchangeLadder).  Observables are recorded per CONTROL-PARAMETER index, as the
reference's ScalarObservableHandlerPT does (mpiobservablehandlerpt.cpp:80-192): a replica contributes to the series of
the parameter it currently holds.  Output layout follows the reference: one sub-directory
p<index>_<name><value> per control parameter (detqmcpt.h:655-660) with <observable>.series, results.values and the
configuration streams, plus exchange-parameters.values / exchange-acceptance.values / exchange-diffusion.values
(detqmcpt.h:596-651) in the working directory.

The bosonic observables (normMeanPhi, associatedEnergy, phiRhoS_Gs, phiRhoS_Gc: the reference's list with
turnoffFermionMeasurements) are always measured; with turnoffFermionMeasurements=False the measurement sweeps are
sweep(true) and the fermionic scalars / vectors of DetSDW::measure are recorded as well.
"""
import os

import numpy as np

from .sdw import DetSDWBatch, ReplicaExchangeLadder

OBSERVABLES = ("normMeanPhi", "associatedEnergy", "phiRhoS_Gs", "phiRhoS_Gc")
FERMIONIC_SCALARS = ("pairPlusMax", "pairMinusMax", "greenK0", "greenLocal", "occDiffSq")      # detsdwopdim.cpp:278-333
FERMIONIC_VECTORS = ("kOccX", "kOccY", "pairPlus", "pairMinus")


def num_to_string(v):
    """tools.h:46-50 numToString: default ostream formatting (6 significant digits)."""
    return "%g" % v


def bosonic_observables(phi, dtau):
    """The observables DetSDW measures with turnoffFermionMeasurements (initMeasurements / measure /
    finishMeasurements, detsdwopdim.cpp:441-560, 903-918; same as DetSDWGpu::measureBosonic in include/detsdw_gpu.h):
    normMeanPhi, associatedEnergy and, for opdim == 2, the bosonic spin stiffness sums phiRhoS_Gs / phiRhoS_Gc
    (zero otherwise).  phi: [m+1][opdim][N]."""
    m1, opdim, N = phi.shape
    L = int(round(np.sqrt(N)))
    ph = phi[1:]
    out = {"normMeanPhi": float(np.linalg.norm(ph.sum(axis=(0, 2)) / (N * (m1 - 1)))),
           "associatedEnergy": float(np.sum(ph * ph) / (2.0 * N * (m1 - 1))), "phiRhoS_Gs": 0.0, "phiRhoS_Gc": 0.0}
    if opdim == 2:
        sites = np.arange(N)
        x, y = sites % L, sites // L
        xp = y * L + (x + 1) % L
        yp = ((y + 1) % L) * L + x
        out["phiRhoS_Gc"] = float(0.5 * dtau * (np.sum(ph * ph[:, :, xp]) + np.sum(ph * ph[:, :, yp])))
        out["phiRhoS_Gs"] = float(dtau * np.sum(ph[:, 0, xp] * ph[:, 1, :] - ph[:, 1, xp] * ph[:, 0, :]))
    return out


class DetQMCPT:
    def __init__(self, model_pars, control_values, thermalization, sweeps, measureInterval=1, exchangeInterval=1,
                 saveConfigurationStreamInterval=0, saveConfigurationStreamBinary=False,
                 saveConfigurationStreamText=False, rngSeed=1020304050, simindex=0, outdir=".",
                 controlParameterName="r", device=None, make_batch=None, turnoffFermionMeasurements=True,
                 resume=False):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.values = np.ascontiguousarray(control_values, dtype=np.float64)
        self.P = len(self.values)
        if controlParameterName != "r":
            raise ValueError("DetSDW exchanges the parameter r (detsdwopdim.cpp:5199-5202)")
        if self.P % self.world:
            raise ValueError("the ladder must divide evenly over the ranks")
        self.n_local = self.P // self.world
        self.lo = self.rank * self.n_local
        self.thermalization, self.sweeps = int(thermalization), int(sweeps)
        self.measureInterval, self.exchangeInterval = int(measureInterval), int(exchangeInterval)
        self.cfgInterval = int(saveConfigurationStreamInterval)
        self.cfgBinary, self.cfgText = bool(saveConfigurationStreamBinary), bool(saveConfigurationStreamText)
        self.outdir = outdir
        self.fermionic = not turnoffFermionMeasurements
        self.name = controlParameterName
        pars = dict(model_pars if isinstance(model_pars, dict) else vars(model_pars))
        pars["seed"] = rngSeed
        # DetQMCPT seeds process p with RngWrapper(rngSeed, (simindex + 1) * (p + 1)) (detqmcpt.h:301)
        idx = [(simindex + 1) * (self.lo + i + 1) for i in range(self.n_local)]
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        # one explicit CUDA stream shared by the library's kernels and torch's copies / collectives (torch's default
        # stream has the handle 0, which dqmc_set_stream reads as "own stream": the two would not be ordered)
        torch.cuda.set_device(device)
        stream = torch.cuda.Stream(device=device)
        make = make_batch or DetSDWBatch
        self.batch = make(pars, n_replicas=self.n_local, device=device, rng_indices=idx,
                          r_values=self.values[self.lo:self.lo + self.n_local], stream=stream.cuda_stream)
        self.stream = stream
        self.ladder = ReplicaExchangeLadder(self.values, self.n_local, self.rank, self.world)
        L = self.ladder
        self.payload = torch.zeros(L.payload_len, dtype=torch.float64, device="cuda")
        self.gathered = torch.zeros(self.world * L.payload_len, dtype=torch.float64, device="cuda")
        self.host_gathered = torch.zeros(self.world * L.payload_len, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()
        self.sweepsDone = 0
        self.sweepsDoneThermalization = 0
        self.swCounter = 0
        # local records: (sweep index, control-parameter index, {observable: value}) and buffered configurations
        self.records = []
        self.configs = []
        self._rng_indices = idx
        if resume and os.path.exists(self.state_path()):
            self.load_state()

    # ------------------------------------------------------------------ saveState / resume, detqmcpt.h:201-253, 447-520
    def state_path(self):
        return os.path.join(self.outdir, "simulation.state.rank%d" % self.rank)

    def save_state(self):
        """Every rank dumps its share of what DetQMCPT::saveContents serialises: the replicas (fields, control data,
        generator position), the sweep counters, the ladder (current_process_par / current_par_process) with its
        exchange statistics, and the measurements recorded so far."""
        import pickle
        L = self.ladder
        st = dict(batch=self.batch.save_state(), sweepsDone=self.sweepsDone,
                  sweepsDoneThermalization=self.sweepsDoneThermalization, swCounter=self.swCounter,
                  ladder={k: getattr(L, k).copy() for k in ("par_process", "process_par", "proposed", "accepted", "going",
                                                            "count_up", "count_down")},
                  records=self.records, configs_written=True, values=self.values.copy())
        os.makedirs(self.outdir, exist_ok=True)
        with open(self.state_path(), "wb") as f:
            pickle.dump(st, f)

    def load_state(self):
        import pickle
        with open(self.state_path(), "rb") as f:
            st = pickle.load(f)
        if not np.array_equal(st["values"], self.values):
            raise ValueError("state file belongs to a different ladder")
        self.batch.load_state(st["batch"], self._rng_indices)
        self.sweepsDone, self.sweepsDoneThermalization = st["sweepsDone"], st["sweepsDoneThermalization"]
        self.swCounter = st["swCounter"]
        for k, v in st["ladder"].items():
            getattr(self.ladder, k)[...] = v
        self.records = st["records"]
        print("State of previous simulation has been loaded.\n  sweepsDoneThermalization: %d\n  sweepsDone: %d"
              % (self.sweepsDoneThermalization, self.sweepsDone))

    # ------------------------------------------------------------------ replicaExchangeStep, detqmcpt.h:962-1118
    def replica_exchange_step(self):
        L = self.ladder
        self.batch.exchange_pack(self.payload.data_ptr(), L.n_uniforms)
        with self.torch.cuda.stream(self.stream):
            if self.world > 1:
                if self.dist.get_backend() == "nccl":
                    self.dist.all_gather_into_tensor(self.gathered, self.payload)
                else:                                # gloo (tests of the multi-rank logic): through the host
                    self.stream.synchronize()
                    parts = [self.torch.zeros(L.payload_len, dtype=self.torch.float64) for _ in range(self.world)]
                    self.dist.all_gather(parts, self.payload.cpu())
                    self.gathered.copy_(self.torch.cat(parts))
                self.host_gathered.copy_(self.gathered, non_blocking=True)
            else:
                self.host_gathered.copy_(self.payload, non_blocking=True)
        self.stream.synchronize()
        r_new, ctrl_new, used = L.walk(self.host_gathered.numpy())
        self.batch.exchange_apply(r_new, ctrl_new, used)

    def local_parameter_indices(self):
        return self.ladder.process_par[self.lo:self.lo + self.n_local].copy()

    def _measure(self):
        b = self.batch
        cpis = self.local_parameter_indices()
        for i in range(self.n_local):
            obs = bosonic_observables(b.phi(i), b.pars["dtau"])
            if self.fermionic:                         # finishMeasurements of the sweep(True) that just ran
                obs.update(b.fermionic_observables(i))
            self.records.append((self.sweepsDone, int(cpis[i]), obs))

    def _buffer_configurations(self):
        # buffer_local_system_configuration, detqmcpt.h:690-700: the configuration goes to the stream of the control
        # parameter the replica holds NOW
        streams = self.batch.config_stream(-1)
        cpis = self.local_parameter_indices()
        for i in range(self.n_local):
            self.configs.append((self.sweepsDone, int(cpis[i]), streams[i].copy()))

    # ------------------------------------------------------------------ run, detqmcpt.h:760-958
    def run(self):
        b = self.batch
        while self.sweepsDoneThermalization < self.thermalization or self.sweepsDone < self.sweeps:
            if self.sweepsDoneThermalization < self.thermalization:
                b.sweepThermalization()
                self.sweepsDoneThermalization += 1
                self.swCounter += 1
                if self.sweepsDoneThermalization == self.thermalization:
                    self.swCounter = 0
            else:
                self.swCounter += 1
                take = self.swCounter % self.measureInterval == 0
                b.sweep(take and self.fermionic)       # sweep(true) accumulates the fermionic observables on the device
                if take:
                    self._measure()
                    if (self.cfgBinary or self.cfgText) and self.cfgInterval and self.swCounter % self.cfgInterval == 0:
                        self._buffer_configurations()
                self.sweepsDone += 1
            # the reference exchanges only while the stage is T or M (detqmcpt.h:948-953): none after the last sweep
            finished = self.sweepsDoneThermalization == self.thermalization and self.sweepsDone == self.sweeps
            if (self.exchangeInterval and not finished
                    and (self.sweepsDone + self.sweepsDoneThermalization) % self.exchangeInterval == 0):
                self.replica_exchange_step()
            # replicaExchangeConsistencyCheck, detqmcpt.h:1120-1135
            want = self.values[self.local_parameter_indices()]
            have = np.array([b.get_exchange_parameter_value(i) for i in range(self.n_local)])
            if np.abs(want - have).max() > 1e-10:
                raise RuntimeError("replica exchange consistency check failed")
        self.save()
        self.save_state()

    # ------------------------------------------------------------------ gather to rank 0 and write
    def _gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(obj, out, dst=0)
        return out

    def subdir(self, cpi):
        return os.path.join(self.outdir, "p%d_%s%s" % (cpi, self.name, num_to_string(self.values[cpi])))

    def save(self):
        recs = self._gather(self.records)
        cfgs = self._gather(self.configs)
        self.configs = []          # configuration streams are appended; the records stay (series are rewritten in full)
        if self.rank != 0:
            return
        scalars = OBSERVABLES + (FERMIONIC_SCALARS if self.fermionic else ())
        vectors = FERMIONIC_VECTORS if self.fermionic else ()
        series = {cpi: {o: [] for o in scalars + vectors} for cpi in range(self.P)}
        flat = sorted((r for part in recs for r in part), key=lambda r: (r[0], r[1]))
        for _, cpi, obs in flat:
            for o in scalars + vectors:
                series[cpi][o].append(obs[o])
        header = "## %s = %s\n## thermalization = %d\n## sweeps = %d\n## exchangeInterval = %d\n"
        for cpi in range(self.P):
            d = self.subdir(cpi)
            os.makedirs(d, exist_ok=True)
            meta = header % (self.name, num_to_string(self.values[cpi]), self.thermalization, self.sweepsDone,
                             self.exchangeInterval)
            with open(os.path.join(d, "results.values"), "w") as res:
                res.write(meta + "## observable \t value \t error\n")
                for o in vectors:                      # vector observables: mean and error per component
                    v = np.asarray(series[cpi][o])
                    with open(os.path.join(d, "results-%s.values" % o), "w") as f:
                        f.write(meta + "## %s: index \t value \t error\n" % o)
                        for idx in range(v.shape[1] if v.ndim == 2 else 0):
                            col = v[:, idx]
                            err = float(col.std(ddof=1) / np.sqrt(len(col))) if len(col) > 1 else 0.0
                            f.write("%d\t%.15g\t%.15g\n" % (idx, float(col.mean()), err))
                for o in scalars:
                    v = np.asarray(series[cpi][o])
                    with open(os.path.join(d, o + ".series"), "w") as f:
                        f.write(meta + "## time series of %s\n" % o)
                        f.write("".join("%.15g\n" % x for x in v))
                    err = float(v.std(ddof=1) / np.sqrt(len(v))) if len(v) > 1 else 0.0
                    res.write("%s\t%.15g\t%.15g\n" % (o, float(v.mean()) if len(v) else 0.0, err))
        # every stream in the order the configurations were buffered (detqmcpt.h:702-757), whichever rank held them
        for _, cpi, cfg in sorted((c for part in cfgs for c in part), key=lambda c: (c[0], c[1])):
            d = self.subdir(cpi)
            if self.cfgBinary:
                with open(os.path.join(d, "configs-phi.binarystream"), "ab") as f:
                    f.write(cfg.tobytes())
            if self.cfgText:
                with open(os.path.join(d, "configs-phi.textstream"), "a") as f:
                    f.write("".join("%.14e\n" % v for v in cfg))
        L = self.ladder
        with open(os.path.join(self.outdir, "exchange-parameters.values"), "w") as f:
            f.write("## Control parameter values\n## control parameter index \t control parameter value\n")
            f.write("".join("%d\t%.15g\n" % (c, self.values[c]) for c in range(self.P)))
        with open(os.path.join(self.outdir, "exchange-acceptance.values"), "w") as f:
            f.write("## Acceptance ratio of exchanging replicas at control parameters (upwards)\n"
                    "## control parameter index \t acceptance ratio\n")
            for c in range(self.P):
                prop = L.proposed[c] if c < self.P - 1 else 0
                f.write("%d\t%.15g\n" % (c, (L.accepted[c] / prop) if prop else 0.0))
        with open(os.path.join(self.outdir, "exchange-diffusion.values"), "w") as f:
            f.write("## Diffusion fraction of replicas at control parameters: df = nUp / (nUp + nDown)\n"
                    "## control parameter index \t diffusion fraction\n")
            for c in range(self.P):
                tot = L.count_up[c] + L.count_down[c]
                f.write("%d\t%.15g\n" % (c, (L.count_up[c] / tot) if tot else 0.0))


def main(argv=None):
    """MPI-free launcher: `python -m torch.distributed.run --nproc-per-node G -m detqmc_b200.pt key=value ...` (or
    plain `python -m detqmc_b200.pt ...` on one GPU).  Keys follow the reference's option names
    (maindetqmcsdwopdim.cpp:93-135, detqmcptparams): model parameters of DetSDW, thermalization, sweeps,
    measureInterval, exchangeInterval, controlParameterValues=v0,v1,..., rngSeed, simindex, outdir,
    saveConfigurationStreamInterval / Binary / Text."""
    import sys
    import torch
    import torch.distributed as dist
    args = dict(a.split("=", 1) for a in (sys.argv[1:] if argv is None else argv))
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1 and not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    values = [float(v) for v in args.pop("controlParameterValues").split(",")]
    mc = {}
    for key, conv in (("thermalization", int), ("sweeps", int), ("measureInterval", int), ("exchangeInterval", int),
                      ("saveConfigurationStreamInterval", int), ("rngSeed", int), ("simindex", int)):
        if key in args:
            mc[key] = conv(args.pop(key))
    for key in ("saveConfigurationStreamBinary", "saveConfigurationStreamText", "turnoffFermionMeasurements", "resume"):
        if key in args:
            mc[key] = args.pop(key).lower() in ("1", "true", "yes")
    outdir = args.pop("outdir", ".")
    model = {}
    for key, val in args.items():
        if val.lower() in ("true", "false"):
            model[key] = val.lower() == "true"
        else:
            try:
                model[key] = int(val)
            except ValueError:
                model[key] = float(val)
    if "beta" in model:                                   # the reference derives m from beta and dtau
        model["m"] = int(round(model.pop("beta") / model.get("dtau", 0.1)))
    mc.setdefault("thermalization", 10)
    mc.setdefault("sweeps", 10)
    pt = DetQMCPT(model, values, outdir=outdir, **mc)
    pt.run()
    if pt.rank == 0:
        print("DetQMCPT finished: %d + %d sweeps of %d replicas on %d rank(s); exchange acceptance %s"
              % (pt.sweepsDoneThermalization, pt.sweepsDone, pt.P, pt.world,
                 ", ".join("%.2f" % (a / max(1, p)) for a, p in zip(pt.ladder.accepted[:pt.P - 1],
                                                                    pt.ladder.proposed[:pt.P - 1]))))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
