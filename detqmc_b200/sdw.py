"""Host-side mirror of the reference's model interface for the DQMC sweep hot path.

`DetSDWBatch` plays the role of DetSDW<CB_ASSAAD_BERG, OPDIM> (detsdwopdim.h:62-153) for a batch of
replicas resident on one GPU; method names follow the reference (sweep, sweepThermalization,
get_exchange_action_contribution, set_exchange_parameter_value, get/set_control_data ...).  All
numerics happen in libdqmc_b200.so through the C ABI; this file only marshals NumPy buffers.

`ReplicaExchangeLadder` plays the role of DetQMCPT::replicaExchangeStep (detqmcpt.h:962-1118) with
replicas partitioned contiguously over ranks: one all-gather of (action, look-ahead uniforms of
replica 0) per exchange step, then the identical serial ladder walk on every rank.
"""
import ctypes
import os

import numpy as np

from .lib import load_library, DqmcParams, ControlData, DqmcError, c_i32, c_vp, c_f64

OP_LEFT, OP_RIGHT, OP_LEFT_INV, OP_RIGHT_INV, OP_LEFT_ADJ = range(5)

_DEFAULTS = dict(opdim=2, L=4, m=20, s=10, dtau=0.1, r=-1.0, c=3.0, u=1.0, lam=1.0, txhor=-1.0, txver=-0.5,
                 tyhor=0.5, tyver=1.0, cdwU=0.0, mu=-0.5, accRatio=0.5, weakZflux=True, bc=0, updateMethod=2,
                 delaySteps=16, globalShift=True, globalUpdateInterval=10, repeatUpdateInSlice=1,
                 wolffClusterUpdate=False, wolffClusterShiftUpdate=False, repeatWolffPerSweep=1,
                 checkerboard=True, seed=1020304050, rngIndex=1)


def _ptr(a):
    return a.ctypes.data_as(c_vp)


def make_params(pars=None, **kw):
    """Build the C struct from an object / dict with the reference's parameter names
    (ModelParamsDetSDW, detsdwparams.h:30-120).  Unsupported settings raise, as check() does."""
    d = dict(_DEFAULTS)
    if pars is not None:
        src = pars if isinstance(pars, dict) else vars(pars)
        d.update({k: v for k, v in src.items() if k in d})
    d.update(kw)
    if d["cdwU"] != 0.0:
        raise DqmcError("cdwU != 0 is outside the accelerated path")
    if d["updateMethod"] not in (0, 1, 2):
        raise DqmcError("updateMethod must be 'iterative' (0), 'woodbury' (1) or 'delayed' (2)")
    p = DqmcParams()
    p.model = 0
    p.opdim, p.L, p.m, p.s, p.bc = d["opdim"], d["L"], d["m"], d["s"], d["bc"]
    p.weakZflux = int(bool(d["weakZflux"]))
    p.denseHopping = 0 if d.get("checkerboard", True) else 1      # checkerboard = false: DetSDW<CB_NONE>, dense e^{-dtau K}
    # woodbury == delayed with 1 step; iterative (detsdwopdim.cpp:2491-2880) evaluates the same ratio and the same
    # rank-MSF update entry by entry (the reference's fields equal woodbury's sweep by sweep): same kernel
    p.delaySteps = d["delaySteps"] if d["updateMethod"] == 2 else 1
    p.globalShift = int(bool(d["globalShift"]))
    p.globalUpdateInterval = d["globalUpdateInterval"]
    p.wolffClusterUpdate = int(bool(d["wolffClusterUpdate"]))
    p.wolffClusterShiftUpdate = int(bool(d["wolffClusterShiftUpdate"]))
    p.repeatWolffPerSweep = int(d["repeatWolffPerSweep"])
    p.repeatUpdateInSlice = int(d["repeatUpdateInSlice"])
    p.dtau, p.r, p.c, p.u, p.lambda_ = d["dtau"], d["r"], d["c"], d["u"], d["lam"]
    p.txhor, p.txver, p.tyhor, p.tyver = d["txhor"], d["txver"], d["tyhor"], d["tyver"]
    p.mux = p.muy = d["mu"]
    p.accRatio = d["accRatio"]
    return p, d


class DetSDWBatch:
    """A batch of DetSDW replicas on one B200."""

    def __init__(self, pars=None, n_replicas=1, device=0, rng_indices=None, r_values=None, init="random",
                 stream=None, full_pivot=False, **kw):
        self.lib = load_library()
        self.cpars, self.pars = make_params(pars, **kw)
        self.R = int(n_replicas)
        h = c_vp()
        rc = self.lib.dqmc_create(ctypes.byref(self.cpars), self.R, int(device), ctypes.byref(h))
        self.h = h
        if rc != 0:
            msg = self.lib.dqmc_last_error(h).decode() if h else "dqmc_create failed"
            if h:
                self.lib.dqmc_destroy(h)
                self.h = None
            raise DqmcError(msg)
        dims = (c_i32 * 8)()
        self._ck(self.lib.dqmc_dims(self.h, dims))
        self.N, self.D, self.m, self.n, self.s, self.ngc, _, self.opdim = list(dims)
        if stream is not None:
            self.set_stream(stream)
        if full_pivot:
            self.set_stabilizer(True)
        if rng_indices is None:
            rng_indices = [self.pars["rngIndex"] + i for i in range(self.R)]
        for rep, idx in enumerate(rng_indices):
            self._ck(self.lib.dqmc_rng_seed(self.h, rep, self.pars["seed"], int(idx)))
        if r_values is not None:
            for rep, r in enumerate(r_values):
                self.set_exchange_parameter_value(r, rep)
        if init == "random":
            # createReplica -> ctor: setupRandomField, then setupUdVStorage_and_calculateGreen
            for rep in range(self.R):
                self._ck(self.lib.dqmc_init_random_fields(self.h, rep))
            self.setup_storage()

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise DqmcError(self.lib.dqmc_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.dqmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.dqmc_set_stream(self.h, c_vp(int(cuda_stream_ptr) if cuda_stream_ptr else None)))

    def set_stabilizer(self, full_pivot):
        """False (default): blocked QR with column pre-pivoting; True: fully pivoted one-CTA QR."""
        self._ck(self.lib.dqmc_set_option(self.h, 0, 1 if full_pivot else 0))

    def set_lanes(self, n):
        """1..64 independent lanes (CUDA streams) per context, default one per replica up to 32; results do not depend on it."""
        self._ck(self.lib.dqmc_set_option(self.h, 1, int(n)))

    def synchronize(self):
        self._ck(self.lib.dqmc_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.lib.dqmc_launch_count(self.h))

    # ------------------------------------------------------------------ state
    def phi(self, rep=0):
        out = np.zeros((self.m + 1, self.opdim, self.N))
        self._ck(self.lib.dqmc_download_fields(self.h, rep, _ptr(out)))
        return out

    def set_phi(self, phi, rep=0):
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        assert phi.shape == (self.m + 1, self.opdim, self.N)
        self._ck(self.lib.dqmc_upload_fields(self.h, rep, _ptr(phi)))

    # ------------------------------------------------------------------ configuration streams
    def config_stream(self, rep=0):
        """getCurrentSystemConfiguration in the order of the reference's configuration streams
        (detsdwopdim.cpp:5000-5010): for ix, iy, k = 1..m, dim.  rep = -1: all replicas, shape [R][N*m*opdim]."""
        n = self.N * self.m * self.opdim
        out = np.zeros((self.R, n) if rep < 0 else n)
        self._ck(self.lib.dqmc_download_config_stream(self.h, int(rep), _ptr(out)))
        return out

    def saveConfigurationStreamBinary(self, directory=".", rep=0):
        """DetSDW::saveConfigurationStreamBinary (detsdwopdim.cpp:4991-5012): append to configs-phi.binarystream."""
        with open(os.path.join(directory, "configs-phi.binarystream"), "ab") as f:
            f.write(self.config_stream(rep).tobytes())

    def saveConfigurationStreamText(self, directory=".", rep=0):
        """DetSDW::saveConfigurationStreamText (detsdwopdim.cpp:4943-4966): precision 14, scientific, one value per line."""
        with open(os.path.join(directory, "configs-phi.textstream"), "a") as f:
            f.write("".join("%.14e\n" % v for v in self.config_stream(rep)))

    def green(self, rep=0):
        out = np.zeros((self.D, self.D), dtype=np.complex128, order="F")
        self._ck(self.lib.dqmc_download_green(self.h, rep, 0, _ptr(out)))
        return out

    def set_green(self, g, rep=0):
        g = np.asfortranarray(g, dtype=np.complex128)
        self._ck(self.lib.dqmc_upload_green(self.h, rep, 0, _ptr(g)))

    def logdet(self, rep=0):
        out = c_f64()
        self._ck(self.lib.dqmc_logdet(self.h, rep, 0, ctypes.byref(out)))
        return out.value

    def control_data(self, rep=0):
        cd = ControlData()
        self._ck(self.lib.dqmc_get_control_data(self.h, rep, ctypes.byref(cd)))
        return cd

    def set_control_data(self, cd, rep=0):
        self._ck(self.lib.dqmc_set_control_data(self.h, rep, ctypes.byref(cd)))

    def phi_delta(self, rep=0):
        return self.control_data(rep).phiDelta

    def sweep_state(self):
        out = (c_i32 * 3)()
        self._ck(self.lib.dqmc_get_sweep_state(self.h, out))
        return dict(currentTimeslice=out[0], lastSweepDir=out[1], performedSweeps=out[2])

    # ------------------------------------------------------------------ RNG
    def rng_draw(self, n, rep=0):
        out = np.zeros(n)
        self._ck(self.lib.dqmc_rng_draw(self.h, rep, n, _ptr(out)))
        return out

    def rng_peek(self, n, rep=0):
        out = np.zeros(n)
        self._ck(self.lib.dqmc_rng_peek(self.h, rep, n, _ptr(out)))
        return out

    def rng_skip(self, n, rep=0):
        self._ck(self.lib.dqmc_rng_skip(self.h, rep, n))

    def rng_consumed(self, rep=0):
        return int(self.lib.dqmc_rng_consumed(self.h, rep))

    # ------------------------------------------------------------------ checkpoint (saveContents / loadContents)
    def save_state(self):
        """State of every replica as the reference serialises it (detsdwopdim.h:1117-1148 + the generator,
        detqmcpt.h:248-250): fields, control data, exchange parameter, position in the random-number stream
        (the generator is re-seeded and advanced on load), sweep counter.  G and the UdV storage are rebuilt."""
        self.synchronize()
        reps = []
        for rep in range(self.R):
            cd = self.control_data(rep)
            reps.append(dict(phi=self.phi(rep), ctrl=bytes(cd), r=self.get_exchange_parameter_value(rep),
                             consumed=self.rng_consumed(rep)))
        return dict(replicas=reps, performedSweeps=self.sweep_state()["performedSweeps"])

    def load_state(self, state, rng_indices):
        assert len(state["replicas"]) == self.R
        for rep, st in enumerate(state["replicas"]):
            self.set_phi(st["phi"], rep)
            self.set_control_data(ControlData.from_buffer_copy(st["ctrl"]), rep)
            self.set_exchange_parameter_value(st["r"], rep)
            self._ck(self.lib.dqmc_rng_seed(self.h, rep, self.pars["seed"], int(rng_indices[rep])))
            left = int(st["consumed"])
            while left > 0:                               # advance the fresh stream to where the saved run stood
                n = min(left, 1 << 20)
                self.rng_skip(n, rep)
                left -= n
        self.setup_storage()                              # loadContents: setupUdVStorage_and_calculateGreen, lastSweepDir = Up
        self._ck(self.lib.dqmc_set_performed_sweeps(self.h, int(state["performedSweeps"])))

    # ------------------------------------------------------------------ operators
    def bmat_mult(self, op, A, k2, k1, rep=0):
        a = np.array(A, dtype=np.complex128, order="F", copy=True)
        self._ck(self.lib.dqmc_bmat_mult(self.h, rep, 0, op, _ptr(a), k2, k1))
        return a

    def setup_storage(self):
        """setupUdVStorage_and_calculateGreen (detmodel.h:678-713)."""
        self._ck(self.lib.dqmc_setup_storage(self.h))

    def wrap_up(self, k):
        self._ck(self.lib.dqmc_wrap_up(self.h, k))

    def wrap_down(self, k):
        self._ck(self.lib.dqmc_wrap_down(self.h, k))

    def advance_up(self, l):
        self._ck(self.lib.dqmc_advance_up(self.h, l))

    def advance_down(self, l):
        self._ck(self.lib.dqmc_advance_down(self.h, l))

    def green_consistency(self):
        out = np.zeros(self.R)
        self._ck(self.lib.dqmc_get_green_consistency(self.h, _ptr(out)))
        return out

    def green_for_timeslice(self, k, rep=0):
        out = np.zeros((self.D, self.D), dtype=np.complex128, order="F")
        self._ck(self.lib.dqmc_green_for_timeslice(self.h, rep, 0, k, _ptr(out)))
        return out

    def green_from_udt(self, Qr, dr, Tr, Ql, dl, Tl):
        f = lambda a: np.asfortranarray(a, dtype=np.complex128)
        Qr, Tr, Ql, Tl = f(Qr), f(Tr), f(Ql), f(Tl)
        dr = np.ascontiguousarray(dr, dtype=np.float64)
        dl = np.ascontiguousarray(dl, dtype=np.float64)
        G = np.zeros((self.D, self.D), dtype=np.complex128, order="F")
        ld = c_f64()
        self._ck(self.lib.dqmc_green_from_udt_host(self.h, _ptr(Qr), _ptr(dr), _ptr(Tr), _ptr(Ql), _ptr(dl),
                                                   _ptr(Tl), _ptr(G), ctypes.byref(ld)))
        return G, ld.value

    def udt_decompose(self, M):
        M = np.asfortranarray(M, dtype=np.complex128)
        Q = np.zeros_like(M, order="F")
        T = np.zeros_like(M, order="F")
        d = np.zeros(self.D)
        self._ck(self.lib.dqmc_udt_decompose_host(self.h, _ptr(M), _ptr(Q), _ptr(d), _ptr(T)))
        return Q, d, T

    def gemm(self, A, B, transa=False, transb=False):
        A = np.asfortranarray(A, dtype=np.complex128)
        B = np.asfortranarray(B, dtype=np.complex128)
        M = A.shape[1] if transa else A.shape[0]
        K = A.shape[0] if transa else A.shape[1]
        N = B.shape[0] if transb else B.shape[1]
        C = np.zeros((M, N), dtype=np.complex128, order="F")
        self._ck(self.lib.dqmc_gemm_host(self.h, int(transa), int(transb), M, N, K, _ptr(A), _ptr(B), _ptr(C)))
        return C

    # ------------------------------------------------------------------ Monte Carlo
    def update_in_slice(self, k, thermalization=False):
        """updateInSlice(k) for every replica; returns the accepted counts."""
        acc = np.zeros(self.R, dtype=np.uint32)
        self._ck(self.lib.dqmc_update_slice(self.h, k, int(thermalization), _ptr(acc)))
        return acc

    def global_shift_move(self):
        acc = np.zeros(self.R, dtype=np.int32)
        self._ck(self.lib.dqmc_global_shift_move(self.h, _ptr(acc)))
        return acc

    def wolff_cluster_move(self, with_shift=False):
        """attemptWolffClusterUpdate / attemptWolffClusterShiftUpdate (detsdwopdim.cpp:3487-3562, 3647-3748) for every
        replica; returns the per-replica accept flags."""
        acc = np.zeros(self.R, dtype=np.int32)
        self._ck(self.lib.dqmc_wolff_cluster_move(self.h, int(bool(with_shift)), _ptr(acc)))
        return acc

    def wolff_statistics(self, rep=0):
        """(attempted, accepted, attemptedShift, acceptedShift, addedWolffClusterSize), detsdwopdim.h:285-299."""
        out = np.zeros(5)
        self._ck(self.lib.dqmc_get_wolff_statistics(self.h, int(rep), _ptr(out)))
        return out

    def bosonic_observables(self, rep=0):
        """normMeanPhi, associatedEnergy, phiRhoS_Gs, phiRhoS_Gc of the current fields (detsdwopdim.cpp:441-560, 903-918)."""
        from .pt import bosonic_observables
        return bosonic_observables(self.phi(rep), self.pars["dtau"])

    def phi_action(self):
        out = np.zeros(self.R)
        self._ck(self.lib.dqmc_phi_action(self.h, _ptr(out)))
        return out

    def sweep(self, takeMeasurements=False):
        """sweep(takeMeasurements), detsdwopdim.cpp:4422-4460.  With takeMeasurements the fermionic observables are
        accumulated on the device after every slice; read them with fermionic_observables()."""
        self._ck(self.lib.dqmc_sweep(self.h, 2 if takeMeasurements else 0))

    def fermionic_observables(self, rep=0):
        """finishMeasurements (detsdwopdim.cpp:903-1000) of the last sweep(True): greenK0, greenLocal, occDiffSq,
        pairPlusMax, pairMinusMax and the vectors kOccX, kOccY, pairPlus, pairMinus."""
        sc, vec = np.zeros(5), np.zeros(4 * self.N)
        self._ck(self.lib.dqmc_get_fermionic_observables(self.h, int(rep), _ptr(sc), _ptr(vec)))
        N = self.N
        return dict(greenK0=sc[0], greenLocal=sc[1], occDiffSq=sc[2], pairPlusMax=sc[3], pairMinusMax=sc[4],
                    kOccX=vec[:N].copy(), kOccY=vec[N:2 * N].copy(), pairPlus=vec[2 * N:3 * N].copy(),
                    pairMinus=vec[3 * N:].copy())

    def sweepThermalization(self):
        self._ck(self.lib.dqmc_sweep(self.h, 1))

    def sweepSimple(self, takeMeasurements=False):
        """greenUpdate = simple (detsdwopdim.cpp:4366-4393): G from scratch at every slice, then the slice update."""
        if takeMeasurements:
            raise DqmcError("fermionic measurements are served by the stabilized sweep (sweep(True)) only")
        self._ck(self.lib.dqmc_sweep_simple(self.h, 0))

    def sweepSimpleThermalization(self):
        self._ck(self.lib.dqmc_sweep_simple(self.h, 1))

    # ------------------------------------------------------------------ replica exchange interface
    def get_exchange_parameter_value(self, rep=0):
        out = c_f64()
        self._ck(self.lib.dqmc_get_exchange_parameter(self.h, rep, ctypes.byref(out)))
        return out.value

    def set_exchange_parameter_value(self, r, rep=0):
        self._ck(self.lib.dqmc_set_exchange_parameter(self.h, rep, float(r)))

    def exchange_pack(self, payload_dev_ptr, n_uniforms):
        self._ck(self.lib.dqmc_exchange_pack(self.h, c_vp(int(payload_dev_ptr)), int(n_uniforms)))

    def exchange_apply(self, r_new, ctrl_new, n_used):
        r_new = np.ascontiguousarray(r_new, dtype=np.float64)
        ctrl_new = np.ascontiguousarray(ctrl_new, dtype=np.float64)
        assert r_new.shape == (self.R,) and ctrl_new.size == self.R * (ctypes.sizeof(ControlData) // 8)
        self._ck(self.lib.dqmc_exchange_apply(self.h, _ptr(r_new), _ptr(ctrl_new), int(n_used)))

    def rng_preload(self, n_sweeps):
        self._ck(self.lib.dqmc_rng_preload(self.h, int(n_sweeps)))

    def rng_release(self):
        self._ck(self.lib.dqmc_rng_release(self.h))

    def profile_enable(self, on=True):
        self._ck(self.lib.dqmc_profile_enable(self.h, int(on)))

    def profile_get(self):
        ncat = 10                                     # DQMC_PROF_NCAT
        ms = np.zeros(ncat)
        cnt = np.zeros(ncat, dtype=np.uint64)
        self._ck(self.lib.dqmc_profile_get(self.h, _ptr(ms), _ptr(cnt)))
        names = [self.lib.dqmc_profile_name(i).decode() for i in range(ncat)]
        return {n: (float(m), int(c)) for n, m, c in zip(names, ms, cnt)}

    def accepted_total(self):
        out = np.zeros(self.R, dtype=np.uint64)
        self._ck(self.lib.dqmc_accepted_total(self.h, _ptr(out)))
        return out

    def get_exchange_action_contribution(self, device_ptr=None):
        out = np.zeros(self.R)
        self._ck(self.lib.dqmc_exchange_actions(self.h, c_vp(device_ptr) if device_ptr else None, _ptr(out)))
        return out


def exchange_walk(control_values, par_process, process_par, actions, uniforms):
    """Serial ladder walk (detqmcpt.h:1031-1079) through the C ABI; returns (n_used, swapped)."""
    lib = load_library()
    cv = np.ascontiguousarray(control_values, dtype=np.float64)
    ac = np.ascontiguousarray(actions, dtype=np.float64)
    un = np.ascontiguousarray(uniforms, dtype=np.float64)
    n = len(cv)
    used = c_i32()
    swapped = np.zeros(max(n - 1, 1), dtype=np.int32)
    rc = lib.dqmc_exchange_walk(n, _ptr(cv), _ptr(par_process), _ptr(process_par), _ptr(ac), _ptr(un),
                                ctypes.byref(used), _ptr(swapped))
    if rc != 0:
        raise DqmcError("dqmc_exchange_walk failed")
    return used.value, swapped[:n - 1]


CTRL_WORDS = ctypes.sizeof(ControlData) // 8      # 108 doubles per control-data blob


class ReplicaExchangeLadder:
    """DetQMCPT::replicaExchangeStep (detqmcpt.h:962-1118) for a ladder of P control-parameter
    values partitioned contiguously over `world` ranks with `n_local` replicas each.

    Per exchange step every rank contributes one payload vector
        [ n_local actions | P-1 look-ahead uniforms of its local replica 0 | n_local control blobs ]
    (dqmc_exchange_pack writes it on the device); the payloads of all ranks are all-gathered (NCCL
    in production, gloo / identity in tests) and every rank then performs the identical serial ladder
    walk with rank 0's uniforms -- the reference walks on MPI rank 0 with rank 0's RngWrapper
    (detqmcpt.h:1031-1079) and scatters the result; here nothing needs to be scattered."""

    def __init__(self, control_values, n_local, rank=0, world=1):
        self.values = np.ascontiguousarray(control_values, dtype=np.float64)
        self.P = len(self.values)
        assert self.P == n_local * world, "ladder length must equal world * n_local"
        self.n_local, self.rank, self.world = n_local, rank, world
        self.par_process = np.arange(self.P, dtype=np.int32)     # current_par_process
        self.process_par = np.arange(self.P, dtype=np.int32)     # current_process_par
        self.proposed = np.zeros(max(self.P - 1, 1), dtype=np.int64)
        self.accepted = np.zeros(max(self.P - 1, 1), dtype=np.int64)
        # ExchangeStatistics (detqmcpt.h:996-1010): replicas are labelled by the end of the ladder they visited last
        self.going = np.zeros(self.P, dtype=np.int8)             # 0 none, +1 up (from index 0), -1 down (from P-1)
        self.count_up = np.zeros(self.P, dtype=np.int64)
        self.count_down = np.zeros(self.P, dtype=np.int64)

    @property
    def n_uniforms(self):
        return self.P - 1

    @property
    def payload_len(self):
        return self.n_local + self.n_uniforms + self.n_local * CTRL_WORDS

    def local_parameters(self):
        lo = self.rank * self.n_local
        return self.values[self.process_par[lo:lo + self.n_local]]

    def walk(self, gathered):
        """gathered: array [world, payload_len].  Returns (r_new[n_local], ctrl_new[n_local, CTRL_WORDS],
        n_uniforms_used_by_this_rank)."""
        g = np.asarray(gathered, dtype=np.float64).reshape(self.world, self.payload_len)
        nl, nu = self.n_local, self.n_uniforms
        actions = np.ascontiguousarray(g[:, :nl].reshape(-1))
        uniforms = np.ascontiguousarray(g[0, nl:nl + nu])
        blobs = g[:, nl + nu:].reshape(self.P, CTRL_WORDS)
        for pi in range(self.P):                                   # diffusion statistics, before the walk
            npar = self.process_par[pi]
            if npar == self.P - 1:
                self.going[pi] = -1
            elif npar == 0:
                self.going[pi] = +1
            if self.going[pi] < 0:
                self.count_down[npar] += 1
            elif self.going[pi] > 0:
                self.count_up[npar] += 1
        old_par_process = self.par_process.copy()
        used, swapped = exchange_walk(self.values, self.par_process, self.process_par, actions, uniforms)
        self.proposed[:self.P - 1] += 1
        self.accepted[:self.P - 1] += swapped
        lo = self.rank * nl
        new_par = self.process_par[lo:lo + nl]
        r_new = self.values[new_par]
        # control data follow the parameter (std::swap of the buffers, detqmcpt.h:1052-1053): replica pi
        # gets the blob of whoever held its new parameter before the walk
        ctrl_new = np.ascontiguousarray(blobs[old_par_process[new_par]])
        return np.ascontiguousarray(r_new), ctrl_new, (used if self.rank == 0 else 0)
