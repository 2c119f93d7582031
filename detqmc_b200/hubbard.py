"""Host-side mirror of DetHubbard (dethubbard.h:60-370) for a batch of replicas on one GPU.

`DetHubbardBatch` follows the reference's member names (sweep, sweepThermalization, ...); all numerics
happen in libdqmc_b200.so through the C ABI (model = DQMC_MODEL_HUBBARD).  Green's functions are real
N x N matrices, component gc = 0 is spin up, gc = 1 spin down (dethubbard.h:281-337)."""
import ctypes

import numpy as np

from .lib import load_library, DqmcParams, DqmcError, c_i32, c_vp, c_f64

_DEFAULTS = dict(L=4, m=40, s=10, dtau=0.1, t=1.0, U=4.0, mu=0.0, checkerboard=False, seed=1020304050, rngIndex=1)


def _ptr(a):
    return a.ctypes.data_as(c_vp)


class DetHubbardBatch:
    def __init__(self, pars=None, n_replicas=1, device=0, rng_indices=None, init="random", **kw):
        self.lib = load_library()
        d = dict(_DEFAULTS)
        if pars is not None:
            src = pars if isinstance(pars, dict) else vars(pars)
            d.update({k: v for k, v in src.items() if k in d})
        d.update(kw)
        self.pars = d
        p = DqmcParams()
        p.model = 1
        p.opdim, p.L, p.m, p.s = 1, d["L"], d["m"], d["s"]
        p.delaySteps = 1
        p.checkerboard = int(bool(d["checkerboard"]))
        p.dtau, p.t, p.U, p.mu = d["dtau"], d["t"], d["U"], d["mu"]
        self.cpars = p
        self.R = int(n_replicas)
        h = c_vp()
        rc = self.lib.dqmc_create(ctypes.byref(p), self.R, int(device), ctypes.byref(h))
        self.h = h
        if rc != 0:
            msg = self.lib.dqmc_last_error(h).decode() if h else "dqmc_create failed"
            if h:
                self.lib.dqmc_destroy(h)
                self.h = None
            raise DqmcError(msg)
        dims = (c_i32 * 8)()
        self._ck(self.lib.dqmc_dims(self.h, dims))
        self.N, self.D, self.m, self.n, self.s, self.ngc, _, _ = list(dims)
        if rng_indices is None:
            rng_indices = [d["rngIndex"] + i for i in range(self.R)]
        for rep, idx in enumerate(rng_indices):
            self._ck(self.lib.dqmc_rng_seed(self.h, rep, d["seed"], int(idx)))
        if init == "random":
            for rep in range(self.R):
                self._ck(self.lib.dqmc_init_random_fields(self.h, rep))      # setupRandomAuxfield
            self._ck(self.lib.dqmc_setup_storage(self.h))                    # setupUdVStorage_and_calculateGreen

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

    @property
    def launch_count(self):
        return int(self.lib.dqmc_launch_count(self.h))

    def auxfield(self, rep=0):
        out = np.zeros((self.m + 1, self.N), dtype=np.int32)
        self._ck(self.lib.dqmc_download_fields(self.h, rep, _ptr(out)))
        return out

    def set_auxfield(self, aux, rep=0):
        aux = np.ascontiguousarray(aux, dtype=np.int32)
        assert aux.shape == (self.m + 1, self.N)
        self._ck(self.lib.dqmc_upload_fields(self.h, rep, _ptr(aux)))

    def green(self, rep=0, gc=0):
        out = np.zeros((self.N, self.N), order="F")
        self._ck(self.lib.dqmc_download_green(self.h, rep, gc, _ptr(out)))
        return out

    def logdet(self, rep=0, gc=0):
        out = c_f64()
        self._ck(self.lib.dqmc_logdet(self.h, rep, gc, ctypes.byref(out)))
        return out.value

    def bmat_mult(self, op, A, k2, k1, rep=0, gc=0):
        a = np.array(A, dtype=np.float64, order="F", copy=True)
        self._ck(self.lib.dqmc_bmat_mult(self.h, rep, gc, op, _ptr(a), k2, k1))
        return a

    def setup_storage(self):
        self._ck(self.lib.dqmc_setup_storage(self.h))

    def green_for_timeslice(self, k, rep=0, gc=0):
        out = np.zeros((self.N, self.N), order="F")
        self._ck(self.lib.dqmc_green_for_timeslice(self.h, rep, gc, k, _ptr(out)))
        return out

    def green_consistency(self):
        out = np.zeros(self.R * self.ngc)
        self._ck(self.lib.dqmc_get_green_consistency(self.h, _ptr(out)))
        return out

    def update_in_slice(self, k):
        acc = np.zeros(self.R, dtype=np.uint32)
        self._ck(self.lib.dqmc_update_slice(self.h, k, 0, _ptr(acc)))
        return acc

    def sweep(self, takeMeasurements=False):
        """sweep(takeMeasurements) (dethubbard.cpp:173-183): with measurements, DetHubbard::measure runs on the device
        after the update of every slice; read the result with observables()."""
        self._ck(self.lib.dqmc_sweep(self.h, 2 if takeMeasurements else 0))

    def sweepThermalization(self):
        self._ck(self.lib.dqmc_sweep(self.h, 0))

    OBSERVABLES = ("occupationUp", "occupationDown", "totalOccupation", "doubleOccupation", "localMoment",
                   "kineticEnergy", "potentialEnergy", "totalEnergy")

    def observables(self, rep=0):
        """finishMeasurements (dethubbard.cpp:601-612) after sweep(True): the reference's scalar observables by name
        plus spinzCorrelationFunction [N]."""
        sc = np.zeros(8)
        zc = np.zeros(self.N)
        self._ck(self.lib.dqmc_get_hubbard_observables(self.h, rep, _ptr(sc), _ptr(zc)))
        out = dict(zip(self.OBSERVABLES, sc.tolist()))
        out["spinzCorrelationFunction"] = zc
        return out

    def synchronize(self):
        self._ck(self.lib.dqmc_synchronize(self.h))

    def rng_draw(self, n, rep=0):
        out = np.zeros(n)
        self._ck(self.lib.dqmc_rng_draw(self.h, rep, n, _ptr(out)))
        return out

    def total_occupation(self, rep=0):
        """<n_up + n_down> per site from the equal-time Green's functions (= 1 at half filling)."""
        return 2.0 - (np.trace(self.green(rep, 0)) + np.trace(self.green(rep, 1))) / self.N
