"""TEST INFRASTRUCTURE ONLY -- ctypes bindings to oracle/_ref/libdetqmc_ref.so (the unmodified
reference compiled by oracle/Makefile).  Never imported by the product path."""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdetqmc_ref.so")

c_i32, c_u32, c_f64, c_vp = ctypes.c_int32, ctypes.c_uint32, ctypes.c_double, ctypes.c_void_p


class RefSdwParams(ctypes.Structure):
    _fields_ = [("opdim", c_i32), ("L", c_i32), ("m", c_i32), ("s", c_i32),
                ("dtau", c_f64), ("r", c_f64), ("c", c_f64), ("u", c_f64), ("lam", c_f64),
                ("txhor", c_f64), ("txver", c_f64), ("tyhor", c_f64), ("tyver", c_f64),
                ("cdwU", c_f64), ("mu", c_f64), ("accRatio", c_f64),
                ("weakZflux", c_i32), ("bc", c_i32), ("updateMethod", c_i32), ("delaySteps", c_i32),
                ("globalShift", c_i32), ("globalUpdateInterval", c_i32), ("repeatUpdateInSlice", c_i32),
                ("seed", c_u32), ("rngIndex", c_u32),
                ("wolffClusterUpdate", c_i32), ("wolffClusterShiftUpdate", c_i32), ("repeatWolffPerSweep", c_i32),
                ("fermionMeasurements", c_i32), ("denseHopping", c_i32)]


class RefHubParams(ctypes.Structure):
    _fields_ = [("L", c_i32), ("m", c_i32), ("s", c_i32), ("checkerboard", c_i32),
                ("dtau", c_f64), ("t", c_f64), ("U", c_f64), ("mu", c_f64),
                ("seed", c_u32), ("rngIndex", c_u32)]


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        l = ctypes.CDLL(LIB_PATH)
        for name in ("ref_rng_create", "ref_sdw_create", "ref_hub_create"):
            getattr(l, name).restype = c_vp
        l.ref_sdw_update_in_slice.restype = c_f64
        l.ref_sdw_exchange_probability.restype = c_f64
        l.ref_sdw_exchange_probability.argtypes = [c_f64] * 4
        l.ref_sdw_set_r.argtypes = [c_vp, c_f64]
        l.ref_sdw_set_phi_delta.argtypes = [c_vp, c_f64]
        if hasattr(l, "ref_sdw_save_config_stream"):
            l.ref_sdw_save_config_stream.argtypes = [c_vp, ctypes.c_char_p, ctypes.c_int]
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(c_vp)


class RefRng:
    def __init__(self, seed, index):
        self.h = c_vp(lib().ref_rng_create(c_u32(seed), c_u32(index)))

    def draw(self, n):
        out = np.zeros(n)
        lib().ref_rng_draw(self.h, c_i32(n), _p(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_rng_destroy(self.h)
            self.h = None


def sdw_params_from(p):
    """Build the ctypes struct from an oracle SdwParams-like object."""
    return RefSdwParams(p.opdim, p.L, p.m, p.s, p.dtau, p.r, p.c, p.u, p.lam, p.txhor, p.txver,
                        p.tyhor, p.tyver, p.cdwU, p.mu, p.accRatio, int(p.weakZflux), p.bc,
                        p.updateMethod, p.delaySteps, int(p.globalShift), p.globalUpdateInterval,
                        p.repeatUpdateInSlice, p.seed, p.rngIndex,
                        int(getattr(p, "wolffClusterUpdate", False)), int(getattr(p, "wolffClusterShiftUpdate", False)),
                        int(getattr(p, "repeatWolffPerSweep", 1)), int(getattr(p, "fermionMeasurements", False)),
                        0 if getattr(p, "checkerboard", True) else 1)


class RefSdw:
    """The reference's DetSDW<CB_ASSAAD_BERG, OPDIM> driven through oracle/ref_harness.cpp."""

    def __init__(self, pars):
        self._cp = sdw_params_from(pars)
        h = lib().ref_sdw_create(ctypes.byref(self._cp))
        if not h:
            raise RuntimeError("reference refused the parameters")
        self.h = c_vp(h)
        d = (c_i32 * 6)()
        lib().ref_sdw_dims(self.h, d)
        self.N, self.D, self.m, self.n, self.s, self.opdim = list(d)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_sdw_destroy(self.h)
            self.h = None

    def phi(self):
        out = np.zeros((self.m + 1, self.opdim, self.N))
        lib().ref_sdw_get_phi(self.h, _p(out))
        return out

    def set_phi(self, phi):
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        assert phi.shape == (self.m + 1, self.opdim, self.N)
        lib().ref_sdw_set_phi(self.h, _p(phi))

    def green(self):
        out = np.zeros((self.D, self.D), dtype=np.complex128, order="F")
        lib().ref_sdw_get_green(self.h, _p(out))
        return out

    def sv(self):
        out = np.zeros(self.D)
        lib().ref_sdw_get_sv(self.h, _p(out))
        return out

    def tables(self):
        c = np.zeros((self.m + 1, self.N))
        s = np.zeros((self.m + 1, self.N))
        lib().ref_sdw_get_tables(self.h, _p(c), _p(s))
        return c, s

    def bmult(self, op, A, k2, k1):
        """op: 0 B*A, 1 A*B, 2 B^-1*A, 3 A*B^-1 with B = B(k2,k1)."""
        a = np.array(A, dtype=np.complex128, order="F", copy=True)
        lib().ref_sdw_bmult(self.h, c_i32(op), _p(a), c_u32(k2), c_u32(k1))
        return a

    def dense_bmat(self, k2, k1):
        """computeBmatSDW(k2, k1) (detsdwopdim.cpp:1307-1497): dense B with the full hopping exponential."""
        out = np.zeros((self.D, self.D), dtype=np.complex128, order="F")
        lib().ref_sdw_dense_bmat(self.h, c_u32(k2), c_u32(k1), _p(out))
        return out

    def update_in_slice(self, k, therm=False):
        return lib().ref_sdw_update_in_slice(self.h, c_u32(k), c_i32(int(therm)))

    def sweep(self, therm=False):
        rc = lib().ref_sdw_sweep(self.h, c_i32(int(therm)))
        if rc:
            raise RuntimeError("reference sweep failed")

    def scalars(self):
        out = np.zeros(16)
        lib().ref_sdw_get_scalars(self.h, _p(out))
        keys = ["phiDelta", "performedSweeps", "currentTimeslice", "lastSweepDir", "acceptedGlobalShifts",
                "attemptedGlobalShifts", "lastAccRatio", "phiAction", "exchangeAction", "r",
                "raSamples", "raAverage"]
        return dict(zip(keys, out[:len(keys)]))

    def set_r(self, r):
        lib().ref_sdw_set_r(self.h, c_f64(r))

    def set_phi_delta(self, d):
        lib().ref_sdw_set_phi_delta(self.h, c_f64(d))

    def rng_draw(self, n):
        out = np.zeros(n)
        lib().ref_sdw_rng_draw(self.h, c_i32(n), _p(out))
        return out

    def green_for_timeslice(self, k):
        out = np.zeros((self.D, self.D), dtype=np.complex128, order="F")
        lib().ref_sdw_green_for_timeslice(self.h, c_u32(k), _p(out))
        return out

    def udv(self, l):
        U = np.zeros((self.D, self.D), dtype=np.complex128, order="F")
        V = np.zeros((self.D, self.D), dtype=np.complex128, order="F")
        d = np.zeros(self.D)
        lib().ref_sdw_get_udv(self.h, c_u32(l), _p(U), _p(d), _p(V))
        return U, d, V

    def green_from_storage(self, l_left, l_right):
        out = np.zeros((self.D, self.D), dtype=np.complex128, order="F")
        sv = np.zeros(self.D)
        lib().ref_sdw_green_from_storage(self.h, c_u32(l_left), c_u32(l_right), _p(out), _p(sv))
        return out, sv


    def measured_sweep_fermionic(self):
        """sweep(true) with fermionic measurements on (construct with fermionMeasurements=True): dict of the scalars
        greenK0, greenLocal, occDiffSq, pairPlusMax, pairMinusMax and the vectors kOccX, kOccY, pairPlus, pairMinus."""
        sc, vec = np.zeros(5), np.zeros(4 * self.N)
        if lib().ref_sdw_measured_sweep_fermionic(self.h, _p(sc), _p(vec)):
            raise RuntimeError("reference sweep failed")
        N = self.N
        return dict(greenK0=sc[0], greenLocal=sc[1], occDiffSq=sc[2], pairPlusMax=sc[3], pairMinusMax=sc[4],
                    kOccX=vec[:N].copy(), kOccY=vec[N:2 * N].copy(), pairPlus=vec[2 * N:3 * N].copy(),
                    pairMinus=vec[3 * N:].copy())

    def shift_green_symmetric(self):
        """shiftGreenSymmetric() of the current g (detsdwopdim.cpp:4505-4612)."""
        out = np.zeros((self.D, self.D), dtype=np.complex128, order="F")
        lib().ref_sdw_shift_green_symmetric(self.h, _p(out))
        return out

    def sweep_simple(self, therm=False):
        """sweepSimple(false) / sweepSimpleThermalization() (greenUpdate = simple, detsdwopdim.cpp:4366-4420)."""
        if lib().ref_sdw_sweep_simple(self.h, c_i32(int(therm))):
            raise RuntimeError("reference sweep failed")

    def measured_sweep(self):
        """sweep(takeMeasurements=True): returns the bosonic observables normMeanPhi, associatedEnergy, phiRhoS_Gs,
        phiRhoS_Gc (detsdwopdim.cpp:441-560, 903-918; the last two are zero unless opdim == 2)."""
        out = np.zeros(4)
        if lib().ref_sdw_measured_sweep(self.h, _p(out)) != 0:
            raise RuntimeError("reference sweep failed")
        return dict(normMeanPhi=out[0], associatedEnergy=out[1], phiRhoS_Gs=out[2], phiRhoS_Gc=out[3])

    def attempt_wolff(self, shift=False):
        """attemptWolffClusterUpdate / attemptWolffClusterShiftUpdate (detsdwopdim.cpp:3487-3562, 3647-3748); returns
        the update statistics (attempted, accepted, attemptedShift, acceptedShift, addedWolffClusterSize)."""
        st = np.zeros(5)
        lib().ref_sdw_attempt_wolff(self.h, 1 if shift else 0, _p(st))
        return st

    def save_config_stream(self, directory, binary=True):
        """Append the current configuration to configs-phi.{binary,text}stream in `directory` with the
        reference's own writers (detsdwopdim.cpp:4943-5036)."""
        lib().ref_sdw_save_config_stream(self.h, str(directory).encode(), 1 if binary else 0)


def exchange_probability(par1, a1, par2, a2):
    return lib().ref_sdw_exchange_probability(par1, a1, par2, a2)


class RefHubbard:
    def __init__(self, pars):
        self._cp = RefHubParams(pars.L, pars.m, pars.s, int(pars.checkerboard), pars.dtau, pars.t,
                                pars.U, pars.mu, pars.seed, pars.rngIndex)
        h = lib().ref_hub_create(ctypes.byref(self._cp))
        if not h:
            raise RuntimeError("reference refused the parameters")
        self.h = c_vp(h)
        d = (c_i32 * 4)()
        lib().ref_hub_dims(self.h, d)
        self.N, self.m, self.n, self.s = list(d)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_hub_destroy(self.h)
            self.h = None

    def aux(self):
        out = np.zeros((self.m + 1, self.N), dtype=np.int32)
        lib().ref_hub_get_aux(self.h, _p(out))
        return out

    def green(self, gc):
        out = np.zeros((self.N, self.N), order="F")
        lib().ref_hub_get_green(self.h, c_i32(gc), _p(out))
        return out

    def sv(self, gc):
        out = np.zeros(self.N)
        lib().ref_hub_get_sv(self.h, c_i32(gc), _p(out))
        return out

    def proptmat(self):
        out = np.zeros((self.N, self.N), order="F")
        lib().ref_hub_get_proptmat(self.h, _p(out))
        return out

    def bmat(self, gc, k2, k1):
        out = np.zeros((self.N, self.N), order="F")
        lib().ref_hub_bmat(self.h, c_i32(gc), c_u32(k2), c_u32(k1), _p(out))
        return out

    def sweep(self, therm=False):
        rc = lib().ref_hub_sweep(self.h, c_i32(int(therm)))
        if rc:
            raise RuntimeError("reference sweep failed")

    def measured_sweep(self):
        """sweep(true): the reference's scalar observables (dethubbard.cpp:88-95 order of finishMeasurements) and zcorr"""
        sc, zc = np.zeros(8), np.zeros(self.N)
        if lib().ref_hub_measured_sweep(self.h, _p(sc), _p(zc)):
            raise RuntimeError("reference sweep failed")
        return sc, zc

    def rng_draw(self, n):
        out = np.zeros(n)
        lib().ref_hub_rng_draw(self.h, c_i32(n), _p(out))
        return out
