"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

NumPy restatement of the reference's DQMC sweep hot path (crstnbr/detqmc).  Every routine cites
the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module, and only as the checker.

Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement against the
reference itself compiled into oracle/_ref/libdetqmc_ref.so (RNG stream, field construction,
checkerboard multiplies, G and singular values after setup, per-slice acceptance, fields and step
sizes after sweeps, global-shift decisions, Hubbard path), and against the committed fixtures in
tests/golden/ that were generated from that library (tools/make_golden.py).  The replica-exchange
ladder walk (detqmcpt.h:1014-1080) cannot be run from the reference here (no MPI): that one
function is a restatement with parity UNPINNED beyond the exchange probability itself
(detsdwopdim.cpp:5251-5264), which is pinned.

Layouts follow Armadillo (column-major):
  phi[k, dim, site]   == arma::Cube(N, OPDIM, m+1) memory   (detsdwopdim.h:455-461)
  G[row, col]         row/col = site + N*bandspin, bandspin XUP=0,YDOWN=1,XDOWN=2,YUP=3
"""

import numpy as np
import scipy.linalg as sla

from dsfmt_oracle import RngOracle

XBAND, YBAND = 0, 1
BAND_OF_BANDSPIN = (XBAND, YBAND, XBAND, YBAND)   # XUP, YDOWN, XDOWN, YUP  (detsdwopdim.h:265-273)
UP, DOWN = +1, -1                                   # SweepDirection (detmodel.h:480)


def udv_decompose(mat):
    """udv.h:68-90: M = U diag(d) V_t^dagger via LAPACK gesvd ('std')."""
    U, d, Vh = sla.svd(mat, lapack_driver="gesvd")
    return U, d, Vh.conj().T


class UdV:
    __slots__ = ("U", "d", "V_t")

    def __init__(self, U, d, V_t):
        self.U, self.d, self.V_t = U, d, V_t

    @staticmethod
    def eye(sz, dtype):
        return UdV(np.eye(sz, dtype=dtype), np.ones(sz), np.eye(sz, dtype=dtype))

    @staticmethod
    def of(mat):
        return UdV(*udv_decompose(mat))

    def copy(self):
        return UdV(self.U.copy(), self.d.copy(), self.V_t.copy())


def green_from_udv(l, r):
    """detmodel.h:768-818.  G = [1 + (U_r d_r V_r^+)(U_l d_l V_l^+)]^-1; returns (G, sv of G^-1)."""
    VU = r.V_t.conj().T @ l.U
    UtVt = r.U.conj().T @ l.V_t
    Ut, sv, Vt = udv_decompose(UtVt + (r.d[:, None] * VU) * l.d[None, :])
    Vp = l.V_t @ Vt
    Up = r.U @ Ut
    return (Vp * (1.0 / sv)[None, :]) @ Up.conj().T, sv


def green_from_eye_and_udv(r):
    """detmodel.h:822-860."""
    Ut, sv, Vt = udv_decompose(r.U.conj().T @ r.V_t + np.diag(r.d))
    Vp = r.V_t @ Vt
    Up = r.U @ Ut
    return (Vp * (1.0 / sv)[None, :]) @ Up.conj().T, sv


class RunningAverage:
    """RunningAverage.h:20-80."""

    def __init__(self, sample_size):
        self.sample_size = sample_size
        self.samples_added = 0
        self.values = []
        self.avg = 0.0

    def add(self, v):
        if self.samples_added < self.sample_size:
            self.values.append(v)
            self.avg += v / self.sample_size
        else:
            self.avg -= self.values.pop(0) / self.sample_size
            self.values.append(v)
            self.avg += v / self.sample_size
        self.samples_added += 1


class SweepSkeleton:
    """Restatement of DetModelGC's stabilised sweep (detmodel.h:678-713, 953-1163, 1261-1440).

    Sub-classes provide left/right B-multiplies (and inverses) with the skeleton signature
    (gc, A, k2, k1), update_in_slice(k) and optionally global_move()."""

    def init_skeleton(self, sz, m, s, n_gc, dtype):
        self.sz, self.m, self.s = sz, m, s
        self.n = int(np.ceil(m / s))
        self.n_gc = n_gc
        self.dtype = dtype
        self.green = [np.zeros((sz, sz), dtype=dtype) for _ in range(n_gc)]
        self.green_inv_sv = [np.zeros(sz) for _ in range(n_gc)]
        self.storage = [[None] * (self.n + 1) for _ in range(n_gc)]
        self.current_timeslice = 0
        self.last_sweep_dir = UP
        self.wrap_vs_advance = []      # (dir, slice, max|G_wrapped - G_advanced|) log

    # detmodel.h:678-713
    def setup_udv_storage_and_calculate_green(self):
        n, s, m = self.n, self.s, self.m
        eye = np.eye(self.sz, dtype=self.dtype)
        for gc in range(self.n_gc):
            st = [None] * (n + 1)
            st[0] = UdV.eye(self.sz, self.dtype)
            st[1] = UdV.of(self.left_multiply_bmat(gc, eye, s, 0))
            for l in range(1, n):
                k_l = s * l
                k_lp1 = s * (l + 1) if l < n - 1 else m
                BU = self.left_multiply_bmat(gc, st[l].U, k_lp1, k_l)
                nu = UdV.of(BU * st[l].d[None, :])
                nu.V_t = st[l].V_t @ nu.V_t
                st[l + 1] = nu
            self.storage[gc] = st
        for gc in range(self.n_gc):
            self.green[gc], self.green_inv_sv[gc] = green_from_eye_and_udv(self.storage[gc][n])
        self.current_timeslice = m
        self.last_sweep_dir = UP

    # detmodel.h:605-674 (used by consistency checks only)
    def green_for_timeslice(self, timeslice, gc=0):
        n, s, m = self.n, self.s, self.m
        if timeslice == m:
            raise ValueError("use setup_udv_storage_and_calculate_green for timeslice == m")
        eye = np.eye(self.sz, dtype=self.dtype)
        lk = timeslice // s
        k_lkp1 = s * (lk + 1) if lk < n - 1 else m
        cur = UdV.of(self.left_multiply_bmat(gc, eye, k_lkp1, timeslice))
        for l in range(lk + 1, n):
            k_l = s * l
            k_lp1 = s * (l + 1) if l < n - 1 else m
            nu = UdV.of(self.left_multiply_bmat(gc, cur.U, k_lp1, k_l) * cur.d[None, :])
            nu.V_t = cur.V_t @ nu.V_t
            cur = nu
        target = lk - 1 if lk * s == timeslice else lk
        for l in range(0, target + 1):
            k_l = s * l
            k_lp1 = s * (l + 1) if l < lk else timeslice
            nu = UdV.of(self.left_multiply_bmat(gc, cur.U, k_lp1, k_l) * cur.d[None, :])
            nu.V_t = cur.V_t @ nu.V_t
            cur = nu
        return green_from_eye_and_udv(cur)

    # detmodel.h:718-758 (greenUpdate = simple)
    def sweep_simple_skeleton(self, update):
        """sweepSimple_skeleton / sweepSimpleThermalization_skeleton: for every slice the Green's function is the plain
        inverse of 1 + B(k, 0) B(m, k) (no stabilisation), then the slice is updated.
        The reference builds these B matrices with computeBmatSDW (detsdwopdim.cpp:1307-1497), i.e. with the DENSE
        hopping exponential even when checkerboard = true; so does this restatement (force_dense) -- pinned by
        tests/golden/sdw_dense_hopping.npz (the reference's own simple sweeps)."""
        eye = np.eye(self.sz, dtype=self.dtype)
        for k in range(1, self.m + 1):
            self.force_dense = True
            for gc in range(len(self.green)):
                b_k0 = self.left_multiply_bmat(gc, eye, k, 0)
                b_mk = self.left_multiply_bmat(gc, eye, self.m, k) if k < self.m else eye
                self.green[gc] = np.linalg.inv(eye + b_k0 @ b_mk)
            self.force_dense = False
            update(k)

    # detmodel.h:953-1017
    def advance_down_green(self, l, gc):
        n, s, m = self.n, self.s, self.m
        st = self.storage[gc]
        k_l = s * l if l < n else m
        k_lm1 = s * (l - 1)
        if l < n:
            L = UdV.of(st[l].d[:, None] * self.right_multiply_bmat(gc, st[l].V_t.conj().T, k_l, k_lm1))
            L.U = st[l].U @ L.U
        else:
            L = UdV.of(self.right_multiply_bmat(gc, np.eye(self.sz, dtype=self.dtype), k_l, k_lm1))
        g_wrapped = self.green[gc]
        if l - 1 > 0:
            self.green[gc], self.green_inv_sv[gc] = green_from_udv(L, st[l - 1])
        else:
            self.green[gc], self.green_inv_sv[gc] = green_from_eye_and_udv(L)
        self.wrap_vs_advance.append((DOWN, k_lm1, float(np.abs(g_wrapped - self.green[gc]).max())))
        st[l - 1] = L
        self.current_timeslice = s * (l - 1)

    # detmodel.h:1106-1163
    def advance_up_green(self, l, gc):
        n, s, m = self.n, self.s, self.m
        st = self.storage[gc]
        k_l = s * l
        k_lp1 = s * (l + 1) if l < n - 1 else m
        assert self.current_timeslice == k_lp1
        g_wrapped = self.green[gc]
        T = UdV.of(self.left_multiply_bmat(gc, st[l].U, k_lp1, k_l) * st[l].d[None, :])
        T.V_t = st[l].V_t @ T.V_t
        if k_lp1 != m:
            self.green[gc], self.green_inv_sv[gc] = green_from_udv(st[l + 1], T)
        else:
            self.green[gc], self.green_inv_sv[gc] = green_from_eye_and_udv(T)
        st[l + 1] = T
        self.wrap_vs_advance.append((UP, k_lp1, float(np.abs(g_wrapped - self.green[gc]).max())))
        self.current_timeslice = k_lp1

    # detmodel.h:1064-1095
    def wrap_down_green(self, k, gc):
        assert self.current_timeslice == k
        self.green[gc] = self.left_multiply_bmat_inv(
            gc, self.right_multiply_bmat(gc, self.green[gc], k, k - 1), k, k - 1)
        self.current_timeslice = k - 1

    # detmodel.h:1234-1259
    def wrap_up_green(self, k, gc):
        assert self.current_timeslice == k
        self.green[gc] = self.left_multiply_bmat(
            gc, self.right_multiply_bmat_inv(gc, self.green[gc], k + 1, k), k + 1, k)
        self.current_timeslice = k + 1

    # detmodel.h:1328-1399
    def sweep_down(self, update):
        n, s, m = self.n, self.s, self.m
        for k in range(m, (n - 1) * s, -1):
            update(k)
            for gc in range(self.n_gc):
                self.wrap_down_green(k, gc) if gc == self.n_gc - 1 else self._wrap_down_keep(k, gc)
        for l in range(n - 1, 0, -1):
            for gc in range(self.n_gc):
                self.advance_down_green(l + 1, gc)
            for k in range(l * s, (l - 1) * s, -1):
                update(k)
                for gc in range(self.n_gc):
                    self.wrap_down_green(k, gc) if gc == self.n_gc - 1 else self._wrap_down_keep(k, gc)
        for gc in range(self.n_gc):
            self.advance_down_green(1, gc)

    def _wrap_down_keep(self, k, gc):
        # multi-component models wrap every gc at the same slice: do not advance the slice
        # counter until the last component has been wrapped
        self.wrap_down_green(k, gc)
        self.current_timeslice = k

    def _wrap_up_keep(self, k, gc):
        self.wrap_up_green(k, gc)
        self.current_timeslice = k

    # detmodel.h:1261-1325
    def sweep_up(self, update):
        n, s, m = self.n, self.s, self.m
        for gc in range(self.n_gc):
            self.storage[gc][0] = UdV.eye(self.sz, self.dtype)
        for l in range(0, n - 1):
            for k in range(l * s + 1, (l + 1) * s + 1):
                for gc in range(self.n_gc):
                    self.wrap_up_green(k - 1, gc) if gc == self.n_gc - 1 else self._wrap_up_keep(k - 1, gc)
                update(k)
            for gc in range(self.n_gc):
                self.advance_up_green(l, gc)
        for k in range((n - 1) * s + 1, m + 1):
            for gc in range(self.n_gc):
                self.wrap_up_green(k - 1, gc) if gc == self.n_gc - 1 else self._wrap_up_keep(k - 1, gc)
            update(k)
        for gc in range(self.n_gc):
            self.advance_up_green(n - 1, gc)

    # detmodel.h:1401-1478
    def sweep_skeleton(self, update):
        if self.last_sweep_dir == UP:
            self.global_move()
            self.sweep_down(update)
            self.last_sweep_dir = DOWN
        else:
            self.sweep_up(update)
            self.last_sweep_dir = UP

    def global_move(self):
        pass


def bosonic_observables(phi, dtau):
    """The observables DetSDW measures with turnoffFermionMeasurements (initMeasurements / measure /
    finishMeasurements, detsdwopdim.cpp:441-560, 903-918), accumulated over the slices k = 1..m:
      normMeanPhi       |sum phi| / (N m)
      associatedEnergy  sum phi.phi / (2 N m)
      phiRhoS_Gc        (dtau / 2) sum_site [phi(site).phi(site + x) + phi(site).phi(site + y)]          (opdim == 2)
      phiRhoS_Gs        dtau sum_site [phi_0(site + x) phi_1(site) - phi_1(site + x) phi_0(site)]          (opdim == 2)
    phi: [m+1][opdim][N]."""
    m1, opdim, N = phi.shape
    L = int(round(np.sqrt(N)))
    ph = phi[1:]
    out = {"normMeanPhi": float(np.linalg.norm(ph.sum(axis=(0, 2)) / (N * (m1 - 1)))),
           "associatedEnergy": float(np.sum(ph * ph) / (2.0 * N * (m1 - 1)))}
    if opdim == 2:
        sites = np.arange(N)
        x, y = sites % L, sites // L
        xp = y * L + (x + 1) % L
        yp = ((y + 1) % L) * L + x
        out["phiRhoS_Gc"] = float(0.5 * dtau * (np.sum(ph * ph[:, :, xp]) + np.sum(ph * ph[:, :, yp])))
        out["phiRhoS_Gs"] = float(dtau * np.sum(ph[:, 0, xp] * ph[:, 1, :] - ph[:, 1, xp] * ph[:, 0, :]))
    return out


def config_stream(phi):
    """One configuration in the order of the reference's configuration streams
    (DetSDW::saveConfigurationStreamBinary, detsdwopdim.cpp:5000-5010; DetSDW_SystemConfig::write_to_disk_phi_*,
    detsdwsystemconfig.cpp:100-134): for ix, for iy (site i = iy*L + ix), for k = 1..m, for dim: phi(i, dim, k).
    `phi` is the [m+1][opdim][N] array of this module; returns a flat float64 array of N*m*opdim values."""
    m1, opdim, N = phi.shape
    L = int(round(np.sqrt(N)))
    out = np.empty(N * (m1 - 1) * opdim)
    pos = 0
    for ix in range(L):
        for iy in range(L):
            i = iy * L + ix
            for k in range(1, m1):
                for dim in range(opdim):
                    out[pos] = phi[k, dim, i]
                    pos += 1
    return out


def config_stream_text(phi):
    """The same configuration as the lines the text stream appends (precision 14, scientific:
    detsdwopdim.cpp:4952-4960)."""
    return "".join("%.14e\n" % v for v in config_stream(phi))


class SdwParams:
    """Subset of ModelParamsDetSDW (detsdwparams.h:30-120) that the hot path reads; defaults follow
    maindetqmcsdwopdim.cpp:93-135 where the reference has any."""

    def __init__(self, **kw):
        self.opdim = 2
        self.L = 4
        self.m = 20
        self.s = 10
        self.dtau = 0.1
        self.r = -1.0
        self.c = 3.0
        self.u = 1.0
        self.lam = 1.0
        self.txhor, self.txver, self.tyhor, self.tyver = -1.0, -0.5, 0.5, 1.0
        self.cdwU = 0.0
        self.mu = -0.5
        self.accRatio = 0.5
        self.weakZflux = True
        self.bc = 0                 # 0 pbc, 1 apbc-x, 2 apbc-y, 3 apbc-xy
        self.updateMethod = 2       # delayed
        self.delaySteps = 16
        self.globalShift = True
        self.globalUpdateInterval = 10
        self.wolffClusterUpdate = False
        self.wolffClusterShiftUpdate = False
        self.repeatWolffPerSweep = 1
        self.fermionMeasurements = False    # not turnoffFermionMeasurements
        self.repeatUpdateInSlice = 1
        self.checkerboard = True            # False: DetSDW<CB_NONE>, dense hopping exponential (reference goldens only)
        self.seed = 1020304050
        self.rngIndex = 1
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)
        # updateTemperatureParameters, detmodelparams.h:68-122
        while self.m <= self.s:
            self.s -= 1
        self.N = self.L * self.L
        self.beta = self.m * self.dtau


class SdwOracle(SweepSkeleton):
    """DetSDW<CB_ASSAAD_BERG, OPDIM> restated (cdwU == 0 only, box proposals, delayed updates)."""

    def __init__(self, pars, rng=None, phi=None):
        p = self.p = pars
        assert p.cdwU == 0.0, "oracle restates the cdwU == 0 path only"
        self.rng = rng if rng is not None else RngOracle(p.seed, p.rngIndex)
        self.N, self.L = p.N, p.L
        self.msf = 4 if p.opdim == 3 else 2                 # MatrixSizeFactor, detsdwopdim.h
        self.init_skeleton(self.msf * p.N, p.m, p.s, 1, np.complex128)
        self.phi = np.zeros((p.m + 1, p.opdim, p.N))
        self.cosh_term = np.zeros((p.m + 1, p.N))
        self.sinh_term = np.zeros((p.m + 1, p.N))
        self.phi_delta = 0.5                                 # AdjustmentData::InitialPhiDelta
        self.ra_box = RunningAverage(100)
        self.last_acc_ratio = 0.0
        self.performed_sweeps = 0
        self.accepted_global_shifts = 0
        self.attempted_global_shifts = 0
        self.wolff_stats = dict(attempted=0, accepted=0, attempted_shift=0, accepted_shift=0, added_size=0.0)
        self.decisions = None                                # optional per-proposal log
        self._setup_lattice()
        if phi is None:
            self._setup_random_field()
        else:
            self.phi[...] = phi
            self.update_cosh_sinh_terms()
        self._setup_checkerboard()
        self.setup_udv_storage_and_calculate_green()

    # ------------------------------------------------------------------ lattice, neighbortable.h
    def _setup_lattice(self):
        L, N = self.L, self.N
        site = np.arange(N)
        x, y = site % L, site // L
        self.neigh = np.stack([y * L + (x + 1) % L, y * L + (x - 1) % L,
                               ((y + 1) % L) * L + x, ((y - 1) % L) * L + x])   # +x,-x,+y,-y

    # ------------------------------------------------------------------ detsdwopdim.cpp:1098-1113
    def _setup_random_field(self):
        p = self.p
        for k in range(1, p.m + 1):
            for site in range(p.N):
                for dim in range(p.opdim):
                    self.phi[k, dim, site] = self.rng.rand_range(-1.0, 1.0)
                self.rng.rand01()                      # cdwl draw, consumed even when cdwU == 0
        self.update_cosh_sinh_terms()

    # detsdwopdim.cpp:1131-1136
    def cosh_sinh_term(self, phivec):
        nrm = np.sqrt(np.sum(np.asarray(phivec) ** 2, axis=0))
        a = self.p.lam * self.p.dtau * nrm
        return np.cosh(a), np.sinh(a) / nrm

    def update_cosh_sinh_terms(self):
        for k in range(1, self.p.m + 1):
            self.cosh_term[k], self.sinh_term[k] = self.cosh_sinh_term(self.phi[k])

    # ------------------------------------------------------------------ checkerboard
    def _plaquette_matrix(self, band, i1, i2, prefactor):
        """4x4 exp(prefactor * h_plaquette) for sites (i, j=i+x, k=i+y, l=k+x);
        detsdwopdim.cpp:1597-1684 (flux) and 1786-1826 (no flux: cosh/sinh closed form)."""
        p, L = self.p, self.L
        hh = (p.txhor, p.tyhor)[band]
        hv = (p.txver, p.tyver)[band]
        if p.bc in (1, 3) and i1 == L - 1:
            hh = -hh
        if p.bc in (2, 3) and i2 == L - 1:
            hv = -hv
        if not p.weakZflux:
            # prefactor = sign*dtau(/2): ch = cosh(-dtau t), sh = sign * sinh(-dtau t) = sinh(-prefactor t)
            chh, shh = np.cosh(prefactor * hh), np.sinh(-prefactor * hh)
            chv, shv = np.cosh(prefactor * hv), np.sinh(-prefactor * hv)
            return np.array([[chh * chv, chv * shh, chh * shv, shh * shv],
                             [chv * shh, chh * chv, shh * shv, chh * shv],
                             [chh * shv, shh * shv, chh * chv, chv * shh],
                             [shh * shv, chh * shv, chv * shh, chh * chv]], dtype=np.complex128)
        zmag = 1.0 / p.N                              # zmag[XUP] = zmag[YDOWN] = +1/N (cpp:218-222)
        j1 = (i1 + 1) % L
        k2 = (i2 + 1) % L
        ph_ij = np.exp(-2j * np.pi * zmag * i2)
        ph_kl = np.exp(-2j * np.pi * zmag * k2)
        ph_ik = ph_jl = 1.0
        if i2 == L - 1:
            ph_ik = np.exp(2j * np.pi * zmag * L * i1)
            ph_jl = np.exp(2j * np.pi * zmag * L * j1)
        h = np.zeros((4, 4), dtype=np.complex128)
        h[0, 1] = ph_ij * hh
        h[0, 2] = ph_ik * hv
        h[1, 3] = ph_jl * hv
        h[2, 3] = ph_kl * hh
        h = -(h + h.conj().T)
        w, v = np.linalg.eigh(h)
        return (v * np.exp(prefactor * w)[None, :]) @ v.conj().T

    def _subgroup_matrix(self, band, subgroup, prefactor):
        """Dense N x N product of the disjoint plaquette exponentials of one subgroup."""
        L, N = self.L, self.N
        E = np.zeros((N, N), dtype=np.complex128)
        for i1 in range(subgroup, L, 2):
            for i2 in range(subgroup, L, 2):
                i = i2 * L + i1
                j = self.neigh[0, i]
                k = self.neigh[2, i]
                l = self.neigh[0, k]
                idx = np.array([i, j, k, l])
                E[np.ix_(idx, idx)] = self._plaquette_matrix(band, i1, i2, prefactor)
        return E

    def _setup_checkerboard(self):
        """cb[band][sign] = E1(half) E0(full) E1(half) approximating exp(sign*dtau*K_band) without the
        chemical potential (detsdwopdim.cpp:1836-1869, 1945-1979)."""
        dtau = self.p.dtau
        self.cb = {}
        for band in (XBAND, YBAND):
            for sign in (-1, +1):
                E1h = self._subgroup_matrix(band, 1, sign * 0.5 * dtau)
                E0 = self._subgroup_matrix(band, 0, sign * dtau)
                self.cb[(band, sign)] = E1h @ E0 @ E1h

    # ------------------------------------------------------------------ potential exponential
    def ev_matrices(self, sign, phi_k, cosh_k, sinh_k):
        """Per-site MSF x MSF blocks of exp(sign*dtau*V) (evMatrix, detsdwopdim.cpp:3188-3229,
        cdwU == 0).  Returns array [msf, msf, nsites]."""
        msf = self.msf
        phi_k = np.asarray(phi_k, dtype=float).reshape(self.p.opdim, -1)
        ns = phi_k.shape[1]
        c = np.broadcast_to(np.asarray(cosh_k, dtype=float), (ns,))
        x = np.broadcast_to(np.asarray(sinh_k, dtype=float), (ns,))
        ev = np.zeros((msf, msf, ns), dtype=np.complex128)
        p0 = phi_k[0]
        p1 = phi_k[1] if self.p.opdim > 1 else np.zeros(ns)
        ev[0, 0] = c
        ev[1, 1] = c
        ev[0, 1] = sign * x * (p0 - 1j * p1)
        ev[1, 0] = sign * x * (p0 + 1j * p1)
        if self.p.opdim == 3:
            p2 = phi_k[2]
            ev[2, 2] = c
            ev[3, 3] = c
            ev[0, 3] = sign * p2 * x
            ev[3, 0] = sign * p2 * x
            ev[2, 1] = -sign * p2 * x
            ev[1, 2] = -sign * p2 * x
            ev[2, 3] = sign * x * (p0 + 1j * p1)
            ev[3, 2] = sign * x * (p0 - 1j * p1)
        return ev

    def _apply_v_left(self, ev, A):
        N, msf = self.N, self.msf
        out = np.zeros_like(A)
        for r in range(msf):
            for c in range(msf):
                out[r * N:(r + 1) * N, :] += ev[r, c][:, None] * A[c * N:(c + 1) * N, :]
        return out

    def _apply_v_right(self, A, ev):
        N, msf = self.N, self.msf
        out = np.zeros_like(A)
        for r in range(msf):
            for c in range(msf):
                out[:, c * N:(c + 1) * N] += A[:, r * N:(r + 1) * N] * ev[r, c][None, :]
        return out

    # setupPropK / computePropagator (detsdwopdim.cpp:1210-1286, detmodel.cpp:31-39): dense e^{-+dtau k_band} of
    # DetSDW<CB_NONE> and of computeBmatSDW (which the reference's sweepSimple uses for every model)
    def dense_propagators(self):
        if getattr(self, "_prop", None) is not None:
            return self._prop
        p, L, N = self.p, self.p.L, self.N
        hops = {0: (p.txhor, p.txver), 1: (p.tyhor, p.tyver)}
        zmag = 1.0 / N if p.weakZflux else 0.0                       # zmag[XUP] = zmag[YDOWN] = +1/N (:219-220)
        prop = {}
        for band in (0, 1):
            hh, hv = hops[band]
            k = -p.mu * np.eye(N, dtype=np.complex128)
            for site in range(N):
                x, y = site % L, site // L
                for d in range(4):                                   # XPLUS, XMINUS, YPLUS, YMINUS
                    nx, ny, hop, phase = x, y, (hh if d < 2 else hv), 1.0 + 0j
                    if d == 0:
                        nx = (x + 1) % L
                        if p.bc in (1, 3) and x == L - 1:
                            hop = -hop
                        phase = np.exp(-2j * np.pi * zmag * y)
                    elif d == 1:
                        nx = (x - 1) % L
                        if p.bc in (1, 3) and x == 0:
                            hop = -hop
                        phase = np.exp(+2j * np.pi * zmag * y)
                    elif d == 2:
                        ny = (y + 1) % L
                        if p.bc in (2, 3) and y == L - 1:
                            hop = -hop
                        if y == L - 1:
                            phase = np.exp(+2j * np.pi * zmag * L * x)
                    else:
                        ny = (y - 1) % L
                        if p.bc in (2, 3) and y == 0:
                            hop = -hop
                        if y == 0:
                            phase = np.exp(-2j * np.pi * zmag * L * x)
                    k[site, ny * L + nx] -= hop * phase
            ev, vec = np.linalg.eigh(k)
            for sign in (-1, +1):                                    # sign = -1: e^{-dtau k} (B), +1: e^{+dtau k} (B^-1)
                prop[(band, sign)] = (vec * np.exp(sign * p.dtau * ev)) @ vec.conj().T
        self._prop = prop
        return prop

    def _dense_now(self):
        return (not getattr(self.p, "checkerboard", True)) or getattr(self, "force_dense", False)

    def _apply_k_left(self, sign, A):
        N = self.N
        mu = self.p.mu
        out = np.empty_like(A)
        if self._dense_now():
            prop = self.dense_propagators()
            for bs in range(self.msf):
                out[bs * N:(bs + 1) * N, :] = prop[(BAND_OF_BANDSPIN[bs], sign)] @ A[bs * N:(bs + 1) * N, :]
            return out
        for bs in range(self.msf):
            f = np.exp(-sign * self.p.dtau * mu)       # sign=-1 (B): e^{+dtau mu}; sign=+1: e^{-dtau mu}
            out[bs * N:(bs + 1) * N, :] = f * (self.cb[(BAND_OF_BANDSPIN[bs], sign)] @ A[bs * N:(bs + 1) * N, :])
        return out

    def _apply_k_right(self, A, sign):
        N = self.N
        mu = self.p.mu
        out = np.empty_like(A)
        if self._dense_now():
            prop = self.dense_propagators()
            for bs in range(self.msf):
                out[:, bs * N:(bs + 1) * N] = A[:, bs * N:(bs + 1) * N] @ prop[(BAND_OF_BANDSPIN[bs], sign)]
            return out
        for bs in range(self.msf):
            f = np.exp(-sign * self.p.dtau * mu)
            out[:, bs * N:(bs + 1) * N] = f * (A[:, bs * N:(bs + 1) * N] @ self.cb[(BAND_OF_BANDSPIN[bs], sign)])
        return out

    # B_k = e^{-dtau V_k} e^{-dtau K};  detsdwopdim.cpp:1994-2070, 2093-2167, 2188-2303, 2326-2402
    def left_multiply_bk(self, A, k):
        ev = self.ev_matrices(-1, self.phi[k], self.cosh_term[k], self.sinh_term[k])
        return self._apply_v_left(ev, self._apply_k_left(-1, A))

    def left_multiply_bk_inv(self, A, k):
        ev = self.ev_matrices(+1, self.phi[k], self.cosh_term[k], self.sinh_term[k])
        return self._apply_k_left(+1, self._apply_v_left(ev, A))

    def right_multiply_bk(self, A, k):
        ev = self.ev_matrices(-1, self.phi[k], self.cosh_term[k], self.sinh_term[k])
        return self._apply_k_right(self._apply_v_right(A, ev), -1)

    def right_multiply_bk_inv(self, A, k):
        ev = self.ev_matrices(+1, self.phi[k], self.cosh_term[k], self.sinh_term[k])
        return self._apply_v_right(self._apply_k_right(A, +1), ev)

    # chains, detsdwopdim.cpp:2074-2090, 2170-2186, 2305-2324, 2404-2420
    def left_multiply_bmat(self, gc, A, k2, k1):
        for k in range(k1 + 1, k2 + 1):
            A = self.left_multiply_bk(A, k)
        return A

    def left_multiply_bmat_inv(self, gc, A, k2, k1):
        for k in range(k2, k1, -1):
            A = self.left_multiply_bk_inv(A, k)
        return A

    def right_multiply_bmat(self, gc, A, k2, k1):
        for k in range(k2, k1, -1):
            A = self.right_multiply_bk(A, k)
        return A

    def right_multiply_bmat_inv(self, gc, A, k2, k1):
        for k in range(k1 + 1, k2 + 1):
            A = self.right_multiply_bk_inv(A, k)
        return A

    # ------------------------------------------------------------------ bosonic action
    # detsdwopdim.cpp:4185-4239
    def delta_s_phi(self, site, k, newphi):
        p = self.p
        old = self.phi[k, :, site]
        diff = newphi - old
        old_sq = old @ old
        new_sq = newphi @ newphi
        sq_diff = new_sq - old_sq
        pow4_diff = new_sq * new_sq - old_sq * old_sq
        k_earlier = k - 1 if k > 1 else p.m
        k_later = k + 1 if k < p.m else 1
        time_neigh = self.phi[k_later, :, site] + self.phi[k_earlier, :, site]
        space_neigh = np.zeros(p.opdim)
        for d in range(4):
            space_neigh = space_neigh + self.phi[k, :, self.neigh[d, site]]
        z = 4
        d1 = (1.0 / (p.c * p.c * p.dtau)) * (sq_diff - time_neigh @ diff)
        d2 = 0.5 * p.dtau * (z * sq_diff - 2.0 * (space_neigh @ diff))
        d3 = p.dtau * (0.5 * p.r * sq_diff + 0.25 * p.u * pow4_diff)
        return d1 + d2 + d3

    # detsdwopdim.cpp:4242-4299
    def phi_action(self):
        p = self.p
        ph = self.phi[1:]                                       # [m, dim, N]
        earlier = np.roll(ph, 1, axis=0)
        td = (ph - earlier) / p.dtau
        action = (p.dtau / (2.0 * p.c * p.c)) * np.sum(td * td)
        xd = ph - ph[:, :, self.neigh[0]]
        yd = ph - ph[:, :, self.neigh[2]]
        action += 0.5 * p.dtau * (np.sum(xd * xd) + np.sum(yd * yd))
        sq = np.sum(ph * ph, axis=1)
        action += 0.5 * p.dtau * p.r * np.sum(sq)
        action += 0.25 * p.dtau * p.u * np.sum(sq * sq)
        return float(action)

    # detsdwopdim.cpp:5204-5216
    def exchange_action(self):
        return 0.5 * self.p.dtau * float(np.sum(self.phi[1:] ** 2))

    # ------------------------------------------------------------------ local updates
    # get_delta_forsite, detsdwopdim.cpp:3177-3289
    def delta_forsite(self, newphi, k, site):
        ev_old = self.ev_matrices(+1, self.phi[k, :, site], self.cosh_term[k, site],
                                  self.sinh_term[k, site])[:, :, 0]
        c_new, s_new = self.cosh_sinh_term(newphi)
        emv_new = self.ev_matrices(-1, newphi, c_new, s_new)[:, :, 0]
        return emv_new @ ev_old - np.eye(self.msf)

    # updateInSlice_delayed with proposeNewPhiBox; detsdwopdim.cpp:3021-3175, 3920-3931
    def update_in_slice(self, k):
        """updateInSlice dispatcher, cpp:2427-2489: the pass over the slice is repeated repeatUpdateInSlice times; the
        recorded acceptance ratio is that of the last pass."""
        acc = None
        for _ in range(max(1, self.p.repeatUpdateInSlice)):
            acc = self.update_in_slice_once(k)
        return acc

    def update_in_slice_once(self, k):
        p, N, msf = self.p, self.N, self.msf
        g = self.green[0]
        rows_of = lambda site: site + N * np.arange(msf)
        accepted = 0
        site = 0
        while site < N:
            delay_now = min(p.delaySteps, N - site)
            X = np.zeros((msf * N, msf * delay_now), dtype=np.complex128)
            Y = np.zeros((msf * delay_now, msf * N), dtype=np.complex128)
            j = 0
            while j < delay_now and site < N:
                newphi = self.phi[k, :, site].copy()
                for dim in range(p.opdim):
                    newphi[dim] += self.rng.rand_range(-self.phi_delta, +self.phi_delta)
                prob_sphi = np.exp(-self.delta_s_phi(site, k, newphi))
                delta = self.delta_forsite(newphi, k, site)
                idx = rows_of(site)
                Rj = g[idx, :] + X[idx, :msf * j] @ Y[:msf * j, :]
                Sj = Rj[:, idx]
                Mj = np.eye(msf) - Sj @ delta + delta
                det = np.linalg.det(Mj)
                prob_fermion = det.real if p.opdim == 3 else abs(det) ** 2
                prob = prob_sphi * prob_fermion
                u = None
                if prob > 1.0:
                    accept = True
                else:
                    u = self.rng.rand01()
                    accept = u < prob
                if self.decisions is not None:
                    self.decisions.append((k, site, float(prob), u, bool(accept)))
                if accept:
                    accepted += 1
                    self.phi[k, :, site] = newphi
                    self.cosh_term[k, site], self.sinh_term[k, site] = self.cosh_sinh_term(newphi)
                    Cj = g[:, idx] + X[:, :msf * j] @ Y[:msf * j, idx]
                    Rj = Rj.copy()
                    Rj[np.arange(msf), idx] -= 1.0
                    X[:, msf * j:msf * (j + 1)] = Cj @ delta
                    Y[msf * j:msf * (j + 1), :] = np.linalg.inv(Mj) @ Rj
                    j += 1
                site += 1
            if j > 0:
                g += X[:, :msf * j] @ Y[:msf * j, :]
        self.last_acc_ratio = accepted / float(N)
        return self.last_acc_ratio

    # updateInSliceThermalization, detsdwopdim.cpp:3293-3375 (box proposals only)
    def update_in_slice_thermalization(self, k):
        self.update_in_slice(k)
        self.ra_box.add(self.last_acc_ratio)
        if self.ra_box.samples_added % 100 == 0:
            if self.ra_box.avg < self.p.accRatio:
                self.phi_delta *= 0.95
            elif self.ra_box.avg > self.p.accRatio:
                self.phi_delta *= 1.05

    # ------------------------------------------------------------------ global shift move
    # globalMove, attemptGlobalShiftMove; detsdwopdim.cpp:3460-3485, 3564-3645, 3755-3763
    def global_move(self):
        """DetSDW::globalMove, detsdwopdim.cpp:3460-3485: shift, Wolff cluster, Wolff cluster + shift, in this order."""
        p = self.p
        if self.performed_sweeps % p.globalUpdateInterval == 0:
            if p.globalShift:
                self.attempt_global_shift_move()
            if p.wolffClusterUpdate:
                self.attempt_wolff_cluster_update()
            if p.wolffClusterShiftUpdate:
                self.attempt_wolff_cluster_shift_update()

    # ------------------------------------------------------------------ Wolff cluster moves, cpp:3487-3562, 3647-3883
    def random_direction(self):
        """randomDirection<OPDIM>::give, cpp:3766-3803 (rngwrapper.h:70-87)."""
        od = self.p.opdim
        if od == 1:
            return np.array([-1.0 if self.rng.rand01() <= 0.5 else 1.0])
        if od == 2:
            ang = self.rng.rand_range(0.0, 2.0 * np.pi)
            return np.array([np.cos(ang), np.sin(ang)])
        ang = self.rng.rand_range(0.0, 2.0 * np.pi)
        costheta = self.rng.rand_range(-1.0, 1.0)
        sintheta = np.sqrt(1.0 - costheta * costheta)
        return np.array([np.cos(ang) * sintheta, np.sin(ang) * sintheta, costheta])

    def build_and_flip_cluster(self, update_cosh_sinh):
        """buildAndFlipCluster, cpp:3805-3883: reflect phi -> phi - 2 (phi.rd) rd on a cluster grown from a random
        seed (site, slice) over space (XPLUS, XMINUS, YPLUS, YMINUS) and time (PLUS, MINUS) bonds with
        p = 1 - exp(min(0, bond_arg)); the stack is LIFO; a uniform is drawn only for bond_arg < 0."""
        p = self.p
        rd = self.random_direction()

        def proj(site, k):
            return float(self.phi[k, :, site] @ rd)

        def flip(site, k):
            ph = self.phi[k, :, site].copy()
            self.phi[k, :, site] = ph - 2.0 * float(ph @ rd) * rd
            if update_cosh_sinh:
                c, x = self.cosh_sinh_term(self.phi[k][:, site:site + 1])
                self.cosh_term[k][site] = c[0]
                self.sinh_term[k][site] = x[0]

        visited = np.zeros((p.N, p.m + 1), dtype=bool)
        k = self.rng.rand_int(1, p.m)
        site = self.rng.rand_int(0, p.N - 1)
        flip(site, k)
        visited[site, k] = True
        stack = [(site, k)]
        size = 1
        while stack:
            site, k = stack.pop()
            for d in range(4):
                nb = int(self.neigh[d, site])
                if not visited[nb, k]:
                    bond = 2.0 * p.dtau * proj(site, k) * proj(nb, k)
                    if bond < 0 and self.rng.rand01() <= (1.0 - np.exp(bond)):
                        flip(nb, k)
                        visited[nb, k] = True
                        stack.append((nb, k))
                        size += 1
            for kn in (k + 1 if k < p.m else 1, k - 1 if k > 1 else p.m):
                if not visited[site, kn]:
                    bond = (2.0 / p.dtau) * proj(site, k) * proj(site, kn)
                    if bond < 0 and self.rng.rand01() <= (1.0 - np.exp(bond)):
                        flip(site, kn)
                        visited[site, kn] = True
                        stack.append((site, kn))
                        size += 1
        return size

    def _global_backup(self):
        return (self.phi.copy(), self.cosh_term.copy(), self.sinh_term.copy(), self.green[0],
                self.green_inv_sv[0], self.storage[0])

    def _global_restore(self, backup):
        (self.phi, self.cosh_term, self.sinh_term, self.green[0], self.green_inv_sv[0], self.storage[0]) = backup

    def _fermion_ratio(self, old_sv):
        log_prob = float(np.sum(np.log(self.green_inv_sv[0]) - np.log(old_sv)))
        prob = np.exp(log_prob)
        return prob ** 2 if self.p.opdim < 3 else prob

    def attempt_wolff_cluster_update(self):
        """attemptWolffClusterUpdate, cpp:3487-3562."""
        p = self.p
        assert self.current_timeslice == p.m
        backup = self._global_backup()
        old_sv = self.green_inv_sv[0]
        sizes = [self.build_and_flip_cluster(True) for _ in range(p.repeatWolffPerSweep)]
        self.setup_udv_storage_and_calculate_green()
        prob = self._fermion_ratio(old_sv)
        self.wolff_stats["attempted"] += 1
        acc = prob >= 1.0 or self.rng.rand01() < prob
        if acc:
            self.wolff_stats["accepted"] += 1
            self.wolff_stats["added_size"] += float(sum(sizes))
        else:
            self._global_restore(backup)
        self.last_wolff = dict(prob=float(prob), sizes=sizes, accepted=bool(acc))

    def attempt_wolff_cluster_shift_update(self):
        """attemptWolffClusterShiftUpdate, cpp:3647-3748: cluster flips, then a global shift; the bosonic action
        difference is that of the shift alone (the cluster flips are rejection free for the bosonic part)."""
        p = self.p
        assert self.current_timeslice == p.m
        backup = self._global_backup()
        old_sv = self.green_inv_sv[0]
        sizes = [self.build_and_flip_cluster(False) for _ in range(p.repeatWolffPerSweep)]
        old_action = self.phi_action()
        for dim in range(p.opdim):
            r = self.rng.rand_range(-self.phi_delta, +self.phi_delta)
            self.phi[:, dim, :] += r
        new_action = self.phi_action()
        prob_scalar = np.exp(-(new_action - old_action))
        self.update_cosh_sinh_terms()
        self.setup_udv_storage_and_calculate_green()
        prob = prob_scalar * self._fermion_ratio(old_sv)
        self.wolff_stats["attempted_shift"] += 1
        acc = prob >= 1.0 or self.rng.rand01() < prob
        if acc:
            self.wolff_stats["accepted_shift"] += 1
            self.wolff_stats["added_size"] += float(sum(sizes))
        else:
            self._global_restore(backup)
        self.last_wolff = dict(prob=float(prob), sizes=sizes, accepted=bool(acc))

    def attempt_global_shift_move(self):
        p = self.p
        old_action = self.phi_action()
        assert self.current_timeslice == p.m
        backup = (self.phi.copy(), self.cosh_term.copy(), self.sinh_term.copy(), self.green[0],
                  self.green_inv_sv[0], self.storage[0])
        old_sv = self.green_inv_sv[0]
        for dim in range(p.opdim):
            r = self.rng.rand_range(-self.phi_delta, +self.phi_delta)
            self.phi[:, dim, :] += r                       # all slices, including the unused k = 0
        self.update_cosh_sinh_terms()
        self.setup_udv_storage_and_calculate_green()
        new_action = self.phi_action()
        prob_scalar = np.exp(-(new_action - old_action))
        log_prob = float(np.sum(np.log(self.green_inv_sv[0]) - np.log(old_sv)))
        prob_fermion = np.exp(log_prob)
        if p.opdim < 3:
            prob_fermion = prob_fermion ** 2
        prob = prob_scalar * prob_fermion
        self.last_global_shift = dict(prob=float(prob), log_det_ratio=log_prob,
                                      dS=float(new_action - old_action))
        self.attempted_global_shifts += 1
        if prob >= 1.0 or self.rng.rand01() < prob:
            self.accepted_global_shifts += 1
            self.last_global_shift["accepted"] = True
        else:
            (self.phi, self.cosh_term, self.sinh_term, self.green[0], self.green_inv_sv[0],
             self.storage[0]) = backup
            self.last_global_shift["accepted"] = False

    # ------------------------------------------------------------------ sweeps, cpp:4422-4502
    def sweep(self):
        self.sweep_skeleton(self.update_in_slice)
        self.performed_sweeps += 1

    def sweep_thermalization(self):
        self.sweep_skeleton(self.update_in_slice_thermalization)
        self.performed_sweeps += 1

    # ------------------------------------------------------------------ fermionic measurements, cpp:508-900, 903-1000
    def shift_green_symmetric(self):
        """shiftGreenSymmetric, cpp:4505-4612 (CB_ASSAAD_BERG): every N x N block (row, col) of g becomes
        E0(-dtau/2) E1(-dtau/2) [band of the row block] . block . E1(+dtau/2) E0(+dtau/2) [band of the column block]
        (two plaquette subgroups, half steps, no chemical potential); blocks 0..3 = XUP, YDOWN, XDOWN, YUP."""
        N, h = self.N, 0.5 * self.p.dtau
        if not hasattr(self, "_shift_mats"):
            self._shift_mats = {b: (self._subgroup_matrix(b, 0, -h) @ self._subgroup_matrix(b, 1, -h),
                                    self._subgroup_matrix(b, 1, +h) @ self._subgroup_matrix(b, 0, +h)) for b in (0, 1)}
        g = self.green[0]
        out = np.zeros_like(g)
        for row in range(self.msf):
            for col in range(self.msf):
                blk = g[row * N:(row + 1) * N, col * N:(col + 1) * N]
                out[row * N:(row + 1) * N, col * N:(col + 1) * N] = \
                    self._shift_mats[row % 2][0] @ blk @ self._shift_mats[col % 2][1]
        return out

    def init_measurements(self):
        N = self.N
        self.fm = dict(greenK0=0.0, greenLocal=0.0, occDiffSq=0.0, kOccX=np.zeros(N), kOccY=np.zeros(N),
                       pairPlus=np.zeros(N), pairMinus=np.zeros(N), slices=0)

    def measure_fermionic(self, k):
        """The fermionic part of measure(timeslice), cpp:540-900, from the symmetrically shifted Green's function of the
        current slice.  Band-spin blocks: XUP = 0, YDOWN = 1, XDOWN = 2, YUP = 3; for opdim < 3 only the XUP / YDOWN
        blocks are stored and the XDOWN / YUP sector is their complex conjugate (gl1, cpp:598-615)."""
        p, N, L = self.p, self.N, self.L
        gs = self.shift_green_symmetric()
        fm = self.fm
        fm["slices"] += 1
        XUP, YDOWN, XDOWN, YUP = 0, 1, 2, 3
        XB, YB, UP, DOWN = 0, 1, 0, 1
        bandspin = {(XB, UP): XUP, (YB, DOWN): YDOWN, (XB, DOWN): XDOWN, (YB, UP): YUP}

        def gl1(s1, bs1, s2, bs2):
            if p.opdim == 3:
                return gs[s1 + N * bs1, s2 + N * bs2]
            if bs1 in (XUP, YDOWN) and bs2 in (XUP, YDOWN):
                return gs[s1 + N * bs1, s2 + N * bs2]
            if bs1 in (XDOWN, YUP) and bs2 in (XDOWN, YUP):
                return np.conj(gs[s1 + N * (bs1 - 2), s2 + N * (bs2 - 2)])
            return 0.0

        def gl(s1, b1, sp1, s2, b2, sp2):
            return gl1(s1, bandspin[(b1, sp1)], s2, bandspin[(b2, sp2)])

        if p.opdim == 3:
            fm["greenK0"] += float(np.real(np.sum(gs)))
            fm["greenLocal"] += float(np.real(np.trace(gs))) / (4.0 * N)
        else:
            fm["greenK0"] += 2.0 * float(np.real(np.sum(gs)))
            fm["greenLocal"] += 2.0 * float(np.real(np.trace(gs))) / (4.0 * N)
        # momentum-space occupation, cpp:623-671
        off_x = 0.5 if p.bc in (1, 3) else 0.0
        off_y = 0.5 if p.bc in (2, 3) else 0.0
        sites = np.arange(N)
        ix, iy = (sites % L).astype(float), (sites // L).astype(float)
        gx = np.array([[gl1(i, XUP, j, XUP) + gl1(i, XDOWN, j, XDOWN) for j in range(N)] for i in range(N)])
        gy = np.array([[gl1(i, YUP, j, YUP) + gl1(i, YDOWN, j, YDOWN) for j in range(N)] for i in range(N)])
        for ks in range(N):
            ky = -np.pi + (float(ks // L) + off_y) * 2 * np.pi / L
            kx = -np.pi + (float(ks % L) + off_x) * 2 * np.pi / L
            phase = np.exp(1j * (kx * (ix[:, None] - ix[None, :]) + ky * (iy[:, None] - iy[None, :])))
            fm["kOccX"][ks] += float(np.real(np.sum(phase * gx)))
            fm["kOccY"][ks] += float(np.real(np.sum(phase * gy)))
        # equal-time pairing correlations, cpp:673-718
        for i in range(N):
            pp = pm = 0.0
            for a, b in ((i, 0), (0, i)):
                for b1 in (XB, YB):
                    for b2 in (XB, YB):
                        t = gl(a, b1, DOWN, b, b2, UP) * gl(a, b1, UP, b, b2, DOWN) - \
                            gl(a, b1, DOWN, b, b2, DOWN) * gl(a, b1, UP, b, b2, UP)
                        pp += -4.0 * t
                        pm += -4.0 * t * (1.0 if b1 == b2 else -1.0)
            fm["pairPlus"][i] += float(np.real(pp))
            fm["pairMinus"][i] += float(np.real(pm))
        # occDiffSq, cpp:744-776
        tot = 0.0
        for i in range(N):
            g = lambda b1, s1, b2, s2: gl(i, b1, s1, i, b2, s2)      # noqa: E731
            tot += (-2.0 * g(XB, DOWN, XB, UP) * g(XB, UP, XB, DOWN) + g(XB, UP, XB, UP)
                    + 2.0 * g(XB, DOWN, YB, DOWN) * g(YB, DOWN, XB, DOWN)
                    + 2.0 * g(XB, UP, YB, DOWN) * g(YB, DOWN, XB, UP) + g(YB, DOWN, YB, DOWN)
                    - 2.0 * g(XB, UP, XB, UP) * g(YB, DOWN, YB, DOWN)
                    + 2.0 * g(XB, DOWN, YB, UP) * g(YB, UP, XB, DOWN)
                    + 2.0 * g(XB, UP, YB, UP) * g(YB, UP, XB, UP)
                    - 2.0 * g(YB, DOWN, YB, UP) * g(YB, UP, YB, DOWN)
                    + g(XB, DOWN, XB, DOWN) * (1.0 + 2.0 * g(XB, UP, XB, UP) - 2.0 * g(YB, DOWN, YB, DOWN)
                                               - 2.0 * g(YB, UP, YB, UP))
                    + g(YB, UP, YB, UP) - 2.0 * g(XB, UP, XB, UP) * g(YB, UP, YB, UP)
                    + 2.0 * g(YB, DOWN, YB, DOWN) * g(YB, UP, YB, UP))
        fm["occDiffSq"] += float(np.real(tot)) / N

    def finish_measurements(self):
        """finishMeasurements, cpp:903-1000 (fermionic part)."""
        p, N, L, m = self.p, self.N, self.L, self.p.m
        fm = self.fm
        assert fm["slices"] == m
        out = dict(greenK0=fm["greenK0"] / m, greenLocal=fm["greenLocal"] / m,
                   kOccX=2.0 - fm["kOccX"] / (m * N), kOccY=2.0 - fm["kOccY"] / (m * N),
                   pairPlus=fm["pairPlus"] / m, pairMinus=fm["pairMinus"] / m)
        far = [yy * L + xx for yy in (L // 2 - 1, L // 2, L // 2 + 1) for xx in (L // 2 - 1, L // 2, L // 2 + 1)]
        out["pairPlusMax"] = float(np.mean(out["pairPlus"][far]))
        out["pairMinusMax"] = float(np.mean(out["pairMinus"][far]))
        out["occDiffSq"] = fm["occDiffSq"] / m
        return out

    def measured_sweep_fermionic(self):
        """sweep(true) with fermionic measurements: measure(k) right after updateInSlice(k) (detmodel.h:1277-1283)."""
        self.init_measurements()

        def update_and_measure(k):
            self.update_in_slice(k)
            self.measure_fermionic(k)
        self.sweep_skeleton(update_and_measure)
        self.performed_sweeps += 1
        return self.finish_measurements()

    # greenUpdate = simple, cpp:4366-4420
    def sweep_simple(self):
        self.sweep_simple_skeleton(self.update_in_slice)
        self.performed_sweeps += 1

    def sweep_simple_thermalization(self):
        self.sweep_simple_skeleton(self.update_in_slice_thermalization)
        self.performed_sweeps += 1


# ---------------------------------------------------------------------------------------------
class HubbardParams:
    def __init__(self, **kw):
        self.L = 4
        self.m = 40
        self.s = 10
        self.dtau = 0.1
        self.t = 1.0
        self.U = 4.0
        self.mu = 0.0
        self.checkerboard = False
        self.seed = 1020304050
        self.rngIndex = 1
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)
        while self.m <= self.s:
            self.s -= 1
        self.N = self.L * self.L


class HubbardOracle(SweepSkeleton):
    """DetHubbard restated (dethubbard.cpp:46-171, 741-765, 823-933; dethubbard.h:281-337)."""

    def __init__(self, pars, rng=None, aux=None):
        p = self.p = pars
        self.rng = rng if rng is not None else RngOracle(p.seed, p.rngIndex)
        self.N, self.L = p.N, p.L
        self.init_skeleton(p.N, p.m, p.s, 2, np.float64)
        self.alpha = np.arccosh(np.exp(p.dtau * p.U * 0.5))
        self.aux = np.zeros((p.m + 1, p.N), dtype=np.int32)
        if aux is None:
            for k in range(1, p.m + 1):
                for site in range(p.N):
                    self.aux[k, site] = +1 if self.rng.rand01() <= 0.5 else -1
        else:
            self.aux[...] = aux
        self._setup_proptmat()
        self.setup_udv_storage_and_calculate_green()

    def _setup_proptmat(self):
        p, L, N = self.p, self.L, self.N
        site = np.arange(N)
        x, y = site % L, site // L
        neigh = np.stack([y * L + (x + 1) % L, y * L + (x - 1) % L,
                          ((y + 1) % L) * L + x, ((y - 1) % L) * L + x])
        if not p.checkerboard:
            tmat = -p.mu * np.eye(N)
            for s_ in range(N):
                for d in range(4):
                    tmat[neigh[d, s_], s_] -= p.t
            w, v = np.linalg.eigh(tmat)
            self.proptmat = (v * np.exp(-p.dtau * w)[None, :]) @ v.T       # detmodel.cpp:21-29
        else:
            # dethubbard.cpp:768-821
            kxa = np.zeros((N, N)); kxb = np.zeros((N, N)); kya = np.zeros((N, N)); kyb = np.zeros((N, N))
            for yy in range(L):
                for xx in range(0, L, 2):
                    a = yy * L + xx
                    na = neigh[0, a]
                    kxa[a, na] = kxa[na, a] = 1.0
                    nb = neigh[0, na]
                    kxb[na, nb] = kxb[nb, na] = 1.0
            for xx in range(L):
                for yy in range(0, L, 2):
                    a = yy * L + xx
                    na = neigh[2, a]
                    kya[a, na] = kya[na, a] = 1.0
                    nb = neigh[2, na]
                    kyb[na, nb] = kyb[nb, na] = 1.0
            ch, sh = np.cosh(p.dtau * p.t), np.sinh(p.dtau * p.t)
            self.proptmat = (ch ** 4 * np.eye(N) + ch ** 3 * sh * (kxa + kxb + kya + kyb)
                             + ch ** 2 * sh ** 2 * (kxa @ kxb + kxa @ kya + kxb @ kya + kxa @ kyb
                                                   + kxb @ kyb + kya @ kyb)
                             + ch * sh ** 3 * (kxa @ kxb @ kya + kxa @ kxb @ kyb + kxa @ kya @ kyb
                                               + kxb @ kya @ kyb)
                             + sh ** 4 * kxa @ kxb @ kya @ kyb)

    # dethubbard.cpp:823-851
    def compute_bmat(self, gc, k2, k1):
        sign = +1.0 if gc == 0 else -1.0
        if k2 == k1:
            return np.eye(self.N)
        single = lambda k: np.exp(sign * self.alpha * self.aux[k].astype(float))[:, None] * self.proptmat
        B = single(k2)
        for k in range(k2 - 1, k1, -1):
            B = B @ single(k)
        return B

    def left_multiply_bmat(self, gc, A, k2, k1):
        return self.compute_bmat(gc, k2, k1) @ A

    def right_multiply_bmat(self, gc, A, k2, k1):
        return A @ self.compute_bmat(gc, k2, k1)

    def left_multiply_bmat_inv(self, gc, A, k2, k1):
        return np.linalg.inv(self.compute_bmat(gc, k2, k1)) @ A

    def right_multiply_bmat_inv(self, gc, A, k2, k1):
        return A @ np.linalg.inv(self.compute_bmat(gc, k2, k1))

    # dethubbard.cpp:141-171, 892-933
    def update_in_slice(self, k):
        N = self.N
        gup, gdn = self.green
        for _ in range(N):
            site = self.rng.rand_int(0, N - 1)
            a = float(self.aux[k, site])
            e_up = np.exp(-2.0 * self.alpha * a)
            e_dn = np.exp(+2.0 * self.alpha * a)
            ratio = (1.0 + (e_up - 1.0) * (1.0 - gup[site, site])) * \
                    (1.0 + (e_dn - 1.0) * (1.0 - gdn[site, site]))
            if ratio > 1.0 or self.rng.rand01() < ratio:
                for g, delta in ((gup, e_up - 1.0), (gdn, e_dn - 1.0)):
                    one_minus = np.eye(N) - g
                    factor = delta / (1.0 + delta * one_minus[site, site])
                    g -= np.outer(g[:, site], factor * one_minus[site, :])
                self.aux[k, site] *= -1

    def sweep(self):
        self.sweep_skeleton(self.update_in_slice)

    sweep_thermalization = sweep


# ---------------------------------------------------------------------------------------------
def exchange_probability(par1, action1, par2, action2):
    """detsdwopdim.cpp:5251-5264."""
    delta = (par1 - par2) * (action2 - action1)
    return 1.0 if delta <= 0.0 else float(np.exp(-delta))


def replica_exchange_walk(control_values, current_par_process, current_process_par, actions, rand01):
    """Serial ladder walk of DetQMCPT::replicaExchangeStep (detqmcpt.h:1031-1079), rank-0 part only.

    control_values[cpi]        -- the ladder (parspt.controlParameterValues)
    current_par_process[cpi]   -- which replica ("process") currently holds parameter index cpi
    current_process_par[pi]    -- which parameter index replica pi currently holds
    actions[pi]                -- get_exchange_action_contribution() of replica pi
    rand01                     -- callable drawing from replica 0's RNG stream
    Returns (new par_process, new process_par, list of (cpi1, prob, accepted)); the caller swaps the
    control data (step sizes / statistics) of the two replicas on every accepted entry."""
    par_process = list(current_par_process)
    process_par = list(current_process_par)
    log = []
    P = len(control_values)
    for cpi1 in range(P - 1):
        cpi2 = cpi1 + 1
        p1, p2 = par_process[cpi1], par_process[cpi2]
        prob = exchange_probability(control_values[cpi1], actions[p1], control_values[cpi2], actions[p2])
        accepted = prob >= 1 or rand01() <= prob
        if accepted:
            process_par[p1] = cpi2
            process_par[p2] = cpi1
            par_process[cpi1] = p2
            par_process[cpi2] = p1
        log.append((cpi1, prob, bool(accepted)))
    return par_process, process_par, log
