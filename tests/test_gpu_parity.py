"""GPU parity tests (run on the B200 box): every operator of the C ABI against the CPU oracle
(oracle/dqmc_oracle.py, pinned to the reference) and against the golden vectors generated from the
unmodified reference (tests/golden, tools/make_golden.py).

Tolerances: the north star asks for G, log-determinants and acceptance ratios within 1e-10
relative (FP64) and identical accept/reject trajectories; operators that are plain sums of
products are held to 1e-12."""
import numpy as np
import pytest

from helpers import load_golden, sdw_params_of, maxabs, relerr

pytestmark = pytest.mark.gpu

SDW_CASES = ["sdw_o2_wolff_L4", "sdw_o2_wolffshift_L4", "sdw_o3_woodbury_L4", "sdw_o2_repeat2_L4",
             "sdw_o2_flux_L4", "sdw_o2_noflux_apbcxy_L4", "sdw_o3_L4", "sdw_o1_apbcx_L4",
             "sdw_o2_flux_L4_delay3_s7", "sdw_o2_flux_L6"]

TOL_G = 1e-10


def make_batch(p, **kw):
    from detqmc_b200 import DetSDWBatch
    return DetSDWBatch(p, n_replicas=kw.pop("n_replicas", 1), **kw)


def test_stabilizers_agree():
    """Blocked pre-pivoted QR (default) and the fully pivoted QR give the same Green's function."""
    from dqmc_oracle import SdwParams
    p = SdwParams(L=6, m=40, s=10)
    b0 = make_batch(p)
    b1 = make_batch(p, full_pivot=True)
    assert relerr(b0.green(), b1.green()) < 1e-11
    assert abs(b0.logdet() - b1.logdet()) < 1e-10 * abs(b1.logdet())
    for b in (b0, b1):
        b.sweepThermalization()
    assert maxabs(b0.phi(), b1.phi()) == 0.0
    assert relerr(b0.green(), b1.green()) < 1e-10


def rand_cplx(shape, seed):
    g = np.random.default_rng(seed)
    return g.standard_normal(shape) + 1j * g.standard_normal(shape)


# ------------------------------------------------------------------------------------------------
def test_library_refuses_without_fallback():
    """The product path must fail loudly rather than fall back: a bogus device id is an error."""
    from detqmc_b200 import DetSDWBatch, DqmcError
    from dqmc_oracle import SdwParams
    with pytest.raises(DqmcError):
        DetSDWBatch(SdwParams(), n_replicas=1, device=4096)


def test_error_behaviour():
    """Parameter combinations the reference's check() rejects (detsdwparams.cpp:38-200) and calls in the wrong state
    come back as errors with a message, never as a silent fallback."""
    from detqmc_b200 import DetSDWBatch, DqmcError
    from dqmc_oracle import SdwParams
    bad = [dict(L=5),                                                  # checkerboard break-up needs an even L
           dict(opdim=3, weakZflux=True),                              # flux only for O(1), O(2) (detsdwparams.cpp:57-60)
           dict(wolffClusterShiftUpdate=True, globalShift=True),       # combined move excludes the others (:94-96)
           dict(delaySteps=0), dict(m=1), dict(globalUpdateInterval=0)]
    for kw in bad:
        with pytest.raises(DqmcError):
            DetSDWBatch(SdwParams(**kw), n_replicas=1)
    with pytest.raises(DqmcError):
        DetSDWBatch(SdwParams(cdwU=1.0), n_replicas=1)
    b = make_batch(SdwParams())
    with pytest.raises(DqmcError):
        b.wrap_down(3)                                                 # currentTimeslice is m after the set-up
    with pytest.raises(DqmcError):
        b.fermionic_observables()                                      # no measured sweep yet
    with pytest.raises(DqmcError):
        b.global_shift_move() if b.wrap_down(b.m) else b.global_shift_move()   # needs currentTimeslice == m
    with pytest.raises(DqmcError):
        b.set_lanes(2)                                                 # more lanes than replicas
    assert b.lib.dqmc_last_error(b.h)                                  # a message is kept for the caller


@pytest.mark.parametrize("shape", [(288, 288, 288), (128, 128, 128), (100, 37, 53), (32, 32, 32), (196, 784, 64)])
@pytest.mark.parametrize("trans", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_dmma(shape, trans):
    from dqmc_oracle import SdwParams
    b = make_batch(SdwParams(), init="none")
    M, N, K = shape
    ta, tb = trans
    A = rand_cplx((K, M) if ta else (M, K), 1)
    B = rand_cplx((N, K) if tb else (K, N), 2)
    ref = (A.conj().T if ta else A) @ (B.conj().T if tb else B)
    C = b.gemm(A, B, ta, tb)
    assert relerr(C, ref) < 1e-13


@pytest.mark.parametrize("name", SDW_CASES)
def test_bmat_mult_vs_golden(name):
    g = load_golden(name)
    p = sdw_params_of(g)
    b = make_batch(p)
    assert maxabs(b.phi(), g["phi0"]) == 0.0          # same dSFMT stream, same construction order
    A = g["A"]
    k2, k1 = [int(v) for v in g["chain"]]
    for op in range(4):
        r1 = g["bmult_op%d_single" % op]
        rc = g["bmult_op%d_chain" % op]
        assert relerr(b.bmat_mult(op, A, 3, 2), r1) < 1e-12, (name, op)
        assert relerr(b.bmat_mult(op, A, k2, k1), rc) < 1e-12, (name, op)


@pytest.mark.parametrize("name", ["sdw_o2_flux_L4", "sdw_o3_L4"])
def test_bmat_adjoint_and_inverse_identities(name):
    from dqmc_oracle import SdwOracle
    g = load_golden(name)
    p = sdw_params_of(g)
    o = SdwOracle(p)
    b = make_batch(p)
    A = g["A"]
    eye = np.eye(o.sz, dtype=np.complex128)
    Bd = o.left_multiply_bmat(0, eye, 9, 4)
    assert relerr(b.bmat_mult(4, A, 9, 4), Bd.conj().T @ A) < 1e-12          # LEFT_ADJ
    back = b.bmat_mult(2, b.bmat_mult(0, A, 9, 4), 9, 4)                     # B^-1 (B A) = A
    assert relerr(back, A) < 1e-12
    back = b.bmat_mult(3, b.bmat_mult(1, A, 9, 4), 9, 4)                     # (A B) B^-1 = A
    assert relerr(back, A) < 1e-12


@pytest.mark.parametrize("full_pivot", [False, True])
@pytest.mark.parametrize("name", ["sdw_o2_flux_L4", "sdw_o3_L4", "sdw_o2_flux_L6", "L12"])
def test_udt_decompose(name, full_pivot):
    if name == "L12":
        from dqmc_oracle import SdwParams
        p = SdwParams(L=12, m=20, s=10)
    else:
        p = sdw_params_of(load_golden(name))
    b = make_batch(p, init="none", full_pivot=full_pivot)
    D = b.D
    # graded matrix like the ones in the chain: (random) * diag(scales over 20 decades), columns shuffled
    M = rand_cplx((D, D), 5) * np.logspace(8, -12, D)[None, :]
    M = M[:, np.random.default_rng(9).permutation(D)]
    Q, d, T = b.udt_decompose(M)
    assert maxabs(Q.conj().T @ Q, np.eye(D)) < 1e-13
    rec = (Q * d[None, :]) @ T
    colnorm = np.linalg.norm(M, axis=0)
    assert np.max(np.linalg.norm(rec - M, axis=0) / colnorm) < 1e-12        # column-wise backward error
    assert np.all(d > 0)
    if full_pivot:
        assert np.all(np.diff(d) <= 1e-12 * d[:-1])                         # pivoting sorts |R_ii|
        assert np.abs(T).max() < 1.0 + 1e-12                                # T = D^-1 R P^T, |T_ij| <= 1
    else:
        # pre-pivoting orders the columns by norm once: |R_ii| decreases up to O(1) factors and T stays
        # well scaled, which is all the stabilised chain needs
        assert np.all(d[1:] <= 50.0 * d[:-1])
        assert np.abs(T).max() < 100.0
        assert np.linalg.cond(T) < 1e6


@pytest.mark.parametrize("name", ["sdw_o2_flux_L4", "sdw_o3_L4", "sdw_o2_flux_L6"])
def test_green_from_udt_vs_reference(name):
    g = load_golden(name)
    p = sdw_params_of(g)
    b = make_batch(p, init="none")
    # reference UdV (M = U d V_t^+): right chain Q d T with T = V_t^+, left chain T^+ d Q^+ with
    # T = U^+, Q = V_t
    Qr, dr, Tr = g["udv1_U"], g["udv1_d"], g["udv1_V"].conj().T
    Ql, dl, Tl = g["udv2_V"], g["udv2_d"], g["udv2_U"].conj().T
    G, logdet = b.green_from_udt(Qr, dr, Tr, Ql, dl, Tl)
    ref = g["green_from_udv_l2_r1"]
    assert relerr(G, ref) < TOL_G
    ref_logdet = np.log(g["sv_from_udv_l2_r1"]).sum()
    assert abs(logdet - ref_logdet) < 1e-10 * max(1.0, abs(ref_logdet))


@pytest.mark.parametrize("name", SDW_CASES)
def test_setup_green_and_logdet(name):
    g = load_golden(name)
    p = sdw_params_of(g)
    b = make_batch(p)
    assert relerr(b.green(), g["green0"]) < TOL_G
    ref_logdet = np.log(g["sv0"]).sum()
    assert abs(b.logdet() - ref_logdet) < 1e-10 * abs(ref_logdet)
    assert abs(b.phi_action()[0] - float(g["phiAction0"])) < 1e-10 * abs(float(g["phiAction0"]))
    assert abs(b.get_exchange_action_contribution()[0] - float(g["exchangeAction0"])) < 1e-11 * float(g["exchangeAction0"])
    assert relerr(b.green_for_timeslice(3), g["green_slice_3"]) < TOL_G
    s = int(p.s)
    assert relerr(b.green_for_timeslice(s), g["green_slice_%d" % s]) < TOL_G


@pytest.mark.parametrize("name", ["sdw_o2_flux_L4", "sdw_o3_L4", "sdw_o2_flux_L4_delay3_s7"])
def test_wrap_and_advance_without_updates(name):
    """The stabilised skeleton alone: a full down- and up-sweep with no field updates, G compared
    with the oracle after every wrap and every advance."""
    from dqmc_oracle import SdwOracle
    g = load_golden(name)
    p = sdw_params_of(g)
    o = SdwOracle(p)
    b = make_batch(p)
    n, s, m = o.n, o.s, o.m
    worst = 0.0

    def check():
        nonlocal worst
        worst = max(worst, relerr(b.green(), o.green[0]))

    for k in range(m, (n - 1) * s, -1):
        o.wrap_down_green(k, 0); b.wrap_down(k); check()
    for l in range(n - 1, 0, -1):
        o.advance_down_green(l + 1, 0); b.advance_down(l + 1); check()
        for k in range(l * s, (l - 1) * s, -1):
            o.wrap_down_green(k, 0); b.wrap_down(k); check()
    o.advance_down_green(1, 0); b.advance_down(1); check()
    assert b.green_consistency()[0] < 1e-9
    # up-sweep (sweepUp resets storage[0] to the identity UdV; advance_up(0) does the same)
    from dqmc_oracle import UdV
    o.storage[0][0] = UdV.eye(o.sz, o.dtype)
    for l in range(0, n - 1):
        for k in range(l * s + 1, (l + 1) * s + 1):
            o.wrap_up_green(k - 1, 0); b.wrap_up(k - 1); check()
        o.advance_up_green(l, 0); b.advance_up(l); check()
    for k in range((n - 1) * s + 1, m + 1):
        o.wrap_up_green(k - 1, 0); b.wrap_up(k - 1); check()
    o.advance_up_green(n - 1, 0); b.advance_up(n - 1); check()
    assert b.green_consistency()[0] < 1e-9
    assert worst < TOL_G


@pytest.mark.parametrize("name", ["sdw_o2_flux_L4", "sdw_o3_L4", "sdw_o1_apbcx_L4", "sdw_o2_flux_L4_delay3_s7"])
def test_update_in_slice_vs_oracle(name):
    from dqmc_oracle import SdwOracle
    g = load_golden(name)
    p = sdw_params_of(g)
    o = SdwOracle(p)
    o.decisions = []
    b = make_batch(p)
    acc = b.update_in_slice(p.m)
    ratio = o.update_in_slice(p.m)
    assert acc[0] == round(ratio * p.N)
    assert maxabs(b.phi()[1:], o.phi[1:]) < 1e-13
    assert relerr(b.green(), o.green[0]) < TOL_G
    assert b.rng_draw(4)[0] == o.rng.rand01()           # same number of values consumed


@pytest.mark.parametrize("name", SDW_CASES)
def test_sweeps_vs_golden(name):
    g = load_golden(name)
    p = sdw_params_of(g)
    b = make_batch(p)
    n = int(g["n_sweeps"])
    for sw in range(n):
        b.sweepThermalization()
        cd = b.control_data()
        assert cd.lastAccRatioLocal_phi == g["lastAccRatio"][sw], (name, sw)
        assert cd.acceptedGlobalShifts == g["acceptedGlobalShifts"][sw], (name, sw)
        assert abs(cd.phiDelta - g["phiDelta"][sw]) < 1e-14
        key = "phi_after_%d" % (sw + 1)
        if key in g.files:
            assert maxabs(b.phi()[1:], g[key][1:]) < 1e-12
            assert relerr(b.green(), g["green_after_%d" % (sw + 1)]) < TOL_G
    assert maxabs(b.phi()[1:], g["phi_final"][1:]) < 1e-12
    assert np.array_equal(b.rng_draw(8), g["rng_next"])


def test_trajectory_100_sweeps():
    """Identical accept/reject trajectory over the first 100 sweeps with the same dSFMT seed."""
    g = load_golden("sdw_o2_flux_L4_traj100")
    p = sdw_params_of(g)
    b = make_batch(p)
    for sw in range(100):
        b.sweepThermalization()
        cd = b.control_data()
        assert cd.lastAccRatioLocal_phi == g["lastAccRatio"][sw], sw
        assert cd.acceptedGlobalShifts == g["acceptedGlobalShifts"][sw], sw
    assert abs(b.phi_delta() - g["phiDelta"][-1]) < 1e-13
    assert maxabs(b.phi()[1:], g["phi_final"][1:]) < 1e-11
    assert relerr(b.green(), g["green_final"]) < TOL_G
    assert np.array_equal(b.rng_draw(8), g["rng_next"])


def test_measurement_sweeps_do_not_adapt_step_size():
    from dqmc_oracle import SdwOracle
    g = load_golden("sdw_o2_flux_L4")
    p = sdw_params_of(g)
    o = SdwOracle(p)
    b = make_batch(p)
    for _ in range(3):
        o.sweep(); b.sweep()
        assert b.control_data().lastAccRatioLocal_phi == o.last_acc_ratio
    assert b.phi_delta() == 0.5
    assert maxabs(b.phi()[1:], o.phi[1:]) < 1e-12


def test_batch_of_replicas_matches_single_replicas():
    """Replicas of one batch are independent: each equals its own single-replica oracle run."""
    from dqmc_oracle import SdwOracle, SdwParams
    rs = [-1.5, -0.5, 0.3]
    idx = [1, 2, 3]
    base = dict(L=4, m=20, s=10)
    b = make_batch(SdwParams(**base), n_replicas=3, rng_indices=idx, r_values=rs)
    oracles = [SdwOracle(SdwParams(r=r, rngIndex=i, **base)) for r, i in zip(rs, idx)]
    for _ in range(3):
        b.sweepThermalization()
        for o in oracles:
            o.sweep_thermalization()
    act = b.get_exchange_action_contribution()
    for rep, o in enumerate(oracles):
        assert maxabs(b.phi(rep)[1:], o.phi[1:]) < 1e-12
        assert relerr(b.green(rep), o.green[0]) < TOL_G
        assert abs(act[rep] - o.exchange_action()) < 1e-11 * o.exchange_action()
        assert b.control_data(rep).lastAccRatioLocal_phi == o.last_acc_ratio


def test_lanes_do_not_change_results():
    """Issuing the replicas as lanes on separate CUDA streams (default: one lane per replica; here also an
    uneven split into 3 lanes) must not change any result compared with a single lane."""
    from dqmc_oracle import SdwParams
    p = SdwParams(L=4, m=20, s=10)
    idx = list(range(1, 9))
    bd = make_batch(p, n_replicas=8, rng_indices=idx)         # default: 8 lanes
    b3 = make_batch(p, n_replicas=8, rng_indices=idx)
    b3.set_lanes(3)
    b1 = make_batch(p, n_replicas=8, rng_indices=idx)
    b1.set_lanes(1)
    for _ in range(3):
        for b in (b1, b3, bd):
            b.sweepThermalization()
    for b in (b3, bd):
        for rep in range(8):
            assert maxabs(b1.phi(rep), b.phi(rep)) == 0.0
            assert maxabs(b1.green(rep), b.green(rep)) == 0.0
            assert b1.control_data(rep).lastAccRatioLocal_phi == b.control_data(rep).lastAccRatioLocal_phi


def test_large_batch_kernel_variants_match_single_replica_runs():
    """The library picks its kernel shapes by the number of matrices in flight: more than 16 replicas (the headline
    batch of 64) run the 96 x 48 rank-K flush, the 32-wide gather tiles, 16-vector checkerboard row tiles, the wide GEMM
    tiles and no programmatic dependent launch; small batches the 32 x 32 flush, 8-wide gather tiles, 8-vector tiles,
    deep k-tiles.  At the headline matrix size (L = 12, D = 288, flux, delaySteps = 16, short imaginary time) replicas of
    an 18-replica batch must follow exactly the trajectories of the same replicas run one at a time."""
    from dqmc_oracle import SdwParams
    base = dict(L=12, m=20, s=10, opdim=2, weakZflux=True, delaySteps=16, updateMethod=2, globalShift=True,
                globalUpdateInterval=2)
    R = 18
    idx = list(range(1, R + 1))
    rs = np.linspace(-1.9, 0.4, R)
    big = make_batch(SdwParams(**base), n_replicas=R, rng_indices=idx, r_values=rs)
    for _ in range(3):
        big.sweepThermalization()
    pick = (0, 7, 17)
    ref = [(big.phi(i).copy(), big.green(i).copy(), big.control_data(i).lastAccRatioLocal_phi) for i in pick]
    big.close()
    for (phi, g, acc), i in zip(ref, pick):
        one = make_batch(SdwParams(**base), n_replicas=1, rng_indices=[idx[i]], r_values=[rs[i]])
        for _ in range(3):
            one.sweepThermalization()
        assert maxabs(one.phi(0)[1:], phi[1:]) == 0.0, i
        assert relerr(one.green(0), g) < TOL_G, i
        assert one.control_data(0).lastAccRatioLocal_phi == acc, i
        one.close()


def test_resident_random_numbers_equal_streamed():
    """dqmc_rng_preload (the random-number stream of several sweeps resident in HBM, what bench.py times as `value`)
    gives the same trajectory as the default streamed mode, across global moves (sweeps 0 and 10, where the host
    draws from the same streams) and with replica-exchange steps consuming look-ahead uniforms in between."""
    from dqmc_oracle import SdwParams
    for wolff in (False, True):                # global shift only: in-place window; with clusters: re-headed window
      p = SdwParams(L=4, m=20, s=10, wolffClusterUpdate=wolff)
      idx = [1, 2, 3]
      a = make_batch(p, n_replicas=3, rng_indices=idx)
      c = make_batch(p, n_replicas=3, rng_indices=idx)
      c.rng_preload(14)
      for sw in range(12):
          a.sweepThermalization()
          c.sweepThermalization()
      c.rng_release()
      for rep in range(3):
          assert maxabs(a.phi(rep), c.phi(rep)) == 0.0
          assert maxabs(a.green(rep), c.green(rep)) == 0.0
          assert a.control_data(rep).phiDelta == c.control_data(rep).phiDelta
          assert a.control_data(rep).acceptedGlobalShifts == c.control_data(rep).acceptedGlobalShifts
          assert list(a.wolff_statistics(rep)) == list(c.wolff_statistics(rep))
          assert maxabs(a.rng_draw(3, rep=rep), c.rng_draw(3, rep=rep)) == 0.0


def test_random_number_modes_interleaved():
    """Streamed sweeps, then a pre-loaded window (re-allocates the device buffer), release, streamed again, a second
    pre-load of a different size: the captured sweep graphs must never replay with the address of a freed buffer
    (dqmc_rng_preload vs the graph cache).  Trajectory equals a plain streamed run."""
    from dqmc_oracle import SdwParams
    p = SdwParams(L=4, m=20, s=10)
    idx = [1, 2]
    a = make_batch(p, n_replicas=2, rng_indices=idx)
    c = make_batch(p, n_replicas=2, rng_indices=idx)
    plan = [("stream", 3), ("preload", 12, 4), ("stream", 3), ("preload", 3, 2), ("preload", 20, 3), ("stream", 2)]
    for step in plan:
        n = step[-1]
        if step[0] == "preload":
            c.rng_preload(step[1])
        for _ in range(n):
            a.sweepThermalization()
            c.sweepThermalization()
        if step[0] == "preload":
            c.rng_release()
        for rep in range(2):
            assert maxabs(a.phi(rep), c.phi(rep)) == 0.0, step
    for rep in range(2):
        assert maxabs(a.green(rep), c.green(rep)) == 0.0
        assert maxabs(a.rng_draw(3, rep=rep), c.rng_draw(3, rep=rep)) == 0.0


def test_global_shift_move_vs_oracle():
    from dqmc_oracle import SdwOracle
    g = load_golden("sdw_o2_flux_L4")
    p = sdw_params_of(g)
    o = SdwOracle(p)
    b = make_batch(p)
    for _ in range(4):
        o.attempt_global_shift_move()
        acc = b.global_shift_move()
        assert bool(acc[0]) == o.last_global_shift["accepted"]
        assert maxabs(b.phi()[1:], o.phi[1:]) < 1e-13
        assert relerr(b.green(), o.green[0]) < TOL_G
        assert abs(b.logdet() - np.log(o.green_inv_sv[0]).sum()) < 1e-9


def test_wolff_cluster_moves_vs_reference_record():
    """dqmc_wolff_cluster_move against the record of the reference's own attemptWolffClusterUpdate /
    attemptWolffClusterShiftUpdate (tests/golden/wolff_moves.npz): identical clusters (fields),
    accept decisions and statistics, Green's function within 1e-10, same position of the random-number stream."""
    import json
    import os
    from dqmc_oracle import SdwParams
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wolff_moves.npz"))
    for tag in ("o2", "o2_shift", "o3_rep2", "o1"):
        d = json.loads(str(g[tag + "_pars"]))
        for key in ("N", "beta"):
            d.pop(key, None)
        p = SdwParams(**d)
        b = make_batch(p, n_replicas=2, rng_indices=[p.rngIndex, p.rngIndex + 10])
        shift = bool(g[tag + "_shift"])
        prev = np.zeros(5)
        for it in range(g[tag + "_phi"].shape[0]):
            acc = b.wolff_cluster_move(shift)
            st = g[tag + "_stats"][it]
            assert list(b.wolff_statistics(0)) == list(st)
            assert bool(acc[0]) == bool((st - prev)[3 if shift else 1])
            prev = st
            assert maxabs(b.phi(0)[1:], g[tag + "_phi"][it][1:]) < 1e-13
            assert relerr(b.green(0), g[tag + "_green"][it]) < TOL_G
        assert maxabs(b.rng_draw(4, rep=0), g[tag + "_rng_next"]) == 0.0
        # the Wolff statistics are part of the control data (UpdateStatistics, detsdwopdim.cpp:5227): they follow the
        # control parameter when the blob is installed on another replica (exchange, checkpoint)
        cd = b.control_data(0)
        assert [cd.attemptedWolffClusterUpdates, cd.acceptedWolffClusterUpdates, cd.attemptedWolffClusterShiftUpdates,
                cd.acceptedWolffClusterShiftUpdates, cd.addedWolffClusterSize] == list(b.wolff_statistics(0))
        b.set_control_data(cd, rep=1)
        assert list(b.wolff_statistics(1)) == list(b.wolff_statistics(0))


def _dense_params(g, tag, **over):
    import json
    from dqmc_oracle import SdwParams
    d = json.loads(str(g["params_" + tag]))
    for k in ("N", "beta"):
        d.pop(k, None)
    d.update(over)
    return SdwParams(**d)


@pytest.mark.parametrize("tag", ["flux", "noflux_apbcx"])
def test_dense_hopping_bmat_and_sweep_simple_vs_reference(tag):
    """SURVEY 8a rows a8 / a26 against the reference itself (tests/golden/sdw_dense_hopping.npz): the dense B matrices of
    computeBmatSDW (detsdwopdim.cpp:1307-1497) and the reference's sweepSimple, which builds its B matrices with the
    dense hopping exponential whatever the checkerboard setting (detsdwopdim.cpp:4366-4420)."""
    g = load_golden("sdw_dense_hopping")
    p = _dense_params(g, tag)
    bd = make_batch(_dense_params(g, tag, checkerboard=False))          # same stream -> same random fields
    D = 2 * p.N
    eye = np.eye(D, dtype=np.complex128)
    for k2, k1 in ((7, 3), (20, 19)):
        want = g["bmat_%s_%d_%d" % (tag, k2, k1)]
        assert relerr(bd.bmat_mult(0, eye, k2, k1), want) < 1e-12      # B I
        assert relerr(bd.bmat_mult(1, eye, k2, k1), want) < 1e-12      # I B
        A = rand_cplx((D, D), 3)
        assert relerr(bd.bmat_mult(2, bd.bmat_mult(0, A, k2, k1), k2, k1), A) < 1e-11      # B^-1 B A
        assert relerr(bd.bmat_mult(3, bd.bmat_mult(1, A, k2, k1), k2, k1), A) < 1e-11      # A B B^-1
    b = make_batch(p)                                                   # checkerboard model, simple sweeps
    for _ in range(2):
        b.sweepSimpleThermalization()
    assert maxabs(b.phi()[1:], g["simple_phi_" + tag]) < 1e-12
    assert relerr(b.green(), g["simple_green_" + tag]) < 1e-9
    assert np.array_equal(b.rng_draw(4), g["simple_rng_next_" + tag])


def test_dense_hopping_model_vs_reference():
    """checkerboard = false: the stabilised sweep of DetSDW<CB_NONE, 2> (dense e^{-dtau K} in every B-matrix product)
    against the reference's own run of that model."""
    g = load_golden("sdw_dense_hopping")
    b = make_batch(_dense_params(g, "cbnone"))
    assert relerr(b.green(), g["cbnone_green0"]) < TOL_G
    assert abs(b.logdet() - float(g["cbnone_logdet0"])) < 1e-10 * abs(float(g["cbnone_logdet0"]))
    for _ in range(4):
        b.sweepThermalization()
    assert maxabs(b.phi()[1:], g["cbnone_phi"]) < 1e-12
    assert relerr(b.green(), g["cbnone_green"]) < 1e-9
    assert np.array_equal(b.rng_draw(4), g["cbnone_rng_next"])


def test_sweep_simple_vs_oracle():
    """greenUpdate = simple (dqmc_sweep_simple): G from scratch at every slice, then the slice update -- against the
    oracle restatement (plain inverse of 1 + B(k,0) B(m,k) with the checkerboard B): identical decisions, fields and
    step sizes, G within the accuracy of the unstabilised inverse at beta = 2."""
    from dqmc_oracle import SdwOracle, SdwParams
    p = SdwParams(L=4, m=20, s=10, rngIndex=15)
    o = SdwOracle(p)
    b = make_batch(p, rng_indices=[15])
    for it in range(3):
        if it < 2:
            o.sweep_simple_thermalization()
            b.sweepSimpleThermalization()
        else:
            o.sweep_simple()
            b.sweepSimple(False)
        assert maxabs(b.phi()[1:], o.phi[1:]) < 1e-13
        assert relerr(b.green(), o.green[0]) < 1e-8
        assert b.control_data().lastAccRatioLocal_phi == o.last_acc_ratio
        assert b.phi_delta() == o.phi_delta
    assert maxabs(b.rng_draw(4), [o.rng.rand01() for _ in range(4)]) == 0.0


@pytest.mark.parametrize("tag", ["o2", "o2_apbc", "o3", "o2_L6"])
def test_fermionic_observables_vs_golden(tag):
    """dqmc_sweep(ctx, 2) + dqmc_get_fermionic_observables against the observables the reference
    measured in its own sweep(true) for the same seed: greenK0, greenLocal, occDiffSq, pairPlusMax, pairMinusMax and the
    vectors kOccX, kOccY, pairPlus, pairMinus after each of three measured sweeps."""
    import json
    import os
    from dqmc_oracle import SdwParams
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fermion_observables.npz"))
    d = json.loads(str(g[tag + "_pars"]))
    for key in ("N", "beta", "fermionMeasurements"):
        d.pop(key, None)
    p = SdwParams(**d)
    b = make_batch(p, n_replicas=2, rng_indices=[p.rngIndex, p.rngIndex + 20])
    for it in range(3):
        b.sweep(True)
        ob = b.fermionic_observables(0)
        sc = [ob["greenK0"], ob["greenLocal"], ob["occDiffSq"], ob["pairPlusMax"], ob["pairMinusMax"]]
        vec = np.concatenate([ob["kOccX"], ob["kOccY"], ob["pairPlus"], ob["pairMinus"]])
        assert np.allclose(sc, g[tag + "_scalars"][it], rtol=1e-8, atol=1e-10)
        assert np.allclose(vec, g[tag + "_vectors"][it], rtol=1e-8, atol=1e-10)
    assert maxabs(b.phi(0)[1:], g[tag + "_phi"][1:]) < 1e-12
    assert relerr(b.green(0), g[tag + "_green"]) < TOL_G
    other = b.fermionic_observables(1)                  # the second replica measured its own trajectory
    assert abs(other["greenLocal"] - ob["greenLocal"]) > 0


def test_bosonic_observables_vs_golden():
    """The observables of a measured sweep (normMeanPhi, associatedEnergy, phiRhoS_Gs, phiRhoS_Gc) from the fields
    on the device against the values the reference measured for the same fields (tests/golden/bosonic_observables.npz)."""
    import json
    import os
    from dqmc_oracle import SdwParams
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bosonic_observables.npz"))
    for tag in ("o2", "o3", "o2_L6"):
        d = json.loads(str(g[tag + "_pars"]))
        for key in ("N", "beta"):
            d.pop(key, None)
        b = make_batch(SdwParams(**d))
        for it in range(g[tag + "_phi"].shape[0]):
            b.set_phi(g[tag + "_phi"][it])
            ob = b.bosonic_observables()
            got = [ob["normMeanPhi"], ob["associatedEnergy"], ob["phiRhoS_Gs"], ob["phiRhoS_Gc"]]
            assert np.allclose(got, g[tag + "_obs"][it], rtol=1e-12, atol=1e-13)


def test_config_stream_vs_golden(tmp_path):
    """dqmc_download_config_stream and the stream writers of the mirror against the bytes / lines the reference's
    own writers produced for the same fields (tests/golden/config_streams.npz, detsdwopdim.cpp:4943-5036)."""
    import json
    import os
    from dqmc_oracle import SdwParams, config_stream
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_streams.npz"))
    for tag in ("o2", "o3", "o2_L6"):
        d = json.loads(str(g[tag + "_pars"]))
        for key in ("N", "beta"):
            d.pop(key, None)
        p = SdwParams(**d)
        b = make_batch(p, n_replicas=2, rng_indices=[1, 2])
        b.set_phi(g[tag + "_phi"], rep=1)
        want = np.frombuffer(g[tag + "_binary"].tobytes(), dtype=np.float64)
        assert np.array_equal(b.config_stream(1), want)
        assert np.array_equal(b.config_stream(0), config_stream(b.phi(0)))
        both = b.config_stream(-1)
        assert both.shape == (2, want.size) and np.array_equal(both[1], want) and np.array_equal(both[0], b.config_stream(0))
        out = tmp_path / tag
        out.mkdir()
        b.saveConfigurationStreamBinary(str(out), rep=1)
        b.saveConfigurationStreamText(str(out), rep=1)
        assert open(str(out / "configs-phi.binarystream"), "rb").read() == g[tag + "_binary"].tobytes()
        assert open(str(out / "configs-phi.textstream"), "rb").read() == g[tag + "_text"].tobytes()


def test_parallel_tempering_driver_vs_oracle_loop(tmp_path):
    """detqmc_b200.DetQMCPT (the MPI-free DetQMCPT::run, detqmcpt.h:760-958) against the same loop
    over oracle replicas: thermalisation and measurement sweeps with a replica-exchange step after every sweep
    (ladder walk with replica 0's stream, control data following the parameter), observables recorded per control
    parameter, configuration streams written to the sub-directory of the parameter held at that time."""
    import copy
    import os
    from detqmc_b200 import DetQMCPT
    from detqmc_b200.pt import bosonic_observables, OBSERVABLES
    from dqmc_oracle import SdwOracle, SdwParams, config_stream, exchange_probability
    values = np.array([-1.6, -1.2, -0.8, -0.4])
    P, therm, sweeps, cfg_interval = len(values), 3, 4, 2
    kw = dict(L=4, m=20, s=10)
    pt = DetQMCPT(SdwParams(**kw), values, thermalization=therm, sweeps=sweeps, exchangeInterval=1,
                  saveConfigurationStreamInterval=cfg_interval, saveConfigurationStreamBinary=True, outdir=str(tmp_path))
    pt.run()

    reps = [SdwOracle(SdwParams(r=float(values[i]), rngIndex=i + 1, **kw)) for i in range(P)]
    par_process, process_par = list(range(P)), list(range(P))
    accepted = np.zeros(P - 1, dtype=int)
    series = {c: {o: [] for o in OBSERVABLES} for c in range(P)}
    cfgs = {c: [] for c in range(P)}

    def get_ctrl(o):
        return (o.phi_delta, o.last_acc_ratio, copy.deepcopy(o.ra_box), o.accepted_global_shifts, o.attempted_global_shifts)

    def set_ctrl(o, c):
        o.phi_delta, o.last_acc_ratio, o.ra_box, o.accepted_global_shifts, o.attempted_global_shifts = c

    for step in range(therm + sweeps):
        for o in reps:
            if step < therm:
                o.sweep_thermalization()
            else:
                o.sweep()
        if step >= therm:
            for pi, o in enumerate(reps):
                ob = bosonic_observables(o.phi, o.p.dtau)
                for name in OBSERVABLES:
                    series[process_par[pi]][name].append(ob[name])
                if (step - therm + 1) % cfg_interval == 0:
                    cfgs[process_par[pi]].append(config_stream(o.phi))
        if step == therm + sweeps - 1:
            break                                        # no exchange after the last sweep (detqmcpt.h:948-953)
        actions = [o.exchange_action() for o in reps]
        ctrl = [get_ctrl(o) for o in reps]
        for c1 in range(P - 1):
            p1, p2 = par_process[c1], par_process[c1 + 1]
            prob = exchange_probability(values[c1], actions[p1], values[c1 + 1], actions[p2])
            if prob >= 1 or reps[0].rng.rand01() <= prob:
                accepted[c1] += 1
                process_par[p1], process_par[p2] = c1 + 1, c1
                par_process[c1], par_process[c1 + 1] = p2, p1
                ctrl[p1], ctrl[p2] = ctrl[p2], ctrl[p1]
        for pi, o in enumerate(reps):
            o.p.r = float(values[process_par[pi]])
            set_ctrl(o, ctrl[pi])

    assert list(pt.ladder.process_par) == process_par
    assert list(pt.ladder.accepted[:P - 1]) == list(accepted)
    for c in range(P):
        d = pt.subdir(c)
        for name in OBSERVABLES:
            got = [float(x) for x in open(os.path.join(d, name + ".series")) if x[0] != "#"]
            assert len(got) == sweeps and np.allclose(got, series[c][name], rtol=1e-10, atol=1e-12)
        want = np.concatenate(cfgs[c]) if cfgs[c] else np.zeros(0)
        path = os.path.join(d, "configs-phi.binarystream")
        got = np.fromfile(path) if os.path.exists(path) else np.zeros(0)
        assert got.shape == want.shape and (got.size == 0 or maxabs(got, want) < 1e-12)
    acc_file = [x.split() for x in open(os.path.join(str(tmp_path), "exchange-acceptance.values")) if x[0] != "#"]
    assert np.allclose([float(a[1]) for a in acc_file][:P - 1], [accepted[c] / (therm + sweeps - 1) for c in range(P - 1)],
                       rtol=1e-12, atol=0)


def test_parallel_tempering_vs_reference_driver(tmp_path):
    """detqmc_b200.DetQMCPT against the reference's OWN replica-exchange driver: tests/golden/pt_reference.npz holds the
    output of the unmodified DetQMCPT<DetSDW<CB_ASSAAD_BERG, 2>> (detqmcpt.h) run with one thread per ladder process
    on a thread-backed stand-in for boost::mpi (oracle/ref_pt_harness.cpp, tools/make_golden.py pt_reference).  Same
    seed, same ladder: the time series of every control parameter (which replica's measurement lands where depends on
    every accepted swap), the exchange acceptance ratios and the diffusion fractions must be the reference's."""
    import os
    from detqmc_b200 import DetQMCPT
    from detqmc_b200.pt import OBSERVABLES
    from dqmc_oracle import SdwParams
    g = load_golden("pt_reference")
    values = g["values"]
    P = len(values)
    m = int(round(float(g["beta"]) / 0.1))
    pt = DetQMCPT(SdwParams(L=int(g["L"]), m=m, s=int(g["s"])), values, thermalization=int(g["thermalization"]),
                  sweeps=int(g["sweeps"]), exchangeInterval=int(g["exchangeInterval"]), outdir=str(tmp_path))
    pt.run()
    for c in range(P):
        d = pt.subdir(c)
        assert os.path.basename(d) == str(g["subdirs"][c])            # the reference's directory names
        for name in OBSERVABLES:
            got = [float(x) for x in open(os.path.join(d, name + ".series")) if x[0] != "#"]
            want = g["series_" + name][c]                             # the reference writes 6 significant digits
            assert len(got) == len(want) and np.allclose(got, want, rtol=2e-5, atol=2e-6), (c, name, got, want)
    for fname, key in (("exchange-acceptance.values", "acceptance"), ("exchange-diffusion.values", "diffusion")):
        rows = [x.split() for x in open(os.path.join(str(tmp_path), fname)) if x[0] != "#"]
        assert np.allclose([float(r[1]) for r in rows], g[key], rtol=1e-12, atol=1e-15), fname


def test_parallel_tempering_resume(tmp_path):
    """DetQMCPT state save / resume (detqmcpt.h:201-253, 447-520): a ladder run stopped after 4 of 8 measurement
    sweeps and resumed from the per-rank state files reproduces the uninterrupted run -- per-parameter series,
    exchange acceptance and diffusion (fields, control data, generator position, ladder assignment and exchange
    statistics are all part of the state)."""
    import os
    from detqmc_b200 import DetQMCPT
    from detqmc_b200.pt import OBSERVABLES
    from dqmc_oracle import SdwParams
    values = np.array([-1.6, -1.2, -0.8, -0.4])
    # exchanges after sweeps 3, 6, 9 of 12: none is due at the stopping point (sweep 8).  Like the reference, a run
    # that stops does not exchange after its last sweep (detqmcpt.h:948-953), so a stop AT an exchange point would
    # differ from the uninterrupted run by that one exchange step -- in the reference as well
    kw = dict(thermalization=4, exchangeInterval=3)
    a, b = os.path.join(str(tmp_path), "straight"), os.path.join(str(tmp_path), "resumed")
    DetQMCPT(SdwParams(L=4, m=20, s=10, globalUpdateInterval=3), values, sweeps=8, outdir=a, **kw).run()
    DetQMCPT(SdwParams(L=4, m=20, s=10, globalUpdateInterval=3), values, sweeps=4, outdir=b, **kw).run()
    pt = DetQMCPT(SdwParams(L=4, m=20, s=10, globalUpdateInterval=3), values, sweeps=8, outdir=b, resume=True, **kw)
    assert pt.sweepsDone == 4 and pt.sweepsDoneThermalization == 4
    pt.run()
    for c in range(len(values)):
        for name in OBSERVABLES:
            sa = [float(x) for x in open(os.path.join(a, os.path.basename(pt.subdir(c)), name + ".series")) if x[0] != "#"]
            sb = [float(x) for x in open(os.path.join(pt.subdir(c), name + ".series")) if x[0] != "#"]
            assert len(sa) == 8 and len(sb) == 8 and np.allclose(sa, sb, rtol=0, atol=1e-9), (c, name)
    for fname in ("exchange-acceptance.values", "exchange-diffusion.values"):
        ra = [x.split()[1] for x in open(os.path.join(a, fname)) if x[0] != "#"]
        rb_ = [x.split()[1] for x in open(os.path.join(b, fname)) if x[0] != "#"]
        assert ra == rb_, fname


def test_parallel_tempering_driver_fermionic_series(tmp_path):
    """DetQMCPT with turnoffFermionMeasurements = False: the measurement sweeps are sweep(true); without exchanges
    (exchangeInterval = 0) every control parameter's greenLocal / occDiffSq series is the oracle's for that replica."""
    import os
    from detqmc_b200 import DetQMCPT
    from dqmc_oracle import SdwOracle, SdwParams
    values = np.array([-1.4, -0.6])
    kw = dict(L=4, m=20, s=10)
    pt = DetQMCPT(SdwParams(**kw), values, thermalization=2, sweeps=2, exchangeInterval=0, outdir=str(tmp_path),
                  turnoffFermionMeasurements=False)
    pt.run()
    for c in range(2):
        o = SdwOracle(SdwParams(r=float(values[c]), rngIndex=c + 1, **kw))
        for _ in range(2):
            o.sweep_thermalization()
        want = {"greenLocal": [], "occDiffSq": []}
        for _ in range(2):
            ob = o.measured_sweep_fermionic()
            for name in want:
                want[name].append(ob[name])
        for name in want:
            got = [float(x) for x in open(os.path.join(pt.subdir(c), name + ".series")) if x[0] != "#"]
            assert np.allclose(got, want[name], rtol=1e-9, atol=1e-11)
        rows = [x.split() for x in open(os.path.join(pt.subdir(c), "results-kOccX.values")) if x[0] != "#"]
        assert len(rows) == 16


def _pt_rank(rank, world, port, outdir):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from detqmc_b200 import DetQMCPT
    from dqmc_oracle import SdwParams
    values = np.array([-1.6, -1.2, -0.8, -0.4])
    pt = DetQMCPT(SdwParams(L=4, m=20, s=10), values, thermalization=3, sweeps=3, exchangeInterval=1,
                  saveConfigurationStreamInterval=1, saveConfigurationStreamBinary=True, outdir=outdir, device=0)
    pt.run()
    if rank == 0:
        np.save(os.path.join(outdir, "process_par.npy"), pt.ladder.process_par)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_parallel_tempering_two_ranks_equal_one_rank(tmp_path):
    """The ladder split over two processes (two contexts on the one GPU of the test box, payloads all-gathered with
    gloo) gives the same files as a single process holding all replicas: the multi-rank path of DetQMCPT -- rank
    offsets of the random-number streams, payload layout, identical ladder walk on every rank, gather of the records
    and configuration streams to rank 0."""
    import os
    import socket
    import torch.multiprocessing as mp
    from detqmc_b200 import DetQMCPT
    from dqmc_oracle import SdwParams
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    two = tmp_path / "two"
    one = tmp_path / "one"
    two.mkdir()
    one.mkdir()
    mp.spawn(_pt_rank, args=(2, port, str(two)), nprocs=2, join=True)
    values = np.array([-1.6, -1.2, -0.8, -0.4])
    pt = DetQMCPT(SdwParams(L=4, m=20, s=10), values, thermalization=3, sweeps=3, exchangeInterval=1,
                  saveConfigurationStreamInterval=1, saveConfigurationStreamBinary=True, outdir=str(one))
    pt.run()
    assert list(np.load(str(two / "process_par.npy"))) == list(pt.ladder.process_par)
    for c in range(len(values)):
        sub = os.path.basename(pt.subdir(c))
        for name in ("normMeanPhi.series", "associatedEnergy.series", "phiRhoS_Gs.series", "phiRhoS_Gc.series",
                     "configs-phi.binarystream"):
            assert open(str(one / sub / name), "rb").read() == open(str(two / sub / name), "rb").read(), (sub, name)
    for name in ("exchange-acceptance.values", "exchange-diffusion.values", "exchange-parameters.values"):
        assert open(str(one / name)).read() == open(str(two / name)).read()


# ---------------------------------------------------------------- full-size checks
@pytest.mark.parametrize("kw", [dict(L=12, m=100, s=10), dict(L=8, m=80, s=10)])
def test_full_size_properties(kw):
    """BASELINE configs C3 (replicas of the L=12, beta=10 ladder) and C2 at full size: G after setup
    and after one sweep against the oracle, plus properties that need no oracle -- B^-1 B = 1, wrap
    up/down round trip, wrapped == recomputed G, G(0) after the sweep == from scratch."""
    from dqmc_oracle import SdwOracle, SdwParams
    p = SdwParams(**kw)
    o = SdwOracle(p)
    b = make_batch(p, n_replicas=2, rng_indices=[1, 2])
    D = b.D
    A = rand_cplx((D, D), 3)
    assert relerr(b.bmat_mult(2, b.bmat_mult(0, A, 17, 7), 17, 7), A) < 1e-11
    assert relerr(b.bmat_mult(3, b.bmat_mult(1, A, 17, 7), 17, 7), A) < 1e-11
    assert relerr(b.green(0), o.green[0]) < TOL_G
    assert abs(b.logdet(0) - np.log(o.green_inv_sv[0]).sum()) < 1e-10 * abs(b.logdet(0))
    G0 = b.green(0)
    b.wrap_down(p.m)
    b.wrap_up(p.m - 1)
    assert relerr(b.green(0), G0) < 1e-10
    b.sweepThermalization()                                         # global move + full down-sweep
    o.sweep_thermalization()
    assert b.control_data(0).lastAccRatioLocal_phi == o.last_acc_ratio
    assert b.control_data(0).acceptedGlobalShifts == o.accepted_global_shifts
    assert maxabs(b.phi(0)[1:], o.phi[1:]) < 1e-12
    assert relerr(b.green(0), o.green[0]) < TOL_G
    assert np.all(b.green_consistency() < 1e-8)                     # wrapped vs advanced G at the last advance
    assert relerr(b.green(1), b.green_for_timeslice(0, rep=1)) < 1e-9
    # a measured sweep at full size through properties that need no oracle: the k grid is complete, so the
    # momentum-space occupations sum to the real-space ones, sum_k (kOccX + kOccY) = 4 N (1 - greenLocal); 0 <= n_k <= 2
    b.sweep(True)
    for rep in range(b.R):
        ob = b.fermionic_observables(rep)
        N = p.N
        assert abs((ob["kOccX"].sum() + ob["kOccY"].sum()) - 4.0 * N * (1.0 - ob["greenLocal"])) < 1e-8 * N
        assert ob["kOccX"].min() > -1e-6 and ob["kOccX"].max() < 2.0 + 1e-6
        assert ob["kOccY"].min() > -1e-6 and ob["kOccY"].max() < 2.0 + 1e-6
        assert np.isfinite(ob["pairPlus"]).all() and np.isfinite(ob["occDiffSq"])
    assert np.all(b.green_consistency() < 1e-8)


# ---------------------------------------------------------------- full-size configs pinned to the reference itself
def _probe_vectors(D):
    """Same fixed probe vectors as tools/make_golden.py::probe_vectors."""
    gen = np.random.default_rng(20261018)
    u = gen.standard_normal(D) + 1j * gen.standard_normal(D)
    v = gen.standard_normal(D) + 1j * gen.standard_normal(D)
    return u / np.linalg.norm(u), v / np.linalg.norm(v)


def assert_matrix_matches_summary(G, g, key, tol, what):
    """Compare a full-size matrix with the reference's committed summary: strided sub-sample (element-wise, relative
    to the largest element), trace, Frobenius norm and a fixed random bilinear probe u^H G v."""
    stride = int(g["stride"])
    scale = float(g[key + "_maxabs"])
    assert np.abs(G[::stride, ::stride] - g[key + "_sub"]).max() < tol * scale, (what, key, "sub-sample")
    assert abs(np.trace(G) - g[key + "_trace"]) < tol * max(abs(g[key + "_trace"]), scale), (what, key, "trace")
    assert abs(np.linalg.norm(G) - g[key + "_fro"]) < tol * float(g[key + "_fro"]), (what, key, "norm")
    u, v = _probe_vectors(G.shape[0])
    assert abs(np.vdot(u, G @ v) - g[key + "_probe"]) < tol * scale, (what, key, "probe")


@pytest.mark.timeout(900)
@pytest.mark.parametrize("name", ["sdw_c3_L12_b10", "sdw_c4_o3_L14_b14", "sdw_c2_L8_b8_traj100"])
def test_full_size_vs_reference_golden(name):
    """BASELINE configs C2 (100-sweep trajectory), C3 (one replica of the ladder, 6 sweeps) and C4 (O(3), L = 14,
    beta = 14, D = 784) at FULL size against goldens generated by the unmodified reference (tools/make_golden.py,
    BIG fixtures): G and log|det| after set-up within 1e-10 relative, identical acceptance ratio / global-shift
    decisions / step size after every sweep, identical fields, G after the sweeps within 1e-10 -- all with the default
    (pre-pivoted blocked QR) stabiliser, the case SURVEY H3 warned about."""
    g = load_golden(name)
    p = sdw_params_of(g)
    b = make_batch(p)
    assert maxabs(b.phi(), g["phi0"]) == 0.0
    assert_matrix_matches_summary(b.green(), g, "green0", TOL_G, name)
    assert abs(b.logdet() - float(g["logdet0"])) < 1e-10 * abs(float(g["logdet0"]))
    assert_matrix_matches_summary(b.green_for_timeslice(p.m), g, "green_slice_m", TOL_G, name)
    n = int(g["n_sweeps"])
    for sw in range(n):
        b.sweepThermalization()
        cd = b.control_data()
        assert cd.lastAccRatioLocal_phi == g["lastAccRatio"][sw], (name, sw)
        assert cd.acceptedGlobalShifts == g["acceptedGlobalShifts"][sw], (name, sw)
        assert abs(cd.phiDelta - g["phiDelta"][sw]) < 1e-14
        key = "phi_after_%d" % (sw + 1)
        if key in g.files:
            assert maxabs(b.phi()[1:], g[key][1:]) < 1e-11, (name, sw)
            # the reference's own wrapped-vs-recomputed deviation at this point bounds what parity can mean
            sd = "ref_selfdev_after_%d" % (sw + 1)
            tol = max(TOL_G, 100.0 * float(g[sd])) if sd in g.files else TOL_G
            assert_matrix_matches_summary(b.green(), g, "green_after_%d" % (sw + 1), tol, (name, sw))
    assert np.array_equal(b.rng_draw(8), g["rng_next"])                # the stream was consumed exactly as in the reference


@pytest.mark.timeout(900)
def test_hubbard_full_size_vs_reference_golden():
    """BASELINE config C5 (DetHubbard L = 20, U = 8, beta = 20) at full size against the unmodified reference:
    identical auxiliary fields after each of two sweeps (80 000 decisions each) and an identically consumed random
    stream; log|det| within 1e-10.  Green's function: the reference's own SVD-based value sits 2.9e-9 (spin up) /
    9e-10 (spin down) away from the s-converged column-pivoted-QR value (tools/c5_accuracy_study.py, which wrote
    tests/golden/hubbard_c5_truth.npz), so the CUDA path is held to that converged value within 2e-10 and to the
    reference within 1e-8 = a few times the reference's own error."""
    from detqmc_b200 import DetHubbardBatch
    from helpers import hubbard_params_of
    g = load_golden("hubbard_c5_L20_U8_b20")
    truth = load_golden("hubbard_c5_truth")
    p = hubbard_params_of(g)
    b = DetHubbardBatch(p)
    assert np.array_equal(b.auxfield()[1:], g["aux0"])
    st = int(truth["stride"])
    for gc in (0, 1):
        G = b.green(0, gc)
        assert np.abs(G[::st, ::st] - truth["green0_%d_sub" % gc]).max() < 2e-10 * float(truth["green0_%d_maxabs" % gc])
        assert float(truth["ref_dev_%d" % gc]) < 1e-8                  # the reference's own distance from the converged value
        assert_matrix_matches_summary(G, g, "green0_%d" % gc, 1e-8, "c5 setup")
        ld = float(g["logdet0_%d" % gc])
        assert abs(b.logdet(0, gc) - ld) < 1e-10 * abs(ld)
    for sw in range(int(g["n_sweeps"])):
        b.sweep()
        assert np.array_equal(b.auxfield()[1:], g["aux_after_%d" % (sw + 1)]), sw
        for gc in (0, 1):
            assert_matrix_matches_summary(b.green(0, gc), g, "green_after_%d_%d" % (sw + 1, gc), 1e-8, ("c5", sw))
    assert np.array_equal(b.rng_draw(8), g["rng_next"])

# ---------------------------------------------------------------- the reference's own driver on top of the C ABI
def test_reference_driver_with_gpu_shim(tmp_path):
    """host/_build/detqmcsdw_gpu = the reference's DetQMC<Model, ModelParams> driver (compiled unmodified from
    the reference tree in the development container) instantiated with include/detsdw_gpu.h.  Its
    thermalised step size, acceptance and time series must equal the oracle's trajectory for the same seed."""
    import os
    import subprocess
    from dqmc_oracle import SdwOracle, SdwParams
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "host", "_build", "detqmcsdw_gpu")
    if not os.path.exists(exe):
        pytest.skip("host/_build/detqmcsdw_gpu not built (needs the reference tree: make -C host)")
    therm, sweeps = 10, 10
    out = subprocess.run([exe, "L=4", "beta=2", "dtau=0.1", "s=10", "r=-1", "thermalization=%d" % therm,
                          "sweeps=%d" % sweeps, "rngSeed=1020304050", "simindex=0", "saveConfigurationStreamInterval=5",
                          "saveConfigurationStreamBinary=1", "saveConfigurationStreamText=1"], cwd=str(tmp_path),
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    series = [float(x) for x in open(os.path.join(str(tmp_path), "normMeanPhi.series")) if x.strip() and x[0] != "#"]
    assert len(series) == sweeps
    others = {name: [float(x) for x in open(os.path.join(str(tmp_path), name + ".series")) if x.strip() and x[0] != "#"]
              for name in ("associatedEnergy", "phiRhoS_Gs", "phiRhoS_Gc")}
    from dqmc_oracle import config_stream, config_stream_text
    o = SdwOracle(SdwParams(L=4, m=20, s=10, r=-1.0))
    for _ in range(therm):
        o.sweep_thermalization()
    from dqmc_oracle import bosonic_observables as oracle_observables
    ref, cfg_bin, cfg_txt = [], [], ""
    ref_others = {name: [] for name in others}
    for sw in range(1, sweeps + 1):
        o.sweep()
        ref.append(float(np.linalg.norm(o.phi[1:].mean(axis=(0, 2)))))
        ob = oracle_observables(o.phi, o.p.dtau)
        for name in others:
            ref_others[name].append(ob[name])
        if sw % 5 == 0:                                              # saveConfigurationStreamInterval (detqmc.h:484-491)
            cfg_bin.append(config_stream(o.phi))
            cfg_txt += config_stream_text(o.phi)
    assert np.allclose(series, ref, rtol=0, atol=2e-6)               # the driver writes 6 significant digits
    for name in others:                                               # the reference's bosonic observable list
        assert len(others[name]) == sweeps and np.allclose(others[name], ref_others[name], rtol=2e-5, atol=2e-6)
    # configuration streams written by the shim through dqmc_download_config_stream
    got = np.fromfile(os.path.join(str(tmp_path), "configs-phi.binarystream"))
    assert got.shape == (2 * 16 * 20 * 2,) and maxabs(got, np.concatenate(cfg_bin)) < 1e-13
    lines = [x for x in open(os.path.join(str(tmp_path), "configs-phi.textstream")) if x[0] != "#"]
    assert len(lines) == got.size and np.allclose([float(x) for x in lines], got, rtol=1e-13, atol=0)
    assert "binary phi configuration stream" in open(os.path.join(str(tmp_path), "configs-phi.infoheader")).read()
    info = open(os.path.join(str(tmp_path), "info.dat")).read()
    assert "libdqmc_b200" in info


def test_reference_driver_with_gpu_shim_fermionic_measurements(tmp_path):
    """The reference's DetQMC driver on the GPU shim with turnoffFermionMeasurements = false: its time series of the
    fermionic observables (greenLocal, occDiffSq, pairPlusMax) must equal the oracle's measured sweeps for the same seed."""
    import os
    import subprocess
    from dqmc_oracle import SdwOracle, SdwParams
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "host", "_build", "detqmcsdw_gpu")
    if not os.path.exists(exe):
        pytest.skip("host/_build/detqmcsdw_gpu not built (needs the reference tree: make -C host)")
    therm, sweeps = 4, 4
    out = subprocess.run([exe, "L=4", "beta=2", "dtau=0.1", "s=10", "r=-1", "thermalization=%d" % therm,
                          "sweeps=%d" % sweeps, "rngSeed=1020304050", "simindex=0", "turnoffFermionMeasurements=0"],
                         cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    o = SdwOracle(SdwParams(L=4, m=20, s=10, r=-1.0))
    for _ in range(therm):
        o.sweep_thermalization()
    ref = {"greenLocal": [], "occDiffSq": [], "pairPlusMax": [], "greenK0": []}
    for _ in range(sweeps):
        ob = o.measured_sweep_fermionic()
        for name in ref:
            ref[name].append(ob[name])
    for name in ref:
        got = [float(x) for x in open(os.path.join(str(tmp_path), name + ".series")) if x.strip() and x[0] != "#"]
        assert len(got) == sweeps and np.allclose(got, ref[name], rtol=2e-5, atol=2e-6), name
    # the vector observables go through the reference's VectorObservableHandler (results-<name>.values)
    assert any("kOccX" in f for f in os.listdir(str(tmp_path))), sorted(os.listdir(str(tmp_path)))


# ---------------------------------------------------------------- DetHubbard (BASELINE configs C1, C5)
@pytest.mark.parametrize("name", ["hubbard_L4_U4_b4", "hubbard_L4_cb"])
def test_hubbard_vs_golden(name):
    from detqmc_b200 import DetHubbardBatch
    from helpers import hubbard_params_of
    g = load_golden(name)
    p = hubbard_params_of(g)
    b = DetHubbardBatch(p)
    assert np.array_equal(b.auxfield()[1:], g["aux0"])                 # same dSFMT stream, same draw order
    N = p.N
    eye = np.eye(N)
    for gc in (0, 1):
        assert maxabs(b.bmat_mult(0, eye, 9, 4, gc=gc), g["bmat_%d_9_4" % gc]) < 1e-12      # B(9,4) I
        assert maxabs(b.bmat_mult(1, eye, 9, 4, gc=gc), g["bmat_%d_9_4" % gc]) < 1e-12      # I B(9,4)
        A = np.random.default_rng(gc).standard_normal((N, N))
        assert relerr(b.bmat_mult(2, b.bmat_mult(0, A, 9, 4, gc=gc), 9, 4, gc=gc), A) < 1e-11   # B^-1 B A
        assert relerr(b.bmat_mult(3, b.bmat_mult(1, A, 9, 4, gc=gc), 9, 4, gc=gc), A) < 1e-11   # A B B^-1
        assert relerr(b.green(0, gc), g["green0_%d" % gc]) < TOL_G
        assert abs(b.logdet(0, gc) - np.log(g["sv0_%d" % gc]).sum()) < 1e-9 * max(1.0, abs(b.logdet(0, gc)))
    n = int(g["n_sweeps"])
    for sw in range(n):
        b.sweep()
        key = "aux_after_%d" % (sw + 1)
        if key in g.files:
            assert np.array_equal(b.auxfield()[1:], g[key])
            for gc in (0, 1):
                assert relerr(b.green(0, gc), g["green_after_%d_%d" % (sw + 1, gc)]) < 1e-9
    assert np.array_equal(b.rng_draw(8), g["rng_next"])                # the stream was consumed exactly as in the reference
    if p.mu == 0.0 and not p.checkerboard:
        assert abs(b.total_occupation() - 1.0) < 1e-10                 # half filling


@pytest.mark.parametrize("tag", ["c1", "cb"])
def test_hubbard_observables_vs_golden(tag):
    """DetHubbard::measure on the device (dqmc_sweep(ctx, 2) + dqmc_get_hubbard_observables) against the observables
    the reference measured itself over six sweeps (tools/make_golden.py hubbard_observables)."""
    import json
    from detqmc_b200 import DetHubbardBatch
    from dqmc_oracle import HubbardParams
    g = load_golden("hubbard_observables")
    d = json.loads(str(g["params_" + tag]))
    d.pop("N", None)
    b = DetHubbardBatch(HubbardParams(**d))
    for _ in range(4):
        b.sweepThermalization()
    for i in range(6):
        b.sweep(True)
        ob = b.observables()
        got = np.array([ob[k] for k in b.OBSERVABLES])
        assert np.allclose(got, g["scalars_" + tag][i], rtol=1e-9, atol=1e-10), (i, got, g["scalars_" + tag][i])
        assert maxabs(ob["spinzCorrelationFunction"], g["zcorr_" + tag][i]) < 1e-9
    assert np.array_equal(b.auxfield()[1:], g["aux_final_" + tag])


def _series(path):
    return [float(x) for x in open(path) if x.strip() and x[0] != "#"]


def test_reference_driver_with_hubbard_shim(tmp_path):
    """host/_build/detqmchubbard_gpu = the reference's DetQMC<Model, ModelParams> driver (compiled unmodified from the
    reference tree) on include/dethubbard_gpu.h: its time series must equal the observables the reference's own
    DetHubbard measured for the same seed (BASELINE config C1)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "host", "_build", "detqmchubbard_gpu")
    if not os.path.exists(exe):
        pytest.skip("host/_build/detqmchubbard_gpu not built (needs the reference tree: make -C host)")
    g = load_golden("hubbard_observables")
    out = subprocess.run([exe, "L=4", "U=4", "beta=4", "dtau=0.1", "s=10", "mu=0", "t=1", "thermalization=4", "sweeps=6",
                          "rngSeed=1020304050", "simindex=0"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    names = ("occupationUp", "occupationDown", "totalOccupation", "doubleOccupation", "localMoment", "kineticEnergy",
             "potentialEnergy", "totalEnergy")
    for i, name in enumerate(names):
        got = _series(os.path.join(str(tmp_path), name + ".series"))
        assert len(got) == 6 and np.allclose(got, g["scalars_c1"][:, i], rtol=2e-5, atol=2e-6), name
    assert "libdqmc_b200" in open(os.path.join(str(tmp_path), "info.dat")).read()
    assert any("spinzCorrelationFunction" in f for f in os.listdir(str(tmp_path)))


@pytest.mark.parametrize("model", ["sdw", "hubbard"])
def test_resume_continues_the_uninterrupted_run(tmp_path, model):
    """saveState / resume through the reference's driver (detqmc.h:264-356): a run stopped after 4 of 8 sweeps and
    resumed from its state file produces the same time series as an uninterrupted run -- the random numbers the
    replica had drawn ahead of consumption are part of the checkpoint (dqmc_rng_look_ahead / dqmc_rng_set_look_ahead),
    and so is the sweep counter that fixes the global-move schedule (dqmc_set_performed_sweeps)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "_build", "detqmcsdw_gpu" if model == "sdw" else "detqmchubbard_gpu")
    if not os.path.exists(exe):
        pytest.skip("host/_build not built (needs the reference tree: make -C host)")
    common = (["L=4", "beta=2", "dtau=0.1", "s=10", "r=-1", "globalUpdateInterval=3"] if model == "sdw"
              else ["L=4", "U=4", "beta=2", "dtau=0.1", "s=10", "mu=0.2"])
    common += ["thermalization=4", "rngSeed=1020304050", "simindex=0"]
    obs = "normMeanPhi" if model == "sdw" else "doubleOccupation"

    def run(d, sweeps):
        os.makedirs(d, exist_ok=True)
        out = subprocess.run([exe] + common + ["sweeps=%d" % sweeps], cwd=d, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        return out.stdout

    a = os.path.join(str(tmp_path), "straight")
    b = os.path.join(str(tmp_path), "resumed")
    run(a, 8)
    run(b, 4)
    assert "will resume simulation" in run(b, 8)
    sa, sb = _series(os.path.join(a, obs + ".series")), _series(os.path.join(b, obs + ".series"))
    assert len(sa) == 8 and len(sb) == 8
    assert np.allclose(sa, sb, rtol=0, atol=1e-9), (sa, sb)


def test_hubbard_full_size_properties():
    """BASELINE config C5 (L=20, U=8, beta=20, m=200, s=10) at full size: properties that need no oracle --
    half-filling identity <n> = 1 for every auxiliary-field configuration, wrapped vs recomputed G at the
    stabilisation points, G(0) after a sweep vs from scratch, batch == single replica."""
    from detqmc_b200 import DetHubbardBatch
    p = dict(L=20, m=200, s=10, dtau=0.1, U=8.0, mu=0.0, t=1.0)
    b = DetHubbardBatch(p, n_replicas=2, rng_indices=[1, 2])
    assert abs(b.total_occupation(0) - 1.0) < 1e-9 and abs(b.total_occupation(1) - 1.0) < 1e-9
    b.sweep()                                                            # full down-sweep
    # wrapping over s = 10 slices at U = 8 amplifies round-off by ~cond(B)^10 ~ 1e11: 1e-5 is the expected size
    assert np.all(b.green_consistency() < 1e-3)
    for rep in (0, 1):
        assert abs(b.total_occupation(rep) - 1.0) < 1e-9
        for gc in (0, 1):
            assert relerr(b.green(rep, gc), b.green_for_timeslice(0, rep=rep, gc=gc)) < 1e-8
    single = DetHubbardBatch(p, n_replicas=1, rng_indices=[2])
    single.sweep()
    assert np.array_equal(single.auxfield(0), b.auxfield(1))
    assert relerr(single.green(0, 1), b.green(1, 1)) < 1e-12


def test_o3_full_size_properties():
    """BASELINE config C4 (DetSDW O(3), L=14, beta=14, s=10: D = 784, m = 140) at full size through properties
    that need no oracle: B^-1 B = 1, wrap round trip, wrapped vs recomputed G at the stabilisation points,
    G(0) after a full sweep vs from scratch, log|det| consistency between setup and from-scratch."""
    from dqmc_oracle import SdwParams
    p = SdwParams(opdim=3, L=14, m=140, s=10, weakZflux=False)
    b = make_batch(p, n_replicas=1)
    D = b.D
    assert D == 784
    A = rand_cplx((D, D), 3)
    assert relerr(b.bmat_mult(2, b.bmat_mult(0, A, 17, 7), 17, 7), A) < 1e-10
    assert relerr(b.bmat_mult(3, b.bmat_mult(1, A, 17, 7), 17, 7), A) < 1e-10
    G0 = b.green(0)
    assert relerr(G0, b.green_for_timeslice(p.m)) < 1e-8
    b.wrap_down(p.m)
    b.wrap_up(p.m - 1)
    assert relerr(b.green(0), G0) < 1e-9
    b.sweepThermalization()                                         # global move + full down-sweep
    acc = b.control_data(0).lastAccRatioLocal_phi
    assert 0.05 < acc < 0.99
    assert np.all(b.green_consistency() < 1e-6)
    assert relerr(b.green(0), b.green_for_timeslice(0)) < 1e-7
