"""CPU tests: the NumPy oracle (oracle/dqmc_oracle.py) against the golden vectors that
tools/make_golden.py produced from the unmodified reference.  This is what pins the oracle."""
import numpy as np
import pytest

from dqmc_oracle import (SdwOracle, HubbardOracle, exchange_probability, replica_exchange_walk)
from dsfmt_oracle import RngOracle
from helpers import load_golden, sdw_params_of, hubbard_params_of, maxabs, relerr

SDW_CASES = ["sdw_o2_wolff_L4", "sdw_o2_wolffshift_L4", "sdw_o3_woodbury_L4", "sdw_o2_repeat2_L4",
             "sdw_o2_flux_L4", "sdw_o2_noflux_apbcxy_L4", "sdw_o3_L4", "sdw_o1_apbcx_L4",
             "sdw_o2_flux_L4_delay3_s7", "sdw_o2_flux_L6"]


def test_rng_streams():
    g = load_golden("rng_streams")
    for key in g.files:
        seed, idx = key[1:].split("_i")
        r = RngOracle(int(seed), int(idx))
        mine = np.array([r.rand01() for _ in range(len(g[key]))])
        assert np.array_equal(mine, g[key]), key      # bit-exact: integer recursion


def test_rng_seed_scramble_kat():
    # seed 1020304050, idx 1 -> 37767 (rngwrapper.cpp:43)
    assert RngOracle(1020304050, 1).my_seed == 37767


@pytest.mark.parametrize("name", SDW_CASES)
def test_sdw_setup_and_bmult(name):
    g = load_golden(name)
    o = SdwOracle(sdw_params_of(g))
    assert maxabs(o.phi, g["phi0"]) == 0.0
    assert maxabs(o.cosh_term[1:], g["cosh0"][1:]) < 1e-14
    assert maxabs(o.sinh_term[1:], g["sinh0"][1:]) < 1e-14
    assert maxabs(o.green[0], g["green0"]) < 1e-12
    assert abs(np.log(o.green_inv_sv[0]).sum() - np.log(g["sv0"]).sum()) < 1e-10
    A = g["A"]
    k2, k1 = [int(v) for v in g["chain"]]
    fns = [o.left_multiply_bmat, o.right_multiply_bmat, o.left_multiply_bmat_inv, o.right_multiply_bmat_inv]
    for op, f in enumerate(fns):
        ref1 = g["bmult_op%d_single" % op]
        refc = g["bmult_op%d_chain" % op]
        assert maxabs(f(0, A, 3, 2), ref1) < 1e-12 * np.abs(ref1).max()
        assert maxabs(f(0, A, k2, k1), refc) < 1e-12 * np.abs(refc).max()
    assert abs(o.phi_action() - float(g["phiAction0"])) < 1e-10
    assert abs(o.exchange_action() - float(g["exchangeAction0"])) < 1e-11
    gs, _ = o.green_for_timeslice(3)
    assert maxabs(gs, g["green_slice_3"]) < 1e-11


@pytest.mark.parametrize("name", SDW_CASES)
def test_sdw_sweeps(name):
    g = load_golden(name)
    o = SdwOracle(sdw_params_of(g))
    n = int(g["n_sweeps"])
    for sw in range(n):
        o.sweep_thermalization()
        assert o.last_acc_ratio == g["lastAccRatio"][sw], (name, sw)
        assert o.accepted_global_shifts == g["acceptedGlobalShifts"][sw]
        assert abs(o.phi_delta - g["phiDelta"][sw]) < 1e-14
        key = "phi_after_%d" % (sw + 1)
        if key in g.files:
            assert maxabs(o.phi[1:], g[key][1:]) < 1e-12
            assert maxabs(o.green[0], g["green_after_%d" % (sw + 1)]) < 1e-10
    assert maxabs(o.phi[1:], g["phi_final"][1:]) < 1e-12
    nxt = np.array([o.rng.rand01() for _ in range(8)])
    assert np.array_equal(nxt, g["rng_next"])          # identical RNG consumption


def test_sdw_trajectory_100_sweeps():
    """North-star requirement: identical accept/reject trajectory over the first 100 sweeps."""
    g = load_golden("sdw_o2_flux_L4_traj100")
    o = SdwOracle(sdw_params_of(g))
    for sw in range(100):
        o.sweep_thermalization()
        assert o.last_acc_ratio == g["lastAccRatio"][sw], sw
        assert o.accepted_global_shifts == g["acceptedGlobalShifts"][sw], sw
    assert abs(o.phi_delta - g["phiDelta"][-1]) < 1e-13
    assert maxabs(o.phi[1:], g["phi_final"][1:]) < 1e-11
    assert maxabs(o.green[0], g["green_final"]) < 1e-10
    nxt = np.array([o.rng.rand01() for _ in range(8)])
    assert np.array_equal(nxt, g["rng_next"])


def test_green_from_udv_kat():
    from dqmc_oracle import UdV, green_from_udv
    g = load_golden("sdw_o2_flux_L4")
    l = UdV(g["udv2_U"], g["udv2_d"], g["udv2_V"])
    r = UdV(g["udv1_U"], g["udv1_d"], g["udv1_V"])
    G, sv = green_from_udv(l, r)
    assert maxabs(G, g["green_from_udv_l2_r1"]) < 1e-11
    assert abs(np.log(sv).sum() - np.log(g["sv_from_udv_l2_r1"]).sum()) < 1e-9


@pytest.mark.parametrize("name", ["hubbard_L4_U4_b4", "hubbard_L4_cb"])
def test_hubbard(name):
    g = load_golden(name)
    p = hubbard_params_of(g)
    o = HubbardOracle(p)
    assert np.array_equal(o.aux[1:], g["aux0"])
    assert maxabs(o.proptmat, g["proptmat"]) < 1e-13
    for gc in (0, 1):
        assert maxabs(o.green[gc], g["green0_%d" % gc]) < 1e-10
        assert maxabs(o.compute_bmat(gc, 9, 4), g["bmat_%d_9_4" % gc]) < 1e-12
    n = int(g["n_sweeps"])
    for sw in range(n):
        o.sweep()
        key = "aux_after_%d" % (sw + 1)
        if key in g.files:
            assert np.array_equal(o.aux[1:], g[key])
            for gc in (0, 1):
                assert maxabs(o.green[gc], g["green_after_%d_%d" % (sw + 1, gc)]) < 1e-9
    nxt = np.array([o.rng.rand01() for _ in range(8)])
    assert np.array_equal(nxt, g["rng_next"])
    if p.mu == 0.0 and not p.checkerboard:
        # half filling: <n> = 1 for every auxiliary-field configuration
        occ = 2.0 - (np.trace(o.green[0]) + np.trace(o.green[1])) / p.N
        assert abs(occ - 1.0) < 1e-10


def test_exchange_probability_golden():
    rows = load_golden("exchange_probability")["rows"]
    for p1, a1, p2, a2, ref in rows:
        assert abs(exchange_probability(p1, a1, p2, a2) - ref) <= 1e-15 * max(1.0, ref)


def test_replica_exchange_walk_properties():
    """Parity of the ladder walk is unpinned (no MPI here); check its invariants instead."""
    gen = np.random.default_rng(3)
    P = 8
    ladder = np.linspace(-1.9, 0.4, P)
    par_proc = list(range(P))
    proc_par = list(range(P))
    rng = RngOracle(1, 1)
    for _ in range(50):
        actions = gen.uniform(10, 20, P)
        par_proc, proc_par, log = replica_exchange_walk(ladder, par_proc, proc_par, actions, rng.rand01)
        assert sorted(par_proc) == list(range(P)) and sorted(proc_par) == list(range(P))
        for cpi in range(P):
            assert proc_par[par_proc[cpi]] == cpi
        assert len(log) == P - 1
    # equal actions -> delta = 0 -> always accepted, no RNG draw
    draws0 = rng.draws
    replica_exchange_walk(ladder, par_proc, proc_par, np.ones(P), rng.rand01)
    assert rng.draws == draws0


def test_config_stream_order_vs_reference_bytes():
    """The restated configuration-stream order (oracle config_stream / config_stream_text) reproduces the bytes and
    lines written by the reference's own saveConfigurationStreamBinary / Text for the same fields."""
    import os
    from dqmc_oracle import config_stream, config_stream_text
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_streams.npz"))
    for tag in ("o2", "o3", "o2_L6"):
        phi = g[tag + "_phi"]
        assert config_stream(phi).tobytes() == g[tag + "_binary"].tobytes()
        assert config_stream_text(phi).encode() == g[tag + "_text"].tobytes()


def test_wolff_cluster_moves_vs_reference_record():
    """attemptWolffClusterUpdate / attemptWolffClusterShiftUpdate of the oracle against the record of the reference's
    own moves (tests/golden/wolff_moves.npz): fields, Green's function, statistics and the position of the
    random-number stream after six attempts, for O(1), O(2), O(3), two clusters per attempt, and the shift variant."""
    import json
    import os
    from dqmc_oracle import SdwParams
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wolff_moves.npz"))
    for tag in ("o2", "o2_shift", "o3_rep2", "o1"):
        d = json.loads(str(g[tag + "_pars"]))
        for key in ("N", "beta"):
            d.pop(key, None)
        o = SdwOracle(SdwParams(**d))
        shift = bool(g[tag + "_shift"])
        for it in range(g[tag + "_phi"].shape[0]):
            if shift:
                o.attempt_wolff_cluster_shift_update()
            else:
                o.attempt_wolff_cluster_update()
            st = g[tag + "_stats"][it]
            ws = o.wolff_stats
            assert [ws["attempted"], ws["accepted"], ws["attempted_shift"], ws["accepted_shift"], ws["added_size"]] == list(st)
            assert maxabs(o.phi[1:], g[tag + "_phi"][it][1:]) < 1e-13
            assert maxabs(o.green[0], g[tag + "_green"][it]) < 1e-10
        assert maxabs([o.rng.rand01() for _ in range(4)], g[tag + "_rng_next"]) == 0.0


def test_bosonic_observables_vs_reference_values():
    """bosonic_observables (oracle restatement and the host mirror's copy) against the values the reference's own
    measured sweeps produced for the same fields."""
    import os
    from dqmc_oracle import bosonic_observables
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bosonic_observables.npz"))
    for tag in ("o2", "o3", "o2_L6"):
        for it in range(g[tag + "_phi"].shape[0]):
            ob = bosonic_observables(g[tag + "_phi"][it], 0.1)
            got = [ob["normMeanPhi"], ob["associatedEnergy"], ob.get("phiRhoS_Gs", 0.0), ob.get("phiRhoS_Gc", 0.0)]
            assert np.allclose(got, g[tag + "_obs"][it], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("tag", ["o2", "o2_apbc", "o3", "o2_L6"])
def test_fermionic_observables_vs_reference_values(tag):
    """measure / finishMeasurements restated (shiftGreenSymmetric, greenK0, greenLocal, occDiffSq, k-space occupation,
    equal-time pairing correlations) against the values of the reference's own measured sweeps."""
    import json
    import os
    from dqmc_oracle import SdwParams
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fermion_observables.npz"))
    d = json.loads(str(g[tag + "_pars"]))
    for key in ("N", "beta"):
        d.pop(key, None)
    o = SdwOracle(SdwParams(**d))
    n = 3 if tag != "o2_L6" else 1                   # L = 6 costs a few seconds per sweep in pure Python
    for it in range(n):
        ob = o.measured_sweep_fermionic()
        sc = [ob["greenK0"], ob["greenLocal"], ob["occDiffSq"], ob["pairPlusMax"], ob["pairMinusMax"]]
        vec = np.concatenate([ob["kOccX"], ob["kOccY"], ob["pairPlus"], ob["pairMinus"]])
        assert np.allclose(sc, g[tag + "_scalars"][it], rtol=1e-9, atol=1e-11)
        assert np.allclose(vec, g[tag + "_vectors"][it], rtol=1e-9, atol=1e-11)
    if n == 3:
        assert maxabs(o.shift_green_symmetric(), g[tag + "_green_shifted"]) < 1e-10


def test_host_mirror_observables_equal_oracle():
    """The host mirror's copy of the bosonic observables (detqmc_b200/pt.py, used by the parallel-tempering loop and
    DetSDWBatch.bosonic_observables) against the reference's values and the oracle restatement; numToString formatting
    of the sub-directory names (tools.h:46-50)."""
    import os
    from detqmc_b200.pt import bosonic_observables as mirror_observables, num_to_string
    from dqmc_oracle import bosonic_observables
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bosonic_observables.npz"))
    for tag in ("o2", "o3", "o2_L6"):
        for it in range(g[tag + "_phi"].shape[0]):
            phi = g[tag + "_phi"][it]
            a, b = mirror_observables(phi, 0.1), bosonic_observables(phi, 0.1)
            got = [a["normMeanPhi"], a["associatedEnergy"], a["phiRhoS_Gs"], a["phiRhoS_Gc"]]
            assert np.allclose(got, g[tag + "_obs"][it], rtol=1e-12, atol=1e-13)
            assert a["normMeanPhi"] == b["normMeanPhi"] and a["associatedEnergy"] == b["associatedEnergy"]
    assert [num_to_string(v) for v in (-1.9, 0.4, -1.0, 1e-7, 123456789.0)] == ["-1.9", "0.4", "-1", "1e-07", "1.23457e+08"]


def test_oracle_full_size_setup_vs_reference():
    """BASELINE config C2 at full size (L = 8, beta = 8, D = 128): G and log|det| of the NumPy oracle after set-up
    against the summary the unmodified reference produced (tools/make_golden.py, BIG fixtures).  (The first sweep of
    C2 and C3 was checked the same way when the fixtures were made -- identical acceptance and fields -- but takes
    minutes in pure Python, too long for this suite.)"""
    g = load_golden("sdw_c2_L8_b8_traj100")
    o = SdwOracle(sdw_params_of(g))
    st = int(g["stride"])
    assert np.abs(o.green[0][::st, ::st] - g["green0_sub"]).max() < 1e-11 * float(g["green0_maxabs"])
    assert abs(np.trace(o.green[0]) - g["green0_trace"]) < 1e-10
    assert abs(np.log(o.green_inv_sv[0]).sum() - float(g["logdet0"])) < 1e-10 * abs(float(g["logdet0"]))
    assert maxabs(o.phi, g["phi0"]) == 0.0


def test_oracle_dense_hopping_vs_reference():
    """The oracle's dense hopping path (setupPropK / computeBmatSDW / sweepSimple / DetSDW<CB_NONE>) against the
    reference's own dense B matrices, simple sweeps and CB_NONE sweeps (tools/make_golden.py dense)."""
    import json
    from dqmc_oracle import SdwOracle, SdwParams
    g = load_golden("sdw_dense_hopping")

    def params(tag, **over):
        d = json.loads(str(g["params_" + tag]))
        for k in ("N", "beta"):
            d.pop(k, None)
        d.update(over)
        return SdwParams(**d)

    for tag in ("flux", "noflux_apbcx"):
        od = SdwOracle(params(tag, checkerboard=False))
        eye = np.eye(od.sz, dtype=np.complex128)
        for k2, k1 in ((7, 3), (20, 19)):
            assert relerr(od.left_multiply_bmat(0, eye, k2, k1), g["bmat_%s_%d_%d" % (tag, k2, k1)]) < 1e-12
        o = SdwOracle(params(tag))
        for _ in range(2):
            o.sweep_simple_thermalization()
        assert maxabs(o.phi[1:], g["simple_phi_" + tag]) < 1e-12
        assert relerr(o.green[0], g["simple_green_" + tag]) < 1e-9
    o = SdwOracle(params("cbnone"))
    assert relerr(o.green[0], g["cbnone_green0"]) < 1e-10
    for _ in range(4):
        o.sweep_thermalization()
    assert maxabs(o.phi[1:], g["cbnone_phi"]) < 1e-12
    assert relerr(o.green[0], g["cbnone_green"]) < 1e-9
