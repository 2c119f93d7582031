#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref
(oracle/Makefile).  Run in the CPU container where /root/reference exists:

    make -C oracle && python tools/make_golden.py

The fixtures are small known-answer vectors for every row of SURVEY.md section 8(a) that the
reference can produce deterministically from a seed: RNG stream, initial fields, checkerboard
B-multiplies, G / singular values after setup, G from two UdVs, fields / step sizes / acceptance
after thermalisation sweeps (incl. a 100-sweep trajectory), exchange probability, and the
DetHubbard path.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_bindings as rb                                   # noqa: E402
from dqmc_oracle import SdwParams, HubbardParams            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def pars_json(p):
    return json.dumps({k: v for k, v in vars(p).items()})


def rng_fixture():
    data = {}
    for seed, idx in ((1020304050, 1), (1020304050, 2), (5, 7), (4242, 64)):
        data["s%d_i%d" % (seed, idx)] = rb.RefRng(seed, idx).draw(2000)
    np.savez_compressed(os.path.join(OUT, "rng_streams.npz"), **data)


def sdw_fixture(name, sweeps, kw, chain=(7, 4)):
    p = SdwParams(**kw)
    r = rb.RefSdw(p)
    d = {"params": pars_json(p)}
    d["phi0"] = r.phi()
    d["green0"] = r.green()
    d["sv0"] = r.sv()
    c, s = r.tables()
    d["cosh0"], d["sinh0"] = c, s
    gen = np.random.default_rng(12345)
    A = gen.standard_normal((r.D, r.D)) + 1j * gen.standard_normal((r.D, r.D))
    d["A"] = A
    for op in range(4):
        d["bmult_op%d_single" % op] = r.bmult(op, A, 3, 2)
        d["bmult_op%d_chain" % op] = r.bmult(op, A, chain[0], chain[1])
    d["chain"] = np.array(chain)
    # G = [1 + M_r M_l]^-1 from two stored UdVs (numerical known-answer for greenFromUdV)
    for l in (1, 2):
        U, dd, V = r.udv(l)
        d["udv%d_U" % l], d["udv%d_d" % l], d["udv%d_V" % l] = U, dd, V
    g12, sv12 = r.green_from_storage(2, 1)
    d["green_from_udv_l2_r1"], d["sv_from_udv_l2_r1"] = g12, sv12
    d["green_slice_3"] = r.green_for_timeslice(3)
    d["green_slice_%d" % p.s] = r.green_for_timeslice(p.s)
    sc = r.scalars()
    d["phiAction0"], d["exchangeAction0"] = sc["phiAction"], sc["exchangeAction"]
    acc, gsa, pdel = [], [], []
    for sw in range(sweeps):
        r.sweep(therm=True)
        sc = r.scalars()
        acc.append(sc["lastAccRatio"])
        gsa.append(sc["acceptedGlobalShifts"])
        pdel.append(sc["phiDelta"])
        if sw + 1 in (1, 2, 6):
            d["phi_after_%d" % (sw + 1)] = r.phi()
            d["green_after_%d" % (sw + 1)] = r.green()
    d["phi_final"], d["green_final"] = r.phi(), r.green()
    d["lastAccRatio"] = np.array(acc)
    d["acceptedGlobalShifts"] = np.array(gsa)
    d["phiDelta"] = np.array(pdel)
    d["n_sweeps"] = sweeps
    sc = r.scalars()
    d["phiAction_final"], d["exchangeAction_final"] = sc["phiAction"], sc["exchangeAction"]
    d["rng_next"] = r.rng_draw(8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "done; acc", acc[-1], "global shifts", gsa[-1], "phiDelta", pdel[-1])


def hubbard_fixture(name, sweeps, kw):
    p = HubbardParams(**kw)
    r = rb.RefHubbard(p)
    d = {"params": pars_json(p)}
    d["aux0"] = r.aux()[1:]
    d["proptmat"] = r.proptmat()
    for gc in (0, 1):
        d["green0_%d" % gc] = r.green(gc)
        d["sv0_%d" % gc] = r.sv(gc)
        d["bmat_%d_9_4" % gc] = r.bmat(gc, 9, 4)
    for sw in range(sweeps):
        r.sweep(False)
        if sw + 1 in (1, 2, sweeps):
            d["aux_after_%d" % (sw + 1)] = r.aux()[1:]
            for gc in (0, 1):
                d["green_after_%d_%d" % (sw + 1, gc)] = r.green(gc)
    d["n_sweeps"] = sweeps
    d["rng_next"] = r.rng_draw(8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "done")


def hubbard_observables_fixture():
    """DetHubbard::measure / finishMeasurements (dethubbard.cpp:511-612) of the reference over a few measured sweeps
    after a short thermalisation, for BASELINE config C1 and an off-half-filling checkerboard case."""
    d = {}
    for tag, kw in (("c1", dict()), ("cb", dict(checkerboard=True, U=6.0, mu=0.3, m=24, s=5))):
        p = HubbardParams(**kw)
        r = rb.RefHubbard(p)
        d["params_" + tag] = pars_json(p)
        for _ in range(4):
            r.sweep(True)
        sc, zc = [], []
        for _ in range(6):
            a, b = r.measured_sweep()
            sc.append(a)
            zc.append(b)
        d["scalars_" + tag], d["zcorr_" + tag] = np.array(sc), np.array(zc)
        d["aux_final_" + tag] = r.aux()[1:]
    np.savez_compressed(os.path.join(OUT, "hubbard_observables.npz"), **d)
    print("hubbard_observables done")


def dense_fixture():
    """Dense hopping path of the reference: computeBmatSDW, sweepSimple (which builds its B
    matrices with the dense hopping exponential whatever the checkerboard setting, detsdwopdim.cpp:4366-4420), and
    stabilised sweeps of DetSDW<CB_NONE, 2> (checkerboard = false)."""
    d = {}
    for tag, kw in (("flux", dict()), ("noflux_apbcx", dict(weakZflux=False, bc=1, rngIndex=4))):
        p = SdwParams(**kw)
        r = rb.RefSdw(p)
        d["params_" + tag] = pars_json(p)
        d["bmat_%s_7_3" % tag] = r.dense_bmat(7, 3)
        d["bmat_%s_20_19" % tag] = r.dense_bmat(20, 19)
        for sw in range(2):
            r.sweep_simple(True)
        d["simple_phi_" + tag] = r.phi()[1:]
        d["simple_green_" + tag] = r.green()
        d["simple_rng_next_" + tag] = r.rng_draw(4)
    p = SdwParams(checkerboard=False, rngIndex=6)
    r = rb.RefSdw(p)
    d["params_cbnone"] = pars_json(p)
    d["cbnone_green0"] = r.green()
    d["cbnone_logdet0"] = float(np.log(r.sv()).sum())
    for sw in range(4):
        r.sweep(True)
    d["cbnone_phi"] = r.phi()[1:]
    d["cbnone_green"] = r.green()
    d["cbnone_rng_next"] = r.rng_draw(4)
    np.savez_compressed(os.path.join(OUT, "sdw_dense_hopping.npz"), **d)
    print("sdw_dense_hopping done")


def pt_reference_fixture():
    """Replica-exchange trajectory of the reference's own DetQMCPT driver (detqmcpt.h, unmodified) run with one thread
    per ladder process on the thread-backed boost::mpi stand-in (oracle/_ref/ref_pt, oracle/ref_pt_harness.cpp):
    per-parameter time series, exchange acceptance and diffusion statistics."""
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_pt")
    values = [-1.9, -1.5, -1.1, -0.7, -0.3, 0.1]
    L, beta, s, therm, sweeps, xint = 4, 2.0, 10, 4, 8, 1
    d = {"values": np.array(values), "L": L, "beta": beta, "s": s, "thermalization": therm, "sweeps": sweeps,
         "exchangeInterval": xint}
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
        subprocess.run([exe, str(L), str(beta), str(s), str(therm), str(sweeps), str(xint)] + [repr(v) for v in values],
                       cwd=tmp, check=True, stdout=subprocess.DEVNULL, env=env)

        def table(name):
            return np.array([[float(x) for x in l.split()] for l in open(os.path.join(tmp, name)) if l[0] != "#" and l.strip()])
        d["acceptance"] = table("exchange-acceptance.values")[:, 1]
        d["diffusion"] = table("exchange-diffusion.values")[:, 1]
        subdirs = sorted(x for x in os.listdir(tmp) if x.startswith("p") and os.path.isdir(os.path.join(tmp, x)))
        assert len(subdirs) == len(values), subdirs
        for obs in ("normMeanPhi", "associatedEnergy", "phiRhoS_Gs", "phiRhoS_Gc"):
            rows = []
            for cpi in range(len(values)):
                sub = [x for x in subdirs if x.startswith("p%d_" % cpi)][0]
                rows.append([float(l) for l in open(os.path.join(tmp, sub, obs + ".series")) if l[0] != "#" and l.strip()])
            d["series_" + obs] = np.array(rows)
        d["subdirs"] = np.array(subdirs)
    np.savez_compressed(os.path.join(OUT, "pt_reference.npz"), **d)
    print("pt_reference done; acceptance", d["acceptance"])


def exchange_fixture():
    gen = np.random.default_rng(7)
    rows = []
    for _ in range(64):
        p1, p2 = gen.uniform(-2, 1, 2)
        a1, a2 = gen.uniform(0, 50, 2)
        rows.append((p1, a1, p2, a2, rb.exchange_probability(p1, a1, p2, a2)))
    np.savez_compressed(os.path.join(OUT, "exchange_probability.npz"), rows=np.array(rows))


def wolff_fixture():
    """Wolff cluster moves: attemptWolffClusterUpdate / attemptWolffClusterShiftUpdate
    (detsdwopdim.cpp:3487-3562, 3647-3883) called directly on a freshly set-up replica, fields / statistics /
    Green's function recorded after every attempt."""
    d = {}
    cases = (("o2", dict(wolffClusterUpdate=True), False),
             ("o2_shift", dict(wolffClusterShiftUpdate=True, globalShift=False, rngIndex=2), True),
             ("o3_rep2", dict(wolffClusterUpdate=True, opdim=3, weakZflux=False, repeatWolffPerSweep=2), False),
             ("o1", dict(wolffClusterUpdate=True, opdim=1, weakZflux=False, rngIndex=3), False))
    for tag, kw, shift in cases:
        p = SdwParams(**kw)
        rep = rb.RefSdw(p)
        n = 6
        phis, stats, greens = [], [], []
        for _ in range(n):
            stats.append(rep.attempt_wolff(shift))
            phis.append(rep.phi())
            greens.append(rep.green())
        d[tag + "_pars"] = pars_json(p)
        d[tag + "_shift"] = int(shift)
        d[tag + "_phi"] = np.array(phis)
        d[tag + "_stats"] = np.array(stats)
        d[tag + "_green"] = np.array(greens)
        d[tag + "_rng_next"] = rep.rng_draw(4)
    np.savez_compressed(os.path.join(OUT, "wolff_moves.npz"), **d)
    print("wolff_moves done")


def observables_fixture():
    """Bosonic observables of a measured sweep (sweep(true) with turnoffFermionMeasurements, detsdwopdim.cpp:441-560,
    903-918): fields after the sweep and normMeanPhi, associatedEnergy, phiRhoS_Gs, phiRhoS_Gc."""
    d = {}
    for tag, kw in (("o2", dict(rngIndex=12)), ("o3", dict(opdim=3, weakZflux=False, rngIndex=13)),
                    ("o2_L6", dict(L=6, m=10, s=5, rngIndex=14))):
        p = SdwParams(**kw)
        rep = rb.RefSdw(p)
        obs, phis = [], []
        for _ in range(3):
            o = rep.measured_sweep()
            obs.append([o["normMeanPhi"], o["associatedEnergy"], o["phiRhoS_Gs"], o["phiRhoS_Gc"]])
            phis.append(rep.phi())
        d[tag + "_pars"] = pars_json(p)
        d[tag + "_obs"] = np.array(obs)
        d[tag + "_phi"] = np.array(phis)
    np.savez_compressed(os.path.join(OUT, "bosonic_observables.npz"), **d)
    print("bosonic_observables done")


def fermion_fixture():
    """Fermionic observables of measured sweeps (sweep(true) with turnoffFermionMeasurements = false,
    detsdwopdim.cpp:508-1000): greenK0, greenLocal, occDiffSq, pairPlusMax, pairMinusMax, kOccX, kOccY, pairPlus,
    pairMinus after each of three sweeps, plus shiftGreenSymmetric of the final G."""
    d = {}
    for tag, kw in (("o2", dict(rngIndex=16)), ("o2_apbc", dict(weakZflux=False, bc=3, rngIndex=17)),
                    ("o3", dict(opdim=3, weakZflux=False, rngIndex=18)), ("o2_L6", dict(L=6, m=10, s=5, rngIndex=19))):
        p = SdwParams(fermionMeasurements=True, **kw)
        rep = rb.RefSdw(p)
        sc, vec = [], []
        for _ in range(3):
            o = rep.measured_sweep_fermionic()
            sc.append([o["greenK0"], o["greenLocal"], o["occDiffSq"], o["pairPlusMax"], o["pairMinusMax"]])
            vec.append(np.concatenate([o["kOccX"], o["kOccY"], o["pairPlus"], o["pairMinus"]]))
        d[tag + "_pars"] = pars_json(p)
        d[tag + "_scalars"] = np.array(sc)
        d[tag + "_vectors"] = np.array(vec)
        d[tag + "_phi"] = rep.phi()
        d[tag + "_green"] = rep.green()
        d[tag + "_green_shifted"] = rep.shift_green_symmetric()
    np.savez_compressed(os.path.join(OUT, "fermion_observables.npz"), **d)
    print("fermion_observables done")


def config_stream_fixture():
    """Configuration streams: fields after two thermalisation sweeps and the bytes / lines the
    reference's own writers append for them (DetSDW::saveConfigurationStreamBinary / Text,
    detsdwopdim.cpp:4943-5036)."""
    import tempfile
    d = {}
    for tag, kw in (("o2", dict()), ("o3", dict(opdim=3, weakZflux=False)), ("o2_L6", dict(L=6, m=10, s=5, rngIndex=4))):
        p = SdwParams(**kw)
        rep = rb.RefSdw(p)
        for _ in range(2):
            rep.sweep(therm=True)
        tmp = tempfile.mkdtemp()
        rep.save_config_stream(tmp, True)
        rep.save_config_stream(tmp, False)
        d[tag + "_pars"] = pars_json(p)
        d[tag + "_phi"] = rep.phi()
        d[tag + "_binary"] = np.fromfile(os.path.join(tmp, "configs-phi.binarystream"), dtype=np.uint8)
        d[tag + "_text"] = np.frombuffer(open(os.path.join(tmp, "configs-phi.textstream"), "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "config_streams.npz"), **d)
    print("config_streams done")


# ------------------------------------------------------------------------------------------------
# Full-size fixtures (BASELINE configs C2-C5).  The matrices are too large to commit whole: a strided
# sub-sample of G plus three full-matrix functionals (trace, Frobenius norm, a fixed random bilinear
# probe u^H G v) are stored; the fields (phi / aux) are stored whole so that "identical accept /
# reject decisions" is checked site by site.
def probe_vectors(D):
    gen = np.random.default_rng(20261018)
    u = gen.standard_normal(D) + 1j * gen.standard_normal(D)
    v = gen.standard_normal(D) + 1j * gen.standard_normal(D)
    return u / np.linalg.norm(u), v / np.linalg.norm(v)


def summarise_matrix(G, stride):
    u, v = probe_vectors(G.shape[0])
    return dict(sub=np.ascontiguousarray(G[::stride, ::stride]), trace=np.trace(G), fro=np.linalg.norm(G),
                probe=np.vdot(u, G @ v), maxabs=np.abs(G).max())


def put(d, key, G, stride):
    for k, v in summarise_matrix(G, stride).items():
        d[key + "_" + k] = v


def big_sdw_fixture(name, sweeps, kw, stride, keep=(1,), selfdev=True):
    import time
    p = SdwParams(**kw)
    t0 = time.time()
    r = rb.RefSdw(p)
    d = {"params": pars_json(p), "stride": stride}
    d["phi0"] = r.phi()
    put(d, "green0", r.green(), stride)
    d["logdet0"] = np.log(r.sv()).sum()
    put(d, "green_slice_m", r.green_for_timeslice(p.m), stride)
    print(name, "setup %.0f s" % (time.time() - t0), flush=True)
    acc, gsa, pdel = [], [], []
    for sw in range(sweeps):
        r.sweep(therm=True)
        sc = r.scalars()
        acc.append(sc["lastAccRatio"]); gsa.append(sc["acceptedGlobalShifts"]); pdel.append(sc["phiDelta"])
        if sw + 1 in keep or sw + 1 == sweeps:
            d["phi_after_%d" % (sw + 1)] = r.phi()
            G = r.green()
            put(d, "green_after_%d" % (sw + 1), G, stride)
            # the reference's own wrapped-vs-recomputed deviation at this point (bounds what parity can mean).
            # computeGreenFromScratch must NOT be called on the replica that keeps sweeping: it overwrites state
            # the next sweep reads (a first version of this fixture diverged from a plain run in sweep 2), so a
            # second replica is set to the same fields and recomputes G there.
            if selfdev:
                k = int(sc["currentTimeslice"]) or p.m          # G(0) = G(beta); computeGreenFromScratch(0) is not defined
                r2 = rb.RefSdw(p)
                r2.set_phi(d["phi_after_%d" % (sw + 1)])
                Gs = r2.green_for_timeslice(k)
                d["ref_selfdev_after_%d" % (sw + 1)] = np.abs(G - Gs).max() / np.abs(Gs).max()
                del r2
        print(name, "sweep", sw + 1, "%.0f s" % (time.time() - t0), "acc", acc[-1], flush=True)
    d["lastAccRatio"], d["acceptedGlobalShifts"], d["phiDelta"] = np.array(acc), np.array(gsa), np.array(pdel)
    d["n_sweeps"] = sweeps
    sc = r.scalars()
    d["phiAction_final"], d["exchangeAction_final"] = sc["phiAction"], sc["exchangeAction"]
    d["rng_next"] = r.rng_draw(8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "done", flush=True)


def big_hubbard_fixture(name, sweeps, kw, stride):
    import time
    p = HubbardParams(**kw)
    t0 = time.time()
    r = rb.RefHubbard(p)
    d = {"params": pars_json(p), "stride": stride}
    d["aux0"] = r.aux()[1:].astype(np.int8)
    for gc in (0, 1):
        put(d, "green0_%d" % gc, r.green(gc), stride)
        d["logdet0_%d" % gc] = np.log(r.sv(gc)).sum()
    print(name, "setup %.0f s" % (time.time() - t0), flush=True)
    for sw in range(sweeps):
        r.sweep(False)
        d["aux_after_%d" % (sw + 1)] = r.aux()[1:].astype(np.int8)
        for gc in (0, 1):
            put(d, "green_after_%d_%d" % (sw + 1, gc), r.green(gc), stride)
        print(name, "sweep", sw + 1, "%.0f s" % (time.time() - t0), flush=True)
    d["n_sweeps"] = sweeps
    d["rng_next"] = r.rng_draw(8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "done", flush=True)


BIG = {
    # C2: 100-sweep trajectory at L = 8, beta = 8
    "sdw_c2_L8_b8_traj100": lambda: big_sdw_fixture("sdw_c2_L8_b8_traj100", 100, dict(L=8, m=80, s=10), 2,
                                                    keep=(1, 2, 10, 50)),
    # C3: one replica of the L = 12, beta = 10 ladder, 6 sweeps (3 down, 3 up)
    "sdw_c3_L12_b10": lambda: big_sdw_fixture("sdw_c3_L12_b10", 6, dict(L=12, m=100, s=10), 3, keep=(1, 2)),
    # C4: O(3), L = 14, beta = 14 (D = 784)
    "sdw_c4_o3_L14_b14": lambda: big_sdw_fixture("sdw_c4_o3_L14_b14", 2, dict(opdim=3, L=14, m=140, s=10,
                                                                                weakZflux=False), 4, keep=(1,), selfdev=False),
    # C5: DetHubbard L = 20, U = 8, beta = 20
    "hubbard_c5_L20_U8_b20": lambda: big_hubbard_fixture("hubbard_c5_L20_U8_b20", 2,
                                                         dict(L=20, m=200, s=10, U=8.0, mu=0.0, t=1.0), 4),
}


if __name__ == "__main__":
    assert rb.available(), "build oracle/_ref first: make -C oracle"

    def wolff_all():
        wolff_fixture()
        sdw_fixture("sdw_o2_wolff_L4", 6, dict(wolffClusterUpdate=True, globalUpdateInterval=2, rngIndex=6))
        sdw_fixture("sdw_o2_wolffshift_L4", 6, dict(wolffClusterShiftUpdate=True, globalShift=False, globalUpdateInterval=2,
                                                    rngIndex=7))

    if len(sys.argv) > 1 and sys.argv[1] in BIG:                     # python tools/make_golden.py <big fixture name>
        BIG[sys.argv[1]]()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "config_streams":
        config_stream_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "observables":
        observables_fixture()
        fermion_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dense":
        dense_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pt_reference":
        pt_reference_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "hubbard_observables":
        hubbard_observables_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "woodbury":
        # (no fixture for updateMethod=iterative: the reference's updateInSlice_iterative corrupts the heap in this
        # build -- "double free or corruption" at teardown -- although its fields equal woodbury's sweep by sweep)
        sdw_fixture("sdw_o3_woodbury_L4", 4, dict(updateMethod=1, opdim=3, weakZflux=False, rngIndex=9))
        sdw_fixture("sdw_o2_repeat2_L4", 4, dict(repeatUpdateInSlice=2, rngIndex=10))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "wolff":
        wolff_all()
        sys.exit(0)
    rng_fixture()
    exchange_fixture()
    sdw_fixture("sdw_o2_flux_L4", 6, dict())
    sdw_fixture("sdw_o2_noflux_apbcxy_L4", 6, dict(weakZflux=False, bc=3))
    sdw_fixture("sdw_o3_L4", 6, dict(opdim=3, weakZflux=False))
    sdw_fixture("sdw_o1_apbcx_L4", 6, dict(opdim=1, weakZflux=False, bc=1))
    sdw_fixture("sdw_o2_flux_L4_delay3_s7", 4, dict(delaySteps=3, s=7, m=24, rngIndex=3, globalUpdateInterval=2))
    sdw_fixture("sdw_o2_flux_L4_traj100", 100, dict(rngIndex=2))
    sdw_fixture("sdw_o2_flux_L6", 2, dict(L=6, m=30, rngIndex=5), chain=(20, 10))
    hubbard_fixture("hubbard_L4_U4_b4", 6, dict())                      # BASELINE config C1
    hubbard_fixture("hubbard_L4_cb", 4, dict(checkerboard=True, U=6.0, mu=0.3, m=24, s=5))
    hubbard_observables_fixture()
    dense_fixture()
    pt_reference_fixture()
    config_stream_fixture()
    observables_fixture()
    fermion_fixture()
    wolff_all()
    sdw_fixture("sdw_o2_repeat2_L4", 4, dict(repeatUpdateInSlice=2, rngIndex=10))
    sdw_fixture("sdw_o3_woodbury_L4", 4, dict(updateMethod=1, opdim=3, weakZflux=False, rngIndex=9))
