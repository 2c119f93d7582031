// dethubbard_gpu.h -- C++ host shim: the reference's Model duck-type for DetHubbard, served by libdqmc_b200.so.
//
// `DetHubbardGpu` is a drop-in for `DetHubbard` (dethubbard.h:52-368, `class DetHubbard : DetModelGC<2, num, false>`)
// as the template argument of the reference's driver `DetQMC<Model, ModelParams>` (detqmc.h:56-159): same member
// names, same argument meaning, same observable names (dethubbard.cpp:88-98), errors surface as the reference's
// GeneralError.  It compiles against the reference's own headers (reference src/ and this directory on the include
// path, link libdqmc_b200.so); nothing of the reference is copied here.  The sweep -- dense B-matrix products,
// wraps, UdV chains, Green's functions, single-flip updates with rank-1 updates of both spin components, and
// DetHubbard::measure after every slice -- runs on the GPU through include/dqmc_gpu.h (model = DQMC_MODEL_HUBBARD).
//
// Differences a maintainer has to know:
//   * greenUpdate = simple (sweepSimple, dethubbard.cpp:936-955) is refused with a GeneralError: the accelerated
//     path is the stabilised sweep, which every configuration of the reference uses by default.
//   * random numbers: the replica consumes the driver's RngWrapper through a pre-drawn FIFO (dqmc_rng_set_source);
//     saveContents stores the values drawn ahead but not yet consumed, loadContents puts them back in front of the
//     stream, so that a resumed run continues with exactly the numbers an uninterrupted run would have used.
#ifndef DETHUBBARD_GPU_H_
#define DETHUBBARD_GPU_H_

#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "dqmc_gpu.h"

// reference headers (crstnbr/detqmc, src/)
#include "detmodel.h"
#include "detmodelloggingparams.h"
#include "dethubbardparams.h"
#include "exceptions.h"
#include "metadata.h"
#include "observable.h"
#include "rngwrapper.h"

class DetHubbardGpu : public DetModel {
public:
    typedef ModelParams<DetHubbard> Params;

    DetHubbardGpu(RngWrapper& rng_, const Params& pars_, int device = 0)
        : rng(rng_), pars(pars_), ctx(nullptr), N(pars_.L * pars_.L), occUp(0), occDn(0), occTotal(0), occDouble(0),
          localMoment(0), eKinetic(0), ePotential(0), eTotal(0), performedSweeps(0) {
        if (pars.d != 2) throw_GeneralError("DetHubbardGpu: the GPU path implements the square lattice (d = 2)");
        if (pars.bc != "pbc") throw_GeneralError("DetHubbardGpu: periodic boundary conditions only");
        zcorr.zeros(N);
        dqmc_params p;
        std::memset(&p, 0, sizeof p);
        p.model = DQMC_MODEL_HUBBARD;
        p.opdim = 1;
        p.L = (int32_t)pars.L;
        p.m = (int32_t)pars.m;
        p.s = (int32_t)pars.s;
        p.delaySteps = 1;
        p.checkerboard = pars.checkerboard ? 1 : 0;
        p.dtau = pars.dtau;
        p.t = pars.t; p.U = pars.U; p.mu = pars.mu;
        int rc = dqmc_create(&p, 1, device, &ctx);
        if (rc != DQMC_OK) {
            const std::string msg = ctx ? dqmc_last_error(ctx) : "no context";
            if (ctx) dqmc_destroy(ctx);
            ctx = nullptr;
            throw_GeneralError("dqmc_create failed: " + msg);
        }
        check(dqmc_rng_set_source(ctx, 0, &DetHubbardGpu::fillFromRng, this), "dqmc_rng_set_source");
        check(dqmc_init_random_fields(ctx, 0), "dqmc_init_random_fields");   // setupRandomAuxfield, dethubbard.cpp:741-751
        check(dqmc_setup_storage(ctx), "dqmc_setup_storage");                // setupUdVStorage_and_calculateGreen
    }
    virtual ~DetHubbardGpu() { if (ctx) dqmc_destroy(ctx); }

    virtual uint32_t getSystemN() const { return N; }

    // dethubbard.cpp:118-139 (same keys)
    virtual MetadataMap prepareModelMetadataMap() const {
        MetadataMap meta;
        meta["model"] = "hubbard";
        meta["checkerboard"] = pars.checkerboard ? "true" : "false";
        meta["t"] = numToString(pars.t);
        meta["U"] = numToString(pars.U);
        meta["mu"] = numToString(pars.mu);
        meta["L"] = numToString(pars.L);
        meta["d"] = numToString(pars.d);
        meta["N"] = numToString(N);
        meta["beta"] = numToString(pars.beta);
        meta["m"] = numToString(pars.m);
        meta["dtau"] = numToString(pars.dtau);
        meta["s"] = numToString(pars.s);
        meta["alpha"] = numToString(std::acosh(std::exp(pars.dtau * pars.U * 0.5)));
        meta["backend"] = "libdqmc_b200 (sm_100a)";
        return meta;
    }

    virtual void sweep(bool takeMeasurements) {
        check(dqmc_sweep(ctx, takeMeasurements ? 2 : 0), "dqmc_sweep");
        ++performedSweeps;
        if (takeMeasurements) fetchObservables();
    }
    virtual void sweepThermalization() {
        check(dqmc_sweep(ctx, 0), "dqmc_sweep");
        ++performedSweeps;
    }
    virtual void sweepSimple(bool) { throw_GeneralError("DetHubbardGpu: greenUpdate = simple is not part of the accelerated path"); }
    virtual void sweepSimpleThermalization() { sweepSimple(false); }

    // the reference's observable list (dethubbard.cpp:88-98): same names, same short names, same order
    virtual std::vector<ScalarObservable> getScalarObservables() {
        std::vector<ScalarObservable> obs;
        obs.push_back(ScalarObservable(std::cref(occUp), "occupationUp", "nUp"));
        obs.push_back(ScalarObservable(std::cref(occDn), "occupationDown", "nDown"));
        obs.push_back(ScalarObservable(std::cref(occTotal), "totalOccupation", "n"));
        obs.push_back(ScalarObservable(std::cref(occDouble), "doubleOccupation", "n2"));
        obs.push_back(ScalarObservable(std::cref(localMoment), "localMoment", "m^2"));
        obs.push_back(ScalarObservable(std::cref(eKinetic), "kineticEnergy", "e_t"));
        obs.push_back(ScalarObservable(std::cref(ePotential), "potentialEnergy", "e_U"));
        obs.push_back(ScalarObservable(std::cref(eTotal), "totalEnergy", "e"));
        return obs;
    }
    virtual std::vector<VectorObservable> getVectorObservables() {
        std::vector<VectorObservable> obs;
        obs.push_back(VectorObservable(std::cref(zcorr), N, "spinzCorrelationFunction", "zcorr"));
        return obs;
    }
    virtual std::vector<KeyValueObservable> getKeyValueObservables() { return std::vector<KeyValueObservable>(); }

    // DetHubbard does not write system configurations (dethubbard.h:83-102): same behaviour
    void saveConfigurationStreamText(const std::string& = ".") {
        throw_GeneralError("DetHubbardGpu::saveConfigurationStreamText not implemented");
    }
    void saveConfigurationStreamBinary(const std::string& = ".") {
        throw_GeneralError("DetHubbardGpu::saveConfigurationStreamBinary not implemented");
    }
    void saveConfigurationStreamTextHeader(const std::string&, const std::string& = ".") {
        throw_GeneralError("DetHubbardGpu::saveConfigurationStreamTextHeader not implemented");
    }
    void saveConfigurationStreamBinaryHeaderfile(const std::string&, const std::string& = ".") {
        throw_GeneralError("DetHubbardGpu::saveConfigurationStreamBinaryHeaderfile not implemented");
    }

    // checkpointing (dethubbard.h:340-366): the auxiliary field, the sweep counter that fixes the direction of the
    // next sweep, and the random numbers drawn ahead of consumption; G and the UdV storage are rebuilt
    template <class Archive>
    void saveContents(Archive& ar) {
        std::vector<int32_t> aux(size_t(pars.m + 1) * N);
        check(dqmc_download_fields(ctx, 0, aux.data()), "dqmc_download_fields");
        std::vector<double> ahead = lookAhead();
        ar & aux & ahead & performedSweeps;
    }
    template <class Archive>
    void loadContents(Archive& ar) {
        std::vector<int32_t> aux;
        std::vector<double> ahead;
        ar & aux & ahead & performedSweeps;
        if (aux.size() != size_t(pars.m + 1) * N) throw_GeneralError("DetHubbardGpu: state does not match the parameters");
        check(dqmc_upload_fields(ctx, 0, aux.data()), "dqmc_upload_fields");
        check(dqmc_rng_set_look_ahead(ctx, 0, ahead.data(), ahead.size()), "dqmc_rng_set_look_ahead");
        check(dqmc_setup_storage(ctx), "dqmc_setup_storage");
        check(dqmc_set_performed_sweeps(ctx, performedSweeps), "dqmc_set_performed_sweeps");
    }

    dqmc_ctx* context() { return ctx; }

private:
    static void fillFromRng(void* self, double* out, size_t n) {
        RngWrapper& g = static_cast<DetHubbardGpu*>(self)->rng;
        for (size_t i = 0; i < n; ++i) out[i] = g.rand01();
    }
    void check(int status, const char* what) {
        if (status != DQMC_OK) throw_GeneralError(std::string(what) + " failed: " + dqmc_last_error(ctx));
    }
    std::vector<double> lookAhead() {
        size_t n = 0;
        check(dqmc_rng_look_ahead(ctx, 0, nullptr, &n), "dqmc_rng_look_ahead");
        std::vector<double> v(n);
        if (n) check(dqmc_rng_look_ahead(ctx, 0, v.data(), &n), "dqmc_rng_look_ahead");
        return v;
    }
    void fetchObservables() {
        double sc[8];
        std::vector<double> zc(N);
        check(dqmc_get_hubbard_observables(ctx, 0, sc, zc.data()), "dqmc_get_hubbard_observables");
        occUp = sc[0]; occDn = sc[1]; occTotal = sc[2]; occDouble = sc[3]; localMoment = sc[4];
        eKinetic = sc[5]; ePotential = sc[6]; eTotal = sc[7];
        for (uint32_t i = 0; i < N; ++i) zcorr[i] = zc[i];
    }

    RngWrapper& rng;
    Params pars;
    dqmc_ctx* ctx;
    uint32_t N;
    num occUp, occDn, occTotal, occDouble, localMoment, eKinetic, ePotential, eTotal;
    VecNum zcorr;
    uint32_t performedSweeps;
};

// same signature as the reference's createReplica (dethubbard.h:52-53, dethubbard.cpp:37-44)
inline void createReplica(std::unique_ptr<DetHubbardGpu>& replica_out, RngWrapper& rng, ModelParams<DetHubbard> pars,
                          DetModelLoggingParams /*ignored*/ = DetModelLoggingParams()) {
    pars = updateTemperatureParameters(pars);
    pars.check();
    replica_out = std::unique_ptr<DetHubbardGpu>(new DetHubbardGpu(rng, pars));
}

#endif  // DETHUBBARD_GPU_H_
