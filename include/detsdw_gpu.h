// detsdw_gpu.h -- C++ host shim: the reference's Model duck-type for DetSDW, served by libdqmc_b200.so.
//
// This is the upper seam of SURVEY.md section 8(b).  `DetSDWGpu<OPDIM>` is a drop-in for
// `DetSDW<CB_ASSAAD_BERG, OPDIM>` (detsdwopdim.h:62-153) as a template argument of the reference's
// drivers `DetQMC<Model, ModelParams>` (detqmc.h:56-159) and `DetQMCPT<Model, ModelParams>`
// (detqmcpt.h): same member names, same argument meaning, errors surface as the reference's
// GeneralError (exceptions.h).  It compiles against the reference's own headers (add the reference's
// src/ and this directory to the include path, link libdqmc_b200.so); nothing of the reference is
// copied here.  The sweep itself -- checkerboard multiplies, wraps, UDT chains, Green's functions,
// delayed updates, global shift moves -- runs on the GPU through include/dqmc_gpu.h.
//
// Differences a maintainer has to know:
//   * fermionic measurements (measure(k), detsdwopdim.cpp:540-900) run on the device during sweep(true) unless
//     turnoffFermionMeasurements is set; the bosonic observables are evaluated on the host from the downloaded
//     field after every measurement sweep.
//   * random numbers: the replica consumes the driver's RngWrapper through a pre-drawn FIFO
//     (dqmc_rng_set_source).  Values drawn ahead but not yet consumed are part of the model's state:
//     saveContents stores them, loadContents puts them back in front of the stream, so a resumed run continues
//     with exactly the numbers an uninterrupted run would have used.
//   * checkpoints are NOT interchangeable with the reference's (different archive layout: fields, control blob,
//     look-ahead, sweep counter).
#ifndef DETSDW_GPU_H_
#define DETSDW_GPU_H_

#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "dqmc_gpu.h"

// reference headers (crstnbr/detqmc, src/)
#include "detmodel.h"
#include "detmodelloggingparams.h"
#include "detsdwparams.h"
#include "detsdwsystemconfig.h"
#include "detsdwsystemconfigfilehandle.h"
#include "exceptions.h"
#include "metadata.h"
#include "observable.h"
#include "rngwrapper.h"

template <int OPDIM>
class DetSDWGpu : public DetModel {
public:
    typedef ModelParamsDetSDW ModelParams;
    typedef DetSDW_SystemConfig SystemConfig;
    typedef DetSDW_SystemConfig_FileHandle SystemConfig_FileHandle;

    DetSDWGpu(RngWrapper& rng_, const ModelParams& pars_, int device = 0)
        : rng(rng_), pars(pars_), ctx(nullptr), normMeanPhi(0), associatedEnergy(0), phiRhoS_Gs(0), phiRhoS_Gc(0),
          greenK0(0), greenLocal(0), occDiffSq(0), pairPlusMax(0), pairMinusMax(0),
          performedSweeps(0) {
        if (pars.opdim != (uint32_t)OPDIM) throw_GeneralError("DetSDWGpu: opdim mismatch");
        kOccX.zeros(pars.L * pars.L); kOccY.zeros(pars.L * pars.L);
        pairPlus.zeros(pars.L * pars.L); pairMinus.zeros(pars.L * pars.L);
        if (!pars.checkerboard && !pars.turnoffFermionMeasurements)
            throw_GeneralError("DetSDWGpu: fermionic measurements are served with the checkerboard break-up only");
        if (pars.turnoffFermions) throw_GeneralError("DetSDWGpu: turnoffFermions is a pure-boson run, use the reference");
        if (pars.cdwU != 0.0) throw_GeneralError("DetSDWGpu: cdwU != 0 is outside the accelerated path");
        // iterative (detsdwopdim.cpp:2491-2880) and woodbury (:2883-3019) evaluate the same ratio and apply the same
        // rank-MSF update immediately: both are served as delayed updates with a block of one (identical decisions)
        if (pars.updateMethod_string != "delayed" && pars.updateMethod_string != "woodbury" &&
            pars.updateMethod_string != "iterative")
            throw_GeneralError("DetSDWGpu: updateMethod must be iterative, woodbury or delayed");
        if (pars.spinProposalMethod_string != "box")
            throw_GeneralError("DetSDWGpu: only box proposals are implemented");
        dqmc_params p;
        std::memset(&p, 0, sizeof p);
        p.model = DQMC_MODEL_SDW;
        p.opdim = OPDIM;
        p.L = (int32_t)pars.L;
        p.m = (int32_t)pars.m;
        p.s = (int32_t)pars.s;
        p.bc = pars.bc_string == "pbc" ? 0 : pars.bc_string == "apbc-x" ? 1 : pars.bc_string == "apbc-y" ? 2 : 3;
        p.weakZflux = pars.weakZflux ? 1 : 0;
        p.denseHopping = pars.checkerboard ? 0 : 1;          // checkerboard = false: DetSDW<CB_NONE, OPDIM>, dense e^{-dtau K}
        p.delaySteps = pars.updateMethod_string == "delayed" ? (int32_t)pars.delaySteps : 1;
        p.globalShift = pars.globalShift ? 1 : 0;
        p.wolffClusterUpdate = pars.wolffClusterUpdate ? 1 : 0;
        p.wolffClusterShiftUpdate = pars.wolffClusterShiftUpdate ? 1 : 0;
        p.repeatWolffPerSweep = (int32_t)pars.repeatWolffPerSweep;
        p.repeatUpdateInSlice = (int32_t)pars.repeatUpdateInSlice;
        p.globalUpdateInterval = (int32_t)pars.globalUpdateInterval;
        p.dtau = pars.dtau;
        p.r = pars.r; p.c = pars.c; p.u = pars.u; p.lambda = pars.lambda;
        p.txhor = pars.txhor; p.txver = pars.txver; p.tyhor = pars.tyhor; p.tyver = pars.tyver;
        // mux and muy supersede mu only when both are given (createReplica, detsdwopdim.cpp:76-80)
        const bool sep = pars.specified.count("mux") && pars.specified.count("muy");
        p.mux = sep ? pars.mux : pars.mu;
        p.muy = sep ? pars.muy : pars.mu;
        p.accRatio = pars.accRatio;
        check(dqmc_create(&p, 1, device, &ctx), "dqmc_create");
        // the driver's generator is THE stream of this replica
        check(dqmc_rng_set_source(ctx, 0, &DetSDWGpu::fillFromRng, this), "dqmc_rng_set_source");
        check(dqmc_init_random_fields(ctx, 0), "dqmc_init_random_fields");   // setupRandomField
        check(dqmc_setup_storage(ctx), "dqmc_setup_storage");                // setupUdVStorage_and_calculateGreen
    }
    virtual ~DetSDWGpu() { dqmc_destroy(ctx); }

    virtual uint32_t getSystemN() const { return pars.L * pars.L; }
    virtual MetadataMap prepareModelMetadataMap() const {
        MetadataMap meta = pars.prepareMetadataMap();
        dqmc_control_data cd;
        dqmc_get_control_data(ctx, 0, &cd);
        meta["phiDelta"] = numToString(cd.phiDelta);
        meta["globalShiftAccRatio"] =
            numToString(cd.attemptedGlobalShifts ? double(cd.acceptedGlobalShifts) / cd.attemptedGlobalShifts : 0.0);
        double ws[5] = {0, 0, 0, 0, 0};                      // detsdwopdim.cpp:406-435
        dqmc_get_wolff_statistics(ctx, 0, ws);
        if (pars.wolffClusterUpdate) {
            meta["wolffClusterUpdateAccRatio"] = numToString(ws[0] > 0 ? ws[1] / ws[0] : 0.0);
            meta["averageAcceptedWolffClusterSize"] = numToString(ws[1] > 0 ? ws[4] / ws[1] : 0.0);
        }
        if (pars.wolffClusterShiftUpdate) {
            meta["wolffClusterShiftUpdateAccRatio"] = numToString(ws[2] > 0 ? ws[3] / ws[2] : 0.0);
            meta["averageAcceptedWolffClusterSize"] = numToString(ws[3] > 0 ? ws[4] / ws[3] : 0.0);
        }
        meta["backend"] = "libdqmc_b200 (sm_100a)";
        return meta;
    }

    virtual void thermalizationOver() {
        dqmc_control_data cd;
        dqmc_get_control_data(ctx, 0, &cd);
        std::cout << "After thermalization: phiDelta = " << cd.phiDelta << '\n'
                  << "recent local accRatio = " << cd.lastAccRatioLocal_phi << std::endl;
    }
    virtual void thermalizationOver(int processIndex) {
        std::cout << "[" << processIndex << "] ";
        thermalizationOver();
    }

    virtual void sweep(bool takeMeasurements) {
        const bool fermionic = takeMeasurements && !pars.turnoffFermionMeasurements;
        check(dqmc_sweep(ctx, fermionic ? 2 : 0), "dqmc_sweep");
        ++performedSweeps;
        if (takeMeasurements) measureBosonic();
        if (fermionic) fetchFermionic();
    }
    virtual void sweepThermalization() {
        check(dqmc_sweep(ctx, 1), "dqmc_sweep");
        ++performedSweeps;
    }
    // greenUpdate = simple (detsdwopdim.cpp:4366-4420): G from scratch at every slice, then the slice update
    virtual void sweepSimple(bool takeMeasurements) {
        if (takeMeasurements && !pars.turnoffFermionMeasurements)
            throw_GeneralError("DetSDWGpu: fermionic measurements are served by the stabilized sweep only");
        check(dqmc_sweep_simple(ctx, 0), "dqmc_sweep_simple");
        ++performedSweeps;
        if (takeMeasurements) measureBosonic();
    }
    virtual void sweepSimpleThermalization() {
        check(dqmc_sweep_simple(ctx, 1), "dqmc_sweep_simple");
        ++performedSweeps;
    }

    virtual std::vector<ScalarObservable> getScalarObservables() {
        std::vector<ScalarObservable> obs;
        // the reference's list for turnoffFermionMeasurements (detsdwopdim.cpp:270-276)
        obs.push_back(ScalarObservable(std::cref(normMeanPhi), "normMeanPhi", "nmp"));
        obs.push_back(ScalarObservable(std::cref(associatedEnergy), "associatedEnergy", ""));
        if (OPDIM == 2) {
            obs.push_back(ScalarObservable(std::cref(phiRhoS_Gs), "phiRhoS_Gs", ""));
            obs.push_back(ScalarObservable(std::cref(phiRhoS_Gc), "phiRhoS_Gc", ""));
        }
        if (!pars.turnoffFermionMeasurements) {              // detsdwopdim.cpp:278-333
            obs.push_back(ScalarObservable(std::cref(pairPlusMax), "pairPlusMax", "ppMax"));
            obs.push_back(ScalarObservable(std::cref(pairMinusMax), "pairMinusMax", "pmMax"));
            obs.push_back(ScalarObservable(std::cref(greenK0), "greenK0", ""));
            obs.push_back(ScalarObservable(std::cref(greenLocal), "greenLocal", ""));
            obs.push_back(ScalarObservable(std::cref(occDiffSq), "occDiffSq", ""));
        }
        return obs;
    }
    virtual std::vector<VectorObservable> getVectorObservables() {
        std::vector<VectorObservable> obs;
        if (!pars.turnoffFermionMeasurements) {              // detsdwopdim.cpp:287-315
            const uint32_t N = pars.L * pars.L;
            obs.push_back(VectorObservable(std::cref(kOccX), N, "kOccX", "nkx"));
            obs.push_back(VectorObservable(std::cref(kOccY), N, "kOccY", "nky"));
            obs.push_back(VectorObservable(std::cref(pairPlus), N, "pairPlus", "pp"));
            obs.push_back(VectorObservable(std::cref(pairMinus), N, "pairMinus", "pm"));
        }
        return obs;
    }
    virtual std::vector<KeyValueObservable> getKeyValueObservables() { return std::vector<KeyValueObservable>(); }

    // configuration streams (detsdwopdim.cpp:4943-5114): the device reorders the fields into the on-disk order
    // (ix, iy, k, dim), the files are written exactly like the reference's (names, append mode, number format)
    std::vector<double> currentConfigurationStream() {
        std::vector<double> cfg(size_t(pars.L) * pars.L * pars.m * OPDIM);
        check(dqmc_download_config_stream(ctx, 0, cfg.data()), "dqmc_download_config_stream");
        return cfg;
    }
    void saveConfigurationStreamText(const std::string& directory = ".") {
        const std::string path = directory + "/configs-phi.textstream";
        std::ofstream out(path.c_str(), std::ios::app);
        if (!out) { std::cerr << "Could not open file " << path << " for writing.\n"; return; }
        out.precision(14);
        out.setf(std::ios::scientific, std::ios::floatfield);
        const std::vector<double> cfg = currentConfigurationStream();
        for (size_t i = 0; i < cfg.size(); ++i) out << cfg[i] << "\n";
        out.flush();
    }
    void saveConfigurationStreamBinary(const std::string& directory = ".") {
        const std::string path = directory + "/configs-phi.binarystream";
        std::ofstream out(path.c_str(), std::ios::binary | std::ios::app);
        if (!out) { std::cerr << "Could not open file " << path << " for writing.\n"; return; }
        const std::vector<double> cfg = currentConfigurationStream();
        out.write(reinterpret_cast<const char*>(cfg.data()), std::streamsize(cfg.size() * sizeof(double)));
        out.flush();
    }
    void saveConfigurationStreamTextHeader(const std::string& simInfoHeaderText, const std::string& directory = ".") {
        const std::string path = directory + "/configs-phi.textstream";
        if (std::ifstream(path.c_str())) return;               // only if the file does not exist yet (:5043)
        std::ofstream out(path.c_str(), std::ios::out);
        if (!out) { std::cerr << "Could not open file " << path << " for writing.\n"; return; }
        out << simInfoHeaderText << "## phi configuration stream\n";
        out.flush();
    }
    void saveConfigurationStreamBinaryHeaderfile(const std::string& simInfoHeaderText, const std::string& directory = ".") {
        const std::string path = directory + "/configs-phi.infoheader";
        if (std::ifstream(path.c_str())) return;
        std::ofstream out(path.c_str(), std::ios::out);
        if (!out) { std::cerr << "Could not open file " << path << " for writing.\n"; return; }
        out << simInfoHeaderText
            << "## binary phi configuration stream (64 bit double precision floats) in file configs-phi.binarystream\n";
        out.flush();
    }

    // replica exchange interface (detsdwopdim.cpp:5189-5242)
    num get_exchange_parameter_value() const {
        double r = 0;
        dqmc_get_exchange_parameter(ctx, 0, &r);
        return r;
    }
    void set_exchange_parameter_value(num r) {
        pars.r = r;
        check(dqmc_set_exchange_parameter(ctx, 0, r), "dqmc_set_exchange_parameter");
    }
    const char* get_exchange_parameter_name() const { return "r"; }
    num get_exchange_action_contribution() const {
        double a = 0;
        const_cast<DetSDWGpu*>(this)->check(dqmc_exchange_actions(ctx, nullptr, &a), "dqmc_exchange_actions");
        return a;
    }
    void get_control_data(std::string& buffer) const {
        dqmc_control_data cd;
        dqmc_get_control_data(ctx, 0, &cd);
        buffer.assign(reinterpret_cast<const char*>(&cd), sizeof cd);
    }
    void set_control_data(const std::string& buffer) {
        if (buffer.size() != sizeof(dqmc_control_data)) throw_GeneralError("DetSDWGpu: bad control data blob");
        dqmc_control_data cd;
        std::memcpy(&cd, buffer.data(), sizeof cd);
        check(dqmc_set_control_data(ctx, 0, &cd), "dqmc_set_control_data");
    }

    // checkpointing: fields + control data + random numbers drawn ahead of consumption + sweep counter (G and the
    // UDT storage are rebuilt, as in the reference: detsdwopdim.h:1117-1148, detmodel.h:493-502)
    template <class Archive>
    void saveContents(Archive& ar) {
        std::vector<double> phi = downloadPhi();
        dqmc_control_data cd;
        dqmc_get_control_data(ctx, 0, &cd);
        std::string blob(reinterpret_cast<const char*>(&cd), sizeof cd);
        size_t n = 0;
        check(dqmc_rng_look_ahead(ctx, 0, nullptr, &n), "dqmc_rng_look_ahead");
        std::vector<double> ahead(n);
        if (n) check(dqmc_rng_look_ahead(ctx, 0, ahead.data(), &n), "dqmc_rng_look_ahead");
        ar & phi & blob & ahead & performedSweeps;
    }
    template <class Archive>
    void loadContents(Archive& ar) {
        std::vector<double> phi, ahead;
        std::string blob;
        ar & phi & blob & ahead & performedSweeps;
        check(dqmc_upload_fields(ctx, 0, phi.data()), "dqmc_upload_fields");
        set_control_data(blob);
        check(dqmc_rng_set_look_ahead(ctx, 0, ahead.data(), ahead.size()), "dqmc_rng_set_look_ahead");
        check(dqmc_setup_storage(ctx), "dqmc_setup_storage");
        check(dqmc_set_performed_sweeps(ctx, performedSweeps), "dqmc_set_performed_sweeps");
    }

    // system configurations for DetQMCPT's buffered configuration streams (detqmcpt.h:679, 697;
    // detsdwopdim.cpp:5116-5160): the reference's own DetSDW_SystemConfig / file-handle types, filled from the device
    SystemConfig getCurrentSystemConfiguration() {
        std::vector<double> phi = downloadPhi();           // [k][dim][site] == Cube(site, dim, k), column major
        const CubeNum cube(phi.data(), pars.L * pars.L, OPDIM, pars.m + 1);
        return SystemConfig(pars, cube);
    }
    SystemConfig_FileHandle prepareSystemConfigurationStreamFileHandle(bool binaryStream, bool textStream,
                                                                       const std::string& directory = ".") {
        if (!binaryStream && !textStream) throw_GeneralError("binaryStream or textStream must be specified to create a file handle");
        typedef SystemConfig_FileHandle::OfstreamPointer Ptr;
        SystemConfig_FileHandle fh;
        if (binaryStream) {
            const std::string path = directory + "/configs-phi.binarystream";
            fh.phi_output_binary = Ptr(new std::ofstream(path.c_str(), std::ios::binary | std::ios::app));
            if (fh.phi_output_binary->fail()) std::cerr << "Could not open file " << path << " for writing.\n";
        }
        if (textStream) {
            const std::string path = directory + "/configs-phi.textstream";
            fh.phi_output_text = Ptr(new std::ofstream(path.c_str(), std::ios::app));
            if (fh.phi_output_text->fail()) std::cerr << "Could not open file " << path << " for writing.\n";
            fh.phi_output_text->precision(14);
            fh.phi_output_text->setf(std::ios::scientific, std::ios::floatfield);
        }
        return fh;
    }

    dqmc_ctx* context() { return ctx; }

private:
    static void fillFromRng(void* self, double* out, size_t n) {
        RngWrapper& g = static_cast<DetSDWGpu*>(self)->rng;
        for (size_t i = 0; i < n; ++i) out[i] = g.rand01();
    }
    void check(int status, const char* what) {
        if (status != DQMC_OK)
            throw_GeneralError(std::string(what) + " failed: " + dqmc_last_error(ctx));
    }
    std::vector<double> downloadPhi() {
        std::vector<double> phi(size_t(pars.m + 1) * OPDIM * pars.L * pars.L);
        check(dqmc_download_fields(ctx, 0, phi.data()), "dqmc_download_fields");
        return phi;
    }
    // bosonic observables of DetSDW::measure (detsdwopdim.cpp:440-506): |mean phi|, mean phi^2, action
    // finishMeasurements of the fermionic observables accumulated on the device during dqmc_sweep(ctx, 2)
    void fetchFermionic() {
        const uint32_t N = pars.L * pars.L;
        double sc[5];
        std::vector<double> vec(size_t(4) * N);
        check(dqmc_get_fermionic_observables(ctx, 0, sc, vec.data()), "dqmc_get_fermionic_observables");
        greenK0 = sc[0]; greenLocal = sc[1]; occDiffSq = sc[2]; pairPlusMax = sc[3]; pairMinusMax = sc[4];
        for (uint32_t i = 0; i < N; ++i) {
            kOccX[i] = vec[i]; kOccY[i] = vec[N + i]; pairPlus[i] = vec[2 * N + i]; pairMinus[i] = vec[3 * N + i];
        }
    }

    // The observables the reference measures with turnoffFermionMeasurements (initMeasurements / measure /
    // finishMeasurements, detsdwopdim.cpp:441-560, 903-918), from the fields of all slices k = 1..m.
    void measureBosonic() {
        const std::vector<double> phi = downloadPhi();
        const uint32_t L = pars.L;
        const size_t N = size_t(L) * L;
        auto at = [&](uint32_t k, int d, size_t s) { return phi[(size_t(k) * OPDIM + d) * N + s]; };
        double mean[3] = {0, 0, 0}, sq = 0, gc = 0, gs = 0;
        for (uint32_t k = 1; k <= pars.m; ++k)
            for (size_t s = 0; s < N; ++s) {
                const size_t x = s % L, y = s / L;
                const size_t xp = y * L + (x + 1) % L, yp = ((y + 1) % L) * L + x;
                for (int d = 0; d < OPDIM; ++d) {
                    const double v = at(k, d, s);
                    mean[d] += v;
                    sq += v * v;
                    if (OPDIM == 2) gc += v * at(k, d, xp) + v * at(k, d, yp);
                }
                if (OPDIM == 2) gs += at(k, 0, xp) * at(k, 1, s) - at(k, 1, xp) * at(k, 0, s);
            }
        double nrm = 0;
        for (int d = 0; d < OPDIM; ++d) { mean[d] /= double(N * pars.m); nrm += mean[d] * mean[d]; }
        normMeanPhi = std::sqrt(nrm);
        associatedEnergy = sq / (2.0 * double(N * pars.m));
        phiRhoS_Gc = gc * (0.5 * pars.dtau);
        phiRhoS_Gs = gs * pars.dtau;
    }

    RngWrapper& rng;
    ModelParams pars;
    dqmc_ctx* ctx;
    num normMeanPhi, associatedEnergy, phiRhoS_Gs, phiRhoS_Gc;
    num greenK0, greenLocal, occDiffSq, pairPlusMax, pairMinusMax;
    VecNum kOccX, kOccY, pairPlus, pairMinus;
    uint32_t performedSweeps;
};

// same signature as the reference's createReplica (detsdwopdim.cpp:48-84)
template <int OPDIM>
void createReplica(std::unique_ptr<DetSDWGpu<OPDIM>>& replica_out, RngWrapper& rng, ModelParamsDetSDW pars,
                   DetModelLoggingParams /*loggingPars*/ = DetModelLoggingParams(), const std::string& /*logfiledir*/ = "") {
    pars = updateTemperatureParameters(pars);
    pars.check();
    replica_out = std::unique_ptr<DetSDWGpu<OPDIM>>(new DetSDWGpu<OPDIM>(rng, pars));
}

// replica-exchange probability for the PT driver: DetQMCPT calls get_replica_exchange_probability<Model>
// (detqmcpt.h:1041), a function template the model specialises (detmodel.h:93-108, detsdwopdim.cpp:5251-5325)
template <int OPDIM>
inline num get_replica_exchange_probability_gpu(num r1, num action1, num r2, num action2) {
    return dqmc_exchange_probability(r1, action1, r2, action2);
}
template <>
inline num get_replica_exchange_probability<DetSDWGpu<1>>(num r1, num action1, num r2, num action2) {
    return dqmc_exchange_probability(r1, action1, r2, action2);
}
template <>
inline num get_replica_exchange_probability<DetSDWGpu<2>>(num r1, num action1, num r2, num action2) {
    return dqmc_exchange_probability(r1, action1, r2, action2);
}
template <>
inline num get_replica_exchange_probability<DetSDWGpu<3>>(num r1, num action1, num r2, num action2) {
    return dqmc_exchange_probability(r1, action1, r2, action2);
}

#endif  // DETSDW_GPU_H_
