// detqmcsdw_gpu -- the reference's single-replica driver DetQMC<Model, ModelParams> (detqmc.h, compiled
// UNMODIFIED from the reference tree) instantiated with the GPU model shim include/detsdw_gpu.h.
// This is the integration example of INTEGRATION.md: what maindetqmcsdwopdim.cpp:252-354 does with
// DetSDW<CB_ASSAAD_BERG, OPDIM>, done with DetSDWGpu<OPDIM>.  Parameters are given as key=value
// arguments with the reference's option names (a stand-in for its boost::program_options front end,
// which is not part of the hot path).
//
//   detqmcsdw_gpu opdim=2 L=4 beta=2 dtau=0.1 s=10 r=-1 thermalization=20 sweeps=20 ...
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>

#include "detqmc.h"
#include "detsdw_gpu.h"

namespace {
std::map<std::string, std::string> parse(int argc, char** argv) {
    std::map<std::string, std::string> kv;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        const size_t eq = a.find('=');
        if (eq == std::string::npos) { std::cerr << "expected key=value, got " << a << "\n"; std::exit(2); }
        kv[a.substr(0, eq)] = a.substr(eq + 1);
    }
    return kv;
}
template <class T, class S>
void take(std::map<std::string, std::string>& kv, S& specified, const char* key, T& field, const T& dflt) {
    auto it = kv.find(key);
    if (it != kv.end()) { field = fromString<T>(it->second); kv.erase(it); }
    else field = dflt;
    specified.insert(key);
}

// a state file in the working directory resumes the simulation, as in maindetqmcsdwopdim.cpp:197-200, 320-340
template <int OPDIM>
int run(ModelParamsDetSDW& pm, DetQMCParams& pq) {
    typedef DetQMC<DetSDWGpu<OPDIM>, ModelParamsDetSDW> Sim;
    if (std::ifstream(pq.stateFileName.c_str())) {
        std::cout << "Found simulation state file " << pq.stateFileName << ", will resume simulation" << std::endl;
        Sim sim(pq.stateFileName, pq);
        sim.run();
    } else {
        Sim sim(pm, pq);
        sim.run();
    }
    return 0;
}
}  // namespace

int main(int argc, char** argv) {
    auto kv = parse(argc, argv);
    ModelParamsDetSDW pm;
    DetQMCParams pq;
    try {
        // model parameters (defaults of maindetqmcsdwopdim.cpp:93-135 unless noted)
        take(kv, pm.specified, "model", pm.model, std::string("sdw"));
        take(kv, pm.specified, "opdim", pm.opdim, uint32_t(2));
        take(kv, pm.specified, "checkerboard", pm.checkerboard, true);
        take(kv, pm.specified, "updateMethod", pm.updateMethod_string, std::string("delayed"));
        take(kv, pm.specified, "spinProposalMethod", pm.spinProposalMethod_string, std::string("box"));
        take(kv, pm.specified, "delaySteps", pm.delaySteps, uint32_t(16));
        take(kv, pm.specified, "turnoffFermionMeasurements", pm.turnoffFermionMeasurements, true);
        take(kv, pm.specified, "r", pm.r, -1.0);
        take(kv, pm.specified, "c", pm.c, 3.0);
        take(kv, pm.specified, "u", pm.u, 1.0);
        take(kv, pm.specified, "lambda", pm.lambda, 1.0);
        take(kv, pm.specified, "txhor", pm.txhor, -1.0);
        take(kv, pm.specified, "txver", pm.txver, -0.5);
        take(kv, pm.specified, "tyhor", pm.tyhor, 0.5);
        take(kv, pm.specified, "tyver", pm.tyver, 1.0);
        take(kv, pm.specified, "cdwU", pm.cdwU, 0.0);
        take(kv, pm.specified, "mu", pm.mu, -0.5);
        take(kv, pm.specified, "weakZflux", pm.weakZflux, true);
        take(kv, pm.specified, "L", pm.L, uint32_t(4));
        take(kv, pm.specified, "d", pm.d, uint32_t(2));
        take(kv, pm.specified, "beta", pm.beta, 2.0);
        take(kv, pm.specified, "dtau", pm.dtau, 0.1);
        take(kv, pm.specified, "s", pm.s, uint32_t(10));
        take(kv, pm.specified, "accRatio", pm.accRatio, 0.5);
        take(kv, pm.specified, "bc", pm.bc_string, std::string("pbc"));
        take(kv, pm.specified, "globalUpdateInterval", pm.globalUpdateInterval, uint32_t(10));
        take(kv, pm.specified, "globalShift", pm.globalShift, true);
        take(kv, pm.specified, "repeatUpdateInSlice", pm.repeatUpdateInSlice, uint32_t(1));
        take(kv, pm.specified, "wolffClusterUpdate", pm.wolffClusterUpdate, false);
        take(kv, pm.specified, "wolffClusterShiftUpdate", pm.wolffClusterShiftUpdate, false);
        // Monte Carlo parameters
        take(kv, pq.specified, "greenUpdate", pq.greenUpdateType_string, std::string("stabilized"));
        take(kv, pq.specified, "sweeps", pq.sweeps, uint32_t(20));
        take(kv, pq.specified, "thermalization", pq.thermalization, uint32_t(20));
        take(kv, pq.specified, "jkBlocks", pq.jkBlocks, uint32_t(1));
        take(kv, pq.specified, "measureInterval", pq.measureInterval, uint32_t(1));
        if (kv.count("saveInterval")) take(kv, pq.specified, "saveInterval", pq.saveInterval, uint32_t(0));   // default: only at the end
        take(kv, pq.specified, "rngSeed", pq.rngSeed, uint32_t(1020304050));
        take(kv, pq.specified, "simindex", pq.simindex, uint32_t(0));
        take(kv, pq.specified, "timeseries", pq.timeseries, true);
        if (kv.count("saveConfigurationStreamInterval")) {
            take(kv, pq.specified, "saveConfigurationStreamInterval", pq.saveConfigurationStreamInterval, uint32_t(0));
            take(kv, pq.specified, "saveConfigurationStreamText", pq.saveConfigurationStreamText, false);
            take(kv, pq.specified, "saveConfigurationStreamBinary", pq.saveConfigurationStreamBinary, false);
        }
        take(kv, pq.specified, "state", pq.stateFileName, std::string("simulation.state"));
        if (!kv.empty()) {
            std::cerr << "unknown option: " << kv.begin()->first << "\n";
            return 2;
        }
        switch (pm.opdim) {
            case 1: return run<1>(pm, pq);
            case 2: return run<2>(pm, pq);
            case 3: return run<3>(pm, pq);
            default: std::cerr << "opdim must be 1, 2 or 3\n"; return 2;
        }
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
}
