// detqmchubbard_gpu -- the reference's single-replica driver DetQMC<Model, ModelParams> (detqmc.h, compiled
// UNMODIFIED from the reference tree) instantiated with the GPU model shim include/dethubbard_gpu.h:
// what maindetqmchubbard.cpp does with DetHubbard, done with DetHubbardGpu.  Parameters are key=value arguments
// with the reference's option names (maindetqmchubbard.cpp:55-90); a state file found in the working directory
// resumes the simulation exactly like the reference's main (maindetqmchubbard.cpp:97-100, 177-181).
//
//   detqmchubbard_gpu L=4 U=4 beta=4 dtau=0.1 s=10 mu=0 thermalization=10 sweeps=10 timeseries=1
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>

#include "detqmc.h"
#include "dethubbard_gpu.h"

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
}  // namespace

int main(int argc, char** argv) {
    auto kv = parse(argc, argv);
    ModelParams<DetHubbard> pm;
    DetQMCParams pq;
    try {
        take(kv, pm.specified, "model", pm.model, std::string("hubbard"));
        take(kv, pm.specified, "checkerboard", pm.checkerboard, false);
        take(kv, pm.specified, "t", pm.t, 1.0);
        take(kv, pm.specified, "U", pm.U, 4.0);
        take(kv, pm.specified, "mu", pm.mu, 0.5);
        take(kv, pm.specified, "L", pm.L, uint32_t(4));
        take(kv, pm.specified, "d", pm.d, uint32_t(2));
        take(kv, pm.specified, "beta", pm.beta, 4.0);
        take(kv, pm.specified, "dtau", pm.dtau, 0.1);
        take(kv, pm.specified, "s", pm.s, uint32_t(1));
        take(kv, pq.specified, "greenUpdate", pq.greenUpdateType_string, std::string("stabilized"));
        take(kv, pq.specified, "sweeps", pq.sweeps, uint32_t(10));
        take(kv, pq.specified, "thermalization", pq.thermalization, uint32_t(10));
        take(kv, pq.specified, "jkBlocks", pq.jkBlocks, uint32_t(1));
        take(kv, pq.specified, "measureInterval", pq.measureInterval, uint32_t(1));
        if (kv.count("saveInterval")) take(kv, pq.specified, "saveInterval", pq.saveInterval, uint32_t(0));
        take(kv, pq.specified, "rngSeed", pq.rngSeed, uint32_t(1020304050));
        take(kv, pq.specified, "simindex", pq.simindex, uint32_t(0));
        take(kv, pq.specified, "timeseries", pq.timeseries, true);
        take(kv, pq.specified, "state", pq.stateFileName, std::string("simulation.state"));
        if (!kv.empty()) {
            std::cerr << "unknown option: " << kv.begin()->first << "\n";
            return 2;
        }
        typedef DetQMC<DetHubbardGpu, ModelParams<DetHubbard>> Sim;
        if (std::ifstream(pq.stateFileName.c_str())) {
            std::cout << "Found simulation state file " << pq.stateFileName << ", will resume simulation" << std::endl;
            Sim sim(pq.stateFileName, pq);
            sim.run();
        } else {
            Sim sim(pm, pq);
            sim.run();
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
}
