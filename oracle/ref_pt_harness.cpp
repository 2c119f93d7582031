// TEST INFRASTRUCTURE ONLY.
//
// ref_pt: the reference's replica-exchange driver DetQMCPT<DetSDW<CB_ASSAAD_BERG, 2>, ModelParamsDetSDW>
// (detqmcpt.h, compiled UNMODIFIED from the reference tree) with one THREAD per ladder process on top of
// oracle/fake_mpi/fake_boost_mpi.hpp (the container has no MPI).  It writes the reference's own output files into the
// working directory; tools/make_golden.py turns them into tests/golden/pt_reference.npz, which pins the exchange
// trajectory of detqmc_b200.DetQMCPT.
//
//   ref_pt <L> <beta> <s> <thermalization> <sweeps> <exchangeInterval> <r_0> <r_1> ...
#include <cstdlib>
#include <iostream>
#include <thread>
#include <vector>

#include "detqmcpt.h"
#include "detsdwopdim.h"

int main(int argc, char** argv) {
    if (argc < 9) { std::cerr << "usage: ref_pt L beta s thermalization sweeps exchangeInterval r_0 r_1 ...\n"; return 2; }
    const uint32_t L = std::atoi(argv[1]);
    const double beta = std::atof(argv[2]);
    const uint32_t s = std::atoi(argv[3]), therm = std::atoi(argv[4]), sweeps = std::atoi(argv[5]), xint = std::atoi(argv[6]);
    std::vector<double> rvals;
    for (int i = 7; i < argc; ++i) rvals.push_back(std::atof(argv[i]));
    const int P = (int)rvals.size();

    ModelParamsDetSDW pm;
#define SETM(name, value) { pm.name = (value); pm.specified.insert(#name); }
    SETM(model, std::string("sdw")); SETM(opdim, 2u); SETM(checkerboard, true);
    SETM(updateMethod_string, std::string("delayed")); pm.specified.insert("updateMethod");
    SETM(spinProposalMethod_string, std::string("box")); pm.specified.insert("spinProposalMethod");
    SETM(delaySteps, 16u); SETM(turnoffFermionMeasurements, true);
    SETM(r, rvals[0]); SETM(c, 3.0); SETM(u, 1.0); SETM(lambda, 1.0);
    SETM(txhor, -1.0); SETM(txver, -0.5); SETM(tyhor, 0.5); SETM(tyver, 1.0);
    SETM(cdwU, 0.0); SETM(mu, -0.5); SETM(weakZflux, true);
    SETM(L, L); SETM(d, 2u); SETM(beta, beta); SETM(dtau, 0.1); SETM(s, s);
    SETM(accRatio, 0.5); SETM(bc_string, std::string("pbc")); pm.specified.insert("bc");
    SETM(globalUpdateInterval, 10u); SETM(globalShift, true); SETM(repeatUpdateInSlice, 1u);
    SETM(wolffClusterUpdate, false); SETM(wolffClusterShiftUpdate, false);
#undef SETM
    DetQMCParams pq;
#define SETQ(name, value) { pq.name = (value); pq.specified.insert(#name); }
    SETQ(greenUpdateType_string, std::string("stabilized")); pq.specified.insert("greenUpdate");
    SETQ(sweeps, sweeps); SETQ(thermalization, therm); SETQ(jkBlocks, 1u); SETQ(measureInterval, 1u);
    SETQ(rngSeed, 1020304050u); SETQ(simindex, 0u); SETQ(timeseries, true);
    SETQ(stateFileName, std::string("simulation.state"));
#undef SETQ
    DetQMCPTParams pp;
    pp.exchangeInterval = xint; pp.specified.insert("exchangeInterval");
    pp.controlParameterName = "r"; pp.specified.insert("controlParameterName");
    pp.controlParameterValues = rvals; pp.specified.insert("controlParameterValues");

    boost::mpi::fake_world::get().init(P);
    std::vector<std::thread> threads;
    std::vector<int> rc(P, 0);
    for (int rank = 0; rank < P; ++rank)
        threads.emplace_back([&, rank] {
            boost::mpi::fake_world::rank() = rank;
            try {
                DetQMCPT<DetSDW<CB_ASSAAD_BERG, 2>, ModelParamsDetSDW> sim(pm, pq, pp);
                sim.run();
            } catch (const std::exception& e) {
                std::cerr << "[" << rank << "] " << e.what() << std::endl;
                rc[rank] = 1;
                std::_Exit(1);                   // the other threads would wait at a barrier forever
            }
        });
    for (auto& t : threads) t.join();
    return 0;
}
