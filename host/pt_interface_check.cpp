// Compile-time check of the PT seam: every member DetQMCPT<Model, ModelParams> calls on its replica
// (detqmcpt.h:187-232, 348-386, 679, 697, 859-898, 971-1123) is called here on DetSDWGpu with the argument types the
// driver uses.  boost::mpi is not available in the development container, so detqmcpt.h itself cannot be compiled
// here; this translation unit instantiates the same duck-type instead (it is compiled by `make -C host`, never run
// on its own -- host/_build/pt_interface_check just prints the exchange probability of a fixed pair).
#include <iostream>
#include <sstream>
#include <string>

#include "boost/archive/binary_iarchive.hpp"
#include "boost/archive/binary_oarchive.hpp"
#include "boost/serialization/vector.hpp"
#include "boost/serialization/string.hpp"

#include "detsdw_gpu.h"

template <class Model, class Params>
void pt_duck_type(std::unique_ptr<Model>& replica, RngWrapper& rng, const Params& pars) {
    typedef typename Model::SystemConfig SystemConfig;                  // detqmcpt.h:187
    typedef typename Model::SystemConfig_FileHandle FileHandle;         // detqmcpt.h:193
    createReplica(replica, rng, pars, DetModelLoggingParams(), std::string("."));    // detqmcpt.h:316
    auto scalarObs = replica->getScalarObservables();                   // :348
    auto vectorObs = replica->getVectorObservables();                   // :355
    auto keyValueObs = replica->getKeyValueObservables();               // :361
    (void)scalarObs; (void)vectorObs; (void)keyValueObs;
    replica->saveConfigurationStreamTextHeader("# header\n", ".");      // :383
    replica->saveConfigurationStreamBinaryHeaderfile("# header\n", ".");    // :386
    FileHandle fh = replica->prepareSystemConfigurationStreamFileHandle(true, true, ".");   // :679
    SystemConfig sc = replica->getCurrentSystemConfiguration();        // :697
    sc.write_to_disk(fh);                                               // :747
    fh.flush();
    std::string buffer;
    serialize_systemConfig_to_buffer(buffer, sc);                       // :716
    replica->sweepThermalization();                                     // :862
    replica->thermalizationOver(0);                                     // :883
    replica->sweep(true);                                               // :898
    std::string control;
    replica->get_control_data(control);                                 // :971
    const double action = replica->get_exchange_action_contribution();  // :998
    const num par = replica->get_exchange_parameter_value();            // :1123
    const num prob = get_replica_exchange_probability<Model>(par, action, par + 0.1, action + 1.0);   // :1041
    replica->set_exchange_parameter_value(par);                         // :1096
    replica->set_control_data(control);                                 // :1115
    std::cout << replica->get_exchange_parameter_name() << " exchange probability " << prob << std::endl;
    std::stringstream ss;
    {
        boost::archive::binary_oarchive oa(ss);
        replica->saveContents(oa);                                      // :232
    }
    {
        boost::archive::binary_iarchive ia(ss);
        replica->loadContents(ia);                                      // :210
    }
}

int main() {
    ModelParamsDetSDW pars;
    RngWrapper rng(1020304050, 1);
    std::unique_ptr<DetSDWGpu<2>> replica;
    if (false) pt_duck_type(replica, rng, pars);        // instantiated, never executed here
    std::cout << get_replica_exchange_probability<DetSDWGpu<2>>(-1.0, 10.0, -0.9, 12.0) << std::endl;
    return 0;
}
