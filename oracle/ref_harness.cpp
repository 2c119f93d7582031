// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// Thin extern "C" harness around the UNMODIFIED reference sources that live under
// /root/reference/src (crstnbr/detqmc).  It is compiled by oracle/Makefile into
// oracle/_ref/libdetqmc_ref.so (git-ignored) together with the reference's own translation
// units; nothing from the reference is copied into this repository.  The harness is built with
// -fno-access-control so it can reach the private state of DetSDW / DetHubbard / DetModelGC
// (G, phi, UdV storage, step sizes) and their private checkerboard multiply routines.
//
// What it pins:
//   * RngWrapper / dSFMT-19937 stream               (rngwrapper.h:43-119, rngwrapper.cpp:43)
//   * DetSDW construction from a seed               (detsdwopdim.cpp:48-84, 157-361)
//   * checkerboard{Left,Right}MultiplyBmat[Inv]     (detsdwopdim.cpp:2074-2420)
//   * G, singular values of G^-1 after setup        (detmodel.h:678-713, 822-860)
//   * updateInSlice / updateInSliceThermalization   (detsdwopdim.cpp:2427-2489, 3293-3375)
//   * sweep / sweepThermalization                   (detsdwopdim.cpp:4422-4502)
//   * exchange action and exchange probability      (detsdwopdim.cpp:5204-5264)
//   * DetHubbard construction, sweep, dense B       (dethubbard.cpp:46-171, 823-962)
//
// Used by: tests/ (as the checker), tools/make_golden.py (fixture generation) and, optionally,
// bench.py's cpu_baseline / --impl reference arm.

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <complex>
#include <iostream>
#include <sstream>

#include "rngwrapper.h"
#include "detmodelparams.h"
#include "detmodelloggingparams.h"
#include "detsdwparams.h"
#include "detsdwopdim.h"
#include "dethubbardparams.h"
#include "dethubbard.h"

typedef std::complex<double> cpx_t;

extern "C" {

struct ref_sdw_params {
    int32_t opdim;
    int32_t L;
    int32_t m;
    int32_t s;
    double dtau;
    double r, c, u, lambda;
    double txhor, txver, tyhor, tyver;
    double cdwU, mu;
    double accRatio;
    int32_t weakZflux;
    int32_t bc;              // 0 pbc, 1 apbc-x, 2 apbc-y, 3 apbc-xy
    int32_t updateMethod;    // 0 iterative, 1 woodbury, 2 delayed
    int32_t delaySteps;
    int32_t globalShift;
    int32_t globalUpdateInterval;
    int32_t repeatUpdateInSlice;
    uint32_t seed;
    uint32_t rngIndex;
    int32_t wolffClusterUpdate;
    int32_t wolffClusterShiftUpdate;
    int32_t repeatWolffPerSweep;
    int32_t fermionMeasurements;   // 0: turnoffFermionMeasurements (default), 1: measure
    int32_t denseHopping;          // 1: checkerboard = false, DetSDW<CB_NONE, OPDIM>
};

struct ref_hub_params {
    int32_t L;
    int32_t m;
    int32_t s;
    int32_t checkerboard;
    double dtau;
    double t, U, mu;
    uint32_t seed;
    uint32_t rngIndex;
};

}  // extern "C"

namespace {

// silence the reference's chatter on stdout while we drive it
struct CoutSilencer {
    std::streambuf* old;
    std::ostringstream sink;
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};

struct SdwBase {
    virtual ~SdwBase() {}
    virtual void dims(int32_t* out) = 0;
    virtual void get_phi(double* out) = 0;
    virtual void set_phi(const double* in) = 0;
    virtual void get_green(double* out) = 0;
    virtual void get_sv(double* out) = 0;
    virtual void get_tables(double* coshOut, double* sinhOut) = 0;
    virtual void bmult(int op, double* A, uint32_t k2, uint32_t k1) = 0;
    virtual double update_in_slice(uint32_t k, int therm) = 0;
    virtual void sweep(int therm) = 0;
    virtual void get_scalars(double* out) = 0;
    virtual void set_r(double r) = 0;
    virtual void set_phi_delta(double d) = 0;
    virtual void rng_draw(int n, double* out) = 0;
    virtual void green_for_timeslice(uint32_t k, double* out) = 0;
    virtual void get_udv(uint32_t l, double* U, double* d, double* Vt) = 0;
    virtual void green_from_storage(uint32_t l_left, uint32_t l_right, double* out, double* sv) = 0;
    virtual void save_config_stream(const char* dir, int binary) = 0;
    virtual void attempt_wolff(int shift, double* stats) = 0;
    virtual void measured_sweep(double* obs) = 0;
    virtual void sweep_simple(int therm) = 0;
    virtual void measured_sweep_fermionic(double* scalars, double* vectors) = 0;
    virtual void shift_green_symmetric(double* out) = 0;
    virtual void dense_bmat(uint32_t k2, uint32_t k1, double* out) = 0;
};

template <int OPDIM, CheckerboardMethod CB = CB_ASSAAD_BERG>
struct SdwImpl : public SdwBase {
    typedef DetSDW<CB, OPDIM> Model;
    RngWrapper rng;
    std::unique_ptr<Model> rep;

    SdwImpl(const ref_sdw_params& p) : rng(p.seed, p.rngIndex), rep() {
        ModelParamsDetSDW pars;
#define SETP(name, value) { pars.name = (value); pars.specified.insert(#name); }
        SETP(opdim, (uint32_t)p.opdim);
        SETP(L, (uint32_t)p.L);
        SETP(m, (uint32_t)p.m);
        SETP(s, (uint32_t)p.s);
        SETP(dtau, p.dtau);
        SETP(r, p.r);
        SETP(c, p.c);
        SETP(u, p.u);
        SETP(lambda, p.lambda);
        SETP(txhor, p.txhor);
        SETP(txver, p.txver);
        SETP(tyhor, p.tyhor);
        SETP(tyver, p.tyver);
        SETP(cdwU, p.cdwU);
        SETP(mu, p.mu);
        SETP(accRatio, p.accRatio);
        SETP(weakZflux, p.weakZflux != 0);
        SETP(checkerboard, CB != CB_NONE);
        SETP(delaySteps, (uint32_t)p.delaySteps);
        SETP(globalShift, p.globalShift != 0);
        SETP(wolffClusterUpdate, p.wolffClusterUpdate != 0);
        SETP(wolffClusterShiftUpdate, p.wolffClusterShiftUpdate != 0);
        if (p.repeatWolffPerSweep > 1) {
            pars.repeatWolffPerSweep_string = std::to_string((long long)p.repeatWolffPerSweep);
            pars.specified.insert("repeatWolffPerSweep");
        }
        SETP(globalUpdateInterval, (uint32_t)p.globalUpdateInterval);
        SETP(repeatUpdateInSlice, (uint32_t)p.repeatUpdateInSlice);
        SETP(turnoffFermionMeasurements, p.fermionMeasurements == 0);
#undef SETP
        static const char* bcs[] = {"pbc", "apbc-x", "apbc-y", "apbc-xy"};
        pars.bc_string = bcs[p.bc];
        pars.specified.insert("bc");
        static const char* ums[] = {"iterative", "woodbury", "delayed"};
        pars.updateMethod_string = ums[p.updateMethod];
        pars.specified.insert("updateMethod");
        pars.spinProposalMethod_string = "box";
        pars.specified.insert("spinProposalMethod");
        DetModelLoggingParams lp;
        createReplica(rep, rng, pars, lp, std::string("/tmp"));
    }

    void dims(int32_t* out) {
        out[0] = (int32_t)rep->pars.N;
        out[1] = (int32_t)rep->sz;
        out[2] = (int32_t)rep->m;
        out[3] = (int32_t)rep->n;
        out[4] = (int32_t)rep->s;
        out[5] = OPDIM;
    }
    void get_phi(double* out) {
        std::memcpy(out, rep->phi.memptr(), sizeof(double) * rep->phi.n_elem);
    }
    void set_phi(const double* in) {
        std::memcpy(rep->phi.memptr(), in, sizeof(double) * rep->phi.n_elem);
        rep->updateCoshSinhTerms();
        rep->setupUdVStorage_and_calculateGreen();
    }
    void get_green(double* out) {
        std::memcpy(out, rep->g.memptr(), sizeof(cpx_t) * rep->g.n_elem);
    }
    void get_sv(double* out) {
        std::memcpy(out, rep->g_inv_sv.memptr(), sizeof(double) * rep->g_inv_sv.n_elem);
    }
    void get_tables(double* coshOut, double* sinhOut) {
        std::memcpy(coshOut, rep->coshTermPhi.memptr(), sizeof(double) * rep->coshTermPhi.n_elem);
        std::memcpy(sinhOut, rep->sinhTermPhi.memptr(), sizeof(double) * rep->sinhTermPhi.n_elem);
    }
    void bmult(int op, double* A, uint32_t k2, uint32_t k1) {
        const uint32_t D = rep->sz;
        typename Model::MatData a(reinterpret_cast<cpx_t*>(A), D, D);  // copies
        typename Model::MatData res;
        switch (op) {
        case 0: res = rep->checkerboardLeftMultiplyBmat(a, k2, k1); break;
        case 1: res = rep->checkerboardRightMultiplyBmat(a, k2, k1); break;
        case 2: res = rep->checkerboardLeftMultiplyBmatInv(a, k2, k1); break;
        case 3: res = rep->checkerboardRightMultiplyBmatInv(a, k2, k1); break;
        default: return;
        }
        std::memcpy(A, res.memptr(), sizeof(cpx_t) * D * D);
    }
    double update_in_slice(uint32_t k, int therm) {
        if (therm) rep->updateInSliceThermalization(k);
        else rep->updateInSlice(k);
        return rep->ad.lastAccRatioLocal_phi;
    }
    void sweep(int therm) {
        if (therm) rep->sweepThermalization();
        else rep->sweep(false);
    }
    void get_scalars(double* out) {
        out[0] = rep->ad.phiDelta;
        out[1] = (double)rep->performedSweeps;
        out[2] = (double)rep->currentTimeslice;
        out[3] = (double)(int)rep->lastSweepDir;
        out[4] = (double)rep->us.acceptedGlobalShifts;
        out[5] = (double)rep->us.attemptedGlobalShifts;
        out[6] = rep->ad.lastAccRatioLocal_phi;
        out[7] = rep->phiAction();
        out[8] = rep->get_exchange_action_contribution();
        out[9] = rep->get_exchange_parameter_value();
        out[10] = (double)rep->ad.accRatioLocal_box_RA.getSamplesAdded();
        out[11] = rep->ad.accRatioLocal_box_RA.get();
    }
    void set_r(double r) { rep->set_exchange_parameter_value(r); }
    void set_phi_delta(double d) { rep->ad.phiDelta = d; }
    void rng_draw(int n, double* out) {
        for (int i = 0; i < n; ++i) out[i] = rng.rand01();
    }
    void green_for_timeslice(uint32_t k, double* out) {
        typename Model::MatData g = rep->computeGreenFromScratch(k, rep->phi);
        std::memcpy(out, g.memptr(), sizeof(cpx_t) * g.n_elem);
    }
    void get_udv(uint32_t l, double* U, double* d, double* Vt) {
        auto& st = (*rep->UdVStorage)[0][l];
        std::memcpy(U, st.U.memptr(), sizeof(cpx_t) * st.U.n_elem);
        std::memcpy(d, st.d.memptr(), sizeof(double) * st.d.n_elem);
        std::memcpy(Vt, st.V_t.memptr(), sizeof(cpx_t) * st.V_t.n_elem);
    }
    void measured_sweep_fermionic(double* scalars, double* vectors) {
        // sweep(true) with fermionic measurements (detsdwopdim.cpp:508-900, 903-1000): scalars = greenK0, greenLocal,
        // occDiffSq, pairPlusMax, pairMinusMax; vectors = kOccX | kOccY | pairPlus | pairMinus (N each)
        rep->sweep(true);
        const uint32_t N = rep->pars.N;
        scalars[0] = rep->greenK0; scalars[1] = rep->greenLocal; scalars[2] = rep->occDiffSq;
        scalars[3] = rep->pairPlusMax; scalars[4] = rep->pairMinusMax;
        for (uint32_t i = 0; i < N; ++i) {
            vectors[i] = rep->kOccX[i]; vectors[N + i] = rep->kOccY[i];
            vectors[2 * N + i] = rep->pairPlus[i]; vectors[3 * N + i] = rep->pairMinus[i];
        }
    }
    void shift_green_symmetric(double* out) {
        typename Model::MatData gs = rep->shiftGreenSymmetric();
        std::memcpy(out, gs.memptr(), sizeof(cpx_t) * gs.n_elem);
    }
    void dense_bmat(uint32_t k2, uint32_t k1, double* out) {
        typename Model::MatData B = rep->computeBmatSDW(k2, k1);
        std::memcpy(out, B.memptr(), sizeof(cpx_t) * B.n_elem);
    }
    void sweep_simple(int therm) {
        // greenUpdate = simple (detsdwopdim.cpp:4366-4420)
        if (therm) rep->sweepSimpleThermalization();
        else rep->sweepSimple(false);
    }
    void measured_sweep(double* obs) {
        // sweep(takeMeasurements = true) with turnoffFermionMeasurements: the bosonic observables of
        // initMeasurements / measure / finishMeasurements (detsdwopdim.cpp:441-560, 903-918)
        rep->sweep(true);
        obs[0] = rep->normMeanPhi;
        obs[1] = rep->associatedEnergy;
        obs[2] = OPDIM == 2 ? rep->phiRhoS_Gs : 0.0;
        obs[3] = OPDIM == 2 ? rep->phiRhoS_Gc : 0.0;
    }
    void attempt_wolff(int shift, double* stats) {
        // attemptWolffClusterUpdate / attemptWolffClusterShiftUpdate (detsdwopdim.cpp:3487-3562, 3647-3748)
        if (shift) rep->attemptWolffClusterShiftUpdate();
        else rep->attemptWolffClusterUpdate();
        stats[0] = rep->us.attemptedWolffClusterUpdates;
        stats[1] = rep->us.acceptedWolffClusterUpdates;
        stats[2] = rep->us.attemptedWolffClusterShiftUpdates;
        stats[3] = rep->us.acceptedWolffClusterShiftUpdates;
        stats[4] = rep->us.addedWolffClusterSize;
    }
    void save_config_stream(const char* dir, int binary) {
        // the reference's own writers (detsdwopdim.cpp:4943-5036): append one configuration to the stream files
        if (binary) rep->saveConfigurationStreamBinary(dir);
        else rep->saveConfigurationStreamText(dir);
    }
    void green_from_storage(uint32_t l_left, uint32_t l_right, double* out, double* sv) {
        // G = [1 + UdV_r * UdV_l]^-1 evaluated by the reference's greenFromUdV (detmodel.h:768-818)
        typename Model::MatData g;
        VecNum svv;
        auto& st = (*rep->UdVStorage)[0];
        rep->greenFromUdV(g, svv, st[l_left], st[l_right]);
        std::memcpy(out, g.memptr(), sizeof(cpx_t) * g.n_elem);
        std::memcpy(sv, svv.memptr(), sizeof(double) * svv.n_elem);
    }
};

struct HubHandle {
    RngWrapper rng;
    std::unique_ptr<DetHubbard> rep;
    HubHandle(const ref_hub_params& p) : rng(p.seed, p.rngIndex), rep() {
        ModelParams<DetHubbard> pars;
#define SETP(name, value) { pars.name = (value); pars.specified.insert(#name); }
        SETP(L, (uint32_t)p.L);
        SETP(d, 2u);
        SETP(m, (uint32_t)p.m);
        SETP(s, (uint32_t)p.s);
        SETP(dtau, p.dtau);
        SETP(t, p.t);
        SETP(U, p.U);
        SETP(mu, p.mu);
        SETP(checkerboard, p.checkerboard != 0);
#undef SETP
        createReplica(rep, rng, pars);
    }
};

}  // namespace

extern "C" {

// ---------------------------------------------------------------- RNG (rngwrapper.h:43-119)
void* ref_rng_create(uint32_t seed, uint32_t index) {
    CoutSilencer q;
    return new RngWrapper(seed, index);
}
void ref_rng_destroy(void* h) { delete static_cast<RngWrapper*>(h); }
void ref_rng_draw(void* h, int n, double* out) {
    RngWrapper* r = static_cast<RngWrapper*>(h);
    for (int i = 0; i < n; ++i) out[i] = r->rand01();
}

// ---------------------------------------------------------------- DetSDW
void* ref_sdw_create(const ref_sdw_params* p) {
    CoutSilencer q;
    try {
        switch (p->opdim) {
        case 1: return static_cast<SdwBase*>(new SdwImpl<1>(*p));
        case 2: return p->denseHopping ? static_cast<SdwBase*>(new SdwImpl<2, CB_NONE>(*p))
                                       : static_cast<SdwBase*>(new SdwImpl<2>(*p));
        case 3: return static_cast<SdwBase*>(new SdwImpl<3>(*p));
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_sdw_create: %s\n", e.what());
    }
    return nullptr;
}
void ref_sdw_destroy(void* h) { delete static_cast<SdwBase*>(h); }
void ref_sdw_dims(void* h, int32_t* out) { static_cast<SdwBase*>(h)->dims(out); }
void ref_sdw_get_phi(void* h, double* out) { static_cast<SdwBase*>(h)->get_phi(out); }
void ref_sdw_set_phi(void* h, const double* in) { CoutSilencer q; static_cast<SdwBase*>(h)->set_phi(in); }
void ref_sdw_get_green(void* h, double* out) { static_cast<SdwBase*>(h)->get_green(out); }
void ref_sdw_get_sv(void* h, double* out) { static_cast<SdwBase*>(h)->get_sv(out); }
void ref_sdw_get_tables(void* h, double* c, double* s) { static_cast<SdwBase*>(h)->get_tables(c, s); }
// computeBmatSDW (detsdwopdim.cpp:1307-1497): the dense B(k2, k1) = prod e^{-dtau V_k} e^{-dtau K}
void ref_sdw_dense_bmat(void* h, uint32_t k2, uint32_t k1, double* out) {
    static_cast<SdwBase*>(h)->dense_bmat(k2, k1, out);
}
void ref_sdw_bmult(void* h, int op, double* A, uint32_t k2, uint32_t k1) {
    static_cast<SdwBase*>(h)->bmult(op, A, k2, k1);
}
double ref_sdw_update_in_slice(void* h, uint32_t k, int therm) {
    CoutSilencer q;
    return static_cast<SdwBase*>(h)->update_in_slice(k, therm);
}
int ref_sdw_sweep(void* h, int therm) {
    CoutSilencer q;
    try {
        static_cast<SdwBase*>(h)->sweep(therm);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_sdw_sweep: %s\n", e.what());
        return 1;
    }
    return 0;
}
void ref_sdw_get_scalars(void* h, double* out) { static_cast<SdwBase*>(h)->get_scalars(out); }
void ref_sdw_set_r(void* h, double r) { static_cast<SdwBase*>(h)->set_r(r); }
void ref_sdw_set_phi_delta(void* h, double d) { static_cast<SdwBase*>(h)->set_phi_delta(d); }
void ref_sdw_rng_draw(void* h, int n, double* out) { static_cast<SdwBase*>(h)->rng_draw(n, out); }
void ref_sdw_green_for_timeslice(void* h, uint32_t k, double* out) {
    CoutSilencer q;
    static_cast<SdwBase*>(h)->green_for_timeslice(k, out);
}
void ref_sdw_get_udv(void* h, uint32_t l, double* U, double* d, double* Vt) {
    static_cast<SdwBase*>(h)->get_udv(l, U, d, Vt);
}
void ref_sdw_green_from_storage(void* h, uint32_t ll, uint32_t lr, double* out, double* sv) {
    static_cast<SdwBase*>(h)->green_from_storage(ll, lr, out, sv);
}

// exchange probability, detsdwopdim.cpp:5251-5264
int ref_sdw_measured_sweep_fermionic(void* h, double* scalars, double* vectors) {
    CoutSilencer q;
    try {
        static_cast<SdwBase*>(h)->measured_sweep_fermionic(scalars, vectors);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_sdw_measured_sweep_fermionic: %s\n", e.what());
        return 1;
    }
    return 0;
}
void ref_sdw_shift_green_symmetric(void* h, double* out) { static_cast<SdwBase*>(h)->shift_green_symmetric(out); }
int ref_sdw_sweep_simple(void* h, int therm) {
    CoutSilencer q;
    try {
        static_cast<SdwBase*>(h)->sweep_simple(therm);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_sdw_sweep_simple: %s\n", e.what());
        return 1;
    }
    return 0;
}
int ref_sdw_measured_sweep(void* h, double* obs) {
    CoutSilencer q;
    try {
        static_cast<SdwBase*>(h)->measured_sweep(obs);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_sdw_measured_sweep: %s\n", e.what());
        return 1;
    }
}
void ref_sdw_attempt_wolff(void* h, int shift, double* stats) {
    CoutSilencer q;
    static_cast<SdwBase*>(h)->attempt_wolff(shift, stats);
}
void ref_sdw_save_config_stream(void* h, const char* dir, int binary) {
    static_cast<SdwBase*>(h)->save_config_stream(dir, binary);
}
double ref_sdw_exchange_probability(double par1, double action1, double par2, double action2) {
    return get_replica_exchange_probability<DetSDW<CB_ASSAAD_BERG, 2> >(par1, action1, par2, action2);
}

// ---------------------------------------------------------------- DetHubbard
void* ref_hub_create(const ref_hub_params* p) {
    CoutSilencer q;
    try {
        return new HubHandle(*p);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_hub_create: %s\n", e.what());
    }
    return nullptr;
}
void ref_hub_destroy(void* h) { delete static_cast<HubHandle*>(h); }
void ref_hub_dims(void* h, int32_t* out) {
    HubHandle* hh = static_cast<HubHandle*>(h);
    out[0] = (int32_t)hh->rep->N;
    out[1] = (int32_t)hh->rep->m;
    out[2] = (int32_t)hh->rep->n;
    out[3] = (int32_t)hh->rep->s;
}
void ref_hub_get_aux(void* h, int32_t* out) {
    HubHandle* hh = static_cast<HubHandle*>(h);
    for (arma::uword i = 0; i < hh->rep->auxfield.n_elem; ++i) out[i] = (int32_t)hh->rep->auxfield[i];
}
void ref_hub_get_green(void* h, int gc, double* out) {
    HubHandle* hh = static_cast<HubHandle*>(h);
    std::memcpy(out, hh->rep->green[gc].memptr(), sizeof(double) * hh->rep->green[gc].n_elem);
}
void ref_hub_get_sv(void* h, int gc, double* out) {
    HubHandle* hh = static_cast<HubHandle*>(h);
    std::memcpy(out, hh->rep->green_inv_sv[gc].memptr(), sizeof(double) * hh->rep->green_inv_sv[gc].n_elem);
}
void ref_hub_get_proptmat(void* h, double* out) {
    HubHandle* hh = static_cast<HubHandle*>(h);
    std::memcpy(out, hh->rep->proptmat.memptr(), sizeof(double) * hh->rep->proptmat.n_elem);
}
void ref_hub_bmat(void* h, int gc, uint32_t k2, uint32_t k1, double* out) {
    HubHandle* hh = static_cast<HubHandle*>(h);
    MatNum B = hh->rep->computeBmat(k2, k1, gc == 0 ? DetHubbard::Spin::Up : DetHubbard::Spin::Down);
    std::memcpy(out, B.memptr(), sizeof(double) * B.n_elem);
}
void ref_hub_update_in_slice(void* h, uint32_t k) {
    static_cast<HubHandle*>(h)->rep->updateInSlice(k);
}
int ref_hub_sweep(void* h, int therm) {
    CoutSilencer q;
    HubHandle* hh = static_cast<HubHandle*>(h);
    try {
        if (therm) hh->rep->sweepThermalization();
        else hh->rep->sweep(false);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_hub_sweep: %s\n", e.what());
        return 1;
    }
    return 0;
}
// sweep(true) (dethubbard.cpp:173-183): measure() after every slice, finishMeasurements at the end.
// scalars[8] = occUp, occDn, occTotal, occDouble, localMoment, eKinetic, ePotential, eTotal; zcorr[N]
int ref_hub_measured_sweep(void* h, double* scalars, double* zcorr) {
    CoutSilencer q;
    HubHandle* hh = static_cast<HubHandle*>(h);
    try {
        hh->rep->sweep(true);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_hub_measured_sweep: %s\n", e.what());
        return 1;
    }
    DetHubbard& r = *hh->rep;
    scalars[0] = r.occUp; scalars[1] = r.occDn; scalars[2] = r.occTotal; scalars[3] = r.occDouble;
    scalars[4] = r.localMoment; scalars[5] = r.eKinetic; scalars[6] = r.ePotential; scalars[7] = r.eTotal;
    for (arma::uword i = 0; i < r.zcorr.n_elem; ++i) zcorr[i] = r.zcorr[i];
    return 0;
}
void ref_hub_rng_draw(void* h, int n, double* out) {
    HubHandle* hh = static_cast<HubHandle*>(h);
    for (int i = 0; i < n; ++i) out[i] = hh->rng.rand01();
}
int ref_hub_current_timeslice(void* h) {
    return (int)static_cast<HubHandle*>(h)->rep->currentTimeslice;
}

}  // extern "C"
