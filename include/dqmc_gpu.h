/*
 * dqmc_gpu.h -- C ABI of the B200-native DQMC sweep hot path (libdqmc_b200.so).
 *
 * This is the lower seam described in SURVEY.md section 8(b): an `extern "C"` layer with an opaque
 * context, int status codes, plain pointers and sizes.  A context owns a BATCH of replicas that
 * live on one GPU (struct-of-arrays in HBM); every operator below acts on all replicas of the
 * batch at once unless it takes a `rep` index.  All calls are ordered on the context's CUDA stream
 * and a context must be driven from one host thread at a time (same contract as the reference:
 * one thread per replica, no re-entrancy; detqmc.h:183-218).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference's
 * src/ directory).  The reference has no FFI of its own -- models are C++ template parameters --
 * so "replaces" means: the C++ shim in detqmc_b200/csrc/detsdw_gpu.h implements the reference's
 * Model duck-type (detqmc.h:183-218, 441-491; detqmcpt.h:316, 857-900, 971-1115) by calling
 * these functions.  See INTEGRATION.md for the binding a maintainer would add.
 *
 * Matrix layout: column-major, complex numbers interleaved (re, im) as in Armadillo's
 * Mat<std::complex<double>>, so buffers can be handed to/from the reference without conversion.
 * Field layout: phi[(m+1)][OPDIM][N] doubles == memory of arma::Cube<double>(N, OPDIM, m+1)
 * (detsdwopdim.h:455-461); slice k = 0 is unused.
 *
 * Status codes: 0 = ok; non-zero = error, text available from dqmc_last_error().  The C++ shim
 * turns a non-zero status into the reference's GeneralError exception (exceptions.h).
 */
#ifndef DQMC_GPU_H_
#define DQMC_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dqmc_ctx dqmc_ctx;

enum { DQMC_OK = 0, DQMC_ERR_PARAM = 1, DQMC_ERR_CUDA = 2, DQMC_ERR_STATE = 3, DQMC_ERR_NUMERIC = 4 };

enum { DQMC_MODEL_SDW = 0, DQMC_MODEL_HUBBARD = 1 };

/* B-matrix operator selector for dqmc_bmat_mult*: B = B(k2, k1) = B_{k2} ... B_{k1+1}. */
enum {
    DQMC_OP_LEFT = 0,      /* B * A        leftMultiplyBmat     detsdwopdim.cpp:2074-2090 */
    DQMC_OP_RIGHT = 1,     /* A * B        rightMultiplyBmat    detsdwopdim.cpp:2305-2324 */
    DQMC_OP_LEFT_INV = 2,  /* B^-1 * A     leftMultiplyBmatInv  detsdwopdim.cpp:2170-2186 */
    DQMC_OP_RIGHT_INV = 3, /* A * B^-1     rightMultiplyBmatInv detsdwopdim.cpp:2404-2420 */
    DQMC_OP_LEFT_ADJ = 4   /* B^dagger * A (used by the left-chain stabilisation; no reference twin) */
};

/* Model parameters: the fields of ModelParamsDetSDW (detsdwparams.h:30-120) and
 * ModelParams<DetHubbard> (dethubbardparams.h:25-60) that the hot path reads.  The exchange
 * parameter r is per replica (dqmc_set_exchange_parameter). */
typedef struct dqmc_params {
    int32_t model;                /* DQMC_MODEL_* */
    int32_t opdim;                /* SDW: 1, 2, 3 */
    int32_t L;                    /* linear lattice size, N = L*L */
    int32_t m;                    /* number of time slices */
    int32_t s;                    /* stabilisation interval */
    int32_t bc;                   /* SDW: 0 pbc, 1 apbc-x, 2 apbc-y, 3 apbc-xy */
    int32_t weakZflux;            /* SDW: weakest magnetic flux (O(1), O(2) only) */
    int32_t delaySteps;           /* SDW: delayed-update block size (1 == Woodbury) */
    int32_t globalShift;          /* SDW: attempt global shift moves */
    int32_t globalUpdateInterval; /* SDW: every # sweeps */
    int32_t checkerboard;         /* Hubbard: checkerboard form of e^{-dtau T} (dethubbard.cpp:768-821) */
    int32_t denseHopping;         /* SDW: 1 = dense hopping exponential e^{-dtau K} (CheckerboardMethod CB_NONE, computeBmatSDW
                                   * detsdwopdim.cpp:1307-1497) instead of the checkerboard break-up */
    double dtau;
    double r;                     /* SDW: initial value of the exchange parameter for all replicas */
    double c, u, lambda;          /* SDW bosonic action + coupling */
    double txhor, txver, tyhor, tyver;
    double mux, muy;              /* SDW chemical potentials (the reference sets both to mu) */
    double accRatio;              /* SDW: target acceptance for the box-size adaptation */
    double t, U, mu;              /* Hubbard */
    int32_t wolffClusterUpdate;       /* SDW: Wolff single-cluster moves every globalUpdateInterval sweeps */
    int32_t wolffClusterShiftUpdate;  /* SDW: combined cluster + global shift move (excludes the two others) */
    int32_t repeatWolffPerSweep;      /* SDW: clusters per attempt (0 or 1: one) */
    int32_t repeatUpdateInSlice;      /* SDW: passes over a slice per updateInSlice (0 or 1: one), detsdwopdim.cpp:2438 */
} dqmc_params;

/* Per-replica control data that follows the exchange parameter in a replica exchange
 * (UpdateStatistics + AdjustmentData, detsdwopdim.h:283-300, 479-560; blob exchanged by
 * get/set_control_data, detsdwopdim.cpp:5218-5242). */
typedef struct dqmc_control_data {
    double phiDelta;
    double lastAccRatioLocal_phi;
    double ra_average;            /* RunningAverage state of accRatioLocal_box_RA */
    int32_t ra_samples_added;
    int32_t ra_count;             /* number of valid entries in ra_values */
    double ra_values[100];
    uint32_t acceptedGlobalShifts;
    uint32_t attemptedGlobalShifts;
    /* Wolff-cluster statistics of UpdateStatistics (detsdwopdim.h:283-300): like the counters above they belong to
     * the control PARAMETER, not to the replica, so they travel with the blob through an exchange and a checkpoint */
    uint32_t acceptedWolffClusterUpdates;
    uint32_t attemptedWolffClusterUpdates;
    uint32_t acceptedWolffClusterShiftUpdates;
    uint32_t attemptedWolffClusterShiftUpdates;
    double addedWolffClusterSize;
} dqmc_control_data;

/* ---- lifecycle ------------------------------------------------------------------------------ */

/* Replaces createReplica(...) (detsdwopdim.cpp:48-84, dethubbard.cpp:30-43) for a batch of
 * n_replicas on CUDA device `device`.  Fails (never falls back to a CPU path) when no device. */
int dqmc_create(const dqmc_params* params, int n_replicas, int device, dqmc_ctx** out);
void dqmc_destroy(dqmc_ctx* ctx);
const char* dqmc_last_error(const dqmc_ctx* ctx);
/* Use an existing CUDA stream (cudaStream_t) for all work of this context; NULL = own (non-blocking) stream.
 * Note that the handle of CUDA's legacy default stream is NULL as well: a host program that wants its own copies,
 * collectives or timing events ordered with the library's kernels must pass an explicitly created stream. */
int dqmc_set_stream(dqmc_ctx* ctx, void* cuda_stream);
int dqmc_synchronize(dqmc_ctx* ctx);
/* out[0..7] = {N, D (Green's-function dimension), m, n, s, n_green_components, n_replicas, opdim} */
int dqmc_dims(const dqmc_ctx* ctx, int32_t* out);
/* Numerical options (no reference twin).  DQMC_OPT_STABILIZER selects the factorisation behind the
 * UDT chains (udvDecompose's role, udv.h:68-90):
 *   DQMC_STAB_PREPIVOT_BLOCKED (default)  columns ordered once by decreasing norm, then blocked
 *                                         Householder QR with tensor-core trailing updates;
 *   DQMC_STAB_FULL_PIVOT                  Householder QR with full column pivoting, one CTA per matrix
 *                                         (slow; kept as the cross-check of the tests).
 * DQMC_OPT_LANES (1..64; default one lane per replica, at most 64): the replicas of a context are split into
 * lanes whose sweeps are issued on separate CUDA streams (inside one CUDA graph).  Replicas of a lane advance in
 * lockstep (a delayed-update round or a QR panel takes as long as its slowest replica); separate lanes remove
 * that coupling and let the latency-bound kernels of one replica (sequential update rounds, QR panels) overlap
 * with the throughput-bound kernels of the others.  Results do not depend on it. */
enum { DQMC_OPT_STABILIZER = 0, DQMC_OPT_LANES = 1 };
enum { DQMC_STAB_PREPIVOT_BLOCKED = 0, DQMC_STAB_FULL_PIVOT = 1 };
int dqmc_set_option(dqmc_ctx* ctx, int option, int value);
/* Number of kernels launched by this context so far (for bench.py's gpu_launches). */
uint64_t dqmc_launch_count(const dqmc_ctx* ctx);

/* ---- random numbers: RngWrapper (rngwrapper.h:43-119) ---------------------------------------- */

/* Seed replica `rep`'s stream exactly like RngWrapper(seed, processIndex) (rngwrapper.cpp:30-50). */
int dqmc_rng_seed(dqmc_ctx* ctx, int rep, uint32_t seed, uint32_t process_index);
/* Alternatively feed the stream from the host program's own generator (e.g. the reference's
 * RngWrapper): fill(user, out, n) must write the next n rand01() values. */
typedef void (*dqmc_rng_fill_fn)(void* user, double* out, size_t n);
int dqmc_rng_set_source(dqmc_ctx* ctx, int rep, dqmc_rng_fill_fn fill, void* user);
/* Consume n values from the head of replica rep's stream (rand01()). */
int dqmc_rng_draw(dqmc_ctx* ctx, int rep, size_t n, double* out);
/* Look at the next n values without consuming them / consume n values. */
int dqmc_rng_peek(dqmc_ctx* ctx, int rep, size_t n, double* out);
int dqmc_rng_skip(dqmc_ctx* ctx, int rep, size_t n);
/* Checkpointing of a stream fed by dqmc_rng_set_source: the values already drawn from the driver's generator but not
 * yet consumed are part of the replica's state (the reference saves model and generator together, detqmc.h:316-356).
 * dqmc_rng_look_ahead: out == NULL -> *n = number of such values; else copy up to *n of them (oldest first).
 * dqmc_rng_set_look_ahead: after a resume (the driver has restored its generator), discard whatever the replica
 * drew from the superseded generator state and install the saved values as the head of the stream. */
int dqmc_rng_look_ahead(dqmc_ctx* ctx, int rep, double* out, size_t* n);
int dqmc_rng_set_look_ahead(dqmc_ctx* ctx, int rep, const double* values, size_t n);
/* performedSweeps of the model state (detsdwopdim.h:1127-1148): restores the phase of the global-move schedule
 * (performedSweeps % globalUpdateInterval, detmodel.h:1422-1424) after loadContents. */
int dqmc_set_performed_sweeps(dqmc_ctx* ctx, uint32_t n);
/* Total number of values consumed from replica rep's stream so far. */
uint64_t dqmc_rng_consumed(const dqmc_ctx* ctx, int rep);
/* Context-free: the first n rand01() values of RngWrapper(seed, process_index) (rngwrapper.cpp:30-50,
 * rngwrapper.h:54-57).  Host only; lets a host program (and the CPU tests) check that its own
 * generator and this library's restatement of dSFMT-19937 produce the same stream. */
int dqmc_rng_stream_sample(uint32_t seed, uint32_t process_index, size_t n, double* out);

/* ---- state ----------------------------------------------------------------------------------- */

/* setupRandomField (detsdwopdim.cpp:1098-1113) / setupRandomAuxfield (dethubbard.cpp:741-751):
 * draws the initial configuration from the replica's stream in the reference's order. */
int dqmc_init_random_fields(dqmc_ctx* ctx, int rep);
/* phi (SDW; doubles) or auxfield (Hubbard; int32 +-1, layout [(m+1)][N]).  Host pointers. */
int dqmc_upload_fields(dqmc_ctx* ctx, int rep, const void* fields);
int dqmc_download_fields(dqmc_ctx* ctx, int rep, void* fields);
/* getCurrentSystemConfiguration / saveConfigurationStreamBinary (detsdwopdim.cpp:5116-5122, 4991-5012;
 * DetSDW_SystemConfig::write_to_disk_phi_binary, detsdwsystemconfig.cpp:123-134): the fields of replica `rep` in
 * the order of the reference's configuration streams, out[((ix*L + iy)*m + (k-1))*opdim + dim] = phi(iy*L + ix, dim, k)
 * for k = 1..m (N*m*opdim doubles, reordered on the device); rep = -1: all replicas, replica-major (what the
 * parallel-tempering driver gathers per exchange-parameter index, detqmcpt.h:702-757). */
int dqmc_download_config_stream(dqmc_ctx* ctx, int rep, double* out);
/* g / green[gc] (detmodel.h:462): D*D values (complex interleaved for SDW, real for Hubbard). */
int dqmc_download_green(dqmc_ctx* ctx, int rep, int gc, double* out);
int dqmc_upload_green(dqmc_ctx* ctx, int rep, int gc, const double* in);
/* get/set_exchange_parameter_value (detsdwopdim.cpp:5189-5197). */
int dqmc_set_exchange_parameter(dqmc_ctx* ctx, int rep, double r);
int dqmc_get_exchange_parameter(dqmc_ctx* ctx, int rep, double* r);
/* get/set_control_data (detsdwopdim.cpp:5218-5242). */
int dqmc_get_control_data(dqmc_ctx* ctx, int rep, dqmc_control_data* out);
int dqmc_set_control_data(dqmc_ctx* ctx, int rep, const dqmc_control_data* in);
/* out = {currentTimeslice, lastSweepDir (+1 up, -1 down), performedSweeps} (detmodel.h:463,481). */
int dqmc_get_sweep_state(const dqmc_ctx* ctx, int32_t* out);

/* ---- operators (each one is parity-tested in isolation) -------------------------------------- */

/* checkerboard{Left,Right}MultiplyBmat[Inv](A, k2, k1) on a host matrix of replica rep, in place
 * (detsdwopdim.cpp:2074-2420); Hubbard: the dense functors of dethubbard.h:281-337. */
int dqmc_bmat_mult(dqmc_ctx* ctx, int rep, int gc, int op, double* A_host, uint32_t k2, uint32_t k1);
/* Same on device memory for the whole batch: A_dev = [n_replicas][D*D] values, in place. */
int dqmc_bmat_mult_device(dqmc_ctx* ctx, int gc, int op, void* A_dev, uint32_t k2, uint32_t k1);
/* Batched device-to-device copy of the operator's input for benchmarking: times `reps`
 * back-to-back launches of op on A_dev and returns the average kernel time in ms. */
int dqmc_bench_bmat_mult(dqmc_ctx* ctx, int op, void* A_dev, uint32_t k2, uint32_t k1, int reps,
                         float* ms_per_launch);

/* setupUdVStorage_and_calculateGreen (detmodel.h:678-713): rebuilds the stabilisation storage
 * from the fields and computes G(beta) for every replica; currentTimeslice = m, lastSweepDir = Up. */
int dqmc_setup_storage(dqmc_ctx* ctx);
/* wrapUpGreen(k): G(k+1) = B_{k+1} G(k) B_{k+1}^-1 (detmodel.h:1234-1259). */
int dqmc_wrap_up(dqmc_ctx* ctx, uint32_t k);
/* wrapDownGreen(k): G(k-1) = B_k^-1 G(k) B_k (detmodel.h:1064-1095). */
int dqmc_wrap_down(dqmc_ctx* ctx, uint32_t k);
/* advanceUpGreen(l) / advanceDownGreen(l) (detmodel.h:1106-1163 / 953-1017). */
int dqmc_advance_up(dqmc_ctx* ctx, uint32_t l);
int dqmc_advance_down(dqmc_ctx* ctx, uint32_t l);
/* max_ij |G_wrapped - G_advanced| recorded by the last advance (the reference's
 * --logGreenConsistency check, detmodel.h:993-1010); one value per replica. */
int dqmc_get_green_consistency(dqmc_ctx* ctx, double* out);
/* log|det G^-1| of replica rep's current G as computed by the last from-scratch evaluation;
 * equals sum_j log(green_inv_sv[j]) of the reference (detsdwopdim.cpp:3613-3620). */
int dqmc_logdet(dqmc_ctx* ctx, int rep, int gc, double* out);
/* G at an arbitrary slice from scratch into a host buffer, leaving the sweep state untouched
 * (computeGreenFromScratch, detsdwopdim.cpp:4905-4933). */
int dqmc_green_for_timeslice(dqmc_ctx* ctx, int rep, int gc, uint32_t k, double* out);
/* sweepSimple(false) / sweepSimpleThermalization() (greenUpdate = simple; detmodel.h:718-758,
 * detsdwopdim.cpp:4366-4420): for every slice k = 1..m the Green's function is recomputed from scratch and the slice is
 * updated.  The reference inverts the plain product 1 + B(k,0) B(beta,k) built from the DENSE hopping exponential
 * (computeBmatSDW, even with checkerboard = true); here the stabilised evaluation with the checkerboard B of the
 * regular sweeps serves it, so G differs from the reference's by the O(dtau^2) break-up error and equals the G of
 * dqmc_sweep.  O(m) from-scratch evaluations per sweep: a consistency tool for
 * small systems.  The UDT storage of the stabilised sweeps is not maintained; call dqmc_setup_storage before going back
 * to dqmc_sweep. */
int dqmc_sweep_simple(dqmc_ctx* ctx, int thermalization);
/* Numerical kernel under greenFromUdV (detmodel.h:768-818): G = [1 + M_r M_l]^-1 for two host
 * D x D matrices given as M_r = Q_r diag(d_r) T_r and M_l = T_l^dagger diag(d_l) Q_l^dagger
 * (all column-major; see DESIGN.md "UDT").  Test surface. */
int dqmc_green_from_udt_host(dqmc_ctx* ctx, const double* Qr, const double* dr, const double* Tr,
                             const double* Ql, const double* dl, const double* Tl,
                             double* G_out, double* logdet_out);
/* Pivoted-QR "UDT" factorisation of a host D x D matrix: M = Q diag(d) T (udvDecompose's role,
 * udv.h:68-90).  Test surface. */
int dqmc_udt_decompose_host(dqmc_ctx* ctx, const double* M, double* Q, double* d, double* T);
/* Batched C = op(A) op(B) on host matrices through the FP64 tensor-core GEMM.  Test surface.
 * transa/transb: 0 = N, 1 = conjugate transpose. */
int dqmc_gemm_host(dqmc_ctx* ctx, int transa, int transb, int M, int N, int K,
                   const double* A, const double* B, double* C);

/* ---- Monte Carlo updates ---------------------------------------------------------------------- */

/* updateInSlice(k) / updateInSliceThermalization(k) for every replica
 * (detsdwopdim.cpp:2427-2489, 3021-3175, 3293-3375; dethubbard.cpp:141-171).  Random numbers are
 * taken from each replica's stream in the reference's order; n_accepted (may be NULL) receives
 * the per-replica number of accepted proposals. */
int dqmc_update_slice(dqmc_ctx* ctx, uint32_t k, int thermalization, uint32_t* n_accepted);
/* attemptGlobalShiftMove() (detsdwopdim.cpp:3564-3645) for every replica; requires
 * currentTimeslice == m.  accepted (may be NULL): per-replica 0/1. */
int dqmc_global_shift_move(dqmc_ctx* ctx, int32_t* accepted);
/* attemptWolffClusterUpdate() (with_shift = 0, detsdwopdim.cpp:3487-3562) / attemptWolffClusterShiftUpdate()
 * (with_shift = 1, :3647-3748) for every replica; requires currentTimeslice == m.  The cluster is grown on the fields
 * on the host with the replica's random-number stream in the reference's order (buildAndFlipCluster, :3805-3883),
 * the re-setup of the UDT storage and the Green's function runs batched on the device.  dqmc_sweep calls them from
 * globalMove (:3460-3485) when the parameters ask for them.  accepted (may be NULL): per-replica 0/1. */
int dqmc_wolff_cluster_move(dqmc_ctx* ctx, int with_shift, int32_t* accepted);
/* UpdateStatistics of the cluster moves (detsdwopdim.h:285-299): out[0..4] = attemptedWolffClusterUpdates,
 * acceptedWolffClusterUpdates, attemptedWolffClusterShiftUpdates, acceptedWolffClusterShiftUpdates,
 * addedWolffClusterSize. */
int dqmc_get_wolff_statistics(dqmc_ctx* ctx, int rep, double* out);
/* phiAction() (detsdwopdim.cpp:4242-4299) per replica. */
int dqmc_phi_action(dqmc_ctx* ctx, double* out);

/* sweep(takeMeasurements) / sweepThermalization() for every replica of the batch
 * (detsdwopdim.cpp:4422-4502 -> detmodel.h:1401-1478): one direction (down or up) including
 * the global move before a down-sweep.  thermalization: 0 = sweep(false), 1 = sweepThermalization() (step-size
 * adaptation), 2 = sweep(true) for DetSDW: measure(k) after the update of every slice (detmodel.h:1277-1283) --
 * the Green's function is shifted symmetrically by half a hopping step (shiftGreenSymmetric, detsdwopdim.cpp:4505-4612)
 * and the fermionic observables of DetSDW::measure (:540-900) are accumulated on the device. */
int dqmc_sweep(dqmc_ctx* ctx, int thermalization);
/* finishMeasurements (detsdwopdim.cpp:903-1000) for replica `rep` after dqmc_sweep(ctx, 2):
 * scalars[5] = greenK0, greenLocal, occDiffSq, pairPlusMax, pairMinusMax;
 * vectors[4 N] = kOccX | kOccY | pairPlus | pairMinus (momentum-space occupation per band, equal-time pairing
 * correlations between site 0 and site i). */
int dqmc_get_fermionic_observables(dqmc_ctx* ctx, int rep, double* scalars, double* vectors);
/* DetHubbard: dqmc_sweep(ctx, 2) accumulates DetHubbard::measure (dethubbard.cpp:511-539) after the update of every
 * slice; finishMeasurements (dethubbard.cpp:601-612) for replica `rep`:
 * scalars[8] = occupationUp, occupationDown, totalOccupation, doubleOccupation, localMoment, kineticEnergy,
 * potentialEnergy, totalEnergy; zcorr[N] = spinzCorrelationFunction. */
int dqmc_get_hubbard_observables(dqmc_ctx* ctx, int rep, double* scalars, double* zcorr);

/* Resident random numbers: upload the next n_sweeps sweeps' worth of every replica's stream once;
 * dqmc_sweep then runs without per-sweep host<->device copies or synchronisation (the consumption
 * cursors live on the device).  dqmc_rng_release reads the cursors back and advances the host
 * streams; it must be called before anything else consumes from them (dqmc_rng_draw ...). */
int dqmc_rng_preload(dqmc_ctx* ctx, int n_sweeps);
int dqmc_rng_release(dqmc_ctx* ctx);

/* ---- measurement hooks (no reference twin; the reference's -DTIMING timers, timing.h:32-78) ---- */
#define DQMC_PROF_NCAT 10
/* Bracket every kernel launch with CUDA events on the context's stream and accumulate device
 * time per kernel family. */
int dqmc_profile_enable(dqmc_ctx* ctx, int on);
int dqmc_profile_get(dqmc_ctx* ctx, double* ms /*[DQMC_PROF_NCAT]*/, uint64_t* counts /*[DQMC_PROF_NCAT]*/);
const char* dqmc_profile_name(int category);
/* Running total of accepted local updates per replica (flop accounting of the delayed updates). */
int dqmc_accepted_total(dqmc_ctx* ctx, uint64_t* out);

/* ---- replica exchange ------------------------------------------------------------------------- */

/* get_exchange_action_contribution() (detsdwopdim.cpp:5204-5216) of every local replica.
 * actions_dev (may be NULL): device buffer of n_replicas doubles -- the NCCL all-gather send
 * buffer; actions_host (may be NULL): host copy. */
int dqmc_exchange_actions(dqmc_ctx* ctx, double* actions_dev, double* actions_host);
/* One-shot packing of everything replicaExchangeStep gathers (detqmcpt.h:968-1012) into a device
 * buffer that can be handed to ncclAllGather as is:
 *   payload[0 .. R)                      exchange action of every local replica
 *   payload[R .. R+n_uniforms)           the next n_uniforms values of LOCAL replica 0's stream
 *                                        (look-ahead only; meaningful on the rank that owns ladder
 *                                        replica 0)
 *   payload[R+n_uniforms .. +R*105)      the control-data blobs (dqmc_control_data, 840 bytes each)
 * No host synchronisation. */
int dqmc_exchange_pack(dqmc_ctx* ctx, double* payload_dev, int n_uniforms);
/* The collective of the path (replaces the two mpi::gather of detqmcpt.h:971-1012).  dqmc_set_comm injects
 * the host's NCCL communicator (an ncclComm_t, passed as void*; nranks / rank its geometry); the library resolves
 * ncclAllGather from the NCCL the process has loaded, it does not link NCCL itself.  dqmc_exchange_allgather packs the
 * local payload (layout of dqmc_exchange_pack, dqmc_exchange_payload_len doubles) into slot `rank` of gathered_dev
 * [nranks * len] and all-gathers in place on the context's stream; gathered_host (may be NULL) receives a copy after
 * the stream has drained.  Without a communicator (single process) it is dqmc_exchange_pack + the optional copy. */
int dqmc_set_comm(dqmc_ctx* ctx, void* nccl_comm, int nranks, int rank);
int dqmc_exchange_payload_len(dqmc_ctx* ctx, int n_uniforms);
int dqmc_exchange_allgather(dqmc_ctx* ctx, int n_uniforms, double* gathered_dev, double* gathered_host);
/* Scatter side of replicaExchangeStep (detqmcpt.h:1082-1115): install the new exchange parameter
 * and control data of every local replica and consume n_uniforms_used values from local replica
 * 0's stream (pass 0 on ranks that do not own ladder replica 0). */
int dqmc_exchange_apply(dqmc_ctx* ctx, const double* r_new, const dqmc_control_data* ctrl_new,
                        int n_uniforms_used);
/* get_replica_exchange_probability (detsdwopdim.cpp:5251-5264). */
double dqmc_exchange_probability(double par1, double action1, double par2, double action2);
/* Serial ladder walk of DetQMCPT::replicaExchangeStep on the gathered actions
 * (detqmcpt.h:1031-1079).  n = ladder length; par_process / process_par are updated in place;
 * uniforms[0..n-2] are the next values of replica 0's stream (dqmc_rng_peek), *n_used returns how
 * many were consumed (dqmc_rng_skip that many on the owner); swapped[cpi] = 1 if the pair
 * (cpi, cpi+1) was exchanged. */
int dqmc_exchange_walk(int n, const double* control_values, int32_t* par_process,
                       int32_t* process_par, const double* actions, const double* uniforms,
                       int32_t* n_used, int32_t* swapped);

#ifdef __cplusplus
}
#endif

#endif /* DQMC_GPU_H_ */
