// Internal declarations shared by the CUDA translation units of libdqmc_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/dqmc_gpu.h"
#include "rng_stream.h"

#define DQMC_MAX_LANES 64

namespace dqmc {
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  A sweep is a chain of ~3500 small dependent kernels per replica; the hand-over from
// one kernel to the next (2-3 us inside a CUDA graph) is a sizeable part of the step.  Every launch carries the
// programmatic-stream-serialization attribute and every kernel starts with pdl_enter() = `griddepcontrol.wait`: the
// successor's launch is processed while the predecessor still runs and its CTAs start the moment the predecessor's
// last CTA has exited, waiting in hardware only for the predecessor's memory to be visible.  The predecessor does NOT
// trigger its dependents early (`griddepcontrol.launch_dependents` at kernel start, the first version, -DDQMC_PDL_EARLY):
// early CTAs sit on shared memory and registers for the whole run time of a long predecessor (a QR panel, a window
// round), which cost more than the overlap gained from 8 replicas per GPU on (measured, same box: 8 replicas 55.4 ->
// 52.7 ms per step, 16: 63.5 -> 57.4, 64: 119.9 -> 117.8 against no attribute; 1 replica 47.5 vs 47.8).
// ------------------------------------------------------------------------------------------------
bool pdl_enabled();
void pdl_set_enabled(bool on);
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef DQMC_PDL_EARLY
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// the same with a thread-block cluster of `cluster_x` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                      unsigned cluster_x, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_x;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

typedef double2 cplx;

// ------------------------------------------------------------------------------------------------
// Checkerboard / B-matrix multiply (cb_kernels.cu)
// ------------------------------------------------------------------------------------------------
struct CbGeom {
    int L, N, msf, D, nplaq, opdim;
    int m;                 // slices
    double lambda_dtau;    // lambda * dtau (only used when tables are rebuilt on device)
};

// One launch applies a chain of single-slice factors to every vector (column or row) of a batch
// of D x D matrices.
struct CbLaunch {
    cplx* A;               // [batch][D*D] in place
    long long strideA;     // elements between consecutive matrices
    const double* phi;     // [batch][(m+1)*opdim*N]
    const double* coshT;   // [batch][(m+1)*N]
    const double* sinhT;   // [batch][(m+1)*N]
    long long stridePhi, strideTab;
    const cplx* cbtab;     // Hermitian-compressed plaquette tables (cb_build_tables)
    int real_tables;       // 1: all plaquette matrices are real (no magnetic flux)
    int kfirst, kstep, kcount;   // slices kfirst, kfirst+kstep, ...
    int rows;              // 0: vectors are columns (left multiply), 1: vectors are rows (right multiply)
    int k_then_v;          // 1: hopping stage first, then potential stage; 0: the other order
    int sign_idx;          // 0: e^{-dtau .} (B), 1: e^{+dtau .} (B^-1)
    int transposed;        // 0: M v, 1: M^T v (right multiplies act on rows)
    const double* colscale;  // optional [batch][D]: scale vector v (column v) by colscale[v] on store
    long long strideScale;
    int batch;
    int skip_hopping = 0;    // 1: potential stage only (the hopping factor is applied as a dense GEMM: CB_NONE path)
    int shift = 0;           // 1: half-step stage of shiftGreenSymmetric (E1 then E0 at +-dtau/2, no potential) instead of B
    cplx* out = nullptr;     // optional output matrices (default: in place)
    long long strideOut = 0;
};

int cb_table_count(const CbGeom& g);                    // number of cplx in the table
// host-side construction of the plaquette tables from the model parameters
void cb_build_tables(const dqmc_params& p, std::vector<cplx>& out);
void cb_build_dense_propagators(const dqmc_params& p, int msf, std::vector<cplx>& P, std::vector<cplx>& Pinv);
void cb_build_shift_matrices(const dqmc_params& p, int msf, std::vector<cplx>& SL, std::vector<cplx>& SR);
cudaError_t cb_launch(const CbGeom& g, const CbLaunch& a, cudaStream_t st);

// dense elementwise helpers (misc_kernels.cu)
cudaError_t launch_set_identity(cplx* A, int D, long long stride, int batch, cudaStream_t st);
cudaError_t launch_conj_transpose(const cplx* A, cplx* B, int D, long long stride, int batch, cudaStream_t st);
cudaError_t launch_scaled_conj_transpose(const cplx* A, long long strideA, cplx* B, long long strideB,
                                         const double* rowscale, long long strideS, int D, int batch, cudaStream_t st);
cudaError_t launch_max_abs_diff(const cplx* A, const cplx* B, int D, long long stride, int batch,
                                double* out, cudaStream_t st);
cudaError_t launch_update_tables(const double* phi, double* coshT, double* sinhT, int N, int opdim, int m,
                                 double lambda_dtau, long long stridePhi, long long strideTab, int batch,
                                 cudaStream_t st);
cudaError_t launch_shift_fields(double* phi, const double* shift, int N, int opdim, int m,
                                long long stridePhi, int batch, cudaStream_t st);
cudaError_t launch_phi_action(const double* phi, const double* rvals, double* out, int L, int opdim, int m,
                              double dtau, double c, double u, long long stridePhi, int batch,
                              cudaStream_t st);
cudaError_t launch_fermion_measure(const cplx* gs, long long strideG, int N, int L, int msf, double* acc, long long strideAcc,
                                   int batch, cudaStream_t st);
cudaError_t launch_config_stream(const double* phi, double* out, int L, int opdim, int m, long long stridePhi,
                                 long long strideOut, int batch, cudaStream_t st);
cudaError_t launch_exchange_action(const double* phi, double* out, int N, int opdim, int m, double dtau,
                                   long long stridePhi, int batch, cudaStream_t st);

cudaError_t launch_exchange_pack(const double* rng, const int* cursor, int window, double* uni_out, int n_uni,
                                 const dqmc_control_data* ctrl, double* ctrl_out, int R, int copy_uniforms,
                                 cudaStream_t st);
cudaError_t launch_cursor_advance(int* cursor, int rep, int n, cudaStream_t st);
cudaError_t launch_cursor_add(int* cursor, const int* add, int n, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// Batched complex GEMM on FP64 tensor cores (gemm_kernels.cu)
//   C = alpha * rowscale .* (op(A) * diag(kscale) * op(B)) .* colscale + beta * C
// ------------------------------------------------------------------------------------------------
struct GemmArgs {
    int M, N, K;
    int transa, transb;           // 0: N, 1: conjugate transpose
    const cplx* A; int lda; long long strideA;
    const cplx* B; int ldb; long long strideB;
    cplx* C; int ldc; long long strideC;
    const double* rowscale; long long strideRow;   // optional, length M
    const double* colscale; long long strideCol;   // optional, length N
    const double* kscale; long long strideK;       // optional, length K
    double alpha;                 // scalar factor of the product
    int b_kmajor;                 // rank-update path only: B element (k, n) is stored at B[k*ldb + n]
    const int* kvec;              // optional per-matrix K (<= K; rank-update path only), device pointer
    double beta;                  // 0 or 1
    int batch;
};
cudaError_t gemm_launch(const GemmArgs& g, cudaStream_t st);
int gemm_matrices_in_flight();
void gemm_set_matrices_in_flight(int n);   // tile-shape hint: matrices of all lanes that are in flight on the device

// ------------------------------------------------------------------------------------------------
// Pivoted Householder QR and friends (qr_kernels.cu)
// ------------------------------------------------------------------------------------------------
// In place: A -> R in the upper triangle, Householder vectors below; tau[D]; perm[D] with
// A_in[:, perm[j]] = (Q R)[:, j].
cudaError_t qrcp_factor_launch(cplx* A, int D, long long strideA, cplx* tau, int* perm, double* colnorm,
                               int batch, cudaStream_t st);
// Q (explicit D x D) from the factored A / tau.
cudaError_t qr_form_q_launch(const cplx* A, const cplx* tau, cplx* Q, int D, long long strideA, int batch,
                             cudaStream_t st);
// d[i] = |R[i,i]|,  T[i, perm[j]] = R[i,j] / d[i] (j >= i), 0 elsewhere.
cudaError_t qr_extract_dt_launch(const cplx* A, const int* perm, double* d, cplx* T, int D, long long strideA,
                                 int batch, cudaStream_t st);
// Solve R Z = Y for upper-triangular R (from a factored A), writing row j of Z to row perm[j] of Zout
// (perm may be null = identity).  Y is overwritten.
cudaError_t trsm_upper_launch(const cplx* A, cplx* Y, cplx* Zout, const int* perm, int D, long long strideA,
                              int batch, cudaStream_t st);
// ---- blocked QR with column pre-pivoting (compact WY, trailing updates on the DMMA GEMM) ----------
struct QrWorkspace {
    int D, batch, nb;
    cplx* V;       // [batch][D*D] explicit unit-lower-trapezoidal reflectors of the last factorisation
    cplx* VT;      // [batch][D*D] V * T of every panel
    cplx* W;       // [batch][nb*D] panel products
    cplx* Rinv;    // [batch][nb*(D+nb)] inverted diagonal blocks of R
    uint64_t launches;   // kernels launched by the drivers below (for dqmc_launch_count)
};
int qr_choose_nb(int D);
cudaError_t qr_workspace_create(QrWorkspace* ws, int D, int batch);
void qr_workspace_destroy(QrWorkspace* ws);
// perm = columns by decreasing norm (ties by index); Aout[:, j] = A[:, perm[j]]
cudaError_t qr_prepivot_launch(const cplx* A, long long strideA, cplx* Aout, long long strideOut, int* perm,
                               double* norms, int D, int batch, cudaStream_t st);
// out[perm[j], :] = in[j, :]
cudaError_t permute_rows_launch(const cplx* in, long long strideIn, cplx* out, long long strideOut, const int* perm,
                                int D, int batch, cudaStream_t st);
// `off` = index of the first matrix of this call inside the workspace
cudaError_t qr_blocked_factor(QrWorkspace& ws, cplx* A, int D, long long strideA, int off, int batch, cudaStream_t st);
cudaError_t qr_blocked_form_q(QrWorkspace& ws, cplx* Q, int D, long long strideQ, int off, int batch, cudaStream_t st);
cudaError_t qr_blocked_apply_qh(QrWorkspace& ws, cplx* C, int D, int ncols, long long strideC, int off, int batch,
                                cudaStream_t st);
cudaError_t trsm_upper_blocked(QrWorkspace& ws, const cplx* A, cplx* Y, cplx* Zout, int D, long long strideA, int off,
                               int batch, cudaStream_t st);
// split scales: big[i] = 1/max(d,1), small[i] = min(d,1); also sum log max(d,1) -> logacc (+=)
cudaError_t scale_split_launch(const double* d, double* inv_big, double* small_, double* logacc, int D,
                               int batch, cudaStream_t st);
cudaError_t logdiag_accumulate_launch(const cplx* A, double* logacc, int D, long long strideA, int batch,
                                      cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// Delayed local updates (update_kernels.cu)
// ------------------------------------------------------------------------------------------------
struct UpdateModel {
    int L, N, msf, D, opdim, m;
    int delaySteps;
    double dtau, c, u, lambda, accRatio;
};
struct UpdateArgs {
    cplx* G; long long strideG;
    double* phi; double* coshT; double* sinhT;
    long long stridePhi, strideTab;
    const double* rvals;          // [batch] exchange parameter r
    cplx* X; cplx* Y;             // delayed-update workspaces [batch][D*KMAX]
    long long strideXY;
    const double* rng;            // [batch][rngWindow]
    long long strideRng;
    int rngWindow;
    int* cursor;                  // [batch] read position in the window
    dqmc_control_data* ctrl;      // [batch] step size + running average (device copy)
    uint32_t* accepted;           // [batch] accepted proposals in this slice
    unsigned long long* acceptedTotal;   // [batch] running total (for flop accounting)
    int* errflag;                 // device error flag (rng window overrun, NaN)
    int k;                        // time slice
    int thermalization;
    int batch;
    int round;                    // 0 .. rounds-1 within the slice
    int inline_flush;             // 1: the CTA applies G += X Y itself and finishes the slice in one launch
    int* site_state;              // [batch] next site to visit (carried from round to round)
    int* kvec;                    // [batch] K = MSF * (#accepted in this round) for the rank-K flush
    int debug;                    // print per-phase clock counts of replica 0 (development)
    int final_pass;               // last of the repeatUpdateInSlice passes: record the acceptance statistics / adapt the step
    int y_in_smem;                // set by the launcher: the CTA keeps the pending Y rows in shared memory
    // ---- window rounds (update_window_kernel + update_build_xy_kernel) ----
    dqmc::cplx* wscratch;         // [batch][strideScratch] per-round record of the accepted updates (see update_kernels.cu)
    long long strideScratch;
    int* whdr;                    // [batch][strideHdr] ints: nacc, site0, w, pad, sites[delaySteps]
    int strideHdr;
    int wmax;                     // sites per window
};
// window rounds: sizes of the per-replica scratch (cplx elements) and header (ints) for a model
int update_window_sites(const UpdateModel& m);            // sites per window, 0 = window path not applicable
size_t update_window_scratch_elems(const UpdateModel& m);
int update_window_hdr_ints(const UpdateModel& m);
cudaError_t update_window_launch(const UpdateModel& m, const UpdateArgs& a, cudaStream_t st);
cudaError_t update_build_xy_launch(const UpdateModel& m, const UpdateArgs& a, cudaStream_t st);
// one round of the delayed local updates; with inline_flush == 0 the caller applies
// G += X[:, :kvec] Y[:kvec, :] (rank-K update GEMM) after every round
int update_rounds_per_slice(const UpdateModel& m, int inline_flush);
cudaError_t update_round_launch(const UpdateModel& m, const UpdateArgs& a, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// DetHubbard (hubbard_kernels.cu)
// ------------------------------------------------------------------------------------------------
void hub_build_propagators(const dqmc_params& p, std::vector<double>& P, std::vector<double>& Pinv);
cudaError_t hub_scales_launch(const int32_t* aux, long long strideAux, int N, int k, double alpha, double sign,
                              double* out, int off, int batch, cudaStream_t st);
cudaError_t hub_real_part_launch(const cplx* in, double* out, size_t n, cudaStream_t st);
cudaError_t hub_to_complex_launch(const double* in, cplx* out, size_t n, cudaStream_t st);
cudaError_t hub_measure_launch(const cplx* G, long long strideG, int N, int L, double* acc, long long strideAcc, int batch,
                               cudaStream_t st);
cudaError_t hub_update_slice_launch(cplx* G, long long strideG, int N, int32_t* aux, long long strideAux, int k,
                                    double alpha, const double* rng, long long strideRng, int rngWindow, int* cursor,
                                    uint32_t* accepted, unsigned long long* acceptedTotal, int* errflag, int batch,
                                    cudaStream_t st);

}  // namespace dqmc

// ------------------------------------------------------------------------------------------------
// The context
// ------------------------------------------------------------------------------------------------
struct dqmc_ctx {
    dqmc_params p;
    int R;                 // replicas
    int N, D, msf, m, s, n, ngc, opdim;
    int nmat;              // R * ngc matrices per batched operator
    int device;
    cudaStream_t stream;
    bool own_stream;
    std::string err;
    uint64_t launches;

    // dense hopping path (CB_NONE): block-diagonal e^{-+dtau K}, a work matrix per batch member; denseNow switches the
    // B-matrix multiplies to it (always for denseHopping models, temporarily inside dqmc_sweep_simple)
    dqmc::cplx* denseP = nullptr; dqmc::cplx* densePinv = nullptr; dqmc::cplx* denseTmp = nullptr; bool denseNow = false;
    // NCCL communicator of the exchange step (dqmc_set_comm); nullptr = single process
    void* comm = nullptr; int commRanks = 0, commRank = 0;
    // sweep state (detmodel.h:463, 481; detsdwopdim.h performedSweeps)
    int currentTimeslice;
    int lastSweepDir;      // +1 up, -1 down
    int performedSweeps;

    dqmc::CbGeom geom;
    dqmc::UpdateModel umodel;

    // device buffers
    dqmc::cplx* G;         // [nmat][D*D]
    dqmc::cplx* Gwrapped;  // copy of G before an advance (green consistency)
    dqmc::cplx* W[5];      // scratch [nmat][D*D]
    dqmc::QrWorkspace qr;  // blocked-QR workspace (V, V*T, panel products)
    int stabilizer;        // 0: blocked QR with column pre-pivoting (default), 1: fully pivoted one-CTA QR
    dqmc::cplx* tQ; dqmc::cplx* tT; double* tD;   // temporary UDT of an advance step
    dqmc::cplx* eyeM; double* onesV;              // shared identity UDT (batch stride 0)
    dqmc::cplx* stQ;       // UDT storage [nmat][n+1][D*D]
    dqmc::cplx* stT;
    double* stD;           // [nmat][n+1][D]
    dqmc::cplx* bkQ;       // backups for the global move (swapped by pointer)
    dqmc::cplx* bkT;
    double* bkD;
    dqmc::cplx* bkG;
    double* bkPhi; double* bkCosh; double* bkSinh;
    double* phi;           // [R][(m+1)*opdim*N]
    double* coshT; double* sinhT;   // [R][(m+1)*N]
    double* rvals;         // [R]
    dqmc::cplx* cbtab;
    dqmc::cplx* tau;       // [nmat][D]
    int* perm;             // [nmat][D]
    double* colnorm;       // [nmat][D]
    double* vecA; double* vecB; double* vecC; double* vecD;   // [nmat][D] scale vectors
    double* dtmp;          // [nmat][D]
    double* logdet;        // [nmat] log|det G^-1| from the last from-scratch evaluation
    double* bkLogdet;
    double* consistency;   // [nmat]
    dqmc::cplx* X; dqmc::cplx* Y;   // [R][D*KMAX]
    dqmc::cplx* winScratch; int* winHdr;   // window rounds: per-round record [R][winScratchStride], header [R][winHdrStride]
    size_t winScratchStride; int winHdrStride; int winSites;
    double* cfgStream;               // [R][N*m*opdim] configuration-stream staging (allocated on first use)
    // fermionic measurements (allocated on first use): block-diagonal shift matrices, per-replica accumulators
    dqmc::cplx* shiftL; dqmc::cplx* shiftR; double* fmAcc; size_t fmAccLen; int fmSlices;
    int kmax;
    double* rngbuf;        // [R][rngCap]
    size_t rngCap;
    int* cursor;           // [R]
    int rngWindow;         // number of values per replica in the uploaded window
    size_t rngAlloc, rngStride;
    size_t hRngAlloc;      // doubles in the pinned staging buffer h_rng
    bool rngResident;      // window pre-loaded for several sweeps (dqmc_rng_preload) or streamed (rngAuto)
    bool rngAuto;          // streamed mode of dqmc_sweep: chunks are uploaded one sweep ahead on copyStream
    size_t rngUploaded;    // values per replica present in the device buffer (streamed mode)
    cudaStream_t copyStream; cudaEvent_t copyEvent[2]; int copyHalf;
    size_t rngResidentUsedBound;
    unsigned long long* acceptedTotal;
    // lanes: the replicas are split into up to two groups whose sweeps are issued on separate streams
    int nlanes; int laneStart[DQMC_MAX_LANES + 1]; int laneOff, laneCnt;
    cudaStream_t laneStream[DQMC_MAX_LANES]; cudaEvent_t laneEvent[DQMC_MAX_LANES];
    // captured sweeps (run_sweep)
    struct SweepGraph { int dir, therm, rngWindow; long long rngStride; int nlanes, stab; cudaStream_t stream;
                        cudaGraphExec_t exec; uint64_t launches; };
    std::vector<SweepGraph> graphs;
    bool graphsOff;
    bool profiling;
    int profForce;         // >= 0: category of the next timed launch (overrides the name-based one)
    struct ProfRec { int cat; cudaEvent_t e0, e1; };
    std::vector<ProfRec> profPending;
    std::vector<cudaEvent_t> profPool;
    double profMs[DQMC_PROF_NCAT];
    uint64_t profCount[DQMC_PROF_NCAT];
    dqmc_control_data* ctrl;      // [R] device
    uint32_t* accepted;    // [R]
    int* siteState;        // [R] delayed-update rounds: next site
    int* cursorAdd;        // [R] values read by the host from the resident window (global moves)
    int* kvec;             // [R] delayed-update rounds: pending rank
    int* errflag;
    double* actions;       // [R]
    double* shiftbuf;      // [R][3]

    // pinned host staging
    double* h_rng;         // [R][rngCap]
    int* h_cursor;
    double* h_scalars;     // generic staging [max(R*8, ...)]
    dqmc_control_data* h_ctrl;
    int* h_err;
    uint32_t* h_acc;
    std::vector<double> lastGlobalProb;

    // Hubbard
    int32_t* aux;          // [R][(m+1)*N]
    dqmc::cplx* propT;     // e^{-dtau T} as complex D x D
    dqmc::cplx* propTinv;
    double* hubScale;      // [nmat][D] scratch row/col scales
    dqmc::cplx* hubTmp;    // [nmat][D*D] ping-pong buffer of the dense B chains
    double* hubReal;       // [D*D] staging for real <-> complex conversion
    double hubAlpha;       // cosh(alpha) = exp(dtau U / 2)

    std::vector<dqmc::RngStream> rng;
    std::vector<double> h_r;
    std::vector<dqmc_control_data> ctrl_host;

    // which storage entries hold a left chain (B(beta,tau) = T^+ d Q^+) vs right chain (Q d T)
    // is implied by the sweep direction exactly as in the reference; nothing to track.
};
