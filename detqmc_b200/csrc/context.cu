// The batch context and the C ABI (include/dqmc_gpu.h): device-resident state of R replicas, the
// stabilised sweep skeleton of DetModelGC (detmodel.h:678-713, 953-1163, 1261-1440) sequenced as
// batched kernel launches on one CUDA stream, UDT storage management, the global shift move and
// the replica-exchange helpers.  There is NO CPU fallback: every numerical step is a kernel launch
// and dqmc_create fails when no CUDA device is usable.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "dqmc_internal.h"

#include <dlfcn.h>

using namespace dqmc;

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return DQMC_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

// kernel launch: counted, and (in profiling mode) bracketed by events on the context's stream
#define CKL(call)                                                                                  \
    do {                                                                                           \
        ctx->launches += 1;                                                                        \
        if (ctx->profiling) {                                                                      \
            cudaEvent_t e0__ = prof_event(ctx), e1__ = prof_event(ctx);                            \
            cudaEventRecord(e0__, ctx->stream);                                                    \
            cudaError_t el__ = (call);                                                             \
            cudaEventRecord(e1__, ctx->stream);                                                    \
            ctx->profPending.push_back({ctx->profForce >= 0 ? ctx->profForce : prof_category(#call), e0__, e1__});                        \
            if (el__ != cudaSuccess) {                                                             \
                ctx->err = std::string(#call) + ": " + cudaGetErrorString(el__);                   \
                return DQMC_ERR_CUDA;                                                              \
            }                                                                                      \
        } else {                                                                                   \
            CK(call);                                                                              \
        }                                                                                          \
    } while (0)

#define RET(call)                                                                                  \
    do {                                                                                           \
        int r__ = (call);                                                                          \
        if (r__ != DQMC_OK) return r__;                                                            \
    } while (0)

// profiling categories (dqmc_profile_get): one per kernel family
const char* const kProfNames[DQMC_PROF_NCAT] = {"cb_mult", "gemm_dmma", "qrcp_factor", "qr_form_q", "trsm_upper",
                                                "update_slice", "other", "cb_chain", "update_flush", "update_build_xy"};
int prof_category(const char* call) {
    if (!std::strncmp(call, "cb_launch", 9)) return 0;
    if (!std::strncmp(call, "gemm_launch", 11)) return 1;
    if (!std::strncmp(call, "qrcp_factor", 11) || !std::strncmp(call, "qr_blocked_factor", 17)) return 2;
    if (!std::strncmp(call, "qr_form_q", 9) || !std::strncmp(call, "qr_blocked_form_q", 17) ||
        !std::strncmp(call, "qr_blocked_apply_qh", 19)) return 3;
    if (!std::strncmp(call, "trsm_upper", 10)) return 4;
    if (!std::strncmp(call, "update_", 7) || !std::strncmp(call, "hub_update", 10)) return 5;
    return 6;
}
cudaEvent_t prof_event(dqmc_ctx* ctx) {
    if (!ctx->profPool.empty()) {
        cudaEvent_t e = ctx->profPool.back();
        ctx->profPool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_collect(dqmc_ctx* ctx) {
    for (auto& p : ctx->profPending) {
        float ms = 0;
        if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            ctx->profMs[p.cat] += ms;
            ctx->profCount[p.cat] += 1;
        }
        ctx->profPool.push_back(p.e0);
        ctx->profPool.push_back(p.e1);
    }
    ctx->profPending.clear();
}

template <class T>
cudaError_t dmalloc(T** p, size_t n) {
    return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
}

inline size_t DD(const dqmc_ctx* c) { return size_t(c->D) * c->D; }
inline size_t phi_stride(const dqmc_ctx* c) { return size_t(c->m + 1) * c->opdim * c->N; }
inline size_t tab_stride(const dqmc_ctx* c) { return size_t(c->m + 1) * c->N; }
inline cplx* stQ(dqmc_ctx* c, int l) { return c->stQ + size_t(l) * DD(c); }
inline cplx* stT(dqmc_ctx* c, int l) { return c->stT + size_t(l) * DD(c); }
inline double* stD(dqmc_ctx* c, int l) { return c->stD + size_t(l) * c->D; }
inline long long st_stride(const dqmc_ctx* c) { return (long long)(c->n + 1) * (long long)DD(c); }
inline long long std_stride(const dqmc_ctx* c) { return (long long)(c->n + 1) * c->D; }

// A "view" of one UDT for a batch: base pointers + strides (stride 0 = shared identity)
struct UdtView {
    const cplx* Q; long long sQ;
    const double* d; long long sd;
    const cplx* T; long long sT;
};

struct OpSpec { int rows, k_then_v, sign_idx, transposed, ascending; };
constexpr int kStreamSweeps = 8;      // capacity of the streamed random-number buffer, in sweeps

const OpSpec kOps[5] = {
    {0, 1, 0, 0, 1},   // LEFT       B A
    {1, 0, 0, 1, 0},   // RIGHT      A B
    {0, 0, 1, 0, 0},   // LEFT_INV   B^-1 A
    {1, 1, 1, 1, 1},   // RIGHT_INV  A B^-1
    {0, 0, 0, 0, 0},   // LEFT_ADJ   B^+ A
};

// ---- batched building blocks; `off` selects the first replica, `batch` how many ---------------

int host_sync_rng(dqmc_ctx* ctx);
int gemm(dqmc_ctx* ctx, int ta, int tb, const cplx* A, long long sA, const cplx* B, long long sB, cplx* C,
         long long sC, const double* rows, long long sRow, const double* cols, long long sCol,
         const double* ks, long long sK, double beta, int batch);
int hub_bmult(dqmc_ctx* ctx, int op, cplx* A, long long strideA, int k2, int k1, const double* colscale,
              long long strideScale, int off, int batch);
int sdw_bmult_dense(dqmc_ctx* ctx, const OpSpec& o, cplx* A, long long strideA, int k2, int k1, const double* colscale,
                    long long strideScale, int off, int batch);

// Dense hopping path (DetSDW<CB_NONE>: computeBmatSDW, detsdwopdim.cpp:1307-1497; the reference's sweepSimple uses it
// for every model, :4366-4420): B_k = e^{-dtau V_k} e^{-dtau K}.  The hopping factor is a D x D GEMM with the
// block-diagonal propagator, the potential factor the potential stage of cb_mult_kernel (no plaquette passes).
int ensure_dense(dqmc_ctx* ctx) {
    if (ctx->denseP) return DQMC_OK;
    std::vector<cplx> P, Pinv;
    cb_build_dense_propagators(ctx->p, ctx->msf, P, Pinv);
    CK(dmalloc(&ctx->denseP, P.size()));
    CK(dmalloc(&ctx->densePinv, Pinv.size()));
    CK(dmalloc(&ctx->denseTmp, DD(ctx) * size_t(ctx->nmat)));
    CK(cudaMemcpy(ctx->denseP, P.data(), P.size() * sizeof(cplx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->densePinv, Pinv.data(), Pinv.size() * sizeof(cplx), cudaMemcpyHostToDevice));
    return DQMC_OK;
}

int sdw_bmult_dense(dqmc_ctx* ctx, const OpSpec& o, cplx* A, long long strideA, int k2, int k1, const double* colscale,
                    long long strideScale, int off, int batch) {
    RET(ensure_dense(ctx));
    const size_t dd = DD(ctx);
    cplx* tmp = ctx->denseTmp + size_t(off) * dd;
    const cplx* Pm = o.sign_idx == 0 ? ctx->denseP : ctx->densePinv;
    CbLaunch a;
    a.strideA = strideA;
    a.phi = ctx->phi + size_t(off) * phi_stride(ctx);
    a.coshT = ctx->coshT + size_t(off) * tab_stride(ctx);
    a.sinhT = ctx->sinhT + size_t(off) * tab_stride(ctx);
    a.stridePhi = (long long)phi_stride(ctx);
    a.strideTab = (long long)tab_stride(ctx);
    a.cbtab = ctx->cbtab;
    a.real_tables = ctx->p.weakZflux ? 0 : 1;
    a.kstep = 1; a.kcount = 1;
    a.rows = o.rows; a.k_then_v = o.k_then_v; a.sign_idx = o.sign_idx; a.transposed = o.transposed;
    a.strideScale = strideScale;
    a.batch = batch;
    a.skip_hopping = 1;
    const int n = k2 - k1;
    for (int i = 0; i < n; ++i) {
        const int k = o.ascending ? k1 + 1 + i : k2 - i;
        const bool last = i == n - 1;
        a.kfirst = k;
        // hopping factor: tmp = Pm * X (columns) or X * Pm (rows)
        auto hop = [&](const cplx* X, long long sX, const double* cols, long long sCol) {
            if (!o.rows) return gemm(ctx, 0, 0, Pm, 0, X, sX, tmp, (long long)dd, nullptr, 0, cols, sCol, nullptr, 0, 0.0, batch);
            return gemm(ctx, 0, 0, X, sX, Pm, 0, tmp, (long long)dd, nullptr, 0, cols, sCol, nullptr, 0, 0.0, batch);
        };
        if (o.k_then_v) {
            RET(hop(A, strideA, nullptr, 0));
            a.A = tmp; a.strideA = (long long)dd;
            a.out = A; a.strideOut = strideA;
            a.colscale = last ? colscale : nullptr;
            CKL(cb_launch(ctx->geom, a, ctx->stream));
        } else {
            a.A = A; a.strideA = strideA;
            a.out = nullptr; a.strideOut = 0;
            a.colscale = nullptr;
            CKL(cb_launch(ctx->geom, a, ctx->stream));
            RET(hop(A, strideA, last ? colscale : nullptr, strideScale));
            CK(cudaMemcpy2DAsync(A, size_t(strideA) * sizeof(cplx), tmp, dd * sizeof(cplx), dd * sizeof(cplx), batch,
                                 cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    return DQMC_OK;
}

// B-matrix multiply of the model: `off` / `batch` count MATRICES (== replicas for DetSDW)
int sdw_bmult(dqmc_ctx* ctx, int op, cplx* A, long long strideA, int k2, int k1, const double* colscale,
              long long strideScale, int off, int batch) {
    if (k2 <= k1) return DQMC_OK;
    if (ctx->p.model == DQMC_MODEL_HUBBARD) return hub_bmult(ctx, op, A, strideA, k2, k1, colscale, strideScale, off, batch);
    const OpSpec& o = kOps[op];
    if (ctx->denseNow) return sdw_bmult_dense(ctx, o, A, strideA, k2, k1, colscale, strideScale, off, batch);
    CbLaunch a;
    a.A = A;
    a.strideA = strideA;
    a.phi = ctx->phi + size_t(off) * phi_stride(ctx);
    a.coshT = ctx->coshT + size_t(off) * tab_stride(ctx);
    a.sinhT = ctx->sinhT + size_t(off) * tab_stride(ctx);
    a.stridePhi = (long long)phi_stride(ctx);
    a.strideTab = (long long)tab_stride(ctx);
    a.cbtab = ctx->cbtab;
    a.real_tables = ctx->p.weakZflux ? 0 : 1;
    a.kcount = k2 - k1;
    if (o.ascending) { a.kfirst = k1 + 1; a.kstep = +1; }
    else { a.kfirst = k2; a.kstep = -1; }
    a.rows = o.rows;
    a.k_then_v = o.k_then_v;
    a.sign_idx = o.sign_idx;
    a.transposed = o.transposed;
    a.colscale = colscale;
    a.strideScale = strideScale;
    a.batch = batch;
    ctx->profForce = a.kcount > 1 ? 7 : -1;          // chains (advance) are accounted separately from the wraps
    CKL(cb_launch(ctx->geom, a, ctx->stream));
    ctx->profForce = -1;
    return DQMC_OK;
}

int hub_bmult(dqmc_ctx* ctx, int op, cplx* A, long long strideA, int k2, int k1, const double* colscale,
              long long strideScale, int off, int batch);

int gemm(dqmc_ctx* ctx, int ta, int tb, const cplx* A, long long sA, const cplx* B, long long sB, cplx* C,
         long long sC, const double* rows, long long sRow, const double* cols, long long sCol,
         const double* ks, long long sK, double beta, int batch) {
    GemmArgs g;
    g.M = g.N = g.K = ctx->D;
    g.transa = ta; g.transb = tb;
    g.A = A; g.lda = ctx->D; g.strideA = sA;
    g.B = B; g.ldb = ctx->D; g.strideB = sB;
    g.C = C; g.ldc = ctx->D; g.strideC = sC;
    g.rowscale = rows; g.strideRow = sRow;
    g.colscale = cols; g.strideCol = sCol;
    g.kscale = ks; g.strideK = sK;
    g.alpha = 1.0; g.kvec = nullptr; g.b_kmajor = 0;
    g.beta = beta;
    g.batch = batch;
    CKL(gemm_launch(g, ctx->stream));
    return DQMC_OK;
}

// M (in `work`, destroyed) -> Q, d, T' with M = Q diag(d) T'
// DetHubbard: B_sigma(k2, k1) applied slice by slice; every slice is one propagator GEMM with the diagonal
// e^{+-sigma alpha s_k} fused as a row / k / column scaling (dethubbard.cpp:823-851, dethubbard.h:281-337)
int hub_bmult(dqmc_ctx* ctx, int op, cplx* A, long long strideA, int k2, int k1, const double* colscale,
              long long strideScale, int off, int batch) {
    const size_t dd = DD(ctx);
    const int N = ctx->N;
    cplx* cur = A;
    long long sCur = strideA;
    cplx* other = ctx->hubTmp + size_t(off) * dd;
    long long sOther = (long long)dd;
    double* sc = ctx->hubScale + size_t(off) * N;
    const bool ascending = (op == DQMC_OP_LEFT || op == DQMC_OP_RIGHT_INV);
    const bool inverse = (op == DQMC_OP_LEFT_INV || op == DQMC_OP_RIGHT_INV);
    const int count = k2 - k1;
    for (int i = 0; i < count; ++i) {
        const int k = ascending ? k1 + 1 + i : k2 - i;
        const bool last = i == count - 1;
        CKL(hub_scales_launch(ctx->aux, (long long)tab_stride(ctx), N, k, ctx->hubAlpha, inverse ? -1.0 : 1.0, sc, off,
                              batch, ctx->stream));
        const double* cs = last ? colscale : nullptr;
        const long long scs = last ? strideScale : 0;
        switch (op) {
            case DQMC_OP_LEFT:        // diag_k (P in)
                RET(gemm(ctx, 0, 0, ctx->propT, 0, cur, sCur, other, sOther, sc, N, cs, scs, nullptr, 0, 0.0, batch));
                break;
            case DQMC_OP_RIGHT:       // (in diag_k) P
                RET(gemm(ctx, 0, 0, cur, sCur, ctx->propT, 0, other, sOther, nullptr, 0, nullptr, 0, sc, N, 0.0, batch));
                break;
            case DQMC_OP_LEFT_INV:    // P^-1 (diag_k^-1 in)
                RET(gemm(ctx, 0, 0, ctx->propTinv, 0, cur, sCur, other, sOther, nullptr, 0, nullptr, 0, sc, N, 0.0, batch));
                break;
            case DQMC_OP_RIGHT_INV:   // (in P^-1) diag_k^-1
                RET(gemm(ctx, 0, 0, cur, sCur, ctx->propTinv, 0, other, sOther, nullptr, 0, sc, N, nullptr, 0, 0.0, batch));
                break;
            default:                  // LEFT_ADJ: P^T (diag_k in)
                RET(gemm(ctx, 1, 0, ctx->propT, 0, cur, sCur, other, sOther, nullptr, 0, cs, scs, sc, N, 0.0, batch));
                break;
        }
        std::swap(cur, other);
        std::swap(sCur, sOther);
    }
    if (cur != A)
        CK(cudaMemcpy2DAsync(A, size_t(strideA) * sizeof(cplx), cur, size_t(sCur) * sizeof(cplx), dd * sizeof(cplx), batch,
                             cudaMemcpyDeviceToDevice, ctx->stream));
    return DQMC_OK;
}

// the rank-K flush of the delayed updates is accounted to the update family, not to the GEMMs
inline cudaError_t update_flush_gemm(const GemmArgs& g, cudaStream_t st) { return gemm_launch(g, st); }

int udt_decompose(dqmc_ctx* ctx, cplx* work, long long sW, cplx* Qout, long long sQ, double* dout, long long sd,
                  cplx* Tout, long long sT, int off, int batch) {
    const int D = ctx->D;
    cplx* tau = ctx->tau + size_t(off) * D;
    int* perm = ctx->perm + size_t(off) * D;
    double* cn = ctx->colnorm + size_t(off) * D;
    if (sW != (long long)DD(ctx) || sQ != sW) {
        // the QR kernels use one stride for A and Q
        ctx->err = "udt_decompose: work and Q must be contiguous batches";
        return DQMC_ERR_STATE;
    }
    cplx* Ttmp = ctx->W[3] + size_t(off) * DD(ctx);
    double* dtmp = ctx->dtmp + size_t(off) * D;
    if (ctx->stabilizer == 0) {
        // column pre-pivoting into W[4], blocked Householder QR there, explicit Q from the block reflectors
        cplx* ap = ctx->W[4] + size_t(off) * DD(ctx);
        CKL(qr_prepivot_launch(work, sW, ap, sW, perm, cn, D, batch, ctx->stream));
        CKL(qr_blocked_factor(ctx->qr, ap, D, sW, off, batch, ctx->stream));
        CKL(qr_blocked_form_q(ctx->qr, Qout, D, sQ, off, batch, ctx->stream));
        CKL(qr_extract_dt_launch(ap, perm, dtmp, Ttmp, D, sW, batch, ctx->stream));
    } else {
        CKL(qrcp_factor_launch(work, D, sW, tau, perm, cn, batch, ctx->stream));
        CKL(qr_form_q_launch(work, tau, Qout, D, sW, batch, ctx->stream));
        // d and T' go to contiguous scratch first (extract kernel uses the A stride), then are copied
        // with the requested strides
        CKL(qr_extract_dt_launch(work, perm, dtmp, Ttmp, D, sW, batch, ctx->stream));
    }
    CK(cudaMemcpy2DAsync(dout, size_t(sd) * sizeof(double), dtmp, size_t(D) * sizeof(double),
                         size_t(D) * sizeof(double), batch, cudaMemcpyDeviceToDevice, ctx->stream));
    if (Tout) {
        CK(cudaMemcpy2DAsync(Tout, size_t(sT) * sizeof(cplx), Ttmp, DD(ctx) * sizeof(cplx), DD(ctx) * sizeof(cplx),
                             batch, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return DQMC_OK;
}

// one stabilised chain step: (Q, d, T) -> (Q', d', T' T) with M = (op(B(k2,k1)) Q) diag(d).
// op = LEFT for right chains B(tau,0) = Q d T, LEFT_ADJ for left chains B(beta,tau) = T^+ d Q^+.
// in == nullptr means the identity UDT.  Output strides are element strides between replicas.
int chain_step(dqmc_ctx* ctx, int op, const UdtView* in, int k2, int k1, cplx* Qout, long long sQo, double* dout,
               long long sdo, cplx* Tout, long long sTo, int off, int batch) {
    const size_t dd = DD(ctx);
    cplx* work = ctx->W[0] + size_t(off) * dd;
    cplx* qtmp = ctx->W[1] + size_t(off) * dd;
    if (in) {
        CK(cudaMemcpy2DAsync(work, dd * sizeof(cplx), in->Q, size_t(in->sQ) * sizeof(cplx), dd * sizeof(cplx), batch,
                             cudaMemcpyDeviceToDevice, ctx->stream));
        RET(sdw_bmult(ctx, op, work, (long long)dd, k2, k1, in->d, in->sd, off, batch));
    } else {
        CKL(launch_set_identity(work, ctx->D, (long long)dd, batch, ctx->stream));
        RET(sdw_bmult(ctx, op, work, (long long)dd, k2, k1, nullptr, 0, off, batch));
    }
    cplx* tprime = ctx->W[2] + size_t(off) * dd;
    RET(udt_decompose(ctx, work, (long long)dd, qtmp, (long long)dd, dout, sdo, in ? tprime : Tout,
                      in ? (long long)dd : sTo, off, batch));
    CK(cudaMemcpy2DAsync(Qout, size_t(sQo) * sizeof(cplx), qtmp, dd * sizeof(cplx), dd * sizeof(cplx), batch,
                         cudaMemcpyDeviceToDevice, ctx->stream));
    if (in) {
        // T_new = T' T_old; computed into scratch first because Tout may alias in->T
        cplx* tnew = ctx->W[3] + size_t(off) * dd;
        RET(gemm(ctx, 0, 0, tprime, (long long)dd, in->T, in->sT, tnew, (long long)dd, nullptr, 0, nullptr, 0, nullptr, 0,
                 0.0, batch));
        CK(cudaMemcpy2DAsync(Tout, size_t(sTo) * sizeof(cplx), tnew, dd * sizeof(cplx), dd * sizeof(cplx), batch,
                             cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return DQMC_OK;
}

// G = [1 + (Q_r d_r T_r)(T_l^+ d_l Q_l^+)]^-1, logdet = log|det G^-1|   (greenFromUdV's role)
int green_from_udts(dqmc_ctx* ctx, const UdtView& r, const UdtView& l, cplx* Gout, long long sG, double* logdet,
                    int off, int batch) {
    const size_t dd = DD(ctx);
    const int D = ctx->D;
    cplx* H = ctx->W[0] + size_t(off) * dd;
    cplx* Qh = ctx->W[1] + size_t(off) * dd;
    cplx* Yw = ctx->W[2] + size_t(off) * dd;
    cplx* Zw = ctx->W[3] + size_t(off) * dd;
    double* rInvBig = ctx->vecA + size_t(off) * D;
    double* rSmall = ctx->vecB + size_t(off) * D;
    double* lInvBig = ctx->vecC + size_t(off) * D;
    double* lSmall = ctx->vecD + size_t(off) * D;
    cplx* tau = ctx->tau + size_t(off) * D;
    int* perm = ctx->perm + size_t(off) * D;
    double* cn = ctx->colnorm + size_t(off) * D;
    CK(cudaMemsetAsync(logdet, 0, sizeof(double) * batch, ctx->stream));
    // scale splitting d = d_big * d_small
    if (r.sd == 0) {
        // shared (identity) scales: replicate once per replica so the kernels can use a stride
        for (int b = 0; b < batch; ++b)
            CK(cudaMemcpyAsync(ctx->dtmp + size_t(off + b) * D, r.d, sizeof(double) * D, cudaMemcpyDeviceToDevice,
                               ctx->stream));
        CKL(scale_split_launch(ctx->dtmp + size_t(off) * D, rInvBig, rSmall, logdet, D, batch, ctx->stream));
    } else {
        // gather with stride into contiguous scratch
        CK(cudaMemcpy2DAsync(ctx->dtmp + size_t(off) * D, sizeof(double) * D, r.d, sizeof(double) * size_t(r.sd),
                             sizeof(double) * D, batch, cudaMemcpyDeviceToDevice, ctx->stream));
        CKL(scale_split_launch(ctx->dtmp + size_t(off) * D, rInvBig, rSmall, logdet, D, batch, ctx->stream));
    }
    if (l.sd == 0) {
        for (int b = 0; b < batch; ++b)
            CK(cudaMemcpyAsync(ctx->dtmp + size_t(off + b) * D, l.d, sizeof(double) * D, cudaMemcpyDeviceToDevice,
                               ctx->stream));
    } else {
        CK(cudaMemcpy2DAsync(ctx->dtmp + size_t(off) * D, sizeof(double) * D, l.d, sizeof(double) * size_t(l.sd),
                             sizeof(double) * D, batch, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CKL(scale_split_launch(ctx->dtmp + size_t(off) * D, lInvBig, lSmall, logdet, D, batch, ctx->stream));
    // H = (1/d_r^b) (Q_r^+ Q_l) (1/d_l^b) + d_r^s (T_r T_l^+) d_l^s
    RET(gemm(ctx, 1, 0, r.Q, r.sQ, l.Q, l.sQ, H, (long long)dd, rInvBig, D, lInvBig, D, nullptr, 0, 0.0, batch));
    RET(gemm(ctx, 0, 1, r.T, r.sT, l.T, l.sT, H, (long long)dd, rSmall, D, lSmall, D, nullptr, 0, 1.0, batch));
    if (ctx->stabilizer == 0) {
        // H P = Q_h R_h with P fixed up front (columns by decreasing norm), blocked Householder QR
        cplx* Hp = ctx->W[4] + size_t(off) * dd;
        CKL(qr_prepivot_launch(H, (long long)dd, Hp, (long long)dd, perm, cn, D, batch, ctx->stream));
        CKL(qr_blocked_factor(ctx->qr, Hp, D, (long long)dd, off, batch, ctx->stream));
        CKL(logdiag_accumulate_launch(Hp, logdet, D, (long long)dd, batch, ctx->stream));
        // Y = Q_h^+ ((1/d_r^b) Q_r^+): the block reflectors are applied directly, Q_h is never formed
        CKL(launch_scaled_conj_transpose(r.Q, r.sQ, Yw, (long long)dd, rInvBig, D, D, batch, ctx->stream));
        CKL(qr_blocked_apply_qh(ctx->qr, Yw, D, D, (long long)dd, off, batch, ctx->stream));
        // Z = P R_h^-1 Y
        CKL(trsm_upper_blocked(ctx->qr, Hp, Yw, Qh, D, (long long)dd, off, batch, ctx->stream));
        CKL(permute_rows_launch(Qh, (long long)dd, Zw, (long long)dd, perm, D, batch, ctx->stream));
    } else {
        // H P = Q_h R_h
        CKL(qrcp_factor_launch(H, D, (long long)dd, tau, perm, cn, batch, ctx->stream));
        CKL(logdiag_accumulate_launch(H, logdet, D, (long long)dd, batch, ctx->stream));
        CKL(qr_form_q_launch(H, tau, Qh, D, (long long)dd, batch, ctx->stream));
        // Y = Q_h^+ (1/d_r^b) Q_r^+
        RET(gemm(ctx, 1, 1, Qh, (long long)dd, r.Q, r.sQ, Yw, (long long)dd, nullptr, 0, nullptr, 0, rInvBig, D, 0.0,
                 batch));
        // Z = P R_h^-1 Y
        CKL(trsm_upper_launch(H, Yw, Zw, perm, D, (long long)dd, batch, ctx->stream));
    }
    // G = Q_l (1/d_l^b) Z
    RET(gemm(ctx, 0, 0, l.Q, l.sQ, Zw, (long long)dd, Gout, sG, nullptr, 0, nullptr, 0, lInvBig, D, 0.0, batch));
    return DQMC_OK;
}

UdtView storage_view(dqmc_ctx* ctx, int l, int off) {
    UdtView v;
    v.Q = stQ(ctx, l) + size_t(off) * st_stride(ctx); v.sQ = st_stride(ctx);
    v.T = stT(ctx, l) + size_t(off) * st_stride(ctx); v.sT = st_stride(ctx);
    v.d = stD(ctx, l) + size_t(off) * std_stride(ctx); v.sd = std_stride(ctx);
    return v;
}

UdtView identity_view(dqmc_ctx* ctx) {
    UdtView v;
    v.Q = ctx->eyeM; v.sQ = 0;
    v.T = ctx->eyeM; v.sT = 0;
    v.d = ctx->onesV; v.sd = 0;
    return v;
}

int set_storage_identity(dqmc_ctx* ctx, int l, int off, int batch) {
    CKL(launch_set_identity(stQ(ctx, l) + size_t(off) * st_stride(ctx), ctx->D, st_stride(ctx), batch, ctx->stream));
    CKL(launch_set_identity(stT(ctx, l) + size_t(off) * st_stride(ctx), ctx->D, st_stride(ctx), batch, ctx->stream));
    for (int b = 0; b < batch; ++b)
        CK(cudaMemcpyAsync(stD(ctx, l) + size_t(off + b) * std_stride(ctx), ctx->onesV, sizeof(double) * ctx->D,
                           cudaMemcpyDeviceToDevice, ctx->stream));
    return DQMC_OK;
}

inline int slice_of(const dqmc_ctx* ctx, int l) { return l < ctx->n ? ctx->s * l : ctx->m; }

// current lane (sub-batch) of the sweep primitives: replicas [laneOff, laneOff + laneCnt) on ctx->stream
inline int lane_mo(const dqmc_ctx* c) { return c->laneOff * c->ngc; }      // first matrix
inline int lane_mc(const dqmc_ctx* c) { return c->laneCnt * c->ngc; }      // matrix count

int setup_storage(dqmc_ctx* ctx, int off, int batch) {
    const int n = ctx->n;
    RET(set_storage_identity(ctx, 0, off, batch));
    for (int l = 0; l < n; ++l) {
        const int k_l = slice_of(ctx, l), k_lp1 = slice_of(ctx, l + 1);
        UdtView in = storage_view(ctx, l, off);
        RET(chain_step(ctx, DQMC_OP_LEFT, l == 0 ? nullptr : &in, k_lp1, k_l,
                       stQ(ctx, l + 1) + size_t(off) * st_stride(ctx), st_stride(ctx),
                       stD(ctx, l + 1) + size_t(off) * std_stride(ctx), std_stride(ctx),
                       stT(ctx, l + 1) + size_t(off) * st_stride(ctx), st_stride(ctx), off, batch));
    }
    UdtView r = storage_view(ctx, n, off);
    UdtView l = identity_view(ctx);
    RET(green_from_udts(ctx, r, l, ctx->G + size_t(off) * DD(ctx), (long long)DD(ctx), ctx->logdet + off, off, batch));
    return DQMC_OK;
}

int record_wrapped(dqmc_ctx* ctx) {
    const size_t o = size_t(lane_mo(ctx)) * DD(ctx);
    CK(cudaMemcpyAsync(ctx->Gwrapped + o, ctx->G + o, sizeof(cplx) * DD(ctx) * lane_mc(ctx), cudaMemcpyDeviceToDevice,
                       ctx->stream));
    return DQMC_OK;
}
int record_consistency(dqmc_ctx* ctx) {
    const size_t o = size_t(lane_mo(ctx)) * DD(ctx);
    CKL(launch_max_abs_diff(ctx->Gwrapped + o, ctx->G + o, ctx->D, (long long)DD(ctx), lane_mc(ctx),
                            ctx->consistency + lane_mo(ctx), ctx->stream));
    return DQMC_OK;
}

// advanceUpGreen(l), detmodel.h:1106-1163 (for the replicas of the current lane)
int advance_up(dqmc_ctx* ctx, int l) {
    const int mo = lane_mo(ctx), R = lane_mc(ctx);
    const size_t dd = DD(ctx);
    const int k_l = ctx->s * l, k_lp1 = slice_of(ctx, l + 1);
    if (ctx->currentTimeslice != k_lp1) { ctx->err = "advance_up: currentTimeslice mismatch"; return DQMC_ERR_STATE; }
    RET(record_wrapped(ctx));
    cplx* tQ = ctx->tQ + size_t(mo) * dd;
    cplx* tT = ctx->tT + size_t(mo) * dd;
    double* tD = ctx->tD + size_t(mo) * ctx->D;
    cplx* G = ctx->G + size_t(mo) * dd;
    UdtView in = storage_view(ctx, l, mo);
    // new right chain B(k_lp1, 0) into the temporary UDT
    // storage[0] is the identity UdV during an up-sweep (detmodel.h:1292-1295)
    RET(chain_step(ctx, DQMC_OP_LEFT, l == 0 ? nullptr : &in, k_lp1, k_l, tQ, (long long)dd, tD, ctx->D, tT, (long long)dd, mo, R));
    UdtView rnew{tQ, (long long)dd, tD, ctx->D, tT, (long long)dd};
    if (k_lp1 != ctx->m) {
        UdtView left = storage_view(ctx, l + 1, mo);         // B(beta, k_lp1) from the last down-sweep
        RET(green_from_udts(ctx, rnew, left, G, (long long)dd, ctx->logdet + mo, mo, R));
    } else {
        UdtView left = identity_view(ctx);
        RET(green_from_udts(ctx, rnew, left, G, (long long)dd, ctx->logdet + mo, mo, R));
    }
    CK(cudaMemcpy2DAsync(stQ(ctx, l + 1) + size_t(mo) * st_stride(ctx), size_t(st_stride(ctx)) * sizeof(cplx), tQ,
                         dd * sizeof(cplx), dd * sizeof(cplx), R, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpy2DAsync(stT(ctx, l + 1) + size_t(mo) * st_stride(ctx), size_t(st_stride(ctx)) * sizeof(cplx), tT,
                         dd * sizeof(cplx), dd * sizeof(cplx), R, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpy2DAsync(stD(ctx, l + 1) + size_t(mo) * std_stride(ctx), size_t(std_stride(ctx)) * sizeof(double), tD,
                         ctx->D * sizeof(double), ctx->D * sizeof(double), R, cudaMemcpyDeviceToDevice, ctx->stream));
    RET(record_consistency(ctx));
    ctx->currentTimeslice = k_lp1;
    return DQMC_OK;
}

// advanceDownGreen(l), detmodel.h:953-1017 (for the replicas of the current lane)
int advance_down(dqmc_ctx* ctx, int l) {
    const int mo = lane_mo(ctx), R = lane_mc(ctx);
    const size_t dd = DD(ctx);
    const int n = ctx->n;
    const int k_l = slice_of(ctx, l), k_lm1 = ctx->s * (l - 1);
    RET(record_wrapped(ctx));
    cplx* tQ = ctx->tQ + size_t(mo) * dd;
    cplx* tT = ctx->tT + size_t(mo) * dd;
    double* tD = ctx->tD + size_t(mo) * ctx->D;
    cplx* G = ctx->G + size_t(mo) * dd;
    UdtView in = storage_view(ctx, l, mo);
    RET(chain_step(ctx, DQMC_OP_LEFT_ADJ, l < n ? &in : nullptr, k_l, k_lm1, tQ, (long long)dd, tD, ctx->D, tT, (long long)dd, mo, R));
    UdtView lnew{tQ, (long long)dd, tD, ctx->D, tT, (long long)dd};
    if (l - 1 > 0) {
        UdtView right = storage_view(ctx, l - 1, mo);        // B(k_lm1, 0) from the last up-sweep
        RET(green_from_udts(ctx, right, lnew, G, (long long)dd, ctx->logdet + mo, mo, R));
    } else {
        UdtView right = identity_view(ctx);
        RET(green_from_udts(ctx, right, lnew, G, (long long)dd, ctx->logdet + mo, mo, R));
    }
    CK(cudaMemcpy2DAsync(stQ(ctx, l - 1) + size_t(mo) * st_stride(ctx), size_t(st_stride(ctx)) * sizeof(cplx), tQ,
                         dd * sizeof(cplx), dd * sizeof(cplx), R, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpy2DAsync(stT(ctx, l - 1) + size_t(mo) * st_stride(ctx), size_t(st_stride(ctx)) * sizeof(cplx), tT,
                         dd * sizeof(cplx), dd * sizeof(cplx), R, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpy2DAsync(stD(ctx, l - 1) + size_t(mo) * std_stride(ctx), size_t(std_stride(ctx)) * sizeof(double), tD,
                         ctx->D * sizeof(double), ctx->D * sizeof(double), R, cudaMemcpyDeviceToDevice, ctx->stream));
    RET(record_consistency(ctx));
    ctx->currentTimeslice = k_lm1;
    return DQMC_OK;
}

int wrap_up(dqmc_ctx* ctx, int k) {
    if (ctx->currentTimeslice != k) { ctx->err = "wrap_up: currentTimeslice mismatch"; return DQMC_ERR_STATE; }
    cplx* G = ctx->G + size_t(lane_mo(ctx)) * DD(ctx);
    RET(sdw_bmult(ctx, DQMC_OP_RIGHT_INV, G, (long long)DD(ctx), k + 1, k, nullptr, 0, lane_mo(ctx), lane_mc(ctx)));
    RET(sdw_bmult(ctx, DQMC_OP_LEFT, G, (long long)DD(ctx), k + 1, k, nullptr, 0, lane_mo(ctx), lane_mc(ctx)));
    ctx->currentTimeslice = k + 1;
    return DQMC_OK;
}

int wrap_down(dqmc_ctx* ctx, int k) {
    if (ctx->currentTimeslice != k) { ctx->err = "wrap_down: currentTimeslice mismatch"; return DQMC_ERR_STATE; }
    cplx* G = ctx->G + size_t(lane_mo(ctx)) * DD(ctx);
    RET(sdw_bmult(ctx, DQMC_OP_RIGHT, G, (long long)DD(ctx), k, k - 1, nullptr, 0, lane_mo(ctx), lane_mc(ctx)));
    RET(sdw_bmult(ctx, DQMC_OP_LEFT_INV, G, (long long)DD(ctx), k, k - 1, nullptr, 0, lane_mo(ctx), lane_mc(ctx)));
    ctx->currentTimeslice = k - 1;
    return DQMC_OK;
}

// ---- random-number window --------------------------------------------------------------------
int upload_rng_window(dqmc_ctx* ctx, size_t per_replica) {
    if (ctx->rngResident) { ctx->err = "release the resident random numbers first (dqmc_rng_release)"; return DQMC_ERR_STATE; }
    if (per_replica > ctx->rngCap) { ctx->err = "rng window larger than capacity"; return DQMC_ERR_STATE; }
    ctx->rngStride = ctx->rngCap;
    for (int r = 0; r < ctx->R; ++r) {
        const double* src = ctx->rng[r].peek(per_replica);
        std::memcpy(ctx->h_rng + size_t(r) * ctx->rngCap, src, per_replica * sizeof(double));
    }
    CK(cudaMemcpy2DAsync(ctx->rngbuf, ctx->rngCap * sizeof(double), ctx->h_rng, ctx->rngCap * sizeof(double),
                         per_replica * sizeof(double), ctx->R, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->cursor, 0, sizeof(int) * ctx->R, ctx->stream));
    ctx->rngWindow = (int)per_replica;
    return DQMC_OK;
}

// read back cursors (+ control data), advance the host streams
int finish_rng_window(dqmc_ctx* ctx) {
    CK(cudaMemcpyAsync(ctx->h_cursor, ctx->cursor, sizeof(int) * ctx->R, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl, sizeof(dqmc_control_data) * ctx->R, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_err, ctx->errflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_err) {
        ctx->err = "device error flag set (random-number window exhausted)";
        return DQMC_ERR_STATE;
    }
    for (int r = 0; r < ctx->R; ++r) {
        ctx->rng[r].skip((size_t)ctx->h_cursor[r]);
        ctx->ctrl_host[r] = ctx->h_ctrl[r];
    }
    ctx->rngWindow = 0;
    return DQMC_OK;
}

// ---- streamed random numbers (the default of dqmc_sweep) -------------------------------------------
// The device keeps a linear buffer of kStreamSweeps sweeps' worth of every replica's stream; the cursors
// stay on the device from sweep to sweep.  While a sweep runs, the host generates the next chunk, stages it
// in pinned memory and copies it on a separate stream, so that the per-sweep host->device transfer and the
// dSFMT generation overlap with the kernels instead of preceding them.  The host streams are reconciled
// (cursors read back, FIFOs advanced) when the buffer is exhausted or the host itself needs to draw.
int stream_chunk_upload(dqmc_ctx* ctx, size_t first, size_t count, cudaStream_t st, int half) {
    double* stage = ctx->h_rng + size_t(half) * ctx->rngCap * ctx->R;
    for (int r = 0; r < ctx->R; ++r) {
        const double* src = ctx->rng[r].peek(first + count) + first;
        std::memcpy(stage + size_t(r) * count, src, count * sizeof(double));
    }
    CK(cudaMemcpy2DAsync(ctx->rngbuf + first, ctx->rngStride * sizeof(double), stage, count * sizeof(double),
                         count * sizeof(double), ctx->R, cudaMemcpyHostToDevice, st));
    return DQMC_OK;
}

int host_sync_rng(dqmc_ctx* ctx) {
    if (!(ctx->rngAuto && ctx->rngResident)) return DQMC_OK;
    CK(cudaStreamSynchronize(ctx->copyStream));
    RET(finish_rng_window(ctx));
    ctx->rngResident = false;
    ctx->rngAuto = false;
    ctx->rngStride = ctx->rngCap;
    return DQMC_OK;
}

int stream_begin_sweep(dqmc_ctx* ctx) {
    if (ctx->rngResident && ctx->rngResidentUsedBound + ctx->rngCap > ctx->rngStride) RET(host_sync_rng(ctx));   // buffer exhausted
    if (!ctx->rngResident) {
        ctx->rngStride = ctx->rngAlloc < ctx->rngCap * kStreamSweeps ? ctx->rngCap : ctx->rngCap * kStreamSweeps;
        ctx->rngResident = true;
        ctx->rngAuto = true;
        ctx->rngWindow = (int)ctx->rngStride;              // constant bound: the graphs are keyed on it
        ctx->rngResidentUsedBound = 0;
        ctx->rngUploaded = 0;
        CK(cudaMemsetAsync(ctx->cursor, 0, sizeof(int) * ctx->R, ctx->stream));
    }
    const size_t need = ctx->rngResidentUsedBound + ctx->rngCap;
    if (ctx->rngUploaded < need) {
        // not (completely) prefetched -- the first sweep after a restart, or an exchange step consumed look-ahead
        // uniforms so that the bound moved past the prefetched chunk: upload the remainder on the main stream.  The
        // chunk that is still in flight on the copy stream must land before the sweep reads it.
        CK(cudaStreamWaitEvent(ctx->stream, ctx->copyEvent[ctx->copyHalf ^ 1], 0));
        CK(cudaEventSynchronize(ctx->copyEvent[ctx->copyHalf]));
        RET(stream_chunk_upload(ctx, ctx->rngUploaded, need - ctx->rngUploaded, ctx->stream, ctx->copyHalf));
        CK(cudaEventRecord(ctx->copyEvent[ctx->copyHalf], ctx->stream));
        ctx->copyHalf ^= 1;
        ctx->rngUploaded = need;
    } else {
        CK(cudaStreamWaitEvent(ctx->stream, ctx->copyEvent[ctx->copyHalf ^ 1], 0));    // the prefetched chunk
    }
    return DQMC_OK;
}

int stream_end_sweep(dqmc_ctx* ctx) {
    ctx->rngResidentUsedBound += ctx->rngCap;
    const size_t need = ctx->rngResidentUsedBound + ctx->rngCap;
    if (need <= ctx->rngStride && ctx->rngUploaded < need) {
        // the device is busy with the sweep just issued: generate, stage and copy the next chunk now
        CK(cudaEventSynchronize(ctx->copyEvent[ctx->copyHalf]));                        // staging half free again
        RET(stream_chunk_upload(ctx, ctx->rngUploaded, need - ctx->rngUploaded, ctx->copyStream, ctx->copyHalf));
        CK(cudaEventRecord(ctx->copyEvent[ctx->copyHalf], ctx->copyStream));
        ctx->copyHalf ^= 1;
        ctx->rngUploaded = need;
    }
    return DQMC_OK;
}

int launch_update(dqmc_ctx* ctx, int k, int therm) {
    const int ro = ctx->laneOff, rc = ctx->laneCnt;           // replicas of the current lane
    const size_t dd = DD(ctx);
    if (ctx->p.model == DQMC_MODEL_HUBBARD) {
        CKL(hub_update_slice_launch(ctx->G + size_t(2 * ro) * dd, (long long)dd, ctx->N, ctx->aux + size_t(ro) * tab_stride(ctx),
                                    (long long)tab_stride(ctx), k, ctx->hubAlpha, ctx->rngbuf + size_t(ro) * ctx->rngStride,
                                    (long long)ctx->rngStride, ctx->rngWindow, ctx->cursor + ro, ctx->accepted + ro,
                                    ctx->acceptedTotal + ro, ctx->errflag, rc, ctx->stream));
        return DQMC_OK;
    }
    UpdateArgs a;
    a.G = ctx->G + size_t(ro) * dd; a.strideG = (long long)dd;
    a.phi = ctx->phi + size_t(ro) * phi_stride(ctx);
    a.coshT = ctx->coshT + size_t(ro) * tab_stride(ctx);
    a.sinhT = ctx->sinhT + size_t(ro) * tab_stride(ctx);
    a.stridePhi = (long long)phi_stride(ctx); a.strideTab = (long long)tab_stride(ctx);
    a.rvals = ctx->rvals + ro;
    a.strideXY = (long long)ctx->D * ctx->kmax;
    a.X = ctx->X + size_t(ro) * a.strideXY; a.Y = ctx->Y + size_t(ro) * a.strideXY;
    a.rng = ctx->rngbuf + size_t(ro) * ctx->rngStride; a.strideRng = (long long)ctx->rngStride; a.rngWindow = ctx->rngWindow;
    a.acceptedTotal = ctx->acceptedTotal + ro;
    a.cursor = ctx->cursor + ro;
    a.ctrl = ctx->ctrl + ro;
    a.accepted = ctx->accepted + ro;
    a.errflag = ctx->errflag;
    a.k = k;
    a.thermalization = therm;
    a.batch = rc;
    a.site_state = ctx->siteState + ro;
    a.debug = std::getenv("DQMC_UPD_DEBUG") ? std::atoi(std::getenv("DQMC_UPD_DEBUG")) : 0;
    a.kvec = ctx->kvec + ro;
    // small delay blocks (Woodbury = 1) are flushed inside the kernel; otherwise every round is
    // followed by the rank-K update G += X Y on all SMs
    a.inline_flush = ctx->p.delaySteps < 8 ? 1 : 0;
    a.y_in_smem = 0;
    const bool window = ctx->winSites > 0 && !a.inline_flush;
    a.wmax = ctx->winSites;
    a.strideScratch = (long long)ctx->winScratchStride; a.strideHdr = ctx->winHdrStride;
    a.wscratch = window ? ctx->winScratch + size_t(ro) * ctx->winScratchStride : nullptr;
    a.whdr = window ? ctx->winHdr + size_t(ro) * ctx->winHdrStride : nullptr;
    const int rounds = update_rounds_per_slice(ctx->umodel, a.inline_flush);
    const int passes = std::max(1, ctx->p.repeatUpdateInSlice);       // updateInSlice repeats the pass, detsdwopdim.cpp:2438
    for (int rd = 0; rd < rounds * passes; ++rd) {
        a.round = rd % rounds;
        a.final_pass = rd / rounds == passes - 1 ? 1 : 0;
        if (window) {
            // decisions of up to delaySteps acceptances on the window block, then X, Y for all rows / columns
            CKL(update_window_launch(ctx->umodel, a, ctx->stream));
            ctx->profForce = 9;
            CKL(update_build_xy_launch(ctx->umodel, a, ctx->stream));
            ctx->profForce = -1;
        } else {
            CKL(update_round_launch(ctx->umodel, a, ctx->stream));
        }
        if (!a.inline_flush) {
            GemmArgs g;
            g.M = g.N = ctx->D; g.K = ctx->kmax;
            g.transa = g.transb = 0;
            g.A = a.X; g.lda = ctx->D; g.strideA = a.strideXY;
            g.B = a.Y; g.ldb = ctx->D; g.strideB = a.strideXY; g.b_kmajor = 1;
            g.C = a.G; g.ldc = ctx->D; g.strideC = (long long)dd;
            g.rowscale = g.colscale = g.kscale = nullptr;
            g.strideRow = g.strideCol = g.strideK = 0;
            g.alpha = 1.0; g.beta = 1.0; g.kvec = a.kvec;
            g.batch = rc;
            ctx->profForce = 8;
            CKL(update_flush_gemm(g, ctx->stream));
            ctx->profForce = -1;
        }
    }
    return DQMC_OK;
}

int upload_ctrl(dqmc_ctx* ctx) {
    std::memcpy(ctx->h_ctrl, ctx->ctrl_host.data(), sizeof(dqmc_control_data) * ctx->R);
    CK(cudaMemcpyAsync(ctx->ctrl, ctx->h_ctrl, sizeof(dqmc_control_data) * ctx->R, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int upload_rvals(dqmc_ctx* ctx) {
    CK(cudaMemcpyAsync(ctx->rvals, ctx->h_r.data(), sizeof(double) * ctx->R, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

// resident mode: bring the host streams up to date with the device cursors (needed before the host
// itself draws, e.g. for a global move), and afterwards re-upload the window from the new head
int sync_resident_cursor(dqmc_ctx* ctx) {
    RET(finish_rng_window(ctx));
    return DQMC_OK;
}
int reupload_resident(dqmc_ctx* ctx) {
    const size_t per = ctx->rngStride;
    for (int r = 0; r < ctx->R; ++r)
        std::memcpy(ctx->h_rng + size_t(r) * per, ctx->rng[r].peek(per), per * sizeof(double));
    CK(cudaMemcpyAsync(ctx->rngbuf, ctx->h_rng, per * ctx->R * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->cursor, 0, sizeof(int) * ctx->R, ctx->stream));
    ctx->rngWindow = (int)per;
    ctx->rngResidentUsedBound = 0;
    return DQMC_OK;
}

// attemptGlobalShiftMove for the whole batch, detsdwopdim.cpp:3564-3645
// buildAndFlipCluster (detsdwopdim.cpp:3805-3883) on the host copy `phi` ([m+1][opdim][N]) of one replica: reflect
// phi -> phi - 2 (phi . rd) rd on a cluster grown from a random seed over space (XPLUS, XMINUS, YPLUS, YMINUS) and
// time (PLUS, MINUS) bonds with p = 1 - exp(min(0, bond_arg)).  The draws come from the replica's host stream in the
// reference's order: direction, seed slice, seed site, then one uniform per bond with bond_arg < 0 (LIFO stack).
int build_and_flip_cluster(dqmc_ctx* ctx, RngStream& rng, double* phi, std::vector<unsigned char>& visited,
                           std::vector<int>& stack) {
    const int N = ctx->N, L = ctx->p.L, m = ctx->m, od = ctx->opdim;
    const double dtau = ctx->p.dtau;
    double rd[3] = {0, 0, 0};
    if (od == 1) {
        rd[0] = rng.draw() <= 0.5 ? -1.0 : 1.0;                            // randomDirection<1>, :3775-3785
    } else if (od == 2) {
        const double ang = rng.draw_range(0.0, 2.0 * M_PI);                // randPointOnCircle, rngwrapper.h:82-87
        rd[0] = std::cos(ang); rd[1] = std::sin(ang);
    } else {
        const double ang = rng.draw_range(0.0, 2.0 * M_PI);                // randPointOnSphere, rngwrapper.h:70-80
        const double costheta = rng.draw_range(-1.0, 1.0);
        const double sintheta = std::sqrt(1.0 - costheta * costheta);
        rd[0] = std::cos(ang) * sintheta; rd[1] = std::sin(ang) * sintheta; rd[2] = costheta;
    }
    auto at = [&](int site, int k, int d) -> double& { return phi[(size_t(k) * od + d) * N + site]; };
    auto proj = [&](int site, int k) {
        double v = 0;
        for (int d = 0; d < od; ++d) v += at(site, k, d) * rd[d];
        return v;
    };
    auto flip = [&](int site, int k) {
        const double pr = proj(site, k);
        for (int d = 0; d < od; ++d) at(site, k, d) = at(site, k, d) - 2.0 * pr * rd[d];
    };
    visited.assign(size_t(N) * (m + 1), 0);
    stack.clear();
    int k = 1 + int((m - 1 + 1.0) * rng.draw());                            // randInt(1, m), rngwrapper.h:66-68
    int site = int((N - 1 + 1.0) * rng.draw());                             // randInt(0, N-1)
    flip(site, k);
    visited[size_t(k) * N + site] = 1;
    stack.push_back(k * N + site);
    int size = 1;
    while (!stack.empty()) {
        const int top = stack.back();
        stack.pop_back();
        site = top % N; k = top / N;
        const int x = site % L, y = site / L;
        const int nbs[4] = {y * L + (x + 1) % L, y * L + (x + L - 1) % L, ((y + 1) % L) * L + x, ((y + L - 1) % L) * L + x};
        for (int d = 0; d < 4; ++d) {
            const int nb = nbs[d];
            if (visited[size_t(k) * N + nb]) continue;
            const double bond = 2.0 * dtau * proj(site, k) * proj(nb, k);
            if (bond < 0 && rng.draw() <= (1.0 - std::exp(bond))) {
                flip(nb, k);
                visited[size_t(k) * N + nb] = 1;
                stack.push_back(k * N + nb);
                ++size;
            }
        }
        const int kns[2] = {k < m ? k + 1 : 1, k > 1 ? k - 1 : m};
        for (int t = 0; t < 2; ++t) {
            const int kn = kns[t];
            if (visited[size_t(kn) * N + site]) continue;
            const double bond = (2.0 / dtau) * proj(site, k) * proj(site, kn);
            if (bond < 0 && rng.draw() <= (1.0 - std::exp(bond))) {
                flip(site, kn);
                visited[size_t(kn) * N + site] = 1;
                stack.push_back(kn * N + site);
                ++size;
            }
        }
    }
    return size;
}

// attemptGlobalShiftMove (kind 0, detsdwopdim.cpp:3564-3645), attemptWolffClusterUpdate (kind 1, :3487-3562) and
// attemptWolffClusterShiftUpdate (kind 2, :3647-3748) for the whole batch.  The cluster construction works on the
// fields (O(N m) per replica, host side, with the replica's own random-number stream); everything that costs
// O(D^3) -- the full re-setup of the UDT storage and G -- is the batched device path.
int global_move_kind(dqmc_ctx* ctx, int kind, int32_t* accepted_out) {
    const int R = ctx->R, D = ctx->D;
    if (ctx->currentTimeslice != ctx->m) { ctx->err = "global move: currentTimeslice != m"; return DQMC_ERR_STATE; }
    double* h = ctx->h_scalars;      // [0,R): old action, [R,2R): new action, [2R,3R) old logdet, [3R,4R) new
    if (kind == 0) {
        CKL(launch_phi_action(ctx->phi, ctx->rvals, ctx->actions, ctx->p.L, ctx->opdim, ctx->m, ctx->p.dtau, ctx->p.c,
                              ctx->p.u, (long long)phi_stride(ctx), R, ctx->stream));
        CK(cudaMemcpyAsync(h, ctx->actions, sizeof(double) * R, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaMemcpyAsync(h + 2 * R, ctx->logdet, sizeof(double) * R, cudaMemcpyDeviceToHost, ctx->stream));
    // backups (globalMoveStoreBackups, :3885-3900): copy the fields, swap everything that is recomputed
    CK(cudaMemcpyAsync(ctx->bkPhi, ctx->phi, sizeof(double) * phi_stride(ctx) * R, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->bkCosh, ctx->coshT, sizeof(double) * tab_stride(ctx) * R, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->bkSinh, ctx->sinhT, sizeof(double) * tab_stride(ctx) * R, cudaMemcpyDeviceToDevice, ctx->stream));
    // everything that is recomputed is copied, not pointer-swapped: the device addresses stay fixed, which the
    // captured sweep graphs rely on (≈ 2 GB of device-to-device copies per attempt at the headline size)
    {
        const size_t nm = ctx->nmat;
        CK(cudaMemcpyAsync(ctx->bkG, ctx->G, sizeof(cplx) * DD(ctx) * nm, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->bkQ, ctx->stQ, sizeof(cplx) * size_t(st_stride(ctx)) * nm, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->bkT, ctx->stT, sizeof(cplx) * size_t(st_stride(ctx)) * nm, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->bkD, ctx->stD, sizeof(double) * size_t(std_stride(ctx)) * nm, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->bkLogdet, ctx->logdet, sizeof(double) * nm, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    std::vector<double> clusterSize(R, 0.0);
    if (kind != 0) {
        // cluster flips on the host copy of the fields (repeatWolffPerSweep clusters per replica), then back to the device
        const size_t ps = phi_stride(ctx);
        std::vector<double> hphi(ps * R);
        CK(cudaMemcpyAsync(hphi.data(), ctx->phi, sizeof(double) * ps * R, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::vector<unsigned char> visited;
        std::vector<int> stack;
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < std::max(1, ctx->p.repeatWolffPerSweep); ++c)
                clusterSize[r] += build_and_flip_cluster(ctx, ctx->rng[r], hphi.data() + ps * r, visited, stack);
        CK(cudaMemcpyAsync(ctx->phi, hphi.data(), sizeof(double) * ps * R, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));            // hphi goes out of scope
    }
    if (kind == 2) {
        // the bosonic action difference of the combined move is that of the shift alone (:3684-3692)
        CKL(launch_phi_action(ctx->phi, ctx->rvals, ctx->actions, ctx->p.L, ctx->opdim, ctx->m, ctx->p.dtau, ctx->p.c,
                              ctx->p.u, (long long)phi_stride(ctx), R, ctx->stream));
        CK(cudaMemcpyAsync(h, ctx->actions, sizeof(double) * R, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (kind != 1) {
        // addGlobalRandomDisplacement (:3755-3763): OPDIM draws of randRange(-phiDelta, +phiDelta)
        for (int r = 0; r < R; ++r) {
            const double pd = ctx->ctrl_host[r].phiDelta;
            for (int d = 0; d < 3; ++d) h[4 * R + 3 * r + d] = 0.0;
            for (int d = 0; d < ctx->opdim; ++d) h[4 * R + 3 * r + d] = ctx->rng[r].draw_range(-pd, +pd);
        }
        CK(cudaMemcpyAsync(ctx->shiftbuf, h + 4 * R, sizeof(double) * 3 * R, cudaMemcpyHostToDevice, ctx->stream));
        CKL(launch_shift_fields(ctx->phi, ctx->shiftbuf, ctx->N, ctx->opdim, ctx->m, (long long)phi_stride(ctx), R, ctx->stream));
    }
    CKL(launch_update_tables(ctx->phi, ctx->coshT, ctx->sinhT, ctx->N, ctx->opdim, ctx->m, ctx->p.lambda * ctx->p.dtau,
                             (long long)phi_stride(ctx), (long long)tab_stride(ctx), R, ctx->stream));
    RET(setup_storage(ctx, 0, R));
    CKL(launch_phi_action(ctx->phi, ctx->rvals, ctx->actions, ctx->p.L, ctx->opdim, ctx->m, ctx->p.dtau, ctx->p.c,
                          ctx->p.u, (long long)phi_stride(ctx), R, ctx->stream));
    CK(cudaMemcpyAsync(h + R, ctx->actions, sizeof(double) * R, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h + 3 * R, ctx->logdet, sizeof(double) * R, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const size_t dd = DD(ctx);
    for (int r = 0; r < R; ++r) {
        const double probScalar = kind == 1 ? 1.0 : std::exp(-(h[R + r] - h[r]));
        double probFermion = std::exp(h[3 * R + r] - h[2 * R + r]);
        if (ctx->opdim < 3) probFermion = probFermion * probFermion;
        const double prob = probScalar * probFermion;
        ctx->lastGlobalProb[r] = prob;
        // UpdateStatistics counters (detsdwopdim.cpp:3487-3562, 3647-3748) live in the control data: they follow the
        // control parameter through dqmc_exchange_apply / dqmc_set_control_data and the checkpoint
        dqmc_control_data& cd = ctx->ctrl_host[r];
        if (kind == 0) cd.attemptedGlobalShifts += 1;
        else if (kind == 1) cd.attemptedWolffClusterUpdates += 1;
        else cd.attemptedWolffClusterShiftUpdates += 1;
        bool acc = prob >= 1.0 || ctx->rng[r].draw() < prob;
        if (accepted_out) accepted_out[r] = acc ? 1 : 0;
        if (acc) {
            if (kind == 0) cd.acceptedGlobalShifts += 1;
            else {
                if (kind == 1) cd.acceptedWolffClusterUpdates += 1;
                else cd.acceptedWolffClusterShiftUpdates += 1;
                cd.addedWolffClusterSize += clusterSize[r];
            }
        } else {
            // globalMoveRestoreBackups (:3902-3917) for this replica
            CK(cudaMemcpyAsync(ctx->phi + size_t(r) * phi_stride(ctx), ctx->bkPhi + size_t(r) * phi_stride(ctx),
                               sizeof(double) * phi_stride(ctx), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->coshT + size_t(r) * tab_stride(ctx), ctx->bkCosh + size_t(r) * tab_stride(ctx),
                               sizeof(double) * tab_stride(ctx), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->sinhT + size_t(r) * tab_stride(ctx), ctx->bkSinh + size_t(r) * tab_stride(ctx),
                               sizeof(double) * tab_stride(ctx), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->G + size_t(r) * dd, ctx->bkG + size_t(r) * dd, sizeof(cplx) * dd,
                               cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->stQ + size_t(r) * st_stride(ctx), ctx->bkQ + size_t(r) * st_stride(ctx),
                               sizeof(cplx) * st_stride(ctx), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->stT + size_t(r) * st_stride(ctx), ctx->bkT + size_t(r) * st_stride(ctx),
                               sizeof(cplx) * st_stride(ctx), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->stD + size_t(r) * std_stride(ctx), ctx->bkD + size_t(r) * std_stride(ctx),
                               sizeof(double) * std_stride(ctx), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->logdet + r, ctx->bkLogdet + r, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    (void)D;
    RET(upload_ctrl(ctx));           // the device copy of the control data carries the move counters too
    ctx->currentTimeslice = ctx->m;
    ctx->lastSweepDir = +1;
    return DQMC_OK;
}

// Run one sweep primitive for every lane (sub-batch) on the lane's own stream.  The lanes are independent
// groups of replicas; issuing them on separate streams lets the latency-bound kernels of one lane (the
// sequential update rounds, the QR panels: one CTA per replica) overlap with the throughput-bound kernels
// of the other (rank-K flushes, GEMMs, checkerboard multiplies) instead of leaving most SMs idle.
template <class F>
int for_each_lane(dqmc_ctx* ctx, F fn) {
    const int ts = ctx->currentTimeslice;
    cudaStream_t main = ctx->stream;
    int rc = DQMC_OK, ts_after = ts;
    for (int ln = 0; ln < ctx->nlanes && rc == DQMC_OK; ++ln) {
        ctx->laneOff = ctx->laneStart[ln];
        ctx->laneCnt = ctx->laneStart[ln + 1] - ctx->laneStart[ln];
        ctx->stream = ln == 0 ? main : ctx->laneStream[ln];
        ctx->currentTimeslice = ts;
        rc = fn();
        ts_after = ctx->currentTimeslice;
    }
    ctx->stream = main;
    ctx->laneOff = 0;
    ctx->laneCnt = ctx->R;
    ctx->currentTimeslice = ts_after;
    return rc;
}

// fork: the extra lane streams wait for everything issued on the main stream so far; join: the reverse
int lanes_fork(dqmc_ctx* ctx) {
    for (int ln = 1; ln < ctx->nlanes; ++ln) {
        CK(cudaEventRecord(ctx->laneEvent[0], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->laneStream[ln], ctx->laneEvent[0], 0));
    }
    return DQMC_OK;
}
int lanes_join(dqmc_ctx* ctx) {
    for (int ln = 1; ln < ctx->nlanes; ++ln) {
        CK(cudaEventRecord(ctx->laneEvent[ln], ctx->laneStream[ln]));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->laneEvent[ln], 0));
    }
    return DQMC_OK;
}

// sweepDown / sweepUp, detmodel.h:1261-1399
// ---- fermionic measurements (DetSDW::measure, detsdwopdim.cpp:540-900) ---------------------------------
size_t fm_acc_len(const dqmc_ctx* ctx) {
    const size_t nb = size_t(2 * ctx->p.L - 1) * (2 * ctx->p.L - 1);
    return 3 + 4 * nb + 2 * size_t(ctx->N);
}

int fm_prepare(dqmc_ctx* ctx) {
    if (ctx->fmAcc) return DQMC_OK;
    if (ctx->p.model == DQMC_MODEL_HUBBARD) {            // DetHubbard::measure needs no shift: five sums + zcorr[N]
        ctx->fmAccLen = 5 + size_t(ctx->N);
        CK(dmalloc(&ctx->fmAcc, ctx->fmAccLen * ctx->R));
        return DQMC_OK;
    }
    ctx->fmAccLen = fm_acc_len(ctx);
    CK(dmalloc(&ctx->fmAcc, ctx->fmAccLen * ctx->R));
    return DQMC_OK;
}

// measure(k) for the replicas of the current lane: gs = shiftGreenSymmetric(G) (:4505-4612) as two sparse checkerboard
// passes of cb_mult_kernel (half-step factors from the left and from the right), then the accumulation kernel
int measure_slice(dqmc_ctx* ctx) {
    const int ro = ctx->laneOff, rc = ctx->laneCnt;
    const size_t dd = DD(ctx);
    if (ctx->p.model == DQMC_MODEL_HUBBARD) {            // dethubbard.cpp:511-539
        CKL(hub_measure_launch(ctx->G + size_t(2 * ro) * dd, (long long)dd, ctx->N, ctx->p.L,
                               ctx->fmAcc + size_t(ro) * ctx->fmAccLen, (long long)ctx->fmAccLen, rc, ctx->stream));
        return DQMC_OK;
    }
    if (ctx->denseNow) { ctx->err = "fermionic measurements are served with the checkerboard break-up only"; return DQMC_ERR_STATE; }
    cplx* G = ctx->G + size_t(ro) * dd;
    cplx* gs = ctx->W[1] + size_t(ro) * dd;
    // gs = E0(-h) E1(-h) G E1(+h) E0(+h), h = dtau / 2: two checkerboard passes (the first one out of place)
    CbLaunch a;
    a.A = G; a.strideA = (long long)dd;
    a.phi = ctx->phi + size_t(ro) * phi_stride(ctx);
    a.coshT = ctx->coshT + size_t(ro) * tab_stride(ctx);
    a.sinhT = ctx->sinhT + size_t(ro) * tab_stride(ctx);
    a.stridePhi = (long long)phi_stride(ctx); a.strideTab = (long long)tab_stride(ctx);
    a.cbtab = ctx->cbtab;
    a.real_tables = ctx->p.weakZflux ? 0 : 1;
    a.kfirst = 1; a.kstep = 1; a.kcount = 1;
    a.rows = 0; a.k_then_v = 1; a.sign_idx = 0; a.transposed = 0;
    a.colscale = nullptr; a.strideScale = 0;
    a.batch = rc;
    a.shift = 1;
    a.out = gs; a.strideOut = (long long)dd;
    CKL(cb_launch(ctx->geom, a, ctx->stream));
    a.A = gs; a.out = nullptr;
    a.rows = 1; a.sign_idx = 1; a.transposed = 1;
    CKL(cb_launch(ctx->geom, a, ctx->stream));
    CKL(launch_fermion_measure(gs, (long long)dd, ctx->N, ctx->p.L, ctx->msf, ctx->fmAcc + size_t(ro) * ctx->fmAccLen,
                               (long long)ctx->fmAccLen, rc, ctx->stream));
    return DQMC_OK;
}

// `mode`: 0 sweep, 1 thermalisation sweep (step-size adaptation), 2 sweep with fermionic measurements after every slice
int sweep_down(dqmc_ctx* ctx, int therm) {
    const int n = ctx->n, s = ctx->s, m = ctx->m;
    RET(lanes_fork(ctx));
    for (int k = m; k >= (n - 1) * s + 1; --k) {
        RET(for_each_lane(ctx, [&] { RET(launch_update(ctx, k, therm == 1)); if (therm == 2) RET(measure_slice(ctx)); return wrap_down(ctx, k); }));
    }
    for (int l = n - 1; l >= 1; --l) {
        RET(for_each_lane(ctx, [&] { return advance_down(ctx, l + 1); }));
        for (int k = l * s; k >= (l - 1) * s + 1; --k) {
            RET(for_each_lane(ctx, [&] { RET(launch_update(ctx, k, therm == 1)); if (therm == 2) RET(measure_slice(ctx)); return wrap_down(ctx, k); }));
        }
    }
    RET(for_each_lane(ctx, [&] { return advance_down(ctx, 1); }));
    RET(lanes_join(ctx));
    return DQMC_OK;
}

int sweep_up(dqmc_ctx* ctx, int therm) {
    const int n = ctx->n, s = ctx->s, m = ctx->m;
    RET(set_storage_identity(ctx, 0, 0, ctx->nmat));
    RET(lanes_fork(ctx));
    for (int l = 0; l <= n - 2; ++l) {
        for (int k = l * s + 1; k <= (l + 1) * s; ++k) {
            RET(for_each_lane(ctx, [&] { RET(wrap_up(ctx, k - 1)); RET(launch_update(ctx, k, therm == 1)); return therm == 2 ? measure_slice(ctx) : DQMC_OK; }));
        }
        RET(for_each_lane(ctx, [&] { return advance_up(ctx, l); }));
    }
    for (int k = (n - 1) * s + 1; k <= m; ++k) {
        RET(for_each_lane(ctx, [&] { RET(wrap_up(ctx, k - 1)); RET(launch_update(ctx, k, therm == 1)); return therm == 2 ? measure_slice(ctx) : DQMC_OK; }));
    }
    RET(for_each_lane(ctx, [&] { return advance_up(ctx, n - 1); }));
    RET(lanes_join(ctx));
    return DQMC_OK;
}

// L2 residency of the Green's functions: G of all replicas is read and written by every update round (window block,
// gather, rank-K flush) and every wrap; with DQMC_L2_PERSIST=1 its address range is marked persisting in the launch
// attributes of the context's streams (stream attributes are recorded into captured kernel nodes), everything else
// streams through.  hitRatio = share of the range that fits the persisting carve-out.
int apply_l2_policy(dqmc_ctx* ctx, cudaStream_t st) {
    static const int on = std::getenv("DQMC_L2_PERSIST") ? std::atoi(std::getenv("DQMC_L2_PERSIST")) : 0;
    if (!on || !ctx->G) return DQMC_OK;
    int dev = 0, maxPersist = 0, maxWindow = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    CK(cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    if (maxPersist <= 0 || maxWindow <= 0) return DQMC_OK;
    const size_t gbytes = sizeof(cplx) * DD(ctx) * size_t(ctx->R) * ctx->ngc;
    const size_t window = std::min(gbytes, size_t(maxWindow));
    static bool limitSet = false;
    if (!limitSet) {
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(window, size_t(maxPersist))));
        limitSet = true;
        if (std::getenv("DQMC_L2_VERBOSE"))
            std::fprintf(stderr, "libdqmc_b200: L2 persisting carve-out %d MB max, window %zu MB of %zu MB\n", maxPersist >> 20,
                         window >> 20, gbytes >> 20);
    }
    cudaStreamAttrValue v;
    std::memset(&v, 0, sizeof v);
    v.accessPolicyWindow.base_ptr = ctx->G;
    v.accessPolicyWindow.num_bytes = window;
    const double ratioEnv = std::getenv("DQMC_L2_RATIO") ? std::atof(std::getenv("DQMC_L2_RATIO")) : 0.0;
    v.accessPolicyWindow.hitRatio = ratioEnv > 0 ? float(ratioEnv) : float(std::min(1.0, double(maxPersist) / double(window)));
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v));
    return DQMC_OK;
}

// One sweep direction as a CUDA graph: the launch sequence of a sweep is static (slice / round / panel loops
// with fixed trip counts, fixed device addresses), so it is captured once per (direction, thermalisation,
// random-number window, lanes, stabiliser) and replayed -- ~5000 kernel launches per sweep become one graph
// launch, which removes the host's launch-issue time from the step.
int run_sweep(dqmc_ctx* ctx, int dir, int therm) {
    auto direct = [&]() { return dir < 0 ? sweep_down(ctx, therm) : sweep_up(ctx, therm); };
    if (ctx->profiling || ctx->graphsOff) return direct();
    for (auto& g : ctx->graphs) {
        if (g.dir == dir && g.therm == therm && g.rngWindow == ctx->rngWindow && g.rngStride == (long long)ctx->rngStride &&
            g.nlanes == ctx->nlanes && g.stab == ctx->stabilizer && g.stream == ctx->stream) {
            CK(cudaGraphLaunch(g.exec, ctx->stream));
            ctx->launches += g.launches;
            ctx->currentTimeslice = dir < 0 ? 0 : ctx->m;
            return DQMC_OK;
        }
    }
    // capture
    const uint64_t l0 = ctx->launches + ctx->qr.launches;
    const int ts0 = ctx->currentTimeslice;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        (void)cudaGetLastError();
        ctx->graphsOff = true;
        return direct();
    }
    const int rc = direct();
    cudaGraph_t graph = nullptr;
    const cudaError_t ee = cudaStreamEndCapture(ctx->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc != DQMC_OK || ee != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        std::fprintf(stderr, "libdqmc_b200: sweep graph capture failed (rc=%d, %s); issuing launches directly\n", rc,
                     cudaGetErrorString(ee));
        ctx->graphsOff = true;                 // fall back to direct issue (same kernels, same order)
        ctx->currentTimeslice = ts0;
        return direct();
    }
    cudaGraphDestroy(graph);
    dqmc_ctx::SweepGraph sg;
    sg.dir = dir; sg.therm = therm; sg.rngWindow = ctx->rngWindow; sg.rngStride = (long long)ctx->rngStride;
    sg.nlanes = ctx->nlanes; sg.stab = ctx->stabilizer; sg.stream = ctx->stream; sg.exec = exec;
    sg.launches = ctx->launches + ctx->qr.launches - l0;
    ctx->graphs.push_back(sg);
    CK(cudaGraphLaunch(exec, ctx->stream));
    return DQMC_OK;
}

bool valid_rep(const dqmc_ctx* ctx, int rep) { return ctx && rep >= 0 && rep < ctx->R; }

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int dqmc_create(const dqmc_params* params, int n_replicas, int device, dqmc_ctx** out) {
    if (!params || !out || n_replicas <= 0) return DQMC_ERR_PARAM;
    *out = nullptr;
    dqmc_ctx* ctx = new (std::nothrow) dqmc_ctx();
    if (!ctx) return DQMC_ERR_PARAM;
    *out = ctx;                      // returned even on failure so the caller can read the error text
    ctx->p = *params;
    const dqmc_params& p = ctx->p;
    const bool hub = p.model == DQMC_MODEL_HUBBARD;
    if (p.model != DQMC_MODEL_SDW && !hub) { ctx->err = "unknown model"; return DQMC_ERR_PARAM; }
    if (hub) {
        // DetHubbard: real N x N Green's functions, two components; fields are the +-1 auxiliary spins
        ctx->p.opdim = 1; ctx->p.delaySteps = 1; ctx->p.globalShift = 0; ctx->p.weakZflux = 0;
        if (p.U < 0) { ctx->err = "DetHubbard needs U >= 0"; return DQMC_ERR_PARAM; }
        if (p.checkerboard && (p.L % 2)) { ctx->err = "checkerboard propagator needs an even L"; return DQMC_ERR_PARAM; }
    }
    if (p.opdim < 1 || p.opdim > 3) { ctx->err = "opdim must be 1, 2 or 3"; return DQMC_ERR_PARAM; }
    if (p.L < 2 || (!hub && p.L % 2)) { ctx->err = "checkerboard decomposition needs an even L >= 2"; return DQMC_ERR_PARAM; }
    if (p.weakZflux && p.opdim == 3) { ctx->err = "weakZflux is only supported for opdim < 3"; return DQMC_ERR_PARAM; }
    if (p.m < 2 || p.s < 1 || p.dtau <= 0) { ctx->err = "need m >= 2, s >= 1, dtau > 0"; return DQMC_ERR_PARAM; }
    if (p.delaySteps < 1 || p.delaySteps > p.L * p.L) { ctx->err = "delaySteps out of range"; return DQMC_ERR_PARAM; }
    if ((p.globalShift || p.wolffClusterUpdate || p.wolffClusterShiftUpdate) && p.globalUpdateInterval < 1) {
        ctx->err = "globalUpdateInterval must be >= 1"; return DQMC_ERR_PARAM;
    }
    if (p.wolffClusterShiftUpdate && (p.globalShift || p.wolffClusterUpdate)) {       // detsdwparams.cpp:94-96
        ctx->err = "either the combined wolffClusterShiftUpdate or the individual global updates"; return DQMC_ERR_PARAM;
    }
    if (hub && (p.wolffClusterUpdate || p.wolffClusterShiftUpdate)) { ctx->err = "Wolff cluster moves are defined for DetSDW"; return DQMC_ERR_PARAM; }
    ctx->R = n_replicas;
    ctx->opdim = p.opdim;
    ctx->N = p.L * p.L;
    ctx->msf = hub ? 1 : (p.opdim == 3 ? 4 : 2);
    ctx->D = ctx->msf * ctx->N;
    ctx->m = p.m;
    ctx->s = p.s;
    while (ctx->m <= ctx->s) ctx->s -= 1;                  // updateTemperatureParameters, detmodelparams.h:108-113
    ctx->n = (ctx->m + ctx->s - 1) / ctx->s;
    ctx->ngc = hub ? 2 : 1;
    ctx->nmat = ctx->R * ctx->ngc;
    ctx->device = device;
    ctx->launches = 0;
    ctx->currentTimeslice = 0;
    ctx->lastSweepDir = +1;
    ctx->performedSweeps = 0;
    ctx->rngWindow = 0;
    ctx->denseNow = !hub && p.denseHopping != 0;

    int ndev = 0;
    (void)cudaGetLastError();        // do not inherit a stale error from an earlier failed call
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { ctx->err = "no such CUDA device (this library has no CPU fallback)"; return DQMC_ERR_CUDA; }
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;

    ctx->geom.L = p.L; ctx->geom.N = ctx->N; ctx->geom.msf = ctx->msf; ctx->geom.D = ctx->D;
    ctx->geom.nplaq = ctx->N / 4; ctx->geom.opdim = p.opdim; ctx->geom.m = ctx->m;
    ctx->geom.lambda_dtau = p.lambda * p.dtau;
    ctx->umodel.L = p.L; ctx->umodel.N = ctx->N; ctx->umodel.msf = ctx->msf; ctx->umodel.D = ctx->D;
    ctx->umodel.opdim = p.opdim; ctx->umodel.m = ctx->m; ctx->umodel.delaySteps = p.delaySteps;
    ctx->umodel.dtau = p.dtau; ctx->umodel.c = p.c; ctx->umodel.u = p.u; ctx->umodel.lambda = p.lambda;
    ctx->umodel.accRatio = p.accRatio;

    const size_t dd = DD(ctx), nm = ctx->nmat, D = ctx->D, R = ctx->R;
    CK(dmalloc(&ctx->G, dd * nm));
    CK(dmalloc(&ctx->bkG, dd * nm));
    CK(dmalloc(&ctx->Gwrapped, dd * nm));
    for (int i = 0; i < 5; ++i) CK(dmalloc(&ctx->W[i], dd * nm));
    CK(qr_workspace_create(&ctx->qr, ctx->D, (int)nm));
    ctx->stabilizer = std::getenv("DQMC_STABILIZER_FULL_PIVOT") ? 1 : 0;
    CK(dmalloc(&ctx->tQ, dd * nm));
    CK(dmalloc(&ctx->tT, dd * nm));
    CK(dmalloc(&ctx->tD, D * nm));
    CK(dmalloc(&ctx->stQ, dd * nm * (ctx->n + 1)));
    CK(dmalloc(&ctx->stT, dd * nm * (ctx->n + 1)));
    CK(dmalloc(&ctx->stD, D * nm * (ctx->n + 1)));
    CK(dmalloc(&ctx->bkQ, dd * nm * (ctx->n + 1)));
    CK(dmalloc(&ctx->bkT, dd * nm * (ctx->n + 1)));
    CK(dmalloc(&ctx->bkD, D * nm * (ctx->n + 1)));
    CK(dmalloc(&ctx->phi, phi_stride(ctx) * R));
    CK(dmalloc(&ctx->coshT, tab_stride(ctx) * R));
    CK(dmalloc(&ctx->sinhT, tab_stride(ctx) * R));
    CK(dmalloc(&ctx->bkPhi, phi_stride(ctx) * R));
    CK(dmalloc(&ctx->bkCosh, tab_stride(ctx) * R));
    CK(dmalloc(&ctx->bkSinh, tab_stride(ctx) * R));
    CK(dmalloc(&ctx->rvals, R));
    CK(dmalloc(&ctx->tau, D * nm));
    CK(dmalloc(&ctx->perm, D * nm));
    CK(dmalloc(&ctx->colnorm, D * nm));
    CK(dmalloc(&ctx->vecA, D * nm));
    CK(dmalloc(&ctx->vecB, D * nm));
    CK(dmalloc(&ctx->vecC, D * nm));
    CK(dmalloc(&ctx->vecD, D * nm));
    CK(dmalloc(&ctx->dtmp, D * nm));
    CK(dmalloc(&ctx->logdet, nm));
    CK(dmalloc(&ctx->bkLogdet, nm));
    CK(dmalloc(&ctx->consistency, nm));
    CK(dmalloc(&ctx->eyeM, dd));
    CK(dmalloc(&ctx->onesV, D));
    ctx->kmax = (ctx->msf * ctx->p.delaySteps + 3) & ~3;        // padded to the 4-term chunks of the update kernel
    CK(dmalloc(&ctx->X, D * ctx->kmax * R));
    CK(dmalloc(&ctx->Y, D * ctx->kmax * R));
    CK(cudaMemsetAsync(ctx->X, 0, sizeof(cplx) * D * ctx->kmax * R, ctx->stream));   // finite everywhere (see extend_xy)
    CK(cudaMemsetAsync(ctx->Y, 0, sizeof(cplx) * D * ctx->kmax * R, ctx->stream));
    // window rounds (update_kernels.cu): per-round record of the accepted updates; DQMC_UPDATE_LEGACY=1 keeps the
    // round kernel of round 1 (development A/B switch)
    ctx->winSites = (hub || std::getenv("DQMC_UPDATE_LEGACY")) ? 0 : update_window_sites(ctx->umodel);
    ctx->winScratchStride = ctx->winSites ? update_window_scratch_elems(ctx->umodel) : 0;
    ctx->winHdrStride = ctx->winSites ? update_window_hdr_ints(ctx->umodel) : 0;
    if (ctx->winSites) {
        CK(dmalloc(&ctx->winScratch, ctx->winScratchStride * R));
        CK(dmalloc(&ctx->winHdr, size_t(ctx->winHdrStride) * R));
        CK(cudaMemsetAsync(ctx->winScratch, 0, sizeof(cplx) * ctx->winScratchStride * R, ctx->stream));
        CK(cudaMemsetAsync(ctx->winHdr, 0, sizeof(int) * size_t(ctx->winHdrStride) * R, ctx->stream));
    }
    ctx->rngCap = size_t(ctx->m) * ctx->N * (ctx->p.opdim + 1) * size_t(std::max(1, hub ? 1 : ctx->p.repeatUpdateInSlice));   // Hubbard: <= 2 values per attempt
    ctx->rngAlloc = ctx->rngCap * kStreamSweeps;               // streamed mode keeps several sweeps' worth on the device
    ctx->rngStride = ctx->rngCap;
    ctx->rngAuto = false;
    ctx->rngUploaded = 0;
    ctx->copyHalf = 0;
    ctx->rngResident = false;
    ctx->rngResidentUsedBound = 0;
    ctx->profiling = false;
    ctx->profForce = -1;
    ctx->graphsOff = std::getenv("DQMC_NO_GRAPHS") != nullptr;
    ctx->laneOff = 0;
    ctx->laneCnt = ctx->R;
    ctx->nlanes = 1;
    for (int i = 0; i <= DQMC_MAX_LANES; ++i) ctx->laneStart[i] = i == 0 ? 0 : ctx->R;
    for (int i = 0; i < DQMC_MAX_LANES; ++i) {
        ctx->laneStream[i] = nullptr;
        CK(cudaEventCreateWithFlags(&ctx->laneEvent[i], cudaEventDisableTiming));
        if (i > 0) CK(cudaStreamCreateWithFlags(&ctx->laneStream[i], cudaStreamNonBlocking));
    }
    {
        // one lane per replica (up to 32 lanes; DQMC_LANES / dqmc_set_option override): replicas of a lane advance in lockstep, so a round or a
        // panel takes as long as its slowest replica; separate lanes remove that coupling and let the latency-
        // bound kernels of one replica overlap with the throughput-bound kernels of the others
        int want = std::min(ctx->R, 32);                  // measured at 64 replicas: 32 lanes 120.6 ms, 64 lanes 121.8 ms per step
        gemm_set_matrices_in_flight(ctx->R * ctx->ngc);
        pdl_set_enabled(true);                            // DQMC_PDL=0 switches the launch attribute off
        if (const char* e = std::getenv("DQMC_LANES")) want = std::atoi(e);
        want = std::max(1, std::min(want, std::min(DQMC_MAX_LANES, ctx->R)));
        ctx->nlanes = want;
        for (int i = 0; i <= DQMC_MAX_LANES; ++i) ctx->laneStart[i] = i >= want ? ctx->R : (ctx->R * i) / want;
    }
    RET(apply_l2_policy(ctx, ctx->stream));
    for (int i = 1; i < DQMC_MAX_LANES; ++i) RET(apply_l2_policy(ctx, ctx->laneStream[i]));
    CK(dmalloc(&ctx->rngbuf, ctx->rngAlloc * R));
    CK(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&ctx->copyEvent[i], cudaEventDisableTiming));
    CK(dmalloc(&ctx->acceptedTotal, R));
    CK(cudaMemsetAsync(ctx->acceptedTotal, 0, sizeof(unsigned long long) * R, ctx->stream));
    CK(dmalloc(&ctx->cursor, R));
    CK(dmalloc(&ctx->ctrl, R));
    CK(dmalloc(&ctx->accepted, R));
    CK(dmalloc(&ctx->siteState, R));
    CK(dmalloc(&ctx->cursorAdd, R));
    CK(dmalloc(&ctx->kvec, R));
    CK(dmalloc(&ctx->errflag, 1));
    CK(dmalloc(&ctx->actions, R));
    CK(dmalloc(&ctx->shiftbuf, 3 * R));
    CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_rng), 2 * ctx->rngCap * R * sizeof(double)));   // two staging chunks
    ctx->hRngAlloc = 2 * ctx->rngCap * R;
    CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_cursor), R * sizeof(int)));
    CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_scalars), (8 * R + 16) * sizeof(double)));
    CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_ctrl), R * sizeof(dqmc_control_data)));
    CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_err), sizeof(int)));
    CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_acc), R * sizeof(uint32_t)));

    ctx->aux = nullptr; ctx->propT = nullptr; ctx->propTinv = nullptr; ctx->hubScale = nullptr; ctx->hubTmp = nullptr;
    ctx->hubReal = nullptr; ctx->cbtab = nullptr; ctx->hubAlpha = 0;
    std::vector<cplx> tab;
    if (!hub) {
        cb_build_tables(p, tab);
        CK(dmalloc(&ctx->cbtab, tab.size()));
        CK(cudaMemcpyAsync(ctx->cbtab, tab.data(), tab.size() * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        ctx->hubAlpha = std::acosh(std::exp(p.dtau * p.U * 0.5));           // dethubbard.cpp:55
        std::vector<double> P, Pinv;
        hub_build_propagators(p, P, Pinv);
        std::vector<cplx> pc(dd), pic(dd);
        for (size_t i = 0; i < dd; ++i) { pc[i] = make_double2(P[i], 0); pic[i] = make_double2(Pinv[i], 0); }
        CK(dmalloc(&ctx->propT, dd));
        CK(dmalloc(&ctx->propTinv, dd));
        CK(dmalloc(&ctx->hubScale, D * nm));
        CK(dmalloc(&ctx->hubTmp, dd * nm));
        CK(dmalloc(&ctx->hubReal, dd));
        CK(dmalloc(&ctx->aux, tab_stride(ctx) * R));
        CK(cudaMemsetAsync(ctx->aux, 0, sizeof(int32_t) * tab_stride(ctx) * R, ctx->stream));
        CK(cudaMemcpyAsync(ctx->propT, pc.data(), dd * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->propTinv, pic.data(), dd * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaMemsetAsync(ctx->phi, 0, sizeof(double) * phi_stride(ctx) * R, ctx->stream));
    CK(cudaMemsetAsync(ctx->coshT, 0, sizeof(double) * tab_stride(ctx) * R, ctx->stream));
    CK(cudaMemsetAsync(ctx->sinhT, 0, sizeof(double) * tab_stride(ctx) * R, ctx->stream));
    CK(cudaMemsetAsync(ctx->errflag, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(ctx->G, 0, sizeof(cplx) * dd * nm, ctx->stream));
    CKL(launch_set_identity(ctx->eyeM, ctx->D, 0, 1, ctx->stream));
    {
        std::vector<double> ones(D, 1.0);
        CK(cudaMemcpyAsync(ctx->onesV, ones.data(), sizeof(double) * D, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->rng.resize(R);
    ctx->h_r.assign(R, p.r);
    ctx->lastGlobalProb.assign(R, 0.0);
    ctx->ctrl_host.resize(R);
    for (size_t r = 0; r < R; ++r) {
        ctx->rng[r].seed(0, (uint32_t)r);
        dqmc_control_data& c = ctx->ctrl_host[r];
        std::memset(&c, 0, sizeof c);
        c.phiDelta = 0.5;                                   // AdjustmentData::InitialPhiDelta
    }
    RET(upload_ctrl(ctx));
    RET(upload_rvals(ctx));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

void dqmc_destroy(dqmc_ctx* ctx) {
    if (!ctx) return;
    if (ctx->stream) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    void* dev[] = {ctx->G, ctx->bkG, ctx->Gwrapped, ctx->W[0], ctx->W[1], ctx->W[2], ctx->W[3], ctx->W[4], ctx->tQ, ctx->tT, ctx->tD,
                   ctx->stQ, ctx->stT, ctx->stD, ctx->bkQ, ctx->bkT, ctx->bkD, ctx->phi, ctx->coshT, ctx->sinhT,
                   ctx->bkPhi, ctx->bkCosh, ctx->bkSinh, ctx->rvals, ctx->tau, ctx->perm, ctx->colnorm, ctx->vecA,
                   ctx->vecB, ctx->vecC, ctx->vecD, ctx->dtmp, ctx->logdet, ctx->bkLogdet, ctx->consistency, ctx->eyeM,
                   ctx->onesV, ctx->X, ctx->Y, ctx->winScratch, ctx->winHdr, ctx->cfgStream, ctx->shiftL, ctx->shiftR, ctx->denseP, ctx->densePinv, ctx->denseTmp, ctx->fmAcc, ctx->rngbuf, ctx->cursor, ctx->ctrl, ctx->accepted, ctx->errflag,
                   ctx->actions, ctx->shiftbuf, ctx->cbtab, ctx->acceptedTotal, ctx->siteState, ctx->cursorAdd, ctx->kvec, ctx->aux,
                   ctx->propT, ctx->propTinv, ctx->hubScale, ctx->hubTmp, ctx->hubReal};
    for (void* p : dev) if (p) cudaFree(p);
    qr_workspace_destroy(&ctx->qr);
    void* host[] = {ctx->h_rng, ctx->h_cursor, ctx->h_scalars, ctx->h_ctrl, ctx->h_err, ctx->h_acc};
    for (void* p : host) if (p) cudaFreeHost(p);
    if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); }
    for (int i = 0; i < 2; ++i) if (ctx->copyEvent[i]) cudaEventDestroy(ctx->copyEvent[i]);
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
    prof_collect(ctx);
    for (cudaEvent_t e : ctx->profPool) cudaEventDestroy(e);
    for (int i = 1; i < DQMC_MAX_LANES; ++i)
        if (ctx->laneStream[i]) { cudaStreamSynchronize(ctx->laneStream[i]); cudaStreamDestroy(ctx->laneStream[i]); }
    for (int i = 0; i < DQMC_MAX_LANES; ++i) if (ctx->laneEvent[i]) cudaEventDestroy(ctx->laneEvent[i]);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* dqmc_last_error(const dqmc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int dqmc_set_stream(dqmc_ctx* ctx, void* cuda_stream) {
    if (!ctx) return DQMC_ERR_PARAM;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    if (cuda_stream) {
        ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    RET(apply_l2_policy(ctx, ctx->stream));
    return DQMC_OK;
}

int dqmc_synchronize(dqmc_ctx* ctx) {
    if (!ctx) return DQMC_ERR_PARAM;
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_dims(const dqmc_ctx* ctx, int32_t* out) {
    if (!ctx || !out) return DQMC_ERR_PARAM;
    out[0] = ctx->N; out[1] = ctx->D; out[2] = ctx->m; out[3] = ctx->n; out[4] = ctx->s; out[5] = ctx->ngc;
    out[6] = ctx->R; out[7] = ctx->opdim;
    return DQMC_OK;
}

int dqmc_set_option(dqmc_ctx* ctx, int option, int value) {
    if (!ctx) return DQMC_ERR_PARAM;
    if (option == DQMC_OPT_STABILIZER && (value == DQMC_STAB_PREPIVOT_BLOCKED || value == DQMC_STAB_FULL_PIVOT)) {
        ctx->stabilizer = value;
        return DQMC_OK;
    }
    if (option == DQMC_OPT_LANES && value >= 1 && value <= DQMC_MAX_LANES) {
        if (value > ctx->R) { ctx->err = "more lanes than replicas"; return DQMC_ERR_PARAM; }
        ctx->nlanes = value;
        for (int i = 0; i <= DQMC_MAX_LANES; ++i) ctx->laneStart[i] = i >= value ? ctx->R : (ctx->R * i) / value;
        return DQMC_OK;
    }
    ctx->err = "dqmc_set_option: unknown option or value";
    return DQMC_ERR_PARAM;
}

uint64_t dqmc_launch_count(const dqmc_ctx* ctx) { return ctx ? ctx->launches + ctx->qr.launches : 0; }

// ---- RNG ---------------------------------------------------------------------------------------
int dqmc_rng_seed(dqmc_ctx* ctx, int rep, uint32_t seed, uint32_t process_index) {
    if (!valid_rep(ctx, rep)) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    ctx->rng[rep].seed(seed, process_index);
    return DQMC_OK;
}
int dqmc_rng_set_source(dqmc_ctx* ctx, int rep, dqmc_rng_fill_fn fill, void* user) {
    if (!valid_rep(ctx, rep) || !fill) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    ctx->rng[rep].set_source(fill, user);
    return DQMC_OK;
}
int dqmc_rng_draw(dqmc_ctx* ctx, int rep, size_t n, double* out) {
    if (!valid_rep(ctx, rep) || (n && !out)) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    for (size_t i = 0; i < n; ++i) out[i] = ctx->rng[rep].draw();
    return DQMC_OK;
}
// Look-ahead of a replica's stream: the values already drawn from the source (the driver's RngWrapper) but not yet
// consumed.  They belong to the replica's state: the reference checkpoints model and generator together
// (detsdwopdim.h:1127-1148 + the driver's rng state), here the generator has run ahead by exactly these values.
// out == NULL: *n receives the count; otherwise up to *n values are copied and *n receives the number copied.
int dqmc_rng_look_ahead(dqmc_ctx* ctx, int rep, double* out, size_t* n) {
    if (!valid_rep(ctx, rep) || !n) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    const size_t have = ctx->rng[rep].buffered();
    if (!out) { *n = have; return DQMC_OK; }
    const size_t cnt = std::min(*n, have);
    std::memcpy(out, ctx->rng[rep].buffered_values(), cnt * sizeof(double));
    *n = cnt;
    return DQMC_OK;
}
int dqmc_rng_set_look_ahead(dqmc_ctx* ctx, int rep, const double* values, size_t n) {
    if (!valid_rep(ctx, rep) || (n && !values)) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    ctx->rng[rep].set_buffered(values, n);
    return DQMC_OK;
}
// performedSweeps of the reference's model state (detsdwopdim.h:1127-1148): fixes the phase of the global-move schedule
// (performedSweeps % globalUpdateInterval, detmodel.h:1422-1424) after a resume
int dqmc_set_performed_sweeps(dqmc_ctx* ctx, uint32_t n) {
    if (!ctx) return DQMC_ERR_PARAM;
    ctx->performedSweeps = n;
    return DQMC_OK;
}
int dqmc_rng_peek(dqmc_ctx* ctx, int rep, size_t n, double* out) {
    if (!valid_rep(ctx, rep) || (n && !out)) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    const double* p = ctx->rng[rep].peek(n);
    std::memcpy(out, p, n * sizeof(double));
    return DQMC_OK;
}
int dqmc_rng_skip(dqmc_ctx* ctx, int rep, size_t n) {
    if (!valid_rep(ctx, rep)) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    ctx->rng[rep].skip(n);
    return DQMC_OK;
}
uint64_t dqmc_rng_consumed(const dqmc_ctx* ctx, int rep) {
    if (!valid_rep(ctx, rep)) return 0;
    if (host_sync_rng(const_cast<dqmc_ctx*>(ctx)) != DQMC_OK) return 0;
    return ctx->rng[rep].consumed();
}
int dqmc_rng_stream_sample(uint32_t seed, uint32_t process_index, size_t n, double* out) {
    if (n && !out) return DQMC_ERR_PARAM;
    RngStream g;
    g.seed(seed, process_index);
    for (size_t i = 0; i < n; ++i) out[i] = g.draw();
    return DQMC_OK;
}

// ---- state -------------------------------------------------------------------------------------
int dqmc_upload_fields(dqmc_ctx* ctx, int rep, const void* fields) {
    if (!valid_rep(ctx, rep) || !fields) return DQMC_ERR_PARAM;
    if (ctx->p.model == DQMC_MODEL_HUBBARD) {
        CK(cudaMemcpyAsync(ctx->aux + size_t(rep) * tab_stride(ctx), fields, sizeof(int32_t) * tab_stride(ctx),
                           cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return DQMC_OK;
    }
    CK(cudaMemcpyAsync(ctx->phi + size_t(rep) * phi_stride(ctx), fields, sizeof(double) * phi_stride(ctx),
                       cudaMemcpyHostToDevice, ctx->stream));
    CKL(launch_update_tables(ctx->phi + size_t(rep) * phi_stride(ctx), ctx->coshT + size_t(rep) * tab_stride(ctx),
                             ctx->sinhT + size_t(rep) * tab_stride(ctx), ctx->N, ctx->opdim, ctx->m,
                             ctx->p.lambda * ctx->p.dtau, (long long)phi_stride(ctx), (long long)tab_stride(ctx), 1,
                             ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_init_random_fields(dqmc_ctx* ctx, int rep) {
    if (!valid_rep(ctx, rep)) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    if (ctx->p.model == DQMC_MODEL_HUBBARD) {
        // setupRandomAuxfield, dethubbard.cpp:741-751: per (k, site) one rand01(), <= 0.5 -> +1
        std::vector<int32_t> aux(tab_stride(ctx), 0);
        RngStream& g = ctx->rng[rep];
        for (int k = 1; k <= ctx->m; ++k)
            for (int site = 0; site < ctx->N; ++site) aux[size_t(k) * ctx->N + site] = g.draw() <= 0.5 ? +1 : -1;
        return dqmc_upload_fields(ctx, rep, aux.data());
    }
    // setupRandomField, detsdwopdim.cpp:1098-1113: per (k, site): OPDIM x randRange(-1, 1), then one
    // rand01() for cdwl (drawn even though cdwU == 0)
    std::vector<double> phi(phi_stride(ctx), 0.0);
    RngStream& g = ctx->rng[rep];
    for (int k = 1; k <= ctx->m; ++k)
        for (int site = 0; site < ctx->N; ++site) {
            for (int d = 0; d < ctx->opdim; ++d)
                phi[(size_t(k) * ctx->opdim + d) * ctx->N + site] = g.draw_range(-1.0, 1.0);
            (void)g.draw();
        }
    return dqmc_upload_fields(ctx, rep, phi.data());
}

int dqmc_download_fields(dqmc_ctx* ctx, int rep, void* fields) {
    if (!valid_rep(ctx, rep) || !fields) return DQMC_ERR_PARAM;
    if (ctx->p.model == DQMC_MODEL_HUBBARD) {
        CK(cudaMemcpyAsync(fields, ctx->aux + size_t(rep) * tab_stride(ctx), sizeof(int32_t) * tab_stride(ctx),
                           cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return DQMC_OK;
    }
    CK(cudaMemcpyAsync(fields, ctx->phi + size_t(rep) * phi_stride(ctx), sizeof(double) * phi_stride(ctx),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

// getCurrentSystemConfiguration / saveConfigurationStream* (detsdwopdim.cpp:4943-5036, 5116-5122): the fields of
// one replica (rep >= 0) or of all replicas (rep = -1, replica-major) in the order of the configuration streams
int dqmc_download_config_stream(dqmc_ctx* ctx, int rep, double* out) {
    if (!ctx || !out || rep < -1 || rep >= ctx->R) return DQMC_ERR_PARAM;
    if (ctx->p.model != DQMC_MODEL_SDW) { ctx->err = "configuration streams are defined for DetSDW"; return DQMC_ERR_STATE; }
    const size_t len = size_t(ctx->N) * ctx->m * ctx->opdim;
    if (!ctx->cfgStream) CK(dmalloc(&ctx->cfgStream, len * ctx->R));
    const int first = rep < 0 ? 0 : rep, count = rep < 0 ? ctx->R : 1;
    CKL(launch_config_stream(ctx->phi + size_t(first) * phi_stride(ctx), ctx->cfgStream + size_t(first) * len, ctx->p.L, ctx->opdim,
                             ctx->m, (long long)phi_stride(ctx), (long long)len, count, ctx->stream));
    CK(cudaMemcpyAsync(out, ctx->cfgStream + size_t(first) * len, sizeof(double) * len * count, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_download_green(dqmc_ctx* ctx, int rep, int gc, double* out) {
    if (!valid_rep(ctx, rep) || gc < 0 || gc >= ctx->ngc || !out) return DQMC_ERR_PARAM;
    const cplx* src = ctx->G + (size_t(rep) * ctx->ngc + gc) * DD(ctx);
    if (ctx->p.model == DQMC_MODEL_HUBBARD) {
        CKL(hub_real_part_launch(src, ctx->hubReal, DD(ctx), ctx->stream));
        CK(cudaMemcpyAsync(out, ctx->hubReal, sizeof(double) * DD(ctx), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        CK(cudaMemcpyAsync(out, src, sizeof(cplx) * DD(ctx), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_upload_green(dqmc_ctx* ctx, int rep, int gc, const double* in) {
    if (!valid_rep(ctx, rep) || gc < 0 || gc >= ctx->ngc || !in) return DQMC_ERR_PARAM;
    cplx* dst = ctx->G + (size_t(rep) * ctx->ngc + gc) * DD(ctx);
    if (ctx->p.model == DQMC_MODEL_HUBBARD) {
        CK(cudaMemcpyAsync(ctx->hubReal, in, sizeof(double) * DD(ctx), cudaMemcpyHostToDevice, ctx->stream));
        CKL(hub_to_complex_launch(ctx->hubReal, dst, DD(ctx), ctx->stream));
    } else {
        CK(cudaMemcpyAsync(dst, in, sizeof(cplx) * DD(ctx), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_set_exchange_parameter(dqmc_ctx* ctx, int rep, double r) {
    if (!valid_rep(ctx, rep)) return DQMC_ERR_PARAM;
    ctx->h_r[rep] = r;
    return upload_rvals(ctx);
}
int dqmc_get_exchange_parameter(dqmc_ctx* ctx, int rep, double* r) {
    if (!valid_rep(ctx, rep) || !r) return DQMC_ERR_PARAM;
    *r = ctx->h_r[rep];
    return DQMC_OK;
}
int dqmc_get_control_data(dqmc_ctx* ctx, int rep, dqmc_control_data* out) {
    if (!valid_rep(ctx, rep) || !out) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    *out = ctx->ctrl_host[rep];
    return DQMC_OK;
}
int dqmc_set_control_data(dqmc_ctx* ctx, int rep, const dqmc_control_data* in) {
    if (!valid_rep(ctx, rep) || !in) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    ctx->ctrl_host[rep] = *in;
    return upload_ctrl(ctx);
}
int dqmc_get_sweep_state(const dqmc_ctx* ctx, int32_t* out) {
    if (!ctx || !out) return DQMC_ERR_PARAM;
    out[0] = ctx->currentTimeslice; out[1] = ctx->lastSweepDir; out[2] = ctx->performedSweeps;
    return DQMC_OK;
}

// ---- operators ---------------------------------------------------------------------------------
int dqmc_bmat_mult(dqmc_ctx* ctx, int rep, int gc, int op, double* A_host, uint32_t k2, uint32_t k1) {
    if (!valid_rep(ctx, rep) || gc < 0 || gc >= ctx->ngc || op < 0 || op > 4 || !A_host || k2 <= k1 || (int)k2 > ctx->m)
        return DQMC_ERR_PARAM;
    const int mat = rep * ctx->ngc + gc;
    cplx* buf = ctx->W[0] + size_t(mat) * DD(ctx);
    const bool hub = ctx->p.model == DQMC_MODEL_HUBBARD;
    if (hub) {
        CK(cudaMemcpyAsync(ctx->hubReal, A_host, sizeof(double) * DD(ctx), cudaMemcpyHostToDevice, ctx->stream));
        CKL(hub_to_complex_launch(ctx->hubReal, buf, DD(ctx), ctx->stream));
    } else {
        CK(cudaMemcpyAsync(buf, A_host, sizeof(cplx) * DD(ctx), cudaMemcpyHostToDevice, ctx->stream));
    }
    RET(sdw_bmult(ctx, op, buf, (long long)DD(ctx), (int)k2, (int)k1, nullptr, 0, mat, 1));
    if (hub) {
        CKL(hub_real_part_launch(buf, ctx->hubReal, DD(ctx), ctx->stream));
        CK(cudaMemcpyAsync(A_host, ctx->hubReal, sizeof(double) * DD(ctx), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        CK(cudaMemcpyAsync(A_host, buf, sizeof(cplx) * DD(ctx), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_bmat_mult_device(dqmc_ctx* ctx, int gc, int op, void* A_dev, uint32_t k2, uint32_t k1) {
    if (!ctx || gc != 0 || op < 0 || op > 4 || !A_dev || k2 <= k1 || (int)k2 > ctx->m) return DQMC_ERR_PARAM;
    return sdw_bmult(ctx, op, static_cast<cplx*>(A_dev), (long long)DD(ctx), (int)k2, (int)k1, nullptr, 0, 0, ctx->nmat);
}

int dqmc_bench_bmat_mult(dqmc_ctx* ctx, int op, void* A_dev, uint32_t k2, uint32_t k1, int reps, float* ms_per_launch) {
    if (!ctx || !A_dev || reps < 1 || !ms_per_launch) return DQMC_ERR_PARAM;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < reps; ++i) RET(dqmc_bmat_mult_device(ctx, 0, op, A_dev, k2, k1));
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return DQMC_OK;
}

int dqmc_setup_storage(dqmc_ctx* ctx) {
    if (!ctx) return DQMC_ERR_PARAM;
    RET(setup_storage(ctx, 0, ctx->nmat));
    ctx->currentTimeslice = ctx->m;
    ctx->lastSweepDir = +1;
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_wrap_up(dqmc_ctx* ctx, uint32_t k) { return ctx ? wrap_up(ctx, (int)k) : DQMC_ERR_PARAM; }
int dqmc_wrap_down(dqmc_ctx* ctx, uint32_t k) { return ctx ? wrap_down(ctx, (int)k) : DQMC_ERR_PARAM; }
int dqmc_advance_up(dqmc_ctx* ctx, uint32_t l) { return ctx ? advance_up(ctx, (int)l) : DQMC_ERR_PARAM; }
int dqmc_advance_down(dqmc_ctx* ctx, uint32_t l) { return ctx ? advance_down(ctx, (int)l) : DQMC_ERR_PARAM; }

int dqmc_get_green_consistency(dqmc_ctx* ctx, double* out) {
    if (!ctx || !out) return DQMC_ERR_PARAM;
    CK(cudaMemcpyAsync(out, ctx->consistency, sizeof(double) * ctx->nmat, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_logdet(dqmc_ctx* ctx, int rep, int gc, double* out) {
    if (!valid_rep(ctx, rep) || gc < 0 || gc >= ctx->ngc || !out) return DQMC_ERR_PARAM;
    CK(cudaMemcpyAsync(out, ctx->logdet + rep * ctx->ngc + gc, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

// scratch of a from-scratch Green's function evaluation (one matrix): right / left chains and the result
struct ScratchG {
    cplx* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // rQ, rT, lQ, lT, G
    double* dv[3] = {nullptr, nullptr, nullptr};                    // rD, lD, logdet
    int alloc(dqmc_ctx* ctx) {
        for (int i = 0; i < 5; ++i) CK(dmalloc(&buf[i], DD(ctx)));
        for (int i = 0; i < 3; ++i) CK(dmalloc(&dv[i], (size_t)ctx->D));
        return DQMC_OK;
    }
    void release() {
        for (int i = 0; i < 5; ++i) cudaFree(buf[i]);
        for (int i = 0; i < 3; ++i) cudaFree(dv[i]);
    }
};

// G(k) = [1 + B(k, 0) B(beta, k)]^-1 of matrix `mat` from the fields, stabilised (right chain B(k, 0) and left chain
// B(beta, k), each in steps of <= s slices), into sc.buf[4]; the sweep state is left untouched
// (computeGreenFromScratch, detsdwopdim.cpp:4905-4933; setupUdVStorage_and_calculateGreen_forTimeslice, detmodel.h:605-674)
int green_for_timeslice_dev(dqmc_ctx* ctx, int mat, int k, ScratchG& sc) {
    const int s = ctx->s, m = ctx->m, D = ctx->D;
    const size_t dd = DD(ctx);
    cplx** buf = sc.buf;
    double** dv = sc.dv;
    bool haveR = false, haveL = false;
    int rc = DQMC_OK;
    for (int k1 = 0; k1 < k && rc == DQMC_OK;) {
        const int k2 = std::min(k, k1 + s);
        UdtView in{buf[0], (long long)dd, dv[0], D, buf[1], (long long)dd};
        rc = chain_step(ctx, DQMC_OP_LEFT, haveR ? &in : nullptr, k2, k1, buf[0], (long long)dd, dv[0], D, buf[1],
                        (long long)dd, mat, 1);
        haveR = true;
        k1 = k2;
    }
    for (int k2 = m; k2 > k && rc == DQMC_OK;) {
        const int k1 = std::max(k, k2 - s);
        UdtView in{buf[2], (long long)dd, dv[1], D, buf[3], (long long)dd};
        rc = chain_step(ctx, DQMC_OP_LEFT_ADJ, haveL ? &in : nullptr, k2, k1, buf[2], (long long)dd, dv[1], D, buf[3],
                        (long long)dd, mat, 1);
        haveL = true;
        k2 = k1;
    }
    if (rc == DQMC_OK) {
        UdtView rv = haveR ? UdtView{buf[0], (long long)dd, dv[0], D, buf[1], (long long)dd} : identity_view(ctx);
        UdtView lv = haveL ? UdtView{buf[2], (long long)dd, dv[1], D, buf[3], (long long)dd} : identity_view(ctx);
        rc = green_from_udts(ctx, rv, lv, buf[4], (long long)dd, dv[2], mat, 1);
    }
    return rc;
}

int dqmc_green_for_timeslice(dqmc_ctx* ctx, int rep, int gc, uint32_t kk, double* out) {
    if (!valid_rep(ctx, rep) || gc < 0 || gc >= ctx->ngc || !out || (int)kk > ctx->m) return DQMC_ERR_PARAM;
    const size_t dd = DD(ctx);
    ScratchG sc;
    RET(sc.alloc(ctx));
    int rc = green_for_timeslice_dev(ctx, rep * ctx->ngc + gc, (int)kk, sc);
    if (rc == DQMC_OK) {
        cudaError_t e;
        if (ctx->p.model == DQMC_MODEL_HUBBARD) {
            e = hub_real_part_launch(sc.buf[4], ctx->hubReal, dd, ctx->stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(out, ctx->hubReal, sizeof(double) * dd, cudaMemcpyDeviceToHost, ctx->stream);
        } else {
            e = cudaMemcpyAsync(out, sc.buf[4], sizeof(cplx) * dd, cudaMemcpyDeviceToHost, ctx->stream);
        }
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = DQMC_ERR_CUDA; }
    }
    cudaStreamSynchronize(ctx->stream);
    sc.release();
    return rc;
}

// sweepSimple / sweepSimpleThermalization (detmodel.h:718-758, detsdwopdim.cpp:4366-4420): for every slice the
// Green's function is recomputed from scratch, then the slice is updated.  The reference inverts the plain product
// 1 + B(k, 0) B(beta, k); here the stabilised evaluation serves it (same G where the plain product is accurate).
// O(m) from-scratch evaluations per sweep: a consistency tool for small systems, as in the reference.
int dqmc_sweep_simple(dqmc_ctx* ctx, int thermalization) {
    if (!ctx) return DQMC_ERR_PARAM;
    if (ctx->p.model != DQMC_MODEL_SDW) { ctx->err = "sweepSimple is served for DetSDW"; return DQMC_ERR_STATE; }
    RET(host_sync_rng(ctx));
    const size_t dd = DD(ctx);
    ScratchG sc;
    RET(sc.alloc(ctx));
    int rc = DQMC_OK;
    const size_t window = size_t(ctx->N) * (ctx->p.opdim + 1) * size_t(std::max(1, ctx->p.repeatUpdateInSlice));
    // the reference builds the B matrices of this sweep with the dense hopping exponential whatever the checkerboard
    // setting (sweepSimple_skeleton with sdwComputeBmat -> computeBmatSDW, detsdwopdim.cpp:4366-4420, 1307-1497)
    const bool denseBefore = ctx->denseNow;
    for (int k = 1; k <= ctx->m && rc == DQMC_OK; ++k) {
        ctx->denseNow = true;
        for (int mat = 0; mat < ctx->nmat && rc == DQMC_OK; ++mat) {
            rc = green_for_timeslice_dev(ctx, mat, k, sc);
            if (rc == DQMC_OK &&
                cudaMemcpyAsync(ctx->G + size_t(mat) * dd, sc.buf[4], sizeof(cplx) * dd, cudaMemcpyDeviceToDevice,
                                ctx->stream) != cudaSuccess) {
                ctx->err = "cudaMemcpyAsync failed";
                rc = DQMC_ERR_CUDA;
            }
        }
        ctx->denseNow = denseBefore;
        if (rc == DQMC_OK) rc = upload_rng_window(ctx, window);
        if (rc == DQMC_OK) rc = launch_update(ctx, k, thermalization);
        if (rc == DQMC_OK) rc = finish_rng_window(ctx);
    }
    ctx->denseNow = denseBefore;
    cudaStreamSynchronize(ctx->stream);
    sc.release();
    if (rc != DQMC_OK) return rc;
    ctx->currentTimeslice = ctx->m;
    ctx->performedSweeps += 1;
    return DQMC_OK;
}

int dqmc_green_from_udt_host(dqmc_ctx* ctx, const double* Qr, const double* dr, const double* Tr, const double* Ql,
                             const double* dl, const double* Tl, double* G_out, double* logdet_out) {
    if (!ctx || !Qr || !dr || !Tr || !Ql || !dl || !Tl || !G_out) return DQMC_ERR_PARAM;
    const size_t dd = DD(ctx);
    const int D = ctx->D;
    cplx* bufs[4];
    double* dv[2];
    for (int i = 0; i < 4; ++i) CK(dmalloc(&bufs[i], dd));
    for (int i = 0; i < 2; ++i) CK(dmalloc(&dv[i], (size_t)D));
    cplx* g = nullptr; double* ld = nullptr;
    CK(dmalloc(&g, dd));
    CK(dmalloc(&ld, 1));
    const double* src[4] = {Qr, Tr, Ql, Tl};
    for (int i = 0; i < 4; ++i) CK(cudaMemcpyAsync(bufs[i], src[i], sizeof(cplx) * dd, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dv[0], dr, sizeof(double) * D, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dv[1], dl, sizeof(double) * D, cudaMemcpyHostToDevice, ctx->stream));
    UdtView rv{bufs[0], (long long)dd, dv[0], D, bufs[1], (long long)dd};
    UdtView lv{bufs[2], (long long)dd, dv[1], D, bufs[3], (long long)dd};
    int rc = green_from_udts(ctx, rv, lv, g, (long long)dd, ld, 0, 1);
    if (rc == DQMC_OK) {
        CK(cudaMemcpyAsync(G_out, g, sizeof(cplx) * dd, cudaMemcpyDeviceToHost, ctx->stream));
        if (logdet_out) CK(cudaMemcpyAsync(logdet_out, ld, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 4; ++i) cudaFree(bufs[i]);
    cudaFree(dv[0]); cudaFree(dv[1]); cudaFree(g); cudaFree(ld);
    return rc;
}

int dqmc_udt_decompose_host(dqmc_ctx* ctx, const double* M, double* Q, double* d, double* T) {
    if (!ctx || !M || !Q || !d || !T) return DQMC_ERR_PARAM;
    const size_t dd = DD(ctx);
    const int D = ctx->D;
    cplx* work = ctx->W[0];
    cplx* q = ctx->W[1];
    cplx* t = ctx->W[2];
    CK(cudaMemcpyAsync(work, M, sizeof(cplx) * dd, cudaMemcpyHostToDevice, ctx->stream));
    RET(udt_decompose(ctx, work, (long long)dd, q, (long long)dd, ctx->vecA, D, t, (long long)dd, 0, 1));
    CK(cudaMemcpyAsync(Q, q, sizeof(cplx) * dd, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(T, t, sizeof(cplx) * dd, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(d, ctx->vecA, sizeof(double) * D, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_gemm_host(dqmc_ctx* ctx, int transa, int transb, int M, int N, int K, const double* A, const double* B,
                   double* C) {
    if (!ctx || !A || !B || !C || M < 1 || N < 1 || K < 1) return DQMC_ERR_PARAM;
    const int ar = transa ? K : M, ac = transa ? M : K, br = transb ? N : K, bc = transb ? K : N;
    cplx *dA, *dB, *dC;
    CK(dmalloc(&dA, size_t(ar) * ac));
    CK(dmalloc(&dB, size_t(br) * bc));
    CK(dmalloc(&dC, size_t(M) * N));
    CK(cudaMemcpyAsync(dA, A, sizeof(cplx) * ar * ac, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dB, B, sizeof(cplx) * br * bc, cudaMemcpyHostToDevice, ctx->stream));
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.transa = transa; g.transb = transb;
    g.A = dA; g.lda = ar; g.strideA = 0;
    g.B = dB; g.ldb = br; g.strideB = 0;
    g.C = dC; g.ldc = M; g.strideC = 0;
    g.rowscale = g.colscale = g.kscale = nullptr;
    g.strideRow = g.strideCol = g.strideK = 0;
    g.alpha = 1.0; g.kvec = nullptr; g.b_kmajor = 0;
    g.beta = 0.0;
    g.batch = 1;
    CKL(gemm_launch(g, ctx->stream));
    CK(cudaMemcpyAsync(C, dC, sizeof(cplx) * M * N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return DQMC_OK;
}

// ---- Monte Carlo -------------------------------------------------------------------------------
int dqmc_update_slice(dqmc_ctx* ctx, uint32_t k, int thermalization, uint32_t* n_accepted) {
    if (!ctx || k < 1 || (int)k > ctx->m) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    RET(upload_rng_window(ctx, size_t(ctx->N) * (ctx->p.opdim + 1) *
                                   size_t(ctx->p.model == DQMC_MODEL_SDW ? std::max(1, ctx->p.repeatUpdateInSlice) : 1)));
    RET(launch_update(ctx, (int)k, thermalization));
    if (n_accepted)
        CK(cudaMemcpyAsync(ctx->h_acc, ctx->accepted, sizeof(uint32_t) * ctx->R, cudaMemcpyDeviceToHost, ctx->stream));
    RET(finish_rng_window(ctx));
    if (n_accepted) std::memcpy(n_accepted, ctx->h_acc, sizeof(uint32_t) * ctx->R);
    return DQMC_OK;
}

int dqmc_global_shift_move(dqmc_ctx* ctx, int32_t* accepted) {
    if (!ctx) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    if (ctx->p.model != DQMC_MODEL_SDW) { ctx->err = "only defined for DetSDW"; return DQMC_ERR_STATE; }
    RET(global_move_kind(ctx, 0, accepted));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_wolff_cluster_move(dqmc_ctx* ctx, int with_shift, int32_t* accepted) {
    if (!ctx) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    if (ctx->p.model != DQMC_MODEL_SDW) { ctx->err = "only defined for DetSDW"; return DQMC_ERR_STATE; }
    RET(global_move_kind(ctx, with_shift ? 2 : 1, accepted));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

// finishMeasurements (detsdwopdim.cpp:903-1000, fermionic part) for one replica after dqmc_sweep(ctx, 2)
int dqmc_get_fermionic_observables(dqmc_ctx* ctx, int rep, double* scalars, double* vectors) {
    if (!valid_rep(ctx, rep) || !scalars || !vectors) return DQMC_ERR_PARAM;
    if (ctx->p.model != DQMC_MODEL_SDW) { ctx->err = "DetSDW observables; use dqmc_get_hubbard_observables"; return DQMC_ERR_STATE; }
    if (!ctx->fmAcc || ctx->fmSlices != ctx->m) { ctx->err = "no measured sweep (dqmc_sweep(ctx, 2)) yet"; return DQMC_ERR_STATE; }
    const int N = ctx->N, L = ctx->p.L, m = ctx->m, w = 2 * L - 1;
    const size_t nb = size_t(w) * w;
    std::vector<double> acc(ctx->fmAccLen);
    CK(cudaMemcpyAsync(acc.data(), ctx->fmAcc + size_t(rep) * ctx->fmAccLen, sizeof(double) * ctx->fmAccLen,
                       cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    scalars[0] = acc[0] / m;                             // greenK0
    scalars[1] = acc[1] / m;                             // greenLocal
    scalars[2] = acc[2] / m;                             // occDiffSq
    const double* binsX = acc.data() + 3;
    const double* binsY = binsX + 2 * nb;
    const double* pairPlus = binsY + 2 * nb;
    const double* pairMinus = pairPlus + N;
    // momentum-space occupation: k = -pi + (index + offset) 2 pi / L, offset 1/2 along antiperiodic directions (:623-671)
    const double offx = (ctx->p.bc == 1 || ctx->p.bc == 3) ? 0.5 : 0.0, offy = (ctx->p.bc == 2 || ctx->p.bc == 3) ? 0.5 : 0.0;
    for (int ks = 0; ks < N; ++ks) {
        const double ky = -M_PI + (double(ks / L) + offy) * 2.0 * M_PI / L;
        const double kx = -M_PI + (double(ks % L) + offx) * 2.0 * M_PI / L;
        double sx = 0, sy = 0;
        for (int dy = 0; dy < w; ++dy)
            for (int dx = 0; dx < w; ++dx) {
                const double arg = kx * (dx - (L - 1)) + ky * (dy - (L - 1));
                const double c = std::cos(arg), sn = std::sin(arg);
                const size_t bi = size_t(dy) * w + dx;
                sx += c * binsX[2 * bi] - sn * binsX[2 * bi + 1];
                sy += c * binsY[2 * bi] - sn * binsY[2 * bi + 1];
            }
        vectors[ks] = 2.0 - sx / (double(m) * N);        // kOccX
        vectors[N + ks] = 2.0 - sy / (double(m) * N);    // kOccY
    }
    for (int i = 0; i < N; ++i) {
        vectors[2 * N + i] = pairPlus[i] / m;
        vectors[3 * N + i] = pairMinus[i] / m;
    }
    // average over the nine sites around the maximum range (L/2, L/2), :973-987
    double pp = 0, pm = 0;
    for (int yy = L / 2 - 1; yy <= L / 2 + 1; ++yy)
        for (int xx = L / 2 - 1; xx <= L / 2 + 1; ++xx) {
            pp += vectors[2 * N + yy * L + xx];
            pm += vectors[3 * N + yy * L + xx];
        }
    scalars[3] = pp / 9.0;
    scalars[4] = pm / 9.0;
    return DQMC_OK;
}

// DetHubbard::finishMeasurements (dethubbard.cpp:601-612) for one replica after dqmc_sweep(ctx, 2).
// scalars: occupationUp, occupationDown, totalOccupation, doubleOccupation, localMoment, kineticEnergy, potentialEnergy,
// totalEnergy; zcorr: spinzCorrelationFunction [N]
int dqmc_get_hubbard_observables(dqmc_ctx* ctx, int rep, double* scalars, double* zcorr) {
    if (!valid_rep(ctx, rep) || !scalars || !zcorr) return DQMC_ERR_PARAM;
    if (ctx->p.model != DQMC_MODEL_HUBBARD) { ctx->err = "only defined for DetHubbard"; return DQMC_ERR_STATE; }
    if (!ctx->fmAcc || ctx->fmSlices != ctx->m) { ctx->err = "no measured sweep (dqmc_sweep(ctx, 2)) yet"; return DQMC_ERR_STATE; }
    std::vector<double> acc(ctx->fmAccLen);
    CK(cudaMemcpyAsync(acc.data(), ctx->fmAcc + size_t(rep) * ctx->fmAccLen, sizeof(double) * ctx->fmAccLen,
                       cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const double nm = double(ctx->N) * ctx->m;
    const double occUp = 1.0 - acc[0] / nm, occDn = 1.0 - acc[1] / nm;
    const double occTotal = occUp + occDn;
    const double occDouble = 1.0 + (acc[2] - acc[0] - acc[1]) / nm;
    const double ePot = ctx->p.U * occDouble;
    const double eKin = (ctx->p.t / nm) * (acc[3] + acc[4]) - ctx->p.mu * occTotal;
    scalars[0] = occUp; scalars[1] = occDn; scalars[2] = occTotal; scalars[3] = occDouble;
    scalars[4] = occTotal - 2 * occDouble;
    scalars[5] = eKin; scalars[6] = ePot; scalars[7] = eKin + ePot;
    for (int i = 0; i < ctx->N; ++i) zcorr[i] = acc[5 + i] / ctx->m;
    return DQMC_OK;
}

int dqmc_get_wolff_statistics(dqmc_ctx* ctx, int rep, double* out) {
    if (!valid_rep(ctx, rep) || !out) return DQMC_ERR_PARAM;
    RET(host_sync_rng(ctx));
    const dqmc_control_data& cd = ctx->ctrl_host[rep];    // attempted, accepted, attemptedShift, acceptedShift, added size
    out[0] = cd.attemptedWolffClusterUpdates; out[1] = cd.acceptedWolffClusterUpdates;
    out[2] = cd.attemptedWolffClusterShiftUpdates; out[3] = cd.acceptedWolffClusterShiftUpdates;
    out[4] = cd.addedWolffClusterSize;
    return DQMC_OK;
}

int dqmc_phi_action(dqmc_ctx* ctx, double* out) {
    if (!ctx || !out) return DQMC_ERR_PARAM;
    if (ctx->p.model != DQMC_MODEL_SDW) { ctx->err = "only defined for DetSDW"; return DQMC_ERR_STATE; }
    CKL(launch_phi_action(ctx->phi, ctx->rvals, ctx->actions, ctx->p.L, ctx->opdim, ctx->m, ctx->p.dtau, ctx->p.c,
                          ctx->p.u, (long long)phi_stride(ctx), ctx->R, ctx->stream));
    CK(cudaMemcpyAsync(out, ctx->actions, sizeof(double) * ctx->R, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DQMC_OK;
}

int dqmc_sweep(dqmc_ctx* ctx, int thermalization) {
    if (!ctx) return DQMC_ERR_PARAM;
    if (thermalization == 2) {                     // sweep(true): fermionic measurements after every slice
        RET(fm_prepare(ctx));
        CK(cudaMemsetAsync(ctx->fmAcc, 0, sizeof(double) * ctx->fmAccLen * ctx->R, ctx->stream));
        ctx->fmSlices = ctx->m;
    }
    const bool preloaded = ctx->rngResident && !ctx->rngAuto;          // explicit dqmc_rng_preload
    if (preloaded && size_t(ctx->rngWindow) < ctx->rngResidentUsedBound + ctx->rngCap + 8) {
        ctx->err = "resident random-number window too small for another sweep";
        return DQMC_ERR_STATE;
    }
    const bool global_now = ctx->lastSweepDir == +1 &&
                            (ctx->p.globalShift || ctx->p.wolffClusterUpdate || ctx->p.wolffClusterShiftUpdate) &&
                            (ctx->performedSweeps % ctx->p.globalUpdateInterval == 0);
    if (global_now) {
        // globalMove() before a down-sweep, detmodel.h:1422-1424 + detsdwopdim.cpp:3460-3485
        // a global shift reads OPDIM + 1 values per replica; a Wolff cluster an unbounded number (it re-heads the window)
        const bool inplace = preloaded && !ctx->p.wolffClusterUpdate && !ctx->p.wolffClusterShiftUpdate;
        if (preloaded && !inplace) {
            RET(sync_resident_cursor(ctx));
        } else if (inplace) {
            // The host draws come from the same streams, right after what the device has consumed.  The window
            // stays where it is: the host reads its values at the device cursors without consuming them, and the
            // cursors are advanced by what it read (no re-upload of the resident window).
            CK(cudaMemcpyAsync(ctx->h_cursor, ctx->cursor, sizeof(int) * ctx->R, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl, sizeof(dqmc_control_data) * ctx->R, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(ctx->h_err, ctx->errflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (*ctx->h_err) { ctx->err = "device error flag set (random-number window exhausted)"; return DQMC_ERR_STATE; }
            for (int r = 0; r < ctx->R; ++r) {
                ctx->ctrl_host[r] = ctx->h_ctrl[r];
                ctx->rng[r].begin_window((size_t)ctx->h_cursor[r]);
            }
        } else {
            RET(host_sync_rng(ctx));
        }
        // DetSDW::globalMove, detsdwopdim.cpp:3460-3485: shift, Wolff cluster, Wolff cluster + shift, in this order
        int grc = DQMC_OK;
        if (ctx->p.globalShift) grc = global_move_kind(ctx, 0, nullptr);
        if (grc == DQMC_OK && ctx->p.wolffClusterUpdate) grc = global_move_kind(ctx, 1, nullptr);
        if (grc == DQMC_OK && ctx->p.wolffClusterShiftUpdate) grc = global_move_kind(ctx, 2, nullptr);
        if (preloaded && !inplace && grc == DQMC_OK) grc = reupload_resident(ctx);
        if (inplace) {
            size_t most = 0;
            for (int r = 0; r < ctx->R; ++r) {
                const size_t used = ctx->rng[r].end_window();
                ctx->h_cursor[r] = (int)used;
                most = std::max(most, used);
            }
            if (grc == DQMC_OK) {
                CK(cudaMemcpyAsync(ctx->cursorAdd, ctx->h_cursor, sizeof(int) * ctx->R, cudaMemcpyHostToDevice, ctx->stream));
                CKL(launch_cursor_add(ctx->cursor, ctx->cursorAdd, ctx->R, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));          // h_cursor is reused
                ctx->rngResidentUsedBound += most;
            }
        }
        RET(grc);
    }
    if (!preloaded) RET(stream_begin_sweep(ctx));
    RET(run_sweep(ctx, ctx->lastSweepDir == +1 ? -1 : +1, thermalization));
    ctx->lastSweepDir = -ctx->lastSweepDir;
    if (!preloaded) RET(stream_end_sweep(ctx));
    else ctx->rngResidentUsedBound += ctx->rngCap;
    ctx->performedSweeps += 1;
    return DQMC_OK;
}

// Resident random numbers: upload the next `n_sweeps` sweeps' worth of every replica's stream once;
// sweeps then run without any host<->device traffic or synchronisation (cursors live on the
// device).  dqmc_rng_release() reads the cursors back and advances the host streams.
int dqmc_rng_preload(dqmc_ctx* ctx, int n_sweeps) {
    if (!ctx || n_sweeps < 1) return DQMC_ERR_PARAM;
    if (ctx->rngResident) RET(dqmc_rng_release(ctx));
    RET(host_sync_rng(ctx));
    const size_t per = ctx->rngCap * size_t(std::max(n_sweeps, 2)) + 16;
    if (per > ctx->rngAlloc) {
        CK(cudaStreamSynchronize(ctx->stream));
        // the captured sweeps hold the address of the old buffer in their kernel arguments: drop them
        for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
        ctx->graphs.clear();
        cudaFree(ctx->rngbuf);
        ctx->rngbuf = nullptr;
        CK(dmalloc(&ctx->rngbuf, per * ctx->R));
        ctx->rngAlloc = per;
    }
    if (per * ctx->R > ctx->hRngAlloc) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->copyStream) CK(cudaStreamSynchronize(ctx->copyStream));
        cudaFreeHost(ctx->h_rng);
        ctx->h_rng = nullptr;
        CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_rng), per * ctx->R * sizeof(double)));
        ctx->hRngAlloc = per * ctx->R;
    }
    ctx->rngStride = per;
    for (int r = 0; r < ctx->R; ++r)
        std::memcpy(ctx->h_rng + size_t(r) * per, ctx->rng[r].peek(per), per * sizeof(double));
    CK(cudaMemcpyAsync(ctx->rngbuf, ctx->h_rng, per * ctx->R * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->cursor, 0, sizeof(int) * ctx->R, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->rngWindow = (int)per;
    ctx->rngResident = true;
    ctx->rngResidentUsedBound = 0;
    return DQMC_OK;
}

int dqmc_rng_release(dqmc_ctx* ctx) {
    if (!ctx) return DQMC_ERR_PARAM;
    if (ctx->rngAuto) return host_sync_rng(ctx);
    if (!ctx->rngResident) return DQMC_OK;
    RET(finish_rng_window(ctx));
    ctx->rngResident = false;
    ctx->rngStride = ctx->rngCap;
    return DQMC_OK;
}

int dqmc_profile_enable(dqmc_ctx* ctx, int on) {
    if (!ctx) return DQMC_ERR_PARAM;
    prof_collect(ctx);
    ctx->profiling = on != 0;
    for (int i = 0; i < DQMC_PROF_NCAT; ++i) { ctx->profMs[i] = 0; ctx->profCount[i] = 0; }
    return DQMC_OK;
}

int dqmc_profile_get(dqmc_ctx* ctx, double* ms, uint64_t* counts) {
    if (!ctx || !ms || !counts) return DQMC_ERR_PARAM;
    CK(cudaStreamSynchronize(ctx->stream));
    prof_collect(ctx);
    for (int i = 0; i < DQMC_PROF_NCAT; ++i) { ms[i] = ctx->profMs[i]; counts[i] = ctx->profCount[i]; }
    return DQMC_OK;
}

const char* dqmc_profile_name(int category) {
    return (category >= 0 && category < DQMC_PROF_NCAT) ? kProfNames[category] : "";
}

int dqmc_accepted_total(dqmc_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return DQMC_ERR_PARAM;
    CK(cudaMemcpyAsync(ctx->h_scalars, ctx->acceptedTotal, sizeof(unsigned long long) * ctx->R, cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::memcpy(out, ctx->h_scalars, sizeof(uint64_t) * ctx->R);
    return DQMC_OK;
}

// ---- replica exchange --------------------------------------------------------------------------
int dqmc_exchange_actions(dqmc_ctx* ctx, double* actions_dev, double* actions_host) {
    if (!ctx) return DQMC_ERR_PARAM;
    if (ctx->p.model != DQMC_MODEL_SDW) { ctx->err = "only defined for DetSDW"; return DQMC_ERR_STATE; }
    double* dst = actions_dev ? actions_dev : ctx->actions;
    CKL(launch_exchange_action(ctx->phi, dst, ctx->N, ctx->opdim, ctx->m, ctx->p.dtau, (long long)phi_stride(ctx),
                               ctx->R, ctx->stream));
    if (actions_host) {
        CK(cudaMemcpyAsync(actions_host, dst, sizeof(double) * ctx->R, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return DQMC_OK;
}

int dqmc_exchange_pack(dqmc_ctx* ctx, double* payload_dev, int n_uniforms) {
    if (!ctx || !payload_dev || n_uniforms < 0) return DQMC_ERR_PARAM;
    if (ctx->p.model != DQMC_MODEL_SDW) { ctx->err = "only defined for DetSDW"; return DQMC_ERR_STATE; }
    const int R = ctx->R;
    if (ctx->rngAuto && ctx->rngResident) {
        // streamed random numbers: the look-ahead uniforms must already be on the device
        if (ctx->rngUploaded < ctx->rngResidentUsedBound + size_t(n_uniforms)) RET(host_sync_rng(ctx));
        else CK(cudaStreamWaitEvent(ctx->stream, ctx->copyEvent[ctx->copyHalf ^ 1], 0));
    }
    CKL(launch_exchange_action(ctx->phi, payload_dev, ctx->N, ctx->opdim, ctx->m, ctx->p.dtau,
                               (long long)phi_stride(ctx), R, ctx->stream));
    double* uni = payload_dev + R;
    double* blobs = uni + n_uniforms;
    if (!ctx->rngResident && n_uniforms > 8 * R + 16) {
        ctx->err = "n_uniforms exceeds the staging buffer (8 R + 16 doubles)";
        return DQMC_ERR_PARAM;
    }
    if (ctx->rngResident && !ctx->rngAuto && n_uniforms > 0) {
        // explicitly preloaded window: the look-ahead must lie inside it (the cursor of replica 0 is on the device;
        // the bound the host tracks is conservative)
        if (ctx->rngResidentUsedBound + size_t(n_uniforms) > size_t(ctx->rngWindow)) {
            ctx->err = "resident random-number window too small for the exchange look-ahead";
            return DQMC_ERR_STATE;
        }
    }
    if (!ctx->rngResident && n_uniforms > 0) {
        const double* src = ctx->rng[0].peek((size_t)n_uniforms);
        std::memcpy(ctx->h_scalars, src, sizeof(double) * n_uniforms);      // h_scalars holds >= 8R+16 doubles
        CK(cudaMemcpyAsync(uni, ctx->h_scalars, sizeof(double) * n_uniforms, cudaMemcpyHostToDevice, ctx->stream));
    }
    CKL(launch_exchange_pack(ctx->rngbuf, ctx->cursor, ctx->rngWindow, uni, n_uniforms, ctx->ctrl, blobs, R,
                             ctx->rngResident ? 1 : 0, ctx->stream));
    return DQMC_OK;
}

int dqmc_exchange_apply(dqmc_ctx* ctx, const double* r_new, const dqmc_control_data* ctrl_new, int n_uniforms_used) {
    if (!ctx || !r_new || !ctrl_new || n_uniforms_used < 0) return DQMC_ERR_PARAM;
    for (int r = 0; r < ctx->R; ++r) {
        ctx->h_r[r] = r_new[r];
        ctx->ctrl_host[r] = ctrl_new[r];
    }
    if (n_uniforms_used > 0) {
        if (ctx->rngResident) {
            CKL(launch_cursor_advance(ctx->cursor, 0, n_uniforms_used, ctx->stream));
            ctx->rngResidentUsedBound += size_t(n_uniforms_used);
        }
        else ctx->rng[0].skip((size_t)n_uniforms_used);
    }
    RET(upload_ctrl(ctx));
    RET(upload_rvals(ctx));
    return DQMC_OK;
}

// ---- NCCL communicator injected by the host: the one collective of the path ---------------------
// The library does not link NCCL: it takes ncclAllGather from the NCCL the host process already uses (so that the
// communicator and the entry point come from the same library), or loads libnccl.so.2 if none is loaded.
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
static nccl_allgather_fn resolve_nccl_allgather() {
    static nccl_allgather_fn fn = nullptr;
    if (fn) return fn;
    void* sym = dlsym(RTLD_DEFAULT, "ncclAllGather");
    if (!sym) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (h) sym = dlsym(h, "ncclAllGather");
    }
    fn = reinterpret_cast<nccl_allgather_fn>(sym);
    return fn;
}

int dqmc_set_comm(dqmc_ctx* ctx, void* nccl_comm, int nranks, int rank) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !nccl_comm)) return DQMC_ERR_PARAM;
    if (nccl_comm && !resolve_nccl_allgather()) { ctx->err = "ncclAllGather not found (no NCCL in this process)"; return DQMC_ERR_STATE; }
    ctx->comm = nccl_comm;
    ctx->commRanks = nranks;
    ctx->commRank = rank;
    return DQMC_OK;
}

int dqmc_exchange_payload_len(dqmc_ctx* ctx, int n_uniforms) {
    if (!ctx || n_uniforms < 0) return -1;
    return ctx->R + n_uniforms + ctx->R * int(sizeof(dqmc_control_data) / sizeof(double));
}

// gather side of replicaExchangeStep (detqmcpt.h:968-1012) in one call: pack the local payload into its slot of
// gathered_dev and all-gather in place on the context's stream; with gathered_host the result is copied back and the
// call returns after the stream has drained
int dqmc_exchange_allgather(dqmc_ctx* ctx, int n_uniforms, double* gathered_dev, double* gathered_host) {
    if (!ctx || !gathered_dev || n_uniforms < 0) return DQMC_ERR_PARAM;
    const int ranks = ctx->commRanks > 0 ? ctx->commRanks : 1, rank = ctx->commRanks > 0 ? ctx->commRank : 0;
    const size_t len = size_t(dqmc_exchange_payload_len(ctx, n_uniforms));
    RET(dqmc_exchange_pack(ctx, gathered_dev + size_t(rank) * len, n_uniforms));
    if (ranks > 1 || ctx->comm) {
        nccl_allgather_fn ag = resolve_nccl_allgather();
        if (!ag) { ctx->err = "ncclAllGather not found"; return DQMC_ERR_STATE; }
        const int rc = ag(gathered_dev + size_t(rank) * len, gathered_dev, len, /* ncclDouble */ 8, ctx->comm, ctx->stream);
        if (rc != 0) { ctx->err = "ncclAllGather failed"; return DQMC_ERR_CUDA; }
        ctx->launches += 1;
    }
    if (gathered_host) {
        CK(cudaMemcpyAsync(gathered_host, gathered_dev, sizeof(double) * len * ranks, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return DQMC_OK;
}

double dqmc_exchange_probability(double par1, double action1, double par2, double action2) {
    const double delta = (par1 - par2) * (action2 - action1);
    return delta <= 0.0 ? 1.0 : std::exp(-delta);
}

int dqmc_exchange_walk(int n, const double* control_values, int32_t* par_process, int32_t* process_par,
                       const double* actions, const double* uniforms, int32_t* n_used, int32_t* swapped) {
    if (n < 1 || !control_values || !par_process || !process_par || !actions || !n_used) return DQMC_ERR_PARAM;
    int used = 0;
    for (int cpi1 = 0; cpi1 < n - 1; ++cpi1) {
        const int cpi2 = cpi1 + 1;
        const int p1 = par_process[cpi1], p2 = par_process[cpi2];
        const double prob = dqmc_exchange_probability(control_values[cpi1], actions[p1], control_values[cpi2], actions[p2]);
        bool acc = prob >= 1.0;
        if (!acc) {
            if (!uniforms) return DQMC_ERR_PARAM;
            const double u = uniforms[used++];
            if (!(u > 0.0 && u < 1.0)) return DQMC_ERR_STATE;      // beyond the look-ahead the owner packed (sentinel -1)
            acc = u <= prob;
        }
        if (acc) {
            process_par[p1] = cpi2;
            process_par[p2] = cpi1;
            par_process[cpi1] = p2;
            par_process[cpi2] = p1;
        }
        if (swapped) swapped[cpi1] = acc ? 1 : 0;
    }
    *n_used = used;
    return DQMC_OK;
}

}  // extern "C"
