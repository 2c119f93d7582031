// Batched Householder QR with column pivoting, explicit Q, UDT extraction and triangular solve.
//
// Replaces udvDecompose (udv.h:68-90, LAPACK zgesvd) as the numerical stabiliser of the B-matrix
// chains: M P = Q R,  M = Q diag(d) T with d = |diag R| and T = diag(d)^-1 R P^T.  Parity with the
// reference is defined on derived quantities (G, log|det|, acceptance ratios), not on the factors
// (SURVEY.md section 0, fact 2).
//
// One CTA factors one matrix of the batch; the 64 replicas of the headline configuration keep 64
// SMs busy.  Per column: block-wide arg-max of the remaining column norms (deterministic ties),
// column swap, LAPACK-style reflector (zlarfg conventions), then one warp per trailing column does
// dot-product + rank-1 update with warp-shuffle reductions and recomputes that column's remaining
// norm exactly (no down-dating, so no cancellation).  Matrices are L2 resident (D x D x 16 B).
#include "dqmc_internal.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

namespace dqmc {
namespace {

constexpr int kQrThreads = 1024;

__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {      // conj(a) * b
    return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// trailing update shared by the factorisation and the Q accumulation:
// for each column c in [c0, D): w = v^H a_c(rows r0..D-1), a_c -= f * w * v; returns new norm^2 of
// rows r0+1..D-1 in colnorm (if not null).  v lives in shared memory, v[0] is row r0.
__device__ __forceinline__ void apply_reflector(cplx* A, int D, const cplx* v, cplx f, int r0, int c0,
                                                double* colnorm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int len = D - r0;
    for (int c = c0 + warp; c < D; c += nwarps) {
        cplx* col = A + size_t(c) * D + r0;
        double wr = 0, wi = 0;
        for (int i = lane; i < len; i += 32) {
            const cplx w = cmulc(v[i], col[i]);
            wr += w.x;
            wi += w.y;
        }
        wr = warp_sum(wr);
        wi = warp_sum(wi);
        const cplx fw = cmul(f, make_double2(wr, wi));
        double nrm = 0;
        for (int i = lane; i < len; i += 32) {
            cplx a = col[i];
            const cplx d = cmul(fw, v[i]);
            a.x -= d.x;
            a.y -= d.y;
            col[i] = a;
            if (i > 0) nrm += a.x * a.x + a.y * a.y;
        }
        if (colnorm) {
            nrm = warp_sum(nrm);
            if (lane == 0) colnorm[c] = nrm;
        }
    }
}

__global__ void __launch_bounds__(kQrThreads) qrcp_factor_kernel(cplx* Aall, int D, long long strideA,
                                                                 cplx* tauAll, int* permAll, double* normAll) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* v = reinterpret_cast<cplx*>(smem_raw);                 // [D]
    __shared__ double red_val[32];
    __shared__ int red_idx[32];
    __shared__ double s_red[32];
    __shared__ int s_pvt;
    __shared__ cplx s_tau, s_scale, s_alpha;
    __shared__ double s_beta;

    const int b = blockIdx.x;
    cplx* A = Aall + size_t(b) * strideA;
    cplx* tau = tauAll + size_t(b) * D;
    int* perm = permAll + size_t(b) * D;
    double* colnorm = normAll + size_t(b) * D;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    for (int c = warp; c < D; c += nwarps) {
        const cplx* col = A + size_t(c) * D;
        double s = 0;
        for (int i = lane; i < D; i += 32) s += col[i].x * col[i].x + col[i].y * col[i].y;
        s = warp_sum(s);
        if (lane == 0) { colnorm[c] = s; perm[c] = c; }
    }
    __syncthreads();

    for (int j = 0; j < D; ++j) {
        // ---- 1. pivot = argmax_{c >= j} colnorm[c], smallest index on ties
        double best = -1.0;
        int bidx = D;
        for (int c = j + tid; c < D; c += blockDim.x) {
            const double val = colnorm[c];
            if (val > best) { best = val; bidx = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
        }
        if (lane == 0) { red_val[warp] = best; red_idx[warp] = bidx; }
        __syncthreads();
        if (warp == 0) {
            best = lane < nwarps ? red_val[lane] : -1.0;
            bidx = lane < nwarps ? red_idx[lane] : D;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
            }
            if (lane == 0) s_pvt = bidx;
        }
        __syncthreads();
        const int pvt = s_pvt;
        // ---- 2. swap columns j <-> pvt
        if (pvt != j) {
            for (int i = tid; i < D; i += blockDim.x) {
                const cplx t = A[size_t(j) * D + i];
                A[size_t(j) * D + i] = A[size_t(pvt) * D + i];
                A[size_t(pvt) * D + i] = t;
            }
            if (tid == 0) {
                const int tp = perm[j]; perm[j] = perm[pvt]; perm[pvt] = tp;
                colnorm[pvt] = colnorm[j];         // colnorm[j] is not needed any more
            }
        }
        __syncthreads();
        // ---- 3. reflector for x = A[j:, j]
        cplx* colj = A + size_t(j) * D;
        double xn = 0;
        for (int i = j + 1 + tid; i < D; i += blockDim.x) xn += colj[i].x * colj[i].x + colj[i].y * colj[i].y;
        xn = warp_sum(xn);
        if (lane == 0) s_red[warp] = xn;
        __syncthreads();
        if (warp == 0) {
            xn = lane < nwarps ? s_red[lane] : 0.0;
            xn = warp_sum(xn);
            if (lane == 0) {
                const cplx alpha = colj[j];
                cplx t, sc;
                double beta;
                if (xn == 0.0 && alpha.y == 0.0) {
                    t = make_double2(0, 0);
                    sc = make_double2(0, 0);
                    beta = alpha.x;
                } else {
                    const double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn);
                    beta = alpha.x >= 0 ? -nrm : nrm;
                    t = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
                    // 1 / (alpha - beta)
                    const double dr = alpha.x - beta, di = alpha.y;
                    const double den = dr * dr + di * di;
                    sc = make_double2(dr / den, -di / den);
                }
                s_tau = t;
                s_scale = sc;
                s_beta = beta;
                s_alpha = alpha;
                tau[j] = t;
            }
        }
        __syncthreads();
        {
            const cplx sc = s_scale;
            for (int i = j + tid; i < D; i += blockDim.x) {
                cplx val;
                if (i == j) {
                    val = make_double2(1, 0);
                    colj[i] = make_double2(s_beta, 0);
                } else {
                    val = cmul(colj[i], sc);
                    colj[i] = val;
                }
                v[i - j] = val;
            }
        }
        __syncthreads();
        // ---- 4. A[j:, j+1:] = (I - conj(tau) v v^H) A[j:, j+1:]
        const cplx ct = make_double2(s_tau.x, -s_tau.y);
        if (s_tau.x != 0.0 || s_tau.y != 0.0) {
            apply_reflector(A, D, v, ct, j, j + 1, colnorm);
        } else {
            // H = I: only the remaining norms change (row j leaves the active part)
            for (int c = j + 1 + tid; c < D; c += blockDim.x) {
                const cplx a = A[size_t(c) * D + j];
                colnorm[c] = fmax(colnorm[c] - (a.x * a.x + a.y * a.y), 0.0);
            }
        }
        __syncthreads();
    }
}

// Q = H_0 H_1 ... H_{D-1} by backward accumulation (LAPACK zung2r).
__global__ void __launch_bounds__(kQrThreads) qr_form_q_kernel(const cplx* Aall, const cplx* tauAll, cplx* Qall,
                                                               int D, long long strideA) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* v = reinterpret_cast<cplx*>(smem_raw);
    const int b = blockIdx.x;
    const cplx* A = Aall + size_t(b) * strideA;
    const cplx* tau = tauAll + size_t(b) * D;
    cplx* Q = Qall + size_t(b) * strideA;
    const int tid = threadIdx.x;
    for (size_t idx = tid; idx < size_t(D) * D; idx += blockDim.x) {
        const int i = idx % D, c = idx / D;
        Q[idx] = make_double2(i == c ? 1.0 : 0.0, 0.0);
    }
    __syncthreads();
    for (int j = D - 1; j >= 0; --j) {
        const cplx* colj = A + size_t(j) * D;
        for (int i = j + tid; i < D; i += blockDim.x) v[i - j] = (i == j) ? make_double2(1, 0) : colj[i];
        __syncthreads();
        const cplx t = tau[j];
        if (t.x != 0.0 || t.y != 0.0) apply_reflector(Q, D, v, t, j, j, nullptr);
        __syncthreads();
    }
}

__global__ void qr_extract_dt_kernel(const cplx* Aall, const int* permAll, double* dAll, cplx* Tall, int D,
                                     long long strideA) {
    pdl_enter();
    const int b = blockIdx.y;
    const cplx* A = Aall + size_t(b) * strideA;
    const int* perm = permAll + size_t(b) * D;
    cplx* T = Tall + size_t(b) * strideA;
    double* d = dAll + size_t(b) * D;
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(D) * D) return;
    const int i = idx % D, j = idx / D;
    const double di = fabs(A[size_t(i) * D + i].x);
    if (j == i) d[i] = di;
    cplx val = make_double2(0, 0);
    if (j >= i) {
        const cplx r = A[size_t(j) * D + i];
        val = make_double2(r.x / di, r.y / di);
    }
    T[size_t(perm[j]) * D + i] = val;
}

// R Z = Y by column-oriented back substitution; a CTA owns kTrsmCols right-hand sides in smem.
constexpr int kTrsmCols = 16;
constexpr int kTrsmThreads = 256;
__global__ void __launch_bounds__(kTrsmThreads) trsm_upper_kernel(const cplx* Aall, cplx* Yall, cplx* Zall,
                                                                  const int* permAll, int D, long long strideA) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* ys = reinterpret_cast<cplx*>(smem_raw);                // [D][kTrsmCols]
    __shared__ cplx zl[kTrsmCols];
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * kTrsmCols;
    const int nc = min(kTrsmCols, D - c0);
    const cplx* A = Aall + size_t(b) * strideA;
    cplx* Y = Yall + size_t(b) * strideA;
    cplx* Z = Zall + size_t(b) * strideA;
    const int* perm = permAll ? permAll + size_t(b) * D : nullptr;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < nc * D; idx += blockDim.x) {
        const int c = idx / D, i = idx - c * D;
        ys[i * kTrsmCols + c] = Y[size_t(c0 + c) * D + i];
    }
    __syncthreads();
    for (int l = D - 1; l >= 0; --l) {
        if (tid < nc) {
            const cplx r = A[size_t(l) * D + l];                 // diagonal of R (real for QR, general here)
            const cplx y = ys[l * kTrsmCols + tid];
            const double den = r.x * r.x + r.y * r.y;
            const cplx z = make_double2((y.x * r.x + y.y * r.y) / den, (y.y * r.x - y.x * r.y) / den);
            zl[tid] = z;
            ys[l * kTrsmCols + tid] = z;
        }
        __syncthreads();
        const cplx* rcol = A + size_t(l) * D;
        for (int idx = tid; idx < l * kTrsmCols; idx += blockDim.x) {
            const int i = idx / kTrsmCols, c = idx - i * kTrsmCols;
            if (c < nc) {
                const cplx r = __ldg(rcol + i);
                const cplx z = zl[c];
                cplx y = ys[idx];
                y.x -= r.x * z.x - r.y * z.y;
                y.y -= r.x * z.y + r.y * z.x;
                ys[idx] = y;
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nc * D; idx += blockDim.x) {
        const int c = idx / D, i = idx - c * D;
        const int orow = perm ? perm[i] : i;
        Z[size_t(c0 + c) * D + orow] = ys[i * kTrsmCols + c];
    }
}

__global__ void scale_split_kernel(const double* dAll, double* invBig, double* small_, double* logacc, int D) {
    pdl_enter();
    const int b = blockIdx.x;
    const double* d = dAll + size_t(b) * D;
    __shared__ double red[32];
    double ls = 0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const double di = d[i];
        const double big = di > 1.0 ? di : 1.0;
        invBig[size_t(b) * D + i] = 1.0 / big;
        small_[size_t(b) * D + i] = di < 1.0 ? di : 1.0;
        ls += log(big);
    }
    ls = warp_sum(ls);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ls;
    __syncthreads();
    if (threadIdx.x < 32) {
        ls = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        ls = warp_sum(ls);
        if (threadIdx.x == 0 && logacc) logacc[b] += ls;
    }
}

__global__ void logdiag_kernel(const cplx* Aall, double* logacc, int D, long long strideA) {
    pdl_enter();
    const int b = blockIdx.x;
    const cplx* A = Aall + size_t(b) * strideA;
    __shared__ double red[32];
    double ls = 0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const cplx r = A[size_t(i) * D + i];
        ls += 0.5 * log(r.x * r.x + r.y * r.y);
    }
    ls = warp_sum(ls);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ls;
    __syncthreads();
    if (threadIdx.x < 32) {
        ls = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        ls = warp_sum(ls);
        if (threadIdx.x == 0) logacc[b] += ls;
    }
}


// =================================================================================================
// Blocked Householder QR (compact WY) with column PRE-pivoting
// =================================================================================================
// Columns are ordered once by decreasing norm (the matrices of the stabilised chains, (B Q) diag(d),
// are column graded, which is what pivoting has to catch; Bai, Lee, Li, Xu, "Stable solutions of
// linear systems involving long chain of matrix multiplications", LAA 2011, sec. 3.2 "pre-pivoting"),
// then a blocked QR without further column exchanges runs as
//     panel factorisation (one CTA per matrix, panel resident in shared memory)
//   + trailing update  A2 -= V (V T)^H A2  as two batched DMMA GEMMs over all SMs.
// The panel kernel leaves R in A, and writes the explicit unit-lower-trapezoidal V and the product
// V*T of every panel to the workspace, so that forming Q, applying Q^H and the trailing updates are
// pure GEMMs.  The fully pivoted one-CTA kernels above stay as the cross-check of the tests.

constexpr int kPanelThreads = 1024;
constexpr int kPanelMaxNb = 32;

__device__ __forceinline__ cplx cfma_(cplx a, cplx b, cplx c) {      // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ cplx cfmac_(cplx a, cplx b, cplx c) {     // conj(a)*b + c
    return make_double2(fma(a.x, b.x, fma(a.y, b.y, c.x)), fma(a.x, b.y, fma(-a.y, b.x, c.y)));
}

// squared column norms, then rank by counting: perm[rank] = column (descending, ties by index).  A cluster of CS CTAs
// per matrix (one SM streams the matrix at ~50 GB/s only): every CTA takes the norms of its share of the columns (same
// summation order per column whichever CTA computes it), CTA 0 collects them through distributed shared memory and ranks.
__global__ void __launch_bounds__(1024) colnorm_rank_kernel(const cplx* Aall, int D, long long strideA, int* permAll,
                                                            double* normAll, int CS) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* nrm = reinterpret_cast<double*>(smem_raw);            // [D]
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const int b = blockIdx.x / CS, crank = blockIdx.x % CS;
    const cplx* A = Aall + size_t(b) * strideA;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int c = crank * nwarps + warp; c < D; c += CS * nwarps) {
        const cplx* col = A + size_t(c) * D;
        double s = 0;
        for (int i = lane; i < D; i += 32) { const cplx v = col[i]; s = fma(v.x, v.x, fma(v.y, v.y, s)); }
        s = warp_sum(s);
        if (lane == 0) nrm[c] = s;
    }
    __syncthreads();
    if (CS > 1) {
        cluster.sync();
        if (crank == 0) {
            for (int c = tid; c < D; c += blockDim.x) {
                const int owner = (c / nwarps) % CS;
                if (owner != 0) nrm[c] = cluster.map_shared_rank(nrm, owner)[c];
            }
        }
        __syncthreads();
        cluster.sync();                                           // peers stay resident until their norms have been read
        if (crank != 0) return;
    }
    for (int j = tid; j < D; j += blockDim.x) {
        const double nj = nrm[j];
        int rank = 0;
        for (int i = 0; i < D; ++i) {
            const double ni = nrm[i];
            rank += (ni > nj || (ni == nj && i < j)) ? 1 : 0;
        }
        permAll[size_t(b) * D + rank] = j;
        if (normAll) normAll[size_t(b) * D + j] = nj;
    }
}

// out[:, j] = in[:, perm[j]]
__global__ void permute_columns_kernel(const cplx* inAll, cplx* outAll, const int* permAll, int D, long long strideIn,
                                       long long strideOut) {
    pdl_enter();
    const int b = blockIdx.y, j = blockIdx.x;
    const cplx* src = inAll + size_t(b) * strideIn + size_t(permAll[size_t(b) * D + j]) * D;
    cplx* dst = outAll + size_t(b) * strideOut + size_t(j) * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) dst[i] = src[i];
}

// out[perm[j], :] = in[j, :]
__global__ void permute_rows_kernel(const cplx* inAll, cplx* outAll, const int* permAll, int D, long long strideIn,
                                    long long strideOut) {
    pdl_enter();
    const int b = blockIdx.y, c = blockIdx.x;
    const cplx* src = inAll + size_t(b) * strideIn + size_t(c) * D;
    cplx* dst = outAll + size_t(b) * strideOut + size_t(c) * D;
    const int* perm = permAll + size_t(b) * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) dst[perm[i]] = src[i];
}

// Householder QR of the panel A[j0:D, j0:j0+nbc] of every matrix of the batch.
__global__ void __launch_bounds__(kPanelThreads) qr_panel_kernel(cplx* Aall, long long strideA, int D, int j0, int nbc,
                                                                 cplx* Vall, cplx* VTall, long long strideV) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = D - j0;
    const int ldp = m | 1;                                          // odd leading dimension
    cplx* P = reinterpret_cast<cplx*>(smem_raw);                    // [nbc][ldp]
    cplx* Gm = P + size_t(nbc) * ldp;                               // [32][33] strict upper Gram of V
    cplx* Tm = Gm + 32 * 33;                                        // [32][33] T factor
    __shared__ cplx s_tau[kPanelMaxNb];
    __shared__ double s_part[32];
    __shared__ double s_norm;                                       // |x|^2 below the diagonal of the next column

    const int b = blockIdx.x;
    cplx* A = Aall + size_t(b) * strideA;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    for (int idx = tid; idx < nbc * m; idx += blockDim.x) {
        const int c = idx / m, i = idx - c * m;
        P[c * ldp + i] = A[size_t(j0 + c) * D + j0 + i];
    }
    __syncthreads();
    // norm^2 of rows 1.. of column 0
    {
        double xn = 0;
        for (int i = 1 + tid; i < m; i += blockDim.x) { const cplx v = P[i]; xn = fma(v.x, v.x, fma(v.y, v.y, xn)); }
        xn = warp_sum(xn);
        if (lane == 0) s_part[warp] = xn;
        __syncthreads();
        if (warp == 0) {
            xn = lane < nwarps ? s_part[lane] : 0.0;
            xn = warp_sum(xn);
            if (lane == 0) s_norm = xn;
        }
        __syncthreads();
    }

    for (int c = 0; c < nbc; ++c) {
        cplx* colc = P + c * ldp;
        // ---- reflector (zlarfg conventions); every thread computes the same scalars
        const double xn = (c + 1 < m) ? s_norm : 0.0;
        const cplx alpha = colc[c];
        cplx tau, sc;
        double beta;
        if (xn == 0.0 && alpha.y == 0.0) {
            tau = make_double2(0, 0);
            sc = make_double2(0, 0);
            beta = alpha.x;
        } else {
            const double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn);
            beta = alpha.x >= 0 ? -nrm : nrm;
            tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
            const double dr = alpha.x - beta, di = alpha.y;
            const double den = dr * dr + di * di;
            sc = make_double2(dr / den, -di / den);
        }
        __syncthreads();                                            // everyone has read alpha / s_norm
        for (int i = c + 1 + tid; i < m; i += blockDim.x) colc[i] = cmul(colc[i], sc);
        if (tid == 0) { colc[c] = make_double2(beta, 0); s_tau[c] = tau; }
        __syncthreads();
        // ---- apply H_c^H = I - conj(tau) v v^H to the remaining panel columns, one warp per column;
        // the warp that owns column c+1 also produces the norm for the next reflector
        if (tau.x != 0.0 || tau.y != 0.0) {
            const cplx ct = make_double2(tau.x, -tau.y);
            for (int k = c + 1 + warp; k < nbc; k += nwarps) {
                cplx* colk = P + k * ldp;
                double wr = 0, wi = 0;
                for (int i = c + lane; i < m; i += 32) {
                    const cplx v = (i == c) ? make_double2(1, 0) : colc[i];
                    const cplx w = cmulc(v, colk[i]);
                    wr += w.x;
                    wi += w.y;
                }
                wr = warp_sum(wr);
                wi = warp_sum(wi);
                const cplx fw = cmul(ct, make_double2(wr, wi));
                double nn = 0;
                for (int i = c + lane; i < m; i += 32) {
                    const cplx v = (i == c) ? make_double2(1, 0) : colc[i];
                    cplx a = colk[i];
                    const cplx d = cmul(fw, v);
                    a.x -= d.x;
                    a.y -= d.y;
                    colk[i] = a;
                    if (i > c + 1) nn = fma(a.x, a.x, fma(a.y, a.y, nn));
                }
                if (k == c + 1) {
                    nn = warp_sum(nn);
                    if (lane == 0) s_norm = nn;
                }
            }
        } else if (c + 1 < nbc && warp == 0) {
            const cplx* colk = P + (c + 1) * ldp;
            double nn = 0;
            for (int i = c + 2 + lane; i < m; i += 32) { const cplx a = colk[i]; nn = fma(a.x, a.x, fma(a.y, a.y, nn)); }
            nn = warp_sum(nn);
            if (lane == 0) s_norm = nn;
        }
        __syncthreads();
    }

    // ---- Gram matrix of the reflectors, strict upper part: G[i][j] = v_i^H v_j  (i < j)
    for (int i = warp; i < nbc; i += nwarps) {
        const int j = lane;
        if (j > i && j < nbc) {
            const cplx* vi = P + i * ldp;
            const cplx* vj = P + j * ldp;
            cplx s = make_double2(vi[j].x, -vi[j].y);               // row j: conj(v_i[j]) * 1
            for (int r = j + 1; r < m; ++r) s = cfmac_(vi[r], vj[r], s);
            Gm[i * 33 + j] = s;
        }
    }
    __syncthreads();
    // ---- T = (diag(1/tau) + strict_upper(G))^-1, column by column (zlarft, forward / columnwise)
    if (tid < nbc) {
        const int c = tid;
        Tm[c * 33 + c] = s_tau[c];
        for (int i = c - 1; i >= 0; --i) {
            cplx s = make_double2(0, 0);
            for (int l = i + 1; l <= c; ++l) s = cfma_(Gm[i * 33 + l], Tm[l * 33 + c], s);
            const cplx ti = s_tau[i];
            Tm[i * 33 + c] = make_double2(-(ti.x * s.x - ti.y * s.y), -(ti.x * s.y + ti.y * s.x));
        }
    }
    __syncthreads();
    // ---- write back: R (and v below it) to A, explicit V and V*T to the workspace
    cplx* V = Vall + size_t(b) * strideV;
    cplx* VT = VTall + size_t(b) * strideV;
    for (int idx = tid; idx < nbc * m; idx += blockDim.x) {
        const int c = idx / m, r = idx - c * m;
        const cplx pv = P[c * ldp + r];
        A[size_t(j0 + c) * D + j0 + r] = pv;
        V[size_t(j0 + c) * D + j0 + r] = r > c ? pv : make_double2(r == c ? 1.0 : 0.0, 0.0);
        // (V T)[r, c] = sum_{l <= min(c, r)} V[r, l] T[l, c]
        cplx s = make_double2(0, 0);
        const int lmax = min(c, r);
        for (int l = 0; l <= lmax; ++l) {
            const cplx vl = (l == r) ? make_double2(1, 0) : P[l * ldp + r];
            s = cfma_(vl, Tm[l * 33 + c], s);
        }
        VT[size_t(j0 + c) * D + j0 + r] = s;
    }
}

// Register-resident variant of the panel factorisation for m <= 32 * MAXT rows: warp w owns panel column
// w in registers (lane l holds rows l, l+32, ...), the current reflector is broadcast through a
// double-buffered shared-memory vector, so a column step is one barrier, one dot product and one update
// per warp.  The warps whose columns are already factored (w < c) use the same dot product to build the
// Gram matrix of the reflectors, and warp c-1 folds its column into the compact-WY factor T while the
// others work, so T costs no extra pass.  ~3x fewer instructions per column than qr_panel_kernel, which
// stays as the fallback for taller panels.
constexpr int kPanelRegThreads = 512;      // 16 warps, two panel columns per warp

template <int MAXT>
__global__ void __launch_bounds__(kPanelRegThreads) qr_panel_reg_kernel(cplx* Aall, long long strideA, int D, int j0,
                                                                        int nbc, cplx* Vall, cplx* VTall, long long strideV) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = D - j0;
    const int ldp = m | 1;
    cplx* P = reinterpret_cast<cplx*>(smem_raw);                    // [nbc][ldp] panel (written after the factorisation)
    cplx* Gm = P + size_t(nbc) * ldp;                               // [32][33] strict upper Gram of V
    cplx* Tm = Gm + 32 * 33;                                        // [32][33] T factor
    cplx* vbuf = Tm + 32 * 33;                                      // [2][m] current reflector
    __shared__ cplx s_tau[kPanelMaxNb];

    const int b = blockIdx.x;
    cplx* A = Aall + size_t(b) * strideA;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;     // w in [0, 16): columns w and w + 16

    cplx col[2][MAXT];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
            const int row = lane + 32 * t, cw = w + 16 * h;
            col[h][t] = (cw < nbc && row < m) ? A[size_t(j0 + cw) * D + j0 + row] : make_double2(0, 0);
        }
    for (int i = tid; i < 32 * 33; i += blockDim.x) { Gm[i] = make_double2(0, 0); Tm[i] = make_double2(0, 0); }
    __syncthreads();

    for (int c = 0; c < nbc; ++c) {
        cplx* vb = vbuf + (c & 1) * m;
        if (w == (c & 15)) {
            // ---- reflector of the own column (zlarfg conventions).  This warp is the serial critical path of
            // the column step, so the code is straight-line: the slot (which of the two columns) is a
            // warp-uniform branch around two copies, rsqrt + one reciprocal replace sqrt + divisions.
            auto reflect = [&](cplx (&a)[MAXT]) {
                double xn = 0;
                cplx alpha = make_double2(0, 0);
#pragma unroll
                for (int t = 0; t < MAXT; ++t) {
                    const int row = lane + 32 * t;
                    const double n2 = fma(a[t].x, a[t].x, a[t].y * a[t].y);
                    xn += (row > c && row < m) ? n2 : 0.0;
                    alpha.x = row == c ? a[t].x : alpha.x;
                    alpha.y = row == c ? a[t].y : alpha.y;
                }
                xn = warp_sum(xn);
                alpha.x = __shfl_sync(0xffffffffu, alpha.x, c & 31);
                alpha.y = __shfl_sync(0xffffffffu, alpha.y, c & 31);
                const bool trivial = xn == 0.0 && alpha.y == 0.0;
                const double x2 = alpha.x * alpha.x + alpha.y * alpha.y + xn;
                const double inrm = trivial ? 0.0 : rsqrt(x2);
                const double nrm = x2 * inrm;
                const double beta = trivial ? alpha.x : (alpha.x >= 0 ? -nrm : nrm);
                const double ib = alpha.x >= 0 ? -inrm : inrm;                    // 1 / beta
                const cplx tau = trivial ? make_double2(0, 0) : make_double2((beta - alpha.x) * ib, -alpha.y * ib);
                const double dr = alpha.x - beta, di = alpha.y;
                const double iden = trivial ? 0.0 : __drcp_rn(dr * dr + di * di);
                const cplx sc = make_double2(dr * iden, -di * iden);
#pragma unroll
                for (int t = 0; t < MAXT; ++t) {
                    const int row = lane + 32 * t;
                    const cplx scaled = cmul(a[t], sc);
                    const cplx vnew = row > c ? scaled : make_double2(row == c ? 1.0 : 0.0, 0.0);
                    if (row < m) vb[row] = vnew;                                  // rows < c are never read
                    a[t].x = row > c ? scaled.x : (row == c ? beta : a[t].x);     // row c keeps the R diagonal
                    a[t].y = row > c ? scaled.y : (row == c ? 0.0 : a[t].y);
                }
                if (lane == 0) s_tau[c] = tau;
            };
            if (c < 16) reflect(col[0]); else reflect(col[1]);
        }
        __syncthreads();
        const cplx tau = s_tau[c];
        if (tau.x != 0.0 || tau.y != 0.0) {
            // dot_h = v_c^H x_h over rows >= c for both own columns (a_w for w > c, v_w for w < c)
            double dr[2] = {0, 0}, di[2] = {0, 0};
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                const int row = lane + 32 * t;
                const cplx v = (row >= c && row < m) ? vb[row] : make_double2(0, 0);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    dr[h] = fma(v.x, col[h][t].x, fma(v.y, col[h][t].y, dr[h]));
                    di[h] = fma(v.x, col[h][t].y, fma(-v.y, col[h][t].x, di[h]));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    dr[h] += __shfl_xor_sync(0xffffffffu, dr[h], o);
                    di[h] += __shfl_xor_sync(0xffffffffu, di[h], o);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cw = w + 16 * h;
                if (cw >= nbc || cw == c) continue;
                if (cw > c) {
                    // a -= conj(tau) (v^H a) v
                    const cplx fw = cmul(make_double2(tau.x, -tau.y), make_double2(dr[h], di[h]));
#pragma unroll
                    for (int t = 0; t < MAXT; ++t) {
                        const int row = lane + 32 * t;
                        const cplx v = (row >= c && row < m) ? vb[row] : make_double2(0, 0);
                        col[h][t].x -= fw.x * v.x - fw.y * v.y;
                        col[h][t].y -= fw.x * v.y + fw.y * v.x;
                    }
                } else if (lane == 0) {
                    Gm[cw * 33 + c] = make_double2(dr[h], -di[h]);   // v_cw^H v_c = conj(v_c^H v_cw)
                }
            }
        }
        // ---- compact-WY factor: column c-1 of T (its Gram column was completed in the previous step)
        if (c > 0 && w == ((c - 1) & 15)) {
            const int cc = c - 1;
            const cplx tcc = s_tau[cc];
            if (lane == cc) Tm[cc * 33 + cc] = tcc;
            if (lane < cc) {
                cplx sacc = make_double2(0, 0);
                for (int l = lane; l < cc; ++l) sacc = cfma_(Tm[lane * 33 + l], Gm[l * 33 + cc], sacc);
                Tm[lane * 33 + cc] = make_double2(-(tcc.x * sacc.x - tcc.y * sacc.y), -(tcc.x * sacc.y + tcc.y * sacc.x));
            }
        }
    }
    __syncthreads();
    if (w == ((nbc - 1) & 15)) {
        const int cc = nbc - 1;
        const cplx tcc = s_tau[cc];
        if (lane == cc) Tm[cc * 33 + cc] = tcc;
        if (lane < cc) {
            cplx sacc = make_double2(0, 0);
            for (int l = lane; l < cc; ++l) sacc = cfma_(Tm[lane * 33 + l], Gm[l * 33 + cc], sacc);
            Tm[lane * 33 + cc] = make_double2(-(tcc.x * sacc.x - tcc.y * sacc.y), -(tcc.x * sacc.y + tcc.y * sacc.x));
        }
    }
    // ---- panel to shared memory (R on / above the diagonal, v below)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int cw = w + 16 * h;
        if (cw < nbc) {
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                const int row = lane + 32 * t;
                if (row < m) P[cw * ldp + row] = col[h][t];
            }
        }
    }
    __syncthreads();
    // ---- write back: R (and v below it) to A, explicit V and V*T to the workspace
    cplx* V = Vall + size_t(b) * strideV;
    cplx* VT = VTall + size_t(b) * strideV;
    for (int idx = tid; idx < nbc * m; idx += blockDim.x) {
        const int c = idx / m, r = idx - c * m;
        const cplx pv = P[c * ldp + r];
        A[size_t(j0 + c) * D + j0 + r] = pv;
        V[size_t(j0 + c) * D + j0 + r] = r > c ? pv : make_double2(r == c ? 1.0 : 0.0, 0.0);
        cplx sacc = make_double2(0, 0);
        const int lmax = min(c, r);
        for (int l = 0; l <= lmax; ++l) {
            const cplx vl = (l == r) ? make_double2(1, 0) : P[l * ldp + r];
            sacc = cfma_(vl, Tm[l * 33 + c], sacc);
        }
        VT[size_t(j0 + c) * D + j0 + r] = sacc;
    }
}

// ------------------------------------------------------------------------------------------------
// Panel factorisation, second generation: the 32-column panel is factored as four 8-column sub-panels.
//
//   * A sub-panel is a chain of eight column steps on eight warps (warp <-> column in registers, lane <-> row):
//     reflector of the own column, one named barrier, one dot product + rank-1 update per warp.  Only the eight
//     columns of the sub-panel take part, so a step costs a quarter of the instructions of the 32-column version
//     and its latency is that of two warp reductions.
//   * The remaining columns of the panel receive the sub-panel's block reflector  A -= V_s (T_s^H (V_s^H A))  on the
//     FP64 tensor cores (DMMA m8n8k4), all sixteen warps, operands in shared memory.
//   * The 32 x 32 compact-WY factor T is assembled from the sub-panel factors and the cross Gram blocks
//     V_s^H V_t (DMMA), and V T is formed on the tensor cores as well.
// Same interface and results (up to summation order) as qr_panel_reg_kernel: R and the reflectors in A, explicit unit
// lower-trapezoidal V and V T in the workspace.
// ------------------------------------------------------------------------------------------------
constexpr int kP2Threads = 512;
constexpr int kP2Sub = 8;                      // columns per sub-panel
constexpr int kP2KSplit = 4;                   // row splits of the V_s^H A products

__device__ __forceinline__ void dmma_qr(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct P2Layout {
    int ldp;            // leading dimension of the panel (== 4 mod 8: conflict-free (column, 4 rows) fragment loads)
    size_t P, vbuf, G, T, Zp, Zs, W, total;   // offsets in cplx elements
};
__host__ __device__ inline P2Layout p2_layout(int m) {
    P2Layout l;
    l.ldp = ((m + 7) & ~7) + 4;
    size_t p = 0;
    l.P = p;    p += size_t(32) * l.ldp;
    l.vbuf = p; p += size_t(2) * 32 * 10;                // two reflector buffers of 32 * MAXT rows (zero beyond row m)
    l.G = p;    p += 32 * 33;
    l.T = p;    p += 32 * 33;
    l.Zp = p;   p += size_t(kP2KSplit) * 3 * 64;     // partial 8 x 24 products (three 8 x 8 blocks per split)
    l.Zs = p;   p += 8 * 24;
    l.W = p;    p += 8 * 24;
    l.total = p;
    return l;
}

// 1 / sqrt(x) and 1 / x for normal positive x: hardware approximation + two Newton steps (the library versions
// handle special cases that cannot occur here and cost ~3x the instructions on the serial path of a column step)
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return y;
}
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * fma(-x, y, 2.0);
    y = y * fma(-x, y, 2.0);
    return y;
}

// element (row r, column c) of the unit lower-trapezoidal V held in the panel buffer (R sits on / above the diagonal)
__device__ __forceinline__ cplx p2_vget(const cplx* P, int ldp, int r, int c, int m) {
    cplx v = make_double2(0, 0);
    if (r < m) {
        v = P[size_t(c) * ldp + r];
        if (r <= c) v = make_double2(r == c ? 1.0 : 0.0, 0.0);
    }
    return v;
}

template <int MAXT>
__global__ void __launch_bounds__(kP2Threads) qr_panel2_kernel(cplx* Aall, long long strideA, int D, int j0, int nbc,
                                                               cplx* Vall, cplx* VTall, long long strideV) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = D - j0;
    const P2Layout lay = p2_layout(m);
    const int ldp = lay.ldp;
    cplx* sm = reinterpret_cast<cplx*>(smem_raw);
    cplx* P = sm + lay.P;
    cplx* vbuf = sm + lay.vbuf;
    cplx* Gm = sm + lay.G;                       // [32][33] Gram matrix of the reflectors, strict upper part
    cplx* Tm = sm + lay.T;                       // [32][33] compact-WY factor
    cplx* Zp = sm + lay.Zp;
    cplx* Zs = sm + lay.Zs;
    cplx* Ws = sm + lay.W;
    __shared__ cplx s_tau[32];

    const int b = blockIdx.x;
    cplx* A = Aall + size_t(b) * strideA;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int grp = lane >> 2, t4 = lane & 3;

    // ---- panel to shared memory (rows beyond m are never read: every fragment load is predicated on r < m)
    for (int idx = tid; idx < nbc * m; idx += kP2Threads) {
        const int c = idx / m, r = idx - c * m;
        P[size_t(c) * ldp + r] = A[size_t(j0 + c) * D + j0 + r];
    }
    for (int i = tid; i < 32 * 33; i += kP2Threads) { Gm[i] = make_double2(0, 0); Tm[i] = make_double2(0, 0); }
    for (int i = tid; i < 2 * 32 * MAXT; i += kP2Threads) vbuf[i] = make_double2(0, 0);
    __syncthreads();

    for (int c0 = 0; c0 < nbc; c0 += kP2Sub) {
        const int ns = min(kP2Sub, nbc - c0);                       // columns of this sub-panel
        {
            // =========================================================== column steps of the sub-panel
            // warps 0..7: warp <-> column c0 + w in registers, lane holds rows c0 + lane + 32 t (zero beyond row m, so
            // no per-row predicates: only the first 32-row block contains rows at or above the diagonal).
            // warps 8..15: cross Gram entries v_j^H v_c with the reflectors of the earlier sub-panels (for T).
            cplx a[MAXT];
            const bool colwarp = w < kP2Sub;
            const bool live = colwarp && w < ns;
            if (colwarp) {
#pragma unroll
                for (int t = 0; t < MAXT; ++t) {
                    const int r = c0 + lane + 32 * t;
                    a[t] = (live && r < m) ? P[size_t(c0 + w) * ldp + r] : make_double2(0, 0);
                }
            }
            for (int cl = 0; cl < ns; ++cl) {
                const int c = c0 + cl;
                cplx* vb = vbuf + (cl & 1) * (32 * MAXT);
                if (w == cl) {
                    // reflector of the own column (zlarfg conventions); row c is lane cl of the first row block
                    double xn0 = (lane > cl) ? fma(a[0].x, a[0].x, a[0].y * a[0].y) : 0.0, xn1 = 0.0;
#pragma unroll
                    for (int t = 1; t < MAXT; ++t) {
                        if (t & 1) xn1 = fma(a[t].x, a[t].x, fma(a[t].y, a[t].y, xn1));
                        else xn0 = fma(a[t].x, a[t].x, fma(a[t].y, a[t].y, xn0));
                    }
                    const double xn = warp_sum(xn0 + xn1);
                    const double alr = __shfl_sync(0xffffffffu, a[0].x, cl);
                    const double ali = __shfl_sync(0xffffffffu, a[0].y, cl);
                    const bool trivial = xn == 0.0 && ali == 0.0;
                    const double x2 = alr * alr + ali * ali + xn;
                    const double inrm = trivial ? 0.0 : fast_rsqrt(x2);
                    const double nrm = x2 * inrm;
                    const double beta = trivial ? alr : (alr >= 0 ? -nrm : nrm);
                    const double ib = alr >= 0 ? -inrm : inrm;                        // 1 / beta
                    const cplx tau = trivial ? make_double2(0, 0) : make_double2((beta - alr) * ib, -ali * ib);
                    const double dr = alr - beta, di = ali;
                    const double iden = trivial ? 0.0 : fast_rcp(dr * dr + di * di);
                    const cplx sc = make_double2(dr * iden, -di * iden);
                    {
                        const cplx scaled = cmul(a[0], sc);
                        const cplx vnew = lane > cl ? scaled : make_double2(lane == cl ? 1.0 : 0.0, 0.0);
                        vb[lane] = vnew;
                        a[0].x = lane > cl ? scaled.x : (lane == cl ? beta : a[0].x);  // row c keeps the R diagonal
                        a[0].y = lane > cl ? scaled.y : (lane == cl ? 0.0 : a[0].y);
                    }
#pragma unroll
                    for (int t = 1; t < MAXT; ++t) {
                        a[t] = cmul(a[t], sc);
                        vb[lane + 32 * t] = a[t];
                    }
                    if (lane == 0) s_tau[c] = tau;
                }
                __syncthreads();
                const cplx tau = s_tau[c];
                const bool nontrivial = tau.x != 0.0 || tau.y != 0.0;
                if (colwarp) {
                    if (w != cl && live && nontrivial) {
                        // dot = v_c^H x (v is zero above row c); x = own column (w > cl) or own reflector (w < cl)
                        double dr0 = 0, di0 = 0, dr1 = 0, di1 = 0;
                        cplx vv[MAXT];
#pragma unroll
                        for (int t = 0; t < MAXT; ++t) vv[t] = vb[lane + 32 * t];
                        {
                            cplx x = a[0];
                            if (w < cl) x = lane > w ? a[0] : make_double2(lane == w ? 1.0 : 0.0, 0.0);   // unit diagonal, zero above
                            dr0 = fma(vv[0].x, x.x, vv[0].y * x.y);
                            di0 = fma(vv[0].x, x.y, -vv[0].y * x.x);
                        }
#pragma unroll
                        for (int t = 1; t < MAXT; ++t) {
                            if (t & 1) {
                                dr1 = fma(vv[t].x, a[t].x, fma(vv[t].y, a[t].y, dr1));
                                di1 = fma(vv[t].x, a[t].y, fma(-vv[t].y, a[t].x, di1));
                            } else {
                                dr0 = fma(vv[t].x, a[t].x, fma(vv[t].y, a[t].y, dr0));
                                di0 = fma(vv[t].x, a[t].y, fma(-vv[t].y, a[t].x, di0));
                            }
                        }
                        double dr = dr0 + dr1, di = di0 + di1;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            dr += __shfl_xor_sync(0xffffffffu, dr, o);
                            di += __shfl_xor_sync(0xffffffffu, di, o);
                        }
                        if (w > cl) {
                            const cplx fw = cmul(make_double2(tau.x, -tau.y), make_double2(dr, di));   // conj(tau) (v^H a)
#pragma unroll
                            for (int t = 0; t < MAXT; ++t) {
                                a[t].x -= fw.x * vv[t].x - fw.y * vv[t].y;
                                a[t].y -= fw.x * vv[t].y + fw.y * vv[t].x;
                            }
                        } else if (lane == 0) {
                            Gm[(c0 + w) * 33 + c] = make_double2(dr, -di);             // v_w^H v_c = conj(v_c^H v_w)
                        }
                    }
                    // compact-WY factor of the sub-panel, column cl - 1 (its Gram column was completed in the previous step)
                    if (cl > 0 && w == cl - 1) {
                        const int cc = c - 1;
                        const cplx tcc = s_tau[cc];
                        if (lane == 0) Tm[cc * 33 + cc] = tcc;
                        const int i = c0 + lane;
                        if (i < cc) {
                            cplx sacc = make_double2(0, 0);
                            for (int l = i; l < cc; ++l) sacc = cfma_(Tm[i * 33 + l], Gm[l * 33 + cc], sacc);
                            Tm[i * 33 + cc] = make_double2(-(tcc.x * sacc.x - tcc.y * sacc.y), -(tcc.x * sacc.y + tcc.y * sacc.x));
                        }
                    }
                } else if (nontrivial) {
                    // cross Gram: G[j, c] = v_j^H v_c for the reflectors j < c0 of the earlier sub-panels (rows >= c only:
                    // v_c vanishes above; V[r, j] = P[j][r] there because r >= c > j)
                    for (int j = w - kP2Sub; j < c0; j += kP2Sub) {
                        const cplx* vj = P + size_t(j) * ldp + c0;
                        double gr0 = 0, gi0 = 0, gr1 = 0, gi1 = 0;
#pragma unroll
                        for (int t = 0; t < MAXT; ++t) {
                            const int rr = lane + 32 * t;
                            if (c0 + rr < m) {
                                const cplx x = vj[rr], v = vb[rr];
                                if (t & 1) {
                                    gr1 = fma(x.x, v.x, fma(x.y, v.y, gr1));           // conj(x) v
                                    gi1 = fma(x.x, v.y, fma(-x.y, v.x, gi1));
                                } else {
                                    gr0 = fma(x.x, v.x, fma(x.y, v.y, gr0));
                                    gi0 = fma(x.x, v.y, fma(-x.y, v.x, gi0));
                                }
                            }
                        }
                        double gr = gr0 + gr1, gi = gi0 + gi1;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            gr += __shfl_xor_sync(0xffffffffu, gr, o);
                            gi += __shfl_xor_sync(0xffffffffu, gi, o);
                        }
                        if (lane == 0) Gm[j * 33 + c] = make_double2(gr, gi);
                    }
                }
            }
            if (live) {
#pragma unroll
                for (int t = 0; t < MAXT; ++t) {
                    const int r = c0 + lane + 32 * t;
                    if (r < m) P[size_t(c0 + w) * ldp + r] = a[t];
                }
            }
        }
        __syncthreads();
        if (w == 0) {
            const int cc = c0 + ns - 1;
            const cplx tcc = s_tau[cc];
            if (lane == 0) Tm[cc * 33 + cc] = tcc;
            const int i = c0 + lane;
            if (i < cc) {
                cplx sacc = make_double2(0, 0);
                for (int l = i; l < cc; ++l) sacc = cfma_(Tm[i * 33 + l], Gm[l * 33 + cc], sacc);
                Tm[i * 33 + cc] = make_double2(-(tcc.x * sacc.x - tcc.y * sacc.y), -(tcc.x * sacc.y + tcc.y * sacc.x));
            }
        }
        const int nrest = nbc - (c0 + ns);                          // panel columns to the right of the sub-panel
        if (nrest <= 0) { __syncthreads(); break; }
        const int nblk = (nrest + 7) / 8;
        const int mloc = m - c0;                                    // rows c0 .. m-1 take part
        const int nk4 = (mloc + 3) / 4;
        const int k4per = (nk4 + kP2KSplit - 1) / kP2KSplit;
        // ---- Z = V_s^H A_rest (8 x nrest): warp <-> (8-column block, row split), partial sums in shared memory
        if (w < 3 * kP2KSplit) {
            const int blk = w % 3, ks = w / 3;
            if (blk < nblk) {
                double zr0 = 0, zr1 = 0, zi0 = 0, zi1 = 0;
                const int cv = c0 + grp;                            // A fragment: conj(V_s[k, i = grp])
                const int cb = c0 + ns + blk * 8 + grp;             // B fragment: A_rest[k, n = grp]
                const bool cbok = blk * 8 + grp < nrest, cvok = grp < ns;
                for (int s4 = ks * k4per; s4 < min(nk4, (ks + 1) * k4per); ++s4) {
                    const int r = c0 + 4 * s4 + t4;
                    const cplx av = cvok ? p2_vget(P, ldp, r, cv, m) : make_double2(0, 0);
                    const cplx bv = (cbok && r < m) ? P[size_t(cb) * ldp + r] : make_double2(0, 0);
                    // conj(a) b = (ax bx + ay by) + i (ax by - ay bx)
                    dmma_qr(zr0, zr1, av.x, bv.x);
                    dmma_qr(zr0, zr1, av.y, bv.y);
                    dmma_qr(zi0, zi1, av.x, bv.y);
                    dmma_qr(zi0, zi1, -av.y, bv.x);
                }
                cplx* zp = Zp + (size_t(ks) * 3 + blk) * 64;        // [i = grp][n = 2 t4 + e]
                zp[grp * 8 + 2 * t4] = make_double2(zr0, zi0);
                zp[grp * 8 + 2 * t4 + 1] = make_double2(zr1, zi1);
            }
        }
        __syncthreads();
        // ---- fixed-order sum of the partial products, then W = T_s^H Z
        if (tid < 8 * 24) {
            const int i = tid / 24, n = tid - i * 24;
            cplx z = make_double2(0, 0);
            if (n < nrest && i < ns) {
                const int blk = n >> 3, nn = n & 7;
#pragma unroll
                for (int ks = 0; ks < kP2KSplit; ++ks) {
                    const cplx p = Zp[(size_t(ks) * 3 + blk) * 64 + i * 8 + nn];
                    z.x += p.x;
                    z.y += p.y;
                }
            }
            Zs[i * 24 + n] = z;
        }
        __syncthreads();
        if (tid < 8 * 24) {
            const int i = tid / 24, n = tid - i * 24;
            cplx wv = make_double2(0, 0);
            if (i < ns)
                for (int l = 0; l <= i; ++l) {                      // (T_s^H)[i, l] = conj(T_s[l, i])
                    const cplx t = Tm[(c0 + l) * 33 + c0 + i];
                    wv = cfmac_(t, Zs[l * 24 + n], wv);
                }
            Ws[i * 24 + n] = wv;
        }
        __syncthreads();
        // ---- A_rest -= V_s W on the tensor cores: task = (8-row block, 8-column block)
        {
            const int nrb = (mloc + 7) / 8;
            for (int task = w; task < nrb * nblk; task += kP2Threads / 32) {
                const int rb = task / nblk, blk = task - rb * nblk;
                const int r = c0 + rb * 8 + grp;                    // accumulator row
                const int cn = c0 + ns + blk * 8 + 2 * t4;          // accumulator columns cn, cn + 1
                const bool ok0 = r < m && blk * 8 + 2 * t4 < nrest, ok1 = r < m && blk * 8 + 2 * t4 + 1 < nrest;
                cplx x0 = ok0 ? P[size_t(cn) * ldp + r] : make_double2(0, 0);
                cplx x1 = ok1 ? P[size_t(cn + 1) * ldp + r] : make_double2(0, 0);
#pragma unroll
                for (int k4 = 0; k4 < kP2Sub; k4 += 4) {
                    const int i = k4 + t4;                          // A fragment: -V_s[r, i]; B fragment: W[i, n = grp]
                    cplx av = i < ns ? p2_vget(P, ldp, r, c0 + i, m) : make_double2(0, 0);
                    av.x = -av.x; av.y = -av.y;
                    const cplx bv = Ws[i * 24 + blk * 8 + grp];
                    dmma_qr(x0.x, x1.x, av.x, bv.x);
                    dmma_qr(x0.x, x1.x, -av.y, bv.y);
                    dmma_qr(x0.y, x1.y, av.x, bv.y);
                    dmma_qr(x0.y, x1.y, av.y, bv.x);
                }
                if (ok0) P[size_t(cn) * ldp + r] = x0;
                if (ok1) P[size_t(cn + 1) * ldp + r] = x1;
            }
        }
        __syncthreads();
    }

    // ---- off-diagonal blocks of T (the cross Gram blocks V_s^H V_t were accumulated during the column steps)
    {
        const int nsub = (nbc + kP2Sub - 1) / kP2Sub;
        // ---- off-diagonal blocks of T, block column by block column:  T[0:c0, c0:c0+8] = -T[0:c0, 0:c0] G[0:c0, c0:c0+8] T_tt
        for (int tb = 1; tb < nsub; ++tb) {
            const int c0 = tb * 8, nt = min(8, nbc - c0);
            // X = G[0:c0, c0:c0+nt] T_tt   (into Zs-like scratch: reuse Zp, c0 x 8)
            for (int e = tid; e < c0 * 8; e += kP2Threads) {
                const int i = e >> 3, cc = e & 7;
                cplx x = make_double2(0, 0);
                if (cc < nt)
                    for (int l = 0; l <= cc; ++l) x = cfma_(Gm[i * 33 + c0 + l], Tm[(c0 + l) * 33 + c0 + cc], x);
                Zp[e] = x;
            }
            __syncthreads();
            for (int e = tid; e < c0 * 8; e += kP2Threads) {
                const int i = e >> 3, cc = e & 7;
                if (cc < nt) {
                    cplx x = make_double2(0, 0);
                    for (int l = i; l < c0; ++l) x = cfma_(Tm[i * 33 + l], Zp[l * 8 + cc], x);
                    Tm[i * 33 + c0 + cc] = make_double2(-x.x, -x.y);
                }
            }
            __syncthreads();
        }
    }

    // ---- write back: R (and v below it) to A, explicit V to the workspace
    cplx* V = Vall + size_t(b) * strideV;
    cplx* VT = VTall + size_t(b) * strideV;
    for (int idx = tid; idx < nbc * m; idx += kP2Threads) {
        const int c = idx / m, r = idx - c * m;
        const cplx pv = P[size_t(c) * ldp + r];
        A[size_t(j0 + c) * D + j0 + r] = pv;
        V[size_t(j0 + c) * D + j0 + r] = r > c ? pv : make_double2(r == c ? 1.0 : 0.0, 0.0);
    }
    // ---- V T on the tensor cores: task = (8-row block, 8-column block), T upper triangular
    {
        const int nrb = (m + 7) / 8, ncb = (nbc + 7) / 8;
        for (int task = w; task < nrb * ncb; task += kP2Threads / 32) {
            const int rb = task / ncb, cb = task - rb * ncb;
            const int r = rb * 8 + grp;
            double xr0 = 0, xr1 = 0, xi0 = 0, xi1 = 0;
            for (int k4 = 0; k4 < (cb + 1) * 8; k4 += 4) {
                const int l = k4 + t4;                              // A fragment: V[r, l]; B fragment: T[l, c = cb 8 + grp]
                const cplx av = l < nbc ? p2_vget(P, ldp, r, l, m) : make_double2(0, 0);
                const cplx bv = (l < nbc && cb * 8 + grp < nbc) ? Tm[l * 33 + cb * 8 + grp] : make_double2(0, 0);
                dmma_qr(xr0, xr1, av.x, bv.x);
                dmma_qr(xr0, xr1, -av.y, bv.y);
                dmma_qr(xi0, xi1, av.x, bv.y);
                dmma_qr(xi0, xi1, av.y, bv.x);
            }
            const int cn = cb * 8 + 2 * t4;
            if (r < m && cn < nbc) VT[size_t(j0 + cn) * D + j0 + r] = make_double2(xr0, xi0);
            if (r < m && cn + 1 < nbc) VT[size_t(j0 + cn + 1) * D + j0 + r] = make_double2(xr1, xi1);
        }
    }
}

// inverse of every nb x nb diagonal block of the upper-triangular R (stored in A): out [batch][P][nb*nb]
__global__ void trtri_blocks_kernel(const cplx* Aall, long long strideA, int D, int nb, cplx* outAll, long long strideOut) {
    pdl_enter();
    __shared__ cplx Rs[32 * 33], Xs[32 * 33];
    const int b = blockIdx.y, p = blockIdx.x;
    const int j0 = p * nb, nbc = min(nb, D - j0);
    const cplx* A = Aall + size_t(b) * strideA;
    for (int idx = threadIdx.x; idx < nbc * nbc; idx += blockDim.x) {
        const int c = idx / nbc, r = idx - c * nbc;
        Rs[r * 33 + c] = A[size_t(j0 + c) * D + j0 + r];
    }
    __syncthreads();
    if (threadIdx.x < nbc) {
        const int c = threadIdx.x;
        for (int i = nbc - 1; i > c; --i) Xs[i * 33 + c] = make_double2(0, 0);
        const cplx rc = Rs[c * 33 + c];
        const double dc = rc.x * rc.x + rc.y * rc.y;
        Xs[c * 33 + c] = make_double2(rc.x / dc, -rc.y / dc);
        for (int i = c - 1; i >= 0; --i) {
            cplx s = make_double2(0, 0);
            for (int l = i + 1; l <= c; ++l) s = cfma_(Rs[i * 33 + l], Xs[l * 33 + c], s);
            const cplx ri = Rs[i * 33 + i];
            const double di = ri.x * ri.x + ri.y * ri.y;
            // -s / r_ii
            Xs[i * 33 + c] = make_double2(-(s.x * ri.x + s.y * ri.y) / di, -(s.y * ri.x - s.x * ri.y) / di);
        }
    }
    __syncthreads();
    cplx* out = outAll + size_t(b) * strideOut + size_t(p) * nb * nb;
    for (int idx = threadIdx.x; idx < nbc * nbc; idx += blockDim.x) {
        const int c = idx / nbc, r = idx - c * nbc;
        out[c * nb + r] = Xs[r * 33 + c];
    }
}

inline GemmArgs gemm_args(int M, int N, int K, int ta, int tb, const cplx* A, int lda, long long sA, const cplx* B,
                          int ldb, long long sB, cplx* C, int ldc, long long sC, double alpha, double beta, int batch) {
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.transa = ta; g.transb = tb;
    g.A = A; g.lda = lda; g.strideA = sA;
    g.B = B; g.ldb = ldb; g.strideB = sB;
    g.C = C; g.ldc = ldc; g.strideC = sC;
    g.rowscale = g.colscale = g.kscale = nullptr;
    g.strideRow = g.strideCol = g.strideK = 0;
    g.alpha = alpha; g.beta = beta; g.batch = batch; g.kvec = nullptr; g.b_kmajor = 0;
    return g;
}

#define QR_TRY(call)                                  \
    do {                                              \
        cudaError_t e__ = (call);                     \
        if (e__ != cudaSuccess) return e__;           \
    } while (0)

}  // namespace

// ---- blocked QR host drivers --------------------------------------------------------------------
int qr_choose_nb(int D) {
    // the panel (D x nb complex) plus the two 32 x 33 factors must fit the 227 KB of one CTA
    for (int nb : {32, 16, 8, 4}) {
        const size_t need = size_t(nb) * (D | 1) * sizeof(cplx) + 2 * 32 * 33 * sizeof(cplx) + 1024;
        if (need <= 220 * 1024) return nb;
    }
    return 2;
}

cudaError_t qr_workspace_create(QrWorkspace* ws, int D, int batch) {
    ws->D = D; ws->batch = batch; ws->nb = qr_choose_nb(D);
    ws->launches = 0;
    const size_t dd = size_t(D) * D;
    QR_TRY(cudaMalloc(reinterpret_cast<void**>(&ws->V), sizeof(cplx) * dd * batch));
    QR_TRY(cudaMalloc(reinterpret_cast<void**>(&ws->VT), sizeof(cplx) * dd * batch));
    QR_TRY(cudaMalloc(reinterpret_cast<void**>(&ws->W), sizeof(cplx) * size_t(ws->nb) * D * batch));
    QR_TRY(cudaMalloc(reinterpret_cast<void**>(&ws->Rinv), sizeof(cplx) * size_t(ws->nb) * (D + ws->nb) * batch));
    QR_TRY(cudaMemset(ws->V, 0, sizeof(cplx) * dd * batch));
    QR_TRY(cudaMemset(ws->VT, 0, sizeof(cplx) * dd * batch));
    return cudaSuccess;
}

void qr_workspace_destroy(QrWorkspace* ws) {
    cudaFree(ws->V); cudaFree(ws->VT); cudaFree(ws->W); cudaFree(ws->Rinv);
    ws->V = ws->VT = ws->W = ws->Rinv = nullptr;
}

cudaError_t qr_prepivot_launch(const cplx* A, long long strideA, cplx* Aout, long long strideOut, int* perm,
                               double* norms, int D, int batch, cudaStream_t st) {
    const int CS = D >= 128 ? 8 : 1;
    QR_TRY(launch_pdl_cluster(colnorm_rank_kernel, dim3(batch * CS), dim3(1024), size_t(D) * sizeof(double), st, (unsigned)CS, A, D,
                              strideA, perm, norms, CS));
    QR_TRY(cudaGetLastError());
    dim3 grid(D, batch);
    launch_pdl(permute_columns_kernel, dim3(grid), dim3(128), 0, st, A, Aout, perm, D, strideA, strideOut);
    return cudaGetLastError();
}

cudaError_t permute_rows_launch(const cplx* in, long long strideIn, cplx* out, long long strideOut, const int* perm,
                                int D, int batch, cudaStream_t st) {
    dim3 grid(D, batch);
    launch_pdl(permute_rows_kernel, dim3(grid), dim3(128), 0, st, in, out, perm, D, strideIn, strideOut);
    return cudaGetLastError();
}

// A (already column-ordered) -> R in the upper triangle; V, V*T of all panels in ws (slices off ..)
cudaError_t qr_blocked_factor(QrWorkspace& ws, cplx* A, int D, long long strideA, int off, int batch, cudaStream_t st) {
    const int nb = ws.nb;
    const size_t dd = size_t(D) * D;
    cplx* V = ws.V + size_t(off) * dd;
    cplx* VT = ws.VT + size_t(off) * dd;
    cplx* W = ws.W + size_t(off) * nb * D;
    const size_t smem = size_t(nb) * (D | 1) * sizeof(cplx) + 2 * 32 * 33 * sizeof(cplx);
    QR_TRY(cudaFuncSetAttribute(qr_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QR_TRY(cudaFuncSetAttribute(qr_panel_reg_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(smem + size_t(2) * D * sizeof(cplx))));
    QR_TRY(cudaFuncSetAttribute(qr_panel2_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(p2_layout(std::min(D, 320)).total * sizeof(cplx))));
    for (int j0 = 0; j0 < D; j0 += nb) {
        const int nbc = std::min(nb, D - j0), m = D - j0, n2 = D - j0 - nbc;
        const size_t sm = size_t(nbc) * (m | 1) * sizeof(cplx) + 2 * 32 * 33 * sizeof(cplx);
        static const bool force_smem_panel = std::getenv("DQMC_QR_SMEM_PANEL") != nullptr;
        static const bool old_reg_panel = std::getenv("DQMC_QR_REG_PANEL") != nullptr;
        if (m <= 320 && nb == 32 && !force_smem_panel && !old_reg_panel) {
            const size_t sm2 = p2_layout(m).total * sizeof(cplx);
            launch_pdl(qr_panel2_kernel<10>, dim3(batch), dim3(kP2Threads), sm2, st, A, strideA, D, j0, nbc, V, VT, (long long)dd);
        } else if (m <= 320 && nb == 32 && !force_smem_panel) {
            launch_pdl(qr_panel_reg_kernel<10>, dim3(batch), dim3(kPanelRegThreads), sm + size_t(2) * m * sizeof(cplx), st, A, strideA, D, j0, nbc, V,
                                                                                                  VT, (long long)dd);
        } else {
            launch_pdl(qr_panel_kernel, dim3(batch), dim3(kPanelThreads), sm, st, A, strideA, D, j0, nbc, V, VT, (long long)dd);
        }
        QR_TRY(cudaGetLastError());
        ws.launches += 1;
        if (n2 > 0) {
            const size_t po = size_t(j0) * D + j0;                  // panel origin
            const size_t co = size_t(j0 + nbc) * D + j0;            // trailing block origin
            // W = (V T)^H C ;  C -= V W
            QR_TRY(gemm_launch(gemm_args(nbc, n2, m, 1, 0, VT + po, D, (long long)dd, A + co, D, strideA, W, nb,
                                         (long long)nb * D, 1.0, 0.0, batch), st));
            QR_TRY(gemm_launch(gemm_args(m, n2, nbc, 0, 0, V + po, D, (long long)dd, W, nb, (long long)nb * D, A + co, D,
                                         strideA, -1.0, 1.0, batch), st));
            ws.launches += 2;
        }
    }
    return cudaSuccess;
}

// Q = H_0 H_1 ... H_{P-1} explicitly (zungqr, backward accumulation of the block reflectors)
cudaError_t qr_blocked_form_q(QrWorkspace& ws, cplx* Q, int D, long long strideQ, int off, int batch, cudaStream_t st) {
    const int nb = ws.nb;
    const size_t dd = size_t(D) * D;
    const cplx* V = ws.V + size_t(off) * dd;
    const cplx* VT = ws.VT + size_t(off) * dd;
    cplx* W = ws.W + size_t(off) * nb * D;
    QR_TRY(launch_set_identity(Q, D, strideQ, batch, st));
    ws.launches += 1;
    const int np = (D + nb - 1) / nb;
    for (int p = np - 1; p >= 0; --p) {
        const int j0 = p * nb, nbc = std::min(nb, D - j0), m = D - j0;
        const size_t po = size_t(j0) * D + j0;
        // W = V^H Q[j0:, j0:] ;  Q[j0:, j0:] -= (V T) W
        QR_TRY(gemm_launch(gemm_args(nbc, m, m, 1, 0, V + po, D, (long long)dd, Q + po, D, strideQ, W, nb, (long long)nb * D,
                                     1.0, 0.0, batch), st));
        QR_TRY(gemm_launch(gemm_args(m, m, nbc, 0, 0, VT + po, D, (long long)dd, W, nb, (long long)nb * D, Q + po, D, strideQ,
                                     -1.0, 1.0, batch), st));
        ws.launches += 2;
    }
    return cudaSuccess;
}

// C <- Q^H C for a D x ncols matrix C (ldc = D), panel by panel: C[j0:, :] -= V ((V T)^H C[j0:, :])
cudaError_t qr_blocked_apply_qh(QrWorkspace& ws, cplx* C, int D, int ncols, long long strideC, int off, int batch,
                                cudaStream_t st) {
    const int nb = ws.nb;
    const size_t dd = size_t(D) * D;
    const cplx* V = ws.V + size_t(off) * dd;
    const cplx* VT = ws.VT + size_t(off) * dd;
    cplx* W = ws.W + size_t(off) * nb * D;
    for (int j0 = 0; j0 < D; j0 += nb) {
        const int nbc = std::min(nb, D - j0), m = D - j0;
        const size_t po = size_t(j0) * D + j0;
        QR_TRY(gemm_launch(gemm_args(nbc, ncols, m, 1, 0, VT + po, D, (long long)dd, C + j0, D, strideC, W, nb,
                                     (long long)nb * D, 1.0, 0.0, batch), st));
        QR_TRY(gemm_launch(gemm_args(m, ncols, nbc, 0, 0, V + po, D, (long long)dd, W, nb, (long long)nb * D, C + j0, D,
                                     strideC, -1.0, 1.0, batch), st));
        ws.launches += 2;
    }
    return cudaSuccess;
}

// Solve R Z = Y (R = upper triangle of A, D x D right-hand sides).  Y is destroyed, Z is written
// to Zout.  Diagonal blocks are inverted once, everything else is GEMM.
cudaError_t trsm_upper_blocked(QrWorkspace& ws, const cplx* A, cplx* Y, cplx* Zout, int D, long long strideA, int off,
                               int batch, cudaStream_t st) {
    const int nb = ws.nb;
    const int np = (D + nb - 1) / nb;
    const long long sR = (long long)nb * (D + nb);
    cplx* Rinv = ws.Rinv + size_t(off) * sR;
    dim3 grid(np, batch);
    launch_pdl(trtri_blocks_kernel, dim3(grid), dim3(128), 0, st, A, strideA, D, nb, Rinv, sR);
    QR_TRY(cudaGetLastError());
    ws.launches += 1;
    for (int p = np - 1; p >= 0; --p) {
        const int j0 = p * nb, nbc = std::min(nb, D - j0);
        // Z_p = R_pp^-1 Y_p
        QR_TRY(gemm_launch(gemm_args(nbc, D, nbc, 0, 0, Rinv + size_t(p) * nb * nb, nb, sR, Y + j0, D, strideA, Zout + j0, D,
                                     strideA, 1.0, 0.0, batch), st));
        ws.launches += 1;
        if (j0 > 0) {
            // Y[0:j0, :] -= R[0:j0, blk] Z_p
            QR_TRY(gemm_launch(gemm_args(j0, D, nbc, 0, 0, A + size_t(j0) * D, D, strideA, Zout + j0, D, strideA, Y, D, strideA,
                                         -1.0, 1.0, batch), st));
            ws.launches += 1;
        }
    }
    return cudaSuccess;
}


cudaError_t qrcp_factor_launch(cplx* A, int D, long long strideA, cplx* tau, int* perm, double* colnorm,
                               int batch, cudaStream_t st) {
    const size_t smem = size_t(D) * sizeof(cplx);
    launch_pdl(qrcp_factor_kernel, dim3(batch), dim3(kQrThreads), smem, st, A, D, strideA, tau, perm, colnorm);
    return cudaGetLastError();
}

cudaError_t qr_form_q_launch(const cplx* A, const cplx* tau, cplx* Q, int D, long long strideA, int batch,
                             cudaStream_t st) {
    const size_t smem = size_t(D) * sizeof(cplx);
    launch_pdl(qr_form_q_kernel, dim3(batch), dim3(kQrThreads), smem, st, A, tau, Q, D, strideA);
    return cudaGetLastError();
}

cudaError_t qr_extract_dt_launch(const cplx* A, const int* perm, double* d, cplx* T, int D, long long strideA,
                                 int batch, cudaStream_t st) {
    dim3 grid((unsigned)((size_t(D) * D + 255) / 256), batch);
    launch_pdl(qr_extract_dt_kernel, dim3(grid), dim3(256), 0, st, A, perm, d, T, D, strideA);
    return cudaGetLastError();
}

cudaError_t trsm_upper_launch(const cplx* A, cplx* Y, cplx* Zout, const int* perm, int D, long long strideA,
                              int batch, cudaStream_t st) {
    const size_t smem = size_t(D) * kTrsmCols * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(trsm_upper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((D + kTrsmCols - 1) / kTrsmCols, batch);
    launch_pdl(trsm_upper_kernel, dim3(grid), dim3(kTrsmThreads), smem, st, A, Y, Zout, perm, D, strideA);
    return cudaGetLastError();
}

cudaError_t scale_split_launch(const double* d, double* inv_big, double* small_, double* logacc, int D,
                               int batch, cudaStream_t st) {
    launch_pdl(scale_split_kernel, dim3(batch), dim3(256), 0, st, d, inv_big, small_, logacc, D);
    return cudaGetLastError();
}

cudaError_t logdiag_accumulate_launch(const cplx* A, double* logacc, int D, long long strideA, int batch,
                                      cudaStream_t st) {
    launch_pdl(logdiag_kernel, dim3(batch), dim3(256), 0, st, A, logacc, D, strideA);
    return cudaGetLastError();
}

}  // namespace dqmc
