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

}  // namespace

cudaError_t qrcp_factor_launch(cplx* A, int D, long long strideA, cplx* tau, int* perm, double* colnorm,
                               int batch, cudaStream_t st) {
    const size_t smem = size_t(D) * sizeof(cplx);
    qrcp_factor_kernel<<<batch, kQrThreads, smem, st>>>(A, D, strideA, tau, perm, colnorm);
    return cudaGetLastError();
}

cudaError_t qr_form_q_launch(const cplx* A, const cplx* tau, cplx* Q, int D, long long strideA, int batch,
                             cudaStream_t st) {
    const size_t smem = size_t(D) * sizeof(cplx);
    qr_form_q_kernel<<<batch, kQrThreads, smem, st>>>(A, tau, Q, D, strideA);
    return cudaGetLastError();
}

cudaError_t qr_extract_dt_launch(const cplx* A, const int* perm, double* d, cplx* T, int D, long long strideA,
                                 int batch, cudaStream_t st) {
    dim3 grid((unsigned)((size_t(D) * D + 255) / 256), batch);
    qr_extract_dt_kernel<<<grid, 256, 0, st>>>(A, perm, d, T, D, strideA);
    return cudaGetLastError();
}

cudaError_t trsm_upper_launch(const cplx* A, cplx* Y, cplx* Zout, const int* perm, int D, long long strideA,
                              int batch, cudaStream_t st) {
    const size_t smem = size_t(D) * kTrsmCols * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(trsm_upper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((D + kTrsmCols - 1) / kTrsmCols, batch);
    trsm_upper_kernel<<<grid, kTrsmThreads, smem, st>>>(A, Y, Zout, perm, D, strideA);
    return cudaGetLastError();
}

cudaError_t scale_split_launch(const double* d, double* inv_big, double* small_, double* logacc, int D,
                               int batch, cudaStream_t st) {
    scale_split_kernel<<<batch, 256, 0, st>>>(d, inv_big, small_, logacc, D);
    return cudaGetLastError();
}

cudaError_t logdiag_accumulate_launch(const cplx* A, double* logacc, int D, long long strideA, int batch,
                                      cudaStream_t st) {
    logdiag_kernel<<<batch, 256, 0, st>>>(A, logacc, D, strideA);
    return cudaGetLastError();
}

}  // namespace dqmc
