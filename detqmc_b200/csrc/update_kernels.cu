// Delayed (Sherman-Morrison-Woodbury) local updates of one time slice, for a batch of replicas.
//
// Replaces DetSDW::updateInSlice / updateInSlice_delayed / updateInSliceThermalization with box
// proposals (detsdwopdim.cpp:2427-2489, 3021-3175, 3293-3375), proposeNewPhiBox (:3920-3931),
// deltaSPhi (:4185-4239) and get_delta_forsite (:3177-3289).
//
// One CTA owns one replica and walks the N sites of the slice in order (the Metropolis chain is
// strictly sequential).  Random numbers come from a window of the replica's host dSFMT stream that
// was uploaded before the launch; the kernel consumes them with a cursor in exactly the
// reference's order: OPDIM values per proposal, plus one more only if the acceptance probability is
// <= 1 (:3113).  The decision needs only the MSF x MSF site block of the effective Green's function
// (G + pending X*Y), which warp 0 gathers with shuffle reductions; the full rows / columns are
// gathered by the whole CTA only on acceptance.  Every `delaySteps` accepted updates (or at the end
// of the slice) the rank-(MSF*j) correction G += X*Y is flushed as a tensor-core GEMM (DMMA
// m8n8k4.f64) by the same CTA, straight out of L1/L2.  Reductions are fixed-order: the kernel is
// deterministic, which the 100-sweep trajectory parity requires.
#include "dqmc_internal.h"

namespace dqmc {
namespace {

constexpr int kUpdThreads = 512;

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    const double den = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// e^{sign*dtau*V} site block (evMatrix, detsdwopdim.cpp:3188-3229, cdwU == 0); x already carries
// sinh(.)/|phi|, sg = sign.
template <int MSF, int OPDIM>
__device__ __forceinline__ void ev_block(cplx* E, double sg, const double* p, double c, double x) {
    const double sx = sg * x;
    const double p0 = p[0];
    const double p1 = OPDIM > 1 ? p[1] : 0.0;
    const cplx e01 = make_double2(sx * p0, -sx * p1);
    const cplx e10 = make_double2(sx * p0, sx * p1);
    const cplx cc = make_double2(c, 0);
    if (MSF == 2) {
        E[0] = cc; E[1] = e01; E[2] = e10; E[3] = cc;
    } else {
        const double p2 = p[OPDIM > 2 ? 2 : 0];
        const cplx z = make_double2(0, 0);
        const cplx a = make_double2(sx * p2, 0), ma = make_double2(-sx * p2, 0);
        const cplx T[16] = {cc, e01, z, a, e10, cc, ma, z, z, ma, cc, e10, a, z, e01, cc};
#pragma unroll
        for (int i = 0; i < 16; ++i) E[i] = T[i];
    }
}

// det and inverse of a small complex matrix (Gauss-Jordan, partial pivoting)
template <int n>
__device__ __forceinline__ cplx small_det_inv(const cplx* Min, cplx* inv) {
    if (n == 2) {
        const cplx det = csub(cmul(Min[0], Min[3]), cmul(Min[1], Min[2]));
        const cplx one = make_double2(1, 0);
        const cplx id = cdiv(one, det);
        inv[0] = cmul(Min[3], id);
        inv[1] = cmul(make_double2(-Min[1].x, -Min[1].y), id);
        inv[2] = cmul(make_double2(-Min[2].x, -Min[2].y), id);
        inv[3] = cmul(Min[0], id);
        return det;
    }
    cplx a[n * n], b[n * n];
    for (int i = 0; i < n * n; ++i) { a[i] = Min[i]; b[i] = make_double2((i % (n + 1)) == 0 ? 1.0 : 0.0, 0.0); }
    cplx det = make_double2(1, 0);
    for (int col = 0; col < n; ++col) {
        int piv = col;
        double best = a[col * n + col].x * a[col * n + col].x + a[col * n + col].y * a[col * n + col].y;
        for (int r = col + 1; r < n; ++r) {
            const double v = a[r * n + col].x * a[r * n + col].x + a[r * n + col].y * a[r * n + col].y;
            if (v > best) { best = v; piv = r; }
        }
        if (piv != col) {
            for (int c = 0; c < n; ++c) {
                cplx t = a[col * n + c]; a[col * n + c] = a[piv * n + c]; a[piv * n + c] = t;
                t = b[col * n + c]; b[col * n + c] = b[piv * n + c]; b[piv * n + c] = t;
            }
            det = make_double2(-det.x, -det.y);
        }
        const cplx p = a[col * n + col];
        det = cmul(det, p);
        const cplx ip = cdiv(make_double2(1, 0), p);
        for (int c = 0; c < n; ++c) { a[col * n + c] = cmul(a[col * n + c], ip); b[col * n + c] = cmul(b[col * n + c], ip); }
        for (int r = 0; r < n; ++r) {
            if (r == col) continue;
            const cplx f = a[r * n + col];
            for (int c = 0; c < n; ++c) {
                a[r * n + c] = csub(a[r * n + c], cmul(f, a[col * n + c]));
                b[r * n + c] = csub(b[r * n + c], cmul(f, b[col * n + c]));
            }
        }
    }
    for (int i = 0; i < n * n; ++i) inv[i] = b[i];
    return det;
}

// G += X[:, 0:K] * Y[0:K, :] on the FP64 tensor cores; X is D x K (lda = D), Y is K x D (ldb = KMAX).
__device__ __forceinline__ void flush_delayed(cplx* G, const cplx* X, const cplx* Y, int D, int K, int KMAX) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int grp = lane >> 2, t4 = lane & 3;
    const int tm = (D + 31) / 32, tn = (D + 15) / 16;     // 32 x 16 output tile per warp
    for (int tile = warp; tile < tm * tn; tile += nwarps) {
        const int m0 = (tile % tm) * 32, n0 = (tile / tm) * 16;
        double acc_re[4][2][2], acc_im[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) { acc_re[i][j][0] = acc_re[i][j][1] = acc_im[i][j][0] = acc_im[i][j][1] = 0.0; }
        for (int k0 = 0; k0 < K; k0 += 4) {
            const int kk = k0 + t4;
            cplx af[4], bf[2];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                const int mrow = m0 + mb * 8 + grp;
                af[mb] = (mrow < D && kk < K) ? X[size_t(kk) * D + mrow] : make_double2(0, 0);
            }
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
                const int ncol = n0 + nb * 8 + grp;
                bf[nb] = (ncol < D && kk < K) ? Y[size_t(ncol) * KMAX + kk] : make_double2(0, 0);
            }
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    dmma(acc_re[mb][nb][0], acc_re[mb][nb][1], af[mb].x, bf[nb].x);
                    dmma(acc_re[mb][nb][0], acc_re[mb][nb][1], -af[mb].y, bf[nb].y);
                    dmma(acc_im[mb][nb][0], acc_im[mb][nb][1], af[mb].x, bf[nb].y);
                    dmma(acc_im[mb][nb][0], acc_im[mb][nb][1], af[mb].y, bf[nb].x);
                }
        }
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
            const int mrow = m0 + mb * 8 + grp;
            if (mrow >= D) continue;
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ncol = n0 + nb * 8 + 2 * t4 + e;
                    if (ncol >= D) continue;
                    cplx* dst = G + size_t(ncol) * D + mrow;
                    cplx g = *dst;
                    g.x += acc_re[mb][nb][e];
                    g.y += acc_im[mb][nb][e];
                    *dst = g;
                }
        }
    }
}

template <int MSF, int OPDIM>
__global__ void __launch_bounds__(kUpdThreads) update_slice_kernel(UpdateModel md, UpdateArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = md.D, N = md.N, L = md.L;
    const int KMAX = MSF * md.delaySteps;
    cplx* sXrow = reinterpret_cast<cplx*>(smem_raw);          // [MSF][KMAX]
    cplx* sYcol = sXrow + MSF * KMAX;                          // [KMAX][MSF]
    __shared__ cplx sDelta[MSF * MSF], sMinv[MSF * MSF];
    __shared__ int sAcceptBuf[2], sAbort;   // decision flag double-buffered by site parity

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k;
    cplx* G = a.G + size_t(b) * a.strideG;
    cplx* X = a.X + size_t(b) * a.strideXY;
    cplx* Y = a.Y + size_t(b) * a.strideXY;
    double* phi = a.phi + size_t(b) * a.stridePhi;
    double* coshT = a.coshT + size_t(b) * a.strideTab;
    double* sinhT = a.sinhT + size_t(b) * a.strideTab;
    const double* rng = a.rng + size_t(b) * a.strideRng;
    const double rpar = a.rvals[b];
    const double phiDelta = a.ctrl[b].phiDelta;
    const double dtau = md.dtau;

    int cursor = a.cursor[b];              // meaningful in thread 0 only
    unsigned accepted = 0;                 // thread 0
    int j = 0;                             // accepted updates pending in X, Y (uniform over the CTA)
    int delayNow = min(md.delaySteps, N);
    if (tid == 0) sAbort = 0;
    __syncthreads();

    double* phik = phi + size_t(k) * OPDIM * N;
    const int kEarlier = k > 1 ? k - 1 : md.m;
    const int kLater = k < md.m ? k + 1 : 1;

    for (int site = 0; site < N; ++site) {
        const int K = MSF * j;
        int& sAccept = sAcceptBuf[site & 1];
        // ------------------------------------------------------------------ phase A: decision
        if (warp == 0) {
            // site block of the effective Green's function: S = G[rows, cols] + X[rows, :K] Y[:K, cols]
            cplx S[MSF * MSF];
#pragma unroll
            for (int r = 0; r < MSF; ++r)
#pragma unroll
                for (int c = 0; c < MSF; ++c) {
                    double sr = 0, si = 0;
                    for (int l = lane; l < K; l += 32) {
                        const cplx x = X[size_t(l) * D + site + r * N];
                        const cplx y = Y[size_t(site + c * N) * KMAX + l];
                        sr += x.x * y.x - x.y * y.y;
                        si += x.x * y.y + x.y * y.x;
                    }
                    sr = warp_sum(sr);
                    si = warp_sum(si);
                    S[r * MSF + c] = make_double2(sr, si);
                }
            if (lane == 0) {
                if (cursor + OPDIM + 1 > a.rngWindow) {
                    sAbort = 1;
                    sAccept = 0;
                } else {
#pragma unroll
                    for (int r = 0; r < MSF; ++r)
#pragma unroll
                        for (int c = 0; c < MSF; ++c)
                            S[r * MSF + c] = cadd(S[r * MSF + c], G[size_t(site + c * N) * D + site + r * N]);
                    // proposeNewPhiBox: OPDIM draws, randRange(-phiDelta, +phiDelta)
                    double oldp[3], newp[3];
#pragma unroll
                    for (int d = 0; d < OPDIM; ++d) {
                        oldp[d] = phik[d * N + site];
                        const double u = rng[cursor + d];
                        newp[d] = oldp[d] + (-phiDelta + (phiDelta - (-phiDelta)) * u);
                    }
                    cursor += OPDIM;
                    // deltaSPhi
                    double oldSq = 0, newSq = 0, tdot = 0, sdot = 0;
                    const int x = site % L, y = site / L;
                    const int nb0 = y * L + (x + 1 == L ? 0 : x + 1);
                    const int nb1 = y * L + (x == 0 ? L - 1 : x - 1);
                    const int nb2 = (y + 1 == L ? 0 : y + 1) * L + x;
                    const int nb3 = (y == 0 ? L - 1 : y - 1) * L + x;
#pragma unroll
                    for (int d = 0; d < OPDIM; ++d) {
                        const double diff = newp[d] - oldp[d];
                        oldSq += oldp[d] * oldp[d];
                        newSq += newp[d] * newp[d];
                        const double tn = phi[(size_t(kLater) * OPDIM + d) * N + site] +
                                          phi[(size_t(kEarlier) * OPDIM + d) * N + site];
                        const double sn = ((phik[d * N + nb0] + phik[d * N + nb1]) + phik[d * N + nb2]) + phik[d * N + nb3];
                        tdot += tn * diff;
                        sdot += sn * diff;
                    }
                    const double sqDiff = newSq - oldSq;
                    const double pow4Diff = newSq * newSq - oldSq * oldSq;
                    const double d1 = (1.0 / (md.c * md.c * dtau)) * (sqDiff - tdot);
                    const double d2 = 0.5 * dtau * (4.0 * sqDiff - 2.0 * sdot);
                    const double d3 = dtau * (0.5 * rpar * sqDiff + 0.25 * md.u * pow4Diff);
                    const double probSPhi = exp(-(d1 + d2 + d3));
                    // get_delta_forsite: Delta = e^{-dtau V(new)} e^{+dtau V(old)} - 1
                    const double cOld = coshT[size_t(k) * N + site], xOld = sinhT[size_t(k) * N + site];
                    const double nrm = sqrt(newSq);
                    const double cNew = cosh(md.lambda * dtau * nrm);
                    const double xNew = sinh(md.lambda * dtau * nrm) / nrm;
                    cplx evOld[MSF * MSF], emvNew[MSF * MSF], Dl[MSF * MSF], M[MSF * MSF], Minv[MSF * MSF];
                    ev_block<MSF, OPDIM>(evOld, +1.0, oldp, cOld, xOld);
                    ev_block<MSF, OPDIM>(emvNew, -1.0, newp, cNew, xNew);
#pragma unroll
                    for (int r = 0; r < MSF; ++r)
#pragma unroll
                        for (int c = 0; c < MSF; ++c) {
                            cplx s = make_double2(r == c ? -1.0 : 0.0, 0.0);
#pragma unroll
                            for (int t = 0; t < MSF; ++t) s = cfma(emvNew[r * MSF + t], evOld[t * MSF + c], s);
                            Dl[r * MSF + c] = s;
                        }
                    // M = 1 - S Delta + Delta
#pragma unroll
                    for (int r = 0; r < MSF; ++r)
#pragma unroll
                        for (int c = 0; c < MSF; ++c) {
                            cplx s = make_double2(r == c ? 1.0 : 0.0, 0.0);
#pragma unroll
                            for (int t = 0; t < MSF; ++t) s = csub(s, cmul(S[r * MSF + t], Dl[t * MSF + c]));
                            M[r * MSF + c] = cadd(s, Dl[r * MSF + c]);
                        }
                    const cplx det = small_det_inv<MSF>(M, Minv);
                    const double probFermion = (OPDIM == 3) ? det.x : (det.x * det.x + det.y * det.y);
                    const double prob = probSPhi * probFermion;
                    bool acc;
                    if (prob > 1.0) {
                        acc = true;
                    } else {
                        const double u = rng[cursor];
                        cursor += 1;
                        acc = u < prob;
                    }
                    if (acc) {
                        accepted += 1;
#pragma unroll
                        for (int d = 0; d < OPDIM; ++d) phik[d * N + site] = newp[d];
                        coshT[size_t(k) * N + site] = cNew;
                        sinhT[size_t(k) * N + site] = xNew;
#pragma unroll
                        for (int i = 0; i < MSF * MSF; ++i) { sDelta[i] = Dl[i]; sMinv[i] = Minv[i]; }
                    }
                    sAccept = acc ? 1 : 0;
                }
            }
        }
        __syncthreads();
        if (sAbort) break;
        if (sAccept) {
            // -------------------------------------------------------------- phase B: extend X, Y
            for (int idx = tid; idx < MSF * K; idx += blockDim.x) {
                const int r = idx / K, l = idx - r * K;
                sXrow[r * KMAX + l] = X[size_t(l) * D + site + r * N];
                sYcol[l * MSF + r] = Y[size_t(site + r * N) * KMAX + l];
            }
            __syncthreads();
            for (int t = tid; t < D; t += blockDim.x) {
                // rows R_j[:, t] of the effective G, then Y_j = M^-1 (R_j - 1_j)
                cplx Rr[MSF];
#pragma unroll
                for (int r = 0; r < MSF; ++r) Rr[r] = G[size_t(t) * D + site + r * N];
                const cplx* ycol = Y + size_t(t) * KMAX;
                for (int l = 0; l < K; ++l) {
                    const cplx yv = ycol[l];
#pragma unroll
                    for (int r = 0; r < MSF; ++r) Rr[r] = cfma(sXrow[r * KMAX + l], yv, Rr[r]);
                }
#pragma unroll
                for (int r = 0; r < MSF; ++r)
                    if (t == site + r * N) Rr[r].x -= 1.0;
                // columns C_j[t, :] of the effective G, then X_j = C_j Delta
                cplx Cc[MSF];
#pragma unroll
                for (int c = 0; c < MSF; ++c) Cc[c] = G[size_t(site + c * N) * D + t];
                for (int l = 0; l < K; ++l) {
                    const cplx xv = X[size_t(l) * D + t];
#pragma unroll
                    for (int c = 0; c < MSF; ++c) Cc[c] = cfma(xv, sYcol[l * MSF + c], Cc[c]);
                }
#pragma unroll
                for (int r = 0; r < MSF; ++r) {
                    cplx yn = make_double2(0, 0), xn = make_double2(0, 0);
#pragma unroll
                    for (int q = 0; q < MSF; ++q) {
                        yn = cfma(sMinv[r * MSF + q], Rr[q], yn);
                        xn = cfma(Cc[q], sDelta[q * MSF + r], xn);
                    }
                    Y[size_t(t) * KMAX + K + r] = yn;
                    X[size_t(K + r) * D + t] = xn;
                }
            }
            __syncthreads();
            j += 1;
            if (j == delayNow) {
                flush_delayed(G, X, Y, D, MSF * j, KMAX);
                __syncthreads();
                j = 0;
                delayNow = min(md.delaySteps, N - (site + 1));
            }
        }
    }
    if (j > 0 && !sAbort) {
        flush_delayed(G, X, Y, D, MSF * j, KMAX);
    }
    if (tid == 0) {
        if (sAbort) atomicExch(a.errflag, 1);
        a.cursor[b] = cursor;
        a.accepted[b] = accepted;
        if (a.acceptedTotal) a.acceptedTotal[b] += accepted;
        dqmc_control_data* cd = a.ctrl + b;
        const double ratio = double(accepted) / double(N);
        cd->lastAccRatioLocal_phi = ratio;
        if (a.thermalization) {
            // RunningAverage::addValue (RunningAverage.h:57-68) on a ring buffer, then the step-size
            // adaptation of updateInSliceThermalization (detsdwopdim.cpp:3329-3341)
            const int pos = cd->ra_samples_added % 100;
            if (cd->ra_samples_added >= 100) cd->ra_average -= cd->ra_values[pos] / 100.0;
            cd->ra_values[pos] = ratio;
            cd->ra_average += ratio / 100.0;
            cd->ra_samples_added += 1;
            if (cd->ra_count < 100) cd->ra_count += 1;
            if (cd->ra_samples_added % 100 == 0) {
                if (cd->ra_average < md.accRatio) cd->phiDelta *= 0.95;
                else if (cd->ra_average > md.accRatio) cd->phiDelta *= 1.05;
            }
        }
    }
}

}  // namespace

cudaError_t update_slice_launch(const UpdateModel& m, const UpdateArgs& a, cudaStream_t st) {
    const int KMAX = m.msf * m.delaySteps;
    const size_t smem = size_t(2) * m.msf * KMAX * sizeof(cplx);
#define LAUNCH(MSF, OPD)                                                                                    \
    {                                                                                                       \
        cudaError_t e = cudaFuncSetAttribute(update_slice_kernel<MSF, OPD>,                                 \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        if (e != cudaSuccess) return e;                                                                     \
        update_slice_kernel<MSF, OPD><<<a.batch, kUpdThreads, smem, st>>>(m, a);                            \
    }
    if (m.opdim == 1) LAUNCH(2, 1)
    else if (m.opdim == 2) LAUNCH(2, 2)
    else LAUNCH(4, 3)
#undef LAUNCH
    return cudaGetLastError();
}

}  // namespace dqmc
