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

// G += X[:, 0:K] * Y[0:K, :] on the FP64 tensor cores; X is D x K (element (t, l) at X[l*D + t]) and Y is
// K x D stored the same way (element (l, t) at Y[l*D + t]).
// In-kernel flush of the small-delaySteps mode (16 x 8 output tile per warp keeps the register
// footprint small; the production path flushes with the rank-K update GEMM on all SMs instead).
__device__ __forceinline__ void flush_delayed(cplx* G, const cplx* X, const cplx* Y, int D, int K, int KMAX) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int grp = lane >> 2, t4 = lane & 3;
    const int tm = (D + 15) / 16, tn = (D + 7) / 8;
    for (int tile = warp; tile < tm * tn; tile += nwarps) {
        const int m0 = (tile % tm) * 16, n0 = (tile / tm) * 8;
        double acc_re[2][2], acc_im[2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i) { acc_re[i][0] = acc_re[i][1] = acc_im[i][0] = acc_im[i][1] = 0.0; }
        for (int k0 = 0; k0 < K; k0 += 4) {
            const int kk = k0 + t4;
            cplx af[2], bf;
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const int mrow = m0 + mb * 8 + grp;
                af[mb] = (mrow < D && kk < K) ? X[size_t(kk) * D + mrow] : make_double2(0, 0);
            }
            const int ncolb = n0 + grp;
            bf = (ncolb < D && kk < K) ? Y[size_t(kk) * D + ncolb] : make_double2(0, 0);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                dmma(acc_re[mb][0], acc_re[mb][1], af[mb].x, bf.x);
                dmma(acc_re[mb][0], acc_re[mb][1], -af[mb].y, bf.y);
                dmma(acc_im[mb][0], acc_im[mb][1], af[mb].x, bf.y);
                dmma(acc_im[mb][0], acc_im[mb][1], af[mb].y, bf.x);
            }
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const int mrow = m0 + mb * 8 + grp;
            if (mrow >= D) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ncol = n0 + 2 * t4 + e;
                if (ncol >= D) continue;
                cplx* dst = G + size_t(ncol) * D + mrow;
                cplx g = *dst;
                g.x += acc_re[mb][e];
                g.y += acc_im[mb][e];
                *dst = g;
            }
        }
    }
}

// Shared-memory carve-up of one CTA (doubles unless noted)
struct UpdSmem {
    double* phik;      // [OPDIM][N]  fields of this slice (kept current as proposals are accepted)
    double* tsum;      // [OPDIM][N]  phi(k+1) + phi(k-1)
    double* ck;        // [N] cosh table of this slice
    double* xk;        // [N] sinh table of this slice
    double* rng;       // [N*(OPDIM+1)] window of the replica's random numbers, starting at the cursor
    cplx* S;           // [MSF*MSF] site block of the effective Green's function (column c, row r at c*MSF+r)
    cplx* Delta;       // [MSF*MSF]
    cplx* Minv;        // [MSF*MSF]
    cplx* xrow;        // [2][MSF][KMAX] pending X rows `site + rN`  (double buffered: this site / next site)
    cplx* ycol;        // [2][MSF][KMAX] pending Y columns `site + cN`
};

constexpr int kDecWarps = 4;        // warp 0: exp(-dS) + decision, 1: cosh, 2: sinh, 3: stager for the next site
constexpr int kDecThreads = 32 * kDecWarps;

// One ROUND of updateInSlice_delayed for every replica: propose/decide site by site until
// `delaySteps` proposals have been accepted (or the slice ends), appending to X, Y.  With
// inline_flush the CTA also applies G += X Y itself and keeps going to the end of the slice
// (small delaySteps, e.g. Woodbury = 1); otherwise it records K = MSF * (#accepted) in kvec and the
// host launches the rank-K update on all SMs before the next round.
//
// The Metropolis chain is strictly sequential, so the kernel is organised around the latency of one
// site.  Warps 0-2 evaluate the three transcendental functions of the proposal in parallel (bosonic
// action difference, cosh, sinh); warp 3 stages the pending X rows / Y columns of the NEXT site into
// shared memory; the remaining warps (thread <-> matrix index t) gather row / column `site` of the
// effective Green's function G + X Y, with the G entries of the next site prefetched one site
// ahead (G is constant during a round) and the pending terms read in batches so that their loads
// overlap.  The site block S = G_eff[site rows, site cols] is a by-product of the row gather; the
// part of the decision that needs it follows after one barrier.
template <int MSF, int OPDIM, int TPT, int MAXT>
__global__ void __launch_bounds__(MAXT) update_round_kernel(UpdateModel md, UpdateArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = md.D, N = md.N, L = md.L;
    const int KMAX = MSF * md.delaySteps;
    UpdSmem sm;
    sm.phik = reinterpret_cast<double*>(smem_raw);
    sm.tsum = sm.phik + OPDIM * N;
    sm.ck = sm.tsum + OPDIM * N;
    sm.xk = sm.ck + N;
    sm.rng = sm.xk + N;
    sm.S = reinterpret_cast<cplx*>(sm.rng + ((N * (OPDIM + 1) + 1) & ~1));
    sm.Delta = sm.S + MSF * MSF;
    sm.Minv = sm.Delta + MSF * MSF;
    sm.xrow = sm.Minv + MSF * MSF;
    sm.ycol = sm.xrow + 2 * MSF * KMAX;
    __shared__ int sAccept, sAbort, sNload;
    __shared__ double sTrans[4];                           // probSPhi, cNew, xNew
    __shared__ double sNewp[3];

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gthreads = blockDim.x - kDecThreads;         // gather threads
    const int gt = tid - kDecThreads;
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

    const int site0 = a.round == 0 ? 0 : a.site_state[b];
    if (site0 >= N) {                                      // slice already finished in an earlier round
        if (tid == 0 && a.kvec) a.kvec[b] = 0;
        return;
    }
    const int cursor0 = a.cursor[b];
    double* phik_g = phi + size_t(k) * OPDIM * N;
    const int kEarlier = k > 1 ? k - 1 : md.m;
    const int kLater = k < md.m ? k + 1 : 1;
    {
        const double* pl = phi + size_t(kLater) * OPDIM * N;
        const double* pe = phi + size_t(kEarlier) * OPDIM * N;
        for (int i = tid; i < OPDIM * N; i += blockDim.x) {
            sm.phik[i] = phik_g[i];
            sm.tsum[i] = pl[i] + pe[i];
        }
        for (int i = tid; i < N; i += blockDim.x) {
            sm.ck[i] = coshT[size_t(k) * N + i];
            sm.xk[i] = sinhT[size_t(k) * N + i];
        }
        const int want = (N - site0) * (OPDIM + 1);
        const int have = max(0, min(want, a.rngWindow - cursor0));
        for (int i = tid; i < have; i += blockDim.x) sm.rng[i] = rng[cursor0 + i];
        if (tid == 0) { sAbort = 0; sNload = have; }
    }
    __syncthreads();

    int cur = 0;                                           // offset into sm.rng (tracked identically by warps 0-2)
    unsigned accepted = 0;                                 // thread 0
    int j = 0;                                             // accepted updates pending in X, Y (uniform)
    int delayNow = min(md.delaySteps, N - site0);
    int site = site0;

    // G entries of the first site (later sites are prefetched one iteration ahead)
    cplx Gr[TPT][MSF], Gc[TPT][MSF];
    if (warp >= kDecWarps) {
#pragma unroll
        for (int q = 0; q < TPT; ++q) {
            const int t = gt + q * gthreads;
            if (t < D) {
#pragma unroll
                for (int r = 0; r < MSF; ++r) {
                    Gr[q][r] = G[size_t(t) * D + site + r * N];
                    Gc[q][r] = G[size_t(site + r * N) * D + t];
                }
            }
        }
    }
    // proposal state carried by thread 0 across the barrier
    double oldp[3] = {0, 0, 0}, newp[3] = {0, 0, 0};
    bool have_rng = true;

    for (; site < N; ++site) {
        const int K = MSF * j;
        const int buf = site & 1;
        const cplx* xrow = sm.xrow + buf * MSF * KMAX;
        const cplx* ycol = sm.ycol + buf * MSF * KMAX;
        cplx Rr[TPT][MSF], Cc[TPT][MSF];
        if (warp < 3) {
            // ---------------------------------------------- proposal (independent of G), lane 0 of warps 0-2
            if (lane == 0) {
                have_rng = cur + OPDIM + 1 <= sNload;
                if (have_rng) {
                    double newSq = 0;
#pragma unroll
                    for (int d = 0; d < OPDIM; ++d) {
                        oldp[d] = sm.phik[d * N + site];
                        const double u = sm.rng[cur + d];
                        newp[d] = oldp[d] + (-phiDelta + (phiDelta - (-phiDelta)) * u);   // randRange(-delta, +delta)
                        newSq += newp[d] * newp[d];
                    }
                    if (warp == 0) {
                        // deltaSPhi (detsdwopdim.cpp:4185-4239)
                        double oldSq = 0, tdot = 0, sdot = 0;
                        const int x = site % L, y = site / L;
                        const int nb0 = y * L + (x + 1 == L ? 0 : x + 1);
                        const int nb1 = y * L + (x == 0 ? L - 1 : x - 1);
                        const int nb2 = (y + 1 == L ? 0 : y + 1) * L + x;
                        const int nb3 = (y == 0 ? L - 1 : y - 1) * L + x;
#pragma unroll
                        for (int d = 0; d < OPDIM; ++d) {
                            const double diff = newp[d] - oldp[d];
                            oldSq += oldp[d] * oldp[d];
                            const double* pk = sm.phik + d * N;
                            const double sn = ((pk[nb0] + pk[nb1]) + pk[nb2]) + pk[nb3];
                            tdot += sm.tsum[d * N + site] * diff;
                            sdot += sn * diff;
                        }
                        const double sqDiff = newSq - oldSq;
                        const double pow4Diff = newSq * newSq - oldSq * oldSq;
                        const double d1 = (1.0 / (md.c * md.c * dtau)) * (sqDiff - tdot);
                        const double d2 = 0.5 * dtau * (4.0 * sqDiff - 2.0 * sdot);
                        const double d3 = dtau * (0.5 * rpar * sqDiff + 0.25 * md.u * pow4Diff);
                        sTrans[0] = exp(-(d1 + d2 + d3));
                    } else if (warp == 1) {
                        sTrans[1] = cosh(md.lambda * dtau * sqrt(newSq));
                    } else {
                        const double nrm = sqrt(newSq);
                        sTrans[2] = sinh(md.lambda * dtau * nrm) / nrm;
                    }
                }
            }
        } else if (warp == 3) {
            // ---------------------------------------------- stage the pending X rows / Y columns of the NEXT site
            if (site + 1 < N) {
                cplx* xn = sm.xrow + (buf ^ 1) * MSF * KMAX;
                cplx* yn = sm.ycol + (buf ^ 1) * MSF * KMAX;
                for (int i = lane; i < MSF * K; i += 32) {
                    const int r = i / K, l = i - r * K;
                    xn[r * KMAX + l] = X[size_t(l) * D + site + 1 + r * N];
                    yn[r * KMAX + l] = Y[size_t(l) * D + site + 1 + r * N];
                }
            }
        } else {
            // ---------------------------------------------- gather rows / columns `site` of G + X Y
#pragma unroll
            for (int q = 0; q < TPT; ++q) {
                const int t = gt + q * gthreads;
                if (t < D) {
#pragma unroll
                    for (int r = 0; r < MSF; ++r) { Rr[q][r] = Gr[q][r]; Cc[q][r] = Gc[q][r]; }
                    if (site + 1 < N) {
                        // prefetch for the next site; consumed one iteration later
#pragma unroll
                        for (int r = 0; r < MSF; ++r) {
                            Gr[q][r] = G[size_t(t) * D + site + 1 + r * N];
                            Gc[q][r] = G[size_t(site + 1 + r * N) * D + t];
                        }
                    }
                    constexpr int CH = 4;
                    for (int l0 = 0; l0 < K; l0 += CH) {
                        cplx yv[CH], xv[CH];
#pragma unroll
                        for (int u = 0; u < CH; ++u) {
                            const int l = l0 + u;
                            if (l < K) {
                                yv[u] = Y[size_t(l) * D + t];
                                xv[u] = X[size_t(l) * D + t];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < CH; ++u) {
                            const int l = l0 + u;
                            if (l < K) {
#pragma unroll
                                for (int r = 0; r < MSF; ++r) {
                                    Rr[q][r] = cfma(xrow[r * KMAX + l], yv[u], Rr[q][r]);
                                    Cc[q][r] = cfma(xv[u], ycol[r * KMAX + l], Cc[q][r]);
                                }
                            }
                        }
                    }
                    // the site block is rows `site + rN` of the columns `site + cN`
                    const int rel = t - site;
                    if (rel >= 0 && rel % N == 0) {
                        const int c = rel / N;
                        if (c < MSF) {
#pragma unroll
                            for (int r = 0; r < MSF; ++r) sm.S[c * MSF + r] = Rr[q][r];
                        }
                    }
                }
            }
        }
        __syncthreads();
        // -------------------------------------------------- decision, thread 0
        bool consumed_extra = false;                       // set identically in lane 0 of warps 0-2 below
        if (tid == 0) {
            if (!have_rng) {
                sAbort = 1;
                sAccept = 0;
            } else {
                const double probSPhi = sTrans[0], cNew = sTrans[1], xNew = sTrans[2];
                // get_delta_forsite: Delta = e^{-dtau V(new)} e^{+dtau V(old)} - 1
                cplx evOld[MSF * MSF], emvNew[MSF * MSF], Dl[MSF * MSF];
                ev_block<MSF, OPDIM>(evOld, +1.0, oldp, sm.ck[site], sm.xk[site]);
                ev_block<MSF, OPDIM>(emvNew, -1.0, newp, cNew, xNew);
#pragma unroll
                for (int r = 0; r < MSF; ++r)
#pragma unroll
                    for (int c = 0; c < MSF; ++c) {
                        cplx sacc = make_double2(r == c ? -1.0 : 0.0, 0.0);
#pragma unroll
                        for (int t = 0; t < MSF; ++t) sacc = cfma(emvNew[r * MSF + t], evOld[t * MSF + c], sacc);
                        Dl[r * MSF + c] = sacc;
                    }
                cplx S[MSF * MSF], M[MSF * MSF], Minv[MSF * MSF];
#pragma unroll
                for (int r = 0; r < MSF; ++r)
#pragma unroll
                    for (int c = 0; c < MSF; ++c) S[r * MSF + c] = sm.S[c * MSF + r];
                // M = 1 - S Delta + Delta
#pragma unroll
                for (int r = 0; r < MSF; ++r)
#pragma unroll
                    for (int c = 0; c < MSF; ++c) {
                        cplx sacc = make_double2(r == c ? 1.0 : 0.0, 0.0);
#pragma unroll
                        for (int t = 0; t < MSF; ++t) sacc = csub(sacc, cmul(S[r * MSF + t], Dl[t * MSF + c]));
                        M[r * MSF + c] = cadd(sacc, Dl[r * MSF + c]);
                    }
                const cplx det = small_det_inv<MSF>(M, Minv);
                const double probFermion = (OPDIM == 3) ? det.x : (det.x * det.x + det.y * det.y);
                const double prob = probSPhi * probFermion;
                bool acc;
                int used = OPDIM;
                if (prob > 1.0) {
                    acc = true;
                } else {
                    acc = sm.rng[cur + OPDIM] < prob;
                    used += 1;
                }
                if (acc) {
                    accepted += 1;
#pragma unroll
                    for (int d = 0; d < OPDIM; ++d) {
                        sm.phik[d * N + site] = newp[d];
                        phik_g[d * N + site] = newp[d];
                    }
                    sm.ck[site] = cNew;
                    sm.xk[site] = xNew;
                    coshT[size_t(k) * N + site] = cNew;
                    sinhT[size_t(k) * N + site] = xNew;
#pragma unroll
                    for (int i = 0; i < MSF * MSF; ++i) { sm.Delta[i] = Dl[i]; sm.Minv[i] = Minv[i]; }
                }
                sAccept = acc ? (used > OPDIM ? 3 : 1) : (used > OPDIM ? 2 : 0);    // bit 0: accepted, bit 1: extra draw
            }
        }
        __syncthreads();
        if (sAbort) break;
        const int decision = sAccept;
        consumed_extra = (decision & 2) != 0;
        cur += OPDIM + (consumed_extra ? 1 : 0);           // every thread tracks the cursor (warps 0-2 use it)
        if (decision & 1) {
            // ---------------------------------------------- extend X, Y:  X_j = C_j Delta,  Y_j = M^-1 (R_j - 1_j)
            if (warp >= kDecWarps) {
#pragma unroll
                for (int q = 0; q < TPT; ++q) {
                    const int t = gt + q * gthreads;
                    if (t < D) {
#pragma unroll
                        for (int r = 0; r < MSF; ++r)
                            if (t == site + r * N) Rr[q][r].x -= 1.0;
                        cplx ynew[MSF], xnew[MSF];
#pragma unroll
                        for (int r = 0; r < MSF; ++r) {
                            cplx yn = make_double2(0, 0), xn = make_double2(0, 0);
#pragma unroll
                            for (int qq = 0; qq < MSF; ++qq) {
                                yn = cfma(sm.Minv[r * MSF + qq], Rr[q][qq], yn);
                                xn = cfma(Cc[q][qq], sm.Delta[qq * MSF + r], xn);
                            }
                            ynew[r] = yn; xnew[r] = xn;
                            Y[size_t(K + r) * D + t] = yn;
                            X[size_t(K + r) * D + t] = xn;
                        }
                        // the new entries of the next site's staged rows / columns
                        const int rel = t - (site + 1);
                        if (site + 1 < N && rel >= 0 && rel % N == 0 && rel / N < MSF) {
                            const int rr = rel / N;
                            cplx* xn2 = sm.xrow + (buf ^ 1) * MSF * KMAX + rr * KMAX;
                            cplx* yn2 = sm.ycol + (buf ^ 1) * MSF * KMAX + rr * KMAX;
#pragma unroll
                            for (int r = 0; r < MSF; ++r) { xn2[K + r] = xnew[r]; yn2[K + r] = ynew[r]; }
                        }
                    }
                }
            }
            j += 1;
            if (j == delayNow) {
                if (!a.inline_flush) { ++site; break; }
                __syncthreads();
                flush_delayed(G, X, Y, D, MSF * j, KMAX);
                j = 0;
                delayNow = min(md.delaySteps, N - (site + 1));
                __syncthreads();
                // G changed: refresh the prefetched entries of the next site
                if (warp >= kDecWarps && site + 1 < N) {
#pragma unroll
                    for (int q = 0; q < TPT; ++q) {
                        const int t = gt + q * gthreads;
                        if (t < D) {
#pragma unroll
                            for (int r = 0; r < MSF; ++r) {
                                Gr[q][r] = G[size_t(t) * D + site + 1 + r * N];
                                Gc[q][r] = G[size_t(site + 1 + r * N) * D + t];
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (a.inline_flush && j > 0 && !sAbort) {
        flush_delayed(G, X, Y, D, MSF * j, KMAX);
        j = 0;
    }
    if (tid == 0) {
        if (sAbort) atomicExch(a.errflag, 1);
        a.cursor[b] = cursor0 + cur;
        if (a.kvec) a.kvec[b] = sAbort ? 0 : MSF * j;
        a.site_state[b] = sAbort ? N : site;
        const unsigned total = (a.round == 0 ? 0u : a.accepted[b]) + accepted;
        a.accepted[b] = total;
        if (a.acceptedTotal) a.acceptedTotal[b] += accepted;
        if (site >= N && !sAbort) {
            // end of the slice: acceptance statistics and step-size adaptation
            dqmc_control_data* cd = a.ctrl + b;
            const double ratio = double(total) / double(N);
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
}

}  // namespace

int update_rounds_per_slice(const UpdateModel& m, int inline_flush) {
    return inline_flush ? 1 : (m.N + m.delaySteps - 1) / m.delaySteps;
}

cudaError_t update_round_launch(const UpdateModel& m, const UpdateArgs& a, cudaStream_t st) {
    const int tpt = m.D > 896 ? 2 : 1;
    if (m.D > 2 * 896) return cudaErrorInvalidValue;
    const int gth = ((((m.D + tpt - 1) / tpt) + 31) / 32) * 32;
    const int threads = kDecThreads + gth;
    const size_t smem = size_t((2 * m.opdim + 2) * m.N + ((m.N * (m.opdim + 1) + 1) & ~1)) * sizeof(double) +
                        size_t(3) * m.msf * m.msf * sizeof(cplx) + size_t(4) * m.msf * m.msf * m.delaySteps * sizeof(cplx);
#define LAUNCH(MSF, OPD, TPT, MAXT)                                                                         \
    {                                                                                                       \
        cudaError_t e = cudaFuncSetAttribute(update_round_kernel<MSF, OPD, TPT, MAXT>,                      \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        if (e != cudaSuccess) return e;                                                                     \
        update_round_kernel<MSF, OPD, TPT, MAXT><<<a.batch, threads, smem, st>>>(m, a);                     \
    }
#define LAUNCH3(MSF, OPD)                                                                                   \
    {                                                                                                       \
        if (threads <= 416) LAUNCH(MSF, OPD, 1, 416)                                                        \
        else if (tpt == 1) LAUNCH(MSF, OPD, 1, 1024)                                                        \
        else LAUNCH(MSF, OPD, 2, 1024)                                                                      \
    }
    if (m.opdim == 1) LAUNCH3(2, 1)
    else if (m.opdim == 2) LAUNCH3(2, 2)
    else LAUNCH3(4, 3)
#undef LAUNCH3
#undef LAUNCH
    return cudaGetLastError();
}

}  // namespace dqmc
