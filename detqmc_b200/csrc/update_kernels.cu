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
// (G + pending X*Y); the full rows / columns are gathered by the whole CTA only on acceptance.  After
// `delaySteps` accepted updates (or at the end of the slice) a ROUND ends and the host launches the rank-K
// correction G += X*Y as a batched tensor-core GEMM (DMMA m8n8k4.f64) on all SMs (context.cu: launch_update);
// only small delay blocks (< 8, e.g. Woodbury = 1) are flushed by the CTA itself.  Reductions are fixed-order:
// the kernel is deterministic, which the 100-sweep trajectory parity requires.
#include "dqmc_internal.h"

#include <cstdlib>

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
        // 1 / det with one division (conj(det) / |det|^2)
        const double rden = 1.0 / (det.x * det.x + det.y * det.y);
        const cplx id = make_double2(det.x * rden, -det.y * rden);
        inv[0] = cmul(Min[3], id);
        inv[1] = cmul(make_double2(-Min[1].x, -Min[1].y), id);
        inv[2] = cmul(make_double2(-Min[2].x, -Min[2].y), id);
        inv[3] = cmul(Min[0], id);
        return det;
    }
    cplx a[n * n], b[n * n];
#pragma unroll
    for (int i = 0; i < n * n; ++i) { a[i] = Min[i]; b[i] = make_double2((i % (n + 1)) == 0 ? 1.0 : 0.0, 0.0); }
    cplx det = make_double2(1, 0);
#pragma unroll
    for (int col = 0; col < n; ++col) {
        // partial pivoting with compile-time indices only (a run-time pivot row would push a and b into local memory):
        // every later row that beats the current candidate is swapped into place.  The row that ends up at `col` is the
        // first one of maximal modulus, as with a single swap; the rows below end up in a different order, which the
        // elimination does not see, and the sign of det follows the parity of the complete permutation either way.
#pragma unroll
        for (int r = col + 1; r < n; ++r) {
            const double best = a[col * n + col].x * a[col * n + col].x + a[col * n + col].y * a[col * n + col].y;
            const double v = a[r * n + col].x * a[r * n + col].x + a[r * n + col].y * a[r * n + col].y;
            if (v > best) {
#pragma unroll
                for (int c = 0; c < n; ++c) {
                    cplx t = a[col * n + c]; a[col * n + c] = a[r * n + c]; a[r * n + c] = t;
                    t = b[col * n + c]; b[col * n + c] = b[r * n + c]; b[r * n + c] = t;
                }
                det = make_double2(-det.x, -det.y);
            }
        }
        const cplx p = a[col * n + col];
        det = cmul(det, p);
        const cplx ip = cdiv(make_double2(1, 0), p);
#pragma unroll
        for (int c = 0; c < n; ++c) { a[col * n + c] = cmul(a[col * n + c], ip); b[col * n + c] = cmul(b[col * n + c], ip); }
#pragma unroll
        for (int r = 0; r < n; ++r) {
            if (r == col) continue;
            const cplx f = a[r * n + col];
#pragma unroll
            for (int c = 0; c < n; ++c) {
                a[r * n + c] = csub(a[r * n + c], cmul(f, a[col * n + c]));
                b[r * n + c] = csub(b[r * n + c], cmul(f, b[col * n + c]));
            }
        }
    }
#pragma unroll
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
    cplx* Gblk;        // [2][MSF*MSF] site block of G (row r, column c at r*MSF+c), this site / next site
    cplx* Spart;       // [MSF*MSF] site block of the pending terms already in X, Y (computed by the stager)
    cplx* Delta;       // [MSF*MSF]
    cplx* Minv;        // [MSF*MSF]
    cplx* xrow;        // [2][MSF][KMAX] pending X rows `site + rN`  (double buffered: this site / next site)
    cplx* ycol;        // [2][MSF][KMAX] pending Y columns `site + cN`
    cplx* Ys;          // [KMAX][D] shared-memory copy of the pending Y rows (nullptr when it does not fit)
};

constexpr int kDecWarps = 2;        // warp 0: proposal + decision, warp 1: stager for the pending rows / columns
constexpr int kDecThreads = 32 * kDecWarps;

// cosh(x) and sinh(x)/x for the small arguments of the model (x = lambda*dtau*|phi|): Horner
// evaluation of the Taylor series in x^2 (error < 1 ulp for x <= 1); library functions otherwise.
__device__ __forceinline__ void cosh_sinhc(double x, double& c, double& sc) {
    if (x <= 1.0) {
        const double z = x * x;
        double pc = 1.0 / 2432902008176640000.0;       // 1/20!
        double ps = 1.0 / 51090942171709440000.0;      // 1/21!
        pc = fma(pc, z, 1.0 / 6402373705728000.0);     // 1/18!
        ps = fma(ps, z, 1.0 / 121645100408832000.0);   // 1/19!
        pc = fma(pc, z, 1.0 / 20922789888000.0);       // 1/16!
        ps = fma(ps, z, 1.0 / 355687428096000.0);      // 1/17!
        pc = fma(pc, z, 1.0 / 87178291200.0);          // 1/14!
        ps = fma(ps, z, 1.0 / 1307674368000.0);        // 1/15!
        pc = fma(pc, z, 1.0 / 479001600.0);            // 1/12!
        ps = fma(ps, z, 1.0 / 6227020800.0);           // 1/13!
        pc = fma(pc, z, 1.0 / 3628800.0);              // 1/10!
        ps = fma(ps, z, 1.0 / 39916800.0);             // 1/11!
        pc = fma(pc, z, 1.0 / 40320.0);                // 1/8!
        ps = fma(ps, z, 1.0 / 362880.0);               // 1/9!
        pc = fma(pc, z, 1.0 / 720.0);                  // 1/6!
        ps = fma(ps, z, 1.0 / 5040.0);                 // 1/7!
        pc = fma(pc, z, 1.0 / 24.0);
        ps = fma(ps, z, 1.0 / 120.0);
        pc = fma(pc, z, 0.5);
        ps = fma(ps, z, 1.0 / 6.0);
        c = fma(pc, z, 1.0);
        sc = fma(ps, z, 1.0);
    } else {
        c = cosh(x);
        sc = sinh(x) / x;
    }
}

__device__ __forceinline__ cplx lds_cplx(uint32_t addr) {
    cplx v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// row / column `site` entries of G for matrix index t (row access is strided in the column-major G)
template <int MSF>
__device__ __forceinline__ void load_g_rowcol(const cplx* __restrict__ G, int D, int N, int site, int t, cplx* Gr, cplx* Gc) {
#pragma unroll
    for (int r = 0; r < MSF; ++r) {
        Gr[r] = G[size_t(t) * D + site + r * N];
        Gc[r] = G[size_t(site + r * N) * D + t];
    }
}

// Gather row / column `site` of the effective Green's function G + X Y (K pending terms) for matrix
// index t and turn them into the new column of X and row of Y:
//   X_j = C_j Delta,  Y_j = M^-1 (R_j - 1_j)      (detsdwopdim.cpp:3123-3138)
template <int MSF>
__device__ __forceinline__ void extend_xy(cplx* __restrict__ X, cplx* __restrict__ Y, const cplx* __restrict__ G,
                                          int D, int N, int KMAX, int K, int site, int t, const cplx* xrow,
                                          const cplx* ycol, const cplx* sDelta, const cplx* sMinv, cplx* xnext,
                                          cplx* ynext, bool have_next, const cplx* Gr, const cplx* Gc, cplx* Ys) {
    cplx Rr[MSF], Cc[MSF];
#pragma unroll
    for (int r = 0; r < MSF; ++r) { Rr[r] = make_double2(0, 0); Cc[r] = make_double2(0, 0); }
    // Pending terms in chunks of 4 without per-term predicates: the staged rows / columns are zero
    // beyond K (kept so by the kernel) and X, Y hold finite values in all KMAX columns, so the padded
    // terms contribute exactly zero.  Shared memory is addressed through 32-bit shared addresses.
    constexpr int CH = 4;
    const uint32_t xs = (uint32_t)__cvta_generic_to_shared(xrow);
    const uint32_t ys = (uint32_t)__cvta_generic_to_shared(ycol);
    const cplx* __restrict__ xp = X + t;
    // pending Y rows of this thread: from the shared-memory copy when the CTA keeps one (Ys), else from global
    const cplx* yp = (Ys ? Ys : Y) + t;
    for (int l0 = 0; l0 < K; l0 += CH) {
        cplx yv[CH], xv[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            yv[u] = yp[size_t(l0 + u) * D];
            xv[u] = xp[size_t(l0 + u) * D];
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
#pragma unroll
            for (int r = 0; r < MSF; ++r) {
                const cplx xr = lds_cplx(xs + uint32_t((r * KMAX + l0 + u) * sizeof(cplx)));
                const cplx yc = lds_cplx(ys + uint32_t((r * KMAX + l0 + u) * sizeof(cplx)));
                Rr[r] = cfma(xr, yv[u], Rr[r]);
                Cc[r] = cfma(xv[u], yc, Cc[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < MSF; ++r) {
        Rr[r] = cadd(Rr[r], Gr[r]);
        Cc[r] = cadd(Cc[r], Gc[r]);
        if (t == site + r * N) Rr[r].x -= 1.0;
    }
    cplx ynew[MSF], xnew[MSF];
#pragma unroll
    for (int r = 0; r < MSF; ++r) {
        cplx yn = make_double2(0, 0), xn = make_double2(0, 0);
#pragma unroll
        for (int q = 0; q < MSF; ++q) {
            yn = cfma(sMinv[r * MSF + q], Rr[q], yn);
            xn = cfma(Cc[q], sDelta[q * MSF + r], xn);
        }
        ynew[r] = yn; xnew[r] = xn;
        Y[size_t(K + r) * D + t] = yn;
        if (Ys) Ys[size_t(K + r) * D + t] = yn;
        X[size_t(K + r) * D + t] = xn;
    }
    // new entries of the staged rows / columns of the next site
    const int rel = t - (site + 1);
    if (have_next && rel >= 0 && rel % N == 0 && rel / N < MSF) {
        const int rr = rel / N;
#pragma unroll
        for (int r = 0; r < MSF; ++r) { xnext[rr * KMAX + K + r] = xnew[r]; ynext[rr * KMAX + K + r] = ynew[r]; }
    }
}

// One ROUND of updateInSlice_delayed for every replica: propose/decide site by site until
// `delaySteps` proposals have been accepted (or the slice ends), appending to X, Y.  With
// inline_flush the CTA also applies G += X Y itself and keeps going to the end of the slice
// (small delaySteps, e.g. Woodbury = 1); otherwise it records K = MSF * (#accepted) in kvec and the
// host launches the rank-K update on all SMs before the next round.
//
// The Metropolis chain is strictly sequential, so the kernel is organised around the latency of one
// site and software-pipelined over two barriers per site:
//   phase 1   warp 0   proposal for `site`: new field, bosonic action difference, cosh / sinh, Delta
//             warp 1   stages the pending X rows / Y columns of `site` in shared memory and prefetches
//                      the G site block of the next site
//             others   (thread <-> matrix index t) finish the PREVIOUS site if it was accepted: gather
//                      its row / column of G + X Y and append X_j, Y_j
//   phase 2   warp 0   site block S of the effective Green's function (lanes over the pending terms),
//                      determinant ratio, Metropolis decision
// so the O(K D) work of an accepted update overlaps with the proposal of the next site, and only the
// O(K) site block sits between the two barriers.
template <int MSF, int OPDIM, int TPT, int MAXT>
__global__ void __launch_bounds__(MAXT) update_round_kernel(UpdateModel md, UpdateArgs a) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = md.D, N = md.N, L = md.L;
    const int KMAX = (MSF * md.delaySteps + 3) & ~3;
    UpdSmem sm;
    sm.phik = reinterpret_cast<double*>(smem_raw);
    sm.tsum = sm.phik + OPDIM * N;
    sm.ck = sm.tsum + OPDIM * N;
    sm.xk = sm.ck + N;
    sm.rng = sm.xk + N;
    sm.Gblk = reinterpret_cast<cplx*>(sm.rng + ((N * (OPDIM + 1) + 1) & ~1));
    sm.Spart = sm.Gblk + 2 * MSF * MSF;
    sm.Delta = sm.Spart + MSF * MSF;
    sm.Minv = sm.Delta + MSF * MSF;
    sm.xrow = sm.Minv + MSF * MSF;
    sm.ycol = sm.xrow + 2 * MSF * KMAX;
    sm.Ys = a.y_in_smem ? sm.ycol + 2 * MSF * KMAX : nullptr;
    __shared__ int sAccept, sAbort, sNload;

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
        if (tid < MSF * MSF) {
            const int r = tid / MSF, c = tid % MSF;
            sm.Gblk[(site0 & 1) * MSF * MSF + tid] = G[size_t(site0 + c * N) * D + site0 + r * N];
        }
        for (int i = tid; i < 2 * MSF * KMAX; i += blockDim.x) {
            sm.xrow[i] = make_double2(0, 0);
            sm.ycol[i] = make_double2(0, 0);
        }
        if (sm.Ys)
            for (int i = tid; i < KMAX * D; i += blockDim.x) sm.Ys[i] = make_double2(0, 0);   // finite beyond K
        if (tid == 0) { sAbort = 0; sNload = have; sAccept = 0; }
    }
    __syncthreads();

    int cur = 0;                                           // lane 0 of warp 0: offset into sm.rng
    unsigned accepted = 0;                                 // lane 0 of warp 0
    int j = 0;                                             // accepted updates whose X, Y columns exist or are in flight
    int delayNow = min(md.delaySteps, N - site0);
    int site = site0;
    int sx = site0 % L, sy = site0 / L;                    // coordinates of `site`
    bool prev_acc = false;                                 // previous site accepted, its X_j / Y_j still to be appended
    cplx Gr[TPT][MSF], Gc[TPT][MSF];                       // gather threads: row / column entries of G for the last decided site

    // proposal state carried by lane 0 of warp 0 across the barrier
    double newp[3] = {0, 0, 0}, cNew = 0, xNew = 0, probSPhi = 0;
    cplx Dl[MSF * MSF];
    bool have_rng = true;
    // per-phase clock counts of replica 0: development builds only (-DDQMC_UPD_TIMING); the counters cost 14 registers
    // in a kernel that sits at the 168-register cap of an 11-warp CTA
#ifdef DQMC_UPD_TIMING
    long long tq[6] = {0, 0, 0, 0, 0, 0};
    long long tmark = clock64();
#define TICK(i) if (a.debug) { long long now__; asm volatile("mov.u64 %0, %%clock64;" : "=l"(now__) :: "memory"); tq[i] += now__ - tmark; tmark = now__; }
#else
#define TICK(i)
#endif

    for (; site < N; ++site, sx = (sx + 1 == L ? 0 : sx + 1), sy += (sx == 0 ? 1 : 0)) {
        // K_done: pending terms already in X, Y;  the update of site-1 (if accepted) adds MSF more in phase 1
        const int K_done = MSF * (j - (prev_acc ? 1 : 0));
        const int buf = site & 1;
        cplx* xrow = sm.xrow + buf * MSF * KMAX;
        cplx* ycol = sm.ycol + buf * MSF * KMAX;
        TICK(5)
        // ================================================================== phase 1
        if (warp == 0) {
            if (lane == 0) {
                // ---------------------------------------------- proposal (independent of G)
                have_rng = cur + OPDIM + 1 <= sNload;
                if (have_rng) {
                    double oldp[3] = {0, 0, 0};
                    double oldSq = 0, newSq = 0, tdot = 0, sdot = 0;
                    const int x = sx, y = sy;              // site % L, site / L, tracked incrementally
                    const int nb0 = y * L + (x + 1 == L ? 0 : x + 1);
                    const int nb1 = y * L + (x == 0 ? L - 1 : x - 1);
                    const int nb2 = (y + 1 == L ? 0 : y + 1) * L + x;
                    const int nb3 = (y == 0 ? L - 1 : y - 1) * L + x;
#pragma unroll
                    for (int d = 0; d < OPDIM; ++d) {
                        oldp[d] = sm.phik[d * N + site];
                        const double u = sm.rng[cur + d];
                        newp[d] = oldp[d] + (-phiDelta + (phiDelta - (-phiDelta)) * u);   // randRange(-delta, +delta)
                        // deltaSPhi (detsdwopdim.cpp:4185-4239)
                        const double diff = newp[d] - oldp[d];
                        oldSq += oldp[d] * oldp[d];
                        newSq += newp[d] * newp[d];
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
                    probSPhi = exp(-(d1 + d2 + d3));
                    // get_delta_forsite: Delta = e^{-dtau V(new)} e^{+dtau V(old)} - 1
                    const double nrm = sqrt(newSq);
                    double sc;
                    cosh_sinhc(md.lambda * dtau * nrm, cNew, sc);
                    xNew = md.lambda * dtau * sc;                   // sinh(lambda dtau |phi|) / |phi|
                    cplx evOld[MSF * MSF], emvNew[MSF * MSF];
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
                }
            }
        } else if (warp == 1) {
            // ---------------------------------------------- stage the pending rows / columns of `site`
            cplx xs[MSF], ys[MSF];
            cplx part[MSF * MSF];
#pragma unroll
            for (int i = 0; i < MSF * MSF; ++i) part[i] = make_double2(0, 0);
            for (int l = lane; l < K_done; l += 32) {
#pragma unroll
                for (int r = 0; r < MSF; ++r) {
                    xs[r] = X[size_t(l) * D + site + r * N];
                    ys[r] = (sm.Ys ? sm.Ys : Y)[size_t(l) * D + site + r * N];
                    xrow[r * KMAX + l] = xs[r];
                    ycol[r * KMAX + l] = ys[r];
                }
#pragma unroll
                for (int r = 0; r < MSF; ++r)
#pragma unroll
                    for (int c = 0; c < MSF; ++c) part[r * MSF + c] = cfma(xs[r], ys[c], part[r * MSF + c]);
            }
            // site block of the terms already in X, Y (fixed-order butterfly: deterministic)
#pragma unroll
            for (int i = 0; i < MSF * MSF; ++i) {
                part[i].x = warp_sum(part[i].x);
                part[i].y = warp_sum(part[i].y);
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < MSF * MSF; ++i) sm.Spart[i] = part[i];
            }
            if (lane < MSF * MSF && site + 1 < N) {
                const int r = lane / MSF, c = lane % MSF;
                sm.Gblk[(buf ^ 1) * MSF * MSF + lane] = G[size_t(site + 1 + c * N) * D + site + 1 + r * N];
            }
        } else if (prev_acc) {
            // ---------------------------------------------- finish the previous (accepted) site
            const int pb = buf ^ 1;
#pragma unroll
            for (int q = 0; q < TPT; ++q) {
                const int t = gt + q * gthreads;
                if (t < D)
                    extend_xy<MSF>(X, Y, G, D, N, KMAX, K_done, site - 1, t, sm.xrow + pb * MSF * KMAX,
                                   sm.ycol + pb * MSF * KMAX, sm.Delta, sm.Minv, xrow, ycol, true, Gr[q], Gc[q], sm.Ys);
            }
        }
        TICK(0)
        __syncthreads();
        TICK(1)
        // ================================================================== phase 2: decision (warp 0)
        const int K = MSF * j;
        if (warp == 0) {
            // site block of the effective Green's function: S = G[blk] + (terms staged in phase 1) + (the
            // terms of the update appended in phase 1, if any)
            TICK(4)
            if (lane == 0) {
                if (!have_rng) {
                    sAbort = 1;
                    sAccept = 0;
                } else {
                    cplx S[MSF * MSF], M[MSF * MSF];
#pragma unroll
                    for (int r = 0; r < MSF; ++r)
#pragma unroll
                        for (int c = 0; c < MSF; ++c) {
                            cplx sacc = cadd(sm.Gblk[buf * MSF * MSF + r * MSF + c], sm.Spart[r * MSF + c]);
                            for (int l = K_done; l < K; ++l) sacc = cfma(xrow[r * KMAX + l], ycol[c * KMAX + l], sacc);
                            S[r * MSF + c] = sacc;
                        }
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
                    cplx Minv[MSF * MSF];
                    const cplx det = small_det_inv<MSF>(M, Minv);
                    const double probFermion = (OPDIM == 3) ? det.x : (det.x * det.x + det.y * det.y);
                    const double prob = probSPhi * probFermion;
                    cur += OPDIM;
                    bool acc;
                    if (prob > 1.0) {
                        acc = true;
                    } else {
                        acc = sm.rng[cur] < prob;
                        cur += 1;
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
                    sAccept = acc ? 1 : 0;
                }
            }
        }
        else if (warp >= kDecWarps) {
            // idle otherwise: request row / column `site` of G now, so that the update of an accepted site
            // does not wait for them in the next phase 1
#pragma unroll
            for (int q = 0; q < TPT; ++q) {
                const int t = gt + q * gthreads;
                if (t < D) load_g_rowcol<MSF>(G, D, N, site, t, Gr[q], Gc[q]);
            }
        }
        TICK(2)
        __syncthreads();
        TICK(3)
        if (sAbort) break;
        prev_acc = sAccept != 0;
        if (prev_acc) {
            j += 1;
            if (j == delayNow) {
                // the delay block is full: append the last update now, then flush (in the kernel or on the host)
                if (warp >= kDecWarps) {
#pragma unroll
                    for (int q = 0; q < TPT; ++q) {
                        const int t = gt + q * gthreads;
                        if (t < D)
                            extend_xy<MSF>(X, Y, G, D, N, KMAX, MSF * (j - 1), site, t, xrow, ycol, sm.Delta, sm.Minv,
                                           nullptr, nullptr, false, Gr[q], Gc[q], sm.Ys);
                    }
                }
                prev_acc = false;
                if (!a.inline_flush) { ++site; break; }
                __syncthreads();
                flush_delayed(G, X, Y, D, MSF * j, KMAX);
                j = 0;
                delayNow = min(md.delaySteps, N - (site + 1));
                __syncthreads();
                // pending terms are gone: clear the staged rows / columns; G changed: refresh the prefetched
                // site block of the next site
                for (int i = tid; i < 2 * MSF * KMAX; i += blockDim.x) {
                    sm.xrow[i] = make_double2(0, 0);
                    sm.ycol[i] = make_double2(0, 0);
                }
                if (tid < MSF * MSF && site + 1 < N) {
                    const int r = tid / MSF, c = tid % MSF;
                    sm.Gblk[(buf ^ 1) * MSF * MSF + tid] = G[size_t(site + 1 + c * N) * D + site + 1 + r * N];
                }
                __syncthreads();
            }
        }
    }
    if (prev_acc && !sAbort) {
        // the slice ended right after an accepted site: append its update
        if (warp >= kDecWarps) {
            const int pb = (site - 1) & 1;
#pragma unroll
            for (int q = 0; q < TPT; ++q) {
                const int t = gt + q * gthreads;
                if (t < D)
                    extend_xy<MSF>(X, Y, G, D, N, KMAX, MSF * (j - 1), site - 1, t, sm.xrow + pb * MSF * KMAX,
                                   sm.ycol + pb * MSF * KMAX, sm.Delta, sm.Minv, nullptr, nullptr, false, Gr[q], Gc[q], sm.Ys);
            }
        }
    }
    __syncthreads();
#ifdef DQMC_UPD_TIMING
    if (a.debug && b == 0 && (tid == 0 || tid == 32 || tid == kDecThreads))
        printf("upd dbg round %d tid %3d sites %d: ph1 %lld waitA %lld ph2(dec) %lld waitB %lld Sred %lld top %lld\n", a.round, tid,
               site - site0, tq[0], tq[1], tq[2], tq[3], tq[4], tq[5]);
#endif
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
        if (site >= N && !sAbort && a.final_pass) {
            // end of the slice (last pass): acceptance statistics and step-size adaptation
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


// =================================================================================================
// Window rounds: the same Metropolis chain, organised so that NO D-wide work sits between two decisions.
//
// A round covers a window of w consecutive sites.  The CTA keeps the (MSF w) x (MSF w) block of the effective
// Green's function that couples the window's sites to each other in shared memory ("Gw") and applies every accepted
// update to that block immediately (rank MSF, O((MSF w)^2)): by the algebra of updateInSlice_delayed
// (detsdwopdim.cpp:3064-3138) the window block of G + X Y evolves by its own rows and columns only, so the
// decision of a site needs nothing but its MSF x MSF diagonal block of Gw.
//
// The D-wide columns X_j = C_j Delta_j and rows Y_j = M_j^-1 (R_j - 1_j) of the delayed update are NOT formed here.
// Every C_j is a combination of columns of the round's INITIAL Green's function G0 at the accepted sites, and every
// R_j - 1_j a combination of its rows:
//     X_l = sum_i G0[:, S_i] Tx[i, l],      Y_l = sum_i Ty[l, i] (G0 - 1)[S_i, :]      (S_i = s_j + q N for term i = MSF j + q)
// with block-triangular K x K coefficient matrices that follow the reference's recurrences
//     Tx[:, (j, r)]  = sum_q ( e_(j,q) + sum_{l < MSF j} Tx[:, l] Y_l[S_(j,q)] ) Delta_j[q, r]
//     Ty[(j, q'), :] = sum_q Minv_j[q', q] ( e_(j,q)^T + sum_{l < MSF j} X_l[S_(j,q)] Ty[l, :] )
// whose couplings Y_l[S_(j,q)], X_l[S_(j,q)] are window entries.  Two helper warps run these recurrences while the
// chain goes on; the round hands A = Tx Ty to `update_gather_kernel`, which copies the K columns of G0 into X and
// forms Y = A (G0 - 1)[S, :] on all SMs; the rank-K GEMM G += X Y follows as before.  The arithmetic is that of the
// reference's delayed update, re-associated.
//
//   warp 0        proposal look-up, decision (all lanes redundantly: no divergence, no shuffles); on acceptance
//                 lane <-> future site: window parts of X_j, Y_j and the diagonal blocks of all future sites
//   warp 1 / 2    Tx / Ty recurrences of the accepted update, fields and tables of the accepted site (polling a counter;
//                 never waited for by warp 0)
//   warps 3..     apply Gw += x_w y_w to the future part of the window while warp 0 goes on with the next sites
//                 (named barriers kBarStart / kBarDone; warp 0 waits only when the NEXT acceptance arrives
//                 before the previous block update has finished)
//
// Output of a round, per replica: header (ints) nacc, site0, w, 0, sites[J]; scratch (cplx) A[i * KM + i'], KM = MSF delaySteps.
// =================================================================================================
constexpr int kBarStart = 1, kBarDone = 2;

__device__ __forceinline__ void named_bar_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int n) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// determinant of the MSF x MSF decision matrix; the inverse is needed only when the proposal is accepted
template <int MSF>
__device__ __forceinline__ cplx small_det(const cplx* M) {
    if (MSF == 2) return csub(cmul(M[0], M[3]), cmul(M[1], M[2]));
    cplx inv[MSF * MSF];
    return small_det_inv<MSF>(M, inv);
}

#ifdef DQMC_UPD_TIMING
#define WTICK(i) { long long now__; asm volatile("mov.u64 %0, %%clock64;" : "=l"(now__) :: "memory"); wq[i] += now__ - wmark; wmark = now__; }
#else
#define WTICK(i)
#endif

// Proposal table of a window: everything about the proposal of window site `pos` that does not depend on earlier
// decisions, for every random-number cursor the site can be reached with (cursor = OPDIM pos + e, e = number of
// acceptance draws consumed so far, 0 <= e <= pos).  Entry (pos, e) sits at index pos (pos + 1) / 2 + e.
template <int MSF, int OPDIM>
struct PropTable {
    static constexpr int DIFF = 1, NEWP = 1 + OPDIM, CNEW = 1 + 2 * OPDIM, XNEW = 2 + 2 * OPDIM;
    static constexpr int DELTA = (3 + 2 * OPDIM + 1) & ~1;             // 16-byte aligned complex block
    static constexpr int STRIDE = DELTA + 2 * MSF * MSF;               // doubles per entry
};

// Shared-memory carve-up of the window kernel (offsets in doubles; every region starts 16-byte aligned)
struct WinLayout {
    int phik, tsum, ck, xk, rngs, ptab, Gw, Sdiag, xh, yh, Tx, Ty, smallD, smallM, total;
};
__host__ __device__ inline int win_even(int x) { return (x + 1) & ~1; }
// packed block-triangular K x K coefficient matrix: term l = MSF j + r holds its first MSF (j + 1) entries
__host__ __device__ inline int win_tri_off(int msf, int l) {
    const int j = l / msf, r = l - j * msf;
    return msf * msf * (j * (j + 1) / 2) + r * msf * (j + 1);
}
__host__ __device__ inline WinLayout win_layout(int msf, int opdim, int N, int wmax, int JM) {
    WinLayout o;
    const int wpm = msf * wmax, ldw = wpm + 1, KM = msf * JM;
    const int pstride = ((3 + 2 * opdim + 1) & ~1) + 2 * msf * msf;
    int p = 0;
    o.phik = p;   p += win_even(opdim * N);
    o.tsum = p;   p += win_even(opdim * wmax);
    o.ck = p;     p += win_even(wmax);
    o.xk = p;     p += win_even(wmax);
    o.rngs = p;   p += win_even(wmax * (opdim + 1));
    o.ptab = p;   p += wmax * (wmax + 1) / 2 * pstride;
    o.Gw = p;     p += 2 * ldw * wpm;
    o.Sdiag = p;  p += 2 * wmax * msf * msf;
    o.xh = p;     p += 2 * KM * wpm;
    o.yh = p;     p += 2 * KM * wpm;
    o.Tx = p;     p += 2 * msf * msf * (JM * (JM + 1) / 2);
    o.Ty = p;     p += 2 * msf * msf * (JM * (JM + 1) / 2);
    o.smallD = p; p += 2 * JM * msf * msf;
    o.smallM = p; p += 2 * JM * msf * msf;
    o.total = p;
    return o;
}

constexpr int kWinMaxJ = 64;

template <int MSF, int OPDIM, int BW>
__global__ void __launch_bounds__(32 * (3 + BW)) update_window_kernel(UpdateModel md, UpdateArgs a) {
    pdl_enter();
    constexpr int NT = 32 * (3 + BW);                      // warp 0: chain, warps 1 / 2: Tx / Ty, BW block-update warps
    constexpr int NBLK = 32 * (1 + BW);                    // participants of kBarStart / kBarDone
    typedef PropTable<MSF, OPDIM> PT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
#ifdef DQMC_UPD_TIMING
    long long wq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long wmark = clock64();
#endif
    const int D = md.D, N = md.N, L = md.L;
    const int wmax = a.wmax, WPM = MSF * wmax, ldw = WPM + 1;
    const int JM = md.delaySteps, KM = MSF * JM;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* hdr = a.whdr + size_t(b) * a.strideHdr;
    const int site0 = a.round == 0 ? 0 : a.site_state[b];
    if (site0 >= N) {                                      // slice already finished in an earlier round
        if (tid == 0) {
            if (a.kvec) a.kvec[b] = 0;
            hdr[0] = 0;
        }
        return;
    }
    const int w = min(wmax, N - site0);                    // sites of this window
    const int WP = MSF * w;                                // window positions: a = i + r * w  <->  matrix index site0 + i + r * N
    const WinLayout lay = win_layout(MSF, OPDIM, N, wmax, JM);
    double* sm = reinterpret_cast<double*>(smem_raw);
    double* phik = sm + lay.phik;                          // [OPDIM][N]   fields of this slice, kept current
    double* tsum = sm + lay.tsum;                          // [OPDIM][wmax] phi(k+1) + phi(k-1) at the window's sites
    double* ck = sm + lay.ck;                              // [wmax] cosh table
    double* xk = sm + lay.xk;                              // [wmax] sinh table
    double* rngs = sm + lay.rngs;                          // [wmax * (OPDIM+1)] random numbers from the cursor on
    double* ptab = sm + lay.ptab;                          // [PT::STRIDE][NE] proposal table, field-major (lanes read different entries)
    const int NE = wmax * (wmax + 1) / 2;
    cplx* Gw = reinterpret_cast<cplx*>(sm + lay.Gw);       // [WPM][ldw] column major
    cplx* Sdiag = reinterpret_cast<cplx*>(sm + lay.Sdiag); // [MSF*MSF][wmax] diagonal blocks, element-major
    cplx* xh = reinterpret_cast<cplx*>(sm + lay.xh);       // [KM][WPM] window parts of X_l, term l = MSF j + q
    cplx* yh = reinterpret_cast<cplx*>(sm + lay.yh);       // [KM][WPM] window parts of Y_l
    cplx* Txp = reinterpret_cast<cplx*>(sm + lay.Tx);      // packed: Tx[i, l] at win_tri_off(l) + i
    cplx* Typ = reinterpret_cast<cplx*>(sm + lay.Ty);      // packed: Ty[l, i] at win_tri_off(l) + i
    cplx* smallD = reinterpret_cast<cplx*>(sm + lay.smallD);   // [JM][MSF*MSF] Delta_j
    cplx* smallM = reinterpret_cast<cplx*>(sm + lay.smallM);   // [JM][MSF*MSF] M_j^-1
    __shared__ int sPos, sTerm, sQuit, sNload, sNacc;
    __shared__ int mbPos[kWinMaxJ], mbEnt[kWinMaxJ];       // accepted updates, in order (chain -> helper warps)
    __shared__ int nPosted, chainDone;

    const int k = a.k;
    const cplx* __restrict__ G = a.G + size_t(b) * a.strideG;
    double* phi = a.phi + size_t(b) * a.stridePhi;
    double* coshT = a.coshT + size_t(b) * a.strideTab;
    double* sinhT = a.sinhT + size_t(b) * a.strideTab;
    const double* rng = a.rng + size_t(b) * a.strideRng;
    cplx* scratch = a.wscratch + size_t(b) * a.strideScratch;
    const double rpar = a.rvals[b];
    const double phiDelta = a.ctrl[b].phiDelta;
    const double dtau = md.dtau;
    const int cursor0 = a.cursor[b];
    double* phik_g = phi + size_t(k) * OPDIM * N;
    {
        // window block of G: column c of the block is MSF contiguous runs of w elements of a column of G
        // (warp <-> column, lane <-> row of a run).  The loads of the first pass stay in flight while the fields,
        // tables and random numbers of the window are fetched: one global round trip instead of two.
        cplx gv[4][MSF];
        auto gw_load = [&](int c0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * (NT / 32);
                if (c < WP) {
                    const int rc = c / w, ic = c - rc * w;
                    const cplx* col = G + size_t(site0 + ic + rc * N) * D + site0;
#pragma unroll
                    for (int r = 0; r < MSF; ++r)
                        if (lane < w) gv[u][r] = col[lane + r * N];
                }
            }
        };
        auto gw_store = [&](int c0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * (NT / 32);
                if (c < WP && lane < w) {
#pragma unroll
                    for (int r = 0; r < MSF; ++r) Gw[lane + r * w + size_t(c) * ldw] = gv[u][r];
                }
            }
        };
        gw_load(warp);
        const int kEarlier = k > 1 ? k - 1 : md.m;
        const int kLater = k < md.m ? k + 1 : 1;
        const double* pl = phi + size_t(kLater) * OPDIM * N;
        const double* pe = phi + size_t(kEarlier) * OPDIM * N;
        for (int i = tid; i < OPDIM * N; i += NT) phik[i] = phik_g[i];
        for (int i = tid; i < OPDIM * w; i += NT) {
            const int d = i / w, pos = i - d * w;
            tsum[d * wmax + pos] = pl[d * N + site0 + pos] + pe[d * N + site0 + pos];
        }
        for (int i = tid; i < w; i += NT) {
            ck[i] = coshT[size_t(k) * N + site0 + i];
            xk[i] = sinhT[size_t(k) * N + site0 + i];
        }
        const int want = w * (OPDIM + 1);
        const int have = max(0, min(want, a.rngWindow - cursor0));
        for (int i = tid; i < have; i += NT) rngs[i] = rng[cursor0 + i];
        gw_store(warp);
        for (int c0 = warp + 4 * (NT / 32); c0 < WP; c0 += 4 * (NT / 32)) {
            gw_load(c0);
            gw_store(c0);
        }
        if (tid == 0) { sQuit = 0; sNload = have; sPos = 0; sTerm = 0; sNacc = 0; nPosted = 0; chainDone = 0; }
    }
    __syncthreads();
    for (int i = tid; i < w * MSF * MSF; i += NT) {
        const int pos = i / (MSF * MSF), e = i - pos * MSF * MSF, r = e / MSF, c = e - r * MSF;
        Sdiag[e * wmax + pos] = Gw[pos + r * w + size_t(pos + c * w) * ldw];
    }
    // ---- proposal table (proposeNewPhiBox, deltaSPhi without the neighbour term, get_delta_forsite): one entry per thread
    {
        const int nload = sNload;
        const double invc2dtau = 1.0 / (md.c * md.c * dtau);
        const double lamdtau = md.lambda * dtau;
        const int nent = w * (w + 1) / 2;
        for (int ent = tid; ent < nent; ent += NT) {
            int pos = int((sqrt(8.0 * ent + 1.0) - 1.0) * 0.5);
            while (pos * (pos + 1) / 2 > ent) --pos;
            while ((pos + 1) * (pos + 2) / 2 <= ent) ++pos;
            const int e = ent - pos * (pos + 1) / 2;
            const int cur = OPDIM * pos + e;
            double* T = ptab + ent;                        // field f at T[f * NE]
            if (cur + OPDIM > nload) continue;             // never reached: the chain aborts before it would read this entry
            const int site = site0 + pos;
            double oldp[3] = {0, 0, 0}, newp[3] = {0, 0, 0};
            double oldSq = 0, newSq = 0, tdot = 0;
#pragma unroll
            for (int d = 0; d < OPDIM; ++d) {
                oldp[d] = phik[d * N + site];
                const double u = rngs[cur + d];
                newp[d] = oldp[d] + (-phiDelta + (phiDelta - (-phiDelta)) * u);   // randRange(-delta, +delta)
                const double diff = newp[d] - oldp[d];
                oldSq += oldp[d] * oldp[d];
                newSq += newp[d] * newp[d];
                tdot += tsum[d * wmax + pos] * diff;
                T[(PT::DIFF + d) * NE] = diff;
                T[(PT::NEWP + d) * NE] = newp[d];
            }
            const double sqDiff = newSq - oldSq;
            const double pow4Diff = newSq * newSq - oldSq * oldSq;
            const double d1 = invc2dtau * (sqDiff - tdot);
            const double d3 = dtau * (0.5 * rpar * sqDiff + 0.25 * md.u * pow4Diff);
            T[0] = (d1 + d3) + 0.5 * dtau * (4.0 * sqDiff);     // deltaSPhi without -dtau * (neighbour sum . diff)
            double cNew, sc;
            cosh_sinhc(lamdtau * sqrt(newSq), cNew, sc);
            const double xNew = lamdtau * sc;                   // sinh(lambda dtau |phi|) / |phi|
            T[PT::CNEW * NE] = cNew;
            T[PT::XNEW * NE] = xNew;
            cplx evOld[MSF * MSF], emvNew[MSF * MSF];
            ev_block<MSF, OPDIM>(evOld, +1.0, oldp, ck[pos], xk[pos]);
            ev_block<MSF, OPDIM>(emvNew, -1.0, newp, cNew, xNew);
#pragma unroll
            for (int r = 0; r < MSF; ++r)
#pragma unroll
                for (int c = 0; c < MSF; ++c) {
                    cplx sacc = make_double2(r == c ? -1.0 : 0.0, 0.0);
#pragma unroll
                    for (int t = 0; t < MSF; ++t) sacc = cfma(emvNew[r * MSF + t], evOld[t * MSF + c], sacc);
                    T[(PT::DELTA + 2 * (r * MSF + c)) * NE] = sacc.x;
                    T[(PT::DELTA + 2 * (r * MSF + c) + 1) * NE] = sacc.y;
                }
        }
    }
    __syncthreads();
    WTICK(0)

    if (warp == 0) {
        // ============================================================ the Metropolis chain
        // A rejected proposal changes nothing but the random-number cursor (by OPDIM + 1), so lane i evaluates site
        // pos + i under the assumption that the sites pos .. pos + i - 1 are rejected; the first lane that accepts ends
        // the batch, the lanes before it were evaluated on the true state and stand, the ones after it are discarded.
        int cur = 0, j = 0, pos = 0;
        unsigned accepted = 0;
        bool aborted = false, outstanding = false;
        const int delayNow = min(JM, N - site0);
        const int nload = sNload;
        while (pos < w) {
            const int mypos = pos + lane;
            const bool in = mypos < w;
            const int mycur = cur + lane * (OPDIM + 1);
            const bool have = in && (mycur + OPDIM + 1 <= nload);
            bool acc = false, draw = false;
            int ent = 0;
            cplx Minv[MSF * MSF];
            if (have) {
                cplx Dl[MSF * MSF];                             // Delta of this proposal: re-read from the table on acceptance
                ent = mypos * (mypos + 1) / 2 + (mycur - OPDIM * mypos);
                const double* T = ptab + ent;
                // bosonic part: the neighbour term of deltaSPhi uses the CURRENT fields of the slice
                double sdot = 0;
                {
                    const int site = site0 + mypos;
                    const int y = site / L, x = site - y * L;
                    const int nb0 = y * L + (x + 1 == L ? 0 : x + 1);
                    const int nb1 = y * L + (x == 0 ? L - 1 : x - 1);
                    const int nb2 = (y + 1 == L ? 0 : y + 1) * L + x;
                    const int nb3 = (y == 0 ? L - 1 : y - 1) * L + x;
#pragma unroll
                    for (int d = 0; d < OPDIM; ++d) {
                        const double* pk = phik + d * N;
                        const double sn = ((pk[nb0] + pk[nb1]) + pk[nb2]) + pk[nb3];
                        sdot += sn * T[(PT::DIFF + d) * NE];
                    }
                }
                const double udraw = rngs[mycur + OPDIM];       // consumed only if the probability is <= 1
                const double probSPhi = exp(-(T[0] - dtau * sdot));       // field 0: deltaSPhi without the neighbour term
#pragma unroll
                for (int i = 0; i < MSF * MSF; ++i)
                    Dl[i] = make_double2(T[(PT::DELTA + 2 * i) * NE], T[(PT::DELTA + 2 * i + 1) * NE]);
                // ------------------------------------------ decision: M = 1 - S Delta + Delta
                cplx M[MSF * MSF];
#pragma unroll
                for (int r = 0; r < MSF; ++r)
#pragma unroll
                    for (int c = 0; c < MSF; ++c) {
                        cplx sacc = make_double2(r == c ? 1.0 : 0.0, 0.0);
#pragma unroll
                        for (int t = 0; t < MSF; ++t)
                            sacc = csub(sacc, cmul(Sdiag[(r * MSF + t) * wmax + mypos], Dl[t * MSF + c]));
                        M[r * MSF + c] = cadd(sacc, Dl[r * MSF + c]);
                    }
                // determinant and inverse together: the reciprocal overlaps with the exponential of the bosonic part
                const cplx det = small_det_inv<MSF>(M, Minv);
                const double probFermion = (OPDIM == 3) ? det.x : (det.x * det.x + det.y * det.y);
                const double prob = probSPhi * probFermion;
                draw = !(prob > 1.0);
                acc = draw ? (udraw < prob) : true;
            }
            const unsigned inMask = __ballot_sync(0xffffffffu, in);
            const unsigned haveMask = __ballot_sync(0xffffffffu, have);
            const unsigned accMask = __ballot_sync(0xffffffffu, acc);
            const unsigned stopMask = inMask & ~haveMask;       // sites the window holds no random numbers for
            const int ia = accMask ? __ffs(accMask) - 1 : 32;
            const int is = stopMask ? __ffs(stopMask) - 1 : 32;
            WTICK(1)
            if (is < ia) { pos += is; aborted = true; break; }
            if (ia == 32) {                                     // every site of the batch rejected
                const int nb = __popc(inMask);
                pos += nb;
                cur += nb * (OPDIM + 1);
                continue;
            }
            // ---------------------------------------------- accepted: lane ia's proposal
            const int apos = pos + ia;
            if (lane == ia) {                                   // record for the other lanes and the helper warps (they poll nPosted)
                const double* T = ptab + ent;
#pragma unroll
                for (int d = 0; d < OPDIM; ++d) phik[d * N + site0 + apos] = T[(PT::NEWP + d) * NE];
                mbPos[j] = apos;
                mbEnt[j] = ent;
#pragma unroll
                for (int i = 0; i < MSF * MSF; ++i) {
                    smallD[j * MSF * MSF + i] = make_double2(T[(PT::DELTA + 2 * i) * NE], T[(PT::DELTA + 2 * i + 1) * NE]);
                    smallM[j * MSF * MSF + i] = Minv[i];
                }
                __threadfence_block();
                *(volatile int*)&nPosted = j + 1;
            }
            cur = __shfl_sync(0xffffffffu, mycur + OPDIM + (draw ? 1 : 0), ia);
            __syncwarp();
            const bool last = (j + 1 == delayNow) || (apos + 1 == w);
            if (!last) {
                cplx Dl[MSF * MSF];
#pragma unroll
                for (int i = 0; i < MSF * MSF; ++i) {
                    Dl[i] = smallD[j * MSF * MSF + i];
                    Minv[i] = smallM[j * MSF * MSF + i];
                }
                WTICK(2)
                if (outstanding) {                          // column / row `apos` of Gw must be current
                    named_bar_sync(kBarDone, NBLK);
                    outstanding = false;
                }
                WTICK(4)
                // window parts of X_j, Y_j and the diagonal blocks of the future sites (lane <-> future site)
                const int f = apos + 1 + lane;
                if (f < w) {
                    cplx* xt = xh + size_t(MSF) * j * WPM;
                    cplx* yt = yh + size_t(MSF) * j * WPM;
                    cplx xv[MSF][MSF], yv[MSF][MSF];        // xv[r][q] = X_(j,q)[f + r w],  yv[q][r] = Y_(j,q)[f + r w]
#pragma unroll
                    for (int r = 0; r < MSF; ++r) {
                        const int ar = f + r * w;
                        cplx cg[MSF], rg[MSF];
#pragma unroll
                        for (int p = 0; p < MSF; ++p) {
                            cg[p] = Gw[ar + size_t(apos + p * w) * ldw];
                            rg[p] = Gw[apos + p * w + size_t(ar) * ldw];
                        }
#pragma unroll
                        for (int q = 0; q < MSF; ++q) {
                            cplx xa = make_double2(0, 0), ya = make_double2(0, 0);
#pragma unroll
                            for (int p = 0; p < MSF; ++p) {
                                xa = cfma(cg[p], Dl[p * MSF + q], xa);
                                ya = cfma(Minv[q * MSF + p], rg[p], ya);
                            }
                            xv[r][q] = xa;
                            yv[q][r] = ya;
                            xt[q * WPM + ar] = xa;
                            yt[q * WPM + ar] = ya;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < MSF; ++r)
#pragma unroll
                        for (int c = 0; c < MSF; ++c) {
                            cplx sacc = Sdiag[(r * MSF + c) * wmax + f];
#pragma unroll
                            for (int q = 0; q < MSF; ++q) sacc = cfma(xv[r][q], yv[q][c], sacc);
                            Sdiag[(r * MSF + c) * wmax + f] = sacc;
                        }
                }
                if (lane == 0) { sPos = apos; sTerm = MSF * j; }
                __threadfence_block();
                __syncwarp();
                named_bar_arrive(kBarStart, NBLK);          // the block update of the window starts
                outstanding = true;
            }
            accepted += 1;
            j += 1;
            pos = apos + 1;
            WTICK(2)
            if (j == delayNow) break;
        }
        if (outstanding) named_bar_sync(kBarDone, NBLK);
        if (lane == 0) {
            sQuit = 1;
            sNacc = aborted ? 0 : j;
            __threadfence_block();
            *(volatile int*)&chainDone = 1;
        }
        __syncwarp();
        named_bar_arrive(kBarStart, NBLK);
        WTICK(3)
#ifdef DQMC_UPD_TIMING
        if (a.debug && b == 0 && lane == 0)
            printf("win dbg round %d sites %d acc %d: prologue %lld site %lld post %lld tail %lld wait %lld\n",
                   a.round, pos, j, wq[0], wq[1], wq[2], wq[3], wq[4]);
#endif
        if (lane == 0) {
            const int site = aborted ? N : site0 + pos;
            if (aborted) atomicExch(a.errflag, 1);
            a.cursor[b] = cursor0 + cur;
            hdr[0] = aborted ? 0 : j;
            hdr[1] = site0;
            hdr[2] = w;
            if (a.kvec) a.kvec[b] = aborted ? 0 : MSF * j;
            a.site_state[b] = site;
            const unsigned total = (a.round == 0 ? 0u : a.accepted[b]) + accepted;
            a.accepted[b] = total;
            if (a.acceptedTotal) a.acceptedTotal[b] += accepted;
            if (site >= N && !aborted && a.final_pass) {
                // end of the slice (last pass): acceptance statistics and step-size adaptation
                // (RunningAverage::addValue, RunningAverage.h:57-68; updateInSliceThermalization, detsdwopdim.cpp:3329-3341)
                dqmc_control_data* cd = a.ctrl + b;
                const double ratio = double(total) / double(N);
                cd->lastAccRatioLocal_phi = ratio;
                if (a.thermalization) {
                    const int ps = cd->ra_samples_added % 100;
                    if (cd->ra_samples_added >= 100) cd->ra_average -= cd->ra_values[ps] / 100.0;
                    cd->ra_values[ps] = ratio;
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
    } else if (warp == 1 || warp == 2) {
        // ============================================================ coefficient recurrences of the accepted updates
        // warp 1: column block j of Tx (lane <-> row i) + fields / tables of the accepted site in global memory
        // warp 2: row block j of Ty (lane <-> column i)
        const bool xside = warp == 1;
        cplx* Tp = xside ? Txp : Typ;
        const cplx* coup = xside ? yh : xh;                // Y_l[S_(j,q)] for Tx, X_l[S_(j,q)] for Ty
        const cplx* smallS = xside ? smallD : smallM;
        for (int j = 0;; ++j) {
            for (;;) {
                if (*(volatile int*)&nPosted > j) break;
                if (*(volatile int*)&chainDone) {
                    if (*(volatile int*)&nPosted > j) break;
                    goto helpers_done;
                }
                __nanosleep(40);
            }
            __threadfence_block();
            const int pos = mbPos[j];
            cplx S[MSF * MSF];
#pragma unroll
            for (int i = 0; i < MSF * MSF; ++i) S[i] = smallS[j * MSF * MSF + i];
            const int nrows = MSF * (j + 1), nprev = MSF * j;
            for (int i = lane; i < nrows; i += 32) {
                cplx c[MSF];
#pragma unroll
                for (int q = 0; q < MSF; ++q) c[q] = make_double2(i == nprev + q ? 1.0 : 0.0, 0.0);
                for (int l = (i / MSF) * MSF; l < nprev; ++l) {
                    const cplx t = Tp[win_tri_off(MSF, l) + i];
                    const cplx* cp = coup + size_t(l) * WPM + pos;
#pragma unroll
                    for (int q = 0; q < MSF; ++q) c[q] = cfma(t, cp[q * w], c[q]);
                }
#pragma unroll
                for (int r = 0; r < MSF; ++r) {
                    cplx v = make_double2(0, 0);
#pragma unroll
                    for (int q = 0; q < MSF; ++q)
                        v = xside ? cfma(c[q], S[q * MSF + r], v)          // Tx[:, (j, r)] = sum_q c_q Delta_j[q, r]
                                  : cfma(S[r * MSF + q], c[q], v);         // Ty[(j, r), :] = sum_q Minv_j[r, q] r_q
                    Tp[win_tri_off(MSF, nprev + r) + i] = v;
                }
            }
            if (xside && lane == 0) {
                const int site = site0 + pos;
                const double* T = ptab + mbEnt[j];
#pragma unroll
                for (int d = 0; d < OPDIM; ++d) phik_g[d * N + site] = T[(PT::NEWP + d) * NE];
                coshT[size_t(k) * N + site] = T[PT::CNEW * NE];
                sinhT[size_t(k) * N + site] = T[PT::XNEW * NE];
                hdr[4 + j] = site;
            }
            __syncwarp();
        }
    helpers_done:;
    } else {
        // ============================================================ block updates of the window
        // thread <-> (window position a of the row, column group cg): rows and columns at or before the accepted site are
        // masked out instead of compacted (no index arithmetic in the loop), eight columns per thread
        const int tb = tid - 96;
        const int ar = tb & 63, cg = tb >> 6;
        constexpr int NCG = BW * 32 / 64;
        int ari = ar;
        while (ari >= w) ari -= w;                          // site index of the row within the window
        for (;;) {
            named_bar_sync(kBarStart, NBLK);
            if (*(volatile int*)&sQuit) break;
            const int pos = *(volatile int*)&sPos;
            const int tb0 = *(volatile int*)&sTerm;
            for (int a0 = ar; a0 < WP; a0 += 64) {
                int ai = a0 == ar ? ari : a0;
                while (ai >= w) ai -= w;
                if (ai <= pos) continue;
                cplx xv[MSF];
#pragma unroll
                for (int q = 0; q < MSF; ++q) xv[q] = xh[size_t(tb0 + q) * WPM + a0];
                int c = cg, ci = cg;
                while (ci >= w) ci -= w;
                for (; c < WP; c += NCG) {
                    if (ci > pos) {
                        cplx gv = Gw[a0 + size_t(c) * ldw];
#pragma unroll
                        for (int q = 0; q < MSF; ++q) gv = cfma(xv[q], yh[size_t(tb0 + q) * WPM + c], gv);
                        Gw[a0 + size_t(c) * ldw] = gv;
                    }
                    ci += NCG;
                    while (ci >= w) ci -= w;
                }
            }
            __threadfence_block();
            named_bar_arrive(kBarDone, NBLK);
        }
    }
    __syncthreads();
    // ---- A = Tx Ty, the K x K core of the round's update  G += G0[:, S] A (G0 - 1)[S, :]
    {
        const int Kc = MSF * sNacc;
        for (int e = tid; e < Kc * Kc; e += NT) {
            const int i = e / Kc, ip = e - i * Kc;
            cplx acc = make_double2(0, 0);
            for (int l = (max(i, ip) / MSF) * MSF; l < Kc; ++l) {
                const int off = win_tri_off(MSF, l);
                acc = cfma(Txp[off + i], Typ[off + ip], acc);
            }
            scratch[size_t(i) * KM + ip] = acc;
        }
    }
}

// X and Y of a window round for all rows / columns, on all SMs (one CTA per tile of kGthTile matrix indices t):
//   X_i[t] = G0[t, S_i]                                   (copy of K columns)
//   Y_i[t] = sum_i' A[i, i'] ( G0[S_i', t] - delta(t, S_i') )
// so that G0 + X Y is the Green's function after the round's accepted updates (detsdwopdim.cpp:3064-3070, 3123-3138, 3156).
constexpr int kGthThreads = 256, kGthTile = 32;

template <int MSF>
__global__ void __launch_bounds__(kGthThreads) update_gather_kernel(UpdateModel md, UpdateArgs a) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.y;
    const int* hdr = a.whdr + size_t(b) * a.strideHdr;
    const int J = hdr[0];
    if (J <= 0) return;
    const int D = md.D, N = md.N, KM = MSF * md.delaySteps;
    const int K = MSF * J;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    constexpr int NW = kGthThreads / 32;
    const int t0 = blockIdx.x * kGthTile;
    cplx* As = reinterpret_cast<cplx*>(smem_raw);          // [K][K]
    cplx* Rs = As + size_t(KM) * KM;                       // [K][kGthTile + 1]
    __shared__ int sidx[kWinMaxJ * 4];
#ifdef DQMC_UPD_TIMING
    long long wq[4] = {0, 0, 0, 0};
    long long wmark = clock64();
#endif
    const cplx* __restrict__ scratch = a.wscratch + size_t(b) * a.strideScratch;
    const cplx* __restrict__ G = a.G + size_t(b) * a.strideG;
    cplx* X = a.X + size_t(b) * a.strideXY;
    cplx* Y = a.Y + size_t(b) * a.strideXY;
    for (int l = tid; l < K; l += kGthThreads) sidx[l] = hdr[4 + l / MSF] + (l % MSF) * N;
    for (int e = tid; e < K * K; e += kGthThreads) {
        const int i = e / K, ip = e - i * K;
        As[e] = scratch[size_t(i) * KM + ip];
    }
    __syncthreads();
    WTICK(0)
    // rows of G0 - 1 at the accepted sites for this tile (lane <-> term: the sites of a window are close together)
    for (int tt = wv; tt < kGthTile; tt += NW) {
        const int t = t0 + tt;
        for (int l = lane; l < K; l += 32) {
            cplx v = make_double2(0, 0);
            if (t < D) {
                const int s = sidx[l];
                v = G[size_t(t) * D + s];
                if (t == s) v.x -= 1.0;
            }
            Rs[l * (kGthTile + 1) + tt] = v;
        }
    }
    // columns of G0 at the accepted sites
    {
        const int t = t0 + lane;
        if (t < D)
            for (int l = wv; l < K; l += NW) X[size_t(l) * D + t] = G[size_t(sidx[l]) * D + t];
    }
    __syncthreads();
    WTICK(1)
    {
        const int t = t0 + lane;
        for (int i0 = wv; i0 < K; i0 += 4 * NW) {
            cplx acc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = make_double2(0, 0);
            for (int l = 0; l < K; ++l) {
                const cplx rv = Rs[l * (kGthTile + 1) + lane];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = min(i0 + u * NW, K - 1);
                    acc[u] = cfma(As[i * K + l], rv, acc[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * NW;
                if (i < K && t < D) Y[size_t(i) * D + t] = acc[u];
            }
        }
    }
    WTICK(2)
#ifdef DQMC_UPD_TIMING
    if (a.debug && b == 0 && blockIdx.x == 0 && tid == 0)
        printf("gth dbg round %d J %d: coef %lld gload %lld product %lld\n", a.round, J, wq[0], wq[1], wq[2]);
#endif
}


// The same for small batches: tiles of 8 matrix indices (four times the CTAs of the kernel above: with one replica per
// launch that one runs as D / 32 CTAs and every phase is a dependent global round trip), one output row of Y per thread.
constexpr int kGthSmallTile = 8;

template <int MSF>
__global__ void __launch_bounds__(kGthThreads) update_gather_small_kernel(UpdateModel md, UpdateArgs a) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE = kGthSmallTile, NG = kGthThreads / TILE;
    const int b = blockIdx.y;
    const int* hdr = a.whdr + size_t(b) * a.strideHdr;
    const int D = md.D, N = md.N, KM = MSF * md.delaySteps;
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * TILE;
    cplx* As = reinterpret_cast<cplx*>(smem_raw);          // [KM][KM]
    cplx* Rs = As + size_t(KM) * KM;                       // [K][TILE + 1]
    __shared__ int sidx[kWinMaxJ * 4];
    const cplx* __restrict__ scratch = a.wscratch + size_t(b) * a.strideScratch;
    const cplx* __restrict__ G = a.G + size_t(b) * a.strideG;
    cplx* X = a.X + size_t(b) * a.strideXY;
    cplx* Y = a.Y + size_t(b) * a.strideXY;
    // the whole coefficient buffer is fetched whatever the number of accepted updates: its loads are in flight during
    // the header round trip (entries beyond K x K are stale and never used)
    cplx av[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = tid + u * kGthThreads;
        if (e < KM * KM) av[u] = scratch[e];
    }
    const int J = hdr[0];
    if (J <= 0) return;
    const int K = MSF * J;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = tid + u * kGthThreads;
        if (e < KM * KM) As[e] = av[u];
    }
    for (int e = tid + 4 * kGthThreads; e < KM * KM; e += kGthThreads) As[e] = scratch[e];
    for (int l = tid; l < K; l += kGthThreads) sidx[l] = hdr[4 + l / MSF] + (l % MSF) * N;
    __syncthreads();
    // rows of G0 - 1 at the accepted sites (thread <-> term fastest: the sites of a window are close together)
    for (int idx = tid; idx < K * TILE; idx += kGthThreads) {
        const int tt = idx / K, l = idx - tt * K;
        const int t = t0 + tt;
        cplx v = make_double2(0, 0);
        if (t < D) {
            const int s = sidx[l];
            v = G[size_t(t) * D + s];
            if (t == s) v.x -= 1.0;
        }
        Rs[l * (TILE + 1) + tt] = v;
    }
    // columns of G0 at the accepted sites
    for (int idx = tid; idx < K * TILE; idx += kGthThreads) {
        const int l = idx / TILE, tt = idx - l * TILE;
        const int t = t0 + tt;
        if (t < D) X[size_t(l) * D + t] = G[size_t(sidx[l]) * D + t];
    }
    __syncthreads();
    {
        const int tt = tid % TILE, ig = tid / TILE;
        const int t = t0 + tt;
        for (int i = ig; i < K; i += NG) {
            cplx acc0 = make_double2(0, 0), acc1 = make_double2(0, 0);
            int l = 0;
            for (; l + 1 < K; l += 2) {
                acc0 = cfma(As[i * KM + l], Rs[l * (TILE + 1) + tt], acc0);
                acc1 = cfma(As[i * KM + l + 1], Rs[(l + 1) * (TILE + 1) + tt], acc1);
            }
            if (l < K) acc0 = cfma(As[i * KM + l], Rs[l * (TILE + 1) + tt], acc0);
            if (t < D) Y[size_t(i) * D + t] = make_double2(acc0.x + acc1.x, acc0.y + acc1.y);
        }
    }
}

}  // namespace

int update_rounds_per_slice(const UpdateModel& m, int inline_flush) {
    if (inline_flush) return 1;
    // a round ends after delaySteps acceptances or (window rounds) at the end of its window
    const int w = update_window_sites(m);
    const int per = (w > 0 && w < m.delaySteps) ? w : m.delaySteps;
    return (m.N + per - 1) / per;
}

cudaError_t update_round_launch(const UpdateModel& m, const UpdateArgs& a, cudaStream_t st) {
    const int tpt = m.D > 896 ? 2 : 1;
    if (m.D > 2 * 896) return cudaErrorInvalidValue;
    const int gth = ((((m.D + tpt - 1) / tpt) + 31) / 32) * 32;
    const int threads = kDecThreads + gth;
    const int kmaxp = (m.msf * m.delaySteps + 3) & ~3;
    size_t smem = size_t((2 * m.opdim + 2) * m.N + ((m.N * (m.opdim + 1) + 1) & ~1)) * sizeof(double) +
                        size_t(5) * m.msf * m.msf * sizeof(cplx) + size_t(4) * m.msf * kmaxp * sizeof(cplx);
    UpdateArgs aa = a;
    const size_t ybytes = size_t(kmaxp) * m.D * sizeof(cplx);
    aa.y_in_smem = (!a.inline_flush && smem + ybytes <= 200 * 1024) ? 1 : 0;     // pending Y rows in shared memory
    if (aa.y_in_smem) smem += ybytes;
#define LAUNCH(MSF, OPD, TPT, MAXT)                                                                         \
    {                                                                                                       \
        cudaError_t e = cudaFuncSetAttribute(update_round_kernel<MSF, OPD, TPT, MAXT>,                      \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        if (e != cudaSuccess) return e;                                                                     \
        launch_pdl(update_round_kernel<MSF, OPD, TPT, MAXT>, dim3(a.batch), dim3(threads), smem, st, m, aa);                     \
    }
#define LAUNCH3(MSF, OPD)                                                                                   \
    {                                                                                                       \
        if (threads <= 352) LAUNCH(MSF, OPD, 1, 352)                                                        \
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


// ---- window rounds ------------------------------------------------------------------------------
// Sites per window: two delay blocks' worth (a round ends after delaySteps acceptances, at ~50 % acceptance that
// takes 2 * delaySteps sites), bounded by the shared memory of the (MSF w)^2 window block.
static size_t window_smem_bytes(const UpdateModel& m, int wmax) {
    return size_t(win_layout(m.msf, m.opdim, m.N, wmax, m.delaySteps).total) * sizeof(double);
}
int update_window_sites(const UpdateModel& m) {
    if (m.delaySteps < 8) return 0;                        // tiny delay blocks are flushed inside the legacy kernel
    if (m.msf * m.delaySteps > 64 || m.delaySteps > kWinMaxJ) return 0;   // K x K coefficient matrices in shared memory
    int w = 2 * m.delaySteps;
    if (w > 32) w = 32;                                    // one lane per future site on acceptance
    if (w > m.N) w = m.N;
    w &= ~1;
    while (w >= 2 && window_smem_bytes(m, w) > size_t(225) * 1024) w -= 2;   // 227 KB per CTA minus the static part
    return w < 4 ? 0 : w;
}
size_t update_window_scratch_elems(const UpdateModel& m) {
    const size_t K = size_t(m.msf) * m.delaySteps;
    return K * K;
}
int update_window_hdr_ints(const UpdateModel& m) { return (4 + m.delaySteps + 3) & ~3; }

cudaError_t update_window_launch(const UpdateModel& m, const UpdateArgs& a, cudaStream_t st) {
    const size_t smem = window_smem_bytes(m, a.wmax);
#define LAUNCHW(MSF, OPD, BW)                                                                               \
    {                                                                                                       \
        cudaError_t e = cudaFuncSetAttribute(update_window_kernel<MSF, OPD, BW>,                            \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        if (e != cudaSuccess) return e;                                                                     \
        launch_pdl(update_window_kernel<MSF, OPD, BW>, dim3(a.batch), dim3(32 * (3 + BW)), smem, st, m, a);                     \
    }
    // 16 block-update warps for the 2 x 2 site blocks (96 registers per thread suffice); the 4 x 4 blocks of O(3)
    // need the registers more than the warps
    if (m.opdim == 1) LAUNCHW(2, 1, 16)
    else if (m.opdim == 2) LAUNCHW(2, 2, 16)
    else LAUNCHW(4, 3, 4)
#undef LAUNCHW
    return cudaGetLastError();
}

cudaError_t update_build_xy_launch(const UpdateModel& m, const UpdateArgs& a, cudaStream_t st) {
    const int KM = m.msf * m.delaySteps;
    if (m.delaySteps > kWinMaxJ) return cudaErrorInvalidValue;
    const size_t smem = (size_t(KM) * KM + size_t(KM) * (kGthTile + 1)) * sizeof(cplx);
    // small batches: 8-wide tiles (DQMC_GATHER_SMALL=0 / 1 overrides)
    static const int smallEnv = std::getenv("DQMC_GATHER_SMALL") ? std::atoi(std::getenv("DQMC_GATHER_SMALL")) : -1;
    const bool small = smallEnv >= 0 ? smallEnv != 0 : gemm_matrices_in_flight() <= 16;
    if (small) {
        const size_t smem_s = (size_t(KM) * KM + size_t(KM) * (kGthSmallTile + 1)) * sizeof(cplx);
        dim3 grid_s((m.D + kGthSmallTile - 1) / kGthSmallTile, a.batch);
#define LAUNCHS(MSF)                                                                                        \
    {                                                                                                       \
        cudaError_t e = cudaFuncSetAttribute(update_gather_small_kernel<MSF>,                               \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);    \
        if (e != cudaSuccess) return e;                                                                     \
        launch_pdl(update_gather_small_kernel<MSF>, dim3(grid_s), dim3(kGthThreads), smem_s, st, m, a);     \
    }
        if (m.msf == 2) LAUNCHS(2)
        else LAUNCHS(4)
#undef LAUNCHS
        return cudaGetLastError();
    }
    dim3 grid((m.D + kGthTile - 1) / kGthTile, a.batch);
#define LAUNCHB(MSF)                                                                                        \
    {                                                                                                       \
        cudaError_t e = cudaFuncSetAttribute(update_gather_kernel<MSF>,                                     \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        if (e != cudaSuccess) return e;                                                                     \
        launch_pdl(update_gather_kernel<MSF>, dim3(grid), dim3(kGthThreads), smem, st, m, a);                                   \
    }
    if (m.msf == 2) LAUNCHB(2)
    else LAUNCHB(4)
#undef LAUNCHB
    return cudaGetLastError();
}

}  // namespace dqmc
