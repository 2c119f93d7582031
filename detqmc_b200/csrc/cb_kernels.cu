// Checkerboard B-matrix multiplies for DetSDW on sm_100a.
//
// Replaces, for a whole batch of replicas and a whole chain of time slices in ONE pass over the
// matrix: cbLMultHoppingExp / cbRMultHoppingExp (detsdwopdim.cpp:1836-1988), the plaquette kernels
// cb_assaad_applyBondFactors{Left,Right}[_precalcedMatrices] (:1686-1756, :1786-1826, :1903-1943),
// leftMultiplyBk / leftMultiplyBkInv / rightMultiplyBk / rightMultiplyBkInv (:1994-2402) and the
// chain wrappers checkerboard{Left,Right}MultiplyBmat[Inv] (:2074-2090, :2170-2186, :2305-2324,
// :2404-2420).
//
// B_k = e^{-dtau V_k} * blockdiag_bandspin( e^{dtau mu_b} E1(1/2) E0(1) E1(1/2) ),  E_g = product of
// the 4-site plaquette exponentials of subgroup g.  A left multiply acts on every COLUMN of A
// independently, a right multiply on every ROW, so a CTA stages a tile of 8 or 16 vectors (each
// of length D = msf*N) in shared memory, applies all slices of the chain there, and writes the tile
// back: HBM traffic is one read + one write of the matrix per chain, whatever its length.  The
// reference recomputes every checkerboard block 2x (O(2)) / 3x (O(3)) and copies the matrix per
// pass; here all band blocks of a vector are transformed once.
//
// Two kernels: `cb_mult_bulk_kernel` for column tiles (bulk-copy engine in and out, natural site order with padded
// lattice rows; see its own comment) and the general `cb_mult_kernel` for row tiles and for shapes the bulk path does
// not take.  The general kernel's
// shared-memory layout: inside every band-spin block the N sites are stored in 4-sublattice order
// (x parity, y parity, then plaquette index), so the four corners of consecutive plaquettes are four
// unit-stride streams -- conflict-free 16-byte accesses for both plaquette subgroups -- and the
// per-site potential stage is unit stride as well.  A thread owns one (plaquette, band-spin) pair
// and keeps that plaquette's 4x4 matrix in registers while it walks over the vectors of the tile;
// the matrices are Hermitian (exponentials of Hermitian hopping blocks), stored as 4 real diagonal
// + 6 complex upper entries, and real symmetric without magnetic flux (REALM variant: half the
// flops).  The transposed matrices needed by the right multiplies are the complex conjugates.
//
// Bound: HBM for single-slice launches (algorithmic bytes = 2 * D^2 * 16 per launch and matrix),
// FP64 pipe for long chains.
#include "dqmc_internal.h"

#include <cmath>
#include <cstdlib>
#include <complex>

namespace dqmc {

// ------------------------------------------------------------------------------------------------
// host: plaquette tables
// ------------------------------------------------------------------------------------------------
// One Hermitian 4x4 matrix = 8 cplx components: {(d0,d1), (d2,d3), o01, o02, o03, o12, o13, o23}, stored
// component-major so that consecutive threads (consecutive plaquettes q) read consecutive addresses:
// index: (((band*2 + sign_idx)*2 + pass) * 8 + component) * nplaq + q
// pass 0: subgroup 1, half step; pass 1: subgroup 0, full step times e^{-+dtau mu_band}.
// Plaquette q of subgroup g sits at x = 2*(q % (L/2)) + g, y = 2*(q / (L/2)) + g.
int cb_table_count(const CbGeom& g) { return 2 * 2 * 3 * g.nplaq * 8; }

namespace {

typedef std::complex<double> zc;

// exp(A) for a 4x4 complex matrix by scaling and squaring with a Taylor series (norms here are
// O(dtau * t) << 1, the series converges to round-off in ~12 terms).
void expm4(const zc* A, zc* out) {
    double nrm = 0;
    for (int i = 0; i < 16; ++i) nrm = std::max(nrm, std::abs(A[i]));
    int sq = 0;
    double scale = 1.0;
    while (nrm * 4 * scale > 0.25) { scale *= 0.5; ++sq; }
    zc S[16], term[16], acc[16], tmp[16];
    for (int i = 0; i < 16; ++i) S[i] = A[i] * scale;
    for (int i = 0; i < 16; ++i) { acc[i] = (i % 5 == 0) ? 1.0 : 0.0; term[i] = acc[i]; }
    for (int k = 1; k <= 24; ++k) {
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                zc s = 0;
                for (int x = 0; x < 4; ++x) s += term[r * 4 + x] * S[x * 4 + c];
                tmp[r * 4 + c] = s / double(k);
            }
        for (int i = 0; i < 16; ++i) { term[i] = tmp[i]; acc[i] += term[i]; }
    }
    for (int q = 0; q < sq; ++q) {
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                zc s = 0;
                for (int x = 0; x < 4; ++x) s += acc[r * 4 + x] * acc[x * 4 + c];
                tmp[r * 4 + c] = s;
            }
        for (int i = 0; i < 16; ++i) acc[i] = tmp[i];
    }
    for (int i = 0; i < 16; ++i) out[i] = acc[i];
}

}  // namespace

// exp(pf * h_plaquette) for the plaquette (i, j = i + x, k = i + y, l = k + x) with lower-left corner (i1, i2) of
// band `band`; pf = sign * dtau (/2): detsdwopdim.cpp:1597-1684 (flux), 1786-1826 (no flux)
static void plaquette_matrix(const dqmc_params& p, int band, int i1, int i2, double pf, zc* M) {
    const int L = p.L, N = L * L;
    const double pi = M_PI;
    double hh = band == 0 ? p.txhor : p.tyhor, hv = band == 0 ? p.txver : p.tyver;
    if ((p.bc == 1 || p.bc == 3) && i1 == L - 1) hh = -hh;
    if ((p.bc == 2 || p.bc == 3) && i2 == L - 1) hv = -hv;
    if (!p.weakZflux) {
        const double chh = std::cosh(pf * hh), shh = std::sinh(-pf * hh);
        const double chv = std::cosh(pf * hv), shv = std::sinh(-pf * hv);
        const double a = chh * chv, b = chv * shh, c = chh * shv, d = shh * shv;
        const double m4[16] = {a, b, c, d, b, a, d, c, c, d, a, b, d, c, b, a};
        for (int i = 0; i < 16; ++i) M[i] = m4[i];
        return;
    }
    const double zmag = 1.0 / N;
    const int j1 = (i1 + 1) % L, k2 = (i2 + 1) % L;
    const zc ph_ij = std::exp(zc(0, -2.0 * pi * zmag * i2));
    const zc ph_kl = std::exp(zc(0, -2.0 * pi * zmag * k2));
    zc ph_ik = 1.0, ph_jl = 1.0;
    if (i2 == L - 1) {
        ph_ik = std::exp(zc(0, 2.0 * pi * zmag * L * i1));
        ph_jl = std::exp(zc(0, 2.0 * pi * zmag * L * j1));
    }
    zc H[16];
    for (int i = 0; i < 16; ++i) H[i] = 0;
    H[0 * 4 + 1] = ph_ij * hh;
    H[0 * 4 + 2] = ph_ik * hv;
    H[1 * 4 + 3] = ph_jl * hv;
    H[2 * 4 + 3] = ph_kl * hh;
    zc Hs[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) Hs[r * 4 + c] = -pf * (H[r * 4 + c] + std::conj(H[c * 4 + r]));
    expm4(Hs, M);
}

// Dense block-diagonal D x D matrices of shiftGreenSymmetric (detsdwopdim.cpp:4505-4612, CB_ASSAAD_BERG): block b
// (band b % 2) of SL is E0(-dtau/2) E1(-dtau/2), of SR it is E1(+dtau/2) E0(+dtau/2), E_s = product of the disjoint
// plaquette exponentials of subgroup s (no chemical potential: it cancels between the two sides).  Column-major.
void cb_build_shift_matrices(const dqmc_params& p, int msf, std::vector<cplx>& SL, std::vector<cplx>& SR) {
    const int L = p.L, N = L * L, D = msf * N;
    SL.assign(size_t(D) * D, make_double2(0, 0));
    SR.assign(size_t(D) * D, make_double2(0, 0));
    auto subgroup = [&](int band, int sg, double pf, std::vector<zc>& E) {
        E.assign(size_t(N) * N, zc(0));
        for (int i2 = sg; i2 < L; i2 += 2)
            for (int i1 = sg; i1 < L; i1 += 2) {
                const int i = i2 * L + i1, j = i2 * L + (i1 + 1) % L, k = ((i2 + 1) % L) * L + i1,
                          l = ((i2 + 1) % L) * L + (i1 + 1) % L;
                const int idx[4] = {i, j, k, l};
                zc M[16];
                plaquette_matrix(p, band, i1, i2, pf, M);
                for (int r = 0; r < 4; ++r)
                    for (int c = 0; c < 4; ++c) E[size_t(idx[c]) * N + idx[r]] = M[r * 4 + c];     // column-major
            }
    };
    auto matmul = [&](const std::vector<zc>& A, const std::vector<zc>& B, std::vector<zc>& C) {
        C.assign(size_t(N) * N, zc(0));
        for (int c = 0; c < N; ++c)
            for (int k = 0; k < N; ++k) {
                const zc b = B[size_t(c) * N + k];
                if (b == zc(0)) continue;
                for (int r = 0; r < N; ++r) C[size_t(c) * N + r] += A[size_t(k) * N + r] * b;
            }
    };
    const double h = 0.5 * p.dtau;
    for (int band = 0; band < 2; ++band) {
        std::vector<zc> E0m, E1m, E0p, E1p, Lm, Rm;
        subgroup(band, 0, -h, E0m); subgroup(band, 1, -h, E1m);
        subgroup(band, 0, +h, E0p); subgroup(band, 1, +h, E1p);
        matmul(E0m, E1m, Lm);
        matmul(E1p, E0p, Rm);
        for (int b = band; b < msf; b += 2)
            for (int c = 0; c < N; ++c)
                for (int r = 0; r < N; ++r) {
                    const zc l = Lm[size_t(c) * N + r], rr = Rm[size_t(c) * N + r];
                    SL[size_t(b * N + c) * D + b * N + r] = make_double2(l.real(), l.imag());
                    SR[size_t(b * N + c) * D + b * N + r] = make_double2(rr.real(), rr.imag());
                }
    }
}

// Dense hopping propagators of DetSDW<CB_NONE> (setupPropK, detsdwopdim.cpp:1210-1286; computePropagator,
// detmodel.cpp:31-39): block b (band b % 2) of P is exp(-dtau k_band), of Pinv exp(+dtau k_band), with the Hermitian
// single-particle matrix k = -mu - hoppings (antiperiodic signs across the boundary, Peierls phases of the weak flux).
// The reference diagonalises k; here exp() is a scaled Taylor series with repeated squaring (|dtau k| < 1).
static void expm_dense(const std::vector<zc>& A, int n, std::vector<zc>& out) {
    double nrm = 0;
    for (int r = 0; r < n; ++r) {
        double row = 0;
        for (int c = 0; c < n; ++c) row += std::abs(A[size_t(r) * n + c]);
        nrm = std::max(nrm, row);
    }
    int sq = 0;
    double scale = 1.0;
    while (nrm * scale > 0.25) { scale *= 0.5; ++sq; }
    std::vector<zc> S(A.size()), term(A.size()), acc(A.size()), tmp(A.size());
    for (size_t i = 0; i < A.size(); ++i) S[i] = A[i] * scale;
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) acc[size_t(r) * n + c] = term[size_t(r) * n + c] = (r == c) ? 1.0 : 0.0;
    auto mul = [&](const std::vector<zc>& X, const std::vector<zc>& Y, std::vector<zc>& Z) {
        std::fill(Z.begin(), Z.end(), zc(0));
        for (int r = 0; r < n; ++r)
            for (int k = 0; k < n; ++k) {
                const zc x = X[size_t(r) * n + k];
                if (x == zc(0)) continue;
                for (int c = 0; c < n; ++c) Z[size_t(r) * n + c] += x * Y[size_t(k) * n + c];
            }
    };
    for (int k = 1; k <= 20; ++k) {
        mul(term, S, tmp);
        for (size_t i = 0; i < tmp.size(); ++i) { term[i] = tmp[i] / double(k); acc[i] += term[i]; }
    }
    for (int q = 0; q < sq; ++q) {
        mul(acc, acc, tmp);
        acc.swap(tmp);
    }
    out = acc;
}

void cb_build_dense_propagators(const dqmc_params& p, int msf, std::vector<cplx>& P, std::vector<cplx>& Pinv) {
    const int L = p.L, N = L * L, D = msf * N;
    const double pi = M_PI;
    P.assign(size_t(D) * D, make_double2(0, 0));
    Pinv.assign(size_t(D) * D, make_double2(0, 0));
    for (int band = 0; band < 2; ++band) {
        const double hh = band == 0 ? p.txhor : p.tyhor, hv = band == 0 ? p.txver : p.tyver;
        const double mu = band == 0 ? p.mux : p.muy;
        const double zmag = p.weakZflux ? 1.0 / N : 0.0;             // zmag[XUP] = zmag[YDOWN] = +1/N (:219-220)
        std::vector<zc> k(size_t(N) * N, zc(0));                      // row-major k(site, neigh)
        for (int site = 0; site < N; ++site) k[size_t(site) * N + site] = -mu;
        for (int site = 0; site < N; ++site) {
            const int x = site % L, y = site / L;
            for (int dir = 0; dir < 4; ++dir) {                       // XPLUS, XMINUS, YPLUS, YMINUS
                int nx = x, ny = y;
                double hop = dir < 2 ? hh : hv;
                zc phase = 1.0;
                if (dir == 0) { nx = (x + 1) % L; if ((p.bc == 1 || p.bc == 3) && x == L - 1) hop = -hop;
                                phase = std::exp(zc(0, -2.0 * pi * zmag * y)); }
                if (dir == 1) { nx = (x + L - 1) % L; if ((p.bc == 1 || p.bc == 3) && x == 0) hop = -hop;
                                phase = std::exp(zc(0, +2.0 * pi * zmag * y)); }
                if (dir == 2) { ny = (y + 1) % L; if ((p.bc == 2 || p.bc == 3) && y == L - 1) hop = -hop;
                                if (y == L - 1) phase = std::exp(zc(0, +2.0 * pi * zmag * L * x)); }
                if (dir == 3) { ny = (y + L - 1) % L; if ((p.bc == 2 || p.bc == 3) && y == 0) hop = -hop;
                                if (y == 0) phase = std::exp(zc(0, -2.0 * pi * zmag * L * x)); }
                k[size_t(site) * N + ny * L + nx] -= hop * phase;
            }
        }
        std::vector<zc> a(k.size()), em, ep;
        for (size_t i = 0; i < k.size(); ++i) a[i] = -p.dtau * k[i];
        expm_dense(a, N, em);
        for (size_t i = 0; i < k.size(); ++i) a[i] = +p.dtau * k[i];
        expm_dense(a, N, ep);
        for (int b = band; b < msf; b += 2)
            for (int r = 0; r < N; ++r)
                for (int c = 0; c < N; ++c) {
                    const zc m1 = em[size_t(r) * N + c], p1 = ep[size_t(r) * N + c];
                    P[size_t(b * N + c) * D + b * N + r] = make_double2(m1.real(), m1.imag());        // column-major
                    Pinv[size_t(b * N + c) * D + b * N + r] = make_double2(p1.real(), p1.imag());
                }
    }
}

void cb_build_tables(const dqmc_params& p, std::vector<cplx>& out) {
    const int L = p.L, N = L * L, nplaq = N / 4, half = L / 2;
    // passes: 0 = subgroup 1, half step; 1 = subgroup 0, full step with the chemical potential; 2 = subgroup 0, half
    // step without it (second factor of shiftGreenSymmetric, detsdwopdim.cpp:4505-4612)
    out.assign(size_t(2) * 2 * 3 * nplaq * 8, make_double2(0, 0));
    for (int band = 0; band < 2; ++band) {
        const double mu = band == 0 ? p.mux : p.muy;
        for (int si = 0; si < 2; ++si) {
            const double sign = si == 0 ? -1.0 : +1.0;
            for (int pass = 0; pass < 3; ++pass) {
                const int subgroup = pass == 0 ? 1 : 0;
                const double pf = sign * p.dtau * (pass == 1 ? 1.0 : 0.5);
                const double ovfac = pass == 1 ? std::exp(-sign * p.dtau * mu) : 1.0;
                for (int q = 0; q < nplaq; ++q) {
                    const int i1 = 2 * (q % half) + subgroup;     // x
                    const int i2 = 2 * (q / half) + subgroup;     // y
                    zc M[16];
                    plaquette_matrix(p, band, i1, i2, pf, M);
                    // exp of a Hermitian block is Hermitian; symmetrise the round-off and compress
                    cplx* dst = out.data() + ((size_t(band) * 2 + si) * 3 + pass) * 8 * nplaq + q;
                    double dg[4];
                    for (int r = 0; r < 4; ++r) dg[r] = M[r * 4 + r].real() * ovfac;
                    dst[0 * nplaq] = make_double2(dg[0], dg[1]);
                    dst[1 * nplaq] = make_double2(dg[2], dg[3]);
                    int o = 2;
                    for (int r = 0; r < 4; ++r)
                        for (int c = r + 1; c < 4; ++c) {
                            const zc v = 0.5 * (M[r * 4 + c] + std::conj(M[c * 4 + r])) * ovfac;
                            dst[size_t(o++) * nplaq] = make_double2(v.real(), v.imag());
                        }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kCbMaxThreads = 256;

__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {    // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ cplx cfmac(cplx a, cplx b, cplx c) {   // conj(a)*b + c
    return make_double2(fma(a.x, b.x, fma(a.y, b.y, c.x)), fma(a.x, b.y, fma(-a.y, b.x, c.y)));
}
__device__ __forceinline__ cplx rfma(double a, cplx b, cplx c) {  // a*b + c, a real
    return make_double2(fma(a, b.x, c.x), fma(a, b.y, c.y));
}
__device__ __forceinline__ cplx rmul(double a, cplx b) { return make_double2(a * b.x, a * b.y); }

// One plaquette matrix in registers.
template <bool REALM>
struct PlaqMat {
    double d[4];
    cplx o[6];      // o01 o02 o03 o12 o13 o23 (imaginary parts unused when REALM)
    // M points at component 0 of this thread's plaquette; components are `stride` apart
    __device__ __forceinline__ void load(const cplx* __restrict__ M, int stride, bool conj) {
        const cplx t0 = __ldg(M), t1 = __ldg(M + stride);
        d[0] = t0.x; d[1] = t0.y; d[2] = t1.x; d[3] = t1.y;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            o[i] = __ldg(M + (2 + i) * stride);
            if (!REALM && conj) o[i].y = -o[i].y;
        }
    }
    __device__ __forceinline__ void apply(cplx& a, cplx& b, cplx& c, cplx& e) const {
        cplx r0 = rmul(d[0], a), r1 = rmul(d[1], b), r2 = rmul(d[2], c), r3 = rmul(d[3], e);
        if (REALM) {
            r0 = rfma(o[0].x, b, r0); r0 = rfma(o[1].x, c, r0); r0 = rfma(o[2].x, e, r0);
            r1 = rfma(o[0].x, a, r1); r1 = rfma(o[3].x, c, r1); r1 = rfma(o[4].x, e, r1);
            r2 = rfma(o[1].x, a, r2); r2 = rfma(o[3].x, b, r2); r2 = rfma(o[5].x, e, r2);
            r3 = rfma(o[2].x, a, r3); r3 = rfma(o[4].x, b, r3); r3 = rfma(o[5].x, c, r3);
        } else {
            r0 = cfma(o[0], b, r0);  r0 = cfma(o[1], c, r0);  r0 = cfma(o[2], e, r0);
            r1 = cfmac(o[0], a, r1); r1 = cfma(o[3], c, r1);  r1 = cfma(o[4], e, r1);
            r2 = cfmac(o[1], a, r2); r2 = cfmac(o[3], b, r2); r2 = cfma(o[5], e, r2);
            r3 = cfmac(o[2], a, r3); r3 = cfmac(o[4], b, r3); r3 = cfmac(o[5], c, r3);
        }
        a = r0; b = r1; c = r2; e = r3;
    }
};

struct CbShape {
    int half, quarter;      // L/2, N/4
    int pairs;              // msf * nplaq
    int G;                  // vector groups of the hopping passes (divides the tile's vector count)
    int Gp;                 // vector groups of the potential stage
};

// position of site s inside its band-spin block of the shared-memory tile
__device__ __forceinline__ int perm_site(int s, int L, int half, int quarter) {
    const int x = s % L, y = s / L;
    return ((x & 1) | ((y & 1) << 1)) * quarter + (y >> 1) * half + (x >> 1);
}

template <int MSF, bool REALM>
__device__ __forceinline__ void hopping_pass(cplx* tile, int ldt, int nv, const cplx* __restrict__ tab,
                                             const CbGeom& g, const CbShape& sh, int sign_idx, int transposed,
                                             int pass) {
    const int items = sh.pairs * sh.G;
    for (int item = threadIdx.x; item < items; item += blockDim.x) {
        const int grp = item / sh.pairs;
        const int p = item - grp * sh.pairs;
        const int bs = p / g.nplaq;
        const int q = p - bs * g.nplaq;
        const int band = bs & 1;       // XUP, YDOWN, XDOWN, YUP -> x, y, x, y
        PlaqMat<REALM> M;
        M.load(tab + ((size_t(band) * 2 + sign_idx) * 3 + pass) * 8 * g.nplaq + q, g.nplaq, transposed != 0);
        int oi, oj, ok, ol;
        const int base = bs * g.N;
        if (pass >= 1) {               // subgroup 0: (even, even) corner
            oi = base + q; oj = oi + sh.quarter; ok = oj + sh.quarter; ol = ok + sh.quarter;
        } else {                       // subgroup 1: (odd, odd) corner, neighbours wrap around
            const int py = q / sh.half, px = q - py * sh.half;
            const int pxp = px + 1 == sh.half ? 0 : px + 1;
            const int pyp = py + 1 == sh.half ? 0 : py + 1;
            oi = base + 3 * sh.quarter + q;
            oj = base + 2 * sh.quarter + py * sh.half + pxp;
            ok = base + 1 * sh.quarter + pyp * sh.half + px;
            ol = base + pyp * sh.half + pxp;
        }
        for (int v = grp; v < nv; v += sh.G) {
            cplx* t = tile + v * ldt;
            cplx a = t[oi], b = t[oj], c = t[ok], e = t[ol];
            M.apply(a, b, c, e);
            t[oi] = a; t[oj] = b; t[ok] = c; t[ol] = e;
        }
    }
}

template <int MSF, bool REALM>
__device__ __forceinline__ void hopping_stage(cplx* tile, int ldt, int nv, const cplx* __restrict__ tab,
                                              const CbGeom& g, const CbShape& sh, int sign_idx, int transposed) {
    hopping_pass<MSF, REALM>(tile, ldt, nv, tab, g, sh, sign_idx, transposed, 0);
    __syncthreads();
    hopping_pass<MSF, REALM>(tile, ldt, nv, tab, g, sh, sign_idx, transposed, 1);
    __syncthreads();
    hopping_pass<MSF, REALM>(tile, ldt, nv, tab, g, sh, sign_idx, transposed, 0);
    __syncthreads();
}

// half-step stage of shiftGreenSymmetric (detsdwopdim.cpp:4505-4612): E1(h) then E0(h), no chemical potential
template <int MSF, bool REALM>
__device__ __forceinline__ void shift_stage(cplx* tile, int ldt, int nv, const cplx* __restrict__ tab, const CbGeom& g,
                                            const CbShape& sh, int sign_idx, int transposed) {
    hopping_pass<MSF, REALM>(tile, ldt, nv, tab, g, sh, sign_idx, transposed, 0);
    __syncthreads();
    hopping_pass<MSF, REALM>(tile, ldt, nv, tab, g, sh, sign_idx, transposed, 2);
    __syncthreads();
}

// coefficients of the per-site block of e^{sign*dtau*V} (evMatrix, detsdwopdim.cpp:3188-3229 with
// cdwU == 0): c on the diagonal, e01 = sign x (phi0 - i phi1), e10 = conj(e01), a = sign x phi2
struct PotCoef { double c, ex, ey, a; };

template <int MSF>
__device__ __forceinline__ PotCoef potential_coef(const double* __restrict__ phi_k, const double* __restrict__ cosh_k,
                                                  const double* __restrict__ sinh_k, const CbGeom& g, int s,
                                                  int sign_idx, int transposed) {
    PotCoef pc;
    const double x = sign_idx == 0 ? -__ldg(sinh_k + s) : __ldg(sinh_k + s);
    pc.c = __ldg(cosh_k + s);
    pc.ex = x * __ldg(phi_k + s);
    const double p1 = g.opdim > 1 ? __ldg(phi_k + g.N + s) : 0.0;
    pc.ey = transposed ? x * p1 : -x * p1;          // imaginary part of e01 (e10 is its conjugate)
    pc.a = MSF == 4 ? x * __ldg(phi_k + 2 * g.N + s) : 0.0;
    return pc;
}

template <int MSF>
__device__ __forceinline__ void potential_apply(cplx* t, int N, int pos, const PotCoef& pc) {
    const cplx e01 = make_double2(pc.ex, pc.ey);
    if (MSF == 2) {
        const cplx v0 = t[pos], v1 = t[N + pos];
        t[pos] = cfma(e01, v1, rmul(pc.c, v0));
        t[N + pos] = cfmac(e01, v0, rmul(pc.c, v1));
    } else {
        // rows of e^{sign dtau V}: (c, e01, 0, a), (e10, c, -a, 0), (0, -a, c, e10), (a, 0, e01, c)
        const cplx v0 = t[pos], v1 = t[N + pos], v2 = t[2 * N + pos], v3 = t[3 * N + pos];
        t[pos] = rfma(pc.a, v3, cfma(e01, v1, rmul(pc.c, v0)));
        t[N + pos] = rfma(-pc.a, v2, cfmac(e01, v0, rmul(pc.c, v1)));
        t[2 * N + pos] = rfma(-pc.a, v1, cfmac(e01, v3, rmul(pc.c, v2)));
        t[3 * N + pos] = rfma(pc.a, v0, cfma(e01, v2, rmul(pc.c, v3)));
    }
}

template <int MSF, bool REALM, int TV>
__global__ void __launch_bounds__(kCbMaxThreads) cb_mult_kernel(CbGeom g, CbLaunch a, CbShape sh) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = g.D, N = g.N;
    const int ldt = D + 1;
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);
    int* sperm = reinterpret_cast<int*>(tile + TV * ldt);      // [D] matrix index -> tile position
    int* sinv = sperm + D;                                               // [N] tile position -> site

    const int b = blockIdx.y;
    const int v0 = blockIdx.x * TV;
    const int nv = min(TV, D - v0);
    cplx* A = a.A + size_t(b) * a.strideA;
    const double* phi = a.phi + size_t(b) * a.stridePhi;
    const double* coshT = a.coshT + size_t(b) * a.strideTab;
    const double* sinhT = a.sinhT + size_t(b) * a.strideTab;

    for (int e = threadIdx.x; e < D; e += blockDim.x) {
        const int bs = e / N, s = e - bs * N;
        const int pos = perm_site(s, g.L, sh.half, sh.quarter);
        sperm[e] = bs * N + pos;
        if (bs == 0) sinv[pos] = s;
    }
    __syncthreads();

    // ---- load the tile (coalesced along the contiguous direction of the column-major matrix); the
    // global loads of a thread are issued back to back (8 in flight) before the shared-memory stores
    if (!a.rows) {
        if (nv == TV) {
            for (int e = threadIdx.x; e < D; e += blockDim.x) {
                cplx buf[TV];
#pragma unroll
                for (int v = 0; v < TV; ++v) buf[v] = A[size_t(v0 + v) * D + e];
                const int pos = sperm[e];
#pragma unroll
                for (int v = 0; v < TV; ++v) tile[v * ldt + pos] = buf[v];
            }
        } else {
            for (int v = 0; v < nv; ++v) {
                const cplx* src = A + size_t(v0 + v) * D;
                cplx* t = tile + v * ldt;
                for (int e = threadIdx.x; e < D; e += blockDim.x) t[sperm[e]] = src[e];
            }
        }
    } else if (nv == TV) {
        constexpr int U = 8;
        for (int base = 0; base < TV * D; base += U * blockDim.x) {
            cplx buf[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = base + u * blockDim.x + threadIdx.x;
                if (idx < TV * D) buf[u] = A[size_t(idx / TV) * D + v0 + idx % TV];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = base + u * blockDim.x + threadIdx.x;
                if (idx < TV * D) tile[(idx % TV) * ldt + sperm[idx / TV]] = buf[u];
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < nv * D; idx += blockDim.x) {
            const int e = idx / nv, v = idx - e * nv;
            tile[v * ldt + sperm[e]] = A[size_t(e) * D + v0 + v];
        }
    }
    __syncthreads();

    // potential stage: item = (site position, vector group)
    const int pot_items = N * sh.Gp;
    if (a.shift) shift_stage<MSF, REALM>(tile, ldt, nv, a.cbtab, g, sh, a.sign_idx, a.transposed);
    for (int step = 0; step < (a.shift ? 0 : a.kcount); ++step) {
        const int k = a.kfirst + step * a.kstep;
        const double* phi_k = phi + size_t(k) * g.opdim * N;
        const double* cosh_k = coshT + size_t(k) * N;
        const double* sinh_k = sinhT + size_t(k) * N;
        // the first item's coefficients are fetched before the hopping stage so that their latency
        // overlaps with it
        PotCoef pc0;
        const int it0 = threadIdx.x;
        const int pos0 = it0 % N, grp0 = it0 / N;
        if (it0 < pot_items) pc0 = potential_coef<MSF>(phi_k, cosh_k, sinh_k, g, sinv[pos0], a.sign_idx, a.transposed);
        if (a.k_then_v && !a.skip_hopping) hopping_stage<MSF, REALM>(tile, ldt, nv, a.cbtab, g, sh, a.sign_idx, a.transposed);
        if (it0 < pot_items)
            for (int v = grp0; v < nv; v += sh.Gp) potential_apply<MSF>(tile + v * ldt, N, pos0, pc0);
        for (int it = it0 + blockDim.x; it < pot_items; it += blockDim.x) {
            const int pos = it % N, grp = it / N;
            const PotCoef pc = potential_coef<MSF>(phi_k, cosh_k, sinh_k, g, sinv[pos], a.sign_idx, a.transposed);
            for (int v = grp; v < nv; v += sh.Gp) potential_apply<MSF>(tile + v * ldt, N, pos, pc);
        }
        __syncthreads();
        if (!a.k_then_v && !a.skip_hopping) hopping_stage<MSF, REALM>(tile, ldt, nv, a.cbtab, g, sh, a.sign_idx, a.transposed);
    }

    // ---- store (in place unless an output matrix is given)
    if (a.out) A = a.out + size_t(b) * a.strideOut;
    if (!a.rows) {
        const double* cs = a.colscale ? a.colscale + size_t(b) * a.strideScale : nullptr;
        if (nv == TV) {
            double sc[TV];
#pragma unroll
            for (int v = 0; v < TV; ++v) sc[v] = cs ? cs[v0 + v] : 1.0;
            for (int e = threadIdx.x; e < D; e += blockDim.x) {
                const int pos = sperm[e];
                cplx buf[TV];
#pragma unroll
                for (int v = 0; v < TV; ++v) buf[v] = tile[v * ldt + pos];
#pragma unroll
                for (int v = 0; v < TV; ++v)
                    A[size_t(v0 + v) * D + e] = cs ? make_double2(buf[v].x * sc[v], buf[v].y * sc[v]) : buf[v];
            }
        } else {
            for (int v = 0; v < nv; ++v) {
                cplx* dst = A + size_t(v0 + v) * D;
                const cplx* t = tile + v * ldt;
                const double sc = cs ? cs[v0 + v] : 1.0;
                for (int e = threadIdx.x; e < D; e += blockDim.x) {
                    cplx val = t[sperm[e]];
                    if (cs) { val.x *= sc; val.y *= sc; }
                    dst[e] = val;
                }
            }
        }
    } else if (nv == TV) {
        for (int idx = threadIdx.x; idx < TV * D; idx += blockDim.x) {
            const int e = idx / TV, v = idx % TV;
            A[size_t(e) * D + v0 + v] = tile[v * ldt + sperm[e]];
        }
    } else {
        for (int idx = threadIdx.x; idx < nv * D; idx += blockDim.x) {
            const int e = idx / nv, v = idx - e * nv;
            A[size_t(e) * D + v0 + v] = tile[v * ldt + sperm[e]];
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Column tiles (left multiplies) through the bulk-copy engine.
//
// A column of the matrix is D contiguous elements; it is brought in as D / L pieces of one lattice row each
// (L sites = L * 16 bytes, `cp.async.bulk.shared.global` completing on an mbarrier: no registers, no LSU instructions)
// and stored the same way (`cp.async.bulk.global.shared`), so the tile keeps the NATURAL site order -- with the lattice
// rows padded to a stride of `rs` elements, rs == L / 2 (mod 4).  With that stride the four corners of consecutive
// plaquettes q = py * L/2 + px sit at 16-byte units 2 q + const (mod 8) for both subgroups, the vectors of the tile
// are 1 (mod 8) units apart, and a quarter warp = 4 consecutive plaquettes x 2 vectors reads eight distinct bank
// groups (only the x-wrapped corner of the odd subgroup deviates).  A thread keeps its plaquette matrix in registers
// and walks over every other vector of the tile.
// ------------------------------------------------------------------------------------------------
struct CbNat {
    int rs;        // stride of a lattice row in the tile
    int blk;       // stride of a band-spin block: L * rs
    int ldv;       // stride of a vector: 1 (mod 8)
    int half;      // L / 2
    int ngrp;      // groups of four plaquettes per band-spin block
    int Gp;        // vector groups of the potential stage
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MSF, bool REALM, int NV>
__device__ __forceinline__ void hopping_pass_nat(cplx* tile, const cplx* __restrict__ tab, const CbGeom& g,
                                                 const CbNat& nt, int sign_idx, int transposed, int pass) {
    const int items = MSF * nt.ngrp * 8;
    for (int item = threadIdx.x; item < items; item += blockDim.x) {
        const int qlo = item & 3, vpar = (item >> 2) & 1, rest = item >> 3;
        const int bs = rest / nt.ngrp;
        const int q = (rest - bs * nt.ngrp) * 4 + qlo;
        if (q >= g.nplaq) continue;
        const int band = bs & 1;
        PlaqMat<REALM> M;
        M.load(tab + ((size_t(band) * 2 + sign_idx) * 3 + pass) * 8 * g.nplaq + q, g.nplaq, transposed != 0);
        const int sg = pass >= 1 ? 0 : 1;
        const int py = q / nt.half, px = q - py * nt.half;
        const int x0 = 2 * px + sg, y0 = 2 * py + sg;
        const int x1 = x0 + 1 == g.L ? 0 : x0 + 1, y1 = y0 + 1 == g.L ? 0 : y0 + 1;
        const int base = bs * nt.blk;
        const int oi = base + y0 * nt.rs + x0, oj = base + y0 * nt.rs + x1;
        const int ok = base + y1 * nt.rs + x0, ol = base + y1 * nt.rs + x1;
#pragma unroll
        for (int v = 0; v < NV; v += 2) {
            cplx* t = tile + (v + vpar) * nt.ldv;
            cplx a = t[oi], b = t[oj], c = t[ok], e = t[ol];
            M.apply(a, b, c, e);
            t[oi] = a; t[oj] = b; t[ok] = c; t[ol] = e;
        }
    }
}

template <int MSF, bool REALM, int NV>
__global__ void __launch_bounds__(kCbMaxThreads) cb_mult_bulk_kernel(CbGeom g, CbLaunch a, CbNat nt) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    const int D = g.D, N = g.N, L = g.L;
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);
    const int b = blockIdx.y;
    const int v0 = blockIdx.x * NV;
    const int rows = D / L;                                 // lattice rows per vector
    const cplx* A = a.A + size_t(b) * a.strideA;
    const double* phi = a.phi + size_t(b) * a.stridePhi;
    const double* coshT = a.coshT + size_t(b) * a.strideTab;
    const double* sinhT = a.sinhT + size_t(b) * a.strideTab;
    const unsigned bar_s = smem_u32(&bar);
    const unsigned rowBytes = unsigned(L) * sizeof(cplx);

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(unsigned(NV) * D * 16u)
                     : "memory");
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NV * rows; c += blockDim.x) {
        const int v = c / rows, r = c - v * rows;
        const int bs = r / L, y = r - bs * L;
        const unsigned dst = smem_u32(tile + v * nt.ldv + bs * nt.blk + y * nt.rs);
        const cplx* src = A + size_t(v0 + v) * D + size_t(r) * L;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(rowBytes), "r"(bar_s) : "memory");
    }
    {
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar_s) : "memory");
        }
    }

    if (a.shift) {
        hopping_pass_nat<MSF, REALM, NV>(tile, a.cbtab, g, nt, a.sign_idx, a.transposed, 0);
        __syncthreads();
        hopping_pass_nat<MSF, REALM, NV>(tile, a.cbtab, g, nt, a.sign_idx, a.transposed, 2);
        __syncthreads();
    }
    const int pot_items = N * nt.Gp;
    for (int step = 0; step < (a.shift ? 0 : a.kcount); ++step) {
        const int k = a.kfirst + step * a.kstep;
        const double* phi_k = phi + size_t(k) * g.opdim * N;
        const double* cosh_k = coshT + size_t(k) * N;
        const double* sinh_k = sinhT + size_t(k) * N;
        PotCoef pc0;
        const int it0 = threadIdx.x;
        const int s0 = it0 % N, grp0 = it0 / N;
        if (it0 < pot_items) pc0 = potential_coef<MSF>(phi_k, cosh_k, sinh_k, g, s0, a.sign_idx, a.transposed);
        if (a.k_then_v && !a.skip_hopping) {
            hopping_pass_nat<MSF, REALM, NV>(tile, a.cbtab, g, nt, a.sign_idx, a.transposed, 0);
            __syncthreads();
            hopping_pass_nat<MSF, REALM, NV>(tile, a.cbtab, g, nt, a.sign_idx, a.transposed, 1);
            __syncthreads();
            hopping_pass_nat<MSF, REALM, NV>(tile, a.cbtab, g, nt, a.sign_idx, a.transposed, 0);
            __syncthreads();
        }
        if (it0 < pot_items) {
            const int pos0 = (s0 / L) * nt.rs + s0 % L;
            for (int v = grp0; v < NV; v += nt.Gp) potential_apply<MSF>(tile + v * nt.ldv, nt.blk, pos0, pc0);
        }
        for (int it = it0 + blockDim.x; it < pot_items; it += blockDim.x) {
            const int s = it % N, grp = it / N;
            const PotCoef pc = potential_coef<MSF>(phi_k, cosh_k, sinh_k, g, s, a.sign_idx, a.transposed);
            const int pos = (s / L) * nt.rs + s % L;
            for (int v = grp; v < NV; v += nt.Gp) potential_apply<MSF>(tile + v * nt.ldv, nt.blk, pos, pc);
        }
        __syncthreads();
        if (!a.k_then_v && !a.skip_hopping) {
            hopping_pass_nat<MSF, REALM, NV>(tile, a.cbtab, g, nt, a.sign_idx, a.transposed, 0);
            __syncthreads();
            hopping_pass_nat<MSF, REALM, NV>(tile, a.cbtab, g, nt, a.sign_idx, a.transposed, 1);
            __syncthreads();
            hopping_pass_nat<MSF, REALM, NV>(tile, a.cbtab, g, nt, a.sign_idx, a.transposed, 0);
            __syncthreads();
        }
    }
    if (a.colscale) {
        const double* cs = a.colscale + size_t(b) * a.strideScale;
        for (int idx = threadIdx.x; idx < NV * D; idx += blockDim.x) {
            const int v = idx / D, e = idx - v * D;
            const int r = e / L, x = e - r * L, bs = r / L, y = r - bs * L;
            cplx* p = tile + v * nt.ldv + bs * nt.blk + y * nt.rs + x;
            const double sc = cs[v0 + v];
            *p = make_double2(p->x * sc, p->y * sc);
        }
    }
    // ---- store: the generic-proxy writes of the tile become visible to the bulk-copy engine, then one piece per thread
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    cplx* O = a.out ? a.out + size_t(b) * a.strideOut : a.A + size_t(b) * a.strideA;
    for (int c = threadIdx.x; c < NV * rows; c += blockDim.x) {
        const int v = c / rows, r = c - v * rows;
        const int bs = r / L, y = r - bs * L;
        const unsigned src = smem_u32(tile + v * nt.ldv + bs * nt.blk + y * nt.rs);
        cplx* dst = O + size_t(v0 + v) * D + size_t(r) * L;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(rowBytes)
                     : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

int pow2_floor(int x) { int p = 1; while (2 * p <= x) p *= 2; return p; }

}  // namespace

// Column tiles through the bulk-copy engine (cb_mult_bulk_kernel); returns cudaErrorNotSupported when the shape does
// not qualify (the caller then uses the general kernel).
template <int NV>
static cudaError_t cb_launch_bulk(const CbGeom& g, const CbLaunch& a, cudaStream_t st) {
    if (a.rows || (g.L & 1) || g.D % NV != 0 || g.D % g.L != 0) return cudaErrorNotSupported;
    CbNat nt;
    nt.half = g.L / 2;
    nt.rs = g.L;
    while (nt.rs % 4 != nt.half % 4) ++nt.rs;
    nt.blk = g.L * nt.rs;
    nt.ldv = g.msf * nt.blk;
    while (nt.ldv % 8 != 1) ++nt.ldv;
    nt.ngrp = (g.nplaq + 3) / 4;
    const int items = g.msf * nt.ngrp * 8;
    int nthreads = std::min(kCbMaxThreads, ((items + 31) / 32) * 32);
    nthreads = std::max(nthreads, 64);
    nt.Gp = std::max(1, std::min(NV, pow2_floor(std::max(1, nthreads / g.N))));
    const size_t smem = size_t(NV) * nt.ldv * sizeof(cplx);
    if (smem > size_t(200) * 1024) return cudaErrorNotSupported;
    dim3 grid(g.D / NV, a.batch);
    const bool realm = a.real_tables != 0;
#define CB_LAUNCHB(MSF, RM)                                                                                   \
    {                                                                                                         \
        cudaError_t e = cudaFuncSetAttribute(cb_mult_bulk_kernel<MSF, RM, NV>,                                \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e != cudaSuccess) return e;                                                                       \
        launch_pdl(cb_mult_bulk_kernel<MSF, RM, NV>, dim3(grid), dim3(nthreads), smem, st, g, a, nt);         \
    }
    if (g.msf == 2) {
        if (realm) CB_LAUNCHB(2, true) else CB_LAUNCHB(2, false)
    } else {
        if (realm) CB_LAUNCHB(4, true) else CB_LAUNCHB(4, false)
    }
#undef CB_LAUNCHB
    return cudaGetLastError();
}

cudaError_t cb_launch(const CbGeom& g, const CbLaunch& a, cudaStream_t st) {
    {
        // DQMC_CB_BULK: 0 = general kernel only, 8 / 16 = vectors per bulk tile
        static const int bulk = std::getenv("DQMC_CB_BULK") ? std::atoi(std::getenv("DQMC_CB_BULK")) : 8;
        cudaError_t e = cudaErrorNotSupported;
        if (bulk == 8) e = cb_launch_bulk<8>(g, a, st);
        else if (bulk == 16) e = cb_launch_bulk<16>(g, a, st);
        if (e != cudaErrorNotSupported) return e;
    }
    CbShape sh;
    sh.half = g.L / 2;
    sh.quarter = g.N / 4;
    sh.pairs = g.msf * g.nplaq;
    // 8-vector tiles for small batches (twice the CTAs: the launch is latency bound), 16 otherwise
    static const int tvEnv = std::getenv("DQMC_CB_TILE") ? std::atoi(std::getenv("DQMC_CB_TILE")) : 0;
    const int tv = tvEnv ? tvEnv : (((g.D + 15) / 16) * a.batch < 2 * 148 ? 8 : 16);
    sh.G = std::max(1, std::min(tv, pow2_floor(std::max(1, kCbMaxThreads / sh.pairs))));
    int nthreads = std::min(kCbMaxThreads, ((sh.pairs * sh.G + 31) / 32) * 32);
    nthreads = std::max(nthreads, 64);
    sh.Gp = std::max(1, std::min(tv, pow2_floor(std::max(1, nthreads / g.N))));
    const size_t smem = size_t(tv) * (g.D + 1) * sizeof(cplx) + size_t(g.D + g.N) * sizeof(int);
    dim3 grid((g.D + tv - 1) / tv, a.batch);
    const bool realm = a.real_tables != 0;
#define CB_LAUNCH2(MSF, RM, TV)                                                                               \
    {                                                                                                         \
        cudaError_t e = cudaFuncSetAttribute(cb_mult_kernel<MSF, RM, TV>,                                     \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e != cudaSuccess) return e;                                                                       \
        launch_pdl(cb_mult_kernel<MSF, RM, TV>, dim3(grid), dim3(nthreads), smem, st, g, a, sh);              \
    }
#define CB_LAUNCH(MSF, RM) { if (tv == 8) CB_LAUNCH2(MSF, RM, 8) else CB_LAUNCH2(MSF, RM, 16) }
    if (g.msf == 2) {
        if (realm) CB_LAUNCH(2, true) else CB_LAUNCH(2, false)
    } else {
        if (realm) CB_LAUNCH(4, true) else CB_LAUNCH(4, false)
    }
#undef CB_LAUNCH
#undef CB_LAUNCH2
    return cudaGetLastError();
}

}  // namespace dqmc
