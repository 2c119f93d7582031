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
// independently, a right multiply on every ROW, so a CTA stages a tile of TV vectors (each of
// length D = msf*N) in shared memory, applies all slices of the chain there, and writes the tile
// back: HBM traffic is one read + one write of the matrix per chain, whatever its length.  The
// reference recomputes every checkerboard block 2x (O(2)) / 3x (O(3)) and copies the matrix per
// pass; here all band blocks of a vector are transformed once.
//
// Bound: HBM (algorithmic bytes = 2 * D^2 * 16 per launch and matrix).
#include "dqmc_internal.h"

#include <cmath>
#include <complex>

namespace dqmc {

// ------------------------------------------------------------------------------------------------
// host: plaquette tables
// ------------------------------------------------------------------------------------------------
// index: ((((band*2 + sign_idx)*2 + transposed)*2 + pass) * nplaq + q) * 16 + r*4 + c
// pass 0: subgroup 1, half step; pass 1: subgroup 0, full step times e^{-+dtau mu_band}.
int cb_table_count(const CbGeom& g) { return 2 * 2 * 2 * 2 * g.nplaq * 16; }

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

void cb_build_tables(const dqmc_params& p, std::vector<cplx>& out) {
    const int L = p.L, N = L * L, nplaq = N / 4, half = L / 2;
    out.assign(size_t(2) * 2 * 2 * 2 * nplaq * 16, make_double2(0, 0));
    const double pi = M_PI;
    for (int band = 0; band < 2; ++band) {
        const double th = band == 0 ? p.txhor : p.tyhor;
        const double tv = band == 0 ? p.txver : p.tyver;
        const double mu = band == 0 ? p.mux : p.muy;
        for (int si = 0; si < 2; ++si) {
            const double sign = si == 0 ? -1.0 : +1.0;
            for (int pass = 0; pass < 2; ++pass) {
                const int subgroup = pass == 0 ? 1 : 0;
                const double pf = sign * p.dtau * (pass == 0 ? 0.5 : 1.0);
                const double ovfac = pass == 1 ? std::exp(-sign * p.dtau * mu) : 1.0;
                for (int q = 0; q < nplaq; ++q) {
                    const int i1 = 2 * (q % half) + subgroup;     // x
                    const int i2 = 2 * (q / half) + subgroup;     // y
                    double hh = th, hv = tv;
                    if ((p.bc == 1 || p.bc == 3) && i1 == L - 1) hh = -hh;
                    if ((p.bc == 2 || p.bc == 3) && i2 == L - 1) hv = -hv;
                    zc M[16];
                    if (!p.weakZflux) {
                        // (2x2 horizontal) x (2x2 vertical), detsdwopdim.cpp:1817-1820
                        const double chh = std::cosh(pf * hh), shh = std::sinh(-pf * hh);
                        const double chv = std::cosh(pf * hv), shv = std::sinh(-pf * hv);
                        const double a = chh * chv, b = chv * shh, c = chh * shv, d = shh * shv;
                        const double m4[16] = {a, b, c, d, b, a, d, c, c, d, a, b, d, c, b, a};
                        for (int i = 0; i < 16; ++i) M[i] = m4[i];
                    } else {
                        // Peierls phases, detsdwopdim.cpp:1647-1666; zmag = +1/N for both bands
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
                            for (int c = 0; c < 4; ++c)
                                Hs[r * 4 + c] = -pf * (H[r * 4 + c] + std::conj(H[c * 4 + r]));
                        expm4(Hs, M);
                    }
                    for (int tr = 0; tr < 2; ++tr) {
                        size_t base = ((((size_t(band) * 2 + si) * 2 + tr) * 2 + pass) * nplaq + q) * 16;
                        for (int r = 0; r < 4; ++r)
                            for (int c = 0; c < 4; ++c) {
                                zc v = (tr ? M[c * 4 + r] : M[r * 4 + c]) * ovfac;
                                out[base + r * 4 + c] = make_double2(v.real(), v.imag());
                            }
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

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {   // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

constexpr int kCbThreads = 256;
constexpr int kCbTileVecs = 8;

template <int MSF>
__device__ __forceinline__ void hopping_pass(cplx* tile, int ldt, int nv, const cplx* __restrict__ tab,
                                             const CbGeom& g, int sign_idx, int transposed, int pass) {
    const int pairs = MSF * g.nplaq;
    const int nthreads = blockDim.x;
    const int tid = threadIdx.x;
    const int half = g.L / 2;
    const int subgroup = pass == 0 ? 1 : 0;
    int p, pstep, v0, vstep;
    if (pairs <= nthreads) {
        const int ngrp = nthreads / pairs;
        p = tid % pairs;
        pstep = pairs;                 // single iteration
        v0 = tid / pairs;
        vstep = ngrp;
        if (v0 >= ngrp) return;
    } else {
        p = tid;
        pstep = nthreads;
        v0 = 0;
        vstep = 1;
    }
    for (; p < pairs; p += pstep) {
        const int bs = p / g.nplaq;
        const int q = p - bs * g.nplaq;
        const int band = bs & 1;       // XUP, YDOWN, XDOWN, YUP -> x, y, x, y
        const cplx* M = tab + ((((size_t(band) * 2 + sign_idx) * 2 + transposed) * 2 + pass) * g.nplaq + q) * 16;
        cplx mm[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) mm[i] = __ldg(M + i);
        const int i1 = 2 * (q % half) + subgroup;
        const int i2 = 2 * (q / half) + subgroup;
        const int i1p = (i1 + 1 == g.L) ? 0 : i1 + 1;
        const int i2p = (i2 + 1 == g.L) ? 0 : i2 + 1;
        const int base = bs * g.N;
        const int si = base + i2 * g.L + i1;
        const int sj = base + i2 * g.L + i1p;
        const int sk = base + i2p * g.L + i1;
        const int sl = base + i2p * g.L + i1p;
        for (int v = v0; v < nv; v += vstep) {
            cplx* t = tile + v * ldt;
            const cplx a = t[si], b = t[sj], c = t[sk], d = t[sl];
            cplx r0 = cmul(mm[0], a), r1 = cmul(mm[4], a), r2 = cmul(mm[8], a), r3 = cmul(mm[12], a);
            r0 = cfma(mm[1], b, r0); r1 = cfma(mm[5], b, r1); r2 = cfma(mm[9], b, r2); r3 = cfma(mm[13], b, r3);
            r0 = cfma(mm[2], c, r0); r1 = cfma(mm[6], c, r1); r2 = cfma(mm[10], c, r2); r3 = cfma(mm[14], c, r3);
            r0 = cfma(mm[3], d, r0); r1 = cfma(mm[7], d, r1); r2 = cfma(mm[11], d, r2); r3 = cfma(mm[15], d, r3);
            t[si] = r0; t[sj] = r1; t[sk] = r2; t[sl] = r3;
        }
    }
}

template <int MSF>
__device__ __forceinline__ void hopping_stage(cplx* tile, int ldt, int nv, const cplx* __restrict__ tab,
                                              const CbGeom& g, int sign_idx, int transposed) {
    hopping_pass<MSF>(tile, ldt, nv, tab, g, sign_idx, transposed, 0);
    __syncthreads();
    hopping_pass<MSF>(tile, ldt, nv, tab, g, sign_idx, transposed, 1);
    __syncthreads();
    hopping_pass<MSF>(tile, ldt, nv, tab, g, sign_idx, transposed, 0);
    __syncthreads();
}

// per-site blocks of e^{sign*dtau*V} (evMatrix, detsdwopdim.cpp:3188-3229 with cdwU == 0)
template <int MSF>
__device__ __forceinline__ void potential_coeffs(cplx* ev, const double* __restrict__ phi_k,
                                                 const double* __restrict__ cosh_k,
                                                 const double* __restrict__ sinh_k, const CbGeom& g,
                                                 int sign_idx, int transposed) {
    const double sg = sign_idx == 0 ? -1.0 : 1.0;
    for (int s = threadIdx.x; s < g.N; s += blockDim.x) {
        const double c = cosh_k[s], x = sinh_k[s] * sg;
        const double p0 = phi_k[s];
        const double p1 = g.opdim > 1 ? phi_k[g.N + s] : 0.0;
        cplx e01 = make_double2(x * p0, -x * p1);     // sign * x * (phi0 - i phi1)
        cplx e10 = make_double2(x * p0, x * p1);
        if (transposed) { cplx t = e01; e01 = e10; e10 = t; }
        if (MSF == 2) {
            ev[0 * g.N + s] = make_double2(c, 0);
            ev[1 * g.N + s] = e01;
            ev[2 * g.N + s] = e10;
            ev[3 * g.N + s] = make_double2(c, 0);
        } else {
            const double p2 = phi_k[2 * g.N + s];
            const cplx z = make_double2(0, 0);
            const cplx cc = make_double2(c, 0);
            const cplx a = make_double2(x * p2, 0);       //  sign * phi2 * x
            const cplx ma = make_double2(-x * p2, 0);
            // rows of e^{sign dtau V}: see detsdwopdim.cpp:3198-3224
            cplx E[16] = {cc, e01, z, a,
                          e10, cc, ma, z,
                          z, ma, cc, e10,
                          a, z, e01, cc};
            // (2,3) = sign x (phi0 + i phi1), (3,2) = sign x (phi0 - i phi1); in the transposed case
            // e01/e10 were swapped above, which is exactly the transpose of those four entries too.
#pragma unroll
            for (int i = 0; i < 16; ++i) ev[i * g.N + s] = E[i];
        }
    }
}

template <int MSF>
__device__ __forceinline__ void potential_stage(cplx* tile, int ldt, int nv, const cplx* ev, const CbGeom& g) {
    const int items = nv * g.N;
    for (int idx = threadIdx.x; idx < items; idx += blockDim.x) {
        const int v = idx / g.N;
        const int s = idx - v * g.N;
        cplx* t = tile + v * ldt;
        cplx old[MSF];
#pragma unroll
        for (int c = 0; c < MSF; ++c) old[c] = t[c * g.N + s];
#pragma unroll
        for (int r = 0; r < MSF; ++r) {
            cplx acc = make_double2(0, 0);
#pragma unroll
            for (int c = 0; c < MSF; ++c) acc = cfma(ev[(r * MSF + c) * g.N + s], old[c], acc);
            t[r * g.N + s] = acc;
        }
    }
}

template <int MSF>
__global__ void __launch_bounds__(kCbThreads) cb_mult_kernel(CbGeom g, CbLaunch a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = g.D;
    const int ldt = D + 1;
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);
    cplx* ev = tile + kCbTileVecs * ldt;

    const int b = blockIdx.y;
    const int v0 = blockIdx.x * kCbTileVecs;
    const int nv = min(kCbTileVecs, D - v0);
    cplx* A = a.A + size_t(b) * a.strideA;
    const double* phi = a.phi + size_t(b) * a.stridePhi;
    const double* coshT = a.coshT + size_t(b) * a.strideTab;
    const double* sinhT = a.sinhT + size_t(b) * a.strideTab;

    // ---- load the tile (coalesced along the contiguous direction of the column-major matrix)
    if (!a.rows) {
        for (int idx = threadIdx.x; idx < nv * D; idx += blockDim.x) {
            const int v = idx / D, e = idx - v * D;
            tile[v * ldt + e] = A[size_t(v0 + v) * D + e];
        }
    } else {
        for (int idx = threadIdx.x; idx < nv * D; idx += blockDim.x) {
            const int e = idx / nv, v = idx - e * nv;
            tile[v * ldt + e] = A[size_t(e) * D + v0 + v];
        }
    }
    __syncthreads();

    for (int step = 0; step < a.kcount; ++step) {
        const int k = a.kfirst + step * a.kstep;
        potential_coeffs<MSF>(ev, phi + size_t(k) * g.opdim * g.N, coshT + size_t(k) * g.N,
                              sinhT + size_t(k) * g.N, g, a.sign_idx, a.transposed);
        if (a.k_then_v) {
            hopping_stage<MSF>(tile, ldt, nv, a.cbtab, g, a.sign_idx, a.transposed);   // ends with a sync
            potential_stage<MSF>(tile, ldt, nv, ev, g);
            __syncthreads();
        } else {
            __syncthreads();
            potential_stage<MSF>(tile, ldt, nv, ev, g);
            __syncthreads();
            hopping_stage<MSF>(tile, ldt, nv, a.cbtab, g, a.sign_idx, a.transposed);
        }
    }

    // ---- store
    if (!a.rows) {
        const double* cs = a.colscale ? a.colscale + size_t(b) * a.strideScale : nullptr;
        for (int idx = threadIdx.x; idx < nv * D; idx += blockDim.x) {
            const int v = idx / D, e = idx - v * D;
            cplx val = tile[v * ldt + e];
            if (cs) { const double sc = cs[v0 + v]; val.x *= sc; val.y *= sc; }
            A[size_t(v0 + v) * D + e] = val;
        }
    } else {
        for (int idx = threadIdx.x; idx < nv * D; idx += blockDim.x) {
            const int e = idx / nv, v = idx - e * nv;
            A[size_t(e) * D + v0 + v] = tile[v * ldt + e];
        }
    }
}

}  // namespace

cudaError_t cb_launch(const CbGeom& g, const CbLaunch& a, cudaStream_t st) {
    const size_t smem = (size_t(kCbTileVecs) * (g.D + 1) + size_t(g.msf) * g.msf * g.N) * sizeof(cplx);
    dim3 grid((g.D + kCbTileVecs - 1) / kCbTileVecs, a.batch);
    cudaError_t e;
    if (g.msf == 2) {
        e = cudaFuncSetAttribute(cb_mult_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cb_mult_kernel<2><<<grid, kCbThreads, smem, st>>>(g, a);
    } else {
        e = cudaFuncSetAttribute(cb_mult_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cb_mult_kernel<4><<<grid, kCbThreads, smem, st>>>(g, a);
    }
    return cudaGetLastError();
}

}  // namespace dqmc
