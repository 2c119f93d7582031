// Batched complex GEMM on the FP64 tensor-core path of sm_100a (DMMA, mma.sync m8n8k4.f64).
//
// Replaces every dense `*` on D x D matrices in the stabilisation path of the reference
// (Armadillo -> zgemm): the UdV chain products (detmodel.h:699-700, 983-987, 1141-1143), the five
// products of greenFromUdV (detmodel.h:784-815) and, for DetHubbard, the dense B-matrix products
// (dethubbard.h:299-337).  tcgen05 has no f64 kind; on Blackwell FP64 matrix math is the warp-level
// mma.sync DMMA path (SASS: DMMA.8x8x4).
//
//   C = alpha * rowscale .* ( op(A) * diag(kscale) * op(B) ) .* colscale + beta * C,   op = N | conj-transpose
//
// Complex arithmetic = 4 real DMMAs per k-step on planar (re / im) shared-memory tiles; the planar
// split, the conjugation, the transposition and the three diagonal scalings are fused into the
// global->shared load and the epilogue, so no intermediate matrices are materialised.
// One CTA computes a (8*MB*WM) x (8*NB*WN) tile, one warp an (8*MB) x (8*NB) sub-tile of DMMA blocks.
// Bound: FP64 tensor pipe.  Algorithmic flops: 8 * M * N * K per matrix.
#include "dqmc_internal.h"

#include <cstdlib>

namespace dqmc {
namespace {

constexpr int BK = 8;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// BKT = depth of a k-tile: 8 for batches that fill the SMs (other CTAs hide the load latency of the one-tile
// register prefetch), 32 for small batches, where a CTA is alone on its SM and every k-tile costs a global round trip
template <int WM, int WN, int MB, int NB, int BKT = 8>
struct GemmCfg {
    static constexpr int BK = BKT;
    static constexpr int TM = 8 * MB * WM, TN = 8 * NB * WN;
    // Leading dimensions == 4 or 12 (mod 16): the 16 lanes of a half warp (grp 0..3 x t4 0..3) read the doubles
    // t4 * LD + grp, which then fall into 16 different 8-byte bank pairs (with LD == 8 (mod 16) rows t4 and t4 + 2
    // collide: 2-way conflicts on every fragment load)
    static constexpr int LDM = TM + 4, LDN = TN + 4;
    static constexpr int NT = 32 * WM * WN;
    static constexpr int A_PER = (TM * BK + NT - 1) / NT;
    static constexpr int B_PER = (TN * BK + NT - 1) / NT;
    static constexpr int SMEM_DOUBLES = 2 * (2 * BK * LDM + 2 * BK * LDN);   // two stages, re+im
};

template <int WM, int WN, int MB, int NB, int BKT>
__global__ void __launch_bounds__(32 * WM * WN) zgemm_dmma_kernel(GemmArgs g) {
    pdl_enter();
    typedef GemmCfg<WM, WN, MB, NB, BKT> C;
    constexpr int BK = BKT;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int grp = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.z;
    const int m0 = blockIdx.x * C::TM, n0 = blockIdx.y * C::TN;

    const cplx* __restrict__ A = g.A + size_t(b) * g.strideA;
    const cplx* __restrict__ B = g.B + size_t(b) * g.strideB;
    const double* __restrict__ ks = g.kscale ? g.kscale + size_t(b) * g.strideK : nullptr;

    double acc_re[MB][NB][2], acc_im[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            acc_re[i][j][0] = acc_re[i][j][1] = 0.0;
            acc_im[i][j][0] = acc_im[i][j][1] = 0.0;
        }

    cplx ra[C::A_PER], rb[C::B_PER];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < C::A_PER; ++i) {
            const int idx = tid + i * C::NT;
            cplx v = make_double2(0, 0);
            if (idx < C::TM * BK) {
                int mm, kk;
                if (!g.transa) { mm = idx % C::TM; kk = idx / C::TM; }
                else { kk = idx % BK; mm = idx / BK; }
                const int gm = m0 + mm, gk = k0 + kk;
                if (gm < g.M && gk < g.K) {
                    if (!g.transa) v = A[size_t(gk) * g.lda + gm];
                    else { v = A[size_t(gm) * g.lda + gk]; v.y = -v.y; }
                    if (ks) { const double sc = ks[gk]; v.x *= sc; v.y *= sc; }
                }
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < C::B_PER; ++i) {
            const int idx = tid + i * C::NT;
            cplx v = make_double2(0, 0);
            if (idx < C::TN * BK) {
                int nn, kk;
                if (!g.transb) { kk = idx % BK; nn = idx / BK; }
                else { nn = idx % C::TN; kk = idx / C::TN; }
                const int gn = n0 + nn, gk = k0 + kk;
                if (gn < g.N && gk < g.K) {
                    if (!g.transb) v = B[size_t(gn) * g.ldb + gk];
                    else { v = B[size_t(gk) * g.ldb + gn]; v.y = -v.y; }
                }
            }
            rb[i] = v;
        }
    };
    auto store_tiles = [&](int stage) {
        double* As_re = smem + stage * (C::SMEM_DOUBLES / 2);
        double* As_im = As_re + BK * C::LDM;
        double* Bs_re = As_im + BK * C::LDM;
        double* Bs_im = Bs_re + BK * C::LDN;
#pragma unroll
        for (int i = 0; i < C::A_PER; ++i) {
            const int idx = tid + i * C::NT;
            if (idx < C::TM * BK) {
                int mm, kk;
                if (!g.transa) { mm = idx % C::TM; kk = idx / C::TM; }
                else { kk = idx % BK; mm = idx / BK; }
                As_re[kk * C::LDM + mm] = ra[i].x;
                As_im[kk * C::LDM + mm] = ra[i].y;
            }
        }
#pragma unroll
        for (int i = 0; i < C::B_PER; ++i) {
            const int idx = tid + i * C::NT;
            if (idx < C::TN * BK) {
                int nn, kk;
                if (!g.transb) { kk = idx % BK; nn = idx / BK; }
                else { nn = idx % C::TN; kk = idx / C::TN; }
                Bs_re[kk * C::LDN + nn] = rb[i].x;
                Bs_im[kk * C::LDN + nn] = rb[i].y;
            }
        }
    };

    const int nk = (g.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();

    for (int kt = 0; kt < nk; ++kt) {
        const int stage = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);          // global loads in flight during the math
        const double* As_re = smem + stage * (C::SMEM_DOUBLES / 2);
        const double* As_im = As_re + BK * C::LDM;
        const double* Bs_re = As_im + BK * C::LDM;
        const double* Bs_im = Bs_re + BK * C::LDN;
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double ar[MB], ai[MB], nai[MB], br[NB], bi[NB];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                const int off = (k4 + t4) * C::LDM + wm * (8 * MB) + mb * 8 + grp;
                ar[mb] = As_re[off];
                ai[mb] = As_im[off];
                nai[mb] = -ai[mb];
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const int off = (k4 + t4) * C::LDN + wn * (8 * NB) + nb * 8 + grp;
                br[nb] = Bs_re[off];
                bi[nb] = Bs_im[off];
            }
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    dmma(acc_re[mb][nb][0], acc_re[mb][nb][1], ar[mb], br[nb]);
                    dmma(acc_re[mb][nb][0], acc_re[mb][nb][1], nai[mb], bi[nb]);
                    dmma(acc_im[mb][nb][0], acc_im[mb][nb][1], ar[mb], bi[nb]);
                    dmma(acc_im[mb][nb][0], acc_im[mb][nb][1], ai[mb], br[nb]);
                }
        }
        if (kt + 1 < nk) store_tiles(stage ^ 1);
        __syncthreads();
    }

    // ---- epilogue
    cplx* __restrict__ Cm = g.C + size_t(b) * g.strideC;
    const double* __restrict__ rs = g.rowscale ? g.rowscale + size_t(b) * g.strideRow : nullptr;
    const double* __restrict__ cs = g.colscale ? g.colscale + size_t(b) * g.strideCol : nullptr;
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        const int gm = m0 + wm * (8 * MB) + mb * 8 + grp;
        if (gm >= g.M) continue;
        const double rsc = rs ? rs[gm] : 1.0;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = n0 + wn * (8 * NB) + nb * 8 + 2 * t4 + e;
                if (gn >= g.N) continue;
                const double sc = g.alpha * rsc * (cs ? cs[gn] : 1.0);
                cplx v = make_double2(acc_re[mb][nb][e] * sc, acc_im[mb][nb][e] * sc);
                cplx* dst = Cm + size_t(gn) * g.ldc + gm;
                if (g.beta != 0.0) { const cplx o = *dst; v.x += o.x; v.y += o.y; }
                *dst = v;
            }
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Batched complex GEMM, second generation: the same tiling and arithmetic as zgemm_dmma_kernel, but the operand tiles
// travel global -> shared with cp.async (16-byte LDGSTS) through a kStages-deep ring, in their interleaved complex
// layout; conjugation and the k-scaling are applied to the fragments in registers, one LDS.128 per complex element.
// The ring keeps kStages - 1 k-tiles in flight, which hides the L2 / HBM latency that the register-staged version
// exposed between its 8-deep k-tiles (long-scoreboard stalls on the staging stores, profiles/prof_gemm_r02_*).
// ------------------------------------------------------------------------------------------------
constexpr int kStages = 4;

template <int WM, int WN, int MB, int NB>
__global__ void __launch_bounds__(32 * WM * WN) zgemm_dmma2_kernel(GemmArgs g) {
    pdl_enter();
    constexpr int TM = 8 * MB * WM, TN = 8 * NB * WN, LDA = TM + 2, LDB = TN + 2, NT = 32 * WM * WN;
    constexpr int STAGE = BK * (LDA + LDB);                 // cplx elements per stage
    extern __shared__ __align__(16) double smem[];
    cplx* ring = reinterpret_cast<cplx*>(smem);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int grp = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.z;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;

    const cplx* __restrict__ A = g.A + size_t(b) * g.strideA;
    const cplx* __restrict__ B = g.B + size_t(b) * g.strideB;
    const double* __restrict__ ks = g.kscale ? g.kscale + size_t(b) * g.strideK : nullptr;
    const cplx zero = make_double2(0, 0);

    double acc_re[MB][NB][2], acc_im[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            acc_re[i][j][0] = acc_re[i][j][1] = 0.0;
            acc_im[i][j][0] = acc_im[i][j][1] = 0.0;
        }

    auto load_tile = [&](int kt) {
        cplx* As = ring + (kt % kStages) * STAGE;
        cplx* Bs = As + BK * LDA;
        const int k0 = kt * BK;
        for (int idx = tid; idx < TM * BK; idx += NT) {
            int mm, kk;
            if (!g.transa) { mm = idx % TM; kk = idx / TM; }
            else { kk = idx % BK; mm = idx / BK; }
            const int gm = m0 + mm, gk = k0 + kk;
            cplx* dst = As + kk * LDA + mm;
            if (gm < g.M && gk < g.K) cp_async16(dst, g.transa ? A + size_t(gm) * g.lda + gk : A + size_t(gk) * g.lda + gm);
            else *dst = zero;
        }
        for (int idx = tid; idx < TN * BK; idx += NT) {
            int nn, kk;
            if (!g.transb) { kk = idx % BK; nn = idx / BK; }
            else { nn = idx % TN; kk = idx / TN; }
            const int gn = n0 + nn, gk = k0 + kk;
            cplx* dst = Bs + kk * LDB + nn;
            if (gn < g.N && gk < g.K) cp_async16(dst, g.transb ? B + size_t(gk) * g.ldb + gn : B + size_t(gn) * g.ldb + gk);
            else *dst = zero;
        }
    };

    const int nk = (g.K + BK - 1) / BK;
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (s < nk) load_tile(s);
        cp_async_commit();
    }
    const double asign = g.transa ? -1.0 : 1.0, bsign = g.transb ? -1.0 : 1.0;
    for (int kt = 0; kt < nk; ++kt) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
        __syncthreads();                                    // tile kt has landed; everyone is done with tile kt - 1
        if (kt + kStages - 1 < nk) load_tile(kt + kStages - 1);
        cp_async_commit();
        const cplx* As = ring + (kt % kStages) * STAGE;
        const cplx* Bs = As + BK * LDA;
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double ar[MB], ai[MB], nai[MB], br[NB], bi[NB];
            const int gk = kt * BK + k4 + t4;
            const double sc = (ks && gk < g.K) ? __ldg(ks + gk) : 1.0;
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                const cplx a = As[(k4 + t4) * LDA + wm * (8 * MB) + mb * 8 + grp];
                ar[mb] = sc * a.x;
                ai[mb] = sc * asign * a.y;
                nai[mb] = -ai[mb];
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const cplx bb = Bs[(k4 + t4) * LDB + wn * (8 * NB) + nb * 8 + grp];
                br[nb] = bb.x;
                bi[nb] = bsign * bb.y;
            }
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    dmma(acc_re[mb][nb][0], acc_re[mb][nb][1], ar[mb], br[nb]);
                    dmma(acc_re[mb][nb][0], acc_re[mb][nb][1], nai[mb], bi[nb]);
                    dmma(acc_im[mb][nb][0], acc_im[mb][nb][1], ar[mb], bi[nb]);
                    dmma(acc_im[mb][nb][0], acc_im[mb][nb][1], ai[mb], br[nb]);
                }
        }
    }

    // ---- epilogue
    cplx* __restrict__ Cm = g.C + size_t(b) * g.strideC;
    const double* __restrict__ rs = g.rowscale ? g.rowscale + size_t(b) * g.strideRow : nullptr;
    const double* __restrict__ cs = g.colscale ? g.colscale + size_t(b) * g.strideCol : nullptr;
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        const int gm = m0 + wm * (8 * MB) + mb * 8 + grp;
        if (gm >= g.M) continue;
        const double rsc = rs ? rs[gm] : 1.0;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = n0 + wn * (8 * NB) + nb * 8 + 2 * t4 + e;
                if (gn >= g.N) continue;
                const double sc = g.alpha * rsc * (cs ? cs[gn] : 1.0);
                cplx v = make_double2(acc_re[mb][nb][e] * sc, acc_im[mb][nb][e] * sc);
                cplx* dst = Cm + size_t(gn) * g.ldc + gm;
                if (g.beta != 0.0) { const cplx o = *dst; v.x += o.x; v.y += o.y; }
                *dst = v;
            }
    }
}

// ------------------------------------------------------------------------------------------------
// Rank-K update  C += alpha * A(M x K) * B(K x N)  for small K (one shared-memory pass per 32): the delayed-update flush G += X*Y
// (detsdwopdim.cpp:3156), the block-reflector updates of the QR (A2 -= V W) and the triangular
// solve.  The whole K extent of both panels is staged in shared memory once (one barrier), the
// accumulators START from the C tile (its global loads are in flight while the panels are staged)
// and the epilogue is a plain store.  `kvec` (optional) gives a per-matrix K <= g.K; K = 0 skips
// the matrix.  Algorithmic traffic is dominated by C (read + write): the kernel is HBM/L2 bound.
// ------------------------------------------------------------------------------------------------
constexpr int KT = 32;

template <int WM, int WN, int MB, int NB>
__global__ void __launch_bounds__(32 * WM * WN, (WM * WN <= 4 ? 3 : (WM * WN <= 8 ? 2 : 1)))
zgemm_rank_update_kernel(GemmArgs g) {
    pdl_enter();
    constexpr int TM = 8 * MB * WM, TN = 8 * NB * WN, LDM = TM + 4, LDN = TN + 4, NT = 32 * WM * WN;
    extern __shared__ __align__(16) double smem[];
    double* As_re = smem;
    double* As_im = As_re + KT * LDM;
    double* Bs_re = As_im + KT * LDM;
    double* Bs_im = Bs_re + KT * LDN;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int grp = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.z;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int K = g.kvec ? min(g.kvec[b], g.K) : g.K;
    if (K <= 0) return;

    const cplx* __restrict__ A = g.A + size_t(b) * g.strideA;
    const cplx* __restrict__ B = g.B + size_t(b) * g.strideB;
    cplx* __restrict__ Cm = g.C + size_t(b) * g.strideC;

    // accumulators start from C
    double acc_re[MB][NB][2], acc_im[MB][NB][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        const int gm = m0 + wm * (8 * MB) + mb * 8 + grp;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = n0 + wn * (8 * NB) + nb * 8 + 2 * t4 + e;
                cplx c = make_double2(0, 0);
                if (gm < g.M && gn < g.N) c = Cm[size_t(gn) * g.ldc + gm];
                acc_re[mb][nb][e] = c.x;
                acc_im[mb][nb][e] = c.y;
            }
    }
    // stage alpha * A[m0: , kc:kc+KT] and B[kc:kc+KT, n0: ] (planar, zero padded to a multiple of 4 in k);
    // K <= KT (the common case) needs a single pass and a single barrier
    for (int kc = 0; kc < K; kc += KT) {
        const int Kc = min(KT, K - kc);
        const int K4 = (Kc + 3) & ~3;
        if (kc > 0) __syncthreads();
        for (int idx = tid; idx < TM * K4; idx += NT) {
            const int mm = idx % TM, kk = idx / TM;
            const int gm = m0 + mm;
            cplx v = make_double2(0, 0);
            if (gm < g.M && kk < Kc) v = A[size_t(kc + kk) * g.lda + gm];
            As_re[kk * LDM + mm] = g.alpha * v.x;
            As_im[kk * LDM + mm] = g.alpha * v.y;
        }
        if (g.b_kmajor) {
            for (int idx = tid; idx < TN * K4; idx += NT) {
                const int nn = idx % TN, kk = idx / TN;
                const int gn = n0 + nn;
                cplx v = make_double2(0, 0);
                if (gn < g.N && kk < Kc) v = B[size_t(kc + kk) * g.ldb + gn];
                Bs_re[kk * LDN + nn] = v.x;
                Bs_im[kk * LDN + nn] = v.y;
            }
        } else {
            for (int idx = tid; idx < TN * K4; idx += NT) {
                const int kk = idx % K4, nn = idx / K4;
                const int gn = n0 + nn;
                cplx v = make_double2(0, 0);
                if (gn < g.N && kk < Kc) v = B[size_t(gn) * g.ldb + kc + kk];
                Bs_re[kk * LDN + nn] = v.x;
                Bs_im[kk * LDN + nn] = v.y;
            }
        }
        __syncthreads();
        for (int k4 = 0; k4 < K4; k4 += 4) {
            double ar[MB], ai[MB], nai[MB], br[NB], bi[NB];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                const int off = (k4 + t4) * LDM + wm * (8 * MB) + mb * 8 + grp;
                ar[mb] = As_re[off];
                ai[mb] = As_im[off];
                nai[mb] = -ai[mb];
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const int off = (k4 + t4) * LDN + wn * (8 * NB) + nb * 8 + grp;
                br[nb] = Bs_re[off];
                bi[nb] = Bs_im[off];
            }
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    dmma(acc_re[mb][nb][0], acc_re[mb][nb][1], ar[mb], br[nb]);
                    dmma(acc_re[mb][nb][0], acc_re[mb][nb][1], nai[mb], bi[nb]);
                    dmma(acc_im[mb][nb][0], acc_im[mb][nb][1], ar[mb], bi[nb]);
                    dmma(acc_im[mb][nb][0], acc_im[mb][nb][1], ai[mb], br[nb]);
                }
        }
    }
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        const int gm = m0 + wm * (8 * MB) + mb * 8 + grp;
        if (gm >= g.M) continue;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = n0 + wn * (8 * NB) + nb * 8 + 2 * t4 + e;
                if (gn >= g.N) continue;
                Cm[size_t(gn) * g.ldc + gm] = make_double2(acc_re[mb][nb][e], acc_im[mb][nb][e]);
            }
    }
}

// ------------------------------------------------------------------------------------------------
// Rank-K update, second generation:  C += alpha * A(M x K) * B(K x N),  K <= 32 per pass.
// The operand panels go from global to shared memory with cp.async (16-byte LDGSTS, no register staging) in their
// natural interleaved complex layout and stay in flight while the accumulators are loaded from C; fragments are read
// with one LDS.128 per complex element (leading dimensions == 2 (mod 8) complex elements: the 4 k-rows x 2 elements
// a quarter warp reads fall into 8 distinct 16-byte bank groups).  alpha is applied to the A fragments in registers.
// ------------------------------------------------------------------------------------------------

template <int WM, int WN, int MB, int NB>
__global__ void __launch_bounds__(32 * WM * WN, (WM * WN <= 4 ? 3 : (WM * WN <= 8 ? (MB * NB <= 6 ? 3 : 2) : 1)))
zgemm_rank_update2_kernel(GemmArgs g) {
    pdl_enter();
    constexpr int TM = 8 * MB * WM, TN = 8 * NB * WN, LDA = TM + 2, LDB = TN + 2, NT = 32 * WM * WN;
    static_assert(LDA % 8 == 2 && LDB % 8 == 2, "fragment loads need leading dimensions == 2 (mod 8)");
    extern __shared__ __align__(16) double smem[];
    cplx* As = reinterpret_cast<cplx*>(smem);              // [KT][LDA]
    cplx* Bs = As + KT * LDA;                              // [KT][LDB]
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int grp = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.z;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int K = g.kvec ? min(g.kvec[b], g.K) : g.K;
    if (K <= 0) return;

    const cplx* __restrict__ A = g.A + size_t(b) * g.strideA;
    const cplx* __restrict__ B = g.B + size_t(b) * g.strideB;
    cplx* __restrict__ Cm = g.C + size_t(b) * g.strideC;
    const cplx zero = make_double2(0, 0);

    auto stage = [&](int kc, int Kc, int K4) {
        for (int idx = tid; idx < TM * K4; idx += NT) {
            const int mm = idx % TM, kk = idx / TM;
            const int gm = m0 + mm;
            cplx* dst = As + kk * LDA + mm;
            if (gm < g.M && kk < Kc) cp_async16(dst, A + size_t(kc + kk) * g.lda + gm);
            else *dst = zero;
        }
        if (g.b_kmajor) {
            for (int idx = tid; idx < TN * K4; idx += NT) {
                const int nn = idx % TN, kk = idx / TN;
                const int gn = n0 + nn;
                cplx* dst = Bs + kk * LDB + nn;
                if (gn < g.N && kk < Kc) cp_async16(dst, B + size_t(kc + kk) * g.ldb + gn);
                else *dst = zero;
            }
        } else {
            for (int idx = tid; idx < TN * K4; idx += NT) {
                const int kk = idx % K4, nn = idx / K4;
                const int gn = n0 + nn;
                cplx* dst = Bs + kk * LDB + nn;
                if (gn < g.N && kk < Kc) cp_async16(dst, B + size_t(gn) * g.ldb + kc + kk);
                else *dst = zero;
            }
        }
        cp_async_commit();
    };

    // the first (usually only) pass of the panels is in flight while the accumulators are loaded from C
    const int Kc0 = min(KT, K), K40 = (Kc0 + 3) & ~3;
    stage(0, Kc0, K40);

    double acc_re[MB][NB][2], acc_im[MB][NB][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        const int gm = m0 + wm * (8 * MB) + mb * 8 + grp;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = n0 + wn * (8 * NB) + nb * 8 + 2 * t4 + e;
                cplx c = zero;
                if (gm < g.M && gn < g.N) c = Cm[size_t(gn) * g.ldc + gm];
                acc_re[mb][nb][e] = c.x;
                acc_im[mb][nb][e] = c.y;
            }
    }
    const double alpha = g.alpha;
    for (int kc = 0; kc < K; kc += KT) {
        const int Kc = min(KT, K - kc);
        const int K4 = (Kc + 3) & ~3;
        if (kc > 0) {
            __syncthreads();
            stage(kc, Kc, K4);
        }
        cp_async_wait_all();
        __syncthreads();
        for (int k4 = 0; k4 < K4; k4 += 4) {
            double ar[MB], ai[MB], nai[MB], br[NB], bi[NB];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                const cplx a = As[(k4 + t4) * LDA + wm * (8 * MB) + mb * 8 + grp];
                ar[mb] = alpha * a.x;
                ai[mb] = alpha * a.y;
                nai[mb] = -ai[mb];
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const cplx bb = Bs[(k4 + t4) * LDB + wn * (8 * NB) + nb * 8 + grp];
                br[nb] = bb.x;
                bi[nb] = bb.y;
            }
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    dmma(acc_re[mb][nb][0], acc_re[mb][nb][1], ar[mb], br[nb]);
                    dmma(acc_re[mb][nb][0], acc_re[mb][nb][1], nai[mb], bi[nb]);
                    dmma(acc_im[mb][nb][0], acc_im[mb][nb][1], ar[mb], bi[nb]);
                    dmma(acc_im[mb][nb][0], acc_im[mb][nb][1], ai[mb], br[nb]);
                }
        }
    }
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        const int gm = m0 + wm * (8 * MB) + mb * 8 + grp;
        if (gm >= g.M) continue;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = n0 + wn * (8 * NB) + nb * 8 + 2 * t4 + e;
                if (gn >= g.N) continue;
                Cm[size_t(gn) * g.ldc + gm] = make_double2(acc_re[mb][nb][e], acc_im[mb][nb][e]);
            }
    }
}

template <int WM, int WN, int MB, int NB>
cudaError_t launch_rank_update(const GemmArgs& g, cudaStream_t st) {
    constexpr int TM = 8 * MB * WM, TN = 8 * NB * WN;
    static const bool legacy = std::getenv("DQMC_RANKUPD_LEGACY") != nullptr;
    // complex leading dimensions of the cp.async variant must be == 2 (mod 8); alignment: 16-byte elements throughout
    if (!legacy && (TM + 2) % 8 == 2 && (TN + 2) % 8 == 2) {
        const size_t smem2 = size_t(KT) * (TM + 2 + TN + 2) * sizeof(cplx);
        cudaError_t e2 = cudaFuncSetAttribute(zgemm_rank_update2_kernel<WM, WN, MB, NB>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e2 != cudaSuccess) return e2;
        dim3 grid2((g.M + TM - 1) / TM, (g.N + TN - 1) / TN, g.batch);
        launch_pdl(zgemm_rank_update2_kernel<WM, WN, MB, NB>, dim3(grid2), dim3(32 * WM * WN), smem2, st, g);
        return cudaGetLastError();
    }
    const size_t smem = size_t(2) * KT * (TM + 4 + TN + 4) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(zgemm_rank_update_kernel<WM, WN, MB, NB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((g.M + TM - 1) / TM, (g.N + TN - 1) / TN, g.batch);
    launch_pdl(zgemm_rank_update_kernel<WM, WN, MB, NB>, dim3(grid), dim3(32 * WM * WN), smem, st, g);
    return cudaGetLastError();
}

template <int WM, int WN, int MB, int NB, int BKT = 8>
cudaError_t launch_cfg(const GemmArgs& g, cudaStream_t st) {
    typedef GemmCfg<WM, WN, MB, NB, BKT> C;
    static const bool legacy = std::getenv("DQMC_GEMM_LEGACY") != nullptr;
    static const bool always = std::getenv("DQMC_GEMM_RING_ALWAYS") != nullptr;
    // the cp.async ring pays for tiles with enough arithmetic per k-step; the skinny panel products of the blocked
    // QR (32-wide tiles) measured slower with it
    if (!legacy && BKT == 8 && (always || (C::TM >= 48 && C::TN >= 48))) {
        const size_t smem2 = size_t(kStages) * BK * (C::TM + 2 + C::TN + 2) * sizeof(cplx);
        cudaError_t e2 = cudaFuncSetAttribute(zgemm_dmma2_kernel<WM, WN, MB, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem2);
        if (e2 != cudaSuccess) return e2;
        dim3 grid2((g.M + C::TM - 1) / C::TM, (g.N + C::TN - 1) / C::TN, g.batch);
        launch_pdl(zgemm_dmma2_kernel<WM, WN, MB, NB>, dim3(grid2), dim3(C::NT), smem2, st, g);
        return cudaGetLastError();
    }
    const size_t smem = C::SMEM_DOUBLES * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(zgemm_dmma_kernel<WM, WN, MB, NB, BKT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((g.M + C::TM - 1) / C::TM, (g.N + C::TN - 1) / C::TN, g.batch);
    launch_pdl(zgemm_dmma_kernel<WM, WN, MB, NB, BKT>, dim3(grid), dim3(C::NT), smem, st, g);
    return cudaGetLastError();
}

}  // namespace

// Matrices in flight on the device (set by dqmc_create): with few of them a launch cannot fill the SMs with
// 96-wide tiles (9 CTAs per 288 x 288 matrix), so smaller tiles are chosen -- more CTAs, shorter latency per
// launch, same arithmetic in the same order (results do not depend on the tile shape).
static int g_matrices_in_flight = 1 << 30;
void gemm_set_matrices_in_flight(int n) { g_matrices_in_flight = n > 0 ? n : 1 << 30; }
int gemm_matrices_in_flight() { return g_matrices_in_flight; }

// programmatic dependent launch (dqmc_internal.h): on unless DQMC_PDL=0 says otherwise
static bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
void pdl_set_enabled(bool on) {
    static const char* env = std::getenv("DQMC_PDL");
    g_pdl = env ? std::atoi(env) != 0 : on;
}

cudaError_t gemm_launch(const GemmArgs& g, cudaStream_t st) {
    if (g.batch <= 0 || g.M <= 0 || g.N <= 0) return cudaSuccess;
    if ((g.K <= KT || g.kvec) && !g.transa && !g.transb && g.beta == 1.0 && !g.rowscale && !g.colscale && !g.kscale) {
        static const int cfgEnv = std::getenv("DQMC_RANKUPD_CFG") ? std::atoi(std::getenv("DQMC_RANKUPD_CFG")) : 0;
        // 32 x 32 tiles for up to 32 matrices in flight (81 CTAs per matrix: the launch is latency bound); otherwise 64 x 48 tiles at
        // 80 registers (three CTAs per SM hide the C-tile round trip better than two 96 x 48 ones: 120.6 -> 119.4 ms per step)
        const int cfg = cfgEnv ? cfgEnv : (g_matrices_in_flight <= 32 ? 5 : 7);   // measured: 32 replicas 78.1 -> 75.4 ms with the small tiles, 64: 121.8 -> 130.5
        if (g.M % 48 == 0 && g.N % 48 == 0 && cfg == 1) return launch_rank_update<2, 2, 3, 3>(g, st);   // 48 x 48
        if (g.M % 96 == 0 && g.N % 48 == 0 && cfg == 2) return launch_rank_update<4, 2, 3, 3>(g, st);   // 96 x 48
        if (g.M % 48 == 0 && g.N % 96 == 0 && cfg == 3) return launch_rank_update<2, 4, 3, 3>(g, st);   // 48 x 96
        if (g.M % 96 == 0 && g.N % 96 == 0 && cfg == 4) return launch_rank_update<4, 4, 3, 3>(g, st);   // 96 x 96, 16 warps
        if (g.M % 96 == 0 && g.N % 32 == 0 && cfg == 6) return launch_rank_update<4, 2, 3, 2>(g, st);   // 96 x 32, 8 warps, 3 CTAs per SM
        if (g.N % 48 == 0 && cfg == 7) return launch_rank_update<4, 2, 2, 3>(g, st);                    // 64 x 48, 8 warps, 3 CTAs per SM
        if (g.M % 32 == 0 && g.N % 32 == 0 && cfg == 5) return launch_rank_update<2, 2, 2, 2>(g, st);   // 32 x 32, 4 warps
        if (g.M % 96 == 0 && g.N % 96 == 0) return launch_rank_update<4, 3, 3, 4>(g, st);
        static const int cfg2 = std::getenv("DQMC_RANKUPD_CFG2") ? std::atoi(std::getenv("DQMC_RANKUPD_CFG2")) : 3;
        if (cfg2 == 1 && g.M > 32 && g.N > 32) return launch_rank_update<4, 4, 2, 2>(g, st);            // 64 x 64, 16 warps
        if (cfg2 == 2 && g.M > 32 && g.N > 32) return launch_rank_update<4, 2, 2, 2>(g, st);            // 64 x 32, 8 warps
        if (cfg2 == 3 && g.M > 32 && g.N > 32) return launch_rank_update<2, 2, 2, 2>(g, st);            // 32 x 32, 4 warps
        if ((g.M > 32 && g.N > 32) || g.kvec) return launch_rank_update<2, 2, 4, 4>(g, st);
    }
    if (g.kvec || g.b_kmajor) return cudaErrorInvalidValue;           // per-matrix K exists on the rank-update path only
    // 96 x 96 tiles when they fit the problem exactly (D = 288), otherwise 64 x 64; skinny shapes
    // (the panel products of the blocked QR / triangular solve) get 32 x 64 and 64 x 32 tiles
    static const int smallEnv = std::getenv("DQMC_GEMM_SMALL") ? std::atoi(std::getenv("DQMC_GEMM_SMALL")) : -1;
    // tile shape of the D x D products by the matrices of THIS launch (with one lane per replica a launch carries one
    // matrix whatever the batch: 96 x 96 tiles would be 9 CTAs of 12 warps, 48 x 48 tiles are 36 CTAs of 4)
    const int small = smallEnv >= 0 ? smallEnv : (g.batch <= 8 ? 1 : 0);
    // small batches: 32-deep k-tiles for the shapes that run as a handful of CTAs (DQMC_GEMM_DEEPK=0 / 1 overrides)
    static const int deepEnv = std::getenv("DQMC_GEMM_DEEPK") ? std::atoi(std::getenv("DQMC_GEMM_DEEPK")) : -1;
    const bool deep = deepEnv >= 0 ? deepEnv != 0 : g_matrices_in_flight <= 32;
    if (small == 3 && g.M % 32 == 0 && g.N % 32 == 0 && g.K >= 64) return launch_cfg<2, 2, 2, 2, 32>(g, st);   // 32 x 32, 4 warps
    if (deep && g.M <= 32 && g.N <= 32) return launch_cfg<2, 2, 2, 2, 32>(g, st);
    // DQMC_SKINNY_CFG: shape of the M <= 32 panel products -- 0: 32 x 32 tiles (8 warps), 1: 32 x 16 (8 warps), 2: 32 x 8 (4 warps)
    static const int skinnyEnv = std::getenv("DQMC_SKINNY_CFG") ? std::atoi(std::getenv("DQMC_SKINNY_CFG")) : -1;
    const int skinny = skinnyEnv >= 0 ? skinnyEnv : (g_matrices_in_flight <= 32 ? 1 : 0);   // measured: 1 replica 49.7 -> 48.3 ms, 32: 70.2 -> 69.4, 64: 117.9 -> 121.0
    if (g.M <= 32 && g.N > 32 && skinny == 1) { if (deep) return launch_cfg<4, 2, 1, 1, 32>(g, st); return launch_cfg<4, 2, 1, 1>(g, st); }
    if (g.M <= 32 && g.N > 32 && skinny == 2) { if (deep) return launch_cfg<4, 1, 1, 1, 32>(g, st); return launch_cfg<4, 1, 1, 1>(g, st); }
    if (deep && g.M <= 32) return launch_cfg<2, 4, 2, 1, 32>(g, st);
    if (deep && g.N <= 32) return launch_cfg<4, 1, 2, 4, 32>(g, st);
    if (small == 1 && g.M % 48 == 0 && g.N % 48 == 0 && g.K >= 64) return launch_cfg<2, 2, 3, 3>(g, st);   // 48 x 48, 4 warps
    if (small == 2 && g.M % 96 == 0 && g.N % 48 == 0 && g.K >= 64) return launch_cfg<4, 2, 3, 3>(g, st);   // 96 x 48, 8 warps
    if (g.M % 96 == 0 && g.N % 96 == 0 && g.K >= 64) return launch_cfg<4, 3, 3, 4>(g, st);   // 12 warps, 24 x 32 each
    if (g.M <= 32 && g.N <= 32) return launch_cfg<1, 1, 4, 4>(g, st);
    static const int cfgw = std::getenv("DQMC_WGEMM_CFG") ? std::atoi(std::getenv("DQMC_WGEMM_CFG")) : 3;
    if (g.M <= 32 && cfgw == 1) return launch_cfg<2, 4, 2, 2>(g, st);          // 32 x 64, 8 warps
    if (g.M <= 32 && cfgw == 2) return launch_cfg<2, 2, 2, 2>(g, st);          // 32 x 32, 4 warps
    if (g.M <= 32 && cfgw == 3) return launch_cfg<2, 4, 2, 1>(g, st);          // 32 x 32, 8 warps
    if (g.M <= 32) return launch_cfg<1, 4, 4, 2>(g, st);                       // 32 x 64, 4 warps
    if (g.N <= 32) return launch_cfg<4, 1, 2, 4>(g, st);                       // 64 x 32, 4 warps
    return launch_cfg<2, 2, 4, 4>(g, st);
}

}  // namespace dqmc
