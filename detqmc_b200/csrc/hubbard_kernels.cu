// DetHubbard on the GPU: dense B matrices through the DMMA GEMM, single-flip local updates.
//
// Replaces DetHubbard::computeBmat and the dense functors (dethubbard.cpp:823-851, dethubbard.h:281-337),
// setupPropTmat (dethubbard.cpp:753-821; detmodel.cpp:21-39), setupRandomAuxfield (:741-751) and
// updateInSlice / weightRatioSingleFlip / updateGreenFunctionWithFlip (:141-171, 892-933).
//
// B_sigma(k) = diag(exp(sigma * alpha * s_k)) * e^{-dtau T}.  The reference multiplies with the dense
// chain product and LU-inverts it on every call; here a chain is applied slice by slice as
// "propagator GEMM with the diagonal fused as a row / column / k scaling", and the inverse propagator
// e^{+dtau T} is precomputed once.  The two Green's-function components (up, down) are two matrices of
// the batch (matrix index = 2 * replica + gc), stored as complex with zero imaginary part so that the
// whole stabilisation code (GEMM, blocked QR, Green's function) is shared with DetSDW.
#include "dqmc_internal.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include <cmath>

namespace dqmc {

// ------------------------------------------------------------------------------------------------
// host: e^{-dtau T} and e^{+dtau T}
// ------------------------------------------------------------------------------------------------
namespace {

// cyclic Jacobi eigenvalue iteration for a real symmetric n x n matrix (row-major a, destroyed);
// eigenvectors in the columns of v.  n <= a few hundred, called once per context.
void jacobi_eigh(std::vector<double>& a, std::vector<double>& v, int n) {
    v.assign(size_t(n) * n, 0.0);
    for (int i = 0; i < n; ++i) v[size_t(i) * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) off += a[size_t(p) * n + q] * a[size_t(p) * n + q];
        if (off < 1e-30) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[size_t(p) * n + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double app = a[size_t(p) * n + p], aqq = a[size_t(q) * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = a[size_t(k) * n + p], akq = a[size_t(k) * n + q];
                    a[size_t(k) * n + p] = c * akp - s * akq;
                    a[size_t(k) * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = a[size_t(p) * n + k], aqk = a[size_t(q) * n + k];
                    a[size_t(p) * n + k] = c * apk - s * aqk;
                    a[size_t(q) * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = v[size_t(k) * n + p], vkq = v[size_t(k) * n + q];
                    v[size_t(k) * n + p] = c * vkp - s * vkq;
                    v[size_t(k) * n + q] = s * vkp + c * vkq;
                }
            }
    }
}

void matmul(const std::vector<double>& A, const std::vector<double>& B, std::vector<double>& C, int n) {
    C.assign(size_t(n) * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) {
            const double aik = A[size_t(i) * n + k];
            if (aik == 0.0) continue;
            for (int j = 0; j < n; ++j) C[size_t(i) * n + j] += aik * B[size_t(k) * n + j];
        }
}

}  // namespace

// P = e^{-dtau T}, Pinv = e^{+dtau T}  (symmetric; row-major == column-major)
void hub_build_propagators(const dqmc_params& p, std::vector<double>& P, std::vector<double>& Pinv) {
    const int L = p.L, N = L * L;
    auto nb = [&](int dir, int s) {
        const int x = s % L, y = s / L;
        switch (dir) {
            case 0: return y * L + (x + 1) % L;
            case 1: return y * L + (x + L - 1) % L;
            case 2: return ((y + 1) % L) * L + x;
            default: return ((y + L - 1) % L) * L + x;
        }
    };
    if (!p.checkerboard) {
        // T = -t * adjacency - mu * 1;  e^{-+dtau T} = V e^{-+dtau w} V^T   (detmodel.cpp:21-39)
        std::vector<double> tm(size_t(N) * N, 0.0), v;
        for (int s = 0; s < N; ++s) {
            tm[size_t(s) * N + s] -= p.mu;
            for (int d = 0; d < 4; ++d) tm[size_t(nb(d, s)) * N + s] -= p.t;
        }
        jacobi_eigh(tm, v, N);
        P.assign(size_t(N) * N, 0.0);
        Pinv.assign(size_t(N) * N, 0.0);
        for (int k = 0; k < N; ++k) {
            const double w = tm[size_t(k) * N + k];
            const double em = std::exp(-p.dtau * w), ep = std::exp(+p.dtau * w);
            for (int i = 0; i < N; ++i) {
                const double vik = v[size_t(i) * N + k];
                for (int j = 0; j < N; ++j) {
                    const double x = vik * v[size_t(j) * N + k];
                    P[size_t(i) * N + j] += em * x;
                    Pinv[size_t(i) * N + j] += ep * x;
                }
            }
        }
    } else {
        // product of the four bond-group exponentials, each (cosh + sinh * k) with k a perfect matching
        // (dethubbard.cpp:768-821; no chemical potential in this form); the inverse reverses the order
        std::vector<double> k4[4];
        for (auto& k : k4) k.assign(size_t(N) * N, 0.0);
        for (int y = 0; y < L; ++y)
            for (int x = 0; x < L; x += 2) {
                const int a = y * L + x, na = nb(0, a), nbb = nb(0, na);
                k4[0][size_t(a) * N + na] = k4[0][size_t(na) * N + a] = 1.0;
                k4[1][size_t(na) * N + nbb] = k4[1][size_t(nbb) * N + na] = 1.0;
            }
        for (int x = 0; x < L; ++x)
            for (int y = 0; y < L; y += 2) {
                const int a = y * L + x, na = nb(2, a), nbb = nb(2, na);
                k4[2][size_t(a) * N + na] = k4[2][size_t(na) * N + a] = 1.0;
                k4[3][size_t(na) * N + nbb] = k4[3][size_t(nbb) * N + na] = 1.0;
            }
        const double ch = std::cosh(p.dtau * p.t), sh = std::sinh(p.dtau * p.t);
        auto factor = [&](int g, double sgn) {
            std::vector<double> e(size_t(N) * N, 0.0);
            for (int i = 0; i < N; ++i) e[size_t(i) * N + i] = ch;
            for (size_t i = 0; i < e.size(); ++i) e[i] += sgn * sh * k4[g][i];
            return e;
        };
        std::vector<double> acc = factor(0, +1.0), tmp;
        for (int g = 1; g < 4; ++g) { matmul(acc, factor(g, +1.0), tmp, N); acc.swap(tmp); }
        P = acc;
        acc = factor(3, -1.0);
        for (int g = 2; g >= 0; --g) { matmul(acc, factor(g, -1.0), tmp, N); acc.swap(tmp); }
        Pinv = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------------
namespace {

// out[mat][i] = exp(sign * sigma(gc) * alpha * aux[rep][k][i]),  mat = off + blockIdx.y, rep = mat / 2
__global__ void hub_scales_kernel(const int32_t* aux, long long strideAux, int N, int k, double alpha, double sign,
                                  double* out, int off) {
    pdl_enter();
    const int mat = off + blockIdx.y;
    const int rep = mat >> 1, gc = mat & 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double sigma = gc == 0 ? 1.0 : -1.0;
    const double s = double(aux[size_t(rep) * strideAux + size_t(k) * N + i]);
    out[size_t(blockIdx.y) * N + i] = exp(sign * sigma * alpha * s);
}

__global__ void hub_real_part_kernel(const cplx* in, double* out, size_t n) {
    pdl_enter();
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i].x;
}
__global__ void hub_to_complex_kernel(const double* in, cplx* out, size_t n) {
    pdl_enter();
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2(in[i], 0.0);
}

// updateInSlice (dethubbard.cpp:141-171): N attempts at random sites, in the reference's draw order
// (site = randInt(0, N-1); a further rand01() only if ratio <= 1).  One CTA per replica.
//
// The rank-1 updates of both components (dethubbard.cpp:910-933) are DELAYED: with the pending updates
// G_eff = G - sum_l u_l v_l^T held in shared memory (KD columns u_l, rows v_l per component), an attempt needs the
// diagonal element G_eff[site, site] only (KD products), an acceptance the column and the row of G_eff at the site
// (2 N KD products instead of N^2), and G itself is touched once per KD acceptances by the rank-KD flush
// G -= U V^T (4 x 4 register tiles per thread).  The arithmetic is the reference's, re-associated: every accepted
// flip contributes the same outer product  G_eff[:, site] * f (e_site - G_eff[site, :]).
//
// A thread-block cluster of CS CTAs serves one replica: streaming the two N x N matrices through ONE SM was the
// bottleneck of the flush, so every CTA of the cluster runs the identical decision chain (same random numbers, same
// arithmetic, private copy of the slice's auxiliary spins: no communication) and flushes its own 1 / CS of the
// columns of G; a cluster barrier on both sides of a flush orders it against the other CTAs' reads of G.
__global__ void __launch_bounds__(512) hub_update_slice_kernel(cplx* Gall, long long strideG, int N, int32_t* auxAll,
                                                                long long strideAux, int k, double alpha,
                                                                const double* rngAll, long long strideRng, int rngWindow,
                                                                int* cursorAll, uint32_t* acceptedAll,
                                                                unsigned long long* acceptedTotal, int* errflag, int KD,
                                                                int CS) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Uu = reinterpret_cast<double*>(smem_raw);       // [KD][N] pending columns, spin up
    double* Vu = Uu + size_t(KD) * N;                       // [KD][N] pending rows (factor included)
    double* Ud = Vu + size_t(KD) * N;
    double* Vd = Ud + size_t(KD) * N;
    int32_t* sAux = reinterpret_cast<int32_t*>(Vd + size_t(KD) * N);   // [N] this slice's auxiliary spins, private copy
    __shared__ int sSite, sAcc, sAbort;
    __shared__ double sFacU, sFacD;
    const int b = blockIdx.x / CS, crank = blockIdx.x % CS, tid = threadIdx.x, lane = tid & 31;
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    cplx* Gu = Gall + size_t(2 * b) * strideG;
    cplx* Gd = Gall + size_t(2 * b + 1) * strideG;
    int32_t* aux = auxAll + size_t(b) * strideAux + size_t(k) * N;
    const double* rng = rngAll + size_t(b) * strideRng;
    int cursor = cursorAll[b];
    unsigned accepted = 0;
    int np = 0;                                             // pending updates (uniform across the CTA)
    if (tid == 0) sAbort = 0;
    for (int i = tid; i < N; i += blockDim.x) sAux[i] = aux[i];
    __syncthreads();

    auto flush = [&]() {
        // G -= U V^T for both components: thread <-> 4 x 4 tile (rows i0.., columns j0..); this CTA's share of the columns
        const int nti = (N + 3) / 4;
        const int tj0 = (nti * crank) / CS, tj1 = (nti * (crank + 1)) / CS;
        if (CS > 1) cluster.sync();                          // every CTA has finished reading G for its pending updates
        for (int tile = tid; tile < nti * (tj1 - tj0); tile += blockDim.x) {
            const int ti = tile % nti, tj = tj0 + tile / nti;
            const int i0 = 4 * ti, j0 = 4 * tj;
            double au[4][4], ad[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int r = 0; r < 4; ++r) { au[c][r] = 0.0; ad[c][r] = 0.0; }
            for (int l = 0; l < np; ++l) {
                double uu[4], vu[4], ud[4], vd[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int i = min(i0 + r, N - 1), j = min(j0 + r, N - 1);
                    uu[r] = Uu[l * N + i]; ud[r] = Ud[l * N + i];
                    vu[r] = Vu[l * N + j]; vd[r] = Vd[l * N + j];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        au[c][r] = fma(uu[r], vu[c], au[c][r]);
                        ad[c][r] = fma(ud[r], vd[c], ad[c][r]);
                    }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (i0 + r < N && j0 + c < N) {
                        const size_t idx = size_t(j0 + c) * N + i0 + r;
                        Gu[idx].x -= au[c][r];
                        Gd[idx].x -= ad[c][r];
                    }
        }
        if (CS > 1) { __threadfence(); cluster.sync(); }     // the other CTAs' shares are visible before G is read again
    };

    for (int attempt = 0; attempt < N; ++attempt) {
        if (tid < 32) {
            bool abort = false, acc = false;
            int site = 0;
            if (cursor + 2 > rngWindow) {
                abort = true;
            } else {
                site = int(double(N) * rng[cursor]);                    // randInt(0, N-1), rngwrapper.h:60-68
                cursor += 1;
                // diagonal elements of G_eff: the pending terms are summed in a fixed order (lane l, then a tree)
                double pu = 0.0, pd = 0.0;
                for (int l = lane; l < np; l += 32) {
                    pu = fma(Uu[l * N + site], Vu[l * N + site], pu);
                    pd = fma(Ud[l * N + site], Vd[l * N + site], pd);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    pu += __shfl_xor_sync(0xffffffffu, pu, o);
                    pd += __shfl_xor_sync(0xffffffffu, pd, o);
                }
                const double a = double(sAux[site]);
                const double dU = exp(-2.0 * alpha * a) - 1.0, dD = exp(+2.0 * alpha * a) - 1.0;
                const double gu = Gu[size_t(site) * N + site].x - pu, gd = Gd[size_t(site) * N + site].x - pd;
                const double ratio = (1.0 + dU * (1.0 - gu)) * (1.0 + dD * (1.0 - gd));
                if (ratio > 1.0) {
                    acc = true;
                } else {
                    acc = rng[cursor] < ratio;
                    cursor += 1;
                }
                __syncwarp();
                if (acc && lane == 0) {
                    sAux[site] = -sAux[site];
                    sFacU = dU / (1.0 + dU * (1.0 - gu));
                    sFacD = dD / (1.0 + dD * (1.0 - gd));
                }
                if (acc) accepted += 1;
            }
            if (lane == 0) {
                sSite = site;
                sAcc = acc ? 1 : 0;
                if (abort) { sAbort = 1; sAcc = 0; }
            }
        }
        __syncthreads();
        if (sAbort) break;
        if (sAcc) {
            const int site = sSite;
            const double fU = sFacU, fD = sFacD;
            // column and row of G_eff at the site become the next pending update
            for (int i = tid; i < N; i += blockDim.x) {
                double cu = Gu[size_t(site) * N + i].x, cd = Gd[size_t(site) * N + i].x;       // G[i, site]
                double ru = Gu[size_t(i) * N + site].x, rd = Gd[size_t(i) * N + site].x;       // G[site, i]
                for (int l = 0; l < np; ++l) {
                    cu = fma(-Uu[l * N + i], Vu[l * N + site], cu);
                    cd = fma(-Ud[l * N + i], Vd[l * N + site], cd);
                    ru = fma(-Uu[l * N + site], Vu[l * N + i], ru);
                    rd = fma(-Ud[l * N + site], Vd[l * N + i], rd);
                }
                const double e = i == site ? 1.0 : 0.0;
                Uu[np * N + i] = cu;
                Ud[np * N + i] = cd;
                Vu[np * N + i] = fU * (e - ru);
                Vd[np * N + i] = fD * (e - rd);
            }
            np += 1;
            __syncthreads();
            if (np == KD) {
                flush();
                np = 0;
                __threadfence_block();
            }
        }
        __syncthreads();
    }
    if (np > 0) flush();
    if (crank != 0) return;
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) aux[i] = sAux[i];
    if (tid == 0) {
        if (sAbort) atomicExch(errflag, 1);
        cursorAll[b] = cursor;
        acceptedAll[b] = accepted;
        if (acceptedTotal) acceptedTotal[b] += accepted;
    }
}

// DetHubbard::measure (dethubbard.cpp:511-539) for one time slice: sums of the diagonal and nearest-neighbour
// elements of both Green's-function components and the spin-z correlations with site 0, accumulated over the
// slices of a sweep.  acc (per replica): sum_GiiUp, sum_GiiDn, sum_GiiUpDn, sum_GneighUp, sum_GneighDn, zcorr[N].
// One CTA per replica; the block sums are reduced in a fixed order (deterministic).
__global__ void __launch_bounds__(256) hub_measure_kernel(const cplx* Gall, long long strideG, int N, int L, double* accAll,
                                                          long long strideAcc) {
    pdl_enter();
    __shared__ double part[5][256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const cplx* Gu = Gall + size_t(2 * b) * strideG;
    const cplx* Gd = Gall + size_t(2 * b + 1) * strideG;
    double* acc = accAll + size_t(b) * strideAcc;
    double s[5] = {0, 0, 0, 0, 0};
    const double gu00 = Gu[0].x, gd00 = Gd[0].x;
    for (int site = tid; site < N; site += blockDim.x) {
        const double gu = Gu[size_t(site) * N + site].x, gd = Gd[size_t(site) * N + site].x;
        s[0] += gu;
        s[1] += gd;
        s[2] += gu * gd;
        // PeriodicCubicLatticeNearestNeighbors (neighbortable.h), d = 2: +x, -x, +y, -y; element (site, neighbour)
        const int x = site % L, y = site / L;
        const int nb[4] = {y * L + (x + 1) % L, y * L + (x + L - 1) % L, ((y + 1) % L) * L + x, ((y + L - 1) % L) * L + x};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            s[3] += Gu[size_t(nb[q]) * N + site].x;
            s[4] += Gd[size_t(nb[q]) * N + site].x;
        }
        if (site == 0) {
            acc[5] += -2.0 * gu00 * gd00 + gu00 + gd00;
        } else {
            const double gu0j = Gu[size_t(site) * N].x, gd0j = Gd[size_t(site) * N].x;
            acc[5 + site] += gu00 * gu - gu00 * gd + gd00 * gd - gd00 * gu - gu0j * gu0j - gd0j * gd0j;
        }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) part[q][tid] = s[q];
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) {
#pragma unroll
            for (int q = 0; q < 5; ++q) part[q][tid] += part[q][tid + off];
        }
        __syncthreads();
    }
    if (tid < 5) acc[tid] += part[tid][0];
}

}  // namespace

cudaError_t hub_scales_launch(const int32_t* aux, long long strideAux, int N, int k, double alpha, double sign,
                              double* out, int off, int batch, cudaStream_t st) {
    dim3 grid((N + 127) / 128, batch);
    launch_pdl(hub_scales_kernel, dim3(grid), dim3(128), 0, st, aux, strideAux, N, k, alpha, sign, out, off);
    return cudaGetLastError();
}

cudaError_t hub_real_part_launch(const cplx* in, double* out, size_t n, cudaStream_t st) {
    launch_pdl(hub_real_part_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, in, out, n);
    return cudaGetLastError();
}
cudaError_t hub_to_complex_launch(const double* in, cplx* out, size_t n, cudaStream_t st) {
    launch_pdl(hub_to_complex_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, in, out, n);
    return cudaGetLastError();
}

cudaError_t hub_update_slice_launch(cplx* G, long long strideG, int N, int32_t* aux, long long strideAux, int k,
                                    double alpha, const double* rng, long long strideRng, int rngWindow, int* cursor,
                                    uint32_t* accepted, unsigned long long* acceptedTotal, int* errflag, int batch,
                                    cudaStream_t st) {
    // delay depth: as many pending updates as 200 KB of shared memory hold (two components, columns and rows), at most 32
    int KD = int((size_t(200) * 1024) / (size_t(4) * N * sizeof(double)));
    KD = std::max(1, std::min(KD, 32));
    static const int kdEnv = std::getenv("DQMC_HUB_DELAY") ? std::atoi(std::getenv("DQMC_HUB_DELAY")) : 0;
    if (kdEnv > 0) KD = std::min(KD, kdEnv);
    const size_t smem = size_t(4) * KD * N * sizeof(double) + size_t(N) * sizeof(int32_t);
    const int threads = N >= 256 ? 512 : 256;
    // cluster size: 8 CTAs per replica for the large lattices (the flush is bandwidth bound), one for the small ones
    static const int csEnv = std::getenv("DQMC_HUB_CLUSTER") ? std::atoi(std::getenv("DQMC_HUB_CLUSTER")) : 0;
    int CS = csEnv > 0 ? csEnv : (N >= 256 ? 8 : 1);
    CS = std::max(1, std::min(CS, 8));
    cudaError_t e = cudaFuncSetAttribute(hub_update_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // halve the cluster until one of them fits the device (co-scheduling CS CTAs of this size on one GPC); the answer
    // for a (size, cluster) pair is remembered
    static size_t cachedSmem = 0;
    static int cachedWant = 0, cachedCS = 0;
    if (cachedSmem == smem && cachedWant == CS) CS = cachedCS;
    const int want = CS;
    while (CS > 1 && !(cachedSmem == smem && cachedWant == want)) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CS);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, hub_update_slice_kernel, &cfg) == cudaSuccess && nclusters >= 1) break;
        (void)cudaGetLastError();
        CS /= 2;
    }
    cachedSmem = smem; cachedWant = want; cachedCS = CS;
    e = launch_pdl_cluster(hub_update_slice_kernel, dim3(batch * CS), dim3(threads), smem, st, (unsigned)CS, G, strideG, N, aux,
                           strideAux, k, alpha, rng, strideRng, rngWindow, cursor, accepted, acceptedTotal, errflag, KD, CS);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t hub_measure_launch(const cplx* G, long long strideG, int N, int L, double* acc, long long strideAcc, int batch,
                               cudaStream_t st) {
    launch_pdl(hub_measure_kernel, dim3(batch), dim3(256), 0, st, G, strideG, N, L, acc, strideAcc);
    return cudaGetLastError();
}

}  // namespace dqmc
