// Small batched helper kernels: identity, conjugate transpose, max|A-B|, cosh/sinh tables,
// global field shift, bosonic action and exchange action reductions.
#include "dqmc_internal.h"

#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace dqmc {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// deterministic block reduction (fixed tree, no atomics): result valid in thread 0
template <bool IS_MAX>
__device__ __forceinline__ double block_reduce(double v, double* red) {
    v = IS_MAX ? warp_max(v) : warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : (IS_MAX ? 0.0 : 0.0);
        v = IS_MAX ? warp_max(v) : warp_sum(v);
    }
    return v;
}

__global__ void set_identity_kernel(cplx* A, int D, long long stride) {
    pdl_enter();
    cplx* a = A + size_t(blockIdx.y) * stride;
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(D) * D) return;
    const int i = idx % D, j = idx / D;
    a[idx] = make_double2(i == j ? 1.0 : 0.0, 0.0);
}

__global__ void conj_transpose_kernel(const cplx* A, cplx* B, int D, long long stride) {
    pdl_enter();
    __shared__ cplx tile[32][33];
    const cplx* a = A + size_t(blockIdx.z) * stride;
    cplx* bm = B + size_t(blockIdx.z) * stride;
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int i = i0 + threadIdx.x, j = j0 + jj;
        if (i < D && j < D) tile[jj][threadIdx.x] = a[size_t(j) * D + i];
    }
    __syncthreads();
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
        const int j = j0 + threadIdx.x, i = i0 + ii;
        if (i < D && j < D) {
            cplx v = tile[threadIdx.x][ii];
            v.y = -v.y;
            bm[size_t(i) * D + j] = v;     // B[j, i] = conj(A[i, j])
        }
    }
}

// B[j, i] = rowscale[j] * conj(A[i, j])   (separate batch strides; strideA = 0 shares one input)
__global__ void scaled_conj_transpose_kernel(const cplx* A, long long strideA, cplx* B, long long strideB,
                                             const double* rowscale, long long strideS, int D) {
    pdl_enter();
    __shared__ cplx tile[32][33];
    const cplx* a = A + size_t(blockIdx.z) * strideA;
    cplx* bm = B + size_t(blockIdx.z) * strideB;
    const double* rs = rowscale ? rowscale + size_t(blockIdx.z) * strideS : nullptr;
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int i = i0 + threadIdx.x, j = j0 + jj;
        if (i < D && j < D) tile[jj][threadIdx.x] = a[size_t(j) * D + i];
    }
    __syncthreads();
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
        const int j = j0 + threadIdx.x, i = i0 + ii;
        if (i < D && j < D) {
            cplx v = tile[threadIdx.x][ii];
            const double sc = rs ? rs[j] : 1.0;
            bm[size_t(i) * D + j] = make_double2(v.x * sc, -v.y * sc);
        }
    }
}

// max |A - B| per matrix; a cluster of CS CTAs per matrix, CTA 0 takes the maximum over its peers' partial results
// (distributed shared memory)
__global__ void max_abs_diff_kernel(const cplx* A, const cplx* B, int D, long long stride, double* out, int CS) {
    pdl_enter();
    __shared__ double red[32];
    __shared__ double part;
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const int m = blockIdx.x / CS, crank = blockIdx.x % CS;
    const cplx* a = A + size_t(m) * stride;
    const cplx* b = B + size_t(m) * stride;
    double mx = 0;
    for (size_t idx = size_t(crank) * blockDim.x + threadIdx.x; idx < size_t(D) * D; idx += size_t(CS) * blockDim.x) {
        const double dx = a[idx].x - b[idx].x, dy = a[idx].y - b[idx].y;
        mx = fmax(mx, sqrt(dx * dx + dy * dy));
    }
    mx = block_reduce<true>(mx, red);
    if (threadIdx.x == 0) part = mx;
    __syncthreads();
    if (CS > 1) cluster.sync();
    if (crank == 0 && threadIdx.x == 0) {
        double t = part;
        for (int r = 1; r < CS; ++r) t = fmax(t, *cluster.map_shared_rank(&part, r));
        out[m] = t;
    }
    if (CS > 1) cluster.sync();
}

// coshTermPhi / sinhTermPhi for every (slice >= 1, site)   (detsdwopdim.cpp:1131-1136, 1174-1181)
__global__ void update_tables_kernel(const double* phi, double* coshT, double* sinhT, int N, int opdim, int m,
                                     double lambda_dtau, long long stridePhi, long long strideTab) {
    pdl_enter();
    const int b = blockIdx.y;
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(m) * N) return;
    const int k = 1 + idx / N, s = idx % N;
    const double* p = phi + size_t(b) * stridePhi + size_t(k) * opdim * N;
    double n2 = 0;
    for (int d = 0; d < opdim; ++d) n2 += p[d * N + s] * p[d * N + s];
    const double nrm = sqrt(n2);
    coshT[size_t(b) * strideTab + size_t(k) * N + s] = cosh(lambda_dtau * nrm);
    sinhT[size_t(b) * strideTab + size_t(k) * N + s] = sinh(lambda_dtau * nrm) / nrm;
}

// addGlobalRandomDisplacement (detsdwopdim.cpp:3755-3763): every slice INCLUDING the unused k = 0
__global__ void shift_fields_kernel(double* phi, const double* shift, int N, int opdim, int m, long long stridePhi) {
    pdl_enter();
    const int b = blockIdx.y;
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(m + 1) * opdim * N) return;
    const int d = (idx / N) % opdim;
    phi[size_t(b) * stridePhi + idx] += shift[b * 3 + d];
}

// phiAction (detsdwopdim.cpp:4242-4299), one CTA per replica, fixed reduction order
__global__ void phi_action_kernel(const double* phi, const double* rvals, double* out, int L, int opdim, int m,
                                  double dtau, double c, double u, long long stridePhi) {
    pdl_enter();
    __shared__ double red[32];
    const int b = blockIdx.x;
    const int N = L * L;
    const double* ph = phi + size_t(b) * stridePhi;
    const double r = rvals[b];
    double acc = 0;
    for (int idx = threadIdx.x; idx < m * N; idx += blockDim.x) {
        const int k = 1 + idx / N, s = idx % N;
        const int ke = k > 1 ? k - 1 : m;
        const int x = s % L, y = s / L;
        const int sx = y * L + (x + 1 == L ? 0 : x + 1);
        const int sy = (y + 1 == L ? 0 : y + 1) * L + x;
        double td2 = 0, xd2 = 0, yd2 = 0, sq = 0;
        for (int d = 0; d < opdim; ++d) {
            const double v = ph[(size_t(k) * opdim + d) * N + s];
            const double td = (v - ph[(size_t(ke) * opdim + d) * N + s]) / dtau;
            const double xd = v - ph[(size_t(k) * opdim + d) * N + sx];
            const double yd = v - ph[(size_t(k) * opdim + d) * N + sy];
            td2 += td * td; xd2 += xd * xd; yd2 += yd * yd; sq += v * v;
        }
        acc += (dtau / (2.0 * c * c)) * td2 + 0.5 * dtau * (xd2 + yd2) + 0.5 * dtau * r * sq +
               0.25 * dtau * u * sq * sq;
    }
    acc = block_reduce<false>(acc, red);
    if (threadIdx.x == 0) out[b] = acc;
}

// Configuration-stream order of the fields (DetSDW::saveConfigurationStreamBinary, detsdwopdim.cpp:5000-5010):
// out[((ix*L + iy)*m + (k-1))*opdim + dim] = phi(site = iy*L + ix, dim, k), k = 1..m; one CTA row per replica.
__global__ void config_stream_kernel(const double* __restrict__ phi, double* __restrict__ out, int L, int opdim, int m,
                                     long long stridePhi, long long strideOut) {
    pdl_enter();
    const int N = L * L;
    const double* p = phi + size_t(blockIdx.y) * stridePhi;
    double* o = out + size_t(blockIdx.y) * strideOut;
    const int total = N * m * opdim;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int dim = idx % opdim;
        const int k = (idx / opdim) % m + 1;
        const int col = idx / (opdim * m);          // ix*L + iy
        const int ix = col / L, iy = col % L;
        o[idx] = p[(size_t(k) * opdim + dim) * N + iy * L + ix];
    }
}


// get_exchange_action_contribution (detsdwopdim.cpp:5204-5216)
__global__ void exchange_action_kernel(const double* phi, double* out, int N, int opdim, int m, double dtau,
                                       long long stridePhi) {
    pdl_enter();
    __shared__ double red[32];
    const int b = blockIdx.x;
    const double* ph = phi + size_t(b) * stridePhi + size_t(opdim) * N;     // skip slice 0
    double acc = 0;
    for (int idx = threadIdx.x; idx < m * opdim * N; idx += blockDim.x) acc += ph[idx] * ph[idx];
    acc = block_reduce<false>(acc, red);
    if (threadIdx.x == 0) out[b] = 0.5 * dtau * acc;
}

// replica-exchange payload: look-ahead uniforms of local replica 0 taken at its device cursor
// (resident mode), and the control-data blobs of all local replicas
__global__ void exchange_pack_kernel(const double* rng, const int* cursor, int window, double* uni_out, int n_uni,
                                     const dqmc_control_data* ctrl, double* ctrl_out, int R, int copy_uniforms) {
    pdl_enter();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    if (copy_uniforms) {
        const int c0 = cursor[0];
        for (int i = tid; i < n_uni; i += nth) uni_out[i] = (c0 + i < window) ? rng[c0 + i] : -1.0;
    }
    const int words = int(sizeof(dqmc_control_data) / sizeof(double));
    const double* src = reinterpret_cast<const double*>(ctrl);
    for (int i = tid; i < R * words; i += nth) ctrl_out[i] = src[i];
}
__global__ void cursor_advance_kernel(int* cursor, int rep, int n) {
    pdl_enter(); cursor[rep] += n; }
__global__ void cursor_add_kernel(int* cursor, const int* add, int n) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cursor[i] += add[i];
}

}  // namespace

cudaError_t launch_exchange_pack(const double* rng, const int* cursor, int window, double* uni_out, int n_uni,
                                 const dqmc_control_data* ctrl, double* ctrl_out, int R, int copy_uniforms,
                                 cudaStream_t st) {
    launch_pdl(exchange_pack_kernel, dim3(8), dim3(256), 0, st, rng, cursor, window, uni_out, n_uni, ctrl, ctrl_out, R, copy_uniforms);
    return cudaGetLastError();
}
cudaError_t launch_cursor_add(int* cursor, const int* add, int n, cudaStream_t st) {
    launch_pdl(cursor_add_kernel, dim3((n + 127) / 128), dim3(128), 0, st, cursor, add, n);
    return cudaGetLastError();
}
cudaError_t launch_cursor_advance(int* cursor, int rep, int n, cudaStream_t st) {
    launch_pdl(cursor_advance_kernel, dim3(1), dim3(1), 0, st, cursor, rep, n);
    return cudaGetLastError();
}

cudaError_t launch_set_identity(cplx* A, int D, long long stride, int batch, cudaStream_t st) {
    dim3 grid((unsigned)((size_t(D) * D + 255) / 256), batch);
    launch_pdl(set_identity_kernel, dim3(grid), dim3(256), 0, st, A, D, stride);
    return cudaGetLastError();
}
cudaError_t launch_scaled_conj_transpose(const cplx* A, long long strideA, cplx* B, long long strideB,
                                         const double* rowscale, long long strideS, int D, int batch, cudaStream_t st) {
    dim3 grid((D + 31) / 32, (D + 31) / 32, batch), block(32, 8);
    launch_pdl(scaled_conj_transpose_kernel, dim3(grid), dim3(block), 0, st, A, strideA, B, strideB, rowscale, strideS, D);
    return cudaGetLastError();
}

cudaError_t launch_conj_transpose(const cplx* A, cplx* B, int D, long long stride, int batch, cudaStream_t st) {
    dim3 grid((D + 31) / 32, (D + 31) / 32, batch);
    launch_pdl(conj_transpose_kernel, dim3(grid), dim3(dim3(32, 8)), 0, st, A, B, D, stride);
    return cudaGetLastError();
}
cudaError_t launch_max_abs_diff(const cplx* A, const cplx* B, int D, long long stride, int batch, double* out,
                                cudaStream_t st) {
    const int CS = D >= 128 ? 8 : 1;
    cudaError_t e = launch_pdl_cluster(max_abs_diff_kernel, dim3(batch * CS), dim3(512), 0, st, (unsigned)CS, A, B, D, stride, out, CS);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}
cudaError_t launch_update_tables(const double* phi, double* coshT, double* sinhT, int N, int opdim, int m,
                                 double lambda_dtau, long long stridePhi, long long strideTab, int batch,
                                 cudaStream_t st) {
    dim3 grid((unsigned)((size_t(m) * N + 255) / 256), batch);
    launch_pdl(update_tables_kernel, dim3(grid), dim3(256), 0, st, phi, coshT, sinhT, N, opdim, m, lambda_dtau, stridePhi, strideTab);
    return cudaGetLastError();
}
cudaError_t launch_shift_fields(double* phi, const double* shift, int N, int opdim, int m, long long stridePhi,
                                int batch, cudaStream_t st) {
    dim3 grid((unsigned)((size_t(m + 1) * opdim * N + 255) / 256), batch);
    launch_pdl(shift_fields_kernel, dim3(grid), dim3(256), 0, st, phi, shift, N, opdim, m, stridePhi);
    return cudaGetLastError();
}
cudaError_t launch_phi_action(const double* phi, const double* rvals, double* out, int L, int opdim, int m,
                              double dtau, double c, double u, long long stridePhi, int batch, cudaStream_t st) {
    launch_pdl(phi_action_kernel, dim3(batch), dim3(512), 0, st, phi, rvals, out, L, opdim, m, dtau, c, u, stridePhi);
    return cudaGetLastError();
}
cudaError_t launch_config_stream(const double* phi, double* out, int L, int opdim, int m, long long stridePhi,
                                 long long strideOut, int batch, cudaStream_t st) {
    dim3 grid(std::max(1, std::min(64, (L * L * m * opdim + 255) / 256)), batch);
    launch_pdl(config_stream_kernel, dim3(grid), dim3(256), 0, st, phi, out, L, opdim, m, stridePhi, strideOut);
    return cudaGetLastError();
}
cudaError_t launch_exchange_action(const double* phi, double* out, int N, int opdim, int m, double dtau,
                                   long long stridePhi, int batch, cudaStream_t st) {
    launch_pdl(exchange_action_kernel, dim3(batch), dim3(512), 0, st, phi, out, N, opdim, m, dtau, stridePhi);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Fermionic measurements of one time slice (DetSDW::measure, detsdwopdim.cpp:540-900) from the symmetrically shifted
// Green's function gs (column-major D x D): accumulates into the per-replica buffer
//   acc = [ greenK0 | greenLocal | occDiffSq | binsX (cplx, (2L-1)^2) | binsY (cplx) | pairPlus (N) | pairMinus (N) ]
// binsX/Y[dy][dx] = sum over site pairs with r_i - r_j = (dx, dy) of the band-diagonal, spin-summed Green's function:
// the momentum-space occupation (an O(N^3) loop in the reference, :623-671) is their Fourier sum, taken at the end.
// Band-spin blocks XUP = 0, YDOWN = 1, XDOWN = 2, YUP = 3; for MSF = 2 only XUP / YDOWN are stored and the
// XDOWN / YUP sector is their complex conjugate (gl1, :598-615).  One CTA per replica.
// ------------------------------------------------------------------------------------------------
template <int MSF>
__device__ __forceinline__ cplx fm_gl1(const cplx* __restrict__ gs, int D, int N, int s1, int bs1, int s2, int bs2) {
    if (MSF == 4) return gs[size_t(s2 + N * bs2) * D + s1 + N * bs1];
    if (bs1 < 2 && bs2 < 2) return gs[size_t(s2 + N * bs2) * D + s1 + N * bs1];
    if (bs1 >= 2 && bs2 >= 2) {
        const cplx v = gs[size_t(s2 + N * (bs2 - 2)) * D + s1 + N * (bs1 - 2)];
        return make_double2(v.x, -v.y);
    }
    return make_double2(0, 0);
}
__device__ __forceinline__ cplx fm_mul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx fm_add(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx fm_sub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx fm_scale(double f, cplx a) { return make_double2(f * a.x, f * a.y); }

__device__ __forceinline__ double fm_block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < nw; ++w) t += red[w];
    return t;
}

// A thread-block cluster of CS CTAs serves one replica (one SM streams the D x D matrix at ~50 GB/s only): every CTA takes
// a 1 / CS share of the matrix elements into its own shared-memory bins, CTA 0 then adds the bins and the partial sums
// of its peers through distributed shared memory in rank order.
template <int MSF>
__global__ void __launch_bounds__(256) fermion_measure_kernel(const cplx* __restrict__ gsAll, long long strideG, int N,
                                                              int L, double* __restrict__ accAll, long long strideAcc,
                                                              int CS) {
    pdl_enter();
    extern __shared__ double fm_smem[];
    __shared__ double red[8];
    __shared__ double part[3];
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const int rep = blockIdx.x / CS, crank = blockIdx.x % CS;
    const int D = MSF * N, nb = (2 * L - 1) * (2 * L - 1);
    const cplx* gs = gsAll + size_t(rep) * strideG;
    double* acc = accAll + size_t(rep) * strideAcc;
    double* bins = fm_smem;                                  // [2][nb] complex
    const int tid = threadIdx.x;
    const int gtid = crank * blockDim.x + tid, gstep = CS * blockDim.x;
    for (int i = tid; i < 4 * nb; i += blockDim.x) bins[i] = 0.0;
    __syncthreads();
    // scalar functions of the Green's function (:569-593)
    double sre = 0, tr = 0;
    for (size_t e = gtid; e < size_t(D) * D; e += gstep) {
        const cplx v = gs[e];
        sre += v.x;
        if (int(e / D) == int(e % D)) tr += v.x;
    }
    const double f = MSF == 4 ? 1.0 : 2.0;
    const double ssum = fm_block_sum(sre, red), strace = fm_block_sum(tr, red);
    // displacement bins of the band-diagonal, spin-summed Green's function
    for (int e = gtid; e < N * N; e += gstep) {
        const int i = e % N, j = e / N;
        const int dx = i % L - j % L + L - 1, dy = i / L - j / L + L - 1;
        const cplx gx = fm_add(fm_gl1<MSF>(gs, D, N, i, 0, j, 0), fm_gl1<MSF>(gs, D, N, i, 2, j, 2));
        const cplx gy = fm_add(fm_gl1<MSF>(gs, D, N, i, 3, j, 3), fm_gl1<MSF>(gs, D, N, i, 1, j, 1));
        const int bi = dy * (2 * L - 1) + dx;
        atomicAdd(&bins[2 * bi], gx.x);
        atomicAdd(&bins[2 * bi + 1], gx.y);
        atomicAdd(&bins[2 * nb + 2 * bi], gy.x);
        atomicAdd(&bins[2 * nb + 2 * bi + 1], gy.y);
    }
    // equal-time pairing correlations (:673-718): band b, spin sp -> block (b == 0 ? (sp == 0 ? 0 : 2) : (sp == 0 ? 3 : 1))
    double* pairPlus = acc + 3 + 4 * nb;
    double* pairMinus = pairPlus + N;
    for (int i = gtid; i < N; i += gstep) {
        cplx pp = make_double2(0, 0), pm = make_double2(0, 0);
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
            const int a = pr == 0 ? i : 0, b = pr == 0 ? 0 : i;
#pragma unroll
            for (int b1 = 0; b1 < 2; ++b1)
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    const int u1 = b1 == 0 ? 0 : 3, d1 = b1 == 0 ? 2 : 1;       // up / down blocks of band b1
                    const int u2 = b2 == 0 ? 0 : 3, d2 = b2 == 0 ? 2 : 1;
                    const cplx t = fm_sub(fm_mul(fm_gl1<MSF>(gs, D, N, a, d1, b, u2), fm_gl1<MSF>(gs, D, N, a, u1, b, d2)),
                                          fm_mul(fm_gl1<MSF>(gs, D, N, a, d1, b, d2), fm_gl1<MSF>(gs, D, N, a, u1, b, u2)));
                    pp = fm_add(pp, fm_scale(-4.0, t));
                    pm = fm_add(pm, fm_scale(b1 == b2 ? -4.0 : 4.0, t));
                }
        }
        pairPlus[i] += pp.x;
        pairMinus[i] += pm.x;
    }
    // occDiffSq (:744-776)
    double occ = 0;
    for (int i = gtid; i < N; i += gstep) {
        // g(b1, s1, b2, s2) at site i; blocks: XU = 0, YD = 1, XD = 2, YU = 3
        auto g = [&](int bs1, int bs2) { return fm_gl1<MSF>(gs, D, N, i, bs1, i, bs2); };
        const int XU = 0, YD = 1, XD = 2, YU = 3;
        cplx t = fm_scale(-2.0, fm_mul(g(XD, XU), g(XU, XD)));
        t = fm_add(t, g(XU, XU));
        t = fm_add(t, fm_scale(2.0, fm_mul(g(XD, YD), g(YD, XD))));
        t = fm_add(t, fm_scale(2.0, fm_mul(g(XU, YD), g(YD, XU))));
        t = fm_add(t, g(YD, YD));
        t = fm_add(t, fm_scale(-2.0, fm_mul(g(XU, XU), g(YD, YD))));
        t = fm_add(t, fm_scale(2.0, fm_mul(g(XD, YU), g(YU, XD))));
        t = fm_add(t, fm_scale(2.0, fm_mul(g(XU, YU), g(YU, XU))));
        t = fm_add(t, fm_scale(-2.0, fm_mul(g(YD, YU), g(YU, YD))));
        cplx br = make_double2(1.0, 0.0);
        br = fm_add(br, fm_scale(2.0, g(XU, XU)));
        br = fm_add(br, fm_scale(-2.0, g(YD, YD)));
        br = fm_add(br, fm_scale(-2.0, g(YU, YU)));
        t = fm_add(t, fm_mul(g(XD, XD), br));
        t = fm_add(t, g(YU, YU));
        t = fm_add(t, fm_scale(-2.0, fm_mul(g(XU, XU), g(YU, YU))));
        t = fm_add(t, fm_scale(2.0, fm_mul(g(YD, YD), g(YU, YU))));
        occ += t.x;
    }
    const double socc = fm_block_sum(occ, red);
    if (tid == 0) { part[0] = ssum; part[1] = strace; part[2] = socc; }
    __syncthreads();
    if (CS > 1) cluster.sync();                              // every CTA's bins and partial sums are complete
    if (crank == 0) {
        for (int i = tid; i < 4 * nb; i += blockDim.x) {
            double t = bins[i];
            for (int r = 1; r < CS; ++r) t += cluster.map_shared_rank(bins, r)[i];
            acc[3 + i] += t;
        }
        if (tid == 0) {
            double t0 = part[0], t1 = part[1], t2 = part[2];
            for (int r = 1; r < CS; ++r) {
                const double* pr = cluster.map_shared_rank(part, r);
                t0 += pr[0]; t1 += pr[1]; t2 += pr[2];
            }
            acc[0] += f * t0;
            acc[1] += f * t1 / (4.0 * N);
            acc[2] += t2 / N;
        }
    }
    if (CS > 1) cluster.sync();                              // peers stay resident until CTA 0 has read their shared memory
}

cudaError_t launch_fermion_measure(const cplx* gs, long long strideG, int N, int L, int msf, double* acc, long long strideAcc,
                                   int batch, cudaStream_t st) {
    const size_t smem = size_t(4) * (2 * L - 1) * (2 * L - 1) * sizeof(double);
    static const int csEnv = std::getenv("DQMC_MEASURE_CLUSTER") ? std::atoi(std::getenv("DQMC_MEASURE_CLUSTER")) : 0;
    int CS = csEnv > 0 ? csEnv : (msf * N >= 128 ? 8 : 1);
    CS = std::max(1, std::min(CS, 8));
    cudaError_t e;
    if (msf == 4) e = launch_pdl_cluster(fermion_measure_kernel<4>, dim3(batch * CS), dim3(256), smem, st, (unsigned)CS, gs, strideG, N, L, acc, strideAcc, CS);
    else e = launch_pdl_cluster(fermion_measure_kernel<2>, dim3(batch * CS), dim3(256), smem, st, (unsigned)CS, gs, strideG, N, L, acc, strideAcc, CS);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace dqmc
