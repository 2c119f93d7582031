// Small batched helper kernels: identity, conjugate transpose, max|A-B|, cosh/sinh tables,
// global field shift, bosonic action and exchange action reductions.
#include "dqmc_internal.h"

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
    cplx* a = A + size_t(blockIdx.y) * stride;
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(D) * D) return;
    const int i = idx % D, j = idx / D;
    a[idx] = make_double2(i == j ? 1.0 : 0.0, 0.0);
}

__global__ void conj_transpose_kernel(const cplx* A, cplx* B, int D, long long stride) {
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

__global__ void max_abs_diff_kernel(const cplx* A, const cplx* B, int D, long long stride, double* out) {
    __shared__ double red[32];
    const cplx* a = A + size_t(blockIdx.x) * stride;
    const cplx* b = B + size_t(blockIdx.x) * stride;
    double mx = 0;
    for (size_t idx = threadIdx.x; idx < size_t(D) * D; idx += blockDim.x) {
        const double dx = a[idx].x - b[idx].x, dy = a[idx].y - b[idx].y;
        mx = fmax(mx, sqrt(dx * dx + dy * dy));
    }
    mx = block_reduce<true>(mx, red);
    if (threadIdx.x == 0) out[blockIdx.x] = mx;
}

// coshTermPhi / sinhTermPhi for every (slice >= 1, site)   (detsdwopdim.cpp:1131-1136, 1174-1181)
__global__ void update_tables_kernel(const double* phi, double* coshT, double* sinhT, int N, int opdim, int m,
                                     double lambda_dtau, long long stridePhi, long long strideTab) {
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
    const int b = blockIdx.y;
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(m + 1) * opdim * N) return;
    const int d = (idx / N) % opdim;
    phi[size_t(b) * stridePhi + idx] += shift[b * 3 + d];
}

// phiAction (detsdwopdim.cpp:4242-4299), one CTA per replica, fixed reduction order
__global__ void phi_action_kernel(const double* phi, const double* rvals, double* out, int L, int opdim, int m,
                                  double dtau, double c, double u, long long stridePhi) {
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
__global__ void cursor_advance_kernel(int* cursor, int rep, int n) { cursor[rep] += n; }

}  // namespace

cudaError_t launch_exchange_pack(const double* rng, const int* cursor, int window, double* uni_out, int n_uni,
                                 const dqmc_control_data* ctrl, double* ctrl_out, int R, int copy_uniforms,
                                 cudaStream_t st) {
    exchange_pack_kernel<<<8, 256, 0, st>>>(rng, cursor, window, uni_out, n_uni, ctrl, ctrl_out, R, copy_uniforms);
    return cudaGetLastError();
}
cudaError_t launch_cursor_advance(int* cursor, int rep, int n, cudaStream_t st) {
    cursor_advance_kernel<<<1, 1, 0, st>>>(cursor, rep, n);
    return cudaGetLastError();
}

cudaError_t launch_set_identity(cplx* A, int D, long long stride, int batch, cudaStream_t st) {
    dim3 grid((unsigned)((size_t(D) * D + 255) / 256), batch);
    set_identity_kernel<<<grid, 256, 0, st>>>(A, D, stride);
    return cudaGetLastError();
}
cudaError_t launch_scaled_conj_transpose(const cplx* A, long long strideA, cplx* B, long long strideB,
                                         const double* rowscale, long long strideS, int D, int batch, cudaStream_t st) {
    dim3 grid((D + 31) / 32, (D + 31) / 32, batch), block(32, 8);
    scaled_conj_transpose_kernel<<<grid, block, 0, st>>>(A, strideA, B, strideB, rowscale, strideS, D);
    return cudaGetLastError();
}

cudaError_t launch_conj_transpose(const cplx* A, cplx* B, int D, long long stride, int batch, cudaStream_t st) {
    dim3 grid((D + 31) / 32, (D + 31) / 32, batch);
    conj_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(A, B, D, stride);
    return cudaGetLastError();
}
cudaError_t launch_max_abs_diff(const cplx* A, const cplx* B, int D, long long stride, int batch, double* out,
                                cudaStream_t st) {
    max_abs_diff_kernel<<<batch, 512, 0, st>>>(A, B, D, stride, out);
    return cudaGetLastError();
}
cudaError_t launch_update_tables(const double* phi, double* coshT, double* sinhT, int N, int opdim, int m,
                                 double lambda_dtau, long long stridePhi, long long strideTab, int batch,
                                 cudaStream_t st) {
    dim3 grid((unsigned)((size_t(m) * N + 255) / 256), batch);
    update_tables_kernel<<<grid, 256, 0, st>>>(phi, coshT, sinhT, N, opdim, m, lambda_dtau, stridePhi, strideTab);
    return cudaGetLastError();
}
cudaError_t launch_shift_fields(double* phi, const double* shift, int N, int opdim, int m, long long stridePhi,
                                int batch, cudaStream_t st) {
    dim3 grid((unsigned)((size_t(m + 1) * opdim * N + 255) / 256), batch);
    shift_fields_kernel<<<grid, 256, 0, st>>>(phi, shift, N, opdim, m, stridePhi);
    return cudaGetLastError();
}
cudaError_t launch_phi_action(const double* phi, const double* rvals, double* out, int L, int opdim, int m,
                              double dtau, double c, double u, long long stridePhi, int batch, cudaStream_t st) {
    phi_action_kernel<<<batch, 512, 0, st>>>(phi, rvals, out, L, opdim, m, dtau, c, u, stridePhi);
    return cudaGetLastError();
}
cudaError_t launch_config_stream(const double* phi, double* out, int L, int opdim, int m, long long stridePhi,
                                 long long strideOut, int batch, cudaStream_t st) {
    dim3 grid(std::max(1, std::min(64, (L * L * m * opdim + 255) / 256)), batch);
    config_stream_kernel<<<grid, 256, 0, st>>>(phi, out, L, opdim, m, stridePhi, strideOut);
    return cudaGetLastError();
}
cudaError_t launch_exchange_action(const double* phi, double* out, int N, int opdim, int m, double dtau,
                                   long long stridePhi, int batch, cudaStream_t st) {
    exchange_action_kernel<<<batch, 512, 0, st>>>(phi, out, N, opdim, m, dtau, stridePhi);
    return cudaGetLastError();
}

}  // namespace dqmc
