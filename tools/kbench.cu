// Kernel micro-benchmarks (development tool, not part of the product): times the internal batched
// kernels of libdqmc_b200 on synthetic data with CUDA events on the launching stream.
//   tools/kbench [cb] [gemm] [qr] [all] [--R 64] [--L 12] [--reps 20]
// Buffers are rotated over enough copies to exceed the 126 MB L2 between timed launches.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <string>
#include <vector>

#include "../detqmc_b200/csrc/dqmc_internal.h"

using namespace dqmc;

#define CHECK(x)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (x);                                                                           \
        if (e__ != cudaSuccess) {                                                                        \
            std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__));     \
            std::exit(1);                                                                                \
        }                                                                                                \
    } while (0)

static cudaStream_t st;

// average ms per call of fn(i) over reps calls (after warm-up)
static double time_ms(const std::function<void(int)>& fn, int reps, int warm = 3) {
    for (int i = 0; i < warm; ++i) fn(i);
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    CHECK(cudaStreamSynchronize(st));
    CHECK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) fn(warm + i);
    CHECK(cudaEventRecord(e1, st));
    CHECK(cudaEventSynchronize(e1));
    CHECK(cudaGetLastError());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms / reps;
}

static void fill_random(std::vector<double>& v, unsigned seed, double scale = 1.0) {
    std::mt19937_64 g(seed);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    for (auto& x : v) x = scale * u(g);
}

__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCH>
__global__ void dfma_latency_kernel(double* out, long long* cyc, int iters, double a, double b) {
    double c[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) c[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main(int argc, char** argv) {
    int R = 64, L = 12, reps = 20, opdim = 2, m = 100;
    bool do_cb = false, do_gemm = false, do_qr = false, do_peak = false;
    int flux = 1;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        if (s == "cb") do_cb = true;
        else if (s == "gemm") do_gemm = true;
        else if (s == "qr") do_qr = true;
        else if (s == "peak") do_peak = true;
        else if (s == "all") do_cb = do_gemm = do_qr = true;
        else if (s == "--R") R = std::atoi(argv[++i]);
        else if (s == "--L") L = std::atoi(argv[++i]);
        else if (s == "--reps") reps = std::atoi(argv[++i]);
        else if (s == "--opdim") opdim = std::atoi(argv[++i]);
        else if (s == "--noflux") flux = 0;
    }
    CHECK(cudaSetDevice(0));
    CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const int N = L * L, msf = opdim == 3 ? 4 : 2, D = msf * N;
    const size_t dd = size_t(D) * D;
    const int ncopies = std::max(2, int(std::ceil(300e6 / (double(dd) * 16 * R))));
    std::printf("# R=%d L=%d opdim=%d D=%d flux=%d copies=%d (%.0f MB rotating)\n", R, L, opdim, D, flux, ncopies,
                ncopies * dd * 16.0 * R / 1e6);

    cplx* A;
    CHECK(cudaMalloc(&A, sizeof(cplx) * dd * R * ncopies));
    {
        std::vector<double> h(2 * dd * R);
        fill_random(h, 1);
        for (int c = 0; c < ncopies; ++c)
            CHECK(cudaMemcpy(A + size_t(c) * dd * R, h.data(), sizeof(cplx) * dd * R, cudaMemcpyHostToDevice));
    }

    if (do_peak) {
        double* out;
        CHECK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
        {
            long long* dcyc;
            CHECK(cudaMalloc(&dcyc, 8));
            long long h = 0;
            const int iters = 4096;
            for (int warps : {1, 2, 4, 8, 16}) {
                dfma_latency_kernel<1><<<1, 32 * warps, 0, st>>>(out, dcyc, iters, 0.999, 1e-3);
                CHECK(cudaMemcpyAsync(&h, dcyc, 8, cudaMemcpyDeviceToHost, st)); CHECK(cudaStreamSynchronize(st));
                const double c1 = double(h) / iters;
                dfma_latency_kernel<2><<<1, 32 * warps, 0, st>>>(out, dcyc, iters, 0.999, 1e-3);
                CHECK(cudaMemcpyAsync(&h, dcyc, 8, cudaMemcpyDeviceToHost, st)); CHECK(cudaStreamSynchronize(st));
                const double c2 = double(h) / iters;
                dfma_latency_kernel<8><<<1, 32 * warps, 0, st>>>(out, dcyc, iters, 0.999, 1e-3);
                CHECK(cudaMemcpyAsync(&h, dcyc, 8, cudaMemcpyDeviceToHost, st)); CHECK(cudaStreamSynchronize(st));
                const double c8 = double(h) / iters;
                std::printf("dfma on ONE SM, %2d warps: cycles per loop iteration with 1 / 2 / 8 independent chains per thread: %6.1f %6.1f %6.1f\n",
                            warps, c1, c2, c8);
            }
        }
        for (int threads : {128, 256, 512, 1024}) {
            const int blocks = 148 * (2048 / threads);
            const int iters = 4096;
            auto f1 = [&](int) { dfma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 0.999, 1e-3); };
            auto f2 = [&](int) { dmma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 0.999, 1e-3); };
            const double ms1 = time_ms(f1, 5), ms2 = time_ms(f2, 5);
            const double fl1 = 2.0 * 8 * iters * double(blocks) * threads;
            const double fl2 = 2.0 * 8 * iters * double(blocks) * (threads / 32) * 256;
            std::printf("fp64 peak threads/CTA=%4d  DFMA %6.2f TFLOP/s   DMMA m8n8k4 %6.2f TFLOP/s\n", threads,
                        fl1 / (ms1 * 1e-3) / 1e12, fl2 / (ms2 * 1e-3) / 1e12);
        }
    }

    if (do_cb) {
        dqmc_params p;
        std::memset(&p, 0, sizeof p);
        p.model = 0; p.opdim = opdim; p.L = L; p.m = m; p.s = 10; p.bc = 0; p.weakZflux = (opdim < 3) ? flux : 0;
        p.dtau = 0.1; p.lambda = 1; p.txhor = -1; p.txver = -0.5; p.tyhor = 0.5; p.tyver = 1; p.mux = p.muy = -0.5;
        CbGeom g;
        g.L = L; g.N = N; g.msf = msf; g.D = D; g.nplaq = N / 4; g.opdim = opdim; g.m = m; g.lambda_dtau = 0.1;
        std::vector<cplx> tab;
        cb_build_tables(p, tab);
        cplx* dtab;
        CHECK(cudaMalloc(&dtab, tab.size() * sizeof(cplx)));
        CHECK(cudaMemcpy(dtab, tab.data(), tab.size() * sizeof(cplx), cudaMemcpyHostToDevice));
        const size_t sphi = size_t(m + 1) * opdim * N, stab = size_t(m + 1) * N;
        std::vector<double> hphi(sphi * R), hc(stab * R), hs(stab * R);
        fill_random(hphi, 2);
        for (int r = 0; r < R; ++r)
            for (int k = 0; k <= m; ++k)
                for (int s = 0; s < N; ++s) {
                    double n2 = 0;
                    for (int d = 0; d < opdim; ++d) { const double v = hphi[r * sphi + (size_t(k) * opdim + d) * N + s]; n2 += v * v; }
                    const double nr = std::sqrt(n2);
                    hc[r * stab + size_t(k) * N + s] = std::cosh(0.1 * nr);
                    hs[r * stab + size_t(k) * N + s] = std::sinh(0.1 * nr) / nr;
                }
        double *dphi, *dc, *ds;
        CHECK(cudaMalloc(&dphi, hphi.size() * 8));
        CHECK(cudaMalloc(&dc, hc.size() * 8));
        CHECK(cudaMalloc(&ds, hs.size() * 8));
        CHECK(cudaMemcpy(dphi, hphi.data(), hphi.size() * 8, cudaMemcpyHostToDevice));
        CHECK(cudaMemcpy(dc, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice));
        CHECK(cudaMemcpy(ds, hs.data(), hs.size() * 8, cudaMemcpyHostToDevice));
        struct Op { const char* name; int rows, k_then_v, sign_idx, transposed; };
        const Op ops[4] = {{"left   B*A", 0, 1, 0, 0}, {"right  A*B", 1, 0, 0, 1}, {"leftinv B^-1*A", 0, 0, 1, 0},
                           {"rightinv A*B^-1", 1, 1, 1, 1}};
        for (int kc : {1, 10}) {
            for (const Op& o : ops) {
                auto fn = [&](int i) {
                    CbLaunch a;
                    a.A = A + size_t(i % ncopies) * dd * R;
                    a.strideA = (long long)dd;
                    a.phi = dphi; a.coshT = dc; a.sinhT = ds;
                    a.stridePhi = (long long)sphi; a.strideTab = (long long)stab;
                    a.cbtab = dtab;
                    a.real_tables = p.weakZflux ? 0 : 1;
                    a.kfirst = 5; a.kstep = 1; a.kcount = kc;
                    a.rows = o.rows; a.k_then_v = o.k_then_v; a.sign_idx = o.sign_idx; a.transposed = o.transposed;
                    a.colscale = nullptr; a.strideScale = 0;
                    a.batch = R;
                    CHECK(cb_launch(g, a, st));
                };
                const double ms = time_ms(fn, reps);
                const double gb = 2.0 * dd * 16 * R / 1e9;
                const double fl = double(dd) * R * kc * (msf == 2 ? (flux && opdim < 3 ? 112.0 : 64.0) : 80.0);
                std::printf("cb %-16s kcount=%2d  %8.1f us  %7.1f GB/s (algorithmic)  ~%5.2f TFLOP/s\n", o.name, kc, ms * 1e3,
                            gb / (ms * 1e-3), fl / (ms * 1e-3) / 1e12);
            }
        }
    }

    if (do_gemm) {
        cplx *B, *C;
        CHECK(cudaMalloc(&B, sizeof(cplx) * dd * R));
        CHECK(cudaMalloc(&C, sizeof(cplx) * dd * R));
        CHECK(cudaMemcpy(B, A, sizeof(cplx) * dd * R, cudaMemcpyDeviceToDevice));
        for (int ta = 0; ta < 2; ++ta)
            for (int tb = 0; tb < 2; ++tb) {
                auto fn = [&](int i) {
                    GemmArgs ga;
                    ga.M = ga.N = ga.K = D;
                    ga.transa = ta; ga.transb = tb;
                    ga.A = A + size_t(i % ncopies) * dd * R; ga.lda = D; ga.strideA = (long long)dd;
                    ga.B = B; ga.ldb = D; ga.strideB = (long long)dd;
                    ga.C = C; ga.ldc = D; ga.strideC = (long long)dd;
                    ga.rowscale = ga.colscale = ga.kscale = nullptr;
                    ga.strideRow = ga.strideCol = ga.strideK = 0;
                    ga.kvec = nullptr; ga.b_kmajor = 0; ga.alpha = 1.0; ga.beta = 0.0; ga.batch = R;
                    CHECK(gemm_launch(ga, st));
                };
                const double ms = time_ms(fn, reps);
                std::printf("gemm ta=%d tb=%d  %8.1f us  %6.2f TFLOP/s\n", ta, tb, ms * 1e3, 8.0 * D * double(dd) * R / (ms * 1e-3) / 1e12);
            }
        // skinny shapes of the blocked QR: W = V^H A2 (nb x n2, K = rows), A2 -= V W
        for (int nb : {16, 32, 48}) {
            auto fn1 = [&](int i) {
                GemmArgs ga;
                ga.M = nb; ga.N = D; ga.K = D;
                ga.transa = 1; ga.transb = 0;
                ga.A = B; ga.lda = D; ga.strideA = (long long)dd;
                ga.B = A + size_t(i % ncopies) * dd * R; ga.ldb = D; ga.strideB = (long long)dd;
                ga.C = C; ga.ldc = nb; ga.strideC = (long long)dd;
                ga.rowscale = ga.colscale = ga.kscale = nullptr;
                ga.strideRow = ga.strideCol = ga.strideK = 0;
                ga.kvec = nullptr; ga.b_kmajor = 0; ga.alpha = 1.0; ga.beta = 0.0; ga.batch = R;
                CHECK(gemm_launch(ga, st));
            };
            auto fn2 = [&](int i) {
                GemmArgs ga;
                ga.M = D; ga.N = D; ga.K = nb;
                ga.transa = 0; ga.transb = 0;
                ga.A = B; ga.lda = D; ga.strideA = (long long)dd;
                ga.B = C; ga.ldb = nb; ga.strideB = (long long)dd;
                ga.C = A + size_t(i % ncopies) * dd * R; ga.ldc = D; ga.strideC = (long long)dd;
                ga.rowscale = ga.colscale = ga.kscale = nullptr;
                ga.strideRow = ga.strideCol = ga.strideK = 0;
                ga.kvec = nullptr; ga.b_kmajor = 0; ga.alpha = -1.0; ga.beta = 1.0; ga.batch = R;
                CHECK(gemm_launch(ga, st));
            };
            const double ms1 = time_ms(fn1, reps), ms2 = time_ms(fn2, reps);
            const double fl = 8.0 * nb * double(dd) * R;
            std::printf("gemm skinny nb=%2d  W=V^H A: %7.1f us %6.2f TF/s   A-=V W: %7.1f us %6.2f TF/s\n", nb, ms1 * 1e3,
                        fl / (ms1 * 1e-3) / 1e12, ms2 * 1e3, fl / (ms2 * 1e-3) / 1e12);
        }
    }

    if (do_qr) {
        cplx *W, *Q, *tau, *T;
        int* perm;
        double *cn, *dv;
        CHECK(cudaMalloc(&W, sizeof(cplx) * dd * R));
        CHECK(cudaMalloc(&Q, sizeof(cplx) * dd * R));
        CHECK(cudaMalloc(&T, sizeof(cplx) * dd * R));
        CHECK(cudaMalloc(&tau, sizeof(cplx) * D * R));
        CHECK(cudaMalloc(&perm, sizeof(int) * D * R));
        CHECK(cudaMalloc(&cn, sizeof(double) * D * R));
        CHECK(cudaMalloc(&dv, sizeof(double) * D * R));
        QrWorkspace ws;
        CHECK(qr_workspace_create(&ws, D, R));
        auto copy_in = [&](int i) {
            CHECK(cudaMemcpyAsync(W, A + size_t(i % ncopies) * dd * R, sizeof(cplx) * dd * R, cudaMemcpyDeviceToDevice, st));
        };
        const double ms_copy = time_ms(copy_in, reps);
        auto f_unblocked = [&](int i) {
            copy_in(i);
            CHECK(qrcp_factor_launch(W, D, (long long)dd, tau, perm, cn, R, st));
        };
        auto f_unblocked_q = [&](int i) { CHECK(qr_form_q_launch(W, tau, Q, D, (long long)dd, R, st)); };
        auto f_blocked = [&](int i) {
            CHECK(qr_prepivot_launch(A + size_t(i % ncopies) * dd * R, (long long)dd, W, (long long)dd, perm, cn, D, R, st));
            CHECK(qr_blocked_factor(ws, W, D, (long long)dd, 0, R, st));
        };
        auto f_blocked_q = [&](int i) { CHECK(qr_blocked_form_q(ws, Q, D, (long long)dd, 0, R, st)); };
        const double fl = 16.0 / 3.0 * D * double(dd) * R;
        double ms = time_ms(f_unblocked, std::max(2, reps / 4), 1) - ms_copy;
        std::printf("qr unblocked factor   %9.1f us  %6.2f TFLOP/s\n", ms * 1e3, fl / (ms * 1e-3) / 1e12);
        ms = time_ms(f_unblocked_q, std::max(2, reps / 4), 1);
        std::printf("qr unblocked form_q   %9.1f us  %6.2f TFLOP/s\n", ms * 1e3, fl / (ms * 1e-3) / 1e12);
        ms = time_ms(f_blocked, reps);
        std::printf("qr blocked   factor (incl. pre-pivot copy)  %9.1f us  %6.2f TFLOP/s\n", ms * 1e3, fl / (ms * 1e-3) / 1e12);
        ms = time_ms(f_blocked_q, reps);
        std::printf("qr blocked   form_q   %9.1f us  %6.2f TFLOP/s\n", ms * 1e3, fl / (ms * 1e-3) / 1e12);
        auto f_applyqh = [&](int i) { CHECK(qr_blocked_apply_qh(ws, Q, D, D, (long long)dd, 0, R, st)); };
        ms = time_ms(f_applyqh, reps);
        std::printf("qr blocked   apply_qh %9.1f us  %6.2f TFLOP/s\n", ms * 1e3, 8.0 * D * double(dd) * R / (ms * 1e-3) / 1e12);
        auto f_trsm = [&](int i) { CHECK(trsm_upper_launch(W, Q, T, perm, D, (long long)dd, R, st)); };
        ms = time_ms(f_trsm, std::max(2, reps / 4), 1);
        std::printf("trsm unblocked        %9.1f us  %6.2f TFLOP/s\n", ms * 1e3, 4.0 * D * double(dd) * R / (ms * 1e-3) / 1e12);
        auto f_trsm_b = [&](int i) { CHECK(trsm_upper_blocked(ws, W, Q, T, D, (long long)dd, 0, R, st)); };
        ms = time_ms(f_trsm_b, reps);
        std::printf("trsm blocked          %9.1f us  %6.2f TFLOP/s\n", ms * 1e3, 4.0 * D * double(dd) * R / (ms * 1e-3) / 1e12);
    }
    CHECK(cudaStreamSynchronize(st));
    std::printf("# done\n");
    return 0;
}
