// Development tool: dependent-chain latencies on one warp of one SM (clocks per operation).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <cstdio>
#include <cuda_runtime.h>

#define N 2048
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }

template <int KIND>
__global__ void chain(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = threadIdx.x & 63 ? 0.0 : 1.0;
    double x = a + threadIdx.x * 1e-9;
    int idx = 0;
    __syncthreads();
    long long t0 = clk();
#pragma unroll 1
    for (int i = 0; i < N / 8; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (KIND == 0) x = fma(x, a, b);
        if (KIND == 1) x = exp(-x) + 0.5;
        if (KIND == 2) x = sqrt(x) + 1.0;
        if (KIND == 3) x = 1.0 / x + 0.5;
        if (KIND == 4) { idx = (int)sm[idx]; }                              // LDS.64 + F2I chain
        if (KIND == 5) x = x + a;
        if (KIND == 6) x = x * a;
        if (KIND == 7) { x = fma(x, a, b); asm volatile("bar.sync 1, 32;" ::: "memory"); }
        if (KIND == 8) { x = fma(x, a, b); __syncwarp(); }
        if (KIND == 9) { volatile double* p = sm; x = fma(x, a, p[threadIdx.x & 63 ? 1 : 2]); }   // LDS feeding DFMA
        if (KIND == 10) { double2 v = make_double2(x, b); v.x = fma(v.x, a, fma(-v.y, b, b)); x = v.x; }
        if (KIND == 11) { asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(x), "+d"(b) : "d"(a), "d"(a)); }
        if (KIND == 12) { x = __shfl_xor_sync(0xffffffffu, x, 1) + a; }
      }
    }
    long long t1 = clk();
    out[threadIdx.x] = x + idx;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    const char* names[] = {"dfma", "exp + add", "sqrt + add", "div + add", "lds + f2i", "dadd", "dmul", "dfma + bar.sync(1 warp)",
                           "dfma + syncwarp", "lds -> dfma", "cfma-like", "dmma m8n8k4", "shfl + dadd"};
#define RUN(K) for (int th : {32, 128}) { chain<K><<<1, th>>>(out, cyc, 0.999, 1e-3); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("%-28s %3d threads: %7.1f clocks per iteration\n", names[K], th, double(h) / N); }
    RUN(0) RUN(5) RUN(6) RUN(1) RUN(2) RUN(3) RUN(4) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
