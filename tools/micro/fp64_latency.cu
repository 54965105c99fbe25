// Microbenchmark: dependent-issue latency and throughput of fp64 ops on the current GPU (diagnostic only).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void chain(double *out, int iters, long long *cyc) {
    double a[ILP];
    for (int j = 0; j < ILP; ++j) a[j] = 1.0 + threadIdx.x * 1e-9 + j;
    const double b = 0.999999, c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) a[j] = fma(a[j], b, c);
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < ILP; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void divchain(double *out, int iters, long long *cyc) {
    double a = 1.0 + threadIdx.x * 1e-9; long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = 1.0 / (a + 0.5);
    long long t1 = clock64(); out[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int it = 4096;
#define RUN(ILP, W) chain<ILP><<<1, 32 * W>>>(out, it, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("warps/SM=%d ILP=%d: %.2f cycles per dependent DFMA step, %.3f DFMA warp-instr/cycle/SM\n", W, ILP, (double)h / it, (double)ILP * W * it / h);
    RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(1, 4) RUN(1, 8) RUN(1, 16) RUN(1, 32) RUN(4, 4) RUN(4, 8) RUN(8, 8) RUN(2, 16)
    divchain<<<1, 32>>>(out, it, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent fp64 (add + divide): %.1f cycles\n", (double)h / it);
    return 0;
}
