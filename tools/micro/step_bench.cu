// Microbenchmark: throughput of the candidate-step arithmetic (du, clamp, stage cost, step) with NO global memory
// traffic in the loop, as a function of resident warps per SM.  Tells how far fwd_kernel is from its compute ceiling.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../differentiable_tube_mpc_b200/csrc/dtmpc_math.cuh"
using namespace dtmpc;
template <int NC> __global__ void bench(const __grid_constant__ EnvT<double> E, const __grid_constant__ OcpT<double> O,
                                        const double *theta, double *out, int iters) {
    Theta<double> th; load_theta(th, theta, O);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    State<double> x[NC]; double J[NC];
    for (int q = 0; q < NC; ++q) { x[q] = State<double>{0.3 + 1e-3 * (tid % 97) + q, 0.2 + 1e-3 * (tid % 89), 0.7, 0.05, 0.05}; J[q] = 0; }
    for (int k = 0; k < iters; ++k) {
        double u0[NC], u1[NC];
        for (int q = 0; q < NC; ++q) {
            const double d0 = x[q].px - 1.0, d1 = x[q].py - 1.0, d2 = x[q].th - 0.5, d3 = x[q].b - 0.1;
            u0[q] = M<double>::fma_(0.5, 0.1 + 0.2 * d0 + 0.1 * d1 + 0.05 * d2 + 0.01 * d3, 5.0);
            u1[q] = M<double>::fma_(0.5, 0.01 * d0 - 0.02 * d1 + 0.03 * d2 + 0.01 * d3, 0.01);
            clamp_u(E, u0[q], u1[q]);
            J[q] += stage_cost(th, x[q].px, x[q].py, x[q].th, x[q].b, u0[q], u1[q], 10.0, 10.0, 0.78, 0.0, 0.0);
        }
        step_multi<double, NC>(E, th, x, u0, u1);
        if ((k & 63) == 63) for (int q = 0; q < NC; ++q) { x[q].px -= 3.0; x[q].py -= 3.0; }   // stay in the arena
    }
    double s = 0; for (int q = 0; q < NC; ++q) s += J[q] + x[q].px;
    out[tid] = s;
}
int main() {
    EnvT<double> E{}; OcpT<double> O{};
    E.dt = 0.01; E.ndt = -0.01; E.umin0 = -10; E.umax0 = 10; E.umin1 = -3.14; E.umax1 = 3.14; E.nbeta = -20; E.ninv_beta = -0.05;
    E.n_obs = 5; E.agg = 0;
    const double cx[5] = {4, 2, 4, 8, 6}, cy[5] = {2, 4, 8, 4, 6};
    for (int i = 0; i < 5; ++i) { E.cx[i] = cx[i]; E.cy[i] = cy[i]; E.r2[i] = 1.0; }
    O.T = 50; O.eps_dyn = O.eps_jac = 1e-4;
    double h_th[12] = {1, 1, 0, 1, 1, 1000, 1000, 1000, 1, 0, 0, 0}, *d_th, *out;
    cudaMalloc(&d_th, sizeof(h_th)); cudaMemcpy(d_th, h_th, sizeof(h_th), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 8 * 148 * 2048);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
#define RUN(NC, W) { bench<NC><<<148, 32 * W>>>(E, O, d_th, out, 100); cudaEventRecord(e0); bench<NC><<<148, 32 * W>>>(E, O, d_th, out, iters); \
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); \
    printf("NC=%d warps/SM=%2d: %.1f G candidate-steps/s (%.0f cycles per warp-step per SMSP-warp)\n", NC, W, \
           1e-6 * NC * 148.0 * 32 * W * iters / ms, ms * 1e-3 * 1.93e9 / iters); }
    RUN(1, 4) RUN(1, 8) RUN(1, 12) RUN(1, 16) RUN(1, 20) RUN(1, 24) RUN(1, 32) RUN(2, 8) RUN(2, 12) RUN(2, 16) RUN(3, 8) RUN(3, 12)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
