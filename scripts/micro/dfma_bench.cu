// Microbenchmark (development aid, round 2): dependent-DFMA latency and per-SM DFMA throughput on sm_100a as a
// function of the independent chains per thread and the warps per SM.  nvcc -arch=sm_100a -O3 dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-9 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void run(int warps_per_sm, double *d)
{
    const int iters = 2000, sms = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 32 * (warps_per_sm > 32 ? 32 : warps_per_sm), blocks = sms * (warps_per_sm > 32 ? warps_per_sm / 32 : 1);
    k<CH><<<blocks, threads>>>(d, 10, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<CH><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * 1.965e9;
    const double dfma_per_warp = (double)iters * 8 * CH;
    printf("chains %d warps/SM %2d: %.1f cycles per dependent DFMA step, %.2f warp-DFMA/cycle/SM\n", CH, warps_per_sm,
           cycles / (iters * 8.0), dfma_per_warp * warps_per_sm / cycles);
}
int main()
{
    double *d;
    cudaMalloc(&d, 148 * 2048 * sizeof(double));
    for (int w : {1, 4, 8, 12, 16, 32, 64}) { run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); run<16>(w, d); }
    return 0;
}
