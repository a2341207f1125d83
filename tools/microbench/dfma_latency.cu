// Dependent-issue latency of DFMA / DMUL / DADD on sm_100a and the throughput of one warp (per scheduler) as a
// function of the number of independent chains: how much ILP does an FP64-bound warp need?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency.bin tools/microbench/dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void chains(double* out, long long* cycles, int iters, double a, double b) {
    double v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = threadIdx.x * 1e-9 + k;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = fma(v[k], a, b);
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += v[k];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int K>
void run(int warps_per_sm_quadrant) {
    double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    const int iters = 1 << 14;
    chains<K><<<1, 32 * 4 * warps_per_sm_quadrant>>>(d, c, iters, 0.999999, 1e-7);   // one SM; w warps per scheduler
    cudaDeviceSynchronize();
    chains<K><<<1, 32 * 4 * warps_per_sm_quadrant>>>(d, c, iters, 0.999999, 1e-7);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%d warp(s)/scheduler, %2d independent DFMA chains per thread: %6.2f cycles per DFMA per warp, pipe use %5.1f %%\n",
           warps_per_sm_quadrant, K, (double)h / ((double)iters * K), 100.0 * 2.0 * iters * K * warps_per_sm_quadrant / (double)h);
    cudaFree(d); cudaFree(c);
}

int main() {
    run<1>(1); run<2>(1); run<4>(1); run<6>(1); run<8>(1); run<12>(1); run<16>(1);
    run<1>(2); run<2>(2); run<4>(2); run<6>(2); run<8>(2); run<12>(2);
    run<4>(4); run<8>(4);
    return 0;
}
