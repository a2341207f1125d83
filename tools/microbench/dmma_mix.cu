// Is the FP64 tensor path (mma.sync m8n8k4 f64, "DMMA") a pipe of its own on sm_100a, i.e. does it run
// concurrently with vector DFMA?   DMMA-only, DFMA-only and mixed loops, 8 CTAs of 256 threads per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_mix.bin tools/microbench/dmma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NM, int NF>
__global__ void __launch_bounds__(256) mix(double* out, int iters, double a, double b) {
    double c[NM > 0 ? NM : 1][2], v[NF > 0 ? NF : 1];
#pragma unroll
    for (int k = 0; k < NM; ++k) { c[k][0] = threadIdx.x * 1e-9; c[k][1] = k; }
#pragma unroll
    for (int k = 0; k < NF; ++k) v[k] = threadIdx.x * 1e-9 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NM; ++k) dmma(c[k][0], c[k][1], a, b);
#pragma unroll
        for (int k = 0; k < NF; ++k) v[k] = fma(v[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NM; ++k) s += c[k][0] + c[k][1];
#pragma unroll
    for (int k = 0; k < NF; ++k) s += v[k];
    if (s == 123.456) out[0] = s;
}

template <int NM, int NF>
void run(int sms) {
    double* d; cudaMalloc(&d, 8);
    const int grid = sms * 8, iters = 1 << 13;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        mix<NM, NF><<<grid, 256>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    const double warps = 8.0 * grid, fma_mma = (double)NM * iters * warps * 256.0, fma_vec = (double)NF * iters * warps * 32.0;
    printf("DMMA x%d + DFMA x%2d per iteration: %8.3f ms   tensor %6.2f TFLOP/s   vector %6.2f TFLOP/s   total %6.2f\n", NM, NF, best,
           2 * fma_mma / (best * 1e-3) / 1e12, 2 * fma_vec / (best * 1e-3) / 1e12, 2 * (fma_mma + fma_vec) / (best * 1e-3) / 1e12);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run<0, 8>(sms); run<8, 0>(sms); run<4, 0>(sms); run<2, 0>(sms);
    run<8, 8>(sms); run<4, 16>(sms); run<2, 16>(sms); run<1, 16>(sms); run<1, 8>(sms);
    return 0;
}
