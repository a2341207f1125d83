// Does a non-FP64 instruction cost FP64 throughput on sm_100a?  K independent DFMA chains per thread
// interleaved with M independent integer (IMAD / LOP3) or FP32 (FFMA) chain steps per iteration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue tools/microbench/fp64_issue.cu && ./fp64_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int M, int KIND>
__global__ void __launch_bounds__(256) mix_kernel(double* out, int iters, double a, double b, unsigned mul, float fa) {
    double v[K];
    unsigned u[M > 0 ? M : 1];
    float f[M > 0 ? M : 1];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = threadIdx.x * 1e-9 + k;
#pragma unroll
    for (int m = 0; m < M; ++m) { u[m] = threadIdx.x + m; f[m] = threadIdx.x * 1e-3f + m; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = fma(v[k], a, b);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            if (KIND == 0) u[m] = u[m] * mul + 12345u;            // IMAD
            else if (KIND == 1) u[m] = (u[m] ^ mul) + (u[m] >> 3);  // LOP3 / SHF / IADD
            else f[m] = fmaf(f[m], fa, 1e-3f);                      // FFMA
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += v[k];
#pragma unroll
    for (int m = 0; m < M; ++m) s += (KIND == 2) ? (double)f[m] : (double)u[m];
    if (s == 123.456) out[0] = s;
}

template <int K, int M, int KIND>
void run(const char* name, int sms) {
    double* d; cudaMalloc(&d, 8);
    const int grid = sms * 8, iters = 1 << 14;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        mix_kernel<K, M, KIND><<<grid, 256>>>(d, iters, 0.999999, 1e-7, 2654435761u, 0.9999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    const double dfma = (double)K * iters * 256.0 * grid;
    printf("%-8s K=%d DFMA + M=%2d other: %8.3f ms  %6.2f TFLOP/s fp64   (other/DFMA = %.2f)\n", name, K, M, best,
           2.0 * dfma / (best * 1e-3) / 1e12, (double)M / K);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    run<8, 0, 0>("none", sms);
    run<8, 2, 0>("imad", sms); run<8, 4, 0>("imad", sms); run<8, 8, 0>("imad", sms); run<8, 16, 0>("imad", sms);
    run<8, 4, 1>("lop/shf", sms); run<8, 8, 1>("lop/shf", sms);
    run<8, 4, 2>("ffma", sms); run<8, 8, 2>("ffma", sms); run<8, 16, 2>("ffma", sms);
    return 0;
}
