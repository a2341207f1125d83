// FP64 tensor (mma.sync.m8n8k4) throughput as a function of resident warps per SM and of independent accumulator
// chains per warp: can 2 warps per scheduler saturate the FP64 units, and how much ILP does one warp need?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_occupancy.bin dmma_occupancy.cu && ./dmma_occupancy.bin
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool INTERLEAVE_DFMA>
__global__ void __launch_bounds__(128) probe(double* out, int iters, double a, double b) {
    double c[ILP][2];
    double v[4] = {1.0, 2.0, 3.0, 4.0};
#pragma unroll
    for (int k = 0; k < ILP; ++k) { c[k][0] = threadIdx.x * 1e-9; c[k][1] = k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
            if (INTERLEAVE_DFMA) v[k & 3] = fma(v[k & 3], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1];
    s += v[0] + v[1] + v[2] + v[3];
    if (s == 123.456) out[0] = s;
}

template <int ILP, bool MIX>
void run(int warps_per_sm, int sms) {
    double* d;
    cudaMalloc(&d, 8);
    const int threads = 128;                       // 4 warps per CTA (one per scheduler), up to 255 registers per thread
    const int ctas_per_sm = warps_per_sm / 4;
    const int iters = 20000 / ILP * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<ILP, MIX><<<sms * ctas_per_sm, threads>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    const double flops = 2.0 * 256.0 * ILP * (double)iters * warps_per_sm * sms;
    printf("warps/SM %2d  ILP %2d  mix %d : %7.2f TFLOP/s (DMMA only)\n", warps_per_sm, ILP, (int)MIX, flops / (best * 1e-3) / 1e12);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    for (int w : {4, 8, 12, 16}) {
        run<1, false>(w, sms); run<2, false>(w, sms); run<4, false>(w, sms); run<8, false>(w, sms); run<20, false>(w, sms);
    }
    for (int w : {4, 8, 16}) { run<4, true>(w, sms); run<20, true>(w, sms); }
    return 0;
}
