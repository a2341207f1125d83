// Latency of cp.async.bulk (TMA 1-D bulk copy, global -> shared, mbarrier completion) as a function of the copy size,
// of how many CTAs copy at once and of whether they read the same or different addresses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_copy_latency.bin bulk_copy_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const double* src, long long stride_doubles, int bytes, int reps, long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
    double* dst = reinterpret_cast<double*>(sm + 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long total = 0, worst = 0;
    const double* my = src + (long long)blockIdx.x * stride_doubles;
    for (int r = 0; r < reps; ++r) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(dst)), "l"(my + (r % 8) * (bytes / 8)), "r"(bytes), "r"(s32(bar)) : "memory");
            asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}"
                         ::"r"(s32(bar)), "r"(r & 1) : "memory");
            const long long dt = clock64() - t0;
            total += dt; worst = dt > worst ? dt : worst;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = total / reps; out[2 * blockIdx.x + 1] = worst; }
}

int main() {
    double* src; long long* out;
    const size_t n = (size_t)64 << 20;
    cudaMalloc(&src, n); cudaMemset(src, 0, n); cudaMalloc(&out, 4096 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    long long h[4096];
    for (int ctas : {1, 148}) for (int same : {1, 0}) for (int bytes : {256, 2560, 10240, 40960}) {
        const long long stride = same ? 0 : (long long)(8 * bytes / 8 + 4096) ;
        probe<<<ctas, 128, 128 + bytes>>>(src, stride, bytes, 200, out);
        probe<<<ctas, 128, 128 + bytes>>>(src, stride, bytes, 200, out);
        cudaMemcpy(h, out, 2 * ctas * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0; long long worst = 0;
        for (int b = 0; b < ctas; ++b) { avg += h[2 * b]; worst = h[2 * b + 1] > worst ? h[2 * b + 1] : worst; }
        printf("CTAs %3d  %s addresses  %6d bytes: avg %7.0f cycles  worst %lld  (%s)\n", ctas, same ? "same     " : "different", bytes,
               avg / ctas, worst, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
