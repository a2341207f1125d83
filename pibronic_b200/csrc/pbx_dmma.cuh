// FP64 tensor-core (mma.sync.m8n8k4.f64) building blocks shared by the blocked kernels (pbx_mid.cuh) and the fused
// large-A kernel (pbx_big.cuh).
#pragma once
#include "pbx_device.cuh"

namespace pbx {

// tile counts of an AT x AT matrix product and of the packed symmetric entries, for m8n8k4 tiles
template <int AT> struct MidShape {
    static constexpr int MT = (AT + 7) / 8, KS = (AT + 3) / 4;      // output tiles per dimension, k-steps
    static constexpr int AA = AT * (AT + 1) / 2, NT = (AA + 7) / 8, AA2 = AT * AT;   // NT: 8-wide mma tiles over the packed entries
};

// D(8x8) += A(8x4) B(4x8): lane (g = lane/4, c = lane%4) supplies A[g][c] and B[c][g] and owns D[g][2c], D[g][2c+1]
__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

}  // namespace pbx
