// Lookup of the compiled register-resident kernels by model shape (see shapes.def).
#include "pbx_fast.cuh"

namespace pbx {
#define PBX_SHAPE(A, N, AR) extern const FastKernelEntry fast_entry_##A##_##N##_##AR;
#include "shapes.def"
#undef PBX_SHAPE

const FastKernelEntry* find_fast_kernel(int A, int N, int AR) {
    static const FastKernelEntry* const table[] = {
#define PBX_SHAPE(A, N, AR) &fast_entry_##A##_##N##_##AR,
#include "shapes.def"
#undef PBX_SHAPE
    };
    for (const FastKernelEntry* e : table)
        if (e->A == A && e->N == N && e->AR == AR) return e;
    return nullptr;
}
}  // namespace pbx
