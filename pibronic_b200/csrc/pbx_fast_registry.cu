// Lookup of the register-resident kernels by model shape: those compiled into the library (shapes.def) and those
// registered at run time (shapes compiled on demand into their own shared library, pibronic_b200/jit.py).
#include <mutex>
#include <vector>

#include "pbx_fast.cuh"

namespace pbx {
#define PBX_SHAPE(A, N, AR) extern const FastKernelEntry fast_entry_##A##_##N##_##AR;
#include "shapes.def"
#undef PBX_SHAPE

namespace {
std::mutex g_dynamic_mutex;
std::vector<const FastKernelEntry*>& dynamic_entries() {
    static std::vector<const FastKernelEntry*> v;
    return v;
}
}  // namespace

const FastKernelEntry* find_fast_kernel(int A, int N, int AR) {
    static const FastKernelEntry* const table[] = {
#define PBX_SHAPE(A, N, AR) &fast_entry_##A##_##N##_##AR,
#include "shapes.def"
#undef PBX_SHAPE
    };
    for (const FastKernelEntry* e : table)
        if (e->A == A && e->N == N && e->AR == AR) return e;
    std::lock_guard<std::mutex> lock(g_dynamic_mutex);
    for (const FastKernelEntry* e : dynamic_entries())
        if (e->A == A && e->N == N && e->AR == AR) return e;
    return nullptr;
}

bool register_fast_kernel(const FastKernelEntry* entry) {
    if (!entry || find_fast_kernel(entry->A, entry->N, entry->AR)) return false;
    std::lock_guard<std::mutex> lock(g_dynamic_mutex);
    dynamic_entries().push_back(entry);
    return true;
}
}  // namespace pbx
