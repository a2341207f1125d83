// Instantiation registry of the register-resident small-A kernels (pbx_fast.cuh).
// One translation unit per (A, N, A_rho) shape so that `make -j` compiles them in parallel:
// this file is compiled once per shape with -DPBX_A=.. -DPBX_N=.. -DPBX_AR=..
#include "pbx_fast_ws.cuh"

namespace pbx {
#define PBX_CAT_(a, b, c, d) a##b##_##c##_##d
#define PBX_CAT(a, b, c, d) PBX_CAT_(a, b, c, d)
extern const FastKernelEntry PBX_CAT(fast_entry_, PBX_A, PBX_N, PBX_AR) = make_entry<PBX_A, PBX_N, PBX_AR>();
}  // namespace pbx

#ifdef PBX_JIT_LIBRARY
// a shape compiled at run time into its own shared library (pibronic_b200/jit.py), loaded with pbx_register_shape_library
extern "C" const pbx::FastKernelEntry* pbx_jit_entry(void) { return &pbx::PBX_CAT(fast_entry_, PBX_A, PBX_N, PBX_AR); }
#endif
