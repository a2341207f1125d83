// Instantiation of the fused large-A kernel (pbx_big.cuh) for one number of surfaces:
// compiled once per A with -DPBX_BIG_AT=.. so that the translation units build in parallel.
#include <algorithm>

#include "pbx_big.cuh"

namespace pbx {
#define PBX_CAT_(a, b) a##b
#define PBX_CAT(a, b) PBX_CAT_(a, b)
extern const BigLauncher PBX_CAT(big_launcher_, PBX_BIG_AT) = &launch_big_at<PBX_BIG_AT>;
}  // namespace pbx
