// kern_lanes_c.cu — instantiations of the small-circuit kernels (solver_lanes.cuh): (lanes per instance, n) = [(4, 9), (4, 10), (4, 11), (4, 12)]
// n is the exact unknown count of the analysis mode: .op uses n_start, .dc and .tran use n_run.
#include "solver_lanes.cuh"

namespace ftsb {
template <int G, int NC>
static KernelFn pick_an(int an)
{
    return an == AN_OP ? (KernelFn)k_sync<G, NC, AN_OP> : an == AN_DC ? (KernelFn)k_sync<G, NC, AN_DC> : (KernelFn)k_lanes<G, NC, AN_TRAN>;
}
KernelFn lanes_kernel_c(int g, int nc, int an)
{
    if (g == 4 && nc == 9) return pick_an<4, 9>(an);
    if (g == 4 && nc == 10) return pick_an<4, 10>(an);
    if (g == 4 && nc == 11) return pick_an<4, 11>(an);
    if (g == 4 && nc == 12) return pick_an<4, 12>(an);
    return nullptr;
}
} // namespace ftsb
