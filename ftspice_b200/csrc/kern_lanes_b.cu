// kern_lanes_b.cu — instantiations of the small-circuit kernels (solver_lanes.cuh): (lanes per instance, n) = [(2, 5), (2, 6), (2, 7), (2, 8)]
// n is the exact unknown count of the analysis mode: .op uses n_start, .dc and .tran use n_run.
#include "solver_lanes.cuh"

namespace ftsb {
template <int G, int NC>
static KernelFn pick_an(int an)
{
    return an == AN_OP ? (KernelFn)k_sync<G, NC, AN_OP> : an == AN_DC ? (KernelFn)k_sync<G, NC, AN_DC> : (KernelFn)k_lanes<G, NC, AN_TRAN>;
}
KernelFn lanes_kernel_b(int g, int nc, int an)
{
    if (g == 2 && nc == 5) return pick_an<2, 5>(an);
    if (g == 2 && nc == 6) return pick_an<2, 6>(an);
    if (g == 2 && nc == 7) return pick_an<2, 7>(an);
    if (g == 2 && nc == 8) return pick_an<2, 8>(an);
    return nullptr;
}
} // namespace ftsb
