// kern_lanes_a.cu — instantiations of the small-circuit kernels (solver_lanes.cuh): (lanes per instance, n) = [(1, 1), (1, 2), (1, 3), (1, 4)]
// n is the exact unknown count of the analysis mode: .op uses n_start, .dc and .tran use n_run.
#include "solver_lanes.cuh"

namespace ftsb {
template <int G, int NC>
static KernelFn pick_an(int an)
{
    return an == AN_OP ? (KernelFn)k_sync<G, NC, AN_OP> : an == AN_DC ? (KernelFn)k_sync<G, NC, AN_DC> : (KernelFn)k_lanes<G, NC, AN_TRAN>;
}
KernelFn lanes_kernel_a(int g, int nc, int an)
{
    if (g == 1 && nc == 1) return pick_an<1, 1>(an);
    if (g == 1 && nc == 2) return pick_an<1, 2>(an);
    if (g == 1 && nc == 3) return pick_an<1, 3>(an);
    if (g == 1 && nc == 4) return pick_an<1, 4>(an);
    return nullptr;
}
} // namespace ftsb
