// kern_lanes_e.cu — instantiations of the small-circuit kernels (solver_lanes.cuh): (lanes per instance, n) = [(1, 14), (2, 14), (1, 5), (4, 5), (8, 14), (8, 5)]
// n is the exact unknown count of the analysis mode: .op uses n_start, .dc and .tran use n_run.
#include "solver_lanes.cuh"

namespace ftsb {
template <int G, int NC>
static KernelFn pick_an(int an)
{
    return an == AN_OP ? (KernelFn)k_sync<G, NC, AN_OP> : an == AN_DC ? (KernelFn)k_sync<G, NC, AN_DC> : (KernelFn)k_lanes<G, NC, AN_TRAN>;
}
KernelFn lanes_kernel_e(int g, int nc, int an)
{
    if (g == 1 && nc == 14) return pick_an<1, 14>(an);
    if (g == 2 && nc == 14) return pick_an<2, 14>(an);
    if (g == 1 && nc == 5) return pick_an<1, 5>(an);
    if (g == 4 && nc == 5) return pick_an<4, 5>(an);
    if (g == 8 && nc == 14) return pick_an<8, 14>(an);
    if (g == 8 && nc == 5) return pick_an<8, 5>(an);
    return nullptr;
}
} // namespace ftsb
