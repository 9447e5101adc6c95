// kern_lanes_d.cu — instantiations of the small-circuit kernels (solver_lanes.cuh): (lanes per instance, n) = [(4, 13), (4, 14), (4, 15), (4, 16)]
// n is the exact unknown count of the analysis mode: .op uses n_start, .dc and .tran use n_run.
#include "solver_lanes.cuh"

namespace ftsb {
template <int G, int NC>
static KernelFn pick_an(int an)
{
    return an == AN_OP ? (KernelFn)k_sync<G, NC, AN_OP> : an == AN_DC ? (KernelFn)k_sync<G, NC, AN_DC> : (KernelFn)k_lanes<G, NC, AN_TRAN>;
}
KernelFn lanes_kernel_d(int g, int nc, int an)
{
    if (g == 4 && nc == 13) return pick_an<4, 13>(an);
    if (g == 4 && nc == 14) return pick_an<4, 14>(an);
    if (g == 4 && nc == 15) return pick_an<4, 15>(an);
    if (g == 4 && nc == 16) return pick_an<4, 16>(an);
    return nullptr;
}
} // namespace ftsb
