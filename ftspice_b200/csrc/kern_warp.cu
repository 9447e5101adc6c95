// kern_warp.cu — instantiations of the warp-per-instance kernels (solver_warp.cuh): 1, 2 or 3 rows per lane
#include "solver_warp.cuh"

namespace ftsb {
template <int RPT>
static KernelFn pick_an(int an)
{
    return an == AN_OP ? (KernelFn)k_warp<RPT, AN_OP> : an == AN_DC ? (KernelFn)k_warp<RPT, AN_DC> : (KernelFn)k_warp<RPT, AN_TRAN>;
}
KernelFn warp_kernel(int rpt, int an)
{
    switch (rpt) {
    case 1: return pick_an<1>(an);
    case 2: return pick_an<2>(an);
    case 3: return pick_an<3>(an);
    default: return nullptr;
    }
}
} // namespace ftsb
