// kern_sparse.cu — instantiations of the structure-exploiting warp-per-instance kernels (solver_sparse.cuh)
#include "solver_sparse.cuh"

namespace ftsb {
template <bool WIN, bool GLOB>
static KernelFn pick(int an)
{
    return an == AN_OP ? (KernelFn)k_spw<AN_OP, WIN, GLOB> : an == AN_DC ? (KernelFn)k_spw<AN_DC, WIN, GLOB> : (KernelFn)k_spw<AN_TRAN, WIN, GLOB>;
}
KernelFn sparse_kernel(int an, bool windowed, bool global_scratch)
{
    if (global_scratch) return windowed ? pick<true, true>(an) : pick<false, true>(an);
    return windowed ? pick<true, false>(an) : pick<false, false>(an);
}
} // namespace ftsb
