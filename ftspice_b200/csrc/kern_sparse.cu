// kern_sparse.cu — instantiations of the structure-exploiting warp-per-instance kernels (solver_sparse.cuh)
#include "solver_sparse.cuh"

namespace ftsb {
template <bool WIN>
static KernelFn pick(int an)
{
    return an == AN_OP ? (KernelFn)k_spw<AN_OP, WIN> : an == AN_DC ? (KernelFn)k_spw<AN_DC, WIN> : (KernelFn)k_spw<AN_TRAN, WIN>;
}
KernelFn sparse_kernel(int an, bool windowed) { return windowed ? pick<true>(an) : pick<false>(an); }
} // namespace ftsb
