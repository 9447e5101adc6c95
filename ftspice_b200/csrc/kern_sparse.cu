// kern_sparse.cu — instantiations of the structure-exploiting warp-per-instance kernels (solver_sparse.cuh)
#include "solver_sparse.cuh"

namespace ftsb {
template <bool WIN, int GLOB>
static KernelFn pick(int an)
{
    return an == AN_OP ? (KernelFn)k_spw<AN_OP, WIN, GLOB> : an == AN_DC ? (KernelFn)k_spw<AN_DC, WIN, GLOB> : (KernelFn)k_spw<AN_TRAN, WIN, GLOB>;
}
// global_scratch: 0 = none, 1 = per-step history only (.tran), 2 = history + slot table + vectors
KernelFn sparse_kernel(int an, bool windowed, int global_scratch)
{
    if (global_scratch == 2) return windowed ? pick<true, 2>(an) : pick<false, 2>(an);
    if (global_scratch == 1 && an == AN_TRAN) return windowed ? pick<true, 1>(an) : pick<false, 1>(an);
    return windowed ? pick<true, 0>(an) : pick<false, 0>(an);
}
} // namespace ftsb
