// kern_sparse.cu — instantiations of the structure-exploiting warp-per-instance kernels (solver_sparse.cuh)
#include "solver_sparse.cuh"

namespace ftsb {
KernelFn sparse_kernel(int an)
{
    return an == AN_OP ? (KernelFn)k_spw<AN_OP> : an == AN_DC ? (KernelFn)k_spw<AN_DC> : (KernelFn)k_spw<AN_TRAN>;
}
} // namespace ftsb
