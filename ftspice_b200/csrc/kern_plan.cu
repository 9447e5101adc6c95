// kern_plan.cu — instantiations of the plan-interpreting kernels (solver_plan.cuh): 1, 2 or 4 lanes per instance
#include "solver_plan.cuh"

namespace ftsb {
template <int G>
static KernelFn pick(int an)
{
    return an == AN_OP ? (KernelFn)k_plan<G, AN_OP> : an == AN_DC ? (KernelFn)k_plan<G, AN_DC> : nullptr;
}
KernelFn plan_kernel(int g, int an) { return g == 1 ? pick<1>(an) : g == 2 ? pick<2>(an) : g == 4 ? pick<4>(an) : nullptr; }
} // namespace ftsb
