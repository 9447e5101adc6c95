// kern_team.cu — instantiations of the per-team .tran kernels (solver_team.cuh)
#include "solver_team.cuh"

namespace ftsb {
template <int PLACE>
static TeamKernelFn pick(int g)
{
    switch (g) {
    case 2:
        return (TeamKernelFn)k_team_tran<2, PLACE>;
    case 4:
        return (TeamKernelFn)k_team_tran<4, PLACE>;
    case 8:
        return (TeamKernelFn)k_team_tran<8, PLACE>;
    case 16:
        return (TeamKernelFn)k_team_tran<16, PLACE>;
    case 32:
        return (TeamKernelFn)k_team_tran<32, PLACE>;
    default:
        return nullptr;
    }
}
// place: bit 0 = slot table in shared memory, bit 1 = factors in shared memory
TeamKernelFn team_kernel(int g, int place)
{
    switch (place & 3) {
    case 0:
        return pick<0>(g);
    case 1:
        return pick<1>(g);
    case 3:
        return pick<3>(g);
    default:
        return nullptr;
    }
}
} // namespace ftsb
