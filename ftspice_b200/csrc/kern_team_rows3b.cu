// kern_team_rows3b.cu — per-team .tran kernels for LINEAR circuits with the rows per lane known at compile time
// (solver_team.cuh, RC > 0): 8 lanes per instance, slot table and factors in the global scratch, ELL width 3, 8..12 rows per lane
#include "solver_team.cuh"

namespace ftsb {
TeamKernelFn team_kernel_rows_w3b(int rows)
{
    switch (rows) {
    case 8:
        return (TeamKernelFn)k_team_tran<8, 0, 3, 8>;
    case 9:
        return (TeamKernelFn)k_team_tran<8, 0, 3, 9>;
    case 10:
        return (TeamKernelFn)k_team_tran<8, 0, 3, 10>;
    case 11:
        return (TeamKernelFn)k_team_tran<8, 0, 3, 11>;
    case 12:
        return (TeamKernelFn)k_team_tran<8, 0, 3, 12>;
    default:
        return nullptr;
    }
}
} // namespace ftsb
