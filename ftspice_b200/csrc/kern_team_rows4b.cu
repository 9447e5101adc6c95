// kern_team_rows4b.cu — per-team .tran kernels for LINEAR circuits with the rows per lane known at compile time
// (solver_team.cuh, RC > 0): 8 lanes per instance, slot table and factors in the global scratch, ELL width 4, 8..12 rows per lane
#include "solver_team.cuh"

namespace ftsb {
TeamKernelFn team_kernel_rows_w4b(int rows)
{
    switch (rows) {
    case 8:
        return (TeamKernelFn)k_team_tran<8, 0, 4, 8>;
    case 9:
        return (TeamKernelFn)k_team_tran<8, 0, 4, 9>;
    case 10:
        return (TeamKernelFn)k_team_tran<8, 0, 4, 10>;
    case 11:
        return (TeamKernelFn)k_team_tran<8, 0, 4, 11>;
    case 12:
        return (TeamKernelFn)k_team_tran<8, 0, 4, 12>;
    default:
        return nullptr;
    }
}
} // namespace ftsb
