// kern_team_rows4a.cu — per-team .tran kernels for LINEAR circuits with the rows per lane known at compile time
// (solver_team.cuh, RC > 0): 8 lanes per instance, slot table and factors in the global scratch, ELL width 4, 3..7 rows per lane
#include "solver_team.cuh"

namespace ftsb {
TeamKernelFn team_kernel_rows_w4a(int rows)
{
    switch (rows) {
    case 3:
        return (TeamKernelFn)k_team_tran<8, 0, 4, 3>;
    case 4:
        return (TeamKernelFn)k_team_tran<8, 0, 4, 4>;
    case 5:
        return (TeamKernelFn)k_team_tran<8, 0, 4, 5>;
    case 6:
        return (TeamKernelFn)k_team_tran<8, 0, 4, 6>;
    case 7:
        return (TeamKernelFn)k_team_tran<8, 0, 4, 7>;
    default:
        return nullptr;
    }
}
} // namespace ftsb
