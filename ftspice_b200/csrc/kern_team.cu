// kern_team.cu — instantiations of the per-team .tran kernels (solver_team.cuh)
#include "solver_team.cuh"

namespace ftsb {
template <int PLACE, int W>
static TeamKernelFn pick(int g)
{
    switch (g) {
    case 4:
        return (TeamKernelFn)k_team_tran<4, PLACE, W>;
    case 8:
        return (TeamKernelFn)k_team_tran<8, PLACE, W>;
    case 16:
        return (TeamKernelFn)k_team_tran<16, PLACE, W>;
    case 32:
        return (TeamKernelFn)k_team_tran<32, PLACE, W>;
    default:
        return nullptr;
    }
}
// place: bit 0 = slot table in shared memory, bit 1 = factors in shared memory; ell_w: ELL slots per row of the residual
TeamKernelFn team_kernel(int g, int place, int ell_w)
{
    if (ell_w != 3 && ell_w != 4) return nullptr;
    switch (place) {
    case 0:
        return ell_w == 3 ? pick<0, 3>(g) : pick<0, 4>(g);
    case 1:
        return ell_w == 3 ? pick<1, 3>(g) : pick<1, 4>(g);
    case 3:
        return ell_w == 3 ? pick<3, 3>(g) : pick<3, 4>(g);
    case 6: // large nonlinear circuits: factors in shared memory, slot table and the entries of A in the global region
        return ell_w == 3 ? pick<6, 3>(g) : pick<6, 4>(g);
    default:
        return nullptr;
    }
}
TeamKernelFn team_kernel_rows_w3a(int rows); // kern_team_rows3a.cu ...
TeamKernelFn team_kernel_rows_w3b(int rows);
TeamKernelFn team_kernel_rows_w4a(int rows);
TeamKernelFn team_kernel_rows_w4b(int rows);
// linear circuits with a chain plan, 8 lanes per instance, slot table and factors in the global scratch (place 0),
// 3 <= rows per lane <= 12 (17 <= n <= 96)
TeamKernelFn team_kernel_rows(int g, int place, int ell_w, int rows)
{
    if (g != 8 || place != 0 || rows < 3 || rows > 12) return nullptr;
    if (ell_w == 3) return rows <= 7 ? team_kernel_rows_w3a(rows) : team_kernel_rows_w3b(rows);
    if (ell_w == 4) return rows <= 7 ? team_kernel_rows_w4a(rows) : team_kernel_rows_w4b(rows);
    return nullptr;
}
} // namespace ftsb

// ---------------------------------------------------------------------------------------------------------
// diagnostics (not part of include/ftsb.h): the back substitution's division by a pivot with a known reciprocal
// (TeamOps::div_known_rcp) and the plain division, element by element, for the bit-equality test
// ---------------------------------------------------------------------------------------------------------
namespace {
__global__ void k_debug_div(const double *y, const double *d, double *fast, double *ref, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double al = 1.0 / d[i];
        fast[i] = ftsb::TeamOps<8>::div_known_rcp(y[i], d[i], al);
        ref[i] = y[i] / d[i];
    }
}
} // namespace

extern "C" int ftsb_debug_div(int device, long long n, const double *y, const double *d, double *fast, double *ref)
{
    double *b = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&b, (size_t)n * 32) != cudaSuccess) return -2;
    cudaMemcpy(b, y, (size_t)n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(b + n, d, (size_t)n * 8, cudaMemcpyHostToDevice);
    k_debug_div<<<592, 256>>>(b, b + n, b + 2 * n, b + 3 * n, n);
    cudaMemcpy(fast, b + 2 * n, (size_t)n * 8, cudaMemcpyDeviceToHost);
    const cudaError_t e = cudaMemcpy(ref, b + 3 * n, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaFree(b);
    return e == cudaSuccess ? 0 : -2;
}
