// jit.h — per-topology specialised kernels for small circuits (n <= 16), .op and .dc.
//
// The generic kernels walk index maps for every stamp, every matrix entry and every elimination update; at n = 14
// that bookkeeping is ~90 % of the executed instructions (profiles/r01_c2_v2_ncu.md).  Here the host turns the
// compiled topology (slot table, gather maps) and the learned elimination plans (plan.h) into straight-line CUDA C++
// — one thread per instance, slot values / x / rhs in registers, the structural matrix entries in shared memory at
// compile-time offsets — compiles it with NVRTC for sm_100a (--fmad=false, like the rest of the library) and loads the
// cubin into the handle's context.  The arithmetic is the generic kernels': same operations in the same order on the
// structural non-zeros, pivot rule verified at every step, zero / non-finite pivots and unknown pivot sequences go to
// the dense kernel through the redo list.  libnvrtc / libcuda are dlopen'ed: when either is missing the generic
// kernels run instead (still on the GPU; there is no CPU path).
#pragma once
#include <string>

#include "compile.h"
#include "plan.h"

namespace ftsb {

// CUDA C++ source of the specialised kernel `ftsb_jit` for analysis `an` (AN_OP or AN_DC)
// fast_log1p: the damping's ln_1p is fts_log1p.h (table in shared memory behind the matrix entries) like every other
// kernel of the library (1: out-of-line, four unknowns per call; 3: out-of-line, one per call; 2: inlined at every unknown); 0 = CUDA's log1p (A/B measurements
// only: results differ in the last bit)
std::string jit_source(const Compiled &comp, int an, const PlanHost &ph, int fast_log1p = 1);
// dynamic shared memory of the generated kernel
int jit_smem_bytes(const PlanHost &ph);

struct JitKernel {
    void *module = nullptr;   // CUmodule
    void *function = nullptr; // CUfunction
    int smem_bytes = 0;
    int regs = 0;
    std::string log;
};

// Compiles `src` for sm_100a with NVRTC.  Returns false (and a message in `err`) when NVRTC is unavailable or the
// compilation fails.  `cubin` receives the binary; no GPU is needed.
// min_blocks: __launch_bounds__(32, min_blocks) — caps the registers so that that many warps fit per SM
// max_regs > 0: __maxnreg__(max_regs) instead of the min_blocks bound
bool jit_compile(const std::string &src, std::string &cubin, std::string &err, int min_blocks = 1, int max_regs = 0);

// Loads a cubin into the current CUDA context (driver API through dlopen) and looks up `ftsb_jit`.
bool jit_load(const std::string &cubin, int smem_bytes, JitKernel &out, std::string &err);
void jit_unload(JitKernel &k);
// cuLaunchKernel(function, grid, 1, 1, 32, 1, 1, smem, stream, params)
bool jit_launch(const JitKernel &k, int grid, void *stream, void **params, std::string &err);
int jit_occupancy(const JitKernel &k); // resident 32-thread blocks per SM, 0 on error

} // namespace ftsb
