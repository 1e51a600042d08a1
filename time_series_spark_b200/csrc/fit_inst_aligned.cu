// EXPERIMENTAL (off unless PB200_ALIGN=1 at pb200_create): the day-table fit kernel as CTAs of 16 one-warp
// engines that pass one CTA barrier per objective evaluation, so the 16 warps of an SM run the per-evaluation
// serial code -- and the point loop -- in phase and share instruction-cache lines (profiles/r1km_table_variants.md:
// instruction fetch is the top stall of the product kernel).  Same header, compiled with PB200_ENGINES into its
// own namespace; the product translation units (fit_inst.cu) are compiled without the macro and are unchanged.
#define PB200_ENGINES 16
#define PB200_ENGINE_SLICE 13440        // bytes per engine: fit_smem_bytes(32, 1, ppad = 42, 2, PTAB_DAY_MAX) = 13376, rounded
#define pb200 pb200_aligned
#include <string.h>
#include "fit_kernel.cuh"
#undef pb200

namespace pb200_aligned {

template <bool LOGI>
static cudaError_t launch_one(const FitArgs& a, int grid, cudaStream_t st, int* occ) {
    auto kern = fit_kernel<32, LOGI, 0, 3, 4, 3>;
    const size_t smem = (size_t)PB200_ENGINES * PB200_ENGINE_SLICE + 16;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, 32 * PB200_ENGINES, smem);
    kern<<<grid, 32 * PB200_ENGINES, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace pb200_aligned

// args: a pb200::FitArgs (same layout as pb200_aligned::FitArgs -- one header), passed as bytes
extern "C" __attribute__((visibility("hidden"))) int pb200_launch_fit_aligned(int logi, const void* args, int grid,
                                                                              void* stream, int* occ) {
    pb200_aligned::FitArgs a;
    memcpy(&a, args, sizeof a);
    return (int)(logi ? pb200_aligned::launch_one<true>(a, grid, (cudaStream_t)stream, occ)
                      : pb200_aligned::launch_one<false>(a, grid, (cudaStream_t)stream, occ));
}

extern "C" __attribute__((visibility("hidden"))) int pb200_aligned_geometry(int* engines, int* slice_bytes, int* args_bytes) {
    *engines = PB200_ENGINES;
    *slice_bytes = PB200_ENGINE_SLICE;
    *args_bytes = (int)sizeof(pb200_aligned::FitArgs);
    return 0;
}
