// Instantiations of the resident wavefront solve kernel for VEC = 2 (see ifk_solve_kernel.cuh).
#include "ifk_solve_kernel.cuh"

namespace ifk {

#define IFK_VARIANTS_V2 X(1, 1) X(1, 2) X(1, 3) X(1, 4) X(1, 6) X(1, 8) X(1, 12) X(2, 1) X(2, 2) X(2, 3) X(2, 4) X(2, 6) X(2, 8) X(2, 12) X(3, 1) X(3, 2) X(3, 3) X(3, 4) X(3, 6) X(3, 8) X(3, 12) X(4, 1) X(4, 2) X(4, 3) X(4, 4) X(4, 6) X(4, 8) X(4, 12) X(6, 1) X(6, 2) X(6, 3) X(6, 4) X(6, 6) X(6, 8) X(6, 12) X(8, 1) X(8, 2) X(8, 3) X(8, 4) X(8, 6) X(8, 8) X(12, 1) X(12, 2) X(12, 3) X(12, 4) X(12, 6)

int launch_solve_vec2(int cc, int nv, const SolveParams &p, dim3 grid, int threads, size_t smem, cudaStream_t s)
{
#define X(CC, NV) if (cc == CC && nv == NV) return launch_solve_variant<CC, NV, 2>(p, grid, threads, smem, s);
    IFK_VARIANTS_V2
#undef X
    return IFK_ERR_UNSUPPORTED;
}

int solve_variant_max_threads_vec2(int cc, int nv)
{
#define X(CC, NV) if (cc == CC && nv == NV) return solve_max_threads<CC, NV, 2>();
    IFK_VARIANTS_V2
#undef X
    return 0;
}

}  // namespace ifk
