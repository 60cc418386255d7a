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
// Eigen's A.ldlt().solve(b) (ct_icp.cpp:914) is replaced by Gauss-Jordan elimination in natural order on the augmented
// matrix [A | b] IN SHARED MEMORY: at pivot p the 12 x 13 entries are updated by all 32 lanes at once (five entries per
// lane: entry (r, c) becomes A[p][c] / A[p][p] on the pivot row and A[r][c] - (A[r][p] / A[p][p]) A[p][c] elsewhere), so a
// pivot step is one reciprocal, ~15 shared loads, 10 FMAs and 5 stores per lane, and no back-substitution is needed.
// (Round 1 kept row r in 13 registers of lane r and broadcast the pivot row with 13 fp64 shuffles per step: ~3.5k
// instructions on one warp per solve — the serial tail of every ICP iteration; this form is ~0.5k.)
// The system is symmetric positive definite (JTJ/n plus the diagonal regularisers), for which elimination without pivoting
// is backward stable; the result agrees with a pivoted LDL^T to ~1e-13 relative.
static __device__ __forceinline__ void warp_ldlt_solve12(SolveScratch &S, int lane) {
    if (lane < 12) S.A[lane][12] = S.b[lane];   // augmented column
    int er[5], ec[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int e = lane + 32 * k;            // 156 entries: rows of 13
        er[k] = e < 156 ? e / 13 : 0;
        ec[k] = e < 156 ? e % 13 : 0;
    }
    __syncwarp();
    // The reciprocal of pivot p + 1 is started during step p: every lane predicts the entry (p+1, p+1) with the very
    // expression its owner uses below, so the division's latency overlaps the store / barrier / load of the update instead
    // of heading every step's dependency chain.
    auto guarded_rcp = [](double pd) { return (fabs(pd) > 2.2250738585072014e-308) ? 1.0 / pd : 0.0; };   // pseudo-inverse like Eigen's D
    double inv = guarded_rcp(S.A[0][0]);
#pragma unroll 1
    for (int p = 0; p < 12; ++p) {
        double next_inv = 0.0;
        if (p < 11) {
            const double f = S.A[p + 1][p] * inv;
            next_inv = guarded_rcp(S.A[p + 1][p + 1] - f * S.A[p][p + 1]);
        }
        double nv[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double arc = S.A[er[k]][ec[k]], arp = S.A[er[k]][p], apc = S.A[p][ec[k]];
            const double f = arp * inv;
            nv[k] = er[k] == p ? apc * inv : arc - f * apc;
        }
        __syncwarp();   // every lane has read the old matrix
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (lane + 32 * k < 156) S.A[er[k]][ec[k]] = nv[k];
        __syncwarp();
        inv = next_inv;
    }
    if (lane < 12) S.x[lane] = S.A[lane][12];
    __syncwarp();
}


}  // namespace cticp
