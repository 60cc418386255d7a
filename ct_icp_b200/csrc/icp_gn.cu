// icp_gn.cu — Gauss-Newton CT-ICP iteration as two kernels per iteration:
//   k_gn_gather : one warp per keypoint — world point from the continuous-time pose pair, 27-voxel stencil gather,
//                 kNN, covariance + eigen, point-to-plane residual and 12-vector Jacobian row, accumulation of
//                 JTJ (78 unique entries) / JTr (12) in registers, block reduction in shared memory.
//   k_gn_solve  : one block — deterministic reduction of the per-block partials, 1/n normalisation, motion-model
//                 regularisers, pivoted LDL^T solve of the 12x12 system, Euler-ZYX pose update, stop test.
// Reference: DoRegisterGaussNewton, src/ct_icp/ct_icp.cpp:709-996 (serial per-keypoint loop :753-857).
#include <cstdio>
#include <cstdlib>

#include <cooperative_groups.h>

#include "gather.cuh"
#include "icp.h"
#include "peer_exchange.cuh"
#include "small_solve.cuh"

namespace cticp {
namespace cg = cooperative_groups;

#define CT_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                            std::to_string(__LINE__));                                                   \
    } while (0)

// Warps per CTA of the gather kernels (one keypoint per warp at a time). Fewer, fatter CTAs make the grid-wide
// barriers and the reduction of the per-CTA partial sums cheaper, and more warps share one SM's instruction cache:
// measured on config 2 (K ~ 1.2k), GN loop per frame: 4 warps 0.27 ms, 8 warps 0.193 ms, 16 warps 0.181 ms
// (16 warps x 128 registers = the whole register file: one CTA per SM, ~77 CTAs).
#ifndef CTICP_GATHER_WARPS
#define CTICP_GATHER_WARPS 16
#endif
constexpr int kGatherWarps = CTICP_GATHER_WARPS;


#ifdef CTICP_DEBUG_TIMERS
// SM-local cycle counter (%globaltimer proved far too slow to read: it tripled the kernel time). Only differences
// taken on the same SM are meaningful: dbg_t[1..3] are all stamped by the solver CTA.
__device__ __forceinline__ unsigned long long global_timer_ns() { return (unsigned long long) clock64(); }
#define CT_STAMP(expr) expr
#else
#define CT_STAMP(expr)
#endif

struct GatherLaunch {
    GatherConfig G;
    GnParams P;
};

__device__ __forceinline__ M3 euler_from_sincos(double sa, double ca, double sb, double cb, double sg, double cg) {
    M3 R;   // ct_icp.cpp:916-932
    R.m[0][0] = cg * cb; R.m[0][1] = -sg * ca + cg * sb * sa; R.m[0][2] = sg * sa + cg * sb * ca;
    R.m[1][0] = sg * cb; R.m[1][1] = cg * ca + sg * sb * sa;  R.m[1][2] = -cg * sa + sg * sb * ca;
    R.m[2][0] = -sb;     R.m[2][1] = cb * sa;                 R.m[2][2] = cb * ca;
    return R;
}

// One warp: accumulator (96 doubles in `acc`) → normal equations → GN step → pose update (ct_icp.cpp:860-980).
// mode 0: full step. mode 1: only emit the linear system into sys_out (debug tap).
__device__ void warp_gn_solve(const double *acc, SolveScratch &S, IcpState *st, const GnParams &P, int mode,
                              double *sys_out, int lane) {
    const int n_used = (int) (acc[kAccUsed] + 0.5);
    if (lane == 0) {
        st->n_used = n_used;
        st->n_keypoints = (int) (acc[kAccKeypoints] + 0.5);
        st->stat_keypoint_iters += (unsigned long long) (acc[kAccKeypoints] + 0.5);
        st->stat_stencil_points += (unsigned long long) (acc[kAccStencil] + 0.5);
        if (sys_out) sys_out[156] = (double) n_used;
    }
    if (n_used < 100) {   // ct_icp.cpp:860-871
        if (lane == 0) {
            st->failed = 1;
            st->done = 1;
        }
        return;
    }
    {
        const double inv = 1.0 / (double) n_used;   // :877-882
        for (int e = lane; e < 78; e += 32) {
            const int i = c_pair_i[e], j = c_pair_j[e];
            const double v = acc[e] * inv;
            S.A[i][j] = v;
            S.A[j][i] = v;
        }
        if (lane < 12) S.b[lane] = acc[78 + lane] * inv;
    }
    __syncwarp();
    if (st->has_motion_model && lane < 3) {   // :885-910
        const int d = lane;
        const double ac = st->beta_location, ae = st->beta_cv;
        const double diff_traj = st->tb[d] - st->te[d];   // the frame's own begin - end (sic, :892)
        S.A[3 + d][3 + d] += ac;
        S.b[3 + d] -= ac * diff_traj;
        const double diff_ego = st->te[d] - st->tb[d] - st->prev_te[d] + st->prev_tb[d];
        S.A[9 + d][9 + d] += ae;
        S.b[9 + d] -= ae * diff_ego;
    }
    __syncwarp();
    if (sys_out) {
        for (int e = lane; e < 144; e += 32) sys_out[e] = S.A[e / 12][e % 12];
        if (lane < 12) sys_out[144 + lane] = S.b[lane];
    }
    if (mode == 1) return;

    warp_ldlt_solve12(S, lane);   // :914

    if (lane < 6) {   // angles x[0..2] (begin) and x[6..8] (end): sin / cos evaluated by six lanes at once
        const double ang = S.x[lane < 3 ? lane : lane + 3];
        S.sn[lane] = sin(ang);
        S.cs[lane] = cos(ang);
    }
    __syncwarp();
    if (lane < 2) {   // lane 0: begin pose, lane 1: end pose (:916-962)
        const int o = 3 * lane;
        const M3 R = euler_from_sincos(S.sn[o], S.cs[o], S.sn[o + 1], S.cs[o + 1], S.sn[o + 2], S.cs[o + 2]);
        double *qp = lane == 0 ? st->qb : st->qe;
        double *tp = lane == 0 ? st->tb : st->te;
        const Q4 q = qnormalized(qfromR(mmul(R, qtoR(Q4{qp[0], qp[1], qp[2], qp[3]}))));
        qp[0] = q.x; qp[1] = q.y; qp[2] = q.z; qp[3] = q.w;
        const int xo = lane == 0 ? 3 : 9;
        for (int d = 0; d < 3; ++d) tp[d] += S.x[xo + d];
    }
    __syncwarp();
    if (lane == 0) {
        double nrm = 0;
        for (int i = 0; i < 12; ++i) nrm += S.x[i] * S.x[i];
        nrm = sqrt(nrm);
        st->x_norm = nrm;
        st->iter += 1;
        if (nrm < P.threshold_norm) st->done = 1;   // :978
        const SlerpConsts sc = slerp_consts(Q4{st->qb[0], st->qb[1], st->qb[2], st->qb[3]},
                                            Q4{st->qe[0], st->qe[1], st->qe[2], st->qe[3]});
        st->slerp_theta = sc.theta;
        st->slerp_inv_sin = sc.inv_sin;
        st->slerp_linear = sc.linear;
        st->slerp_negate = sc.negate;
    }
}

// mode 0: gather + (last CTA) reduce + solve + pose update          [single GPU: ONE launch per ICP iteration]
// mode 1: gather + (last CTA) reduce + emit linear system to sys_out [debug tap]
// mode 2: gather + (last CTA) reduce into acc_out                    [multi-GPU: all-reduce then k_gn_solve_acc]
__global__ void __launch_bounds__(kGatherWarps * 32)
k_gn_iterate(GatherLaunch cfg, const float4 *__restrict__ keypoints, const int *__restrict__ d_num_keypoints,
             IcpState *st, double *__restrict__ partials, unsigned int *ticket, int mode, double *acc_out,
             double *sys_out) {
    __shared__ KnnStage s_stage[kGatherWarps][64];
    __shared__ double s_u[kGatherWarps][16];
    __shared__ double s_acc[kGatherWarps][kAcc];
    __shared__ int s_stencil[kMaxStencil];
    __shared__ SolveScratch s_solve;
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const GatherConfig &G = cfg.G;
    const GnParams &P = cfg.P;

    double acc0 = 0, acc1 = 0, acc2 = 0;              // entries lane, lane+32, lane+64 of [A upper | b]
    double n_used = 0, sum_sq = 0, n_stencil = 0, n_kp = 0, n_valid = 0;
    const bool active = (mode == 1) || !st->done;
    CT_STAMP(if (blockIdx.x == 0 && threadIdx.x == 0) st->dbg_t[0] = global_timer_ns();)

    if (active) {
        const int *stencil = stencil_table_fill(s_stencil, G.r);
        __syncthreads();
        const Q4 qb{st->qb[0], st->qb[1], st->qb[2], st->qb[3]}, qe{st->qe[0], st->qe[1], st->qe[2], st->qe[3]};
        const V3 tb{st->tb[0], st->tb[1], st->tb[2]}, te{st->te[0], st->te[1], st->te[2]};
        const SlerpConsts sc{st->slerp_theta, st->slerp_inv_sin, st->slerp_linear, st->slerp_negate};
        const int K = *d_num_keypoints;
        const int lo = (int) ((long long) K * P.shard_rank / P.shard_world);
        const int hi = (int) ((long long) K * (P.shard_rank + 1) / P.shard_world);
        const int warps_total = gridDim.x * kGatherWarps;
        const int i0 = lane, i1 = lane + 32, i2 = lane + 64;
        const int pi0 = c_pair_i[i0], pj0 = c_pair_j[i0], pi1 = c_pair_i[i1], pj1 = c_pair_j[i1];
        const int pi2 = i2 < kAccUsed ? c_pair_i[i2] : 0, pj2 = i2 < kAccUsed ? c_pair_j[i2] : 0;

        for (int kp = lo + blockIdx.x * kGatherWarps + w; kp < hi; kp += warps_total) {
            const float4 kraw = __ldg(keypoints + kp);   // raw xyz (sensor frame) + alpha timestamp
            const V3 raw{(double) kraw.x, (double) kraw.y, (double) kraw.z};
            const double alpha = (double) kraw.w;
            // world_kpts[i] = InterpolatePose(begin, end, t_i) * raw_i  (ct_icp.cpp:964-966, types.h:361-366)
            const V3 p = ct_transform_c(qb, tb, qe, te, alpha, raw, sc);
            const QueryCtx ctx = make_query(p, G.L.res, lane);

            KnnEntry best;
            unsigned spts = 0;
            const int n = warp_gather_knn(G, stencil, ctx, lane, s_stage[w], best, spts);
            n_kp += 1;
            n_stencil += (double) spts;
            if (n < P.kmin || n < 5) continue;   // ct_icp.cpp:769 ; neighborhood.h:227
            n_valid += 1;

            const NeighborhoodDesc nd = warp_describe(G, stencil, ctx, best, n, lane);
            V3 normal = nd.normal;
            // orient towards the sensor position at frame begin (ct_icp.cpp:782-784)
            if (dot(normal, tb - p) < 0) normal = -1.0 * normal;
            const double weight = nd.a2D * nd.a2D;                     // :787-788
            // p - closest_point, closest_point = points[0] = farthest kept (:791)
            const V3 diff{-nd.far_rel.x, -nd.far_rel.y, -nd.far_rel.z};
            const double dist_to_plane = dot(normal, diff);
            if (!(fabs(dist_to_plane) < P.max_dist_to_plane)) continue;   // :803
            const V3 nw = weight * normal;
            const double scalar = dot(nw, diff);

            if (lane == 0) {
                const V3 ob = qrot(qb, raw), oe = qrot(qe, raw);         // :813-816
                const double am = 1.0 - alpha, a = alpha;
                const V3 cb = cross(ob, nw), ce = cross(oe, nw);
                double *u = s_u[w];
                u[0] = am * cb.x; u[1] = am * cb.y; u[2] = am * cb.z;
                u[3] = am * nw.x; u[4] = am * nw.y; u[5] = am * nw.z;
                u[6] = a * ce.x;  u[7] = a * ce.y;  u[8] = a * ce.z;
                u[9] = a * nw.x;  u[10] = a * nw.y; u[11] = a * nw.z;
                u[12] = -scalar;   // b -= u * scalar (:849)
            }
            __syncwarp();
            {
                const double *u = s_u[w];
                acc0 += u[pi0] * u[pj0];
                acc1 += u[pi1] * u[pj1];
                if (i2 < kAccUsed) acc2 += u[pi2] * u[pj2];
            }
            __syncwarp();
            n_used += 1;
            sum_sq += scalar * scalar;
        }
    }

    // block reduction (fixed order → run-to-run deterministic)
    s_acc[w][lane] = acc0;
    s_acc[w][lane + 32] = acc1;
    s_acc[w][lane + 64] = acc2;
    if (lane == 0) {
        s_acc[w][kAccUsed] = n_used;
        s_acc[w][kAccSumSq] = sum_sq;
        s_acc[w][kAccStencil] = n_stencil;
        s_acc[w][kAccKeypoints] = n_kp;
        s_acc[w][kAccValidNb] = n_valid;
        s_acc[w][95] = 0;
    }
    __syncthreads();
    if (threadIdx.x < kAcc) {
        double s = 0;
#pragma unroll
        for (int ww = 0; ww < kGatherWarps; ++ww) s += s_acc[ww][threadIdx.x];
        partials[(size_t) blockIdx.x * kAcc + threadIdx.x] = s;
    }
    // ---- last CTA to finish reduces the partials (fixed order) and takes the Gauss-Newton step -----------------
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        *ticket = 0;
        CT_STAMP(st->dbg_t[1] = global_timer_ns();)
    }
    if (!active) return;
    double *acc = &s_acc[0][0];
    {
        // deterministic reduction: warp g sums rows b = g (mod 4) of three columns per lane; fixed-order combine
        double a0 = 0, a1 = 0, a2 = 0;
        const int nb = gridDim.x;
        for (int b = w; b < nb; b += kGatherWarps) {
            const double *row = partials + (size_t) b * kAcc;
            a0 += __ldcg(row + lane);
            a1 += __ldcg(row + lane + 32);
            a2 += __ldcg(row + lane + 64);
        }
        __syncthreads();
        s_acc[w][lane] = a0;
        s_acc[w][lane + 32] = a1;
        s_acc[w][lane + 64] = a2;
        __syncthreads();
        double sum = 0;
        if (threadIdx.x < kAcc) {
#pragma unroll
            for (int ww = 0; ww < kGatherWarps; ++ww) sum += s_acc[ww][threadIdx.x];
        }
        __syncthreads();
        if (threadIdx.x < kAcc) {
            acc[threadIdx.x] = sum;
            if (mode == 2) acc_out[threadIdx.x] = sum;
        }
    }
    __syncthreads();
    CT_STAMP(if (threadIdx.x == 0) st->dbg_t[2] = global_timer_ns();)
    if (mode == 2 || w != 0) return;
    warp_gn_solve(acc, s_solve, st, P, mode, sys_out, lane);
    CT_STAMP(if (lane == 0) st->dbg_t[3] = global_timer_ns();)
}

// ---- persistent variant: the WHOLE Gauss-Newton loop in one cooperative launch --------------------------------
// CTA 0 is the solver CTA (deterministic reduction of the partials + 12x12 solve + pose update, by the same warp on
// the same SM every iteration, so its instructions stay in that SM's instruction cache: executed cold, the ~1.5k
// instructions of the serial tail cost ~45 us per iteration, warm ~8 us); CTAs 1..G gather. Two grid-wide barriers
// per iteration replace two kernel launches. While the gather CTAs work on iteration 0, the solver warp runs the
// solve once on a dummy system to pull its code into the instruction cache.
__device__ __forceinline__ void load_state_volatile(const IcpState *st, Q4 &qb, V3 &tb, Q4 &qe, V3 &te, SlerpConsts &sc) {
    qb = Q4{__ldcg(&st->qb[0]), __ldcg(&st->qb[1]), __ldcg(&st->qb[2]), __ldcg(&st->qb[3])};
    qe = Q4{__ldcg(&st->qe[0]), __ldcg(&st->qe[1]), __ldcg(&st->qe[2]), __ldcg(&st->qe[3])};
    tb = V3{__ldcg(&st->tb[0]), __ldcg(&st->tb[1]), __ldcg(&st->tb[2])};
    te = V3{__ldcg(&st->te[0]), __ldcg(&st->te[1]), __ldcg(&st->te[2])};
    sc = SlerpConsts{__ldcg(&st->slerp_theta), __ldcg(&st->slerp_inv_sin), __ldcg(&st->slerp_linear), __ldcg(&st->slerp_negate)};
}

//
// kPeers (multi-GPU, keypoints sharded): between its reduction and its solve the solver CTA exchanges the accumulator
// with the other ranks' solver CTAs through NVLink peer memory (peer_exchange.cuh) — the all-reduce of SURVEY §8e
// happens INSIDE the loop, so the sharded loop is still one launch and costs one NVLink round trip per iteration.
#ifdef CTICP_HANDOFF
// Experiment (-DCTICP_HANDOFF, to be measured): the two grid-wide barriers of an iteration become two one-directional
// hand-offs — gather CTAs → solver CTA through an arrive counter (only the solver polls it), solver CTA → gather CTAs
// through an epoch word (one thread per gather CTA polls it, the rest of the CTA waits at a hardware barrier).
__device__ __forceinline__ unsigned int handoff_load(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void handoff_store(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded poll (one thread per CTA): false after ~1 s, so a protocol error ends the kernel instead of hanging the GPU
__device__ __forceinline__ bool handoff_wait(const unsigned int *p, unsigned int target) {
    const long long t0 = clock64();
    while (handoff_load(p) < target)
        if (clock64() - t0 > 2000000000LL) return false;
    return true;
}
#endif

template <bool kPeers>
__global__ void __launch_bounds__(kGatherWarps * 32)
k_gn_persistent(GatherLaunch cfg, const float4 *__restrict__ keypoints, const int *__restrict__ d_num_keypoints,
                IcpState *st, double *__restrict__ partials, int num_iters, PeerLinks links) {
    // Measured alternatives (config 2, 5 iterations, ICP ms): this design 0.180; arrive-counter / epoch-word hand-off
    // instead of grid.sync() 0.383; ONE barrier per iteration with every CTA redundantly reducing + solving 0.343 —
    // although its empty loop is 2.4x cheaper (0.029 vs 0.070): when every SM alternates between the gather code and
    // the solve code each iteration, the instruction working set no longer fits the SM's instruction cache, while a
    // dedicated solver CTA keeps the ~1.5k-instruction serial tail hot on one SM (6.6 us vs ~40 us per iteration).
    cg::grid_group grid = cg::this_grid();
    __shared__ KnnStage s_stage[kGatherWarps][64];
    __shared__ double s_u[kGatherWarps][16];
    __shared__ double s_acc[kGatherWarps][kAcc];
    __shared__ int s_stencil[kMaxStencil];
    __shared__ SolveScratch s_solve;
    __shared__ IcpState s_dummy;
    __shared__ int s_peer_ok;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const GatherConfig &G = cfg.G;
    const GnParams &P = cfg.P;
    const bool solver_cta = blockIdx.x == 0;
    unsigned int peer_seq = 0;   // sequence number of the last exchange (solver CTA only)
    if (kPeers && solver_cta) peer_seq = *links.seq;
    const int gather_ctas = gridDim.x - 1;
    const int *stencil = stencil_table_fill(s_stencil, G.r);
    __syncthreads();

    if (solver_cta && w == 0 && !(P.debug_flags & 4)) {
        // instruction-cache warm-up of the serial tail on a dummy well-posed system (results discarded)
        CT_STAMP(if (lane == 0) st->dbg_t[0] = global_timer_ns();)
        for (int i = lane; i < kAcc; i += 32) s_acc[1][i] = 0.0;
        __syncwarp();
        for (int e = lane; e < 78; e += 32)
            if (c_pair_i[e] == c_pair_j[e]) s_acc[1][e] = 200.0 * (1.0 + c_pair_i[e]);
        if (lane < 12) s_acc[1][78 + lane] = 1e-3 * (lane + 1);
        if (lane == 0) {
            s_acc[1][kAccUsed] = 200.0;
            s_dummy = *st;
        }
        __syncwarp();
        warp_gn_solve(s_acc[1], s_solve, &s_dummy, P, 0, nullptr, lane);
        __syncwarp();
    }

    for (int it = 0; it < num_iters; ++it) {
#ifdef CTICP_HANDOFF
        if (!solver_cta && it > 0) {   // wait until the solver CTA has published the pose of iteration it - 1
            if (threadIdx.x == 0) s_peer_ok = handoff_wait(&st->handoff_epoch, (unsigned int) it);
            __syncthreads();
            if (!s_peer_ok) break;   // timed out (the solver CTA flags the failure)
        }
#endif
        if (__ldcg(&st->done)) break;   // uniform: written before the previous grid barrier
        if (!solver_cta) {
            double acc0 = 0, acc1 = 0, acc2 = 0;
            double n_used = 0, sum_sq = 0, n_stencil = 0, n_kp = 0, n_valid = 0;
            Q4 qb, qe;
            V3 tb, te;
            SlerpConsts sc;
            load_state_volatile(st, qb, tb, qe, te, sc);
            const int K = *d_num_keypoints;
            const int lo = (int) ((long long) K * P.shard_rank / P.shard_world);
            const int hi = (int) ((long long) K * (P.shard_rank + 1) / P.shard_world);
            const int warps_total = gather_ctas * kGatherWarps;
            const int i0 = lane, i1 = lane + 32, i2 = lane + 64;
            const int pi0 = c_pair_i[i0], pj0 = c_pair_j[i0], pi1 = c_pair_i[i1], pj1 = c_pair_j[i1];
            const int pi2 = i2 < kAccUsed ? c_pair_i[i2] : 0, pj2 = i2 < kAccUsed ? c_pair_j[i2] : 0;
            for (int kp = lo + (blockIdx.x - 1) * kGatherWarps + w; kp < hi; kp += warps_total) {
                if (P.debug_flags & 2) break;
                const float4 kraw = __ldg(keypoints + kp);
                const V3 raw{(double) kraw.x, (double) kraw.y, (double) kraw.z};
                const double alpha = (double) kraw.w;
                const V3 p = ct_transform_c(qb, tb, qe, te, alpha, raw, sc);
                const QueryCtx ctx = make_query(p, G.L.res, lane);
                KnnEntry best;
                unsigned spts = 0;
                const int n = warp_gather_knn(G, stencil, ctx, lane, s_stage[w], best, spts);
                n_kp += 1;
                n_stencil += (double) spts;
                if (n < P.kmin || n < 5) continue;
                n_valid += 1;
                const NeighborhoodDesc nd = warp_describe(G, stencil, ctx, best, n, lane);
                V3 normal = nd.normal;
                if (dot(normal, tb - p) < 0) normal = -1.0 * normal;
                const double weight = nd.a2D * nd.a2D;
                const V3 diff{-nd.far_rel.x, -nd.far_rel.y, -nd.far_rel.z};
                const double dist_to_plane = dot(normal, diff);
                if (!(fabs(dist_to_plane) < P.max_dist_to_plane)) continue;
                const V3 nw = weight * normal;
                const double scalar = dot(nw, diff);
                if (lane == 0) {
                    const V3 ob = qrot(qb, raw), oe = qrot(qe, raw);
                    const double am = 1.0 - alpha, a = alpha;
                    const V3 cb = cross(ob, nw), ce = cross(oe, nw);
                    double *u = s_u[w];
                    u[0] = am * cb.x; u[1] = am * cb.y; u[2] = am * cb.z;
                    u[3] = am * nw.x; u[4] = am * nw.y; u[5] = am * nw.z;
                    u[6] = a * ce.x;  u[7] = a * ce.y;  u[8] = a * ce.z;
                    u[9] = a * nw.x;  u[10] = a * nw.y; u[11] = a * nw.z;
                    u[12] = -scalar;
                }
                __syncwarp();
                {
                    const double *u = s_u[w];
                    acc0 += u[pi0] * u[pj0];
                    acc1 += u[pi1] * u[pj1];
                    if (i2 < kAccUsed) acc2 += u[pi2] * u[pj2];
                }
                __syncwarp();
                n_used += 1;
                sum_sq += scalar * scalar;
            }
            s_acc[w][lane] = acc0;
            s_acc[w][lane + 32] = acc1;
            s_acc[w][lane + 64] = acc2;
            if (lane == 0) {
                s_acc[w][kAccUsed] = n_used;
                s_acc[w][kAccSumSq] = sum_sq;
                s_acc[w][kAccStencil] = n_stencil;
                s_acc[w][kAccKeypoints] = n_kp;
                s_acc[w][kAccValidNb] = n_valid;
                s_acc[w][95] = 0;
            }
            __syncthreads();
            if (threadIdx.x < kAcc) {
                double s = 0;
#pragma unroll
                for (int ww = 0; ww < kGatherWarps; ++ww) s += s_acc[ww][threadIdx.x];
                __stcg(&partials[(size_t) (blockIdx.x - 1) * kAcc + threadIdx.x], s);
            }
        }
#ifdef CTICP_HANDOFF
        if (!solver_cta) {   // deliver: row written → fence → count this CTA in
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(&st->handoff_arrive, 1u);
        } else {             // collect: all gather CTAs of this iteration have delivered
            if (threadIdx.x == 0) {
                s_peer_ok = handoff_wait(&st->handoff_arrive, (unsigned int) gather_ctas * (unsigned int) (it + 1));
                if (!s_peer_ok) {
                    st->failed = 3;
                    st->done = 1;
                }
            }
            __syncthreads();
            if (!s_peer_ok) break;
        }
#else
        grid.sync();
#endif
        if (solver_cta) {
            CT_STAMP(if (threadIdx.x == 0) st->dbg_t[1] = global_timer_ns();)
            // deterministic reduction: warp g sums the rows b = g (mod kGatherWarps) of three columns per lane, then
            // the per-warp sums are combined in fixed order
            double a0 = 0, a1 = 0, a2 = 0;
#ifdef CTICP_REDUCE_MLP
            // Experiment (-DCTICP_REDUCE_MLP, to be measured): all of a warp's rows are requested before the first is
            // added (one L2 round trip instead of one per four rows); the additions keep their order
            constexpr int kRowsInFlight = 8;
            for (int b0 = w; b0 < gather_ctas; b0 += kGatherWarps * kRowsInFlight) {
                double v0[kRowsInFlight], v1[kRowsInFlight], v2[kRowsInFlight];
#pragma unroll
                for (int u = 0; u < kRowsInFlight; ++u) {
                    const int b = b0 + u * kGatherWarps;
                    if (b < gather_ctas) {
                        const double *row = partials + (size_t) b * kAcc;
                        v0[u] = __ldcg(row + lane);
                        v1[u] = __ldcg(row + lane + 32);
                        v2[u] = __ldcg(row + lane + 64);
                    }
                }
#pragma unroll
                for (int u = 0; u < kRowsInFlight; ++u)
                    if (b0 + u * kGatherWarps < gather_ctas) {
                        a0 += v0[u];
                        a1 += v1[u];
                        a2 += v2[u];
                    }
            }
#else
            for (int b = w; b < gather_ctas; b += kGatherWarps) {
                const double *row = partials + (size_t) b * kAcc;
                a0 += __ldcg(row + lane);
                a1 += __ldcg(row + lane + 32);
                a2 += __ldcg(row + lane + 64);
            }
#endif
            s_acc[w][lane] = a0;
            s_acc[w][lane + 32] = a1;
            s_acc[w][lane + 64] = a2;
            __syncthreads();
            if (threadIdx.x < kAcc) {
                double s = 0;
#pragma unroll
                for (int ww = 0; ww < kGatherWarps; ++ww) s += s_acc[ww][threadIdx.x];
                s_u[0][0] = 0;   // (keeps s_u referenced in the solver CTA)
                s_acc[0][threadIdx.x] = s;
            }
            __syncthreads();
            bool peers_ok = true;
            if (kPeers) {
                // Σ over ranks, in rank order (bit-identical on every rank); the stencil staging area of this CTA is
                // unused by the solver CTA and serves as scratch
                static_assert(sizeof(KnnStage) * 64 * kGatherWarps >= sizeof(unsigned int) * kMaxPeers * kPeerWords, "scratch");
                peers_ok = peer_allreduce(links, ++peer_seq, s_acc[0], reinterpret_cast<unsigned int *>(&s_stage[0][0]), &s_peer_ok);
            }
            CT_STAMP(if (threadIdx.x == 0) st->dbg_t[2] = global_timer_ns();)
            if (w == 0) {
                if (!peers_ok) {
                    if (lane == 0) {   // a peer never answered: give up instead of hanging the device
                        st->failed = 3;
                        st->done = 1;
                    }
                } else if (P.debug_flags & 1) {
                    if (lane == 0) st->iter += 1;
                } else
                    warp_gn_solve(s_acc[0], s_solve, st, P, 0, nullptr, lane);
                CT_STAMP(if (lane == 0) st->dbg_t[3] = global_timer_ns();)
            }
            __threadfence();
        }
#ifdef CTICP_HANDOFF
        if (solver_cta) {   // publish: the pose update (fenced above by every thread of this CTA) is visible → epoch
            __syncthreads();
            if (threadIdx.x == 0) handoff_store(&st->handoff_epoch, (unsigned int) (it + 1));
        }
#else
        grid.sync();
#endif
    }
    if (kPeers && solver_cta && threadIdx.x == 0) *links.seq = peer_seq;
}

// Stand-alone exchange for the launch-per-step paths (solvers CERES / ROBUST: one per LM evaluation; GN with
// CTICP_PERSISTENT=0): acc ← Σ over ranks of acc, in place, one CTA.
__global__ void __launch_bounds__(256) k_peer_allreduce(PeerLinks links, double *__restrict__ acc, IcpState *st) {
    __shared__ double s_acc[kAcc];
    __shared__ unsigned int s_half[kMaxPeers * kPeerWords];
    __shared__ int s_ok;
    if (threadIdx.x < kAcc) s_acc[threadIdx.x] = acc[threadIdx.x];
    const unsigned int seq = *links.seq + 1;
    const bool ok = peer_allreduce(links, seq, s_acc, s_half, &s_ok);
    if (threadIdx.x < kAcc) acc[threadIdx.x] = s_acc[threadIdx.x];
    if (threadIdx.x == 0) {
        *links.seq = seq;
        if (!ok && st) {
            st->failed = 3;
            st->done = 1;
        }
    }
}

// multi-GPU tail: the all-reduced accumulator → GN step (one warp)
__global__ void k_gn_solve_acc(const double *__restrict__ acc_in, IcpState *st, GnParams P) {
    __shared__ SolveScratch s_solve;
    __shared__ double s_acc[kAcc];
    if (st->done) return;
    for (int i = threadIdx.x; i < kAcc; i += 32) s_acc[i] = acc_in[i];
    __syncwarp();
    warp_gn_solve(s_acc, s_solve, st, P, 0, nullptr, threadIdx.x);
}

// Neighbor lists for arbitrary queries (parity tests of the map search; ComputeNeighborhoods, map.h:532-540)
__global__ void __launch_bounds__(kGatherWarps * 32)
k_neighborhoods(GatherConfig G, const double *__restrict__ queries, int n, double *__restrict__ out_points,
                int *__restrict__ out_counts) {
    __shared__ KnnStage s_stage[kGatherWarps][64];
    __shared__ int s_stencil[kMaxStencil];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int *stencil = stencil_table_fill(s_stencil, G.r);
    __syncthreads();
    for (int i = blockIdx.x * kGatherWarps + w; i < n; i += gridDim.x * kGatherWarps) {
        const V3 q{queries[3 * i], queries[3 * i + 1], queries[3 * i + 2]};
        const QueryCtx ctx = make_query(q, G.L.res, lane);
        KnnEntry best;
        unsigned spts;
        const int cnt = warp_gather_knn(G, stencil, ctx, lane, s_stage[w], best, spts);
        if (lane == 0) out_counts[i] = cnt;
        if (lane < cnt) {
            const V3 rel = knn_rel_position(G, stencil, ctx, best);
            double *o = out_points + ((size_t) i * G.kmax + (cnt - 1 - lane)) * 3;   // farthest first
            o[0] = q.x + rel.x; o[1] = q.y + rel.y; o[2] = q.z + rel.z;
        }
        __syncwarp();
    }
}

// ComputeNeighborhoods(queries, radiuses, max_num_neighbors, true, sensor_location), map.h:434-447: per-query radius →
// per-query level and stencil; optional normal filter.
struct RadiusSearchLevels {
    int num_levels, filter;
    V3 sensor;
    MapLevel levels[CTICP_MAX_RESOLUTIONS];
};
__global__ void __launch_bounds__(kGatherWarps * 32)
k_radius_search(const RadiusSearchLevels *__restrict__ R, int kmax, const double *__restrict__ queries,
                const double *__restrict__ radiuses, int n, double *__restrict__ out_points, int *__restrict__ out_counts) {
    __shared__ KnnStage s_stage[kGatherWarps][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = blockIdx.x * kGatherWarps + w; i < n; i += gridDim.x * kGatherWarps) {
        const V3 q{queries[3 * i], queries[3 * i + 1], queries[3 * i + 2]};
        const double radius = radiuses[i];
        int it = 0;
        while (it < R->num_levels && R->levels[it].res <= radius) ++it;
        GatherConfig G;
        G.L = R->levels[it > 0 ? it - 1 : 0];
        G.r = (int) ceil(radius / G.L.res);
        G.radius2 = radius * radius;
        G.kmax = kmax;
        const QueryCtx ctx = make_query(q, G.L.res, lane);
        KnnEntry best;
        unsigned spts;
        int cnt;
        if (R->filter)
            cnt = warp_gather_knn<true>(G, nullptr, ctx, lane, s_stage[w], best, spts,
                                        V3{R->sensor.x - q.x, R->sensor.y - q.y, R->sensor.z - q.z});
        else
            cnt = warp_gather_knn<false>(G, nullptr, ctx, lane, s_stage[w], best, spts);
        if (lane == 0) out_counts[i] = cnt;
        if (lane < cnt) {
            const V3 rel = knn_rel_position(G, nullptr, ctx, best);
            double *o = out_points + ((size_t) i * kmax + (cnt - 1 - lane)) * 3;   // farthest first
            o[0] = q.x + rel.x; o[1] = q.y + rel.y; o[2] = q.z + rel.z;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
void IcpSolver::SetPeerLinks(const PeerLinksHost &links) {
    links_host_ = links;
    peers_ready_ = links.world > 1 && links.seq != nullptr;
}
void IcpSolver::AllReduceAccumulator(void *nccl_comm, IcpState *d_state) {
    if (peers_ready_) {
        k_peer_allreduce<<<1, 256, 0, stream_>>>(PeerLinksOf(links_host_), d_acc_, d_state);
        launches_ += 1;
        return;
    }
    NcclAllReduceAccumulator(nccl_comm);
}

IcpSolver::IcpSolver(cudaStream_t stream) : stream_(stream) {
    UploadPairTables();
    if (const char *e = getenv("CTICP_PERSISTENT")) use_persistent_ = atoi(e) != 0;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, dev);
    CT_CUDA_CHECK(cudaMalloc(&d_sys_, sizeof(double) * 160));
    CT_CUDA_CHECK(cudaMalloc(&d_acc_, sizeof(double) * kAcc));
    CT_CUDA_CHECK(cudaMalloc(&d_ticket_, sizeof(unsigned int)));
    CT_CUDA_CHECK(cudaMemset(d_ticket_, 0, sizeof(unsigned int)));
    CT_CUDA_CHECK(cudaMalloc(&d_sync_words_, sizeof(unsigned int) * 2));
    for (int i = 0; i < kMaxEvents; ++i) {
        CT_CUDA_CHECK(cudaEventCreate(&ev_begin_[i]));
        CT_CUDA_CHECK(cudaEventCreate(&ev_end_[i]));
    }
}
IcpSolver::~IcpSolver() {
    cudaFree(d_partials_);
    cudaFree(d_sys_);
    cudaFree(d_acc_);
    cudaFree(d_ticket_);
    cudaFree(d_sync_words_);
    FreeLmBuffers();
    for (int i = 0; i < kMaxEvents; ++i) {
        cudaEventDestroy(ev_begin_[i]);
        cudaEventDestroy(ev_end_[i]);
    }
}
void IcpSolver::EnsurePartials(int blocks) {
    if (blocks <= partial_blocks_) return;
    CT_CUDA_CHECK(cudaStreamSynchronize(stream_));
    cudaFree(d_partials_);
    CT_CUDA_CHECK(cudaMalloc(&d_partials_, sizeof(double) * kAcc * (size_t) blocks));
    partial_blocks_ = blocks;
}
GnParams IcpSolver::MakeParams(const DeviceMap &map, const cticp_icp_options &opt) const {
    GnParams P{};
    map.SearchParams(map.Options().default_radius, &P.level, &P.r);
    P.radius = map.Options().default_radius;
    P.kmax = opt.max_number_neighbors;
    P.kmin = opt.min_number_neighbors;
    P.max_dist_to_plane = opt.max_dist_to_plane_ct_icp;
    P.threshold_norm = opt.threshold_orientation_norm;
    P.shard_rank = 0;
    P.shard_world = 1;
    P.debug_flags = 0;
    if (const char *e = getenv("CTICP_DEBUG_FLAGS")) P.debug_flags = atoi(e);
    return P;
}
static int GatherBlocks(size_t k_hint, int num_sms) {
    // one warp per keypoint, kGatherWarps warps per CTA, at most 8 CTAs per SM; beyond that warps loop.
    // k_hint is an ESTIMATE of the keypoint count (the exact count lives on the device): too small only makes
    // warps iterate, too large only adds idle CTAs whose zero partials the last CTA has to sum.
    size_t want = (k_hint + kGatherWarps - 1) / kGatherWarps;
    size_t cap = (size_t) num_sms * 8;
    return (int) std::max<size_t>(1, std::min(want, cap));
}

void IcpSolver::CollectGatherTiming() {
    for (int i = 0; i < ev_used_; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev_begin_[i], ev_end_[i]) == cudaSuccess) gather_ms_ += ms;
    }
    ev_used_ = 0;
}

void IcpSolver::EnqueueGaussNewton(const DeviceMap &map, const cticp_icp_options &opt, const float4 *d_keypoints,
                                   const int *d_num_keypoints, size_t k_upper, int num_iters, IcpState *d_state,
                                   int shard_rank, int shard_world, void *nccl_comm) {
    if (opt.max_number_neighbors > 32 || opt.max_number_neighbors < 1)
        throw std::invalid_argument("max_number_neighbors must be in [1, 32]");
    GatherLaunch cfg;
    cfg.P = MakeParams(map, opt);
    cfg.P.shard_rank = shard_rank;
    cfg.P.shard_world = shard_world;
    cfg.G.L = map.Level(cfg.P.level);
    cfg.G.r = cfg.P.r;
    cfg.G.radius2 = cfg.P.radius * cfg.P.radius;
    cfg.G.kmax = cfg.P.kmax;
    const int blocks = GatherBlocks((k_upper + shard_world - 1) / shard_world + 16, num_sms_);
    EnsurePartials(blocks + 1);
    const bool peers = nccl_comm && shard_world > 1 && peers_ready_;
    if (use_persistent_ && (!nccl_comm || peers)) {
        // one cooperative launch for the whole loop; the grid must be co-resident (grid-wide barriers)
        void *kernel = peers ? (void *) k_gn_persistent<true> : (void *) k_gn_persistent<false>;
        int &coresident = max_coresident_[peers ? 1 : 0];
        if (coresident == 0) {
            int per_sm = 0;
            CT_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kGatherWarps * 32, 0));
            coresident = std::max(1, per_sm * num_sms_);
        }
        int grid = std::min(blocks + 1, coresident);
        grid = std::max(grid, 2);
        const float4 *kp = d_keypoints;
        const int *nk = d_num_keypoints;
        double *parts = d_partials_;
        int iters = num_iters;
        PeerLinks links = peers ? PeerLinksOf(links_host_) : PeerLinks{};
        void *args[] = {&cfg, &kp, &nk, &d_state, &parts, &iters, &links};
        const bool timed = time_gather_ && ev_used_ < kMaxEvents;
        if (timed) cudaEventRecord(ev_begin_[ev_used_], stream_);
        CT_CUDA_CHECK(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kGatherWarps * 32), args, 0, stream_));
        if (timed) cudaEventRecord(ev_end_[ev_used_++], stream_);
        gather_launches_ += 1;
        launches_ += 1;
        return;
    }
    for (int it = 0; it < num_iters; ++it) {
        const bool timed = time_gather_ && ev_used_ < kMaxEvents;
        if (timed) cudaEventRecord(ev_begin_[ev_used_], stream_);
        k_gn_iterate<<<blocks, kGatherWarps * 32, 0, stream_>>>(cfg, d_keypoints, d_num_keypoints, d_state, d_partials_,
                                                               d_ticket_, nccl_comm ? 2 : 0, d_acc_, nullptr);
        if (timed) cudaEventRecord(ev_end_[ev_used_++], stream_);
        ++gather_launches_;
        launches_ += 1;
        if (nccl_comm) {
            AllReduceAccumulator(nccl_comm, d_state);   // in-place sum of d_acc_ over ranks (peer mailboxes, else NCCL)
            k_gn_solve_acc<<<1, 32, 0, stream_>>>(d_acc_, d_state, cfg.P);
            launches_ += 1;
        }
    }
    CT_CUDA_CHECK(cudaGetLastError());
}

void IcpSolver::NormalEquations(const DeviceMap &map, const cticp_icp_options &opt, const float4 *d_keypoints,
                                const int *d_num_keypoints, size_t k_upper, IcpState *d_state, double *h_A144,
                                double *h_b12, int *h_n_used) {
    GatherLaunch cfg;
    cfg.P = MakeParams(map, opt);
    cfg.G.L = map.Level(cfg.P.level);
    cfg.G.r = cfg.P.r;
    cfg.G.radius2 = cfg.P.radius * cfg.P.radius;
    cfg.G.kmax = cfg.P.kmax;
    const int blocks = GatherBlocks(k_upper, num_sms_);
    EnsurePartials(blocks);
    CT_CUDA_CHECK(cudaMemsetAsync(d_sys_, 0, sizeof(double) * 160, stream_));
    k_gn_iterate<<<blocks, kGatherWarps * 32, 0, stream_>>>(cfg, d_keypoints, d_num_keypoints, d_state, d_partials_,
                                                           d_ticket_, 1, d_acc_, d_sys_);
    launches_ += 1;
    double h[160];
    CT_CUDA_CHECK(cudaMemcpyAsync(h, d_sys_, sizeof(h), cudaMemcpyDeviceToHost, stream_));
    CT_CUDA_CHECK(cudaStreamSynchronize(stream_));
    for (int i = 0; i < 144; ++i) h_A144[i] = h[i];
    for (int i = 0; i < 12; ++i) h_b12[i] = h[144 + i];
    *h_n_used = (int) (h[156] + 0.5);
}

void IcpSolver::Neighborhoods(const DeviceMap &map, const double *d_queries, size_t n, int kmax, double *d_out_points,
                              int *d_out_counts) {
    if (kmax > 32 || kmax < 1) throw std::invalid_argument("max_num_neighbors must be in [1, 32]");
    GatherConfig G;
    int level, r;
    map.SearchParams(map.Options().default_radius, &level, &r);
    G.L = map.Level(level);
    G.r = r;
    G.radius2 = map.Options().default_radius * map.Options().default_radius;
    G.kmax = kmax;
    const int blocks = GatherBlocks(n, num_sms_);
    k_neighborhoods<<<blocks, kGatherWarps * 32, 0, stream_>>>(G, d_queries, (int) n, d_out_points, d_out_counts);
    launches_ += 1;
    CT_CUDA_CHECK(cudaGetLastError());
}

void IcpSolver::RadiusSearch(const DeviceMap &map, const double *d_queries, const double *d_radiuses, size_t n, int kmax,
                             const double *sensor_location, double *d_out_points, int *d_out_counts) {
    if (kmax > 32 || kmax < 1) throw std::invalid_argument("max_num_neighbors must be in [1, 32]");
    RadiusSearchLevels R{};
    R.num_levels = map.NumLevels();
    R.filter = (sensor_location && map.Options().select_valid_normals_direction && map.HasNormals()) ? 1 : 0;
    if (sensor_location) R.sensor = V3{sensor_location[0], sensor_location[1], sensor_location[2]};
    for (int i = 0; i < R.num_levels; ++i) R.levels[i] = map.Level(i);
    RadiusSearchLevels *d_R = nullptr;
    CT_CUDA_CHECK(cudaMalloc(&d_R, sizeof(R)));
    CT_CUDA_CHECK(cudaMemcpyAsync(d_R, &R, sizeof(R), cudaMemcpyHostToDevice, stream_));
    const int blocks = GatherBlocks(n, num_sms_);
    k_radius_search<<<blocks, kGatherWarps * 32, 0, stream_>>>(d_R, kmax, d_queries, d_radiuses, (int) n, d_out_points, d_out_counts);
    launches_ += 1;
    CT_CUDA_CHECK(cudaGetLastError());
    CT_CUDA_CHECK(cudaStreamSynchronize(stream_));
    cudaFree(d_R);
}

}  // namespace cticp
