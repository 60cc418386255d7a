// icp_gn.cu — Gauss-Newton CT-ICP iteration as two kernels per iteration:
//   k_gn_gather : one warp per keypoint — world point from the continuous-time pose pair, 27-voxel stencil gather,
//                 kNN, covariance + eigen, point-to-plane residual and 12-vector Jacobian row, accumulation of
//                 JTJ (78 unique entries) / JTr (12) in registers, block reduction in shared memory.
//   k_gn_solve  : one block — deterministic reduction of the per-block partials, 1/n normalisation, motion-model
//                 regularisers, pivoted LDL^T solve of the 12x12 system, Euler-ZYX pose update, stop test.
// Reference: DoRegisterGaussNewton, src/ct_icp/ct_icp.cpp:709-996 (serial per-keypoint loop :753-857).
#include <cstdio>

#include "gather.cuh"
#include "icp.h"

namespace cticp {

#define CT_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                            std::to_string(__LINE__));                                                   \
    } while (0)

constexpr int kGatherWarps = 8;   // warps per CTA of the gather kernel

// (i,j) of the idx-th entry of the row-major upper triangle of a 12x12 matrix; entries 78..89 are b[0..11]
__constant__ unsigned char c_pair_i[kAccUsed];
__constant__ unsigned char c_pair_j[kAccUsed];

struct GatherLaunch {
    GatherConfig G;
    GnParams P;
};

__global__ void __launch_bounds__(kGatherWarps * 32)
k_gn_gather(GatherLaunch cfg, const float4 *__restrict__ keypoints, const int *__restrict__ d_num_keypoints,
            const IcpState *__restrict__ st, double *__restrict__ partials) {
    __shared__ KnnStage s_stage[kGatherWarps][64];
    __shared__ double s_u[kGatherWarps][16];
    __shared__ double s_acc[kGatherWarps][kAcc];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const GatherConfig &G = cfg.G;
    const GnParams &P = cfg.P;

    double acc0 = 0, acc1 = 0, acc2 = 0;              // entries lane, lane+32, lane+64 of [A upper | b]
    double n_used = 0, sum_sq = 0, n_stencil = 0, n_kp = 0, n_valid = 0;

    if (!st->done) {
        const Q4 qb{st->qb[0], st->qb[1], st->qb[2], st->qb[3]}, qe{st->qe[0], st->qe[1], st->qe[2], st->qe[3]};
        const V3 tb{st->tb[0], st->tb[1], st->tb[2]}, te{st->te[0], st->te[1], st->te[2]};
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
            const V3 p = ct_transform(qb, tb, qe, te, alpha, raw);

            KnnEntry best;
            unsigned spts = 0;
            const int n = warp_gather_knn(G, p, lane, s_stage[w], best, spts);
            n_kp += 1;
            n_stencil += (double) spts;
            if (n < P.kmin || n < 5) continue;   // ct_icp.cpp:769 ; neighborhood.h:227
            n_valid += 1;

            const NeighborhoodDesc nd = warp_describe(G, p, best, n, lane);
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
}

// ---- 12x12 pivoted LDL^T (stand-in for Eigen's A.ldlt().solve(b), ct_icp.cpp:914) ----------------------------
__device__ void ldlt_solve12(double A[12][12], const double b[12], double x[12]) {
    int perm[12];
    double D[12], y[12];
    for (int i = 0; i < 12; ++i) perm[i] = i;
    for (int k = 0; k < 12; ++k) {
        int p = k;
        double best = fabs(A[k][k]);
        for (int i = k + 1; i < 12; ++i)
            if (fabs(A[i][i]) > best) { best = fabs(A[i][i]); p = i; }
        if (p != k) {
            for (int j = 0; j < 12; ++j) { double t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
            for (int i = 0; i < 12; ++i) { double t = A[i][k]; A[i][k] = A[i][p]; A[i][p] = t; }
            int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        D[k] = A[k][k];
        if (D[k] == 0.0) {
            for (int i = k + 1; i < 12; ++i) A[i][k] = 0.0;
            continue;
        }
        for (int i = k + 1; i < 12; ++i) A[i][k] /= D[k];
        for (int i = k + 1; i < 12; ++i)
            for (int j = k + 1; j <= i; ++j) {
                A[i][j] -= A[i][k] * D[k] * A[j][k];
                A[j][i] = A[i][j];
            }
    }
    for (int i = 0; i < 12; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
    for (int i = 0; i < 12; ++i) y[i] = (fabs(D[i]) > 2.2250738585072014e-308) ? y[i] / D[i] : 0.0;
    for (int i = 11; i >= 0; --i)
        for (int j = i + 1; j < 12; ++j) y[i] -= A[j][i] * y[j];
    for (int i = 0; i < 12; ++i) x[perm[i]] = y[i];
}

// Sum the per-block partials into one accumulator vector (multi-GPU path: followed by an all-reduce)
__global__ void k_reduce_partials(const double *__restrict__ partials, int blocks, double *__restrict__ acc) {
    const int t = threadIdx.x;
    if (t < kAcc) {
        double s = 0;
        for (int b = 0; b < blocks; ++b) s += partials[(size_t) b * kAcc + t];
        acc[t] = s;
    }
}

// mode 0: full GN step (solve + pose update). mode 1: only emit the linear system (debug tap).
__global__ void k_gn_solve(const double *__restrict__ partials, int blocks, IcpState *st, GnParams P, int mode,
                           double *__restrict__ sys_out) {
    __shared__ double s_acc[kAcc];
    if (st->done && mode == 0) return;
    const int t = threadIdx.x;
    if (t < kAcc) {
        double s = 0;
        for (int b = 0; b < blocks; ++b) s += partials[(size_t) b * kAcc + t];
        s_acc[t] = s;
    }
    __syncthreads();
    if (t != 0) return;

    const int n_used = (int) (s_acc[kAccUsed] + 0.5);
    st->n_used = n_used;
    st->n_keypoints = (int) (s_acc[kAccKeypoints] + 0.5);
    st->stat_keypoint_iters += (unsigned long long) (s_acc[kAccKeypoints] + 0.5);
    st->stat_stencil_points += (unsigned long long) (s_acc[kAccStencil] + 0.5);
    if (sys_out) sys_out[156] = (double) n_used;
    if (n_used < 100) {   // ct_icp.cpp:860-871
        st->failed = 1;
        st->done = 1;
        return;
    }
    double A[12][12], b[12], x[12];
    {
        int idx = 0;
        const double inv = 1.0 / (double) n_used;   // :877-882
        for (int i = 0; i < 12; ++i)
            for (int j = i; j < 12; ++j) {
                A[i][j] = s_acc[idx] * inv;
                A[j][i] = A[i][j];
                ++idx;
            }
        for (int i = 0; i < 12; ++i) b[i] = s_acc[78 + i] * inv;
    }
    if (st->has_motion_model) {   // :885-910
        const double ac = st->beta_location, ae = st->beta_cv;
        for (int d = 0; d < 3; ++d) {
            const double diff_traj = st->tb[d] - st->te[d];   // the frame's own begin - end (sic, :892)
            A[3 + d][3 + d] += ac;
            b[3 + d] -= ac * diff_traj;
            const double diff_ego = st->te[d] - st->tb[d] - st->prev_te[d] + st->prev_tb[d];
            A[9 + d][9 + d] += ae;
            b[9 + d] -= ae * diff_ego;
        }
    }
    if (sys_out) {
        for (int i = 0; i < 12; ++i) {
            for (int j = 0; j < 12; ++j) sys_out[i * 12 + j] = A[i][j];
            sys_out[144 + i] = b[i];
        }
    }
    if (mode == 1) return;

    ldlt_solve12(A, b, x);   // :914

    Q4 qb{st->qb[0], st->qb[1], st->qb[2], st->qb[3]}, qe{st->qe[0], st->qe[1], st->qe[2], st->qe[3]};
    qb = qnormalized(qfromR(mmul(eulerZYX(x[0], x[1], x[2]), qtoR(qb))));   // :916-955 then normalize :961-962
    qe = qnormalized(qfromR(mmul(eulerZYX(x[6], x[7], x[8]), qtoR(qe))));
    st->qb[0] = qb.x; st->qb[1] = qb.y; st->qb[2] = qb.z; st->qb[3] = qb.w;
    st->qe[0] = qe.x; st->qe[1] = qe.y; st->qe[2] = qe.z; st->qe[3] = qe.w;
    for (int d = 0; d < 3; ++d) {
        st->tb[d] += x[3 + d];
        st->te[d] += x[9 + d];
    }
    double nrm = 0;
    for (int i = 0; i < 12; ++i) nrm += x[i] * x[i];
    nrm = sqrt(nrm);
    st->x_norm = nrm;
    st->iter += 1;
    if (nrm < P.threshold_norm) st->done = 1;   // :978
}

// Neighbor lists for arbitrary queries (parity tests of the map search; ComputeNeighborhoods, map.h:532-540)
__global__ void __launch_bounds__(kGatherWarps * 32)
k_neighborhoods(GatherConfig G, const double *__restrict__ queries, int n, double *__restrict__ out_points,
                int *__restrict__ out_counts) {
    __shared__ KnnStage s_stage[kGatherWarps][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = blockIdx.x * kGatherWarps + w; i < n; i += gridDim.x * kGatherWarps) {
        const V3 q{queries[3 * i], queries[3 * i + 1], queries[3 * i + 2]};
        KnnEntry best;
        unsigned spts;
        const int cnt = warp_gather_knn(G, q, lane, s_stage[w], best, spts);
        if (lane == 0) out_counts[i] = cnt;
        if (lane < cnt) {
            const V3 rel = knn_rel_position(G, q, best);
            double *o = out_points + ((size_t) i * G.kmax + (cnt - 1 - lane)) * 3;   // farthest first
            o[0] = q.x + rel.x; o[1] = q.y + rel.y; o[2] = q.z + rel.z;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
static bool g_pairs_uploaded = false;
static void UploadPairs() {
    if (g_pairs_uploaded) return;
    unsigned char pi[kAccUsed], pj[kAccUsed];
    int idx = 0;
    for (int i = 0; i < 12; ++i)
        for (int j = i; j < 12; ++j) {
            pi[idx] = (unsigned char) i;
            pj[idx] = (unsigned char) j;
            ++idx;
        }
    for (int i = 0; i < 12; ++i) {   // b[i] = Σ u[i] * u[12]  (u[12] = -scalar)
        pi[78 + i] = (unsigned char) i;
        pj[78 + i] = 12;
    }
    CT_CUDA_CHECK(cudaMemcpyToSymbol(c_pair_i, pi, sizeof(pi)));
    CT_CUDA_CHECK(cudaMemcpyToSymbol(c_pair_j, pj, sizeof(pj)));
    g_pairs_uploaded = true;
}

IcpSolver::IcpSolver(cudaStream_t stream) : stream_(stream) {
    UploadPairs();
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, dev);
    CT_CUDA_CHECK(cudaMalloc(&d_sys_, sizeof(double) * 160));
    for (int i = 0; i < kMaxEvents; ++i) {
        CT_CUDA_CHECK(cudaEventCreate(&ev_begin_[i]));
        CT_CUDA_CHECK(cudaEventCreate(&ev_end_[i]));
    }
}
IcpSolver::~IcpSolver() {
    cudaFree(d_partials_);
    cudaFree(d_sys_);
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
    return P;
}
static int GatherBlocks(size_t k_upper, int num_sms) {
    // one warp per keypoint; persistent-style cap of 4 CTAs (32 warps) per SM
    size_t want = (k_upper + kGatherWarps - 1) / kGatherWarps;
    size_t cap = (size_t) num_sms * 4;
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
    const int blocks = GatherBlocks((k_upper + shard_world - 1) / shard_world, num_sms_);
    EnsurePartials(blocks);
    for (int it = 0; it < num_iters; ++it) {
        const bool timed = time_gather_ && ev_used_ < kMaxEvents;
        if (timed) cudaEventRecord(ev_begin_[ev_used_], stream_);
        k_gn_gather<<<blocks, kGatherWarps * 32, 0, stream_>>>(cfg, d_keypoints, d_num_keypoints, d_state, d_partials_);
        if (timed) cudaEventRecord(ev_end_[ev_used_++], stream_);
        ++gather_launches_;
        if (nccl_comm) {
            AllReducePartials(nccl_comm, blocks);   // defined in nccl_shard.cu
            k_gn_solve<<<1, 128, 0, stream_>>>(d_partials_, 1, d_state, cfg.P, 0, nullptr);
            launches_ += 3;
        } else {
            k_gn_solve<<<1, 128, 0, stream_>>>(d_partials_, blocks, d_state, cfg.P, 0, nullptr);
            launches_ += 2;
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
    k_gn_gather<<<blocks, kGatherWarps * 32, 0, stream_>>>(cfg, d_keypoints, d_num_keypoints, d_state, d_partials_);
    k_gn_solve<<<1, 128, 0, stream_>>>(d_partials_, blocks, d_state, cfg.P, 1, d_sys_);
    launches_ += 2;
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

}  // namespace cticp
