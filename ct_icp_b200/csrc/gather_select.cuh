// gather_select.cuh — the ICP kernels' neighbor gather: one warp per query, k-nearest SELECTION without a sort.
//
// Replaces MultipleResolutionVoxelMap::RadiusSearchInPlace (include/ct_icp/map.h:449-514) + the sums of
// TNeighborhood::ComputeNeighborhood (include/SlamCore/experimental/neighborhood.h:226-257) for the callers that only
// need WHICH k neighbors are kept, their first and second moments, and the reference's points[0] (the farthest kept,
// map.h:508-513) — i.e. the GN / CERES / ROBUST residual assembly (src/ct_icp/ct_icp.cpp:561-604, 753-857, 1229-1289).
// (gather.cuh's warp_gather_knn keeps the SORTED variant for the ComputeNeighborhoods API entry points.)
//
// Round-1's kernel kept the k nearest in a register-resident sorted list with a bitonic sort/merge network over
// shuffles: ~50 dependent compare-exchange steps (4 shuffles each) per keypoint, half of its ~2.9k warp instructions.
// The order of the kept neighbors is never used by the residual assembly, so this version SELECTS instead:
//   1. the stencil's points are read as one flattened list (prefix sum of the voxel counts; up to kSelPrefetch chunks of
//      32 float4 loads in flight), distances in fp64, in-radius candidates compacted IN SCAN ORDER into shared memory;
//   2. a 32-bucket histogram of d2 (bucket = floor(32 d2 / radius^2): monotone in d2, uniform for points on a surface
//      through the query) gives, by one prefix sum over the lanes, the bucket x* that holds the k-th smallest; every
//      candidate of a lower bucket is kept, and only the handful of candidates INSIDE x* are ranked exactly by
//      (d2, scan order) — the reference's strict `<` replacement (map.h:495) keeps the earlier-scanned point on a tie;
//   3. the kept candidates' first / second moments are accumulated per lane and reduced over the warp.
// The result is the exact set the reference keeps. A stencil with more in-radius candidates than the staging area holds
// is compacted to its k best (same selection) and pruned with the k-th distance from then on.
#pragma once
#include <cstddef>

#include "gather.cuh"

// -DCTICP_SEL_V1: the first cut of this file's load path, kept for A/B (tools/ab_variants.sh): owner lane of a flat candidate
// index by a binary search over shuffles, histogram counted in a separate pass, moments reduced by nine butterfly sums.
// The default path below replaces the three: owner by a start-bit mask (one popc per chunk), buckets recorded and counted
// when a candidate is staged, moments reduced through shared memory (~300 warp instructions less per query).
#if defined(CTICP_SEL_BULK) && !defined(CTICP_SEL_V1)
#define CTICP_SEL_V1 1
#endif

namespace cticp {

#ifdef CTICP_SEL_BULK
// Variant (-DCTICP_SEL_BULK, A/B of the north star's "shared-memory/TMA staging of each voxel's neighbor list"): every
// occupied voxel of the stencil is one contiguous run of <= B float4 — its owner lane issues ONE cp.async.bulk (the TMA
// engine's 1-D bulk copy, UBLKCP) into the warp's staging area at its prefix offset, one mbarrier per warp collects the
// bytes, and the distance pass reads the points from shared memory instead of issuing per-lane global loads.
constexpr int kSelCap = 128;       // staged in-radius candidates per warp (compaction checked per 32-point chunk)
constexpr int kBulkCap = 256;      // points the staging area holds; a 32-cell stencil slice beyond it uses the load path
#else
#ifndef CTICP_SEL_CAP
#define CTICP_SEL_CAP 192
#endif
constexpr int kSelCap = CTICP_SEL_CAP;   // staged in-radius candidates per warp (compaction when a batch might not fit)
#endif
constexpr int kSelBuckets = 32;    // one histogram bucket per lane
#ifndef CTICP_SEL_PREFETCH
#define CTICP_SEL_PREFETCH 4
#endif
constexpr int kSelPrefetch = CTICP_SEL_PREFETCH;

struct __align__(16) SelScratch {   // per-warp shared memory (6.5 KB)
    double d2[kSelCap];             // SoA: lane i touches word i — no bank conflicts
    double rx[kSelCap], ry[kSelCap], rz[kSelCap];   // candidate position relative to the query (fp64)
    unsigned int hist[kSelBuckets];
    unsigned char eidx[kSelCap];    // staged positions of the boundary bucket's candidates
    unsigned char eflag[kSelCap];   // per staged position (boundary bucket only): 0 dropped, 1 kept, 2 kept & farthest
    double far[4];                  // rel xyz, d2 of the farthest kept candidate (= reference points[0])
#ifndef CTICP_SEL_V1
    unsigned char bkt[kSelCap];     // histogram bucket of each staged candidate (S.hist counts the staged ones at all times)
    double mom[9];                  // reduced moments
    // the occupied voxels of the current stencil slice, in scan order: first flat index, slot, origin relative to the query
    int own_excl[32], own_slot[32];
    double own_ox[32], own_oy[32], own_oz[32];
    unsigned int starts[64];        // bit f set: a voxel's run starts at flat index f (32 cells x <= 64 points per voxel)
#endif
#ifdef CTICP_SEL_BULK
    float4 pts[kBulkCap];           // the stencil slice's points, bulk-copied
#endif
};
static_assert(kSelCap <= 256, "eidx is a byte");
#ifndef CTICP_SEL_BULK
static_assert(kSelCap >= 32 * kSelPrefetch + 32, "a batch must fit next to the k kept candidates");
#endif

#ifdef CTICP_SEL_BULK
__device__ __forceinline__ unsigned sel_smem_addr(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
// The warp's mbarrier (8 bytes of shared memory OUTSIDE any area that is reused between phases) and its phase parity.
struct SelBulk {
    unsigned long long *mbar;
    unsigned phase;
};
// once per kernel and warp, before the first query (all lanes call; lane 0 initialises)
__device__ __forceinline__ void sel_bulk_init(SelBulk &B, unsigned long long *mbar, int lane) {
    B.mbar = mbar;
    B.phase = 0;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sel_smem_addr(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
}
// bounded wait (a protocol error must end the kernel, not hang the GPU): false after ~0.5 s
__device__ __forceinline__ bool sel_bulk_wait(const SelBulk &B, unsigned parity) {
    const unsigned bar = sel_smem_addr(B.mbar);
    const long long t0 = clock64();
    unsigned done = 0;
    while (true) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return true;
        if (clock64() - t0 > 1000000000LL) return false;
    }
}
#endif

struct NeighborSums {
    int n;                                             // neighbors kept (min(kmax, in-radius candidates))
    double sx, sy, sz, sxx, sxy, sxz, syy, syz, szz;   // Σ rel, Σ rel rel^T over the kept neighbors (query-centred)
    double fx, fy, fz, fd2;                            // farthest kept neighbor: rel position, squared distance
};

__device__ __forceinline__ int sel_bucket(double d2, double scale) {
    const int b = f64_floor_nonneg(d2 * scale);   // (= (int) (d2 * scale) for d2 >= 0, off the XU pipe: se3.cuh)
    return b < kSelBuckets - 1 ? b : kSelBuckets - 1;
}

struct SelPlan {
    int k_eff;   // neighbors to keep
    int xstar;   // boundary bucket: lower buckets are kept whole
};

// Steps 2 of the header comment on the M staged candidates: fills S.eflag for the candidates of the boundary bucket.
__device__ __forceinline__ SelPlan sel_plan(SelScratch &S, int M, int kmax, double scale, int lane) {
    const unsigned lt_mask = (1u << lane) - 1u;
#ifdef CTICP_SEL_V1
    S.hist[lane] = 0;
    __syncwarp();
    for (int i = lane; i < M; i += 32) atomicAdd(&S.hist[sel_bucket(S.d2[i], scale)], 1u);
#endif
    __syncwarp();
    const int h = (int) S.hist[lane];
    int cum = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, cum, o);
        if (lane >= o) cum += y;
    }
    SelPlan P;
    P.k_eff = M < kmax ? M : kmax;
    const unsigned ge = __ballot_sync(0xffffffffu, cum >= P.k_eff);   // non-zero: cum[31] = M >= k_eff
    P.xstar = __ffs(ge) - 1;
    const int c_less = __shfl_sync(0xffffffffu, cum - h, P.xstar);    // candidates in lower buckets
    const int m = P.k_eff - c_less;                                   // to keep from the boundary bucket (>= 1)
    // the boundary bucket's candidates, in scan order
    int ne = 0;
    for (int base = 0; base < M; base += 32) {
        const int i = base + lane;
#ifdef CTICP_SEL_V1
        const bool is_e = i < M && sel_bucket(S.d2[i], scale) == P.xstar;
#else
        const bool is_e = i < M && (int) S.bkt[i] == P.xstar;
#endif
        const unsigned mask = __ballot_sync(0xffffffffu, is_e);
        if (is_e) S.eidx[ne + __popc(mask & lt_mask)] = (unsigned char) i;
        ne += __popc(mask);
    }
    __syncwarp();
    // exact rank inside the bucket by (d2, scan order): kept if rank < m, rank m-1 is the farthest kept
    for (int t0 = 0; t0 < ne; t0 += 32) {
        const int t = t0 + lane;
        const int i = t < ne ? (int) S.eidx[t] : 0;
        const double di = S.d2[i];
        int rank = 0;
        for (int u = 0; u < ne; ++u) {
            const int j = (int) S.eidx[u];
            const double dj = S.d2[j];
            rank += (int) ((dj < di) | ((dj == di) & (j < i)));
        }
        if (t < ne) S.eflag[i] = (unsigned char) (rank < m ? (rank == m - 1 ? 2 : 1) : 0);
    }
    __syncwarp();
    return P;
}

// Staging area full: keep only the k best (in scan order, at the front of the staging area); returns the new fill and
// the prune threshold (the k-th distance: a later candidate at or beyond it can never be kept, map.h:495).
__device__ __forceinline__ int sel_compact(SelScratch &S, int M, int kmax, double scale, int lane, double &prune_d2) {
    const unsigned lt_mask = (1u << lane) - 1u;
    const SelPlan P = sel_plan(S, M, kmax, scale, lane);
#ifndef CTICP_SEL_V1
    S.hist[lane] = 0;   // (sel_plan has read it) recounted below for the candidates that stay
    __syncwarp();
#endif
    int out = 0;
    for (int base = 0; base < M; base += 32) {
        const int i = base + lane;
        double d = 0, x = 0, y = 0, z = 0;
        bool kept = false;
        int b = 0;
        if (i < M) {
            d = S.d2[i]; x = S.rx[i]; y = S.ry[i]; z = S.rz[i];
#ifdef CTICP_SEL_V1
            b = sel_bucket(d, scale);
#else
            b = (int) S.bkt[i];
#endif
            const int flag = b == P.xstar ? (int) S.eflag[i] : 0;
            kept = b < P.xstar || flag != 0;
            if (flag == 2) S.far[3] = d;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, kept);
        __syncwarp();   // every lane has read its entry before any lane overwrites one
        if (kept) {
            const int o = out + __popc(mask & lt_mask);   // o <= i: never clobbers an unread entry of a later round
            S.d2[o] = d; S.rx[o] = x; S.ry[o] = y; S.rz[o] = z;
#ifndef CTICP_SEL_V1
            S.bkt[o] = (unsigned char) b;
            atomicAdd(&S.hist[b], 1u);
#endif
        }
        out += __popc(mask);
        __syncwarp();
    }
    if (out >= kmax) prune_d2 = S.far[3];
    return out;
}

// One query, one warp. `origin_*`: all lanes pass the same query (world position q, its voxel kx ky kz).
// need: callers ignore neighborhoods with fewer than `need` neighbors — the moments are then skipped (out.n is set).
// kFilter: RadiusSearchInPlace with a sensor_location (map.h:482-490), `to_sensor` = sensor_location - query.
template <bool kFilter>
__device__ __forceinline__ void warp_gather_sums(const GatherConfig &G, double bucket_scale, const int *stencil,
                                                 const V3 &q, int kx, int ky, int kz, int need, int lane,
                                                 SelScratch &S, NeighborSums &out, unsigned &stencil_points,
                                                 V3 to_sensor = V3{0, 0, 0}, void *bulk_ptr = nullptr,
                                                 double *ranked = nullptr, int n_ranked = 0) {
#ifdef CTICP_SEL_BULK
    SelBulk *bulk = static_cast<SelBulk *>(bulk_ptr);   // nullptr: this caller uses the load path
#endif
    const MapLevel &L = G.L;
    const int side = 2 * G.r + 1;
    const int nst = side * side * side;
    const unsigned lt_mask = (1u << lane) - 1u;
    int fill = 0;
    unsigned pts_total = 0;
    double prune_d2 = kKnnInf;
#ifndef CTICP_SEL_V1
    S.hist[lane] = 0;   // counts the staged candidates per bucket as they arrive (visible after the slice's first __syncwarp)
#endif

    for (int base = 0; base < nst; base += 32) {
        const int s = base + lane;
        int slot = 0, cnt = 0;
        double ox = 0, oy = 0, oz = 0;   // my voxel's origin relative to the query (fp64)
        double sdn = 0;                  // kFilter: to_sensor . voxel normal
        int has_normal = 0;
        int cx = 0, cy = 0, cz = 0;      // my voxel's coordinates
        if (s < nst) {
            int dx, dy, dz;
            stencil_lookup(stencil, s, G.r, dx, dy, dz);
            cx = kx + dx; cy = ky + dy; cz = kz + dz;
            uint32_t c = 0;
            const int found = map_find(L, pack_voxel(kx + dx, ky + dy, kz + dz), &c);
            if (found >= 0) {
                slot = found;
                cnt = (int) c;
                if (kFilter && L.normals && c > 0) {
                    const double *nrm = L.normals + 4 * (size_t) found;
                    if (nrm[3] != 0.0) {
                        has_normal = 1;
                        sdn = to_sensor.x * nrm[0] + to_sensor.y * nrm[1] + to_sensor.z * nrm[2];
                    }
                }
            }
            ox = i32_to_f64(kx + dx) * L.res - q.x;
            oy = i32_to_f64(ky + dy) * L.res - q.y;
            oz = i32_to_f64(kz + dz) * L.res - q.z;
        }
        pts_total += (unsigned) __reduce_add_sync(0xffffffffu, cnt);   // (statistics: every point of the stencil, map.h:470-506)
#ifndef CTICP_NO_VOXEL_PRUNE
        // Exact prune: a voxel whose box lies farther than the radius from the query cannot hold a candidate (every one of
        // its points would fail the radius test below), so its points are not loaded. Voxel c of Voxel::Coordinates'
        // truncation holds offsets in [0, res) for c > 0, (-res, 0] for c < 0 and (-res, res) for c == 0 (types.h:65-86);
        // a margin of 1e-6 res covers the rounding of the fp32 offsets and of the quotient that assigned the voxel.
        if (cnt > 0) {
            const double m = 1e-6 * L.res;
            const double ax = ox + (cx > 0 ? 0.0 : -L.res) - m, bx = ox + (cx < 0 ? 0.0 : L.res) + m;
            const double ay = oy + (cy > 0 ? 0.0 : -L.res) - m, by = oy + (cy < 0 ? 0.0 : L.res) + m;
            const double az = oz + (cz > 0 ? 0.0 : -L.res) - m, bz = oz + (cz < 0 ? 0.0 : L.res) + m;
            const double gx = ax > 0 ? ax : (bx < 0 ? -bx : 0.0), gy = ay > 0 ? ay : (by < 0 ? -by : 0.0),
                         gz = az > 0 ? az : (bz < 0 ? -bz : 0.0);
            if (gx * gx + gy * gy + gz * gz > G.radius2) cnt = 0;
        }
#endif
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - cnt;

#ifdef CTICP_SEL_BULK
        bool staged = false;
        if (bulk && total > 0 && total <= kBulkCap) {
            // one bulk copy per occupied voxel into the staging area at its prefix offset; the mbarrier counts the bytes
            const unsigned bar = sel_smem_addr(bulk->mbar);
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned) total * 16u) : "memory");
            __syncwarp();
            if (cnt > 0)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sel_smem_addr(&S.pts[excl])), "l"(L.points + (size_t) slot * L.B), "r"((unsigned) cnt * 16u), "r"(bar)
                             : "memory");
            const unsigned parity = bulk->phase & 1u;
            staged = sel_bulk_wait(*bulk, parity);
            bulk->phase = parity ^ 1u;   // (per-lane copy of the same value)
            __syncwarp();
        }
        if (staged) {
            for (int c0 = 0; c0 < total; c0 += 32) {
                if (fill + 32 > kSelCap) fill = sel_compact(S, fill, G.kmax, bucket_scale, lane, prune_d2);
                const int f = c0 + lane;
                int lo = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int probe = lo + step;
                    const int ex = __shfl_sync(0xffffffffu, excl, probe & 31);
                    if (probe < 32 && ex <= f) lo = probe;
                }
                const bool valid = f < total;
                const float4 p4 = valid ? S.pts[f] : make_float4(0.f, 0.f, 0.f, 0.f);
                const double vx = __shfl_sync(0xffffffffu, ox, lo), vy = __shfl_sync(0xffffffffu, oy, lo),
                             vz = __shfl_sync(0xffffffffu, oz, lo);
                const double rx = vx + f32_to_f64(p4.x), ry = vy + f32_to_f64(p4.y), rz = vz + f32_to_f64(p4.z);
                const double d2 = rx * rx + ry * ry + rz * rz;
                bool in = valid && !(d2 > G.radius2) && d2 < prune_d2;
                if (kFilter) {
                    const double vs = __shfl_sync(0xffffffffu, sdn, lo);
                    const int vh = __shfl_sync(0xffffffffu, has_normal, lo);
                    const double scalar = signbit(p4.w) ? -vs : vs;
                    if (vh && scalar < 0.0) in = false;
                }
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const int o = fill + __popc(m & lt_mask);
                    S.d2[o] = d2; S.rx[o] = rx; S.ry[o] = ry; S.rz[o] = rz;
                }
                fill += __popc(m);
            }
            __syncwarp();   // the staging area is read before the next slice's copies overwrite it
            continue;
        }
        // slices beyond the staging area (or a timed-out copy: never expected) take the load path below
        for (int c0 = 0; c0 < total; c0 += 32) {
            if (fill + 32 > kSelCap) fill = sel_compact(S, fill, G.kmax, bucket_scale, lane, prune_d2);
            const int f = c0 + lane;
            int lo = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int probe = lo + step;
                const int ex = __shfl_sync(0xffffffffu, excl, probe & 31);
                if (probe < 32 && ex <= f) lo = probe;
            }
            const int o_excl = __shfl_sync(0xffffffffu, excl, lo);
            const int o_slot = __shfl_sync(0xffffffffu, slot, lo);
            const bool valid = f < total;
            const float4 p4 = valid ? __ldg(L.points + (size_t) o_slot * L.B + (f - o_excl)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const double vx = __shfl_sync(0xffffffffu, ox, lo), vy = __shfl_sync(0xffffffffu, oy, lo),
                         vz = __shfl_sync(0xffffffffu, oz, lo);
            const double rx = vx + f32_to_f64(p4.x), ry = vy + f32_to_f64(p4.y), rz = vz + f32_to_f64(p4.z);
            const double d2 = rx * rx + ry * ry + rz * rz;
            bool in = valid && !(d2 > G.radius2) && d2 < prune_d2;
            if (kFilter) {
                const double vs = __shfl_sync(0xffffffffu, sdn, lo);
                const int vh = __shfl_sync(0xffffffffu, has_normal, lo);
                const double scalar = signbit(p4.w) ? -vs : vs;
                if (vh && scalar < 0.0) in = false;
            }
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) {
                const int o = fill + __popc(m & lt_mask);
                S.d2[o] = d2; S.rx[o] = rx; S.ry[o] = ry; S.rz[o] = rz;
            }
            fill += __popc(m);
        }
#elif !defined(CTICP_SEL_V1)
        // the slice's occupied voxels in scan order (rank = position among them) and the start bits of their runs: the
        // owner of flat index f is voxel number popc(start bits up to f) - 1 — one LDS + popc per chunk instead of a
        // binary search over shuffles
        const unsigned occ = __ballot_sync(0xffffffffu, cnt > 0);
        S.starts[lane] = 0u;
        S.starts[lane + 32] = 0u;
        __syncwarp();
        if (cnt > 0) {
            const int rk = __popc(occ & lt_mask);
            S.own_excl[rk] = excl; S.own_slot[rk] = slot;
            S.own_ox[rk] = ox; S.own_oy[rk] = oy; S.own_oz[rk] = oz;
            atomicOr(&S.starts[excl >> 5], 1u << (excl & 31));
        }
        __syncwarp();
        int started = 0;   // runs that start before the current batch (warp-uniform)
        for (int c0 = 0; c0 < total; c0 += 32 * kSelPrefetch) {
            // room for a whole batch (checked once per batch: one copy of the compaction code, off the common path)
            if (fill + 32 * kSelPrefetch > kSelCap) fill = sel_compact(S, fill, G.kmax, bucket_scale, lane, prune_d2);
            float4 pv[kSelPrefetch];
            int owner[kSelPrefetch];   // rank of the owning voxel, -1: no candidate
            // phase 1: locate and issue every load of this batch
#pragma unroll
            for (int u = 0; u < kSelPrefetch; ++u) {
                const int f = c0 + 32 * u + lane;   // flat candidate index
                owner[u] = -1;
                pv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + 32 * u < total) {           // warp-uniform
                    const unsigned word = S.starts[(c0 >> 5) + u];
                    const int rk = started + __popc(word & (0xffffffffu >> (31 - lane))) - 1;
                    started += __popc(word);
                    if (f < total) {
                        owner[u] = rk;
                        pv[u] = __ldg(L.points + (size_t) S.own_slot[rk] * L.B + (f - S.own_excl[rk]));
                    }
                }
            }
            // phase 2: distances, radius test, compaction (in scan order) into the staging area
#pragma unroll
            for (int u = 0; u < kSelPrefetch; ++u) {
                if (c0 + 32 * u < total) {           // warp-uniform
                    const int ol = owner[u] < 0 ? 0 : owner[u];
                    const double rx = S.own_ox[ol] + f32_to_f64(pv[u].x), ry = S.own_oy[ol] + f32_to_f64(pv[u].y),
                                 rz = S.own_oz[ol] + f32_to_f64(pv[u].z);
                    const double d2 = rx * rx + ry * ry + rz * rz;
                    bool in = owner[u] >= 0 && !(d2 > G.radius2) && d2 < prune_d2;
                    if (kFilter) {
                        // (the filter's per-voxel terms stay with the voxel's lane: found again through the scan order)
                        const int vl = __fns(occ, 0, ol + 1);
                        const double vs = __shfl_sync(0xffffffffu, sdn, vl & 31);
                        const int vh = __shfl_sync(0xffffffffu, has_normal, vl & 31);
                        // this point's copy of the normal is -n when its w is negative
                        const double scalar = signbit(pv[u].w) ? -vs : vs;
                        if (vh && scalar < 0.0) in = false;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const int o = fill + __popc(m & lt_mask);
                        const int b = sel_bucket(d2, bucket_scale);
                        S.d2[o] = d2; S.rx[o] = rx; S.ry[o] = ry; S.rz[o] = rz;
                        S.bkt[o] = (unsigned char) b;
                        atomicAdd(&S.hist[b], 1u);
                    }
                    fill += __popc(m);
                }
            }
        }
        __syncwarp();   // the owner tables are read before the next slice rewrites them
#else
        for (int c0 = 0; c0 < total; c0 += 32 * kSelPrefetch) {
            // room for a whole batch (checked once per batch: one copy of the compaction code, off the common path)
            if (fill + 32 * kSelPrefetch > kSelCap) fill = sel_compact(S, fill, G.kmax, bucket_scale, lane, prune_d2);
            float4 pv[kSelPrefetch];
            int owner[kSelPrefetch];
            // phase 1: locate and issue every load of this batch
#pragma unroll
            for (int u = 0; u < kSelPrefetch; ++u) {
                const int f = c0 + 32 * u + lane;   // flat candidate index
                owner[u] = -1;
                pv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + 32 * u < total) {           // warp-uniform
                    int lo = 0;                      // owner = last lane whose exclusive prefix is <= f
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const int probe = lo + step;
                        const int ex = __shfl_sync(0xffffffffu, excl, probe & 31);
                        if (probe < 32 && ex <= f) lo = probe;
                    }
                    const int o_excl = __shfl_sync(0xffffffffu, excl, lo);
                    const int o_slot = __shfl_sync(0xffffffffu, slot, lo);
                    if (f < total) {
                        owner[u] = lo;
                        pv[u] = __ldg(L.points + (size_t) o_slot * L.B + (f - o_excl));
                    }
                }
            }
            // phase 2: distances, radius test, compaction (in scan order) into the staging area
#pragma unroll
            for (int u = 0; u < kSelPrefetch; ++u) {
                if (c0 + 32 * u < total) {           // warp-uniform
                    const int ol = owner[u] < 0 ? 0 : owner[u];
                    const double vx = __shfl_sync(0xffffffffu, ox, ol), vy = __shfl_sync(0xffffffffu, oy, ol),
                                 vz = __shfl_sync(0xffffffffu, oz, ol);
                    const double rx = vx + f32_to_f64(pv[u].x), ry = vy + f32_to_f64(pv[u].y), rz = vz + f32_to_f64(pv[u].z);
                    const double d2 = rx * rx + ry * ry + rz * rz;
                    bool in = owner[u] >= 0 && !(d2 > G.radius2) && d2 < prune_d2;
                    if (kFilter) {
                        const double vs = __shfl_sync(0xffffffffu, sdn, ol);
                        const int vh = __shfl_sync(0xffffffffu, has_normal, ol);
                        // this point's copy of the normal is -n when its w is negative
                        const double scalar = signbit(pv[u].w) ? -vs : vs;
                        if (vh && scalar < 0.0) in = false;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const int o = fill + __popc(m & lt_mask);
                        S.d2[o] = d2; S.rx[o] = rx; S.ry[o] = ry; S.rz[o] = rz;
                    }
                    fill += __popc(m);
                }
            }
        }
    #endif
    }
    __syncwarp();
    stencil_points = pts_total;
    out.n = fill < G.kmax ? fill : G.kmax;
    if (out.n < need || fill == 0) return;

    const SelPlan P = sel_plan(S, fill, G.kmax, bucket_scale, lane);
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int base = 0; base < fill; base += 32) {
        const int i = base + lane;
        if (i < fill) {
            const double d = S.d2[i];
#ifdef CTICP_SEL_V1
            const int b = sel_bucket(d, bucket_scale);
#else
            const int b = (int) S.bkt[i];
#endif
            const int flag = b == P.xstar ? (int) S.eflag[i] : 0;
            if (b < P.xstar || flag != 0) {
                const double x = S.rx[i], y = S.ry[i], z = S.rz[i];
                a[0] += x; a[1] += y; a[2] += z;
                a[3] += x * x; a[4] += x * y; a[5] += x * z;
                a[6] += y * y; a[7] += y * z; a[8] += z * z;
                if (flag == 2) { S.far[0] = x; S.far[1] = y; S.far[2] = z; S.far[3] = d; }
            }
        }
    }
    if (ranked) {
        // ranked[3 i .. 3 i + 2]: position (relative to the query) of the i-th farthest kept neighbor, i < n_ranked — the head
        // of the reference's neighbor list, which comes out of its max-heap farthest first (map.h:508-513). Only solver CERES
        // with num_closest_neighbors > 1 asks for more than points[0] (ct_icp.cpp:593-601): an O(M^2 / 32) ranking of the kept
        // candidates by (d2, scan order) descending, off every default path.
        for (int base = 0; base < fill; base += 32) {
            const int i = base + lane;
            bool kept_i = false;
            double di = 0;
            if (i < fill) {
                di = S.d2[i];
#ifdef CTICP_SEL_V1
                const int b = sel_bucket(di, bucket_scale);
#else
                const int b = (int) S.bkt[i];
#endif
                kept_i = b < P.xstar || (b == P.xstar && S.eflag[i] != 0);
            }
            int rank = 0;
            for (int j = 0; j < fill; ++j) {
                const double dj = S.d2[j];
#ifdef CTICP_SEL_V1
                const int bj = sel_bucket(dj, bucket_scale);
#else
                const int bj = (int) S.bkt[j];
#endif
                const bool kept_j = bj < P.xstar || (bj == P.xstar && S.eflag[j] != 0);
                rank += (int) (kept_j && ((dj > di) || (dj == di && j > i)));
            }
            if (kept_i && rank < n_ranked) {
                ranked[3 * rank] = S.rx[i];
                ranked[3 * rank + 1] = S.ry[i];
                ranked[3 * rank + 2] = S.rz[i];
            }
        }
        __syncwarp();
    }
#ifdef CTICP_SEL_V1
#pragma unroll
    for (int v = 0; v < 9; ++v) a[v] = warp_sum(a[v]);
    __syncwarp();
#else
    // Σ over the lanes through shared memory: the staging arrays are free now (rows of 33 words, conflict-free both ways;
    // d2 .. rz are adjacent: 9 x 33 = 297 <= 4 kSelCap); lane 3v + part sums a third of row v, three partials per moment
    // meet by two shuffles — 9 stores + 11 loads per lane instead of 45 double shuffles. Fixed order: deterministic.
    {
        static_assert(offsetof(SelScratch, rx) == offsetof(SelScratch, d2) + sizeof(double) * kSelCap &&
                          offsetof(SelScratch, rz) == offsetof(SelScratch, d2) + 3 * sizeof(double) * kSelCap &&
                          4 * kSelCap >= 9 * 33,
                      "reduction scratch");
        __syncwarp();   // every lane is done reading the staged candidates
        double *red = S.d2;
#pragma unroll
        for (int v = 0; v < 9; ++v) red[v * 33 + lane] = a[v];
        __syncwarp();
        const int v = lane / 3, part = lane - 3 * v;   // lanes 27..31: v = 9, idle
        double t = 0;
        if (v < 9) {
            const double *row = red + v * 33 + part * 11;
            const int cntp = part < 2 ? 11 : 10;
            for (int e = 0; e < cntp; ++e) t += row[e];
        }
        const double t1 = __shfl_down_sync(0xffffffffu, t, 1), t2 = __shfl_down_sync(0xffffffffu, t, 2);
        if (v < 9 && part == 0) S.mom[v] = (t + t1) + t2;
        __syncwarp();
#pragma unroll
        for (int v2 = 0; v2 < 9; ++v2) a[v2] = S.mom[v2];
    }
#endif
    out.sx = a[0]; out.sy = a[1]; out.sz = a[2];
    out.sxx = a[3]; out.sxy = a[4]; out.sxz = a[5];
    out.syy = a[6]; out.syz = a[7]; out.szz = a[8];
    out.fx = S.far[0]; out.fy = S.far[1]; out.fz = S.far[2]; out.fd2 = S.far[3];
    __syncwarp();   // S.far is read before the next query's selection rewrites it
}

// 1.0 / n for the neighbor counts (n <= 32): a table of the correctly rounded quotients instead of I2F + MUFU.RCP64H + Newton
__constant__ double c_inv_count[33] = {0.0, 1.0 / 1, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10,
                                       1.0 / 11, 1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19,
                                       1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28,
                                       1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32};
__device__ __forceinline__ double inv_count(int n) { return (n >= 1 && n <= 32) ? c_inv_count[n] : 1.0 / (double) n; }

// TNeighborhood::ComputeNeighborhood + ComputeNeighborhoodInfo (neighborhood.h:226-257, 286-316) from the moments: one
// THREAD per neighborhood (the epilogue of a tile of queries runs lane-per-query).
__device__ __forceinline__ NeighborhoodDesc describe_from_sums(const NeighborSums &s) {
    const double inv = inv_count(s.n);
    const double mx = s.sx * inv, my = s.sy * inv, mz = s.sz * inv;
    const Eig3 e = sym_eig3_fast(s.sxx * inv - mx * mx, s.sxy * inv - mx * my, s.sxz * inv - mx * mz,
                                 s.syy * inv - my * my, s.syz * inv - my * mz, s.szz * inv - mz * mz);
    NeighborhoodDesc d;
    d.normal = e.normal;
    d.a2D = (sqrt(e.sv1) - sqrt(e.sv2)) * rsqrt(e.sv0);   // (:302; one MUFU seed less than a division by sqrt)
    d.far_rel = V3{s.fx, s.fy, s.fz};
    d.far_d2 = s.fd2;
    return d;
}
__device__ __forceinline__ NeighborhoodDescFull describe_full_from_sums(const NeighborSums &s) {
    const double inv = inv_count(s.n);
    const double mx = s.sx * inv, my = s.sy * inv, mz = s.sz * inv;
    NeighborhoodDescFull d;
    d.cov[0] = s.sxx * inv - mx * mx;
    d.cov[1] = s.sxy * inv - mx * my;
    d.cov[2] = s.sxz * inv - mx * mz;
    d.cov[3] = s.syy * inv - my * my;
    d.cov[4] = s.syz * inv - my * mz;
    d.cov[5] = s.szz * inv - mz * mz;
    const Eig3Full e = sym_eig3_full(d.cov[0], d.cov[1], d.cov[2], d.cov[3], d.cov[4], d.cov[5]);
    d.normal = e.normal;
    d.line = e.line;
    d.linearity = (e.sv0 - e.sv1) / e.sv0;
    d.planarity = (e.sv1 - e.sv2) / e.sv0;
    d.mean_rel = V3{mx, my, mz};
    d.far_rel = V3{s.fx, s.fy, s.fz};
    return d;
}

// lane `src`'s copy of the sums to every lane's registers is not needed: the tile loops keep the result of query j in
// lane j only (all lanes hold the same totals after the warp reductions).
__device__ __forceinline__ void sums_keep_if(bool mine, NeighborSums &dst, const NeighborSums &src) {
    if (mine) dst = src;
}

}  // namespace cticp
