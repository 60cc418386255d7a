"""Host model of the k-nearest SELECTION the ICP kernels use instead of a sort (ct_icp_b200/csrc/gather_select.cuh).

The device code keeps, per query, the k nearest in-radius map points of the voxel stencil exactly as the reference's
bounded max-heap does (include/ct_icp/map.h:491-500: strict `<` replacement, so on equal distances the earlier-scanned
point stays) — but by a 32-bucket histogram of d2 + an exact ranking inside the boundary bucket, and with a compaction
when the staging area fills up. This model restates those steps one-to-one (same bucket function, same staging
capacity / batch size, same prune rule) and checks them against a plain stable sort on adversarial inputs: ties,
everything in one bucket, more candidates than the staging area, fewer than k.
"""
import random

import numpy as np
import pytest

K_SEL_CAP = 192       # kSelCap
K_BATCH = 128         # 32 * kSelPrefetch
K_BUCKETS = 32


def bucket(d2, scale):
    return np.minimum((d2 * scale).astype(np.int64), K_BUCKETS - 1)


def sel_plan(d2, kmax, scale):
    """-> (kept mask over the staged candidates, index of the farthest kept)"""
    m_total = len(d2)
    k_eff = min(m_total, kmax)
    b = bucket(d2, scale)
    hist = np.bincount(b, minlength=K_BUCKETS)
    cum = np.cumsum(hist)
    xstar = int(np.argmax(cum >= k_eff))
    c_less = int(cum[xstar] - hist[xstar])
    m = k_eff - c_less
    assert m >= 1
    e = np.flatnonzero(b == xstar)            # scan order
    flag = np.zeros(m_total, dtype=np.int64)
    for i in e:
        rank = 0
        for j in e:
            rank += int((d2[j] < d2[i]) or (d2[j] == d2[i] and j < i))
        flag[i] = (2 if rank == m - 1 else 1) if rank < m else 0
    kept = (b < xstar) | (flag != 0)
    far = int(np.flatnonzero(flag == 2)[0])
    assert kept.sum() == k_eff
    return kept, far


def device_selection(d2_all, kmax, radius2):
    """d2_all: squared distances of every stencil point in scan order. Returns (sorted original indices kept, far)."""
    scale = K_BUCKETS / radius2
    staged_d2, staged_src = [], []
    prune = np.inf
    for c0 in range(0, len(d2_all), K_BATCH):
        if len(staged_d2) + K_BATCH > K_SEL_CAP:      # sel_compact
            d = np.array(staged_d2)
            kept, far = sel_plan(d, kmax, scale)
            idx = np.flatnonzero(kept)
            if len(idx) >= kmax:
                prune = d[far]
            staged_d2 = [staged_d2[i] for i in idx]
            staged_src = [staged_src[i] for i in idx]
        for f in range(c0, min(c0 + K_BATCH, len(d2_all))):
            v = d2_all[f]
            if (not v > radius2) and v < prune:
                staged_d2.append(v)
                staged_src.append(f)
        assert len(staged_d2) <= K_SEL_CAP
    if not staged_d2:
        return [], None
    kept, far = sel_plan(np.array(staged_d2), kmax, scale)
    return [staged_src[i] for i in np.flatnonzero(kept)], staged_src[far]


def reference_selection(d2_all, kmax, radius2):
    """The reference's heap: k smallest by (distance, scan order); points[0] is the largest of them."""
    cand = [i for i, v in enumerate(d2_all) if not v > radius2]
    cand.sort(key=lambda i: (d2_all[i], i))
    kept = cand[:kmax]
    return sorted(kept), (kept[-1] if kept else None)


CASES = []
rng = np.random.default_rng(7)
for n in (0, 1, 5, 19, 20, 21, 35, 64, 150, 193, 540, 1500):
    CASES.append(("uniform%d" % n, rng.uniform(0, 1.6, n), 20))
for n in (40, 300, 900):
    CASES.append(("ties%d" % n, np.round(rng.uniform(0, 1.2, n), 1), 20))            # many exactly equal distances
    CASES.append(("one_bucket%d" % n, 0.5 + 1e-9 * rng.integers(0, 50, n), 20))       # everything in the boundary bucket
    CASES.append(("all_equal%d" % n, np.full(n, 0.25), 20))
    CASES.append(("descending%d" % n, np.linspace(0.99, 0.01, n), 20))                # every batch beats the kept set
    CASES.append(("k32_%d" % n, rng.uniform(0, 1.0, n), 32))
CASES.append(("on_radius", np.array([1.0, 1.0, 0.5, 1.0000000001, 1.0]), 20))       # d == radius is in, beyond is out
CASES.append(("k1", rng.uniform(0, 1, 50), 1))


@pytest.mark.parametrize("name,d2,kmax", CASES, ids=[c[0] for c in CASES])
def test_selection_matches_the_heap(name, d2, kmax):
    radius2 = 1.0
    got, got_far = device_selection(np.asarray(d2, dtype=np.float64), kmax, radius2)
    want, want_far = reference_selection(list(d2), kmax, radius2)
    assert sorted(got) == want
    assert got_far == want_far


def test_bucket_is_monotone():
    d = np.sort(rng.uniform(0, 1, 10000))
    b = bucket(d, K_BUCKETS / 1.0)
    assert np.all(np.diff(b) >= 0)


def test_owner_lookup_by_start_bits():
    """warp_gather_sums (gather_select.cuh, default path): the stencil slice's points form one flat list (prefix sum of the
    voxel counts); the voxel that owns flat index f is found as popc(start bits up to f) - 1 among the OCCUPIED voxels in
    scan order, the start-bit words consumed batch by batch with a running count. Model of that index arithmetic."""
    rng = random.Random(5)
    for _ in range(500):
        B = rng.choice([1, 5, 20, 24, 64])
        cnt = [rng.choice([0, 0, rng.randint(1, B)]) for _ in range(32)]
        excl = [sum(cnt[:i]) for i in range(32)]
        total = sum(cnt)
        occupied = [i for i in range(32) if cnt[i] > 0]
        own_excl, own_cnt = [excl[i] for i in occupied], [cnt[i] for i in occupied]
        starts = [0] * 64
        for e in own_excl:
            starts[e >> 5] |= 1 << (e & 31)
        for prefetch in (2, 4, 6):
            started, c0 = 0, 0
            while c0 < total:
                for u in range(prefetch):
                    if c0 + 32 * u < total:
                        word = starts[(c0 >> 5) + u]
                        for lane in range(32):
                            f = c0 + 32 * u + lane
                            rk = started + bin(word & (0xFFFFFFFF >> (31 - lane))).count("1") - 1
                            if f < total:
                                assert own_excl[rk] <= f < own_excl[rk] + own_cnt[rk]
                        started += bin(word).count("1")
                c0 += 32 * prefetch


def test_moment_reduction_partition():
    """the nine moment sums leave the lanes through shared memory: lane 3v + part adds entries [11 part, 11 part + 11 or 10)
    of row v — the three parts tile the 32 lanes exactly"""
    covered = []
    for part in range(3):
        covered += list(range(part * 11, part * 11 + (11 if part < 2 else 10)))
    assert covered == list(range(32))
    assert [lane // 3 for lane in range(27)] == [v for v in range(9) for _ in range(3)]


# ---- the exact voxel prune in front of the selection (gather_select.cuh, warp_gather_sums) -------------------------------
def voxel_box_gap2(q, c, res):
    """Squared distance from the query to the box that voxel c's points can occupy under Voxel::Coordinates' truncation
    (include/SlamCore/types.h:65-86): offsets in [0, res) for c > 0, (-res, 0] for c < 0, (-res, res) for c == 0 — with
    the device code's margin of 1e-6 res."""
    o = c * res - q
    m = 1e-6 * res
    a = o + np.where(c > 0, 0.0, -res) - m
    b = o + np.where(c < 0, 0.0, res) + m
    g = np.where(a > 0, a, np.where(b < 0, -b, 0.0))
    return float((g * g).sum())


@pytest.mark.parametrize("res,radius", [(1.0, 0.8), (0.8, 0.75), (0.2, 0.8), (1.0, 0.25)])
def test_voxel_prune_never_drops_an_in_radius_point(res, radius):
    rng = np.random.default_rng(5)
    r = int(np.ceil(radius / res))
    pruned = kept = 0
    for _ in range(3000):
        q = rng.uniform(-3 * res, 3 * res, 3)
        if rng.random() < 0.2:
            q[rng.integers(3)] = rng.choice([-res, 0.0, res]) + rng.choice([-1e-12, 0.0, 1e-12])   # on a voxel boundary
        k = np.trunc(q / res).astype(np.int64)
        c = k + rng.integers(-r, r + 1, 3)
        pts = (c * res) + rng.uniform(-res, res, (64, 3))
        pts = pts[np.all(np.trunc(pts / res).astype(np.int64) == c, axis=1)]   # the points the map files under voxel c
        if len(pts) == 0:
            continue
        # stored as fp32 offsets from the voxel origin (device_map.cuh)
        pts = c * res + (pts - c * res).astype(np.float32).astype(np.float64)
        d2 = ((pts - q) ** 2).sum(axis=1)
        if voxel_box_gap2(q, c, res) > radius * radius:
            pruned += 1
            assert not (d2 <= radius * radius).any()
        else:
            kept += 1
    assert pruned > 100 and kept > 100
