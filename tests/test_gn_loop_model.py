"""Host models of the two pieces of k_gn_persistent (ct_icp_b200/csrc/icp_gn.cu) that decide WHO does WHAT and WHEN — the
parts a numerical parity test cannot see when they go wrong only under an unlucky schedule:

* work distribution: a CTA owns a balanced contiguous range of the keypoints, its warps grab tiles of that range from a
  counter (chunks of kRowCap rows), the CTA reduces the rows in keypoint order — every keypoint exactly once, the reduced
  value independent of which warp took which tile;
* loop synchronisation: gather CTAs ARRIVE on a counter without waiting and poll the epoch the solver CTA bumps after
  publishing the state; the solver polls the counter. Modelled with real threads and random delays: the solver must only
  ever read partial rows of the iteration it is reducing, a gather CTA must only ever read the state of the iteration it is
  gathering for, early convergence and the last iteration must terminate every thread.

These restate the device code's index arithmetic and protocol one-to-one (same formulas, same order of operations)."""
import random
import threading
import time

import numpy as np
import pytest

K_GATHER_WARPS = 16   # CTICP_GATHER_WARPS
K_TILE_MAX = 16       # kTileMax
K_ROW_CAP = 256       # kRowCap
K_ROW_PARTS = 5       # kRowParts


def cta_range(lo, hi, g, num_ctas):
    """k_gn_persistent: c_lo / c_hi of gather CTA g."""
    n = hi - lo
    return lo + n * g // num_ctas, lo + n * (g + 1) // num_ctas


def tile_width(span):
    """gn_tile_width"""
    if span <= 2 * K_GATHER_WARPS:
        return 1
    return min((span + K_GATHER_WARPS - 1) // K_GATHER_WARPS, K_TILE_MAX)


def cta_gather(c_lo, c_hi, rng, rows_of):
    """gn_cta_gather + gn_gather_tiles + gn_cta_reduce_rows for one CTA. The warps are simulated by picking, at every grab,
    a random warp to be the next one to reach the counter. Returns (visits per keypoint, reduced value)."""
    w = tile_width(c_hi - c_lo)
    visits = {}
    carry = 0.0
    base = c_lo
    while True:
        top = base + K_ROW_CAP if (c_hi - base) > K_ROW_CAP else c_hi
        nxt = 0                       # R.next
        table = {}                    # R.u / R.used of the chunk
        warps_done = set()
        while len(warps_done) < K_GATHER_WARPS:
            warp = rng.choice([x for x in range(K_GATHER_WARPS) if x not in warps_done])
            j0, nxt = nxt, nxt + w    # atomicAdd(&R.next, W)
            t0 = base + j0
            if t0 >= top or top <= base:
                warps_done.add(warp)
                continue
            wt = min(w, top - t0)
            for lane in range(wt):
                assert 0 <= j0 + lane < K_ROW_CAP
                assert (j0 + lane) not in table
                table[j0 + lane] = rows_of(t0 + lane)
                visits[t0 + lane] = visits.get(t0 + lane, 0) + 1
        n = max(top - base, 0)
        assert sorted(table) == list(range(n))
        # gn_cta_reduce_rows: kRowParts interleaved classes, each in row order, the classes then in fixed order
        parts = []
        for part in range(K_ROW_PARTS):
            s = 0.0
            for r in range(part, n, K_ROW_PARTS):
                s += table[r]
            parts.append(s)
        s = parts[0]
        for q in range(1, K_ROW_PARTS):
            s += parts[q]
        carry += s
        if top >= c_hi:
            break
        base += K_ROW_CAP
    return visits, carry


@pytest.mark.parametrize("K,num_ctas,world", [(0, 147, 1), (1, 147, 1), (146, 147, 1), (2352, 147, 1), (2443, 147, 1),
                                               (4181, 147, 1), (31917, 147, 1), (131072, 147, 1), (2443, 147, 8), (700, 9, 2)])
def test_every_keypoint_once_and_order_independent(K, num_ctas, world):
    vals = np.random.default_rng(K + 1).standard_normal(max(K, 1))
    for rank in range(world):
        lo, hi = K * rank // world, K * (rank + 1) // world
        seen = {}
        totals = []
        for seed in (1, 2):           # two different schedules of the warps
            rng = random.Random(seed * 7919 + K)
            per_cta = []
            seen = {}
            for g in range(num_ctas):
                c_lo, c_hi = cta_range(lo, hi, g, num_ctas)
                assert lo <= c_lo <= c_hi <= hi
                v, s = cta_gather(c_lo, c_hi, rng, lambda k: float(vals[k]))
                for k, c in v.items():
                    seen[k] = seen.get(k, 0) + c
                per_cta.append(s)
            assert sorted(seen) == list(range(lo, hi)) and set(seen.values()) <= {1}
            totals.append(per_cta)
        assert totals[0] == totals[1]     # bit-identical partial rows whatever the schedule
    assert cta_range(0, K, 0, num_ctas)[0] == 0 and cta_range(0, K, num_ctas - 1, num_ctas)[1] == K


def test_tile_width_rule():
    assert [tile_width(n) for n in (0, 1, 16, 17, 32)] == [1, 1, 1, 1, 1]
    assert tile_width(33) == 3 and tile_width(216) == 14 and tile_width(10 ** 6) == K_TILE_MAX


# ---- the arrive / epoch protocol ----------------------------------------------------------------------------------------
class _Loop:
    def __init__(self, gather_ctas, num_iters, done_after, seed):
        self.G, self.I, self.done_after = gather_ctas, num_iters, done_after
        self.arrive = 0                 # sync.arrive
        self.epoch = 0                  # sync.epoch
        self.lock = threading.Lock()    # (atomicAdd)
        self.state_iter = 0             # the published state: which iteration it is the input of
        self.state_done = False
        self.partials = [(-1, None)] * gather_ctas   # (iteration, value) per gather CTA
        self.errors = []
        self.rng = random.Random(seed)
        self.sums = []

    def jitter(self):
        if self.rng.random() < 0.3:
            time.sleep(self.rng.random() * 2e-4)

    def gather(self, c):
        for it in range(self.I):
            if it > 0:
                deadline = time.time() + 10
                while self.epoch < it:                       # loop_wait_at_least(sync.epoch, it)
                    if time.time() > deadline:
                        self.errors.append("gather %d timed out at %d" % (c, it))
                        return
                    time.sleep(0)
            if self.state_iter != it and not self.state_done:   # the pose it gathers with must be iteration it's
                self.errors.append("gather %d read the state of %d in iteration %d" % (c, self.state_iter, it))
            if self.state_done:
                return
            self.jitter()
            self.partials[c] = (it, float(c + 1) * (it + 1))     # __stcg of the partial row; fence; barrier
            with self.lock:
                self.arrive += 1                                 # atomicAdd(sync.arrive, 1)
            if it == self.I - 1:
                return

    def solver(self):
        done = False
        for it in range(self.I):
            if done:
                break
            deadline = time.time() + 10
            while self.arrive < self.G * (it + 1):           # loop_wait_at_least(sync.arrive, G * (it + 1))
                if time.time() > deadline:
                    self.errors.append("solver timed out at %d" % it)
                    return
                time.sleep(0)
            rows = list(self.partials)                       # gn_reduce_rows
            if any(r[0] != it for r in rows):
                self.errors.append("solver reduced rows of %s in iteration %d" % (sorted({r[0] for r in rows}), it))
            self.sums.append(sum(r[1] for r in rows))
            self.jitter()
            done = self.done_after is not None and it + 1 >= self.done_after
            self.state_done = done                           # publish the state …
            self.state_iter = it + 1
            self.epoch = it + 1                              # … then the epoch (fence in between on the device)


@pytest.mark.parametrize("gather_ctas,num_iters,done_after", [(7, 5, None), (7, 5, 2), (12, 1, None), (3, 15, 15), (16, 6, 1)])
def test_arrive_epoch_protocol_under_random_schedules(gather_ctas, num_iters, done_after):
    for seed in range(6):
        L = _Loop(gather_ctas, num_iters, done_after, seed)
        threads = [threading.Thread(target=L.gather, args=(c,)) for c in range(gather_ctas)] + [threading.Thread(target=L.solver)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(30)
        assert not any(t.is_alive() for t in threads), "a thread never terminated"
        assert not L.errors, L.errors
        ran = num_iters if done_after is None else min(num_iters, done_after)
        expect = [sum(float(c + 1) * (it + 1) for c in range(gather_ctas)) for it in range(ran)]
        assert L.sums == expect
