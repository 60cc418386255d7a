// Host-side unit test of cticp::HostPool (engine.h) — the fork-join team behind RegisterFrame's O(N) host passes.
// No CUDA device needed: exercised by the CPU test-suite (tests/test_host_pool.py).
//   * ParallelFor covers [0, n) exactly once for sizes around the threading threshold;
//   * ParallelRegion runs every part concurrently (a spin barrier inside the region would deadlock otherwise);
//   * the team survives the poll → sleep → wake transitions (jobs issued back to back and after pauses > 1 ms).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
#include <cstdlib>
#include <cstring>
#include <random>

#include "engine.h"

#define EXPECT(cond)                                                   \
    do {                                                               \
        if (!(cond)) {                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                  \
        }                                                              \
    } while (0)

int main() {
    for (int threads : {1, 2, 5, 8}) {
        cticp::HostPool pool(threads);
        EXPECT(pool.size() == threads);
        EXPECT(pool.PartsFor(100) == 1);
        EXPECT(pool.PartsFor(1 << 20) == threads);

        // exact coverage, each index once
        for (size_t n : {size_t(0), size_t(1), size_t(16383), size_t(16384), size_t(100003), size_t(1) << 20}) {
            std::vector<unsigned char> hit(n, 0);
            std::atomic<int> bad_part{0};
            pool.ParallelFor(n, [&](size_t b, size_t e, int part) {
                if (part < 0 || part >= pool.size()) bad_part++;
                for (size_t i = b; i < e; ++i) hit[i]++;
            });
            EXPECT(bad_part.load() == 0);
            size_t wrong = 0;
            for (size_t i = 0; i < n; ++i) wrong += hit[i] != 1;
            EXPECT(wrong == 0);
        }

        // all parts of a region run at the same time: a team barrier inside it completes
        for (int rep = 0; rep < 200; ++rep) {
            std::atomic<int> arrived{0};
            std::atomic<long long> sum{0};
            pool.ParallelRegion(size_t(1) << 20, [&](int part, int parts) {
                arrived.fetch_add(1);
                while (arrived.load() < parts) std::this_thread::yield();
                sum.fetch_add(part + 1);
            });
            EXPECT(arrived.load() == threads);
            EXPECT(sum.load() == (long long) threads * (threads + 1) / 2);
            if (rep % 50 == 49) std::this_thread::sleep_for(std::chrono::milliseconds(3));   // workers go to sleep
        }
    }
    // the AVX2 packer of contiguous float64 scans (host_pack.cpp) against the scalar definition: float32-representable
    // and not, spans / no span, ragged ends, NaN
    if (cticp::HostPackHasAvx2()) {
        std::mt19937_64 rng(3);
        std::uniform_real_distribution<double> u(-80.0, 80.0);
        for (int trial = 0; trial < 60; ++trial) {
            const size_t n = 4 * (size_t) (1 + trial * 7) + (size_t) (trial % 4);
            std::vector<double> xyz(3 * n), t(n);
            for (size_t i = 0; i < n; ++i) {
                for (int d = 0; d < 3; ++d) xyz[3 * i + d] = trial % 3 ? (double) (float) u(rng) : u(rng);
                t[i] = 100.0 + 0.1 * (double) i / (double) n;
            }
            if (trial == 7) xyz[5] = std::nan("");
            const bool spans = trial % 5 != 0;
            const double mn = 100.0, inv = spans ? 10.0 : 0.0;
            float *got = static_cast<float *>(aligned_alloc(64, 16 * n)), *want = static_cast<float *>(aligned_alloc(64, 16 * n));
            bool any_got = false, any_want = false;
            cticp::PackBlockF64Avx2(xyz.data(), t.data(), 0, n, mn, inv, spans, reinterpret_cast<float4 *>(got), &any_got);
            for (size_t i = 0; i < n; ++i) {
                const double a = spans ? (t[i] - mn) * inv : 1.0;
                for (int d = 0; d < 3; ++d) {
                    want[4 * i + d] = (float) xyz[3 * i + d];
                    if ((double) want[4 * i + d] != xyz[3 * i + d]) any_want = true;
                }
                want[4 * i + 3] = (float) a;
            }
            if (std::memcmp(got, want, 16 * n) != 0 || any_got != any_want) {
                std::printf("PackBlockF64Avx2 mismatch in trial %d (n = %zu, any %d / %d)\n", trial, n, (int) any_got, (int) any_want);
                return 1;
            }
            free(got);
            free(want);
        }
    }
    std::printf("HOST POOL OK\n");
    return 0;
}
