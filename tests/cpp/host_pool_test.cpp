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
    std::printf("HOST POOL OK\n");
    return 0;
}
