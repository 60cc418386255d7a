// Host MODEL of the multi-GPU exchange protocol of ct_icp_b200/csrc/peer_exchange.cuh (threads stand in for ranks,
// std::atomic<uint64_t> words for the NVLink-mapped mailboxes). It checks the protocol's design claims, not the CUDA
// code (that is covered on 2 GPUs by tools/multigpu_check.py and the gpu-marked sharding test):
//   * LL words: a word whose upper 32 bits equal the exchange's sequence number carries that exchange's payload;
//   * two slots per source rank (parity of the sequence number) are enough: no word is overwritten before its
//     reader consumed it, for any interleaving (ranks run with random delays; a checker verifies payload integrity);
//   * the rank-ordered sum is bit-identical on every rank.
// Layout and constants mirror the device code: [parity][source rank][2 * kAcc words].
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

constexpr int kAcc = 96, kPeerWords = 2 * kAcc, kMaxPeers = 8;

struct Mailbox {
    std::vector<std::atomic<uint64_t>> words;
    Mailbox() : words(size_t(2) * kMaxPeers * kPeerWords) {
        for (auto &w : words) w.store(0, std::memory_order_relaxed);
    }
};

static double contribution(int rank, uint32_t seq, int e) {   // what rank `rank` adds in exchange `seq`, entry e
    return std::ldexp(double((rank + 1) * 1000003u ^ (seq * 2654435761u + uint32_t(e) * 40503u)), -20) - 1000.0 * e;
}

int main(int argc, char **argv) {
    const int exchanges = argc > 1 ? std::atoi(argv[1]) : 3000;
    for (int world : {2, 3, 8}) {
        std::vector<Mailbox> box(world);
        std::vector<std::vector<double>> results(world, std::vector<double>(size_t(exchanges) * kAcc));
        std::atomic<int> failures{0};
        auto rank_main = [&](int rank) {
            std::mt19937 rng(1234 + rank);
            uint32_t seq = 0;
            for (int x = 0; x < exchanges; ++x) {
                if (rng() % 7 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 50));   // skew
                ++seq;
                const size_t slot = size_t(seq & 1u) * kMaxPeers * kPeerWords;
                double acc[kAcc];
                for (int e = 0; e < kAcc; ++e) acc[e] = contribution(rank, seq, e);
                uint32_t halves[kPeerWords];
                std::memcpy(halves, acc, sizeof(acc));
                for (int wd = 0; wd < kPeerWords; ++wd)     // 1. store into every rank's mailbox
                    for (int p = 0; p < world; ++p)
                        box[p].words[slot + size_t(rank) * kPeerWords + wd].store((uint64_t(seq) << 32) | halves[wd],
                                                                                  std::memory_order_relaxed);
                double sum[kAcc];
                for (int e = 0; e < kAcc; ++e) sum[e] = 0;
                for (int r = 0; r < world; ++r) {           // 2. poll own mailbox, 3. sum in rank order
                    uint32_t got[kPeerWords];
                    for (int wd = 0; wd < kPeerWords; ++wd) {
                        uint64_t w;
                        while (uint32_t((w = box[rank].words[slot + size_t(r) * kPeerWords + wd].load(std::memory_order_relaxed)) >> 32) != seq)
                            std::this_thread::yield();
                        got[wd] = uint32_t(w);
                    }
                    double in[kAcc];
                    std::memcpy(in, got, sizeof(in));
                    for (int e = 0; e < kAcc; ++e) {
                        if (in[e] != contribution(r, seq, e)) failures++;   // payload of THIS exchange, intact
                        sum[e] += in[e];
                    }
                }
                std::memcpy(&results[rank][size_t(x) * kAcc], sum, sizeof(sum));
            }
        };
        std::vector<std::thread> ranks;
        for (int r = 0; r < world; ++r) ranks.emplace_back(rank_main, r);
        for (auto &t : ranks) t.join();
        for (int r = 1; r < world; ++r)
            if (std::memcmp(results[0].data(), results[r].data(), results[0].size() * sizeof(double)) != 0) failures++;
        if (failures.load()) {
            std::printf("FAILED world=%d failures=%d\n", world, failures.load());
            return 1;
        }
    }
    std::printf("PEER PROTOCOL OK\n");
    return 0;
}
