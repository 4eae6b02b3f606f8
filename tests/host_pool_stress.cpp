// host_pool_stress.cpp -- TEST: the asynchronous host thread pool (patternmatching_b200/csrc/host_pool.hpp) under random job mixes:
// several jobs in flight, waited out of order, empty jobs, sleeping workers; built and run by tests/test_host_compiler.py.
#include "host_pool.hpp"
#include <cstdio>
#include <random>
#include <numeric>
int main(int argc, char** argv) {
    int threads = argc > 1 ? atoi(argv[1]) : 8;
    pm::HostPool pool(threads);
    std::mt19937 rng(1);
    std::vector<uint32_t> a(1 << 22), b(1 << 22);
    long checks = 0;
    for (int round = 0; round < 400; ++round) {
        // several jobs in flight, waited in a different order
        const int nj = 1 + rng() % 4;
        std::vector<pm::HostPool::Ticket> t;
        std::vector<size_t> off, len;
        size_t o = 0;
        for (int j = 0; j < nj; ++j) {
            size_t n = rng() % 300000; if (rng() % 7 == 0) n = 0;
            if (o + n > a.size()) n = a.size() - o;
            const size_t grain = 1 + rng() % 40000;
            uint32_t* pa = a.data() + o; const uint32_t tag = round * 16 + j;
            t.push_back(pool.submit(n, grain, [pa, tag](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) pa[i] = tag + uint32_t(i); }));
            off.push_back(o); len.push_back(n); o += n;
        }
        for (int j = nj - 1; j >= 0; --j) {
            pool.wait(t[j]);
            for (size_t i = 0; i < len[j]; ++i) if (a[off[j] + i] != uint32_t(round * 16 + j) + uint32_t(i)) { printf("MISMATCH round %d job %d i %zu\n", round, j, i); return 1; }
            checks += len[j];
        }
        if (round % 50 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(3));   // let the workers fall asleep
    }
    std::vector<uint16_t> pids(1 << 20); std::vector<uint64_t> table(65536), out(1 << 20);
    for (auto& x : pids) x = rng(); for (size_t i = 0; i < table.size(); ++i) table[i] = i * 3 + 1;
    const uint16_t* pp = pids.data(); const uint64_t* tt = table.data(); uint64_t* oo = out.data();
    pool.run(pids.size(), 4096, [pp, tt, oo](size_t lo, size_t hi) { pm::HostPool::expand_range(pp, lo, hi, tt, oo); });
    for (size_t i = 0; i < pids.size(); ++i) if (out[i] != table[pids[i]]) { printf("expand mismatch\n"); return 1; }
    printf("ok threads=%d checks=%ld\n", threads, checks);
    return 0;
}
