// host_mem.cpp -- what the GPU box's HOST can do for the plugin's 8-byte-per-position contract (gpu_read_block writes a
// pattern_id_t per stream byte, Core/src/mps.h:41-42): streaming-store bandwidth, memcpy bandwidth and the pid -> id
// expansion rate, per thread count.  Build: g++ -O3 -march=native -pthread host_mem.cpp -o host_mem
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <pthread.h>
#include <sched.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class F> double run_threads(int T, F f) {
    std::vector<std::thread> th;
    double t0 = now();
    for (int t = 0; t < T; ++t) {
        th.emplace_back([=] {
            cpu_set_t one; CPU_ZERO(&one); CPU_SET(t, &one); pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
            f(t, T);
        });
    }
    for (auto& x : th) x.join();
    return now() - t0;
}

static void expand_sse(const uint16_t* pids, size_t lo, size_t hi, const uint64_t* table, uint64_t* out) {
    size_t j = lo;
    for (; j + 4 <= hi; j += 4) {
        const __m128i a = _mm_set_epi64x((long long)table[pids[j + 1]], (long long)table[pids[j]]);
        const __m128i b = _mm_set_epi64x((long long)table[pids[j + 3]], (long long)table[pids[j + 2]]);
        _mm_stream_si128(reinterpret_cast<__m128i*>(out + j), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(out + j + 2), b);
    }
    _mm_sfence();
}
#ifdef __AVX512F__
static void expand_avx512(const uint16_t* pids, size_t lo, size_t hi, const uint64_t* table, uint64_t* out) {
    size_t j = lo;
    for (; j + 16 <= hi; j += 16) {
        const __m256i p16 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(pids + j));
        const __m512i idx = _mm512_cvtepu16_epi32(p16);
        const __m512i a = _mm512_i32gather_epi64(_mm512_castsi512_si256(idx), table, 8);
        const __m512i b = _mm512_i32gather_epi64(_mm512_extracti64x4_epi64(idx, 1), table, 8);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(out + j), a);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(out + j + 8), b);
    }
    _mm_sfence();
}
#endif
static void expand_plain(const uint16_t* pids, size_t lo, size_t hi, const uint64_t* table, uint64_t* out) {
    for (size_t j = lo; j < hi; ++j) out[j] = table[pids[j]];
}

int main(int argc, char** argv) {
    const size_t n = size_t(argc > 1 ? atol(argv[1]) : 256) << 20;   // positions
    const int P = 55580;
    std::vector<uint64_t> table(P + 1);
    for (int i = 1; i <= P; ++i) table[i] = 0x7F0000000000ull + 64ull * i;
    uint16_t* pids = static_cast<uint16_t*>(aligned_alloc(4096, n * 2));
    uint64_t* out = static_cast<uint64_t*>(aligned_alloc(4096, n * 8));
    uint8_t* src = static_cast<uint8_t*>(aligned_alloc(4096, n));
    // pid mix of uniform traffic: 29% none, the rest spread over ~2000 short patterns, a few long ones
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const uint32_t r = uint32_t(x >> 33);
        pids[i] = (r % 100 < 29) ? 0 : uint16_t(1 + (r >> 8) % ((r & 255) ? 2000 : P));
    }
    memset(out, 1, n * 8); memset(src, 2, n);
    const unsigned hw = std::thread::hardware_concurrency();
    printf("{\"hardware_threads\": %u, \"positions\": %zu, \"rows\": [\n", hw, n);
    const int Ts[] = {1, 2, 4, 8, 12, 16, 24, 32};
    bool first = true;
    for (int T : Ts) {
        if (unsigned(T) > hw) break;
        auto slice = [&](int t, int TT, size_t& lo, size_t& hi) { lo = n / TT * t / 64 * 64; hi = (t == TT - 1) ? n : n / TT * (t + 1) / 64 * 64; };
        double best[6] = {1e9, 1e9, 1e9, 1e9, 1e9, 1e9};
        for (int rep = 0; rep < 3; ++rep) {
            double s;
            s = run_threads(T, [&](int t, int TT) { size_t lo, hi; slice(t, TT, lo, hi);
                const __m128i v = _mm_set1_epi32(rep);
                for (size_t j = lo; j < hi; j += 2) _mm_stream_si128(reinterpret_cast<__m128i*>(out + j), v);
                _mm_sfence(); });
            if (s < best[0]) best[0] = s;
            s = run_threads(T, [&](int t, int TT) { size_t lo, hi; slice(t, TT, lo, hi); memcpy(reinterpret_cast<uint8_t*>(out) + lo, src + lo, hi - lo); });
            if (s < best[1]) best[1] = s;
            s = run_threads(T, [&](int t, int TT) { size_t lo, hi; slice(t, TT, lo, hi); expand_sse(pids, lo, hi, table.data(), out); });
            if (s < best[2]) best[2] = s;
            s = run_threads(T, [&](int t, int TT) { size_t lo, hi; slice(t, TT, lo, hi); expand_plain(pids, lo, hi, table.data(), out); });
            if (s < best[3]) best[3] = s;
#ifdef __AVX512F__
            s = run_threads(T, [&](int t, int TT) { size_t lo, hi; slice(t, TT, lo, hi); expand_avx512(pids, lo, hi, table.data(), out); });
            if (s < best[4]) best[4] = s;
#endif
            s = run_threads(T, [&](int t, int TT) { size_t lo, hi; slice(t, TT, lo, hi); memset(out + lo, rep, (hi - lo) * 8); });
            if (s < best[5]) best[5] = s;
        }
        printf("%s {\"threads\": %d, \"nt_store_GBps\": %.1f, \"memcpy_GBps_copied\": %.1f, \"expand_sse_nt_GBps_written\": %.1f, \"expand_plain_GBps_written\": %.1f, \"expand_avx512_nt_GBps_written\": %.1f, \"memset_GBps\": %.1f}",
               first ? "" : ",\n", T, n * 8 / best[0] / 1e9, n / best[1] / 1e9, n * 8 / best[2] / 1e9, n * 8 / best[3] / 1e9,
               best[4] < 1e8 ? n * 8 / best[4] / 1e9 : 0.0, n * 8 / best[5] / 1e9);
        first = false;
        fflush(stdout);
    }
    printf("\n]}\n");
    return 0;
}
