// host_pool.hpp -- a small persistent pool of host threads for the staging work around the GPU pipeline: copying a
// pageable stream into pinned memory, copying dense results out, and translating pids to the caller's pattern ids
// (8 bytes per stream byte -- the reference's read_char contract, Core/src/mps.h:41-42).  One `run` at a time; the
// workers are pinned to distinct CPUs and sleep on a condition variable between runs.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif

namespace pm {

class HostPool {
  public:
    explicit HostPool(int n_threads) : n_(n_threads < 1 ? 1 : n_threads) {
        if (n_ > 1) for (int i = 0; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
#if defined(__linux__)
        // One worker per allowed CPU.  A thread woken through a condition variable tends to start on the waker's CPU
        // and schedulers can take a long time to spread such threads (measured in the build sandbox: four runnable
        // workers shared one CPU for a second); the copies here last milliseconds, so the workers are placed once.
        cpu_set_t allowed;
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
            std::vector<int> cpus;
            for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
            for (size_t w = 0; w < workers_.size() && !cpus.empty(); ++w) {
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(cpus[w % cpus.size()], &one);
                pthread_setaffinity_np(workers_[w].native_handle(), sizeof(one), &one);
            }
        }
#endif
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return n_; }
    // fn(part, n_parts) on every worker of the pool (on the caller when the pool has one thread); returns when all
    // parts are done.  The caller sleeps meanwhile: it is not pinned and would share a CPU with one of the workers.
    void run(const std::function<void(int, int)>& fn) {
        if (n_ == 1) { fn(0, 1); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            pending_ = n_;
            ++gen_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
    // [0, n) cut into n_parts pieces whose boundaries are multiples of `align`
    static void slice(size_t n, int part, int n_parts, size_t align, size_t* lo, size_t* hi) {
        const size_t units = (n + align - 1) / align;
        *lo = std::min(n, units * size_t(part) / size_t(n_parts) * align);
        *hi = std::min(n, units * size_t(part + 1) / size_t(n_parts) * align);
    }
    void copy(void* dst, const void* src, size_t bytes) {
        if (bytes < (size_t(256) << 10) || n_ == 1) { memcpy(dst, src, bytes); return; }
        run([&](int part, int parts) {
            size_t lo, hi;
            slice(bytes, part, parts, 4096, &lo, &hi);
            if (hi > lo) memcpy(static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, hi - lo);
        });
    }
    // out[j] = table[pids[j]]: 8-byte ids from 2-byte pids, written with streaming stores (the result is not read
    // back by this thread, and at 8 bytes per position it would only evict the caller's working set)
    static void expand_range(const uint16_t* pids, size_t lo, size_t hi, const uint64_t* table, uint64_t* out) {
        size_t j = lo;
#if defined(__SSE2__)
        for (; j < hi && (reinterpret_cast<uintptr_t>(out + j) & 15); ++j) out[j] = table[pids[j]];
        for (; j + 4 <= hi; j += 4) {
            const __m128i a = _mm_set_epi64x((long long)table[pids[j + 1]], (long long)table[pids[j]]);
            const __m128i b = _mm_set_epi64x((long long)table[pids[j + 3]], (long long)table[pids[j + 2]]);
            _mm_stream_si128(reinterpret_cast<__m128i*>(out + j), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(out + j + 2), b);
        }
        _mm_sfence();
#endif
        for (; j < hi; ++j) out[j] = table[pids[j]];
    }
    void expand(const uint16_t* pids, size_t n, const uint64_t* table, uint64_t* out) {
        if (n < (size_t(64) << 10) || n_ == 1) { expand_range(pids, 0, n, table, out); return; }
        run([&](int part, int parts) {
            size_t lo, hi;
            slice(n, part, parts, 512, &lo, &hi);
            expand_range(pids, lo, hi, table, out);
        });
    }

  private:
    void loop(int idx) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)>* fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(idx, n_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    const int n_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int, int)>* fn_ = nullptr;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

}  // namespace pm
