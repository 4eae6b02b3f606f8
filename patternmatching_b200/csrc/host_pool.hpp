// host_pool.hpp -- a small persistent pool of host threads for the staging work around the GPU pipeline: copying a
// pageable stream into pinned memory, copying dense results out, and translating pids to the caller's pattern ids
// (8 bytes per stream byte -- the reference's read_char contract, Core/src/mps.h:41-42).
//
// Jobs are ASYNCHRONOUS: submit() cuts [0, n) into units, queues them and returns a ticket at once; wait() makes the
// caller work on the job's remaining units and then waits for the stragglers.  The engine keeps several jobs open
// (staging piece k+1 while piece k-2 is still being translated), so the workers never idle at a per-piece barrier --
// with the barrier design of the first version the host path reached half of what scripts/microbench/host_mem.cpp
// measures for the same loops on the same box.  Idle workers poll for ~100 us before they sleep on the condition
// variable: a wake-up through the kernel costs tens of microseconds, as much as a unit of work.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
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
    using Fn = std::function<void(size_t, size_t)>;   // [lo, hi) of the job's items
    struct Job {
        Fn fn;
        size_t n = 0, grain = 1, units = 0;
        size_t next = 0;               // next unclaimed unit (under mu_)
        std::atomic<size_t> done{0};   // finished units
    };
    using Ticket = std::shared_ptr<Job>;

    // cpu_first / cpu_count: the slice of the allowed CPUs this pool may pin its workers to (0 / 0 = all of them)
    explicit HostPool(int n_threads, int cpu_first = 0, int cpu_count = 0) : n_(n_threads < 1 ? 1 : n_threads) {
        // the caller takes part in every job it waits for, so n_ - 1 workers make n_ threads
        for (int i = 0; i + 1 < n_; ++i) workers_.emplace_back([this] { loop(); });
#if defined(__linux__)
        // One worker per allowed CPU, starting from the last one (the caller and the CUDA driver's own threads tend
        // to sit on the first).  A thread woken through a condition variable starts on the waker's CPU and schedulers
        // can take a long time to spread such threads; the copies here last milliseconds, so the workers are placed once.
        cpu_set_t allowed;
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
            std::vector<int> cpus;
            for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
            if (cpu_count > 0 && size_t(cpu_first) < cpus.size()) {
                const size_t hi = std::min(cpus.size(), size_t(cpu_first) + size_t(cpu_count));
                cpus = std::vector<int>(cpus.begin() + cpu_first, cpus.begin() + hi);
            }
            for (size_t w = 0; w < workers_.size() && !cpus.empty(); ++w) {
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(cpus[cpus.size() - 1 - (w % cpus.size())], &one);
                pthread_setaffinity_np(workers_[w].native_handle(), sizeof(one), &one);
            }
        }
#endif
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            epoch_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return n_; }

    // fn(lo, hi) over [0, n) in units of `grain` items (the last one may be shorter); returns at once
    Ticket submit(size_t n, size_t grain, Fn fn) {
        Ticket j = std::make_shared<Job>();
        j->fn = std::move(fn); j->n = n; j->grain = grain ? grain : 1;
        j->units = (n + j->grain - 1) / j->grain;
        if (j->units == 0) return j;
        {
            // the counter changes under the lock a worker holds from its last look at it until it sleeps: no lost wake-up
            std::lock_guard<std::mutex> lk(mu_);
            open_.push_back(j);
            epoch_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        return j;
    }
    // the caller works on the job's unclaimed units, then waits until the units other threads took are finished
    void wait(const Ticket& j) {
        if (!j || j->units == 0) return;
        for (;;) {
            size_t u;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (j->next >= j->units) break;
                u = j->next++;
                if (j->next >= j->units) drop(j.get());
            }
            run_unit(*j, u);
        }
        int spins = 0;
        while (j->done.load(std::memory_order_acquire) < j->units) {
            if (++spins < 2000) cpu_relax(); else std::this_thread::yield();
        }
    }
    void run(size_t n, size_t grain, Fn fn) { wait(submit(n, grain, std::move(fn))); }

    void copy(void* dst, const void* src, size_t bytes) {
        if (bytes < (size_t(256) << 10) || n_ == 1) { memcpy(dst, src, bytes); return; }
        run(bytes, size_t(256) << 10, [=](size_t lo, size_t hi) {
            memcpy(static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, hi - lo);
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

  private:
    static void cpu_relax() {
#if defined(__SSE2__)
        _mm_pause();
#endif
    }
    static void run_unit(Job& j, size_t u) {
        const size_t lo = u * j.grain, hi = std::min(j.n, lo + j.grain);
        j.fn(lo, hi);
        j.done.fetch_add(1, std::memory_order_release);
    }
    void drop(const Job* j) {   // under mu_: the job has no unclaimed units left
        for (auto it = open_.begin(); it != open_.end(); ++it)
            if (it->get() == j) { open_.erase(it); return; }
    }
    // oldest job first; returns false when nothing is queued
    bool claim(Ticket* j, size_t* u) {
        std::lock_guard<std::mutex> lk(mu_);
        while (!open_.empty() && open_.front()->next >= open_.front()->units) open_.pop_front();
        if (open_.empty()) return false;
        *j = open_.front();
        *u = (*j)->next++;
        if ((*j)->next >= (*j)->units) open_.pop_front();
        return true;
    }
    void loop() {
        for (;;) {
            Ticket j;
            size_t u;
            const uint64_t seen = epoch_.load(std::memory_order_acquire);   // read BEFORE looking at the queue
            if (claim(&j, &u)) { run_unit(*j, u); continue; }
            // nothing queued: poll the submit counter for a while, then sleep
            const auto t0 = std::chrono::steady_clock::now();
            bool changed = false;
            while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(100)) {
                for (int k = 0; k < 64; ++k) cpu_relax();
                if (epoch_.load(std::memory_order_acquire) != seen) { changed = true; break; }
            }
            if (changed) continue;
            std::unique_lock<std::mutex> lk(mu_);
            if (stop_) return;
            if (epoch_.load(std::memory_order_acquire) != seen) continue;
            sleepers_.fetch_add(1, std::memory_order_release);
            cv_.wait(lk, [&] { return stop_ || epoch_.load(std::memory_order_acquire) != seen; });
            sleepers_.fetch_sub(1, std::memory_order_release);
            if (stop_) return;
        }
    }
    const int n_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Ticket> open_;            // jobs with unclaimed units, oldest first
    std::atomic<uint64_t> epoch_{0};     // bumped by every submit (and by the destructor)
    std::atomic<int> sleepers_{0};
    bool stop_ = false;
};

}  // namespace pm
