// hostcopy.cpp — worker pool + streaming copy used to stage pageable host fields into pinned
// slots (the host half of RegridFilter.forward's data movement; the reference keeps every
// field in ordinary numpy memory, filters/fields/regrid.py:204-208).
#include "hostcopy.h"

#include <immintrin.h>

#include <cstdlib>
#include <cstring>

namespace at {

namespace {

bool nontemporal_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("AT_B200_STAGE_NT");
        if (e != nullptr && e[0] == '0') return false;
        return __builtin_cpu_supports("avx2") != 0;
    }();
    return on;
}

int prefetch_distance() {
    static const int d = [] {
        const char* e = std::getenv("AT_B200_STAGE_PREFETCH");
        return e != nullptr ? std::atoi(e) : 0;  // measured: no gain on the B200 hosts
    }();
    return d;
}

__attribute__((target("avx2"))) void copy_nt_avx2(char* dst, const char* src, size_t n) {
    // head: bring dst to a 32-byte boundary
    size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head > n) head = n;
    if (head) {
        std::memcpy(dst, src, head);
        dst += head;
        src += head;
        n -= head;
    }
    size_t i = 0;
    const bool prefetch = prefetch_distance() > 0;
    const size_t ahead = static_cast<size_t>(prefetch_distance());
    for (; i + 128 <= n; i += 128) {
        // the hardware prefetcher stops at 4 KB page boundaries of the (pageable) source
        if (prefetch && i + ahead + 128 <= n) {
            _mm_prefetch(src + i + ahead, _MM_HINT_NTA);
            _mm_prefetch(src + i + ahead + 64, _MM_HINT_NTA);
        }
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    if (i < n) std::memcpy(dst + i, src + i, n - i);
    _mm_sfence();
}

}  // namespace

bool copy_streaming_is_nontemporal() { return nontemporal_enabled(); }

void copy_streaming(void* dst, const void* src, size_t bytes) {
    if (bytes >= (64u << 10) && nontemporal_enabled()) {
        copy_nt_avx2(static_cast<char*>(dst), static_cast<const char*>(src), bytes);
        return;
    }
    std::memcpy(dst, src, bytes);
}

WorkerPool::WorkerPool(int n_threads, std::function<void()> on_thread_start) : on_start_(std::move(on_thread_start)) {
    for (int t = 1; t < n_threads; ++t) workers_.emplace_back([this] { worker_main(); });
}

WorkerPool::~WorkerPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    wake_.notify_all();
    for (auto& t : workers_) t.join();
}

void WorkerPool::run_tasks() {
    for (;;) {
        const int64_t i = next_.fetch_add(1, std::memory_order_relaxed);
        if (i >= job_n_) return;
        (*job_)(i);
        if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            std::lock_guard<std::mutex> lk(mu_);
            done_.notify_all();
        }
    }
}

void WorkerPool::worker_main() {
    if (on_start_) on_start_();
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
        wake_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
        if (job_ == nullptr) continue;
        ++active_;
        lk.unlock();
        run_tasks();
        lk.lock();
        if (--active_ == 0) done_.notify_all();
    }
}

void WorkerPool::parallel_for(int64_t n, const std::function<void(int64_t)>& fn) {
    if (n <= 0) return;
    if (workers_.empty() || n == 1) {
        for (int64_t i = 0; i < n; ++i) fn(i);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        job_ = &fn;
        job_n_ = n;
        next_.store(0, std::memory_order_relaxed);
        pending_.store(n, std::memory_order_release);
        ++generation_;
    }
    wake_.notify_all();
    run_tasks();
    std::unique_lock<std::mutex> lk(mu_);
    // no worker may still be inside run_tasks when the next job is published
    done_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) == 0 && active_ == 0; });
    job_ = nullptr;
}

}  // namespace at
