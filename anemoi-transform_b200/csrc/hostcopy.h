// hostcopy.h — host-side helpers of the field I/O engine (hostio.cu): a persistent worker
// pool and a streaming memory copy.  Plain C++ (compiled by the host compiler, no CUDA).
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace at {

// dst[0:bytes) = src[0:bytes).  Large copies use non-temporal stores (the destination is a
// pinned staging slot that the DMA engine reads next, never the CPU), followed by a store
// fence so the data is globally visible before the copy is handed to the device.
void copy_streaming(void* dst, const void* src, size_t bytes);

// Whether copy_streaming uses the AVX2 non-temporal path on this CPU (AT_B200_STAGE_NT=0 disables it).
bool copy_streaming_is_nontemporal();

// n_threads - 1 persistent workers; the caller of parallel_for takes a share of the tasks.
class WorkerPool {
   public:
    // `on_thread_start` runs once in every worker (used to bind the worker to a CUDA device).
    WorkerPool(int n_threads, std::function<void()> on_thread_start);
    ~WorkerPool();
    WorkerPool(const WorkerPool&) = delete;
    WorkerPool& operator=(const WorkerPool&) = delete;

    int size() const { return static_cast<int>(workers_.size()) + 1; }
    // fn(i) for i in [0, n), spread dynamically over the threads; returns when all are done.
    void parallel_for(int64_t n, const std::function<void(int64_t)>& fn);

   private:
    void worker_main();
    void run_tasks();

    std::vector<std::thread> workers_;
    std::function<void()> on_start_;
    std::mutex mu_;
    std::condition_variable wake_, done_;
    const std::function<void(int64_t)>* job_ = nullptr;
    int64_t job_n_ = 0;
    std::atomic<int64_t> next_{0};
    std::atomic<int64_t> pending_{0};
    uint64_t generation_ = 0;
    int active_ = 0;
    bool stop_ = false;
};

}  // namespace at
