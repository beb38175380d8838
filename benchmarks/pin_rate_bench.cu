// How fast can the pinned pool grow?  cudaHostAlloc from 1 / 4 / 8 threads, and
// mmap + MADV_HUGEPAGE + touch + cudaHostRegister.  (Design evidence for csrc/hostio.cu.)
//   nvcc -O2 -Xcompiler -pthread benchmarks/pin_rate_bench.cu -o /tmp/pin_rate && /tmp/pin_rate
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
    cudaFree(0);
    const size_t slab = 64u << 20, total = size_t(4) << 30;
    const int n = int(total / slab);
    for (int threads : {1, 4, 8}) {
        std::vector<void*> p(n, nullptr);
        double t0 = now();
        std::vector<std::thread> ts;
        for (int t = 0; t < threads; ++t)
            ts.emplace_back([&, t] { for (int i = t; i < n; i += threads) cudaHostAlloc(&p[i], slab, cudaHostAllocPortable); });
        for (auto& th : ts) th.join();
        double dt = now() - t0;
        printf("cudaHostAlloc 64 MB slabs, %d thread(s): %.2f GB/s\n", threads, total / dt / 1e9);
        for (void* q : p) cudaFreeHost(q);
    }
    for (size_t big : {size_t(1) << 30}) {
        void* q;
        double t0 = now();
        for (int i = 0; i < 4; ++i) { cudaHostAlloc(&q, big, cudaHostAllocPortable); cudaFreeHost(q); }
        printf("cudaHostAlloc 1 GB blocks: %.2f GB/s (alloc + free)\n", 4.0 * big / (now() - t0) / 1e9);
    }
    for (int huge : {0, 1}) {
        for (int threads : {1, 4}) {
            std::vector<void*> p(n, nullptr);
            double t0 = now();
            std::vector<std::thread> ts;
            for (int t = 0; t < threads; ++t)
                ts.emplace_back([&, t] {
                    for (int i = t; i < n; i += threads) {
                        void* m = mmap(nullptr, slab, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
                        if (huge) madvise(m, slab, MADV_HUGEPAGE);
                        for (size_t o = 0; o < slab; o += 4096) ((volatile char*)m)[o] = 0;
                        cudaHostRegister(m, slab, cudaHostRegisterPortable);
                        p[i] = m;
                    }
                });
            for (auto& th : ts) th.join();
            double dt = now() - t0;
            printf("mmap%s + touch + cudaHostRegister, %d thread(s): %.2f GB/s\n", huge ? " + MADV_HUGEPAGE" : "", threads, total / dt / 1e9);
            for (void* q : p) { cudaHostUnregister(q); munmap(q, slab); }
        }
    }
    return 0;
}
