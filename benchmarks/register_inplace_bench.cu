// Can the caller's pageable field arrays be page-locked in place instead of staged?
// malloc'd, touched 4 152 960-byte arrays (what numpy gives for one 0.25 degree float32 field):
//   (1) cudaHostRegister / cudaHostUnregister alone, T threads
//   (2) register -> cudaMemcpyAsync H2D -> unregister per array, T threads (whole upload leg)
//   (3) cudaMemcpyAsync straight from pageable memory, T threads (the driver's own staging)
// (Design evidence for csrc/hostio.cu.)
//   nvcc -O2 -Xcompiler -pthread benchmarks/register_inplace_bench.cu -o /tmp/reg_bench && /tmp/reg_bench
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class F>
static double run_threads(int threads, F f) {
    double t0 = now();
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; ++t) ts.emplace_back([&, t] { f(t); });
    for (auto& th : ts) th.join();
    return now() - t0;
}

int main() {
    cudaFree(0);
    const size_t bytes = 1038240u * 4u;
    const int n = 768;  // 3.2 GB
    const double total = double(bytes) * n;
    std::vector<char*> a(n);
    for (int i = 0; i < n; ++i) {
        a[i] = static_cast<char*>(malloc(bytes));
        memset(a[i], i & 0xff, bytes);
    }
    char* d = nullptr;
    cudaMalloc(&d, bytes * 16);

    for (unsigned flags : {unsigned(cudaHostRegisterPortable), unsigned(cudaHostRegisterPortable | cudaHostRegisterReadOnly)}) {
        for (int threads : {1, 2, 4, 8, 12}) {
            int bad = 0;
            double tr = run_threads(threads, [&](int t) {
                for (int i = t; i < n; i += threads)
                    if (cudaHostRegister(a[i], bytes, flags) != cudaSuccess) ++bad;
            });
            double tu = run_threads(threads, [&](int t) {
                for (int i = t; i < n; i += threads) cudaHostUnregister(a[i]);
            });
            cudaGetLastError();
            printf("flags %u, %2d thread(s): register %.2f GB/s (%.0f us each), unregister %.2f GB/s (%.0f us each), failures %d\n",
                   flags, threads, total / tr / 1e9, tr / n * threads * 1e6, total / tu / 1e9, tu / n * threads * 1e6, bad);
            fflush(stdout);
        }
    }
    for (int threads : {1, 2, 4, 8}) {
        std::vector<cudaStream_t> s(threads);
        for (auto& x : s) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
        double t = run_threads(threads, [&](int k) {
            for (int i = k; i < n; i += threads) {
                cudaHostRegister(a[i], bytes, cudaHostRegisterPortable);
                cudaMemcpyAsync(d + size_t(k) * bytes, a[i], bytes, cudaMemcpyHostToDevice, s[k]);
                cudaStreamSynchronize(s[k]);
                cudaHostUnregister(a[i]);
            }
        });
        printf("register -> H2D -> unregister, %d thread(s): %.2f GB/s\n", threads, total / t / 1e9);
        t = run_threads(threads, [&](int k) {
            for (int i = k; i < n; i += threads) {
                cudaMemcpyAsync(d + size_t(k) * bytes, a[i], bytes, cudaMemcpyHostToDevice, s[k]);
                cudaStreamSynchronize(s[k]);
            }
        });
        printf("cudaMemcpyAsync from pageable memory, %d thread(s): %.2f GB/s\n", threads, total / t / 1e9);
        fflush(stdout);
        for (auto& x : s) cudaStreamDestroy(x);
    }
    // one big registration over a contiguous range that holds many fields (an arena allocator's view)
    {
        const size_t big = size_t(1) << 30;
        char* m = static_cast<char*>(aligned_alloc(4096, big));
        memset(m, 1, big);
        double t0 = now();
        cudaError_t e = cudaHostRegister(m, big, cudaHostRegisterPortable);
        double t1 = now();
        cudaHostUnregister(m);
        double t2 = now();
        printf("1 GB contiguous touched range: register %.2f GB/s, unregister %.2f GB/s (%s)\n", big / (t1 - t0) / 1e9, big / (t2 - t1) / 1e9,
               cudaGetErrorString(e));
        free(m);
    }
    FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    if (f) {
        char buf[128] = {0};
        if (fgets(buf, sizeof buf, f)) printf("transparent_hugepage/enabled: %s", buf);
        fclose(f);
    }
    return 0;
}
