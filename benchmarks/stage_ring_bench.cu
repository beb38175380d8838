// Design experiment for csrc/hostio.cu: does a small, cache-resident staging ring (regular
// stores; the DMA engine reads the lines out of the CPU caches) move pageable fields to the
// GPU faster than large DRAM-resident slots written with non-temporal stores?
//   nvcc -O2 -Xcompiler -mavx2,-pthread benchmarks/stage_ring_bench.cu -o /tmp/stage_ring && /tmp/stage_ring
#include <cuda_runtime.h>
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void copy_nt(char* dst, const char* src, size_t n) {
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a); _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c); _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
    _mm_sfence();
}
static void copy_regular(char* dst, const char* src, size_t n) {
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_store_si256((__m256i*)(dst + i), a); _mm256_store_si256((__m256i*)(dst + i + 32), b);
        _mm256_store_si256((__m256i*)(dst + i + 64), c); _mm256_store_si256((__m256i*)(dst + i + 96), d);
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
}

int main(int argc, char** argv) {
    const size_t field = 4152960;
    const int n_fields = argc > 1 ? atoi(argv[1]) : 1536;
    const bool duplex = argc > 2 && atoi(argv[2]) != 0;
    std::vector<char*> src(n_fields);
    for (auto& p : src) { p = (char*)malloc(field); memset(p, 1, field); }
    char* dev;
    cudaMalloc(&dev, field * 64);
    // optional D2H traffic in the other direction (pinned destination), as the regrid produces
    char *d2h_dev, *d2h_host;
    cudaMalloc(&d2h_dev, 256u << 20);
    cudaHostAlloc(&d2h_host, 256u << 20, cudaHostAllocPortable);
    cudaStream_t s_out;
    cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);

    for (int threads : {4, 8, 12}) {
        for (size_t piece : {size_t(1) << 20, size_t(2) << 20, field}) {
            for (int mode = 0; mode < 2; ++mode) {  // 0: per-thread ring of 2 pieces, regular stores; 1: same ring, NT stores
                for (int ring = 2; ring <= 3; ++ring) {
                    if (piece == field && ring == 3) continue;
                    std::atomic<bool> stop{false};
                    std::thread d2h;
                    if (duplex)
                        d2h = std::thread([&] {
                            while (!stop.load()) {
                                cudaMemcpyAsync(d2h_host, d2h_dev, 256u << 20, cudaMemcpyDeviceToHost, s_out);
                                cudaStreamSynchronize(s_out);
                            }
                        });
                    auto t0 = std::chrono::steady_clock::now();
                    std::vector<std::thread> ts;
                    for (int t = 0; t < threads; ++t)
                        ts.emplace_back([&, t] {
                            cudaStream_t st;
                            cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
                            std::vector<char*> buf(ring);
                            std::vector<cudaEvent_t> ev(ring);
                            for (int r = 0; r < ring; ++r) {
                                cudaHostAlloc(&buf[r], piece, cudaHostAllocPortable);
                                cudaEventCreateWithFlags(&ev[r], cudaEventDisableTiming);
                            }
                            size_t k = 0;
                            for (int f = t; f < n_fields; f += threads)
                                for (size_t off = 0; off < field; off += piece, ++k) {
                                    const size_t len = std::min(piece, field - off);
                                    const int r = k % ring;
                                    cudaEventSynchronize(ev[r]);
                                    if (mode == 0) copy_regular(buf[r], src[f] + off, len);
                                    else copy_nt(buf[r], src[f] + off, len);
                                    cudaMemcpyAsync(dev + (size_t)(f % 64) * field + off, buf[r], len, cudaMemcpyHostToDevice, st);
                                    cudaEventRecord(ev[r], st);
                                }
                            cudaStreamSynchronize(st);
                            for (int r = 0; r < ring; ++r) { cudaFreeHost(buf[r]); cudaEventDestroy(ev[r]); }
                            cudaStreamDestroy(st);
                        });
                    for (auto& th : ts) th.join();
                    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    stop.store(true);
                    if (duplex) d2h.join();
                    printf("threads %2d piece %4zu KB ring %d %s%s: %.1f GB/s H2D\n", threads, piece >> 10, ring, mode == 0 ? "regular" : "nt     ",
                           duplex ? " +d2h" : "", (double)n_fields * field / s / 1e9);
                    fflush(stdout);
                }
            }
        }
    }
    return 0;
}
