// Host-memory copy bandwidth on the GPU box: how fast can T threads move pageable field arrays
// into a staging area?  (Design evidence for csrc/hostio.cu; not part of the library.)
//   g++ -O2 -mavx2 -pthread benchmarks/hostcopy_bench.cpp -o /tmp/hostcopy_bench && /tmp/hostcopy_bench
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void copy_nt(char* dst, const char* src, size_t n) {
    for (size_t i = 0; i + 128 <= n; i += 128) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a); _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c); _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    _mm_sfence();
}
static void read_only(const char* src, size_t n, __m256i* sink) {
    __m256i acc = _mm256_setzero_si256();
    for (size_t i = 0; i + 128 <= n; i += 128) {
        acc = _mm256_xor_si256(acc, _mm256_loadu_si256((const __m256i*)(src + i)));
        acc = _mm256_xor_si256(acc, _mm256_loadu_si256((const __m256i*)(src + i + 32)));
        acc = _mm256_xor_si256(acc, _mm256_loadu_si256((const __m256i*)(src + i + 64)));
        acc = _mm256_xor_si256(acc, _mm256_loadu_si256((const __m256i*)(src + i + 96)));
    }
    *sink = acc;
}

int main(int argc, char** argv) {
    const size_t field = 4152960, n_fields = argc > 1 ? atoi(argv[1]) : 1024;
    std::vector<char*> src(n_fields);
    for (auto& p : src) { p = (char*)malloc(field); memset(p, 1, field); }
    char* big = (char*)aligned_alloc(4096, field * 256);  // DRAM-sized staging (1 GB)
    memset(big, 0, field * 256);
    for (int threads : {1, 2, 4, 8, 12, 16}) {
        for (int mode = 0; mode < 4; ++mode) {  // 0 memcpy->big, 1 NT->big, 2 regular->small per-thread ring (cache resident), 3 read only
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> ts;
            for (int t = 0; t < threads; ++t)
                ts.emplace_back([&, t] {
                    char* ring = (char*)aligned_alloc(4096, 2 << 20);
                    __m256i sink;
                    long long total = 0;
                    for (size_t f = t; f < n_fields; f += threads) {
                        char* dst = big + (f % 256) * field;
                        if (mode == 0) memcpy(dst, src[f], field);
                        else if (mode == 1) copy_nt(dst, src[f], field);
                        else if (mode == 2) for (size_t o = 0; o < field; o += 1 << 20) memcpy(ring + ((o >> 20) & 1) * (1 << 20), src[f] + o, std::min<size_t>(1 << 20, field - o));
                        else { read_only(src[f], field, &sink); total += _mm256_extract_epi64(sink, 0); }
                    }
                    free(ring);
                    if (total == 42) printf("!");
                });
            for (auto& th : ts) th.join();
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("threads %2d mode %s: %.1f GB/s\n", threads, mode == 0 ? "memcpy->dram " : mode == 1 ? "nt->dram     " : mode == 2 ? "memcpy->cache" : "read only    ", n_fields * field / s / 1e9);
        }
    }
    return 0;
}
