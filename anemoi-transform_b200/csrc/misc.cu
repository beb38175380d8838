// misc.cu — error plumbing, device queries, host-memory pinning.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace at {

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

namespace {

struct DeviceCache {
    std::mutex mu;
    std::unordered_map<void*, std::pair<size_t, int>> live;                  // block -> (size class, device)
    std::map<std::pair<int, size_t>, std::vector<void*>> free_blocks;        // (device, size class) -> blocks
    size_t cached_bytes = 0;
    size_t cap() {
        static const size_t c = [] {
            const char* e = std::getenv("AT_B200_DEVICE_CACHE_MB");
            return (e != nullptr && std::atoll(e) >= 0) ? static_cast<size_t>(std::atoll(e)) << 20 : (size_t(4) << 30);
        }();
        return c;
    }
    void trim_locked() {
        for (auto& kv : free_blocks) {
            for (void* p : kv.second) {
                cudaFree(p);
                live.erase(p);
            }
            kv.second.clear();
        }
        free_blocks.clear();
        cached_bytes = 0;
    }
};

DeviceCache& device_cache() {
    static DeviceCache* c = new DeviceCache();  // leaked on purpose: handles may be destroyed during exit
    return *c;
}

}  // namespace

cudaError_t device_alloc(void** p, size_t bytes) {
    const size_t cls = (std::max<size_t>(bytes, 16) + 511) / 512 * 512;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    DeviceCache& c = device_cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.free_blocks.find({dev, cls});
        if (it != c.free_blocks.end() && !it->second.empty()) {
            *p = it->second.back();
            it->second.pop_back();
            c.cached_bytes -= cls;
            return cudaSuccess;
        }
    }
    e = cudaMalloc(p, cls);
    if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lk(c.mu);
            c.trim_locked();
        }
        e = cudaMalloc(p, cls);
    }
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(c.mu);
    c.live[*p] = {cls, dev};
    return cudaSuccess;
}

void device_free(void* p) {
    if (p == nullptr) return;
    DeviceCache& c = device_cache();
    size_t cls = 0;
    int dev = 0;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.live.find(p);
        if (it == c.live.end()) {  // not ours (should not happen): let the driver decide
            cudaFree(p);
            return;
        }
        cls = it->second.first;
        dev = it->second.second;
    }
    // like cudaFree: kernels queued on any stream may still be using the block
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != dev) cudaSetDevice(dev);
    cudaDeviceSynchronize();
    if (cur != dev) cudaSetDevice(cur);
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.cached_bytes + cls > c.cap()) {
        cudaFree(p);
        c.live.erase(p);
        return;
    }
    c.free_blocks[{dev, cls}].push_back(p);
    c.cached_bytes += cls;
}

}  // namespace at

extern "C" int at_device_cache_trim(void) {
    std::lock_guard<std::mutex> lk(at::device_cache().mu);
    at::device_cache().trim_locked();
    return AT_OK;
}

extern "C" const char* at_last_error(void) { return at::error_buffer(); }

extern "C" int at_version(void) { return 100; }

extern "C" int at_device_count(int* count) {
    AT_REQUIRE(count != nullptr, "at_device_count: null argument");
    *count = 0;
    AT_CUDA_TRY(cudaGetDeviceCount(count));
    if (*count <= 0) return at::set_error(AT_ERR_CUDA, "no CUDA device is visible");
    return AT_OK;
}

extern "C" int at_set_device(int device) {
    AT_CUDA_TRY(cudaSetDevice(device));
    return AT_OK;
}

extern "C" int at_host_register(void* ptr, size_t bytes) {
    AT_REQUIRE(ptr != nullptr && bytes > 0, "at_host_register: empty range");
    AT_CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return AT_OK;
}

extern "C" int at_host_unregister(void* ptr) {
    AT_REQUIRE(ptr != nullptr, "at_host_unregister: null pointer");
    AT_CUDA_TRY(cudaHostUnregister(ptr));
    return AT_OK;
}
