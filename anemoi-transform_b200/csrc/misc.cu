// misc.cu — error plumbing, device queries, host-memory pinning.
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

}  // namespace at

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
