// matrix.cu — interpolation-matrix construction on the device, replacing the external `mir`
// binary of `anemoi-transform make-regrid-file` (commands/make-regrid-file.py:142-160) and
// earthkit-regrid's matrix inventory (filters/fields/regrid.py:211-259) for the case that can
// be built without either: 4-point bilinear weights from a regular lat-lon source grid.
//
// One thread per target point.  All arithmetic is float64 with separate multiplies and adds
// (the library is compiled with -fmad=false), in the order of the numpy restatement
// oracle/matrix.py::bilinear_matrix, so weights (rounded to float32 at the end) and column
// indices are bitwise those of the scipy-built matrix.
#include "common.cuh"

namespace at {

// numpy's float remainder (npy_divmod): fmod, then the sign of the divisor.
__device__ __forceinline__ double np_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0.0) != (m < 0.0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

__global__ void __launch_bounds__(256)
    bilinear_matrix_kernel(double lat0, double dlat, long long n_lat, double lon0, double dlon, long long n_lon,
                           const double* __restrict__ tlat, const double* __restrict__ tlon, long long n,
                           float* __restrict__ data, int* __restrict__ indices) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double fy = (tlat[t] - lat0) / dlat;
    long long j = static_cast<long long>(floor(fy));
    j = j < 0 ? 0 : (j > n_lat - 2 ? n_lat - 2 : j);
    const double wy = fy - static_cast<double>(j);
    const double fx = np_mod(tlon[t] - lon0, 360.0) / dlon;
    long long i0 = static_cast<long long>(floor(fx));
    const double wx = fx - static_cast<double>(i0);
    i0 %= n_lon;
    const long long i1 = (i0 + 1) % n_lon;
    long long c[4] = {j * n_lon + i0, j * n_lon + i1, (j + 1) * n_lon + i0, (j + 1) * n_lon + i1};
    double w[4] = {(1.0 - wy) * (1.0 - wx), (1.0 - wy) * wx, wy * (1.0 - wx), wy * wx};
    // stable insertion sort of the four entries by column (numpy argsort(kind="stable"))
#pragma unroll
    for (int a = 1; a < 4; ++a) {
#pragma unroll
        for (int b = a; b > 0; --b) {
            if (c[b - 1] > c[b]) {
                const long long tc = c[b];
                c[b] = c[b - 1];
                c[b - 1] = tc;
                const double tw = w[b];
                w[b] = w[b - 1];
                w[b - 1] = tw;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        data[4 * t + a] = static_cast<float>(w[a]);
        indices[4 * t + a] = static_cast<int>(c[a]);
    }
}

}  // namespace at

using namespace at;

extern "C" int at_bilinear_matrix(double lat0, double dlat, int64_t n_lat, double lon0, double dlon, int64_t n_lon,
                                  const double* tgt_lat, const double* tgt_lon, int64_t n_tgt, float* data_out,
                                  int32_t* indices_out, void* stream) {
    AT_REQUIRE(n_lat >= 2 && n_lon >= 1 && dlat != 0.0 && dlon > 0.0, "at_bilinear_matrix: bad source grid");
    AT_REQUIRE(n_lat * n_lon < (1ll << 31), "at_bilinear_matrix: source grid too large for int32 columns");
    AT_REQUIRE(n_tgt >= 0 && 4 * n_tgt < (1ll << 31), "at_bilinear_matrix: bad target count");
    if (n_tgt == 0) return AT_OK;
    AT_REQUIRE(tgt_lat != nullptr && tgt_lon != nullptr && data_out != nullptr && indices_out != nullptr,
               "at_bilinear_matrix: null argument");
    const int64_t blocks = (n_tgt + 255) / 256;
    bilinear_matrix_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(
        lat0, dlat, n_lat, lon0, dlon, n_lon, tgt_lat, tgt_lon, n_tgt, data_out, indices_out);
    AT_LAUNCH_CHECK("bilinear_matrix_kernel");
    return AT_OK;
}
