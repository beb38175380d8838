// grib.cuh — internal interface between grib.cu (parser + unpack kernel) and hostio.cu (the
// streamed upload of packed messages).
#pragma once

#include <cstdint>

#include "common.cuh"

namespace at {

// One field of a packed chunk as the kernel sees it (device table, 40 bytes).
struct GribColumn {
    long long byte_offset;  // of the field's packed values inside the chunk's device buffer
    double reference;       // R
    double binary;          // 2^E
    double decimal;         // 10^-D
    int nbits;
    int reserved;
};

// Validate one scanned message against the batch (no bitmap, value count, section length) and
// fill its kernel parameters.
int grib_column_of(const at_grib_field_t& info, int64_t n_points, int64_t byte_offset, GribColumn* out);

// out[p, f] = value p of field f for f < n_fields (out_dtype AT_F32 | AT_F64, leading dimension ld).
int grib_unpack_launch(const uint8_t* d_packed, const GribColumn* d_cols, int n_fields, int64_t n_points, int out_dtype, void* d_pm,
                       int64_t ld, cudaStream_t st);

}  // namespace at
