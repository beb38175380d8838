// grib.cuh — internal interface between grib.cu (parser + unpack kernel) and hostio.cu (the
// streamed upload of packed messages).
#pragma once

#include <cstdint>

#include "common.cuh"

namespace at {

// One field of a packed chunk as the kernel sees it (device table, 56 bytes).
struct GribColumn {
    long long byte_offset;    // of the field's packed values inside the chunk's device buffer
    long long bitmap_offset;  // of its bitmap (one bit per grid point, MSB first), -1 without one
    long long rank_offset;    // of its rank table (uint32 per 256 points: values before the tile), device scratch
    double reference;         // R
    double binary;            // 2^E
    double decimal;           // 10^-D
    int nbits;
    int reserved;
};

// Octets a field occupies in a chunk buffer: packed values, bitmap, rank table — each part
// starts on a 256-byte boundary; bitmap and ranks are 0 for a field without a bitmap.
struct GribFootprint {
    size_t values = 0, bitmap = 0, ranks = 0;
    size_t total() const { return values + bitmap + ranks; }
    size_t value_octets = 0, bitmap_octets = 0;  // what has to be copied from the message
};
GribFootprint grib_footprint(const at_grib_field_t& info, int64_t n_points);

// Validate one scanned message against the batch (value count, section lengths) and fill its
// kernel parameters for a field region that starts at `byte_offset` of the chunk buffer.
int grib_column_of(const at_grib_field_t& info, int64_t n_points, int64_t byte_offset, GribColumn* out);

// out[p, f] = value p of field f for f < n_fields (out_dtype AT_F32 | AT_F64, leading dimension ld).
// `any_bitmap`: some field has one (the rank tables are then built first, inside d_packed).
int grib_unpack_launch(uint8_t* d_packed, const GribColumn* d_cols, int n_fields, int64_t n_points, int out_dtype, void* d_pm,
                       int64_t ld, bool any_bitmap, cudaStream_t st);

}  // namespace at
