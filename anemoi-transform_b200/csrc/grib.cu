// grib.cu — GRIB simple packing decoded on the device, straight into the point-major batch.
//
// Where this sits: a GRIB-backed FieldList reaches RegridFilter.forward as earthkit-data
// GribFields, and the per-field `field.to_numpy(flatten=True)` of the reference
// (src/anemoi/transform/filters/fields/regrid.py:309, matching.py:242-246) is where ecCodes
// decodes each message to float64 on one CPU core.  Here the packed octets (2 bytes per point
// at 16 bits, instead of 8 bytes of decoded float64) cross PCIe and the kernel below unpacks
// them into columns of the [points x fields] batch the SpMM reads.
//
//   at_grib_scan      host: locate packing parameters, the packed values and the bitmap of one
//                     message (editions 1 and 2, grid-point simple packing)
//   at_grib_unpack    device: Y = ((X · 2^E) + R) · 10^-D for a batch of fields, fused with the
//                     field-major -> point-major transposition
//
// Arithmetic: float64, one multiplication, one addition, one multiplication, never fused —
// the order ecCodes' data_simple_packing uses (restated in oracle/grib.py; parity with ecCodes
// itself is unpinned offline).  For D = 0 the result is the correctly rounded
// R + X·2^E whatever the order.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "grib.cuh"

namespace at {

namespace {

// ---- host: message scan ----------------------------------------------------------------
inline uint64_t be(const uint8_t* p, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v = (v << 8) | p[i];
    return v;
}

inline int sign_magnitude16(const uint8_t* p) {
    const int u = static_cast<int>(be(p, 2));
    return (u & 0x8000) ? -(u & 0x7fff) : (u & 0x7fff);
}

inline double ibm32(const uint8_t* p) {
    const uint32_t u = static_cast<uint32_t>(be(p, 4));
    const uint32_t mantissa = u & 0xffffffu;
    const int exponent = static_cast<int>((u >> 24) & 0x7f);
    if (mantissa == 0) return 0.0;
    const double v = std::ldexp(static_cast<double>(mantissa), 4 * (exponent - 64 - 6));  // 16^(e-70), exact
    return (u >> 31) ? -v : v;
}

inline double ieee32(const uint8_t* p) {
    const uint32_t u = static_cast<uint32_t>(be(p, 4));
    float f;
    std::memcpy(&f, &u, 4);
    return static_cast<double>(f);
}

#define GRIB_NEED(cond, ...)                                             \
    do {                                                                 \
        if (!(cond)) return set_error(AT_ERR_INVALID, __VA_ARGS__);      \
    } while (0)
#define GRIB_UNSUPPORTED(...) return set_error(AT_ERR_UNSUPPORTED, __VA_ARGS__)

int scan_edition2(const uint8_t* m, size_t len, at_grib_field_t* o) {
    GRIB_NEED(len >= 20, "at_grib_scan: truncated edition-2 message");
    const uint64_t total = be(m + 8, 8);
    GRIB_NEED(total <= len && total >= 20, "at_grib_scan: message length %llu exceeds the buffer (%zu)", (unsigned long long)total, len);
    GRIB_NEED(std::memcmp(m + total - 4, "7777", 4) == 0, "at_grib_scan: end section missing");
    o->edition = 2;
    o->message_length = static_cast<int64_t>(total);
    size_t pos = 16;
    int seen5 = 0, seen7 = 0, seen3 = 0;
    while (pos + 5 <= total - 4) {
        const uint64_t sl = be(m + pos, 4);
        const int number = m[pos + 4];
        GRIB_NEED(sl >= 5 && pos + sl <= total - 4, "at_grib_scan: section %d overruns the message", number);
        switch (number) {
            case 3:
                GRIB_NEED(sl >= 14, "at_grib_scan: section 3 too short");
                o->n_points = static_cast<int64_t>(be(m + pos + 6, 4));
                ++seen3;
                break;
            case 5: {
                GRIB_NEED(sl >= 11, "at_grib_scan: section 5 too short");
                o->n_values = static_cast<int64_t>(be(m + pos + 5, 4));
                const int tmpl = static_cast<int>(be(m + pos + 9, 2));
                if (tmpl != 0) GRIB_UNSUPPORTED("at_grib_scan: data representation template 5.%d (only 5.0, simple packing)", tmpl);
                GRIB_NEED(sl >= 20, "at_grib_scan: template 5.0 too short");
                o->reference_value = ieee32(m + pos + 11);
                o->binary_scale = sign_magnitude16(m + pos + 15);
                o->decimal_scale = sign_magnitude16(m + pos + 17);
                o->bits_per_value = m[pos + 19];
                ++seen5;
                break;
            }
            case 6: {
                GRIB_NEED(sl >= 6, "at_grib_scan: section 6 too short");
                const int indicator = m[pos + 5];
                if (indicator == 0) {
                    o->has_bitmap = 1;
                    o->bitmap_offset = static_cast<int64_t>(pos + 6);
                } else if (indicator != 255) {
                    GRIB_UNSUPPORTED("at_grib_scan: bitmap indicator %d", indicator);
                }
                break;
            }
            case 7:
                o->data_offset = static_cast<int64_t>(pos + 5);
                o->data_length = static_cast<int64_t>(sl - 5);
                ++seen7;
                break;
            default:
                break;
        }
        pos += sl;
    }
    if (seen7 != 1 || seen5 != 1 || seen3 != 1) GRIB_UNSUPPORTED("at_grib_scan: expected one field per message (sections 3/5/7 seen %d/%d/%d times)", seen3, seen5, seen7);
    return AT_OK;
}

int scan_edition1(const uint8_t* m, size_t len, at_grib_field_t* o) {
    const uint32_t len3 = static_cast<uint32_t>(be(m + 4, 3));
    size_t pos = 8;
    GRIB_NEED(len >= pos + 28, "at_grib_scan: truncated edition-1 message");
    const size_t pds_len = be(m + pos, 3);
    GRIB_NEED(pds_len >= 28 && pos + pds_len <= len, "at_grib_scan: bad PDS length");
    const int flag = m[pos + 7];
    o->edition = 1;
    o->decimal_scale = sign_magnitude16(m + pos + 26);
    pos += pds_len;
    if (flag & 0x80) {
        GRIB_NEED(pos + 3 <= len, "at_grib_scan: truncated GDS");
        pos += be(m + pos, 3);
    }
    int64_t n_points = -1;
    if (flag & 0x40) {
        GRIB_NEED(pos + 6 <= len, "at_grib_scan: truncated BMS");
        const size_t bms_len = be(m + pos, 3);
        if (be(m + pos + 4, 2) != 0) GRIB_UNSUPPORTED("at_grib_scan: predefined bitmaps");
        o->has_bitmap = 1;
        o->bitmap_offset = static_cast<int64_t>(pos + 6);
        n_points = static_cast<int64_t>(bms_len - 6) * 8 - m[pos + 3];
        pos += bms_len;
    }
    GRIB_NEED(pos + 11 <= len, "at_grib_scan: truncated BDS");
    const uint64_t stored4 = be(m + pos, 3);
    uint64_t total, bds_len;
    const bool long_message = (len3 & 0x800000u) != 0;
    if (long_message) {  // ECMWF convention: units of 120 octets, stored section-4 length subtracted
        total = static_cast<uint64_t>(len3 & 0x7fffffu) * 120u - stored4;
        GRIB_NEED(total > pos + 15, "at_grib_scan: bad long-message length");
        bds_len = total - 4 - pos;
    } else {
        total = len3;
        bds_len = stored4;
    }
    GRIB_NEED(total <= len && bds_len >= 11 && pos + bds_len + 4 <= total, "at_grib_scan: message length %llu exceeds the buffer (%zu)", (unsigned long long)total, len);
    GRIB_NEED(std::memcmp(m + total - 4, "7777", 4) == 0, "at_grib_scan: end section missing");
    const int flags4 = m[pos + 3];
    if (flags4 & 0xf0) GRIB_UNSUPPORTED("at_grib_scan: BDS flags 0x%02x (only grid-point simple packing of float values)", flags4 & 0xf0);
    o->binary_scale = sign_magnitude16(m + pos + 4);
    o->reference_value = ibm32(m + pos + 6);
    o->bits_per_value = m[pos + 10];
    o->data_offset = static_cast<int64_t>(pos + 11);
    o->data_length = static_cast<int64_t>(bds_len - 11);
    // the value count is implied by the section length; a long message pads without recording it
    o->n_values = (o->bits_per_value > 0 && !long_message) ? (o->data_length * 8 - (flags4 & 0x0f)) / o->bits_per_value : -1;
    o->n_points = n_points >= 0 ? n_points : o->n_values;
    o->message_length = static_cast<int64_t>(total);
    return AT_OK;
}

// n^s by repeated multiplication / division from 1.0 (how ecCodes forms 2^E and 10^-D).
double repeated_power(long s, long n) {
    double v = 1.0;
    if (s == 0) return 1.0;
    if (s == 1) return static_cast<double>(n);
    while (s < 0) {
        v /= static_cast<double>(n);
        ++s;
    }
    while (s > 0) {
        v *= static_cast<double>(n);
        --s;
    }
    return v;
}

}  // namespace

// ---- device: unpack + transpose ---------------------------------------------------------
constexpr int kTileFields = 32;   // columns of a CTA tile (one warp-wide row segment)
constexpr int kTilePoints = 256;  // points of a CTA tile: 32 lanes x 8 values

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// The 8 values of point group g (points 8g .. 8g+7) of one field start on an octet boundary:
// 8·nbits bits = nbits octets.  `avail` = values of the group that exist (8 except at the end).
template <typename TOut>
__device__ __forceinline__ void unpack_group(const uint8_t* __restrict__ src, const GribColumn& c, int avail, TOut* __restrict__ dst) {
    const int nbits = c.nbits;
    if (nbits == 0) {
#pragma unroll
        for (int v = 0; v < 8; ++v)
            if (v < avail) dst[v] = static_cast<TOut>(c.reference);
        return;
    }
    uint32_t x[8];
    if (nbits == 16 && avail == 8 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(src));
        const uint32_t ws[4] = {bswap32(w.x), bswap32(w.y), bswap32(w.z), bswap32(w.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[2 * j] = ws[j] >> 16;
            x[2 * j + 1] = ws[j] & 0xffffu;
        }
    } else if ((nbits & 3) == 0 && avail == 8 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
        // whole 32-bit words: nbits / 4 of them hold the 8 values
        const uint32_t* w = reinterpret_cast<const uint32_t*>(src);
        const int n_words = nbits >> 2;
        unsigned long long acc = 0;
        int have = 0, next = 0;
        const uint32_t mask = nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1u);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            if (have < nbits) {
                acc = (acc << 32) | (next < n_words ? bswap32(__ldg(w + next)) : 0u);
                ++next;
                have += 32;
            }
            x[v] = static_cast<uint32_t>(acc >> (have - nbits)) & mask;
            have -= nbits;
        }
    } else {
        // any width, any alignment, partial last group: octet by octet
        const int n_octets = (avail * nbits + 7) >> 3;
        unsigned long long acc = 0;
        int have = 0, next = 0;
        const uint32_t mask = nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1u);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            x[v] = 0;
            if (v < avail) {
                while (have < nbits) {
                    acc = (acc << 8) | (next < n_octets ? __ldg(src + next) : 0u);
                    ++next;
                    have += 8;
                }
                x[v] = static_cast<uint32_t>(acc >> (have - nbits)) & mask;
                have -= nbits;
            }
        }
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        if (v < avail) {
            const double y = __dmul_rn(__dadd_rn(__dmul_rn(static_cast<double>(x[v]), c.binary), c.reference), c.decimal);
            dst[v] = static_cast<TOut>(y);
        }
    }
}

// ---- bitmaps: points without a value ------------------------------------------------------
// A field with a bitmap packs only the points whose bit is set; point p's value is number
// rank(p) = (set bits before p) of the stream.  grib_bitmap_rank_kernel writes, per field and per
// tile of 256 points, the rank of the tile's first point (one warp per field, 32 tiles per step,
// warp scan with carry); the unpack kernel adds the set bits of the lanes before it inside the
// tile.  A missing point decodes to NaN (what earthkit-data's to_numpy() hands to the filters).
__global__ void __launch_bounds__(32)
    grib_bitmap_rank_kernel(uint8_t* __restrict__ packed, const GribColumn* __restrict__ cols, long long n_points) {
    const GribColumn c = cols[blockIdx.x];
    if (c.bitmap_offset < 0) return;
    const int lane = threadIdx.x;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(packed + c.bitmap_offset);  // 256-byte aligned
    uint32_t* rank = reinterpret_cast<uint32_t*>(packed + c.rank_offset);
    const long long n_tiles = (n_points + kTilePoints - 1) / kTilePoints;
    uint32_t carry = 0;
    for (long long t0 = 0; t0 < n_tiles; t0 += 32) {
        const long long t = t0 + lane;
        uint32_t count = 0;
        if (t < n_tiles) {
            const long long first = t * kTilePoints;
            const int valid = static_cast<int>(min(static_cast<long long>(kTilePoints), n_points - first));
#pragma unroll
            for (int w = 0; w < kTilePoints / 32; ++w) {
                const int v = valid - 32 * w;  // points of this word that exist
                if (v <= 0) break;
                uint32_t word = words[t * (kTilePoints / 32) + w];
                if (v < 32) {
                    // point i of the word is bit 7 - i % 8 of octet i / 8; octet j is bits 8j .. 8j+7 of the little-endian word
                    const int whole = v >> 3, rest = v & 7;
                    uint32_t mask = whole >= 4 ? 0xffffffffu : ((1u << (8 * whole)) - 1u);
                    if (rest != 0) mask |= ((0xffu << (8 - rest)) & 0xffu) << (8 * whole);
                    word &= mask;
                }
                count += __popc(word);
            }
        }
        uint32_t inc = count;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (t < n_tiles) rank[t] = carry + inc - count;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// Value number `index` of a stream of nbits-wide big-endian fields that starts at `values`.
__device__ __forceinline__ uint32_t extract_bits(const uint8_t* __restrict__ values, long long index, int nbits) {
    const long long bit = index * nbits;
    const uint8_t* b = values + (bit >> 3);
    const int shift = static_cast<int>(bit & 7);
    // shift + nbits <= 39 bits: five octets (the regions carry 16 octets of slack)
    unsigned long long acc = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) acc = (acc << 8) | __ldg(b + i);
    const uint32_t mask = nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1u);
    return static_cast<uint32_t>(acc >> (40 - shift - nbits)) & mask;
}

template <typename TOut>
__device__ __forceinline__ void unpack_group_bitmap(const uint8_t* __restrict__ packed, const GribColumn& c, long long p0,
                                                    long long n_points, int lane, int avail, TOut* __restrict__ dst) {
    // this lane's octet of the bitmap: points p0 + 8·lane .. + 7, MSB first
    uint32_t bits = 0;
    if (avail > 0) {
        bits = __ldg(packed + c.bitmap_offset + ((p0 >> 3) + lane));
        if (avail < 8) bits &= (0xffu << (8 - avail)) & 0xffu;
    }
    uint32_t inc = __popc(bits);
    const uint32_t mine = inc;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    const uint32_t* rank = reinterpret_cast<const uint32_t*>(packed + c.rank_offset);
    long long index = static_cast<long long>(__ldg(rank + p0 / kTilePoints)) + (inc - mine);
    const uint8_t* values = packed + c.byte_offset;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k >= avail) break;
        if (bits & (0x80u >> k)) {
            const uint32_t x = c.nbits == 0 ? 0u : extract_bits(values, index, c.nbits);
            ++index;
            const double y = c.nbits == 0 ? c.reference : __dmul_rn(__dadd_rn(__dmul_rn(static_cast<double>(x), c.binary), c.reference), c.decimal);
            dst[k] = static_cast<TOut>(y);
        } else {
            dst[k] = static_cast<TOut>(__longlong_as_double(0x7ff8000000000000ll));
        }
    }
}

// One CTA per (tile of 256 points, group of 32 fields).  Warp w unpacks fields w, w+8, w+16,
// w+24 of the group (every lane 8 consecutive points: coalesced reads of nbits octets per lane)
// into shared memory; then every warp writes rows of 32 adjacent columns of the batch.
template <typename TOut, bool BITMAPS>
__global__ void __launch_bounds__(256)
    grib_unpack_kernel(const uint8_t* __restrict__ packed, const GribColumn* __restrict__ cols, int n_fields, long long n_points,
                       TOut* __restrict__ out, long long ld, unsigned n_field_groups) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TOut* tile = reinterpret_cast<TOut*>(smem_raw);  // [kTileFields][kTilePoints + 1]
    constexpr int kPitch = kTilePoints + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // field groups vary fastest: CTAs that run together write adjacent column segments of the
    // same rows, so whole rows of the batch reach DRAM together
    const long long p0 = static_cast<long long>(blockIdx.x / n_field_groups) * kTilePoints;
    const int f0 = static_cast<int>(blockIdx.x % n_field_groups) * kTileFields;
    const long long p = p0 + lane * 8;
    const int avail = static_cast<int>(max(0ll, min(8ll, n_points - p)));
#pragma unroll 1
    for (int j = warp; j < kTileFields; j += 8) {
        const int f = f0 + j;
        if (f >= n_fields) continue;  // warp-uniform
        const GribColumn c = cols[f];
        TOut v[8];
        if (BITMAPS && c.bitmap_offset >= 0) {  // warp-uniform: the whole warp takes part in the scan
            unpack_group_bitmap<TOut>(packed, c, p0, n_points, lane, avail, v);
        } else {
            if (avail <= 0) continue;
            unpack_group<TOut>(packed + c.byte_offset + (p >> 3) * c.nbits, c, avail, v);
        }
        // point 8·lane + k of the tile is kept at position 32·k + lane: consecutive lanes write
        // consecutive words (position 8·lane + k put 32 lanes on 2 banks: ncu counted 121 M
        // bank conflicts per launch and the kernel sat at 0.52 of the HBM peak)
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < avail) tile[j * kPitch + k * 32 + lane] = v[k];
    }
    __syncthreads();
    const int f = f0 + lane;
    const int rows = static_cast<int>(min(static_cast<long long>(kTilePoints), n_points - p0));
    if (f < n_fields)
        for (int r = warp; r < rows; r += 8) out[(p0 + r) * ld + f] = tile[lane * kPitch + (r & 7) * 32 + (r >> 3)];
}

template <typename TOut, bool BITMAPS>
static int launch_unpack(const uint8_t* d_packed, const GribColumn* d_cols, int n_fields, int64_t n_points, TOut* out, int64_t ld,
                         unsigned grid, unsigned n_field_groups, cudaStream_t st) {
    const size_t smem = sizeof(TOut) * kTileFields * (kTilePoints + 1);
    if (smem > 48 * 1024) {  // above the default: opt in (idempotent)
        AT_CUDA_TRY(cudaFuncSetAttribute(grib_unpack_kernel<TOut, BITMAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    }
    grib_unpack_kernel<TOut, BITMAPS><<<grid, 256, smem, st>>>(d_packed, d_cols, n_fields, n_points, out, ld, n_field_groups);
    AT_LAUNCH_CHECK("grib_unpack_kernel");
    return AT_OK;
}

int grib_unpack_launch(uint8_t* d_packed, const GribColumn* d_cols, int n_fields, int64_t n_points, int out_dtype, void* d_pm,
                       int64_t ld, bool any_bitmap, cudaStream_t st) {
    if (n_fields == 0 || n_points == 0) return AT_OK;
    const unsigned n_field_groups = static_cast<unsigned>((n_fields + kTileFields - 1) / kTileFields);
    const long long blocks = ((n_points + kTilePoints - 1) / kTilePoints) * n_field_groups;
    if (blocks >= (1ll << 31)) return set_error(AT_ERR_UNSUPPORTED, "GRIB unpack: batch too large for one launch");
    const unsigned grid = static_cast<unsigned>(blocks);
    if (any_bitmap) {
        grib_bitmap_rank_kernel<<<static_cast<unsigned>(n_fields), 32, 0, st>>>(d_packed, d_cols, n_points);
        AT_LAUNCH_CHECK("grib_bitmap_rank_kernel");
        if (out_dtype == AT_F32) return launch_unpack<float, true>(d_packed, d_cols, n_fields, n_points, static_cast<float*>(d_pm), ld, grid, n_field_groups, st);
        return launch_unpack<double, true>(d_packed, d_cols, n_fields, n_points, static_cast<double*>(d_pm), ld, grid, n_field_groups, st);
    }
    if (out_dtype == AT_F32) return launch_unpack<float, false>(d_packed, d_cols, n_fields, n_points, static_cast<float*>(d_pm), ld, grid, n_field_groups, st);
    return launch_unpack<double, false>(d_packed, d_cols, n_fields, n_points, static_cast<double*>(d_pm), ld, grid, n_field_groups, st);
}

static size_t round_up_256(size_t n) { return (n + 255) / 256 * 256; }

GribFootprint grib_footprint(const at_grib_field_t& info, int64_t n_points) {
    GribFootprint fp;
    const int64_t n_values = info.has_bitmap ? (info.n_values >= 0 ? info.n_values : n_points) : n_points;
    int64_t octets = (n_values * std::max(info.bits_per_value, 0) + 7) / 8;
    if (info.has_bitmap && info.n_values < 0) octets = std::max<int64_t>(info.data_length, 0);  // count unknown: the whole section
    fp.value_octets = static_cast<size_t>(std::min<int64_t>(octets, std::max<int64_t>(info.data_length, 0)));
    fp.values = round_up_256(fp.value_octets + 16);
    if (info.has_bitmap) {
        fp.bitmap_octets = static_cast<size_t>((n_points + 7) / 8);
        fp.bitmap = round_up_256(fp.bitmap_octets + 32);
        fp.ranks = round_up_256(static_cast<size_t>((n_points + kTilePoints - 1) / kTilePoints + 1) * sizeof(uint32_t));
    }
    return fp;
}

int grib_column_of(const at_grib_field_t& info, int64_t n_points, int64_t byte_offset, GribColumn* out) {
    AT_REQUIRE(info.bits_per_value >= 0 && info.bits_per_value <= 32, "GRIB unpack: bitsPerValue %d out of range", info.bits_per_value);
    const GribFootprint fp = grib_footprint(info, n_points);
    if (info.has_bitmap) {
        AT_REQUIRE(info.n_points < 0 || info.n_points == n_points, "GRIB unpack: the bitmap covers %lld points, the batch has %lld",
                   (long long)info.n_points, (long long)n_points);
        AT_REQUIRE(info.n_values <= n_points, "GRIB unpack: %lld values for %lld points", (long long)info.n_values, (long long)n_points);
        AT_REQUIRE(info.bitmap_offset >= 0 && info.bitmap_offset + static_cast<int64_t>(fp.bitmap_octets) <= info.message_length,
                   "GRIB unpack: the bitmap runs past the end of the message");
        if (info.n_values >= 0)
            AT_REQUIRE(info.data_length >= (info.n_values * info.bits_per_value + 7) / 8, "GRIB unpack: %lld octets of packed values, %lld needed",
                       (long long)info.data_length, (long long)((info.n_values * info.bits_per_value + 7) / 8));
    } else {
        AT_REQUIRE(info.n_values < 0 || info.n_values == n_points, "GRIB unpack: message holds %lld values, the batch has %lld points",
                   (long long)info.n_values, (long long)n_points);
        const int64_t need = (n_points * info.bits_per_value + 7) / 8;
        AT_REQUIRE(info.data_length >= need, "GRIB unpack: %lld octets of packed values, %lld needed", (long long)info.data_length, (long long)need);
    }
    out->byte_offset = byte_offset;
    out->bitmap_offset = info.has_bitmap ? byte_offset + static_cast<int64_t>(fp.values) : -1;
    out->rank_offset = info.has_bitmap ? byte_offset + static_cast<int64_t>(fp.values + fp.bitmap) : -1;
    out->reference = info.reference_value;
    out->binary = repeated_power(info.binary_scale, 2);
    out->decimal = repeated_power(-info.decimal_scale, 10);
    out->nbits = info.bits_per_value;
    out->reserved = 0;
    return AT_OK;
}

}  // namespace at

using namespace at;

extern "C" int at_grib_scan(const void* message, size_t length, at_grib_field_t* out) {
    AT_REQUIRE(message != nullptr && out != nullptr, "at_grib_scan: null argument");
    std::memset(out, 0, sizeof(*out));
    out->bitmap_offset = -1;
    out->n_points = out->n_values = -1;
    const uint8_t* m = static_cast<const uint8_t*>(message);
    AT_REQUIRE(length >= 8 && std::memcmp(m, "GRIB", 4) == 0, "at_grib_scan: not a GRIB message");
    const int edition = m[7];
    if (edition == 2) return scan_edition2(m, length, out);
    if (edition == 1) return scan_edition1(m, length, out);
    return set_error(AT_ERR_UNSUPPORTED, "at_grib_scan: GRIB edition %d", edition);
}

extern "C" int at_grib_scan_many(const void* const* messages, const size_t* lengths, int64_t n, at_grib_field_t* out, int32_t* status) {
    AT_REQUIRE(n >= 0, "at_grib_scan_many: negative count");
    if (n == 0) return AT_OK;
    AT_REQUIRE(messages != nullptr && lengths != nullptr && out != nullptr && status != nullptr, "at_grib_scan_many: null argument");
    for (int64_t i = 0; i < n; ++i) status[i] = messages[i] != nullptr ? at_grib_scan(messages[i], lengths[i], out + i) : AT_ERR_INVALID;
    return AT_OK;
}

extern "C" int at_grib_unpack(const void* d_packed, const int64_t* byte_offsets, const at_grib_field_t* fields, int64_t n_fields,
                              int64_t n_points, int out_dtype, void* d_pm, int64_t ld, void* stream) {
    AT_REQUIRE(n_fields >= 0 && n_points >= 0, "at_grib_unpack: negative size");
    AT_REQUIRE(out_dtype == AT_F32 || out_dtype == AT_F64, "at_grib_unpack: bad dtype code");
    if (n_fields == 0 || n_points == 0) return AT_OK;
    AT_REQUIRE(d_packed != nullptr && byte_offsets != nullptr && fields != nullptr && d_pm != nullptr, "at_grib_unpack: null argument");
    AT_REQUIRE(ld >= n_fields && n_fields < (1ll << 31), "at_grib_unpack: bad leading dimension");
    std::vector<GribColumn> cols(static_cast<size_t>(n_fields));
    for (int64_t f = 0; f < n_fields; ++f) {
        AT_REQUIRE(byte_offsets[f] >= 0, "at_grib_unpack: negative offset");
        AT_REQUIRE(fields[f].has_bitmap == 0, "at_grib_unpack: messages with a bitmap go through at_hostio_upload_grib / at_hostio_regrid_grib (they stage the bitmap too)");
        const int rc = grib_column_of(fields[f], n_points, byte_offsets[f], &cols[static_cast<size_t>(f)]);
        if (rc != AT_OK) return rc;
    }
    GribColumn* d_cols = nullptr;
    AT_CUDA_TRY(device_alloc(reinterpret_cast<void**>(&d_cols), cols.size() * sizeof(GribColumn)));
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemcpyAsync(d_cols, cols.data(), cols.size() * sizeof(GribColumn), cudaMemcpyHostToDevice, st);
    int rc = AT_OK;
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // `cols` is pageable and leaves scope
    if (e != cudaSuccess) rc = set_error(AT_ERR_CUDA, "at_grib_unpack: %s", cudaGetErrorString(e));
    if (rc == AT_OK)
        rc = grib_unpack_launch(const_cast<uint8_t*>(static_cast<const uint8_t*>(d_packed)), d_cols, static_cast<int>(n_fields), n_points, out_dtype, d_pm, ld, false, st);
    device_free(d_cols);  // waits for the device, like cudaFree
    return rc;
}
