// knn.cu — bucketed nearest-neighbour search on float64 xyz, replacing scipy's cKDTree on
// the spatial path (reference src/anemoi/transform/spatial.py:96, 396, 501, 533, 628-632).
//
// Structure:
//   - a hierarchy of uniform 3-D grids; level l has cell edge h0·2^l and its cell
//     coordinates are the level-0 integer coordinates shifted right by l (exact nesting);
//   - each level is a counting-sort bucket table keyed by hash(cell) & (M_l - 1)
//     (no stored keys: colliding cells simply share a bucket — candidates are filtered by
//     their true distance, so collisions cost time, never correctness);
//   - a last pseudo-level has one bucket holding every point (brute force) so the search
//     always terminates, e.g. for queries far outside the source domain;
//   - one warp per query: 27 lanes look up the ring-1 buckets of the query's cell, duplicate
//     buckets are dropped with __match_any_sync, the candidate ranges are flattened with a
//     warp scan and all 32 lanes stride over the candidates; the k nearest are extracted by
//     k successive warp-shuffle min-reductions on the key (d², index) — ties break to the
//     lowest index — and the search moves to a coarser level until the k-th d² is inside
//     the radius the ring is guaranteed to cover, (1-1e-8)·h_l.
//
// Exactness: d² = ((dx·dx)+(dy·dy))+(dz·dz) in float64 with __dmul_rn/__dadd_rn (no FMA
// contraction), bitwise what cKDTree's sqeuclidean_distance_double gives for 3-vectors;
// returned distance = sqrt(d²) (correctly rounded on both sides).
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace at {

constexpr int kMaxLevels = 28;
constexpr int kMaxK = 32;
constexpr double kRingSafety = 1.0 - 1e-8;

struct KnnLevel {
    const int32_t* start;  // bucket offsets, mask + 2 entries
    const int32_t* perm;   // source index per slot (nullptr: identity, brute-force level)
    uint32_t mask;         // bucket count - 1
    int shift;             // cell coordinate shift (level index); < 0 for the brute-force level
    double cover2;         // (kRingSafety · h_l)²: every source closer than this was scanned
    double h;              // cell edge h0 · 2^l
    double inv_h;          // scale of the integer cell coordinates before `shift` (1/h0, or 1/h of a fine level)
    const double *sx, *sy, *sz;  // sources in this level's slot order (coalesced candidate reads), or nullptr
};

// Levels that get their own slot-ordered copy of the coordinates: the ones a query may start at
// (level 0 for k = 1, level 1 for k <= 6, level 2 beyond — see start_level()).
constexpr int kSlotOrderedLevels = 3;

// Fine levels: cell edges h0/2, h0/4, … built only when some level-0 cell is crowded (a regular
// lat-lon grid packs 1440 points into every ring around the poles, and 1440 copies of the pole
// itself).  They are not part of the octree the far search walks; a query whose level-0 ring is
// crowded starts at the fine level where that ring shrinks to ≈ 100 candidates.
constexpr int kMaxFine = 4;
constexpr int kDenseTarget = 64;  // candidates a ring should hold at least before a query goes below its start level (x4: when it does)

struct KnnDev {
    const double *x, *y, *z;     // sources, original order
    long long n;
    double ox, oy, oz;  // grid origin
    double inv_h0;
    int n_levels;       // including the brute-force level
    int n_fine;         // fine levels below level 0 (0 for evenly spaced sources)
    KnnLevel lv[kMaxLevels];
    KnnLevel fine[kMaxFine];  // fine[j]: cell edge h0 / 2^(j+1)
};

// Level by signed index: v >= 0 is lv[v], v < 0 is fine[-1 - v].
__device__ __forceinline__ const KnnLevel& level_at(const KnnDev& d, int v) { return v < 0 ? d.fine[-1 - v] : d.lv[v]; }

__device__ __forceinline__ uint32_t cell_hash(int cx, int cy, int cz) {
    uint32_t h = static_cast<uint32_t>(cx) * 73856093u ^ static_cast<uint32_t>(cy) * 19349663u ^
                 static_cast<uint32_t>(cz) * 83492791u;
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

__device__ __forceinline__ int cell_coord(double p, double origin, double inv_h) {
    // floor + clamp to [-2^30, 2^30): the conversion rounds towards -inf and saturates, the
    // clamp is two integer instructions (fmin / fmax on doubles were 6 % of the query kernel)
    const int c = __double2int_rd(__dmul_rn(__dsub_rn(p, origin), inv_h));
    return min(max(c, -1073741824), 1073741823);
}

__device__ __forceinline__ double dist2(double qx, double qy, double qz, double px, double py, double pz) {
    const double dx = __dsub_rn(px, qx), dy = __dsub_rn(py, qy), dz = __dsub_rn(pz, qz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// ---- build ---------------------------------------------------------------------------
__global__ void bbox_kernel(const double* __restrict__ x, const double* __restrict__ y,
                            const double* __restrict__ z, long long n, double* __restrict__ out6) {
    // out6 = {minx, miny, minz, maxx, maxy, maxz}; atomics on ordered bit patterns.
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double v[3] = {x[i], y[i], z[i]};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mn[a] = fmin(mn[a], v[a]);
            mx[a] = fmax(mx[a], v[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fmin(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmax(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        // order-preserving map double -> uint64
        auto enc = [](double d) {
            unsigned long long u = __double_as_longlong(d);
            return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
        };
        unsigned long long* o = reinterpret_cast<unsigned long long*>(out6);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(o + a, enc(mn[a]));
            atomicMax(o + 3 + a, enc(mx[a]));
        }
    }
}

// Occupancy of a coarse 64³ grid over the bounding box: estimates the area the points cover.
__global__ void coarse_occupancy_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                        const double* __restrict__ z, long long n, double ox, double oy,
                                        double oz, double inv_h, uint32_t* __restrict__ bits) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = min(63, max(0, cell_coord(x[i], ox, inv_h)));
    const int cy = min(63, max(0, cell_coord(y[i], oy, inv_h)));
    const int cz = min(63, max(0, cell_coord(z[i], oz, inv_h)));
    const int c = (cz * 64 + cy) * 64 + cx;
    atomicOr(bits + (c >> 5), 1u << (c & 31));
}

__global__ void popcount_kernel(const uint32_t* __restrict__ bits, int n_words, unsigned int* __restrict__ out) {
    unsigned int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x)
        c += __popc(bits[i]);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void bucket_count_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                    const double* __restrict__ z, long long n, double ox, double oy,
                                    double oz, double inv_h0, int shift, uint32_t mask,
                                    int32_t* __restrict__ counts, int32_t* __restrict__ bucket_of) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool live = i < n;
    uint32_t b = 0xffffffffu;
    if (live) {
        const int cx = cell_coord(x[i], ox, inv_h0) >> shift;
        const int cy = cell_coord(y[i], oy, inv_h0) >> shift;
        const int cz = cell_coord(z[i], oz, inv_h0) >> shift;
        b = cell_hash(cx, cy, cz) & mask;
        bucket_of[i] = static_cast<int32_t>(b);
    }
    // Grid points arrive in spatial order, so the lanes of a warp mostly fall into a handful of
    // buckets (at the coarse levels: one).  One atomic per distinct bucket per warp instead of
    // one per point: the coarse levels of a 6.6 M-point build were serialised on 256 counters.
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    if (live && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(counts + b, __popc(peers));
}

// Largest bucket of a level (start[] is the exclusive scan, m + 1 entries).
__global__ void max_bucket_kernel(const int32_t* __restrict__ start, uint32_t m, unsigned int* __restrict__ out) {
    unsigned int c = 0;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < m; b += gridDim.x * blockDim.x)
        c = max(c, static_cast<unsigned int>(start[b + 1] - start[b]));
    for (int o = 16; o > 0; o >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
    if ((threadIdx.x & 31) == 0 && c) atomicMax(out, c);
}

__global__ void bucket_scatter_kernel(const int32_t* __restrict__ bucket_of, long long n,
                                      const int32_t* __restrict__ start, int32_t* __restrict__ cursor,
                                      int32_t* __restrict__ perm) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool live = i < n;
    const int b = live ? bucket_of[i] : -1;
    // warp-aggregated as in bucket_count_kernel: the leader reserves the group's slots
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int base = 0;
    if (live && lane == leader) base = atomicAdd(cursor + b, __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live) perm[start[b] + base + __popc(peers & ((1u << lane) - 1u))] = static_cast<int32_t>(i);
}

__global__ void permute_points_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                      const double* __restrict__ z, const int32_t* __restrict__ perm,
                                      long long n, double* __restrict__ sx, double* __restrict__ sy,
                                      double* __restrict__ sz) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = perm[i];
    sx[i] = x[s];
    sy[i] = y[s];
    sz[i] = z[s];
}

// ---- exclusive scan of int32 (three small kernels) ---------------------------------------
constexpr int kScanBlock = 1024;  // elements per block (256 threads x 4)

__global__ void __launch_bounds__(256) scan_block_kernel(const int32_t* __restrict__ in, long long n,
                                                         int32_t* __restrict__ out,
                                                         int32_t* __restrict__ block_sums) {
    __shared__ int warp_tot[8];
    const long long base = static_cast<long long>(blockIdx.x) * kScanBlock + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = base + j < n ? in[base + j] : 0;
    const int t = v[0] + v[1] + v[2] + v[3];
    int inc = t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    int run = woff + inc - t;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (threadIdx.x == 255) block_sums[blockIdx.x] = woff + inc;
}

// Single block: exclusive scan of the block sums in place (sequential chunks of 1024).
__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ sums, int n_blocks,
                                                         int32_t* __restrict__ total) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n_blocks ? sums[i] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_tot[w];
        const int carry = carry_s;
        if (i < n_blocks) sums[i] = carry + woff + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total != nullptr) *total = carry_s;
}

__global__ void __launch_bounds__(256) scan_add_kernel(int32_t* __restrict__ out, long long n,
                                                       const int32_t* __restrict__ block_sums) {
    const long long base = static_cast<long long>(blockIdx.x) * kScanBlock + threadIdx.x * 4;
    const int off = block_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (base + j < n) out[base + j] += off;
}

// out[0..n) = exclusive scan of in[0..n); out[n] = total.  scratch: ceil(n/1024) int32.
static int exclusive_scan(const int32_t* in, long long n, int32_t* out, int32_t* scratch, cudaStream_t st) {
    const long long blocks = (n + kScanBlock - 1) / kScanBlock;
    if (blocks >= (1ll << 31)) return set_error(AT_ERR_UNSUPPORTED, "scan too large");
    scan_block_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(in, n, out, scratch);
    AT_LAUNCH_CHECK("scan_block_kernel");
    scan_sums_kernel<<<1, 1024, 0, st>>>(scratch, static_cast<int>(blocks), out + n);
    AT_LAUNCH_CHECK("scan_sums_kernel");
    scan_add_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(out, n, scratch);
    AT_LAUNCH_CHECK("scan_add_kernel");
    return AT_OK;
}

// ---- search --------------------------------------------------------------------------
struct Cand {
    double d2;
    long long idx;
};

__device__ __forceinline__ bool key_less(double d2a, long long ia, double d2b, long long ib) {
    return d2a < d2b || (d2a == d2b && ia < ib);
}

// Warp-wide minimum of the keys (d², index) held one per lane, d² >= 0 or +inf, index < 2^31:
// a non-negative double orders like its bit pattern, so three integer warp reductions (REDUX:
// high word, low word among the lanes that hold the smallest high word, index among the lanes
// that hold the smallest d²) replace five rounds of four shuffles and a 128-bit comparison.
__device__ __forceinline__ void warp_min_key(double& d2, long long& idx) {
    const unsigned hi = static_cast<unsigned>(__double2hiint(d2)), lo = static_cast<unsigned>(__double2loint(d2));
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    const bool holds = hi == mhi && lo == mlo;
    const unsigned midx = __reduce_min_sync(0xffffffffu, holds ? static_cast<unsigned>(idx) : 0xffffffffu);
    d2 = __hiloint2double(static_cast<int>(mhi), static_cast<int>(mlo));
    idx = static_cast<long long>(midx);
}

// Flattened ring-1 candidate ranges of one query at one level, held across the warp:
// lane j owns range [s0, s0 + cnt) at flat offset `off`; `total` candidates in all.
struct Ring {
    int s0, cnt, off, total;
};

__device__ __forceinline__ Ring ring_ranges(const KnnDev& d, const KnnLevel& L, double qx, double qy,
                                            double qz, int lane) {
    Ring r;
    r.s0 = 0;
    r.cnt = 0;
    if (L.shift < 0) {  // brute-force level: a single range over everything
        if (lane == 0) r.cnt = static_cast<int>(d.n);
    } else {
        int b = -1 - lane;  // distinct dummy values for lanes >= 27
        if (lane < 27) {
            const int cx = (cell_coord(qx, d.ox, L.inv_h) >> L.shift) + (lane % 3) - 1;
            const int cy = (cell_coord(qy, d.oy, L.inv_h) >> L.shift) + ((lane / 3) % 3) - 1;
            const int cz = (cell_coord(qz, d.oz, L.inv_h) >> L.shift) + (lane / 9) - 1;
            b = static_cast<int>(cell_hash(cx, cy, cz) & L.mask);
        }
        const unsigned same = __match_any_sync(0xffffffffu, b);
        const bool leader = lane < 27 && (__ffs(same) - 1) == lane;
        if (leader) {
            r.s0 = __ldg(L.start + b);
            r.cnt = __ldg(L.start + b + 1) - r.s0;
        }
    }
    int inc = r.cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    r.off = inc - r.cnt;
    r.total = __shfl_sync(0xffffffffu, inc, 31);
    return r;
}

// Slot of flat candidate t (t < total): binary search over the 32 lane offsets.
__device__ __forceinline__ int ring_slot(const Ring& r, int t) {
    // find the last lane whose off <= t and cnt > 0 covering t
    int lo = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const int probe = lo + step;
        const int off_p = __shfl_sync(0xffffffffu, r.off, probe & 31);
        if (probe < 32 && off_p <= t) lo = probe;
    }
    // lanes with cnt == 0 share the offset of the next non-empty lane; `lo` is the last lane
    // with off <= t, and since offsets are non-decreasing the owner is the last such lane
    // with cnt > 0 — which is `lo` itself unless trailing empty lanes follow the owner.
    const int s0 = __shfl_sync(0xffffffffu, r.s0, lo);
    const int off = __shfl_sync(0xffffffffu, r.off, lo);
    const int cnt = __shfl_sync(0xffffffffu, r.cnt, lo);
    return (t - off < cnt) ? s0 + (t - off) : -1;
}

// Load candidate at slot s of level L.
__device__ __forceinline__ void load_cand(const KnnDev& d, const KnnLevel& L, int s, double& px, double& py, double& pz,
                                          long long& idx) {
    if (L.perm == nullptr) {
        idx = s;
        px = d.x[s];
        py = d.y[s];
        pz = d.z[s];
    } else {
        idx = __ldg(L.perm + s);
        if (L.sx != nullptr) {
            px = __ldg(L.sx + s);
            py = __ldg(L.sy + s);
            pz = __ldg(L.sz + s);
        } else {
            px = __ldg(d.x + idx);
            py = __ldg(d.y + idx);
            pz = __ldg(d.z + idx);
        }
    }
}

// One selection pass over the ring: smallest key strictly greater than (prev_d2, prev_idx)
// with d2 < ub2.  Returns the warp-wide minimum (d2 = +inf when none).
__device__ __forceinline__ Cand select_next(const KnnDev& d, const KnnLevel& L, const Ring& r,
                                            double qx, double qy, double qz, double prev_d2,
                                            long long prev_idx, double ub2, int lane) {
    Cand best;
    best.d2 = INFINITY;
    best.idx = d.n;
    // every lane runs the same number of iterations (ring_slot uses full-warp shuffles)
    for (int t0 = 0; t0 < r.total; t0 += 32) {
        const int t = t0 + lane;
        const int s = ring_slot(r, t < r.total ? t : r.total - 1);
        if (t < r.total && s >= 0) {
            double px, py, pz;
            long long idx;
            load_cand(d, L, s, px, py, pz, idx);
            const double c2 = dist2(qx, qy, qz, px, py, pz);
            if (c2 < ub2 && key_less(prev_d2, prev_idx, c2, idx) && key_less(c2, idx, best.d2, best.idx)) {
                best.d2 = c2;
                best.idx = idx;
            }
        }
    }
    warp_min_key(best.d2, best.idx);
    return best;
}

// For k > 1 the k selection passes would evaluate every candidate k times.  When the ring holds
// at most kCandCap candidates their keys (d2, index) are computed once into the warp's slice of
// shared memory and the passes select from there (no loads, no ring_slot shuffles, no float64
// arithmetic per pass).
constexpr int kCandCap = 512;
struct CandCache {
    double* d2;  // [kCandCap] of this warp
    int* idx;
};

__device__ __forceinline__ void fill_cache(const KnnDev& d, const KnnLevel& L, const Ring& r, double qx,
                                           double qy, double qz, double ub2, int lane, const CandCache& c) {
    for (int t0 = 0; t0 < r.total; t0 += 32) {
        const int t = t0 + lane;
        const int s = ring_slot(r, t < r.total ? t : r.total - 1);
        if (t < r.total) {
            double c2 = INFINITY;
            long long idx = d.n;
            if (s >= 0) {
                double px, py, pz;
                load_cand(d, L, s, px, py, pz, idx);
                c2 = dist2(qx, qy, qz, px, py, pz);
                if (!(c2 < ub2)) c2 = INFINITY;
            }
            c.d2[t] = c2;
            c.idx[t] = static_cast<int>(idx);
        }
    }
    __syncwarp();
}

// Smallest cached key strictly greater than (prev_d2, prev_idx); d2 = +inf when none.
__device__ __forceinline__ Cand select_cached(const KnnDev& d, const CandCache& c, int total, double prev_d2,
                                              long long prev_idx, int lane) {
    Cand best;
    best.d2 = INFINITY;
    best.idx = d.n;
    for (int t = lane; t < total; t += 32) {
        const double c2 = c.d2[t];
        const long long idx = c.idx[t];
        if (c2 < INFINITY && key_less(prev_d2, prev_idx, c2, idx) && key_less(c2, idx, best.d2, best.idx)) {
            best.d2 = c2;
            best.idx = idx;
        }
    }
    warp_min_key(best.d2, best.idx);
    return best;
}

// Level a search for k neighbours starts at: the ring of level l (cell edge h0·2^l, h0 = the mean
// point spacing) is sure to hold every source within h0·2^l, a disc of about pi·4^l points.
__host__ __device__ inline int start_level(int k, int n_grid_levels) {
    const int want = k <= 1 ? 0 : (k <= 6 ? 1 : 2);
    return want < n_grid_levels ? want : (n_grid_levels > 0 ? n_grid_levels - 1 : 0);
}

// Full k-NN of one query by one warp.  Lane j < k ends up holding the j-th neighbour.
// Returns tie bits (valid when want_tie).
//
// The search starts at `first_level` (start_level(k)) — lower, down into the fine levels, when
// that ring is crowded — and tries levels below `max_levels`; `done` tells whether the answer
// is final (the caller hands unfinished queries to the tree search).
template <bool CACHE>
__device__ __forceinline__ unsigned knn_one(const KnnDev& d, double qx, double qy, double qz, int k,
                                            double ub2, bool want_tie, int lane, int first_level, int max_levels, bool& done,
                                            double& my_d2, long long& my_idx, const CandCache& cache) {
    unsigned tie = 0;
    done = true;
    // Crowded neighbourhood (warp-uniform: the ring total is the same in every lane): go down to
    // the level where the ring holds about `target` candidates and walk back up from there.
    int first = first_level;
    Ring r0;
    r0.s0 = r0.cnt = r0.off = r0.total = 0;
    const bool can_descend = (d.n_fine > 0 || first_level > 0) && first_level < max_levels;
    if (can_descend) {
        r0 = ring_ranges(d, d.lv[first_level], qx, qy, qz, lane);
        const int target = max(kDenseTarget, 16 * k);
        if (r0.total > 4 * target) {
            int halvings = 1;  // each halving of the cell edge divides a surface ring by about 4
            while (first_level - halvings > -d.n_fine && (r0.total >> (2 * halvings)) > target) ++halvings;
            first = first_level - halvings;
        }
    }
    for (int level = first; level < max_levels; ++level) {
        const KnnLevel& L = level_at(d, level);
        const bool last = level == d.n_levels - 1;
        // A level whose ring cannot decide anything (cover radius below the best possible
        // distance) is still scanned: cheap, and usually terminates at level 0.
        const Ring r = (level == first_level && can_descend) ? r0 : ring_ranges(d, L, qx, qy, qz, lane);
        my_d2 = INFINITY;
        my_idx = d.n;
        tie = 0;
        if (r.total == 0 && !last && !(ub2 <= L.cover2)) continue;
        double pd2 = -1.0;
        long long pidx = -1;
        int found = 0;
        const bool cached = CACHE && k > 1 && r.total <= kCandCap;  // warp-uniform
        if (cached) fill_cache(d, L, r, qx, qy, qz, ub2, lane, cache);
        for (int j = 0; j < k; ++j) {
            const Cand c = cached ? select_cached(d, cache, r.total, pd2, pidx, lane)
                                  : select_next(d, L, r, qx, qy, qz, pd2, pidx, ub2, lane);
            if (!(c.d2 < INFINITY)) break;
            if (j > 0 && c.d2 == pd2) tie |= 1u;
            if (lane == j) {
                my_d2 = c.d2;
                my_idx = c.idx;
            }
            pd2 = c.d2;
            pidx = c.idx;
            ++found;
        }
        const bool complete = found == k && pd2 < L.cover2;
        if (complete || last || ub2 <= L.cover2) {
            if (want_tie && found == k) {
                const Cand c = cached ? select_cached(d, cache, r.total, pd2, pidx, lane)
                                      : select_next(d, L, r, qx, qy, qz, pd2, pidx, ub2, lane);
                if (c.d2 == pd2) tie |= 2u;
            }
            __syncwarp();  // the cache is rewritten by the next level / query
            return tie;
        }
        __syncwarp();
    }
    done = false;
    return tie;
}

constexpr long long kPending = -1;  // idx_out[q*k] of a query the near search left to the tree search

// Gather buffers of the other ranks (P2P-mapped over NVLink), already offset to this rank's slice:
// the query kernels store every finished index there as well, so the all-gather of a sharded
// query happens inside the search instead of in a collective after it (at_knn_query_gather).
constexpr int kMaxPeers = 15;
struct PeerOut {
    long long* ptr[kMaxPeers];
    int n;
};

// CACHE: k > 1, candidate keys cached in shared memory (knn_one); PEERS: also store to peers
template <bool CACHE, bool PEERS>
__global__ void __launch_bounds__(256, 4)
    knn_query_kernel(const KnnDev d, const double* __restrict__ qx, const double* __restrict__ qy,
                     const double* __restrict__ qz, long long nq, int k, double ub2, int first_level, int near_levels,
                     long long* __restrict__ idx_out, double* __restrict__ dist_out,
                     uint8_t* __restrict__ tie_out, const PeerOut peers, const int* __restrict__ list,
                     const unsigned int* __restrict__ list_n) {
    // `list` != nullptr: only the *list_n queries named there (what knn1_thread_kernel left over)
    if (list != nullptr) nq = *list_n;
    // per-warp candidate cache (k > 1 only; the k = 1 launch passes no shared memory)
    extern __shared__ __align__(16) unsigned char s_cache[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CandCache cache = {nullptr, nullptr};
    if (CACHE) {
        cache.d2 = reinterpret_cast<double*>(s_cache) + warp * kCandCap;
        cache.idx = reinterpret_cast<int*>(s_cache + (blockDim.x >> 5) * kCandCap * sizeof(double)) + warp * kCandCap;
    }
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < nq; w += warps) {
        // Grids run pole to pole, and a regular lat-lon source is crowded at both poles: the
        // expensive queries sit at the two ends of the array.  Taking the queries from both ends
        // inwards starts them first, so they overlap the cheap ones instead of forming the tail.
        long long q = (w & 1) ? nq - 1 - (w >> 1) : (w >> 1);
        if (list != nullptr) q = list[q];
        const double x = qx[q], y = qy[q], z = qz[q];
        double d2;
        long long idx;
        bool done;
        const unsigned tie = knn_one<CACHE>(d, x, y, z, k, ub2, tie_out != nullptr, lane, first_level, near_levels, done, d2, idx, cache);
        if (!done) {
            if (lane == 0) idx_out[q * k] = kPending;
            continue;
        }
        if (lane < k) {
            idx_out[q * k + lane] = idx;
            if (dist_out != nullptr) dist_out[q * k + lane] = sqrt(d2);
            if (PEERS)
                for (int p = 0; p < peers.n; ++p) peers.ptr[p][q * k + lane] = idx;
        }
        if (tie_out != nullptr && lane == 0) tie_out[q] = static_cast<uint8_t>(tie);
    }
}

// ---- k = 1, one THREAD per query ------------------------------------------------------------
// With about one source per level-0 cell a ring holds a dozen candidates: a warp per query spends
// its instructions on the ring bookkeeping (27 hashes spread over lanes, a scan, a slot search, a
// warp reduction — 446 warp instructions per query) rather than on distances.  Here a thread
// walks the 27 cells of its own query and keeps the smallest key (d², index) — the same exact
// answer, the minimum does not depend on the order — at a fraction of the issue slots.
// Consecutive queries are neighbours on the sphere, so the lanes of a warp read the same buckets
// (L1 hits).  Level 0 keeps, for this kernel, one {start, end} pair per bucket (one 8-byte load)
// and one 32-byte record {x, y, z, index} per source in slot order (one sector per candidate).
// A query whose ring is crowded (more than kThreadRingMax candidates: the poles of a regular
// lat-lon grid) or whose nearest source is not inside the level-0 cover radius is appended to
// `pending` and finished by knn_query_kernel, one warp per query, with its fine and coarse levels.
// (Letting each warp finish its own leftovers instead put up to 32 crowded queries in sequence
// on the polar warps: 0.36 ms against 0.19 + 0.06 ms for the two kernels.)
constexpr int kThreadRingMax = 96;

struct __align__(32) KnnRecord {
    double x, y, z;
    long long idx;
};

template <bool PEERS>
__global__ void __launch_bounds__(256)
    knn1_thread_kernel(const KnnDev d, const int2* __restrict__ range0, const KnnRecord* __restrict__ rec0,
                       const double* __restrict__ qx, const double* __restrict__ qy, const double* __restrict__ qz,
                       long long nq, double ub2, long long* __restrict__ idx_out, double* __restrict__ dist_out,
                       const PeerOut peers, int* __restrict__ pending, unsigned int* __restrict__ n_pending) {
    const KnnLevel& L = d.lv[0];
    const int lane = threadIdx.x & 31;
    const long long n_blocks32 = (nq + 31) >> 5;
    const long long wi = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (wi >= n_blocks32) return;
    // blocks of 32 consecutive queries taken from both ends of the array inwards (the crowded
    // poles first, see knn_query_kernel)
    const long long b32 = (wi & 1) ? n_blocks32 - 1 - (wi >> 1) : (wi >> 1);
    const long long q = (b32 << 5) + lane;
    bool leftover = false;
    if (q < nq) {
        const double x = qx[q], y = qy[q], z = qz[q];
        const int cx = cell_coord(x, d.ox, L.inv_h), cy = cell_coord(y, d.oy, L.inv_h), cz = cell_coord(z, d.oz, L.inv_h);
        double best = INFINITY;
        long long best_idx = 0x7fffffffffffffffll;
        int scanned = 0;
#pragma unroll 1
        for (int dz = -1; dz <= 1 && scanned <= kThreadRingMax; ++dz) {
            // the nine buckets of this plane of cells: all lookups in flight together, then ONE
            // loop over their candidates (a loop per bucket left most lanes idle: lanes disagree
            // on which buckets are empty, the warp ran the sum of the per-bucket maxima)
            int first[9], upto[9];  // slot offset and cumulative count per bucket (registers: fully unrolled)
            int total = 0;
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                const int2 r = __ldg(range0 + (cell_hash(cx + (c % 3) - 1, cy + (c / 3) - 1, cz + dz) & L.mask));
                first[c] = r.x - total;  // slot = first[c] + i for the candidates i of bucket c
                total += r.y - r.x;
                upto[c] = total;
            }
            scanned += total;
            if (scanned > kThreadRingMax) break;
            for (int i = 0; i < total; ++i) {
                int base = first[0];
#pragma unroll
                for (int c = 1; c < 9; ++c) base = i >= upto[c - 1] ? first[c] : base;
                const int s = base + i;
                const double2 xy = __ldg(reinterpret_cast<const double2*>(rec0 + s));
                const double2 zi = __ldg(reinterpret_cast<const double2*>(rec0 + s) + 1);
                const double c2 = dist2(x, y, z, xy.x, xy.y, zi.x);
                const long long idx = __double_as_longlong(zi.y);
                // (a bucket shared by two cells is read twice: same key, no change)
                if (c2 < ub2 && (c2 < best || (c2 == best && idx < best_idx))) {
                    best = c2;
                    best_idx = idx;
                }
            }
        }
        const bool decided = scanned <= kThreadRingMax && (best < L.cover2 || ub2 <= L.cover2);
        if (decided) {
            const long long found = best < INFINITY ? best_idx : d.n;
            idx_out[q] = found;
            if (dist_out != nullptr) dist_out[q] = sqrt(best);
            if (PEERS)
                for (int p = 0; p < peers.n; ++p) peers.ptr[p][q] = found;
        } else {
            leftover = true;
        }
    }
    // warp-aggregated append to the list of queries the warp kernel finishes
    const unsigned who = __ballot_sync(0xffffffffu, leftover);
    if (who != 0) {
        unsigned base = 0;
        if (lane == __ffs(who) - 1) base = atomicAdd(n_pending, __popc(who));
        base = __shfl_sync(0xffffffffu, base, __ffs(who) - 1);
        if (leftover) pending[base + __popc(who & ((1u << lane) - 1u))] = static_cast<int>(q);
    }
}

// Level-0 tables of knn1_thread_kernel, built once per index.
__global__ void knn_level0_tables_kernel(const int32_t* __restrict__ start, const int32_t* __restrict__ perm,
                                         const double* __restrict__ x, const double* __restrict__ y,
                                         const double* __restrict__ z, uint32_t n_buckets, long long n,
                                         int2* __restrict__ range0, KnnRecord* __restrict__ rec0) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n_buckets) range0[i] = make_int2(start[i], start[i + 1]);
    if (i < n) {
        const int s = perm[i];
        KnnRecord r;
        r.x = x[s];
        r.y = y[s];
        r.z = z[s];
        r.idx = s;
        rec0[i] = r;
    }
}

// ---- tree search for the queries the near search could not finish -----------------------
// The grid levels form an implicit octree (cell c of level l has the 8 children 2c+{0,1} of
// level l-1; every source lies in cell (0,0,0) of the top level).  One thread per pending
// query walks it depth first, nearest child first, pruning cells whose box is farther than
// the current k-th best, and scans a cell's bucket once it holds at most kLeafMax points.
// Hash collisions can make a bucket hold points of other cells: they are just more (valid)
// candidates; a point met twice is recognised by its (d², index).  The result is the exact
// k smallest keys (d², index), the same as the near search.
constexpr int kLeafMax = 48;
constexpr int kStackMax = 8 * kMaxLevels;

__device__ __forceinline__ double box_min_d2(const KnnDev& d, double h, int cx, int cy, int cz, double qx, double qy,
                                             double qz) {
    const double eps = 1e-8 * h;  // the cell assignment is rounded: widen the box a little
    const double lx = d.ox + cx * h - eps, ly = d.oy + cy * h - eps, lz = d.oz + cz * h - eps;
    const double w = h + 2.0 * eps;
    const double dx = fmax(fmax(lx - qx, qx - (lx + w)), 0.0);
    const double dy = fmax(fmax(ly - qy, qy - (ly + w)), 0.0);
    const double dz = fmax(fmax(lz - qz, qz - (lz + w)), 0.0);
    return (dx * dx + dy * dy + dz * dz) * (1.0 - 1e-10);  // a lower bound despite rounding
}

template <bool PEERS>
__global__ void __launch_bounds__(128)
    knn_tree_kernel(const KnnDev d, const double* __restrict__ qx, const double* __restrict__ qy,
                    const double* __restrict__ qz, long long nq, int k, double ub2,
                    long long* __restrict__ idx_out, double* __restrict__ dist_out, uint8_t* __restrict__ tie_out,
                    const PeerOut peers) {
    const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= nq || idx_out[q * k] != kPending) return;
    const double x = qx[q], y = qy[q], z = qz[q];
    const int kk = k + (tie_out != nullptr ? 1 : 0);  // one extra to see a tie across the k-th place

    double bd[kMaxK + 1];
    long long bi[kMaxK + 1];
    int cnt = 0;
    unsigned long long stack[kStackMax];
    int sp = 0;
    const int root = d.n_levels - 2;  // top grid level; the last level is the brute-force bucket
    stack[sp++] = static_cast<unsigned long long>(root) << 60;

    while (sp > 0) {
        const unsigned long long e = stack[--sp];
        const int level = static_cast<int>(e >> 60);
        const int cx = static_cast<int>((e >> 40) & 0xfffff), cy = static_cast<int>((e >> 20) & 0xfffff),
                  cz = static_cast<int>(e & 0xfffff);
        const KnnLevel& L = d.lv[level];
        const double worst = cnt == kk ? bd[kk - 1] : INFINITY;
        const double m = box_min_d2(d, L.h, cx, cy, cz, x, y, z);
        if (m > worst || m >= ub2) continue;  // pruned since it was pushed
        const uint32_t b = cell_hash(cx, cy, cz) & L.mask;
        const int s0 = __ldg(L.start + b), s1 = __ldg(L.start + b + 1);
        if (s1 == s0) continue;
        if (level == 0 || s1 - s0 <= kLeafMax) {
            for (int s = s0; s < s1; ++s) {
                double px, py, pz;
                long long idx;
                load_cand(d, L, s, px, py, pz, idx);
                const double c2 = dist2(x, y, z, px, py, pz);
                if (!(c2 < ub2)) continue;
                if (cnt == kk && !key_less(c2, idx, bd[kk - 1], bi[kk - 1])) continue;
                int pos = cnt < kk ? cnt : kk - 1;  // insertion sort, ascending by (d², index)
                bool dup = false;
                while (pos > 0 && !key_less(bd[pos - 1], bi[pos - 1], c2, idx)) {
                    if (bd[pos - 1] == c2 && bi[pos - 1] == idx) {
                        dup = true;
                        break;
                    }
                    --pos;
                }
                if (dup) continue;
                const int end = cnt < kk ? cnt : kk - 1;
                for (int j = end; j > pos; --j) {
                    bd[j] = bd[j - 1];
                    bi[j] = bi[j - 1];
                }
                bd[pos] = c2;
                bi[pos] = idx;
                if (cnt < kk) ++cnt;
            }
            continue;
        }
        // children of level - 1, farthest pushed first so the nearest is popped first
        const KnnLevel& C = d.lv[level - 1];
        double cm[8];
        int order[8];
        int n_child = 0;
        for (int c = 0; c < 8; ++c) {
            const double mc = box_min_d2(d, C.h, 2 * cx + (c & 1), 2 * cy + ((c >> 1) & 1), 2 * cz + (c >> 2), x, y, z);
            if (mc > worst || mc >= ub2) continue;
            int p = n_child++;
            while (p > 0 && cm[order[p - 1]] < mc) {
                order[p] = order[p - 1];
                --p;
            }
            order[p] = c;
            cm[c] = mc;
        }
        for (int j = 0; j < n_child && sp < kStackMax; ++j) {
            const int c = order[j];
            stack[sp++] = (static_cast<unsigned long long>(level - 1) << 60) |
                          (static_cast<unsigned long long>(2 * cx + (c & 1)) << 40) |
                          (static_cast<unsigned long long>(2 * cy + ((c >> 1) & 1)) << 20) |
                          static_cast<unsigned long long>(2 * cz + (c >> 2));
        }
    }

    unsigned tie = 0;
    for (int j = 0; j < k; ++j) {
        const bool have = j < cnt;
        const long long found = have ? bi[j] : d.n;
        idx_out[q * k + j] = found;
        if (PEERS)
            for (int p = 0; p < peers.n; ++p) peers.ptr[p][q * k + j] = found;
        if (dist_out != nullptr) dist_out[q * k + j] = have ? sqrt(bd[j]) : INFINITY;
        if (have && j > 0 && bd[j] == bd[j - 1]) tie |= 1u;
    }
    if (tie_out != nullptr) {
        if (cnt > k && bd[k] == bd[k - 1]) tie |= 2u;
        tie_out[q] = static_cast<uint8_t>(tie);
    }
}

// min over sources of the distance to the 2nd nearest source (k = 2 self query).
__global__ void __launch_bounds__(256)
    min_second_nn_kernel(const KnnDev d, long long first, long long end, unsigned long long* __restrict__ out_bits) {
    const int lane = threadIdx.x & 31;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    double best = INFINITY;
    for (long long q = first + ((static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5); q < end; q += warps) {
        double d2;
        long long idx;
        bool done;
        const CandCache no_cache = {nullptr, nullptr};
        knn_one<false>(d, d.x[q], d.y[q], d.z[q], 2, INFINITY, false, lane, start_level(2, d.n_levels - 1), d.n_levels, done, d2, idx, no_cache);
        const double second = __shfl_sync(0xffffffffu, d2, 1);
        best = fmin(best, second);
    }
    if (lane == 0) {
        const double dist = sqrt(best);  // monotone: min of sqrt == sqrt of min
        atomicMin(out_bits, static_cast<unsigned long long>(__double_as_longlong(dist)));
    }
}

// Ball union: mark every source within r of any query.
__global__ void __launch_bounds__(256)
    ball_mark_kernel(const KnnDev d, int level, const double* __restrict__ qx,
                     const double* __restrict__ qy, const double* __restrict__ qz, long long nq, double r2,
                     uint8_t* __restrict__ mark) {
    const int lane = threadIdx.x & 31;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const KnnLevel& L = level_at(d, level);
    for (long long q = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; q < nq; q += warps) {
        const double x = qx[q], y = qy[q], z = qz[q];
        const Ring r = ring_ranges(d, L, x, y, z, lane);
        for (int t0 = 0; t0 < r.total; t0 += 32) {
            const int t = t0 + lane;
            const int s = ring_slot(r, t < r.total ? t : r.total - 1);
            if (t < r.total && s >= 0) {
                double px, py, pz;
                long long idx;
                load_cand(d, L, s, px, py, pz, idx);
                if (dist2(x, y, z, px, py, pz) <= r2) mark[idx] = 1;
            }
        }
    }
}

}  // namespace at

struct at_knn {
    long long n = 0;
    int device = 0;
    double h0 = 0;
    double* d_x = nullptr;
    double* d_y = nullptr;
    double* d_z = nullptr;
    std::vector<void*> owned;  // every device allocation, for destroy
    at::KnnDev dev;
    // level-0 tables of the one-thread-per-query kernel (k = 1): {start, end} per bucket and a
    // 32-byte {x, y, z, index} record per source in slot order
    int2* d_range0 = nullptr;
    at::KnnRecord* d_rec0 = nullptr;
    // scratch of the k = 1 query: ids of the queries the thread kernel leaves to the warp kernel,
    // and their count (grown on demand; queries on one index are issued one at a time)
    int* d_pending = nullptr;
    unsigned int* d_n_pending = nullptr;
    long long pending_cap = 0;
};

using namespace at;

namespace {

int dev_alloc(at_knn* k, void** p, size_t bytes) {
    cudaError_t e = device_alloc(p, bytes > 0 ? bytes : 16);
    if (e != cudaSuccess)
        return set_error(e == cudaErrorMemoryAllocation ? AT_ERR_NOMEM : AT_ERR_CUDA,
                         "at_knn_create: device allocation (%zu) failed: %s", bytes, cudaGetErrorString(e));
    k->owned.push_back(*p);
    return AT_OK;
}

double dec_ordered(unsigned long long u) {
    u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
    memcpy(&d, &u, 8);
    return d;
}

}  // namespace

extern "C" int at_knn_destroy(at_knn_t* k) {
    if (k == nullptr) return AT_OK;
    if (k->d_pending != nullptr) device_free(k->d_pending);
    if (k->d_n_pending != nullptr) device_free(k->d_n_pending);
    for (void* p : k->owned) device_free(p);
    delete k;
    return AT_OK;
}

extern "C" int at_knn_create(const double* x, const double* y, const double* z, int64_t n, int on_device,
                             double cell_size, at_knn_t** out) {
    AT_REQUIRE(out != nullptr, "at_knn_create: out is null");
    *out = nullptr;
    AT_REQUIRE(x != nullptr && y != nullptr && z != nullptr, "at_knn_create: null coordinates");
    AT_REQUIRE(n >= 1 && n < (1ll << 31) - 64, "at_knn_create: need 1 <= n < 2^31 (n=%lld)", (long long)n);

    at_knn* k = new at_knn();
    k->n = n;
    cudaGetDevice(&k->device);
    int rc = AT_OK;
#define KNN_TRY(expr)             \
    do {                          \
        rc = (expr);              \
        if (rc != AT_OK) {        \
            at_knn_destroy(k);    \
            return rc;            \
        }                         \
    } while (0)
#define KNN_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            at_knn_destroy(k);                                                                  \
            return set_error(AT_ERR_CUDA, "at_knn_create: %s failed: %s", #expr, cudaGetErrorString(_e)); \
        }                                                                                       \
    } while (0)

    const size_t nb = static_cast<size_t>(n) * 8;
    KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&k->d_x), nb));
    KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&k->d_y), nb));
    KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&k->d_z), nb));
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    KNN_CUDA(cudaMemcpy(k->d_x, x, nb, kind));
    KNN_CUDA(cudaMemcpy(k->d_y, y, nb, kind));
    KNN_CUDA(cudaMemcpy(k->d_z, z, nb, kind));

    const unsigned pt_blocks = static_cast<unsigned>((n + 255) / 256);
    double extent_cells = 0.0;  // bounding-box extent in level-0 cells
    k->dev.n_fine = 0;

    // bounding box
    double* d_box = nullptr;
    KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&d_box), 6 * 8));
    {
        unsigned long long init[6];
        for (int a = 0; a < 3; ++a) {
            init[a] = ~0ull;
            init[3 + a] = 0ull;
        }
        KNN_CUDA(cudaMemcpy(d_box, init, sizeof(init), cudaMemcpyHostToDevice));
        bbox_kernel<<<std::min(pt_blocks, 1024u), 256>>>(k->d_x, k->d_y, k->d_z, n, d_box);
        KNN_CUDA(cudaGetLastError());
        KNN_CUDA(cudaMemcpy(init, d_box, sizeof(init), cudaMemcpyDeviceToHost));
        double mn[3], mx[3];
        for (int a = 0; a < 3; ++a) {
            mn[a] = dec_ordered(init[a]);
            mx[a] = dec_ordered(init[3 + a]);
            if (!std::isfinite(mn[a]) || !std::isfinite(mx[a])) {
                at_knn_destroy(k);
                return set_error(AT_ERR_INVALID, "at_knn_create: coordinates must be finite");
            }
        }
        const double extent = std::max({mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2], 1e-300});
        k->dev.ox = mn[0];
        k->dev.oy = mn[1];
        k->dev.oz = mn[2];

        double h0 = cell_size;
        if (!(h0 > 0)) {
            // Estimate the covered surface from the occupancy of a coarse 64³ grid: points of a
            // grid on a sphere fill ~area / hc² coarse cells; spacing ≈ sqrt(area / n).
            uint32_t* d_bits = nullptr;
            unsigned int* d_cnt = nullptr;
            const int n_words = 64 * 64 * 64 / 32;
            KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&d_bits), n_words * 4));
            KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&d_cnt), 4));
            KNN_CUDA(cudaMemset(d_bits, 0, n_words * 4));
            KNN_CUDA(cudaMemset(d_cnt, 0, 4));
            const double hc = extent / 64.0 * (1.0 + 1e-9);
            coarse_occupancy_kernel<<<pt_blocks, 256>>>(k->d_x, k->d_y, k->d_z, n, mn[0], mn[1], mn[2],
                                                       1.0 / hc, d_bits);
            KNN_CUDA(cudaGetLastError());
            popcount_kernel<<<32, 256>>>(d_bits, n_words, d_cnt);
            KNN_CUDA(cudaGetLastError());
            unsigned int occ = 0;
            KNN_CUDA(cudaMemcpy(&occ, d_cnt, 4, cudaMemcpyDeviceToHost));
            // A surface crossing a cubic lattice occupies ~1.5 cells per hc² of area.
            const double area = std::max(1.0, static_cast<double>(occ)) * hc * hc / 1.5;
            const double spacing = std::sqrt(area / static_cast<double>(n));
            // one point per cell on average: a k = 1 query reads ~11 candidates in its ring;
            // larger k start one or two levels up (start_level), where the cells are 2x / 4x this
            h0 = 1.0 * spacing;
        }
        h0 = std::max(h0, extent / 1.0e6);  // keeps level-0 integer coordinates small
        h0 = std::max(h0, 1e-300);
        k->h0 = h0;
        k->dev.inv_h0 = 1.0 / h0;
        extent_cells = extent / h0;

        // levels until one cell spans the whole bounding box, then the brute-force level
        int n_grid = 1;
        while (h0 * std::ldexp(1.0, n_grid - 1) < extent * 2.0 && n_grid < kMaxLevels - 1) ++n_grid;
        k->dev.n_levels = n_grid + 1;
    }

    // bucket tables
    // Four buckets per point at level 0 (about one point per cell there) and a quarter of that
    // per level up, like the cells: colliding cells share a bucket, and at one bucket per cell
    // every ring read twice the candidates it needed (and overflowed the k > 1 cache).
    uint32_t m0 = 1024;
    while (m0 < 4 * static_cast<uint64_t>(n) && m0 < (1u << 26)) m0 <<= 1;
    int32_t* d_bucket_of = nullptr;
    int32_t* d_counts = nullptr;
    int32_t* d_scratch = nullptr;
    KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&d_bucket_of), static_cast<size_t>(n) * 4));
    KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&d_counts), (static_cast<size_t>(m0) + 2) * 4));
    KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&d_scratch), (static_cast<size_t>(m0) / kScanBlock + 2) * 4));

    const int n_grid = k->dev.n_levels - 1;
    for (int l = 0; l < n_grid; ++l) {
        uint32_t m = m0 >> std::min(2 * l, 30);
        m = std::max(m, 256u);
        int32_t* d_start = nullptr;
        int32_t* d_perm = nullptr;
        KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&d_start), (static_cast<size_t>(m) + 2) * 4));
        KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&d_perm), static_cast<size_t>(n) * 4));
        KNN_CUDA(cudaMemset(d_counts, 0, (static_cast<size_t>(m) + 2) * 4));
        bucket_count_kernel<<<pt_blocks, 256>>>(k->d_x, k->d_y, k->d_z, n, k->dev.ox, k->dev.oy, k->dev.oz,
                                               k->dev.inv_h0, l, m - 1, d_counts, d_bucket_of);
        KNN_CUDA(cudaGetLastError());
        KNN_TRY(exclusive_scan(d_counts, m, d_start, d_scratch, nullptr));
        KNN_CUDA(cudaMemset(d_counts, 0, (static_cast<size_t>(m) + 2) * 4));
        bucket_scatter_kernel<<<pt_blocks, 256>>>(d_bucket_of, n, d_start, d_counts, d_perm);
        KNN_CUDA(cudaGetLastError());
        KnnLevel& L = k->dev.lv[l];
        L.sx = L.sy = L.sz = nullptr;
        if (l < kSlotOrderedLevels) {  // coordinates in this level's slot order: coalesced candidate reads
            double *sx = nullptr, *sy = nullptr, *sz = nullptr;
            KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&sx), nb));
            KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&sy), nb));
            KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&sz), nb));
            permute_points_kernel<<<pt_blocks, 256>>>(k->d_x, k->d_y, k->d_z, d_perm, n, sx, sy, sz);
            KNN_CUDA(cudaGetLastError());
            L.sx = sx;
            L.sy = sy;
            L.sz = sz;
        }
        L.start = d_start;
        L.perm = d_perm;
        L.mask = m - 1;
        L.shift = l;
        L.h = k->h0 * std::ldexp(1.0, l);
        L.inv_h = k->dev.inv_h0;
        const double cover = kRingSafety * L.h;
        L.cover2 = cover * cover;
        if (l == 0) {
            KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&k->d_range0), static_cast<size_t>(m) * sizeof(int2)));
            KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&k->d_rec0), static_cast<size_t>(n) * sizeof(KnnRecord)));
            const long long items = std::max<long long>(m, n);
            knn_level0_tables_kernel<<<static_cast<unsigned>((items + 255) / 256), 256>>>(d_start, d_perm, k->d_x, k->d_y, k->d_z, m, n, k->d_range0, k->d_rec0);
            KNN_CUDA(cudaGetLastError());
            // crowded cells? (sources that are far from evenly spaced: the poles of a regular
            // lat-lon grid) — then build levels finer than h0 for the queries that land there
            unsigned int* d_max = reinterpret_cast<unsigned int*>(d_scratch);
            KNN_CUDA(cudaMemset(d_max, 0, 4));
            max_bucket_kernel<<<std::min((m + 255u) / 256u, 1024u), 256>>>(d_start, m, d_max);
            KNN_CUDA(cudaGetLastError());
            unsigned int c_max = 0;
            KNN_CUDA(cudaMemcpy(&c_max, d_max, 4, cudaMemcpyDeviceToHost));
            int n_fine = 0;
            if (c_max > 64)
                while (n_fine < kMaxFine && (c_max >> (2 * n_fine)) > 16) ++n_fine;
            // the finest integer cell coordinates must stay well inside int32
            while (n_fine > 0 && extent_cells * std::ldexp(1.0, n_fine) > 5.0e8) --n_fine;
            k->dev.n_fine = n_fine;
        }
    }
    for (int j = 0; j < k->dev.n_fine; ++j) {
        int32_t* d_start = nullptr;
        int32_t* d_perm = nullptr;
        KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&d_start), (static_cast<size_t>(m0) + 2) * 4));
        KNN_TRY(dev_alloc(k, reinterpret_cast<void**>(&d_perm), static_cast<size_t>(n) * 4));
        const double inv_h = k->dev.inv_h0 * std::ldexp(1.0, j + 1);
        KNN_CUDA(cudaMemset(d_counts, 0, (static_cast<size_t>(m0) + 2) * 4));
        bucket_count_kernel<<<pt_blocks, 256>>>(k->d_x, k->d_y, k->d_z, n, k->dev.ox, k->dev.oy, k->dev.oz, inv_h, 0,
                                               m0 - 1, d_counts, d_bucket_of);
        KNN_CUDA(cudaGetLastError());
        KNN_TRY(exclusive_scan(d_counts, m0, d_start, d_scratch, nullptr));
        KNN_CUDA(cudaMemset(d_counts, 0, (static_cast<size_t>(m0) + 2) * 4));
        bucket_scatter_kernel<<<pt_blocks, 256>>>(d_bucket_of, n, d_start, d_counts, d_perm);
        KNN_CUDA(cudaGetLastError());
        KnnLevel& L = k->dev.fine[j];
        L.start = d_start;
        L.perm = d_perm;
        L.mask = m0 - 1;
        L.shift = 0;
        L.h = k->h0 * std::ldexp(1.0, -(j + 1));
        L.inv_h = inv_h;
        L.sx = L.sy = L.sz = nullptr;
        const double cover = kRingSafety * L.h;
        L.cover2 = cover * cover;
    }
    {
        KnnLevel& L = k->dev.lv[n_grid];
        L.start = nullptr;
        L.perm = nullptr;
        L.mask = 0;
        L.shift = -1;
        L.cover2 = INFINITY;
        L.h = INFINITY;
        L.inv_h = k->dev.inv_h0;
        L.sx = L.sy = L.sz = nullptr;
    }
    k->dev.x = k->d_x;
    k->dev.y = k->d_y;
    k->dev.z = k->d_z;
    k->dev.n = n;
    KNN_CUDA(cudaDeviceSynchronize());
#undef KNN_TRY
#undef KNN_CUDA
    *out = k;
    return AT_OK;
}

static unsigned query_blocks(long long nq) {
    const long long want = (nq + 7) / 8;  // 8 warps per block
    const long long cap = static_cast<long long>(sm_count()) * 64;
    return static_cast<unsigned>(std::max(1ll, std::min(want, cap)));
}

static int launch_knn_query(const at_knn_t* k, const double* qx, const double* qy, const double* qz, int64_t nq, int kk,
                            double upper_bound, long long* idx_out, double* dist_out, uint8_t* tie_out,
                            const PeerOut& peers, cudaStream_t st) {
    const double ub2 = upper_bound * upper_bound;
    // near search: ring 1 of the two finest levels, one warp per query; whatever it cannot
    // decide (queries far from every source) goes to the per-thread tree search
    const int n_grid = k->dev.n_levels - 1;
    const int first_level = start_level(kk, n_grid);
    const int near_levels = std::min(first_level + 2, n_grid);  // exclusive
    // k > 1: 8 warps x kCandCap cached keys (float64 d2 + int32 index) = 48 KB per CTA
    const size_t cache_bytes = kk > 1 ? static_cast<size_t>(8) * kCandCap * (sizeof(double) + sizeof(int)) : 0;
    const unsigned blocks = query_blocks(nq);
    const int* list = nullptr;
    const unsigned int* list_n = nullptr;
    unsigned list_blocks = blocks;
    if (kk == 1 && tie_out == nullptr && k->d_range0 != nullptr && nq < (1ll << 31) - 64) {
        // one thread per query first; the warp kernel then finishes what that left over
        at_knn* mk = const_cast<at_knn*>(k);
        if (mk->pending_cap < nq) {
            if (mk->d_pending != nullptr) device_free(mk->d_pending);
            mk->d_pending = nullptr;
            mk->pending_cap = 0;
            AT_CUDA_TRY(device_alloc(reinterpret_cast<void**>(&mk->d_pending), static_cast<size_t>(nq) * sizeof(int)));
            mk->pending_cap = nq;
        }
        if (mk->d_n_pending == nullptr) AT_CUDA_TRY(device_alloc(reinterpret_cast<void**>(&mk->d_n_pending), 16));
        AT_CUDA_TRY(cudaMemsetAsync(mk->d_n_pending, 0, sizeof(unsigned int), st));
        const long long threads = ((nq + 31) / 32) * 32;
        const unsigned tblocks = static_cast<unsigned>((threads + 255) / 256);
        if (peers.n > 0)
            knn1_thread_kernel<true><<<tblocks, 256, 0, st>>>(k->dev, k->d_range0, k->d_rec0, qx, qy, qz, nq, ub2, idx_out, dist_out, peers, mk->d_pending, mk->d_n_pending);
        else
            knn1_thread_kernel<false><<<tblocks, 256, 0, st>>>(k->dev, k->d_range0, k->d_rec0, qx, qy, qz, nq, ub2, idx_out, dist_out, peers, mk->d_pending, mk->d_n_pending);
        AT_LAUNCH_CHECK("knn1_thread_kernel");
        list = mk->d_pending;
        list_n = mk->d_n_pending;
        list_blocks = std::min(blocks, static_cast<unsigned>(sm_count()) * 8u);  // grid-stride over *list_n
    }
    if (kk > 1 && peers.n > 0)
        knn_query_kernel<true, true><<<blocks, 256, cache_bytes, st>>>(k->dev, qx, qy, qz, nq, kk, ub2, first_level, near_levels, idx_out, dist_out, tie_out, peers, nullptr, nullptr);
    else if (kk > 1)
        knn_query_kernel<true, false><<<blocks, 256, cache_bytes, st>>>(k->dev, qx, qy, qz, nq, kk, ub2, first_level, near_levels, idx_out, dist_out, tie_out, peers, nullptr, nullptr);
    else if (peers.n > 0)
        knn_query_kernel<false, true><<<list_blocks, 256, 0, st>>>(k->dev, qx, qy, qz, nq, kk, ub2, first_level, near_levels, idx_out, dist_out, tie_out, peers, list, list_n);
    else
        knn_query_kernel<false, false><<<list_blocks, 256, 0, st>>>(k->dev, qx, qy, qz, nq, kk, ub2, first_level, near_levels, idx_out, dist_out, tie_out, peers, list, list_n);
    AT_LAUNCH_CHECK("knn_query_kernel");
    const long long tree_blocks = (nq + 127) / 128;
    AT_REQUIRE(tree_blocks < (1ll << 31), "at_knn_query: too many queries");
    if (peers.n > 0)
        knn_tree_kernel<true><<<static_cast<unsigned>(tree_blocks), 128, 0, st>>>(k->dev, qx, qy, qz, nq, kk, ub2, idx_out, dist_out, tie_out, peers);
    else
        knn_tree_kernel<false><<<static_cast<unsigned>(tree_blocks), 128, 0, st>>>(k->dev, qx, qy, qz, nq, kk, ub2, idx_out, dist_out, tie_out, peers);
    AT_LAUNCH_CHECK("knn_tree_kernel");
    return AT_OK;
}

extern "C" int at_knn_query(const at_knn_t* k, const double* qx, const double* qy, const double* qz,
                            int64_t nq, int kk, double upper_bound, int64_t* idx_out, double* dist_out,
                            uint8_t* tie_out, void* stream) {
    AT_REQUIRE(k != nullptr, "at_knn_query: null index");
    AT_REQUIRE(kk >= 1 && kk <= kMaxK, "at_knn_query: k must be in [1, %d] (k=%d)", kMaxK, kk);
    AT_REQUIRE(nq >= 0, "at_knn_query: negative query count");
    AT_REQUIRE(!(upper_bound != upper_bound) && upper_bound >= 0, "at_knn_query: bad distance_upper_bound");
    if (nq == 0) return AT_OK;  // an empty query set has null pointers (zero-length device arrays)
    AT_REQUIRE(qx != nullptr && qy != nullptr && qz != nullptr && idx_out != nullptr, "at_knn_query: null argument");
    PeerOut none;
    none.n = 0;
    return launch_knn_query(k, qx, qy, qz, nq, kk, upper_bound, reinterpret_cast<long long*>(idx_out), dist_out, tie_out, none,
                            as_stream(stream));
}

// ---- sharded query with the all-gather fused into the search (NVLink peer stores) --------
namespace {

struct PeerFlags {
    unsigned long long* ptr[kMaxPeers + 1];  // flag array (uint64[world]) of every rank, own included
    int world, rank;
};

// Thread r tells rank r "rank `rank` has finished epoch `epoch`" (release: the peer stores of the
// kernels before this one on the stream are visible first), then waits until rank r has said the
// same to us.  A rank that never arrives trips the timeout instead of hanging the GPU.
// `epoch` = 0: the epoch is this rank's own call counter, kept in device memory behind its flags
// (every rank makes the same sequence of calls), so the launch has no per-call arguments and a
// whole step can be replayed as a CUDA graph.
__device__ __forceinline__ unsigned long long next_epoch(unsigned long long epoch, unsigned long long* counter) {
    __shared__ unsigned long long shared_epoch;
    if (threadIdx.x == 0) shared_epoch = epoch != 0 ? epoch : ++(*counter);
    __syncthreads();
    return shared_epoch;
}

__global__ void peer_signal_wait_kernel(const PeerFlags f, unsigned long long epoch_arg, unsigned long long* __restrict__ counter,
                                        int* __restrict__ error) {
    const unsigned long long epoch = next_epoch(epoch_arg, counter);
    const int r = threadIdx.x;
    if (r >= f.world) return;
    __threadfence_system();
    unsigned long long* theirs = f.ptr[r] + f.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    const unsigned long long* mine = f.ptr[f.rank] + r;
    unsigned long long t0, now, seen;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
        if (seen >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > 2000000000ull) {  // 2 s
            *error = 1;
            break;
        }
    }
}

// The exchange as one coalesced pass after the search: every thread copies 16-byte pieces of
// this rank's slice into every peer's gather buffer (long contiguous NVLink writes instead of
// one 8-byte store per query and peer from inside the search, which at 8 GPUs tripled the
// search kernel's time); the last CTA to finish then runs the arrival-flag exchange.
__global__ void __launch_bounds__(256)
    peer_broadcast_kernel(const PeerOut peers, const long long* __restrict__ mine, long long n_items,
                          const PeerFlags f, unsigned long long epoch_arg, unsigned long long* __restrict__ counter,
                          unsigned int* __restrict__ done_ctas, int* __restrict__ error) {
    const long long n2 = n_items / 2;  // int64 pairs (the slice starts 16-byte aligned)
    const longlong2* src = reinterpret_cast<const longlong2*>(mine);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const longlong2 v = src[i];
        for (int p = 0; p < peers.n; ++p) reinterpret_cast<longlong2*>(peers.ptr[p])[i] = v;
    }
    if ((n_items & 1) && blockIdx.x == 0 && threadIdx.x == 0)
        for (int p = 0; p < peers.n; ++p) peers.ptr[p][n_items - 1] = mine[n_items - 1];
    __threadfence_system();
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(done_ctas, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *done_ctas = 0;  // ready for the next call on this stream
    const unsigned long long epoch = next_epoch(epoch_arg, counter);
    const int r = threadIdx.x;
    if (r >= f.world) return;
    __threadfence_system();
    unsigned long long* theirs = f.ptr[r] + f.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    const unsigned long long* own = f.ptr[f.rank] + r;
    unsigned long long t0, now, seen;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(own) : "memory");
        if (seen >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > 2000000000ull) {
            *error = 1;
            break;
        }
    }
}

}  // namespace

extern "C" int at_knn_query_gather(const at_knn_t* k, const double* qx, const double* qy, const double* qz,
                                   int64_t nq_local, int kk, double upper_bound, int64_t* const* gather_bufs,
                                   uint64_t* const* flag_bufs, int world, int rank, int64_t row_offset,
                                   double* dist_out, uint8_t* tie_out, uint64_t epoch, int32_t* error_flag,
                                   int exchange, void* stream) {
    AT_REQUIRE(k != nullptr && gather_bufs != nullptr && flag_bufs != nullptr && error_flag != nullptr,
               "at_knn_query_gather: null argument");
    AT_REQUIRE(world >= 1 && world <= kMaxPeers + 1 && rank >= 0 && rank < world, "at_knn_query_gather: bad rank %d of %d",
               rank, world);
    AT_REQUIRE(kk >= 1 && kk <= kMaxK, "at_knn_query_gather: k must be in [1, %d] (k=%d)", kMaxK, kk);
    AT_REQUIRE(nq_local >= 0 && row_offset >= 0, "at_knn_query_gather: negative size");
    AT_REQUIRE(!(upper_bound != upper_bound) && upper_bound >= 0, "at_knn_query_gather: bad distance_upper_bound");
    for (int r = 0; r < world; ++r)
        AT_REQUIRE(gather_bufs[r] != nullptr && flag_bufs[r] != nullptr, "at_knn_query_gather: rank %d has no buffer", r);
    AT_REQUIRE(exchange == AT_EXCHANGE_INLINE || exchange == AT_EXCHANGE_BULK, "at_knn_query_gather: unknown exchange mode %d", exchange);
    cudaStream_t st = as_stream(stream);
    PeerOut peers, none;
    peers.n = none.n = 0;
    for (int r = 0; r < world; ++r)
        if (r != rank) peers.ptr[peers.n++] = reinterpret_cast<long long*>(gather_bufs[r]) + row_offset * kk;
    long long* mine = reinterpret_cast<long long*>(gather_bufs[rank]) + row_offset * kk;
    const bool bulk = exchange == AT_EXCHANGE_BULK && world > 1 && (reinterpret_cast<uintptr_t>(mine) & 15) == 0;
    if (nq_local > 0) {
        AT_REQUIRE(qx != nullptr && qy != nullptr && qz != nullptr, "at_knn_query_gather: null queries");
        int rc = launch_knn_query(k, qx, qy, qz, nq_local, kk, upper_bound, mine, dist_out, tie_out, bulk ? none : peers, st);
        if (rc != AT_OK) return rc;
    }
    if (world > 1) {
        PeerFlags f;
        f.world = world;
        f.rank = rank;
        for (int r = 0; r < world; ++r) f.ptr[r] = reinterpret_cast<unsigned long long*>(flag_bufs[r]);
        // behind the uint64[world] flags of this rank's own buffer: the CTA counter of the bulk
        // exchange, then the call counter that serves as the epoch when the caller passes 0
        unsigned long long* epoch_counter = f.ptr[rank] + world + 1;
        if (bulk) {
            unsigned int* counter = reinterpret_cast<unsigned int*>(f.ptr[rank] + world);
            const long long n_items = nq_local * kk;
            const unsigned blocks = static_cast<unsigned>(std::max<long long>(1, std::min<long long>((n_items / 2 + 255) / 256, 4ll * sm_count())));
            peer_broadcast_kernel<<<blocks, 256, 0, st>>>(peers, mine, n_items, f, epoch, epoch_counter, counter, error_flag);
            AT_LAUNCH_CHECK("peer_broadcast_kernel");
        } else {
            peer_signal_wait_kernel<<<1, 32, 0, st>>>(f, epoch, epoch_counter, error_flag);
            AT_LAUNCH_CHECK("peer_signal_wait_kernel");
        }
    }
    return AT_OK;
}

// Device memory another process of this box can map (cudaIpc): the gather buffers and flags of
// at_knn_query_gather.  Zero-filled.  `handle` receives the 64-byte cudaIpcMemHandle_t.
extern "C" int at_peer_alloc(size_t bytes, void** ptr, void* handle) {
    AT_REQUIRE(ptr != nullptr && handle != nullptr && bytes > 0, "at_peer_alloc: bad arguments");
    *ptr = nullptr;
    void* p = nullptr;
    AT_CUDA_TRY(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        cudaGetLastError();
        return set_error(AT_ERR_CUDA, "at_peer_alloc: %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handle, &h, sizeof(h));
    *ptr = p;
    return AT_OK;
}

extern "C" int at_peer_free(void* ptr) {
    if (ptr != nullptr) AT_CUDA_TRY(cudaFree(ptr));
    return AT_OK;
}

// Map a buffer another rank made with at_peer_alloc (enables peer access to its GPU).
extern "C" int at_peer_open(const void* handle, void** ptr) {
    AT_REQUIRE(handle != nullptr && ptr != nullptr, "at_peer_open: null argument");
    *ptr = nullptr;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(AT_ERR_CUDA, "at_peer_open: %s", cudaGetErrorString(e));
    }
    return AT_OK;
}

extern "C" int at_peer_close(void* ptr) {
    if (ptr != nullptr) AT_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return AT_OK;
}

extern "C" int at_ball_mark(const at_knn_t* k, const double* qx, const double* qy, const double* qz,
                            int64_t nq, double r, uint8_t* mark, void* stream) {
    AT_REQUIRE(k != nullptr, "at_ball_mark: null index");
    AT_REQUIRE(nq >= 0 && r >= 0, "at_ball_mark: bad arguments");
    if (nq == 0) return AT_OK;  // an empty query set has null pointers (zero-length device arrays)
    AT_REQUIRE(qx != nullptr && qy != nullptr && qz != nullptr && mark != nullptr, "at_ball_mark: null argument");
    const double r2 = r * r;
    // the finest level (fine levels included, index -1 - j) whose ring-1 covers the radius
    int level = k->dev.n_levels - 1;
    for (int l = -k->dev.n_fine; l < k->dev.n_levels; ++l) {
        const KnnLevel& L = l < 0 ? k->dev.fine[-1 - l] : k->dev.lv[l];
        if (r2 < L.cover2) {
            level = l;
            break;
        }
    }
    ball_mark_kernel<<<query_blocks(nq), 256, 0, as_stream(stream)>>>(k->dev, level, qx, qy, qz, nq, r2, mark);
    AT_LAUNCH_CHECK("ball_mark_kernel");
    return AT_OK;
}

extern "C" int at_min_nn_distance(const at_knn_t* k, int64_t first, int64_t count, double* out_host,
                                  void* stream) {
    AT_REQUIRE(k != nullptr && out_host != nullptr, "at_min_nn_distance: null argument");
    AT_REQUIRE(first >= 0 && first <= k->n, "at_min_nn_distance: first out of range");
    const long long end = count < 0 ? k->n : std::min<long long>(k->n, first + count);
    if (end <= first) {
        *out_host = INFINITY;
        return AT_OK;
    }
    unsigned long long* d_bits = nullptr;
    AT_CUDA_TRY(device_alloc(reinterpret_cast<void**>(&d_bits), 8));
    const double inf = INFINITY;
    unsigned long long init;
    memcpy(&init, &inf, 8);
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemcpyAsync(d_bits, &init, 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        min_second_nn_kernel<<<query_blocks(end - first), 256, 0, st>>>(k->dev, first, end, d_bits);
        e = cudaGetLastError();
    }
    unsigned long long bits = init;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bits, d_bits, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    device_free(d_bits);
    if (e != cudaSuccess) return set_error(AT_ERR_CUDA, "at_min_nn_distance: %s", cudaGetErrorString(e));
    memcpy(out_host, &bits, 8);
    return AT_OK;
}
