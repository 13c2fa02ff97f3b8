// Stable LSD radix sort: 64-bit keys + 32-bit payload, 8-bit digits.
//
// Per pass: (1) per-tile digit histograms, (2) exclusive scan of the digit-major
// [256 x tiles] table, (3) stable scatter.  Ranking inside a tile is warp-level: every warp
// owns a contiguous slab of the tile, ranks its 32 keys per step with __match_any_sync, and
// the per-warp digit counters are prefix-summed across warps, so equal digits keep their
// input order (that is what lets one (row*N+col) sort double as `coalesce()` ordering and
// lets the edge->CSR grouping keep the reference's accumulation order).
#include "common.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;  // 4096 keys per CTA
constexpr int kRadix = 256;

__device__ __forceinline__ unsigned digit_of(unsigned long long key, int shift) {
    return (unsigned)(key >> shift) & (kRadix - 1);
}

__global__ void __launch_bounds__(kThreads) rs_hist_kernel(const unsigned long long *__restrict__ keys, int64_t n,
                                                           int shift, int64_t tiles, int64_t *__restrict__ tile_hist) {
    __shared__ unsigned hist[kRadix];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        int64_t i = base + (int64_t)k * kThreads + threadIdx.x;
        if (i < n) atomicAdd(&hist[digit_of(keys[i], shift)], 1u);
    }
    __syncthreads();
    tile_hist[(int64_t)threadIdx.x * tiles + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(kThreads) rs_scatter_kernel(const unsigned long long *__restrict__ keys_in,
                                                              const uint32_t *__restrict__ vals_in,
                                                              unsigned long long *__restrict__ keys_out,
                                                              uint32_t *__restrict__ vals_out, int64_t n, int shift,
                                                              int64_t tiles, const int64_t *__restrict__ tile_off) {
    __shared__ unsigned wcount[kWarps][kRadix];
    __shared__ int64_t goff[kRadix];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kWarps * kRadix; i += kThreads) (&wcount[0][0])[i] = 0;
    goff[threadIdx.x] = tile_off[(int64_t)threadIdx.x * tiles + blockIdx.x];
    __syncthreads();

    // warp `warp` owns keys [base, base + 32*kItems) of the tile, step k covers 32 consecutive keys
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)warp * (32 * kItems);
    unsigned long long key[kItems];
    unsigned rank[kItems];
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int64_t i = base + k * 32 + lane;
        key[k] = (i < n) ? keys_in[i] : ~0ull;  // padding sorts last and is never written
    }
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const unsigned d = digit_of(key[k], shift);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (lane == leader) {
            old = wcount[warp][d];
            wcount[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    {  // exclusive prefix over warps, one digit per thread
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            unsigned c = wcount[w][threadIdx.x];
            wcount[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int64_t i = base + k * 32 + lane;
        if (i < n) {
            const unsigned d = digit_of(key[k], shift);
            const int64_t pos = goff[d] + wcount[warp][d] + rank[k];
            keys_out[pos] = key[k];
            vals_out[pos] = vals_in[i];
        }
    }
}
}  // namespace

extern "C" size_t pg_sort_pairs_ws_bytes(int64_t n) {
    if (n <= 0) return 256;
    const int64_t tiles = pg_ceil_div(n, kTile);
    return pg_align_up((size_t)tiles * kRadix * sizeof(int64_t), 256) + pg_scan_ws_bytes(tiles * kRadix) + 256;
}

extern "C" int pg_sort_pairs(unsigned long long *d_keys, unsigned long long *d_keys_alt, uint32_t *d_vals,
                             uint32_t *d_vals_alt, int64_t n, int key_bits, void *d_ws, size_t ws_bytes,
                             pg_stream_t stream_) {
    cudaStream_t stream = pg_cu(stream_);
    PG_CHECK_ARG(n >= 0 && key_bits >= 0 && key_bits <= 64, "pg_sort_pairs: bad n=%lld key_bits=%d", (long long)n,
                 key_bits);
    if (n <= 1 || key_bits == 0) return PG_OK;
    PG_CHECK_ARG(d_keys && d_keys_alt && d_vals && d_vals_alt, "pg_sort_pairs: null buffer");
    const int64_t tiles = pg_ceil_div(n, kTile);
    PgArena arena(d_ws, ws_bytes);
    int64_t *tile_hist = arena.take<int64_t>((size_t)tiles * kRadix);
    if (!arena.ok || ws_bytes < pg_sort_pairs_ws_bytes(n)) {
        pg_set_error("pg_sort_pairs: workspace too small (%zu < %zu)", ws_bytes, pg_sort_pairs_ws_bytes(n));
        return PG_EWORKSPACE;
    }
    unsigned long long *kin = d_keys, *kout = d_keys_alt;
    uint32_t *vin = d_vals, *vout = d_vals_alt;
    const int passes = (key_bits + 7) / 8;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * 8;
        rs_hist_kernel<<<(unsigned)tiles, kThreads, 0, stream>>>(kin, n, shift, tiles, tile_hist);
        PG_CUDA_LAUNCH_CHECK("rs_hist_kernel");
        int rc = pg_exclusive_scan_i64(tile_hist, tile_hist, tiles * kRadix, nullptr, arena.base + arena.used,
                                       arena.cap - arena.used, stream);
        if (rc != PG_OK) return rc;
        rs_scatter_kernel<<<(unsigned)tiles, kThreads, 0, stream>>>(kin, vin, kout, vout, n, shift, tiles, tile_hist);
        PG_CUDA_LAUNCH_CHECK("rs_scatter_kernel");
        unsigned long long *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    if (kin != d_keys) {
        PG_CUDA_CALL(cudaMemcpyAsync(d_keys, kin, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
        PG_CUDA_CALL(cudaMemcpyAsync(d_vals, vin, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    }
    return PG_OK;
}
