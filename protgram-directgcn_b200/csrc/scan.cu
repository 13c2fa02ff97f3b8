// Device-wide exclusive scan (int64) -- reduce-then-scan over 2048-element tiles, recursing on
// the tile sums.  Used by the builder (node ids, edge compaction) and the graph kernels.
#include <stdarg.h>

#include "common.cuh"

// ---------------------------------------------------------------------------- error plumbing
static thread_local char g_err[512] = "";
void pg_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char *pg_last_error(void) { return g_err; }
extern "C" int pg_version(void) { return 100; }
static unsigned long long g_launches = 0;
void pg_count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" unsigned long long pg_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// ---------------------------------------------------------------------------- scan
namespace {
constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

__device__ __forceinline__ int64_t warp_inclusive_scan(int64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum.
__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total) {
    __shared__ int64_t warp_sums[kThreads / 32];
    __shared__ int64_t block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t inc = warp_inclusive_scan(v, lane);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < kThreads / 32 ? warp_sums[lane] : 0;
        int64_t winc = warp_inclusive_scan(w, lane);
        if (lane < kThreads / 32) warp_sums[lane] = winc - w;
        if (lane == kThreads / 32 - 1) block_total = winc;
    }
    __syncthreads();
    *total = block_total;
    int64_t r = inc - v + warp_sums[warp];
    __syncthreads();  // smem reused by the next call
    return r;
}

__global__ void __launch_bounds__(kThreads) tile_reduce_kernel(const int64_t *__restrict__ in, int64_t n,
                                                               int64_t *__restrict__ tile_sums) {
    const int64_t base = (int64_t)blockIdx.x * kTile;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        int64_t i = base + (int64_t)k * kThreads + threadIdx.x;  // striped: coalesced
        if (i < n) s += in[i];
    }
    int64_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kThreads) tile_scan_kernel(const int64_t *__restrict__ in, int64_t *__restrict__ out,
                                                             int64_t n, const int64_t *__restrict__ tile_offsets,
                                                             int64_t *__restrict__ d_total) {
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;  // blocked
    int64_t v[kItems];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int64_t total;
    int64_t run = block_exclusive_scan(s, &total) + (tile_offsets ? tile_offsets[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (d_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
        *d_total = total + (tile_offsets ? tile_offsets[blockIdx.x] : 0);
}
}  // namespace

size_t pg_scan_ws_bytes(int64_t n) {
    size_t bytes = 0;
    while (n > kTile) {
        n = pg_ceil_div(n, kTile);
        bytes += pg_align_up((size_t)n * sizeof(int64_t), 256);
    }
    return bytes + 256;
}

int pg_exclusive_scan_i64(const int64_t *d_in, int64_t *d_out, int64_t n, int64_t *d_total, void *d_ws,
                          size_t ws_bytes, cudaStream_t stream) {
    if (n <= 0) {
        if (d_total) PG_CUDA_CALL(cudaMemsetAsync(d_total, 0, sizeof(int64_t), stream));
        return PG_OK;
    }
    const int64_t tiles = pg_ceil_div(n, kTile);
    if (tiles == 1) {
        tile_scan_kernel<<<1, kThreads, 0, stream>>>(d_in, d_out, n, nullptr, d_total);
        PG_CUDA_LAUNCH_CHECK("tile_scan_kernel");
        return PG_OK;
    }
    PgArena arena(d_ws, ws_bytes);
    int64_t *tile_sums = arena.take<int64_t>((size_t)tiles);
    if (!arena.ok) {
        pg_set_error("exclusive_scan: workspace too small (%zu bytes)", ws_bytes);
        return PG_EWORKSPACE;
    }
    tile_reduce_kernel<<<(unsigned)tiles, kThreads, 0, stream>>>(d_in, n, tile_sums);
    PG_CUDA_LAUNCH_CHECK("tile_reduce_kernel");
    int rc = pg_exclusive_scan_i64(tile_sums, tile_sums, tiles, nullptr, arena.base + arena.used,
                                   arena.cap - arena.used, stream);
    if (rc != PG_OK) return rc;
    tile_scan_kernel<<<(unsigned)tiles, kThreads, 0, stream>>>(d_in, d_out, n, tile_sums, d_total);
    PG_CUDA_LAUNCH_CHECK("tile_scan_kernel");
    return PG_OK;
}
