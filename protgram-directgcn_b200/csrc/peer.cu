// Halo exchange of a row-partitioned graph over NVLink peer memory (SURVEY.md 8(e), row "Propagation"; one process per GPU on
// one NVSwitch node).  The NCCL form of the exchange (host/partitioned.py) packs the rows a peer asked for into a send buffer
// (pg_gather_rows) and lets all_to_all_single move it; here ONE kernel does both: it reads the rows straight out of X and
// stores them into the peers' receive buffers through CUDA-IPC mapped pointers -- no send buffer, no second pass over the
// rows, no collective launch -- then publishes an epoch flag on every peer (release at system scope).  The consumer side is
// a one-CTA kernel in front of the SpMM that waits (acquire, bounded spin) until every peer's flag has reached the epoch.
//
// Buffers: every rank owns a ring of receive slots; exchange number e uses slot e % ring.  The exchanges are in lockstep (a
// rank can post exchange e + 1 only after it consumed e, which needs every peer's push of e), so a ring of 4 is enough for
// two exchanges posted per consume phase (dY rows + gate rows in the layer's backward); host/partitioned.py uses 4.
#include "common.cuh"

namespace {

struct PushArgs {
    float *dst[PG_MAX_PEERS];        // peer p: base of the slot region that receives THIS rank's rows (already offset)
    uint32_t *flag[PG_MAX_PEERS];    // peer p: the flag word this rank sets there
    int64_t row_begin[PG_MAX_PEERS + 1];   // rows [row_begin[p], row_begin[p+1]) of d_idx go to peer p
    int64_t rot_begin[PG_MAX_PEERS + 1];   // the same row counts, prefix-summed in ROTATED peer order (self + 1, self + 2, ...)
    int world, self;
};

// grid-stride over (row, float4) items; the last CTA to finish publishes the epoch to every peer
template <bool VEC>
__global__ void __launch_bounds__(256) halo_push_kernel(const float *__restrict__ src, int64_t ld_src, const int64_t *__restrict__ idx,
                                                        PushArgs a, int w, int64_t ld_dst, uint32_t epoch, unsigned int *done_counter) {
    const int per_row = VEC ? (w >> 2) : w;
    const int64_t total = a.rot_begin[a.world] * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // work is laid out in rotated peer order: rank r's first CTAs write to peer r + 1, the next ones to r + 2, ... so that at any
        // moment the ranks store into DIFFERENT peers (all ranks starting with peer 0 made one GPU's ingress the bottleneck:
        // 247 GB/s per GPU at 8 ranks against 585 GB/s at 2)
        const int64_t rr = i / per_row;
        const int c = (int)(i - rr * per_row);
        int q = 0;
#pragma unroll
        for (int t = 1; t < PG_MAX_PEERS; ++t) q += (t < a.world && rr >= a.rot_begin[t]) ? 1 : 0;
        int p = a.self + 1 + q;
        if (p >= a.world) p -= a.world;
        const int64_t r = a.row_begin[p] + (rr - a.rot_begin[q]);
        const int64_t s = __ldg(idx + r);
        float *drow = a.dst[p] + (r - a.row_begin[p]) * ld_dst;
        if (VEC) reinterpret_cast<float4 *>(drow)[c] = __ldg(reinterpret_cast<const float4 *>(src + s * ld_src) + c);
        else drow[c] = __ldg(src + s * ld_src + c);
    }
    __threadfence_system();                       // this thread's peer stores are visible system-wide before the counter moves
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1) {              // every CTA has fenced its stores
            *done_counter = 0u;                   // re-armed for the next push on this stream
            __threadfence_system();
            for (int p = 0; p < a.world; ++p)
                if (a.flag[p] != nullptr) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flag[p]), "r"(epoch) : "memory");
        }
    }
}

// one CTA, thread p watches peer p's flag; bounded: ~4 s of SM clock, then the error word is raised instead of hanging the GPU
__global__ void __launch_bounds__(32) halo_wait_kernel(const uint32_t *__restrict__ flags, int world, int self, uint32_t epoch,
                                                       int *__restrict__ error_flag) {
    const int p = threadIdx.x;
    if (p >= world || p == self) return;
    const long long t0 = clock64();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + p) : "memory");
        if ((int32_t)(v - epoch) >= 0) return;    // epochs only grow (wrap-safe comparison)
        if (clock64() - t0 > 8000000000LL) {
            atomicExch(error_flag, 1 + p);
            return;
        }
        __nanosleep(200);
    }
}
}  // namespace

extern "C" int pg_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle64) {
    PG_CHECK_ARG(d_ptr && handle64 && bytes > 0, "pg_peer_alloc: bad argument");
    PG_CUDA_CALL(cudaMalloc(d_ptr, bytes));
    PG_CUDA_CALL(cudaMemset(*d_ptr, 0, bytes));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    PG_CUDA_CALL(cudaIpcGetMemHandle(&h, *d_ptr));
    memcpy(handle64, &h, 64);
    return PG_OK;
}

extern "C" int pg_peer_open(const unsigned char *handle64, void **d_ptr) {
    PG_CHECK_ARG(d_ptr && handle64, "pg_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    PG_CUDA_CALL(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PG_OK;
}

extern "C" int pg_peer_close(void *d_ptr) {
    if (d_ptr) PG_CUDA_CALL(cudaIpcCloseMemHandle(d_ptr));
    return PG_OK;
}

extern "C" int pg_peer_free(void *d_ptr) {
    if (d_ptr) PG_CUDA_CALL(cudaFree(d_ptr));
    return PG_OK;
}

extern "C" int pg_halo_push(const float *d_src, int64_t ld_src, const int64_t *d_idx, const int64_t *h_row_begin, float *const *h_peer_dst,
                            uint32_t *const *h_peer_flag, int world, int self, int w, int64_t ld_dst, uint32_t epoch,
                            unsigned int *d_done_counter, pg_stream_t stream) {
    PG_CHECK_ARG(world >= 1 && world <= PG_MAX_PEERS && self >= 0 && self < world && w >= 1 && ld_src >= w && ld_dst >= w,
                 "pg_halo_push: bad shape (world <= %d)", PG_MAX_PEERS);
    PG_CHECK_ARG(h_row_begin && h_peer_dst && h_peer_flag && d_done_counter, "pg_halo_push: null argument");
    PushArgs a;
    a.world = world;
    bool vec = w % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 && (((uintptr_t)d_src) & 15) == 0;
    for (int p = 0; p < PG_MAX_PEERS; ++p) {
        a.dst[p] = p < world ? h_peer_dst[p] : nullptr;
        a.flag[p] = p < world ? h_peer_flag[p] : nullptr;
        if (p < world && a.dst[p] && (((uintptr_t)a.dst[p]) & 15) != 0) vec = false;
    }
    for (int p = 0; p <= PG_MAX_PEERS; ++p) a.row_begin[p] = h_row_begin[p < world ? p : world];
    a.self = self;
    a.rot_begin[0] = 0;
    for (int q = 0; q < PG_MAX_PEERS; ++q) {
        const int p = (self + 1 + q) % world;
        a.rot_begin[q + 1] = a.rot_begin[q] + (q < world ? a.row_begin[p + 1] - a.row_begin[p] : 0);
    }
    const int64_t rows = a.row_begin[world];
    PG_CHECK_ARG(rows >= 0 && (rows == 0 || (d_src && d_idx)), "pg_halo_push: null buffer");
    for (int p = 0; p < world; ++p)
        PG_CHECK_ARG(a.row_begin[p + 1] >= a.row_begin[p] && (a.row_begin[p + 1] == a.row_begin[p] || a.dst[p]), "pg_halo_push: peer %d has rows but no buffer", p);
    const int64_t total = rows * (vec ? w / 4 : w);
    int64_t want = pg_ceil_div(total, 256);
    const int64_t cap = (int64_t)PG_NUM_SMS * 8;
    const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
    cudaStream_t st = pg_cu(stream);
    if (vec) halo_push_kernel<true><<<grid, 256, 0, st>>>(d_src, ld_src, d_idx, a, w, ld_dst, epoch, d_done_counter);
    else halo_push_kernel<false><<<grid, 256, 0, st>>>(d_src, ld_src, d_idx, a, w, ld_dst, epoch, d_done_counter);
    PG_CUDA_LAUNCH_CHECK("halo_push_kernel");
    return PG_OK;
}

extern "C" int pg_halo_wait(const uint32_t *d_flags, int world, int self, uint32_t epoch, int *d_error_flag, pg_stream_t stream) {
    PG_CHECK_ARG(d_flags && d_error_flag && world >= 1 && world <= PG_MAX_PEERS && self >= 0 && self < world, "pg_halo_wait: bad argument");
    halo_wait_kernel<<<1, 32, 0, pg_cu(stream)>>>(d_flags, world, self, epoch, d_error_flag);
    PG_CUDA_LAUNCH_CHECK("halo_wait_kernel");
    return PG_OK;
}
