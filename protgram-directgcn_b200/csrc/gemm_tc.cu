// Hot path B, dense transform on the 5th-gen tensor cores (tcgen05 + TMEM), used when the hidden
// width makes it a real contraction (F_out >= 128 by default; north_star).
//
//   H = leaky_relu( A_ext @ W_ext (+X) + constant )        same contract as layer_gemm_fwd (gemm.cu)
//
// Precision: the parity bar is 1e-4 in fp32, which single-pass TF32 (10-bit mantissa) misses.  Each
// operand is split x = hi + lo with hi = x truncated to TF32 and lo = x - hi (exact in fp32), and
//   D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi        (3 x kind::tf32 MMAs, fp32 accumulate in TMEM)
// which leaves ~2^-21 relative error per product.
//
// Structure (one CTA = 128 threads = one 128 x N output tile, N = F_out <= 256 TMEM columns):
//   * A_ext is virtual (gates, [Z|X|gate columns]); the 128 threads build the 128 x 32 k-tile in
//     shared memory themselves: coalesced float4 global reads, gate, hi/lo split, 16-byte stores
//     into the canonical K-major / no-swizzle UMMA layout (core matrix = 8 rows x 16 B;
//     SBO = 128 B between 8-row groups, LBO = (rows+1)*16 B between K chunks -- the +1 row of
//     padding makes the transposing stores bank-conflict free).
//   * W_ext is pre-split / pre-transposed once per call into that same tile image (wprep kernel),
//     so the B tile is a straight coalesced copy.
//   * two shared-memory stages; one elected thread issues the 12 tcgen05.mma of a k-tile and
//     tcgen05.commit's them to the stage's mbarrier, so the loads of tile k+1 overlap the MMAs of
//     tile k.  Epilogue: tcgen05.ld (32 lanes x 16 columns per warp) -> +X, +constant, leaky_relu.
//   * every mbarrier wait is bounded (watchdog): a wrong descriptor must fail a test, not hang a GPU.
#include "common.cuh"

namespace {

// ---- the same virtual A operand as gemm.cu (kept in sync by tests/test_gpu_parity.py) ----
struct AExtTc {
    const float *z, *x, *ga, *gb, *gc;
    int gate_stride;
    int64_t ldz, ldx, M;
    int F_in, has_res, k_data, k_ext;

    __device__ __forceinline__ float gate(int seg, int64_t i) const {
        const float *g = seg == 0 ? ga : (seg == 1 ? gb : gc);
        return g[i * gate_stride];
    }
    __device__ __forceinline__ float at(int64_t i, int k) const {
        if (i >= M || k >= k_ext) return 0.f;
        if (k < 3 * F_in) return z[i * ldz + k] * gate(k / F_in, i);
        if (k < k_data) return x[i * ldx + (k - 3 * F_in)];
        const int j = k - k_data;
        return j < 3 ? gate(j, i) : 1.f;
    }
    __device__ __forceinline__ float4 at4(int64_t i, int k0) const {  // F_in % 4 == 0, 16 B aligned rows
        if (i < M && k0 + 3 < k_data) {
            if (k0 < 3 * F_in) {
                float4 v = __ldg(reinterpret_cast<const float4 *>(z + i * ldz + k0));
                const float g = gate(k0 / F_in, i);
                return make_float4(v.x * g, v.y * g, v.z * g, v.w * g);
            }
            return __ldg(reinterpret_cast<const float4 *>(x + i * ldx + (k0 - 3 * F_in)));
        }
        return make_float4(at(i, k0), at(i, k0 + 1), at(i, k0 + 2), at(i, k0 + 3));
    }
};

constexpr int TC_BM = 128;      // rows per CTA = UMMA M
constexpr int TC_BK = 32;       // k per stage = 8 chunks of 16 B = 4 MMAs of K = 8
constexpr int TC_CHUNKS = 8;
constexpr int TC_THREADS = 128;
constexpr uint32_t TC_LBO_A = (TC_BM + 1) * 16;

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start [0,14) | LBO [16,30) | SBO [32,46) | version=1 [46,48) | layout NONE [61,64)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// bounded wait: returns false when the watchdog expires (caller flags the error and bails out)
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}

// W_ext [k_ext, F_out] -> per k-tile image [kt][half(hi,lo)][chunk 0..7][n 0..F_out-1] of float4 (4 consecutive k)
__global__ void __launch_bounds__(256) wprep_kernel(const float *__restrict__ w_ext, int k_ext, int F_out, int k_tiles,
                                                    float4 *__restrict__ wp) {
    const int64_t total = (int64_t)k_tiles * TC_CHUNKS * F_out;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(t % F_out);
        const int kc = (int)((t / F_out) % TC_CHUNKS);
        const int kt = (int)(t / ((int64_t)F_out * TC_CHUNKS));
        float v[4], h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = kt * TC_BK + kc * 4 + j;
            v[j] = k < k_ext ? w_ext[(int64_t)k * F_out + n] : 0.f;
            h[j] = tf32_hi(v[j]);
            l[j] = v[j] - h[j];
        }
        const int64_t base = (int64_t)kt * 2 * TC_CHUNKS * F_out;
        wp[base + (int64_t)kc * F_out + n] = make_float4(h[0], h[1], h[2], h[3]);
        wp[base + (int64_t)(TC_CHUNKS + kc) * F_out + n] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1) layer_gemm_fwd_tc_kernel(AExtTc A, const float4 *__restrict__ wp, int F_out, int k_tiles,
                                                                          const float *__restrict__ constant, int64_t ldconst,
                                                                          int add_identity, float slope, float *__restrict__ h,
                                                                          int64_t ldh, int *__restrict__ error_flag) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mma_done[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ int bail;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
    const uint32_t lbo_b = (uint32_t)(F_out + 1) * 16u;
    const uint32_t a_bytes = TC_CHUNKS * TC_LBO_A, b_bytes = TC_CHUNKS * lbo_b;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;  // A_hi | A_lo | B_hi | B_lo
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < F_out) tmem_cols <<= 1;

    if (tid == 0) {
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        bail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // one warp allocates the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    // instruction descriptor: D = f32, A = B = tf32, K-major both, N = F_out, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(F_out >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

    for (int kt = 0; kt < k_tiles; ++kt) {
        const int s = kt & 1;
        uint8_t *st = smem + (size_t)s * stage_bytes;
        if (kt >= 2) {  // the MMAs that read this stage two tiles ago must have retired
            if (!mbar_wait(&mma_done[s], (uint32_t)(((kt >> 1) - 1) & 1))) bail = 1;
        }
        // ---- A tile: 128 rows x 8 chunks, gate + hi/lo split on the way in
        const int k0 = kt * TC_BK;
#pragma unroll
        for (int i = 0; i < (TC_BM * TC_CHUNKS) / TC_THREADS; ++i) {
            const int idx = i * TC_THREADS + tid;
            const int r = idx >> 3, kc = idx & 7;
            const float4 v = A.at4(m0 + r, k0 + kc * 4);
            const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
            *reinterpret_cast<float4 *>(st + kc * TC_LBO_A + r * 16) = hi;
            *reinterpret_cast<float4 *>(st + a_bytes + kc * TC_LBO_A + r * 16) = lo;
        }
        // ---- B tile: straight copy of the pre-split image
        const float4 *src = wp + (int64_t)kt * 2 * TC_CHUNKS * F_out;
        for (int idx = tid; idx < 2 * TC_CHUNKS * F_out; idx += TC_THREADS) {
            const int n = idx % F_out, kc = (idx / F_out) % TC_CHUNKS, half = idx / (F_out * TC_CHUNKS);
            *reinterpret_cast<float4 *>(st + 2 * a_bytes + half * b_bytes + kc * lbo_b + n * 16) = __ldg(src + idx);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0 && !bail) {
                const uint32_t a_hi = smem_u32(st), a_lo = a_hi + a_bytes, b_hi = a_hi + 2 * a_bytes, b_lo = b_hi + b_bytes;
#pragma unroll
                for (int ks = 0; ks < TC_BK / 8; ++ks) {  // one MMA consumes 2 chunks (K = 8 tf32)
                    const uint64_t dah = umma_desc(a_hi + ks * 2 * TC_LBO_A, TC_LBO_A, 128);
                    const uint64_t dal = umma_desc(a_lo + ks * 2 * TC_LBO_A, TC_LBO_A, 128);
                    const uint64_t dbh = umma_desc(b_hi + ks * 2 * lbo_b, lbo_b, 128);
                    const uint64_t dbl = umma_desc(b_lo + ks * 2 * lbo_b, lbo_b, 128);
                    umma_tf32(tmem_d, dah, dbh, idesc, (kt | ks) != 0);
                    umma_tf32(tmem_d, dah, dbl, idesc, 1);
                    umma_tf32(tmem_d, dal, dbh, idesc, 1);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done[s]))
                             : "memory");
            }
            __syncwarp();
        }
    }
    // ---- wait for the last commit (commits retire in order), then the epilogue
    {
        const int last = k_tiles - 1;
        if (!mbar_wait(&mma_done[last & 1], (uint32_t)((last >> 1) & 1))) bail = 1;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (bail) {
        if (tid == 0) atomicExch(error_flag, 1);
    } else {
        const int64_t row = m0 + warp * 32 + lane;
        for (int c0 = 0; c0 < F_out; c0 += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < A.M) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float y[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = c0 + q * 4 + j;
                        float v = __uint_as_float(r[q * 4 + j]);
                        if (add_identity) v += A.x[row * A.ldx + c];
                        if (constant) v += constant[row * ldconst + c];
                        if (slope != 1.f) v = v > 0.f ? v : v * slope;
                        y[j] = v;
                    }
                    *reinterpret_cast<float4 *>(h + row * ldh + c0 + q * 4) = make_float4(y[0], y[1], y[2], y[3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}

inline bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }
inline int k_tiles_of(int F_in, int has_res) {
    const int k_ext = 3 * F_in + (has_res ? F_in : 0) + 3 + (has_res ? 1 : 0);
    return (k_ext + TC_BK - 1) / TC_BK;
}
}  // namespace

extern "C" int pg_layer_gemm_fwd_tc_supported(int F_in, int F_out) {
    return F_in % 4 == 0 && F_out % 16 == 0 && F_out >= 16 && F_out <= 256;
}

extern "C" size_t pg_layer_gemm_fwd_tc_ws_bytes(int F_in, int F_out, int has_res) {
    return (size_t)k_tiles_of(F_in, has_res) * 2 * TC_CHUNKS * F_out * sizeof(float4) + 256;
}

extern "C" int pg_layer_gemm_fwd_tc(const float *d_z, int64_t ldz, const float *d_x, int64_t ldx, const float *d_gate_a,
                                    const float *d_gate_b, const float *d_gate_c, int gate_stride, const float *d_w_ext,
                                    const float *d_constant, int64_t ldconst, int64_t num_rows, int F_in, int F_out, int has_res,
                                    int add_identity, float slope, float *d_h, int64_t ldh, void *d_ws, size_t ws_bytes,
                                    pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F_in >= 1 && F_out >= 1 && (gate_stride == 0 || gate_stride == 1), "pg_layer_gemm_fwd_tc: bad shape");
    PG_CHECK_ARG(pg_layer_gemm_fwd_tc_supported(F_in, F_out), "pg_layer_gemm_fwd_tc: needs F_in %% 4 == 0, F_out %% 16 == 0, F_out <= 256");
    PG_CHECK_ARG(!(has_res && add_identity) && (!add_identity || F_in == F_out), "pg_layer_gemm_fwd_tc: bad residual mode");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_z && d_gate_a && d_gate_b && d_gate_c && d_w_ext && d_h && d_ws, "pg_layer_gemm_fwd_tc: null buffer");
    PG_CHECK_ARG(!(has_res || add_identity) || d_x, "pg_layer_gemm_fwd_tc: residual needs x");
    PG_CHECK_ARG(al16(d_z) && ldz % 4 == 0 && (!d_x || (al16(d_x) && ldx % 4 == 0)) && al16(d_h) && ldh % 4 == 0 && al16(d_ws),
                 "pg_layer_gemm_fwd_tc: operands must be 16-byte aligned with row strides %% 4 == 0");
    PG_CHECK_ARG(ldz >= 3 * (int64_t)F_in && ldh >= F_out && (!d_constant || ldconst >= F_out), "pg_layer_gemm_fwd_tc: bad stride");
    const int kt = k_tiles_of(F_in, has_res);
    const size_t need = pg_layer_gemm_fwd_tc_ws_bytes(F_in, F_out, has_res);
    if (ws_bytes < need) {
        pg_set_error("pg_layer_gemm_fwd_tc: workspace too small (%zu < %zu)", ws_bytes, need);
        return PG_EWORKSPACE;
    }
    cudaStream_t st = pg_cu(stream);
    AExtTc A;
    A.z = d_z; A.x = d_x; A.ga = d_gate_a; A.gb = d_gate_b; A.gc = d_gate_c; A.gate_stride = gate_stride;
    A.ldz = ldz; A.ldx = ldx; A.M = num_rows; A.F_in = F_in; A.has_res = has_res;
    A.k_data = 3 * F_in + (has_res ? F_in : 0);
    A.k_ext = A.k_data + 3 + (has_res ? 1 : 0);
    float4 *wp = reinterpret_cast<float4 *>(d_ws);
    int *err = reinterpret_cast<int *>(reinterpret_cast<char *>(d_ws) + need - 256);
    PG_CUDA_CALL(cudaMemsetAsync(err, 0, sizeof(int), st));
    {
        const int64_t total = (int64_t)kt * TC_CHUNKS * F_out;
        wprep_kernel<<<(unsigned)pg_ceil_div(total, 256), 256, 0, st>>>(d_w_ext, A.k_ext, F_out, kt, wp);
        PG_CUDA_LAUNCH_CHECK("wprep_kernel");
    }
    const size_t stage = 2 * (size_t)TC_CHUNKS * TC_LBO_A + 2 * (size_t)TC_CHUNKS * (F_out + 1) * 16;
    const size_t smem = 2 * stage;
    static bool attr_set = false;
    if (!attr_set) {
        PG_CUDA_CALL(cudaFuncSetAttribute(layer_gemm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    layer_gemm_fwd_tc_kernel<<<(unsigned)pg_ceil_div(num_rows, TC_BM), TC_THREADS, smem, st>>>(A, wp, F_out, kt, d_constant, ldconst,
                                                                                              add_identity, slope, d_h, ldh, err);
    PG_CUDA_LAUNCH_CHECK("layer_gemm_fwd_tc_kernel");
    return PG_OK;
}

// reads back the watchdog flag of the last pg_layer_gemm_fwd_tc call on this workspace (host sync; tests only)
extern "C" int pg_layer_gemm_fwd_tc_check(const void *d_ws, int F_in, int F_out, int has_res, pg_stream_t stream) {
    const size_t need = pg_layer_gemm_fwd_tc_ws_bytes(F_in, F_out, has_res);
    int flag = 0;
    PG_CUDA_CALL(cudaMemcpyAsync(&flag, reinterpret_cast<const char *>(d_ws) + need - 256, sizeof(int), cudaMemcpyDeviceToHost, pg_cu(stream)));
    PG_CUDA_CALL(cudaStreamSynchronize(pg_cu(stream)));
    if (flag) {
        pg_set_error("pg_layer_gemm_fwd_tc: tensor-core pipeline watchdog expired (MMA never signalled completion)");
        return PG_ECUDA;
    }
    return PG_OK;
}
