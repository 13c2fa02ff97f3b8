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
// k per stage = CHUNKS chunks of 16 B (4 fp32): CHUNKS = 8 -> BK = 32 (4 MMA k-steps), CHUNKS = 4 -> BK = 16.
// The smaller stage lets 2 CTAs share an SM, so one CTA's prologue / epilogue hides under the other's mainloop.
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
__global__ void __launch_bounds__(256) wprep_kernel(const float *__restrict__ w_ext, int k_ext, int F_out, int k_tiles, int TC_CHUNKS,
                                                    float4 *__restrict__ wp) {
    const int TC_BK = TC_CHUNKS * 4;
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

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Warp roles: warps 0..7 (256 threads) build the A tiles (and run the epilogue), warp 8 lane 0 streams
// the B tiles with 1-D bulk copies (TMA engine, no register staging) and issues the MMAs.
//   full[s]  : 256 producer arrivals + 1 expect_tx arrival + the B bytes  -> stage s is ready
//   empty[s] : tcgen05.commit of the MMAs that read stage s               -> stage s may be refilled
constexpr int TC_PRODUCERS = 256;
constexpr int TC_THREADS2 = TC_PRODUCERS + 32;

template <int TC_CHUNKS>
__global__ void __launch_bounds__(TC_THREADS2, (TC_CHUNKS == 4) ? 2 : 1) layer_gemm_fwd_tc_kernel(AExtTc A, const float4 *__restrict__ wp, int F_out, int k_tiles,
                                                                           const float *__restrict__ constant, int64_t ldconst,
                                                                           int add_identity, float slope, float *__restrict__ h,
                                                                           int64_t ldh, int *__restrict__ error_flag) {
    constexpr int TC_BK = TC_CHUNKS * 4;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[2], empty[2], done;
    __shared__ uint32_t tmem_base_slot;
    __shared__ volatile int bail;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
    const uint32_t lbo_b = (uint32_t)F_out * 16u;              // B chunks are written by the copy engine: no padding needed
    const uint32_t a_bytes = TC_CHUNKS * TC_LBO_A, b_bytes = TC_CHUNKS * lbo_b;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;  // A_hi | A_lo | B_hi | B_lo
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < F_out) tmem_cols <<= 1;

    if (tid == 0) {
        mbar_init(&full[0], TC_PRODUCERS + 1);
        mbar_init(&full[1], TC_PRODUCERS + 1);
        mbar_init(&empty[0], 1);
        mbar_init(&empty[1], 1);
        mbar_init(&done, 1);
        bail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(F_out >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

    if (warp < TC_PRODUCERS / 32) {
        // ------------------------------------------------------------------ A producers
        constexpr int PER = (TC_BM * TC_CHUNKS) / TC_PRODUCERS;  // float4 per thread per tile
        constexpr int PF = 2;                                    // k-tiles of global loads kept in flight (registers); 4 measured slower (C3 0.59 -> 0.69 ms)
        float4 v[PF][PER];
#pragma unroll
        for (int d = 0; d < PF; ++d)
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int idx = i * TC_PRODUCERS + tid;
                v[d][i] = (d < k_tiles) ? A.at4(m0 + idx / TC_CHUNKS, d * TC_BK + (idx % TC_CHUNKS) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        for (int kt = 0; kt < k_tiles; kt += PF) {
#pragma unroll
            for (int d = 0; d < PF; ++d) {
                const int t = kt + d;
                if (t < k_tiles) {
                    const int sg = d & 1;                              // stage == t & 1 (kt is a multiple of PF, PF even)
                    uint8_t *st = smem + (size_t)sg * stage_bytes;
                    if (t >= 2 && !bail) {
                        if (!mbar_wait(&empty[sg], (uint32_t)(((t >> 1) - 1) & 1))) bail = 1;
                    }
#pragma unroll
                    for (int i = 0; i < PER; ++i) {
                        const int idx = i * TC_PRODUCERS + tid;
                        const int r = idx / TC_CHUNKS, kc = idx % TC_CHUNKS;
                        const float4 x4 = v[d][i];
                        const float4 hi = make_float4(tf32_hi(x4.x), tf32_hi(x4.y), tf32_hi(x4.z), tf32_hi(x4.w));
                        const float4 lo = make_float4(x4.x - hi.x, x4.y - hi.y, x4.z - hi.z, x4.w - hi.w);
                        *reinterpret_cast<float4 *>(st + kc * TC_LBO_A + r * 16) = hi;
                        *reinterpret_cast<float4 *>(st + a_bytes + kc * TC_LBO_A + r * 16) = lo;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
                    mbar_arrive(&full[sg]);
                    if (t + PF < k_tiles) {  // refill the register slot with the tile PF steps ahead
                        const int k0 = (t + PF) * TC_BK;
#pragma unroll
                        for (int i = 0; i < PER; ++i) {
                            const int idx = i * TC_PRODUCERS + tid;
                            v[d][i] = A.at4(m0 + idx / TC_CHUNKS, k0 + (idx % TC_CHUNKS) * 4);
                        }
                    }
                }
            }
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------------ B copies + MMA issue (one thread)
        for (int kt = 0; kt < k_tiles; ++kt) {
            const int s = kt & 1;
            uint8_t *st = smem + (size_t)s * stage_bytes;
            if (kt >= 2 && !bail) {
                if (!mbar_wait(&empty[s], (uint32_t)(((kt >> 1) - 1) & 1))) bail = 1;
            }
            if (bail) break;
            const uint32_t a_hi = smem_u32(st), a_lo = a_hi + a_bytes, b_hi = a_hi + 2 * a_bytes, b_lo = b_hi + b_bytes;
            mbar_arrive_expect_tx(&full[s], 2 * b_bytes);
            const float4 *src = wp + (int64_t)kt * 2 * TC_CHUNKS * F_out;  // [half][chunk][n] contiguous == B_hi | B_lo image
            bulk_g2s(b_hi, src, b_bytes, &full[s]);
            bulk_g2s(b_lo, src + (int64_t)TC_CHUNKS * F_out, b_bytes, &full[s]);
            if (!mbar_wait(&full[s], (uint32_t)((kt >> 1) & 1))) { bail = 1; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    }
    // ---------------------------------------------------------------------- epilogue (warps 0..7)
    if (warp < TC_PRODUCERS / 32) {
        if (!bail && !mbar_wait(&done, 0)) bail = 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!bail) {
            // TMEM lanes are owned per warp quadrant (warp % 4); warps 4..7 take the upper half of the columns
            const int quad = warp & 3;
            const int64_t row = m0 + quad * 32 + lane;
            const int half_cols = ((F_out / 16 + 1) / 2) * 16;
            const int c_begin = (warp < 4) ? 0 : half_cols, c_end = (warp < 4) ? half_cols : F_out;
            const bool in_range = row < A.M;
            auto load_side = [&](int c0, float4 (&dst)[4]) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (in_range && c0 < c_end) {
                        if (constant) acc4 = __ldg(reinterpret_cast<const float4 *>(constant + row * ldconst + c0 + q * 4));
                        if (add_identity) {
                            const float4 xv = __ldg(reinterpret_cast<const float4 *>(A.x + row * A.ldx + c0 + q * 4));
                            acc4.x += xv.x; acc4.y += xv.y; acc4.z += xv.z; acc4.w += xv.w;
                        }
                    }
                    dst[q] = acc4;
                }
            };
            auto process = [&](int c0, const float4 (&cur)[4], float4 (&nxt)[4]) {
                load_side(c0 + 16, nxt);  // next chunk's side loads fly under this chunk's TMEM read + stores
                uint32_t r[16];
                const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (in_range) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 sd = cur[q];
                        float y[4] = {__uint_as_float(r[q * 4 + 0]) + sd.x, __uint_as_float(r[q * 4 + 1]) + sd.y,
                                      __uint_as_float(r[q * 4 + 2]) + sd.z, __uint_as_float(r[q * 4 + 3]) + sd.w};
                        if (slope != 1.f) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) y[j] = y[j] > 0.f ? y[j] : y[j] * slope;
                        }
                        *reinterpret_cast<float4 *>(h + row * ldh + c0 + q * 4) = make_float4(y[0], y[1], y[2], y[3]);
                    }
                }
            };
            float4 side_a[4], side_b[4];
            load_side(c_begin, side_a);
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                process(c0, side_a, side_b);
                if (c0 + 16 < c_end) process(c0 + 16, side_b, side_a);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (bail && tid == 0) atomicExch(error_flag, 1);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}

inline bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }
inline int k_ext_of(int F_in, int has_res) { return 3 * F_in + (has_res ? F_in : 0) + 3 + (has_res ? 1 : 0); }
// chunks per stage: 4 (two CTAs per SM) unless the reduction is long and the tile wide
inline int chunks_for(int F_in, int F_out, int has_res) {
    (void)F_in; (void)has_res;
    return F_out <= 256 ? 4 : 8;
}
inline int k_tiles_of(int F_in, int F_out, int has_res) {
    const int bk = 4 * chunks_for(F_in, F_out, has_res);
    return (k_ext_of(F_in, has_res) + bk - 1) / bk;
}
}  // namespace

extern "C" int pg_layer_gemm_fwd_tc_supported(int F_in, int F_out) {
    return F_in % 4 == 0 && F_out % 16 == 0 && F_out >= 16 && F_out <= 256;
}

extern "C" size_t pg_layer_gemm_fwd_tc_ws_bytes(int F_in, int F_out, int has_res) {
    return (size_t)k_tiles_of(F_in, F_out, has_res) * 2 * chunks_for(F_in, F_out, has_res) * F_out * sizeof(float4) + 256;
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
    PG_CHECK_ARG(!d_constant || (al16(d_constant) && ldconst % 4 == 0), "pg_layer_gemm_fwd_tc: constant must be 16-byte aligned, ldconst %% 4 == 0");
    const int kt = k_tiles_of(F_in, F_out, has_res);
    const int chunks = chunks_for(F_in, F_out, has_res);
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
        const int64_t total = (int64_t)kt * chunks * F_out;
        wprep_kernel<<<(unsigned)pg_ceil_div(total, 256), 256, 0, st>>>(d_w_ext, A.k_ext, F_out, kt, chunks, wp);
        PG_CUDA_LAUNCH_CHECK("wprep_kernel");
    }
    const size_t stage = 2 * (size_t)chunks * TC_LBO_A + 2 * (size_t)chunks * F_out * 16;
    const size_t smem = 2 * stage;
    static bool attr_set = false;
    if (!attr_set) {
        PG_CUDA_CALL(cudaFuncSetAttribute(layer_gemm_fwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        PG_CUDA_CALL(cudaFuncSetAttribute(layer_gemm_fwd_tc_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        PG_CUDA_CALL(cudaFuncSetAttribute(layer_gemm_fwd_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    const unsigned grid = (unsigned)pg_ceil_div(num_rows, TC_BM);
    if (chunks == 4)
        layer_gemm_fwd_tc_kernel<4><<<grid, TC_THREADS2, smem, st>>>(A, wp, F_out, kt, d_constant, ldconst, add_identity, slope, d_h, ldh, err);
    else
        layer_gemm_fwd_tc_kernel<8><<<grid, TC_THREADS2, smem, st>>>(A, wp, F_out, kt, d_constant, ldconst, add_identity, slope, d_h, ldh, err);
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
