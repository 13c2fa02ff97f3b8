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
#include "tc_common.cuh"

namespace {
using namespace pgtc;

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

// W_ext [k_ext, F_out] -> per column block (n_full columns, the last one narrower) and k-tile the image
// [block][kt][half(hi,lo)][chunk][n] of float4 (4 consecutive k), dense in the block's own width
__global__ void __launch_bounds__(256) wprep_kernel(const float *__restrict__ w_ext, int k_ext, int F_out, int k_tiles, int TC_CHUNKS,
                                                    int n_full, int blocks, float4 *__restrict__ wp) {
    const int TC_BK = TC_CHUNKS * 4;
    const int64_t per_block = (int64_t)k_tiles * 2 * TC_CHUNKS * n_full;
    const int64_t slots = (int64_t)k_tiles * TC_CHUNKS * n_full;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < slots * blocks; t += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / slots);
        const int64_t u = t - (int64_t)b * slots;
        const int n0 = b * n_full;
        const int n_blk = min(n_full, ((F_out - n0 + 15) / 16) * 16);
        if (u >= (int64_t)k_tiles * TC_CHUNKS * n_blk) continue;
        const int n = (int)(u % n_blk);
        const int kc = (int)((u / n_blk) % TC_CHUNKS);
        const int kt = (int)(u / ((int64_t)n_blk * TC_CHUNKS));
        float v[4], h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = kt * TC_BK + kc * 4 + j;
            v[j] = (k < k_ext && n0 + n < F_out) ? w_ext[(int64_t)k * F_out + n0 + n] : 0.f;
            h[j] = tf32_hi(v[j]);
            l[j] = v[j] - h[j];
        }
        float4 *img = wp + (int64_t)b * per_block + (int64_t)kt * 2 * TC_CHUNKS * n_blk;
        img[(int64_t)kc * n_blk + n] = make_float4(h[0], h[1], h[2], h[3]);
        img[(int64_t)(TC_CHUNKS + kc) * n_blk + n] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

// Warp roles: warps 0..7 (256 threads) build the A tiles (and run the epilogue), warp 8 lane 0 streams
// the B tiles with 1-D bulk copies (TMA engine, no register staging) and issues the MMAs.
//   full[s]  : 8 producer-warp arrivals + 1 expect_tx arrival + the B bytes -> stage s is ready
//   empty[s] : tcgen05.commit of the MMAs that read stage s               -> stage s may be refilled
constexpr int TC_PRODUCERS = 256;
constexpr int TC_THREADS2 = TC_PRODUCERS + 32;

// Epilogue of the forward transform: H = leaky_relu(acc (+X) + constant).  The kernel hands the epilogues float4s of one row
// in an order that makes a warp's stores cover whole 128-byte lines (see the staging in tc_rows_gemm_kernel).
struct EpiFwdTc {
    static constexpr bool kRowDots = false;
    const float *constant, *x;
    int64_t ldconst, ldx;
    int add_identity;
    float slope;
    float *h;
    int64_t ldh;
    // 4 consecutive columns c..c+3 of one row (c % 4 == 0, inside the matrix)
    __device__ __forceinline__ void store4(int64_t row, int c, float4 v) const {
        if (constant) {
            const float4 cv = __ldg(reinterpret_cast<const float4 *>(constant + row * ldconst + c));
            v.x += cv.x; v.y += cv.y; v.z += cv.z; v.w += cv.w;
        }
        if (add_identity) {
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + row * ldx + c));
            v.x += xv.x; v.y += xv.y; v.z += xv.z; v.w += xv.w;
        }
        if (slope != 1.f) {
            v.x = v.x > 0.f ? v.x : v.x * slope;
            v.y = v.y > 0.f ? v.y : v.y * slope;
            v.z = v.z > 0.f ? v.z : v.z * slope;
            v.w = v.w > 0.f ? v.w : v.w * slope;
        }
        *reinterpret_cast<float4 *>(h + row * ldh + c) = v;
    }
};

// D[rows m0..m0+127, columns n0..n0+n_blk) = A[rows, K] @ B[K, columns]:  one CTA per (row tile, column block).
// AOp: the A operand, built by the producer warps (at4(row, k0) = 4 consecutive k; zero outside the matrix).
// wp : pre-split B image per column block [block][kt][hi|lo][chunk][n] (wprep kernels), streamed with bulk copies;
//      every block but the last is n_full columns wide.  Epi: what happens to the accumulators.
template <int TC_CHUNKS, class AOp, class Epi>
__global__ void __launch_bounds__(TC_THREADS2, (TC_CHUNKS == 4) ? 2 : 1) tc_rows_gemm_kernel(AOp A, Epi epi, const float4 *__restrict__ wp_all,
                                                                                           int n_full, int n_total, int k_tiles,
                                                                                           int *__restrict__ error_flag) {
    constexpr int TC_BK = TC_CHUNKS * 4;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[2], empty[2], done;
    __shared__ uint32_t tmem_base_slot;
    __shared__ volatile int bail;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
    const int n0 = (int)blockIdx.y * n_full;
    const int n_blk = min(n_full, ((n_total - n0 + 15) / 16) * 16);   // columns of this CTA (multiple of 16; the image is zero-padded)
    const float4 *wp = wp_all + (int64_t)blockIdx.y * k_tiles * 2 * TC_CHUNKS * n_full;
    const uint32_t lbo_b = (uint32_t)n_blk * 16u;              // B chunks are written by the copy engine: no padding needed
    const uint32_t a_bytes = TC_CHUNKS * TC_LBO_A, b_bytes = TC_CHUNKS * lbo_b;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * (uint32_t)TC_CHUNKS * (uint32_t)n_full * 16u;  // A_hi | A_lo | B_hi | B_lo
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < n_blk) tmem_cols <<= 1;

    if (tid == 0) {
        mbar_init(&full[0], TC_PRODUCERS / 32 + 1);
        mbar_init(&full[1], TC_PRODUCERS / 32 + 1);
        mbar_init(&empty[0], 1);
        mbar_init(&empty[1], 1);
        mbar_init(&done, 1);
        bail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, tmem_cols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_blk >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

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
                    fence_async_smem();  // generic-proxy stores -> visible to the tensor core
                    __syncwarp();        // one arrival per warp: 256 per-thread arrivals on one mbarrier serialise
                    if (lane == 0) mbar_arrive(&full[sg]);
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
            const float4 *src = wp + (int64_t)kt * 2 * TC_CHUNKS * n_blk;  // [half][chunk][n] contiguous == B_hi | B_lo image
            bulk_g2s(b_hi, src, b_bytes, &full[s]);
            bulk_g2s(b_lo, src + (int64_t)TC_CHUNKS * n_blk, b_bytes, &full[s]);
            if (!mbar_wait(&full[s], (uint32_t)((kt >> 1) & 1))) { bail = 1; break; }
            fence_after_sync();
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
            umma_commit(&empty[s]);
        }
        umma_commit(&done);
    }
    // ---------------------------------------------------------------------- epilogue (warps 0..7)
    if (warp < TC_PRODUCERS / 32) {
        if (!bail && !mbar_wait(&done, 0)) bail = 1;
        fence_after_sync();
        if (!bail) {
            // TMEM lanes are owned per warp quadrant (warp % 4); warps 4..7 take the upper half of the columns.  tcgen05.ld
            // gives every lane 32 consecutive columns of ITS row; written out like that, one store instruction would touch 32
            // different rows (32 lines, 16 bytes each).  Each warp therefore transposes through its own 32 x 32 patch of the
            // (now idle) pipeline stages: rows padded to 36 floats (conflict-free float4 rows), read back as 8 lanes per row,
            // so a store instruction covers 4 rows x 128 contiguous bytes.
            const int quad = warp & 3;
            float *stg = reinterpret_cast<float *>(smem) + warp * (32 * 36);
            const int half_cols = ((n_blk / 16 + 1) / 2) * 16;
            const int c_begin = (warp < 4) ? 0 : half_cols, c_end = (warp < 4) ? half_cols : n_blk;
            const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16);
            float dots[Epi::kRowDots ? 8 : 1][3];
#pragma unroll
            for (int i = 0; i < (Epi::kRowDots ? 8 : 1); ++i) dots[i][0] = dots[i][1] = dots[i][2] = 0.f;
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                uint32_t r[32];
                tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
                if (c0 + 16 < c_end) {
                    tmem_ld16(taddr + (uint32_t)(c0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
                } else {
#pragma unroll
                    for (int j = 16; j < 32; ++j) r[j] = 0u;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4 *>(stg + lane * 36 + 4 * q) = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                                                      __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rl = 4 * i + (lane >> 3);
                    const int cc = c0 + 4 * (lane & 7);
                    const int64_t row = m0 + quad * 32 + rl;
                    const float4 v = *reinterpret_cast<const float4 *>(stg + rl * 36 + 4 * (lane & 7));
                    if (row < A.M && cc < c_end) {
                        if constexpr (Epi::kRowDots) epi.dot4(row, n0 + cc, v, dots[i]);
                        else epi.store4(row, n0 + cc, v);
                    }
                }
                __syncwarp();
            }
            if constexpr (Epi::kRowDots) {   // 8 lanes share a row: sum their parts, lane 0 of the group writes
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int v = 0; v < 3; ++v) {
                        float d = dots[i][v];
                        d += __shfl_xor_sync(0xffffffffu, d, 1);
                        d += __shfl_xor_sync(0xffffffffu, d, 2);
                        d += __shfl_xor_sync(0xffffffffu, d, 4);
                        dots[i][v] = d;
                    }
                    const int64_t row = m0 + quad * 32 + 4 * i + (lane >> 3);
                    if ((lane & 7) == 0 && row < A.M) epi.write_dots(row, (int)blockIdx.y * 2 + (warp >= 4 ? 1 : 0), dots[i]);
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (bail && tid == 0) atomicExch(error_flag, 1);
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// backward, data gradient:  dA[:, :k_data] = dY @ W_ext[:k_data, :]^T   (gating + gate gradients: gate_grad_kernel, gemm.cu)
// Same kernel: A = rows of dY, B image = W_ext^T cut into column blocks of <= 256 (TMEM columns), epilogue = plain stores.
// ------------------------------------------------------------------------------------------------
struct APlainTc {
    const float *p;
    int64_t ld, M;
    int K;
    __device__ __forceinline__ float4 at4(int64_t i, int k0) const {  // K % 4 == 0, 16 B aligned rows
        if (i < M && k0 + 3 < K) return __ldg(reinterpret_cast<const float4 *>(p + i * ld + k0));
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
};

struct EpiBwdDataTc {
    static constexpr bool kRowDots = false;  // columns [0, 3 F_in) -> dZ (raw, gated later), [3 F_in, k_data) -> dXres
    float *dz, *dxres;
    int64_t lddz, lddxres;
    int f3, k_data;
    __device__ __forceinline__ void store4(int64_t row, int c, float4 v) const {
        if (c >= k_data) return;   // f3 and k_data are multiples of 4: a float4 never straddles a boundary
        if (c < f3) *reinterpret_cast<float4 *>(dz + row * lddz + c) = v;
        else *reinterpret_cast<float4 *>(dxres + row * lddxres + (c - f3)) = v;
    }
};

// ---- input gradient through the transposed structure:  dX = [T_in | T_out | T_und | dY?] @ Wcat (+ dY),  T_v = A_v (g_v * dY)
struct ACat2Tc {   // A(i, k) = k < K1 ? t[i, k] : dy[i, k - K1]     (K1 and K multiples of 4, 16 B aligned rows)
    const float *t, *dy;
    int64_t ldt, lddy, M;
    int K1, K;
    __device__ __forceinline__ float4 at4(int64_t i, int k0) const {
        if (i >= M || k0 + 3 >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 < K1) return __ldg(reinterpret_cast<const float4 *>(t + i * ldt + k0));
        return __ldg(reinterpret_cast<const float4 *>(dy + i * lddy + (k0 - K1)));
    }
};

struct EpiDxTc {
    static constexpr bool kRowDots = false;   // dX[row, c] = acc (+ dY[row, c] for the identity residual); c < F_in
    float *dx;
    const float *dy;   // null unless add_identity
    int64_t lddx, lddy;
    int F_in;
    __device__ __forceinline__ void store4(int64_t row, int c, float4 v) const {
        if (c >= F_in) return;
        if (dy) {
            const float4 r = __ldg(reinterpret_cast<const float4 *>(dy + row * lddy + c));
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        *reinterpret_cast<float4 *>(dx + row * lddx + c) = v;
    }
};

// B image of the dX GEMM: B(k, n = c) = W_ext[(k / F_out) * F_in + c][k % F_out] for the three propagation blocks and, with a
// residual projection, W_ext[3 F_in + c][k - 3 F_out] for k >= 3 F_out; [block][kt][half][chunk][n] like wprep_kernel
__global__ void __launch_bounds__(256) wprep_dx_kernel(const float *__restrict__ w_ext, int F_in, int F_out, int K, int k_tiles, int n_full,
                                                       int blocks, float4 *__restrict__ wp) {
    constexpr int CH = 4;
    const int64_t per_block = (int64_t)k_tiles * 2 * CH * n_full;
    const int64_t slots = (int64_t)k_tiles * CH * n_full;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < slots * blocks; t += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / slots);
        const int64_t u = t - (int64_t)b * slots;
        const int n0 = b * n_full;
        const int n_blk = min(n_full, ((F_in - n0 + 15) / 16) * 16);
        if (u >= (int64_t)k_tiles * CH * n_blk) continue;
        const int n = (int)(u % n_blk);
        const int kc = (int)((u / n_blk) % CH);
        const int kt = (int)(u / ((int64_t)n_blk * CH));
        const int c = n0 + n;
        float v[4], h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = kt * 16 + kc * 4 + j;
            v[j] = (k < K && c < F_in) ? w_ext[((int64_t)(k / F_out) * F_in + c) * F_out + (k % F_out)] : 0.f;
            h[j] = tf32_hi(v[j]);
            l[j] = v[j] - h[j];
        }
        float4 *img = wp + (int64_t)b * per_block + (int64_t)kt * 2 * CH * n_blk;
        img[(int64_t)kc * n_blk + n] = make_float4(h[0], h[1], h[2], h[3]);
        img[(int64_t)(CH + kc) * n_blk + n] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

// ---- gate gradients without materialising dZ:  dgate_v[i] = <dY[i] W_v^T, Z_v[i]> + <dY[i], beta_v>
// GEMM columns [0, 3 F_in) are the rows of dA = dY W_ext[:3 F_in]^T, columns 3 F_in + v the bias rows; the epilogue only forms
// per-row dot products with Z (kRowDots: the kernel keeps 3 sums per row group, reduces them over the 8 lanes that share a row
// and hands them to write_dots; a second kernel sums the column-block parts in fixed order).
struct EpiGateDotTc {
    static constexpr bool kRowDots = true;
    const float *z;
    int64_t ldz, M;
    int F_in, f3;
    float *partial;   // [parts][3][M]
    __device__ __forceinline__ void store4(int64_t, int, float4) const {}
    __device__ __forceinline__ void dot4(int64_t row, int c, float4 v, float (&acc)[3]) const {
        if (c < f3) {
            const float4 zv = __ldg(reinterpret_cast<const float4 *>(z + row * ldz + c));
            const float d = fmaf(v.x, zv.x, fmaf(v.y, zv.y, fmaf(v.z, zv.z, v.w * zv.w)));
            const int seg = c / F_in;    // F_in % 4 == 0: a float4 lies in one segment
            acc[0] += seg == 0 ? d : 0.f;
            acc[1] += seg == 1 ? d : 0.f;
            acc[2] += seg == 2 ? d : 0.f;
        } else if (c == f3) {            // the three bias-row columns share one float4
            acc[0] += v.x;
            acc[1] += v.y;
            acc[2] += v.z;
        }
    }
    __device__ __forceinline__ void write_dots(int64_t row, int part, const float (&acc)[3]) const {
#pragma unroll
        for (int v = 0; v < 3; ++v) partial[((int64_t)part * 3 + v) * M + row] = acc[v];
    }
};

// images of [W_ext[:3 F_in]; W_ext[k_data .. k_data+2]]^T per column block (columns 3 F_in + 3 .. are zero padding)
__global__ void __launch_bounds__(256) wprep_gate_kernel(const float *__restrict__ w_ext, int f3, int k_data, int F_out, int k_tiles,
                                                         int n_full, int blocks, float4 *__restrict__ wp) {
    constexpr int CH = 4;
    const int cols = f3 + 3;
    const int64_t per_block = (int64_t)k_tiles * 2 * CH * n_full;
    const int64_t slots = (int64_t)k_tiles * CH * n_full;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < slots * blocks; t += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / slots);
        const int64_t u = t - (int64_t)b * slots;
        const int n0 = b * n_full;
        const int n_blk = min(n_full, ((cols - n0 + 15) / 16) * 16);
        if (u >= (int64_t)k_tiles * CH * n_blk) continue;
        const int n = (int)(u % n_blk);
        const int kc = (int)((u / n_blk) % CH);
        const int kt = (int)(u / ((int64_t)n_blk * CH));
        const int c = n0 + n;
        const int64_t src_row = c < f3 ? c : (c < cols ? k_data + (c - f3) : -1);
        const int f0 = kt * 16 + kc * 4;
        float v[4], h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = (src_row >= 0 && f0 + j < F_out) ? w_ext[src_row * F_out + f0 + j] : 0.f;
            h[j] = tf32_hi(v[j]);
            l[j] = v[j] - h[j];
        }
        float4 *img = wp + (int64_t)b * per_block + (int64_t)kt * 2 * CH * n_blk;
        img[(int64_t)kc * n_blk + n] = make_float4(h[0], h[1], h[2], h[3]);
        img[(int64_t)(CH + kc) * n_blk + n] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

__global__ void __launch_bounds__(256) gate_parts_reduce_kernel(const float *__restrict__ partial, int parts, int64_t numel,
                                                                float *__restrict__ dgate) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int p = 0; p < parts; ++p) acc += partial[(int64_t)p * numel + i];   // fixed order
        dgate[i] = acc;
    }
}

struct EpiLinearTc {
    static constexpr bool kRowDots = false;  // out[row, c] = acc + bias[c]   (plain Linear; c < n_total, rows padded to ldo % 4 == 0)
    float *out;
    int64_t ldo;
    const float *bias;
    int n_total;
    __device__ __forceinline__ void store4(int64_t row, int c, float4 v) const {
        if (c >= n_total) return;
        float y[4] = {v.x, v.y, v.z, v.w};
        if (bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + j < n_total) y[j] += __ldg(bias + c + j);
        }
        if (c + 3 < n_total) {
            *reinterpret_cast<float4 *>(out + row * ldo + c) = make_float4(y[0], y[1], y[2], y[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c + j < n_total) out[row * ldo + c + j] = y[j];
        }
    }
};

// W_ext [k_ext, F_out] -> images of W_ext[:k_data]^T per column block: [block][kt][half][chunk][n] float4 over 4 consecutive f
__global__ void __launch_bounds__(256) wprep_t_kernel(const float *__restrict__ w_ext, int k_data, int F_out, int k_tiles, int n_full,
                                                      int blocks, float4 *__restrict__ wp) {
    constexpr int CH = 4;
    const int64_t per_block = (int64_t)k_tiles * 2 * CH * n_full;
    const int64_t total = (int64_t)blocks * k_tiles * CH * n_full;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / ((int64_t)k_tiles * CH * n_full));
        const int64_t u = t - (int64_t)b * k_tiles * CH * n_full;
        const int n0 = b * n_full;
        const int n_blk = min(n_full, ((k_data - n0 + 15) / 16) * 16);
        // enumerate (kt, kc, n) over the block's own width so that the image is dense in n_blk
        const int64_t dense = (int64_t)k_tiles * CH * n_blk;
        if (u >= dense) continue;
        const int n = (int)(u % n_blk);
        const int kc = (int)((u / n_blk) % CH);
        const int kt = (int)(u / ((int64_t)n_blk * CH));
        const int c = n0 + n;
        const int f0 = kt * 16 + kc * 4;
        float v[4], h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = (c < k_data && f0 + j < F_out) ? w_ext[(int64_t)c * F_out + f0 + j] : 0.f;
            h[j] = tf32_hi(v[j]);
            l[j] = v[j] - h[j];
        }
        float4 *img = wp + (int64_t)b * per_block + (int64_t)kt * 2 * CH * n_blk;
        img[(int64_t)kc * n_blk + n] = make_float4(h[0], h[1], h[2], h[3]);
        img[(int64_t)(CH + kc) * n_blk + n] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// backward, weight gradient:  dW_ext = A_ext^T @ dY  -- the reduction runs over the graph rows.
// Both operands are contiguous along their M / N dimension in memory (A_ext rows hold the ext columns, dY rows the
// output features), i.e. MN-major for the MMA.  kind::tf32 does not take MN-major operands on this part: with bit 15
// or 16 of the instruction descriptor set, every shared-memory layout (no swizzle / 32 / 64 / 128-byte swizzle, LBO and
// SBO either way round) returns exact zeros while the K-major forms of the same probe are correct (tools/probe_umma.cu,
// profiles/r01_probe_umma_tf32_major.txt).  So the producers transpose: a thread gathers 4 consecutive graph rows of ONE
// column (scalar loads, coalesced across the warp) into one 16-byte chunk of the K-major layout the forward uses.
// Grid = (ext-column tiles of 128) x (row splits, ~2 CTAs per SM); every CTA writes its 128 x F_out partial, a second
// kernel sums the splits in fixed order.
// ------------------------------------------------------------------------------------------------
constexpr int TCW_BK = 16;   // graph rows per stage = 4 chunks of 4

__global__ void __launch_bounds__(TC_THREADS2, 2) tc_bwd_weight_kernel(AExtTc A, const float *__restrict__ dy, int64_t lddy, int F_out,
                                                                       int64_t rows_per_split, float *__restrict__ partial,
                                                                       int *__restrict__ error_flag) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[2], empty[2], done;
    __shared__ uint32_t tmem_base_slot;
    __shared__ volatile int bail;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = (int)blockIdx.x * TC_BM;                         // first ext column of this tile
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_end = min(r_begin + rows_per_split, A.M);
    const int k_tiles = (int)((r_end - r_begin + TCW_BK - 1) / TCW_BK);
    const uint32_t b_lbo = (uint32_t)(F_out + 1) * 16u;             // +1 row of padding like TC_LBO_A
    const uint32_t a_bytes = 4 * TC_LBO_A, b_bytes = 4 * b_lbo;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;         // A_hi | A_lo | B_hi | B_lo
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < F_out) tmem_cols <<= 1;

    if (tid == 0) {
        mbar_init(&full[0], TC_PRODUCERS / 32);
        mbar_init(&full[1], TC_PRODUCERS / 32);
        mbar_init(&empty[0], 1);
        mbar_init(&empty[1], 1);
        mbar_init(&done, 1);
        bail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, tmem_cols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(F_out >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

    if (warp < TC_PRODUCERS / 32) {
        // ------------------------------------------------------------------ producers: 2 A chunks and <= 4 B chunks per thread and k-tile
        int a_row[2], a_col[2], b_row[4], b_col[4];   // slot = (first of 4 graph rows inside the tile, column)
        uint32_t a_off[2], b_off[4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int idx = j * TC_PRODUCERS + tid;
            a_row[j] = 4 * (idx >> 7);
            a_col[j] = m0 + (idx & 127);
            a_off[j] = (uint32_t)(idx >> 7) * TC_LBO_A + (uint32_t)(idx & 127) * 16u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = j * TC_PRODUCERS + tid;
            const bool on = idx < 4 * F_out;
            b_row[j] = on ? 4 * (idx / F_out) : -1;
            b_col[j] = on ? idx % F_out : 0;
            b_off[j] = (uint32_t)(on ? idx / F_out : 0) * b_lbo + (uint32_t)b_col[j] * 16u;
        }
        // The loads below are latency-bound (ncu: 6.6 long-scoreboard stall cycles per issued instruction with two k-tiles
        // in flight); more register prefetch costs the second CTA per SM, so tiles further ahead are pulled into L2
        // instead: 64 threads touch the 128-byte lines of an A tile (Z columns only), 128 threads those of a dY tile.
        constexpr int kL2Ahead = 8;
        auto prefetch_l2 = [&](int t) {
            const int64_t r0 = r_begin + (int64_t)(t + kL2Ahead) * TCW_BK;
            if (tid < 64) {
                const int64_t r = r0 + (tid >> 2);
                const int col = m0 + (tid & 3) * 32;
                if (r < r_end && col + 31 < 3 * A.F_in) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.z + r * A.ldz + col));
            } else if (tid < 192) {
                const int u = tid - 64;
                const int64_t r = r0 + (u >> 3);
                const int col = (u & 7) * 32;
                if (r < r_end && col < F_out) asm volatile("prefetch.global.L2 [%0];" ::"l"(dy + r * lddy + col));
            }
        };
        auto load_tile = [&](int t, float4 (&va)[2], float4 (&vb)[4]) {
            const int64_t r0 = r_begin + (int64_t)t * TCW_BK;
            prefetch_l2(t);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int64_t r = r0 + a_row[j];
                va[j] = make_float4(A.at(r, a_col[j]), A.at(r + 1, a_col[j]), A.at(r + 2, a_col[j]), A.at(r + 3, a_col[j]));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t r = r0 + b_row[j];
                float e[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) e[q] = (b_row[j] >= 0 && r + q < A.M) ? __ldg(dy + (r + q) * lddy + b_col[j]) : 0.f;
                vb[j] = make_float4(e[0], e[1], e[2], e[3]);
            }
        };
        auto split_store = [&](uint8_t *hi_base, uint8_t *lo_base, uint32_t off, const float4 &x4) {
            const float4 hi = make_float4(tf32_hi(x4.x), tf32_hi(x4.y), tf32_hi(x4.z), tf32_hi(x4.w));
            const float4 lo = make_float4(x4.x - hi.x, x4.y - hi.y, x4.z - hi.z, x4.w - hi.w);
            *reinterpret_cast<float4 *>(hi_base + off) = hi;
            *reinterpret_cast<float4 *>(lo_base + off) = lo;
        };
        float4 va[2][2], vb[2][4];   // two k-tiles of global loads in flight
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            if (d < k_tiles) load_tile(d, va[d], vb[d]);
        }
        for (int kt = 0; kt < k_tiles; kt += 2) {
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int t = kt + d;
                if (t < k_tiles) {
                    uint8_t *st = smem + (size_t)d * stage_bytes;     // stage == t & 1 == d
                    if (t >= 2 && !bail) {
                        if (!mbar_wait(&empty[d], (uint32_t)(((t >> 1) - 1) & 1))) bail = 1;
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) split_store(st, st + a_bytes, a_off[j], va[d][j]);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (b_row[j] >= 0) split_store(st + 2 * a_bytes, st + 2 * a_bytes + b_bytes, b_off[j], vb[d][j]);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[d]);
                    if (t + 2 < k_tiles) load_tile(t + 2, va[d], vb[d]);
                }
            }
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------------ MMA issue (one thread)
        for (int kt = 0; kt < k_tiles; ++kt) {
            const int s = kt & 1;
            if (bail) break;
            const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes), a_lo = a_hi + a_bytes, b_hi = a_hi + 2 * a_bytes, b_lo = b_hi + b_bytes;
            if (!mbar_wait(&full[s], (uint32_t)((kt >> 1) & 1))) { bail = 1; break; }
            fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < TCW_BK / 8; ++ks) {  // one MMA consumes 2 chunks (K = 8 tf32)
                const uint64_t dah = umma_desc(a_hi + ks * 2 * TC_LBO_A, TC_LBO_A, 128);
                const uint64_t dal = umma_desc(a_lo + ks * 2 * TC_LBO_A, TC_LBO_A, 128);
                const uint64_t dbh = umma_desc(b_hi + ks * 2 * b_lbo, b_lbo, 128);
                const uint64_t dbl = umma_desc(b_lo + ks * 2 * b_lbo, b_lbo, 128);
                umma_tf32(tmem_d, dah, dbh, idesc, (kt | ks) != 0);
                umma_tf32(tmem_d, dah, dbl, idesc, 1);
                umma_tf32(tmem_d, dal, dbh, idesc, 1);
            }
            umma_commit(&empty[s]);
        }
        umma_commit(&done);
    }
    // ---------------------------------------------------------------------- epilogue: TMEM -> this split's partial
    if (warp < TC_PRODUCERS / 32) {
        if (!bail && !mbar_wait(&done, 0)) bail = 1;
        fence_after_sync();
        if (!bail) {
            const int quad = warp & 3;
            const int m = m0 + quad * 32 + lane;
            const int half_cols = ((F_out / 16 + 1) / 2) * 16;
            const int c_begin = (warp < 4) ? 0 : half_cols, c_end = (warp < 4) ? half_cols : F_out;
            float *out = partial + ((int64_t)blockIdx.y * A.k_ext + m) * F_out;
            for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
                if (m < A.k_ext) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4 *>(out + c0 + q * 4) = k_tiles > 0
                            ? make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3]))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (bail && tid == 0) atomicExch(error_flag, 1);
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

__global__ void __launch_bounds__(256) tc_reduce_splits_kernel(const float *__restrict__ partial, int splits, int64_t numel,
                                                               float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * numel + i];  // fixed order
        out[i] = s;
    }
}

template <int CHUNKS, class AOp, class Epi>
int launch_rows_gemm(const AOp &A, const Epi &epi, const float4 *wp, int n_full, int n_total, int k_tiles, int *err, dim3 grid,
                     size_t smem, cudaStream_t st) {
    static bool attr_set = false;   // one flag per instantiation
    if (!attr_set) {
        PG_CUDA_CALL(cudaFuncSetAttribute(tc_rows_gemm_kernel<CHUNKS, AOp, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          CHUNKS == 4 ? 110 * 1024 : 200 * 1024));
        if (CHUNKS == 4)
            PG_CUDA_CALL(cudaFuncSetAttribute(tc_rows_gemm_kernel<CHUNKS, AOp, Epi>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    if (smem < (size_t)8 * 32 * 36 * sizeof(float)) smem = (size_t)8 * 32 * 36 * sizeof(float);   // the epilogue's staging patches
    tc_rows_gemm_kernel<CHUNKS, AOp, Epi><<<grid, TC_THREADS2, smem, st>>>(A, epi, wp, n_full, n_total, k_tiles, err);
    return PG_OK;
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
    // + 16 columns: the output may be cut into two column blocks whose widths are rounded up to 16 (few row tiles, see below)
    return (size_t)k_tiles_of(F_in, F_out, has_res) * 2 * chunks_for(F_in, F_out, has_res) * (F_out + 16) * sizeof(float4) + 256;
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
    // Few row tiles (n-gram graphs of n <= 3: 66 tiles at config C2) leave most SMs idle and the kernel latency-bound:
    // cut the output into two column blocks so that twice as many CTAs run (each rebuilds the A tile, from L2).
    const int64_t row_tiles = pg_ceil_div(num_rows, TC_BM);
    int n_full = F_out, blocks = 1;
    if (row_tiles < PG_NUM_SMS && F_out >= 64) {
        n_full = ((F_out + 1) / 2 + 15) / 16 * 16;
        blocks = (F_out + n_full - 1) / n_full;
    }
    {
        const int64_t total = (int64_t)blocks * kt * chunks * n_full;
        wprep_kernel<<<(unsigned)pg_ceil_div(total, 256), 256, 0, st>>>(d_w_ext, A.k_ext, F_out, kt, chunks, n_full, blocks, wp);
        PG_CUDA_LAUNCH_CHECK("wprep_kernel");
    }
    const size_t stage = 2 * (size_t)chunks * TC_LBO_A + 2 * (size_t)chunks * n_full * 16;
    const size_t smem = 2 * stage;
    EpiFwdTc epi{d_constant, d_x, ldconst, ldx, add_identity, slope, d_h, ldh};
    const dim3 grid((unsigned)row_tiles, (unsigned)blocks, 1);
    int rc = chunks == 4 ? launch_rows_gemm<4>(A, epi, wp, n_full, F_out, kt, err, grid, smem, st)
                         : launch_rows_gemm<8>(A, epi, wp, n_full, F_out, kt, err, grid, smem, st);
    if (rc != PG_OK) return rc;
    PG_CUDA_LAUNCH_CHECK("tc_rows_gemm_kernel (forward)");
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

// ---------------------------------------------------------------------------------------------- backward on tensor cores
namespace {
struct BwdDataPlan { int k_data, k_tiles, n_full, blocks; size_t image_bytes; };
// output columns `cols` (the B image's n), reduction length K: column blocks of <= 256, k-tiles of 16
inline BwdDataPlan bwd_data_plan_raw(int cols, int K);
inline BwdDataPlan bwd_data_plan(int F_in, int F_out, int has_res) { return bwd_data_plan_raw(3 * F_in + (has_res ? F_in : 0), F_out); }
inline BwdDataPlan bwd_data_plan_raw(int cols, int K) {
    BwdDataPlan p;
    p.k_data = cols;
    p.k_tiles = (K + 15) / 16;
    const int padded = (p.k_data + 15) / 16 * 16;
    p.n_full = padded < 256 ? padded : 256;
    p.blocks = (p.k_data + p.n_full - 1) / p.n_full;
    p.image_bytes = (size_t)p.blocks * p.k_tiles * 2 * 4 * p.n_full * sizeof(float4);
    return p;
}
struct BwdWeightPlan { int k_ext, m_tiles, splits; int64_t rows_per_split; };
inline BwdWeightPlan bwd_weight_plan(int64_t num_rows, int F_in, int has_res) {
    BwdWeightPlan p;
    p.k_ext = k_ext_of(F_in, has_res);
    p.m_tiles = (p.k_ext + TC_BM - 1) / TC_BM;
    int64_t s = (2 * PG_NUM_SMS) / p.m_tiles;                 // two CTAs per SM
    const int64_t max_s = pg_ceil_div(num_rows, 512);         // >= 32 k-tiles per CTA so the epilogue amortises
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    p.rows_per_split = pg_ceil_div(pg_ceil_div(num_rows, s), TCW_BK) * TCW_BK;
    p.splits = (int)pg_ceil_div(num_rows, p.rows_per_split);
    if (p.splits < 1) p.splits = 1;
    return p;
}
}  // namespace

extern "C" size_t pg_layer_gemm_bwd_data_tc_ws_bytes(int F_in, int F_out, int has_res) {
    return bwd_data_plan(F_in, F_out, has_res).image_bytes + 256;
}

extern "C" int pg_layer_gemm_bwd_data_tc(const float *d_dy, int64_t lddy, const float *d_w_ext, const float *d_z, int64_t ldz,
                                         const float *d_gate_a, const float *d_gate_b, const float *d_gate_c, int gate_stride,
                                         int64_t num_rows, int F_in, int F_out, int has_res, float *d_dz, int64_t lddz, float *d_dxres,
                                         int64_t lddxres, float *d_dgate, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F_in >= 1 && F_out >= 1 && (gate_stride == 0 || gate_stride == 1), "pg_layer_gemm_bwd_data_tc: bad shape");
    PG_CHECK_ARG(pg_layer_gemm_fwd_tc_supported(F_in, F_out), "pg_layer_gemm_bwd_data_tc: needs F_in %% 4 == 0, F_out %% 16 == 0, F_out <= 256");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_dy && d_w_ext && d_z && d_gate_a && d_gate_b && d_gate_c && d_dz && d_dgate && d_ws, "pg_layer_gemm_bwd_data_tc: null buffer");
    PG_CHECK_ARG(!has_res || d_dxres, "pg_layer_gemm_bwd_data_tc: has_res needs d_dxres");
    PG_CHECK_ARG(lddy >= F_out && ldz >= 3 * (int64_t)F_in && lddz >= 3 * (int64_t)F_in && (!has_res || lddxres >= F_in),
                 "pg_layer_gemm_bwd_data_tc: bad stride");
    PG_CHECK_ARG(al16(d_dy) && lddy % 4 == 0 && al16(d_dz) && lddz % 4 == 0 && (!has_res || (al16(d_dxres) && lddxres % 4 == 0)) && al16(d_ws),
                 "pg_layer_gemm_bwd_data_tc: operands must be 16-byte aligned with row strides %% 4 == 0");
    const BwdDataPlan p = bwd_data_plan(F_in, F_out, has_res);
    const size_t need = p.image_bytes + 256;
    if (ws_bytes < need) {
        pg_set_error("pg_layer_gemm_bwd_data_tc: workspace too small (%zu < %zu)", ws_bytes, need);
        return PG_EWORKSPACE;
    }
    cudaStream_t st = pg_cu(stream);
    float4 *wp = reinterpret_cast<float4 *>(d_ws);
    int *err = reinterpret_cast<int *>(reinterpret_cast<char *>(d_ws) + need - 256);
    PG_CUDA_CALL(cudaMemsetAsync(err, 0, sizeof(int), st));
    {
        const int64_t total = (int64_t)p.blocks * p.k_tiles * 4 * p.n_full;
        wprep_t_kernel<<<(unsigned)pg_ceil_div(total, 256), 256, 0, st>>>(d_w_ext, p.k_data, F_out, p.k_tiles, p.n_full, p.blocks, wp);
        PG_CUDA_LAUNCH_CHECK("wprep_t_kernel");
    }
    APlainTc A{d_dy, lddy, num_rows, F_out};
    EpiBwdDataTc epi{d_dz, d_dxres, lddz, lddxres, 3 * F_in, p.k_data};
    const size_t stage = 2 * (size_t)4 * TC_LBO_A + 2 * (size_t)4 * p.n_full * 16;
    const dim3 grid((unsigned)pg_ceil_div(num_rows, TC_BM), (unsigned)p.blocks, 1);
    int rc = launch_rows_gemm<4>(A, epi, wp, p.n_full, p.k_data, p.k_tiles, err, grid, 2 * stage, st);
    if (rc != PG_OK) return rc;
    PG_CUDA_LAUNCH_CHECK("tc_rows_gemm_kernel (data gradient)");
    return pg_launch_gate_grad(d_dz, lddz, d_z, ldz, d_dy, lddy, d_w_ext, d_gate_a, d_gate_b, d_gate_c, gate_stride, num_rows, F_in, F_out,
                               p.k_data, d_dgate, st);
}

extern "C" size_t pg_layer_gemm_bwd_weight_tc_ws_bytes(int64_t num_rows, int F_in, int F_out, int has_res) {
    const BwdWeightPlan p = bwd_weight_plan(num_rows, F_in, has_res);
    return pg_align_up((size_t)p.splits * p.k_ext * F_out * sizeof(float), 256) + 256;
}

extern "C" int pg_layer_gemm_bwd_weight_tc(const float *d_z, int64_t ldz, const float *d_x, int64_t ldx, const float *d_gate_a,
                                           const float *d_gate_b, const float *d_gate_c, int gate_stride, const float *d_dy,
                                           int64_t lddy, int64_t num_rows, int F_in, int F_out, int has_res, float *d_dw_ext,
                                           void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F_in >= 1 && F_out >= 1 && (gate_stride == 0 || gate_stride == 1), "pg_layer_gemm_bwd_weight_tc: bad shape");
    PG_CHECK_ARG(pg_layer_gemm_fwd_tc_supported(F_in, F_out), "pg_layer_gemm_bwd_weight_tc: needs F_in %% 4 == 0, F_out %% 16 == 0, F_out <= 256");
    PG_CHECK_ARG(d_dw_ext, "pg_layer_gemm_bwd_weight_tc: null output");
    cudaStream_t st = pg_cu(stream);
    const BwdWeightPlan p = bwd_weight_plan(num_rows, F_in, has_res);
    const int64_t numel = (int64_t)p.k_ext * F_out;
    if (num_rows == 0) {
        PG_CUDA_CALL(cudaMemsetAsync(d_dw_ext, 0, (size_t)numel * sizeof(float), st));
        return PG_OK;
    }
    PG_CHECK_ARG(d_z && d_gate_a && d_gate_b && d_gate_c && d_dy && d_ws && (!has_res || d_x), "pg_layer_gemm_bwd_weight_tc: null buffer");
    PG_CHECK_ARG(al16(d_z) && ldz % 4 == 0 && (!has_res || (al16(d_x) && ldx % 4 == 0)) && al16(d_dy) && lddy % 4 == 0 && al16(d_ws),
                 "pg_layer_gemm_bwd_weight_tc: operands must be 16-byte aligned with row strides %% 4 == 0");
    PG_CHECK_ARG(ldz >= 3 * (int64_t)F_in && lddy >= F_out, "pg_layer_gemm_bwd_weight_tc: bad stride");
    const size_t need = pg_layer_gemm_bwd_weight_tc_ws_bytes(num_rows, F_in, F_out, has_res);
    if (ws_bytes < need) {
        pg_set_error("pg_layer_gemm_bwd_weight_tc: workspace too small (%zu < %zu)", ws_bytes, need);
        return PG_EWORKSPACE;
    }
    AExtTc A;
    A.z = d_z; A.x = d_x; A.ga = d_gate_a; A.gb = d_gate_b; A.gc = d_gate_c; A.gate_stride = gate_stride;
    A.ldz = ldz; A.ldx = ldx; A.M = num_rows; A.F_in = F_in; A.has_res = has_res;
    A.k_data = 3 * F_in + (has_res ? F_in : 0);
    A.k_ext = p.k_ext;
    float *partial = reinterpret_cast<float *>(d_ws);
    int *err = reinterpret_cast<int *>(reinterpret_cast<char *>(d_ws) + need - 256);
    PG_CUDA_CALL(cudaMemsetAsync(err, 0, sizeof(int), st));
    const size_t stage = 2 * (size_t)4 * TC_LBO_A + 2 * (size_t)4 * (F_out + 1) * 16;
    static bool attr_set = false;
    if (!attr_set) {
        PG_CUDA_CALL(cudaFuncSetAttribute(tc_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        PG_CUDA_CALL(cudaFuncSetAttribute(tc_bwd_weight_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    const dim3 grid((unsigned)p.m_tiles, (unsigned)p.splits, 1);
    tc_bwd_weight_kernel<<<grid, TC_THREADS2, 2 * stage, st>>>(A, d_dy, lddy, F_out, p.rows_per_split, partial, err);
    PG_CUDA_LAUNCH_CHECK("tc_bwd_weight_kernel");
    tc_reduce_splits_kernel<<<(unsigned)pg_ceil_div(numel, 256), 256, 0, st>>>(partial, p.splits, numel, d_dw_ext);
    PG_CUDA_LAUNCH_CHECK("tc_reduce_splits_kernel");
    return PG_OK;
}

// watchdog flag of the last tensor-core call that used this workspace; `need` = the ws_bytes query of that call (host sync; tests only)
extern "C" int pg_tc_check(const void *d_ws, size_t need, pg_stream_t stream) {
    int flag = 0;
    PG_CUDA_CALL(cudaMemcpyAsync(&flag, reinterpret_cast<const char *>(d_ws) + need - 256, sizeof(int), cudaMemcpyDeviceToHost, pg_cu(stream)));
    PG_CUDA_CALL(cudaStreamSynchronize(pg_cu(stream)));
    if (flag) {
        pg_set_error("tensor-core pipeline watchdog expired (an MMA never signalled completion)");
        return PG_ECUDA;
    }
    return PG_OK;
}

// ---------------------------------------------------------------------------------------------- plain Linear on tensor cores
// out[N, C] = x[N, K] @ W[C, K]^T + bias  (torch.nn.Linear layout) -- the decoder's output layer with C = N classes
// (row f1: 8401 x 8401 logits from K = 32 at config C2).  Same kernel as the data gradient: rows of x against the
// pre-split image of W^T in column blocks of 256.
extern "C" size_t pg_linear_tc_ws_bytes(int K, int C) {
    if (K < 1 || C < 1) return 0;
    const BwdDataPlan p = bwd_data_plan_raw(C, K);
    return p.image_bytes + 256;
}

extern "C" int pg_linear_tc(const float *d_x, int64_t ldx, int64_t num_rows, int K, const float *d_w, const float *d_bias, int C,
                            float *d_out, int64_t ldo, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && K >= 4 && K % 4 == 0 && C >= 1, "pg_linear_tc: needs K %% 4 == 0");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_x && d_w && d_out && d_ws, "pg_linear_tc: null buffer");
    PG_CHECK_ARG(al16(d_x) && ldx % 4 == 0 && ldx >= K && al16(d_out) && ldo % 4 == 0 && ldo >= C && al16(d_ws),
                 "pg_linear_tc: x / out must be 16-byte aligned with row strides %% 4 == 0 (pad the output rows)");
    const BwdDataPlan p = bwd_data_plan_raw(C, K);
    const size_t need = p.image_bytes + 256;
    if (ws_bytes < need) {
        pg_set_error("pg_linear_tc: workspace too small (%zu < %zu)", ws_bytes, need);
        return PG_EWORKSPACE;
    }
    cudaStream_t st = pg_cu(stream);
    float4 *wp = reinterpret_cast<float4 *>(d_ws);
    int *err = reinterpret_cast<int *>(reinterpret_cast<char *>(d_ws) + need - 256);
    PG_CUDA_CALL(cudaMemsetAsync(err, 0, sizeof(int), st));
    const int64_t total = (int64_t)p.blocks * p.k_tiles * 4 * p.n_full;
    wprep_t_kernel<<<(unsigned)pg_ceil_div(total, 256), 256, 0, st>>>(d_w, C, K, p.k_tiles, p.n_full, p.blocks, wp);
    PG_CUDA_LAUNCH_CHECK("wprep_t_kernel");
    APlainTc A{d_x, ldx, num_rows, K};
    EpiLinearTc epi{d_out, ldo, d_bias, C};
    const size_t stage = 2 * (size_t)4 * TC_LBO_A + 2 * (size_t)4 * p.n_full * 16;
    const dim3 grid((unsigned)pg_ceil_div(num_rows, TC_BM), (unsigned)p.blocks, 1);
    int rc = launch_rows_gemm<4>(A, epi, wp, p.n_full, C, p.k_tiles, err, grid, 2 * stage, st);
    if (rc != PG_OK) return rc;
    PG_CUDA_LAUNCH_CHECK("tc_rows_gemm_kernel (linear)");
    return PG_OK;
}

// ---------------------------------------------------------------------------------------------- dX through the fan-out kernel
// dX = sum_v (A_v (g_v * dY)) W'_v^T (+ dY W_res^T | + dY): d_t = [T_in | T_out | T_und] from pg_spmm_fanout_scaled on dY
// (symmetric structure), then this GEMM with K = 3 F_out (+ F_out for a residual projection).
extern "C" size_t pg_layer_gemm_bwd_dx_tc_ws_bytes(int F_in, int F_out, int has_res) {
    const BwdDataPlan p = bwd_data_plan_raw(F_in, (3 + (has_res ? 1 : 0)) * F_out);
    return p.image_bytes + 256;
}

extern "C" int pg_layer_gemm_bwd_dx_tc(const float *d_t, int64_t ldt, const float *d_dy, int64_t lddy, const float *d_w_ext,
                                       int64_t num_rows, int F_in, int F_out, int has_res, int add_identity, float *d_dx, int64_t lddx,
                                       void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F_in >= 4 && F_in % 4 == 0 && F_out >= 4 && F_out % 4 == 0, "pg_layer_gemm_bwd_dx_tc: needs F_in, F_out %% 4 == 0");
    PG_CHECK_ARG(!(has_res && add_identity) && (!add_identity || F_in == F_out), "pg_layer_gemm_bwd_dx_tc: bad residual mode");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_t && d_dy && d_w_ext && d_dx && d_ws, "pg_layer_gemm_bwd_dx_tc: null buffer");
    PG_CHECK_ARG(al16(d_t) && ldt % 4 == 0 && ldt >= 3 * (int64_t)F_out && al16(d_dy) && lddy % 4 == 0 && lddy >= F_out && al16(d_dx) &&
                     lddx % 4 == 0 && lddx >= F_in && al16(d_ws),
                 "pg_layer_gemm_bwd_dx_tc: operands must be 16-byte aligned with row strides %% 4 == 0");
    const int K = (3 + (has_res ? 1 : 0)) * F_out;
    const BwdDataPlan p = bwd_data_plan_raw(F_in, K);
    const size_t need = p.image_bytes + 256;
    if (ws_bytes < need) {
        pg_set_error("pg_layer_gemm_bwd_dx_tc: workspace too small (%zu < %zu)", ws_bytes, need);
        return PG_EWORKSPACE;
    }
    cudaStream_t st = pg_cu(stream);
    float4 *wp = reinterpret_cast<float4 *>(d_ws);
    int *err = reinterpret_cast<int *>(reinterpret_cast<char *>(d_ws) + need - 256);
    PG_CUDA_CALL(cudaMemsetAsync(err, 0, sizeof(int), st));
    const int64_t total = (int64_t)p.blocks * p.k_tiles * 4 * p.n_full;
    wprep_dx_kernel<<<(unsigned)pg_ceil_div(total, 256), 256, 0, st>>>(d_w_ext, F_in, F_out, K, p.k_tiles, p.n_full, p.blocks, wp);
    PG_CUDA_LAUNCH_CHECK("wprep_dx_kernel");
    ACat2Tc A{d_t, d_dy, ldt, lddy, num_rows, 3 * F_out, K};
    EpiDxTc epi{d_dx, add_identity ? d_dy : nullptr, lddx, lddy, F_in};
    const size_t stage = 2 * (size_t)4 * TC_LBO_A + 2 * (size_t)4 * p.n_full * 16;
    const dim3 grid((unsigned)pg_ceil_div(num_rows, TC_BM), (unsigned)p.blocks, 1);
    int rc = launch_rows_gemm<4>(A, epi, wp, p.n_full, F_in, p.k_tiles, err, grid, 2 * stage, st);
    if (rc != PG_OK) return rc;
    PG_CUDA_LAUNCH_CHECK("tc_rows_gemm_kernel (input gradient)");
    return PG_OK;
}

// ---------------------------------------------------------------------------------------------- gate gradients, dZ never stored
namespace {
struct GatePlan { BwdDataPlan g; int parts; size_t partial_off, need; };
inline GatePlan gate_plan(int64_t num_rows, int F_in, int F_out) {
    GatePlan p;
    p.g = bwd_data_plan_raw(3 * F_in + 3, F_out);
    p.parts = 2 * p.g.blocks;
    p.partial_off = pg_align_up(p.g.image_bytes, 256);
    p.need = p.partial_off + pg_align_up((size_t)p.parts * 3 * (size_t)num_rows * sizeof(float), 256) + 256;
    return p;
}
}  // namespace

extern "C" size_t pg_layer_gate_grad_tc_ws_bytes(int64_t num_rows, int F_in, int F_out) { return gate_plan(num_rows, F_in, F_out).need; }

extern "C" int pg_layer_gate_grad_tc(const float *d_dy, int64_t lddy, const float *d_w_ext, const float *d_z, int64_t ldz, int64_t num_rows,
                                     int F_in, int F_out, int has_res, float *d_dgate, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F_in >= 4 && F_in % 4 == 0 && F_out >= 4 && F_out % 4 == 0, "pg_layer_gate_grad_tc: needs F_in, F_out %% 4 == 0");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_dy && d_w_ext && d_z && d_dgate && d_ws, "pg_layer_gate_grad_tc: null buffer");
    PG_CHECK_ARG(al16(d_dy) && lddy % 4 == 0 && lddy >= F_out && al16(d_z) && ldz % 4 == 0 && ldz >= 3 * (int64_t)F_in && al16(d_ws),
                 "pg_layer_gate_grad_tc: operands must be 16-byte aligned with row strides %% 4 == 0");
    const GatePlan p = gate_plan(num_rows, F_in, F_out);
    if (ws_bytes < p.need) {
        pg_set_error("pg_layer_gate_grad_tc: workspace too small (%zu < %zu)", ws_bytes, p.need);
        return PG_EWORKSPACE;
    }
    cudaStream_t st = pg_cu(stream);
    float4 *wp = reinterpret_cast<float4 *>(d_ws);
    float *partial = reinterpret_cast<float *>(reinterpret_cast<char *>(d_ws) + p.partial_off);
    int *err = reinterpret_cast<int *>(reinterpret_cast<char *>(d_ws) + p.need - 256);
    PG_CUDA_CALL(cudaMemsetAsync(err, 0, sizeof(int), st));
    const int k_data = 3 * F_in + (has_res ? F_in : 0);
    const int64_t total = (int64_t)p.g.blocks * p.g.k_tiles * 4 * p.g.n_full;
    wprep_gate_kernel<<<(unsigned)pg_ceil_div(total, 256), 256, 0, st>>>(d_w_ext, 3 * F_in, k_data, F_out, p.g.k_tiles, p.g.n_full, p.g.blocks, wp);
    PG_CUDA_LAUNCH_CHECK("wprep_gate_kernel");
    APlainTc A{d_dy, lddy, num_rows, F_out};
    EpiGateDotTc epi{d_z, ldz, num_rows, F_in, 3 * F_in, partial};
    const size_t stage = 2 * (size_t)4 * TC_LBO_A + 2 * (size_t)4 * p.g.n_full * 16;
    const dim3 grid((unsigned)pg_ceil_div(num_rows, TC_BM), (unsigned)p.g.blocks, 1);
    int rc = launch_rows_gemm<4>(A, epi, wp, p.g.n_full, 3 * F_in + 3, p.g.k_tiles, err, grid, 2 * stage, st);
    if (rc != PG_OK) return rc;
    PG_CUDA_LAUNCH_CHECK("tc_rows_gemm_kernel (gate gradients)");
    const int64_t numel = 3 * num_rows;
    gate_parts_reduce_kernel<<<(unsigned)pg_ceil_div(numel, 256), 256, 0, st>>>(partial, p.parts, numel, d_dgate);
    PG_CUDA_LAUNCH_CHECK("gate_parts_reduce_kernel");
    return PG_OK;
}
