// Row f1 of SURVEY.md section 8: the loss of the next-node task, log_softmax over C = N classes
// followed by nll_loss (reference protgram_directgcn.py:221 + protgram_directgcn_trainer.py:90-94)
// and their autograd backward.  The reference makes 10 passes over the N x C matrix (282 MB at
// C2): log_softmax fwd (R+W), nll gather, zero-fill of grad (W), nll scatter, log_softmax bwd
// (2R+W), bias-grad reduce (R), plus the GEMMs.  Here ONE pass reads each logit row into shared
// memory, reduces max / sum(exp), and overwrites the row with d(loss)/d(logits); the column sums of
// that gradient (= bias gradient) are accumulated per CTA in shared memory and combined in fixed
// order, so the result is bitwise reproducible.  HBM traffic: 8 B per logit (read + write).
#include "common.cuh"

namespace {

constexpr int kThreads = 512;

__device__ __forceinline__ float block_reduce(float v, float *red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, v, s);
        v = is_max ? fmaxf(v, o) : v + o;
    }
    __syncthreads();  // red[] may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];  // fixed order
    return r;
}

// CTA b owns rows b, b + grid, ...; thread t owns columns t, t + 512, ... of every row, so the
// per-CTA column sums need no atomics.  smem: row[c] (exp values) + colsum[c].
__global__ void __launch_bounds__(kThreads) softmax_nll_kernel(float *__restrict__ logits, int64_t ld, int64_t n, int c,
                                                               const int64_t *__restrict__ labels, float grad_scale,
                                                               float *__restrict__ row_loss, float *__restrict__ colsum_partial) {
    extern __shared__ float sm[];
    float *row = sm, *colsum = sm + c;
    __shared__ float red[kThreads / 32];
    for (int j = threadIdx.x; j < c; j += kThreads) colsum[j] = 0.f;
    for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
        float *x = logits + r * ld;
        const int64_t y = labels[r];
        if (y < 0 || y >= c) {  // ignored row (nll_loss ignore_index semantics): no loss, zero gradient
            for (int j = threadIdx.x; j < c; j += kThreads) x[j] = 0.f;
            if (threadIdx.x == 0) row_loss[r] = 0.f;
            continue;
        }
        float m = -INFINITY, xy = 0.f;
        for (int j = threadIdx.x; j < c; j += kThreads) {
            const float v = x[j];
            row[j] = v;
            m = fmaxf(m, v);
            if (j == (int)y) xy = v;  // only the thread that owns column y keeps it
        }
        m = block_reduce(m, red, true);
        float s = 0.f;
        for (int j = threadIdx.x; j < c; j += kThreads) {
            const float e = expf(row[j] - m);
            row[j] = e;
            s += e;
        }
        s = block_reduce(s, red, false);
        const float inv = 1.f / s;
        for (int j = threadIdx.x; j < c; j += kThreads) {
            const float g = (row[j] * inv - (j == (int)y ? 1.f : 0.f)) * grad_scale;
            x[j] = g;
            colsum[j] += g;
        }
        if (threadIdx.x == (int)(y % kThreads)) row_loss[r] = (m - xy) + logf(s);  // -log_softmax(x)[y]
    }
    float *out = colsum_partial + (int64_t)blockIdx.x * c;
    for (int j = threadIdx.x; j < c; j += kThreads) out[j] = colsum[j];
}

// Few classes (c <= 128: the community / closest_aa tasks, 21 classes at config C3): one WARP per row, the row lives in
// registers (4 columns per lane), 8 rows per CTA in flight.  Same outputs, same fixed summation orders.
constexpr int kSmallC = 128, kSmallThreads = 256;
__global__ void __launch_bounds__(kSmallThreads) softmax_nll_small_kernel(float *__restrict__ logits, int64_t ld, int64_t n, int c,
                                                                          const int64_t *__restrict__ labels, float grad_scale,
                                                                          float *__restrict__ row_loss, float *__restrict__ colsum_partial) {
    __shared__ float cs[kSmallThreads / 32][kSmallC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t warps = (int64_t)gridDim.x * (kSmallThreads / 32);
    for (int64_t r = (int64_t)blockIdx.x * (kSmallThreads / 32) + warp; r < n; r += warps) {
        float *x = logits + r * ld;
        const int64_t y = labels[r];
        if (y < 0 || y >= c) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (lane + 32 * k < c) x[lane + 32 * k] = 0.f;
            if (lane == 0) row_loss[r] = 0.f;
            continue;
        }
        float v[4], m = -INFINITY, xy = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = lane + 32 * k < c ? x[lane + 32 * k] : -INFINITY;
            m = fmaxf(m, v[k]);
            if (lane + 32 * k == (int)y) xy = v[k];  // only the lane that owns column y keeps it
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = lane + 32 * k < c ? expf(v[k] - m) : 0.f;
            sum += v[k];
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
        const float inv = 1.f / sum;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = lane + 32 * k;
            if (j < c) {
                const float g = (v[k] * inv - (j == (int)y ? 1.f : 0.f)) * grad_scale;
                x[j] = g;
                acc[k] += g;
            }
        }
        if (lane == (int)(y & 31)) row_loss[r] = (m - xy) + logf(sum);  // -log_softmax(x)[y]
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) cs[warp][lane + 32 * k] = acc[k];
    __syncthreads();
    for (int j = threadIdx.x; j < c; j += kSmallThreads) {
        float t = cs[0][j];
#pragma unroll
        for (int w = 1; w < kSmallThreads / 32; ++w) t += cs[w][j];  // fixed order
        colsum_partial[(int64_t)blockIdx.x * c + j] = t;
    }
}

// colsum[j] = sum_p partial[p][j]: a CTA owns 32 columns, its 8 warps take every 8th partial (coalesced 128 B
// rows), the 8 sub-sums are combined in warp order -> fixed summation order
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float *__restrict__ partial, int parts, int c,
                                                            float *__restrict__ colsum) {
    __shared__ float sub[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (j < c)
        for (int p = warp; p < parts; p += 8) s += partial[(int64_t)p * c + j];
    sub[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && j < c) {
        float t = sub[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) t += sub[w][lane];
        colsum[j] = t;
    }
}

// loss = grad_scale * sum_r row_loss[r], one CTA, fixed order
__global__ void __launch_bounds__(kThreads) loss_reduce_kernel(const float *__restrict__ row_loss, int64_t n, float grad_scale,
                                                               float *__restrict__ loss) {
    __shared__ float red[kThreads / 32];
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < n; r += kThreads) s += (double)row_loss[r];
    const float tot = block_reduce((float)s, red, false);
    if (threadIdx.x == 0) *loss = tot * grad_scale;
}

int nll_grid(int64_t n, int c) {
    if (c <= kSmallC) {  // warp per row
        const int64_t g = pg_ceil_div(n, kSmallThreads / 32);
        return (int)(g < (int64_t)PG_NUM_SMS * 8 ? g : (int64_t)PG_NUM_SMS * 8);
    }
    const size_t smem = (size_t)c * 8;
    int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    int64_t g = (int64_t)PG_NUM_SMS * per_sm;
    if (g > n) g = n;
    return (int)g;
}
}  // namespace

extern "C" size_t pg_softmax_nll_ws_bytes(int64_t n, int64_t c) {
    if (n < 1 || c < 1 || c > PG_SOFTMAX_NLL_MAX_CLASSES) return 0;
    return pg_align_up((size_t)nll_grid(n, (int)c) * (size_t)c * sizeof(float), 256) + 256;
}

extern "C" int pg_softmax_nll(float *d_logits, int64_t ld, int64_t n, int64_t c, const int64_t *d_labels, float grad_scale,
                              float *d_row_loss, float *d_colsum, float *d_loss, void *d_ws, size_t ws_bytes,
                              pg_stream_t stream) {
    PG_CHECK_ARG(n >= 0 && c >= 1 && ld >= c, "pg_softmax_nll: bad shape");
    if (c > PG_SOFTMAX_NLL_MAX_CLASSES) {
        pg_set_error("pg_softmax_nll: %lld classes exceed the shared-memory row cache (max %d)", (long long)c,
                     PG_SOFTMAX_NLL_MAX_CLASSES);
        return PG_ERANGE;
    }
    PG_CHECK_ARG(d_row_loss && d_colsum && d_loss, "pg_softmax_nll: null output");
    cudaStream_t st = pg_cu(stream);
    if (n == 0) {
        PG_CUDA_CALL(cudaMemsetAsync(d_colsum, 0, (size_t)c * sizeof(float), st));
        PG_CUDA_CALL(cudaMemsetAsync(d_loss, 0, sizeof(float), st));
        return PG_OK;
    }
    PG_CHECK_ARG(d_logits && d_labels && d_ws, "pg_softmax_nll: null buffer");
    const int grid = nll_grid(n, (int)c);
    if (ws_bytes < (size_t)grid * (size_t)c * sizeof(float)) {
        pg_set_error("pg_softmax_nll: workspace too small (%zu < %zu)", ws_bytes, pg_softmax_nll_ws_bytes(n, c));
        return PG_EWORKSPACE;
    }
    if (c <= kSmallC) {
        softmax_nll_small_kernel<<<grid, kSmallThreads, 0, st>>>(d_logits, ld, n, (int)c, d_labels, grad_scale, d_row_loss, (float *)d_ws);
        PG_CUDA_LAUNCH_CHECK("softmax_nll_small_kernel");
    } else {
        const size_t smem = (size_t)c * 8;
        PG_CUDA_CALL(cudaFuncSetAttribute(softmax_nll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        softmax_nll_kernel<<<grid, kThreads, smem, st>>>(d_logits, ld, n, (int)c, d_labels, grad_scale, d_row_loss, (float *)d_ws);
        PG_CUDA_LAUNCH_CHECK("softmax_nll_kernel");
    }
    const int g2 = (int)pg_ceil_div(c, 32);
    colsum_reduce_kernel<<<g2, 256, 0, st>>>((const float *)d_ws, grid, (int)c, d_colsum);
    PG_CUDA_LAUNCH_CHECK("colsum_reduce_kernel");
    loss_reduce_kernel<<<1, kThreads, 0, st>>>(d_row_loss, n, grad_scale, d_loss);
    PG_CUDA_LAUNCH_CHECK("loss_reduce_kernel");
    return PG_OK;
}

// -------------------------------------------------------------------------------------------------------------------------
// Row f1, the two gradient GEMMs of the decoder's output layer (reference protgram_directgcn.py:177-180 under autograd):
//     dd  [N x K] = g  [N x C] @ W2 [C x K]          (input gradient of the layer)
//     dW2 [C x K] = g^T        @ d  [N x K]          (weight gradient)
// with g = d(loss)/d(logits), the N x C matrix pg_softmax_nll leaves behind (282 MB at config C2, K = 32).  Two library
// SGEMMs read g twice; this kernel streams it ONCE: a CTA owns a (row group x column group) block of g, walks it in 64 x 64
// tiles through shared memory, and every tile feeds both products -- lane k of a warp owns output column k (K = 32 q), the g
// element is a shared-memory broadcast, so a tile costs 2 FMAs per element and lane with ~0.3 LDS per FMA.
//   dd : accumulated in registers over the CTA's column tiles, written per row strip to dd_part[column group]
//   dW2: accumulated in shared memory over the CTA's row strips, written once to dw_part[row group]
// two small kernels sum the parts in fixed order (bitwise reproducible).  Any N, C; K a multiple of 32 up to 128.
// -------------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int DG_T = 64;            // tile edge
constexpr int DG_LD = DG_T + 4;     // padded row of the g tile (16-byte aligned rows)
constexpr int DG_THREADS = 256;

struct DecoderGradPlan {
    int row_groups, col_groups;
    int64_t rows_per_group;   // multiple of 64
    int cols_per_group;       // multiple of 64
    size_t smem_bytes;
};

inline DecoderGradPlan decoder_grad_plan(int64_t N, int C, int K) {
    DecoderGradPlan p;
    const int cols_cap = ((96 * 1024) / (4 * K)) / DG_T * DG_T;                 // dW accumulator of <= 96 KB
    int cg = (int)pg_ceil_div(C, cols_cap);
    const int cg_want = (int)pg_ceil_div(C, DG_T) < 12 ? (int)pg_ceil_div(C, DG_T) : 12;
    if (cg < cg_want) cg = cg_want;
    p.cols_per_group = (int)(pg_ceil_div(pg_ceil_div(C, cg), DG_T) * DG_T);
    p.col_groups = (int)pg_ceil_div(C, p.cols_per_group);
    int64_t rg = PG_NUM_SMS / p.col_groups;
    if (rg < 1) rg = 1;
    if (rg > pg_ceil_div(N, DG_T)) rg = pg_ceil_div(N, DG_T);
    p.rows_per_group = pg_ceil_div(pg_ceil_div(N, rg), DG_T) * DG_T;
    p.row_groups = (int)pg_ceil_div(N, p.rows_per_group);
    p.smem_bytes = ((size_t)p.cols_per_group * K + 2 * DG_T * DG_LD + 2 * DG_T * K) * sizeof(float);
    return p;
}

template <int KR>   // K = 32 * KR
__global__ void __launch_bounds__(DG_THREADS) decoder_grads_kernel(const float *__restrict__ g, int64_t ldg, const float *__restrict__ d,
                                                                   int64_t ldd, const float *__restrict__ w2, int64_t N, int C,
                                                                   int64_t rows_per_group, int cols_per_group,
                                                                   float *__restrict__ dd_part, float *__restrict__ dw_part, int vec) {
    constexpr int K = 32 * KR;
    extern __shared__ __align__(16) float sm[];
    float *dw_acc = sm;                                         // [cols_per_group][K]
    float *gs = dw_acc + (size_t)cols_per_group * K;            // [2][64][DG_LD]
    float *w2s = gs + 2 * DG_T * DG_LD;                         // [64][K]
    float *ds = w2s + DG_T * K;                                 // [64][K]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_group, r_end = min(N, r_begin + rows_per_group);
    const int c_begin = blockIdx.x * cols_per_group, c_end = min(C, c_begin + cols_per_group);
    for (int i = tid; i < cols_per_group * K; i += DG_THREADS) dw_acc[i] = 0.f;

    auto load_g = [&](int buf, int64_t r0, int c0) {           // 64 x 64 tile, zero beyond the matrix
        float *dst = gs + buf * DG_T * DG_LD;
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            const int idx = tid + rep * DG_THREADS;             // 1024 float4 slots
            const int rr = idx >> 4, cc = (idx & 15) * 4;
            const int64_t r = r0 + rr;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < r_end) {
                const float *src = g + r * ldg + c0 + cc;
                if (vec && c0 + cc + 3 < c_end) v = __ldg(reinterpret_cast<const float4 *>(src));
                else {
                    if (c0 + cc + 0 < c_end) v.x = src[0];
                    if (c0 + cc + 1 < c_end) v.y = src[1];
                    if (c0 + cc + 2 < c_end) v.z = src[2];
                    if (c0 + cc + 3 < c_end) v.w = src[3];
                }
            }
            *reinterpret_cast<float4 *>(dst + rr * DG_LD + cc) = v;
        }
    };
    auto load_rows = [&](float *dst, const float *src, int64_t ld, int64_t r0, int64_t rmax) {   // 64 rows x K, zero beyond rmax
        for (int i = tid; i < DG_T * K; i += DG_THREADS) {
            const int rr = i / K, kk = i - rr * K;
            dst[i] = (r0 + rr < rmax) ? __ldg(src + (r0 + rr) * ld + kk) : 0.f;
        }
    };

    for (int64_t r0 = r_begin; r0 < r_end; r0 += DG_T) {
        __syncthreads();                                        // previous strip's readers of ds / the last g tile are done
        load_rows(ds, d, ldd, r0, r_end);
        float dd_acc[8][KR];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < KR; ++q) dd_acc[i][q] = 0.f;
        load_g(0, r0, c_begin);
        int buf = 0;
        for (int c0 = c_begin; c0 < c_end; c0 += DG_T, buf ^= 1) {
            __syncthreads();                                    // g tile `buf` + ds complete; w2s free again
            load_rows(w2s, w2, K, c0, c_end);
            if (c0 + DG_T < c_end) load_g(buf ^ 1, r0, c0 + DG_T);   // next tile in flight under this tile's FMAs
            __syncthreads();
            const float *gt = gs + buf * DG_T * DG_LD;
            // (A) dd rows 8 warp .. 8 warp + 7 of the strip, all 64 columns of the tile
#pragma unroll 2
            for (int c4 = 0; c4 < DG_T; c4 += 4) {
                float gv[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 t = *reinterpret_cast<const float4 *>(gt + (8 * warp + i) * DG_LD + c4);
                    gv[i][0] = t.x; gv[i][1] = t.y; gv[i][2] = t.z; gv[i][3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int q = 0; q < KR; ++q) {
                        const float wv = w2s[(c4 + j) * K + lane + 32 * q];
#pragma unroll
                        for (int i = 0; i < 8; ++i) dd_acc[i][q] = fmaf(gv[i][j], wv, dd_acc[i][q]);
                    }
                }
            }
            // (B) dW2 columns 8 warp .. 8 warp + 7 of the tile, all 64 rows of the strip
            float dw[8][KR];
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int q = 0; q < KR; ++q) dw[j][q] = 0.f;
#pragma unroll 2
            for (int i = 0; i < DG_T; ++i) {
                const float4 t0 = *reinterpret_cast<const float4 *>(gt + i * DG_LD + 8 * warp);
                const float4 t1 = *reinterpret_cast<const float4 *>(gt + i * DG_LD + 8 * warp + 4);
                const float gj[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                for (int q = 0; q < KR; ++q) {
                    const float dv = ds[i * K + lane + 32 * q];
#pragma unroll
                    for (int j = 0; j < 8; ++j) dw[j][q] = fmaf(gj[j], dv, dw[j][q]);
                }
            }
            const int col_local = (c0 - c_begin) + 8 * warp;    // this warp alone owns these accumulator rows
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int q = 0; q < KR; ++q) dw_acc[(size_t)(col_local + j) * K + lane + 32 * q] += dw[j][q];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t r = r0 + 8 * warp + i;
            if (r < r_end)
#pragma unroll
                for (int q = 0; q < KR; ++q) dd_part[((int64_t)blockIdx.x * N + r) * K + lane + 32 * q] = dd_acc[i][q];
        }
    }
    __syncthreads();
    for (int i = tid; i < (c_end - c_begin) * K; i += DG_THREADS)
        dw_part[((int64_t)blockIdx.y * C + c_begin) * K + i] = dw_acc[i];
}

__global__ void __launch_bounds__(256) sum_parts_kernel(const float *__restrict__ part, int parts, int64_t numel, float scale,
                                                        float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < parts; ++p) s += part[(int64_t)p * numel + i];   // fixed order
        out[i] = s * scale;
    }
}
}  // namespace

extern "C" int pg_decoder_grads_supported(int K) { return (K == 32 || K == 64 || K == 128) ? 1 : 0; }

extern "C" size_t pg_decoder_grads_ws_bytes(int64_t N, int C, int K) {
    if (!pg_decoder_grads_supported(K) || N < 1 || C < 1) return 0;
    const DecoderGradPlan p = decoder_grad_plan(N, C, K);
    return ((size_t)p.col_groups * N + (size_t)p.row_groups * C) * K * sizeof(float) + 256;
}

extern "C" int pg_decoder_grads(const float *d_g, int64_t ldg, const float *d_d, int64_t ldd, const float *d_w2, int64_t N, int C, int K,
                                float scale, float *d_dd, float *d_dw2, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(pg_decoder_grads_supported(K), "pg_decoder_grads: K must be 32, 64 or 128 (got %d)", K);
    PG_CHECK_ARG(N >= 0 && C >= 1 && ldg >= C && ldd >= K, "pg_decoder_grads: bad shape");
    cudaStream_t st = pg_cu(stream);
    if (N == 0) {
        if (d_dw2) PG_CUDA_CALL(cudaMemsetAsync(d_dw2, 0, (size_t)C * K * sizeof(float), st));
        return PG_OK;
    }
    PG_CHECK_ARG(d_g && d_d && d_w2 && d_dd && d_dw2 && d_ws, "pg_decoder_grads: null buffer");
    const DecoderGradPlan p = decoder_grad_plan(N, C, K);
    const size_t need = ((size_t)p.col_groups * N + (size_t)p.row_groups * C) * K * sizeof(float);
    if (ws_bytes < need) {
        pg_set_error("pg_decoder_grads: workspace too small (%zu < %zu)", ws_bytes, need);
        return PG_EWORKSPACE;
    }
    float *dd_part = (float *)d_ws, *dw_part = dd_part + (size_t)p.col_groups * N * K;
    const int vec = (((uintptr_t)d_g & 15) == 0 && ldg % 4 == 0) ? 1 : 0;
    const dim3 grid((unsigned)p.col_groups, (unsigned)p.row_groups, 1);
#define PG_DG_LAUNCH(KR)                                                                                                             \
    do {                                                                                                                             \
        static bool attr_set = false;                                                                                                \
        if (!attr_set) {                                                                                                             \
            PG_CUDA_CALL(cudaFuncSetAttribute(decoder_grads_kernel<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   \
            attr_set = true;                                                                                                         \
        }                                                                                                                            \
        decoder_grads_kernel<KR><<<grid, DG_THREADS, p.smem_bytes, st>>>(d_g, ldg, d_d, ldd, d_w2, N, C, p.rows_per_group,           \
                                                                         p.cols_per_group, dd_part, dw_part, vec);                  \
    } while (0)
    if (K == 32) PG_DG_LAUNCH(1);
    else if (K == 64) PG_DG_LAUNCH(2);
    else PG_DG_LAUNCH(4);
#undef PG_DG_LAUNCH
    PG_CUDA_LAUNCH_CHECK("decoder_grads_kernel");
    const int64_t n_dd = N * K, n_dw = (int64_t)C * K;
    sum_parts_kernel<<<(unsigned)pg_ceil_div(n_dd, 256), 256, 0, st>>>(dd_part, p.col_groups, n_dd, scale, d_dd);
    PG_CUDA_LAUNCH_CHECK("sum_parts_kernel(dd)");
    sum_parts_kernel<<<(unsigned)pg_ceil_div(n_dw, 256), 256, 0, st>>>(dw_part, p.row_groups, n_dw, scale, d_dw2);
    PG_CUDA_LAUNCH_CHECK("sum_parts_kernel(dW2)");
    return PG_OK;
}
