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
