// Hot path B: the dense half of a DirectGCN layer, fused around the collapsed algebra
// (reference src/models/protgram_directgcn.py:93-135 + :210-215):
//
//   A_ext[i,:] = [ a_i Z_in[i] | b_i Z_out[i] | c_i Z_und[i] | X[i] (res) | a_i b_i c_i | 1 (res) ]
//   Y = A_ext @ W_ext (+ X) + constant ;  H = leaky_relu(Y)
//
// A_ext is never materialised: the A-tile loader reads Z / X / the gate vectors and applies the
// gates on the way into shared memory, so 4 Linears + 6 bias adds + 5 gate multiplies + the
// residual add + the activation of the reference are one kernel.  fp32 FFMA throughout (the
// 1e-4 parity bar rules out single-pass TF32; see DESIGN.md "dense transform").
#include "common.cuh"

namespace {

struct AExt {
    const float *z;      // [M, 3*F_in] (ldz)
    const float *x;      // [M, F_in]   (ldx)  (only read when has_res)
    const float *ga, *gb, *gc;
    int gate_stride;     // 1 = per-row vectors, 0 = scalars
    int64_t ldz, ldx, M;
    int F_in, has_res;
    int k_data;          // 3*F_in + (has_res ? F_in : 0)
    int k_ext;           // k_data + 3 + has_res
    bool vec;            // float4 loads legal (F_in % 4 == 0, 16 B aligned rows)

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
    // 4 consecutive k (k0 % 4 == 0)
    __device__ __forceinline__ float4 at4(int64_t i, int k0) const {
        if (vec && i < M && k0 + 3 < k_data) {
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

__device__ __forceinline__ float4 ld4_guard(const float *p, int64_t ld, int64_t r, int64_t nrows, int c0, int ncols, bool vec) {
    if (r >= nrows) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && c0 + 3 < ncols) return __ldg(reinterpret_cast<const float4 *>(p + r * ld + c0));
    float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = (c0 + j < ncols) ? p[r * ld + c0 + j] : 0.f;
    return make_float4(t[0], t[1], t[2], t[3]);
}

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int APAD = 4;

// -------------------------------------------------------------------------------------------
// forward:  H[M, F_out] = epilogue(A_ext @ W_ext)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layer_gemm_fwd_kernel(AExt A, const float *__restrict__ w_ext, int F_out, bool w_vec,
                                                             const float *__restrict__ constant, int64_t ldconst,
                                                             int add_identity, float slope, float *__restrict__ h, int64_t ldh,
                                                             bool h_vec) {
    __shared__ __align__(16) float As[BK][BM + APAD];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;          // 16 x 16 threads, 8 x 4 outputs each
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < A.k_ext; k0 += BK) {
        // A tile: 128 rows x 16 k = 512 float4, two per thread, stored k-major
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
            const int idx = tid + rep * 256;
            const int r = idx >> 2, kq = (idx & 3) * 4;
            const float4 v = A.at4(m0 + r, k0 + kq);
            As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
        }
        {   // B tile: 16 k x 64 n = 256 float4, one per thread
            const int kr = tid >> 4, nq = (tid & 15) * 4;
            const float4 v = ld4_guard(w_ext, F_out, k0 + kr, A.k_ext, n0 + nq, F_out, w_vec);
            *reinterpret_cast<float4 *>(&Bs[kr][nq]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][tm * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][tm * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tn * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    // epilogue: + identity residual + constant, leaky_relu
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + tm * 8 + i;
        if (r >= A.M) continue;
        const int c0 = n0 + tn * 4;
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            float v = acc[i][j];
            if (c < F_out) {
                if (add_identity) v += A.x[r * A.ldx + c];
                if (constant) v += constant[r * ldconst + c];
                if (slope != 1.f) v = v > 0.f ? v : v * slope;
            }
            y[j] = v;
        }
        if (h_vec && c0 + 3 < F_out) {
            *reinterpret_cast<float4 *>(h + r * ldh + c0) = make_float4(y[0], y[1], y[2], y[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < F_out) h[r * ldh + c0 + j] = y[j];
        }
    }
}

// -------------------------------------------------------------------------------------------
// backward (data):  dA[M, k_data] = dY[M, F_out] @ W_ext[:k_data, :]^T   -> dZ (raw) | dXres
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layer_gemm_bwd_data_kernel(const float *__restrict__ dy, int64_t lddy, bool dy_vec,
                                                                  const float *__restrict__ w_ext, bool w_vec, int64_t M, int F_in,
                                                                  int F_out, int k_data, float *__restrict__ dz, int64_t lddz,
                                                                  float *__restrict__ dxres, int64_t lddxres) {
    __shared__ __align__(16) float As[BK][BM + APAD];   // dY tile, k (= f) major
    __shared__ __align__(16) float Bs[BK][BN + APAD];   // W_ext^T tile: Bs[f][kd]
    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;                     // column of dA = row of W_ext
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < F_out; k0 += BK) {
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
            const int idx = tid + rep * 256;
            const int r = idx >> 2, kq = (idx & 3) * 4;
            const float4 v = ld4_guard(dy, lddy, m0 + r, M, k0 + kq, F_out, dy_vec);
            As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
        }
        {   // 64 rows of W_ext (kd) x 16 f = 256 float4, one per thread, stored f-major
            const int r = tid >> 2, kq = (tid & 3) * 4;
            const float4 v = ld4_guard(w_ext, F_out, n0 + r, k_data, k0 + kq, F_out, w_vec);
            Bs[kq + 0][r] = v.x; Bs[kq + 1][r] = v.y; Bs[kq + 2][r] = v.z; Bs[kq + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][tm * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][tm * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tn * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + tm * 8 + i;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tn * 4 + j;
            if (c >= k_data) continue;
            if (c < 3 * F_in) dz[r * lddz + c] = acc[i][j];
            else dxres[r * lddxres + (c - 3 * F_in)] = acc[i][j];
        }
    }
}

// gate gradients + gating of dZ:  one warp per row
//   dgate_v[i] = <dZraw_v[i], Z_v[i]> + <dY[i], W_ext[k_data + v]>;   dZ_v[i] *= gate_v[i]
__global__ void __launch_bounds__(256) gate_grad_kernel(float *__restrict__ dz, int64_t lddz, const float *__restrict__ z, int64_t ldz,
                                                        const float *__restrict__ dy, int64_t lddy, const float *__restrict__ w_ext,
                                                        const float *__restrict__ ga, const float *__restrict__ gb,
                                                        const float *__restrict__ gc, int gate_stride, int64_t M, int F_in, int F_out,
                                                        int k_data, float *__restrict__ dgate) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += warps) {
        float dot[3] = {0.f, 0.f, 0.f};
        const float g[3] = {ga[r * gate_stride], gb[r * gate_stride], gc[r * gate_stride]};
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            for (int f = lane; f < F_in; f += 32) {
                const int64_t o = (int64_t)v * F_in + f;
                const float d = dz[r * lddz + o];
                dot[v] = fmaf(d, z[r * ldz + o], dot[v]);
                dz[r * lddz + o] = d * g[v];
            }
        }
        for (int f = lane; f < F_out; f += 32) {
            const float d = dy[r * lddy + f];
#pragma unroll
            for (int v = 0; v < 3; ++v) dot[v] = fmaf(d, w_ext[(int64_t)(k_data + v) * F_out + f], dot[v]);
        }
#pragma unroll
        for (int v = 0; v < 3; ++v) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) dot[v] += __shfl_xor_sync(0xffffffffu, dot[v], s);
            if (lane == 0) dgate[(int64_t)v * M + r] = dot[v];
        }
    }
}

// -------------------------------------------------------------------------------------------
// backward (weights):  dW_ext[k_ext, F_out] = A_ext^T @ dY, rows split into `splits` slices
// -------------------------------------------------------------------------------------------
constexpr int WM = 64, WN = 64, WR = 16;  // output tile 64 (k_ext) x 64 (f), 16 rows per step

__global__ void __launch_bounds__(256) layer_gemm_bwd_weight_kernel(AExt A, const float *__restrict__ dy, int64_t lddy, bool dy_vec,
                                                                    int F_out, int64_t rows_per_split, float *__restrict__ partial) {
    __shared__ __align__(16) float As[WR][WM];
    __shared__ __align__(16) float Bs[WR][WN];
    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;          // 16 x 16 threads, 4 x 4 outputs each
    const int kx0 = blockIdx.x * WM;
    const int n0 = blockIdx.y * WN;
    const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r_end = min(A.M, r_begin + rows_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    AExt Ac = A;
    Ac.M = r_end;  // rows beyond the slice read as zero
    for (int64_t r0 = r_begin; r0 < r_end; r0 += WR) {
        {   // 16 rows x 64 kx = 256 float4 each
            const int rr = tid >> 4, q = (tid & 15) * 4;
            *reinterpret_cast<float4 *>(&As[rr][q]) = Ac.at4(r0 + rr, kx0 + q);
            *reinterpret_cast<float4 *>(&Bs[rr][q]) = ld4_guard(dy, lddy, r0 + rr, r_end, n0 + q, F_out, dy_vec);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < WR; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][tm * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tn * 4]);
            const float aa[4] = {a.x, a.y, a.z, a.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *out = partial + (int64_t)blockIdx.z * A.k_ext * F_out;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int kx = kx0 + tm * 4 + i;
        if (kx >= A.k_ext) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tn * 4 + j;
            if (c < F_out) out[(int64_t)kx * F_out + c] = acc[i][j];
        }
    }
}

__global__ void __launch_bounds__(256) reduce_splits_kernel(const float *__restrict__ partial, int splits, int64_t numel,
                                                            float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * numel + i];  // fixed order
        out[i] = s;
    }
}

// -------------------------------------------------------------------------------------------
// small elementwise / row kernels
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lrelu_bwd_kernel(const float *__restrict__ dh, const float *__restrict__ h, float slope,
                                                        int64_t numel, float *__restrict__ dy) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x)
        dy[i] = h[i] > 0.f ? dh[i] : dh[i] * slope;
}

__global__ void __launch_bounds__(256) l2_normalize_kernel(const float *__restrict__ h, int64_t ldh, int64_t M, int F, float eps,
                                                           float *__restrict__ out, int64_t ldout) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += warps) {
        float ss = 0.f;
        for (int f = lane; f < F; f += 32) {
            const float v = h[r * ldh + f];
            ss = fmaf(v, v, ss);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
        const float inv = 1.f / (sqrtf(ss) + eps);
        for (int f = lane; f < F; f += 32) out[r * ldout + f] = h[r * ldh + f] * inv;
    }
}

inline bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }
inline unsigned grid_for(int64_t n, int threads = 256, int per_sm = 8) {
    int64_t want = pg_ceil_div(n, threads);
    int64_t cap = (int64_t)PG_NUM_SMS * per_sm;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

AExt make_aext(const float *z, int64_t ldz, const float *x, int64_t ldx, const float *ga, const float *gb, const float *gc,
               int gate_stride, int64_t M, int F_in, int has_res) {
    AExt A;
    A.z = z; A.x = x; A.ga = ga; A.gb = gb; A.gc = gc;
    A.gate_stride = gate_stride; A.ldz = ldz; A.ldx = ldx; A.M = M; A.F_in = F_in; A.has_res = has_res;
    A.k_data = 3 * F_in + (has_res ? F_in : 0);
    A.k_ext = A.k_data + 3 + (has_res ? 1 : 0);
    A.vec = (F_in % 4 == 0) && al16(z) && ldz % 4 == 0 && (!has_res || (al16(x) && ldx % 4 == 0));
    return A;
}

int split_count(int64_t M, int k_ext, int F_out) {
    const int64_t tiles = pg_ceil_div(k_ext, WM) * pg_ceil_div(F_out, WN);
    int64_t s = pg_ceil_div(2 * PG_NUM_SMS, tiles);
    const int64_t max_s = pg_ceil_div(M, 4 * WR);
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 1024) s = 1024;
    return (int)s;
}
}  // namespace

extern "C" int pg_layer_gemm_fwd(const float *d_z, int64_t ldz, const float *d_x, int64_t ldx, const float *d_gate_a,
                                 const float *d_gate_b, const float *d_gate_c, int gate_stride, const float *d_w_ext,
                                 const float *d_constant, int64_t ldconst, int64_t num_rows, int F_in, int F_out, int has_res,
                                 int add_identity, float slope, float *d_h, int64_t ldh, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F_in >= 1 && F_out >= 1 && (gate_stride == 0 || gate_stride == 1), "pg_layer_gemm_fwd: bad shape");
    PG_CHECK_ARG(!(has_res && add_identity), "pg_layer_gemm_fwd: has_res and add_identity are exclusive");
    PG_CHECK_ARG(!add_identity || F_in == F_out, "pg_layer_gemm_fwd: identity residual needs F_in == F_out");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_z && d_gate_a && d_gate_b && d_gate_c && d_w_ext && d_h, "pg_layer_gemm_fwd: null buffer");
    PG_CHECK_ARG(!(has_res || add_identity) || d_x, "pg_layer_gemm_fwd: residual needs x");
    PG_CHECK_ARG(ldz >= 3 * (int64_t)F_in && ldh >= F_out && (!d_constant || ldconst >= F_out), "pg_layer_gemm_fwd: bad stride");
    AExt A = make_aext(d_z, ldz, d_x, ldx, d_gate_a, d_gate_b, d_gate_c, gate_stride, num_rows, F_in, has_res);
    const bool w_vec = F_out % 4 == 0 && al16(d_w_ext);
    const bool h_vec = F_out % 4 == 0 && al16(d_h) && ldh % 4 == 0;
    dim3 grid((unsigned)pg_ceil_div(num_rows, BM), (unsigned)pg_ceil_div(F_out, BN));
    layer_gemm_fwd_kernel<<<grid, 256, 0, pg_cu(stream)>>>(A, d_w_ext, F_out, w_vec, d_constant, ldconst, add_identity, slope, d_h,
                                                           ldh, h_vec);
    PG_CUDA_LAUNCH_CHECK("layer_gemm_fwd_kernel");
    return PG_OK;
}

extern "C" int pg_lrelu_bwd(const float *d_dh, const float *d_h, float slope, int64_t numel, float *d_dy, pg_stream_t stream) {
    PG_CHECK_ARG(numel >= 0, "pg_lrelu_bwd: bad size");
    if (numel == 0) return PG_OK;
    PG_CHECK_ARG(d_dh && d_h && d_dy, "pg_lrelu_bwd: null buffer");
    lrelu_bwd_kernel<<<grid_for(numel), 256, 0, pg_cu(stream)>>>(d_dh, d_h, slope, numel, d_dy);
    PG_CUDA_LAUNCH_CHECK("lrelu_bwd_kernel");
    return PG_OK;
}

extern "C" int pg_layer_gemm_bwd_data(const float *d_dy, int64_t lddy, const float *d_w_ext, const float *d_z, int64_t ldz,
                                      const float *d_gate_a, const float *d_gate_b, const float *d_gate_c, int gate_stride,
                                      int64_t num_rows, int F_in, int F_out, int has_res, float *d_dz, int64_t lddz, float *d_dxres,
                                      int64_t lddxres, float *d_dgate, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F_in >= 1 && F_out >= 1 && (gate_stride == 0 || gate_stride == 1), "pg_layer_gemm_bwd_data: bad shape");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_dy && d_w_ext && d_z && d_gate_a && d_gate_b && d_gate_c && d_dz && d_dgate, "pg_layer_gemm_bwd_data: null buffer");
    PG_CHECK_ARG(!has_res || d_dxres, "pg_layer_gemm_bwd_data: has_res needs d_dxres");
    PG_CHECK_ARG(lddy >= F_out && ldz >= 3 * (int64_t)F_in && lddz >= 3 * (int64_t)F_in && (!has_res || lddxres >= F_in),
                 "pg_layer_gemm_bwd_data: bad stride");
    const int k_data = 3 * F_in + (has_res ? F_in : 0);
    const bool dy_vec = F_out % 4 == 0 && al16(d_dy) && lddy % 4 == 0;
    const bool w_vec = F_out % 4 == 0 && al16(d_w_ext);
    cudaStream_t st = pg_cu(stream);
    dim3 grid((unsigned)pg_ceil_div(num_rows, BM), (unsigned)pg_ceil_div(k_data, BN));
    layer_gemm_bwd_data_kernel<<<grid, 256, 0, st>>>(d_dy, lddy, dy_vec, d_w_ext, w_vec, num_rows, F_in, F_out, k_data, d_dz, lddz,
                                                     d_dxres, lddxres);
    PG_CUDA_LAUNCH_CHECK("layer_gemm_bwd_data_kernel");
    gate_grad_kernel<<<grid_for(num_rows * 32), 256, 0, st>>>(d_dz, lddz, d_z, ldz, d_dy, lddy, d_w_ext, d_gate_a, d_gate_b, d_gate_c,
                                                              gate_stride, num_rows, F_in, F_out, k_data, d_dgate);
    PG_CUDA_LAUNCH_CHECK("gate_grad_kernel");
    return PG_OK;
}

extern "C" size_t pg_layer_gemm_bwd_weight_ws_bytes(int64_t num_rows, int F_in, int F_out, int has_res) {
    const int k_ext = 3 * F_in + (has_res ? F_in : 0) + 3 + (has_res ? 1 : 0);
    return (size_t)split_count(num_rows, k_ext, F_out) * k_ext * F_out * sizeof(float) + 256;
}

extern "C" int pg_layer_gemm_bwd_weight(const float *d_z, int64_t ldz, const float *d_x, int64_t ldx, const float *d_gate_a,
                                        const float *d_gate_b, const float *d_gate_c, int gate_stride, const float *d_dy,
                                        int64_t lddy, int64_t num_rows, int F_in, int F_out, int has_res, float *d_dw_ext,
                                        void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F_in >= 1 && F_out >= 1 && (gate_stride == 0 || gate_stride == 1), "pg_layer_gemm_bwd_weight: bad shape");
    PG_CHECK_ARG(d_dw_ext, "pg_layer_gemm_bwd_weight: null output");
    cudaStream_t st = pg_cu(stream);
    AExt A = make_aext(d_z, ldz, d_x, ldx, d_gate_a, d_gate_b, d_gate_c, gate_stride, num_rows, F_in, has_res);
    const int64_t numel = (int64_t)A.k_ext * F_out;
    if (num_rows == 0) {
        PG_CUDA_CALL(cudaMemsetAsync(d_dw_ext, 0, (size_t)numel * sizeof(float), st));
        return PG_OK;
    }
    PG_CHECK_ARG(d_z && d_gate_a && d_gate_b && d_gate_c && d_dy && d_ws && (!has_res || d_x), "pg_layer_gemm_bwd_weight: null buffer");
    const int splits = split_count(num_rows, A.k_ext, F_out);
    if (ws_bytes < (size_t)splits * numel * sizeof(float)) {
        pg_set_error("pg_layer_gemm_bwd_weight: workspace too small (%zu < %zu)", ws_bytes, (size_t)splits * numel * sizeof(float));
        return PG_EWORKSPACE;
    }
    int64_t rows_per_split = pg_ceil_div(num_rows, splits);
    rows_per_split = pg_ceil_div(rows_per_split, WR) * WR;
    const bool dy_vec = F_out % 4 == 0 && al16(d_dy) && lddy % 4 == 0;
    dim3 grid((unsigned)pg_ceil_div(A.k_ext, WM), (unsigned)pg_ceil_div(F_out, WN), (unsigned)splits);
    layer_gemm_bwd_weight_kernel<<<grid, 256, 0, st>>>(A, d_dy, lddy, dy_vec, F_out, rows_per_split, (float *)d_ws);
    PG_CUDA_LAUNCH_CHECK("layer_gemm_bwd_weight_kernel");
    reduce_splits_kernel<<<grid_for(numel), 256, 0, st>>>((const float *)d_ws, splits, numel, d_dw_ext);
    PG_CUDA_LAUNCH_CHECK("reduce_splits_kernel");
    return PG_OK;
}

extern "C" int pg_l2_normalize_rows(const float *d_h, int64_t ldh, int64_t num_rows, int F, float eps, float *d_out, int64_t ldout,
                                    pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F >= 1 && ldh >= F && ldout >= F, "pg_l2_normalize_rows: bad shape");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_h && d_out, "pg_l2_normalize_rows: null buffer");
    l2_normalize_kernel<<<grid_for(num_rows * 32), 256, 0, pg_cu(stream)>>>(d_h, ldh, num_rows, F, eps, d_out, ldout);
    PG_CUDA_LAUNCH_CHECK("l2_normalize_kernel");
    return PG_OK;
}
