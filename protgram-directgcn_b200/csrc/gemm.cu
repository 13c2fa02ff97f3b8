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

constexpr int BN = 64, BK = 16, PAD = 4;

// -------------------------------------------------------------------------------------------
// One SIMT fp32 mainloop for the three GEMMs of a layer: C[m, n] = sum_k A(m, k) * B(k, n).
// 256 threads, (BM x 64) output tile, BK = 16, operands staged k-major in shared memory with a
// register-prefetched double buffer (one __syncthreads per k-tile).  The operand functors do the
// layout work (gating, virtual columns, transposes), the epilogue functors the fusion.
// -------------------------------------------------------------------------------------------
struct OpA_Ext {  // forward: A(m, k) = A_ext[m, k]
    AExt A;
    template <int BM_>
    __device__ __forceinline__ void load(float4 (&r)[BM_ / 64], int64_t m0, int64_t k0, int tid) const {
#pragma unroll
        for (int rep = 0; rep < BM_ / 64; ++rep) {
            const int idx = tid + rep * 256;
            r[rep] = A.at4(m0 + (idx >> 2), (int)k0 + (idx & 3) * 4);
        }
    }
    template <int BM_>
    __device__ __forceinline__ void store(float (*As)[BM_ + PAD], const float4 (&r)[BM_ / 64], int tid) const {
#pragma unroll
        for (int rep = 0; rep < BM_ / 64; ++rep) {
            const int idx = tid + rep * 256;
            const int m = idx >> 2, kq = (idx & 3) * 4;
            As[kq + 0][m] = r[rep].x; As[kq + 1][m] = r[rep].y; As[kq + 2][m] = r[rep].z; As[kq + 3][m] = r[rep].w;
        }
    }
};

struct OpA_Plain {  // backward data: A(m, k) = dY[m, k]
    const float *p;
    int64_t ld, M;
    int ncols;
    bool vec;
    template <int BM_>
    __device__ __forceinline__ void load(float4 (&r)[BM_ / 64], int64_t m0, int64_t k0, int tid) const {
#pragma unroll
        for (int rep = 0; rep < BM_ / 64; ++rep) {
            const int idx = tid + rep * 256;
            r[rep] = ld4_guard(p, ld, m0 + (idx >> 2), M, (int)k0 + (idx & 3) * 4, ncols, vec);
        }
    }
    template <int BM_>
    __device__ __forceinline__ void store(float (*As)[BM_ + PAD], const float4 (&r)[BM_ / 64], int tid) const {
#pragma unroll
        for (int rep = 0; rep < BM_ / 64; ++rep) {
            const int idx = tid + rep * 256;
            const int m = idx >> 2, kq = (idx & 3) * 4;
            As[kq + 0][m] = r[rep].x; As[kq + 1][m] = r[rep].y; As[kq + 2][m] = r[rep].z; As[kq + 3][m] = r[rep].w;
        }
    }
};

struct OpA_ExtT {  // backward weights: A(m = ext column, k = node row) = A_ext[k, m]
    AExt A;        // A.M = end of this split's row range (rows beyond read as zero)
    template <int BM_>
    __device__ __forceinline__ void load(float4 (&r)[BM_ / 64], int64_t m0, int64_t k0, int tid) const {
#pragma unroll
        for (int rep = 0; rep < BM_ / 64; ++rep) {
            const int idx = tid + rep * 256;
            r[rep] = A.at4(k0 + idx / (BM_ / 4), (int)m0 + (idx % (BM_ / 4)) * 4);
        }
    }
    template <int BM_>
    __device__ __forceinline__ void store(float (*As)[BM_ + PAD], const float4 (&r)[BM_ / 64], int tid) const {
#pragma unroll
        for (int rep = 0; rep < BM_ / 64; ++rep) {
            const int idx = tid + rep * 256;
            *reinterpret_cast<float4 *>(&As[idx / (BM_ / 4)][(idx % (BM_ / 4)) * 4]) = r[rep];
        }
    }
};

struct OpB_Rows {  // B(k, n) = p[k, n]  (W_ext for the forward, dY for the weight gradient)
    const float *p;
    int64_t ld, nrows;
    int ncols;
    bool vec;
    __device__ __forceinline__ void load(float4 &r, int n0, int64_t k0, int tid) const {
        r = ld4_guard(p, ld, k0 + (tid >> 4), nrows, n0 + (tid & 15) * 4, ncols, vec);
    }
    __device__ __forceinline__ void store(float (*Bs)[BN + PAD], const float4 &r, int tid) const {
        *reinterpret_cast<float4 *>(&Bs[tid >> 4][(tid & 15) * 4]) = r;
    }
};

struct OpB_Trans {  // B(k, n) = p[n, k]  (W_ext^T for the data gradient)
    const float *p;
    int64_t ld, nrows;  // nrows = extent of n
    int ncols;          // extent of k
    bool vec;
    __device__ __forceinline__ void load(float4 &r, int n0, int64_t k0, int tid) const {
        r = ld4_guard(p, ld, n0 + (tid >> 2), nrows, (int)k0 + (tid & 3) * 4, ncols, vec);
    }
    __device__ __forceinline__ void store(float (*Bs)[BN + PAD], const float4 &r, int tid) const {
        const int n = tid >> 2, kq = (tid & 3) * 4;
        Bs[kq + 0][n] = r.x; Bs[kq + 1][n] = r.y; Bs[kq + 2][n] = r.z; Bs[kq + 3][n] = r.w;
    }
};

struct Epi_Fwd {  // H = leaky_relu(acc (+ X) + constant)
    const float *x;
    int64_t ldx;
    const float *constant;
    int64_t ldconst;
    float *h;
    int64_t ldh, M;
    int F_out, add_identity;
    float slope;
    bool h_vec;
    __device__ __forceinline__ void operator()(const float (&acc)[4], int64_t r, int c0, int) const {
        if (r >= M) return;
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            float v = acc[j];
            if (c < F_out) {
                if (add_identity) v += x[r * ldx + c];
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
};

struct Epi_BwdData {  // dA columns -> dZ (raw, gated later) | dXres
    float *dz;
    int64_t lddz;
    float *dxres;
    int64_t lddxres, M;
    int F_in, k_data;
    __device__ __forceinline__ void operator()(const float (&acc)[4], int64_t r, int c0, int) const {
        if (r >= M) return;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            if (c >= k_data) continue;
            if (c < 3 * F_in) dz[r * lddz + c] = acc[j];
            else dxres[r * lddxres + (c - 3 * F_in)] = acc[j];
        }
    }
};

struct Epi_Partial {  // split-K partial of dW_ext
    float *partial;
    int rows, cols;  // k_ext, F_out
    __device__ __forceinline__ void operator()(const float (&acc)[4], int64_t r, int c0, int z) const {
        if (r >= rows) return;
        float *out = partial + (int64_t)z * rows * cols + r * cols;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c0 + j < cols) out[c0 + j] = acc[j];
    }
};

// ---- plain Linear layers (the decoder MLP of reference protgram_directgcn.py:173-180; row f1) on the same mainloop ----
struct OpA_PlainT {  // A(m, k) = p[k, m]: the transposed operand of a weight gradient (m = column of p, k = row of p)
    const float *p;
    int64_t ld, nrows;   // nrows = extent of k (rows beyond read as zero: split-K tails)
    int ncols;           // extent of m
    bool vec;
    template <int BM_>
    __device__ __forceinline__ void load(float4 (&r)[BM_ / 64], int64_t m0, int64_t k0, int tid) const {
#pragma unroll
        for (int rep = 0; rep < BM_ / 64; ++rep) {
            const int idx = tid + rep * 256;
            r[rep] = ld4_guard(p, ld, k0 + idx / (BM_ / 4), nrows, (int)m0 + (idx % (BM_ / 4)) * 4, ncols, vec);
        }
    }
    template <int BM_>
    __device__ __forceinline__ void store(float (*As)[BM_ + PAD], const float4 (&r)[BM_ / 64], int tid) const {
#pragma unroll
        for (int rep = 0; rep < BM_ / 64; ++rep) {
            const int idx = tid + rep * 256;
            *reinterpret_cast<float4 *>(&As[idx / (BM_ / 4)][(idx % (BM_ / 4)) * 4]) = r[rep];
        }
    }
};

struct Epi_BiasAct {  // out = act(acc + bias)
    const float *bias;
    float *out;
    int64_t ldo, M;
    int ncols, relu;
    __device__ __forceinline__ void operator()(const float (&acc)[4], int64_t r, int c0, int) const {
        if (r >= M) return;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            if (c < ncols) {
                float v = acc[j] + (bias ? bias[c] : 0.f);
                if (relu) v = v > 0.f ? v : 0.f;
                out[r * ldo + c] = v;
            }
        }
    }
};

// column sums of a row-major matrix in fixed order: CTA (x, y) sums 32 columns over row slice y (8 row-strided partials per column
// combined by the first warp), a second pass adds the slices in order
__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ g, int64_t ld, int64_t M, int ncols, int64_t rows_per_slice,
                                                     float *__restrict__ partial) {
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int slice = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice, r1 = min(M, r0 + rows_per_slice);
    float acc = 0.f;
    if (c < ncols)
        for (int64_t r = r0 + slice; r < r1; r += 8) acc += g[r * ld + c];
    part[slice][threadIdx.x & 31] = acc;
    __syncthreads();
    if (slice == 0 && c < ncols) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x & 31];
        partial[(int64_t)blockIdx.y * ncols + c] = t;
    }
}

// ---- operands / epilogues of the regrouped backward on the SIMT path (the tensor-core path has its own in gemm_tc.cu):
//   gate gradients   dgate_v[i] = <dY[i] W'_v^T, Z_v[i]> + <dY[i], beta_v>      (data-gradient GEMM with a dot-product epilogue:
//                                                                               the 3 F_in-wide dZ is never written)
//   input gradient   dX = [T_in | T_out | T_und | dY] @ [W'_in^T; W'_out^T; W'_und^T; W_res^T] (+ dY),  T_v = A_v^T (g_v * dY)
struct OpA_Cat2 {  // A(m, k) = k < k_first ? p0[m, k] : p1[m, k - k_first]
    const float *p0;
    int64_t ld0;
    int k_first;
    const float *p1;
    int64_t ld1;
    int k_total;
    int64_t M;
    bool vec;      // k_first % 4 == 0 and both matrices 16 B aligned with row strides % 4 == 0
    __device__ __forceinline__ float4 at4(int64_t m, int k0) const {
        if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec && k0 + 3 < k_total)
            return k0 < k_first ? __ldg(reinterpret_cast<const float4 *>(p0 + m * ld0 + k0))
                                : __ldg(reinterpret_cast<const float4 *>(p1 + m * ld1 + (k0 - k_first)));
        float t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            t[j] = k >= k_total ? 0.f : (k < k_first ? p0[m * ld0 + k] : p1[m * ld1 + (k - k_first)]);
        }
        return make_float4(t[0], t[1], t[2], t[3]);
    }
    template <int BM_>
    __device__ __forceinline__ void load(float4 (&r)[BM_ / 64], int64_t m0, int64_t k0, int tid) const {
#pragma unroll
        for (int rep = 0; rep < BM_ / 64; ++rep) {
            const int idx = tid + rep * 256;
            r[rep] = at4(m0 + (idx >> 2), (int)k0 + (idx & 3) * 4);
        }
    }
    template <int BM_>
    __device__ __forceinline__ void store(float (*As)[BM_ + PAD], const float4 (&r)[BM_ / 64], int tid) const {
#pragma unroll
        for (int rep = 0; rep < BM_ / 64; ++rep) {
            const int idx = tid + rep * 256;
            const int m = idx >> 2, kq = (idx & 3) * 4;
            As[kq + 0][m] = r[rep].x; As[kq + 1][m] = r[rep].y; As[kq + 2][m] = r[rep].z; As[kq + 3][m] = r[rep].w;
        }
    }
};

struct OpB_WextBlocksT {  // B(k = v * F_out + o, n) = W_ext[v * F_in + n, o]   (v < blocks; n < F_in)
    const float *w;       // [k_ext, F_out] row-major
    int F_in, F_out, blocks;
    bool vec;             // F_out % 4 == 0, w 16 B aligned
    __device__ __forceinline__ void load(float4 &r, int n0, int64_t k0, int tid) const {
        const int n = n0 + (tid >> 2);
        const int k = (int)k0 + (tid & 3) * 4;
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < F_in) {
            if (vec && k + 3 < blocks * F_out) {
                const int v = k / F_out, o = k - v * F_out;
                r = __ldg(reinterpret_cast<const float4 *>(w + ((int64_t)v * F_in + n) * F_out + o));
                return;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kk = k + j;
                if (kk < blocks * F_out) {
                    const int v = kk / F_out, o = kk - v * F_out;
                    t[j] = w[((int64_t)v * F_in + n) * F_out + o];
                }
            }
        }
        r = make_float4(t[0], t[1], t[2], t[3]);
    }
    __device__ __forceinline__ void store(float (*Bs)[BN + PAD], const float4 &r, int tid) const {
        const int n = tid >> 2, kq = (tid & 3) * 4;
        Bs[kq + 0][n] = r.x; Bs[kq + 1][n] = r.y; Bs[kq + 2][n] = r.z; Bs[kq + 3][n] = r.w;
    }
};

struct Epi_Dx {  // dX = acc (+ dY)
    float *dx;
    int64_t lddx;
    const float *dy;
    int64_t lddy, M;
    int F_in, add_identity;
    __device__ __forceinline__ void operator()(const float (&acc)[4], int64_t r, int c0, int) const {
        if (r >= M) return;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            if (c < F_in) dx[r * lddx + c] = acc[j] + (add_identity ? dy[r * lddy + c] : 0.f);
        }
    }
};

struct Epi_GateDot {  // per row and 64-column block: the three <dA_v[i], Z_v[i]> parts (fixed-order reduction over the blocks afterwards)
    const float *z;
    int64_t ldz, M;
    int F_in;
    float *partial;      // [gridDim.y][3][M]
    __device__ __forceinline__ void operator()(const float (&acc)[4], int64_t r, int c0, int) const {
        float dot[3] = {0.f, 0.f, 0.f};
        if (r < M) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j;
                if (c < 3 * F_in) {
                    const float p = acc[j] * z[r * ldz + c];
                    const int v = c / F_in;
                    dot[0] += v == 0 ? p : 0.f;
                    dot[1] += v == 1 ? p : 0.f;
                    dot[2] += v == 2 ? p : 0.f;
                }
            }
        }
        // the 16 threads that share a row of the tile are 16 consecutive lanes (tid = tm * 16 + tn)
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int s = 8; s > 0; s >>= 1) dot[v] += __shfl_xor_sync(0xffffffffu, dot[v], s);
        if ((threadIdx.x & 15) == 0 && r < M) {
#pragma unroll
            for (int v = 0; v < 3; ++v) partial[((int64_t)blockIdx.y * 3 + v) * M + r] = dot[v];
        }
    }
};

// dgate[v][r] = sum over column blocks of partial + <dY[r], W_ext[k_data + v]>   (one warp per row, fixed order)
__global__ void __launch_bounds__(256) gate_dot_reduce_kernel(const float *__restrict__ partial, int blocks, const float *__restrict__ dy,
                                                              int64_t lddy, const float *__restrict__ w_ext, int64_t M, int F_out,
                                                              int k_data, float *__restrict__ dgate) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += warps) {
        float dot[3] = {0.f, 0.f, 0.f};
        for (int f = lane; f < F_out; f += 32) {
            const float d = dy[r * lddy + f];
#pragma unroll
            for (int v = 0; v < 3; ++v) dot[v] = fmaf(d, w_ext[(int64_t)(k_data + v) * F_out + f], dot[v]);
        }
#pragma unroll
        for (int v = 0; v < 3; ++v) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) dot[v] += __shfl_xor_sync(0xffffffffu, dot[v], s);
            if (lane == 0) {
                float acc = 0.f;
                for (int b = 0; b < blocks; ++b) acc += partial[((int64_t)b * 3 + v) * M + r];
                dgate[(int64_t)v * M + r] = acc + dot[v];
            }
        }
    }
}

template <int BM_, class OpA, class OpB, class Epi>
__global__ void __launch_bounds__(256) gemm_kernel(OpA opa, OpB opb, Epi epi, int64_t k_total, int64_t k_per_z) {
    constexpr int TM = BM_ / 16;
    __shared__ __align__(16) float As[2][BK][BM_ + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;  // 16 x 16 threads, TM x 4 outputs each
    const int64_t m0 = (int64_t)blockIdx.x * BM_;
    const int n0 = blockIdx.y * BN;
    const int64_t kb = (int64_t)blockIdx.z * k_per_z;
    const int64_t ke = min(k_total, kb + k_per_z);
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float4 ra[BM_ / 64], rb;
    if (kb < ke) {
        opa.template load<BM_>(ra, m0, kb, tid);
        opb.load(rb, n0, kb, tid);
        opa.template store<BM_>(As[0], ra, tid);
        opb.store(Bs[0], rb, tid);
    }
    __syncthreads();
    int cur = 0;
    for (int64_t k0 = kb; k0 < ke; k0 += BK) {
        const bool more = k0 + BK < ke;
        if (more) {
            opa.template load<BM_>(ra, m0, k0 + BK, tid);
            opb.load(rb, n0, k0 + BK, tid);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 av = *reinterpret_cast<const float4 *>(&As[cur][k][tm * TM + i]);
                a[i] = av.x; a[i + 1] = av.y; a[i + 2] = av.z; a[i + 3] = av.w;
            }
            const float4 bv = *reinterpret_cast<const float4 *>(&Bs[cur][k][tn * 4]);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (more) {
            opa.template store<BM_>(As[cur ^ 1], ra, tid);
            opb.store(Bs[cur ^ 1], rb, tid);
        }
        __syncthreads();
        cur ^= 1;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) epi(acc[i], m0 + tm * TM + i, n0 + tn * 4, (int)blockIdx.z);
}

template <class OpA, class OpB, class Epi>
int launch_gemm(OpA opa, OpB opb, Epi epi, int64_t M, int64_t N, int64_t k_total, int64_t k_per_z, int z, cudaStream_t st) {
    const int64_t tiles128 = pg_ceil_div(M, 128) * pg_ceil_div(N, BN) * z;
    if (tiles128 >= 2 * PG_NUM_SMS) {
        dim3 grid((unsigned)pg_ceil_div(M, 128), (unsigned)pg_ceil_div(N, BN), (unsigned)z);
        gemm_kernel<128, OpA, OpB, Epi><<<grid, 256, 0, st>>>(opa, opb, epi, k_total, k_per_z);
    } else {  // small problems: more, smaller tiles to fill 148 SMs
        dim3 grid((unsigned)pg_ceil_div(M, 64), (unsigned)pg_ceil_div(N, BN), (unsigned)z);
        gemm_kernel<64, OpA, OpB, Epi><<<grid, 256, 0, st>>>(opa, opb, epi, k_total, k_per_z);
    }
    PG_CUDA_LAUNCH_CHECK("gemm_kernel");
    return PG_OK;
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
    const int64_t tiles = pg_ceil_div(k_ext, 64) * pg_ceil_div(F_out, BN);
    int64_t s = pg_ceil_div(2 * PG_NUM_SMS, tiles);
    const int64_t max_s = pg_ceil_div(M, 4 * BK);
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 1024) s = 1024;
    return (int)s;
}
}  // namespace

// shared with the tensor-core data gradient (gemm_tc.cu): gating of dZ + gate gradients
int pg_launch_gate_grad(float *d_dz, int64_t lddz, const float *d_z, int64_t ldz, const float *d_dy, int64_t lddy, const float *d_w_ext,
                        const float *d_gate_a, const float *d_gate_b, const float *d_gate_c, int gate_stride, int64_t num_rows, int F_in,
                        int F_out, int k_data, float *d_dgate, cudaStream_t st) {
    gate_grad_kernel<<<grid_for(num_rows * 32), 256, 0, st>>>(d_dz, lddz, d_z, ldz, d_dy, lddy, d_w_ext, d_gate_a, d_gate_b, d_gate_c,
                                                              gate_stride, num_rows, F_in, F_out, k_data, d_dgate);
    PG_CUDA_LAUNCH_CHECK("gate_grad_kernel");
    return PG_OK;
}

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
    OpA_Ext opa{A};
    OpB_Rows opb{d_w_ext, F_out, A.k_ext, F_out, w_vec};
    Epi_Fwd epi{d_x, ldx, d_constant, ldconst, d_h, ldh, num_rows, F_out, add_identity, slope, h_vec};
    return launch_gemm(opa, opb, epi, num_rows, F_out, A.k_ext, A.k_ext, 1, pg_cu(stream));
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
    {
        OpA_Plain opa{d_dy, lddy, num_rows, F_out, dy_vec};
        OpB_Trans opb{d_w_ext, F_out, k_data, F_out, w_vec};
        Epi_BwdData epi{d_dz, lddz, d_dxres, lddxres, num_rows, F_in, k_data};
        int rc = launch_gemm(opa, opb, epi, num_rows, k_data, F_out, F_out, 1, st);
        if (rc != PG_OK) return rc;
    }
    return pg_launch_gate_grad(d_dz, lddz, d_z, ldz, d_dy, lddy, d_w_ext, d_gate_a, d_gate_b, d_gate_c, gate_stride, num_rows, F_in, F_out,
                               k_data, d_dgate, st);
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
    rows_per_split = pg_ceil_div(rows_per_split, BK) * BK;
    const bool dy_vec = F_out % 4 == 0 && al16(d_dy) && lddy % 4 == 0;
    {
        OpA_ExtT opa{A};
        OpB_Rows opb{d_dy, lddy, num_rows, F_out, dy_vec};
        Epi_Partial epi{(float *)d_ws, A.k_ext, F_out};
        int rc = launch_gemm(opa, opb, epi, A.k_ext, F_out, num_rows, rows_per_split, splits, st);
        if (rc != PG_OK) return rc;
    }
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

// ---- regrouped backward, SIMT path (same contracts as pg_layer_gate_grad_tc / pg_layer_gemm_bwd_dx_tc) ----
extern "C" size_t pg_layer_gate_grad_ws_bytes(int64_t num_rows, int F_in, int F_out) {
    (void)F_out;
    return (size_t)pg_ceil_div(3 * (int64_t)F_in, BN) * 3 * (size_t)num_rows * sizeof(float) + 256;
}

extern "C" int pg_layer_gate_grad(const float *d_dy, int64_t lddy, const float *d_w_ext, const float *d_z, int64_t ldz,
                                  int64_t num_rows, int F_in, int F_out, int has_res, float *d_dgate, void *d_ws, size_t ws_bytes,
                                  pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F_in >= 1 && F_out >= 1 && lddy >= F_out && ldz >= 3 * (int64_t)F_in, "pg_layer_gate_grad: bad shape");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_dy && d_w_ext && d_z && d_dgate && d_ws, "pg_layer_gate_grad: null buffer");
    if (ws_bytes < pg_layer_gate_grad_ws_bytes(num_rows, F_in, F_out) - 256) {
        pg_set_error("pg_layer_gate_grad: workspace too small (%zu bytes)", ws_bytes);
        return PG_EWORKSPACE;
    }
    const int k_data = 3 * F_in + (has_res ? F_in : 0);
    const int blocks = (int)pg_ceil_div(3 * (int64_t)F_in, BN);
    const bool dy_vec = F_out % 4 == 0 && al16(d_dy) && lddy % 4 == 0;
    const bool w_vec = F_out % 4 == 0 && al16(d_w_ext);
    cudaStream_t st = pg_cu(stream);
    OpA_Plain opa{d_dy, lddy, num_rows, F_out, dy_vec};
    OpB_Trans opb{d_w_ext, F_out, 3 * F_in, F_out, w_vec};
    Epi_GateDot epi{d_z, ldz, num_rows, F_in, (float *)d_ws};
    int rc = launch_gemm(opa, opb, epi, num_rows, 3 * (int64_t)F_in, F_out, F_out, 1, st);
    if (rc != PG_OK) return rc;
    gate_dot_reduce_kernel<<<grid_for(num_rows * 32), 256, 0, st>>>((const float *)d_ws, blocks, d_dy, lddy, d_w_ext, num_rows, F_out, k_data,
                                                                   d_dgate);
    PG_CUDA_LAUNCH_CHECK("gate_dot_reduce_kernel");
    return PG_OK;
}

extern "C" int pg_layer_gemm_bwd_dx(const float *d_t, int64_t ldt, const float *d_dy, int64_t lddy, const float *d_w_ext,
                                    int64_t num_rows, int F_in, int F_out, int has_res, int add_identity, float *d_dx, int64_t lddx,
                                    pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && F_in >= 1 && F_out >= 1 && ldt >= 3 * (int64_t)F_out && lddy >= F_out && lddx >= F_in,
                 "pg_layer_gemm_bwd_dx: bad shape");
    PG_CHECK_ARG(!(has_res && add_identity) && (!add_identity || F_in == F_out), "pg_layer_gemm_bwd_dx: bad residual flags");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_t && d_dy && d_w_ext && d_dx, "pg_layer_gemm_bwd_dx: null buffer");
    const int blocks = has_res ? 4 : 3;
    const int k_total = blocks * F_out;
    const bool a_vec = F_out % 4 == 0 && al16(d_t) && ldt % 4 == 0 && al16(d_dy) && lddy % 4 == 0;
    const bool w_vec = F_out % 4 == 0 && al16(d_w_ext);
    OpA_Cat2 opa{d_t, ldt, 3 * F_out, d_dy, lddy, k_total, num_rows, a_vec};
    OpB_WextBlocksT opb{d_w_ext, F_in, F_out, blocks, w_vec};
    Epi_Dx epi{d_dx, lddx, d_dy, lddy, num_rows, F_in, add_identity};
    return launch_gemm(opa, opb, epi, num_rows, F_in, k_total, k_total, 1, pg_cu(stream));
}

// ---- plain Linear (decoder MLP, row f1): forward with fused bias + ReLU, both gradients, column sums (bias gradient) ----
extern "C" int pg_linear_fwd(const float *d_x, int64_t ldx, int64_t num_rows, int K, const float *d_w, const float *d_bias, int C, int relu,
                             float *d_out, int64_t ldo, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && K >= 1 && C >= 1 && ldx >= K && ldo >= C, "pg_linear_fwd: bad shape");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_x && d_w && d_out, "pg_linear_fwd: null buffer");
    OpA_Plain opa{d_x, ldx, num_rows, K, K % 4 == 0 && al16(d_x) && ldx % 4 == 0};
    OpB_Trans opb{d_w, K, C, K, K % 4 == 0 && al16(d_w)};
    Epi_BiasAct epi{d_bias, d_out, ldo, num_rows, C, relu};
    return launch_gemm(opa, opb, epi, num_rows, C, K, K, 1, pg_cu(stream));
}

extern "C" int pg_linear_bwd_data(const float *d_g, int64_t ldg, int64_t num_rows, int C, const float *d_w, int K, float *d_dx,
                                  int64_t lddx, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && K >= 1 && C >= 1 && ldg >= C && lddx >= K, "pg_linear_bwd_data: bad shape");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_g && d_w && d_dx, "pg_linear_bwd_data: null buffer");
    OpA_Plain opa{d_g, ldg, num_rows, C, C % 4 == 0 && al16(d_g) && ldg % 4 == 0};
    OpB_Rows opb{d_w, K, C, K, K % 4 == 0 && al16(d_w)};
    Epi_BiasAct epi{nullptr, d_dx, lddx, num_rows, K, 0};
    return launch_gemm(opa, opb, epi, num_rows, K, C, C, 1, pg_cu(stream));
}

extern "C" size_t pg_linear_bwd_weight_ws_bytes(int64_t num_rows, int C, int K) {
    return (size_t)split_count(num_rows, C, K) * C * K * sizeof(float) + 256;
}

extern "C" int pg_linear_bwd_weight(const float *d_g, int64_t ldg, const float *d_x, int64_t ldx, int64_t num_rows, int C, int K,
                                    float *d_dw, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && K >= 1 && C >= 1 && ldg >= C && ldx >= K, "pg_linear_bwd_weight: bad shape");
    PG_CHECK_ARG(d_dw, "pg_linear_bwd_weight: null output");
    cudaStream_t st = pg_cu(stream);
    const int64_t numel = (int64_t)C * K;
    if (num_rows == 0) {
        PG_CUDA_CALL(cudaMemsetAsync(d_dw, 0, (size_t)numel * sizeof(float), st));
        return PG_OK;
    }
    PG_CHECK_ARG(d_g && d_x && d_ws, "pg_linear_bwd_weight: null buffer");
    const int splits = split_count(num_rows, C, K);
    if (ws_bytes < (size_t)splits * numel * sizeof(float)) {
        pg_set_error("pg_linear_bwd_weight: workspace too small (%zu < %zu)", ws_bytes, (size_t)splits * numel * sizeof(float));
        return PG_EWORKSPACE;
    }
    int64_t rows_per_split = pg_ceil_div(num_rows, splits);
    rows_per_split = pg_ceil_div(rows_per_split, BK) * BK;
    OpA_PlainT opa{d_g, ldg, num_rows, C, C % 4 == 0 && al16(d_g) && ldg % 4 == 0};
    OpB_Rows opb{d_x, ldx, num_rows, K, K % 4 == 0 && al16(d_x) && ldx % 4 == 0};
    Epi_Partial epi{(float *)d_ws, C, K};
    int rc = launch_gemm(opa, opb, epi, C, K, num_rows, rows_per_split, splits, st);
    if (rc != PG_OK) return rc;
    reduce_splits_kernel<<<grid_for(numel), 256, 0, st>>>((const float *)d_ws, splits, numel, d_dw);
    PG_CUDA_LAUNCH_CHECK("reduce_splits_kernel");
    return PG_OK;
}

extern "C" size_t pg_colsum_ws_bytes(int64_t num_rows, int C) {
    int64_t slices = pg_ceil_div(num_rows, 512);
    if (slices > 256) slices = 256;
    if (slices < 1) slices = 1;
    return (size_t)slices * C * sizeof(float) + 256;
}

extern "C" int pg_colsum(const float *d_g, int64_t ldg, int64_t num_rows, int C, float *d_out, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(num_rows >= 0 && C >= 1 && ldg >= C, "pg_colsum: bad shape");
    PG_CHECK_ARG(d_out && d_ws && (num_rows == 0 || d_g), "pg_colsum: null buffer");
    int64_t slices = pg_ceil_div(num_rows, 512);
    if (slices > 256) slices = 256;
    if (slices < 1) slices = 1;
    if (ws_bytes < (size_t)slices * C * sizeof(float)) {
        pg_set_error("pg_colsum: workspace too small (%zu bytes)", ws_bytes);
        return PG_EWORKSPACE;
    }
    const int64_t rows_per_slice = pg_ceil_div(num_rows > 0 ? num_rows : 1, slices);
    cudaStream_t st = pg_cu(stream);
    const dim3 grid((unsigned)pg_ceil_div(C, 32), (unsigned)slices, 1);
    colsum_kernel<<<grid, 256, 0, st>>>(d_g, ldg, num_rows, C, rows_per_slice, (float *)d_ws);
    PG_CUDA_LAUNCH_CHECK("colsum_kernel");
    reduce_splits_kernel<<<grid_for(C), 256, 0, st>>>((const float *)d_ws, (int)slices, C, d_out);
    PG_CUDA_LAUNCH_CHECK("reduce_splits_kernel");
    return PG_OK;
}
