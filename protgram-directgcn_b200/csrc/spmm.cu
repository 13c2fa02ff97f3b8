// Hot path B: the propagate() of DirectGCNLayer (reference src/models/protgram_directgcn.py:
// 101-112,137-140) as CSR SpMM.  The reference runs 6 gather -> scale -> scatter_add passes
// that each materialise an [E, F] temporary; here one pass gathers each neighbour row ONCE and
// applies up to three edge values (A_in | A_out | U share one sparsity pattern on
// reference-built graphs).
//
//   fan-out  Z[i, v*F:(v+1)*F] = sum_k val_v[k] * X[col[k], :]            (forward)
//   fan-in   Y[i, :]           = init[i,:] + sum_v sum_k val_v[k] * G[col[k], v*F:(v+1)*F]   (backward)
//
// Mapping: a group of LPR lanes owns one row; every lane holds CHUNKS float4 of the feature
// row (128-bit loads).  The group stages LPR (col, val) entries with one coalesced load each
// and broadcasts them with width-limited shuffles, so index traffic is read once per row.
// Accumulation order is the CSR order within the row -> bitwise reproducible run to run.
// HBM-bound: algorithmic bytes per nnz = 4 (col) + 4*nv (vals) + 4*F (row gather when X is
// not L2 resident); see DESIGN.md.
#include "common.cuh"

namespace {

__device__ __forceinline__ void fma4(float4 &acc, float s, const float4 &x) {
    acc.x = fmaf(s, x.x, acc.x);
    acc.y = fmaf(s, x.y, acc.y);
    acc.z = fmaf(s, x.z, acc.z);
    acc.w = fmaf(s, x.w, acc.w);
}

// Work mapping.  Rows mode: group g owns row g; rows longer than `skip_above` nnz are left to the
// long-row pass.  Items mode (long_rows != nullptr): group g owns one `chunk`-nnz slice of a long
// row and writes a partial sum to its own output row g (combined afterwards in item order, so the
// result stays bitwise reproducible however skewed the degrees are).
struct RowMap {
    int64_t num_groups;
    int64_t skip_above;           // rows mode: 0 = no limit
    const int32_t *long_rows;     // items mode
    const int64_t *item_ptr;
    const int32_t *item_row;
    int32_t chunk;
};

__device__ __forceinline__ bool map_group(const RowMap &m, const int64_t *__restrict__ rowptr, int64_t g, int64_t *rb,
                                          int64_t *re, int64_t *out_row) {
    if (m.long_rows != nullptr) {
        const int32_t li = m.item_row[g];
        const int64_t row = m.long_rows[li];
        const int64_t b = rowptr[row] + (g - m.item_ptr[li]) * (int64_t)m.chunk;
        *rb = b;
        *re = min(b + m.chunk, rowptr[row + 1]);
        *out_row = g;
        return true;
    }
    *rb = rowptr[g];
    *re = rowptr[g + 1];
    *out_row = g;
    return !(m.skip_above > 0 && *re - *rb > m.skip_above);
}

template <int NV, int LPR, int CHUNKS>
__global__ void __launch_bounds__(256) spmm_fanout_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                          const float *__restrict__ val0, const float *__restrict__ val1,
                                                          const float *__restrict__ val2, RowMap map, int F,
                                                          const float *__restrict__ x, int64_t ldx, float *__restrict__ z,
                                                          int64_t ldz, int64_t z_off, const float *__restrict__ s0,
                                                          const float *__restrict__ s1, const float *__restrict__ s2, int sstride) {
    // s0..s2 (optional): per-SOURCE-row scales, z_v[i] = sum_j val_v[i,j] * s_v[j] * x[j]  (backward of the gated layer:
    // x = dY, s_v = gate_v, so that the 3F-wide gated gradient never has to be gathered)
    constexpr int UNROLL = (CHUNKS == 1) ? 4 : 2;
    const int lane = threadIdx.x & 31;
    const int lg = lane & (LPR - 1);                       // lane inside the row group
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (lane & ~(LPR - 1)));
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    if (gid >= map.num_groups) return;
    int64_t rb, re, row;
    if (!map_group(map, rowptr, gid, &rb, &re, &row)) return;
    const int nvec = F >> 2;                               // float4 per feature row
    float4 acc[NV][CHUNKS];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) acc[v][c] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int64_t base = rb; base < re; base += LPR) {
        const int cnt = (int)min((int64_t)LPR, re - base);
        int my_col = 0;
        float my_v[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) my_v[v] = 0.f;
        if (lg < cnt) {
            my_col = col[base + lg];
            my_v[0] = val0[base + lg];
            if (NV > 1) my_v[1] = val1[base + lg];
            if (NV > 2) my_v[NV - 1] = val2[base + lg];
            if (s0 != nullptr) {
                const int64_t so = (int64_t)my_col * sstride;
                my_v[0] *= __ldg(s0 + so);
                if (NV > 1) my_v[1] *= __ldg(s1 + so);
                if (NV > 2) my_v[NV - 1] *= __ldg(s2 + so);
            }
        }
        int t = 0;
        for (; t + UNROLL <= cnt; t += UNROLL) {
            int c_[UNROLL];
            float s_[UNROLL][NV];
            float4 xr[UNROLL][CHUNKS];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                c_[u] = __shfl_sync(gmask, my_col, t + u, LPR);
#pragma unroll
                for (int v = 0; v < NV; ++v) s_[u][v] = __shfl_sync(gmask, my_v[v], t + u, LPR);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float4 *xp = reinterpret_cast<const float4 *>(x + (int64_t)c_[u] * ldx);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    const int f4 = lg + c * LPR;
                    xr[u][c] = (f4 < nvec) ? __ldg(xp + f4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) fma4(acc[v][c], s_[u][v], xr[u][c]);
        }
        for (; t < cnt; ++t) {
            const int cc = __shfl_sync(gmask, my_col, t, LPR);
            float s[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) s[v] = __shfl_sync(gmask, my_v[v], t, LPR);
            const float4 *xp = reinterpret_cast<const float4 *>(x + (int64_t)cc * ldx);
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const int f4 = lg + c * LPR;
                if (f4 < nvec) {
                    const float4 xv = __ldg(xp + f4);
#pragma unroll
                    for (int v = 0; v < NV; ++v) fma4(acc[v][c], s[v], xv);
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        float4 *zp = reinterpret_cast<float4 *>(z + row * ldz + z_off + (int64_t)v * F);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int f4 = lg + c * LPR;
            if (f4 < nvec) zp[f4] = acc[v][c];
        }
    }
}

template <int NV, int LPR, int CHUNKS>
__global__ void __launch_bounds__(256) spmm_fanin_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                         const float *__restrict__ val0, const float *__restrict__ val1,
                                                         const float *__restrict__ val2, RowMap map, int F,
                                                         const float *__restrict__ g, int64_t ldg, int64_t g_off,
                                                         const float *__restrict__ init, int64_t ldinit, float *__restrict__ y,
                                                         int64_t ldy, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int lg = lane & (LPR - 1);
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (lane & ~(LPR - 1)));
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    if (gid >= map.num_groups) return;
    int64_t rb, re, row;
    if (!map_group(map, rowptr, gid, &rb, &re, &row)) return;
    const int nvec = F >> 2;
    float4 acc[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int f4 = lg + c * LPR;
        acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f4 < nvec) {
            if (init) acc[c] = __ldg(reinterpret_cast<const float4 *>(init + row * ldinit) + f4);
            if (accumulate) {
                const float4 o = reinterpret_cast<const float4 *>(y + row * ldy)[f4];
                acc[c].x += o.x; acc[c].y += o.y; acc[c].z += o.z; acc[c].w += o.w;
            }
        }
    }
    for (int64_t base = rb; base < re; base += LPR) {
        const int cnt = (int)min((int64_t)LPR, re - base);
        int my_col = 0;
        float my_v[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) my_v[v] = 0.f;
        if (lg < cnt) {
            my_col = col[base + lg];
            my_v[0] = val0[base + lg];
            if (NV > 1) my_v[1] = val1[base + lg];
            if (NV > 2) my_v[NV - 1] = val2[base + lg];
        }
        int t = 0;
        for (; t + 2 <= cnt; t += 2) {
            int c_[2];
            float s_[2][NV];
            float4 gr[2][NV][CHUNKS];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                c_[u] = __shfl_sync(gmask, my_col, t + u, LPR);
#pragma unroll
                for (int v = 0; v < NV; ++v) s_[u][v] = __shfl_sync(gmask, my_v[v], t + u, LPR);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float4 *gp = reinterpret_cast<const float4 *>(g + (int64_t)c_[u] * ldg + g_off + (int64_t)v * F);
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) {
                        const int f4 = lg + c * LPR;
                        gr[u][v][c] = (f4 < nvec) ? __ldg(gp + f4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) fma4(acc[c], s_[u][v], gr[u][v][c]);
        }
        for (; t < cnt; ++t) {
            const int cc = __shfl_sync(gmask, my_col, t, LPR);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float s = __shfl_sync(gmask, my_v[v], t, LPR);
                const float4 *gp = reinterpret_cast<const float4 *>(g + (int64_t)cc * ldg + g_off + (int64_t)v * F);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    const int f4 = lg + c * LPR;
                    if (f4 < nvec) fma4(acc[c], s, __ldg(gp + f4));
                }
            }
        }
    }
    float4 *yp = reinterpret_cast<float4 *>(y + row * ldy);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int f4 = lg + c * LPR;
        if (f4 < nvec) yp[f4] = acc[c];
    }
}

// Generic scalar kernels: any F / any alignment (one warp per row, lanes stride the features).
template <int NV>
__global__ void __launch_bounds__(256) spmm_fanout_scalar_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                                 const float *__restrict__ val0, const float *__restrict__ val1,
                                                                 const float *__restrict__ val2, int64_t num_rows, int F,
                                                                 const float *__restrict__ x, int64_t ldx, float *__restrict__ z,
                                                                 int64_t ldz, int64_t z_off, const float *__restrict__ s0,
                                                                 const float *__restrict__ s1, const float *__restrict__ s2, int sstride) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= num_rows) return;
    const int64_t rb = rowptr[row], re = rowptr[row + 1];
    for (int f = lane; f < F; f += 32) {
        float acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0.f;
        for (int64_t k = rb; k < re; ++k) {
            const int64_t c = col[k];
            const float xv = x[c * ldx + f];
            acc[0] = fmaf(s0 ? val0[k] * s0[c * sstride] : val0[k], xv, acc[0]);
            if (NV > 1) acc[1] = fmaf(s0 ? val1[k] * s1[c * sstride] : val1[k], xv, acc[1]);
            if (NV > 2) acc[NV - 1] = fmaf(s0 ? val2[k] * s2[c * sstride] : val2[k], xv, acc[NV - 1]);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) z[row * ldz + z_off + (int64_t)v * F + f] = acc[v];
    }
}

template <int NV>
__global__ void __launch_bounds__(256) spmm_fanin_scalar_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                                const float *__restrict__ val0, const float *__restrict__ val1,
                                                                const float *__restrict__ val2, int64_t num_rows, int F,
                                                                const float *__restrict__ g, int64_t ldg, int64_t g_off,
                                                                const float *__restrict__ init, int64_t ldinit, float *__restrict__ y,
                                                                int64_t ldy, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= num_rows) return;
    const int64_t rb = rowptr[row], re = rowptr[row + 1];
    for (int f = lane; f < F; f += 32) {
        float acc = init ? init[row * ldinit + f] : 0.f;
        if (accumulate) acc += y[row * ldy + f];
        for (int64_t k = rb; k < re; ++k) {
            const float *gp = g + (int64_t)col[k] * ldg + g_off + f;
            acc = fmaf(val0[k], gp[0], acc);
            if (NV > 1) acc = fmaf(val1[k], gp[F], acc);
            if (NV > 2) acc = fmaf(val2[k], gp[2 * (int64_t)F], acc);
        }
        y[row * ldy + f] = acc;
    }
}

// Combine the partial sums of the long rows in item order (fixed order => reproducible).
__global__ void __launch_bounds__(256) spmm_reduce_long_kernel(const int32_t *__restrict__ long_rows, const int64_t *__restrict__ item_ptr,
                                                               int64_t n_long, int width, const float *__restrict__ partials,
                                                               float *__restrict__ out, int64_t ldo, int64_t o_off,
                                                               const float *__restrict__ init, int64_t ldinit, int accumulate) {
    const int64_t li = blockIdx.x;
    if (li >= n_long) return;
    const int64_t row = long_rows[li];
    const int64_t ib = item_ptr[li], ie = item_ptr[li + 1];
    for (int f = threadIdx.x; f < width; f += blockDim.x) {
        float acc = init ? init[row * ldinit + f] : 0.f;
        if (accumulate) acc += out[row * ldo + o_off + f];
        for (int64_t it = ib; it < ie; ++it) acc += partials[it * width + f];
        out[row * ldo + o_off + f] = acc;
    }
}

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

// pick the row-group width: smallest power of two >= F/4, capped at 32 lanes x 4 chunks (F <= 512)
inline bool pick_shape(int F, int *lpr, int *chunks) {
    if (F % 4 != 0 || F > 512 || F < 4) return false;
    const int nvec = F / 4;
    int l = 4;
    while (l < nvec && l < 32) l <<= 1;
    int c = (nvec + l - 1) / l;
    if (c == 3) c = 4;
    *lpr = l;
    *chunks = c;
    return true;
}

inline RowMap rows_map(int64_t num_rows, const pg_spmm_plan *plan) {
    RowMap m;
    m.num_groups = num_rows;
    m.skip_above = (plan && plan->n_long > 0) ? plan->chunk : 0;
    m.long_rows = nullptr; m.item_ptr = nullptr; m.item_row = nullptr; m.chunk = 0;
    return m;
}
inline RowMap items_map(const pg_spmm_plan *plan) {
    RowMap m;
    m.num_groups = plan->n_items;
    m.skip_above = 0;
    m.long_rows = plan->d_long_rows; m.item_ptr = plan->d_item_ptr; m.item_row = plan->d_item_row; m.chunk = plan->chunk;
    return m;
}
inline bool plan_ok(const pg_spmm_plan *plan) {
    return !plan || plan->n_long == 0 ||
           (plan->chunk > 0 && plan->n_items >= plan->n_long && plan->d_long_rows && plan->d_item_ptr && plan->d_item_row && plan->d_partials);
}
}  // namespace

#define PG_SPMM_DISPATCH(KERNEL, NV, GROUPS, ...)                                                     \
    do {                                                                                              \
        const unsigned grid = (unsigned)pg_ceil_div((GROUPS) * lpr, 256);                             \
        if (lpr == 4) KERNEL<NV, 4, 1><<<grid, 256, 0, st>>>(__VA_ARGS__);                            \
        else if (lpr == 8) KERNEL<NV, 8, 1><<<grid, 256, 0, st>>>(__VA_ARGS__);                       \
        else if (lpr == 16) KERNEL<NV, 16, 1><<<grid, 256, 0, st>>>(__VA_ARGS__);                     \
        else if (chunks == 1) KERNEL<NV, 32, 1><<<grid, 256, 0, st>>>(__VA_ARGS__);                   \
        else if (chunks == 2) KERNEL<NV, 32, 2><<<grid, 256, 0, st>>>(__VA_ARGS__);                   \
        else KERNEL<NV, 32, 4><<<grid, 256, 0, st>>>(__VA_ARGS__);                                    \
    } while (0)

extern "C" int pg_spmm_fanout(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0, const float *d_val1,
                              const float *d_val2, int nv, int64_t num_rows, int F, const float *d_x, int64_t ldx, float *d_z,
                              int64_t ldz, int64_t z_off, const pg_spmm_plan *plan, pg_stream_t stream) {
    return pg_spmm_fanout_scaled(d_rowptr, d_col, d_val0, d_val1, d_val2, nv, num_rows, F, d_x, ldx, d_z, ldz, z_off, nullptr, nullptr,
                                 nullptr, 0, plan, stream);
}

extern "C" int pg_spmm_fanout_scaled(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0, const float *d_val1,
                                     const float *d_val2, int nv, int64_t num_rows, int F, const float *d_x, int64_t ldx, float *d_z,
                                     int64_t ldz, int64_t z_off, const float *d_s0, const float *d_s1, const float *d_s2,
                                     int scale_stride, const pg_spmm_plan *plan, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(!d_s0 || ((nv == 1 || (d_s1 && d_s2)) && (scale_stride == 0 || scale_stride == 1)), "pg_spmm_fanout_scaled: bad scales");
    const float *s1 = nv == 3 ? d_s1 : d_s0, *s2 = nv == 3 ? d_s2 : d_s0;
    PG_CHECK_ARG(nv == 1 || nv == 3, "pg_spmm_fanout: nv must be 1 or 3 (got %d)", nv);
    PG_CHECK_ARG(num_rows >= 0 && F >= 1 && ldx >= F && ldz >= z_off + (int64_t)nv * F && z_off >= 0, "pg_spmm_fanout: bad shape");
    PG_CHECK_ARG(plan_ok(plan), "pg_spmm_fanout: malformed plan");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_rowptr && d_col && d_val0 && d_x && d_z && (nv == 1 || (d_val1 && d_val2)), "pg_spmm_fanout: null buffer");
    const float *v1 = nv == 3 ? d_val1 : d_val0, *v2 = nv == 3 ? d_val2 : d_val0;
    int lpr = 0, chunks = 0;
    const bool vec = pick_shape(F, &lpr, &chunks) && aligned16(d_x) && aligned16(d_z) && ldx % 4 == 0 && ldz % 4 == 0 && z_off % 4 == 0;
    if (vec) {
        const RowMap rm = rows_map(num_rows, plan);
        if (nv == 3) PG_SPMM_DISPATCH(spmm_fanout_kernel, 3, num_rows, d_rowptr, d_col, d_val0, v1, v2, rm, F, d_x, ldx, d_z, ldz, z_off, d_s0, s1, s2, scale_stride);
        else PG_SPMM_DISPATCH(spmm_fanout_kernel, 1, num_rows, d_rowptr, d_col, d_val0, v1, v2, rm, F, d_x, ldx, d_z, ldz, z_off, d_s0, s1, s2, scale_stride);
        PG_CUDA_LAUNCH_CHECK("spmm_fanout_kernel");
        if (plan && plan->n_long > 0) {
            const RowMap im = items_map(plan);
            const int64_t w = (int64_t)nv * F;
            if (nv == 3) PG_SPMM_DISPATCH(spmm_fanout_kernel, 3, plan->n_items, d_rowptr, d_col, d_val0, v1, v2, im, F, d_x, ldx, plan->d_partials, w, 0, d_s0, s1, s2, scale_stride);
            else PG_SPMM_DISPATCH(spmm_fanout_kernel, 1, plan->n_items, d_rowptr, d_col, d_val0, v1, v2, im, F, d_x, ldx, plan->d_partials, w, 0, d_s0, s1, s2, scale_stride);
            PG_CUDA_LAUNCH_CHECK("spmm_fanout_kernel(long rows)");
            spmm_reduce_long_kernel<<<(unsigned)plan->n_long, 256, 0, st>>>(plan->d_long_rows, plan->d_item_ptr, plan->n_long, (int)w,
                                                                            plan->d_partials, d_z, ldz, z_off, nullptr, 0, 0);
            PG_CUDA_LAUNCH_CHECK("spmm_reduce_long_kernel");
        }
    } else {
        const unsigned grid = (unsigned)pg_ceil_div(num_rows * 32, 256);
        if (nv == 3) spmm_fanout_scalar_kernel<3><<<grid, 256, 0, st>>>(d_rowptr, d_col, d_val0, v1, v2, num_rows, F, d_x, ldx, d_z, ldz, z_off, d_s0, s1, s2, scale_stride);
        else spmm_fanout_scalar_kernel<1><<<grid, 256, 0, st>>>(d_rowptr, d_col, d_val0, v1, v2, num_rows, F, d_x, ldx, d_z, ldz, z_off, d_s0, s1, s2, scale_stride);
        PG_CUDA_LAUNCH_CHECK("spmm_fanout_scalar_kernel");
    }
    return PG_OK;
}

extern "C" int pg_spmm_fanin(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0, const float *d_val1,
                             const float *d_val2, int nv, int64_t num_rows, int F, const float *d_g, int64_t ldg, int64_t g_off,
                             const float *d_init, int64_t ldinit, float *d_y, int64_t ldy, int accumulate,
                             const pg_spmm_plan *plan, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(nv == 1 || nv == 3, "pg_spmm_fanin: nv must be 1 or 3 (got %d)", nv);
    PG_CHECK_ARG(num_rows >= 0 && F >= 1 && ldg >= g_off + (int64_t)nv * F && g_off >= 0 && ldy >= F, "pg_spmm_fanin: bad shape");
    PG_CHECK_ARG(!d_init || ldinit >= F, "pg_spmm_fanin: bad init stride");
    PG_CHECK_ARG(plan_ok(plan), "pg_spmm_fanin: malformed plan");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_rowptr && d_col && d_val0 && d_g && d_y && (nv == 1 || (d_val1 && d_val2)), "pg_spmm_fanin: null buffer");
    const float *v1 = nv == 3 ? d_val1 : d_val0, *v2 = nv == 3 ? d_val2 : d_val0;
    int lpr = 0, chunks = 0;
    const bool vec = pick_shape(F, &lpr, &chunks) && aligned16(d_g) && aligned16(d_y) && ldg % 4 == 0 && ldy % 4 == 0 &&
                     g_off % 4 == 0 && (!d_init || (aligned16(d_init) && ldinit % 4 == 0));
    if (vec) {
        const RowMap rm = rows_map(num_rows, plan);
        if (nv == 3) PG_SPMM_DISPATCH(spmm_fanin_kernel, 3, num_rows, d_rowptr, d_col, d_val0, v1, v2, rm, F, d_g, ldg, g_off, d_init, ldinit, d_y, ldy, accumulate);
        else PG_SPMM_DISPATCH(spmm_fanin_kernel, 1, num_rows, d_rowptr, d_col, d_val0, v1, v2, rm, F, d_g, ldg, g_off, d_init, ldinit, d_y, ldy, accumulate);
        PG_CUDA_LAUNCH_CHECK("spmm_fanin_kernel");
        if (plan && plan->n_long > 0) {
            const RowMap im = items_map(plan);
            if (nv == 3) PG_SPMM_DISPATCH(spmm_fanin_kernel, 3, plan->n_items, d_rowptr, d_col, d_val0, v1, v2, im, F, d_g, ldg, g_off, nullptr, 0, plan->d_partials, (int64_t)F, 0);
            else PG_SPMM_DISPATCH(spmm_fanin_kernel, 1, plan->n_items, d_rowptr, d_col, d_val0, v1, v2, im, F, d_g, ldg, g_off, nullptr, 0, plan->d_partials, (int64_t)F, 0);
            PG_CUDA_LAUNCH_CHECK("spmm_fanin_kernel(long rows)");
            spmm_reduce_long_kernel<<<(unsigned)plan->n_long, 256, 0, st>>>(plan->d_long_rows, plan->d_item_ptr, plan->n_long, F, plan->d_partials,
                                                                            d_y, ldy, 0, d_init, ldinit, accumulate);
            PG_CUDA_LAUNCH_CHECK("spmm_reduce_long_kernel");
        }
    } else {
        const unsigned grid = (unsigned)pg_ceil_div(num_rows * 32, 256);
        if (nv == 3) spmm_fanin_scalar_kernel<3><<<grid, 256, 0, st>>>(d_rowptr, d_col, d_val0, v1, v2, num_rows, F, d_g, ldg, g_off, d_init, ldinit, d_y, ldy, accumulate);
        else spmm_fanin_scalar_kernel<1><<<grid, 256, 0, st>>>(d_rowptr, d_col, d_val0, v1, v2, num_rows, F, d_g, ldg, g_off, d_init, ldinit, d_y, ldy, accumulate);
        PG_CUDA_LAUNCH_CHECK("spmm_fanin_scalar_kernel");
    }
    return PG_OK;
}
