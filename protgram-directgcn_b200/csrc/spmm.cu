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

// The gathered matrix.  Rows below `split` come from `lo`, the others from `hi` (rows are then numbered from `split`):
// a row-partitioned graph keeps its own rows where they are and receives the halo rows it references into a second,
// compact buffer (host/partitioned.py: HaloExchange); the single-GPU calls pass split = INT64_MAX.
struct Operand {
    const float *lo;
    int64_t ld_lo;
    const float *hi;
    int64_t ld_hi;
    int64_t split;
};
__device__ __forceinline__ const float *operand_row(const Operand &x, int c) {
    return c < x.split ? x.lo + (int64_t)c * x.ld_lo : x.hi + ((int64_t)c - x.split) * x.ld_hi;
}

// Work mapping of ONE launch.  The first n_items groups each own one `chunk`-nnz slice of a long row and write a partial sum
// to the plan's scratch (combined afterwards in item order, so the result stays bitwise reproducible however skewed the
// degrees are); the following num_rows groups own one row each and skip the long ones.  Slices first: they are the
// longest-running groups, and the short rows that follow fill the SMs' gaps while they drain.
struct RowMap {
    int64_t num_rows;
    int64_t skip_above;           // rows longer than this are handled as slices (0 = no long-row plan)
    int64_t n_items;
    const int32_t *long_rows;
    const int64_t *item_ptr;
    const int32_t *item_row;
    float *partials;
    int32_t chunk;
};

// -> false: nothing to do for this group.  *item: the group is a slice (output row = slice index in the scratch).
__device__ __forceinline__ bool map_group(const RowMap &m, const int64_t *__restrict__ rowptr, int64_t g, int64_t *rb,
                                          int64_t *re, int64_t *out_row, bool *item) {
    if (g < m.n_items) {
        const int32_t li = m.item_row[g];
        const int64_t row = m.long_rows[li];
        const int64_t b = rowptr[row] + (g - m.item_ptr[li]) * (int64_t)m.chunk;
        *rb = b;
        *re = min(b + m.chunk, rowptr[row + 1]);
        *out_row = g;
        *item = true;
        return true;
    }
    g -= m.n_items;
    *item = false;
    if (g >= m.num_rows) return false;
    *rb = rowptr[g];
    *re = rowptr[g + 1];
    *out_row = g;
    return !(m.skip_above > 0 && *re - *rb > m.skip_above);
}

#ifndef PG_SPMM_MIN_BLOCKS
#define PG_SPMM_MIN_BLOCKS 6   // CTAs per SM the fan-out kernel is compiled for when a lane holds ONE float4 (register cap 80; wider rows: 4 / 3).  Measured on the R-MAT leg (fan-out, ms): 4 blocks x 8 gathers 8.1, 5 x 8 7.2, 6 x 4 6.0, 8 x 4 7.5, 4 x 4 7.1
#endif
#ifndef PG_SPMM_HALF_WARP_ROWS
#define PG_SPMM_HALF_WARP_ROWS 0
#endif
#ifndef PG_SPMM_UNROLL1
#define PG_SPMM_UNROLL1 4      // gathers in flight per lane when a lane holds one float4 of the row
#endif
constexpr int SPMM_THREADS = 128;   // small CTAs: a CTA lives as long as its longest row, the other warps' slots idle meanwhile
constexpr int SPMM_WARPS = SPMM_THREADS / 32;

// One staged CSR entry: the address of the gathered row (split operand resolved by the lane that owns the entry) and the
// up-to-three edge values (gate scales folded in).  32 bytes: one LDS.64 + one LDS.128 broadcast per entry and warp.
struct __align__(16) StagedEntry {
    unsigned long long row;
    float pad0, pad1;
    float v0, v1, v2, pad2;
};

// What bounds these kernels (ncu, round 2).  Version 1: latency -- 80 % of the stall samples long-scoreboard at 19 resident
// warps.  Version 2 (register prefetch of the next index chunk, 8 predicated gathers in flight): the stalls went away but
// the kernel then ISSUED 52 warp instructions per stored entry (shuffles of col and 3 values, the split-operand address
// arithmetic in all 32 lanes, two predicate branches per entry) and ran no faster.  This version cuts the issue count:
//   * the lane that loads an entry resolves its row address ONCE and stages (address, values) in shared memory; the
//     consumers fetch them with two broadcast LDS instead of 4-5 shuffles + address arithmetic per lane;
//   * full batches of UNROLL entries run unpredicated, only a row's tail batch is predicated;
//   * the next chunk's entries are fetched into registers before the current chunk's gathers are issued and staged after
//     them (double-buffered stage), so the index latency stays hidden.
// Accumulation stays in CSR order (bitwise reproducible; a partitioned block equals the same rows of the whole matrix).
template <int NV, int LPR, int CHUNKS, bool FULL>   // FULL: F == 4 * LPR * CHUNKS, no lane is idle -> unpredicated gathers
__global__ void __launch_bounds__(SPMM_THREADS, CHUNKS == 1 ? PG_SPMM_MIN_BLOCKS : (CHUNKS == 2 ? 4 : 3)) spmm_fanout_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                                   const float *__restrict__ val0, const float *__restrict__ val1,
                                                                   const float *__restrict__ val2, RowMap map, int F, Operand x,
                                                                   float *__restrict__ z, int64_t ldz, int64_t z_off, int64_t z_vstride,
                                                                   const float *__restrict__ s0, const float *__restrict__ s1,
                                                                   const float *__restrict__ s2, int sstride) {
    // s0..s2 (optional): per-SOURCE-row scales, z_v[i] = sum_j val_v[i,j] * s_v[j] * x[j]  (backward of the gated layer:
    // x = dY, s_v = gate_v, so that the 3F-wide gated gradient never has to be gathered)
    constexpr int UNROLL = (CHUNKS == 1) ? PG_SPMM_UNROLL1 : (CHUNKS == 2 ? 4 : 2);
    __shared__ StagedEntry stage[SPMM_WARPS][2][32];
    const int lane = threadIdx.x & 31;
    const int lg = lane & (LPR - 1);                       // lane inside the row group
    const int gbase = lane & ~(LPR - 1);                   // first lane of the group = first stage slot of the group
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << gbase);
    StagedEntry(*st)[32] = stage[threadIdx.x >> 5];
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    int64_t rb, re, row;
    bool item;
    if (!map_group(map, rowptr, gid, &rb, &re, &row, &item)) return;
    const int nvec = F >> 2;                               // float4 per feature row
    float4 acc[NV][CHUNKS];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) acc[v][c] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load_entry = [&](int64_t base, unsigned long long &e_row, float &e0, float &e1, float &e2) {
        e_row = 0ull; e0 = e1 = e2 = 0.f;
        if (base + lg < re) {
            const int c = __ldg(col + base + lg);
            e0 = __ldg(val0 + base + lg);
            if (NV > 1) e1 = __ldg(val1 + base + lg);
            if (NV > 2) e2 = __ldg(val2 + base + lg);
            if (s0 != nullptr) {
                const int64_t so = (int64_t)c * sstride;
                e0 *= __ldg(s0 + so);
                if (NV > 1) e1 *= __ldg(s1 + so);
                if (NV > 2) e2 *= __ldg(s2 + so);
            }
            e_row = (unsigned long long)operand_row(x, c);
        }
    };
    auto put_entry = [&](int buf, unsigned long long e_row, float e0, float e1, float e2) {
        *reinterpret_cast<uint2 *>(&st[buf][lane].row) = make_uint2((unsigned)e_row, (unsigned)(e_row >> 32));
        *reinterpret_cast<float4 *>(&st[buf][lane].v0) = make_float4(e0, e1, e2, 0.f);
    };
    auto batch = [&](int buf, int t, int live) {          // live >= UNROLL: all entries of the batch exist (no predicates)
        float4 xr[UNROLL][CHUNKS];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (u < live) {
                const uint2 a = *reinterpret_cast<const uint2 *>(&st[buf][gbase + t + u].row);
                const float4 *xp = reinterpret_cast<const float4 *>(((unsigned long long)a.y << 32) | a.x) + lg;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
                    xr[u][c] = (FULL || lg + c * LPR < nvec) ? __ldg(xp + c * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("" ::: "memory");   // all gathers of the batch are issued before the first value fetch + FMA (ptxas otherwise sinks them)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (u < live) {
                const float4 sv = *reinterpret_cast<const float4 *>(&st[buf][gbase + t + u].v0);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    fma4(acc[0][c], sv.x, xr[u][c]);
                    if (NV > 1) fma4(acc[1][c], sv.y, xr[u][c]);
                    if (NV > 2) fma4(acc[NV - 1][c], sv.z, xr[u][c]);
                }
            }
        }
    };

    unsigned long long n_row;
    float n0, n1, n2;
    load_entry(rb, n_row, n0, n1, n2);
    put_entry(0, n_row, n0, n1, n2);
    __syncwarp(gmask);
    int buf = 0;
    for (int64_t base = rb; base < re; base += LPR, buf ^= 1) {
        const int cnt = (int)min((int64_t)LPR, re - base);
        const bool more = base + LPR < re;
        if (more) load_entry(base + LPR, n_row, n0, n1, n2);   // in flight under this chunk's gathers
        int t = 0;
        for (; t + UNROLL <= cnt; t += UNROLL) batch(buf, t, UNROLL);
        if (t < cnt) batch(buf, t, cnt - t);
        if (more) put_entry(buf ^ 1, n_row, n0, n1, n2);
        __syncwarp(gmask);                                  // stage[buf^1] complete; every lane is done reading stage[buf]
    }
    float *zrow = item ? map.partials + row * (int64_t)(NV * F) : z + row * ldz + z_off;
    const int64_t vstride = item ? (int64_t)F : z_vstride;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        float4 *zp = reinterpret_cast<float4 *>(zrow + (int64_t)v * vstride);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int f4 = lg + c * LPR;
            if (f4 < nvec) zp[f4] = acc[v][c];
        }
    }
}

template <int NV, int LPR, int CHUNKS, bool FULL>
__global__ void __launch_bounds__(SPMM_THREADS, NV * CHUNKS <= 3 ? 6 : (NV * CHUNKS <= 6 ? 4 : 2)) spmm_fanin_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                                  const float *__restrict__ val0, const float *__restrict__ val1,
                                                                  const float *__restrict__ val2, RowMap map, int F, Operand g,
                                                                  int64_t g_off, int64_t g_vstride, const float *__restrict__ init,
                                                                  int64_t ldinit, float *__restrict__ y, int64_t ldy, int accumulate) {
    constexpr int UNROLL = (NV * CHUNKS <= 2) ? 4 : 2;
    __shared__ StagedEntry stage[SPMM_WARPS][2][32];
    const int lane = threadIdx.x & 31;
    const int lg = lane & (LPR - 1);
    const int gbase = lane & ~(LPR - 1);
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << gbase);
    StagedEntry(*st)[32] = stage[threadIdx.x >> 5];
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    int64_t rb, re, row;
    bool item;
    if (!map_group(map, rowptr, gid, &rb, &re, &row, &item)) return;
    const int nvec = F >> 2;
    float4 acc[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int f4 = lg + c * LPR;
        acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f4 < nvec && !item) {
            if (init) acc[c] = __ldg(reinterpret_cast<const float4 *>(init + row * ldinit) + f4);
            if (accumulate) {
                const float4 o = reinterpret_cast<const float4 *>(y + row * ldy)[f4];
                acc[c].x += o.x; acc[c].y += o.y; acc[c].z += o.z; acc[c].w += o.w;
            }
        }
    }
    auto load_entry = [&](int64_t base, unsigned long long &e_row, float &e0, float &e1, float &e2) {
        e_row = 0ull; e0 = e1 = e2 = 0.f;
        if (base + lg < re) {
            const int c = __ldg(col + base + lg);
            e0 = __ldg(val0 + base + lg);
            if (NV > 1) e1 = __ldg(val1 + base + lg);
            if (NV > 2) e2 = __ldg(val2 + base + lg);
            e_row = (unsigned long long)(operand_row(g, c) + g_off);
        }
    };
    auto put_entry = [&](int buf, unsigned long long e_row, float e0, float e1, float e2) {
        *reinterpret_cast<uint2 *>(&st[buf][lane].row) = make_uint2((unsigned)e_row, (unsigned)(e_row >> 32));
        *reinterpret_cast<float4 *>(&st[buf][lane].v0) = make_float4(e0, e1, e2, 0.f);
    };
    auto batch = [&](int buf, int t, int live) {
        float4 gr[UNROLL][NV][CHUNKS];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (u < live) {
                const uint2 a = *reinterpret_cast<const uint2 *>(&st[buf][gbase + t + u].row);
                const float *gb = reinterpret_cast<const float *>(((unsigned long long)a.y << 32) | a.x);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float4 *gp = reinterpret_cast<const float4 *>(gb + (int64_t)v * g_vstride) + lg;
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c)
                        gr[u][v][c] = (FULL || lg + c * LPR < nvec) ? __ldg(gp + c * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        asm volatile("" ::: "memory");
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (u < live) {
                const float4 sv = *reinterpret_cast<const float4 *>(&st[buf][gbase + t + u].v0);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    fma4(acc[c], sv.x, gr[u][0][c]);
                    if (NV > 1) fma4(acc[c], sv.y, gr[u][1][c]);
                    if (NV > 2) fma4(acc[c], sv.z, gr[u][NV - 1][c]);
                }
            }
        }
    };
    unsigned long long n_row;
    float n0, n1, n2;
    load_entry(rb, n_row, n0, n1, n2);
    put_entry(0, n_row, n0, n1, n2);
    __syncwarp(gmask);
    int buf = 0;
    for (int64_t base = rb; base < re; base += LPR, buf ^= 1) {
        const int cnt = (int)min((int64_t)LPR, re - base);
        const bool more = base + LPR < re;
        if (more) load_entry(base + LPR, n_row, n0, n1, n2);
        int t = 0;
        for (; t + UNROLL <= cnt; t += UNROLL) batch(buf, t, UNROLL);
        if (t < cnt) batch(buf, t, cnt - t);
        if (more) put_entry(buf ^ 1, n_row, n0, n1, n2);
        __syncwarp(gmask);
    }
    float4 *yp = reinterpret_cast<float4 *>(item ? map.partials + row * (int64_t)F : y + row * ldy);
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
                                                                 const float *__restrict__ val2, int64_t num_rows, int F, Operand x,
                                                                 float *__restrict__ z, int64_t ldz, int64_t z_off, int64_t z_vstride,
                                                                 const float *__restrict__ s0, const float *__restrict__ s1,
                                                                 const float *__restrict__ s2, int sstride) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= num_rows) return;
    const int64_t rb = rowptr[row], re = rowptr[row + 1];
    for (int f = lane; f < F; f += 32) {
        float acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0.f;
        for (int64_t k = rb; k < re; ++k) {
            const int c = col[k];
            const float xv = operand_row(x, c)[f];
            acc[0] = fmaf(s0 ? val0[k] * s0[(int64_t)c * sstride] : val0[k], xv, acc[0]);
            if (NV > 1) acc[1] = fmaf(s0 ? val1[k] * s1[(int64_t)c * sstride] : val1[k], xv, acc[1]);
            if (NV > 2) acc[NV - 1] = fmaf(s0 ? val2[k] * s2[(int64_t)c * sstride] : val2[k], xv, acc[NV - 1]);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) z[row * ldz + z_off + (int64_t)v * z_vstride + f] = acc[v];
    }
}

template <int NV>
__global__ void __launch_bounds__(256) spmm_fanin_scalar_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                                const float *__restrict__ val0, const float *__restrict__ val1,
                                                                const float *__restrict__ val2, int64_t num_rows, int F, Operand g,
                                                                int64_t g_off, int64_t g_vstride, const float *__restrict__ init,
                                                                int64_t ldinit, float *__restrict__ y, int64_t ldy, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= num_rows) return;
    const int64_t rb = rowptr[row], re = rowptr[row + 1];
    for (int f = lane; f < F; f += 32) {
        float acc = init ? init[row * ldinit + f] : 0.f;
        if (accumulate) acc += y[row * ldy + f];
        for (int64_t k = rb; k < re; ++k) {
            const float *gp = operand_row(g, col[k]) + g_off + f;
            acc = fmaf(val0[k], gp[0], acc);
            if (NV > 1) acc = fmaf(val1[k], gp[g_vstride], acc);
            if (NV > 2) acc = fmaf(val2[k], gp[2 * g_vstride], acc);
        }
        y[row * ldy + f] = acc;
    }
}

// Combine the partial sums of the long rows in item order (fixed order => reproducible).  `width` = nv * F floats per
// slice, laid out [v][F]; output segment v goes to out + o_off + v * o_vstride.
__global__ void __launch_bounds__(256) spmm_reduce_long_kernel(const int32_t *__restrict__ long_rows, const int64_t *__restrict__ item_ptr,
                                                               int64_t n_long, int width, int F, const float *__restrict__ partials,
                                                               float *__restrict__ out, int64_t ldo, int64_t o_off, int64_t o_vstride,
                                                               const float *__restrict__ init, int64_t ldinit, int accumulate) {
    const int64_t li = blockIdx.x;
    if (li >= n_long) return;
    const int64_t row = long_rows[li];
    const int64_t ib = item_ptr[li], ie = item_ptr[li + 1];
    for (int f = threadIdx.x; f < width; f += blockDim.x) {
        const int v = f / F, ff = f - v * F;
        float *o = out + row * ldo + o_off + (int64_t)v * o_vstride + ff;
        float acc = init ? init[row * ldinit + f] : 0.f;
        if (accumulate) acc += *o;
        for (int64_t it = ib; it < ie; ++it) acc += partials[it * width + f];
        *o = acc;
    }
}

// Pack step of the halo exchange: dst[i, :] = src[idx[i], :w] (the rows of this rank that a peer's block references).
template <bool VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ src, int64_t ld_src, const int64_t *__restrict__ idx,
                                                          int64_t count, int w, float *__restrict__ dst, int64_t ld_dst) {
    const int per_row = VEC ? (w >> 2) : w;
    const int64_t total = count * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / per_row;
        const int c = (int)(i - r * per_row);
        const int64_t s = __ldg(idx + r);
        if (VEC) reinterpret_cast<float4 *>(dst + r * ld_dst)[c] = __ldg(reinterpret_cast<const float4 *>(src + s * ld_src) + c);
        else dst[r * ld_dst + c] = __ldg(src + s * ld_src + c);
    }
}

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

// pick the row-group width: smallest power of two >= F/4, capped at 32 lanes x 4 chunks (F <= 512)
inline bool pick_shape(int F, int *lpr, int *chunks) {
    if (F % 4 != 0 || F > 512 || F < 4) return false;
    const int nvec = F / 4;
    int l = 4;
    while (l < nvec && l < 32) l <<= 1;
    if (l == 32 && nvec == 32 && PG_SPMM_HALF_WARP_ROWS) l = 16;   // F = 128: two rows per warp, two float4 per lane
    int c = (nvec + l - 1) / l;
    if (c == 3) c = 4;
    *lpr = l;
    *chunks = c;
    return true;
}

inline RowMap work_map(int64_t num_rows, const pg_spmm_plan *plan) {
    RowMap m;
    const bool lng = plan && plan->n_long > 0;
    m.num_rows = num_rows;
    m.skip_above = lng ? plan->chunk : 0;
    m.n_items = lng ? plan->n_items : 0;
    m.long_rows = lng ? plan->d_long_rows : nullptr;
    m.item_ptr = lng ? plan->d_item_ptr : nullptr;
    m.item_row = lng ? plan->d_item_row : nullptr;
    m.partials = lng ? plan->d_partials : nullptr;
    m.chunk = lng ? plan->chunk : 0;
    return m;
}
inline bool plan_ok(const pg_spmm_plan *plan) {
    return !plan || plan->n_long == 0 ||
           (plan->chunk > 0 && plan->n_items >= plan->n_long && plan->d_long_rows && plan->d_item_ptr && plan->d_item_row && plan->d_partials);
}
inline bool operand_ok(const pg_spmm_operand *x, int64_t min_ld) {
    return x && x->lo && x->ld_lo >= min_ld && x->split >= 0 && (x->hi == nullptr || x->ld_hi >= min_ld);
}
inline bool operand_vec(const pg_spmm_operand *x) {
    return aligned16(x->lo) && x->ld_lo % 4 == 0 && (x->hi == nullptr || (aligned16(x->hi) && x->ld_hi % 4 == 0));
}
inline Operand to_operand(const pg_spmm_operand *x) {
    Operand o;
    o.lo = x->lo; o.ld_lo = x->ld_lo; o.hi = x->hi; o.ld_hi = x->ld_hi;
    o.split = x->hi ? x->split : INT64_MAX;
    return o;
}
}  // namespace

#define PG_SPMM_LAUNCH(KERNEL, NV, L, C, ...)                                                         \
    do {                                                                                              \
        if (full) KERNEL<NV, L, C, true><<<grid, SPMM_THREADS, 0, st>>>(__VA_ARGS__);                 \
        else KERNEL<NV, L, C, false><<<grid, SPMM_THREADS, 0, st>>>(__VA_ARGS__);                     \
    } while (0)
#define PG_SPMM_DISPATCH(KERNEL, NV, GROUPS, ...)                                                     \
    do {                                                                                              \
        const unsigned grid = (unsigned)pg_ceil_div((GROUPS) * lpr, SPMM_THREADS);                    \
        const bool full = (F == 4 * lpr * chunks);                                                    \
        if (lpr == 4) PG_SPMM_LAUNCH(KERNEL, NV, 4, 1, __VA_ARGS__);                                  \
        else if (lpr == 8) PG_SPMM_LAUNCH(KERNEL, NV, 8, 1, __VA_ARGS__);                             \
        else if (lpr == 16 && chunks == 2) PG_SPMM_LAUNCH(KERNEL, NV, 16, 2, __VA_ARGS__);            \
        else if (lpr == 16) PG_SPMM_LAUNCH(KERNEL, NV, 16, 1, __VA_ARGS__);                           \
        else if (chunks == 1) PG_SPMM_LAUNCH(KERNEL, NV, 32, 1, __VA_ARGS__);                         \
        else if (chunks == 2) PG_SPMM_LAUNCH(KERNEL, NV, 32, 2, __VA_ARGS__);                         \
        else PG_SPMM_LAUNCH(KERNEL, NV, 32, 4, __VA_ARGS__);                                          \
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
    pg_spmm_operand x;
    x.lo = d_x; x.ld_lo = ldx; x.hi = nullptr; x.ld_hi = 0; x.split = 0;
    PG_CHECK_ARG(ldz >= z_off + (int64_t)nv * F, "pg_spmm_fanout: bad shape");
    return pg_spmm_fanout_split(d_rowptr, d_col, d_val0, d_val1, d_val2, nv, num_rows, F, &x, d_z, ldz, z_off, F, d_s0, d_s1, d_s2,
                                scale_stride, plan, stream);
}

extern "C" int pg_spmm_fanout_split(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0, const float *d_val1,
                                    const float *d_val2, int nv, int64_t num_rows, int F, const pg_spmm_operand *x, float *d_z,
                                    int64_t ldz, int64_t z_off, int64_t z_vstride, const float *d_s0, const float *d_s1,
                                    const float *d_s2, int scale_stride, const pg_spmm_plan *plan, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(!d_s0 || ((nv == 1 || (d_s1 && d_s2)) && scale_stride >= 0), "pg_spmm_fanout_scaled: bad scales");
    const float *s1 = nv == 3 ? d_s1 : d_s0, *s2 = nv == 3 ? d_s2 : d_s0;
    PG_CHECK_ARG(nv == 1 || nv == 3, "pg_spmm_fanout: nv must be 1 or 3 (got %d)", nv);
    PG_CHECK_ARG(num_rows >= 0 && F >= 1 && z_vstride >= F && z_off >= 0 && ldz >= z_off + (int64_t)(nv - 1) * z_vstride + F,
                 "pg_spmm_fanout: bad shape");
    PG_CHECK_ARG(plan_ok(plan), "pg_spmm_fanout: malformed plan");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(operand_ok(x, F), "pg_spmm_fanout: bad operand (null buffer or row stride < F)");
    PG_CHECK_ARG(d_rowptr && d_col && d_val0 && d_z && (nv == 1 || (d_val1 && d_val2)), "pg_spmm_fanout: null buffer");
    const float *v1 = nv == 3 ? d_val1 : d_val0, *v2 = nv == 3 ? d_val2 : d_val0;
    const Operand xo = to_operand(x);
    int lpr = 0, chunks = 0;
    const bool vec = pick_shape(F, &lpr, &chunks) && operand_vec(x) && aligned16(d_z) && ldz % 4 == 0 && z_off % 4 == 0 && z_vstride % 4 == 0;
    if (vec) {
        const RowMap wm = work_map(num_rows, plan);                    // long-row slices and plain rows in ONE launch
        const int64_t groups = wm.n_items + num_rows;
        if (nv == 3) PG_SPMM_DISPATCH(spmm_fanout_kernel, 3, groups, d_rowptr, d_col, d_val0, v1, v2, wm, F, xo, d_z, ldz, z_off, z_vstride, d_s0, s1, s2, scale_stride);
        else PG_SPMM_DISPATCH(spmm_fanout_kernel, 1, groups, d_rowptr, d_col, d_val0, v1, v2, wm, F, xo, d_z, ldz, z_off, z_vstride, d_s0, s1, s2, scale_stride);
        PG_CUDA_LAUNCH_CHECK("spmm_fanout_kernel");
        if (wm.n_items > 0) {
            const int64_t w = (int64_t)nv * F;
            spmm_reduce_long_kernel<<<(unsigned)plan->n_long, 256, 0, st>>>(plan->d_long_rows, plan->d_item_ptr, plan->n_long, (int)w, F,
                                                                            plan->d_partials, d_z, ldz, z_off, z_vstride, nullptr, 0, 0);
            PG_CUDA_LAUNCH_CHECK("spmm_reduce_long_kernel");
        }
    } else {
        const unsigned grid = (unsigned)pg_ceil_div(num_rows * 32, 256);
        if (nv == 3) spmm_fanout_scalar_kernel<3><<<grid, 256, 0, st>>>(d_rowptr, d_col, d_val0, v1, v2, num_rows, F, xo, d_z, ldz, z_off, z_vstride, d_s0, s1, s2, scale_stride);
        else spmm_fanout_scalar_kernel<1><<<grid, 256, 0, st>>>(d_rowptr, d_col, d_val0, v1, v2, num_rows, F, xo, d_z, ldz, z_off, z_vstride, d_s0, s1, s2, scale_stride);
        PG_CUDA_LAUNCH_CHECK("spmm_fanout_scalar_kernel");
    }
    return PG_OK;
}

extern "C" int pg_spmm_fanin(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0, const float *d_val1,
                             const float *d_val2, int nv, int64_t num_rows, int F, const float *d_g, int64_t ldg, int64_t g_off,
                             const float *d_init, int64_t ldinit, float *d_y, int64_t ldy, int accumulate,
                             const pg_spmm_plan *plan, pg_stream_t stream) {
    pg_spmm_operand g;
    g.lo = d_g; g.ld_lo = ldg; g.hi = nullptr; g.ld_hi = 0; g.split = 0;
    return pg_spmm_fanin_split(d_rowptr, d_col, d_val0, d_val1, d_val2, nv, num_rows, F, &g, g_off, F, d_init, ldinit, d_y, ldy,
                               accumulate, plan, stream);
}

extern "C" int pg_spmm_fanin_split(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0, const float *d_val1,
                                   const float *d_val2, int nv, int64_t num_rows, int F, const pg_spmm_operand *g, int64_t g_off,
                                   int64_t g_vstride, const float *d_init, int64_t ldinit, float *d_y, int64_t ldy, int accumulate,
                                   const pg_spmm_plan *plan, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(nv == 1 || nv == 3, "pg_spmm_fanin: nv must be 1 or 3 (got %d)", nv);
    PG_CHECK_ARG(num_rows >= 0 && F >= 1 && g_off >= 0 && g_vstride >= F && ldy >= F, "pg_spmm_fanin: bad shape");
    PG_CHECK_ARG(!d_init || ldinit >= F, "pg_spmm_fanin: bad init stride");
    PG_CHECK_ARG(plan_ok(plan), "pg_spmm_fanin: malformed plan");
    if (num_rows == 0) return PG_OK;
    PG_CHECK_ARG(operand_ok(g, g_off + (int64_t)(nv - 1) * g_vstride + F), "pg_spmm_fanin: bad operand (null buffer or row stride too small)");
    PG_CHECK_ARG(d_rowptr && d_col && d_val0 && d_y && (nv == 1 || (d_val1 && d_val2)), "pg_spmm_fanin: null buffer");
    const float *v1 = nv == 3 ? d_val1 : d_val0, *v2 = nv == 3 ? d_val2 : d_val0;
    const Operand go = to_operand(g);
    int lpr = 0, chunks = 0;
    const bool vec = pick_shape(F, &lpr, &chunks) && operand_vec(g) && aligned16(d_y) && ldy % 4 == 0 && g_off % 4 == 0 &&
                     g_vstride % 4 == 0 && (!d_init || (aligned16(d_init) && ldinit % 4 == 0));
    if (vec) {
        const RowMap wm = work_map(num_rows, plan);
        const int64_t groups = wm.n_items + num_rows;
        if (nv == 3) PG_SPMM_DISPATCH(spmm_fanin_kernel, 3, groups, d_rowptr, d_col, d_val0, v1, v2, wm, F, go, g_off, g_vstride, d_init, ldinit, d_y, ldy, accumulate);
        else PG_SPMM_DISPATCH(spmm_fanin_kernel, 1, groups, d_rowptr, d_col, d_val0, v1, v2, wm, F, go, g_off, g_vstride, d_init, ldinit, d_y, ldy, accumulate);
        PG_CUDA_LAUNCH_CHECK("spmm_fanin_kernel");
        if (wm.n_items > 0) {
            spmm_reduce_long_kernel<<<(unsigned)plan->n_long, 256, 0, st>>>(plan->d_long_rows, plan->d_item_ptr, plan->n_long, F, F, plan->d_partials,
                                                                            d_y, ldy, 0, 0, d_init, ldinit, accumulate);
            PG_CUDA_LAUNCH_CHECK("spmm_reduce_long_kernel");
        }
    } else {
        const unsigned grid = (unsigned)pg_ceil_div(num_rows * 32, 256);
        if (nv == 3) spmm_fanin_scalar_kernel<3><<<grid, 256, 0, st>>>(d_rowptr, d_col, d_val0, v1, v2, num_rows, F, go, g_off, g_vstride, d_init, ldinit, d_y, ldy, accumulate);
        else spmm_fanin_scalar_kernel<1><<<grid, 256, 0, st>>>(d_rowptr, d_col, d_val0, v1, v2, num_rows, F, go, g_off, g_vstride, d_init, ldinit, d_y, ldy, accumulate);
        PG_CUDA_LAUNCH_CHECK("spmm_fanin_scalar_kernel");
    }
    return PG_OK;
}

extern "C" int pg_gather_rows(const float *d_src, int64_t ld_src, const int64_t *d_idx, int64_t count, int w, float *d_dst,
                              int64_t ld_dst, pg_stream_t stream) {
    PG_CHECK_ARG(count >= 0 && w >= 1 && ld_src >= w && ld_dst >= w, "pg_gather_rows: bad shape");
    if (count == 0) return PG_OK;
    PG_CHECK_ARG(d_src && d_idx && d_dst, "pg_gather_rows: null buffer");
    cudaStream_t st = pg_cu(stream);
    const bool vec = w % 4 == 0 && aligned16(d_src) && aligned16(d_dst) && ld_src % 4 == 0 && ld_dst % 4 == 0;
    const int64_t total = count * (vec ? w / 4 : w);
    const int64_t want = pg_ceil_div(total, 256), cap = (int64_t)PG_NUM_SMS * 16;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (vec) gather_rows_kernel<true><<<grid, 256, 0, st>>>(d_src, ld_src, d_idx, count, w, d_dst, ld_dst);
    else gather_rows_kernel<false><<<grid, 256, 0, st>>>(d_src, ld_src, d_idx, count, w, d_dst, ld_dst);
    PG_CUDA_LAUNCH_CHECK("gather_rows_kernel");
    return PG_OK;
}
