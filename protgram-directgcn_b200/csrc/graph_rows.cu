// Hot path A, part 2, for a graph whose rows are partitioned over ranks (SURVEY.md 8(e), row
// "Normalisation (a7-a9)"): the adjacency + propagation matrices of DirectedNgramGraph
// (reference src/utils/graph_utils.py:140-287) for ONE row block [row_lo, row_lo + rows).
//
// The owner of a row needs, besides its own out-edges (i -> j), the transposed entries (j -> i)
// that other ranks own; the host side (host/partitioned.py) delivers those with one all-to-all and
// all-gathers three per-node vectors (weighted out-/in-degree, undirected degree).  Everything
// else is local and follows graph.cu item by item, so that a block computed here is bitwise equal
// to the same rows of pg_normalize_* on the whole graph:
//
//   tag 0: (i,j) own out-edge    tag 1: (i,j) from a received in-edge (j -> i)    tag 2: (i,i)
//
// one stable radix sort of nnz_out + nnz_in + rows keys (i - row_lo) * N + j, unique keys = the
// pattern rows of the block, tag-1 items in order = the block's rows of A_in_w.
#include "common.cuh"

namespace {

inline unsigned grid_for(int64_t n, int threads = 256, int per_sm = 8) {
    int64_t want = pg_ceil_div(n, threads);
    int64_t cap = (int64_t)PG_NUM_SMS * per_sm;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

inline int bits_for(unsigned long long max_value) {
    int b = 0;
    while (max_value) {
        ++b;
        max_value >>= 1;
    }
    return b;
}

#define PG_GRID_STRIDE(i, n) \
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

// weighted degrees of the block's rows: out-degree from the own out-edges, in-degree from the received
// in-edges.  fp64 atomics as in graph.cu (exact for integer transition counts).
__global__ void __launch_bounds__(256) rows_degree_sums_kernel(const int64_t *__restrict__ o_src, const float *__restrict__ o_w,
                                                               int64_t nnz_o, const int64_t *__restrict__ i_dst,
                                                               const float *__restrict__ i_w, int64_t nnz_i, int64_t row_lo,
                                                               int64_t rows, double *__restrict__ rs_out, double *__restrict__ rs_in) {
    PG_GRID_STRIDE(t, nnz_o + nnz_i) {
        if (t < nnz_o) {
            const int64_t r = o_src[t] - row_lo;
            if (r >= 0 && r < rows) atomicAdd(&rs_out[r], (double)o_w[t]);
        } else {
            const int64_t e = t - nnz_o;
            const int64_t r = i_dst[e] - row_lo;
            if (r >= 0 && r < rows) atomicAdd(&rs_in[r], (double)i_w[e]);
        }
    }
}

__global__ void __launch_bounds__(256) rows_tagged_keys_kernel(const int64_t *__restrict__ o_src, const int64_t *__restrict__ o_dst,
                                                               int64_t nnz_o, const int64_t *__restrict__ i_src,
                                                               const int64_t *__restrict__ i_dst, int64_t nnz_i, int64_t num_nodes,
                                                               int64_t row_lo, int64_t rows, unsigned long long *__restrict__ keys,
                                                               uint32_t *__restrict__ payload, int64_t *__restrict__ bad) {
    const unsigned long long N = (unsigned long long)num_nodes;
    PG_GRID_STRIDE(t, nnz_o + nnz_i + rows) {
        int64_t r, c;
        if (t < nnz_o) {
            r = o_src[t] - row_lo;
            c = o_dst[t];
        } else if (t < nnz_o + nnz_i) {
            r = i_dst[t - nnz_o] - row_lo;
            c = i_src[t - nnz_o];
        } else {
            r = t - nnz_o - nnz_i;
            c = row_lo + r;
        }
        if (r < 0 || r >= rows || c < 0 || c >= num_nodes) {  // not an edge of this block: report, keep the sort in range
            *bad = 1;
            r = 0;
            c = row_lo;
        }
        keys[t] = (unsigned long long)r * N + (unsigned long long)c;
        payload[t] = (uint32_t)t;
    }
}

__global__ void __launch_bounds__(256) rows_union_flags_kernel(const unsigned long long *__restrict__ keys,
                                                               const uint32_t *__restrict__ payload, int64_t n, int64_t nnz_o,
                                                               int64_t nnz_i, int64_t *__restrict__ head, int64_t *__restrict__ is_t) {
    PG_GRID_STRIDE(i, n) {
        head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
        const int64_t p = payload[i];
        is_t[i] = (p >= nnz_o && p < nnz_o + nnz_i) ? 1 : 0;
    }
}

// pass 1 over the sorted list: pattern rows of the block, its rows of A_in_w, native-self-loop flags
__global__ void __launch_bounds__(256) rows_structure_kernel(
    const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ payload, const int64_t *__restrict__ upos,
    const int64_t *__restrict__ tpos, const float *__restrict__ i_w, int64_t n, int64_t nnz_o, int64_t nnz_i, int64_t num_nodes,
    int64_t row_lo, int64_t rows, int64_t pattern_nnz, int64_t *__restrict__ in_row, int64_t *__restrict__ in_col,
    float *__restrict__ in_w, int64_t *__restrict__ rowptr, int32_t *__restrict__ col, uint8_t *__restrict__ native_loop) {
    const unsigned long long N = (unsigned long long)num_nodes;
    PG_GRID_STRIDE(i, n) {
        const unsigned long long k = keys[i];
        const int64_t rl = (int64_t)(k / N), c = (int64_t)(k % N);
        const int64_t p = payload[i];
        if (p >= nnz_o && p < nnz_o + nnz_i) {  // received in-edge -> A_in_w entry (row = its dst, col = its src)
            const int64_t t = tpos[i];
            in_row[t] = row_lo + rl;
            in_col[t] = c;
            in_w[t] = i_w[p - nnz_o];
        }
        const bool head = (i == 0) || keys[i - 1] != k;
        if (head) {
            const int64_t u = upos[i];
            col[u] = (int32_t)c;
            // every row holds its diagonal, so a row starts where the previous item belongs to another row
            if (i == 0 || (int64_t)(keys[i - 1] / N) != rl) rowptr[rl] = u;
            if (row_lo + rl == c) native_loop[rl] = (p < nnz_o + nnz_i) ? 1 : 0;  // first of the group is tag 0/1
        }
        if (i == n - 1) rowptr[rows] = pattern_nnz;
    }
}

// pass 2: the three value arrays (same fp32 op sequence as union_values_kernel in graph.cu)
__global__ void __launch_bounds__(256) rows_values_kernel(
    const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ payload, const int64_t *__restrict__ upos,
    const float *__restrict__ o_w, const float *__restrict__ i_w, int64_t n, int64_t nnz_o, int64_t nnz_i, int64_t num_nodes,
    int64_t row_lo, const double *__restrict__ rs_out, const double *__restrict__ rs_in, const int32_t *__restrict__ deg,
    const uint8_t *__restrict__ native_loop, float eps, float *__restrict__ val_out, float *__restrict__ val_in,
    float *__restrict__ val_und) {
    const unsigned long long N = (unsigned long long)num_nodes;
    PG_GRID_STRIDE(i, n) {
        const unsigned long long k = keys[i];
        if (i != 0 && keys[i - 1] == k) continue;
        const int64_t rl = (int64_t)(k / N), c = (int64_t)(k % N);
        const int64_t r = row_lo + rl;
        float w_rc = 0.f, w_cr = 0.f;  // A_out[r,c], A_out[c,r]
        bool has_sym = false;
        for (int64_t j = i; j < n && keys[j] == k; ++j) {
            const int64_t p = payload[j];
            if (p < nnz_o) { w_rc = o_w[p]; has_sym = true; }
            else if (p < nnz_o + nnz_i) { w_cr = i_w[p - nnz_o]; has_sym = true; }
        }
        const int64_t u = upos[i];
        const float so_r = (float)rs_out[r], so_c = (float)rs_out[c];
        const float si_r = (float)rs_in[r], si_c = (float)rs_in[c];
        const float io_r = so_r != 0.f ? __frcp_rn(so_r) : 0.f, io_c = so_c != 0.f ? __frcp_rn(so_c) : 0.f;
        const float ii_r = si_r != 0.f ? __frcp_rn(si_r) : 0.f, ii_c = si_c != 0.f ? __frcp_rn(si_c) : 0.f;
        float vo, vi;
        if (has_sym) {
            float a = __fmul_rn(w_rc, io_r), b = __fmul_rn(w_cr, io_c);
            float s = (r == c) ? __fadd_rn(__fmul_rn(a, a), __fmul_rn(a, a)) : __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
            vo = __fsqrt_rn(__fadd_rn(__fmul_rn(s, 0.5f), eps));
            a = __fmul_rn(w_cr, ii_r);
            b = __fmul_rn(w_rc, ii_c);
            s = (r == c) ? __fadd_rn(__fmul_rn(a, a), __fmul_rn(a, a)) : __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
            vi = __fsqrt_rn(__fadd_rn(__fmul_rn(s, 0.5f), eps));
            if (r == c) { vo = __fadd_rn(vo, 1.f); vi = __fadd_rn(vi, 1.f); }
        } else {
            vo = vi = 1.f;
        }
        val_out[u] = vo;
        val_in[u] = vi;
        const float dr = __frcp_rn(__fsqrt_rn((float)deg[r])), dc = __frcp_rn(__fsqrt_rn((float)deg[c]));
        float vu = __fmul_rn(dr, dc);
        if (r == c && native_loop[rl]) vu = __fadd_rn(vu, vu);
        val_und[u] = vu;
    }
}

struct RowsWs {
    unsigned long long *keys, *keys_alt;
    uint32_t *pay, *pay_alt;
    void *sort_ws;
    size_t sort_bytes;
    int64_t *upos, *tpos;
    void *scan_ws;
    size_t scan_bytes;
};

bool carve_rows(void *d_ws, size_t ws_bytes, int64_t n, RowsWs *w) {
    PgArena a(d_ws, ws_bytes);
    w->keys = a.take<unsigned long long>((size_t)n);
    w->keys_alt = a.take<unsigned long long>((size_t)n);
    w->pay = a.take<uint32_t>((size_t)n);
    w->pay_alt = a.take<uint32_t>((size_t)n);
    w->sort_bytes = pg_sort_pairs_ws_bytes(n);
    w->sort_ws = a.take<char>(w->sort_bytes);
    w->upos = a.take<int64_t>((size_t)n);
    w->tpos = a.take<int64_t>((size_t)n);
    w->scan_bytes = pg_scan_ws_bytes(n);
    w->scan_ws = a.take<char>(w->scan_bytes);
    return a.ok;
}

}  // namespace

extern "C" int pg_degree_sums_rows(const int64_t *d_out_src, const float *d_out_w, int64_t nnz_out, const int64_t *d_in_dst,
                                   const float *d_in_w, int64_t nnz_in, int64_t row_lo, int64_t rows, double *d_rs_out,
                                   double *d_rs_in, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(nnz_out >= 0 && nnz_in >= 0 && row_lo >= 0 && rows >= 0, "pg_degree_sums_rows: bad argument");
    if (rows == 0) return PG_OK;
    PG_CHECK_ARG(d_rs_out && d_rs_in, "pg_degree_sums_rows: null output");
    PG_CHECK_ARG((nnz_out == 0 || (d_out_src && d_out_w)) && (nnz_in == 0 || (d_in_dst && d_in_w)),
                 "pg_degree_sums_rows: null edge buffer");
    PG_CUDA_CALL(cudaMemsetAsync(d_rs_out, 0, (size_t)rows * sizeof(double), st));
    PG_CUDA_CALL(cudaMemsetAsync(d_rs_in, 0, (size_t)rows * sizeof(double), st));
    if (nnz_out + nnz_in > 0) {
        rows_degree_sums_kernel<<<grid_for(nnz_out + nnz_in), 256, 0, st>>>(d_out_src, d_out_w, nnz_out, d_in_dst, d_in_w, nnz_in,
                                                                            row_lo, rows, d_rs_out, d_rs_in);
        PG_CUDA_LAUNCH_CHECK("rows_degree_sums_kernel");
    }
    return PG_OK;
}

extern "C" size_t pg_normalize_rows_ws_bytes(int64_t nnz_out, int64_t nnz_in, int64_t rows) {
    const int64_t n = nnz_out + nnz_in + rows;
    if (n <= 0) return 256;
    return pg_align_up((size_t)n * 8, 256) * 4 + pg_align_up((size_t)n * 4, 256) * 2 + pg_align_up(pg_sort_pairs_ws_bytes(n), 256) +
           pg_align_up(pg_scan_ws_bytes(n), 256) + 2048;
}

extern "C" int pg_normalize_rows_sizes(const int64_t *d_out_src, const int64_t *d_out_dst, int64_t nnz_out, const int64_t *d_in_src,
                                       const int64_t *d_in_dst, int64_t nnz_in, int64_t num_nodes, int64_t row_lo, int64_t rows,
                                       int64_t *d_sizes, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(nnz_out >= 0 && nnz_in >= 0 && num_nodes > 0 && rows > 0 && row_lo >= 0 && row_lo + rows <= num_nodes && d_sizes && d_ws,
                 "pg_normalize_rows_sizes: bad argument");
    PG_CHECK_ARG((nnz_out == 0 || (d_out_src && d_out_dst)) && (nnz_in == 0 || (d_in_src && d_in_dst)),
                 "pg_normalize_rows_sizes: null edge buffer");
    PG_CHECK_ARG(num_nodes <= (1ll << 31) - 1, "pg_normalize_rows_sizes: num_nodes must fit int32 columns");
    const int64_t n = nnz_out + nnz_in + rows;
    PG_CHECK_ARG(n < (1ll << 32), "pg_normalize_rows_sizes: nnz_out + nnz_in + rows must be < 2^32");
    RowsWs w;
    if (!carve_rows(d_ws, ws_bytes, n, &w)) {
        pg_set_error("pg_normalize_rows_sizes: workspace too small (%zu < %zu)", ws_bytes,
                     pg_normalize_rows_ws_bytes(nnz_out, nnz_in, rows));
        return PG_EWORKSPACE;
    }
    PG_CUDA_CALL(cudaMemsetAsync(d_sizes, 0, 2 * sizeof(int64_t), st));
    rows_tagged_keys_kernel<<<grid_for(n), 256, 0, st>>>(d_out_src, d_out_dst, nnz_out, d_in_src, d_in_dst, nnz_in, num_nodes, row_lo,
                                                         rows, w.keys, w.pay, d_sizes + 1);
    PG_CUDA_LAUNCH_CHECK("rows_tagged_keys_kernel");
    const unsigned long long top = (unsigned long long)rows * (unsigned long long)num_nodes - 1ull;
    int rc = pg_sort_pairs(w.keys, w.keys_alt, w.pay, w.pay_alt, n, bits_for(top), w.sort_ws, w.sort_bytes, stream);
    if (rc != PG_OK) return rc;
    rows_union_flags_kernel<<<grid_for(n), 256, 0, st>>>(w.keys, w.pay, n, nnz_out, nnz_in, w.upos, w.tpos);
    PG_CUDA_LAUNCH_CHECK("rows_union_flags_kernel");
    rc = pg_exclusive_scan_i64(w.upos, w.upos, n, d_sizes, w.scan_ws, w.scan_bytes, st);
    if (rc != PG_OK) return rc;
    return pg_exclusive_scan_i64(w.tpos, w.tpos, n, nullptr, w.scan_ws, w.scan_bytes, st);
}

extern "C" int pg_normalize_rows_structure(const float *d_in_w, int64_t nnz_out, int64_t nnz_in, int64_t num_nodes, int64_t row_lo,
                                           int64_t rows, int64_t pattern_nnz, int64_t *d_ain_row, int64_t *d_ain_col, float *d_ain_w,
                                           int64_t *d_rowptr, int32_t *d_col, uint8_t *d_native_loop, void *d_ws, size_t ws_bytes,
                                           pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(nnz_out >= 0 && nnz_in >= 0 && num_nodes > 0 && rows > 0 && row_lo >= 0 && pattern_nnz >= rows && d_ws,
                 "pg_normalize_rows_structure: bad argument");
    PG_CHECK_ARG(d_rowptr && d_col && d_native_loop, "pg_normalize_rows_structure: null output");
    PG_CHECK_ARG(nnz_in == 0 || (d_in_w && d_ain_row && d_ain_col && d_ain_w), "pg_normalize_rows_structure: null edge buffer");
    const int64_t n = nnz_out + nnz_in + rows;
    RowsWs w;
    if (!carve_rows(d_ws, ws_bytes, n, &w)) {
        pg_set_error("pg_normalize_rows_structure: workspace too small");
        return PG_EWORKSPACE;
    }
    rows_structure_kernel<<<grid_for(n), 256, 0, st>>>(w.keys, w.pay, w.upos, w.tpos, d_in_w, n, nnz_out, nnz_in, num_nodes, row_lo,
                                                       rows, pattern_nnz, d_ain_row, d_ain_col, d_ain_w, d_rowptr, d_col,
                                                       d_native_loop);
    PG_CUDA_LAUNCH_CHECK("rows_structure_kernel");
    return PG_OK;
}

extern "C" int pg_normalize_rows_values(const float *d_out_w, const float *d_in_w, int64_t nnz_out, int64_t nnz_in, int64_t num_nodes,
                                        int64_t row_lo, int64_t rows, const double *d_rs_out, const double *d_rs_in,
                                        const int32_t *d_deg, const uint8_t *d_native_loop, float eps, float *d_val_out,
                                        float *d_val_in, float *d_val_und, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(nnz_out >= 0 && nnz_in >= 0 && num_nodes > 0 && rows > 0 && row_lo >= 0 && d_ws,
                 "pg_normalize_rows_values: bad argument");
    PG_CHECK_ARG(d_rs_out && d_rs_in && d_deg && d_native_loop && d_val_out && d_val_in && d_val_und,
                 "pg_normalize_rows_values: null buffer");
    PG_CHECK_ARG((nnz_out == 0 || d_out_w) && (nnz_in == 0 || d_in_w), "pg_normalize_rows_values: null edge buffer");
    const int64_t n = nnz_out + nnz_in + rows;
    RowsWs w;
    if (!carve_rows(d_ws, ws_bytes, n, &w)) {
        pg_set_error("pg_normalize_rows_values: workspace too small");
        return PG_EWORKSPACE;
    }
    rows_values_kernel<<<grid_for(n), 256, 0, st>>>(w.keys, w.pay, w.upos, d_out_w, d_in_w, n, nnz_out, nnz_in, num_nodes, row_lo,
                                                    d_rs_out, d_rs_in, d_deg, d_native_loop, eps, d_val_out, d_val_in, d_val_und);
    PG_CUDA_LAUNCH_CHECK("rows_values_kernel");
    return PG_OK;
}
