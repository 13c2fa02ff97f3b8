// Hot path A, part 2: adjacency + propagation matrices of DirectedNgramGraph
// (reference src/utils/graph_utils.py:140-287), rebuilt as ONE stable radix sort of
// 2E+N tagged keys plus per-item passes, instead of ~14 sparse-COO coalesces.
//
//   tag 0: (i,j) from A_out_w      tag 1: (j,i) the transposed copy      tag 2: (i,i) identity
//
// Sorting key = row*N + col (payload = position in the tagged list, so the tag is implied and
// the sort's stability orders duplicates tag0 < tag1 < tag2).  Unique keys = the shared pattern
// sym(A) U I of all three propagation matrices; the tag-1 items in sorted order ARE A_in_w
// coalesced.  Every fp32 step mirrors the reference op sequence (explicit _rn intrinsics so
// ptxas cannot contract mul+add into an FMA and change the rounding).
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

// ------------------------------------------------------------------ CSR <-> COO glue
__global__ void __launch_bounds__(256) rowptr_kernel(const int64_t *__restrict__ rows, int64_t nnz, int64_t num_rows,
                                                     int64_t *__restrict__ rowptr) {
    // rowptr[r] = first position whose row id >= r.  Thread e fills the gap (rows[e-1], rows[e]].
    PG_GRID_STRIDE(e, nnz + 1) {
        const int64_t lo = (e == 0) ? -1 : rows[e - 1];
        const int64_t hi = (e == nnz) ? num_rows : rows[e];
        for (int64_t r = lo + 1; r <= hi; ++r) rowptr[r] = e;
    }
}

__global__ void __launch_bounds__(256) coo_from_csr_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                           int64_t num_rows, int64_t *__restrict__ row_out,
                                                           int64_t *__restrict__ col_out) {
    // one warp per row: rows of n-gram graphs hold <= 2*sigma+1 entries
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < num_rows; r += warps) {
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        for (int64_t k = b + lane; k < e; k += 32) {
            row_out[k] = r;
            col_out[k] = col[k];
        }
    }
}

// ------------------------------------------------------------------ coalesce
__global__ void __launch_bounds__(256) make_keys_kernel(const int64_t *__restrict__ a, const int64_t *__restrict__ b,
                                                        int64_t nnz, unsigned long long stride,
                                                        unsigned long long *__restrict__ keys, uint32_t *__restrict__ payload) {
    PG_GRID_STRIDE(e, nnz) {
        keys[e] = (unsigned long long)a[e] * stride + (unsigned long long)b[e];
        payload[e] = (uint32_t)e;
    }
}

__global__ void __launch_bounds__(256) head_flags_kernel(const unsigned long long *__restrict__ keys, int64_t n,
                                                         int64_t *__restrict__ flags) {
    PG_GRID_STRIDE(i, n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void __launch_bounds__(256) coalesce_emit_kernel(const unsigned long long *__restrict__ keys,
                                                            const uint32_t *__restrict__ payload, const int64_t *__restrict__ upos,
                                                            const float *__restrict__ w, int64_t n, unsigned long long stride,
                                                            int64_t *__restrict__ src_out, int64_t *__restrict__ dst_out,
                                                            float *__restrict__ w_out) {
    PG_GRID_STRIDE(i, n) {
        if (i == 0 || keys[i] != keys[i - 1]) {
            const unsigned long long k = keys[i];
            float acc = w[payload[i]];
            for (int64_t j = i + 1; j < n && keys[j] == k; ++j) acc = __fadd_rn(acc, w[payload[j]]);  // input order
            const int64_t u = upos[i];
            src_out[u] = (int64_t)(k / stride);
            dst_out[u] = (int64_t)(k % stride);
            w_out[u] = acc;
        }
    }
}

// ------------------------------------------------------------------ normalisation
__global__ void __launch_bounds__(256) degree_sums_kernel(const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                                          const float *__restrict__ w, int64_t nnz, double *__restrict__ rs_out,
                                                          double *__restrict__ rs_in) {
    // weighted out-/in-degree.  fp64 atomics: exact (hence order-independent) for the integer
    // transition counts of reference-built graphs up to 2^53.
    PG_GRID_STRIDE(e, nnz) {
        const double x = (double)w[e];
        atomicAdd(&rs_out[src[e]], x);
        atomicAdd(&rs_in[dst[e]], x);
    }
}

__global__ void __launch_bounds__(256) tagged_keys_kernel(const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                                          int64_t nnz, int64_t num_nodes, unsigned long long *__restrict__ keys,
                                                          uint32_t *__restrict__ payload) {
    const unsigned long long N = (unsigned long long)num_nodes;
    PG_GRID_STRIDE(t, 2 * nnz + num_nodes) {
        unsigned long long k;
        if (t < nnz) k = (unsigned long long)src[t] * N + (unsigned long long)dst[t];
        else if (t < 2 * nnz) k = (unsigned long long)dst[t - nnz] * N + (unsigned long long)src[t - nnz];
        else k = (unsigned long long)(t - 2 * nnz) * (N + 1);
        keys[t] = k;
        payload[t] = (uint32_t)t;
    }
}

__global__ void __launch_bounds__(256) union_flags_kernel(const unsigned long long *__restrict__ keys,
                                                          const uint32_t *__restrict__ payload, int64_t n, int64_t nnz,
                                                          int64_t *__restrict__ head, int64_t *__restrict__ is_t) {
    PG_GRID_STRIDE(i, n) {
        head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
        const int64_t p = payload[i];
        is_t[i] = (p >= nnz && p < 2 * nnz) ? 1 : 0;
    }
}

// pass 1 over the sorted list: pattern structure, A_in_w, native-self-loop flags
__global__ void __launch_bounds__(256) union_structure_kernel(
    const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ payload, const int64_t *__restrict__ upos,
    const int64_t *__restrict__ tpos, const float *__restrict__ w, int64_t n, int64_t nnz, int64_t num_nodes,
    int64_t pattern_nnz, int64_t *__restrict__ in_src, int64_t *__restrict__ in_dst, float *__restrict__ in_w,
    int64_t *__restrict__ rowptr, int32_t *__restrict__ col, uint8_t *__restrict__ native_loop) {
    const unsigned long long N = (unsigned long long)num_nodes;
    PG_GRID_STRIDE(i, n) {
        const unsigned long long k = keys[i];
        const int64_t r = (int64_t)(k / N), c = (int64_t)(k % N);
        const int64_t p = payload[i];
        if (p >= nnz && p < 2 * nnz) {  // transposed copy -> A_in_w entry (row = dst, col = src)
            const int64_t t = tpos[i];
            in_src[t] = r;
            in_dst[t] = c;
            in_w[t] = w[p - nnz];
        }
        const bool head = (i == 0) || keys[i - 1] != k;
        if (head) {
            const int64_t u = upos[i];
            col[u] = (int32_t)c;
            // every row holds its diagonal, so row r starts where the previous item has another row
            if (i == 0 || (int64_t)(keys[i - 1] / N) != r) rowptr[r] = u;
            if (r == c) native_loop[r] = (p < 2 * nnz) ? 1 : 0;  // first of the group is tag 0/1 => native loop
        }
        if (i == n - 1) rowptr[num_nodes] = pattern_nnz;
    }
}

// pass 2: the three value arrays
__global__ void __launch_bounds__(256) union_values_kernel(
    const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ payload, const int64_t *__restrict__ upos,
    const float *__restrict__ w, int64_t n, int64_t nnz, int64_t num_nodes, const double *__restrict__ rs_out,
    const double *__restrict__ rs_in, const int64_t *__restrict__ rowptr, const uint8_t *__restrict__ native_loop, float eps,
    float *__restrict__ val_out, float *__restrict__ val_in, float *__restrict__ val_und) {
    const unsigned long long N = (unsigned long long)num_nodes;
    PG_GRID_STRIDE(i, n) {
        const unsigned long long k = keys[i];
        if (i != 0 && keys[i - 1] == k) continue;
        const int64_t r = (int64_t)(k / N), c = (int64_t)(k % N);
        float w_rc = 0.f, w_cr = 0.f;  // A_out[r,c], A_out[c,r]
        bool has_sym = false;
        for (int64_t j = i; j < n && keys[j] == k; ++j) {
            const int64_t p = payload[j];
            if (p < nnz) { w_rc = w[p]; has_sym = true; }
            else if (p < 2 * nnz) { w_cr = w[p - nnz]; has_sym = true; }
        }
        const int64_t u = upos[i];
        // D^-1 (graph_utils.py:231-241): 1/rowsum, 0 for empty rows
        const float so_r = (float)rs_out[r], so_c = (float)rs_out[c];
        const float si_r = (float)rs_in[r], si_c = (float)rs_in[c];
        const float io_r = so_r != 0.f ? __frcp_rn(so_r) : 0.f, io_c = so_c != 0.f ? __frcp_rn(so_c) : 0.f;
        const float ii_r = si_r != 0.f ? __frcp_rn(si_r) : 0.f, ii_c = si_c != 0.f ? __frcp_rn(si_c) : 0.f;
        float vo, vi;
        if (has_sym) {
            // mathcal_A_out: A = A_out_w.  An[r,c] = w_rc/out(r), An[c,r] = w_cr/out(c)
            float a = __fmul_rn(w_rc, io_r), b = __fmul_rn(w_cr, io_c);
            float s = (r == c) ? __fadd_rn(__fmul_rn(a, a), __fmul_rn(a, a)) : __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
            vo = __fsqrt_rn(__fadd_rn(__fmul_rn(s, 0.5f), eps));
            // mathcal_A_in: A = A_in_w = A_out_w^T.  An[r,c] = w_cr/in(r), An[c,r] = w_rc/in(c)
            a = __fmul_rn(w_cr, ii_r);
            b = __fmul_rn(w_rc, ii_c);
            s = (r == c) ? __fadd_rn(__fmul_rn(a, a), __fmul_rn(a, a)) : __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
            vi = __fsqrt_rn(__fadd_rn(__fmul_rn(s, 0.5f), eps));
            if (r == c) { vo = __fadd_rn(vo, 1.f); vi = __fadd_rn(vi, 1.f); }
        } else {
            vo = vi = 1.f;  // identity only (graph_utils.py:268-269)
        }
        val_out[u] = vo;
        val_in[u] = vi;
        // undirected (graph_utils.py:168-195): deg counts sym-pattern entries of the column plus the
        // appended self loop; a native self loop is stored twice and summed by coalesce().
        const float deg_r = (float)(rowptr[r + 1] - rowptr[r] + native_loop[r]);
        const float deg_c = (float)(rowptr[c + 1] - rowptr[c] + native_loop[c]);
        const float dr = __frcp_rn(__fsqrt_rn(deg_r)), dc = __frcp_rn(__fsqrt_rn(deg_c));
        float vu = __fmul_rn(dr, dc);
        if (r == c && native_loop[r]) vu = __fadd_rn(vu, vu);
        val_und[u] = vu;
    }
}

// ------------------------------------------------------------------ edge list -> CSR
__global__ void __launch_bounds__(256) group_keys_kernel(const int64_t *__restrict__ group, int64_t nnz,
                                                         unsigned long long *__restrict__ keys, uint32_t *__restrict__ payload) {
    PG_GRID_STRIDE(e, nnz) {
        keys[e] = (unsigned long long)group[e];
        payload[e] = (uint32_t)e;
    }
}

__global__ void __launch_bounds__(256) csr_gather_kernel(const unsigned long long *__restrict__ keys,
                                                         const uint32_t *__restrict__ payload, const int64_t *__restrict__ other,
                                                         const float *__restrict__ w, int64_t nnz, int64_t *__restrict__ rows_tmp,
                                                         int32_t *__restrict__ col, float *__restrict__ val) {
    PG_GRID_STRIDE(i, nnz) {
        const uint32_t p = payload[i];
        rows_tmp[i] = (int64_t)keys[i];
        col[i] = (int32_t)other[p];
        val[i] = w ? w[p] : 1.f;
    }
}

struct SortWs {
    unsigned long long *keys, *keys_alt;
    uint32_t *pay, *pay_alt;
    void *sort_ws;
    size_t sort_bytes;
};
bool carve_sort(PgArena &a, int64_t n, SortWs *s) {
    s->keys = a.take<unsigned long long>((size_t)n);
    s->keys_alt = a.take<unsigned long long>((size_t)n);
    s->pay = a.take<uint32_t>((size_t)n);
    s->pay_alt = a.take<uint32_t>((size_t)n);
    s->sort_bytes = pg_sort_pairs_ws_bytes(n);
    s->sort_ws = a.take<char>(s->sort_bytes);
    return a.ok;
}
size_t sort_ws_total(int64_t n) {
    return pg_align_up((size_t)n * 8, 256) * 2 + pg_align_up((size_t)n * 4, 256) * 2 +
           pg_align_up(pg_sort_pairs_ws_bytes(n), 256);
}

struct NormWs {
    SortWs s;
    int64_t *upos, *tpos;
    double *rs_out, *rs_in;
    uint8_t *native_loop;
    void *scan_ws;
    size_t scan_bytes;
};
bool carve_norm(void *d_ws, size_t ws_bytes, int64_t nnz, int64_t num_nodes, NormWs *w) {
    PgArena a(d_ws, ws_bytes);
    const int64_t n = 2 * nnz + num_nodes;
    carve_sort(a, n, &w->s);
    w->upos = a.take<int64_t>((size_t)n);
    w->tpos = a.take<int64_t>((size_t)n);
    w->rs_out = a.take<double>((size_t)num_nodes);
    w->rs_in = a.take<double>((size_t)num_nodes);
    w->native_loop = a.take<uint8_t>((size_t)num_nodes);
    w->scan_bytes = pg_scan_ws_bytes(n);
    w->scan_ws = a.take<char>(w->scan_bytes);
    return a.ok;
}
}  // namespace

extern "C" int pg_rowptr_from_sorted(const int64_t *d_rows, int64_t nnz, int64_t num_rows, int64_t *d_rowptr, pg_stream_t stream) {
    PG_CHECK_ARG(d_rowptr && nnz >= 0 && num_rows >= 0 && (nnz == 0 || d_rows), "pg_rowptr_from_sorted: bad argument");
    rowptr_kernel<<<grid_for(nnz + 1), 256, 0, pg_cu(stream)>>>(d_rows, nnz, num_rows, d_rowptr);
    PG_CUDA_LAUNCH_CHECK("rowptr_kernel");
    return PG_OK;
}

extern "C" int pg_coo_from_csr(const int64_t *d_rowptr, const int32_t *d_col, int64_t num_rows, int64_t nnz, int64_t *d_row_out,
                               int64_t *d_col_out, pg_stream_t stream) {
    PG_CHECK_ARG(d_rowptr && num_rows >= 0 && nnz >= 0, "pg_coo_from_csr: bad argument");
    if (nnz == 0 || num_rows == 0) return PG_OK;
    PG_CHECK_ARG(d_col && d_row_out && d_col_out, "pg_coo_from_csr: null buffer");
    coo_from_csr_kernel<<<grid_for(num_rows * 32), 256, 0, pg_cu(stream)>>>(d_rowptr, d_col, num_rows, d_row_out, d_col_out);
    PG_CUDA_LAUNCH_CHECK("coo_from_csr_kernel");
    return PG_OK;
}

extern "C" size_t pg_coo_coalesce_ws_bytes(int64_t nnz) {
    if (nnz <= 0) return 256;
    return sort_ws_total(nnz) + pg_align_up((size_t)nnz * 8, 256) + pg_align_up(pg_scan_ws_bytes(nnz), 256) + 1024;
}

extern "C" int pg_coo_coalesce(const int64_t *d_src, const int64_t *d_dst, const float *d_w, int64_t nnz, int64_t num_nodes,
                               int64_t *d_src_out, int64_t *d_dst_out, float *d_w_out, int64_t *d_sizes, void *d_ws,
                               size_t ws_bytes, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(nnz >= 0 && num_nodes >= 0 && d_sizes, "pg_coo_coalesce: bad argument");
    PG_CHECK_ARG(num_nodes <= (1ll << 31), "pg_coo_coalesce: num_nodes > 2^31");
    PG_CHECK_ARG(nnz < (1ll << 32), "pg_coo_coalesce: nnz must be < 2^32");
    if (nnz == 0) {
        PG_CUDA_CALL(cudaMemsetAsync(d_sizes, 0, sizeof(int64_t), st));
        return PG_OK;
    }
    PG_CHECK_ARG(d_src && d_dst && d_w && d_src_out && d_dst_out && d_w_out && d_ws, "pg_coo_coalesce: null buffer");
    PgArena a(d_ws, ws_bytes);
    SortWs s;
    carve_sort(a, nnz, &s);
    int64_t *upos = a.take<int64_t>((size_t)nnz);
    size_t scan_bytes = pg_scan_ws_bytes(nnz);
    void *scan_ws = a.take<char>(scan_bytes);
    if (!a.ok) {
        pg_set_error("pg_coo_coalesce: workspace too small (%zu < %zu)", ws_bytes, pg_coo_coalesce_ws_bytes(nnz));
        return PG_EWORKSPACE;
    }
    const unsigned long long N = (unsigned long long)num_nodes;
    make_keys_kernel<<<grid_for(nnz), 256, 0, st>>>(d_src, d_dst, nnz, N, s.keys, s.pay);
    PG_CUDA_LAUNCH_CHECK("make_keys_kernel");
    int rc = pg_sort_pairs(s.keys, s.keys_alt, s.pay, s.pay_alt, nnz, bits_for(N * N - 1), s.sort_ws, s.sort_bytes, stream);
    if (rc != PG_OK) return rc;
    head_flags_kernel<<<grid_for(nnz), 256, 0, st>>>(s.keys, nnz, upos);
    PG_CUDA_LAUNCH_CHECK("head_flags_kernel");
    rc = pg_exclusive_scan_i64(upos, upos, nnz, d_sizes, scan_ws, scan_bytes, st);
    if (rc != PG_OK) return rc;
    coalesce_emit_kernel<<<grid_for(nnz), 256, 0, st>>>(s.keys, s.pay, upos, d_w, nnz, N, d_src_out, d_dst_out, d_w_out);
    PG_CUDA_LAUNCH_CHECK("coalesce_emit_kernel");
    return PG_OK;
}

extern "C" size_t pg_normalize_ws_bytes(int64_t nnz, int64_t num_nodes) {
    const int64_t n = 2 * nnz + num_nodes;
    if (n <= 0) return 256;
    return sort_ws_total(n) + pg_align_up((size_t)n * 8, 256) * 2 + pg_align_up((size_t)num_nodes * 8, 256) * 2 +
           pg_align_up((size_t)num_nodes, 256) + pg_align_up(pg_scan_ws_bytes(n), 256) + 2048;
}

extern "C" int pg_normalize_sizes(const int64_t *d_src, const int64_t *d_dst, const float *d_w, int64_t nnz, int64_t num_nodes,
                                  int64_t *d_sizes, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(nnz >= 0 && num_nodes > 0 && d_sizes && d_ws, "pg_normalize_sizes: bad argument");
    PG_CHECK_ARG(nnz == 0 || (d_src && d_dst && d_w), "pg_normalize_sizes: null edge buffer");
    PG_CHECK_ARG(num_nodes <= (1ll << 31) - 1, "pg_normalize_sizes: num_nodes must fit int32 columns");
    const int64_t n = 2 * nnz + num_nodes;
    PG_CHECK_ARG(n < (1ll << 32), "pg_normalize_sizes: 2*nnz + num_nodes must be < 2^32");
    NormWs w;
    if (!carve_norm(d_ws, ws_bytes, nnz, num_nodes, &w)) {
        pg_set_error("pg_normalize_sizes: workspace too small (%zu < %zu)", ws_bytes, pg_normalize_ws_bytes(nnz, num_nodes));
        return PG_EWORKSPACE;
    }
    PG_CUDA_CALL(cudaMemsetAsync(w.rs_out, 0, (size_t)num_nodes * sizeof(double), st));
    PG_CUDA_CALL(cudaMemsetAsync(w.rs_in, 0, (size_t)num_nodes * sizeof(double), st));
    if (nnz > 0) {
        degree_sums_kernel<<<grid_for(nnz), 256, 0, st>>>(d_src, d_dst, d_w, nnz, w.rs_out, w.rs_in);
        PG_CUDA_LAUNCH_CHECK("degree_sums_kernel");
    }
    tagged_keys_kernel<<<grid_for(n), 256, 0, st>>>(d_src, d_dst, nnz, num_nodes, w.s.keys, w.s.pay);
    PG_CUDA_LAUNCH_CHECK("tagged_keys_kernel");
    const unsigned long long N = (unsigned long long)num_nodes;
    int rc = pg_sort_pairs(w.s.keys, w.s.keys_alt, w.s.pay, w.s.pay_alt, n, bits_for(N * N - 1), w.s.sort_ws, w.s.sort_bytes, stream);
    if (rc != PG_OK) return rc;
    union_flags_kernel<<<grid_for(n), 256, 0, st>>>(w.s.keys, w.s.pay, n, nnz, w.upos, w.tpos);
    PG_CUDA_LAUNCH_CHECK("union_flags_kernel");
    rc = pg_exclusive_scan_i64(w.upos, w.upos, n, d_sizes, w.scan_ws, w.scan_bytes, st);
    if (rc != PG_OK) return rc;
    return pg_exclusive_scan_i64(w.tpos, w.tpos, n, nullptr, w.scan_ws, w.scan_bytes, st);
}

extern "C" int pg_normalize_fill(const int64_t *d_src, const int64_t *d_dst, const float *d_w, int64_t nnz, int64_t num_nodes,
                                 float eps, int64_t pattern_nnz, int64_t *d_in_src, int64_t *d_in_dst, float *d_in_w,
                                 int64_t *d_rowptr, int32_t *d_col, float *d_val_out, float *d_val_in, float *d_val_und,
                                 void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    (void)d_src;
    (void)d_dst;
    PG_CHECK_ARG(nnz >= 0 && num_nodes > 0 && pattern_nnz >= num_nodes && d_ws, "pg_normalize_fill: bad argument");
    PG_CHECK_ARG(d_rowptr && d_col && d_val_out && d_val_in && d_val_und, "pg_normalize_fill: null output");
    PG_CHECK_ARG(nnz == 0 || (d_w && d_in_src && d_in_dst && d_in_w), "pg_normalize_fill: null edge buffer");
    const int64_t n = 2 * nnz + num_nodes;
    NormWs w;
    if (!carve_norm(d_ws, ws_bytes, nnz, num_nodes, &w)) {
        pg_set_error("pg_normalize_fill: workspace too small");
        return PG_EWORKSPACE;
    }
    union_structure_kernel<<<grid_for(n), 256, 0, st>>>(w.s.keys, w.s.pay, w.upos, w.tpos, d_w, n, nnz, num_nodes, pattern_nnz,
                                                         d_in_src, d_in_dst, d_in_w, d_rowptr, d_col, w.native_loop);
    PG_CUDA_LAUNCH_CHECK("union_structure_kernel");
    union_values_kernel<<<grid_for(n), 256, 0, st>>>(w.s.keys, w.s.pay, w.upos, d_w, n, nnz, num_nodes, w.rs_out, w.rs_in,
                                                      d_rowptr, w.native_loop, eps, d_val_out, d_val_in, d_val_und);
    PG_CUDA_LAUNCH_CHECK("union_values_kernel");
    return PG_OK;
}

extern "C" size_t pg_edges_to_csr_ws_bytes(int64_t nnz) {
    if (nnz <= 0) return 256;
    return sort_ws_total(nnz) + pg_align_up((size_t)nnz * 8, 256) + 1024;
}

extern "C" int pg_edges_to_csr(const int64_t *d_group, const int64_t *d_other, const float *d_w, int64_t nnz, int64_t num_nodes,
                               int64_t *d_rowptr, int32_t *d_col, float *d_val, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    cudaStream_t st = pg_cu(stream);
    PG_CHECK_ARG(nnz >= 0 && num_nodes >= 0 && d_rowptr, "pg_edges_to_csr: bad argument");
    PG_CHECK_ARG(num_nodes <= (1ll << 31) - 1 && nnz < (1ll << 32), "pg_edges_to_csr: sizes exceed int32 columns / 2^32 edges");
    if (nnz == 0) {
        PG_CUDA_CALL(cudaMemsetAsync(d_rowptr, 0, (size_t)(num_nodes + 1) * sizeof(int64_t), st));
        return PG_OK;
    }
    PG_CHECK_ARG(d_group && d_other && d_col && d_val && d_ws, "pg_edges_to_csr: null buffer");
    PgArena a(d_ws, ws_bytes);
    SortWs s;
    carve_sort(a, nnz, &s);
    int64_t *rows_tmp = a.take<int64_t>((size_t)nnz);
    if (!a.ok) {
        pg_set_error("pg_edges_to_csr: workspace too small (%zu < %zu)", ws_bytes, pg_edges_to_csr_ws_bytes(nnz));
        return PG_EWORKSPACE;
    }
    group_keys_kernel<<<grid_for(nnz), 256, 0, st>>>(d_group, nnz, s.keys, s.pay);
    PG_CUDA_LAUNCH_CHECK("group_keys_kernel");
    int rc = pg_sort_pairs(s.keys, s.keys_alt, s.pay, s.pay_alt, nnz, bits_for((unsigned long long)(num_nodes > 0 ? num_nodes - 1 : 0)),
                           s.sort_ws, s.sort_bytes, stream);
    if (rc != PG_OK) return rc;
    csr_gather_kernel<<<grid_for(nnz), 256, 0, st>>>(s.keys, s.pay, d_other, d_w, nnz, rows_tmp, d_col, d_val);
    PG_CUDA_LAUNCH_CHECK("csr_gather_kernel");
    return pg_rowptr_from_sorted(rows_tmp, nnz, num_nodes, d_rowptr, stream);
}
