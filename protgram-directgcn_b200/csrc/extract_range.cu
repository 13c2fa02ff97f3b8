// Hot path A, key-range form of the table -> graph step for tables merged by an NCCL REDUCE-SCATTER
// over key ranges (north star; SURVEY.md 8(e) row "Builder"): after the reduce-scatter rank r holds
// the summed bins of the source n-gram codes [code_lo, code_lo + codes), i.e. the (n+1)-gram keys
// [code_lo * sigma, (code_lo + codes) * sigma) -- whole rows of A_out_w.  Reference steps replaced:
// data_builder.py:151-177 (distinct n-grams, sorted ids) and :267-286 (the edge table).
//
//   pg_graph_extract_range_mark   local keys -> presence of their source / target n-grams in the GLOBAL
//                                 presence table (uint8[sigma^n]; the host MAX-reduces it over the ranks),
//                                 edge offsets of the local keys, number of local edges
//   pg_node_ids_from_presence     reduced presence -> node id of every code (rank among the present
//                                 codes = rank of the n-gram in sorted order), number of nodes
//   pg_node_codes_emit            id -> code list (ascending)
//   pg_graph_extract_range_fill   local keys -> (src id, dst id, count), sorted by (src, dst); the ranks'
//                                 lists concatenated in rank order are the whole coalesced edge table
#include "common.cuh"

namespace {

inline unsigned grid_for(int64_t n, int threads = 256, int per_sm = 8) {
    int64_t want = pg_ceil_div(n, threads);
    int64_t cap = (int64_t)PG_NUM_SMS * per_sm;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

// sigma^n and sigma^(n+1) with overflow guard (tables beyond 2^40 bins are out of range anyway)
bool powers(int n, int sigma, int64_t *pow_n, int64_t *pow_m) {
    if (n < 1 || sigma < 1 || sigma > 255) return false;
    int64_t p = 1;
    for (int i = 0; i < n; ++i) {
        p *= sigma;
        if (p > (1ll << 40)) return false;
    }
    *pow_n = p;
    *pow_m = p * sigma;
    return *pow_m <= (1ll << 40);
}

__global__ void __launch_bounds__(256) range_mark_kernel(const unsigned long long *__restrict__ bins, int64_t key_lo, int64_t nkeys,
                                                         uint32_t sigma, int64_t sigma_pow_n, uint8_t *__restrict__ present,
                                                         int64_t *__restrict__ edge_flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nkeys; i += (int64_t)gridDim.x * blockDim.x) {
        const bool nz = bins[i] != 0ull;
        edge_flag[i] = nz ? 1 : 0;
        if (nz) {
            const int64_t k = key_lo + i;
            present[k / sigma] = 1;        // source n-gram = leading n symbols (always inside the own code range)
            present[k % sigma_pow_n] = 1;  // target n-gram = trailing n symbols (anywhere)
        }
    }
}

__global__ void __launch_bounds__(256) widen_presence_kernel(const uint8_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i] ? 1 : 0;
}

__global__ void __launch_bounds__(256) node_codes_kernel(const uint8_t *__restrict__ present, const int64_t *__restrict__ node_id,
                                                         int64_t ngrams, int64_t *__restrict__ node_code) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngrams; g += (int64_t)gridDim.x * blockDim.x)
        if (present[g]) node_code[node_id[g]] = g;
}

__global__ void __launch_bounds__(256) range_edges_kernel(const unsigned long long *__restrict__ bins, int64_t key_lo, int64_t nkeys,
                                                          const int64_t *__restrict__ edge_off, const int64_t *__restrict__ node_id,
                                                          uint32_t sigma, int64_t sigma_pow_n, int64_t *__restrict__ src,
                                                          int64_t *__restrict__ dst, int64_t *__restrict__ count) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nkeys; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = bins[i];
        if (c) {
            const int64_t k = key_lo + i;
            const int64_t e = edge_off[i];
            src[e] = node_id[k / sigma];
            dst[e] = node_id[k % sigma_pow_n];
            count[e] = (int64_t)c;
        }
    }
}

// keys of the range that lie inside the table (a padded last range may reach past sigma^(n+1): all zero there)
inline int64_t keys_in_table(int64_t code_lo, int64_t codes, int64_t pow_n, int sigma) {
    if (code_lo >= pow_n) return 0;
    const int64_t c = (code_lo + codes <= pow_n) ? codes : pow_n - code_lo;
    return c * sigma;
}

struct RangeWs {
    int64_t *edge_off;
    void *scan_ws;
    size_t scan_bytes;
};
bool carve_range(void *d_ws, size_t ws_bytes, int64_t nkeys, RangeWs *w) {
    PgArena a(d_ws, ws_bytes);
    w->edge_off = a.take<int64_t>((size_t)(nkeys > 0 ? nkeys : 1));
    w->scan_bytes = pg_scan_ws_bytes(nkeys > 0 ? nkeys : 1);
    w->scan_ws = a.take<char>(w->scan_bytes);
    return a.ok;
}

}  // namespace

extern "C" size_t pg_graph_extract_range_ws_bytes(int sigma, int64_t codes) {
    const int64_t nkeys = codes > 0 && sigma > 0 ? codes * sigma : 1;
    return pg_align_up((size_t)nkeys * 8, 256) + pg_align_up(pg_scan_ws_bytes(nkeys), 256) + 1024;
}

extern "C" int pg_graph_extract_range_mark(const unsigned long long *d_bins_local, int n, int sigma, int64_t code_lo, int64_t codes,
                                           uint8_t *d_present, int64_t *d_sizes, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    int64_t pow_n, pow_m;
    PG_CHECK_ARG(d_present && d_sizes && d_ws && code_lo >= 0 && codes >= 0, "pg_graph_extract_range_mark: bad argument");
    if (!powers(n, sigma, &pow_n, &pow_m)) {
        pg_set_error("pg_graph_extract_range_mark: table out of range (n=%d sigma=%d)", n, sigma);
        return PG_ERANGE;
    }
    cudaStream_t st = pg_cu(stream);
    const int64_t nkeys = keys_in_table(code_lo, codes, pow_n, sigma);
    if (nkeys == 0) {
        PG_CUDA_CALL(cudaMemsetAsync(d_sizes, 0, sizeof(int64_t), st));
        return PG_OK;
    }
    PG_CHECK_ARG(d_bins_local, "pg_graph_extract_range_mark: null table");
    RangeWs w;
    if (!carve_range(d_ws, ws_bytes, nkeys, &w)) {
        pg_set_error("pg_graph_extract_range_mark: workspace too small (%zu < %zu)", ws_bytes,
                     pg_graph_extract_range_ws_bytes(sigma, codes));
        return PG_EWORKSPACE;
    }
    range_mark_kernel<<<grid_for(nkeys), 256, 0, st>>>(d_bins_local, code_lo * sigma, nkeys, (uint32_t)sigma, pow_n, d_present, w.edge_off);
    PG_CUDA_LAUNCH_CHECK("range_mark_kernel");
    return pg_exclusive_scan_i64(w.edge_off, w.edge_off, nkeys, d_sizes, w.scan_ws, w.scan_bytes, st);
}

extern "C" size_t pg_node_ids_ws_bytes(int64_t ngrams) { return pg_align_up(pg_scan_ws_bytes(ngrams > 0 ? ngrams : 1), 256) + 512; }

extern "C" int pg_node_ids_from_presence(const uint8_t *d_present, int64_t ngrams, int64_t *d_node_id, int64_t *d_sizes, void *d_ws,
                                         size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(d_present && d_node_id && d_sizes && d_ws && ngrams > 0, "pg_node_ids_from_presence: bad argument");
    PG_CHECK_ARG(ws_bytes >= pg_scan_ws_bytes(ngrams), "pg_node_ids_from_presence: workspace too small");
    cudaStream_t st = pg_cu(stream);
    widen_presence_kernel<<<grid_for(ngrams), 256, 0, st>>>(d_present, ngrams, d_node_id);
    PG_CUDA_LAUNCH_CHECK("widen_presence_kernel");
    return pg_exclusive_scan_i64(d_node_id, d_node_id, ngrams, d_sizes, d_ws, ws_bytes, st);
}

extern "C" int pg_node_codes_emit(const uint8_t *d_present, const int64_t *d_node_id, int64_t ngrams, int64_t *d_node_code,
                                  pg_stream_t stream) {
    PG_CHECK_ARG(d_present && d_node_id && d_node_code && ngrams > 0, "pg_node_codes_emit: bad argument");
    node_codes_kernel<<<grid_for(ngrams), 256, 0, pg_cu(stream)>>>(d_present, d_node_id, ngrams, d_node_code);
    PG_CUDA_LAUNCH_CHECK("node_codes_kernel");
    return PG_OK;
}

extern "C" int pg_graph_extract_range_fill(const unsigned long long *d_bins_local, int n, int sigma, int64_t code_lo, int64_t codes,
                                           const int64_t *d_node_id, int64_t num_edges_local, int64_t *d_src, int64_t *d_dst,
                                           int64_t *d_count, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    int64_t pow_n, pow_m;
    PG_CHECK_ARG(d_ws && code_lo >= 0 && codes >= 0 && num_edges_local >= 0, "pg_graph_extract_range_fill: bad argument");
    if (!powers(n, sigma, &pow_n, &pow_m)) {
        pg_set_error("pg_graph_extract_range_fill: table out of range (n=%d sigma=%d)", n, sigma);
        return PG_ERANGE;
    }
    const int64_t nkeys = keys_in_table(code_lo, codes, pow_n, sigma);
    if (nkeys == 0 || num_edges_local == 0) return PG_OK;
    PG_CHECK_ARG(d_bins_local && d_node_id && d_src && d_dst && d_count, "pg_graph_extract_range_fill: null buffer");
    RangeWs w;
    if (!carve_range(d_ws, ws_bytes, nkeys, &w)) {
        pg_set_error("pg_graph_extract_range_fill: workspace too small");
        return PG_EWORKSPACE;
    }
    range_edges_kernel<<<grid_for(nkeys), 256, 0, pg_cu(stream)>>>(d_bins_local, code_lo * sigma, nkeys, w.edge_off, d_node_id,
                                                                    (uint32_t)sigma, pow_n, d_src, d_dst, d_count);
    PG_CUDA_LAUNCH_CHECK("range_edges_kernel");
    return PG_OK;
}
