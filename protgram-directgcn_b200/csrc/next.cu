// Rows f2 and f4 of SURVEY.md section 8 (the callers on either side of the two hot paths):
//   f4  next_node labels      reference src/pipeline/protgram_directgcn_trainer.py:222-237  (O(N*E) Python masks)
//   f2a feature hand-off      reference protgram_directgcn_trainer.py:312-330  (Python loop over all nodes, dict lookups)
//   f2b protein pooling       reference src/utils/models_utils.py:210-262      (Python loop over all residues)
//   f4b cluster mini-batches  reference protgram_directgcn_trainer.py:179-197  (3 x torch_geometric.utils.subgraph per cluster)
// The first three are gather kernels over the packed base-sigma n-gram codes hot path A already computes.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ f4: arg-max successor per row
__global__ void __launch_bounds__(256) next_node_labels_kernel(const int64_t *__restrict__ rowptr, const int64_t *__restrict__ dst,
                                                               const float *__restrict__ w, int64_t n, int64_t *__restrict__ labels) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lo = rowptr[i], hi = rowptr[i + 1];
        int64_t best = i;  // no successor: the node itself (reference :229-230)
        float bw = -INFINITY;
        for (int64_t e = lo; e < hi; ++e) {
            const float v = w[e];
            if (v > bw) {  // strict: the first maximal successor in (src, dst) order
                bw = v;
                best = dst[e];
            }
        }
        labels[i] = best;
    }
}

// ------------------------------------------------------------------ f2a: features of level n from level n-1
__device__ __forceinline__ int64_t find_code(const int64_t *__restrict__ sorted, int64_t count, int64_t key) {
    int64_t lo = 0, hi = count;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return (lo < count && sorted[lo] == key) ? lo : -1;
}

// one warp per node; prefix = code / sigma, suffix = code % sigma^(n-1) are level-(n-1) codes
__global__ void __launch_bounds__(256) feature_init_kernel(const int64_t *__restrict__ code, int64_t num, const int64_t *__restrict__ prev_code,
                                                           int64_t num_prev, int64_t sigma, int64_t sigma_pow_nm1,
                                                           const float *__restrict__ prev_emb, int64_t ld, int F, float *__restrict__ x,
                                                           int64_t ldx) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < num; i += warps) {
        const int64_t c = code[i];
        const int64_t p = find_code(prev_code, num_prev, c / sigma);
        const int64_t s = find_code(prev_code, num_prev, c % sigma_pow_nm1);
        for (int f = lane; f < F; f += 32) {
            float v = 0.f;
            if (p >= 0 && s >= 0) v = __fdiv_rn(__fadd_rn(prev_emb[p * ld + f], prev_emb[s * ld + f]), 2.f);  // np.mean of two fp32 rows
            else if (p >= 0) v = prev_emb[p * ld + f];
            else if (s >= 0) v = prev_emb[s * ld + f];
            x[i * ldx + f] = v;
        }
    }
}

// ------------------------------------------------------------------ f2b: protein-level pooling
// One CTA per protein.  Phase 1 marks the DISTINCT known n-grams of the protein in a shared-memory bitmap indexed by
// the packed code (the reference's fancy-index `+=` applies once per distinct n-gram).  Phase 2 walks the bitmap in
// ascending code order (= ascending node id) chunk by chunk: an exclusive scan of the word popcounts turns a chunk into
// an ordered id list, then thread f adds the rows' feature f one after the other -- the reference's fp32 summation order.
constexpr int kPoolThreads = 128;
constexpr int kPoolChunkWords = 256;                       // 8192 bits per chunk
constexpr int kPoolListCap = kPoolChunkWords * 32;

__global__ void __launch_bounds__(kPoolThreads) pool_proteins_kernel(const uint8_t *__restrict__ seqs, const int64_t *__restrict__ offsets,
                                                                     int64_t num_proteins, int n, const uint8_t *__restrict__ rank_of_byte,
                                                                     uint32_t sigma, uint32_t sigma_pow_n, const int32_t *__restrict__ code_to_id,
                                                                     const float *__restrict__ emb, int64_t ld, int F, float *__restrict__ out,
                                                                     int64_t ldout, uint8_t *__restrict__ valid) {
    extern __shared__ unsigned smem_u[];
    const uint32_t words = (sigma_pow_n + 31u) / 32u;
    unsigned *bitmap = smem_u;                  // [words]
    int *list = (int *)(smem_u + words);         // [kPoolListCap]
    __shared__ uint8_t lut[256];
    __shared__ int warp_tot[kPoolThreads / 32];
    __shared__ int chunk_count;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256; i += kPoolThreads) lut[i] = rank_of_byte[i];
    for (int64_t p = blockIdx.x; p < num_proteins; p += gridDim.x) {
        __syncthreads();
        for (uint32_t i = tid; i < words; i += kPoolThreads) bitmap[i] = 0u;
        __syncthreads();
        const int64_t b = offsets[p], len = offsets[p + 1] - b;
        for (int64_t i = tid; i + n <= len; i += kPoolThreads) {
            uint32_t code = 0;
            bool ok = true;
            for (int k = 0; k < n; ++k) {
                const uint32_t r = lut[seqs[b + i + k]];
                ok = ok && r != 255u;
                code = code * sigma + r;
            }
            if (ok && code_to_id[code] >= 0) atomicOr(&bitmap[code >> 5], 1u << (code & 31u));
        }
        __syncthreads();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};  // features tid, tid+128, tid+256, tid+384 (F <= 512)
        int total = 0;
        for (uint32_t w0 = 0; w0 < words; w0 += kPoolChunkWords) {
            // contiguous words per thread so that list order == code order
            constexpr int WPT = kPoolChunkWords / kPoolThreads;
            unsigned mine[WPT];
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < WPT; ++j) {
                const uint32_t wi = w0 + tid * WPT + j;
                mine[j] = wi < words ? bitmap[wi] : 0u;
                cnt += __popc(mine[j]);
            }
            int incl = cnt;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= s) incl += o;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            int base = 0;
            for (int q = 0; q < warp; ++q) base += warp_tot[q];
            if (tid == kPoolThreads - 1) chunk_count = base + incl;
            int pos = base + incl - cnt;
#pragma unroll
            for (int j = 0; j < WPT; ++j) {
                unsigned m = mine[j];
                const uint32_t code0 = (w0 + tid * WPT + j) * 32u;
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    list[pos++] = code_to_id[code0 + bit];
                }
            }
            __syncthreads();
            const int m_count = chunk_count;
            total += m_count;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int f = tid + q * kPoolThreads;
                if (f < F) {
                    float a = acc[q];
                    for (int e = 0; e < m_count; ++e) a = __fadd_rn(a, emb[(int64_t)list[e] * ld + f]);
                    acc[q] = a;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int f = tid + q * kPoolThreads;
            if (f < F) out[p * ldout + f] = total > 0 ? __fdiv_rn(acc[q], (float)total) : 0.f;
        }
        if (tid == 0) valid[p] = total > 0 ? 1 : 0;
    }
}

// ------------------------------------------------------------------ f4b: cluster mini-batch = induced subgraph of the shared-pattern CSR
__global__ void __launch_bounds__(256) subgraph_mark_kernel(const int64_t *__restrict__ subset, int64_t n_sub, int64_t num_nodes,
                                                            int32_t *__restrict__ new_id, int *__restrict__ bad) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sub; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = subset[s];
        if (r < 0 || r >= num_nodes) *bad = 1;
        else new_id[r] = (int32_t)s;   // duplicates in `subset`: the last position wins, like node_idx[subset] = arange(...)
    }
}

// FILL = false: kept[s] = stored entries of row subset[s] whose column is in the subset;  FILL = true: write them
template <bool FILL>
__global__ void __launch_bounds__(256) subgraph_rows_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                            const float *__restrict__ va, const float *__restrict__ vb,
                                                            const float *__restrict__ vc, const int64_t *__restrict__ subset, int64_t n_sub,
                                                            int64_t num_nodes, const int32_t *__restrict__ new_id,
                                                            int64_t *__restrict__ kept, const int64_t *__restrict__ sub_rowptr,
                                                            int32_t *__restrict__ sub_col, float *__restrict__ sa, float *__restrict__ sb,
                                                            float *__restrict__ sc, int64_t *__restrict__ coo_row, int64_t *__restrict__ coo_col) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sub; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = subset[s];
        int64_t out = FILL ? sub_rowptr[s] : 0;
        if (r >= 0 && r < num_nodes && new_id[r] == (int32_t)s) {   // a duplicated node keeps its edges at its last position only
            for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
                const int32_t c = new_id[col[e]];
                if (c < 0) continue;
                if (FILL) {
                    sub_col[out] = c;
                    sa[out] = va[e];
                    if (vb) sb[out] = vb[e];
                    if (vc) sc[out] = vc[e];
                    if (coo_row) {
                        coo_row[out] = s;
                        coo_col[out] = c;
                    }
                }
                ++out;
            }
        }
        if (!FILL) kept[s] = out;
    }
}

inline unsigned grid_for(int64_t n, int threads = 256, int per_sm = 8) {
    int64_t want = pg_ceil_div(n, threads);
    const int64_t cap = (int64_t)PG_NUM_SMS * per_sm;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}
}  // namespace

extern "C" int pg_next_node_labels(const int64_t *d_rowptr, const int64_t *d_dst, const float *d_weight, int64_t num_nodes,
                                   int64_t *d_labels, pg_stream_t stream) {
    PG_CHECK_ARG(num_nodes >= 0, "pg_next_node_labels: bad shape");
    if (num_nodes == 0) return PG_OK;
    PG_CHECK_ARG(d_rowptr && d_labels, "pg_next_node_labels: null buffer");
    next_node_labels_kernel<<<grid_for(num_nodes), 256, 0, pg_cu(stream)>>>(d_rowptr, d_dst, d_weight, num_nodes, d_labels);
    PG_CUDA_LAUNCH_CHECK("next_node_labels_kernel");
    return PG_OK;
}

extern "C" int pg_ngram_feature_init(const int64_t *d_code, int64_t num_nodes, const int64_t *d_prev_code, int64_t num_prev, int sigma,
                                     int n, const float *d_prev_emb, int64_t ld, int F, float *d_x, int64_t ldx, pg_stream_t stream) {
    PG_CHECK_ARG(num_nodes >= 0 && num_prev >= 0 && sigma >= 1 && n >= 2 && n <= 7 && F >= 1 && ld >= F && ldx >= F,
                 "pg_ngram_feature_init: bad shape (needs n >= 2)");
    if (num_nodes == 0) return PG_OK;
    PG_CHECK_ARG(d_code && d_x && (num_prev == 0 || (d_prev_code && d_prev_emb)), "pg_ngram_feature_init: null buffer");
    int64_t pow_nm1 = 1;
    for (int k = 0; k < n - 1; ++k) pow_nm1 *= sigma;
    feature_init_kernel<<<grid_for(num_nodes * 32), 256, 0, pg_cu(stream)>>>(d_code, num_nodes, d_prev_code, num_prev, sigma, pow_nm1,
                                                                              d_prev_emb, ld, F, d_x, ldx);
    PG_CUDA_LAUNCH_CHECK("feature_init_kernel");
    return PG_OK;
}

extern "C" int pg_pool_proteins(const uint8_t *d_seqs, const int64_t *d_offsets, int64_t num_proteins, int n,
                                const uint8_t *d_rank_of_byte, int sigma, const int32_t *d_code_to_id, const float *d_emb, int64_t ld,
                                int F, float *d_out, int64_t ldout, uint8_t *d_valid, pg_stream_t stream) {
    PG_CHECK_ARG(num_proteins >= 0 && n >= 1 && n <= 6 && sigma >= 1 && sigma <= 254 && F >= 1 && F <= 4 * kPoolThreads && ld >= F &&
                     ldout >= F,
                 "pg_pool_proteins: bad shape (F <= 512, sigma <= 254)");
    if (num_proteins == 0) return PG_OK;
    PG_CHECK_ARG(d_seqs && d_offsets && d_rank_of_byte && d_code_to_id && d_emb && d_out && d_valid, "pg_pool_proteins: null buffer");
    int64_t pow_n = 1;
    for (int k = 0; k < n; ++k) pow_n *= sigma;
    const size_t smem = (size_t)((pow_n + 31) / 32) * 4 + (size_t)kPoolListCap * 4;
    if (smem > 220 * 1024) {
        pg_set_error("pg_pool_proteins: sigma^n = %lld codes do not fit the shared-memory bitmap (n=%d sigma=%d)", (long long)pow_n, n, sigma);
        return PG_ERANGE;
    }
    PG_CUDA_CALL(cudaFuncSetAttribute(pool_proteins_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((size_t)(220 * 1024) / (smem + 2048));
    if (per_sm > 12) per_sm = 12;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)PG_NUM_SMS * per_sm;
    if (grid > num_proteins) grid = num_proteins;
    pool_proteins_kernel<<<(unsigned)grid, kPoolThreads, smem, pg_cu(stream)>>>(d_seqs, d_offsets, num_proteins, n, d_rank_of_byte,
                                                                                (uint32_t)sigma, (uint32_t)pow_n, d_code_to_id, d_emb, ld,
                                                                                F, d_out, ldout, d_valid);
    PG_CUDA_LAUNCH_CHECK("pool_proteins_kernel");
    return PG_OK;
}

extern "C" size_t pg_subgraph_ws_bytes(int64_t n_sub) {
    return pg_align_up((size_t)(n_sub > 0 ? n_sub : 1) * sizeof(int64_t), 256) + pg_scan_ws_bytes(n_sub > 0 ? n_sub : 1) + 512;
}

extern "C" int pg_subgraph_sizes(const int64_t *d_rowptr, const int32_t *d_col, int64_t num_nodes, const int64_t *d_subset, int64_t n_sub,
                                 int32_t *d_new_id, int64_t *d_sub_rowptr, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    PG_CHECK_ARG(num_nodes >= 0 && n_sub >= 0 && num_nodes < (1ll << 31), "pg_subgraph_sizes: bad shape");
    PG_CHECK_ARG(d_sub_rowptr && (num_nodes == 0 || d_new_id), "pg_subgraph_sizes: null output");
    cudaStream_t st = pg_cu(stream);
    if (num_nodes > 0) PG_CUDA_CALL(cudaMemsetAsync(d_new_id, 0xFF, (size_t)num_nodes * sizeof(int32_t), st));   // -1
    if (n_sub == 0) {
        PG_CUDA_CALL(cudaMemsetAsync(d_sub_rowptr, 0, sizeof(int64_t), st));
        return PG_OK;
    }
    PG_CHECK_ARG(d_rowptr && d_col && d_subset && d_ws, "pg_subgraph_sizes: null buffer");
    if (ws_bytes < pg_subgraph_ws_bytes(n_sub)) {
        pg_set_error("pg_subgraph_sizes: workspace too small (%zu < %zu)", ws_bytes, pg_subgraph_ws_bytes(n_sub));
        return PG_EWORKSPACE;
    }
    PgArena a(d_ws, ws_bytes);
    int *bad = a.take<int>(64);
    int64_t *kept = a.take<int64_t>((size_t)n_sub);
    const size_t scan_bytes = pg_scan_ws_bytes(n_sub);
    void *scan_ws = a.take<char>(scan_bytes);
    PG_CUDA_CALL(cudaMemsetAsync(bad, 0, sizeof(int), st));
    subgraph_mark_kernel<<<grid_for(n_sub), 256, 0, st>>>(d_subset, n_sub, num_nodes, d_new_id, bad);
    PG_CUDA_LAUNCH_CHECK("subgraph_mark_kernel");
    subgraph_rows_kernel<false><<<grid_for(n_sub), 256, 0, st>>>(d_rowptr, d_col, nullptr, nullptr, nullptr, d_subset, n_sub, num_nodes, d_new_id,
                                                                 kept, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    PG_CUDA_LAUNCH_CHECK("subgraph_rows_kernel<count>");
    return pg_exclusive_scan_i64(kept, d_sub_rowptr, n_sub, d_sub_rowptr + n_sub, scan_ws, scan_bytes, st);
}

extern "C" int pg_subgraph_fill(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val_a, const float *d_val_b,
                                const float *d_val_c, int64_t num_nodes, const int64_t *d_subset, int64_t n_sub, const int32_t *d_new_id,
                                const int64_t *d_sub_rowptr, int32_t *d_sub_col, float *d_sub_a, float *d_sub_b, float *d_sub_c,
                                int64_t *d_coo_row, int64_t *d_coo_col, pg_stream_t stream) {
    PG_CHECK_ARG(num_nodes >= 0 && n_sub >= 0, "pg_subgraph_fill: bad shape");
    if (n_sub == 0 || num_nodes == 0) return PG_OK;
    PG_CHECK_ARG(d_rowptr && d_col && d_val_a && d_subset && d_new_id && d_sub_rowptr && d_sub_col && d_sub_a, "pg_subgraph_fill: null buffer");
    PG_CHECK_ARG((!d_val_b || d_sub_b) && (!d_val_c || d_sub_c) && (!d_coo_row == !d_coo_col), "pg_subgraph_fill: inconsistent optional outputs");
    subgraph_rows_kernel<true><<<grid_for(n_sub), 256, 0, pg_cu(stream)>>>(d_rowptr, d_col, d_val_a, d_val_b, d_val_c, d_subset, n_sub, num_nodes,
                                                                           d_new_id, nullptr, d_sub_rowptr, d_sub_col, d_sub_a, d_sub_b,
                                                                           d_sub_c, d_coo_row, d_coo_col);
    PG_CUDA_LAUNCH_CHECK("subgraph_rows_kernel<fill>");
    return PG_OK;
}
