/* pgb200.h -- C ABI of libpgb200.so: the B200 (sm_100a) hot paths of ProtGram-DirectGCN.
 *
 * The reference (iebeid/ProtGram-DirectGCN) is pure Python and has no FFI registry; the drop-in
 * boundary is its Python class API.  Each entry point below replaces the *body* of the reference
 * function cited next to it; the same-named Python classes under protgram-directgcn_b200/host/
 * bind these with ctypes (see INTEGRATION.md for the stub a maintainer adds to the reference).
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller (torch's allocator); the
 *     library borrows it for the call, allocates nothing and keeps no global state.
 *   - scratch is passed explicitly: (d_ws, ws_bytes); the matching *_ws_bytes() query is pure.
 *   - `stream` is a cudaStream_t (passed as void*); every call is asynchronous on it.
 *   - return 0 on success, a negative PG_E* code otherwise; pg_last_error() gives the
 *     thread-local message.  There is no CPU fallback anywhere.
 *
 * Corpus buffer (input of hot path A): the byte string
 *       [' ' only before global sequence #0]  seq_0 ' ' 0xFF  seq_1 ' ' 0xFF ...
 *   i.e. each padded sequence of reference src/pipeline/data_builder.py:29-35 followed by one
 *   separator byte PG_SEP.  Windows never contain a separator, so they never cross sequences
 *   (data_builder.py:40-42,47-50); sequence bytes must be 7-bit ASCII.
 *
 * Dense n-gram tables: symbols are ranked by ascending byte value among the bytes present in
 * the corpus (rank_of_byte), so the base-sigma number of an n-gram's ranks orders n-grams
 * exactly like Python's sorted() on the strings (data_builder.py:164,172-173).
 */
#ifndef PGB200_H
#define PGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_SEP 0xFF

#define PG_OK 0
#define PG_EINVAL (-1)   /* bad argument (size, alignment, null pointer)          */
#define PG_ECUDA (-2)    /* a CUDA runtime call / kernel launch failed            */
#define PG_EWORKSPACE (-3) /* workspace too small                                 */
#define PG_ERANGE (-4)   /* table would not fit the index type (sigma^(n+1) etc.) */

typedef void *pg_stream_t; /* cudaStream_t */

int pg_version(void);
const char *pg_last_error(void);
/* diagnostics only: kernels launched by this library since it was loaded (process-wide) */
unsigned long long pg_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Row f3: FASTA text -> corpus buffer on the HOST (no CUDA call; replaces the Python line loop of
 * data_utils.py:182-213 fused with the padding rule of data_builder.py:29-35,97-102).
 * pg_fasta_open mmaps the file (NULL + pg_last_error() on failure).  pg_fasta_next_chunk packs whole records
 * "[' ' before global sequence #0] SEQUENCE ' ' 0xFF" into out[0, cap) until the next record would not fit and
 * returns the bytes written (0 = end of file) or a negative code; records are dealt to `world` ranks in blocks
 * of `block` records (rank 0 of 1 = everything).  Same record rules as the reference parser, including its
 * early stop at a bare ">" header (pg_fasta_stopped_early); sequence bytes >= 0x80 are refused.
 * ---------------------------------------------------------------------------------------- */
#define PG_FASTA_ETOOSMALL (-10) /* not even one record fits the caller's buffer */
#define PG_FASTA_ENONASCII (-11) /* non-ASCII byte in sequence text */
typedef struct pg_fasta_reader pg_fasta_reader;
pg_fasta_reader *pg_fasta_open(const char *path);
void pg_fasta_close(pg_fasta_reader *reader);
int64_t pg_fasta_file_bytes(const pg_fasta_reader *reader);
int64_t pg_fasta_records(const pg_fasta_reader *reader); /* records emitted so far, all ranks */
int pg_fasta_stopped_early(const pg_fasta_reader *reader);
int64_t pg_fasta_next_chunk(pg_fasta_reader *reader, uint8_t *out, int64_t cap, int rank, int world, int block);
/* The whole file (this rank's records) in one call, parsed by `threads` host threads over file ranges cut at header
 * lines; out needs file size + 16 bytes at most.  Returns the bytes written (same bytes as draining pg_fasta_next_chunk)
 * or a negative code; *n_records = records of all ranks, *stopped_early as above. */
int64_t pg_fasta_pack_parallel(const char *path, uint8_t *out, int64_t cap, int threads, int rank, int world, int block,
                               int64_t *n_records, int *stopped_early);
/* The same over ONE WINDOW of an open file, for files that do not fit host memory as a single corpus buffer: the caller walks
 * the file (start with *pos = 0 and first_index = 0; add *n_records to first_index after every call).  The window ends at the
 * first header line at or after *pos + window_bytes, or at the end of the file; *pos is advanced to it.  out needs the window's
 * file bytes + 16 at most; PG_FASTA_ETOOSMALL leaves *pos unchanged. */
int64_t pg_fasta_pack_window(pg_fasta_reader *reader, int64_t *pos, int64_t window_bytes, int64_t first_index, uint8_t *out,
                             int64_t cap, int threads, int rank, int world, int block, int64_t *n_records,
                             int *stopped_early);

/* 5-bit host format of the corpus buffer (what crosses PCIe when a corpus is streamed from host memory): 8 symbols in
 * 5 bytes, fixed code ' ' = 0, 'A'..'Z' = 1..26, '*' = 27, '-' = 28, '.' = 29, separator = 31 (also the tail padding).
 * pg_pack5_host (host code) returns the packed size = pg_pack5_bytes(n) or PG_EPACK when a byte has no code (keep the
 * byte format then); pg_unpack5 restores the n_symbols corpus bytes on the device (d_out 16-byte aligned). */
#define PG_EPACK (-12)
int64_t pg_pack5_bytes(int64_t n_symbols);
int64_t pg_pack5_host(const uint8_t *bytes, int64_t n_symbols, uint8_t *out);
int pg_unpack5(const uint8_t *d_packed, int64_t n_symbols, uint8_t *d_out, pg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Hot path A, part 1: n-gram transition counting
 * ---------------------------------------------------------------------------------------- */

/* Alphabet discovery: d_present256[b] = 1 for every byte value b != PG_SEP that occurs.
 * (The reference's alphabet is data-defined: data_utils.py:207 upper-cases but never filters.)
 * d_present256 must be zeroed by the caller (so shards can be OR-reduced across GPUs). */
int pg_byte_presence(const uint8_t *d_buf, int64_t nbytes, uint32_t *d_present256, pg_stream_t stream);

/* Synthetic protein-like corpus generated in place: nseq sequences of seq_len
 * residues, residue (s, j) a pure function of (seed, first_seq + s, j) so any sharding sees the
 * same corpus.  Layout = corpus buffer; bytes written = nseq*(seq_len+2) + leading_space. */
int pg_synth_corpus(uint8_t *d_buf, int64_t first_seq, int64_t nseq, int seq_len, uint32_t seed,
                    int leading_space, pg_stream_t stream);

/* Replaces data_builder.py:45-54 (+ :203-220 text spill, :267-273 CSV re-parse + groupby.size()):
 * for every separator-free window of n+1 bytes, d_bins[code(window)] += 1, where code is the
 * base-sigma number of the n+1 symbol ranks.  d_bins has sigma^(n+1) uint64 entries and is
 * ACCUMULATED into (zero it first; call once per corpus chunk / merge shards by summation).
 * d_short_present (sigma^n bytes, caller-zeroed) receives a 1 for n-grams that occur only as a
 * whole padded sequence of length exactly n (a node without any edge, data_builder.py:40 vs :47). */
int pg_ngram_count(const uint8_t *d_buf, int64_t nbytes, int n, const uint8_t *d_rank_of_byte,
                   int sigma, unsigned long long *d_bins, uint8_t *d_short_present,
                   void *d_ws, size_t ws_bytes, pg_stream_t stream);
/* Workspace of pg_ngram_count (256-byte aligned device memory, contents irrelevant): a status
 * word, one packed shared-memory table per CTA (written with plain stores and summed by a reduce
 * kernel: no flush atomics) and, for the 8-bit-lane variant, a scratch table for drained lanes;
 * that variant's result is only merged into d_bins when the kernel proved it exact, else a
 * strictly exact variant recounts (all on the stream, no host round trip).
 * d_ws may be NULL: every window is then one L2 atomic on d_bins (slower, same result). */
size_t pg_ngram_count_ws_bytes(int n, int sigma);
/* Same, sized for a corpus buffer of `nbytes`: tables too large for one CTA's shared memory (n >= 4 at
 * sigma = 21) are counted by a partition pass (windows appended to sigma^2 buckets keyed by their first two
 * symbols, 4 B per window) followed by shared-memory counting per bucket; the workspace then also holds the
 * bucket entries of one corpus chunk (<= 2^30 windows = 4 GiB) and, for 8-bit lanes, the scratch table.
 * pg_ngram_count picks that variant when the workspace it is given is large enough, else falls back to
 * L2 atomics on d_bins (same result). */
size_t pg_ngram_count_ws_bytes_for(int n, int sigma, int64_t nbytes);

/* Test hook: pins the variant of pg_ngram_count so parity tests can cover every kernel.
 * AUTO: widest shared-memory lanes that fit one CTA (32/16 bit: strict; 8 bit: scratch + hazard
 * check + gated strict recount), L2 REDs for tables beyond 4 key-range splits or tiny corpora.
 * STRICT: the 32/16-bit variants even for tiny corpora; FAST8*: the 8-bit variant where it applies. */
enum { PG_COUNT_AUTO = 0, PG_COUNT_GLOBAL = 1, PG_COUNT_STRICT = 2, PG_COUNT_FAST8 = 3, PG_COUNT_FAST8_FORCE_HAZARD = 4,
       PG_COUNT_PARTITIONED = 5 /* partition + shared-memory count even for tiny corpora (needs the _for workspace) */,
       PG_COUNT_PARTITIONED_FORCE_HAZARD = 6 };
void pg_debug_count_variant(int variant);

/* Replaces data_builder.py:151-177 (distinct + sorted ids) and :281-286 (edge table).
 * Step 1: d_sizes[0] = #nodes (distinct n-grams), d_sizes[1] = #unique transitions; also leaves
 *         the node-id table and edge offsets in the workspace for step 2.
 * Step 2: node codes (ascending = string order; id = position), and the edge table sorted by
 *         (src, dst): exactly the coalesced row-major order of A_out_w (graph_utils.py:154). */
size_t pg_graph_extract_ws_bytes(int n, int sigma);
int pg_graph_extract_sizes(const unsigned long long *d_bins, const uint8_t *d_short_present, int n,
                           int sigma, int64_t *d_sizes, void *d_ws, size_t ws_bytes, pg_stream_t stream);
int pg_graph_extract_fill(const unsigned long long *d_bins, int n, int sigma, int64_t num_nodes,
                          int64_t num_edges, int64_t *d_node_code, int64_t *d_src, int64_t *d_dst,
                          int64_t *d_count, void *d_ws, size_t ws_bytes, pg_stream_t stream);

/* Key-range form of the two steps above for tables merged by an NCCL REDUCE-SCATTER over key ranges
 * (SURVEY.md 8(e) "Builder"; replaces data_builder.py:151-177 and :267-286 on several GPUs): the caller
 * holds the summed bins of the source n-gram codes [code_lo, code_lo + codes), i.e. of the keys
 * [code_lo * sigma, (code_lo + codes) * sigma) -- whole rows of A_out_w -- in d_bins_local (a padded
 * last range may reach past the table; those bins are never read).
 *   _range_mark : ORs the presence of the local keys' source / target n-grams into d_present
 *                 (uint8[sigma^n], preset with the short-sequence flags; the host MAX-reduces it over
 *                 the ranks afterwards) and reports d_sizes[0] = number of local edges
 *   pg_node_ids_from_presence : reduced presence -> d_node_id[code] = rank among the present codes
 *                 (= id of the n-gram in sorted order), d_sizes[0] = number of nodes
 *   pg_node_codes_emit : d_node_code[id] = code (ascending)
 *   _range_fill : the local keys' edges (src id, dst id, count) sorted by (src, dst); the ranks' lists
 *                 concatenated in rank order are the whole coalesced edge table.
 * The workspace carries the edge offsets from _range_mark to _range_fill. */
size_t pg_graph_extract_range_ws_bytes(int sigma, int64_t codes);
int pg_graph_extract_range_mark(const unsigned long long *d_bins_local, int n, int sigma, int64_t code_lo,
                                int64_t codes, uint8_t *d_present, int64_t *d_sizes, void *d_ws,
                                size_t ws_bytes, pg_stream_t stream);
size_t pg_node_ids_ws_bytes(int64_t ngrams);
int pg_node_ids_from_presence(const uint8_t *d_present, int64_t ngrams, int64_t *d_node_id,
                              int64_t *d_sizes, void *d_ws, size_t ws_bytes, pg_stream_t stream);
int pg_node_codes_emit(const uint8_t *d_present, const int64_t *d_node_id, int64_t ngrams,
                       int64_t *d_node_code, pg_stream_t stream);
int pg_graph_extract_range_fill(const unsigned long long *d_bins_local, int n, int sigma, int64_t code_lo,
                                int64_t codes, const int64_t *d_node_id, int64_t num_edges_local,
                                int64_t *d_src, int64_t *d_dst, int64_t *d_count, void *d_ws,
                                size_t ws_bytes, pg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Hot path A, part 2: adjacency + propagation matrices (graph_utils.py:140-287)
 * ---------------------------------------------------------------------------------------- */

/* Stable LSD radix sort of 64-bit keys (low `key_bits` bits significant) carrying a 32-bit
 * payload.  Result is left in d_keys / d_vals (the alt buffers are scratch). */
size_t pg_sort_pairs_ws_bytes(int64_t n);
int pg_sort_pairs(unsigned long long *d_keys, unsigned long long *d_keys_alt, uint32_t *d_vals,
                  uint32_t *d_vals_alt, int64_t n, int key_bits, void *d_ws, size_t ws_bytes,
                  pg_stream_t stream);

/* Coalesce an arbitrary edge table (graph_utils.py:154 `.coalesce()`): sort by (src, dst) and
 * sum duplicates.  d_sizes[0] receives the number of unique edges E'; outputs are written to
 * the first E' slots of d_src_out/d_dst_out/d_w_out (each sized for `nnz`). */
size_t pg_coo_coalesce_ws_bytes(int64_t nnz);
int pg_coo_coalesce(const int64_t *d_src, const int64_t *d_dst, const float *d_w, int64_t nnz,
                    int64_t num_nodes, int64_t *d_src_out, int64_t *d_dst_out, float *d_w_out,
                    int64_t *d_sizes, void *d_ws, size_t ws_bytes, pg_stream_t stream);

/* From the coalesced A_out_w (sorted, unique) build everything DirectedNgramGraph holds:
 *   A_in_w            = A_out_w^T, coalesced                              (graph_utils.py:158)
 *   shared pattern    = pattern(A) U pattern(A^T) U I, row-major sorted   (:173-195, :252-269)
 *   mathcal_A_out/in  = sqrt(0.5*(An_ij^2 + An_ji^2) + eps) + [i==j],  An = D^-1 A   (:198-273)
 *   A_undirected_norm = deg^-1/2 (sym(pattern) + I) deg^-1/2 with the reference's
 *                       duplicate-self-loop semantics                      (:160-196)
 * Step 1 sorts/merges and reports d_sizes[0] = P (pattern nnz); step 2 fills.
 * d_rowptr is int64[num_nodes+1]; pattern columns are int32 (kernel format) -- the host side
 * widens to the reference's int64 COO where that API is exposed. */
size_t pg_normalize_ws_bytes(int64_t nnz, int64_t num_nodes);
int pg_normalize_sizes(const int64_t *d_src, const int64_t *d_dst, const float *d_w, int64_t nnz,
                       int64_t num_nodes, int64_t *d_sizes, void *d_ws, size_t ws_bytes,
                       pg_stream_t stream);
int pg_normalize_fill(const int64_t *d_src, const int64_t *d_dst, const float *d_w, int64_t nnz,
                      int64_t num_nodes, float eps, int64_t pattern_nnz,
                      int64_t *d_in_src, int64_t *d_in_dst, float *d_in_w,   /* A_in_w, nnz each */
                      int64_t *d_rowptr, int32_t *d_col,                     /* shared pattern   */
                      float *d_val_out, float *d_val_in, float *d_val_und,   /* P each           */
                      void *d_ws, size_t ws_bytes, pg_stream_t stream);

/* The same matrices for ONE ROW BLOCK [row_lo, row_lo + rows) of a graph whose rows are partitioned
 * over ranks (SURVEY.md 8(e) "Normalisation (a7-a9)": graph_utils.py:140-287 with one exchange).
 * The owner of the block passes its own out-edges (src in the block; d_out_*) and the in-edges of
 * its rows (dst in the block; d_in_*, delivered by the host side's all-to-all), both coalesced
 * (unique pairs), in any order.  Per-node vectors cross ranks between the calls:
 *   pg_degree_sums_rows         -> weighted out-/in-degree of the block's rows (fp64[rows]); all-gather
 *   pg_normalize_rows_sizes     -> d_sizes[0] = pattern nnz of the block, d_sizes[1] != 0 if an edge
 *                                  does not belong to the block or has an id >= num_nodes
 *   pg_normalize_rows_structure -> rowptr int64[rows+1] (local offsets), col int32 (GLOBAL ids), the
 *                                  block's rows of A_in_w (row-major sorted), native self-loop flags;
 *                                  undirected degree of row r = rowptr[r+1] - rowptr[r] + native_loop[r]; all-gather
 *   pg_normalize_rows_values    -> the three value arrays; d_rs_out / d_rs_in / d_deg are GLOBAL [num_nodes]
 * The workspace carries the sorted keys from _sizes to _values.  Rows computed here are bitwise
 * equal to the same rows of pg_normalize_fill on the whole graph. */
int pg_degree_sums_rows(const int64_t *d_out_src, const float *d_out_w, int64_t nnz_out,
                        const int64_t *d_in_dst, const float *d_in_w, int64_t nnz_in, int64_t row_lo,
                        int64_t rows, double *d_rs_out, double *d_rs_in, pg_stream_t stream);
size_t pg_normalize_rows_ws_bytes(int64_t nnz_out, int64_t nnz_in, int64_t rows);
int pg_normalize_rows_sizes(const int64_t *d_out_src, const int64_t *d_out_dst, int64_t nnz_out,
                            const int64_t *d_in_src, const int64_t *d_in_dst, int64_t nnz_in,
                            int64_t num_nodes, int64_t row_lo, int64_t rows, int64_t *d_sizes,
                            void *d_ws, size_t ws_bytes, pg_stream_t stream);
int pg_normalize_rows_structure(const float *d_in_w, int64_t nnz_out, int64_t nnz_in, int64_t num_nodes,
                                int64_t row_lo, int64_t rows, int64_t pattern_nnz,
                                int64_t *d_ain_row, int64_t *d_ain_col, float *d_ain_w, /* nnz_in each */
                                int64_t *d_rowptr, int32_t *d_col, uint8_t *d_native_loop,
                                void *d_ws, size_t ws_bytes, pg_stream_t stream);
int pg_normalize_rows_values(const float *d_out_w, const float *d_in_w, int64_t nnz_out, int64_t nnz_in,
                             int64_t num_nodes, int64_t row_lo, int64_t rows, const double *d_rs_out,
                             const double *d_rs_in, const int32_t *d_deg, const uint8_t *d_native_loop,
                             float eps, float *d_val_out, float *d_val_in, float *d_val_und,
                             void *d_ws, size_t ws_bytes, pg_stream_t stream);

/* CSR row pointers from sorted row ids; COO row ids from row pointers (int64 <-> CSR glue). */
int pg_rowptr_from_sorted(const int64_t *d_rows, int64_t nnz, int64_t num_rows, int64_t *d_rowptr,
                          pg_stream_t stream);
int pg_coo_from_csr(const int64_t *d_rowptr, const int32_t *d_col, int64_t num_rows, int64_t nnz,
                    int64_t *d_row_out, int64_t *d_col_out, pg_stream_t stream);

/* Group an arbitrary edge list by one endpoint (the layer's general-edge-list contract,
 * gnn_benchmarker.py:297-305): stable sort by `group` id -> CSR (rowptr, other endpoint, weight).
 * d_w may be NULL (unweighted: protgram_directgcn.py:138-139) -> weights of 1. */
size_t pg_edges_to_csr_ws_bytes(int64_t nnz);
int pg_edges_to_csr(const int64_t *d_group, const int64_t *d_other, const float *d_w, int64_t nnz,
                    int64_t num_nodes, int64_t *d_rowptr, int32_t *d_col, float *d_val,
                    void *d_ws, size_t ws_bytes, pg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Hot path B: DirectGCN propagation (protgram_directgcn.py:93-140) forward and backward
 * ---------------------------------------------------------------------------------------- */

/* Load-balancing plan for skewed degree distributions (optional, NULL = one lane group per row).
 * Rows with more than `chunk` stored entries ("long" rows) are cut into n_items slices of `chunk`
 * entries, one lane group each; slices write partial sums into d_partials, which a second kernel
 * adds up in slice order, so results stay bitwise reproducible.  The plan only depends on rowptr
 * and is built once per graph by the host (host/protgram_directgcn.py: SpmmPlan). */
typedef struct pg_spmm_plan {
    int32_t chunk;              /* entries per slice; rows with nnz > chunk are long          */
    int64_t n_long;             /* number of long rows (0 => plan ignored)                   */
    int64_t n_items;            /* total number of slices                                    */
    const int32_t *d_long_rows; /* [n_long]   row ids, ascending                             */
    const int64_t *d_item_ptr;  /* [n_long+1] first slice of each long row                   */
    const int32_t *d_item_row;  /* [n_items]  index into d_long_rows                         */
    float *d_partials;          /* scratch: n_items x max(nv*F) floats                       */
} pg_spmm_plan;

/* Fan-out SpMM (forward aggregation): for v < nv:
 *     Z[i, v*F : (v+1)*F] = sum_k val_v[k] * X[col[k], :]      k in row i of the shared CSR
 * nv = 3 with one shared pattern is the fused dual-direction + undirected propagate
 * (A_in X | A_out X | U X gathered once); nv = 1 is one propagate on its own CSR.
 * X is [num_cols x F] with row stride ldx, Z is [num_rows x *] with row stride ldz; the nv
 * output segments start at column z_off + v*F. */
int pg_spmm_fanout(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0,
                   const float *d_val1, const float *d_val2, int nv, int64_t num_rows, int F,
                   const float *d_x, int64_t ldx, float *d_z, int64_t ldz, int64_t z_off,
                   const pg_spmm_plan *plan, pg_stream_t stream);
/* Same with per-SOURCE-row scales: Z_v[i] = sum_j val_v[i,j] * s_v[j] * X[j] (d_s0 NULL = unscaled; scale_stride 1 = one
 * scale per row, 0 = a scalar, k > 1 = one scale every k floats, e.g. gate triples kept as rows of 4).  The layer's backward uses it with X = dY and s_v = gate_v on the symmetric structure:
 * dX = sum_v (A_v (g_v * dY)) W_v^T gathers F_out-wide rows once instead of the 3 F_in-wide gated gradient. */
int pg_spmm_fanout_scaled(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0,
                          const float *d_val1, const float *d_val2, int nv, int64_t num_rows, int F,
                          const float *d_x, int64_t ldx, float *d_z, int64_t ldz, int64_t z_off,
                          const float *d_s0, const float *d_s1, const float *d_s2, int scale_stride,
                          const pg_spmm_plan *plan, pg_stream_t stream);

/* The gathered matrix of a ROW-PARTITIONED graph: rows below `split` are read from `lo` (row stride ld_lo), the others from
 * `hi` (numbered from `split`, row stride ld_hi) -- the rank's own rows stay where they are, the halo rows it references
 * arrive in a compact second buffer (host/partitioned.py: HaloExchange) and the block's columns are renumbered once into
 * [own rows | halo rows].  hi = NULL: everything from `lo`. */
typedef struct pg_spmm_operand {
    const float *lo;
    int64_t ld_lo;
    const float *hi;
    int64_t ld_hi;
    int64_t split;
} pg_spmm_operand;
/* pg_spmm_fanout_scaled on a split operand; `z_vstride` = distance between the nv output segments (F for the plain calls;
 * the full feature width when a call handles one COLUMN CHUNK of the features -- pass x->lo / d_z advanced to the chunk,
 * F = chunk width: the exchange of chunk k+1 then overlaps the SpMM of chunk k, and each output element is still summed
 * in CSR order). */
int pg_spmm_fanout_split(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0, const float *d_val1,
                         const float *d_val2, int nv, int64_t num_rows, int F, const pg_spmm_operand *x, float *d_z,
                         int64_t ldz, int64_t z_off, int64_t z_vstride, const float *d_s0, const float *d_s1,
                         const float *d_s2, int scale_stride, const pg_spmm_plan *plan, pg_stream_t stream);

/* Fan-in SpMM (backward of the above over the transposed structure; forward of nothing else):
 *     Y[i, :] = (d_init ? init[i, :] : 0) + sum_v sum_k val_v[k] * G[col[k], g_off + v*F : +F]
 * With the symmetric shared pattern of reference-built graphs the transposed structure is the
 * structure itself; general edge lists pass the CSR grouped by source. */
int pg_spmm_fanin(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0,
                  const float *d_val1, const float *d_val2, int nv, int64_t num_rows, int F,
                  const float *d_g, int64_t ldg, int64_t g_off, const float *d_init, int64_t ldinit,
                  float *d_y, int64_t ldy, int accumulate, const pg_spmm_plan *plan, pg_stream_t stream);

/* pg_spmm_fanin on a split operand; the nv gradient segments of a row start at g_off + v * g_vstride. */
int pg_spmm_fanin_split(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val0, const float *d_val1,
                        const float *d_val2, int nv, int64_t num_rows, int F, const pg_spmm_operand *g, int64_t g_off,
                        int64_t g_vstride, const float *d_init, int64_t ldinit, float *d_y, int64_t ldy, int accumulate,
                        const pg_spmm_plan *plan, pg_stream_t stream);

/* Pack step of the halo exchange of a row-partitioned graph: d_dst[i, 0:w] = d_src[d_idx[i], 0:w] -- the rows of this rank
 * that a peer's block references, in the order the peer asked for them (one call per feature-column chunk). */
int pg_gather_rows(const float *d_src, int64_t ld_src, const int64_t *d_idx, int64_t count, int w, float *d_dst,
                   int64_t ld_dst, pg_stream_t stream);

/* ---- halo exchange over NVLink peer memory (csrc/peer.cu; one process per GPU, one node) ------------------------------------
 * pg_peer_alloc: device buffer (zeroed) + its 64-byte CUDA-IPC handle, which the host side all-gathers; pg_peer_open maps a
 * peer's buffer into this process.  pg_halo_push: ONE kernel stores the rows d_src[d_idx[i], 0:w] that peer p asked for
 * (rows h_row_begin[p] .. h_row_begin[p+1] of d_idx) into h_peer_dst[p] (row stride ld_dst) and, when all stores are fenced,
 * writes `epoch` to h_peer_flag[p] for every peer (NULL = no flag: the rank itself); the work is ordered by peer self + 1,
 * self + 2, ... so that the ranks never all store into the same GPU.  pg_halo_wait: one CTA on the consumer's
 * stream that returns once every peer's flag word in d_flags has reached `epoch`; it gives up after ~4 s and raises
 * *d_error_flag (1 + peer) instead of hanging.  h_* arrays are HOST arrays of `world` (row_begin: world + 1) entries. */
#define PG_MAX_PEERS 16
int pg_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle64);
int pg_peer_open(const unsigned char *handle64, void **d_ptr);
int pg_peer_close(void *d_ptr);
int pg_peer_free(void *d_ptr);
int pg_halo_push(const float *d_src, int64_t ld_src, const int64_t *d_idx, const int64_t *h_row_begin, float *const *h_peer_dst,
                 uint32_t *const *h_peer_flag, int world, int self, int w, int64_t ld_dst, uint32_t epoch,
                 unsigned int *d_done_counter, pg_stream_t stream);
int pg_halo_wait(const uint32_t *d_flags, int world, int self, uint32_t epoch, int *d_error_flag, pg_stream_t stream);

/* Fused dense transform of one DirectGCN layer (the collapsed algebra of SURVEY.md 7.2):
 *   A_ext[i, :] = [ a_i*Z_in[i] | b_i*Z_out[i] | c_i*Z_und[i] | X[i] (if has_res) | a_i b_i c_i | 1 (if has_res) ]
 *   Y = A_ext @ W_ext (+ X if add_identity) + constant[i]        W_ext: [K_ext x F_out] row-major
 *   H = leaky_relu(Y, slope) if slope != 1 else Y
 * gates a,b,c are per-row vectors (gate_stride = 1) or scalars broadcast (gate_stride = 0).
 * K_ext = 3*F_in + (has_res ? F_in : 0) + 3 + (has_res ? 1 : 0). */
int pg_layer_gemm_fwd(const float *d_z, int64_t ldz, const float *d_x, int64_t ldx,
                      const float *d_gate_a, const float *d_gate_b, const float *d_gate_c,
                      int gate_stride, const float *d_w_ext, const float *d_constant,
                      int64_t ldconst, int64_t num_rows, int F_in, int F_out, int has_res,
                      int add_identity, float slope, float *d_h, int64_t ldh, pg_stream_t stream);

/* Same contract as pg_layer_gemm_fwd, computed on the tcgen05 tensor cores with a 3 x TF32 operand
 * split (fp32-level accuracy, see csrc/gemm_tc.cu).  Requires F_in % 4 == 0, F_out % 16 == 0,
 * F_out <= 256 (pg_layer_gemm_fwd_tc_supported) and 16-byte aligned operands; d_ws holds the
 * pre-split weight image.  pg_layer_gemm_fwd_tc_check (host-synchronising, for tests) reports a
 * tensor-pipeline watchdog expiry instead of letting a mis-programmed MMA hang the device. */
int pg_layer_gemm_fwd_tc_supported(int F_in, int F_out);
size_t pg_layer_gemm_fwd_tc_ws_bytes(int F_in, int F_out, int has_res);
int pg_layer_gemm_fwd_tc(const float *d_z, int64_t ldz, const float *d_x, int64_t ldx,
                         const float *d_gate_a, const float *d_gate_b, const float *d_gate_c,
                         int gate_stride, const float *d_w_ext, const float *d_constant,
                         int64_t ldconst, int64_t num_rows, int F_in, int F_out, int has_res,
                         int add_identity, float slope, float *d_h, int64_t ldh, void *d_ws,
                         size_t ws_bytes, pg_stream_t stream);
int pg_layer_gemm_fwd_tc_check(const void *d_ws, int F_in, int F_out, int has_res, pg_stream_t stream);

/* dY = dH * leaky_relu'(H)   (elementwise; also the gradient of `constant`). */
int pg_lrelu_bwd(const float *d_dh, const float *d_h, float slope, int64_t numel, float *d_dy,
                 pg_stream_t stream);

/* dA = dY @ W_ext[:K_data]^T, then split: for the three Z segments
 *     dgate_v[i] = <dA_v[i], Z_v[i]> + <dY[i], beta_v>,   dZ_v[i] = gate_v[i] * dA_v[i]
 * and the residual segment (if has_res) goes to d_dxres.  Scalar gates (gate_stride 0) still
 * get per-row dgate values; the host sums them. */
int pg_layer_gemm_bwd_data(const float *d_dy, int64_t lddy, const float *d_w_ext, const float *d_z,
                           int64_t ldz, const float *d_gate_a, const float *d_gate_b,
                           const float *d_gate_c, int gate_stride, int64_t num_rows, int F_in,
                           int F_out, int has_res, float *d_dz, int64_t lddz, float *d_dxres,
                           int64_t lddxres, float *d_dgate /* [3 x num_rows] */, pg_stream_t stream);

/* dW_ext = A_ext^T @ dY   ([K_ext x F_out]; split over rows, reduced deterministically). */
size_t pg_layer_gemm_bwd_weight_ws_bytes(int64_t num_rows, int F_in, int F_out, int has_res);
int pg_layer_gemm_bwd_weight(const float *d_z, int64_t ldz, const float *d_x, int64_t ldx,
                             const float *d_gate_a, const float *d_gate_b, const float *d_gate_c,
                             int gate_stride, const float *d_dy, int64_t lddy, int64_t num_rows,
                             int F_in, int F_out, int has_res, float *d_dw_ext, void *d_ws,
                             size_t ws_bytes, pg_stream_t stream);

/* The two backward GEMMs on the tcgen05 tensor cores (3 x TF32 operand split like the forward; same contracts,
 * shapes as pg_layer_gemm_fwd_tc_supported, 16-byte aligned operands).  Data gradient: rows of dY against the
 * pre-split image of W_ext[:K_data]^T, output columns cut into blocks of <= 256 TMEM columns, then the same
 * gating / gate-gradient pass as the SIMT path.  Weight gradient: producers transpose both operands into the
 * K-major layout (kind::tf32 takes no MN-major operands), rows split over CTAs, partials summed in fixed order.  pg_tc_check (host-synchronising,
 * tests only) reads the watchdog flag a call leaves at d_ws + need - 256, need = that call's *_ws_bytes. */
size_t pg_layer_gemm_bwd_data_tc_ws_bytes(int F_in, int F_out, int has_res);
int pg_layer_gemm_bwd_data_tc(const float *d_dy, int64_t lddy, const float *d_w_ext, const float *d_z,
                              int64_t ldz, const float *d_gate_a, const float *d_gate_b,
                              const float *d_gate_c, int gate_stride, int64_t num_rows, int F_in,
                              int F_out, int has_res, float *d_dz, int64_t lddz, float *d_dxres,
                              int64_t lddxres, float *d_dgate /* [3 x num_rows] */, void *d_ws,
                              size_t ws_bytes, pg_stream_t stream);
size_t pg_layer_gemm_bwd_weight_tc_ws_bytes(int64_t num_rows, int F_in, int F_out, int has_res);
int pg_layer_gemm_bwd_weight_tc(const float *d_z, int64_t ldz, const float *d_x, int64_t ldx,
                                const float *d_gate_a, const float *d_gate_b, const float *d_gate_c,
                                int gate_stride, const float *d_dy, int64_t lddy, int64_t num_rows,
                                int F_in, int F_out, int has_res, float *d_dw_ext, void *d_ws,
                                size_t ws_bytes, pg_stream_t stream);
/* Input gradient through the transposed structure (symmetric shared pattern): with d_t = [T_in | T_out | T_und],
 * T_v = A_v (gate_v * dY) from pg_spmm_fanout_scaled,  dX = d_t @ [W'_in^T; W'_out^T; W'_und^T] (+ dY @ W_res^T | + dY).
 * Replaces gathering the 3 F_in-wide gated gradient (pg_spmm_fanin) by gathering F_out-wide rows of dY. */
size_t pg_layer_gemm_bwd_dx_tc_ws_bytes(int F_in, int F_out, int has_res);
int pg_layer_gemm_bwd_dx_tc(const float *d_t, int64_t ldt, const float *d_dy, int64_t lddy, const float *d_w_ext,
                            int64_t num_rows, int F_in, int F_out, int has_res, int add_identity, float *d_dx,
                            int64_t lddx, void *d_ws, size_t ws_bytes, pg_stream_t stream);
/* Gate gradients alone (the input gradient going through pg_spmm_fanout_scaled + pg_layer_gemm_bwd_dx_tc needs no dZ):
 * d_dgate[v][i] = <dY[i] W_v^T, Z_v[i]> + <dY[i], beta_v>, the data-gradient GEMM with a dot-product epilogue -- dZ is
 * never written.  Same values as pg_layer_gemm_bwd_data(_tc)'s d_dgate. */
size_t pg_layer_gate_grad_tc_ws_bytes(int64_t num_rows, int F_in, int F_out);
int pg_layer_gate_grad_tc(const float *d_dy, int64_t lddy, const float *d_w_ext, const float *d_z, int64_t ldz,
                          int64_t num_rows, int F_in, int F_out, int has_res, float *d_dgate /* [3 x num_rows] */,
                          void *d_ws, size_t ws_bytes, pg_stream_t stream);
int pg_tc_check(const void *d_ws, size_t need, pg_stream_t stream);
/* The same two steps of the regrouped backward on the SIMT fp32 path (layers narrower than the tensor-core threshold, i.e. every
 * n <= 3 graph's last layers and the benchmarker's small models): identical contracts, any F_in / F_out / alignment. */
size_t pg_layer_gate_grad_ws_bytes(int64_t num_rows, int F_in, int F_out);
int pg_layer_gate_grad(const float *d_dy, int64_t lddy, const float *d_w_ext, const float *d_z, int64_t ldz,
                       int64_t num_rows, int F_in, int F_out, int has_res, float *d_dgate /* [3 x num_rows] */,
                       void *d_ws, size_t ws_bytes, pg_stream_t stream);
int pg_layer_gemm_bwd_dx(const float *d_t, int64_t ldt, const float *d_dy, int64_t lddy, const float *d_w_ext,
                         int64_t num_rows, int F_in, int F_out, int has_res, int add_identity, float *d_dx,
                         int64_t lddx, pg_stream_t stream);

/* The two gradient GEMMs of the decoder's output layer in ONE pass over the N x C gradient matrix g that pg_softmax_nll leaves
 * behind (row f1; reference protgram_directgcn.py:177-180 under autograd):  dd = scale * g @ W2 [C x K],  dW2 = scale * g^T @ d [N x K].
 * K in {32, 64, 128} (pg_decoder_grads_supported); any N, C, row strides.  Fixed-order partial sums: bitwise reproducible. */
int pg_decoder_grads_supported(int K);
size_t pg_decoder_grads_ws_bytes(int64_t N, int C, int K);
int pg_decoder_grads(const float *d_g, int64_t ldg, const float *d_d, int64_t ldd, const float *d_w2, int64_t N, int C, int K,
                     float scale, float *d_dd, float *d_dw2, void *d_ws, size_t ws_bytes, pg_stream_t stream);

/* Plain Linear layers on the SIMT fp32 mainloop (the decoder MLP, reference protgram_directgcn.py:173-180,219: row f1), torch.nn.Linear
 * layout W [C x K]:  out = act(x W^T + bias);  dx = g W;  dW = g^T x (rows split over CTAs, fixed-order reduction);  column sums of
 * g (the bias gradient).  Any shape / alignment. */
int pg_linear_fwd(const float *d_x, int64_t ldx, int64_t num_rows, int K, const float *d_w, const float *d_bias, int C, int relu,
                  float *d_out, int64_t ldo, pg_stream_t stream);
int pg_linear_bwd_data(const float *d_g, int64_t ldg, int64_t num_rows, int C, const float *d_w, int K, float *d_dx, int64_t lddx,
                       pg_stream_t stream);
size_t pg_linear_bwd_weight_ws_bytes(int64_t num_rows, int C, int K);
int pg_linear_bwd_weight(const float *d_g, int64_t ldg, const float *d_x, int64_t ldx, int64_t num_rows, int C, int K, float *d_dw,
                         void *d_ws, size_t ws_bytes, pg_stream_t stream);
size_t pg_colsum_ws_bytes(int64_t num_rows, int C);
int pg_colsum(const float *d_g, int64_t ldg, int64_t num_rows, int C, float *d_out, void *d_ws, size_t ws_bytes, pg_stream_t stream);

/* Plain Linear on the same tensor-core kernel: out[N, C] = x[N, K] @ W[C, K]^T + bias (torch.nn.Linear layout, bias may be
 * NULL).  Used for the decoder's output layer (protgram_directgcn.py:177-180 with C = N classes, row f1).  K % 4 == 0,
 * x / out 16-byte aligned with row strides % 4 == 0 (pad the output rows to a multiple of 4 columns). */
size_t pg_linear_tc_ws_bytes(int K, int C);
int pg_linear_tc(const float *d_x, int64_t ldx, int64_t num_rows, int K, const float *d_w, const float *d_bias, int C,
                 float *d_out, int64_t ldo, void *d_ws, size_t ws_bytes, pg_stream_t stream);

/* Row-wise L2 normalisation  out = h / (||h||_2 + eps)   (models_utils.py:139-147). */
int pg_l2_normalize_rows(const float *d_h, int64_t ldh, int64_t num_rows, int F, float eps,
                         float *d_out, int64_t ldout, pg_stream_t stream);

/* Parameter side of one DirectGCNLayer (protgram_directgcn.py:34-66): the reference's own tensors,
 * row-major [F_out, F_in] weights, [F_out] biases, gate vectors of num_gate entries (N, or 1 for the
 * scalar-coefficient variant).  w_res / b_res: the stack's res_proj Linear (:159-160), NULL without. */
typedef struct pg_layer_params {
    const float *w_in, *w_out, *w_und, *w_sh;                       /* lin_main_in/out, lin_undirected, lin_shared .weight */
    const float *b_in, *b_out, *b_und, *bs_in, *bs_out, *bs_und;   /* bias_main_*, bias_undirected, bias_*_shared* */
    const float *w_res, *b_res;
    const float *c_in, *c_out, *c_dir, *c_und, *c_all;             /* C_in, C_out, C_directed, C_undirected, C_all (_vec) */
} pg_layer_params;
typedef struct pg_layer_param_grads {
    float *w_in, *w_out, *w_und, *w_sh, *b_in, *b_out, *b_und, *bs_in, *bs_out, *bs_und, *w_res, *b_res;
    float *c_in, *c_out, *c_dir, *c_und, *c_all;
} pg_layer_param_grads;

/* -> the operands of pg_layer_gemm_*: W_ext [3F_in (+F_in) + 3 (+1), F_out] and the gates
 * a = (C_all*C_dir)*C_in, b = (C_all*C_dir)*C_out, c = C_all*C_und.  Same fp32 operations as the
 * reference's tensor ops (:101-133), so bit-identical to composing them. */
int pg_pack_layer_params(const pg_layer_params *params, int64_t num_gate, int F_in, int F_out,
                         int has_res, float *d_w_ext, float *d_gate_a, float *d_gate_b,
                         float *d_gate_c, pg_stream_t stream);
/* Backward of the packing: dW_ext, da, db, dc -> gradients of every reference parameter. */
int pg_unpack_layer_param_grads(const pg_layer_params *params, const float *d_dw_ext,
                                const float *d_dgate_a, const float *d_dgate_b,
                                const float *d_dgate_c, int64_t num_gate, int F_in, int F_out,
                                int has_res, const pg_layer_param_grads *grads, pg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Row f1: loss of the next-node task, forward AND backward in one pass
 * ---------------------------------------------------------------------------------------- */

/* Replaces F.log_softmax(task_logits, dim=-1) (protgram_directgcn.py:221) + F.nll_loss(.., y)
 * (protgram_directgcn_trainer.py:90,94; reduction 'mean') and their autograd backward.
 * d_logits [n, ld] (c <= ld columns used) is OVERWRITTEN with d(loss)/d(logits) =
 * (softmax(row) - onehot(label)) * grad_scale; rows whose label is outside [0, c) are ignored
 * (zero gradient, zero loss: ignore_index semantics), grad_scale = 1 / #counted rows.
 * d_row_loss[n] = -log_softmax(row)[label]; d_loss[0] = grad_scale * sum(d_row_loss);
 * d_colsum[c] = column sums of the gradient (= gradient of the decoder bias).
 * All reductions run in a fixed order (bitwise reproducible). */
#define PG_SOFTMAX_NLL_MAX_CLASSES 28672
size_t pg_softmax_nll_ws_bytes(int64_t n, int64_t c);
int pg_softmax_nll(float *d_logits, int64_t ld, int64_t n, int64_t c, const int64_t *d_labels,
                   float grad_scale, float *d_row_loss, float *d_colsum, float *d_loss, void *d_ws,
                   size_t ws_bytes, pg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Rows f2 / f4: the callers on either side of the two hot paths
 * ---------------------------------------------------------------------------------------- */

/* f4, replaces _generate_next_node_labels (protgram_directgcn_trainer.py:222-237): labels[i] = the
 * successor of node i with the largest A_out_w weight (rows of the coalesced matrix given as
 * d_rowptr[n+1] over d_dst / d_weight); ties -> the first in (src, dst) order (the reference draws
 * one of the maximal successors at random: every one of them is admissible); no successor -> i. */
int pg_next_node_labels(const int64_t *d_rowptr, const int64_t *d_dst, const float *d_weight,
                        int64_t num_nodes, int64_t *d_labels, pg_stream_t stream);

/* f2a, replaces the feature hand-off between levels (protgram_directgcn_trainer.py:312-330):
 * x[i] = mean of the level-(n-1) embeddings of node i's prefix and suffix (n-1)-grams (those that
 * exist; zeros if neither).  d_code / d_prev_code: packed base-sigma node codes of level n / n-1
 * (ascending = node-id order, as pg_graph_extract_fill emits them), same alphabet. */
int pg_ngram_feature_init(const int64_t *d_code, int64_t num_nodes, const int64_t *d_prev_code,
                          int64_t num_prev, int sigma, int n, const float *d_prev_emb, int64_t ld,
                          int F, float *d_x, int64_t ldx, pg_stream_t stream);

/* f2b, replaces EmbeddingProcessor.pool_ngram_embeddings_for_protein_fast (models_utils.py:210-262):
 * d_out[p] = mean over the DISTINCT known n-grams of protein p of their embedding rows, summed in
 * ascending code order in fp32 (the reference's order when ids are ranks of the sorted n-grams),
 * d_valid[p] = 0 (and a zero row) when the protein holds no known n-gram.  d_seqs: the raw sequence
 * bytes back to back, d_offsets[P+1] their bounds; d_rank_of_byte[256]: symbol rank or 255 for bytes
 * outside the alphabet; d_code_to_id[sigma^n]: node id of a packed code or -1.  F <= 512. */
int pg_pool_proteins(const uint8_t *d_seqs, const int64_t *d_offsets, int64_t num_proteins, int n,
                     const uint8_t *d_rank_of_byte, int sigma, const int32_t *d_code_to_id,
                     const float *d_emb, int64_t ld, int F, float *d_out, int64_t ldout,
                     uint8_t *d_valid, pg_stream_t stream);

/* f4b, replaces the three torch_geometric.utils.subgraph(..., relabel_nodes=True) calls per cluster of
 * _create_clustered_subgraphs (protgram_directgcn_trainer.py:179-197) by ONE pass over the shared-pattern CSR:
 * keeps the stored entries whose row AND column are in `subset`, relabelled by position in `subset`
 * (node_idx[subset] = arange(len(subset))), rows in subset order, columns in stored order -- for an ascending
 * subset that is exactly the reference's edge order, and the result is again a sorted symmetric CSR.
 * Step 1 fills d_new_id[num_nodes] (-1 = outside) and d_sub_rowptr[n_sub + 1]; step 2 (after the caller read
 * d_sub_rowptr[n_sub] = kept entries and allocated) writes int32 columns, up to three value arrays and,
 * optionally, the int64 COO rows / columns of the reference layout. */
size_t pg_subgraph_ws_bytes(int64_t n_sub);
int pg_subgraph_sizes(const int64_t *d_rowptr, const int32_t *d_col, int64_t num_nodes, const int64_t *d_subset,
                      int64_t n_sub, int32_t *d_new_id, int64_t *d_sub_rowptr, void *d_ws, size_t ws_bytes,
                      pg_stream_t stream);
int pg_subgraph_fill(const int64_t *d_rowptr, const int32_t *d_col, const float *d_val_a, const float *d_val_b,
                     const float *d_val_c, int64_t num_nodes, const int64_t *d_subset, int64_t n_sub,
                     const int32_t *d_new_id, const int64_t *d_sub_rowptr, int32_t *d_sub_col, float *d_sub_a,
                     float *d_sub_b, float *d_sub_c, int64_t *d_coo_row, int64_t *d_coo_col, pg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PGB200_H */
