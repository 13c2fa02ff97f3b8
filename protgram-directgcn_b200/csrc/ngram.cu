// Hot path A, part 1: residue n-grams -> dense transition table -> node ids + edge table.
//
// Replaces the per-residue Python of reference src/pipeline/data_builder.py:38-54 and the
// Dask distinct()/groupby().size() of :151-177,267-273.  An edge IS an (n+1)-gram
// (window i -> window i+1), so transition counting is an (n+1)-gram histogram over the
// corpus buffer; node ids are ranks of present n-grams in base-sigma (= string) order.
//
// HBM traffic: 1 B/residue read (16 B vector loads, fully coalesced); the table updates are
// L2 atomics (RED.ADD.64), which is the real limiter -- see DESIGN.md "count kernel".
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ alphabet discovery
__global__ void __launch_bounds__(256) byte_presence_kernel(const uint8_t *__restrict__ buf, int64_t nbytes,
                                                            uint32_t *__restrict__ present256) {
    __shared__ uint32_t seen[256];
    seen[threadIdx.x] = 0;
    __syncthreads();
    const int64_t nvec = nbytes / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(buf);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 q = v[i];
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            seen[w[k] & 0xFF] = 1;
            seen[(w[k] >> 8) & 0xFF] = 1;
            seen[(w[k] >> 16) & 0xFF] = 1;
            seen[w[k] >> 24] = 1;
        }
    }
    if (blockIdx.x == 0) {  // tail bytes
        for (int64_t i = nvec * 16 + threadIdx.x; i < nbytes; i += blockDim.x) seen[buf[i]] = 1;
    }
    __syncthreads();
    if (seen[threadIdx.x] && threadIdx.x != PG_SEP) present256[threadIdx.x] = 1;
}

// ------------------------------------------------------------------ 5-bit host format -> corpus bytes (see csrc/fasta.cu)
__constant__ uint8_t kCode5ToByte[32] = {' ', 'A', 'B', 'C', 'D', 'E', 'F', 'G', 'H', 'I', 'J', 'K', 'L', 'M', 'N', 'O',
                                         'P', 'Q', 'R', 'S', 'T', 'U', 'V', 'W', 'X', 'Y', 'Z', '*', '-', '.', PG_SEP, PG_SEP};

// one thread = two 40-bit groups = 10 packed bytes -> 16 corpus bytes (one 16-byte store)
__global__ void __launch_bounds__(256) unpack5_kernel(const uint8_t *__restrict__ packed, int64_t n_symbols, uint8_t *__restrict__ out) {
    __shared__ uint8_t lut[32];
    if (threadIdx.x < 32) lut[threadIdx.x] = kCode5ToByte[threadIdx.x];
    __syncthreads();
    const int64_t pairs = (n_symbols + 15) / 16;
    const int64_t groups = (n_symbols + 7) / 8;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < pairs; t += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t g = 2 * t + h;
            unsigned long long word = ~0ull;   // a missing second group decodes to separators (never stored)
            if (g < groups) {
                const uint8_t *src = packed + g * 5;
                word = (unsigned long long)src[0] | ((unsigned long long)src[1] << 8) | ((unsigned long long)src[2] << 16) |
                       ((unsigned long long)src[3] << 24) | ((unsigned long long)src[4] << 32);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) w[2 * h + i / 4] |= (uint32_t)lut[(word >> (5 * i)) & 31u] << (8 * (i % 4));
        }
        const int64_t p0 = t * 16;
        if (p0 + 16 <= n_symbols) {
            *reinterpret_cast<uint4 *>(out + p0) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int i = 0; i < 16 && p0 + i < n_symbols; ++i) out[p0 + i] = (uint8_t)(w[i / 4] >> (8 * (i % 4)));
        }
    }
}

// ------------------------------------------------------------------ synthetic corpus
__constant__ uint32_t kAaCum16[20] = {5408,  6308,  9885,  14308, 16838, 21473, 22961, 26847, 30662, 37140,
                                      38723, 41383, 44486, 47063, 50689, 54995, 58502, 63002, 63720, 65536};
__constant__ char kAa[21] = "ACDEFGHIKLMNPQRSTVWY";

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(256) synth_corpus_kernel(uint8_t *__restrict__ buf, int64_t first_seq, int64_t nseq,
                                                           int seq_len, uint32_t seed, int leading_space) {
    const int64_t stride = seq_len + 2;
    const int64_t total = nseq * stride;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = t / stride;
        const int j = (int)(t - s * stride);
        uint8_t out;
        if (j == seq_len) out = ' ';
        else if (j == seq_len + 1) out = PG_SEP;
        else {
            const unsigned long long g = (unsigned long long)(first_seq + s) * (unsigned long long)seq_len + j;
            const uint32_t h = mix32((uint32_t)g ^ mix32((uint32_t)(g >> 32) ^ seed) ^ 0x9E3779B9u);
            const uint32_t u = h >> 16;
            int idx = 0;
#pragma unroll
            for (int k = 0; k < 20; ++k) idx += (kAaCum16[k] <= u);
            out = (uint8_t)kAa[idx];
        }
        buf[t + leading_space] = out;
    }
    if (leading_space && blockIdx.x == 0 && threadIdx.x == 0) buf[0] = ' ';
}

// ------------------------------------------------------------------ (n+1)-gram count
// One thread owns the 16 windows that END in its 16-byte vector; the M-1 bytes of look-back
// come from the previous vector (an L1 hit: the neighbouring thread loads it as its own).
// The window code rolls: code = code*sigma + r_new - r_old*sigma^M, exact modulo 2^32 because
// the true value is < sigma^M <= 2^32.  `emit(code)` is called once per separator-free window.
// Separator positions and window validity are handled as bit masks over the 24 loaded bytes
// (bit p <-> buffer position p0-8+p): a byte is a separator iff its top bit is set (sequence bytes
// are 7-bit ASCII by contract), a window ending at p is invalid iff any of its M bytes is one.
__device__ __forceinline__ uint32_t sep_nibble(uint32_t w) {
    // top bits of the 4 bytes -> 4 adjacent bits (multiply gathers b0..b3 into bits 28..31)
    return (((w >> 7) & 0x01010101u) * 0x10204080u) >> 28;
}

// Output: codes[k] = code of the window ending at byte p0+k, bit k of the returned mask = that
// window is separator-free.  (Branch-free so the 16 table updates of a thread can be in flight
// together.)
// The 24 bytes a thread works on: prev.x prev.y cur.x cur.y cur.z cur.w (buffer positions p0-8 .. p0+15);
// bytes outside the buffer read as separators.
__device__ __forceinline__ void load_vector(const uint8_t *__restrict__ buf, int64_t nbytes, int64_t p0, uint32_t (&w)[6]) {
    if (p0 + 16 <= nbytes) {
        const uint4 cur = *reinterpret_cast<const uint4 *>(buf + p0);
        w[2] = cur.x; w[3] = cur.y; w[4] = cur.z; w[5] = cur.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t x = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int64_t i = p0 + k * 4 + t;
                x |= (uint32_t)(i < nbytes ? buf[i] : (uint8_t)PG_SEP) << (8 * t);
            }
            w[2 + k] = x;
        }
    }
    if (p0 > 0) {
        const uint2 prev = *reinterpret_cast<const uint2 *>(buf + p0 - 8);
        w[0] = prev.x; w[1] = prev.y;
    } else {
        w[0] = w[1] = 0xFFFFFFFFu;  // before the buffer = separator
    }
}

// SPLIT2 (partitioned variant below): codes[k] = (code of the first two symbols) << 18 | code of the other M-2
// symbols, i.e. (key / sub_bins) << 18 | key % sub_bins with sub_bins = sigma^(M-2) <= 2^18.
template <int M, bool SPLIT2 = false>
__device__ __forceinline__ uint32_t scan_words(const uint32_t (&w)[6], int64_t nbytes, int64_t p0, const uint8_t *lut,
                                               uint32_t sigma, uint32_t sigma_pow_m, uint32_t sigma_pow_n,
                                               uint8_t *__restrict__ short_present, uint32_t (&codes)[16],
                                               uint32_t sub_bins = 0) {
    constexpr int LB = M - 1;  // look-back bytes (= n)
    uint32_t sep = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) sep |= sep_nibble(w[k]) << (4 * k);
    uint32_t inv = sep;  // bit p: the M-window ending at p holds a separator
#pragma unroll
    for (int k = 1; k < M; ++k) inv |= sep << k;
    const uint32_t valid = ~inv;

    uint32_t code = 0;
    uint32_t r_hist[M];
    const uint32_t neg_pow_m = 0u - sigma_pow_m;
#pragma unroll
    for (int i = 0; i < LB + 16; ++i) {
        const int pos = 8 - LB + i;
        const uint32_t r = lut[(w[pos >> 2] >> (8 * (pos & 3))) & 0xFFu];
        code = code * sigma + r;
        if (i >= M) code += r_hist[i % M] * neg_pow_m;  // - r_old * sigma^M (mod 2^32)
        r_hist[i % M] = r;
        if (i >= LB) {
            if (SPLIT2) {  // the window's oldest two ranks sit at (i+1) % M and (i+2) % M of the ring
                const uint32_t pre = r_hist[(i + 1) % M] * sigma + r_hist[(i + 2) % M];
                codes[i - LB] = (pre << 18) | (code - pre * sub_bins);
            } else {
                codes[i - LB] = code;
            }
        }
    }
    if (short_present != nullptr) {
        // a padded sequence of exactly n bytes: separator at s (inside this vector), n clean bytes
        // before it, separator (or buffer start) before those  -> a node without any edge
        uint32_t inv_n = sep;
#pragma unroll
        for (int k = 1; k < LB; ++k) inv_n |= sep << k;
        uint32_t cand = sep & ((~inv_n) << 1) & (sep << (LB + 1)) & 0x00FFFF00u;
        while (cand) {  // rare
            const int s_pos = __ffs(cand) - 1;
            cand &= cand - 1;
            if (p0 - 8 + s_pos >= nbytes) continue;  // padding past the end of the buffer
            uint32_t c = 0;
            for (int k = s_pos - LB; k < s_pos; ++k) c = c * sigma + lut[(w[k >> 2] >> (8 * (k & 3))) & 0xFFu];
            short_present[c % sigma_pow_n] = 1;
        }
    }
    return (valid >> 8) & 0xFFFFu;
}

template <int M>
__device__ __forceinline__ uint32_t scan_vector(const uint8_t *__restrict__ buf, int64_t nbytes, int64_t p0,
                                                const uint8_t *lut, uint32_t sigma, uint32_t sigma_pow_m,
                                                uint32_t sigma_pow_n, uint8_t *__restrict__ short_present,
                                                uint32_t (&codes)[16]) {
    uint32_t w[6];
    load_vector(buf, nbytes, p0, w);
    return scan_words<M>(w, nbytes, p0, lut, sigma, sigma_pow_m, sigma_pow_n, short_present, codes);
}

// Variant G: every window is one RED.ADD.64 into the dense table in L2.  Measured on B200:
// ~200 G updates/s for >= 194k bins, collapsing to 12 G/s at 441 bins (same-address serialisation),
// so it is only used when the table does not fit the shared-memory variant below.
// `gate` (optional): the kernel is a no-op unless *gate != 0 (strict recount after a flagged hazard).
template <int M>
__global__ void __launch_bounds__(256) ngram_count_kernel(const uint8_t *__restrict__ buf, int64_t nbytes,
                                                          const uint8_t *__restrict__ rank_of_byte, uint32_t sigma,
                                                          uint32_t sigma_pow_m, uint32_t sigma_pow_n,
                                                          unsigned long long *__restrict__ bins,
                                                          uint8_t *__restrict__ short_present, const int *__restrict__ gate) {
    if (gate != nullptr && *gate == 0) return;
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = rank_of_byte[threadIdx.x];
    __syncthreads();
    const int64_t nvec = (nbytes + 15) / 16;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        uint32_t codes[16];
        const uint32_t valid = scan_vector<M>(buf, nbytes, v * 16, lut, sigma, sigma_pow_m, sigma_pow_n, short_present, codes);
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if ((valid >> k) & 1u) atomicAdd(&bins[codes[k]], 1ull);
    }
}

// Variant S: privatised table in shared memory, persistent CTAs (one per SM, 1024 threads).
// Shared-memory atomics run at ~1.3 T updates/s on B200 (vs 0.2 T/s for L2 REDs).  The table is
// packed into LB-bit lanes of 32-bit words, LB = the widest of 32/16/8 whose table fits one CTA, so
// that ONE CTA holds the whole key space and every corpus byte is scanned once:
//   LB = 32  no overflow possible, fire-and-forget adds                     (<= 56k bins)
//   LB = 16  the single add that observes a lane at 32767 moves 32768 to `out` and subtracts it again;
//            a lane can only carry into its neighbour after 32768 further in-flight adds, more than
//            1024 threads x 16 windows can have pending: STRICTLY exact        (<= 112k bins)
//   LB = 8   same scheme with 127 -> 128.  The slack is only 128 adds, which a pathological corpus
//            (one window repeated >= 128 times inside a 16 KB tile, e.g. homopolymers) can exhaust
//            before the subtraction lands.  A carry needs some add to OBSERVE a lane >= 192 first
//            (adds move a lane by +1), and until the first carry every observation is exact, so
//            "no add observed >= 192" proves the result exact.  Such an observation raises *status;
//            the 8-bit kernel therefore counts into a zeroed SCRATCH table, which is merged into
//            the real table only if *status stayed 0, else a strict variant recounts (gate).
// Larger tables are cut into `splits` contiguous key ranges (SPLIT = true): CTA b serves range
// b % splits and walks the tiles of group b / splits (second read of a tile = L2 hit).
constexpr int kSmemCountThreads = 1024;
constexpr int64_t kSmemTableBytes = 224 * 1024;  // of the 227 KB per CTA (lut + counters take the rest)

template <int M, int LB, bool SPLIT>
__global__ void __launch_bounds__(kSmemCountThreads, 1) ngram_count_smem_kernel(
    const uint8_t *__restrict__ buf, int64_t nbytes, const uint8_t *__restrict__ rank_of_byte, uint32_t sigma,
    uint32_t sigma_pow_m, uint32_t sigma_pow_n, unsigned long long *__restrict__ out, uint8_t *__restrict__ short_present,
    int splits, uint32_t lanes_per_split, int *__restrict__ status, const int *__restrict__ gate,
    unsigned *__restrict__ partials) {
    constexpr uint32_t PER_WORD = 32 / LB;                 // lanes per 32-bit word
    constexpr uint32_t LSH = LB == 32 ? 0 : (LB == 16 ? 1 : 2);
    constexpr uint32_t LANE_MASK = LB == 32 ? 0xFFFFFFFFu : ((1u << (LB % 32)) - 1u);
    constexpr uint32_t HALF = LB == 32 ? 0u : (1u << ((LB - 1) % 32));  // drain quantum (128 / 32768)
    extern __shared__ unsigned tbl[];
    __shared__ uint8_t lut[256];
    if (gate != nullptr && *gate == 0) return;
    if (threadIdx.x < 256) lut[threadIdx.x] = rank_of_byte[threadIdx.x];
    const uint32_t words = (lanes_per_split + PER_WORD - 1) / PER_WORD;
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) tbl[i] = 0;
    __syncthreads();
    const int split = SPLIT ? blockIdx.x % splits : 0;
    const int64_t group = SPLIT ? blockIdx.x / splits : blockIdx.x, groups = SPLIT ? gridDim.x / splits : gridDim.x;
    const uint32_t lo = (uint32_t)split * lanes_per_split;
    const uint32_t hi = min(lo + lanes_per_split, sigma_pow_m);
    const uint32_t span = hi > lo ? hi - lo : 0u;
    uint8_t *sp = (split == 0) ? short_present : nullptr;
    const int64_t nvec = (nbytes + 15) / 16;
    bool hazard = false;
    // the loads of the next vector are issued before this one is processed (the loop body is ~400
    // instructions per warp: one iteration of lookahead hides the whole HBM latency)
    const int64_t vstep = groups * blockDim.x;
    int64_t vec = group * blockDim.x + threadIdx.x;
    uint32_t wn[6];
    if (vec < nvec) load_vector(buf, nbytes, vec * 16, wn);
    for (; vec < nvec; vec += vstep) {
        uint32_t wc[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) wc[k] = wn[k];
        if (vec + vstep < nvec) load_vector(buf, nbytes, (vec + vstep) * 16, wn);
        uint32_t codes[16];
        uint32_t valid = scan_words<M>(wc, nbytes, vec * 16, lut, sigma, sigma_pow_m, sigma_pow_n, sp, codes);
        if (SPLIT) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                codes[k] -= lo;
                if (codes[k] >= span) {
                    valid &= ~(1u << k);
                    codes[k] = 0;
                }
            }
        }
        if (LB == 32) {
#pragma unroll
            for (int k = 0; k < 16; ++k) atomicAdd(&tbl[codes[k]], (valid >> k) & 1u);
        } else {
            // Branch-free: invalid windows add 0.  `att` collects "this add saw its lane at >= HALF-1"
            // (top lane bit of old+inc); only then (rare) the slow path looks at the lane values.
            unsigned old[16];
            uint32_t att = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {  // 16 independent shared-memory atomics, issued back to back
                const uint32_t sh = (codes[k] * LB) & 31u;
                const uint32_t inc = ((valid >> k) & 1u) << sh;
                old[k] = atomicAdd(&tbl[codes[k] >> LSH], inc);
                att |= (old[k] + inc) & (inc << (LB - 1));
            }
            if (att) {
                // re-derive the codes (cheaper than keeping 16 more registers live in the hot loop)
                uint32_t valid2 = scan_vector<M>(buf, nbytes, vec * 16, lut, sigma, sigma_pow_m, sigma_pow_n, nullptr, codes);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    uint32_t c = codes[k];
                    if (SPLIT) {
                        c -= lo;
                        if (c >= span) continue;
                    }
                    if (!((valid2 >> k) & 1u)) continue;
                    const uint32_t sh = (c * LB) & 31u;
                    const uint32_t lane = (old[k] >> sh) & LANE_MASK;
                    if (lane == HALF - 1u) {  // this add took the lane to HALF: move HALF counts to the 64-bit table
                        atomicSub(&tbl[c >> LSH], HALF << sh);
                        atomicAdd(&out[c + lo], (unsigned long long)HALF);
                    }
                    if (LB == 8 && lane >= 192u) hazard = true;
                }
            }
        }
    }
    if (LB == 8 && hazard) *status = 1;
    __syncthreads();
    // The packed table goes to this CTA's slice of the workspace with plain coalesced stores; the
    // reduce kernel below sums the slices.  (Flushing with one RED per non-zero bin and CTA cost
    // 148 x 194k = 28.8 M L2 atomics = half of the kernel's time at C2.)
    unsigned *mine = partials + (size_t)blockIdx.x * words;
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) mine[i] = tbl[i];
}

// bins[k] += sum over the CTAs' packed tables (+ the drained counts in `scratch`, 8-bit variant).
// run_if: -1 always; 0 only if *status == 0 (8-bit result proven exact); 1 only if *status != 0
// (strict recount after a flagged hazard).  One thread per 32-bit word = 32/LB bins; fixed order.
template <int LB>
__global__ void __launch_bounds__(256) reduce_partials_kernel(const unsigned *__restrict__ partials, int groups, int splits,
                                                              uint32_t lanes_per_split, uint32_t sigma_pow_m,
                                                              const unsigned long long *__restrict__ scratch,
                                                              const int *__restrict__ status, int run_if,
                                                              unsigned long long *__restrict__ bins) {
    if (run_if >= 0 && (*status != 0) != (run_if != 0)) return;
    constexpr uint32_t PER_WORD = 32 / LB;
    constexpr uint32_t LANE_MASK = LB == 32 ? 0xFFFFFFFFu : ((1u << (LB % 32)) - 1u);
    constexpr int SLICES = 8;  // a CTA = 32 consecutive words x 8 slices of the CTA tables (coalesced 128 B rows)
    __shared__ unsigned long long sub[SLICES][32][PER_WORD];
    const uint32_t words = lanes_per_split / PER_WORD;
    const int64_t total = (int64_t)splits * words;
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    for (int64_t w0 = (int64_t)blockIdx.x * 32; w0 < total; w0 += (int64_t)gridDim.x * 32) {
        const int64_t w = w0 + lane;
        unsigned long long acc[PER_WORD];
#pragma unroll
        for (uint32_t j = 0; j < PER_WORD; ++j) acc[j] = 0ull;
        int split = 0;
        uint32_t wi = 0;
        if (w < total) {
            split = (int)(w / words);
            wi = (uint32_t)(w - (int64_t)split * words);
            const unsigned *src = partials + (size_t)split * words + wi;
#pragma unroll 4
            for (int g = slice; g < groups; g += SLICES) {
                const unsigned v = src[(size_t)g * splits * words];
#pragma unroll
                for (uint32_t j = 0; j < PER_WORD; ++j) acc[j] += LB == 32 ? v : ((v >> ((j * LB) % 32)) & LANE_MASK);
            }
        }
#pragma unroll
        for (uint32_t j = 0; j < PER_WORD; ++j) sub[slice][lane][j] = acc[j];
        __syncthreads();
        if (slice == 0 && w < total) {
            const uint32_t k0 = (uint32_t)split * lanes_per_split + wi * PER_WORD;
#pragma unroll
            for (uint32_t j = 0; j < PER_WORD; ++j) {
                unsigned long long c = 0ull;
#pragma unroll
                for (int sl = 0; sl < SLICES; ++sl) c += sub[sl][lane][j];
                const uint32_t k = k0 + j;
                if (k < sigma_pow_m) {
                    if (scratch != nullptr) c += scratch[k];
                    if (c) bins[k] += c;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ Variant P: partition + shared-memory count
// Tables that do not fit one CTA (n >= 4 at sigma = 21: 4 M / 86 M bins).  L2 REDs cap at ~200 G
// updates/s while the table sits in L2 (n = 4) and drop to ~30 G/s once it does not (n = 5, 686 MB).
// Here the corpus is cut into chunks; per chunk
//   P1  every CTA scans a contiguous range of vectors and histograms its windows by the code of their
//       first two symbols ("bucket", nb = sigma^2 <= 1024),
//   P2  one CTA turns the [CTA][bucket] histogram into exact write cursors (bucket-major) and a work list,
//   P3  the same CTAs rescan their ranges tile by tile (1024 vectors), counting-sort the tile's windows by
//       bucket in shared memory and append the remaining (M-2)-symbol codes to their buckets: runs of one
//       bucket are contiguous in the tile AND continue where the CTA's previous tile stopped, so the
//       4 B/window stores coalesce,
//   P4  persistent CTAs take (bucket, slice) items from the work list and count the slice's codes in a
//       shared-memory table of sub_bins = sigma^(M-2) lanes (32/16/8 bit as above), then add the non-zero
//       lanes to the dense table (one RED per lane and item).
// Traffic: 2 B/residue read + 4 B written + 4 B read; every table update is a shared-memory atomic.
// This is the north star's "radix pass + segmented reduce" with the reduce done by a histogram.
constexpr int kPartThreads = 1024;                         // P2, P4
constexpr int kPartScanThreads = 512;                      // P1, P3: two CTAs per SM, so one's barriers hide under the other's work
constexpr int kPartScanCtasPerSm = 2;
constexpr uint32_t kPartTileVecs = kPartScanThreads;       // one 16-byte vector (16 windows) per thread and tile
constexpr uint32_t kPartSuffixBits = 18;
constexpr uint32_t kPartSuffixMask = (1u << kPartSuffixBits) - 1u;
constexpr uint32_t kPartMaxBuckets = 1024;

struct PartRange {  // the vectors [lo, hi) of this chunk are split evenly over the CTAs of P1 / P3
    int64_t lo, hi;
    __device__ __forceinline__ void mine(int64_t &a, int64_t &b) const {
        const int64_t per = (hi - lo + gridDim.x - 1) / gridDim.x;
        a = lo + (int64_t)blockIdx.x * per;
        b = a + per < hi ? a + per : hi;
    }
};

template <int M>
__global__ void __launch_bounds__(kPartScanThreads, kPartScanCtasPerSm) part_hist_kernel(
    const uint8_t *__restrict__ buf, int64_t nbytes, PartRange range, const uint8_t *__restrict__ rank_of_byte, uint32_t sigma,
    uint32_t sigma_pow_m, uint32_t sigma_pow_n, uint32_t sub_bins, uint32_t nb, uint8_t *__restrict__ short_present,
    uint32_t *__restrict__ hist) {
    __shared__ uint32_t cnt[kPartMaxBuckets];
    __shared__ uint8_t lut[256];
    if (threadIdx.x < 256) lut[threadIdx.x] = rank_of_byte[threadIdx.x];
    for (uint32_t b = threadIdx.x; b < kPartMaxBuckets; b += kPartScanThreads) cnt[b] = 0;
    __syncthreads();
    int64_t lo, hi;
    range.mine(lo, hi);
    int64_t vec = lo + threadIdx.x;
    uint32_t wn[6];
    if (vec < hi) load_vector(buf, nbytes, vec * 16, wn);
    for (; vec < hi; vec += kPartScanThreads) {
        uint32_t wc[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) wc[k] = wn[k];
        if (vec + kPartScanThreads < hi) load_vector(buf, nbytes, (vec + kPartScanThreads) * 16, wn);
        uint32_t codes[16];
        const uint32_t valid = scan_words<M, true>(wc, nbytes, vec * 16, lut, sigma, sigma_pow_m, sigma_pow_n, short_present, codes, sub_bins);
#pragma unroll
        for (int k = 0; k < 16; ++k) atomicAdd(&cnt[codes[k] >> kPartSuffixBits], (valid >> k) & 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += kPartScanThreads) hist[(size_t)blockIdx.x * nb + b] = cnt[b];
}

// block-wide exclusive scan of one value per thread (blockDim.x a multiple of 32, <= 1024); *total gets the sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t x, uint32_t *wsum /*[33] shared*/, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t v = lane < (blockDim.x >> 5) ? wsum[lane] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, d);
            if (lane >= d) s += y;
        }
        wsum[lane] = s - v;
        if (lane == 31) wsum[32] = s;
    }
    __syncthreads();
    const uint32_t r = incl - x + wsum[warp];
    if (total != nullptr) *total = wsum[32];
    __syncthreads();  // wsum may be reused by the caller's next scan
    return r;
}

// ctrl[0] = number of work items, ctrl[1] = next item (P4's fetch counter).
// offs[c][b] = first entry slot of CTA c's bucket-b run; items = (bucket, begin, end) triples, slices of
// `slice` entries; bucket regions start at multiples of 4 entries (16-byte loads in P4).
__global__ void __launch_bounds__(kPartThreads, 1) part_offsets_kernel(const uint32_t *__restrict__ hist, int ctas, uint32_t nb,
                                                                       uint32_t slice, uint32_t *__restrict__ offs,
                                                                       uint32_t *__restrict__ items, uint32_t *__restrict__ ctrl) {
    __shared__ uint32_t wsum[33];
    const uint32_t b = threadIdx.x;
    uint32_t total = 0;
    if (b < nb)
        for (int c = 0; c < ctas; ++c) {
            const uint32_t h = hist[(size_t)c * nb + b];
            offs[(size_t)c * nb + b] = total;
            total += h;
        }
    const uint32_t padded = (total + 3u) & ~3u;
    const uint32_t start = block_excl_scan(padded, wsum, nullptr);
    const uint32_t n_items = (total + slice - 1) / slice;
    uint32_t all_items;
    const uint32_t item0 = block_excl_scan(n_items, wsum, &all_items);
    if (b < nb) {
        for (int c = 0; c < ctas; ++c) offs[(size_t)c * nb + b] += start;
        for (uint32_t j = 0; j < n_items; ++j) {
            uint32_t *it = items + 3 * (size_t)(item0 + j);
            it[0] = b;
            it[1] = start + j * slice;
            it[2] = start + min(total, (j + 1) * slice);
        }
    }
    if (b == 0) {
        ctrl[0] = all_items;
        ctrl[1] = 0;
    }
}

template <int M>
__global__ void __launch_bounds__(kPartScanThreads, kPartScanCtasPerSm) part_scatter_kernel(
    const uint8_t *__restrict__ buf, int64_t nbytes, PartRange range, const uint8_t *__restrict__ rank_of_byte, uint32_t sigma,
    uint32_t sigma_pow_m, uint32_t sigma_pow_n, uint32_t sub_bins, uint32_t nb, const uint32_t *__restrict__ offs,
    uint32_t *__restrict__ entries) {
    extern __shared__ uint32_t staging[];           // kPartTileVecs * 16 entries (dynamic: static + this exceeds 48 KB)
    __shared__ uint32_t cnt[kPartMaxBuckets];       // windows of this tile per bucket
    __shared__ uint32_t base[kPartMaxBuckets];      // first staging slot of the bucket's run
    __shared__ uint32_t gdelta[kPartMaxBuckets];    // entries index = staging slot + gdelta[bucket]  (mod 2^32)
    __shared__ uint32_t cursor[kPartMaxBuckets];    // next free slot of this CTA's run in the bucket
    __shared__ uint32_t wsum[33];
    __shared__ uint8_t lut[256];
    const uint32_t tid = threadIdx.x;
    static_assert(kPartMaxBuckets == 2 * kPartScanThreads, "thread t owns buckets 2t and 2t+1");
    if (tid < 256) lut[tid] = rank_of_byte[tid];
#pragma unroll
    for (uint32_t j = 0; j < 2; ++j) {
        const uint32_t b = 2 * tid + j;
        cursor[b] = b < nb ? offs[(size_t)blockIdx.x * nb + b] : 0u;
        cnt[b] = 0;
    }
    __syncthreads();
    int64_t lo, hi;
    range.mine(lo, hi);
    uint32_t wn[6];
    if (lo + tid < hi) load_vector(buf, nbytes, (lo + tid) * 16, wn);
    for (int64_t tile = lo; tile < hi; tile += kPartTileVecs) {
        const int64_t vec = tile + tid;
        uint32_t codes[16];
        unsigned short slot[16];
        uint32_t valid = 0;
        if (vec < hi) {
            uint32_t wc[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) wc[k] = wn[k];
            if (vec + kPartTileVecs < hi) load_vector(buf, nbytes, (vec + kPartTileVecs) * 16, wn);
            valid = scan_words<M, true>(wc, nbytes, vec * 16, lut, sigma, sigma_pow_m, sigma_pow_n, nullptr, codes, sub_bins);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if ((valid >> k) & 1u) slot[k] = (unsigned short)atomicAdd(&cnt[codes[k] >> kPartSuffixBits], 1u);
        }
        __syncthreads();
        uint32_t total;
        const uint32_t c0 = cnt[2 * tid], c1 = cnt[2 * tid + 1];
        const uint32_t excl = block_excl_scan(c0 + c1, wsum, &total);
        base[2 * tid] = excl;
        base[2 * tid + 1] = excl + c0;
        gdelta[2 * tid] = cursor[2 * tid] - excl;
        gdelta[2 * tid + 1] = cursor[2 * tid + 1] - (excl + c0);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if ((valid >> k) & 1u) staging[base[codes[k] >> kPartSuffixBits] + slot[k]] = codes[k];
        __syncthreads();
        for (uint32_t i = tid; i < total; i += kPartScanThreads) {
            const uint32_t c = staging[i];
            entries[gdelta[c >> kPartSuffixBits] + i] = c & kPartSuffixMask;
        }
        cursor[2 * tid] += c0;
        cursor[2 * tid + 1] += c1;
        cnt[2 * tid] = 0;
        cnt[2 * tid + 1] = 0;
        __syncthreads();
    }
}

// P4.  LB as in the shared-memory variant: 32 = plain adds; 16 = strict drain at 32767; 8 = drain at 127 into `out`
// (a zeroed scratch table) + *status on a lane observed >= 192 (result then discarded, see pg_ngram_count).
template <int LB>
__global__ void __launch_bounds__(kPartThreads, 1) part_count_kernel(const uint32_t *__restrict__ entries,
                                                                     const uint32_t *__restrict__ items, uint32_t *__restrict__ ctrl,
                                                                     uint32_t sub_bins, unsigned long long *__restrict__ out,
                                                                     int *__restrict__ status) {
    constexpr uint32_t PER_WORD = 32 / LB;
    constexpr uint32_t LSH = LB == 32 ? 0 : (LB == 16 ? 1 : 2);
    constexpr uint32_t LANE_MASK = LB == 32 ? 0xFFFFFFFFu : ((1u << (LB % 32)) - 1u);
    constexpr uint32_t HALF = LB == 32 ? 0u : (1u << ((LB - 1) % 32));
    extern __shared__ unsigned tbl[];
    __shared__ uint32_t s_item;
    const uint32_t tid = threadIdx.x;
    const uint32_t words = (sub_bins + PER_WORD - 1) / PER_WORD;
    const uint32_t n_items = ctrl[0];
    bool hazard = false;
    for (;;) {
        if (tid == 0) s_item = atomicAdd(&ctrl[1], 1u);
        for (uint32_t i = tid; i < words; i += kPartThreads) tbl[i] = 0;
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const uint32_t bucket = items[3 * (size_t)item], begin = items[3 * (size_t)item + 1], end = items[3 * (size_t)item + 2];
        unsigned long long *dst = out + (size_t)bucket * sub_bins;
        for (uint32_t b0 = begin; b0 < end; b0 += 16u * kPartThreads) {
            uint32_t e[16];
            uint32_t vm = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t idx = b0 + ((uint32_t)j * kPartThreads + tid) * 4u;
                if (idx + 4u <= end) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(entries + idx);
                    e[4 * j] = q.x; e[4 * j + 1] = q.y; e[4 * j + 2] = q.z; e[4 * j + 3] = q.w;
                    vm |= 0xFu << (4 * j);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const bool in = idx + (uint32_t)t < end;
                        e[4 * j + t] = in ? entries[idx + t] : 0u;
                        vm |= (in ? 1u : 0u) << (4 * j + t);
                    }
                }
            }
            if (LB == 32) {
#pragma unroll
                for (int k = 0; k < 16; ++k) atomicAdd(&tbl[e[k]], (vm >> k) & 1u);
            } else {
                unsigned old[16];
                uint32_t att = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t sh = (e[k] * LB) & 31u;
                    const uint32_t inc = ((vm >> k) & 1u) << sh;
                    old[k] = atomicAdd(&tbl[e[k] >> LSH], inc);
                    att |= (old[k] + inc) & (inc << (LB - 1));
                }
                if (att) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        if (!((vm >> k) & 1u)) continue;
                        const uint32_t sh = (e[k] * LB) & 31u;
                        const uint32_t lane = (old[k] >> sh) & LANE_MASK;
                        if (lane == HALF - 1u) {
                            atomicSub(&tbl[e[k] >> LSH], HALF << sh);
                            atomicAdd(&dst[e[k]], (unsigned long long)HALF);
                        }
                        if (LB == 8 && lane >= 192u) hazard = true;
                    }
                }
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < words; i += kPartThreads) {
            const unsigned v = tbl[i];
            if (v == 0u) continue;
#pragma unroll
            for (uint32_t j = 0; j < PER_WORD; ++j) {
                const unsigned c = LB == 32 ? v : ((v >> ((j * LB) % 32)) & LANE_MASK);
                const uint32_t k = i * PER_WORD + j;
                if (c != 0u && k < sub_bins) atomicAdd(&dst[k], (unsigned long long)c);
            }
        }
        __syncthreads();  // table is re-zeroed at the top; s_item is rewritten by thread 0 only after this barrier
    }
    if (LB == 8 && hazard) *status = 1;
}

// bins[k] += scratch[k] if the 8-bit result was proven exact (*status == 0)
__global__ void __launch_bounds__(256) merge_scratch_kernel(const unsigned long long *__restrict__ scratch, int64_t nbins,
                                                            const int *__restrict__ status, unsigned long long *__restrict__ bins) {
    if (*status != 0) return;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nbins; k += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = scratch[k];
        if (c) bins[k] += c;
    }
}

__global__ void set_flag_kernel(int *flag, int v) { *flag = v; }

// ------------------------------------------------------------------ table -> ids / edges
__global__ void __launch_bounds__(256) mark_present_kernel(const unsigned long long *__restrict__ bins, int64_t nbins,
                                                           uint32_t sigma, uint32_t sigma_pow_n,
                                                           int64_t *__restrict__ present, int64_t *__restrict__ edge_flag) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nbins; k += (int64_t)gridDim.x * blockDim.x) {
        const bool nz = bins[k] != 0ull;
        edge_flag[k] = nz ? 1 : 0;
        if (nz) {
            present[k / sigma] = 1;        // source n-gram  = leading n symbols
            present[k % sigma_pow_n] = 1;  // target n-gram  = trailing n symbols
        }
    }
}

__global__ void __launch_bounds__(256) widen_flags_kernel(const uint8_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i] ? 1 : 0;
}

__global__ void __launch_bounds__(256) emit_nodes_kernel(const int64_t *__restrict__ present, const int64_t *__restrict__ node_id,
                                                         int64_t ngrams, int64_t *__restrict__ node_code) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngrams; g += (int64_t)gridDim.x * blockDim.x)
        if (present[g]) node_code[node_id[g]] = g;
}

__global__ void __launch_bounds__(256) emit_edges_kernel(const unsigned long long *__restrict__ bins, int64_t nbins,
                                                         const int64_t *__restrict__ edge_off, const int64_t *__restrict__ node_id,
                                                         uint32_t sigma, uint32_t sigma_pow_n, int64_t *__restrict__ src,
                                                         int64_t *__restrict__ dst, int64_t *__restrict__ count) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nbins; k += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = bins[k];
        if (c) {
            const int64_t e = edge_off[k];
            src[e] = node_id[k / sigma];
            dst[e] = node_id[k % sigma_pow_n];
            count[e] = (int64_t)c;
        }
    }
}

inline unsigned grid_for(int64_t n, int threads = 256, int per_sm = 8) {
    int64_t want = pg_ceil_div(n, threads);
    int64_t cap = (int64_t)PG_NUM_SMS * per_sm;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

bool table_sizes(int n, int sigma, int64_t *pow_n, int64_t *pow_m) {
    if (n < 1 || n > 6 || sigma < 1 || sigma > 255) return false;
    int64_t p = 1;
    for (int k = 0; k < n; ++k) p *= sigma;
    *pow_n = p;
    *pow_m = p * sigma;
    return *pow_m <= (1ll << 32);
}

struct ExtractWs {
    int64_t *present, *node_id, *edge_off;
    void *scan_ws;
    size_t scan_bytes;
};

bool carve_extract(void *d_ws, size_t ws_bytes, int64_t pow_n, int64_t pow_m, ExtractWs *w) {
    PgArena a(d_ws, ws_bytes);
    w->present = a.take<int64_t>((size_t)pow_n);
    w->node_id = a.take<int64_t>((size_t)pow_n);
    w->edge_off = a.take<int64_t>((size_t)pow_m);
    w->scan_bytes = pg_scan_ws_bytes(pow_m);
    w->scan_ws = a.take<char>(w->scan_bytes);
    return a.ok;
}
}  // namespace

extern "C" int pg_byte_presence(const uint8_t *d_buf, int64_t nbytes, uint32_t *d_present256, pg_stream_t stream) {
    PG_CHECK_ARG(d_buf && d_present256 && nbytes >= 0, "pg_byte_presence: bad arguments");
    PG_CHECK_ARG(((uintptr_t)d_buf & 15) == 0, "pg_byte_presence: d_buf must be 16-byte aligned");
    if (nbytes == 0) return PG_OK;
    byte_presence_kernel<<<grid_for(nbytes / 16 + 1), 256, 0, pg_cu(stream)>>>(d_buf, nbytes, d_present256);
    PG_CUDA_LAUNCH_CHECK("byte_presence_kernel");
    return PG_OK;
}

extern "C" int pg_unpack5(const uint8_t *d_packed, int64_t n_symbols, uint8_t *d_out, pg_stream_t stream) {
    PG_CHECK_ARG(n_symbols >= 0, "pg_unpack5: bad size");
    if (n_symbols == 0) return PG_OK;
    PG_CHECK_ARG(d_packed && d_out && ((uintptr_t)d_out & 15) == 0, "pg_unpack5: null buffer or output not 16-byte aligned");
    unpack5_kernel<<<grid_for((n_symbols + 15) / 16, 256, 16), 256, 0, pg_cu(stream)>>>(d_packed, n_symbols, d_out);
    PG_CUDA_LAUNCH_CHECK("unpack5_kernel");
    return PG_OK;
}

extern "C" int pg_synth_corpus(uint8_t *d_buf, int64_t first_seq, int64_t nseq, int seq_len, uint32_t seed,
                               int leading_space, pg_stream_t stream) {
    PG_CHECK_ARG(d_buf && nseq >= 0 && seq_len >= 1 && first_seq >= 0, "pg_synth_corpus: bad arguments");
    if (nseq == 0) return PG_OK;
    synth_corpus_kernel<<<grid_for(nseq * (seq_len + 2), 256, 16), 256, 0, pg_cu(stream)>>>(
        d_buf, first_seq, nseq, seq_len, seed, leading_space ? 1 : 0);
    PG_CUDA_LAUNCH_CHECK("synth_corpus_kernel");
    return PG_OK;
}

// test hook: pin the count variant so the parity tests cover every kernel (see pgb200.h)
static int g_count_variant = PG_COUNT_AUTO;
extern "C" void pg_debug_count_variant(int v) { g_count_variant = v; }

namespace {
struct CountArgs {
    const uint8_t *buf;
    int64_t nbytes;
    const uint8_t *rank;
    uint32_t sigma, pow_m, pow_n;
    uint8_t *short_present;
    cudaStream_t st;
};

// Workspace of the shared-memory variants: [status word, 256 B][drain scratch: sigma^(n+1) x u64, 8-bit
// variant only][one packed table per CTA: 148 x <= 224 KB].
constexpr size_t kPartialBytes = (size_t)PG_NUM_SMS * kSmemTableBytes;
constexpr int kMaxSplits = 4;
inline int splits_for(int64_t pow_m, int lane_bits) { return (int)pg_ceil_div(pow_m * lane_bits / 8, kSmemTableBytes); }
inline bool uses_fast8(int64_t pow_m) { return splits_for(pow_m, 16) > 1 && splits_for(pow_m, 8) <= kMaxSplits; }
inline size_t scratch_bytes(int64_t pow_m) { return uses_fast8(pow_m) ? pg_align_up((size_t)pow_m * 8, 256) : 0; }

struct CountWs {
    int *status;
    unsigned long long *scratch;  // null unless the 8-bit variant applies
    unsigned *partials;
};

// count into per-CTA packed tables + reduce them into `bins`.  spill = where drained lanes go (bins, or
// the scratch for the 8-bit variant); run_if / gate as in reduce_partials_kernel.
template <int LB, bool SPLIT>
int launch_smem_count(int m, const CountArgs &a, unsigned long long *bins, unsigned long long *spill, int splits, const CountWs &ws,
                      int run_if) {
    const uint32_t per_word = 32 / LB;
    uint32_t lanes = (uint32_t)pg_ceil_div(a.pow_m, splits);
    lanes = (lanes + per_word - 1) / per_word * per_word;  // whole words per split
    const uint32_t words = lanes / per_word;
    const size_t smem = (size_t)words * sizeof(unsigned);
    const int64_t nvec = pg_ceil_div(a.nbytes, 16);
    int64_t groups = PG_NUM_SMS / splits;
    const int64_t max_groups = pg_ceil_div(nvec, kSmemCountThreads);
    if (groups > max_groups) groups = max_groups;
    const unsigned grid = (unsigned)(groups * splits);
    const int *gate = run_if == 1 ? ws.status : nullptr;
    int *status = LB == 8 ? ws.status : nullptr;
#define PG_LAUNCH_SMEM(MM)                                                                                                  \
    do {                                                                                                                    \
        PG_CUDA_CALL(cudaFuncSetAttribute(ngram_count_smem_kernel<MM, LB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem));                                                                      \
        ngram_count_smem_kernel<MM, LB, SPLIT><<<grid, kSmemCountThreads, smem, a.st>>>(                                     \
            a.buf, a.nbytes, a.rank, a.sigma, a.pow_m, a.pow_n, spill, a.short_present, splits, lanes, status, gate,         \
            ws.partials);                                                                                                    \
    } while (0)
    switch (m) {
        case 2: PG_LAUNCH_SMEM(2); break;
        case 3: PG_LAUNCH_SMEM(3); break;
        case 4: PG_LAUNCH_SMEM(4); break;
        case 5: PG_LAUNCH_SMEM(5); break;
        case 6: PG_LAUNCH_SMEM(6); break;
        case 7: PG_LAUNCH_SMEM(7); break;
        default: pg_set_error("pg_ngram_count: unsupported n=%d", m - 1); return PG_EINVAL;
    }
#undef PG_LAUNCH_SMEM
    PG_CUDA_LAUNCH_CHECK("ngram_count_smem_kernel");
    const int64_t total_words = (int64_t)splits * words;
    reduce_partials_kernel<LB><<<grid_for(total_words * 8, 256, 16), 256, 0, a.st>>>(
        ws.partials, (int)groups, splits, lanes, a.pow_m, LB == 8 ? ws.scratch : nullptr, ws.status, run_if, bins);
    PG_CUDA_LAUNCH_CHECK("reduce_partials_kernel");
    return PG_OK;
}

int launch_global_count(int m, const CountArgs &a, unsigned long long *bins, const int *gate) {
    const unsigned grid = grid_for(pg_ceil_div(a.nbytes, 16), 256, 8);   // grid-stride: a gated (no-op) launch costs ~1200 CTAs, not one per 4 KB
#define PG_LAUNCH_COUNT(MM) \
    ngram_count_kernel<MM><<<grid, 256, 0, a.st>>>(a.buf, a.nbytes, a.rank, a.sigma, a.pow_m, a.pow_n, bins, a.short_present, gate)
    switch (m) {
        case 2: PG_LAUNCH_COUNT(2); break;
        case 3: PG_LAUNCH_COUNT(3); break;
        case 4: PG_LAUNCH_COUNT(4); break;
        case 5: PG_LAUNCH_COUNT(5); break;
        case 6: PG_LAUNCH_COUNT(6); break;
        case 7: PG_LAUNCH_COUNT(7); break;
        default: pg_set_error("pg_ngram_count: unsupported n=%d", m - 1); return PG_EINVAL;
    }
#undef PG_LAUNCH_COUNT
    PG_CUDA_LAUNCH_CHECK("ngram_count_kernel");
    return PG_OK;
}

// the strictly exact variants: 32-bit lanes, 16-bit lanes (<= 4 key-range splits), else L2 REDs.
// run_if = -1: unconditional; 1: only if *ws.status != 0 (recount after a flagged hazard).
int launch_strict_count(int m, const CountArgs &a, unsigned long long *bins, const CountWs *ws, int run_if, bool allow_smem) {
    allow_smem = allow_smem && ws != nullptr;
    if (allow_smem && splits_for(a.pow_m, 32) == 1) return launch_smem_count<32, false>(m, a, bins, bins, 1, *ws, run_if);
    const int s16 = splits_for(a.pow_m, 16);
    if (allow_smem && s16 == 1) return launch_smem_count<16, false>(m, a, bins, bins, 1, *ws, run_if);
    if (allow_smem && s16 <= kMaxSplits) return launch_smem_count<16, true>(m, a, bins, bins, s16, *ws, run_if);
    return launch_global_count(m, a, bins, run_if == 1 ? ws->status : nullptr);
}

// ---- variant P (partition + shared-memory count)
constexpr int64_t kPartMinChunk = 16 * (int64_t)kPartTileVecs;   // one tile of windows
constexpr int64_t kPartMaxChunk = 1ll << 30;                     // entry indices stay 32-bit
constexpr int64_t kPartAutoMinBytes = 32ll << 20;                // AUTO: smaller corpora stay on the L2 REDs

struct PartPlan {
    bool ok;
    int lane_bits;
    uint32_t nb, sub_bins;
    size_t fixed_bytes;  // everything but the entries
    size_t scratch_off, hist_off, offs_off, items_off, entries_off;
    int64_t max_items_cap;
};

// lane_bits = widest of 32/16/8 whose sub-table (sigma^(m-2) lanes) fits one CTA; ok = false if none does
PartPlan part_plan(int m, int sigma, int64_t pow_m) {
    PartPlan p{};
    if (m < 5 || sigma * sigma > (int)kPartMaxBuckets) return p;
    p.nb = (uint32_t)(sigma * sigma);
    const int64_t sub = pow_m / p.nb;
    if (sub > (int64_t)kPartSuffixMask + 1) return p;
    p.sub_bins = (uint32_t)sub;
    p.lane_bits = sub * 4 <= kSmemTableBytes ? 32 : (sub * 2 <= kSmemTableBytes ? 16 : (sub <= kSmemTableBytes ? 8 : 0));
    if (p.lane_bits == 0) return p;
    p.max_items_cap = kPartMaxChunk / (64 * kPartMinChunk) + p.nb + 8;   // slices are never smaller than 64 tiles at the largest chunk
    size_t off = 512;  // [status 256 B][ctrl 256 B]
    p.scratch_off = off;
    if (p.lane_bits == 8) off += pg_align_up((size_t)pow_m * 8, 256);
    p.hist_off = off;
    off += pg_align_up((size_t)PG_NUM_SMS * kPartScanCtasPerSm * p.nb * 4, 256);
    p.offs_off = off;
    off += pg_align_up((size_t)PG_NUM_SMS * kPartScanCtasPerSm * p.nb * 4, 256);
    p.items_off = off;
    off += pg_align_up((size_t)p.max_items_cap * 12, 256);
    p.entries_off = off;
    p.fixed_bytes = off;
    p.ok = true;
    return p;
}

inline size_t part_entry_bytes(const PartPlan &p, int64_t chunk_windows) { return pg_align_up((size_t)(chunk_windows + 4 * p.nb + 64) * 4, 256); }

// windows per chunk the workspace allows (0 = variant P cannot run with this workspace)
int64_t part_chunk_windows(const PartPlan &p, size_t ws_bytes, int64_t nbytes) {
    if (!p.ok || ws_bytes <= p.fixed_bytes + part_entry_bytes(p, kPartMinChunk)) return 0;
    int64_t cap = (int64_t)((ws_bytes - p.fixed_bytes) / 4) - 4 * p.nb - 128;
    if (cap > kPartMaxChunk) cap = kPartMaxChunk;
    const int64_t need = pg_ceil_div(nbytes, kPartMinChunk) * kPartMinChunk;
    if (cap > need) cap = need;
    return cap / kPartMinChunk * kPartMinChunk;
}

int launch_partitioned_count(int m, const CountArgs &a, unsigned long long *bins, void *d_ws, const PartPlan &p, int64_t chunk_windows) {
    char *ws = (char *)d_ws;
    int *status = (int *)ws;
    uint32_t *ctrl = (uint32_t *)(ws + 256);
    unsigned long long *scratch = p.lane_bits == 8 ? (unsigned long long *)(ws + p.scratch_off) : nullptr;
    uint32_t *hist = (uint32_t *)(ws + p.hist_off), *offs = (uint32_t *)(ws + p.offs_off), *items = (uint32_t *)(ws + p.items_off);
    uint32_t *entries = (uint32_t *)(ws + p.entries_off);
    unsigned long long *out = p.lane_bits == 8 ? scratch : bins;
    PG_CUDA_CALL(cudaMemsetAsync(ws, 0, p.lane_bits == 8 ? p.hist_off : 512, a.st));   // status, ctrl (+ scratch)
    if (g_count_variant == PG_COUNT_PARTITIONED_FORCE_HAZARD) {
        set_flag_kernel<<<1, 1, 0, a.st>>>(status, 1);
        PG_CUDA_LAUNCH_CHECK("set_flag_kernel");
    }
    const int64_t nvec = pg_ceil_div(a.nbytes, 16);
    const int64_t chunk_vecs = chunk_windows / 16;
    // slices: large enough that zero + flush of the sub-table (sub_bins lanes) stays a small part of an item
    int64_t slice = pg_ceil_div(chunk_windows, 256 * kPartMinChunk) * kPartMinChunk;
    if (slice < 4 * kPartMinChunk) slice = 4 * kPartMinChunk;
    if (slice > 64 * kPartMinChunk * 8) slice = 64 * kPartMinChunk * 8;   // 4 M entries
    if (chunk_windows / slice + p.nb + 1 > p.max_items_cap) {
        pg_set_error("pg_ngram_count: internal error (work list capacity)");
        return PG_EINVAL;
    }
    const size_t stage_bytes = (size_t)kPartTileVecs * 16 * 4;
    const uint32_t words = (p.sub_bins + 32 / p.lane_bits - 1) / (32 / p.lane_bits);
    const size_t tbl_bytes = (size_t)words * 4;
    for (int64_t v0 = 0; v0 < nvec; v0 += chunk_vecs) {
        PartRange r{v0, v0 + chunk_vecs < nvec ? v0 + chunk_vecs : nvec};
        int64_t ctas = pg_ceil_div(r.hi - r.lo, kPartTileVecs);
        if (ctas > PG_NUM_SMS * kPartScanCtasPerSm) ctas = PG_NUM_SMS * kPartScanCtasPerSm;
#define PG_PART_SCAN(MM)                                                                                                          \
    do {                                                                                                                          \
        part_hist_kernel<MM><<<(unsigned)ctas, kPartScanThreads, 0, a.st>>>(a.buf, a.nbytes, r, a.rank, a.sigma, a.pow_m, a.pow_n, \
                                                                       p.sub_bins, p.nb, a.short_present, hist);                  \
        PG_CUDA_LAUNCH_CHECK("part_hist_kernel");                                                                                 \
        part_offsets_kernel<<<1, kPartThreads, 0, a.st>>>(hist, (int)ctas, p.nb, (uint32_t)slice, offs, items, ctrl);              \
        PG_CUDA_LAUNCH_CHECK("part_offsets_kernel");                                                                              \
        PG_CUDA_CALL(cudaFuncSetAttribute(part_scatter_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes)); \
        part_scatter_kernel<MM><<<(unsigned)ctas, kPartScanThreads, stage_bytes, a.st>>>(a.buf, a.nbytes, r, a.rank, a.sigma,     \
                                                                                        a.pow_m, a.pow_n, p.sub_bins, p.nb, offs, \
                                                                                        entries);                                 \
        PG_CUDA_LAUNCH_CHECK("part_scatter_kernel");                                                                              \
    } while (0)
        switch (m) {
            case 5: PG_PART_SCAN(5); break;
            case 6: PG_PART_SCAN(6); break;
            case 7: PG_PART_SCAN(7); break;
            default: pg_set_error("pg_ngram_count: variant P needs n >= 4"); return PG_EINVAL;
        }
#undef PG_PART_SCAN
        const unsigned grid4 = (unsigned)(PG_NUM_SMS * (tbl_bytes <= 96 * 1024 ? 2 : 1));
#define PG_PART_COUNT(LL)                                                                                                     \
    do {                                                                                                                      \
        PG_CUDA_CALL(cudaFuncSetAttribute(part_count_kernel<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tbl_bytes)); \
        part_count_kernel<LL><<<grid4, kPartThreads, tbl_bytes, a.st>>>(entries, items, ctrl, p.sub_bins, out, status);         \
    } while (0)
        if (p.lane_bits == 32) PG_PART_COUNT(32);
        else if (p.lane_bits == 16) PG_PART_COUNT(16);
        else PG_PART_COUNT(8);
#undef PG_PART_COUNT
        PG_CUDA_LAUNCH_CHECK("part_count_kernel");
    }
    if (p.lane_bits == 8) {
        merge_scratch_kernel<<<grid_for(a.pow_m), 256, 0, a.st>>>(scratch, (int64_t)a.pow_m, status, bins);
        PG_CUDA_LAUNCH_CHECK("merge_scratch_kernel");
        CountArgs strict = a;
        strict.short_present = a.short_present;   // idempotent flags
        return launch_global_count(m, strict, bins, status);   // no-op unless a hazard was flagged
    }
    return PG_OK;
}
}  // namespace

extern "C" size_t pg_ngram_count_ws_bytes(int n, int sigma) {
    int64_t pow_n, pow_m;
    if (!table_sizes(n, sigma, &pow_n, &pow_m)) return 0;
    if (splits_for(pow_m, 8) > kMaxSplits) return 256;  // L2 REDs only: no workspace needed
    return 256 + scratch_bytes(pow_m) + kPartialBytes;
}

extern "C" size_t pg_ngram_count_ws_bytes_for(int n, int sigma, int64_t nbytes) {
    int64_t pow_n, pow_m;
    if (!table_sizes(n, sigma, &pow_n, &pow_m)) return 0;
    const size_t base = pg_ngram_count_ws_bytes(n, sigma);
    const PartPlan p = part_plan(n + 1, sigma, pow_m);
    if (!p.ok || splits_for(pow_m, 8) <= kMaxSplits || nbytes <= 0) return base;
    int64_t chunk = pg_ceil_div(nbytes, kPartMinChunk) * kPartMinChunk;
    if (chunk > kPartMaxChunk) chunk = kPartMaxChunk;
    const size_t want = p.fixed_bytes + part_entry_bytes(p, chunk) + 1024;
    return want > base ? want : base;
}

extern "C" int pg_ngram_count(const uint8_t *d_buf, int64_t nbytes, int n, const uint8_t *d_rank_of_byte, int sigma,
                              unsigned long long *d_bins, uint8_t *d_short_present, void *d_ws, size_t ws_bytes,
                              pg_stream_t stream) {
    int64_t pow_n, pow_m;
    PG_CHECK_ARG(d_buf && d_rank_of_byte && d_bins && d_short_present && nbytes >= 0, "pg_ngram_count: null/negative argument");
    PG_CHECK_ARG(((uintptr_t)d_buf & 15) == 0, "pg_ngram_count: d_buf must be 16-byte aligned");
    PG_CHECK_ARG(nbytes < (1ll << 40), "pg_ngram_count: at most 2^40 bytes per call (count the corpus in chunks)");
    if (!table_sizes(n, sigma, &pow_n, &pow_m)) {
        pg_set_error("pg_ngram_count: table sigma^(n+1) out of range (n=%d sigma=%d; need 1<=n<=6, sigma^(n+1)<=2^32)", n, sigma);
        return PG_ERANGE;
    }
    if (nbytes == 0) return PG_OK;
    const int64_t nvec = pg_ceil_div(nbytes, 16);
    CountArgs a{d_buf, nbytes, d_rank_of_byte, (uint32_t)sigma, (uint32_t)pow_m, (uint32_t)pow_n, d_short_present, pg_cu(stream)};
    const int m = n + 1;
    const int variant = g_count_variant;
    // tiny corpora: zeroing and reducing 148 tables costs more than the L2 REDs
    const bool big_enough = nvec >= 4 * kSmemCountThreads;
    const bool ws_usable = d_ws != nullptr && ((uintptr_t)d_ws & 255) == 0;
    const bool part_forced = variant == PG_COUNT_PARTITIONED || variant == PG_COUNT_PARTITIONED_FORCE_HAZARD;
    if (ws_usable && splits_for(pow_m, 8) > kMaxSplits && (part_forced || (variant == PG_COUNT_AUTO && nbytes >= kPartAutoMinBytes))) {
        const PartPlan p = part_plan(m, sigma, pow_m);
        const int64_t chunk = part_chunk_windows(p, ws_bytes, nbytes);
        // AUTO wants chunks of >= 1/16 of the corpus or 64 M windows (the dense-table flush per chunk and bucket must amortise)
        if (chunk > 0 && (part_forced || chunk >= (nbytes < (64ll << 20) * 16 ? nbytes / 16 : (64ll << 20))))
            return launch_partitioned_count(m, a, d_bins, d_ws, p, chunk);
    }
    const bool have_ws = ws_usable && ws_bytes >= pg_ngram_count_ws_bytes(n, sigma) && splits_for(pow_m, 8) <= kMaxSplits;
    if (variant == PG_COUNT_GLOBAL || part_forced || !have_ws) return launch_global_count(m, a, d_bins, nullptr);
    CountWs ws;
    ws.status = (int *)d_ws;
    ws.scratch = uses_fast8(pow_m) ? (unsigned long long *)((char *)d_ws + 256) : nullptr;
    ws.partials = (unsigned *)((char *)d_ws + 256 + scratch_bytes(pow_m));
    const bool fast8 = uses_fast8(pow_m) && variant != PG_COUNT_STRICT;
    if (!fast8) return launch_strict_count(m, a, d_bins, &ws, -1, variant == PG_COUNT_STRICT || big_enough);
    if (variant == PG_COUNT_AUTO && !big_enough) return launch_global_count(m, a, d_bins, nullptr);
    // 8-bit lanes: drained counts go to the zeroed scratch; the reduce merges tables + scratch into d_bins
    // only if no hazard was observed, else the gated strict variant recounts into d_bins
    PG_CUDA_CALL(cudaMemsetAsync(d_ws, 0, 256 + scratch_bytes(pow_m), a.st));
    if (variant == PG_COUNT_FAST8_FORCE_HAZARD) {
        set_flag_kernel<<<1, 1, 0, a.st>>>(ws.status, 1);  // the count below can only set it as well
        PG_CUDA_LAUNCH_CHECK("set_flag_kernel");
    }
    const int s8 = splits_for(pow_m, 8);
    int rc = s8 == 1 ? launch_smem_count<8, false>(m, a, d_bins, ws.scratch, 1, ws, 0)
                     : launch_smem_count<8, true>(m, a, d_bins, ws.scratch, s8, ws, 0);
    if (rc != PG_OK) return rc;
    return launch_strict_count(m, a, d_bins, &ws, 1, true);  // no-op kernels unless *status != 0
}

extern "C" size_t pg_graph_extract_ws_bytes(int n, int sigma) {
    int64_t pow_n, pow_m;
    if (!table_sizes(n, sigma, &pow_n, &pow_m)) return 0;
    return pg_align_up((size_t)pow_n * 8, 256) * 2 + pg_align_up((size_t)pow_m * 8, 256) +
           pg_align_up(pg_scan_ws_bytes(pow_m), 256) + 1024;
}

extern "C" int pg_graph_extract_sizes(const unsigned long long *d_bins, const uint8_t *d_short_present, int n, int sigma,
                                      int64_t *d_sizes, void *d_ws, size_t ws_bytes, pg_stream_t stream) {
    int64_t pow_n, pow_m;
    PG_CHECK_ARG(d_bins && d_short_present && d_sizes && d_ws, "pg_graph_extract_sizes: null argument");
    if (!table_sizes(n, sigma, &pow_n, &pow_m)) {
        pg_set_error("pg_graph_extract_sizes: table out of range (n=%d sigma=%d)", n, sigma);
        return PG_ERANGE;
    }
    ExtractWs w;
    if (!carve_extract(d_ws, ws_bytes, pow_n, pow_m, &w)) {
        pg_set_error("pg_graph_extract_sizes: workspace too small (%zu < %zu)", ws_bytes, pg_graph_extract_ws_bytes(n, sigma));
        return PG_EWORKSPACE;
    }
    cudaStream_t st = pg_cu(stream);
    widen_flags_kernel<<<grid_for(pow_n), 256, 0, st>>>(d_short_present, pow_n, w.present);
    PG_CUDA_LAUNCH_CHECK("widen_flags_kernel");
    mark_present_kernel<<<grid_for(pow_m), 256, 0, st>>>(d_bins, pow_m, (uint32_t)sigma, (uint32_t)pow_n, w.present, w.edge_off);
    PG_CUDA_LAUNCH_CHECK("mark_present_kernel");
    int rc = pg_exclusive_scan_i64(w.present, w.node_id, pow_n, d_sizes + 0, w.scan_ws, w.scan_bytes, st);
    if (rc != PG_OK) return rc;
    rc = pg_exclusive_scan_i64(w.edge_off, w.edge_off, pow_m, d_sizes + 1, w.scan_ws, w.scan_bytes, st);
    return rc;
}

extern "C" int pg_graph_extract_fill(const unsigned long long *d_bins, int n, int sigma, int64_t num_nodes, int64_t num_edges,
                                     int64_t *d_node_code, int64_t *d_src, int64_t *d_dst, int64_t *d_count, void *d_ws,
                                     size_t ws_bytes, pg_stream_t stream) {
    int64_t pow_n, pow_m;
    PG_CHECK_ARG(d_bins && d_ws && num_nodes >= 0 && num_edges >= 0, "pg_graph_extract_fill: bad argument");
    PG_CHECK_ARG(num_nodes == 0 || d_node_code, "pg_graph_extract_fill: null node output");
    PG_CHECK_ARG(num_edges == 0 || (d_src && d_dst && d_count), "pg_graph_extract_fill: null edge output");
    if (!table_sizes(n, sigma, &pow_n, &pow_m)) {
        pg_set_error("pg_graph_extract_fill: table out of range (n=%d sigma=%d)", n, sigma);
        return PG_ERANGE;
    }
    ExtractWs w;
    if (!carve_extract(d_ws, ws_bytes, pow_n, pow_m, &w)) {
        pg_set_error("pg_graph_extract_fill: workspace too small");
        return PG_EWORKSPACE;
    }
    cudaStream_t st = pg_cu(stream);
    if (num_nodes > 0) {
        emit_nodes_kernel<<<grid_for(pow_n), 256, 0, st>>>(w.present, w.node_id, pow_n, d_node_code);
        PG_CUDA_LAUNCH_CHECK("emit_nodes_kernel");
    }
    if (num_edges > 0) {
        emit_edges_kernel<<<grid_for(pow_m), 256, 0, st>>>(d_bins, pow_m, w.edge_off, w.node_id, (uint32_t)sigma,
                                                            (uint32_t)pow_n, d_src, d_dst, d_count);
        PG_CUDA_LAUNCH_CHECK("emit_edges_kernel");
    }
    return PG_OK;
}
