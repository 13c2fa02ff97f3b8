// Row f3 of SURVEY.md section 8: FASTA text -> corpus buffer, natively (host code only).
//
// Replaces the Python line loop of reference src/utils/data_utils.py:182-213 (DataLoader.parse_sequences)
// fused with the padding rule of src/pipeline/data_builder.py:29-35,97-102: the file is mmap'ed and
// every record goes straight into the byte layout the count kernel reads,
//     [' ' before global sequence #0] SEQUENCE ' ' 0xFF
// so no Python string is ever created (UniRef50: 17.5 GB of residues).  Record rules, restated from the
// reference (the parity tests compare against its Python twin, host/data_utils.py):
//   * lines end at \n, \r\n or a lone \r (Python text mode, universal newlines); each line is stripped of
//     leading / trailing white space (str.strip(): \t \n \v \f \r, \x1c-\x1f, space); blank lines are skipped
//   * a line starting with '>' opens a record and closes the previous one; the previous one is emitted only
//     if it collected at least one sequence line (`if protein_id and sequence_parts`)
//   * a header that is exactly ">" makes the reference raise inside `header.split()[0]`; its generator
//     prints and STOPS there (everything before it was already yielded) -- so does this parser
//   * text before the first header is ignored; sequence lines are upper-cased, interior bytes kept as is
//   * bytes >= 0x80 inside sequence text are refused (PG_FASTA_ENONASCII): the reference windows per code
//     point, the byte kernels cannot, and the Python path already refuses them
// Multi-GPU: sequences are dealt to ranks in blocks of `block` records, like host/corpus.py:stream_chunks.
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <new>

#include "common.cuh"

struct pg_fasta_reader {
    const uint8_t *data;
    int64_t size;
    int fd;
    int64_t pos;          // next unread byte: the start of a header line, or of text that belongs to no record
    int64_t records;      // records emitted so far (all ranks) = global index of the next one
    bool stopped;         // the reference's generator died on a bare ">" header
    bool done;
};

namespace {
inline bool is_space(uint8_t c) { return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f); }

// [b, e) = the next line without its terminator, stripped; returns the position after the terminator.
// memchr finds the '\n'; a '\r' before it (CRLF, or a lone CR that ends a line by itself) is looked for inside that span.
inline int64_t next_line(const uint8_t *d, int64_t pos, int64_t size, int64_t *b, int64_t *e) {
    const uint8_t *nl = (const uint8_t *)memchr(d + pos, '\n', (size_t)(size - pos));
    int64_t end = nl ? (int64_t)(nl - d) : size;
    const uint8_t *cr = (const uint8_t *)memchr(d + pos, '\r', (size_t)(end - pos));
    int64_t next;
    if (cr) {
        end = (int64_t)(cr - d);
        next = (end + 1 < size && d[end + 1] == '\n') ? end + 2 : end + 1;
    } else {
        next = end < size ? end + 1 : size;
    }
    int64_t lo = pos, hi = end;
    while (lo < hi && is_space(d[lo])) ++lo;
    while (hi > lo && is_space(d[hi - 1])) --hi;
    *b = lo;
    *e = hi;
    return next;
}

// dst[i] = upper(src[i]); returns the OR of all bytes (top bit set = some byte was not ASCII).  Branch-free: vectorises.
inline uint8_t copy_upper(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int64_t n) {
    uint8_t acc = 0;
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t c = src[i];
        acc |= c;
        dst[i] = (uint8_t)(c - (((uint8_t)(c - 'a') < 26u) << 5));
    }
    return acc;
}
inline uint8_t or_bytes(const uint8_t *src, int64_t n) {
    uint8_t acc = 0;
    for (int64_t i = 0; i < n; ++i) acc |= src[i];
    return acc;
}
}  // namespace

extern "C" pg_fasta_reader *pg_fasta_open(const char *path) {
    if (path == nullptr) {
        pg_set_error("pg_fasta_open: null path");
        return nullptr;
    }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        pg_set_error("pg_fasta_open: cannot open '%s'", path);
        return nullptr;
    }
    struct stat st;
    if (fstat(fd, &st) != 0) {
        close(fd);
        pg_set_error("pg_fasta_open: cannot stat '%s'", path);
        return nullptr;
    }
    const uint8_t *data = nullptr;
    if (st.st_size > 0) {
        void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) {
            close(fd);
            pg_set_error("pg_fasta_open: mmap of '%s' failed", path);
            return nullptr;
        }
        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
        data = (const uint8_t *)m;
    }
    pg_fasta_reader *r = new (std::nothrow) pg_fasta_reader{data, (int64_t)st.st_size, fd, 0, 0, false, st.st_size == 0};
    if (r == nullptr) {
        if (data) munmap((void *)data, (size_t)st.st_size);
        close(fd);
        pg_set_error("pg_fasta_open: out of memory");
    }
    return r;
}

extern "C" void pg_fasta_close(pg_fasta_reader *r) {
    if (r == nullptr) return;
    if (r->data) munmap((void *)r->data, (size_t)r->size);
    close(r->fd);
    delete r;
}

extern "C" int64_t pg_fasta_file_bytes(const pg_fasta_reader *r) { return r ? r->size : -1; }
extern "C" int64_t pg_fasta_records(const pg_fasta_reader *r) { return r ? r->records : -1; }
extern "C" int pg_fasta_stopped_early(const pg_fasta_reader *r) { return r && r->stopped ? 1 : 0; }

// Packs whole records into out[0, cap) until the next one would not fit or the file ends.
// Returns the bytes written (0 = end of file), PG_FASTA_ETOOSMALL if not even one record fits, PG_FASTA_ENONASCII.
extern "C" int64_t pg_fasta_next_chunk(pg_fasta_reader *r, uint8_t *out, int64_t cap, int rank, int world, int block) {
    if (r == nullptr || out == nullptr || cap < 4 || world < 1 || rank < 0 || rank >= world || block < 1) {
        pg_set_error("pg_fasta_next_chunk: bad arguments");
        return PG_EINVAL;
    }
    const uint8_t *d = r->data;
    int64_t w = 0;             // bytes written and committed (whole records)
    bool open_rec = false;            // a header has been consumed and its record is being collected
    int64_t rec_start_pos = r->pos;   // file position of that header: where the next call resumes if the record does not fit
    int64_t rec_w = w;                // where the open record's bytes start in `out`
    int64_t rec_len = 0;              // sequence bytes of the open record so far
    auto mine = [&](int64_t idx) { return (idx / block) % world == rank; };
    auto begin_record = [&](int64_t pos_of_header) {
        open_rec = true;
        rec_start_pos = pos_of_header;
        rec_w = w;
        rec_len = 0;
    };
    // returns false if the record did not fit
    auto finish_record = [&]() -> bool {
        if (open_rec && rec_len > 0) {
            if (mine(r->records)) {
                if (rec_w + (r->records == 0 ? 1 : 0) + rec_len + 2 > cap) return false;
                out[rec_w + (r->records == 0 ? 1 : 0) + rec_len] = ' ';
                out[rec_w + (r->records == 0 ? 1 : 0) + rec_len + 1] = PG_SEP;
                w = rec_w + (r->records == 0 ? 1 : 0) + rec_len + 2;
            }
            r->records += 1;
        }
        open_rec = false;
        return true;
    };
    if (r->done || r->stopped) return 0;
    int64_t pos = r->pos;
    while (pos < r->size) {
        int64_t b, e;
        const int64_t line_pos = pos;
        const int64_t next = next_line(d, pos, r->size, &b, &e);
        if (b == e) {
            pos = next;
            continue;
        }
        if (d[b] == '>') {
            if (!finish_record()) goto no_room;
            // the committed prefix is safe; from here on a failure must come back to THIS header
            if (e - b == 1) {  // bare ">": the reference's generator raises and stops here
                r->stopped = true;
                r->pos = r->size;
                return w;
            }
            begin_record(line_pos);
            pos = next;
            continue;
        }
        if (open_rec) {
            const bool take = mine(r->records);
            const int64_t lead = r->records == 0 ? 1 : 0;
            if (take && rec_w + lead + rec_len + (e - b) + 2 > cap) goto no_room;
            if (take && rec_len == 0 && lead) out[rec_w] = ' ';
            const uint8_t any = take ? copy_upper(out + rec_w + lead + rec_len, d + b, e - b) : or_bytes(d + b, e - b);
            if (any & 0x80u) {
                int64_t i = b;
                while (d[i] < 0x80) ++i;
                pg_set_error("pg_fasta_next_chunk: non-ASCII byte 0x%02x in sequence text at file offset %lld", d[i], (long long)i);
                return PG_FASTA_ENONASCII;
            }
            rec_len += e - b;
        }
        pos = next;
    }
    if (!finish_record()) goto no_room;
    r->pos = r->size;
    r->done = true;
    return w;

no_room:
    if (w == 0) {
        pg_set_error("pg_fasta_next_chunk: a single record does not fit %lld bytes", (long long)cap);
        return PG_FASTA_ETOOSMALL;
    }
    r->pos = rec_start_pos;
    return w;
}

// ------------------------------------------------------------------------------------------------
// 5-bit host format of the corpus buffer: what crosses PCIe when a corpus is streamed from host memory
// (end-to-end step; corpora larger than HBM that are re-uploaded per n level).  Protein text needs 5 bits
// per symbol, so 8 symbols travel in 5 bytes (0.625 B/symbol instead of 1): symbol i of a group sits in bits
// [5i, 5i+5) of the group's 40-bit little-endian word.  Fixed code (no alphabet pass needed on the host):
//   0 = ' ', 1..26 = 'A'..'Z', 27 = '*', 28 = '-', 29 = '.', 31 = 0xFF (sequence separator; also the tail padding,
//   which adds no window).  Any other byte is refused (PG_EPACK): the caller keeps the byte format.
// pg_unpack5 (device) restores the byte buffer the kernels read; csrc/ngram.cu holds its kernel.
// ------------------------------------------------------------------------------------------------
namespace {
inline int code5(uint8_t c) {
    if (c == ' ') return 0;
    if (c >= 'A' && c <= 'Z') return c - 'A' + 1;
    if (c == '*') return 27;
    if (c == '-') return 28;
    if (c == '.') return 29;
    if (c == PG_SEP) return 31;
    return -1;
}
}  // namespace

extern "C" int64_t pg_pack5_bytes(int64_t n_symbols) { return n_symbols < 0 ? -1 : (n_symbols + 7) / 8 * 5; }

extern "C" int64_t pg_pack5_host(const uint8_t *bytes, int64_t n_symbols, uint8_t *out) {
    if (n_symbols < 0 || (n_symbols > 0 && (bytes == nullptr || out == nullptr))) {
        pg_set_error("pg_pack5_host: bad arguments");
        return PG_EINVAL;
    }
    uint8_t lut[256];
    for (int c = 0; c < 256; ++c) lut[c] = (uint8_t)(code5((uint8_t)c) < 0 ? 0x80 : code5((uint8_t)c));
    const int64_t groups = (n_symbols + 7) / 8, full = n_symbols / 8;
    uint8_t bad = 0;
    for (int64_t g = 0; g < full; ++g) {   // branch-free: invalid bytes only raise the top bit of `bad`
        const uint8_t *src = bytes + g * 8;
        uint64_t word = 0;
        for (int i = 0; i < 8; ++i) {
            const uint8_t c = lut[src[i]];
            bad |= c;
            word |= (uint64_t)(c & 31u) << (5 * i);
        }
        for (int b = 0; b < 5; ++b) out[g * 5 + b] = (uint8_t)(word >> (8 * b));
    }
    if (full < groups) {                   // tail group, padded with separators
        uint64_t word = 0;
        for (int i = 0; i < 8; ++i) {
            const int64_t p = full * 8 + i;
            const uint8_t c = p < n_symbols ? lut[bytes[p]] : (uint8_t)31;
            bad |= c;
            word |= (uint64_t)(c & 31u) << (5 * i);
        }
        for (int b = 0; b < 5; ++b) out[full * 5 + b] = (uint8_t)(word >> (8 * b));
    }
    if (bad & 0x80u) {
        int64_t p = 0;
        while (p < n_symbols && !(lut[bytes[p]] & 0x80u)) ++p;
        pg_set_error("pg_pack5_host: byte 0x%02x at offset %lld has no 5-bit code", bytes[p], (long long)p);
        return PG_EPACK;
    }
    return groups * 5;
}

// ------------------------------------------------------------------------------------------------
// Whole-file variant on several host threads.  The file is cut into ranges that start at header lines ('>' right after
// a line break), so no record spans two ranges; pass 1 counts the records of every range, a prefix sum gives every
// range its first global record index (rank dealing and the leading space of record #0 depend on it), pass 2 sizes
// this rank's bytes per range, pass 3 writes them -- all three passes in parallel over the ranges.  Same output as
// draining pg_fasta_next_chunk into one buffer.
// ------------------------------------------------------------------------------------------------
#include <thread>
#include <vector>

namespace {
struct RangeStat {
    int64_t lo, hi;         // file range
    int64_t records;        // records emitted in this range
    int64_t stop_at;        // file offset of a bare ">" header inside the range, or -1
    int64_t first_index;    // global index of the range's first record
    int64_t bytes;          // packed bytes of this rank's records
    int64_t out_off;        // where they go
    int bad;                // non-ASCII sequence byte seen
};

// mode 0: count records (+ find the stop);  1: size this rank's bytes;  2: write them
template <int MODE>
void parse_range(const uint8_t *d, RangeStat &r, uint8_t *out, int rank, int world, int block) {
    int64_t pos = r.lo;
    const int64_t end = r.stop_at >= 0 ? r.stop_at : r.hi;
    bool open_rec = false;
    int64_t rec_len = 0, idx = r.first_index, w = r.out_off, bytes = 0, records = 0;
    auto mine = [&](int64_t i) { return (i / block) % world == rank; };
    auto finish = [&]() {
        if (open_rec && rec_len > 0) {
            if (MODE == 0) bytes += rec_len + 2;   // world == 1 needs no sizing pass: every record is this rank's
            if (MODE >= 1 && mine(idx)) {
                const int64_t lead = idx == 0 ? 1 : 0;
                if (MODE == 2) {
                    out[w + lead + rec_len] = ' ';
                    out[w + lead + rec_len + 1] = PG_SEP;
                    w += lead + rec_len + 2;
                }
                bytes += lead + rec_len + 2;
            }
            ++idx;
            ++records;
        }
        open_rec = false;
    };
    while (pos < end) {
        int64_t b, e;
        const int64_t line_pos = pos;
        const int64_t next = next_line(d, pos, r.hi, &b, &e);
        pos = next;
        if (b == e) continue;
        if (d[b] == '>') {
            finish();
            if (e - b == 1) {   // the reference's generator stops here
                if (MODE == 0) r.stop_at = line_pos;
                break;
            }
            open_rec = true;
            rec_len = 0;
            continue;
        }
        if (!open_rec) continue;
        if (MODE == 2 && mine(idx)) {
            const int64_t lead = idx == 0 ? 1 : 0;
            if (rec_len == 0 && lead) out[w] = ' ';
            if (copy_upper(out + w + lead + rec_len, d + b, e - b) & 0x80u) r.bad = 1;
        } else if (MODE == 0) {
            if (or_bytes(d + b, e - b) & 0x80u) r.bad = 1;
        }
        rec_len += e - b;
    }
    finish();
    if (MODE == 0) r.records = records;
    if (MODE <= 1) r.bytes = bytes;
}

template <class F>
void run_parallel(int n, int threads, F f) {
    if (threads <= 1 || n <= 1) {
        for (int i = 0; i < n; ++i) f(i);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([=] { for (int i = t; i < n; i += threads) f(i); });
    for (auto &th : pool) th.join();
}
}  // namespace

namespace {
// Packs the records of the file window [win_lo, win_hi) -- win_lo at a header line (or 0), win_hi at a header line or the
// end of the file -- whose first record has the global index first_index.  Returns bytes written or a negative code.
int64_t pack_window(const uint8_t *d, int64_t win_lo, int64_t win_hi, int64_t first_index, uint8_t *out, int64_t cap, int threads,
                    int rank, int world, int block, int64_t *n_records, int *stopped_early) {
    const int64_t size = win_hi;
    const int64_t span = win_hi - win_lo;
    // ranges: nominal cuts moved forward to the next header line
    const int n_ranges = (int)(span < (1 << 20) ? 1 : (threads * 4 < 256 ? threads * 4 : 256));
    std::vector<RangeStat> rs;
    auto next_header = [&](int64_t from) -> int64_t {   // first '>' at or after `from` that directly follows a line break
        int64_t p = from;
        while (p < size) {
            const uint8_t *q = (const uint8_t *)memchr(d + p, '>', (size_t)(size - p));
            if (q == nullptr) return size;
            p = (int64_t)(q - d);
            if (p > 0 && (d[p - 1] == '\n' || d[p - 1] == '\r')) return p;
            ++p;
        }
        return size;
    };
    int64_t lo = win_lo;
    for (int i = 1; i <= n_ranges && lo < size; ++i) {
        const int64_t nominal = win_lo + span / n_ranges * i;
        const int64_t cut = i == n_ranges ? size : next_header(nominal > lo ? nominal : lo + 1);
        if (cut > lo) {
            rs.push_back(RangeStat{lo, cut, 0, -1, 0, 0, 0, 0});
            lo = cut;
        }
    }
    const int n = (int)rs.size();
    run_parallel(n, threads, [&](int i) { parse_range<0>(d, rs[i], nullptr, rank, world, block); });
    int stop_range = -1;
    int64_t total_records = 0;
    for (int i = 0; i < n; ++i) {
        if (stop_range >= 0) {      // everything after the stop is ignored
            rs[i].hi = rs[i].lo;
            rs[i].records = 0;
            continue;
        }
        rs[i].first_index = first_index + total_records;
        total_records += rs[i].records;
        if (rs[i].stop_at >= 0) stop_range = i;
    }
    int64_t rcode = 0;
    for (int i = 0; i < n; ++i)
        if (rs[i].bad && rs[i].hi > rs[i].lo) rcode = PG_FASTA_ENONASCII;
    if (rcode == 0) {
        if (world > 1) {
            run_parallel(n, threads, [&](int i) { if (rs[i].hi > rs[i].lo) parse_range<1>(d, rs[i], nullptr, rank, world, block); });
        } else {
            for (int i = 0; i < n; ++i)   // pass 1 sized every record; the leading space goes to the range that holds record #0
                if (rs[i].hi > rs[i].lo && rs[i].records > 0 && rs[i].first_index == 0) rs[i].bytes += 1;
        }
        int64_t off = 0;
        for (int i = 0; i < n; ++i) {
            rs[i].out_off = off;
            off += rs[i].hi > rs[i].lo ? rs[i].bytes : 0;
        }
        if (off > cap) {
            pg_set_error("pg_fasta_pack_parallel: %lld bytes do not fit the buffer (%lld)", (long long)off, (long long)cap);
            rcode = PG_FASTA_ETOOSMALL;
        } else {
            run_parallel(n, threads, [&](int i) { if (rs[i].hi > rs[i].lo) parse_range<2>(d, rs[i], out, rank, world, block); });
            rcode = off;
        }
    } else {
        pg_set_error("pg_fasta_pack_parallel: non-ASCII byte in sequence text");
    }
    if (n_records) *n_records = total_records;
    if (stopped_early) *stopped_early = stop_range >= 0 ? 1 : 0;
    return rcode;
}
}  // namespace

extern "C" int64_t pg_fasta_pack_parallel(const char *path, uint8_t *out, int64_t cap, int threads, int rank, int world, int block,
                                          int64_t *n_records, int *stopped_early) {
    if (out == nullptr || cap < 4 || world < 1 || rank < 0 || rank >= world || block < 1 || threads < 1) {
        pg_set_error("pg_fasta_pack_parallel: bad arguments");
        return PG_EINVAL;
    }
    pg_fasta_reader *rd = pg_fasta_open(path);
    if (rd == nullptr) return PG_EINVAL;
    const int64_t rcode = pack_window(rd->data, 0, rd->size, 0, out, cap, threads, rank, world, block, n_records, stopped_early);
    pg_fasta_close(rd);
    return rcode;
}

// The same over ONE WINDOW of an open file, for files that do not fit host memory as a single corpus buffer: the caller walks
// the file window by window (start at *pos = 0, first_index = 0; add *n_records to first_index after every call).  The window
// ends at the first header line at or after *pos + window_bytes (or the end of the file); *pos is advanced to it.  `out` needs
// (window end - window start) + 16 bytes at most; PG_FASTA_ETOOSMALL leaves *pos unchanged so the call can be repeated.
extern "C" int64_t pg_fasta_pack_window(pg_fasta_reader *rd, int64_t *pos, int64_t window_bytes, int64_t first_index, uint8_t *out,
                                        int64_t cap, int threads, int rank, int world, int block, int64_t *n_records,
                                        int *stopped_early) {
    if (rd == nullptr || pos == nullptr || out == nullptr || cap < 4 || world < 1 || rank < 0 || rank >= world || block < 1 ||
        threads < 1 || window_bytes < 1 || first_index < 0 || *pos < 0 || *pos > rd->size) {
        pg_set_error("pg_fasta_pack_window: bad arguments");
        return PG_EINVAL;
    }
    const uint8_t *d = rd->data;
    const int64_t size = rd->size, lo = *pos;
    int64_t hi = size;
    if (lo + window_bytes < size) {     // first header line at or after the nominal end
        int64_t p = lo + window_bytes;
        hi = size;
        while (p < size) {
            const uint8_t *q = (const uint8_t *)memchr(d + p, '>', (size_t)(size - p));
            if (q == nullptr) break;
            p = (int64_t)(q - d);
            if (d[p - 1] == '\n' || d[p - 1] == '\r') {
                hi = p;
                break;
            }
            ++p;
        }
    }
    const int64_t rcode = pack_window(d, lo, hi, first_index, out, cap, threads, rank, world, block, n_records, stopped_early);
    if (rcode >= 0) *pos = hi;
    return rcode;
}
