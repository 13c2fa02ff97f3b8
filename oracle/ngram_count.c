/* CPU oracle, plain C -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Restates the transition counting of the reference builder for one n level on the SAME
 * corpus buffer layout the CUDA path consumes (include/pgb200.h, "corpus buffer"):
 *
 *     [' ' only before global sequence #0] seq_0 ' ' 0xFF seq_1 ' ' 0xFF ...
 *
 * i.e. every padded sequence of /root/reference/src/pipeline/data_builder.py:29-35 followed by
 * one separator byte 0xFF.  For every position i whose m = n+1 bytes hold no separator this is
 * one transition  window(i, n) -> window(i+1, n)   (data_builder.py:45-54); transitions are
 * counted (data_builder.py:267-273) in a dense table indexed by the base-sigma number of the
 * m symbol ranks.  rank_of_byte maps a byte to its rank among the bytes present in the corpus
 * (ascending byte value == Python string order for ASCII, data_builder.py:164).
 * n-gram presence (data_builder.py:38-42) is recorded for every separator-free n-window.
 *
 * Pinned: tests/test_oracle_golden.py compares this against the reference-generated goldens.
 */
#include <stdint.h>
#include <string.h>

#define SEP 0xFF

/* bins: sigma^m uint64 (zeroed by caller); present: sigma^n uint8 (zeroed by caller). */
int oracle_ngram_count(const uint8_t *buf, int64_t nbytes, int n, const uint8_t *rank_of_byte,
                       int sigma, uint64_t *bins, uint8_t *present)
{
    const int m = n + 1;
    uint64_t pow_n = 1;
    for (int k = 0; k < n; ++k) pow_n *= (uint64_t)sigma;
    /* rolling base-sigma code of the last n symbols; `run` = separator-free bytes ending here */
    uint64_t code_n = 0;
    int64_t run = 0;
    for (int64_t i = 0; i < nbytes; ++i) {
        uint8_t b = buf[i];
        if (b == SEP) { run = 0; code_n = 0; continue; }
        uint64_t r = rank_of_byte[b];
        uint64_t prev = code_n; /* code of the n symbols ending at i-1 (valid when run >= n) */
        code_n = (code_n % (pow_n / (uint64_t)sigma)) * (uint64_t)sigma + r;
        if (n == 1) code_n = r;
        ++run;
        if (run >= n) present[code_n] = 1;
        if (run >= m) bins[prev * (uint64_t)sigma + r] += 1;
    }
    return 0;
}

/* 256-entry byte presence (separator excluded) -> alphabet discovery. */
int oracle_byte_presence(const uint8_t *buf, int64_t nbytes, uint8_t *present256)
{
    memset(present256, 0, 256);
    for (int64_t i = 0; i < nbytes; ++i) present256[buf[i]] = 1;
    present256[SEP] = 0;
    return 0;
}
