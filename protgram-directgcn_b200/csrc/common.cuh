// Shared helpers for libpgb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pgb200.h"

#define PG_NUM_SMS 148  // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this

void pg_set_error(const char *fmt, ...);
void pg_count_launch();  // diagnostics: number of kernels this library has launched (pg_launch_count)

#define PG_CHECK_ARG(cond, ...)        \
    do {                               \
        if (!(cond)) {                 \
            pg_set_error(__VA_ARGS__); \
            return PG_EINVAL;          \
        }                              \
    } while (0)

#define PG_CUDA_LAUNCH_CHECK(name)                                                    \
    do {                                                                              \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            pg_set_error("%s: kernel launch failed: %s", name, cudaGetErrorString(_e)); \
            return PG_ECUDA;                                                          \
        }                                                                             \
        pg_count_launch();                                                            \
    } while (0)

#define PG_CUDA_CALL(expr)                                                          \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            pg_set_error("%s failed: %s", #expr, cudaGetErrorString(_e));           \
            return PG_ECUDA;                                                        \
        }                                                                           \
    } while (0)

static inline size_t pg_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t pg_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Bump allocator over the caller's workspace (256 B aligned slices).
struct PgArena {
    char *base;
    size_t cap, used;
    bool ok;
    PgArena(void *p, size_t bytes) : base((char *)p), cap(bytes), used(0), ok(true) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = pg_align_up(count * sizeof(T), 256);
        if (used + bytes > cap) {
            ok = false;
            return nullptr;
        }
        T *r = (T *)(base + used);
        used += bytes;
        return r;
    }
};

// ---- device-wide exclusive scan of int64 (scan.cu) ----
size_t pg_scan_ws_bytes(int64_t n);
// d_out[i] = sum_{j<i} d_in[j]; d_total (optional, device) = sum of all.  In-place allowed.
int pg_exclusive_scan_i64(const int64_t *d_in, int64_t *d_out, int64_t n, int64_t *d_total, void *d_ws,
                          size_t ws_bytes, cudaStream_t stream);

// gating of dZ + gate gradients after a data-gradient GEMM (gemm.cu; also used by gemm_tc.cu)
int pg_launch_gate_grad(float *d_dz, int64_t lddz, const float *d_z, int64_t ldz, const float *d_dy, int64_t lddy, const float *d_w_ext,
                        const float *d_gate_a, const float *d_gate_b, const float *d_gate_c, int gate_stride, int64_t num_rows, int F_in,
                        int F_out, int k_data, float *d_dgate, cudaStream_t st);

static inline cudaStream_t pg_cu(pg_stream_t s) { return (cudaStream_t)s; }
