// Probe (not part of the product): which shared-memory word does tcgen05.mma.kind::tf32 read for the logical element
// (m, k) of A -- or (n, k) of B -- under a given smem descriptor / major bit?  The probed operand's buffer holds its own
// word index (split into low / high 10 bits so the values are exact in tf32), the other operand is a K-major selector
// (known-good layout), so D[m][n] comes back as the index of the word used for k = n.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I protgram-directgcn_b200/csrc -I include -o /tmp/probe tools/probe_umma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

void pg_set_error(const char *, ...) {}
void pg_count_launch() {}

using namespace pgtc;

constexpr int WORDS = 8192;   // probed operand buffer: 32 KB

__global__ void __launch_bounds__(128) probe_kernel(int probe_b, uint32_t lbo, uint32_t sbo, uint32_t layout_type, int major_bit, int high_part,
                                                    int n_cols, float *__restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_slot;
    float *probe = reinterpret_cast<float *>(smem);                  // WORDS floats
    float *sel = reinterpret_cast<float *>(smem + WORDS * 4);        // selector, K-major no-swizzle: chunk c at c*LBO + r*16
    const int tid = threadIdx.x, warp = tid >> 5;
    const int sel_rows = probe_b ? 128 : n_cols;                     // selector is A (128 rows) when B is probed
    const uint32_t sel_lbo = (uint32_t)sel_rows * 16u;
    for (int i = tid; i < WORDS; i += 128) probe[i] = (float)(high_part ? (i >> 10) : (i & 1023));
    for (int i = tid; i < sel_rows * 8; i += 128) {
        const int r = i / 8, k = i % 8;
        sel[(k / 4) * (sel_lbo / 4) + r * 4 + (k % 4)] = (k == (r % 8)) ? 1.f : 0.f;
    }
    if (tid == 0) {
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 32);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = tmem_slot;
    if (tid == 0) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (major_bit) idesc |= probe_b ? (1u << 16) : (1u << 15);
        const uint64_t d_probe = umma_desc(smem_u32(probe), lbo, sbo) | ((uint64_t)layout_type << 61);
        const uint64_t d_sel = umma_desc(smem_u32(sel), sel_lbo, 128);
        umma_tf32(tmem_d, probe_b ? d_sel : d_probe, probe_b ? d_probe : d_sel, idesc, 0);
        umma_commit(&done);
    }
    const bool ok = mbar_wait(&done, 0);
    fence_after_sync();
    uint32_t r[16];
    tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16), r);
    for (int c = 0; c < 16; ++c) out[tid * 16 + c] = ok ? __uint_as_float(r[c]) : -1.f;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 32);
}

int main() {
    float *d_out;
    cudaMalloc(&d_out, 128 * 16 * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    struct Cfg { const char *name; int probe_b; uint32_t lbo, sbo, layout; int major; };
    const Cfg cfgs[] = {
        {"A K-major  none   lbo=2048 sbo=128 (sanity: word = (k/4)*512 + m*4 + k%4)", 0, 2048, 128, 0, 0},
        {"A MN-major none   lbo=4096 sbo=128", 0, 4096, 128, 0, 1},
        {"A MN-major none   lbo=128  sbo=4096", 0, 128, 4096, 0, 1},
        {"A MN-major none   lbo=256  sbo=128", 0, 256, 128, 0, 1},
        {"A MN-major sw128  lbo=1024 sbo=4096", 0, 1024, 4096, 2, 1},
        {"A MN-major sw128  lbo=4096 sbo=1024", 0, 4096, 1024, 2, 1},
        {"A MN-major sw64   lbo=512  sbo=2048", 0, 512, 2048, 4, 1},
        {"A MN-major sw32   lbo=256  sbo=1024", 0, 256, 1024, 6, 1},
        {"B K-major  none   lbo=256  sbo=128 (sanity, N=16)", 1, 256, 128, 0, 0},
        {"B MN-major none   lbo=512  sbo=128 (N=16)", 1, 512, 128, 0, 1},
        {"B MN-major sw128  lbo=1024 sbo=1024 (N=16)", 1, 1024, 1024, 2, 1},
    };
    std::vector<float> lo(128 * 16), hi(128 * 16);
    for (const Cfg &c : cfgs) {
        for (int part = 0; part < 2; ++part) {
            probe_kernel<<<1, 128, 48 * 1024>>>(c.probe_b, c.lbo, c.sbo, c.layout, c.major, part, 16, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e));
                return 1;
            }
            cudaMemcpy(part ? hi.data() : lo.data(), d_out, 128 * 16 * 4, cudaMemcpyDeviceToHost);
        }
        printf("== %s\n", c.name);
        const int rows[] = {0, 1, 2, 3, 4, 5, 8, 9, 16, 32, 33, 64, 127};
        for (int m : rows) {
            printf("  row %3d:", m);
            for (int n = 0; n < 8; ++n) printf(" %6d", (int)hi[m * 16 + n] * 1024 + (int)lo[m * 16 + n]);
            printf("\n");
        }
    }
    return 0;
}
