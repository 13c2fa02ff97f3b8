// Microbenchmark (not part of the product): shared-memory atomics routed through a 2-CTA cluster
// (DSMEM).  Each thread adds to the table of the CTA that owns the key's half: 50% remote.
#include <cstdint>
#include <cstdio>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16; return x;
}

template <int MODE>  // 0: all local, 1: route by key half (50% remote, no return), 2: same with return consumed
__global__ void __launch_bounds__(1024, 1) dsmem_atomics(unsigned *out, uint32_t sb, long ops_per_cta) {
    extern __shared__ unsigned tbl[];
    cg::cluster_group cluster = cg::this_cluster();
    for (uint32_t i = threadIdx.x; i < sb; i += blockDim.x) tbl[i] = 0;
    cluster.sync();
    unsigned *t0 = cluster.map_shared_rank(tbl, 0);
    unsigned *t1 = cluster.map_shared_rank(tbl, 1);
    unsigned acc = 0;
    const long base = (long)blockIdx.x * ops_per_cta;
    for (long i = threadIdx.x; i < ops_per_cta; i += blockDim.x) {
        uint32_t h = mix32((uint32_t)(base + i));
        uint32_t k = (h >> 1) % sb;
        unsigned *t = (MODE == 0) ? tbl : ((h & 1) ? t1 : t0);
        if (MODE == 2) acc += atomicAdd(&t[k], 1u);
        else atomicAdd(&t[k], 1u);
    }
    cluster.sync();
    unsigned s = acc;
    for (uint32_t i = threadIdx.x; i < sb; i += blockDim.x) s += tbl[i];
    if (s == 0xdeadbeef) out[blockIdx.x] = s;
}

template <typename K>
float run(K kern, unsigned *out, uint32_t sb, long per) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = sb * 4UL;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaLaunchKernelEx(&cfg, kern, out, sb, per);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) cudaLaunchKernelEx(&cfg, kern, out, sb, per);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    const long n_ops = 175000000L; long per = n_ops / 148;
    unsigned *out; cudaMalloc(&out, 4096);
    for (uint32_t sb : {9261u, 48621u}) {
        float a = run(dsmem_atomics<0>, out, sb, per), b = run(dsmem_atomics<1>, out, sb, per), c = run(dsmem_atomics<2>, out, sb, per);
        printf("words=%6u local: %.3f ms (%.0f Gop/s)  routed 50%% remote: %.3f ms (%.0f Gop/s)  routed+return: %.3f ms (%.0f Gop/s)\n",
               sb, a, n_ops / a / 1e6, b, n_ops / b / 1e6, c, n_ops / c / 1e6);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
