// Microbenchmark (not part of the product): what bounds the (n+1)-gram count kernel on B200?
//   A) global RED.ADD.u64 / u32 into an L2-resident dense table of NB bins, uniform random keys
//   B) shared-memory atomicAdd (no return) into a per-CTA table of SB 32-bit words, 1 CTA per SM
//   C) shared-memory atomicAdd with the return value consumed
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench_atomics.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16; return x;
}

template <typename T>
__global__ void __launch_bounds__(256) global_red(T *bins, uint32_t nb, long n_ops) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    for (; i < n_ops; i += stride) {
        uint32_t k = mix32((uint32_t)i) % nb;
        atomicAdd(&bins[k], (T)1);
    }
}

template <int MODE>  // 0 = no return, 1 = return consumed, 2 = predicated quarter (only keys in my quarter)
__global__ void __launch_bounds__(1024, 1) smem_atomics(unsigned *out, uint32_t sb, long ops_per_cta, uint32_t nb_total) {
    extern __shared__ unsigned tbl[];
    for (uint32_t i = threadIdx.x; i < sb; i += blockDim.x) tbl[i] = 0;
    __syncthreads();
    unsigned acc = 0;
    const long base = (long)blockIdx.x * ops_per_cta;
    for (long i = threadIdx.x; i < ops_per_cta; i += blockDim.x) {
        uint32_t h = mix32((uint32_t)(base + i));
        if (MODE == 2) {
            uint32_t k = h % nb_total;
            uint32_t q = blockIdx.x & 3;
            if (k / sb == q) atomicAdd(&tbl[k - q * sb], 1u);
        } else {
            uint32_t k = h % sb;
            if (MODE == 0) atomicAdd(&tbl[k], 1u);
            else acc += atomicAdd(&tbl[k], 1u);
        }
    }
    __syncthreads();
    unsigned s = acc;
    for (uint32_t i = threadIdx.x; i < sb; i += blockDim.x) s += tbl[i];
    if (s == 0xdeadbeef) out[blockIdx.x] = s;
}

template <typename F>
float time_ms(F f, int iters = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main() {
    const long n_ops = 175000000L;
    const uint32_t nbs[] = {441, 9261, 194481, 4084101};
    void *bins; cudaMalloc(&bins, 4084101UL * 8);
    for (uint32_t nb : nbs) {
        cudaMemset(bins, 0, nb * 8UL);
        float m64 = time_ms([&] { global_red<unsigned long long><<<148 * 16, 256>>>((unsigned long long *)bins, nb, n_ops); });
        float m32 = time_ms([&] { global_red<unsigned><<<148 * 16, 256>>>((unsigned *)bins, nb, n_ops); });
        printf("global RED  bins=%8u  u64: %7.3f ms (%6.1f Gop/s)   u32: %7.3f ms (%6.1f Gop/s)\n", nb, m64, n_ops / m64 / 1e6, m32, n_ops / m32 / 1e6);
    }
    unsigned *out; cudaMalloc(&out, 148 * 4 * 4);
    const uint32_t sbs[] = {441, 9261, 48621};
    for (uint32_t sb : sbs) {
        size_t smem = sb * 4UL;
        cudaFuncSetAttribute(smem_atomics<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(smem_atomics<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(smem_atomics<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        long per = n_ops / 148;
        float a = time_ms([&] { smem_atomics<0><<<148, 1024, smem>>>(out, sb, per, sb * 4); });
        float b = time_ms([&] { smem_atomics<1><<<148, 1024, smem>>>(out, sb, per, sb * 4); });
        // mode 2: 4 CTAs (a cluster in the real kernel) each see the SAME ops and keep their quarter
        float c = time_ms([&] { smem_atomics<2><<<148, 1024, smem>>>(out, sb, per * 4, sb * 4); });
        printf("smem ATOMS  words=%6u  noret: %7.3f ms (%6.1f Gop/s)  ret: %7.3f ms (%6.1f Gop/s)  quarter-filter(4x keys): %7.3f ms (%6.1f Gres/s)\n",
               sb, a, n_ops / a / 1e6, b, n_ops / b / 1e6, c, n_ops / c / 1e6);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
