// How long does the hardware need to replace one CTA of the network kernel's shape (480 threads, 128 registers,
// 232 448 B of dynamic shared memory, i.e. one CTA per SM) by the next one?  Empty CTAs, many more than SMs, timed
// with CUDA events: time / (CTAs per SM) = turn-over per CTA.  With and without a tensor-memory allocation.
// Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cta_turnover tools/cta_turnover.cu && /tmp/cta_turnover
#include <cstdio>
#include <cuda_runtime.h>

template <bool kTmem, bool kBarriers>
__global__ void __launch_bounds__(480, 1) k_empty(int* sink, int spin) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned tmem_ptr;
    const int warp = threadIdx.x >> 5;
    if (kBarriers && threadIdx.x == 0) {
        for (int i = 0; i < 70; ++i) {
            unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(smem + 8 * i));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        }
    }
    if (kTmem && warp == 13) {
        unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(&tmem_ptr));
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    __syncthreads();
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if (sink && threadIdx.x == 0 && smem[threadIdx.x] == 77) sink[0] = 1;
    __syncthreads();
    if (kTmem && warp == 13) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512));
    }
}

template <bool kTmem, bool kBarriers>
static void run(const char* what, int sms, int spin) {
    const int smem = 232448 - 1024;
    cudaFuncSetAttribute(k_empty<kTmem, kBarriers>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int per_sm = 200;
    k_empty<kTmem, kBarriers><<<sms * 4, 480, smem>>>(nullptr, spin);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_empty<kTmem, kBarriers><<<sms * per_sm, 480, smem>>>(nullptr, spin);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double us = 1e3 * ms / per_sm;
    printf("%-44s spin %6d cycles: %7.2f us per CTA = %7.0f cycles at %d MHz, turn-over %6.0f cycles (%s)\n", what, spin, us,
           us * clk / 1e3, clk / 1000, us * clk / 1e3 - spin, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int spin : {0, 20000, 70000}) {
        run<false, false>("480 threads, 227 KB smem", sms, spin);
        run<false, true>("... + 70 mbarrier.init by one thread", sms, spin);
        run<true, true>("... + tcgen05.alloc / dealloc of 512 columns", sms, spin);
    }
    return 0;
}
