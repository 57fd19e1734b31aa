// mma_rate.cu -- micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of
// N, A source (TMEM .ts / smem .ss) and accumulator reuse.  Used to size the tensor FIR tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) k(long long *out, int iters) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(128, N);
        const uint64_t bdesc = make_desc(base), adesc = make_desc(base + 64 * 1024);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const uint32_t d = tmem + 256 + (uint32_t)((i % NACC) * N);
            const uint32_t a = tmem + (uint32_t)((i % 24) * 8);
            const uint64_t b = bdesc + (uint64_t)(((i % 4) * 32) >> 4);
            if (TS) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(1u));
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(b), "r"(idesc), "r"(1u));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u));
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, bool TS, int NACC> void run(const char *name, int grid) {
    long long *d; cudaMalloc(&d, 8);
    const int iters = 4000;
    auto kern = k<N, TS, NACC>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    kern<<<grid, 128, 200 * 1024>>>(d, iters);
    kern<<<grid, 128, 200 * 1024>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s grid=%3d  %.1f cycles/MMA  (ideal %d)  %s\n", name, grid, (double)h / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<64, true, 1>("N=64  .ts same-acc", grid);
        run<64, true, 2>("N=64  .ts 2-acc alternating", grid);
        run<128, true, 1>("N=128 .ts same-acc", grid);
        run<256, true, 1>("N=256 .ts same-acc", grid);
        run<64, false, 1>("N=64  .ss same-acc", grid);
        run<128, false, 1>("N=128 .ss same-acc", grid);
        run<256, false, 1>("N=256 .ss same-acc", grid);
        run<32, true, 1>("N=32  .ts same-acc", grid);
        run<16, true, 1>("N=16  .ts same-acc", grid);
    }
    return 0;
}
