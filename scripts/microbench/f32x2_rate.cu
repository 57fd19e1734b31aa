// f32x2_rate.cu -- issue / pipe rate of the packed FP32 instructions (FFMA2, FADD2) against the scalar ones on sm_100a.
// Each thread runs CH independent dependency chains; the printed figure is lane-FLOP/s and warp-instructions per clock
// per SM sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CH = 8, ITERS = 4096;

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
    float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ float add1(float a, float b) {
    float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d;
}

template <int MODE> __global__ void __launch_bounds__(256) rate(float *out, float seed) {
    float2 acc[CH]; float2 m = make_float2(seed, seed * 0.5f), c = make_float2(0.25f, 0.125f);
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) { acc[i].x = fma1(acc[i].x, m.x, c.x); acc[i].y = fma1(acc[i].y, m.y, c.y); }
            if (MODE == 1) {
                unsigned long long r = fma2(*reinterpret_cast<unsigned long long *>(&acc[i]), *reinterpret_cast<unsigned long long *>(&m),
                                            *reinterpret_cast<unsigned long long *>(&c));
                acc[i] = *reinterpret_cast<float2 *>(&r);
            }
            if (MODE == 2) { acc[i].x = add1(acc[i].x, c.x); acc[i].y = add1(acc[i].y, c.y); }
            if (MODE == 3) {
                unsigned long long r = add2(*reinterpret_cast<unsigned long long *>(&acc[i]), *reinterpret_cast<unsigned long long *>(&c));
                acc[i] = *reinterpret_cast<float2 *>(&r);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, int warps_per_smsp) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int threads = 256, blocks = sms * warps_per_smsp * 4 * 32 / threads;
    float *out; cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    rate<MODE><<<blocks, threads>>>(out, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) rate<MODE><<<blocks, threads>>>(out, 1.0001f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    const double lane_ops = (double)blocks * threads * ITERS * CH * 2;          // f32 lanes processed
    const double winst = (double)blocks * (threads / 32) * ITERS * CH * ((MODE & 1) ? 1 : 2);
    printf("{\"op\": \"%s\", \"warps_per_smsp\": %d, \"ms\": %.4f, \"lane_Gops\": %.1f, \"warp_inst_per_clk_per_smsp_at_max_clock\": %.3f}\n",
           name, warps_per_smsp, ms, lane_ops / ms * 1e-6, winst / (ms * 1e-3) / ((double)khz * 1e3) / (sms * 4));
    cudaFree(out);
}

int main() {
    for (int w : {2, 4, 8}) {
        run<0>("FFMA", w); run<1>("FFMA2", w); run<2>("FADD", w); run<3>("FADD2", w);
    }
    return 0;
}
