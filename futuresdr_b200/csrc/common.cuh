// common.cuh -- context object, error plumbing and small device helpers shared by all
// translation units of libb200sdr.so.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "b200sdr.h"

struct b2s_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    // side streams + events for the host-slice pipeline (b2s_fir_filter_host)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    int sm_count = 0;
    size_t smem_optin = 0;
    std::atomic<uint64_t> launches{0};
    std::string err;
    // workspace for *_host calls (grown on demand, freed with the context)
    void *ws_dev = nullptr;
    size_t ws_bytes = 0;
    cudaEvent_t hev[16] = {};      // events of the host-slice pipeline, created once
    bool hev_ready = false;
    std::mutex host_mu;            // serialises the *_host pipelines that share the workspace and side streams
    // device status word (bit0: a cross-GPU flag wait timed out); checked by b2s_ctx_sync when flag_ops > 0
    unsigned *d_status = nullptr;
    uint64_t flag_ops = 0;
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE and a process may hold contexts on several: a
// `static PerDeviceOnce` next to each kernel instantiation remembers which devices have been opted in.
struct PerDeviceOnce {
    std::atomic<uint64_t> mask{0};
    bool need(int dev) const { return ((mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull) == 0; }
    void done(int dev) { mask.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

#ifdef __CUDACC__
// acc += x * t for a Complex<f32> sample and a REAL tap as ONE packed instruction (sm_100 FFMA2: both lanes are IEEE
// fused multiply-adds, bit-identical to the two scalar FFMAs).  The packed form runs at the same lane rate as FFMA
// (scripts/microbench/f32x2_rate.cu: 35 T lane-op/s either way) but takes one issue slot instead of two, which is what
// the sliding-window kernels are short of (ncu: 80 % of the issue slots busy, 43 % of them FFMA).
// tt = (t, t); build it once per tap with dup2() so the register pair is reused by every MAC of that tap.
__device__ __forceinline__ unsigned long long dup2(float t) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(t));
    return d;
}
__device__ __forceinline__ void cmac2(float2 &acc, float2 x, unsigned long long tt) {
    unsigned long long a = *reinterpret_cast<unsigned long long *>(&acc);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(*reinterpret_cast<unsigned long long *>(&x)), "l"(tt));
    acc = *reinterpret_cast<float2 *>(&a);
}
#endif

extern thread_local std::string g_b2s_last_error;

int32_t b2s_fail(b2s_ctx *ctx, int32_t code, const char *fmt, ...);

#define B2S_CUDA(ctx, expr)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return b2s_fail((ctx), B2S_ECUDA, "%s failed: %s (%s:%d)", #expr,                 \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                     \
    } while (0)

#define B2S_CHECK_LAUNCH(ctx)                                                                 \
    do {                                                                                      \
        (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                              \
        B2S_CUDA((ctx), cudaGetLastError());                                                  \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// peer.cu: NVTX ranges (visible to nsys / ncu --nvtx) and the cross-GPU flag kernels
void nvtx_push(const char *name);
void nvtx_pop();
struct NvtxRange { explicit NvtxRange(const char *n) { nvtx_push(n); } ~NvtxRange() { nvtx_pop(); } };
int32_t peer_flag_set_launch(b2s_ctx *ctx, unsigned *flag, unsigned value, cudaStream_t st);
int32_t peer_flag_wait_launch(b2s_ctx *ctx, const unsigned *flag, unsigned value, cudaStream_t st);

static inline size_t sat_sub(size_t a, size_t b) { return a > b ? a - b : 0; }
static inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }
static inline size_t round_up(size_t a, size_t b) { return ceil_div(a, b) * b; }

static inline size_t kind_in_bytes(b2s_kind k) { return k == B2S_F32_F32 ? 4 : 8; }   // F64_F64: 8 as well
static inline size_t kind_tap_floats(b2s_kind k) { return k == B2S_C32_C32 ? 2 : 1; }
