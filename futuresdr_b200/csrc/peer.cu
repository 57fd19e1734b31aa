// peer.cu -- cross-GPU plumbing of the sharded stream (new design, the reference is single-device):
// CUDA-IPC export / import of device allocations so that one process per GPU can read its left neighbour's chunk
// tail straight over NVLink, and the system-scope flags that order those reads without any host round trip.
//
//   producer (rank r)                               consumer (rank r+1)
//   ... fill chunk t ...                            b2s_fir_exec_hist(hist = peer tail of rank r,
//   b2s_flag_set(ready_r, t+1)   -- release.sys -->      wait_flag = ready_r, wait_value = t+1,
//                                                         done_flag = consumed_r, done_value = t+1)
//   before refilling that slot:                     (the FIR kernel's TMA loader spins on ready_r, fetches the tail,
//   b2s_flag_wait(consumed_r, t+1) <-- release.sys --  a converter thread stores consumed_r once it is in smem)
//
// Flags are 32-bit counters in DEVICE memory (wrap-safe signed comparison); every wait has a 4 s time-out that sets
// bit 0 of the context's status word instead of hanging the GPU (reported by b2s_ctx_sync as B2S_ETIMEOUT).
#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

void nvtx_push(const char *name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

namespace {

__global__ void flag_set_kernel(unsigned *flag, unsigned value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void flag_wait_kernel(const unsigned *flag, unsigned value, unsigned *status) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - value) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) { atomicOr(status, 1u); break; }
        __nanosleep(128);
    }
}

}  // namespace

int32_t peer_flag_set_launch(b2s_ctx *ctx, unsigned *flag, unsigned value, cudaStream_t st) {
    flag_set_kernel<<<1, 1, 0, st>>>(flag, value);
    B2S_CHECK_LAUNCH(ctx);
    return B2S_OK;
}
int32_t peer_flag_wait_launch(b2s_ctx *ctx, const unsigned *flag, unsigned value, cudaStream_t st) {
    ctx->flag_ops++;
    flag_wait_kernel<<<1, 1, 0, st>>>(flag, value, ctx->d_status);
    B2S_CHECK_LAUNCH(ctx);
    return B2S_OK;
}

extern "C" {

int32_t b2s_flag_set(b2s_ctx *ctx, uint32_t *d_flag, uint32_t value) {
    if (!ctx || !d_flag) return b2s_fail(ctx, B2S_EINVAL, "b2s_flag_set: NULL argument");
    DeviceGuard g(ctx->device);
    return peer_flag_set_launch(ctx, d_flag, value, ctx->stream);
}

int32_t b2s_flag_wait(b2s_ctx *ctx, const uint32_t *d_flag, uint32_t value) {
    if (!ctx || !d_flag) return b2s_fail(ctx, B2S_EINVAL, "b2s_flag_wait: NULL argument");
    DeviceGuard g(ctx->device);
    return peer_flag_wait_launch(ctx, d_flag, value, ctx->stream);
}

int32_t b2s_flag_read(b2s_ctx *ctx, const uint32_t *d_flag, uint32_t *value) {
    if (!ctx || !d_flag || !value) return b2s_fail(ctx, B2S_EINVAL, "b2s_flag_read: NULL argument");
    DeviceGuard g(ctx->device);
    B2S_CUDA(ctx, cudaMemcpy(value, d_flag, sizeof(uint32_t), cudaMemcpyDefault));
    return B2S_OK;
}

int32_t b2s_ipc_export(b2s_ctx *ctx, void *d_base, uint8_t handle[B2S_IPC_HANDLE_BYTES]) {
    if (!ctx || !d_base || !handle) return b2s_fail(ctx, B2S_EINVAL, "b2s_ipc_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == B2S_IPC_HANDLE_BYTES, "IPC handle size");
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    B2S_CUDA(ctx, cudaIpcGetMemHandle(&h, d_base));
    memcpy(handle, &h, sizeof(h));
    return B2S_OK;
}

int32_t b2s_ipc_open(b2s_ctx *ctx, const uint8_t handle[B2S_IPC_HANDLE_BYTES], void **d_peer) {
    if (!ctx || !handle || !d_peer) return b2s_fail(ctx, B2S_EINVAL, "b2s_ipc_open: NULL argument");
    *d_peer = nullptr;
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    B2S_CUDA(ctx, cudaIpcOpenMemHandle(d_peer, h, cudaIpcMemLazyEnablePeerAccess));
    return B2S_OK;
}

int32_t b2s_ipc_close(b2s_ctx *ctx, void *d_peer) {
    if (!ctx) return b2s_fail(ctx, B2S_EINVAL, "ctx is NULL");
    if (!d_peer) return B2S_OK;
    DeviceGuard g(ctx->device);
    B2S_CUDA(ctx, cudaIpcCloseMemHandle(d_peer));
    return B2S_OK;
}

// same-process peers (one process driving several GPUs): enable direct access ctx -> peer_device
int32_t b2s_peer_enable(b2s_ctx *ctx, int32_t peer_device) {
    if (!ctx) return b2s_fail(ctx, B2S_EINVAL, "ctx is NULL");
    if (peer_device == ctx->device) return B2S_OK;
    DeviceGuard g(ctx->device);
    int can = 0;
    B2S_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) return b2s_fail(ctx, B2S_EUNSUPPORTED, "device %d cannot access device %d", ctx->device, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return B2S_OK; }
    B2S_CUDA(ctx, e);
    return B2S_OK;
}

int32_t b2s_memcpy_d2d(b2s_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return b2s_fail(ctx, B2S_EINVAL, "ctx is NULL");
    DeviceGuard g(ctx->device);
    B2S_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    return B2S_OK;
}

int32_t b2s_memset(b2s_ctx *ctx, void *dst, int32_t byte, size_t bytes) {
    if (!ctx) return b2s_fail(ctx, B2S_EINVAL, "ctx is NULL");
    DeviceGuard g(ctx->device);
    B2S_CUDA(ctx, cudaMemsetAsync(dst, byte, bytes, ctx->stream));
    return B2S_OK;
}

}  // extern "C"
