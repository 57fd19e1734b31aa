// mavg.cu -- MovingAvg<WIDTH> (src/blocks/moving_avg.rs:24-116), the tail of the spectrum pipe
// (examples/spectrum/src/bin/cpu.rs:21-28: Fft(shift) -> |x|^2 -> MovingAvg -> sink): SURVEY §8f-3.
//
// Per bin i an exponential average over consecutive WIDTH-item chunks:
//     avg[i] = (1 - decay) * avg[i] + decay * t        (t finite)      (:87)
//     avg[i] *= 1 - decay                              (t not finite)  (:89)
// and every `history_size` chunks the WIDTH averages are emitted (:95-99).  The recurrence is
// sequential over chunks but independent across bins: one thread per bin walks the chunks in order
// with un-fused IEEE multiplies/adds (identical to the reference bit for bit), reads are coalesced
// across bins.  The state (avg[], chunk counter) lives on the device between calls.
#include <cmath>

#include "common.cuh"

struct b2s_mavg {
    b2s_ctx *ctx = nullptr;
    size_t width = 0, history = 1;
    float decay = 0.1f;
    float *d_avg = nullptr;
    size_t i = 0;                 // chunks since the last emission (host mirror; data-independent)
};

namespace {

__global__ void mavg_kernel(const float *__restrict__ in, float *__restrict__ out, float *avg, int width,
                            long long nchunks, int history, int i0, float decay, long long max_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= width) return;
    float a = avg[b];
    const float keep = __fsub_rn(1.0f, decay);
    int i = i0;
    long long produced = 0;
    for (long long c = 0; c < nchunks; c++) {
        const float t = __ldg(in + c * width + b);
        if (isfinite(t)) a = __fadd_rn(__fmul_rn(keep, a), __fmul_rn(decay, t));
        else a = __fmul_rn(a, keep);
        if (++i == history) {
            if (produced < max_out) out[produced * width + b] = a;
            produced++;
            i = 0;
        }
    }
    avg[b] = a;
}

}  // namespace

extern "C" {

int32_t b2s_mavg_create(b2s_ctx *ctx, size_t width, float decay_factor, size_t history_size, b2s_mavg **out) {
    if (!ctx || !out) return b2s_fail(ctx, B2S_EINVAL, "b2s_mavg_create: NULL argument");
    *out = nullptr;
    if (width == 0 || history_size == 0) return b2s_fail(ctx, B2S_EINVAL, "b2s_mavg_create: width and history_size must be > 0");
    if (!(decay_factor >= 0.0f && decay_factor <= 1.0f))                        // moving_avg.rs:58-61
        return b2s_fail(ctx, B2S_EINVAL, "decay_factor must be in [0, 1]");
    DeviceGuard g(ctx->device);
    b2s_mavg *m = new b2s_mavg();
    m->ctx = ctx; m->width = width; m->history = history_size; m->decay = decay_factor;
    if (cudaMalloc((void **)&m->d_avg, width * sizeof(float)) != cudaSuccess) { delete m; return b2s_fail(ctx, B2S_ENOMEM, "mavg state"); }
    B2S_CUDA(ctx, cudaMemsetAsync(m->d_avg, 0, width * sizeof(float), ctx->stream));
    *out = m;
    return B2S_OK;
}

void b2s_mavg_destroy(b2s_mavg *m) {
    if (!m) return;
    DeviceGuard g(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    if (m->d_avg) cudaFree(m->d_avg);
    delete m;
}

// One Kernel::work call (moving_avg.rs:72-115).  consumed / produced are in ITEMS (multiples of WIDTH).
int32_t b2s_mavg_exec(b2s_mavg *m, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                      size_t *consumed, size_t *produced) {
    if (!m || !consumed || !produced) return b2s_fail(m ? m->ctx : nullptr, B2S_EINVAL, "b2s_mavg_exec: NULL argument");
    const size_t W = m->width;
    // while (consumed+1)*W <= in.len() && (produced+1)*W <= out.len()   (:82)
    size_t c = 0, p = 0, i = m->i;
    const size_t cin = n_in / W, cout = n_out_cap / W;
    while (c + 1 <= cin && p + 1 <= cout) {
        // jump to the next emission (or to the end of the input)
        const size_t to_emit = m->history - i;
        if (c + to_emit <= cin) { c += to_emit; p += 1; i = 0; }
        else { i += cin - c; c = cin; }
    }
    *consumed = c * W; *produced = p * W;
    if (c == 0) return B2S_OK;
    if (!d_in || (!d_out && p)) return b2s_fail(m->ctx, B2S_EINVAL, "b2s_mavg_exec: NULL buffer");
    DeviceGuard g(m->ctx->device);
    const int th = 128;
    mavg_kernel<<<(unsigned)ceil_div(W, (size_t)th), th, 0, m->ctx->stream>>>(
        (const float *)d_in, (float *)d_out, m->d_avg, (int)W, (long long)c, (int)m->history, (int)m->i, m->decay,
        (long long)p);
    B2S_CHECK_LAUNCH(m->ctx);
    m->i = i;
    return B2S_OK;
}

}  // extern "C"
