// synth.cu -- polyphase synthesizer (src/blocks/pfb/synthesizer.rs:52-144), the dual of chan.cu:
// SURVEY.md §8f row 2.
//
// Reference: per input vector v (one sample from each of the N input streams) an un-normalised
// N-point inverse FFT "spins" the vector, element w is pushed into window w, and once the windows
// are filled arm w (taps[w::N]) filters window w into the next output item -- N outputs per vector.
// Device form: (1) gather the channel-major inputs into vectors and run the batched inverse FFT of
// fft.cu; (2) one thread per (vector, arm) dots the T newest spun samples of its slot (history
// buffer + this call) with its arm and writes out[(v - v0) * N + w] -- coalesced in w.
// All windows move in lockstep, so the WindowBuffer bookkeeping (including the scattered start-up
// order of window_buffer.rs:24-32) is two host scalars; the loop condition of :95-97
// (`out.len() - produced > N || !all_windows_filled`) is evaluated in closed form.
#include <cmath>

#include "common.cuh"

struct b2s_synth {
    b2s_ctx *ctx = nullptr;
    size_t N = 0, T = 0;
    float *d_arms = nullptr;        // [T][N] tap-major: d_arms[j*N + w] = arm_w[j] = taps[w + j*N]; newest sample <-> j = 0
    float2 *d_circ = nullptr;       // [N][T] window positions while filling
    float2 *d_hist = nullptr;       // [N][T] FIFO order once filled
    size_t start_idx = 0, missing = 0;
    bool all_filled = false;
    b2s_fft *ifft = nullptr;
    float2 *d_tmp = nullptr;        // 2 * tmp_items: gathered vectors, spun vectors
    size_t tmp_items = 0;
};

namespace {

// vec[v * N + w] = in[w * stride + v]
__global__ void synth_gather_kernel(const float2 *__restrict__ in, float2 *__restrict__ vec, int N, long long nv,
                                    long long stride) {
    __shared__ float2 tile[32][33];
    const long long v0 = (long long)blockIdx.x * 32;
    const int w0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int w = w0 + i; const long long v = v0 + threadIdx.x;
        if (w < N && v < nv) tile[i][threadIdx.x] = in[(long long)w * stride + v];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long v = v0 + i; const int w = w0 + threadIdx.x;
        if (w < N && v < nv) vec[v * N + w] = tile[threadIdx.x][i];
    }
}

// replay of the fill pushes: window position of push #k is pos[k] (same for every window)
__global__ void synth_fill_kernel(const float2 *__restrict__ spun, float2 *circ, int N, int T, int start_idx,
                                  int missing, int count) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= N) return;
    for (int k = 0; k < count; k++) {
        int idx = (start_idx - missing) % T;
        if (idx < 0) idx += T;
        circ[(size_t)w * T + idx] = spun[(size_t)k * N + w];
        if (missing > 0) missing--;
        start_idx = (start_idx + 1) % T;
    }
}

__global__ void synth_hist_from_circ(const float2 *__restrict__ circ, float2 *hist, int N, int T, int start_idx) {
    const int w = blockIdx.x;
    for (int t = threadIdx.x; t < T; t += blockDim.x) hist[(size_t)w * T + t] = circ[(size_t)w * T + (start_idx + t) % T];
}

// outputs for steady vectors u in [u0, k2): u = -1 is the vector that completed the fill (history only)
__global__ void synth_bank_kernel(const float2 *__restrict__ spun /* steady vectors, u = 0 first */,
                                  const float2 *__restrict__ hist, const float *__restrict__ arms,
                                  float2 *__restrict__ out, int N, int T, int u0, long long k2) {
    const long long total = (k2 - u0) * N;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long u = u0 + g / N;
        const int w = (int)(g % N);
        const float *a = arms + w;                                // tap-major table: arm w, tap j at a[j * N] (coalesced across w)
        float re = 0.f, im = 0.f;
        for (int j = T - 1; j >= 0; j--) {                       // oldest first, like the reference's t = 0..T-1
            const long long up = u - j;
            const float2 x = up >= 0 ? __ldg(spun + up * N + w) : hist[(size_t)w * T + (T + up)];
            const float tap = __ldg(a + (size_t)j * N);
            re = fmaf(x.x, tap, re); im = fmaf(x.y, tap, im);
        }
        out[g] = make_float2(re, im);
    }
}

__global__ void synth_hist_update(float2 *hist, const float2 *__restrict__ spun, int N, int T, long long k2) {
    extern __shared__ float2 tmp[];
    const int w = blockIdx.x;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const long long up = k2 - T + t;                          // new hist[t] = S(k2 - T + t)
        tmp[t] = up >= 0 ? spun[up * N + w] : hist[(size_t)w * T + (T + up)];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) hist[(size_t)w * T + t] = tmp[t];
}

}  // namespace

extern "C" {

void b2s_synth_destroy(b2s_synth *s);

int32_t b2s_synth_plan_c32(b2s_ctx *ctx, size_t num_channels, const float *taps, size_t ntaps, b2s_synth **out) {
    if (!ctx || !out || !taps) return b2s_fail(ctx, B2S_EINVAL, "b2s_synth_plan_c32: NULL argument");
    *out = nullptr;
    if (num_channels < 2 || ntaps == 0) return b2s_fail(ctx, B2S_EINVAL, "b2s_synth_plan_c32: need >= 2 channels and taps");
    DeviceGuard g(ctx->device);
    b2s_synth *s = new b2s_synth();
    s->ctx = ctx; s->N = num_channels;
    const size_t N = s->N, T = (size_t)std::ceil((float)ntaps / (float)N);     // utilities.rs:9
    s->T = T; s->missing = T;
    std::vector<float> arms(N * T, 0.0f);
    for (size_t i = 0; i < N; i++) { size_t j = 0; for (size_t idx = i; idx < ntaps; idx += N) arms[(j++) * N + i] = taps[idx]; }
    int32_t rc = b2s_fft_plan_c32(ctx, N, 1, 0, 0, 1.0f, &s->ifft);            // plan_fft(n, Inverse) (synthesizer.rs:65)
    if (rc != B2S_OK) { delete s; return rc; }
    if (cudaMalloc((void **)&s->d_arms, arms.size() * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void **)&s->d_circ, N * T * sizeof(float2)) != cudaSuccess ||
        cudaMalloc((void **)&s->d_hist, N * T * sizeof(float2)) != cudaSuccess) {
        b2s_synth_destroy(s);
        return b2s_fail(ctx, B2S_ENOMEM, "synthesizer buffers");
    }
    B2S_CUDA(ctx, cudaMemcpyAsync(s->d_arms, arms.data(), arms.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    B2S_CUDA(ctx, cudaMemsetAsync(s->d_circ, 0, N * T * sizeof(float2), ctx->stream));
    B2S_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = s;
    return B2S_OK;
}

void b2s_synth_destroy(b2s_synth *s) {
    if (!s) return;
    DeviceGuard g(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->ifft) b2s_fft_destroy(s->ifft);
    if (s->d_arms) cudaFree(s->d_arms);
    if (s->d_circ) cudaFree(s->d_circ);
    if (s->d_hist) cudaFree(s->d_hist);
    if (s->d_tmp) cudaFree(s->d_tmp);
    delete s;
}

// One Kernel::work call (synthesizer.rs:80-144).  d_in is channel-major (stream w at d_in + w*in_stride),
// n_in the shortest input slice, d_out the single output slice of n_out_cap items.
int32_t b2s_synth_exec(b2s_synth *s, const void *d_in, size_t in_stride, size_t n_in, void *d_out, size_t n_out_cap,
                       size_t *consumed_per_channel, size_t *produced) {
    if (!s || !consumed_per_channel || !produced) return b2s_fail(s ? s->ctx : nullptr, B2S_EINVAL, "b2s_synth_exec: NULL argument");
    b2s_ctx *ctx = s->ctx;
    *consumed_per_channel = 0; *produced = 0;
    const size_t N = s->N, T = s->T;
    // closed form of `while n_in - c > 0 && (cap - p > N || !all_filled)`
    size_t k1 = 0, k2 = 0, p = 0;
    bool completes = false;
    if (!s->all_filled) {
        k1 = std::min(n_in, s->missing);
        completes = (k1 == s->missing) && k1 > 0;
        if (completes) p = N;                       // the completing vector writes its N outputs unconditionally
    }
    if (s->all_filled || completes) {
        const size_t remaining = n_in - k1;
        if (n_out_cap > p + N) {
            const size_t room = n_out_cap - N - p;  // vectors while cap - p > N  <=>  p < cap - N
            k2 = std::min(remaining, ceil_div(room, N));
        }
        p += k2 * N;
    }
    if (completes && n_out_cap < N)
        return b2s_fail(ctx, B2S_EINVAL, "b2s_synth_exec: output slice (%zu) shorter than num_channels while the windows fill (the reference would index out of bounds)", n_out_cap);
    const size_t nv = k1 + k2;
    if (nv == 0) return B2S_OK;
    if (!d_in || (!d_out && p)) return b2s_fail(ctx, B2S_EINVAL, "b2s_synth_exec: NULL buffer");
    DeviceGuard g(ctx->device);
    NvtxRange nvtx("b2s_synth_exec");
    const size_t items = nv * N;
    if (s->tmp_items < items) {
        B2S_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (s->d_tmp) cudaFree(s->d_tmp);
        s->tmp_items = items * 5 / 4 + 1024;
        if (cudaMalloc((void **)&s->d_tmp, 2 * s->tmp_items * sizeof(float2)) != cudaSuccess) {
            s->d_tmp = nullptr; s->tmp_items = 0; cudaGetLastError();
            return b2s_fail(ctx, B2S_ENOMEM, "synthesizer workspace");
        }
    }
    float2 *vec = s->d_tmp, *spun = s->d_tmp + s->tmp_items;
    dim3 gg((unsigned)ceil_div(nv, (size_t)32), (unsigned)ceil_div(N, (size_t)32));
    synth_gather_kernel<<<gg, dim3(32, 8), 0, ctx->stream>>>((const float2 *)d_in, vec, (int)N, (long long)nv, (long long)in_stride);
    B2S_CHECK_LAUNCH(ctx);
    size_t fc = 0, fp = 0;
    int32_t rc = b2s_fft_exec(s->ifft, vec, items, spun, items, &fc, &fp);
    if (rc != B2S_OK) return rc;
    if (k1) {
        synth_fill_kernel<<<(unsigned)ceil_div(N, (size_t)128), 128, 0, ctx->stream>>>(spun, s->d_circ, (int)N, (int)T,
                                                                                         (int)s->start_idx, (int)s->missing, (int)k1);
        B2S_CHECK_LAUNCH(ctx);
        s->missing -= k1;
        s->start_idx = (s->start_idx + k1) % T;
        if (completes) {
            synth_hist_from_circ<<<(unsigned)N, 64, 0, ctx->stream>>>(s->d_circ, s->d_hist, (int)N, (int)T, (int)s->start_idx);
            B2S_CHECK_LAUNCH(ctx);
            s->all_filled = true;
        }
    }
    if (p) {
        const int u0 = completes ? -1 : 0;
        const size_t total = (k2 - (long long)u0) * N;
        const unsigned grid = (unsigned)std::min<size_t>(ceil_div(total, (size_t)256), (size_t)ctx->sm_count * 32);
        synth_bank_kernel<<<grid, 256, 0, ctx->stream>>>(spun + k1 * N, s->d_hist, s->d_arms, (float2 *)d_out, (int)N, (int)T,
                                                         u0, (long long)k2);
        B2S_CHECK_LAUNCH(ctx);
    }
    if (k2) {
        synth_hist_update<<<(unsigned)N, 64, T * sizeof(float2), ctx->stream>>>(s->d_hist, spun + k1 * N, (int)N, (int)T, (long long)k2);
        B2S_CHECK_LAUNCH(ctx);
    }
    *consumed_per_channel = nv; *produced = p;
    return B2S_OK;
}

}  // extern "C"
