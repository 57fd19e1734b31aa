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
#include <cstdlib>

#include "common.cuh"
#include "fft_common.cuh"

const float2 *b2s_fft_twiddles(const b2s_fft *p);   // fft.cu
int b2s_fft_log2n(const b2s_fft *p);

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
    float *d_arms_pad = nullptr;    // [TPAD][N]: d_arms zero-padded to the fused kernel's tap count
    int tpad = 0;
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

// ---------------------------------------------------------------------------------------------------------------
// FUSED steady state (N a power of two <= 256, T <= 32): gather + N-point inverse FFT + FIR bank in ONE kernel --
// 8 B/sample in, 8 B/sample out, the spun vectors never touch HBM (the three kernels above move 40 B/sample).  A CTA
// owns a contiguous range of tiles of OB = 4096/N vectors and keeps the last TPAD-1 spun vectors of the previous tile
// in a shared-memory ring (its first tile is preceded by a warm-up tile that only fills the ring):
//   A. the tile's OB vectors are read channel-major (OB consecutive items per input stream: coalesced) into a staging
//      array with an odd pitch;
//   B. every vector is spun by the Stockham passes of fft_common.cuh (inverse = conj o FFT o conj), the first pass
//      reading the staging array transposed, the last one writing row (TPAD-1+v) of the ring;
//   C. thread (arm w, run of RL vectors) streams column w of the ring through registers once -- each spun sample is
//      multiplied into every output of the run that contains it (taps in registers, oldest first like
//      synthesizer.rs:106-118) -- and stores out[u*N + w], coalesced in w;
//   D. the last TPAD-1 rows move to the top of the ring; the CTA that owns the last tile leaves the T newest spun
//      vectors in the plan's history buffer for the next call.
// The first vectors of a call (windows still reaching into the previous call's history; one tile, more when a tile is
// shorter than the history), the window fill and other bank shapes take the three-kernel path.
// ---------------------------------------------------------------------------------------------------------------
template <int LOG2N, int TPAD>
__global__ void __launch_bounds__(256) synth_fused_kernel(const float2 *__restrict__ in, long long in_stride,
                                                          const float *__restrict__ arms_pad, const float2 *__restrict__ tw,
                                                          float2 *__restrict__ out, float2 *__restrict__ hist, int T,
                                                          long long k2, int ntiles, int tiles_per_cta) {
    using namespace fftk;
    constexpr int N = 1 << LOG2N;
    constexpr int TT = (N / 16 < 1) ? 1 : N / 16;            // threads per transform
    constexpr int OB = 256 / TT;                             // vectors per tile
    constexpr int RUNS = 256 / N, RL = OB / RUNS;
    constexpr int NP = N + N / 16;
    constexpr int XP = OB + 1;                               // odd pitch of the gather staging
    constexpr int WARM = (TPAD - 1 + OB - 1) / OB;           // warm-up tiles that fill TPAD-1 rows of history (1 unless OB < TPAD-1)
    extern __shared__ __align__(16) unsigned char ysm[];
    float2 *Xs = reinterpret_cast<float2 *>(ysm);            // [N][XP]
    float2 *Vf = Xs + (size_t)N * XP;                        // [OB][NP]
    float2 *Sb = Vf + (size_t)OB * NP;                       // [TPAD-1+OB][N]  spun vectors, oldest row first
    const int tid = threadIdx.x;
    const int w = tid % N, run = tid / N;
    unsigned long long tap[TPAD];            // (t, t) pairs: one FFMA2 per complex x real MAC (common.cuh cmac2)
#pragma unroll
    for (int j = 0; j < TPAD; j++) tap[j] = dup2(__ldg(arms_pad + (size_t)j * N + w));

    const int t0 = blockIdx.x * tiles_per_cta, t1 = min(t0 + tiles_per_cta, ntiles);
    for (int t = t0 - WARM; t < t1; t++) {                   // t < t0: warm-up tiles (fill the ring, emit nothing)
        const long long v0 = (long long)OB * (t + WARM);     // tile t covers vectors [OB*(t+WARM), +OB); vectors < OB*WARM: generic path
        // ---- A: gather
        for (int e = tid; e < N * OB; e += 256) {
            const int ch = e / OB, v = e % OB;
            Xs[(size_t)ch * XP + v] = (v0 + v < k2) ? __ldg(in + (long long)ch * in_stride + v0 + v) : make_float2(0.f, 0.f);
        }
        __syncthreads();
        // ---- B: spin
        {
            const int ol = tid / TT, tt = tid % TT;
            float2 *sm = Vf + (size_t)ol * NP;
            fft_passes<LOG2N, TT>([&](int idx) { const float2 x = Xs[(size_t)idx * XP + ol]; return make_float2(x.x, -x.y); },
                                  [&](int idx, float2 v) { Sb[(size_t)(TPAD - 1 + ol) * N + idx] = make_float2(v.x, -v.y); },
                                  sm, tw, tt, false);
        }
        // (fft_passes ends with a CTA barrier)
        if (t >= t0) {
            // ---- C: FIR bank
            float2 acc[RL];
#pragma unroll
            for (int u = 0; u < RL; u++) acc[u] = make_float2(0.f, 0.f);
            const float2 *col = Sb + (size_t)(run * RL) * N + w;   // ring row (run*RL + k) <-> vector v0 + run*RL + k - (TPAD-1)
#pragma unroll
            for (int k = 0; k < RL + TPAD - 1; k++) {
                const float2 x = col[(size_t)k * N];
#pragma unroll
                for (int u = 0; u < RL; u++) {
                    const int j = u + TPAD - 1 - k;          // output u sees this row as its j-th newest spun sample
                    if (j >= 0 && j < TPAD) cmac2(acc[u], x, tap[j]);
                }
            }
#pragma unroll
            for (int u = 0; u < RL; u++) {
                const long long uu = v0 + run * RL + u;
                if (uu < k2) out[uu * N + w] = acc[u];
            }
            if (t == ntiles - 1) {                            // the T newest spun vectors of the call -> history of the next one
                for (int e = tid; e < T * N; e += 256) {
                    const int tt = e / N, ch = e % N;
                    const long long vv = k2 - T + tt;         // >= v0 - (TPAD-1): the last tile holds vector k2-1 and T <= TPAD
                    hist[(size_t)ch * T + tt] = Sb[(size_t)(vv - v0 + TPAD - 1) * N + ch];
                }
            }
        }
        __syncthreads();
        // ---- D: keep the newest TPAD-1 rows for the next tile (through registers: source and destination overlap
        // when a tile is shorter than the history, e.g. 256 channels x 32 taps)
        {
            constexpr int ITER = ((TPAD - 1) * N + 255) / 256;
            float2 keep[ITER];
#pragma unroll
            for (int i = 0; i < ITER; i++) {
                const int e = tid + i * 256;
                keep[i] = e < (TPAD - 1) * N ? Sb[(size_t)OB * N + e] : make_float2(0.f, 0.f);
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < ITER; i++) {
                const int e = tid + i * 256;
                if (e < (TPAD - 1) * N) Sb[e] = keep[i];
            }
        }
        __syncthreads();
    }
}

template <int LOG2N, int TPAD> constexpr size_t synth_fused_smem() {
    constexpr int N = 1 << LOG2N;
    constexpr int TT = (N / 16 < 1) ? 1 : N / 16;
    constexpr int OB = 256 / TT;
    return ((size_t)N * (OB + 1) + (size_t)OB * (N + N / 16) + (size_t)(TPAD - 1 + OB) * N) * sizeof(float2);
}

template <int LOG2N, int TPAD>
int32_t synth_fused_launch(b2s_synth *s, const float2 *in, long long in_stride, float2 *out, long long k2) {
    constexpr int N = 1 << LOG2N;
    constexpr int TT = (N / 16 < 1) ? 1 : N / 16;
    constexpr int OB = 256 / TT;
    constexpr size_t smem = synth_fused_smem<LOG2N, TPAD>();
    auto kern = synth_fused_kernel<LOG2N, TPAD>;
    static PerDeviceOnce optin;
    if (smem > 48 * 1024 && optin.need(s->ctx->device)) {
        B2S_CUDA(s->ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        optin.done(s->ctx->device);
    }
    static int resident = 0;
    if (!resident) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, 256, smem) != cudaSuccess || resident < 1) { cudaGetLastError(); resident = 1; }
    }
    constexpr int WARM = (TPAD - 1 + OB - 1) / OB;
    const long long ntiles = (k2 - (long long)WARM * OB + OB - 1) / OB;   // tiles over vectors [WARM*OB, k2)
    if (ntiles > 0x7fffff00ll) return b2s_fail(s->ctx, B2S_EUNSUPPORTED, "synthesizer: too many vectors in one call");
    const long long grid = std::min<long long>(ntiles, (long long)s->ctx->sm_count * resident);
    const long long tpc = (ntiles + grid - 1) / grid;
    const long long grid2 = (ntiles + tpc - 1) / tpc;         // no empty CTAs (the last tile must be owned by the last CTA)
    kern<<<(unsigned)grid2, 256, smem, s->ctx->stream>>>(in, in_stride, s->d_arms_pad, b2s_fft_twiddles(s->ifft), out, s->d_hist,
                                                          (int)s->T, k2, (int)ntiles, (int)tpc);
    B2S_CHECK_LAUNCH(s->ctx);
    return B2S_OK;
}

template <int TPAD>
int32_t synth_fused_dispatch(b2s_synth *s, int log2n, const float2 *in, long long in_stride, float2 *out, long long k2) {
    switch (log2n) {
        case 2: return synth_fused_launch<2, TPAD>(s, in, in_stride, out, k2);
        case 3: return synth_fused_launch<3, TPAD>(s, in, in_stride, out, k2);
        case 4: return synth_fused_launch<4, TPAD>(s, in, in_stride, out, k2);
        case 5: return synth_fused_launch<5, TPAD>(s, in, in_stride, out, k2);
        case 6: return synth_fused_launch<6, TPAD>(s, in, in_stride, out, k2);
        case 7: return synth_fused_launch<7, TPAD>(s, in, in_stride, out, k2);
        case 8: return synth_fused_launch<8, TPAD>(s, in, in_stride, out, k2);
    }
    return B2S_EAGAIN;
}

int synth_fused_tpad(const b2s_synth *s) {
    const int l2 = b2s_fft_log2n(s->ifft);
    if (getenv("B2S_SYNTH_NO_FUSED")) return 0;
    if (l2 < 2 || l2 > 8 || s->T > 32) return 0;
    return s->T <= 8 ? 8 : (s->T <= 16 ? 16 : 32);
}

int synth_fused_ob(int log2n) { const int n = 1 << log2n; return 256 / ((n / 16 < 1) ? 1 : n / 16); }
// vectors at the start of a call that stay on the generic path: the warm-up tiles of the first CTA
size_t synth_fused_lead(int log2n, int tpad) { const int ob = synth_fused_ob(log2n); return (size_t)((tpad - 1 + ob - 1) / ob) * ob; }

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
    s->tpad = synth_fused_tpad(s);
    std::vector<float> apad;
    if (s->tpad) {
        apad.assign((size_t)s->tpad * N, 0.0f);                                  // taps beyond T are zero (older samples)
        std::copy(arms.begin(), arms.end(), apad.begin());
        if (cudaMalloc((void **)&s->d_arms_pad, apad.size() * sizeof(float)) != cudaSuccess) {
            cudaGetLastError(); b2s_synth_destroy(s); return b2s_fail(ctx, B2S_ENOMEM, "synthesizer buffers");
        }
        B2S_CUDA(ctx, cudaMemcpyAsync(s->d_arms_pad, apad.data(), apad.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    }
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
    if (s->d_arms_pad) cudaFree(s->d_arms_pad);
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
    // steady calls long enough for two tiles: the first OB vectors (their windows reach into the previous call's
    // history) through the three kernels below, everything after them through the fused kernel
    const int l2n = b2s_fft_log2n(s->ifft);
    const size_t ob = s->tpad ? (size_t)synth_fused_ob(l2n) : 0;
    const size_t lead = s->tpad ? synth_fused_lead(l2n, s->tpad) : 0;
    const bool fused = s->tpad && s->all_filled && k1 == 0 && k2 >= lead + ob && k2 >= lead + T;
    const size_t k2_all = k2;
    if (fused) k2 = lead;                                 // the generic part
    const size_t items = (k1 + k2) * N;
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
    const size_t nvg = k1 + k2;                           // vectors of the generic part
    dim3 gg((unsigned)ceil_div(nvg, (size_t)32), (unsigned)ceil_div(N, (size_t)32));
    synth_gather_kernel<<<gg, dim3(32, 8), 0, ctx->stream>>>((const float2 *)d_in, vec, (int)N, (long long)nvg, (long long)in_stride);
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
    if (fused) {
        // (the generic bank kernel above has read the old history; the fused kernel writes the new one)
        int32_t frc = B2S_EAGAIN;
        if (s->tpad == 8) frc = synth_fused_dispatch<8>(s, l2n, (const float2 *)d_in, (long long)in_stride, (float2 *)d_out, (long long)k2_all);
        else if (s->tpad == 16) frc = synth_fused_dispatch<16>(s, l2n, (const float2 *)d_in, (long long)in_stride, (float2 *)d_out, (long long)k2_all);
        else if (s->tpad == 32) frc = synth_fused_dispatch<32>(s, l2n, (const float2 *)d_in, (long long)in_stride, (float2 *)d_out, (long long)k2_all);
        if (frc != B2S_OK) return frc == B2S_EAGAIN ? b2s_fail(ctx, B2S_ESTATE, "synthesizer: fused shape mismatch") : frc;
    } else if (k2) {
        synth_hist_update<<<(unsigned)N, 64, T * sizeof(float2), ctx->stream>>>(s->d_hist, spun + k1 * N, (int)N, (int)T, (long long)k2);
        B2S_CHECK_LAUNCH(ctx);
    }
    *consumed_per_channel = nv; *produced = p;
    return B2S_OK;
}

}  // extern "C"
