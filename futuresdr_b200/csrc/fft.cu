// fft.cu -- batched power-of-two Complex<f32> FFT, one pass through shared memory (sm_100a).
//
// Device version of the reference's Fft block (src/blocks/fft.rs:160-221), whose arithmetic is
// rustfft 6.4 (crates.io): forward X[k] = sum_n x[n] e^{-2 pi i kn/N}, inverse un-normalised
// e^{+...}; optional fftshift (after a forward transform fft.rs:196-204, before an inverse one
// :179-185) and optional scalar normalisation (:206-210).
//
// Algorithm: Stockham autosort, mixed radix 16/8/4/2 butterflies held in registers; a transform
// lives in shared memory between passes (padded: idx + idx/16, so the stride-R scatter of the
// first pass is bank-conflict free), the first pass reads global memory and the last pass
// writes it, both fully coalesced (element j + r*N/R for consecutive j).  HBM traffic is the
// algorithmic minimum, 8 B in + 8 B out per sample (the reference's WGSL/CubeCL prior art makes
// one global pass per radix-2 stage, perf/burn/src/bin/fft-wgpu-hack.rs:270-397).
// fftshift is an index rotation on the first-pass load / last-pass store, normalisation a
// multiply on the store; the inverse transform is conj(FFT(conj(x))) (conjugations are free on
// load/store).  Twiddles come from a table W_N[k] evaluated in f64 on the host.
#include <cmath>

#include "common.cuh"
#include "fft_common.cuh"

struct b2s_fft {
    b2s_ctx *ctx = nullptr;
    size_t n = 0;
    int log2n = 0;
    int inverse = 0, shift = 0, has_norm = 0;
    float norm = 1.0f;
    float2 *d_tw = nullptr;     // W_N[k] = exp(-2 pi i k / N), k in [0, N)
};

namespace {

using namespace fftk;

struct FftArgs {
    const float2 *in;
    float2 *out;
    const float2 *tw;
    long long nfft;
    int inverse, shift, has_norm;
    float norm;
};

// One Stockham pass of radix R at sub-transform size NS (NS = product of the radices of the
// earlier passes).  FIRST reads global, LAST writes global, otherwise shared memory `sm`.
// A middle pass is in place in shared memory: every thread first pulls ALL its butterflies'
// inputs into registers, the CTA synchronises, then results are scattered.
template <int N, int R, int NS, bool FIRST, bool LAST, int T>
__device__ __forceinline__ void fft_pass(const FftArgs &a, const float2 *gin, float2 *gout, float2 *sm, int t,
                                         bool active) {
    constexpr int NB = N / R;                      // butterflies per transform
    constexpr int ITER = (NB + T - 1) / T;
    static_assert(NB % T == 0 || ITER == 1, "butterflies must tile the threads");
    float2 v[ITER][R];
#pragma unroll
    for (int it = 0; it < ITER; it++) {
        const int j = t + it * T;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int idx = j + r * NB;
            if constexpr (FIRST) {
                // inverse + shift: buff[k] = i[(k + N/2) % N]   (fft.rs:179-185)
                const int src = (a.inverse && a.shift) ? ((idx + N / 2) & (N - 1)) : idx;
                float2 x = (j < NB) ? __ldg(gin + src) : make_float2(0.f, 0.f);
                if (a.inverse) x.y = -x.y;
                v[it][r] = x;
            } else {
                v[it][r] = (j < NB) ? sm[pad(idx)] : make_float2(0.f, 0.f);
            }
        }
    }
    if constexpr (!FIRST && !LAST) __syncthreads();
#pragma unroll
    for (int it = 0; it < ITER; it++) {
        const int j = t + it * T;
        if constexpr (NS > 1) {
            const int k = j & (NS - 1);
            constexpr int STEP = N / (NS * R);
            apply_twiddles<R>(v[it], __ldg(a.tw + k * STEP));
        }
        Dft<R>::run(v[it]);
        const int j0 = (j / NS) * NS * R + (j & (NS - 1));
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int idx = j0 + r * NS;
            if (j >= NB) continue;
            if constexpr (LAST) {
                float2 y = v[it][r];
                if (a.inverse) y.y = -y.y;
                if (a.has_norm) { y.x *= a.norm; y.y *= a.norm; }
                // forward + shift: o[k] = X[(k + N/2) % N]   (fft.rs:196-204)
                const int dst = (!a.inverse && a.shift) ? ((idx + N / 2) & (N - 1)) : idx;
                if (active) gout[dst] = y;
            } else {
                sm[pad(idx)] = v[it][r];
            }
        }
    }
    if constexpr (!LAST) __syncthreads();
}

template <int LOG2N> struct Plan;   // radices per pass
template <> struct Plan<1>  { static constexpr int P = 1, R0 = 2,  R1 = 1,  R2 = 1, R3 = 1; };
template <> struct Plan<2>  { static constexpr int P = 1, R0 = 4,  R1 = 1,  R2 = 1, R3 = 1; };
template <> struct Plan<3>  { static constexpr int P = 1, R0 = 8,  R1 = 1,  R2 = 1, R3 = 1; };
template <> struct Plan<4>  { static constexpr int P = 1, R0 = 16, R1 = 1,  R2 = 1, R3 = 1; };
template <> struct Plan<5>  { static constexpr int P = 2, R0 = 8,  R1 = 4,  R2 = 1, R3 = 1; };
template <> struct Plan<6>  { static constexpr int P = 2, R0 = 8,  R1 = 8,  R2 = 1, R3 = 1; };
template <> struct Plan<7>  { static constexpr int P = 2, R0 = 16, R1 = 8,  R2 = 1, R3 = 1; };
template <> struct Plan<8>  { static constexpr int P = 2, R0 = 16, R1 = 16, R2 = 1, R3 = 1; };
template <> struct Plan<9>  { static constexpr int P = 3, R0 = 8,  R1 = 8,  R2 = 8, R3 = 1; };
template <> struct Plan<10> { static constexpr int P = 3, R0 = 16, R1 = 8,  R2 = 8, R3 = 1; };
template <> struct Plan<11> { static constexpr int P = 3, R0 = 16, R1 = 16, R2 = 8, R3 = 1; };
template <> struct Plan<12> { static constexpr int P = 3, R0 = 16, R1 = 16, R2 = 16, R3 = 1; };
template <> struct Plan<13> { static constexpr int P = 4, R0 = 16, R1 = 8,  R2 = 8, R3 = 8; };
template <> struct Plan<14> { static constexpr int P = 4, R0 = 16, R1 = 16, R2 = 8, R3 = 8; };

constexpr int kFftThreads = 256;

template <int LOG2N>
__global__ void __launch_bounds__(kFftThreads) fft_kernel(const FftArgs a) {
    constexpr int N = 1 << LOG2N;
    using PL = Plan<LOG2N>;
    constexpr int T = (N / 16 < 1) ? 1 : ((N / 16 > kFftThreads) ? kFftThreads : N / 16);   // threads per transform
    constexpr int FPB = kFftThreads / T;                                                    // transforms per CTA
    constexpr int NP = N + N / 16;                                                          // padded length
    extern __shared__ __align__(16) unsigned char fsm[];
    const int t = threadIdx.x % T, fl = threadIdx.x / T;
    const long long f = (long long)blockIdx.x * FPB + fl;
    const bool active = f < a.nfft;
    float2 *sm = reinterpret_cast<float2 *>(fsm) + (size_t)fl * NP;
    // idle transform slots of the last CTA recompute transform nfft-1 and skip the store
    const long long fc = active ? f : a.nfft - 1;
    const float2 *gin = a.in + fc * N;
    float2 *gout = a.out + fc * N;

    if constexpr (PL::P == 1) {
        fft_pass<N, PL::R0, 1, true, true, T>(a, gin, gout, sm, t, active);
    } else if constexpr (PL::P == 2) {
        fft_pass<N, PL::R0, 1, true, false, T>(a, gin, gout, sm, t, active);
        fft_pass<N, PL::R1, PL::R0, false, true, T>(a, gin, gout, sm, t, active);
    } else if constexpr (PL::P == 3) {
        fft_pass<N, PL::R0, 1, true, false, T>(a, gin, gout, sm, t, active);
        fft_pass<N, PL::R1, PL::R0, false, false, T>(a, gin, gout, sm, t, active);
        fft_pass<N, PL::R2, PL::R0 * PL::R1, false, true, T>(a, gin, gout, sm, t, active);
    } else {
        fft_pass<N, PL::R0, 1, true, false, T>(a, gin, gout, sm, t, active);
        fft_pass<N, PL::R1, PL::R0, false, false, T>(a, gin, gout, sm, t, active);
        fft_pass<N, PL::R2, PL::R0 * PL::R1, false, false, T>(a, gin, gout, sm, t, active);
        fft_pass<N, PL::R3, PL::R0 * PL::R1 * PL::R2, false, true, T>(a, gin, gout, sm, t, active);
    }
}

template <int LOG2N>
int32_t launch_fft(b2s_fft *p, const FftArgs &a, cudaStream_t stream) {
    constexpr int N = 1 << LOG2N;
    constexpr int T = (N / 16 < 1) ? 1 : ((N / 16 > kFftThreads) ? kFftThreads : N / 16);
    constexpr int FPB = kFftThreads / T;
    constexpr size_t smem = (size_t)FPB * (N + N / 16) * sizeof(float2);
    auto kern = fft_kernel<LOG2N>;
    if (smem > 48 * 1024) {
        static thread_local bool set = false;
        if (!set) {
            B2S_CUDA(p->ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set = true;
        }
    }
    const unsigned grid = (unsigned)ceil_div((size_t)a.nfft, (size_t)FPB);
    kern<<<grid, kFftThreads, smem, stream>>>(a);
    B2S_CHECK_LAUNCH(p->ctx);
    return B2S_OK;
}

}  // namespace

extern "C" {

int32_t b2s_fft_plan_c32(b2s_ctx *ctx, size_t n, int32_t inverse, int32_t fft_shift, int32_t has_normalize,
                         float normalize, b2s_fft **out) {
    if (!ctx || !out) return b2s_fail(ctx, B2S_EINVAL, "b2s_fft_plan_c32: NULL argument");
    *out = nullptr;
    if (n < 2) return b2s_fail(ctx, B2S_EINVAL, "b2s_fft_plan_c32: n must be >= 2");
    if (n & (n - 1))
        return b2s_fail(ctx, B2S_EUNSUPPORTED,
                        "b2s_fft_plan_c32: n = %zu is not a power of two (rustfft handles any n; this build 2..16384)", n);
    if (n > 16384) return b2s_fail(ctx, B2S_EUNSUPPORTED, "b2s_fft_plan_c32: n = %zu > 16384", n);
    DeviceGuard g(ctx->device);
    b2s_fft *p = new b2s_fft();
    p->ctx = ctx; p->n = n;
    while (((size_t)1 << p->log2n) < n) p->log2n++;
    p->inverse = inverse != 0; p->shift = fft_shift != 0; p->has_norm = has_normalize != 0; p->norm = normalize;
    std::vector<float2> tw(n);
    const double PI = 3.14159265358979323846264338327950288;
    for (size_t k = 0; k < n; k++) {
        const double ang = -2.0 * PI * (double)k / (double)n;
        tw[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    cudaError_t e = cudaMalloc((void **)&p->d_tw, n * sizeof(float2));
    if (e != cudaSuccess) { delete p; return b2s_fail(ctx, B2S_ENOMEM, "fft twiddles"); }
    B2S_CUDA(ctx, cudaMemcpyAsync(p->d_tw, tw.data(), n * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    B2S_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = p;
    return B2S_OK;
}

void b2s_fft_destroy(b2s_fft *p) {
    if (!p) return;
    DeviceGuard g(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    if (p->d_tw) cudaFree(p->d_tw);
    delete p;
}

size_t b2s_fft_length(const b2s_fft *p) { return p ? p->n : 0; }

int32_t b2s_fft_exec(b2s_fft *p, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                     size_t *consumed, size_t *produced) {
    if (!p || !consumed || !produced) return b2s_fail(p ? p->ctx : nullptr, B2S_EINVAL, "b2s_fft_exec: NULL argument");
    // m = min(i.len(), o.len()) rounded down to a multiple of len (fft.rs:169-170)
    size_t m = n_in < n_out_cap ? n_in : n_out_cap;
    m = (m / p->n) * p->n;
    *consumed = m; *produced = m;
    if (m == 0) return B2S_OK;
    if (!d_in || !d_out) return b2s_fail(p->ctx, B2S_EINVAL, "b2s_fft_exec: NULL buffer");
    if (d_in == d_out) return b2s_fail(p->ctx, B2S_EINVAL, "b2s_fft_exec: in-place is not supported");
    DeviceGuard g(p->ctx->device);
    FftArgs a;
    a.in = (const float2 *)d_in; a.out = (float2 *)d_out; a.tw = p->d_tw; a.nfft = (long long)(m / p->n);
    a.inverse = p->inverse; a.shift = p->shift; a.has_norm = p->has_norm; a.norm = p->norm;
    cudaStream_t s = p->ctx->stream;
    switch (p->log2n) {
        case 1: return launch_fft<1>(p, a, s);
        case 2: return launch_fft<2>(p, a, s);
        case 3: return launch_fft<3>(p, a, s);
        case 4: return launch_fft<4>(p, a, s);
        case 5: return launch_fft<5>(p, a, s);
        case 6: return launch_fft<6>(p, a, s);
        case 7: return launch_fft<7>(p, a, s);
        case 8: return launch_fft<8>(p, a, s);
        case 9: return launch_fft<9>(p, a, s);
        case 10: return launch_fft<10>(p, a, s);
        case 11: return launch_fft<11>(p, a, s);
        case 12: return launch_fft<12>(p, a, s);
        case 13: return launch_fft<13>(p, a, s);
        case 14: return launch_fft<14>(p, a, s);
    }
    return b2s_fail(p->ctx, B2S_EUNSUPPORTED, "b2s_fft_exec: unsupported size");
}

}  // extern "C"
