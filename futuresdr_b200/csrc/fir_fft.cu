// fir_fft.cu -- overlap-save FFT convolution for LONG filters (ntaps > 257) on Complex<f32>
// streams, real or complex taps, no decimation.
//
// Same result as crates/futuredsp/src/fir.rs:77-88,  o[k] = sum_t i[k+t] * taps[N-1-t], evaluated
// per block of NF = 4096 inputs as  y = IFFT( FFT(x_block) . H ),  H[f] = (1/NF) sum_t g[t]
// e^{+2 pi i f t / NF}  (g[t] = taps[N-1-t]; the correlation theorem), keeping the V = NF-(N-1)
// outputs that do not wrap.  One CTA owns one block: forward transform, spectrum product and
// inverse transform all happen in ONE kernel with the block resident in (padded) shared memory
// -- HBM sees 8 B in (x NF/V overlap, mostly L2 hits) + 8 B out per sample, instead of the
// 4*N FLOP per sample of the direct form (4096 FLOP/sample at 1024 taps: 15 Gsamples/s on CUDA
// cores).  Radix-16 Stockham passes from fft_common.cuh; H is computed in f64 on the host.
// Parity: |err| <~ 1e-6 * rms(y) * sqrt(log2 NF), far inside 1e-5 * ||taps||_1 * max|x|.
#include <cmath>
#include <cstdlib>

#include "fft_common.cuh"
#include "fir.cuh"

using namespace fftk;

namespace {

constexpr int kLog2NF = 12;
constexpr int kNF = 1 << kLog2NF;
constexpr int kFfThreads = 256;

struct FftFirArgs {
    const float2 *in;
    float2 *out;
    const float2 *H;    // [NF]
    const float2 *tw;   // W_NF[k] = exp(-2 pi i k / NF)
    long long n_in, n_out;
    int V;              // valid outputs per block
};

// MINB = CTAs per SM the register allocation must allow: the unconstrained build took 171 registers = ONE 256-thread
// CTA per SM (ncu: 12 % of the warp slots active); 128 registers (no spills) give two, 80 (288 B of spills) three.
template <int MINB>
__global__ void __launch_bounds__(kFfThreads, MINB) fir_fft_kernel(const FftFirArgs a) {
    constexpr int N = kNF, T = kFfThreads;
    extern __shared__ __align__(16) unsigned char ffsm[];
    float2 *sm = reinterpret_cast<float2 *>(ffsm);
    const int t = threadIdx.x;
    const long long s = (long long)blockIdx.x * a.V;
    const float2 *in = a.in + s;
    const long long avail = a.n_in - s;                 // items readable from `in`

    auto ld_sm = [&](int idx) { return sm[pad(idx)]; };
    auto st_sm = [&](int idx, float2 v) { sm[pad(idx)] = v; };
    // one butterfly per thread and pass: the base twiddle of a pass depends on the thread alone, so it is fetched one
    // pass AHEAD (the load is in flight across the barrier instead of being waited for right after it)
    static_assert(N / 16 == T, "one radix-16 butterfly per thread");
    const float2 *tw2p = a.tw + (t & 15) * 16, *tw3p = a.tw + t;
    float2 w = __ldg(tw2p);
    // forward: 16 x 16 x 16
    ss_pass<N, 16, 1, T>([&](int idx) { return idx < avail ? __ldg(in + idx) : make_float2(0.f, 0.f); }, st_sm, a.tw, t, false);
    TwPre twa{w};
    w = __ldg(tw3p);
    ss_pass_tw<N, 16, 16, T>(ld_sm, st_sm, twa, t, true);
    TwPre twb{w};
    ss_pass_tw<N, 16, 256, T>(ld_sm, st_sm, twb, t, true);
    // inverse = conj(FFT(conj(X . H))): spectrum product + conjugation fused into the first load
    w = __ldg(tw2p);
    ss_pass<N, 16, 1, T>([&](int idx) {
        const float2 y = cmul(sm[pad(idx)], __ldg(a.H + idx));
        return make_float2(y.x, -y.y);
    }, st_sm, a.tw, t, true);
    TwPre twc{w};
    w = __ldg(tw3p);
    ss_pass_tw<N, 16, 16, T>(ld_sm, st_sm, twc, t, true);
    const TwPre tw3{w};
    const long long room = a.n_out - s;
    float2 *out = a.out + s;
    const int V = a.V;
    ss_pass_tw<N, 16, 256, T>(ld_sm, [&](int idx, float2 v) {
        if (idx < V && idx < room) out[idx] = make_float2(v.x, -v.y);
    }, tw3, t, true);
}

}  // namespace

bool fir_fft_supported(const b2s_fir *f) {
    if (f->decim != 1) return false;
    if (f->kind != B2S_C32_F32 && f->kind != B2S_C32_C32) return false;
    return f->ntaps >= 64 && f->ntaps <= kNF / 2 + 1;      // V >= NF/2
}

int32_t fir_fft_prepare(b2s_fir *f) {
    b2s_ctx *ctx = f->ctx;
    if (f->fft_ready) return B2S_OK;
    if (!fir_fft_supported(f)) return b2s_fail(ctx, B2S_EUNSUPPORTED, "FFT FIR: unsupported plan");
    const size_t N = f->ntaps;
    const bool ctap = f->kind == B2S_C32_C32;
    const double PI = 3.14159265358979323846264338327950288;
    std::vector<float2> H(kNF), tw(kNF);
    // H[f] = (1/NF) sum_t g[t] e^{+2 pi i f t / NF},  g[t] = taps[N-1-t]
    std::vector<double> cs(kNF), sn(kNF);
    for (int k = 0; k < kNF; k++) {
        const double ang = 2.0 * PI * (double)k / (double)kNF;
        cs[k] = std::cos(ang); sn[k] = std::sin(ang);
        tw[k] = make_float2((float)cs[k], (float)-sn[k]);
    }
    for (int fr = 0; fr < kNF; fr++) {
        double re = 0.0, im = 0.0;
        for (size_t t = 0; t < N; t++) {
            const size_t src = N - 1 - t;
            const double gr = ctap ? f->taps_host[2 * src] : f->taps_host[src];
            const double gi = ctap ? f->taps_host[2 * src + 1] : 0.0;
            const int k = (int)(((size_t)fr * t) & (kNF - 1));
            re += gr * cs[k] - gi * sn[k];
            im += gr * sn[k] + gi * cs[k];
        }
        H[fr] = make_float2((float)(re / kNF), (float)(im / kNF));
    }
    B2S_CUDA(ctx, cudaMalloc((void **)&f->d_fftH, 2 * kNF * sizeof(float2)));
    B2S_CUDA(ctx, cudaMemcpyAsync(f->d_fftH, H.data(), kNF * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    B2S_CUDA(ctx, cudaMemcpyAsync(f->d_fftH + kNF, tw.data(), kNF * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    B2S_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    f->fft_ready = true;
    return B2S_OK;
}

void fir_fft_release(b2s_fir *f) {
    if (f->d_fftH) cudaFree(f->d_fftH);
    f->d_fftH = nullptr;
    f->fft_ready = false;
}

int32_t fir_fft_launch(b2s_fir *f, const void *d_in, size_t n_in, void *d_out, size_t n_out, cudaStream_t stream) {
    b2s_ctx *ctx = f->ctx;
    if (n_out == 0) return B2S_OK;
    if (!f->fft_ready) return b2s_fail(ctx, B2S_ESTATE, "FFT FIR not prepared");
    if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 7)
        return fir_direct_launch(f, d_in, n_in, d_out, n_out, stream);
    FftFirArgs a;
    a.in = (const float2 *)d_in; a.out = (float2 *)d_out;
    a.H = f->d_fftH; a.tw = f->d_fftH + kNF;
    a.n_in = (long long)n_in; a.n_out = (long long)n_out;
    a.V = kNF - (int)(f->ntaps - 1);
    const unsigned grid = (unsigned)ceil_div(n_out, (size_t)a.V);
    const size_t smem = (size_t)(kNF + kNF / 16) * sizeof(float2);
    static const int minb = [] { const char *e = getenv("B2S_FFTFIR_MINB"); const int v = e ? atoi(e) : 2; return v >= 1 && v <= 3 ? v : 2; }();
    if (minb == 1) fir_fft_kernel<1><<<grid, kFfThreads, smem, stream>>>(a);
    else if (minb == 3) fir_fft_kernel<3><<<grid, kFfThreads, smem, stream>>>(a);
    else fir_fft_kernel<2><<<grid, kFfThreads, smem, stream>>>(a);
    B2S_CHECK_LAUNCH(ctx);
    return B2S_OK;
}
