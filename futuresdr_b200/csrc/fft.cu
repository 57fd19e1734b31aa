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
#include <cstdlib>

#include "common.cuh"
#include "fft_common.cuh"

struct b2s_fft {
    b2s_ctx *ctx = nullptr;
    size_t n = 0;
    int log2n = 0;
    int inverse = 0, shift = 0, has_norm = 0;
    float norm = 1.0f;
    float2 *d_tw = nullptr;     // W_N[k] = exp(-2 pi i k / N), k in [0, N)  (W_M for Bluestein)
    // Bluestein (chirp-z) path for lengths that are not a power of two
    bool bluestein = false;
    int log2m = 0;              // M = 2^log2m >= 2n - 1
    float2 *d_chirp = nullptr;  // w[k] = exp(-i pi k^2 / n), k in [0, n)
    float2 *d_bhat = nullptr;   // FFT_M of the wrapped conjugate chirp, pre-divided by M
    // LARGE transforms (n > 16384, or Bluestein with M > 16384): four-step through HBM on top of two shared-memory plans
    bool big = false;
    size_t big_m = 0, big_n1 = 0, big_n2 = 0;     // M = n1 * n2 (M = n for powers of two)
    b2s_fft *sub1 = nullptr, *sub2 = nullptr;     // forward n1- and n2-point plans (no shift, no scale)
    float2 *d_work = nullptr;                     // 2 * M (four-step scratch) [+ 2 * M for Bluestein]
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

constexpr int kFftThreads = 256;

// threads per CTA / CTAs per SM the register allocation must allow, per size.  8192: the unconstrained build takes
// 171 registers = one 256-thread CTA per SM although shared memory admits three -> cap at 128 (16 B of spills), two
// CTAs.  16384: one transform fills 139 KiB of shared memory, so one CTA per SM whatever we do -- 512 threads halve
// the butterflies (and registers) per thread and double the warps that hide latency.
// (<= 4096: three CTAs per SM as before -- naming a minimum of 1 let ptxas take 111 registers and cost 15 % at 4096.)
template <int LOG2N> struct FftCfg { static constexpr int THREADS = LOG2N >= 14 ? 512 : 256, MINB = LOG2N >= 14 ? 1 : (LOG2N == 13 ? 2 : 3); };

template <int LOG2N, int TH, int MINB>
__global__ void __launch_bounds__(TH, MINB) fft_kernel(const FftArgs a) {
    constexpr int N = 1 << LOG2N;
    using PL = Plan<LOG2N>;
    constexpr int T = (N / 16 < 1) ? 1 : ((N / 16 > TH) ? TH : N / 16);                     // threads per transform
    constexpr int FPB = TH / T;                                                             // transforms per CTA
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

template <int LOG2N, int TH, int MINB>
int32_t launch_fft_cfg(b2s_fft *p, const FftArgs &a, cudaStream_t stream) {
    constexpr int N = 1 << LOG2N;
    constexpr int T = (N / 16 < 1) ? 1 : ((N / 16 > TH) ? TH : N / 16);
    constexpr int FPB = TH / T;
    constexpr size_t smem = (size_t)FPB * (N + N / 16) * sizeof(float2);
    auto kern = fft_kernel<LOG2N, TH, MINB>;
    if (smem > 48 * 1024) {
        static PerDeviceOnce optin;              // per template instantiation, per device
        if (optin.need(p->ctx->device)) {
            B2S_CUDA(p->ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            optin.done(p->ctx->device);
        }
    }
    const unsigned grid = (unsigned)ceil_div((size_t)a.nfft, (size_t)FPB);
    kern<<<grid, TH, smem, stream>>>(a);
    B2S_CHECK_LAUNCH(p->ctx);
    return B2S_OK;
}

template <int LOG2N>
int32_t launch_fft(b2s_fft *p, const FftArgs &a, cudaStream_t stream) {
    if constexpr (LOG2N == 14) {
        // one transform = one CTA = one SM (139 KiB of shared memory): 1024 threads run ONE butterfly each per pass at
        // 64 registers (32 B of spills) and give the SM 32 warps to hide latency with; 512 threads (two butterflies,
        // 128 registers) is the A/B alternative (B2S_FFT16K_THREADS=512)
        static const int th = [] { const char *e = getenv("B2S_FFT16K_THREADS"); return e ? atoi(e) : 1024; }();
        if (th == 512) return launch_fft_cfg<14, 512, 1>(p, a, stream);
        return launch_fft_cfg<14, 1024, 1>(p, a, stream);
    } else if constexpr (LOG2N == 13) {
        static const int th = [] { const char *e = getenv("B2S_FFT8K_THREADS"); return e ? atoi(e) : 256; }();
        if (th == 512) return launch_fft_cfg<13, 512, 2>(p, a, stream);      // A/B: one butterfly per thread and pass
        return launch_fft_cfg<13, 256, 2>(p, a, stream);
    } else {
        return launch_fft_cfg<LOG2N, FftCfg<LOG2N>::THREADS, FftCfg<LOG2N>::MINB>(p, a, stream);
    }
}


// ---- Bluestein: X[k] = w[k] * sum_n (x[n] w[n]) * conj(w[k-n]),  w[n] = exp(-i pi n^2 / N) ----------
// One transform per thread group: a = x.w zero-padded to M, A = FFT_M(a), C = A . Bhat,
// c = IFFT_M(C), X = c . w -- all inside one kernel with the M-point buffer in shared memory.
struct BsArgs {
    const float2 *in;
    float2 *out;
    const float2 *tw, *chirp, *bhat;
    long long nfft;
    int n, inverse, shift, has_norm;
    float norm;
};

template <int LOG2M>
__global__ void __launch_bounds__(kFftThreads) bluestein_kernel(const BsArgs a) {
    constexpr int M = 1 << LOG2M;
    constexpr int T = (M / 16 < 1) ? 1 : ((M / 16 > kFftThreads) ? kFftThreads : M / 16);
    constexpr int FPB = kFftThreads / T;
    constexpr int MP = M + M / 16;
    extern __shared__ __align__(16) unsigned char fsm[];
    const int t = threadIdx.x % T, fl = threadIdx.x / T;
    const long long f = (long long)blockIdx.x * FPB + fl;
    const bool active = f < a.nfft;
    const long long fc = active ? f : a.nfft - 1;
    float2 *sm = reinterpret_cast<float2 *>(fsm) + (size_t)fl * MP;
    const float2 *gin = a.in + fc * a.n;
    float2 *gout = a.out + fc * a.n;
    const int n = a.n, half = n / 2;
    auto st_sm = [&](int idx, float2 v) { sm[pad(idx)] = v; };
    // forward M-point FFT of a[j] = x'[j] * w[j]  (x' = conj / pre-shifted input for the inverse direction)
    fft_passes<LOG2M, T>(
        [&](int idx) {
            if (idx >= n) return make_float2(0.f, 0.f);
            const int src = (a.inverse && a.shift) ? (idx + half) % n : idx;        // fft.rs:179-185
            float2 x = __ldg(gin + src);
            if (a.inverse) x.y = -x.y;
            return cmul(x, __ldg(a.chirp + idx));
        },
        st_sm, sm, a.tw, t, false);
    // inverse M-point FFT of A . Bhat as conj(FFT(conj(.))), then the post-chirp
    fft_passes<LOG2M, T>(
        [&](int idx) {
            const float2 y = cmul(sm[pad(idx)], __ldg(a.bhat + idx));
            return make_float2(y.x, -y.y);
        },
        [&](int idx, float2 v) {
            if (idx >= n || !active) return;
            float2 y = cmul(make_float2(v.x, -v.y), __ldg(a.chirp + idx));
            if (a.inverse) y.y = -y.y;
            if (a.has_norm) { y.x *= a.norm; y.y *= a.norm; }
            const int dst = (!a.inverse && a.shift) ? (idx + n - half) % n : idx;   // o[k] = X[(k + n/2) % n]  (fft.rs:196-204)
            gout[dst] = y;
        },
        sm, a.tw, t, true);
}

template <int LOG2M>
int32_t launch_bluestein(b2s_fft *p, const BsArgs &a, cudaStream_t stream) {
    constexpr int M = 1 << LOG2M;
    constexpr int T = (M / 16 < 1) ? 1 : ((M / 16 > kFftThreads) ? kFftThreads : M / 16);
    constexpr int FPB = kFftThreads / T;
    constexpr size_t smem = (size_t)FPB * (M + M / 16) * sizeof(float2);
    auto kern = bluestein_kernel<LOG2M>;
    if (smem > 48 * 1024) {
        static PerDeviceOnce optin;              // per template instantiation, per device
        if (optin.need(p->ctx->device)) {
            B2S_CUDA(p->ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            optin.done(p->ctx->device);
        }
    }
    const unsigned grid = (unsigned)ceil_div((size_t)a.nfft, (size_t)FPB);
    kern<<<grid, kFftThreads, smem, stream>>>(a);
    B2S_CHECK_LAUNCH(p->ctx);
    return B2S_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// LARGE transforms: rustfft plans any length (src/blocks/fft.rs:98-103); beyond what one shared-memory transform holds
// the classic four-step algorithm runs through HBM:  M = n1 * n2,  i = i1*n2 + i2,  k = k1 + n1*k2,
//   X[k1 + n1*k2] = sum_{i2} W_{n2}^{i2 k2} * ( W_M^{i2 k1} * sum_{i1} x[i1*n2 + i2] W_{n1}^{i1 k1} )
// as  transpose -> n2 transforms of length n1 -> twiddle + transpose -> n1 transforms of length n2 -> transpose,
// every transform being the batched shared-memory kernel above.  Five passes over the data instead of one: this path
// exists for completeness (spectrum analysers with 64 Ki+ bins, odd lengths), not for speed.
// ---------------------------------------------------------------------------------------------------------------
struct BigT {                 // dst[c * R + r] = f(src[r * C + c]) with optional index rotations / twiddle / scale
    const float2 *src;
    float2 *dst;
    long long R, C, M;
    int conj_in, conj_out, twiddle;
    long long src_rot, dst_rot;   // src index (idx + src_rot) % M ; dst index (idx + dst_rot) % M
    float scale;
};

__global__ void big_transpose_kernel(const BigT a) {
    __shared__ float2 tile[32][33];
    const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long r = r0 + i, c = c0 + threadIdx.x;
        if (r < a.R && c < a.C) {
            long long idx = r * a.C + c;
            if (a.src_rot) { idx += a.src_rot; if (idx >= a.M) idx -= a.M; }
            float2 v = a.src[idx];
            if (a.conj_in) v.y = -v.y;
            if (a.twiddle) {                                  // W_M^{r c}, exponent reduced exactly, angle in f64
                const unsigned long long t = ((unsigned long long)r * (unsigned long long)c) % (unsigned long long)a.M;
                double sn, cs;
                sincospi(-2.0 * (double)t / (double)a.M, &sn, &cs);
                const float wr = (float)cs, wi = (float)sn;
                v = make_float2(v.x * wr - v.y * wi, v.x * wi + v.y * wr);
            }
            tile[i][threadIdx.x] = v;
        }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long c = c0 + i, r = r0 + threadIdx.x;
        if (r < a.R && c < a.C) {
            float2 v = tile[threadIdx.x][i];
            if (a.conj_out) v.y = -v.y;
            v.x *= a.scale; v.y *= a.scale;
            long long idx = c * a.R + r;
            if (a.dst_rot) { idx += a.dst_rot; if (idx >= a.M) idx -= a.M; }
            a.dst[idx] = v;
        }
    }
}

int32_t big_transpose(b2s_ctx *ctx, const BigT &a, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div((size_t)a.C, (size_t)32), (unsigned)ceil_div((size_t)a.R, (size_t)32));
    big_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(a);
    B2S_CHECK_LAUNCH(ctx);
    return B2S_OK;
}

// Bluestein element-wise stages for M > 16384
__global__ void big_bs_pre(const float2 *__restrict__ in, const float2 *__restrict__ chirp, float2 *a, long long n, long long M,
                           int inverse, long long src_rot) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += stride) {
        float2 v = make_float2(0.f, 0.f);
        if (j < n) {
            long long sidx = j + src_rot; if (sidx >= n) sidx -= n;
            float2 x = in[sidx];
            if (inverse) x.y = -x.y;
            v = cmul(x, chirp[j]);
        }
        a[j] = v;
    }
}
__global__ void big_bs_mul(float2 *A, const float2 *__restrict__ bhat, long long M) {   // A <- conj(A . Bhat): input of the inverse M-point FFT
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += stride) {
        const float2 y = cmul(A[j], bhat[j]);
        A[j] = make_float2(y.x, -y.y);
    }
}
__global__ void big_bs_post(const float2 *__restrict__ c, const float2 *__restrict__ chirp, float2 *out, long long n, int inverse,
                            long long dst_rot, float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        float2 y = cmul(make_float2(c[k].x, -c[k].y), chirp[k]);      // conj closes the inverse M-point FFT
        if (inverse) y.y = -y.y;
        y.x *= scale; y.y *= scale;
        long long d = k + dst_rot; if (d >= n) d -= n;
        out[d] = y;
    }
}

// plan internals for the kernels that embed an N-point transform (chan.cu's fused channelizer)
const float2 *b2s_fft_twiddles(const b2s_fft *p) { return p ? p->d_tw : nullptr; }
int b2s_fft_log2n(const b2s_fft *p) { return (p && !p->bluestein && !p->big) ? p->log2n : -1; }

// one M-point forward transform src -> dst through the four-step scratch (d_work[0 .. 2M))
static int32_t big_fft(b2s_fft *p, const float2 *src, float2 *dst, int conj_in, int conj_out, long long src_rot,
                       long long dst_rot, float scale, cudaStream_t st) {
    b2s_ctx *ctx = p->ctx;
    const long long M = (long long)p->big_m, n1 = (long long)p->big_n1, n2 = (long long)p->big_n2;
    float2 *A = p->d_work, *B = p->d_work + M;
    size_t c = 0, q = 0;
    int32_t rc;
    BigT t{};
    t.M = M; t.scale = 1.0f;
    t.src = src; t.dst = A; t.R = n1; t.C = n2; t.conj_in = conj_in; t.src_rot = src_rot;       // x[i1][i2] -> A[i2][i1]
    if ((rc = big_transpose(ctx, t, st))) return rc;
    if ((rc = b2s_fft_exec(p->sub1, A, (size_t)M, B, (size_t)M, &c, &q))) return rc;               // n2 transforms of length n1
    t = BigT{}; t.M = M; t.scale = 1.0f;
    t.src = B; t.dst = A; t.R = n2; t.C = n1; t.twiddle = 1;                                       // Y[i2][k1] W_M^{i2 k1} -> A[k1][i2]
    if ((rc = big_transpose(ctx, t, st))) return rc;
    if ((rc = b2s_fft_exec(p->sub2, A, (size_t)M, B, (size_t)M, &c, &q))) return rc;               // n1 transforms of length n2
    t = BigT{}; t.M = M;
    t.src = B; t.dst = dst; t.R = n1; t.C = n2; t.conj_out = conj_out; t.dst_rot = dst_rot; t.scale = scale;   // Z[k1][k2] -> X[k2*n1 + k1]
    return big_transpose(ctx, t, st);
}

extern "C" {

int32_t b2s_fft_plan_c32(b2s_ctx *ctx, size_t n, int32_t inverse, int32_t fft_shift, int32_t has_normalize,
                         float normalize, b2s_fft **out) {
    if (!ctx || !out) return b2s_fail(ctx, B2S_EINVAL, "b2s_fft_plan_c32: NULL argument");
    *out = nullptr;
    if (n < 2) return b2s_fail(ctx, B2S_EINVAL, "b2s_fft_plan_c32: n must be >= 2");
    const bool pow2 = (n & (n - 1)) == 0;
    // one transform in shared memory up to 16384 points (Bluestein: M >= 2n-1 <= 16384); beyond that four-step
    // through HBM, bounded by the scratch it needs (2 M / 4 M items)
    const bool big = pow2 ? n > 16384 : n > 8192;
    if (pow2 && n > ((size_t)1 << 26)) return b2s_fail(ctx, B2S_EUNSUPPORTED, "b2s_fft_plan_c32: n = %zu > 2^26", n);
    if (!pow2 && n > ((size_t)1 << 24)) return b2s_fail(ctx, B2S_EUNSUPPORTED, "b2s_fft_plan_c32: non-power-of-two n = %zu > 2^24", n);
    DeviceGuard g(ctx->device);
    b2s_fft *p = new b2s_fft();
    p->ctx = ctx; p->n = n; p->big = big;
    p->inverse = inverse != 0; p->shift = fft_shift != 0; p->has_norm = has_normalize != 0; p->norm = normalize;
    const double PI = 3.14159265358979323846264338327950288;
    size_t tw_n = n;
    if (pow2) {
        while (((size_t)1 << p->log2n) < n) p->log2n++;
    } else {
        p->bluestein = true;
        while (((size_t)1 << p->log2m) < 2 * n - 1) p->log2m++;
        tw_n = (size_t)1 << p->log2m;
    }
    if (big) {
        // M = n1 * n2 with both factors <= 16384; the two shared-memory plans do the actual transforms
        p->big_m = tw_n;
        int l2 = 0;
        while (((size_t)1 << l2) < tw_n) l2++;
        p->big_n1 = (size_t)1 << ((l2 + 1) / 2);
        p->big_n2 = tw_n / p->big_n1;
        int32_t rc = b2s_fft_plan_c32(ctx, p->big_n1, 0, 0, 0, 1.0f, &p->sub1);
        if (rc == B2S_OK) rc = b2s_fft_plan_c32(ctx, p->big_n2, 0, 0, 0, 1.0f, &p->sub2);
        if (rc != B2S_OK) { b2s_fft_destroy(p); return rc; }
        if (cudaMalloc((void **)&p->d_work, (p->bluestein ? 4 : 2) * tw_n * sizeof(float2)) != cudaSuccess) {
            cudaGetLastError(); b2s_fft_destroy(p); return b2s_fail(ctx, B2S_ENOMEM, "fft four-step scratch (%zu items)", (p->bluestein ? 4 : 2) * tw_n);
        }
    }
    std::vector<float2> tw(big ? 1 : tw_n);
    for (size_t k = 0; k < tw.size(); k++) {
        const double ang = -2.0 * PI * (double)k / (double)tw_n;
        tw[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    cudaError_t e = cudaMalloc((void **)&p->d_tw, tw.size() * sizeof(float2));
    if (e != cudaSuccess) { b2s_fft_destroy(p); return b2s_fail(ctx, B2S_ENOMEM, "fft twiddles"); }
    B2S_CUDA(ctx, cudaMemcpyAsync(p->d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<float2> chirp, bhat;
    if (p->bluestein) {
        const size_t M = tw_n;
        chirp.resize(n); bhat.resize(M);
        std::vector<double> br(M, 0.0), bi(M, 0.0);
        for (size_t k = 0; k < n; k++) {
            const unsigned long long k2 = ((unsigned long long)k * k) % (2ull * n);     // k^2 mod 2n, exact
            const double ang = -PI * (double)k2 / (double)n;
            chirp[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
            br[k] = std::cos(ang); bi[k] = -std::sin(ang);                             // conj(w[k])
            if (k) { br[M - k] = br[k]; bi[M - k] = bi[k]; }
        }
        // Bhat = FFT_M(b) / M in f64 (iterative radix-2), once per plan
        std::vector<double> xr(br), xi(bi);
        for (size_t i = 1, j = 0; i < M; i++) {
            size_t bit = M >> 1;
            for (; j & bit; bit >>= 1) j ^= bit;
            j ^= bit;
            if (i < j) { std::swap(xr[i], xr[j]); std::swap(xi[i], xi[j]); }
        }
        for (size_t len = 2; len <= M; len <<= 1)
            for (size_t j = 0; j < len / 2; j++) {
                const double ang = -2.0 * PI * (double)j / (double)len, wr = std::cos(ang), wi = std::sin(ang);
                for (size_t s0 = 0; s0 < M; s0 += len) {
                    const size_t u = s0 + j, v = u + len / 2;
                    const double tr = xr[v] * wr - xi[v] * wi, ti = xr[v] * wi + xi[v] * wr;
                    xr[v] = xr[u] - tr; xi[v] = xi[u] - ti; xr[u] += tr; xi[u] += ti;
                }
            }
        for (size_t k = 0; k < M; k++) bhat[k] = make_float2((float)(xr[k] / (double)M), (float)(xi[k] / (double)M));
        if (cudaMalloc((void **)&p->d_chirp, n * sizeof(float2)) != cudaSuccess ||
            cudaMalloc((void **)&p->d_bhat, M * sizeof(float2)) != cudaSuccess) {
            b2s_fft_destroy(p);
            return b2s_fail(ctx, B2S_ENOMEM, "fft bluestein tables");
        }
        B2S_CUDA(ctx, cudaMemcpyAsync(p->d_chirp, chirp.data(), n * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
        B2S_CUDA(ctx, cudaMemcpyAsync(p->d_bhat, bhat.data(), M * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    }
    B2S_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = p;
    return B2S_OK;
}

void b2s_fft_destroy(b2s_fft *p) {
    if (!p) return;
    DeviceGuard g(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    if (p->d_tw) cudaFree(p->d_tw);
    if (p->d_chirp) cudaFree(p->d_chirp);
    if (p->d_bhat) cudaFree(p->d_bhat);
    if (p->d_work) cudaFree(p->d_work);
    if (p->sub1) b2s_fft_destroy(p->sub1);
    if (p->sub2) b2s_fft_destroy(p->sub2);
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
    NvtxRange nvtx("b2s_fft_exec");
    if (p->big) {
        cudaStream_t st = p->ctx->stream;
        const long long n = (long long)p->n, M = (long long)p->big_m, half = n / 2;
        const float scale = p->has_norm ? p->norm : 1.0f;
        const int gridE = p->ctx->sm_count * 8;
        for (size_t fr = 0; fr < m / p->n; fr++) {
            const float2 *in = (const float2 *)d_in + fr * p->n;
            float2 *out = (float2 *)d_out + fr * p->n;
            int32_t rc;
            if (!p->bluestein) {
                // inverse = conj o FFT o conj; shift: pre-rotation for the inverse, post-rotation for the forward transform
                rc = big_fft(p, in, out, p->inverse, p->inverse, (p->inverse && p->shift) ? half : 0,
                             (!p->inverse && p->shift) ? half : 0, scale, st);
                if (rc) return rc;
                continue;
            }
            float2 *b0 = p->d_work + 2 * M, *b1 = p->d_work + 3 * M;
            big_bs_pre<<<gridE, 256, 0, st>>>(in, p->d_chirp, b0, n, M, p->inverse, (p->inverse && p->shift) ? half : 0);
            B2S_CHECK_LAUNCH(p->ctx);
            if ((rc = big_fft(p, b0, b1, 0, 0, 0, 0, 1.0f, st))) return rc;
            big_bs_mul<<<gridE, 256, 0, st>>>(b1, p->d_bhat, M);
            B2S_CHECK_LAUNCH(p->ctx);
            if ((rc = big_fft(p, b1, b0, 0, 0, 0, 0, 1.0f, st))) return rc;
            big_bs_post<<<gridE, 256, 0, st>>>(b0, p->d_chirp, out, n, p->inverse, (!p->inverse && p->shift) ? n - half : 0, scale);
            B2S_CHECK_LAUNCH(p->ctx);
        }
        return B2S_OK;
    }
    if (p->bluestein) {
        BsArgs b;
        b.in = (const float2 *)d_in; b.out = (float2 *)d_out; b.tw = p->d_tw; b.chirp = p->d_chirp; b.bhat = p->d_bhat;
        b.nfft = (long long)(m / p->n); b.n = (int)p->n;
        b.inverse = p->inverse; b.shift = p->shift; b.has_norm = p->has_norm; b.norm = p->norm;
        cudaStream_t bs = p->ctx->stream;
        switch (p->log2m) {
            case 2: return launch_bluestein<2>(p, b, bs);
            case 3: return launch_bluestein<3>(p, b, bs);
            case 4: return launch_bluestein<4>(p, b, bs);
            case 5: return launch_bluestein<5>(p, b, bs);
            case 6: return launch_bluestein<6>(p, b, bs);
            case 7: return launch_bluestein<7>(p, b, bs);
            case 8: return launch_bluestein<8>(p, b, bs);
            case 9: return launch_bluestein<9>(p, b, bs);
            case 10: return launch_bluestein<10>(p, b, bs);
            case 11: return launch_bluestein<11>(p, b, bs);
            case 12: return launch_bluestein<12>(p, b, bs);
            case 13: return launch_bluestein<13>(p, b, bs);
            case 14: return launch_bluestein<14>(p, b, bs);
        }
        return b2s_fail(p->ctx, B2S_EUNSUPPORTED, "b2s_fft_exec: unsupported Bluestein size");
    }
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
