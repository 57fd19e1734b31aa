// fft_common.cuh -- in-register radix-2/4/8/16 forward DFTs and the padded shared-memory index
// shared by fft.cu (Fft block) and fir_fft.cu (overlap-save FIR).
#pragma once
#include <cuda_runtime.h>

namespace fftk {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// ---- in-register forward DFTs, natural-order output ------------------------------------------
template <int R> struct Dft;
template <> struct Dft<2> {
    __device__ static __forceinline__ void run(float2 (&v)[2]) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    }
};
template <> struct Dft<4> {
    __device__ static __forceinline__ void run(float2 (&v)[4]) {
        const float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        const float2 t2 = cadd(v[1], v[3]), t3 = mul_mi(csub(v[1], v[3]));
        v[0] = cadd(t0, t2); v[1] = cadd(t1, t3); v[2] = csub(t0, t2); v[3] = csub(t1, t3);
    }
};
template <> struct Dft<8> {
    __device__ static __forceinline__ void run(float2 (&v)[8]) {
        float2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
        Dft<4>::run(e); Dft<4>::run(o);
        constexpr float h = 0.70710678118654752440f;
        o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));      // * W8^1 = (1-i)/sqrt2
        o[2] = mul_mi(o[2]);                                                   // * W8^2 = -i
        o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));     // * W8^3 = (-1-i)/sqrt2
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = cadd(e[k], o[k]); v[k + 4] = csub(e[k], o[k]); }
    }
};
template <> struct Dft<16> {
    __device__ static __forceinline__ void run(float2 (&v)[16]) {
        float2 e[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
        Dft<8>::run(e); Dft<8>::run(o);
        // W16^k = exp(-2 pi i k / 16)
        constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
        constexpr float h = 0.70710678118654752440f;
        o[1] = cmul(o[1], make_float2(c1, -s1));
        o[2] = make_float2(h * (o[2].x + o[2].y), h * (o[2].y - o[2].x));
        o[3] = cmul(o[3], make_float2(s1, -c1));
        o[4] = mul_mi(o[4]);
        o[5] = cmul(o[5], make_float2(-s1, -c1));
        o[6] = make_float2(h * (o[6].y - o[6].x), -h * (o[6].x + o[6].y));
        o[7] = cmul(o[7], make_float2(-c1, -s1));
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] = cadd(e[k], o[k]); v[k + 8] = csub(e[k], o[k]); }
    }
};

__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }


// v[r] *= w^r for r = 1..R-1, powers built by a log-depth product tree from ONE table load
// (15 dependent-free complex multiplies instead of 15 scattered 8-byte loads per butterfly: the
// table loads of a radix-16 pass cost ~16 L1 wavefronts each and made the FFT LSU-bound).
// Rounding: every power is at most 4 multiplies deep -> |error| <= ~4 ulp of the twiddle.
template <int R>
__device__ __forceinline__ void apply_twiddles(float2 (&v)[R], float2 w1) {
    if constexpr (R >= 2) v[1] = cmul(v[1], w1);
    if constexpr (R >= 4) {
        const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
        v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3);
        if constexpr (R >= 8) {
            const float2 w4 = cmul(w2, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
            v[4] = cmul(v[4], w4); v[5] = cmul(v[5], w5); v[6] = cmul(v[6], w6); v[7] = cmul(v[7], w7);
            if constexpr (R >= 16) {
                const float2 w8 = cmul(w4, w4);
                v[8] = cmul(v[8], w8); v[9] = cmul(v[9], cmul(w8, w1)); v[10] = cmul(v[10], cmul(w8, w2));
                v[11] = cmul(v[11], cmul(w8, w3)); v[12] = cmul(v[12], cmul(w8, w4)); v[13] = cmul(v[13], cmul(w8, w5));
                v[14] = cmul(v[14], cmul(w8, w6)); v[15] = cmul(v[15], cmul(w8, w7));
            }
        }
    }
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


// One Stockham pass: loads through `load`, optional CTA barrier (in-place passes), twiddles, DFT,
// stores through `store`, CTA barrier.
// The pass's base twiddle of butterfly j: a table lookup, or (TwPre) a value the caller fetched ahead of time -- with
// one butterfly per thread the index depends on the thread alone, so a kernel that runs the same plan repeatedly loads
// it once instead of waiting for a global load after every barrier.
struct TwTable {
    const float2 *__restrict__ tw;
    template <int N, int R, int NS> __device__ __forceinline__ float2 get(int j, int) const {
        return __ldg(tw + (j & (NS - 1)) * (N / (NS * R)));
    }
};
struct TwPre {
    float2 w;
    template <int N, int R, int NS> __device__ __forceinline__ float2 get(int, int) const { return w; }
};
// Base twiddles of one whole pass (ITER butterflies per thread), fetched before the PREVIOUS pass runs.
template <int N, int R, int NS, int T> struct TwAhead {
    static constexpr int ITER = (N / R) / T > 0 ? (N / R) / T : 1;
    float2 w[ITER];
    __device__ __forceinline__ void fetch(const float2 *__restrict__ tw, int t) {
#pragma unroll
        for (int it = 0; it < ITER; it++) w[it] = __ldg(tw + ((t + it * T) & (NS - 1)) * (N / (NS * R)));
    }
    template <int N_, int R_, int NS_> __device__ __forceinline__ float2 get(int, int it) const { return w[it]; }
};

template <int N, int R, int NS, int T, typename LoadF, typename StoreF, typename TwF>
__device__ __forceinline__ void ss_pass_tw(LoadF load, StoreF store, TwF twf, int t, bool sync_between) {
    constexpr int NB = N / R, ITER = NB / T;
    static_assert(NB % T == 0 || NB < T, "butterflies must tile the threads");
    float2 v[ITER][R];
#pragma unroll
    for (int it = 0; it < ITER; it++) {
        const int j = t + it * T;
#pragma unroll
        for (int r = 0; r < R; r++) v[it][r] = load(j + r * NB);
    }
    if (sync_between) __syncthreads();
#pragma unroll
    for (int it = 0; it < ITER; it++) {
        const int j = t + it * T;
        if constexpr (NS > 1) apply_twiddles<R>(v[it], twf.template get<N, R, NS>(j, it));
        Dft<R>::run(v[it]);
        const int j0 = (j / NS) * NS * R + (j & (NS - 1));
#pragma unroll
        for (int r = 0; r < R; r++) store(j0 + r * NS, v[it][r]);
    }
    __syncthreads();
}

template <int N, int R, int NS, int T, typename LoadF, typename StoreF>
__device__ __forceinline__ void ss_pass(LoadF load, StoreF store, const float2 *__restrict__ tw, int t,
                                        bool sync_between) {
    ss_pass_tw<N, R, NS, T>(load, store, TwTable{tw}, t, sync_between);
}


// Runs a whole LOG2N-point forward FFT as Stockham passes through the padded smem buffer `sm`.
// Pass 0 reads through `first_load(idx)`, the last pass writes through `last_store(idx, v)`, the
// passes in between read and write `sm`.  first_reads_smem: pass 0's source is `sm` itself.
template <int LOG2N, int T, typename LoadF, typename StoreF>
__device__ __forceinline__ void fft_passes(LoadF first_load, StoreF last_store, float2 *sm,
                                           const float2 *__restrict__ tw, int t, bool first_reads_smem) {
    constexpr int N = 1 << LOG2N;
    using PL = Plan<LOG2N>;
    auto ld_sm = [&](int idx) { return sm[pad(idx)]; };
    auto st_sm = [&](int idx, float2 v) { sm[pad(idx)] = v; };
    // every pass's base twiddles are fetched one pass ahead: the load is in flight across the barrier
    if constexpr (PL::P == 1) {
        ss_pass<N, PL::R0, 1, T>(first_load, last_store, tw, t, first_reads_smem);
    } else if constexpr (PL::P == 2) {
        TwAhead<N, PL::R1, PL::R0, T> w1; w1.fetch(tw, t);
        ss_pass<N, PL::R0, 1, T>(first_load, st_sm, tw, t, first_reads_smem);
        ss_pass_tw<N, PL::R1, PL::R0, T>(ld_sm, last_store, w1, t, true);
    } else if constexpr (PL::P == 3) {
        TwAhead<N, PL::R1, PL::R0, T> w1; w1.fetch(tw, t);
        ss_pass<N, PL::R0, 1, T>(first_load, st_sm, tw, t, first_reads_smem);
        TwAhead<N, PL::R2, PL::R0 * PL::R1, T> w2; w2.fetch(tw, t);
        ss_pass_tw<N, PL::R1, PL::R0, T>(ld_sm, st_sm, w1, t, true);
        ss_pass_tw<N, PL::R2, PL::R0 * PL::R1, T>(ld_sm, last_store, w2, t, true);
    } else {
        TwAhead<N, PL::R1, PL::R0, T> w1; w1.fetch(tw, t);
        ss_pass<N, PL::R0, 1, T>(first_load, st_sm, tw, t, first_reads_smem);
        TwAhead<N, PL::R2, PL::R0 * PL::R1, T> w2; w2.fetch(tw, t);
        ss_pass_tw<N, PL::R1, PL::R0, T>(ld_sm, st_sm, w1, t, true);
        TwAhead<N, PL::R3, PL::R0 * PL::R1 * PL::R2, T> w3; w3.fetch(tw, t);
        ss_pass_tw<N, PL::R2, PL::R0 * PL::R1, T>(ld_sm, st_sm, w2, t, true);
        ss_pass_tw<N, PL::R3, PL::R0 * PL::R1 * PL::R2, T>(ld_sm, last_store, w3, t, true);
    }
}

// Same, with a hook that runs right after the first pass (its source buffer is free from then on: the caller
// starts the asynchronous fetch of the NEXT transform's input there).  Plans with at least two passes (N >= 32).
template <int LOG2N, int T, typename LoadF, typename HookF, typename StoreF>
__device__ __forceinline__ void fft_passes_hook(LoadF first_load, HookF after_first, StoreF last_store, float2 *sm,
                                                const float2 *__restrict__ tw, int t) {
    constexpr int N = 1 << LOG2N;
    using PL = Plan<LOG2N>;
    static_assert(PL::P >= 2, "fft_passes_hook needs a multi-pass plan");
    auto ld_sm = [&](int idx) { return sm[pad(idx)]; };
    auto st_sm = [&](int idx, float2 v) { sm[pad(idx)] = v; };
    TwAhead<N, PL::R1, PL::R0, T> w1; w1.fetch(tw, t);
    ss_pass<N, PL::R0, 1, T>(first_load, st_sm, tw, t, false);
    after_first();
    if constexpr (PL::P == 2) {
        ss_pass_tw<N, PL::R1, PL::R0, T>(ld_sm, last_store, w1, t, true);
    } else if constexpr (PL::P == 3) {
        TwAhead<N, PL::R2, PL::R0 * PL::R1, T> w2; w2.fetch(tw, t);
        ss_pass_tw<N, PL::R1, PL::R0, T>(ld_sm, st_sm, w1, t, true);
        ss_pass_tw<N, PL::R2, PL::R0 * PL::R1, T>(ld_sm, last_store, w2, t, true);
    } else {
        TwAhead<N, PL::R2, PL::R0 * PL::R1, T> w2; w2.fetch(tw, t);
        ss_pass_tw<N, PL::R1, PL::R0, T>(ld_sm, st_sm, w1, t, true);
        TwAhead<N, PL::R3, PL::R0 * PL::R1 * PL::R2, T> w3; w3.fetch(tw, t);
        ss_pass_tw<N, PL::R2, PL::R0 * PL::R1, T>(ld_sm, st_sm, w2, t, true);
        ss_pass_tw<N, PL::R3, PL::R0 * PL::R1 * PL::R2, T>(ld_sm, last_store, w3, t, true);
    }
}

}  // namespace fftk
