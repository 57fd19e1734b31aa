// pfbarb.cu -- polyphase arbitrary-rate resampler (src/blocks/pfb/arb_resampler.rs:90-231,
// pfb/utilities.rs:5-25, pfb/window_buffer.rs:13-44) on the device.
//
// The reference is a strictly sequential state machine: a f32 timing recurrence
// (update_timing_state :132-140, `tau -= 1.0` :184-186) decides, sample by sample, how many
// outputs are produced and which two polyphase arms are blended.  The recurrence does not
// depend on the data, only on (rate, num_filters, number of samples), so it is split off:
//   * host: replays the reference's f32 recurrence once per call, bit-for-bit (plain C float
//     ops, no contraction), and records the timing state at the start of every 32-sample
//     sub-block -- 16 bytes per 32 samples, this is what makes output COUNTS and arm indices
//     identical to the reference;
//   * device: every sub-block is replayed by one thread from its recorded state (same IEEE
//     operations via __fadd_rn/__fmul_rn), which yields one descriptor per output in shared
//     memory; then all threads of the CTA evaluate the outputs: two T-tap dot products on the
//     window ending at the right sample, blended with (1-mu, mu) (:159-176, Boundary :147-156).
// The input history (the reference's WindowBuffer) is a T-sample device buffer carried between
// calls, including the reference's start-up behaviour: while the window fills, push() writes
// sample j at slot (start_idx - missing) mod T (window_buffer.rs:24-32), which scatters the
// first T samples (it is not a plain append); this is reproduced so the first outputs match.
#include <cmath>

#include "common.cuh"

namespace {
constexpr int kSB = 32;            // input samples per recorded sub-block
constexpr int kPaThreads = 256;
constexpr int kDescCap = 4096;     // outputs per CTA (descriptor slots in shared memory)

struct SubRec {                    // timing state at a sub-block boundary
    uint32_t out0;                 // index (within the call) of the sub-block's first output
    float tau;
    float mu;
    uint32_t base_flag;            // base_index | boundary << 31
};
}  // namespace

struct b2s_pfbarb {
    b2s_ctx *ctx = nullptr;
    size_t num_filters = 0, T = 0, ntaps = 0;
    float rate = 1.f, delay = 1.f;
    float *d_arms = nullptr;       // [num_filters][T], time-reversed: arm_b[T-1-j]
    float2 *d_circ = nullptr;      // 2*T, the reference's circular buffer (only used while filling)
    float2 *d_hist = nullptr;      // T samples of history once filled
    // WindowBuffer bookkeeping (host)
    size_t start_idx = 0, missing = 0;
    // State (host): arb_resampler.rs:40-52
    float tau = 0.f, bf = 0.f, mu = 0.f;
    size_t base_index = 0;
    bool boundary = false;
    // per-call records
    SubRec *h_recs = nullptr;      // pinned
    SubRec *d_recs = nullptr;
    size_t recs_cap = 0;
};

namespace {

// ---- window fill: the reference's push() while num_samples_missing > 0 ------------------------
__global__ void pfb_fill_kernel(const float2 *__restrict__ in, float2 *circ, int L, int start_idx, int missing,
                                int count) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int c = 0; c < count; c++) {
        int idx = (start_idx - missing) % L;
        if (idx < 0) idx += L;                                   // rem_euclid
        circ[idx] = in[c];
        circ[idx + L] = in[c];
        if (missing > 0) missing--;
        start_idx = (start_idx + 1) % L;
    }
}

__global__ void pfb_hist_from_circ(const float2 *__restrict__ circ, float2 *hist, int L, int start_idx) {
    for (int j = threadIdx.x; j < L; j += blockDim.x) hist[j] = circ[start_idx + j];
}

// hist <- last L samples of [hist | in[0..n))
__global__ void pfb_hist_update(float2 *hist, const float2 *__restrict__ in, int L, long long n) {
    extern __shared__ float2 tmp[];
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const long long idx = n + j;                             // position in [hist | in]
        tmp[j] = (idx < L) ? hist[idx] : in[idx - L];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < L; j += blockDim.x) hist[j] = tmp[j];
}

__device__ __forceinline__ float2 pfb_x(const float2 *__restrict__ hist, const float2 *__restrict__ in, int L,
                                        long long idx) {
    return idx < L ? hist[idx] : __ldg(in + (idx - L));
}

__global__ void __launch_bounds__(kPaThreads)
pfb_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, const float2 *__restrict__ hist,
           const float *__restrict__ arms, const SubRec *__restrict__ recs, long long n_in, int nsub,
           int sub_per_cta, int N, int T, float delay, int arms_in_smem) {
    extern __shared__ __align__(16) unsigned char psm[];
    uint32_t *d_s1 = reinterpret_cast<uint32_t *>(psm);          // window start of y1 | boundary << 31
    uint32_t *d_b0 = d_s1 + kDescCap;                            // arm of y0
    float *d_mu = reinterpret_cast<float *>(d_b0 + kDescCap);
    float *s_arms = d_mu + kDescCap;

    const int sb0 = blockIdx.x * sub_per_cta;
    const int sb1 = min(sb0 + sub_per_cta, nsub);
    const uint32_t o_first = recs[sb0].out0;
    const uint32_t o_end = recs[sb1].out0;                       // recs has nsub + 1 entries
    if (arms_in_smem)
        for (int j = threadIdx.x; j < N * T; j += kPaThreads) s_arms[j] = arms[j];

    // ---- phase 1: replay the timing recurrence of each sub-block (one thread per sub-block)
    for (int sb = sb0 + threadIdx.x; sb < sb1; sb += kPaThreads) {
        const SubRec r = recs[sb];
        float tau = r.tau, mu = r.mu;
        uint32_t base = r.base_flag & 0x7fffffffu;
        bool boundary = (r.base_flag >> 31) != 0;
        uint32_t o = r.out0 - o_first;
        const long long s_beg = (long long)sb * kSB, s_end = min(s_beg + kSB, n_in);
        const float fN = (float)N;
        for (long long s = s_beg; s < s_end; s++) {
            while (base < (uint32_t)N) {
                if (boundary) {
                    d_s1[o] = (uint32_t)(s + 1) | 0x80000000u; d_b0[o] = (uint32_t)(N - 1); d_mu[o] = mu; o++;
                    tau = __fadd_rn(tau, delay);
                    const float bf = __fmul_rn(tau, fN);
                    base = (uint32_t)floorf(bf);
                    mu = __fsub_rn(bf, (float)base);
                    boundary = false;
                } else if (base == (uint32_t)(N - 1)) {
                    boundary = true;
                    base = (uint32_t)N;
                } else {
                    d_s1[o] = (uint32_t)(s + 1); d_b0[o] = base; d_mu[o] = mu; o++;
                    tau = __fadd_rn(tau, delay);
                    const float bf = __fmul_rn(tau, fN);
                    base = (uint32_t)floorf(bf);
                    mu = __fsub_rn(bf, (float)base);
                }
            }
            tau = __fsub_rn(tau, 1.0f);
            base -= (uint32_t)N;
        }
    }
    __syncthreads();

    // ---- phase 2: evaluate the outputs
    const float *A = arms_in_smem ? s_arms : arms;
    const uint32_t cnt = o_end - o_first;
    for (uint32_t o = threadIdx.x; o < cnt; o += kPaThreads) {
        const uint32_t w = d_s1[o];
        const bool boundary = (w >> 31) != 0;
        const long long s1 = (long long)(w & 0x7fffffffu);
        const long long s0 = boundary ? s1 - 1 : s1;
        const uint32_t b0 = d_b0[o], b1 = boundary ? 0u : b0 + 1u;
        const float mu = d_mu[o];
        const float *a0 = A + (size_t)b0 * T, *a1 = A + (size_t)b1 * T;
        float2 y0 = make_float2(0.f, 0.f), y1 = make_float2(0.f, 0.f);
        if (!boundary) {
            for (int j = 0; j < T; j++) {
                const float2 x = pfb_x(hist, in, T, s1 + j);
                const float t0 = a0[j], t1 = a1[j];
                y0.x = fmaf(x.x, t0, y0.x); y0.y = fmaf(x.y, t0, y0.y);
                y1.x = fmaf(x.x, t1, y1.x); y1.y = fmaf(x.y, t1, y1.y);
            }
        } else {
            for (int j = 0; j < T; j++) {
                const float2 xa = pfb_x(hist, in, T, s0 + j), xb = pfb_x(hist, in, T, s1 + j);
                const float t0 = a0[j], t1 = a1[j];
                y0.x = fmaf(xa.x, t0, y0.x); y0.y = fmaf(xa.y, t0, y0.y);
                y1.x = fmaf(xb.x, t1, y1.x); y1.y = fmaf(xb.y, t1, y1.y);
            }
        }
        // (1.0 - mu) * buff[0] + mu * buff[1]   (arb_resampler.rs:153,:176)
        const float a = __fsub_rn(1.0f, mu);
        float2 r;
        r.x = __fadd_rn(__fmul_rn(a, y0.x), __fmul_rn(mu, y1.x));
        r.y = __fadd_rn(__fmul_rn(a, y0.y), __fmul_rn(mu, y1.y));
        out[(size_t)o_first + o] = r;
    }
}

// Host replay of State::consume_single's control flow (arb_resampler.rs:142-188) for n samples,
// recording the state at sub-block starts.  Must stay bit-identical to the reference: plain
// float ops, this translation unit is compiled without fast-math / contraction on the host side.
size_t host_schedule(b2s_pfbarb *p, size_t n, SubRec *recs) {
    const uint32_t N = (uint32_t)p->num_filters;
    // Plain float locals: the host side of this file is built by g++ for x86-64 without -ffast-math, so
    // every + and * is one IEEE binary32 SSE operation (no x87 excess precision, no FMA contraction) and
    // the sequence is the reference's.  (They used to be `volatile`, which put a store-to-load round
    // trip on the tau chain and halved the replay rate.)
    float tau = p->tau, bf = p->bf, mu = p->mu;
    size_t base = p->base_index;
    bool boundary = p->boundary;
    size_t o = 0;
    const float delay = p->delay, fN = (float)N;
    for (size_t s = 0; s < n; s++) {
        if ((s % kSB) == 0) {
            SubRec &r = recs[s / kSB];
            r.out0 = (uint32_t)o; r.tau = tau; r.mu = mu;
            r.base_flag = (uint32_t)base | (boundary ? 0x80000000u : 0u);
        }
        while (base < N) {
            if (boundary) {
                o++;
                tau = tau + delay; bf = tau * fN; base = (size_t)floorf(bf); mu = bf - (float)base;
                boundary = false;
            } else if (base == N - 1) {
                boundary = true;
                base = N;
            } else {
                o++;
                tau = tau + delay; bf = tau * fN; base = (size_t)floorf(bf); mu = bf - (float)base;
            }
        }
        tau = tau - 1.0f;
        bf = bf - fN;
        base -= N;
    }
    SubRec &e = recs[ceil_div(n, (size_t)kSB)];
    e.out0 = (uint32_t)o; e.tau = tau; e.mu = mu; e.base_flag = (uint32_t)base | (boundary ? 0x80000000u : 0u);
    p->tau = tau; p->bf = bf; p->mu = mu; p->base_index = base; p->boundary = boundary;
    return o;
}

}  // namespace

extern "C" {

int32_t b2s_pfbarb_plan_c32(b2s_ctx *ctx, const float *taps, size_t ntaps, size_t num_filters, float rate,
                            b2s_pfbarb **out) {
    if (!ctx || !out || !taps) return b2s_fail(ctx, B2S_EINVAL, "b2s_pfbarb_plan_c32: NULL argument");
    *out = nullptr;
    // the reference asserts these (arb_resampler.rs:92-104)
    if (!(rate > 0.f)) return b2s_fail(ctx, B2S_EINVAL, "PfbArbResampler: resampling rate must be greater than zero");
    if (num_filters == 0) return b2s_fail(ctx, B2S_EINVAL, "PfbArbResampler: number of filter banks must be greater than zero");
    if (ntaps < num_filters) return b2s_fail(ctx, B2S_EINVAL, "PfbArbResampler: prototype filter length must be at least num_filters");
    if (num_filters > (1u << 20) || rate > 1024.f) return b2s_fail(ctx, B2S_EUNSUPPORTED, "PfbArbResampler: num_filters / rate too large");
    DeviceGuard g(ctx->device);
    b2s_pfbarb *p = new b2s_pfbarb();
    p->ctx = ctx; p->num_filters = num_filters; p->ntaps = ntaps; p->rate = rate;
    p->delay = 1.0f / rate;
    // partition_filter_taps (utilities.rs:9-19): T = ceil(len as f32 / n as f32); arm i = taps[i::n] zero padded
    const size_t T = (size_t)std::ceil((float)ntaps / (float)num_filters);
    p->T = T;
    std::vector<float> arms(num_filters * T, 0.0f);
    for (size_t i = 0; i < num_filters; i++) {
        size_t j = 0;
        for (size_t idx = i; idx < ntaps; idx += num_filters, j++) arms[i * T + (T - 1 - j)] = taps[idx];   // reversed
    }
    cudaError_t e1 = cudaMalloc((void **)&p->d_arms, arms.size() * sizeof(float));
    cudaError_t e2 = cudaMalloc((void **)&p->d_circ, 2 * T * sizeof(float2));
    cudaError_t e3 = cudaMalloc((void **)&p->d_hist, T * sizeof(float2));
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { b2s_pfbarb_destroy(p); return b2s_fail(ctx, B2S_ENOMEM, "pfbarb buffers"); }
    B2S_CUDA(ctx, cudaMemcpyAsync(p->d_arms, arms.data(), arms.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    B2S_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = p;
    return b2s_pfbarb_reset(p);
}

void b2s_pfbarb_destroy(b2s_pfbarb *p) {
    if (!p) return;
    DeviceGuard g(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    if (p->d_arms) cudaFree(p->d_arms);
    if (p->d_circ) cudaFree(p->d_circ);
    if (p->d_hist) cudaFree(p->d_hist);
    if (p->d_recs) cudaFree(p->d_recs);
    if (p->h_recs) cudaFreeHost(p->h_recs);
    delete p;
}

int32_t b2s_pfbarb_reset(b2s_pfbarb *p) {
    if (!p) return b2s_fail(nullptr, B2S_EINVAL, "pfbarb is NULL");
    DeviceGuard g(p->ctx->device);
    p->start_idx = 0; p->missing = p->T;                           // WindowBuffer::new(len, pad_start=false)
    p->tau = 0.f; p->bf = 0.f; p->mu = 0.f; p->base_index = 0; p->boundary = false;
    B2S_CUDA(p->ctx, cudaMemsetAsync(p->d_circ, 0, 2 * p->T * sizeof(float2), p->ctx->stream));
    B2S_CUDA(p->ctx, cudaMemsetAsync(p->d_hist, 0, p->T * sizeof(float2), p->ctx->stream));
    return B2S_OK;
}

int32_t b2s_pfbarb_exec(b2s_pfbarb *p, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                        size_t *consumed, size_t *produced, int32_t *call_again) {
    if (!p || !consumed || !produced || !call_again) return b2s_fail(p ? p->ctx : nullptr, B2S_EINVAL, "b2s_pfbarb_exec: NULL argument");
    b2s_ctx *ctx = p->ctx;
    *consumed = 0; *produced = 0; *call_again = 0;
    DeviceGuard g(ctx->device);
    const int T = (int)p->T;
    const float2 *in = (const float2 *)d_in;
    // fill filter history (arb_resampler.rs:199-215)
    if (p->missing != 0) {
        const size_t c = std::min(p->missing, n_in);
        if (c) {
            pfb_fill_kernel<<<1, 32, 0, ctx->stream>>>(in, p->d_circ, T, (int)p->start_idx, (int)p->missing, (int)c);
            B2S_CHECK_LAUNCH(ctx);
            p->missing -= c;
            p->start_idx = (p->start_idx + c) % p->T;
            if (p->missing == 0) {
                pfb_hist_from_circ<<<1, 256, 0, ctx->stream>>>(p->d_circ, p->d_hist, T, (int)p->start_idx);
                B2S_CHECK_LAUNCH(ctx);
            }
        }
        *consumed = c;
        if (n_in - c > 0) *call_again = 1;
        return B2S_OK;
    }
    // nitem_to_process = min(ninput_items, (noutput_items as f32 / rate) as usize)   (:218)
    const size_t cap = (size_t)((float)n_out_cap / p->rate);
    const size_t n = std::min(n_in, cap);
    if (n == 0) return B2S_OK;
    if (!d_in || !d_out) return b2s_fail(ctx, B2S_EINVAL, "b2s_pfbarb_exec: NULL buffer");
    if (n >= (1ull << 31)) return b2s_fail(ctx, B2S_EUNSUPPORTED, "b2s_pfbarb_exec: more than 2^31 items per call");
    const size_t nsub = ceil_div(n, (size_t)kSB);
    if (p->recs_cap < nsub + 1) {
        B2S_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (p->d_recs) cudaFree(p->d_recs);
        if (p->h_recs) cudaFreeHost(p->h_recs);
        p->recs_cap = (nsub + 1) * 5 / 4 + 16;
        B2S_CUDA(ctx, cudaMalloc((void **)&p->d_recs, p->recs_cap * sizeof(SubRec)));
        B2S_CUDA(ctx, cudaHostAlloc((void **)&p->h_recs, p->recs_cap * sizeof(SubRec), cudaHostAllocDefault));
    } else {
        // the pinned records of the previous call may still be in flight to the device
        B2S_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const size_t nout = host_schedule(p, n, p->h_recs);
    if (nout > n_out_cap)
        return b2s_fail(ctx, B2S_ESTATE, "pfbarb: schedule produced %zu > capacity %zu (the reference would overrun its slice)", nout, n_out_cap);
    B2S_CUDA(ctx, cudaMemcpyAsync(p->d_recs, p->h_recs, (nsub + 1) * sizeof(SubRec), cudaMemcpyHostToDevice, ctx->stream));
    // CTA tiling: sub-blocks per CTA so that a CTA never exceeds kDescCap outputs
    const size_t per_sample_max = (size_t)std::ceil(p->rate) + 1;
    size_t sub_per_cta = kDescCap / (per_sample_max * kSB);
    if (sub_per_cta == 0) return b2s_fail(ctx, B2S_EUNSUPPORTED, "pfbarb: rate %f too high for the descriptor tile", (double)p->rate);
    sub_per_cta = std::min<size_t>(sub_per_cta, kPaThreads);
    const unsigned grid = (unsigned)ceil_div(nsub, sub_per_cta);
    const size_t arms_bytes = p->num_filters * p->T * sizeof(float);
    const int arms_smem = arms_bytes <= 64 * 1024;
    const size_t smem = 3 * kDescCap * sizeof(uint32_t) + (arms_smem ? arms_bytes : 0);
    B2S_CUDA(ctx, cudaFuncSetAttribute(pfb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kDescCap * 4 + 64 * 1024));
    pfb_kernel<<<grid, kPaThreads, smem, ctx->stream>>>(in, (float2 *)d_out, p->d_hist, p->d_arms, p->d_recs,
                                                        (long long)n, (int)nsub, (int)sub_per_cta,
                                                        (int)p->num_filters, T, p->delay, arms_smem);
    B2S_CHECK_LAUNCH(ctx);
    pfb_hist_update<<<1, 256, T * sizeof(float2), ctx->stream>>>(p->d_hist, in, T, (long long)n);
    B2S_CHECK_LAUNCH(ctx);
    *consumed = n; *produced = nout;
    return B2S_OK;
}

}  // extern "C"
