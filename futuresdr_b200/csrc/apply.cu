// apply.cu -- element-wise Apply block on the device (src/blocks/apply.rs:100-131).
//
// The reference applies an arbitrary Rust closure `FnMut(&A) -> B` per sample; a device
// backend cannot run host closures, so this is the closed catalogue of the closures that appear
// on the hot path / in the reference's GPU examples (b2s_op).  Stateful closures keep their
// state in device memory: the FM demodulator's `last` sample
// (examples/fm-receiver/src/main.rs:99-104) is the previous input item, so item j reads
// in[j-1] and item 0 reads the carried sample; after the launch the carry is refreshed from
// in[m-1] on the same stream.
#include "common.cuh"

struct b2s_apply {
    b2s_ctx *ctx = nullptr;
    b2s_op op = B2S_OP_SCALE_F32;
    float param = 1.0f;
    float2 *d_carry = nullptr;   // closure state (QUAD_DEMOD*: last sample)
};

namespace {

// arg(v * conj(last)) with num_complex's Mul: re = a.re*b.re - a.im*b.im, im = a.re*b.im + a.im*b.re,
// b = conj(last) = (lr, -li).  __fmul_rn/__fsub_rn keep the products un-fused like the Rust code.
// atan2 for finite inputs in ~25 instructions (CUDA's atan2f is ~60 on its fast path and made the
// demodulator issue-bound at half the HBM roofline): a = min/max of the magnitudes (MUFU.RCP based
// division), atan(a) = a * P(a^2) with a degree-8 minimax P (max error 1.1e-7 rad on [0,1] in f32 Horner
// form, coefficients fitted in scripts -- see DESIGN.md 4.6), then the octant is unfolded with the SIGN
// BITS so that +-0 behave like libm: atan2(+0,-0) = pi, atan2(0,+0) = 0 (the demodulator's first
// sample multiplies by conj(0)).  Total error < 3e-7 rad; parity bar 1e-5*pi (tests/test_gpu_blocks.py).
__device__ __forceinline__ float atan2_finite(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
    const float s = a * a;
    float p = 0.0029408063273876905f;
    p = fmaf(p, s, -0.0164317823946476f);
    p = fmaf(p, s, 0.04328067600727081f);
    p = fmaf(p, s, -0.07554050534963608f);
    p = fmaf(p, s, 0.10664203763008118f);
    p = fmaf(p, s, -0.14209550619125366f);
    p = fmaf(p, s, 0.19993355870246887f);
    p = fmaf(p, s, -0.33333107829093933f);
    p = fmaf(p, s, 1.0f);
    float r = p * a;
    if (ay > ax) r = 1.57079632679489662f - r;
    if (__float_as_int(x) < 0) r = 3.14159265358979324f - r;
    return copysignf(r, y);
}

__device__ __forceinline__ float quad_demod_one(float2 v, float2 last) {
    const float cr = last.x, ci = -last.y;
    const float pr = __fsub_rn(__fmul_rn(v.x, cr), __fmul_rn(v.y, ci));
    const float pi = __fadd_rn(__fmul_rn(v.x, ci), __fmul_rn(v.y, cr));
    return atan2_finite(pi, pr);
}

template <int OP>
__global__ void apply_kernel(const void *__restrict__ vin, void *__restrict__ vout, long long n, float param,
                             const float2 *__restrict__ carry) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        if constexpr (OP == B2S_OP_SCALE_F32) {
            ((float *)vout)[j] = ((const float *)vin)[j] * param;
        } else if constexpr (OP == B2S_OP_SCALE_C32) {
            const float2 v = ((const float2 *)vin)[j];
            ((float2 *)vout)[j] = make_float2(v.x * param, v.y * param);
        } else if constexpr (OP == B2S_OP_QUAD_DEMOD || OP == B2S_OP_QUAD_DEMOD_C32) {
            const float2 *in = (const float2 *)vin;
            const float2 v = in[j];
            const float2 last = (j == 0) ? *carry : in[j - 1];
            const float ph = quad_demod_one(v, last);
            if constexpr (OP == B2S_OP_QUAD_DEMOD) ((float *)vout)[j] = ph;
            else ((float2 *)vout)[j] = make_float2(ph, 0.0f);
        } else if constexpr (OP == B2S_OP_NORM_SQR) {
            const float2 v = ((const float2 *)vin)[j];
            ((float *)vout)[j] = __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y));   // Complex::norm_sqr
        } else if constexpr (OP == B2S_OP_EXP_F32) {
            ((float *)vout)[j] = expf(((const float *)vin)[j]);
        } else if constexpr (OP == B2S_OP_MAG_C32) {
            const float2 v = ((const float2 *)vin)[j];
            ((float *)vout)[j] = hypotf(v.x, v.y);                                         // Complex::norm
        } else if constexpr (OP == B2S_OP_LOG10_F32) {
            ((float *)vout)[j] = param * log10f(((const float *)vin)[j]);
        }
    }
}

template <int OP>
int32_t launch(b2s_apply *a, const void *in, void *out, size_t n) {
    b2s_ctx *ctx = a->ctx;
    const int th = 256;
    const size_t want = ceil_div(n, (size_t)th);
    const unsigned grid = (unsigned)std::min<size_t>(want, (size_t)ctx->sm_count * 16);
    apply_kernel<OP><<<grid, th, 0, ctx->stream>>>(in, out, (long long)n, a->param, a->d_carry);
    B2S_CHECK_LAUNCH(ctx);
    return B2S_OK;
}

bool in_is_complex(b2s_op op) {
    return op == B2S_OP_SCALE_C32 || op == B2S_OP_QUAD_DEMOD || op == B2S_OP_NORM_SQR ||
           op == B2S_OP_QUAD_DEMOD_C32 || op == B2S_OP_MAG_C32;
}

}  // namespace

extern "C" {

int32_t b2s_apply_create(b2s_ctx *ctx, b2s_op op, float param, b2s_apply **out) {
    if (!ctx || !out) return b2s_fail(ctx, B2S_EINVAL, "b2s_apply_create: NULL argument");
    *out = nullptr;
    if ((int)op < 0 || (int)op > (int)B2S_OP_LOG10_F32) return b2s_fail(ctx, B2S_EINVAL, "b2s_apply_create: bad op %d", (int)op);
    DeviceGuard g(ctx->device);
    b2s_apply *a = new b2s_apply();
    a->ctx = ctx; a->op = op; a->param = param;
    cudaError_t e = cudaMalloc((void **)&a->d_carry, sizeof(float2));
    if (e != cudaSuccess) { delete a; return b2s_fail(ctx, B2S_ENOMEM, "apply state"); }
    *out = a;
    return b2s_apply_reset(a);
}

void b2s_apply_destroy(b2s_apply *a) {
    if (!a) return;
    DeviceGuard g(a->ctx->device);
    cudaStreamSynchronize(a->ctx->stream);
    cudaFree(a->d_carry);
    delete a;
}

// `let mut last = Complex32::new(0.0, 0.0)` (examples/fm-receiver/src/main.rs:98)
int32_t b2s_apply_reset(b2s_apply *a) {
    if (!a) return b2s_fail(nullptr, B2S_EINVAL, "apply is NULL");
    DeviceGuard g(a->ctx->device);
    B2S_CUDA(a->ctx, cudaMemsetAsync(a->d_carry, 0, sizeof(float2), a->ctx->stream));
    return B2S_OK;
}

int32_t b2s_apply_exec(b2s_apply *a, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                       size_t *consumed, size_t *produced) {
    if (!a || !consumed || !produced) return b2s_fail(a ? a->ctx : nullptr, B2S_EINVAL, "b2s_apply_exec: NULL argument");
    const size_t m = n_in < n_out_cap ? n_in : n_out_cap;        // apply.rs:109
    *consumed = m; *produced = m;
    if (m == 0) return B2S_OK;
    if (!d_in || !d_out) return b2s_fail(a->ctx, B2S_EINVAL, "b2s_apply_exec: NULL buffer");
    // the demodulators read in[j-1] while a neighbour thread writes out[j-1]: the slices must not overlap
    // (the element-wise ops may run in place)
    if (a->op == B2S_OP_QUAD_DEMOD || a->op == B2S_OP_QUAD_DEMOD_C32) {
        const char *i0 = (const char *)d_in, *i1 = i0 + m * sizeof(float2);
        const char *o0 = (const char *)d_out, *o1 = o0 + m * (a->op == B2S_OP_QUAD_DEMOD ? sizeof(float) : sizeof(float2));
        if (i0 < o1 && o0 < i1)
            return b2s_fail(a->ctx, B2S_EINVAL, "b2s_apply_exec: the quadrature demodulator cannot run in place (input and output overlap)");
    }
    DeviceGuard g(a->ctx->device);
    NvtxRange nvtx("b2s_apply_exec");
    int32_t rc = B2S_EINVAL;
    switch (a->op) {
        case B2S_OP_SCALE_F32: rc = launch<B2S_OP_SCALE_F32>(a, d_in, d_out, m); break;
        case B2S_OP_SCALE_C32: rc = launch<B2S_OP_SCALE_C32>(a, d_in, d_out, m); break;
        case B2S_OP_QUAD_DEMOD: rc = launch<B2S_OP_QUAD_DEMOD>(a, d_in, d_out, m); break;
        case B2S_OP_NORM_SQR: rc = launch<B2S_OP_NORM_SQR>(a, d_in, d_out, m); break;
        case B2S_OP_QUAD_DEMOD_C32: rc = launch<B2S_OP_QUAD_DEMOD_C32>(a, d_in, d_out, m); break;
        case B2S_OP_EXP_F32: rc = launch<B2S_OP_EXP_F32>(a, d_in, d_out, m); break;
        case B2S_OP_MAG_C32: rc = launch<B2S_OP_MAG_C32>(a, d_in, d_out, m); break;
        case B2S_OP_LOG10_F32: rc = launch<B2S_OP_LOG10_F32>(a, d_in, d_out, m); break;
    }
    if (rc != B2S_OK) return rc;
    if (a->op == B2S_OP_QUAD_DEMOD || a->op == B2S_OP_QUAD_DEMOD_C32) {
        // last = in[m-1] for the next call (stream-ordered after the kernel that read the old carry)
        B2S_CUDA(a->ctx, cudaMemcpyAsync(a->d_carry, (const float2 *)d_in + (m - 1), sizeof(float2),
                                         cudaMemcpyDeviceToDevice, a->ctx->stream));
    }
    (void)in_is_complex;
    return B2S_OK;
}

}  // extern "C"
