/*
 * oracle.c -- CPU restatement of FutureSDR's FIR / decimator / resampler / FFT / Apply /
 * PfbArbResampler hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for futuresdr_b200.  It is NOT part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (futuresdr_b200/ + libb200sdr.so) never links or calls it and
 * fails loudly when the CUDA library is missing.
 *
 * The reference (Rust, /root/reference) cannot be compiled in this image (no cargo/rustc),
 * so every function below restates the cited reference lines in plain C with the same
 * evaluation order in f32 (compile with -ffp-contract=off so gcc does not fuse a*b+c,
 * which stable Rust never does).  Pinning: the known-answer vectors of the reference's own
 * unit tests (fir.rs:283-365, decimating_fir.rs:313-489, polyphase_resampling_fir.rs:174-260,
 * tests/fir.rs:7-31, firdes/basic.rs:467-760, special_funs.rs:53-123) are replayed against
 * this file by tests/test_oracle_golden.py.  FFT, PfbArbResampler and the Apply demod have
 * no value-pinning test in the reference: "parity unpinned" for those three.
 *
 * Complex<f32> is interleaved {re, im} (num_complex repr(C)), passed here as float*.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* futuredsp::ComputationStatus, crates/futuredsp/src/lib.rs:33-45 */
enum { ORC_INSUFFICIENT_INPUT = 0, ORC_INSUFFICIENT_OUTPUT = 1, ORC_BOTH_SUFFICIENT = 2 };

static size_t sat_sub(size_t a, size_t b) { return a > b ? a - b : 0; }

/* ------------------------------------------------------------------------------------------
 * FirFilter  -- crates/futuredsp/src/fir.rs:52-91 (fir_kernel_core)
 *   n = min(len(i)+1-ntaps, len(o)); o[k] = sum_t mac(sum, i[k+t], taps[ntaps-1-t])
 *   status: fir.rs:70-74
 * ---------------------------------------------------------------------------------------- */
static int fir_counts(size_t n_in, size_t ntaps, size_t n_out_cap, size_t *n) {
    size_t producable = sat_sub(n_in + 1, ntaps);            /* fir.rs:69 */
    if (producable > n_out_cap) { *n = n_out_cap; return ORC_INSUFFICIENT_OUTPUT; }
    if (producable == n_out_cap) { *n = producable; return ORC_BOTH_SUFFICIENT; }
    *n = producable; return ORC_INSUFFICIENT_INPUT;
}

/* f32 x f32, stable-Rust mac `accum + sample * tap`  (fir.rs:206-215) */
int orc_fir_f32_f32(const float *taps, size_t ntaps, const float *in, size_t n_in,
                    float *out, size_t n_out_cap, size_t *consumed, size_t *produced) {
    size_t n; int st = fir_counts(n_in, ntaps, n_out_cap, &n);
    for (size_t k = 0; k < n; k++) {
        float sum = 0.0f;
        for (size_t t = 0; t < ntaps; t++) sum = sum + in[k + t] * taps[ntaps - 1 - t];
        out[k] = sum;
    }
    *consumed = n; *produced = n; return st;
}

/* f64 x f64 (fir.rs:217-226) -- only used for the reference's f64 known-answer test */
int orc_fir_f64_f64(const double *taps, size_t ntaps, const double *in, size_t n_in,
                    double *out, size_t n_out_cap, size_t *consumed, size_t *produced) {
    size_t n; int st = fir_counts(n_in, ntaps, n_out_cap, &n);
    for (size_t k = 0; k < n; k++) {
        double sum = 0.0;
        for (size_t t = 0; t < ntaps; t++) sum = sum + in[k + t] * taps[ntaps - 1 - t];
        out[k] = sum;
    }
    *consumed = n; *produced = n; return st;
}

/* Complex<f32> x f32: re/im accumulate separately (fir.rs:228-255) */
int orc_fir_c32_f32(const float *taps, size_t ntaps, const float *in, size_t n_in,
                    float *out, size_t n_out_cap, size_t *consumed, size_t *produced) {
    size_t n; int st = fir_counts(n_in, ntaps, n_out_cap, &n);
    for (size_t k = 0; k < n; k++) {
        float re = 0.0f, im = 0.0f;
        for (size_t t = 0; t < ntaps; t++) {
            float tap = taps[ntaps - 1 - t];
            re = re + in[2 * (k + t)] * tap;
            im = im + in[2 * (k + t) + 1] * tap;
        }
        out[2 * k] = re; out[2 * k + 1] = im;
    }
    *consumed = n; *produced = n; return st;
}

/* Complex<f32> x Complex<f32>: `accum + sample * tap` with num_complex Mul
 * (re = a.re*b.re - a.im*b.im, im = a.re*b.im + a.im*b.re)  (fir.rs:257-276) */
int orc_fir_c32_c32(const float *taps, size_t ntaps, const float *in, size_t n_in,
                    float *out, size_t n_out_cap, size_t *consumed, size_t *produced) {
    size_t n; int st = fir_counts(n_in, ntaps, n_out_cap, &n);
    for (size_t k = 0; k < n; k++) {
        float re = 0.0f, im = 0.0f;
        for (size_t t = 0; t < ntaps; t++) {
            float tr = taps[2 * (ntaps - 1 - t)], ti = taps[2 * (ntaps - 1 - t) + 1];
            float sr = in[2 * (k + t)], si = in[2 * (k + t) + 1];
            float pr = sr * tr - si * ti;
            float pi = sr * ti + si * tr;
            re = re + pr; im = im + pi;
        }
        out[2 * k] = re; out[2 * k + 1] = im;
    }
    *consumed = n; *produced = n; return st;
}

/* f64-accumulating arbiter of the same sum (not a reference function): used by tests to
 * show both the reference-order f32 result and the GPU result sit within tolerance of the
 * exact value. */
void orc_fir_c32_f32_exact(const float *taps, size_t ntaps, const float *in, size_t n_out,
                           size_t decim, double *out) {
    for (size_t k = 0; k < n_out; k++) {
        double re = 0.0, im = 0.0;
        size_t base = decim - 1 + k * decim;
        for (size_t t = 0; t < ntaps; t++) {
            double tap = taps[ntaps - 1 - t];
            re += (double)in[2 * (base + t)] * tap;
            im += (double)in[2 * (base + t) + 1] * tap;
        }
        out[2 * k] = re; out[2 * k + 1] = im;
    }
}

/* ------------------------------------------------------------------------------------------
 * DecimatingFirFilter -- crates/futuredsp/src/decimating_fir.rs:53-95
 *   consumable = (len+1-ntaps)/D; input index D-1 + k*D + t (:86); returns (n*D, n) (:94)
 * ---------------------------------------------------------------------------------------- */
static int decim_counts(size_t n_in, size_t ntaps, size_t decim, size_t n_out_cap, size_t *n) {
    size_t filterable = sat_sub(n_in + 1, ntaps);             /* :70 */
    size_t consumable = filterable / decim;                    /* :71 */
    if (consumable > n_out_cap) { *n = n_out_cap; return ORC_INSUFFICIENT_OUTPUT; }
    if (consumable == n_out_cap) { *n = n_out_cap; return ORC_BOTH_SUFFICIENT; }
    *n = consumable; return ORC_INSUFFICIENT_INPUT;
}

int orc_decim_f32_f32(const float *taps, size_t ntaps, size_t decim, const float *in,
                      size_t n_in, float *out, size_t n_out_cap, size_t *consumed,
                      size_t *produced) {
    size_t n; int st = decim_counts(n_in, ntaps, decim, n_out_cap, &n);
    for (size_t k = 0; k < n; k++) {
        float sum = 0.0f;
        for (size_t t = 0; t < ntaps; t++)
            sum = sum + in[decim - 1 + k * decim + t] * taps[ntaps - 1 - t];
        out[k] = sum;
    }
    *consumed = n * decim; *produced = n; return st;
}

int orc_decim_c32_f32(const float *taps, size_t ntaps, size_t decim, const float *in,
                      size_t n_in, float *out, size_t n_out_cap, size_t *consumed,
                      size_t *produced) {
    size_t n; int st = decim_counts(n_in, ntaps, decim, n_out_cap, &n);
    for (size_t k = 0; k < n; k++) {
        float re = 0.0f, im = 0.0f;
        size_t base = decim - 1 + k * decim;
        for (size_t t = 0; t < ntaps; t++) {
            float tap = taps[ntaps - 1 - t];
            re = re + in[2 * (base + t)] * tap;
            im = im + in[2 * (base + t) + 1] * tap;
        }
        out[2 * k] = re; out[2 * k + 1] = im;
    }
    *consumed = n * decim; *produced = n; return st;
}

int orc_decim_c32_c32(const float *taps, size_t ntaps, size_t decim, const float *in,
                      size_t n_in, float *out, size_t n_out_cap, size_t *consumed,
                      size_t *produced) {
    size_t n; int st = decim_counts(n_in, ntaps, decim, n_out_cap, &n);
    for (size_t k = 0; k < n; k++) {
        float re = 0.0f, im = 0.0f;
        size_t base = decim - 1 + k * decim;
        for (size_t t = 0; t < ntaps; t++) {
            float tr = taps[2 * (ntaps - 1 - t)], ti = taps[2 * (ntaps - 1 - t) + 1];
            float sr = in[2 * (base + t)], si = in[2 * (base + t) + 1];
            float pr = sr * tr - si * ti;
            float pi = sr * ti + si * tr;
            re = re + pr; im = im + pi;
        }
        out[2 * k] = re; out[2 * k + 1] = im;
    }
    *consumed = n * decim; *produced = n; return st;
}

/* ------------------------------------------------------------------------------------------
 * PolyphaseResamplingFir -- crates/futuredsp/src/polyphase_resampling_fir.rs:70-124
 * ---------------------------------------------------------------------------------------- */
static int resamp_counts(size_t n_in, size_t ntaps_total, size_t interp, size_t decim,
                         size_t n_out_cap, size_t *np, size_t *ncons) {
    size_t num_taps = ntaps_total / interp;                                     /* :92 */
    size_t p = sat_sub(sat_sub(n_in + 1, num_taps) * interp, 1) / decim;        /* :93-94 */
    p = (p / interp) * interp;                                                  /* :96 */
    int st;
    if (p > n_out_cap) { p = (n_out_cap / interp) * interp; st = ORC_INSUFFICIENT_OUTPUT; }
    else if (p == n_out_cap) st = ORC_BOTH_SUFFICIENT;
    else st = ORC_INSUFFICIENT_INPUT;
    *np = p; *ncons = (p / interp) * decim;                                     /* :106 */
    return st;
}

int orc_resamp_f32_f32(const float *taps, size_t ntaps_total, size_t interp, size_t decim,
                       const float *in, size_t n_in, float *out, size_t n_out_cap,
                       size_t *consumed, size_t *produced) {
    size_t p, c; int st = resamp_counts(n_in, ntaps_total, interp, decim, n_out_cap, &p, &c);
    size_t num_taps = ntaps_total / interp;
    for (size_t k = 0; k < p; k++) {
        size_t bank = (k * decim) % interp;                                     /* :110 */
        size_t in0 = k * decim / interp;                                        /* :111 */
        float sum = 0.0f;
        for (size_t t = 0; t < num_taps; t++) {
            size_t tap_idx = interp * (num_taps - t - 1) + bank;                /* :114 */
            sum = sum + in[in0 + t] * taps[tap_idx];
        }
        out[k] = sum;
    }
    *consumed = c; *produced = p; return st;
}

int orc_resamp_c32_f32(const float *taps, size_t ntaps_total, size_t interp, size_t decim,
                       const float *in, size_t n_in, float *out, size_t n_out_cap,
                       size_t *consumed, size_t *produced) {
    size_t p, c; int st = resamp_counts(n_in, ntaps_total, interp, decim, n_out_cap, &p, &c);
    size_t num_taps = ntaps_total / interp;
    for (size_t k = 0; k < p; k++) {
        size_t bank = (k * decim) % interp;
        size_t in0 = k * decim / interp;
        float re = 0.0f, im = 0.0f;
        for (size_t t = 0; t < num_taps; t++) {
            float tap = taps[interp * (num_taps - t - 1) + bank];
            re = re + in[2 * (in0 + t)] * tap;
            im = im + in[2 * (in0 + t) + 1] * tap;
        }
        out[2 * k] = re; out[2 * k + 1] = im;
    }
    *consumed = c; *produced = p; return st;
}

/* ------------------------------------------------------------------------------------------
 * Tap design (f64, runs once)
 *   besseli0      crates/futuredsp/src/math/special_funs.rs:22-45
 *   windows::kaiser  crates/futuredsp/src/windows.rs:144-152
 *   firdes::lowpass  crates/futuredsp/src/firdes/basic.rs:25-42
 *   kaiser::{lowpass,multirate,compute_kaiser_beta,design_kaiser_window} basic.rs:310-459
 * ---------------------------------------------------------------------------------------- */
static double powi(double x, int n) {          /* f64::powi: repeated multiplication */
    if (n < 0) return 1.0 / powi(x, -n);
    double r = 1.0;
    while (n) { if (n & 1) r *= x; x *= x; n >>= 1; }
    return r;
}

double orc_besseli0(double x) {
    double t = x / 3.75;
    if (fabs(x) <= 3.75) {
        return 1.0 + 3.5156229 * powi(t, 2) + 3.0899424 * powi(t, 4) + 1.2067492 * powi(t, 6)
             + 0.2659732 * powi(t, 8) + 0.0360768 * powi(t, 10) + 0.0045813 * powi(t, 12);
    }
    return (1.0 / (sqrt(fabs(x)) * exp(-x)))
         * (0.39894228 + 0.01328592 * powi(t, -1) + 0.00225319 * powi(t, -2)
            - 0.00157565 * powi(t, -3) + 0.00916281 * powi(t, -4) - 0.02057706 * powi(t, -5)
            + 0.02635537 * powi(t, -6) - 0.01647633 * powi(t, -7) + 0.00392377 * powi(t, -8));
}

void orc_window_kaiser(size_t len, double beta, double *w) {
    double alpha = (double)(len - 1) / 2.0;
    for (size_t n = 0; n < len; n++) {
        double r = ((double)n - alpha) / alpha;
        double x = beta * sqrt(1.0 - r * r);
        w[n] = orc_besseli0(x) / orc_besseli0(beta);
    }
}

/* firdes::lowpass(cutoff, window) -> f64 taps (cast to f32 by caller: T::from_f64) */
void orc_firdes_lowpass(double cutoff, const double *window, size_t len, double *taps) {
    const double PI = 3.14159265358979323846264338327950288;
    double omega_c = 2.0 * PI * cutoff;
    double alpha = (double)(len - 1) / 2.0;
    for (size_t n = 0; n < len; n++) {
        double x = (double)n - alpha;
        double ft = (x == 0.0) ? omega_c / PI : sin(omega_c * x) / (PI * x);
        taps[n] = window[n] * ft;
    }
}

double orc_kaiser_beta(double max_ripple) {                       /* basic.rs:444-452 */
    double ripple_db = -20.0 * log10(max_ripple);
    if (ripple_db > 50.0) return 0.1102 * (ripple_db - 8.7);
    if (ripple_db >= 21.0) return 0.5842 * pow(ripple_db - 21.0, 0.4) + 0.07886 * (ripple_db - 21.0);
    return 0.0;
}

size_t orc_kaiser_num_taps(double transition_bw, double max_ripple) {   /* basic.rs:454-459 */
    double ripple_db = -20.0 * log10(max_ripple);
    return (size_t)(ceil((ripple_db - 7.95) / (14.36 * transition_bw)) + 1.0);
}

/* kaiser::lowpass (basic.rs:310-321). Returns ntaps; writes up to cap taps (f64). */
size_t orc_kaiser_lowpass(double cutoff, double transition_bw, double max_ripple,
                          double *taps, size_t cap) {
    size_t n = orc_kaiser_num_taps(transition_bw, max_ripple);
    if (taps == NULL || cap < n) return n;
    double beta = orc_kaiser_beta(max_ripple);
    double *win = (double *)malloc(n * sizeof(double));
    orc_window_kaiser(n, beta, win);
    double omega_c = (2.0 * cutoff + transition_bw) / 2.0;
    orc_firdes_lowpass(omega_c, win, n, taps);
    free(win);
    return n;
}

/* kaiser::multirate (basic.rs:412-442). Returns ntaps. */
size_t orc_kaiser_multirate(size_t interp, size_t decim, size_t half_len, double max_ripple,
                            double *taps, size_t cap) {
    if (interp == 1 && decim == 1) { if (taps && cap >= 1) taps[0] = 1.0; return 1; }
    size_t band = (interp == 1) ? decim : interp;
    size_t n = 2 * half_len * band;
    if (taps == NULL || cap < n) return n;
    double beta = orc_kaiser_beta(max_ripple);
    double *win = (double *)malloc((n + 1) * sizeof(double));
    double *full = (double *)malloc((n + 1) * sizeof(double));
    orc_window_kaiser(n + 1, beta, win);
    for (size_t i = 0; i < n + 1; i++) win[i] = (double)interp * win[i];
    size_t mx = interp > decim ? interp : decim;
    double omega_c = 1.0 / (2.0 * (double)mx);
    orc_firdes_lowpass(omega_c, win, n + 1, full);
    memcpy(taps, full, n * sizeof(double));                          /* truncate(num_taps) */
    free(win); free(full);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Fft block -- src/blocks/fft.rs:160-221.  Arithmetic lives in rustfft 6.4 (crates.io, not
 * vendored): forward X[k] = sum_n x[n] e^{-2 pi i k n / N}, inverse unnormalised e^{+...}.
 * This oracle evaluates the DFT in f64 (radix-2 when N is a power of two, direct O(N^2)
 * otherwise) and applies the block's shift / normalize exactly as fft.rs:179-210.
 * "Parity unpinned": the reference has no test that pins an Fft value.
 * ---------------------------------------------------------------------------------------- */
static void dft_f64(const double *xr, const double *xi, double *yr, double *yi, size_t n,
                    int inverse) {
    const double PI = 3.14159265358979323846264338327950288;
    double sgn = inverse ? 1.0 : -1.0;
    if ((n & (n - 1)) == 0 && n > 1) {
        /* iterative radix-2 DIT in f64 with exactly-evaluated twiddles */
        size_t lg = 0; while (((size_t)1 << lg) < n) lg++;
        for (size_t i = 0; i < n; i++) {
            size_t r = 0;
            for (size_t b = 0; b < lg; b++) if (i & ((size_t)1 << b)) r |= (size_t)1 << (lg - 1 - b);
            yr[r] = xr[i]; yi[r] = xi[i];
        }
        for (size_t len = 2; len <= n; len <<= 1) {
            size_t half = len >> 1;
            for (size_t j = 0; j < half; j++) {
                double ang = sgn * 2.0 * PI * (double)j / (double)len;
                double wr = cos(ang), wi = sin(ang);
                for (size_t s = 0; s < n; s += len) {
                    size_t a = s + j, b = a + half;
                    double tr = yr[b] * wr - yi[b] * wi, ti = yr[b] * wi + yi[b] * wr;
                    yr[b] = yr[a] - tr; yi[b] = yi[a] - ti;
                    yr[a] += tr; yi[a] += ti;
                }
            }
        }
        return;
    }
    for (size_t k = 0; k < n; k++) {
        double sr = 0.0, si = 0.0;
        for (size_t m = 0; m < n; m++) {
            size_t km = (k * m) % n;
            double ang = sgn * 2.0 * PI * (double)km / (double)n;
            double wr = cos(ang), wi = sin(ang);
            sr += xr[m] * wr - xi[m] * wi;
            si += xr[m] * wi + xi[m] * wr;
        }
        yr[k] = sr; yi[k] = si;
    }
}

/* One work() call over everything available: m = min(n_in, n_out_cap) rounded down to a
 * multiple of len (fft.rs:169-170).  The 32-FFT cap (:171) only limits how much ONE call
 * does; a Mocker-style loop repeats until m == 0, so the stream result is the same.
 * normalize: pass has_norm = 0 for None.  Returns m (= consumed = produced). */
size_t orc_fft_block_c32(size_t len, int inverse, int fft_shift, int has_norm, float norm,
                         const float *in, size_t n_in, float *out, size_t n_out_cap) {
    size_t m = n_in < n_out_cap ? n_in : n_out_cap;
    m = (m / len) * len;
    double *xr = (double *)malloc(4 * len * sizeof(double));
    double *xi = xr + len, *yr = xi + len, *yi = yr + len;
    for (size_t f = 0; f < m / len; f++) {
        const float *src = in + 2 * f * len;
        float *dst = out + 2 * f * len;
        for (size_t k = 0; k < len; k++) {
            size_t s = (inverse && fft_shift) ? (k + len / 2) % len : k;      /* :179-185 */
            xr[k] = src[2 * s]; xi[k] = src[2 * s + 1];
        }
        dft_f64(xr, xi, yr, yi, len, inverse);
        for (size_t k = 0; k < len; k++) {
            size_t s = (!inverse && fft_shift) ? (k + len / 2) % len : k;     /* :196-204 */
            float re = (float)yr[s], im = (float)yi[s];
            if (has_norm) { re = re * norm; im = im * norm; }                 /* :206-210 */
            dst[2 * k] = re; dst[2 * k + 1] = im;
        }
    }
    free(xr);
    return m;
}

/* ------------------------------------------------------------------------------------------
 * Apply closures used on the path (src/blocks/apply.rs:100-131 calls f per sample)
 * ---------------------------------------------------------------------------------------- */
/* FM quadrature demod, examples/fm-receiver/src/main.rs:99-104:
 *   arg = (v * last.conj()).arg(); last = *v;      carry = last (2 floats, in/out) */
void orc_apply_quad_demod(const float *in, size_t n, float *out, float *carry) {
    float lr = carry[0], li = carry[1];
    for (size_t j = 0; j < n; j++) {
        float vr = in[2 * j], vi = in[2 * j + 1];
        float cr = lr, ci = -li;                       /* conj */
        float pr = vr * cr - vi * ci;                  /* num_complex Mul */
        float pi = vr * ci + vi * cr;
        out[j] = atan2f(pi, pr);                       /* Complex::arg = im.atan2(re) */
        lr = vr; li = vi;
    }
    carry[0] = lr; carry[1] = li;
}

/* tests/vulkan.rs:16-27, blocks/wgpu.rs:22-32: x *= 12.0 (generalised to a constant) */
void orc_apply_scale_f32(const float *in, size_t n, float k, float *out) {
    for (size_t j = 0; j < n; j++) out[j] = in[j] * k;
}

/* examples/spectrum/src/bin/cpu.rs: |x|^2 (Complex::norm_sqr = re*re + im*im) */
void orc_apply_norm_sqr(const float *in, size_t n, float *out) {
    for (size_t j = 0; j < n; j++) out[j] = in[2 * j] * in[2 * j] + in[2 * j + 1] * in[2 * j + 1];
}

/* ------------------------------------------------------------------------------------------
 * PfbArbResampler -- src/blocks/pfb/arb_resampler.rs:90-231, pfb/utilities.rs:5-25,
 * pfb/window_buffer.rs:13-44.  Restated as a streaming object so tests can drive it
 * call-by-call exactly like Kernel::work.  "Parity unpinned" (no reference test).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    size_t num_filters, taps_per_filter;
    float *arms;                 /* [num_filters][taps_per_filter], utilities.rs order (NOT reversed) */
    /* WindowBuffer */
    size_t buffer_len, start_idx, num_samples_missing;
    float *circ;                 /* 2*buffer_len complex */
    /* State */
    float rate, delay, buff[4], tau, bf, mu;
    int boundary;                /* ResampState::Boundary */
    size_t base_index;
} orc_pfbarb;

orc_pfbarb *orc_pfbarb_new(float rate, const float *taps, size_t ntaps, size_t num_filters) {
    orc_pfbarb *s = (orc_pfbarb *)calloc(1, sizeof(orc_pfbarb));
    /* utilities.rs:9: taps_per_filter = ceil(len as f32 / n as f32) */
    size_t tpf = (size_t)ceilf((float)ntaps / (float)num_filters);
    s->num_filters = num_filters; s->taps_per_filter = tpf;
    s->arms = (float *)calloc(num_filters * tpf, sizeof(float));
    for (size_t i = 0; i < num_filters; i++) {
        size_t j = 0;
        for (size_t idx = i; idx < ntaps; idx += num_filters) s->arms[i * tpf + j++] = taps[idx];
        /* zero pad to tpf (utilities.rs:11-19) -- calloc did it */
    }
    s->buffer_len = tpf; s->start_idx = 0; s->num_samples_missing = tpf;   /* pad_start=false */
    s->circ = (float *)calloc(4 * tpf, sizeof(float));
    s->rate = rate; s->delay = 1.0f / rate;
    s->tau = 0.0f; s->bf = 0.0f; s->base_index = 0; s->mu = 0.0f; s->boundary = 0;
    return s;
}

void orc_pfbarb_free(orc_pfbarb *s) { if (s) { free(s->arms); free(s->circ); free(s); } }

static void wb_push(orc_pfbarb *s, float re, float im) {          /* window_buffer.rs:24-32 */
    long L = (long)s->buffer_len;
    long idx = ((long)s->start_idx - (long)s->num_samples_missing) % L;
    if (idx < 0) idx += L;                                         /* rem_euclid */
    s->circ[2 * idx] = re; s->circ[2 * idx + 1] = im;
    s->circ[2 * (idx + L)] = re; s->circ[2 * (idx + L) + 1] = im;
    if (s->num_samples_missing > 0) s->num_samples_missing--;
    s->start_idx = (s->start_idx + 1) % s->buffer_len;
}

/* FirFilter::filter on the window with 1 output: sum_t win[t] * arm[T-1-t], strict order */
static void arm_filter(const orc_pfbarb *s, size_t arm, float *o) {
    const float *win = s->circ + 2 * s->start_idx;                 /* get_as_slice */
    const float *a = s->arms + arm * s->taps_per_filter;
    size_t T = s->taps_per_filter;
    float re = 0.0f, im = 0.0f;
    for (size_t t = 0; t < T; t++) {
        float tap = a[T - 1 - t];
        re = re + win[2 * t] * tap; im = im + win[2 * t + 1] * tap;
    }
    o[0] = re; o[1] = im;
}

static void update_timing_state(orc_pfbarb *s) {                   /* arb_resampler.rs:132-140 */
    s->tau += s->delay;
    s->bf = s->tau * (float)s->num_filters;
    s->base_index = (size_t)floorf(s->bf);
    s->mu = s->bf - (float)s->base_index;
}

static size_t consume_single(orc_pfbarb *s, float re, float im, float *out) {   /* :142-188 */
    wb_push(s, re, im);
    size_t produced = 0;
    while (s->base_index < s->num_filters) {
        if (s->boundary) {
            arm_filter(s, 0, s->buff + 2);
            /* (1.0 - mu) * buff[0] + mu * buff[1]: f32 * Complex = {f*re, f*im} */
            float a = 1.0f - s->mu;
            out[2 * produced] = a * s->buff[0] + s->mu * s->buff[2];
            out[2 * produced + 1] = a * s->buff[1] + s->mu * s->buff[3];
            produced++;
            update_timing_state(s);
            s->boundary = 0;
        } else {
            arm_filter(s, s->base_index, s->buff);
            if (s->base_index == s->num_filters - 1) {
                s->boundary = 1;
                s->base_index = s->num_filters;
            } else {
                arm_filter(s, s->base_index + 1, s->buff + 2);
                float a = 1.0f - s->mu;
                out[2 * produced] = a * s->buff[0] + s->mu * s->buff[2];
                out[2 * produced + 1] = a * s->buff[1] + s->mu * s->buff[3];
                produced++;
                update_timing_state(s);
            }
        }
    }
    s->tau -= 1.0f;
    s->bf -= (float)s->num_filters;
    s->base_index -= s->num_filters;
    return produced;
}

/* One Kernel::work call (arb_resampler.rs:193-231).  Returns via pointers; *call_again is
 * io.call_again.  Output capacity n_out_cap in complex items. */
void orc_pfbarb_work(orc_pfbarb *s, const float *in, size_t n_in, float *out, size_t n_out_cap,
                     size_t *consumed, size_t *produced, int *call_again) {
    *consumed = 0; *produced = 0; *call_again = 0;
    if (s->num_samples_missing != 0) {
        size_t c = 0;
        while (s->num_samples_missing != 0 && c < n_in) { wb_push(s, in[2 * c], in[2 * c + 1]); c++; }
        *consumed = c;
        if (n_in - c > 0) *call_again = 1;
        return;
    }
    size_t cap = (size_t)((float)n_out_cap / s->rate);             /* :218 */
    size_t n = n_in < cap ? n_in : cap;
    size_t p = 0;
    for (size_t j = 0; j < n; j++) p += consume_single(s, in[2 * j], in[2 * j + 1], out + 2 * p);
    *consumed = n; *produced = p;
}

/* ------------------------------------------------------------------------------------------
 * CPU baseline legs (bench.py cpu_baseline / --impl reference): the same strict-order loop
 * as orc_fir_c32_f32 over contiguous shards on all host threads (≙ `smoln` over pipes,
 * perf/fir/fir.rs:80-84).  Shard s produces outputs [s*per, (s+1)*per).
 * ---------------------------------------------------------------------------------------- */
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_fir_c32_f32_mt(const float *taps, size_t ntaps, const float *in, size_t n_in,
                        float *out, int threads) {
    if (n_in + 1 <= ntaps) return;
    size_t n = n_in + 1 - ntaps;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static)
#endif
    for (long k = 0; k < (long)n; k++) {
        float re = 0.0f, im = 0.0f;
        for (size_t t = 0; t < ntaps; t++) {
            float tap = taps[ntaps - 1 - t];
            re = re + in[2 * ((size_t)k + t)] * tap;
            im = im + in[2 * ((size_t)k + t) + 1] * tap;
        }
        out[2 * k] = re; out[2 * k + 1] = im;
    }
    (void)threads;
}

/* ------------------------------------------------------------------------------------------
 * Rotator -- crates/futuredsp/src/rotator.rs:13-48 ("next" row).
 *   new(phase_incr): incr = Complex32::from_polar(1.0, phase_incr) = (cos, sin) in f32; phase = (1, 0)
 *   per sample: phase *= incr; out = in * phase   (num_complex Mul, separate roundings)
 * state: phase (2 floats, in/out).  "Parity unpinned": no reference test pins a Rotator value.
 * ---------------------------------------------------------------------------------------- */
void orc_rotator_incr(float phase_incr, float *incr) {
    incr[0] = 1.0f * cosf(phase_incr);         /* from_polar(r, theta) = (r*cos(theta), r*sin(theta)) */
    incr[1] = 1.0f * sinf(phase_incr);
}

void orc_rotator_rotate(const float *in, size_t n, float *out, const float *incr, float *phase) {
    float pr = phase[0], pi = phase[1];
    const float ir = incr[0], ii = incr[1];
    for (size_t j = 0; j < n; j++) {
        const float nr = pr * ir - pi * ii;    /* phase *= phase_incr */
        const float ni = pr * ii + pi * ir;
        pr = nr; pi = ni;
        const float xr = in[2 * j], xi = in[2 * j + 1];
        out[2 * j] = xr * pr - xi * pi;        /* *v *= phase */
        out[2 * j + 1] = xr * pi + xi * pr;
    }
    phase[0] = pr; phase[1] = pi;
}

/* XlatingFir band-pass taps -- src/blocks/xlating_fir.rs:80-86:
 *   bpf[i] = Complex32::from_polar(1.0, i as f32 * TAU * offset / sample_rate) * tap[i] */
void orc_xlating_taps(const float *taps, size_t ntaps, float offset, float sample_rate, float *bpf) {
    const float TAU = 6.28318530717958647692f;
    for (size_t i = 0; i < ntaps; i++) {
        const float th = (float)i * TAU * offset / sample_rate;
        bpf[2 * i] = (1.0f * cosf(th)) * taps[i];
        bpf[2 * i + 1] = (1.0f * sinf(th)) * taps[i];
    }
}

/* ------------------------------------------------------------------------------------------
 * PfbChannelizer -- src/blocks/pfb/channelizer.rs:88-223 ("next" row), with
 * partition_filter_taps (pfb/utilities.rs:5-25) and WindowBuffer (pfb/window_buffer.rs:13-44).
 * Streaming object driven call-by-call like Kernel::work.  Outputs are written channel-major:
 * out[ch * out_stride + k].  The IFFT (rustfft, un-normalised inverse) is evaluated in f64.
 * "Parity unpinned": no reference test constructs this block.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    size_t N, D, T;              /* channels, decimation_factor, taps per arm */
    float *arms;                 /* [N][T] utilities.rs order */
    float *circ;                 /* [N][2T] complex */
    size_t *start_idx, *missing; /* per window */
    size_t base_index;
    int all_filled;
} orc_chan;

orc_chan *orc_chan_new(size_t num_channels, const float *taps, size_t ntaps, float oversample_rate) {
    orc_chan *s = (orc_chan *)calloc(1, sizeof(orc_chan));
    size_t N = num_channels;
    s->N = N; s->D = (size_t)((float)N / oversample_rate);                    /* channelizer.rs:106 */
    size_t T = (size_t)ceilf((float)ntaps / (float)N);
    s->T = T;
    s->arms = (float *)calloc(N * T, sizeof(float));
    for (size_t i = 0; i < N; i++) { size_t j = 0; for (size_t idx = i; idx < ntaps; idx += N) s->arms[i * T + j++] = taps[idx]; }
    s->circ = (float *)calloc(N * 4 * T, sizeof(float));
    s->start_idx = (size_t *)calloc(N, sizeof(size_t));
    s->missing = (size_t *)calloc(N, sizeof(size_t));
    for (size_t i = 0; i < N; i++) s->missing[i] = T;
    s->base_index = N - 1; s->all_filled = 0;
    return s;
}
void orc_chan_free(orc_chan *s) { if (s) { free(s->arms); free(s->circ); free(s->start_idx); free(s->missing); free(s); } }

static void chan_push(orc_chan *s, size_t w, float re, float im) {              /* window_buffer.rs:24-32 */
    long L = (long)s->T;
    long idx = ((long)s->start_idx[w] - (long)s->missing[w]) % L; if (idx < 0) idx += L;
    float *c = s->circ + w * 4 * s->T;
    c[2 * idx] = re; c[2 * idx + 1] = im; c[2 * (idx + L)] = re; c[2 * (idx + L) + 1] = im;
    if (s->missing[w] > 0) s->missing[w]--;
    s->start_idx[w] = (s->start_idx[w] + 1) % s->T;
}
static int chan_all_filled(const orc_chan *s) { for (size_t i = 0; i < s->N; i++) if (s->missing[i]) return 0; return 1; }
static void chan_dec(orc_chan *s) { s->base_index = s->base_index == 0 ? s->N - 1 : s->base_index - 1; }

/* one work() call; n_out_cap = min over channels of the output slice length.  Returns consumed and
 * produced-per-channel; *call_again as io.call_again. */
void orc_chan_work(orc_chan *s, const float *in, size_t n_in, float *out, size_t out_stride, size_t n_out_cap,
                   size_t *consumed, size_t *produced, int *call_again) {
    *consumed = 0; *produced = 0; *call_again = 0;
    size_t nprod = n_in / s->D; if (nprod > n_out_cap) nprod = n_out_cap;       /* :155-158 */
    if (!s->all_filled) {                                                      /* :160-181 */
        size_t c = 0;
        while (!chan_all_filled(s)) {
            if (c == n_in) { *consumed = c; return; }
            chan_push(s, s->base_index, in[2 * c], in[2 * c + 1]);
            chan_dec(s); c++;
        }
        s->all_filled = 1;
        if (n_in >= s->D) *call_again = 1;
        /* NB: the reference returns WITHOUT consuming here (:176-180 never call input.consume), so the
         * samples pushed during this call are presented again and pushed a second time. */
        return;
    }
    size_t N = s->N, T = s->T;
    double *br = (double *)malloc(4 * N * sizeof(double)), *bi = br + N, *yr = bi + N, *yi = yr + N;
    for (size_t o = 0; o < nprod; o++) {
        for (size_t j = 0; j < s->D; j++) {
            chan_push(s, s->base_index, in[2 * (o * s->D + j)], in[2 * (o * s->D + j) + 1]);
            chan_dec(s);
        }
        float *fb = (float *)malloc(2 * N * sizeof(float));
        for (size_t i = 0; i < N; i++) {
            size_t bidx = (s->base_index + i + 1) % N;                          /* :190 */
            const float *win = s->circ + bidx * 4 * T + 2 * s->start_idx[bidx];
            const float *a = s->arms + i * T;
            float re = 0.0f, im = 0.0f;
            for (size_t t = 0; t < T; t++) { float tap = a[T - 1 - t]; re = re + win[2 * t] * tap; im = im + win[2 * t + 1] * tap; }
            fb[2 * bidx] = re; fb[2 * bidx + 1] = im;
        }
        for (size_t i = 0; i < N; i++) { br[i] = fb[2 * i]; bi[i] = fb[2 * i + 1]; }
        dft_f64(br, bi, yr, yi, N, 1);                                          /* ifft.process (:201) */
        for (size_t ch = 0; ch < N; ch++) { out[2 * (ch * out_stride + o)] = (float)yr[ch]; out[2 * (ch * out_stride + o) + 1] = (float)yi[ch]; }
        free(fb);
    }
    free(br);
    *consumed = nprod * s->D; *produced = nprod;
}

/* ------------------------------------------------------------------------------------------
 * PfbSynthesizer -- src/blocks/pfb/synthesizer.rs:52-144.  Inputs are channel-major:
 * in[ch * in_stride + v]; n_in = the shortest input slice (:90).  "Parity unpinned".
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    size_t N, T;
    float *arms;                 /* [N][T] utilities.rs order */
    float *circ;                 /* [N][2T] complex */
    size_t start_idx, missing;   /* all windows move in lockstep (one push each per vector) */
    int all_filled;
} orc_synth;

orc_synth *orc_synth_new(size_t num_channels, const float *taps, size_t ntaps) {
    orc_synth *s = (orc_synth *)calloc(1, sizeof(orc_synth));
    size_t N = num_channels, T = (size_t)ceilf((float)ntaps / (float)N);
    s->N = N; s->T = T;
    s->arms = (float *)calloc(N * T, sizeof(float));
    for (size_t i = 0; i < N; i++) { size_t j = 0; for (size_t idx = i; idx < ntaps; idx += N) s->arms[i * T + j++] = taps[idx]; }
    s->circ = (float *)calloc(N * 4 * T, sizeof(float));
    s->start_idx = 0; s->missing = T; s->all_filled = 0;
    return s;
}
void orc_synth_free(orc_synth *s) { if (s) { free(s->arms); free(s->circ); free(s); } }

void orc_synth_work(orc_synth *s, const float *in, size_t in_stride, size_t n_in, float *out, size_t n_out_cap,
                    size_t *consumed_per_channel, size_t *produced) {
    size_t N = s->N, T = s->T, c = 0, p = 0;
    double *br = (double *)malloc(4 * N * sizeof(double)), *bi = br + N, *yr = bi + N, *yi = yr + N;
    while (n_in - c > 0 && (n_out_cap - p > N || !s->all_filled)) {                /* :95-97 */
        for (size_t w = 0; w < N; w++) { br[w] = in[2 * (w * in_stride + c)]; bi[w] = in[2 * (w * in_stride + c) + 1]; }
        c++;
        dft_f64(br, bi, yr, yi, N, 1);                                             /* ifft.process (:104) */
        long L = (long)T;
        long idx = ((long)s->start_idx - (long)s->missing) % L; if (idx < 0) idx += L;
        size_t missing_after = s->missing > 0 ? s->missing - 1 : 0;
        size_t start_after = (s->start_idx + 1) % T;
        for (size_t w = 0; w < N; w++) {
            float *cw = s->circ + w * 4 * T;
            float re = (float)yr[w], im = (float)yi[w];
            cw[2 * idx] = re; cw[2 * idx + 1] = im; cw[2 * (idx + L)] = re; cw[2 * (idx + L) + 1] = im;   /* window.push */
            if (missing_after == 0) {                                              /* window.filled() */
                const float *win = cw + 2 * start_after;
                const float *a = s->arms + w * T;
                float ore = 0.0f, oim = 0.0f;
                for (size_t t = 0; t < T; t++) { float tap = a[T - 1 - t]; ore = ore + win[2 * t] * tap; oim = oim + win[2 * t + 1] * tap; }
                out[2 * p] = ore; out[2 * p + 1] = oim; p++;
            }
        }
        s->missing = missing_after; s->start_idx = start_after;
        if (!s->all_filled) s->all_filled = (s->missing == 0);
    }
    free(br);
    *consumed_per_channel = c; *produced = p;
}

/* ------------------------------------------------------------------------------------------
 * MovingAvg<WIDTH> -- src/blocks/moving_avg.rs:72-115 (spectrum pipe tail).
 * state: avg[width] and the chunk counter *i (in/out).  Returns consumed/produced in items.
 * ---------------------------------------------------------------------------------------- */
void orc_mavg_work(float *avg, size_t *i, size_t width, float decay, size_t history, const float *in,
                   size_t n_in, float *out, size_t n_out_cap, size_t *consumed, size_t *produced) {
    size_t c = 0, p = 0;
    while ((c + 1) * width <= n_in && (p + 1) * width <= n_out_cap) {
        for (size_t b = 0; b < width; b++) {
            float t = in[c * width + b];
            if (isfinite(t)) avg[b] = (1.0f - decay) * avg[b] + decay * t;
            else avg[b] *= 1.0f - decay;
        }
        *i += 1;
        if (*i == history) {
            memcpy(out + p * width, avg, width * sizeof(float));
            *i = 0; p++;
        }
        c++;
    }
    *consumed = c * width; *produced = p * width;
}
