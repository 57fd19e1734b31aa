/*
 * oracle_fast.c -- the reference's NIGHTLY arithmetic for the c32 x f32 FIR, as a CPU baseline.
 * TEST/BENCH INFRASTRUCTURE ONLY (same rules as oracle.c).
 *
 * On nightly rustc the reference uses `algebraic_add/algebraic_mul`
 * (crates/futuredsp/src/fir.rs:119-141), which licenses reassociation and lets LLVM
 * vectorise the tap loop.  The closest C statement is the same loop compiled with
 * -ffast-math (this file is built with -O3 -march=native -ffast-math).  Results differ from
 * the strict-order oracle only by summation order; bench.py reports whichever CPU variant
 * is faster so the GPU is compared against the reference's best case.
 */
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_fir_c32_f32_fast_mt(const float *taps, size_t ntaps, const float *in, size_t n_in,
                             float *out, int threads) {
    if (n_in + 1 <= ntaps) return;
    size_t n = n_in + 1 - ntaps;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static)
#endif
    for (long k = 0; k < (long)n; k++) {
        float re = 0.0f, im = 0.0f;
        const float *x = in + 2 * (size_t)k;
        for (size_t t = 0; t < ntaps; t++) {
            float tap = taps[ntaps - 1 - t];
            re += x[2 * t] * tap;
            im += x[2 * t + 1] * tap;
        }
        out[2 * k] = re; out[2 * k + 1] = im;
    }
    (void)threads;
}
