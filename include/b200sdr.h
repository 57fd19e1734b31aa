/*
 * b200sdr.h -- C ABI of the B200-native streaming-DSP backend for FutureSDR's
 * FIR / decimator / resampler / FFT / Apply / PfbArbResampler hot path.
 *
 * The reference (FutureSDR, Rust) has no FFI for this path: its boundary is a set of Rust
 * traits.  Every entry point below names the reference interface it replaces (file:line is
 * relative to the FutureSDR tree).  INTEGRATION.md shows the Rust `extern "C"` block and the
 * `impl futuredsp::Filter` / `impl Kernel` shims a maintainer would add on top of this header.
 *
 * Conventions
 *   - every function returns int32_t: 0 = B2S_OK, <0 = B2S_E*; b2s_last_error() gives text.
 *     No exception crosses the boundary, no torch/CUDA type appears in a signature
 *     (a CUDA stream is passed as void*).
 *   - handles are opaque; a handle is used by one caller at a time (thread-compatible),
 *     different handles may be used concurrently.
 *   - Complex<f32> is interleaved {re, im} (num_complex is repr(C)); item counts are in
 *     ITEMS (samples), never bytes, exactly like the Rust slices they replace.
 *   - *_exec calls take DEVICE pointers, are asynchronous and ordered on the context's
 *     stream; the (consumed, produced, status) triple is a pure function of the sizes and is
 *     returned immediately.  *_host calls take HOST pointers and return when the output is
 *     in host memory (they are the literal drop-in for `Filter::filter(&[In], &mut [Out])`).
 */
#ifndef B200SDR_H
#define B200SDR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_VERSION 100 /* 0.1.0 */

/* ---- status codes ---------------------------------------------------------------------- */
#define B2S_OK            0
#define B2S_EINVAL       (-1) /* bad argument (the reference would panic/assert)            */
#define B2S_ECUDA        (-2) /* CUDA runtime error; sticky, see b2s_last_error             */
#define B2S_ENOMEM       (-3)
#define B2S_EAGAIN       (-4) /* ring: no buffer available right now (not an error)         */
#define B2S_EUNSUPPORTED (-5) /* combination the reference supports but this build does not */
#define B2S_ESTATE       (-6) /* slot/ring used out of order                                */
#define B2S_ETIMEOUT     (-7) /* a cross-GPU flag wait gave up (the peer never published)   */

/* futuredsp::ComputationStatus  (crates/futuredsp/src/lib.rs:33-45) */
#define B2S_INSUFFICIENT_INPUT  0
#define B2S_INSUFFICIENT_OUTPUT 1
#define B2S_BOTH_SUFFICIENT     2

/* sample x tap kinds = the Filter impls that exist in futuredsp
 * (fir.rs:206-276, decimating_fir.rs:99-300, polyphase_resampling_fir.rs:126-167) */
typedef enum {
    B2S_F32_F32 = 0, /* f32 samples, f32 taps                 */
    B2S_C32_F32 = 1, /* Complex<f32> samples, f32 taps        */
    B2S_C32_C32 = 2, /* Complex<f32> samples, Complex<f32> taps */
    B2S_F64_F64 = 3  /* f64 samples, f64 taps (fir.rs:217-226; plan with b2s_fir_plan_f64_f64) */
} b2s_kind;

/* FIR algorithm selection (no reference equivalent; AUTO picks by tap count / kind) */
typedef enum {
    B2S_ALGO_AUTO   = 0,
    B2S_ALGO_DIRECT = 1, /* CUDA-core register-blocked direct form (any kind, any decimation) */
    B2S_ALGO_TENSOR = 2, /* tcgen05 block-Toeplitz GEMM, split-bf16 (real taps, 16..257, decim divides 128).
                          * Numerics: operands are split into bf16 hi + lo and three of the four partial products are
                          * summed in FP32: ~2^-18 rms per product, worst case ~3e-5 of ||taps||_1 max|x| when EVERY
                          * product errs the same way (constant taps on constant input -- AUTO keeps constant tap
                          * vectors on DIRECT).  Non-finite input: a NaN/Inf sample at index i makes every output of
                          * the 128-sample blocks whose K-range contains it non-finite (inside [i-K, i+131],
                          * K = 128*ceil((ntaps+127)/128), a superset of the reference's [i-ntaps+1, i]); all other
                          * outputs are unaffected and no finite output is ever wrong.  f32 DENORMAL samples count as
                          * zero (tensor-core operands are flush-to-zero).  Streams that may carry
                          * non-finite samples and need the reference's exact propagation: use B2S_ALGO_DIRECT. */
    B2S_ALGO_FFT    = 3  /* overlap-save FFT convolution (c32 samples, 64..2049 taps, decim == 1)  */
} b2s_algo;

typedef struct b2s_ctx    b2s_ctx;
typedef struct b2s_fir    b2s_fir;
typedef struct b2s_resamp b2s_resamp;
typedef struct b2s_pfbarb b2s_pfbarb;
typedef struct b2s_fft    b2s_fft;
typedef struct b2s_apply  b2s_apply;
typedef struct b2s_ring   b2s_ring;
typedef struct b2s_slot   b2s_slot;

/* ---- context (replaces runtime::buffer::vulkan::Instance, buffer/vulkan/mod.rs:45-153) -- */
int32_t     b2s_version(void);
/* b2s_ctx_create: the context creates (and owns) a non-blocking stream.
 * b2s_ctx_create_on_stream: `stream` is a cudaStream_t owned by the caller (e.g. torch's current
 * stream; 0 / NULL means the legacy default stream) and all work is ordered on it. */
int32_t     b2s_ctx_create(int device, b2s_ctx **out);
int32_t     b2s_ctx_create_on_stream(int device, void *stream, b2s_ctx **out);
void        b2s_ctx_destroy(b2s_ctx *ctx);
const char *b2s_last_error(const b2s_ctx *ctx); /* ctx may be NULL: last global error       */
int32_t     b2s_ctx_sync(b2s_ctx *ctx);         /* ≙ awaiting the fence, blocks/vulkan.rs:157-162 */
void       *b2s_ctx_stream(b2s_ctx *ctx);
int32_t     b2s_ctx_sm_count(b2s_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t    b2s_ctx_launch_count(const b2s_ctx *ctx);

/* device / pinned-host memory (≙ Instance::create_buffer, buffer/vulkan/mod.rs:132) */
int32_t b2s_malloc(b2s_ctx *ctx, size_t bytes, void **dptr);
int32_t b2s_free(b2s_ctx *ctx, void *dptr);
int32_t b2s_host_alloc(b2s_ctx *ctx, size_t bytes, void **hptr); /* pinned */
int32_t b2s_host_free(b2s_ctx *ctx, void *hptr);
int32_t b2s_memcpy_h2d(b2s_ctx *ctx, void *dptr, const void *hptr, size_t bytes); /* async */
int32_t b2s_memcpy_d2h(b2s_ctx *ctx, void *hptr, const void *dptr, size_t bytes); /* async */

/* ---- FIR plans (≙ FirFilter::new fir.rs:39-46, DecimatingFirFilter::new decimating_fir.rs:41-49)
 * taps: ntaps items of f32 (or interleaved Complex<f32> for B2S_C32_C32), in the SAME order the
 * reference takes them (the filter applies them reversed, fir.rs:84); copied at creation.
 * decim == 1 gives FirFilter, decim > 1 DecimatingFirFilter. */
int32_t b2s_fir_plan(b2s_ctx *ctx, b2s_kind kind, const float *taps, size_t ntaps, size_t decim,
                     b2s_fir **out);
int32_t b2s_fir_plan_f32_f32(b2s_ctx *ctx, const float *taps, size_t ntaps, size_t decim, b2s_fir **out);
int32_t b2s_fir_plan_c32_f32(b2s_ctx *ctx, const float *taps, size_t ntaps, size_t decim, b2s_fir **out);
int32_t b2s_fir_plan_c32_c32(b2s_ctx *ctx, const float *taps, size_t ntaps, size_t decim, b2s_fir **out);
/* f64 samples x f64 taps (fir.rs:217-226, decimating_fir.rs:117-130): plain CUDA-core form, un-fused multiply/add in
 * tap order -- bit-identical to the stable-Rust loop; b2s_fir_exec / b2s_fir_exec_hist take f64 items. */
int32_t b2s_fir_plan_f64_f64(b2s_ctx *ctx, const double *taps, size_t ntaps, size_t decim, b2s_fir **out);
void    b2s_fir_destroy(b2s_fir *f);
size_t  b2s_fir_length(const b2s_fir *f);           /* ≙ Filter::length, lib.rs:65-67 */
int32_t b2s_fir_set_algo(b2s_fir *f, b2s_algo algo);
int32_t b2s_fir_get_algo(const b2s_fir *f);         /* the algorithm AUTO resolved to */

/* ≙ Filter::filter (lib.rs:58-64; cores fir.rs:52-91, decimating_fir.rs:53-95): device slices.
 * Elements of d_out beyond *produced are left unspecified, as in the reference. */
int32_t b2s_fir_exec(b2s_fir *f, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                     size_t *consumed, size_t *produced, int32_t *status);
/* Same contract on the LOGICAL slice  d_hist[0..n_hist) ++ d_in[0..n_in):  the history a FIR block keeps in front
 * of new samples (blocks/fir.rs:49 min_items, slab.rs:370-398) handed over as a separate pointer, so that it can live
 * in another ring slot -- or in ANOTHER GPU's memory (the last ntaps-1 samples of the left neighbour's
 * shard; `d_hist` is then a peer address from b2s_ipc_open / b2s_peer_enable).  On the tensor path the kernel's TMA
 * loader fetches the history itself (over NVLink for peer memory): one launch, no copy, no host synchronisation.
 * Other paths copy it into the n_hist items in FRONT of d_in, which must therefore be writable scratch of the same
 * allocation (a ring slot's halo region; b2s_ring_create(..., halo_items >= n_hist, ...)).
 * Optional cross-GPU handshake `hs` (NULL = none; every flag pointer inside may be NULL as well), all system scope,
 * see b2s_flag_*:  publish: *publish_flag = publish_value is stored when the call starts executing -- "my chunk
 * (everything queued on this context before the call) is in HBM", what the right neighbour waits for;  wait: before
 * reading the history the device spins until *wait_flag >= wait_value;  done: *done_flag = done_value is stored
 * once the history has been read, so its owner may overwrite it.  On the tensor path all three happen INSIDE the FIR
 * kernel (no extra launch). */
typedef struct b2s_handshake {
    uint32_t       *publish_flag; uint32_t publish_value;
    const uint32_t *wait_flag;    uint32_t wait_value;
    uint32_t       *done_flag;    uint32_t done_value;
} b2s_handshake;
int32_t b2s_fir_exec_hist(b2s_fir *f, const void *d_hist, size_t n_hist, const void *d_in, size_t n_in,
                          void *d_out, size_t n_out_cap, const b2s_handshake *hs, size_t *consumed, size_t *produced,
                          int32_t *status);
/* Same contract with host slices (pageable or pinned): chunked H2D -> kernel -> D2H pipeline
 * through an internal device ring; returns after the last D2H completed. */
int32_t b2s_fir_filter_host(b2s_fir *f, const void *h_in, size_t n_in, void *h_out, size_t n_out_cap,
                            size_t *consumed, size_t *produced, int32_t *status);

/* ---- rational polyphase resampler (≙ PolyphaseResamplingFir, polyphase_resampling_fir.rs:42-124)
 * kinds B2S_F32_F32 and B2S_C32_F32 (the only impls, :126-167); ntaps % interp == 0 (:56). */
int32_t b2s_resamp_plan(b2s_ctx *ctx, b2s_kind kind, const float *taps, size_t ntaps, size_t interp,
                        size_t decim, b2s_resamp **out);
void    b2s_resamp_destroy(b2s_resamp *r);
size_t  b2s_resamp_length(const b2s_resamp *r);
int32_t b2s_resamp_exec(b2s_resamp *r, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                        size_t *consumed, size_t *produced, int32_t *status);

/* ---- PfbArbResampler (≙ src/blocks/pfb/arb_resampler.rs:90-231; Complex<f32> only).
 * Stateful like the block: one exec == one Kernel::work call (window fill first, :199-215). */
int32_t b2s_pfbarb_plan_c32(b2s_ctx *ctx, const float *taps, size_t ntaps, size_t num_filters,
                            float rate, b2s_pfbarb **out);
void    b2s_pfbarb_destroy(b2s_pfbarb *p);
int32_t b2s_pfbarb_reset(b2s_pfbarb *p);
/* Host-only diagnostic: the timing recurrence (update_timing_state :130-140, `tau -= 1.0` :184-186) is periodic; this
 * returns the period in input items and the outputs it produces (0, 0 when no cycle shorter than 2^25 items starts at
 * tau = 0 -- such plans replay the recurrence on the host per call).  Needs no device. */
int32_t b2s_pfbarb_period(float rate, size_t num_filters, uint64_t *period_items, uint64_t *outputs_per_period);
int32_t b2s_pfbarb_exec(b2s_pfbarb *p, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                        size_t *consumed, size_t *produced, int32_t *call_again);

/* ---- FFT (≙ Fft::with_options, src/blocks/fft.rs:66-121, and Fft::work :160-221) ---------
 * inverse: FftDirection; fft_shift, normalize as in with_options (has_normalize = Option::is_some).
 * One exec processes m = floor(min(n_in, n_out_cap)/n)*n items (no 32-FFT cap: the cap only
 * splits work across calls, fft.rs:171).
 * Lengths: any n >= 2 like rustfft.  Powers of two <= 16384 and other lengths <= 8192 run as ONE pass through shared
 * memory (Stockham / fused Bluestein: 16 B/sample of HBM traffic); larger ones -- powers of two up to 2^26, other
 * lengths up to 2^24 -- use the four-step algorithm through HBM on top of two shared-memory plans (five passes;
 * the plan owns 2 n / 4 M items of scratch).  Changing the length = destroy + create (fft.rs:124-151's handler). */
int32_t b2s_fft_plan_c32(b2s_ctx *ctx, size_t n, int32_t inverse, int32_t fft_shift,
                         int32_t has_normalize, float normalize, b2s_fft **out);
void    b2s_fft_destroy(b2s_fft *f);
size_t  b2s_fft_length(const b2s_fft *f);
int32_t b2s_fft_exec(b2s_fft *f, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                     size_t *consumed, size_t *produced);

/* ---- Apply (≙ src/blocks/apply.rs:100-131).  The reference takes an arbitrary Rust closure;
 * the device version is a closed catalogue of the closures used on the path. */
typedef enum {
    B2S_OP_SCALE_F32     = 0, /* f32 -> f32: x * param          (tests/vulkan.rs:16-27, blocks/wgpu.rs:22-32) */
    B2S_OP_SCALE_C32     = 1, /* c32 -> c32: x * param                                                       */
    B2S_OP_QUAD_DEMOD    = 2, /* c32 -> f32: arg(x[n] * conj(x[n-1])), stateful (examples/fm-receiver/src/main.rs:99-104) */
    B2S_OP_NORM_SQR      = 3, /* c32 -> f32: re^2 + im^2        (examples/spectrum/src/bin/cpu.rs)            */
    B2S_OP_QUAD_DEMOD_C32 = 4, /* c32 -> c32 {re: phase, im: 0}: demod packed for PfbArbResampler */
    B2S_OP_EXP_F32       = 5, /* f32 -> f32: exp(x)            (examples/vulkan/src/main.rs:17-29)            */
    B2S_OP_MAG_C32       = 6, /* c32 -> f32: sqrt(re^2 + im^2)                                               */
    B2S_OP_LOG10_F32     = 7  /* f32 -> f32: param * log10(x)   (spectrum dB stage)                           */
} b2s_op;
int32_t b2s_apply_create(b2s_ctx *ctx, b2s_op op, float param, b2s_apply **out);
void    b2s_apply_destroy(b2s_apply *a);
int32_t b2s_apply_reset(b2s_apply *a); /* closure state back to its initial value */
/* m = min(n_in, n_out_cap) items are processed (apply.rs:109).  Element-wise ops may run in place (d_in == d_out);
 * the stateful demodulators (B2S_OP_QUAD_DEMOD*) need disjoint slices (B2S_EINVAL otherwise) and FINITE input
 * (their atan2 is a finite-input polynomial: inf/inf gives NaN where libm gives +-pi/4, +-3pi/4). */
int32_t b2s_apply_exec(b2s_apply *a, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                       size_t *consumed, size_t *produced);

/* ---- Rotator + XlatingFir helper (≙ futuredsp::Rotator, crates/futuredsp/src/rotator.rs:13-48;
 * XlatingFir = DecimatingFirFilter with band-pass Complex<f32> taps + Rotator at the output rate,
 * src/blocks/xlating_fir.rs:72-126 -- build the FIR with b2s_fir_plan(B2S_C32_C32, bpf, n, decim)).
 * The rotator keeps its phase across calls exactly like the reference object. */
typedef struct b2s_rotator b2s_rotator;
int32_t b2s_rotator_create(b2s_ctx *ctx, float phase_incr, b2s_rotator **out);
void    b2s_rotator_destroy(b2s_rotator *r);
int32_t b2s_rotator_reset(b2s_rotator *r);
/* ≙ Rotator::rotate (:32-47); d_in == d_out is rotate_inplace (:24-29). n = min(n_in, n_out_cap). */
int32_t b2s_rotator_exec(b2s_rotator *r, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                         size_t *processed, int32_t *status);
/* xlating_fir.rs:80-86 and :97-99: band-pass taps (2*ntaps floats) and the rotator's phase increment */
int32_t b2s_xlating_taps(const float *taps, size_t ntaps, float offset, float sample_rate, size_t decimation,
                         float *bpf_interleaved, float *rotator_phase_incr);

/* ---- PfbChannelizer (≙ src/blocks/pfb/channelizer.rs:88-223): polyphase FIR bank +
 * N-point inverse FFT per output vector.  Stateful like the block: one exec == one Kernel::work
 * call (window fill first; note the reference does not consume on the call that completes the fill).
 * d_out is channel-major: output stream ch starts at d_out + ch * out_stride items. */
typedef struct b2s_chan b2s_chan;
int32_t b2s_chan_plan_c32(b2s_ctx *ctx, size_t num_channels, const float *taps, size_t ntaps, float oversample_rate,
                          b2s_chan **out);
void    b2s_chan_destroy(b2s_chan *c);
size_t  b2s_chan_decimation(const b2s_chan *c);
int32_t b2s_chan_exec(b2s_chan *c, const void *d_in, size_t n_in, void *d_out, size_t out_stride, size_t n_out_cap,
                      size_t *consumed, size_t *produced_per_channel, int32_t *call_again);

/* ---- PfbSynthesizer (≙ src/blocks/pfb/synthesizer.rs:52-144): N-point inverse FFT per
 * input vector + polyphase FIR bank, N outputs per vector.  d_in is channel-major (stream w starts at
 * d_in + w * in_stride items), n_in = the shortest input slice.  One exec == one Kernel::work call.
 * Power-of-two banks up to 256 channels with <= 32 taps per arm run as ONE fused kernel in the steady state. */
typedef struct b2s_synth b2s_synth;
int32_t b2s_synth_plan_c32(b2s_ctx *ctx, size_t num_channels, const float *taps, size_t ntaps, b2s_synth **out);
void    b2s_synth_destroy(b2s_synth *s);
int32_t b2s_synth_exec(b2s_synth *s, const void *d_in, size_t in_stride, size_t n_in, void *d_out, size_t n_out_cap,
                       size_t *consumed_per_channel, size_t *produced);

/* ---- MovingAvg<WIDTH> (≙ src/blocks/moving_avg.rs:24-116; SURVEY §8f-3, tail of the spectrum pipe
 * Fft(shift) -> |x|^2 -> MovingAvg).  f32 items; state (avg[WIDTH], chunk counter) kept on the device. */
typedef struct b2s_mavg b2s_mavg;
int32_t b2s_mavg_create(b2s_ctx *ctx, size_t width, float decay_factor, size_t history_size, b2s_mavg **out);
void    b2s_mavg_destroy(b2s_mavg *m);
int32_t b2s_mavg_exec(b2s_mavg *m, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                      size_t *consumed, size_t *produced);

/* ---- fused spectrum pipe: Fft::with_options(n, Forward, fft_shift, None) -> Apply(|x|^2) ->
 * MovingAvg<n>::new(decay_factor, history_size) [-> log10_scale * log10(.) when log10_scale != 0] of
 * examples/spectrum/src/bin/cpu.rs:21-28 in ONE pass over the samples: 8 B/sample in, n floats per `history_size`
 * frames out (the three separate blocks above move 32 B/sample through HBM).  Stateful like MovingAvg (avg[], i).
 * The moving average is evaluated as a blocked scan, so values agree with the sequential f32 reference to rounding
 * (~1e-6 relative), not bit for bit -- use b2s_fft + b2s_apply + b2s_mavg where bit equality matters.
 * n: power of two, 32..8192.  consumed counts Complex<f32> items, produced counts f32 items. */
typedef struct b2s_spectrum b2s_spectrum;
int32_t b2s_spectrum_plan(b2s_ctx *ctx, size_t n, int32_t fft_shift, float decay_factor, size_t history_size,
                          float log10_scale, b2s_spectrum **out);
void    b2s_spectrum_destroy(b2s_spectrum *s);
int32_t b2s_spectrum_reset(b2s_spectrum *s);
int32_t b2s_spectrum_exec(b2s_spectrum *s, const void *d_in, size_t n_in, void *d_out, size_t n_out_cap,
                          size_t *consumed, size_t *produced);

/* ---- device-resident buffer ring (≙ buffer/vulkan/{h2d,d2h}.rs + circuit.rs + slab.rs history)
 * n_slots buffers of `halo_items + chunk_items` items each stay in HBM; ownership of a slot
 * moves source-edge -> GPU block(s) -> sink-edge -> back (circuit), exactly like
 * vulkan::Buffer (h2d.rs:161-232, d2h.rs:66-74, :270-299).  Each slot has pinned host staging
 * for the H2D / D2H edges when with_host_staging != 0. */
int32_t b2s_ring_create(b2s_ctx *ctx, size_t item_bytes, size_t chunk_items, size_t halo_items,
                        int32_t n_slots, int32_t with_host_staging, b2s_ring **out);
void    b2s_ring_destroy(b2s_ring *r);
/* ≙ H2DWriter: pop an empty buffer from `inbound` (h2d.rs:178-197); B2S_EAGAIN if none */
int32_t b2s_ring_acquire_empty(b2s_ring *r, b2s_slot **slot);
/* ≙ H2DWriter::produce -> outbound.push + notify (h2d.rs:199-232).  from_host != 0 first
 * enqueues the pinned->device copy of valid_items. */
int32_t b2s_ring_submit_full(b2s_ring *r, b2s_slot *slot, size_t valid_items, int32_t from_host);
/* ≙ H2DReader::buffers() / D2HReader (h2d.rs:276, d2h.rs:247-268); B2S_EAGAIN if none */
int32_t b2s_ring_acquire_full(b2s_ring *r, b2s_slot **slot, size_t *valid_items);
/* ≙ D2HReader::consume -> buffer back to the circuit start (d2h.rs:270-299) */
int32_t b2s_ring_release(b2s_ring *r, b2s_slot *slot);
/* slab.rs:370-398: copy the unconsumed tail (`tail_items` <= halo_items) of `from` in front of
 * `to`'s data so the next block sees contiguous history. */
int32_t b2s_ring_carry_halo(b2s_ring *r, const b2s_slot *from, size_t from_valid, size_t tail_items,
                            b2s_slot *to);
void   *b2s_slot_device_ptr(const b2s_slot *slot); /* first data item; halo lives just below   */
void   *b2s_slot_host_ptr(const b2s_slot *slot);   /* pinned staging (NULL without staging)    */
size_t  b2s_slot_halo_valid(const b2s_slot *slot); /* items of history currently in front      */
int32_t b2s_slot_fetch_to_host(b2s_slot *slot, size_t items); /* async D2H into staging + event */
int32_t b2s_slot_wait(b2s_slot *slot);             /* host wait on the slot's event            */
size_t  b2s_ring_free_slots(const b2s_ring *r);
size_t  b2s_ring_full_slots(const b2s_ring *r);
/* geometry of the ring's single device allocation, for exporting it to a peer process (b2s_ipc_export):
 * a peer addresses slot i's first data item at  peer_base + b2s_ring_slot_offset(i)  and the ring's two u32
 * counters {ready, consumed} at  peer_base + b2s_ring_flags_offset(). */
void   *b2s_ring_base(const b2s_ring *r);
size_t  b2s_ring_bytes(const b2s_ring *r);
size_t  b2s_ring_slot_offset(const b2s_ring *r, int32_t slot_index);
size_t  b2s_ring_flags_offset(const b2s_ring *r);
int32_t b2s_slot_index(const b2s_slot *slot);

/* ---- cross-GPU plumbing of a sharded stream (no reference equivalent -- FutureSDR is single-device).
 * One process per GPU: export a device allocation (d_base = pointer returned by b2s_malloc or b2s_ring_base) as an
 * opaque 64-byte handle, ship the handle to the neighbour by any means (torch.distributed, a pipe), and map it
 * there.  One process driving several GPUs uses b2s_peer_enable instead and passes raw pointers. */
#define B2S_IPC_HANDLE_BYTES 64
int32_t b2s_ipc_export(b2s_ctx *ctx, void *d_base, uint8_t handle[B2S_IPC_HANDLE_BYTES]);
int32_t b2s_ipc_open(b2s_ctx *ctx, const uint8_t handle[B2S_IPC_HANDLE_BYTES], void **d_peer);
int32_t b2s_ipc_close(b2s_ctx *ctx, void *d_peer);
int32_t b2s_peer_enable(b2s_ctx *ctx, int32_t peer_device);
/* Stream-ordered system-scope flags (u32 counters in device memory, local or peer):
 * set = release store after everything queued before it; wait = the stream stalls until *flag >= value (wrap-safe),
 * giving up after 4 s (b2s_ctx_sync then returns B2S_ETIMEOUT).  b2s_flag_read is a blocking host read. */
int32_t b2s_flag_set(b2s_ctx *ctx, uint32_t *d_flag, uint32_t value);
int32_t b2s_flag_wait(b2s_ctx *ctx, const uint32_t *d_flag, uint32_t value);
int32_t b2s_flag_read(b2s_ctx *ctx, const uint32_t *d_flag, uint32_t *value);
int32_t b2s_memcpy_d2d(b2s_ctx *ctx, void *dst, const void *src, size_t bytes); /* async, any two devices */
int32_t b2s_memset(b2s_ctx *ctx, void *dst, int32_t byte, size_t bytes);         /* async */

/* ---- tap design, host side, f64 then cast (≙ futuredsp::firdes::kaiser, firdes/basic.rs:310-459)
 * Return the tap count; write taps only if cap is large enough (call with taps=NULL to size). */
size_t b2s_firdes_kaiser_lowpass(double cutoff, double transition_bw, double max_ripple,
                                 float *taps, size_t cap);
size_t b2s_firdes_kaiser_multirate(size_t interp, size_t decim, size_t half_polyphase_len,
                                   double max_ripple, float *taps, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* B200SDR_H */
