// b200sdr.hpp -- C++ host layer above the C ABI (include/b200sdr.h), mirroring the reference's
// Rust interfaces for the hot path so call sites and tests read like the reference's own:
//
//   futuredsp::Filter::filter(&self, &[In], &mut [Out]) -> (usize, usize, ComputationStatus)
//                                     (crates/futuredsp/src/lib.rs:48-68)
//   futuredsp::{FirFilter, DecimatingFirFilter, PolyphaseResamplingFir}
//   futuresdr::blocks::{Fir, FirBuilder, Fft, Apply, PfbArbResampler}   (src/blocks/*.rs)
//   futuresdr::runtime::{WorkIo, mocker::Mocker}                        (work_io.rs, mocker.rs)
//
// The reference is Rust; no Rust toolchain exists in this image, so this header is the
// compiled-language host side (INTEGRATION.md carries the Rust shim source).  Header-only,
// C++17, links against libb200sdr.so.  Errors the reference panics/asserts on throw b2s::Error.
#pragma once

#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#include "b200sdr.h"

namespace b2s {

using Complex32 = std::complex<float>;

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &m) : std::runtime_error(m), code(c) {}
};

inline void check(int32_t rc, const b2s_ctx *ctx = nullptr) {
    if (rc < 0) throw Error(rc, std::string("libb200sdr: ") + b2s_last_error(ctx));
}

// futuredsp::ComputationStatus (lib.rs:33-45)
enum class ComputationStatus : int32_t { InsufficientInput = 0, InsufficientOutput = 1, BothSufficient = 2 };
using FilterResult = std::tuple<size_t, size_t, ComputationStatus>;

// ≙ runtime::buffer::vulkan::Instance (buffer/vulkan/mod.rs:45-153)
class Instance {
public:
    explicit Instance(int device = 0) { check(b2s_ctx_create(device, &ctx_)); }
    ~Instance() { b2s_ctx_destroy(ctx_); }
    Instance(const Instance &) = delete;
    Instance &operator=(const Instance &) = delete;
    b2s_ctx *get() const { return ctx_; }
    void sync() const { check(b2s_ctx_sync(ctx_), ctx_); }
    uint64_t launch_count() const { return b2s_ctx_launch_count(ctx_); }

    template <typename T> T *device_alloc(size_t items) const {
        void *p = nullptr;
        check(b2s_malloc(ctx_, items * sizeof(T), &p), ctx_);
        return static_cast<T *>(p);
    }
    void device_free(void *p) const { b2s_free(ctx_, p); }
    template <typename T> void upload(T *dst, const T *src, size_t items) const {
        check(b2s_memcpy_h2d(ctx_, dst, src, items * sizeof(T)), ctx_);
    }
    template <typename T> void download(T *dst, const T *src, size_t items) const {
        check(b2s_memcpy_d2h(ctx_, dst, src, items * sizeof(T)), ctx_);
        sync();
    }

private:
    b2s_ctx *ctx_ = nullptr;
};

template <typename Sample, typename Tap> constexpr b2s_kind kind_of() {
    if constexpr (std::is_same_v<Sample, float> && std::is_same_v<Tap, float>) return B2S_F32_F32;
    else if constexpr (std::is_same_v<Sample, Complex32> && std::is_same_v<Tap, float>) return B2S_C32_F32;
    else {
        static_assert(std::is_same_v<Sample, Complex32> && std::is_same_v<Tap, Complex32>,
                      "no futuredsp impl for this sample/tap combination");
        return B2S_C32_C32;
    }
}

// ---- futuredsp::Filter ----------------------------------------------------------------------
template <typename Sample> class Filter {
public:
    virtual ~Filter() = default;
    // host slices (Filter::filter(&[In], &mut [Out]))
    virtual FilterResult filter(const Sample *input, size_t n_in, Sample *output, size_t n_out) const = 0;
    // device slices (samples already in HBM; asynchronous on the instance's stream)
    virtual FilterResult filter_device(const Sample *d_in, size_t n_in, Sample *d_out, size_t n_out) const = 0;
    virtual size_t length() const = 0;
    FilterResult filter(const std::vector<Sample> &i, std::vector<Sample> &o) const {
        return filter(i.data(), i.size(), o.data(), o.size());
    }
};

// ≙ DecimatingFirFilter (decimating_fir.rs:31-95); decimation 1 == FirFilter (fir.rs:31-91)
template <typename Sample, typename Tap> class DecimatingFirFilter : public Filter<Sample> {
public:
    DecimatingFirFilter(const Instance &inst, size_t decimation, const std::vector<Tap> &taps, b2s_algo algo = B2S_ALGO_AUTO)
        : inst_(inst) {
        check(b2s_fir_plan(inst.get(), kind_of<Sample, Tap>(), reinterpret_cast<const float *>(taps.data()), taps.size(),
                           decimation, &plan_), inst.get());
        if (algo != B2S_ALGO_AUTO) check(b2s_fir_set_algo(plan_, algo), inst.get());
    }
    ~DecimatingFirFilter() override { b2s_fir_destroy(plan_); }
    FilterResult filter(const Sample *i, size_t n_in, Sample *o, size_t n_out) const override {
        size_t c = 0, p = 0; int32_t st = 0;
        check(b2s_fir_filter_host(plan_, i, n_in, o, n_out, &c, &p, &st), inst_.get());
        return {c, p, static_cast<ComputationStatus>(st)};
    }
    FilterResult filter_device(const Sample *i, size_t n_in, Sample *o, size_t n_out) const override {
        size_t c = 0, p = 0; int32_t st = 0;
        check(b2s_fir_exec(plan_, i, n_in, o, n_out, &c, &p, &st), inst_.get());
        return {c, p, static_cast<ComputationStatus>(st)};
    }
    using Filter<Sample>::filter;
    size_t length() const override { return b2s_fir_length(plan_); }
    int algo() const { return b2s_fir_get_algo(plan_); }

protected:
    const Instance &inst_;
    b2s_fir *plan_ = nullptr;
};

template <typename Sample, typename Tap> class FirFilter : public DecimatingFirFilter<Sample, Tap> {
public:
    FirFilter(const Instance &inst, const std::vector<Tap> &taps, b2s_algo algo = B2S_ALGO_AUTO)
        : DecimatingFirFilter<Sample, Tap>(inst, 1, taps, algo) {}
};

// ≙ PolyphaseResamplingFir (polyphase_resampling_fir.rs:42-124); device slices only
template <typename Sample> class PolyphaseResamplingFir : public Filter<Sample> {
public:
    PolyphaseResamplingFir(const Instance &inst, size_t interp, size_t decim, const std::vector<float> &taps)
        : inst_(inst) {
        check(b2s_resamp_plan(inst.get(), kind_of<Sample, float>(), taps.data(), taps.size(), interp, decim, &plan_),
              inst.get());
    }
    ~PolyphaseResamplingFir() override { b2s_resamp_destroy(plan_); }
    FilterResult filter(const Sample *i, size_t n_in, Sample *o, size_t n_out) const override {
        // host slices: stage through device memory (no internal pipeline for this core yet)
        Sample *di = inst_.device_alloc<Sample>(n_in + 1), *dout = inst_.device_alloc<Sample>(n_out + 1);
        inst_.upload(di, i, n_in);
        auto r = filter_device(di, n_in, dout, n_out);
        inst_.download(o, dout, std::get<1>(r));
        inst_.device_free(di); inst_.device_free(dout);
        return r;
    }
    FilterResult filter_device(const Sample *i, size_t n_in, Sample *o, size_t n_out) const override {
        size_t c = 0, p = 0; int32_t st = 0;
        check(b2s_resamp_exec(plan_, i, n_in, o, n_out, &c, &p, &st), inst_.get());
        return {c, p, static_cast<ComputationStatus>(st)};
    }
    using Filter<Sample>::filter;
    size_t length() const override { return b2s_resamp_length(plan_); }

private:
    const Instance &inst_;
    b2s_resamp *plan_ = nullptr;
};

// ---- firdes (futuredsp::firdes::kaiser, firdes/basic.rs:310-459) ---------------------------------
namespace firdes::kaiser {
inline std::vector<float> lowpass(double cutoff, double transition_bw, double max_ripple) {
    std::vector<float> t(b2s_firdes_kaiser_lowpass(cutoff, transition_bw, max_ripple, nullptr, 0));
    if (t.empty()) throw Error(B2S_EINVAL, "firdes::kaiser::lowpass: bad specification");
    b2s_firdes_kaiser_lowpass(cutoff, transition_bw, max_ripple, t.data(), t.size());
    return t;
}
inline std::vector<float> multirate(size_t interp, size_t decim, size_t half_len, double max_ripple) {
    std::vector<float> t(b2s_firdes_kaiser_multirate(interp, decim, half_len, max_ripple, nullptr, 0));
    if (t.empty()) throw Error(B2S_EINVAL, "firdes::kaiser::multirate: bad specification");
    b2s_firdes_kaiser_multirate(interp, decim, half_len, max_ripple, t.data(), t.size());
    return t;
}
}  // namespace firdes::kaiser

// ---- runtime pieces the blocks need -----------------------------------------------------------
struct WorkIo { bool call_again = false, finished = false; };   // work_io.rs:11-34

// mocker::Reader / mocker::Writer (mocker.rs:213-400) over device memory
template <typename T> class Reader {
public:
    explicit Reader(const Instance &i) : inst_(i) {}
    ~Reader() { if (d_) inst_.device_free(d_); }
    void set(const std::vector<T> &v) {
        if (d_) inst_.device_free(d_);
        d_ = inst_.device_alloc<T>(v.size() + 1); n_ = v.size(); pos_ = 0;
        inst_.upload(d_, v.data(), v.size());
    }
    const T *slice() const { return d_ + pos_; }
    size_t len() const { return n_ - pos_; }
    void consume(size_t n) { pos_ += n; }
    bool finished() const { return true; }
private:
    const Instance &inst_; T *d_ = nullptr; size_t n_ = 0, pos_ = 0;
};
template <typename T> class Writer {
public:
    explicit Writer(const Instance &i) : inst_(i) {}
    ~Writer() { if (d_) inst_.device_free(d_); }
    void reserve(size_t n) { if (d_) inst_.device_free(d_); d_ = inst_.device_alloc<T>(n + 1); cap_ = n; len_ = 0; }
    T *slice() { return d_ + len_; }
    size_t capacity() const { return cap_ - len_; }
    void produce(size_t n) { len_ += n; }
    std::vector<T> get() const { std::vector<T> v(len_); if (len_) inst_.download(v.data(), d_, len_); return v; }
private:
    const Instance &inst_; T *d_ = nullptr; size_t cap_ = 0, len_ = 0;
};

// ≙ blocks::Fir (src/blocks/fir.rs:13-95)
template <typename Sample> class Fir {
public:
    Fir(const Instance &inst, std::unique_ptr<Filter<Sample>> core) : input(inst), output(inst), filter_(std::move(core)) {}
    size_t n_taps() const { return filter_->length(); }
    void work(WorkIo &io) {                                                        // fir.rs:75-94
        auto [consumed, produced, status] = filter_->filter_device(input.slice(), input.len(), output.slice(), output.capacity());
        input.consume(consumed);
        output.produce(produced);
        if (input.finished() && status != ComputationStatus::InsufficientOutput) io.finished = true;
    }
    Reader<Sample> input;
    Writer<Sample> output;
private:
    std::unique_ptr<Filter<Sample>> filter_;
};

// ≙ blocks::FirBuilder (src/blocks/fir.rs:126-233)
struct FirBuilder {
    template <typename Sample, typename Tap>
    static Fir<Sample> fir(const Instance &i, const std::vector<Tap> &taps) {
        return Fir<Sample>(i, std::make_unique<FirFilter<Sample, Tap>>(i, taps));
    }
    template <typename Sample> static Fir<Sample> decimating(const Instance &i, size_t decim) {
        return decimating_with_taps<Sample, float>(i, decim, firdes::kaiser::lowpass(1.0 / (double)decim, 0.1, 0.0001));   // fir.rs:154
    }
    template <typename Sample, typename Tap>
    static Fir<Sample> decimating_with_taps(const Instance &i, size_t decim, const std::vector<Tap> &taps) {
        return Fir<Sample>(i, std::make_unique<DecimatingFirFilter<Sample, Tap>>(i, decim, taps));
    }
    template <typename Sample> static Fir<Sample> resampling(const Instance &i, size_t interp, size_t decim) {
        size_t a = interp, b = decim;
        while (b) { size_t t = a % b; a = b; b = t; }                               // gcd (fir.rs:197-199)
        interp /= a; decim /= a;
        return resampling_with_taps<Sample>(i, interp, decim, firdes::kaiser::multirate(interp, decim, 12, 0.0001));
    }
    template <typename Sample>
    static Fir<Sample> resampling_with_taps(const Instance &i, size_t interp, size_t decim, const std::vector<float> &taps) {
        if (taps.size() % interp) throw Error(B2S_EINVAL, "taps.num_taps().is_multiple_of(interp)");   // :56
        return Fir<Sample>(i, std::make_unique<PolyphaseResamplingFir<Sample>>(i, interp, decim, taps));
    }
};

// ≙ blocks::Fft (src/blocks/fft.rs:30-221)
enum class FftDirection { Forward, Inverse };
class Fft {
public:
    Fft(const Instance &inst, size_t len, FftDirection dir = FftDirection::Forward, bool fft_shift = false,
        bool has_normalize = false, float normalize = 1.0f)
        : input(inst), output(inst), inst_(inst), len_(len) {
        check(b2s_fft_plan_c32(inst.get(), len, dir == FftDirection::Inverse, fft_shift, has_normalize, normalize, &plan_), inst.get());
    }
    ~Fft() { b2s_fft_destroy(plan_); }
    void work(WorkIo &io) {                                                        // fft.rs:160-221
        size_t c = 0, p = 0;
        check(b2s_fft_exec(plan_, input.slice(), input.len(), output.slice(), output.capacity(), &c, &p), inst_.get());
        input.consume(c); output.produce(p);
        if (input.finished() && c == (c / len_) * len_) io.finished = true;
    }
    Reader<Complex32> input;
    Writer<Complex32> output;
private:
    const Instance &inst_; size_t len_; b2s_fft *plan_ = nullptr;
};

// ≙ blocks::Apply (src/blocks/apply.rs:100-131) for the device op catalogue
template <typename A, typename B> class Apply {
public:
    Apply(const Instance &inst, b2s_op op, float param = 1.0f) : input(inst), output(inst), inst_(inst) {
        check(b2s_apply_create(inst.get(), op, param, &h_), inst.get());
    }
    ~Apply() { b2s_apply_destroy(h_); }
    void work(WorkIo &io) {
        const size_t i_len = input.len();
        size_t c = 0, p = 0;
        check(b2s_apply_exec(h_, input.slice(), i_len, output.slice(), output.capacity(), &c, &p), inst_.get());
        input.consume(c); output.produce(p);
        if (input.finished() && c == i_len) io.finished = true;                     // apply.rs:126-128
    }
    Reader<A> input;
    Writer<B> output;
private:
    const Instance &inst_; b2s_apply *h_ = nullptr;
};

// ≙ blocks::PfbArbResampler (src/blocks/pfb/arb_resampler.rs:72-231)
class PfbArbResampler {
public:
    PfbArbResampler(const Instance &inst, float rate, const std::vector<float> &taps, size_t num_filters)
        : input(inst), output(inst), inst_(inst) {
        check(b2s_pfbarb_plan_c32(inst.get(), taps.data(), taps.size(), num_filters, rate, &h_), inst.get());
    }
    ~PfbArbResampler() { b2s_pfbarb_destroy(h_); }
    void work(WorkIo &io) {
        size_t c = 0, p = 0; int32_t again = 0;
        const size_t n = input.len();
        check(b2s_pfbarb_exec(h_, input.slice(), n, output.slice(), output.capacity(), &c, &p, &again), inst_.get());
        input.consume(c); output.produce(p);
        if (again) io.call_again = true;
        else if (n - c == 0 && input.finished()) io.finished = true;
    }
    Reader<Complex32> input;
    Writer<Complex32> output;
private:
    const Instance &inst_; b2s_pfbarb *h_ = nullptr;
};

// ≙ futuredsp::Rotator (crates/futuredsp/src/rotator.rs:13-48): phase recurrence replayed bit for bit
class Rotator {
public:
    Rotator(const Instance &inst, float phase_incr) : inst_(inst) { check(b2s_rotator_create(inst.get(), phase_incr, &h_), inst.get()); }
    ~Rotator() { b2s_rotator_destroy(h_); }
    Rotator(const Rotator &) = delete;
    // Rotator::rotate (:32-47) on device slices; d_in == d_out is rotate_inplace (:24-29)
    std::pair<size_t, ComputationStatus> rotate_device(const Complex32 *d_in, size_t n_in, Complex32 *d_out, size_t n_out) {
        size_t n = 0; int32_t st = 0;
        check(b2s_rotator_exec(h_, d_in, n_in, d_out, n_out, &n, &st), inst_.get());
        return {n, static_cast<ComputationStatus>(st)};
    }
    void reset() { check(b2s_rotator_reset(h_), inst_.get()); }
private:
    const Instance &inst_; b2s_rotator *h_ = nullptr;
};

// ≙ blocks::XlatingFir (src/blocks/xlating_fir.rs:22-126): band-pass complex taps (:80-86), decimating FIR,
// Rotator at the output rate (:97-99, :118)
class XlatingFir {
public:
    XlatingFir(const Instance &inst, size_t decimation, float offset, float sample_rate)          // XlatingFir::new (:42-48)
        : XlatingFir(inst, default_taps(decimation), decimation, offset, sample_rate) {}
    XlatingFir(const Instance &inst, const std::vector<float> &taps, size_t decimation, float offset, float sample_rate)
        : input(inst), output(inst), inst_(inst) {
        if (decimation == 0) throw Error(B2S_EINVAL, "Xlating FIR: decimation must be > 0");
        std::vector<Complex32> bpf(taps.size());
        float incr = 0.f;
        check(b2s_xlating_taps(taps.data(), taps.size(), offset, sample_rate, decimation,
                               reinterpret_cast<float *>(bpf.data()), &incr));
        filter_ = std::make_unique<DecimatingFirFilter<Complex32, Complex32>>(inst, decimation, bpf);
        rotator_ = std::make_unique<Rotator>(inst, incr);
    }
    size_t n_taps() const { return filter_->length(); }
    void work(WorkIo &io) {                                                        // xlating_fir.rs:105-126
        auto [consumed, produced, status] = filter_->filter_device(input.slice(), input.len(), output.slice(), output.capacity());
        if (produced) rotator_->rotate_device(output.slice(), produced, output.slice(), produced);
        input.consume(consumed);
        output.produce(produced);
        if (input.finished() && status != ComputationStatus::InsufficientOutput) io.finished = true;
    }
    Reader<Complex32> input;
    Writer<Complex32> output;
private:
    static std::vector<float> default_taps(size_t decimation) {
        if (decimation < 2) throw Error(B2S_EINVAL, "Xlating FIR: Decimation has to be >= 2");   // :43
        const double transition_bw = 0.1;
        const double cutoff = std::min(0.5 - transition_bw - 2.220446049250313e-16, 1.0 / (double)decimation);
        return firdes::kaiser::lowpass(cutoff, transition_bw, 0.0001);
    }
    const Instance &inst_;
    std::unique_ptr<DecimatingFirFilter<Complex32, Complex32>> filter_;
    std::unique_ptr<Rotator> rotator_;
};

// ≙ blocks::MovingAvg<WIDTH> (src/blocks/moving_avg.rs:24-116)
class MovingAvg {
public:
    MovingAvg(const Instance &inst, size_t width, float decay_factor, size_t history_size)
        : input(inst), output(inst), inst_(inst), width_(width) {
        check(b2s_mavg_create(inst.get(), width, decay_factor, history_size, &h_), inst.get());   // asserts of :58-61 -> EINVAL
    }
    ~MovingAvg() { b2s_mavg_destroy(h_); }
    void work(WorkIo &io) {                                                        // moving_avg.rs:72-115
        const size_t n = input.len();
        size_t c = 0, p = 0;
        check(b2s_mavg_exec(h_, input.slice(), n, output.slice(), output.capacity(), &c, &p), inst_.get());
        if (input.finished() && c / width_ == n / width_) io.finished = true;       // :106-108
        input.consume(c); output.produce(p);
    }
    Reader<float> input;
    Writer<float> output;
private:
    const Instance &inst_; size_t width_; b2s_mavg *h_ = nullptr;
};

// The spectrum flowgraph's compute chain as one block: Fft::with_options(n, Forward, fft_shift, None) ->
// Apply(norm_sqr) -> MovingAvg<n>::new(decay, history) of examples/spectrum/src/bin/cpu.rs:21-28 in ONE pass over the
// samples (b2s_spectrum_*); counts follow MovingAvg::work, values agree with the three blocks to rounding.
class SpectrumPipe {
public:
    SpectrumPipe(const Instance &inst, size_t n, float decay_factor, size_t history_size, bool fft_shift = true,
                 float log10_scale = 0.0f)
        : input(inst), output(inst), inst_(inst), n_(n) {
        check(b2s_spectrum_plan(inst.get(), n, fft_shift ? 1 : 0, decay_factor, history_size, log10_scale, &h_), inst.get());
    }
    ~SpectrumPipe() { b2s_spectrum_destroy(h_); }
    void work(WorkIo &io) {
        const size_t n = input.len();
        size_t c = 0, p = 0;
        check(b2s_spectrum_exec(h_, input.slice(), n, output.slice(), output.capacity(), &c, &p), inst_.get());
        if (input.finished() && c / n_ == n / n_) io.finished = true;               // moving_avg.rs:106-108
        input.consume(c); output.produce(p);
    }
    Reader<Complex32> input;
    Writer<float> output;
private:
    const Instance &inst_; size_t n_; b2s_spectrum *h_ = nullptr;
};

// ≙ runtime::mocker::Mocker (mocker.rs:33-190): run one block without a scheduler
template <typename Block> class Mocker {
public:
    explicit Mocker(Block &b) : b_(b) {}
    template <typename T> void input(const std::vector<T> &v) { b_.input.set(v); }
    void init_output(size_t n) { b_.output.reserve(n); }
    WorkIo run() {
        WorkIo io;
        for (int guard = 0; guard < (1 << 20); guard++) {
            io = WorkIo{};
            b_.work(io);
            if (io.finished || !io.call_again) break;
        }
        return io;
    }
    auto output() { return b_.output.get(); }
private:
    Block &b_;
};

}  // namespace b2s
