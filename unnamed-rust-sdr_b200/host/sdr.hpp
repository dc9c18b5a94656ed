// sdr.hpp -- C++ host-side mirror of the reference crate's public API for the hot path, sitting on
// the C ABI of include/sdr_b200.h.  (The reference is Rust; this image has no Rust toolchain, so the
// host side above the boundary is written in C++ -- same names, argument meaning and error behaviour.
// The Rust binding a maintainer would add is in ../rust/ and INTEGRATION.md.)
//
//   reference                                          here
//   ------------------------------------------------   -------------------------------------------
//   num::Complex<f32>                                  sdr::Complex
//   trait Signal {next, rate, + combinators}           sdr::signal::Signal<Sample> (block pull) with the same
//     (src/signal/mod.rs:13-123)                         combinators: block decimate filter map resample
//                                                        resample_with skip take iter
//   signal::from_iter (sources.rs:31-36)               sdr::signal::from_iter(rate, vector)
//   rtltcp::RtlTcpSignal (rtltcp.rs:151-168)           sdr::signal::from_u8iq(rate, bytes)
//   filter::Fir<C,A>, FilterDesign for Vec<C>          sdr::filter::Fir<C,A>  (+ block process(), mirroring
//     (src/filter/fir.rs)                                SampleRate::process, since a per-sample GPU call is absurd)
//   filter::BiquadD, Biquad, Identity, PllDesign, Pll  sdr::filter::BiquadD, Biquad<A>, Identity, PllDesign, Pll
//   resample::SampleRate<A>, ConverterType, Error      sdr::resample::SampleRate<A>, ConverterType, Error
//     (src/resample.rs)
//   fft::fft, fft::rfft (src/fft.rs)                   sdr::fft::fft, sdr::fft::rfft
//
// Errors: the reference returns Result and its adaptors unwrap() (panic); here constructors / process
// throw sdr::Error / sdr::resample::Error, which is the same "abort the pipeline" behaviour.
#pragma once
#include <complex>
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/sdr_b200.h"

namespace sdr {

using Complex = std::complex<float>;

struct Error : std::runtime_error {
    int code;
    Error(int c, const char *what) : std::runtime_error(std::string(what) + ": " + sdr_strerror(c)), code(c) {}
};
inline void check(int rc, const char *what) {
    if (rc != SDR_OK) throw Error(rc, what);
}

template <class A> struct SampleTraits;
template <> struct SampleTraits<float> { static constexpr int fmt = SDR_FMT_F32; static constexpr int channels = 1; };
template <> struct SampleTraits<Complex> { static constexpr int fmt = SDR_FMT_C64; static constexpr int channels = 2; };
template <class A, class B> struct SampleTraits<std::pair<A, B>> {  // resample.rs:280-282
    static constexpr int channels = SampleTraits<A>::channels + SampleTraits<B>::channels;
};

// =========================================================================================
namespace resample {

enum class ConverterType {  // resample.rs:112-119
    SincBestQuality = SDR_SRC_SINC_BEST_QUALITY,
    SincMediumQuality = SDR_SRC_SINC_MEDIUM_QUALITY,
    SincFastest = SDR_SRC_SINC_FASTEST,
    ZeroOrderHold = SDR_SRC_ZERO_ORDER_HOLD,
    Linear = SDR_SRC_LINEAR,
};
inline const char *name(ConverterType t) { return sdr_src_get_name((int)t); }
inline const char *description(ConverterType t) { return sdr_src_get_description((int)t); }
inline const char *version() { return sdr_src_get_version(); }

struct Error : std::runtime_error {  // resample.rs:151-270
    int code;
    explicit Error(int c) : std::runtime_error(sdr_src_strerror(c) ? sdr_src_strerror(c) : "Unknown"), code(c) {}
};

// SampleRate<A>: A is memory-identical to [f32; channels] (resample.rs:27-30)
template <class A>
class SampleRate {
    SDR_SRC_STATE *state_ = nullptr;
    explicit SampleRate(SDR_SRC_STATE *s) : state_(s) {}

  public:
    explicit SampleRate(ConverterType typ) {  // SampleRate::new (resample.rs:33-44)
        int err = 0;
        state_ = sdr_src_new((int)typ, SampleTraits<A>::channels, &err);
        if (!state_ || err) throw Error(err);
    }
    SampleRate(SampleRate &&o) noexcept : state_(o.state_) { o.state_ = nullptr; }
    SampleRate &operator=(SampleRate &&o) noexcept { std::swap(state_, o.state_); return *this; }
    SampleRate(const SampleRate &o) : state_(nullptr) { *this = o.try_clone(); }  // impl Clone (resample.rs:17-21)
    ~SampleRate() { if (state_) state_ = sdr_src_delete(state_); }              // Drop (resample.rs:101-110)

    // process(ratio, &input, &mut output) -> input_frames_used; output.len = output_frames_gen;
    // output_frames = output.capacity(); end_of_input = input.is_empty()   (resample.rs:46-67)
    size_t process(double ratio, const std::vector<A> &input, std::vector<A> &output) {
        const size_t cap = output.capacity();
        output.resize(cap);
        SDR_SRC_DATA cmd;
        cmd.data_in = reinterpret_cast<const float *>(input.data());
        cmd.data_out = reinterpret_cast<float *>(output.data());
        cmd.input_frames = (long)input.size();
        cmd.output_frames = (long)cap;
        cmd.input_frames_used = 0;
        cmd.output_frames_gen = 0;
        cmd.end_of_input = input.empty() ? 1 : 0;
        cmd.src_ratio = ratio;
        const int rc = sdr_src_process(state_, &cmd);
        if (rc) { output.resize(0); throw Error(rc); }
        output.resize((size_t)cmd.output_frames_gen);
        return (size_t)cmd.input_frames_used;
    }
    void reset() { int rc = sdr_src_reset(state_); if (rc) throw Error(rc); }
    SampleRate try_clone() const {
        int err = 0;
        SDR_SRC_STATE *s = sdr_src_clone(state_, &err);
        if (!s || err) throw Error(err);
        return SampleRate(s);
    }
    size_t channels() const { return (size_t)sdr_src_get_channels(state_); }
    void set_ratio(double r) { int rc = sdr_src_set_ratio(state_, r); if (rc) throw Error(rc); }
};

}  // namespace resample

// =========================================================================================
namespace filter {

// Fir<C,A> (src/filter/fir.rs:7-33).  C in {float, Complex}; A in {float, Complex}; also usable with
// input_u8iq = true, where the samples arrive as rtl_tcp bytes and are unpacked on the GPU.
template <class C, class A>
class Fir {
    sdr_fir_t *h_ = nullptr;
    std::vector<C> coef_;
    explicit Fir(sdr_fir_t *h, std::vector<C> c) : h_(h), coef_(std::move(c)) {}

  public:
    explicit Fir(std::vector<C> coef, size_t decimation = 1, bool input_u8iq = false, unsigned flags = 0)
        : coef_(std::move(coef)) {
        static_assert(!(std::is_same<C, Complex>::value && std::is_same<A, float>::value),
                      "f32 * Complex<f32> is not a Convolve impl in the reference");
        sdr_fir_config_t cfg{};
        cfg.taps = reinterpret_cast<const float *>(coef_.data());
        cfg.n_taps = coef_.size();
        cfg.taps_complex = std::is_same<C, Complex>::value;
        cfg.input_format = input_u8iq ? SDR_FMT_U8IQ : SampleTraits<A>::fmt;
        cfg.decimation = decimation;
        cfg.n_channels = 1;
        cfg.flags = flags;
        cfg.device = 0;
        cfg.stream = nullptr;
        int err = 0;
        h_ = sdr_fir_create(&cfg, &err);
        if (!h_) throw sdr::Error(err, "Fir::new");
    }
    Fir(Fir &&o) noexcept : h_(o.h_), coef_(std::move(o.coef_)) { o.h_ = nullptr; }
    Fir(const Fir &o) : coef_(o.coef_) {  // #[derive(Clone)] (fir.rs:6)
        int err = 0;
        h_ = sdr_fir_clone(o.h_, &err);
        if (!h_) throw sdr::Error(err, "Fir::clone");
    }
    ~Fir() { sdr_fir_destroy(h_); }

    // block form of Filter::apply: out gets one element per kept input (all of them when decimation == 1)
    void process(const void *in, size_t n_in, std::vector<A> &out) {
        const size_t n_out = sdr_fir_output_count(h_, n_in);
        out.resize(n_out);
        size_t used = 0, got = 0;
        check(sdr_fir_process(h_, in, n_in, n_in, out.data(), n_out, n_out, &used, &got), "Fir::process");
        out.resize(got);
    }
    A apply(A value) {  // Filter::apply (filter/mod.rs:23-26); one launch per sample: correct, not fast
        std::vector<A> o;
        process(&value, 1, o);
        return o.empty() ? A() : o[0];
    }
    void reset() { check(sdr_fir_reset(h_), "Fir::reset"); }
    const std::vector<C> &coef() const { return coef_; }
};

struct BiquadD {  // biquad.rs:74-81 + simple.rs Identity
    sdr_biquad_design_t d;
    static BiquadD LowPass(float f, float q) { return {{SDR_BQ_LOWPASS, f, q}}; }
    static BiquadD HighPass(float f, float q) { return {{SDR_BQ_HIGHPASS, f, q}}; }
    static BiquadD BandPass(float f, float q) { return {{SDR_BQ_BANDPASS, f, q}}; }
    static BiquadD Notch(float f, float q) { return {{SDR_BQ_NOTCH, f, q}}; }
    static BiquadD Lr(float decayrate) { return {{SDR_BQ_LR, decayrate, 0.0f}}; }
    static BiquadD Identity() { return {{SDR_BQ_IDENTITY, 0.0f, 0.0f}}; }
};

// Biquad<f32, A> (biquad.rs:5-71) as a block stream filter; BiquadD::design(rate) (biquad.rs:83-154) builds it
template <class A>
class Biquad {
    sdr_biquad_t *h_ = nullptr;

  public:
    Biquad(const BiquadD &d, float rate) {
        sdr_biquad_config_t cfg{};
        cfg.designs = &d.d;
        cfg.n_designs = 1;
        cfg.n_streams = 1;
        cfg.rate = rate;
        cfg.sample_complex = SampleTraits<A>::channels == 2;
        int err = 0;
        h_ = sdr_biquad_create(&cfg, &err);
        if (!h_) throw sdr::Error(err, "BiquadD::design");
    }
    Biquad(Biquad &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Biquad(const Biquad &o) {
        int err = 0;
        h_ = sdr_biquad_clone(o.h_, &err);
        if (!h_) throw sdr::Error(err, "Biquad::clone");
    }
    ~Biquad() { sdr_biquad_destroy(h_); }
    void process(const A *in, size_t n, std::vector<A> &out) {
        out.resize(n);
        check(sdr_biquad_process(h_, reinterpret_cast<const float *>(in), n, n, reinterpret_cast<float *>(out.data()), n),
              "Biquad::apply");
    }
    void reset() { check(sdr_biquad_reset(h_), "Biquad::reset"); }
};

class Pll;
struct PllDesign {  // pll.rs:4-36
    float reference, gain;
    BiquadD loopfilter, outputfilter, lockfilter;
    PllDesign(float reference_, float gain_, BiquadD loop, BiquadD output, BiquadD lock)
        : reference(reference_), gain(gain_), loopfilter(loop), outputfilter(output), lockfilter(lock) {}
    Pll design(float rate) const;  // FilterDesign::design (pll.rs:48-60)
};

class Pll {  // pll.rs:13-85 ; Output = Option<f32>
    sdr_pll_t *h_ = nullptr;

  public:
    Pll(const PllDesign &d, float rate, unsigned flags = 0) {
        sdr_pll_design_t cd{d.reference, d.gain, d.loopfilter.d, d.outputfilter.d, d.lockfilter.d};
        sdr_pll_config_t cfg{};
        cfg.designs = &cd;
        cfg.n_designs = 1;
        cfg.n_streams = 1;
        cfg.rate = rate;
        cfg.flags = flags;
        int err = 0;
        h_ = sdr_pll_create(&cfg, &err);
        if (!h_) throw sdr::Error(err, "PllDesign::design");
    }
    Pll(Pll &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Pll(const Pll &o) {
        int err = 0;
        h_ = sdr_pll_clone(o.h_, &err);
        if (!h_) throw sdr::Error(err, "Pll::clone");
    }
    ~Pll() { sdr_pll_destroy(h_); }
    void process(const Complex *in, size_t n, std::vector<std::optional<float>> &out) {
        std::vector<float> v(n);
        std::vector<uint8_t> lk(n);
        check(sdr_pll_process(h_, reinterpret_cast<const float *>(in), n, n, v.data(), lk.data(), n), "Pll::process");
        out.resize(n);
        for (size_t i = 0; i < n; ++i) out[i] = lk[i] ? std::optional<float>(v[i]) : std::nullopt;
    }
    std::optional<float> apply(Complex value) {
        std::vector<std::optional<float>> o;
        process(&value, 1, o);
        return o[0];
    }
    // the closure of main.rs:62-71 with this Pll as `pllpilot`: v -> (mono, diff)
    void stereo_decode(const float *v, size_t n, std::vector<std::pair<float, float>> &out) {
        out.resize(n);
        check(sdr_pll_stereo_decode(h_, v, n, n, reinterpret_cast<float *>(out.data()), n), "Pll::stereo_decode");
    }
    float nphase() { float a, b, c; check(sdr_pll_get_state(h_, 0, &a, &b, &c), "Pll::nphase"); return a; }  // pub nphase
    Complex value() { float a, b, c; check(sdr_pll_get_state(h_, 0, &a, &b, &c), "Pll::value"); return {b, c}; }  // pub value
};
inline Pll PllDesign::design(float rate) const { return Pll(*this, rate); }

}  // namespace filter

// =========================================================================================
namespace signal {

inline size_t decimate_wait(float rate_in, float rate_out) { return sdr_decimate_wait(rate_in, rate_out); }

// A Signal yields blocks: next_block appends up to n samples to `out` and returns how many (0 = end).
template <class SampleT>
class Signal : public std::enable_shared_from_this<Signal<SampleT>> {
  public:
    using Sample = SampleT;
    virtual ~Signal() = default;
    virtual size_t next_block(size_t n, std::vector<Sample> &out) = 0;
    virtual float rate() const = 0;
    // rtl_tcp sources can hand out raw bytes so the consumer kernel unpacks on the GPU
    virtual bool has_raw_u8iq() const { return false; }
    virtual size_t next_raw(size_t, std::vector<uint8_t> &) { return 0; }

    std::vector<Sample> collect(size_t block = (size_t)1 << 20) {  // .iter().collect()
        std::vector<Sample> all, b;
        for (;;) {
            b.clear();
            if (next_block(block, b) == 0) break;
            all.insert(all.end(), b.begin(), b.end());
        }
        return all;
    }
};
template <class A> using Sig = std::shared_ptr<Signal<A>>;

template <class A>
class FromIter : public Signal<A> {  // sources.rs:6-36
    std::vector<A> data_;
    size_t pos_ = 0;
    float rate_;

  public:
    FromIter(float rate, std::vector<A> d) : data_(std::move(d)), rate_(rate) {}
    size_t next_block(size_t n, std::vector<A> &out) override {
        const size_t k = std::min(n, data_.size() - pos_);
        out.insert(out.end(), data_.begin() + pos_, data_.begin() + pos_ + k);
        pos_ += k;
        return k;
    }
    float rate() const override { return rate_; }
};
template <class A> Sig<A> from_iter(float rate, std::vector<A> d) { return std::make_shared<FromIter<A>>(rate, std::move(d)); }

class RtlTcpBytes : public Signal<Complex> {  // RtlTcpSignal (rtltcp.rs:151-168) over a captured byte buffer
    std::vector<uint8_t> raw_;
    size_t pos_ = 0;  // samples
    float rate_;

  public:
    RtlTcpBytes(float rate, std::vector<uint8_t> raw) : raw_(std::move(raw)), rate_(rate) {}
    bool has_raw_u8iq() const override { return true; }
    size_t next_raw(size_t n, std::vector<uint8_t> &out) override {
        const size_t k = std::min(n, raw_.size() / 2 - pos_);
        out.insert(out.end(), raw_.begin() + 2 * pos_, raw_.begin() + 2 * (pos_ + k));
        pos_ += k;
        return k;
    }
    size_t next_block(size_t n, std::vector<Complex> &out) override {
        std::vector<uint8_t> b;
        const size_t k = next_raw(n, b);
        if (!k) return 0;
        const size_t o = out.size();
        out.resize(o + k);
        check(sdr_unpack_u8iq(b.data(), k, reinterpret_cast<float *>(out.data() + o), 0), "RtlTcpSignal::next");
        return k;
    }
    float rate() const override { return rate_; }
};
inline Sig<Complex> from_u8iq(float rate, std::vector<uint8_t> raw) { return std::make_shared<RtlTcpBytes>(rate, std::move(raw)); }

// signal::Filter (adapters/mod.rs:67-100) for a Fir design; decimation > 1 = Filter followed by Decimate, fused
template <class C, class A>
class FirFilter : public Signal<A> {
    Sig<A> up_;
    std::vector<C> taps_;
    size_t D_;
    std::unique_ptr<filter::Fir<C, A>> fir_;
    bool raw_;

  public:
    FirFilter(Sig<A> up, std::vector<C> taps, size_t D = 1) : up_(std::move(up)), taps_(std::move(taps)), D_(D) {
        raw_ = std::is_same<A, Complex>::value && up_->has_raw_u8iq();
    }
    const std::vector<C> &taps() const { return taps_; }
    Sig<A> upstream() const { return up_; }
    bool started() const { return (bool)fir_; }
    size_t next_block(size_t n, std::vector<A> &out) override {
        if (!fir_) fir_.reset(new filter::Fir<C, A>(taps_, D_, raw_));
        for (;;) {
            std::vector<A> y;
            if (raw_) {
                std::vector<uint8_t> b;
                const size_t k = up_->next_raw(n * D_, b);
                if (!k) return 0;
                fir_->process(b.data(), k, y);
            } else {
                std::vector<A> x;
                const size_t k = up_->next_block(n * D_, x);
                if (!k) return 0;
                fir_->process(x.data(), k, y);
            }
            if (y.empty()) continue;
            out.insert(out.end(), y.begin(), y.end());
            return y.size();
        }
    }
    float rate() const override { return up_->rate(); }
};

// signal::Decimate (adapters/mod.rs:14-41); rate() is the upstream rate (:38-40, reference quirk)
template <class A>
class Decimate : public Signal<A> {
    Sig<A> up_;
    size_t wait_, phase_ = 0;

  public:
    Decimate(Sig<A> up, float rate) : up_(std::move(up)) {
        wait_ = decimate_wait(up_->rate(), rate);
        if (wait_ == 0) throw sdr::Error(SDR_ERR_INVALID_ARG, "Decimate (wait == 0 underflows in the reference)");
    }
    size_t wait() const { return wait_; }
    size_t next_block(size_t n, std::vector<A> &out) override {
        for (;;) {
            std::vector<A> x;
            const size_t k = up_->next_block(n * wait_, x);
            if (!k) return 0;
            size_t made = 0;
            for (size_t i = wait_ - 1 - phase_; i < k; i += wait_) { out.push_back(x[i]); ++made; }
            phase_ = (phase_ + k) % wait_;
            if (made) return made;
        }
    }
    float rate() const override { return up_->rate(); }
};

// signal::Resample (adapters/resample.rs:5-86)
template <class A>
class Resample : public Signal<A> {
    Sig<A> up_;
    resample::SampleRate<A> sr_;
    float rate_;
    double ratio_;
    std::vector<A> buffer_, resampled_;
    size_t buffer_size_ = 4096, next_ = 4096;
    bool done_ = false;

  public:
    Resample(Sig<A> up, resample::ConverterType typ, float rate)
        : up_(std::move(up)), sr_(typ), rate_(rate), ratio_((double)rate / (double)up_->rate()) {
        buffer_.reserve(buffer_size_);
        resampled_.reserve(buffer_size_);
    }
    size_t next_block(size_t n, std::vector<A> &out) override {
        size_t made = 0;
        while (made < n) {
            if (done_) break;
            while (next_ >= resampled_.size()) {
                if (buffer_.size() < buffer_size_) up_->next_block(buffer_size_ - buffer_.size(), buffer_);  // refill (:46-52)
                resampled_.reserve(buffer_size_);
                const size_t used = sr_.process(ratio_, buffer_, resampled_);                              // (:55-59)
                if (buffer_.empty() && resampled_.empty()) { done_ = true; break; }                          // (:62-65)
                buffer_.erase(buffer_.begin(), buffer_.begin() + used);                                      // (:68)
                if (resampled_.empty()) continue;                                                            // (:70-73)
                next_ = 0;
            }
            if (done_) break;
            const size_t k = std::min(n - made, resampled_.size() - next_);
            out.insert(out.end(), resampled_.begin() + next_, resampled_.begin() + next_ + k);
            next_ += k;
            made += k;
            if (made) break;  // hand back what one chunk produced; callers loop
        }
        return made;
    }
    float rate() const override { return rate_; }
};

template <class A>
class Take : public Signal<A> {  // adapters/mod.rs:241-268
    Sig<A> up_;
    size_t left_;

  public:
    Take(Sig<A> up, float duration) : up_(std::move(up)) { left_ = sdr_duration_samples(up_->rate(), duration); }
    size_t next_block(size_t n, std::vector<A> &out) override {
        n = std::min(n, left_);
        if (!n) return 0;
        const size_t k = up_->next_block(n, out);
        left_ -= k;
        return k;
    }
    float rate() const override { return up_->rate(); }
};

template <class A>
class Skip : public Signal<A> {  // adapters/mod.rs:166-194
    Sig<A> up_;
    size_t left_;

  public:
    Skip(Sig<A> up, float duration) : up_(std::move(up)) { left_ = sdr_duration_samples(up_->rate(), duration); }
    size_t next_block(size_t n, std::vector<A> &out) override {
        while (left_ > 0) {
            std::vector<A> junk;
            const size_t k = up_->next_block(std::min<size_t>(left_, 1 << 20), junk);
            if (!k) return 0;
            left_ -= k;
        }
        return up_->next_block(n, out);
    }
    float rate() const override { return up_->rate(); }
};

// signal::Block (adapters/block.rs:106-207).  The upstream is pulled in blocks of ceil(size * rate) samples (:117),
// one block ahead of the reader (target = 1, :165-189); clone() is a tee (:129-140): clones share the upstream and one
// deque of blocks with a per-reader `available` count (TeeDeque, :7-103) -- a new reader starts with available =
// data.len(), so it also sees the blocks the deque still held.  Faithful to the reference's quirk that `next` returns
// current[0] without advancing `i` when it moves to a new block (:197-199): the first sample of every block is
// delivered twice unless dedup is set (a deliberate divergence).
template <class A>
class Block : public Signal<A> {
    struct Shared {
        Sig<A> up;
        std::deque<std::vector<A>> data;   // samples; newest block at the front
        std::deque<std::vector<uint8_t>> raw;  // same blocks as raw rtl_tcp bytes when the upstream hands those out
        std::vector<size_t> available{0};
        size_t block_size = 0;
        bool is_raw = false, dedup = false;
        float rate = 0.0f;
    };
    std::shared_ptr<Shared> sh_;
    size_t id_ = 0, i_ = 0, cur_len_ = 0;
    std::vector<A> cur_;
    std::vector<uint8_t> cur_raw_;
    bool have_cur_ = false;

    void push() {  // TeeDequePush::push (:76-89) around the fill closure (:176-187)
        Shared &s = *sh_;
        const size_t held = s.is_raw ? s.raw.size() : s.data.size();
        if (*std::max_element(s.available.begin(), s.available.end()) < held) {
            if (s.is_raw) s.raw.pop_back(); else s.data.pop_back();
        }
        if (s.is_raw) { std::vector<uint8_t> v; s.up->next_raw(s.block_size, v); s.raw.push_front(std::move(v)); }
        else { std::vector<A> v; s.up->next_block(s.block_size, v); s.data.push_front(std::move(v)); }
        for (auto &a : s.available) ++a;
    }
    void load(size_t idx) {
        Shared &s = *sh_;
        if (s.is_raw) { cur_raw_ = s.raw[idx]; cur_len_ = cur_raw_.size() / 2; }
        else { cur_ = s.data[idx]; cur_len_ = cur_.size(); }
    }
    size_t fetch() {  // the else-branch of Block::next (:154-195)
        Shared &s = *sh_;
        cur_.clear(); cur_raw_.clear(); cur_len_ = 0; i_ = 0;
        bool needs_extra = true;
        size_t avail = 0;
        if (s.available[id_] > 0) { avail = --s.available[id_]; load(avail); needs_extra = false; }
        if (avail < 1) {
            const size_t jobs = 1 - avail + (needs_extra ? 1 : 0);
            for (size_t j = 0; j < jobs; ++j) push();
            if (needs_extra) load(--s.available[id_]);
        }
        have_cur_ = true;
        return cur_len_;
    }
    template <class V, class T>
    size_t serve(size_t n, std::vector<T> &out, const V &cur, size_t per) {
        size_t made = 0;
        if (!have_cur_ || i_ >= cur_len_) {
            // note: fetch() re-fills `cur`, which the caller passed by reference
            if (fetch() == 0 || n == 0) return 0;
            if (!sh_->dedup) { out.insert(out.end(), cur.begin(), cur.begin() + per); made = 1; }  // (:197-199)
        }
        const size_t k = std::min(n - made, cur_len_ - i_);
        out.insert(out.end(), cur.begin() + per * i_, cur.begin() + per * (i_ + k));
        i_ += k;
        return made + k;
    }

  public:
    Block(Sig<A> up, float size, bool dedup = false) : sh_(std::make_shared<Shared>()) {
        sh_->rate = up->rate();
        sh_->block_size = sdr_block_samples(size, up->rate());
        sh_->is_raw = up->has_raw_u8iq();
        sh_->dedup = dedup;
        sh_->up = std::move(up);
    }
    // Clone for Block (:129-140) + Clone for TeeDeque (:92-103)
    std::shared_ptr<Block<A>> clone() const {
        auto b = std::shared_ptr<Block<A>>(new Block<A>(*this));
        b->cur_.clear(); b->cur_raw_.clear(); b->cur_len_ = 0; b->i_ = 0; b->have_cur_ = false;
        Shared &s = *sh_;
        s.available.push_back(s.is_raw ? s.raw.size() : s.data.size());
        b->id_ = s.available.size() - 1;
        return b;
    }
    size_t block_size() const { return sh_->block_size; }
    size_t next_block(size_t n, std::vector<A> &out) override {
        if (!sh_->is_raw) return serve(n, out, cur_, 1);
        // raw upstream, sample consumer: unpack the served bytes (RtlTcpSignal::next, rtltcp.rs:158-164)
        std::vector<uint8_t> b;
        const size_t k = serve(n, b, cur_raw_, 2);
        if (!k) return 0;
        const size_t o = out.size();
        out.resize(o + k);
        if (std::is_same<A, Complex>::value)
            check(sdr_unpack_u8iq(b.data(), k, reinterpret_cast<float *>(out.data() + o), 0), "Block::next");
        return k;
    }
    bool has_raw_u8iq() const override { return sh_->is_raw; }
    size_t next_raw(size_t n, std::vector<uint8_t> &out) override { return serve(n, out, cur_raw_, 2); }
    float rate() const override { return sh_->rate; }
};

template <class A, class B>
class Map : public Signal<B> {  // adapters/mod.rs:139-163
    Sig<A> up_;
    std::function<B(A)> f_;

  public:
    Map(Sig<A> up, std::function<B(A)> f) : up_(std::move(up)), f_(std::move(f)) {}
    size_t next_block(size_t n, std::vector<B> &out) override {
        std::vector<A> x;
        const size_t k = up_->next_block(n, x);
        for (size_t i = 0; i < k; ++i) out.push_back(f_(x[i]));
        return k;
    }
    float rate() const override { return up_->rate(); }
};

// ---- the combinators of trait Signal (src/signal/mod.rs:18-122) as free functions over Sig<A> ----
template <class A, class C> Sig<A> filter(Sig<A> s, std::vector<C> taps) { return std::make_shared<FirFilter<C, A>>(std::move(s), std::move(taps)); }
template <class A> Sig<A> decimate(Sig<A> s, float rate) {
    // Filter followed by Decimate: fuse so only kept outputs are computed
    if (auto f = std::dynamic_pointer_cast<FirFilter<float, A>>(s))
        if (!f->started()) return std::make_shared<FirFilter<float, A>>(f->upstream(), f->taps(), decimate_wait(s->rate(), rate));
    if (auto f = std::dynamic_pointer_cast<FirFilter<Complex, A>>(s))
        if (!f->started()) return std::make_shared<FirFilter<Complex, A>>(f->upstream(), f->taps(), decimate_wait(s->rate(), rate));
    return std::make_shared<Decimate<A>>(std::move(s), rate);
}
template <class A> Sig<A> resample_with(Sig<A> s, resample::ConverterType typ, float rate) { return std::make_shared<Resample<A>>(std::move(s), typ, rate); }
template <class A> Sig<A> resample(Sig<A> s, float rate) { return resample_with(std::move(s), resample::ConverterType::SincBestQuality, rate); }  // mod.rs:83
template <class A> Sig<A> take(Sig<A> s, float duration) { return std::make_shared<Take<A>>(std::move(s), duration); }
template <class A> Sig<A> skip(Sig<A> s, float duration) { return std::make_shared<Skip<A>>(std::move(s), duration); }
template <class A> std::shared_ptr<Block<A>> block(Sig<A> s, float size, bool dedup = false) { return std::make_shared<Block<A>>(std::move(s), size, dedup); }
template <class A, class F> auto map(Sig<A> s, F f) -> Sig<decltype(f(std::declval<A>()))> {
    using B = decltype(f(std::declval<A>()));
    return std::make_shared<Map<A, B>>(std::move(s), std::function<B(A)>(f));
}

}  // namespace signal

// =========================================================================================
namespace fft {

// fft::fft (src/fft.rs:3-28): drains the signal, one N-point transform, (label, value) pairs
inline std::vector<std::pair<float, Complex>> fft(signal::Sig<Complex> input) {
    const float rate = input->rate();
    std::vector<std::pair<float, Complex>> collated;
    sdr_fft_config_t cfg{};
    cfg.flags = SDR_FFT_SHIFT | SDR_FFT_NORM;
    std::vector<uint8_t> raw;
    std::vector<Complex> data;
    const void *src;
    if (input->has_raw_u8iq()) {
        input->next_raw((size_t)1 << 62, raw);
        cfg.n = raw.size() / 2;
        cfg.input_format = SDR_FMT_U8IQ;
        src = raw.data();
    } else {
        data = input->collect();
        cfg.n = data.size();
        cfg.input_format = SDR_FMT_C64;
        src = data.data();
    }
    if (cfg.n == 0) return collated;
    int err = 0;
    sdr_fft_t *plan = sdr_fft_create(&cfg, &err);
    if (!plan) throw Error(err, "fft::fft");
    std::vector<Complex> vals(cfg.n);
    std::vector<float> labels(cfg.n);
    const int rc = sdr_fft_exec(plan, src, 1, reinterpret_cast<float *>(vals.data()));
    sdr_fft_destroy(plan);
    check(rc, "fft::fft");
    check(sdr_fft_labels(cfg.n, rate, 0, labels.data()), "fft::fft labels");
    collated.reserve(cfg.n);
    for (size_t i = 0; i < cfg.n; ++i) collated.emplace_back(labels[i], vals[i]);
    return collated;
}

// fft::rfft (src/fft.rs:30-37)
inline std::vector<std::pair<float, Complex>> rfft(signal::Sig<float> input) {
    const float rate = input->rate();
    std::vector<float> data = input->collect();
    std::vector<std::pair<float, Complex>> collated;
    if (data.empty()) return collated;
    sdr_fft_config_t cfg{};
    cfg.n = data.size();
    cfg.input_format = SDR_FMT_F32;
    cfg.flags = SDR_FFT_SHIFT | SDR_FFT_NORM | SDR_FFT_RFFT;
    int err = 0;
    sdr_fft_t *plan = sdr_fft_create(&cfg, &err);
    if (!plan) throw Error(err, "fft::rfft");
    const size_t keep = sdr_fft_output_len(plan);
    std::vector<Complex> vals(keep);
    std::vector<float> labels(keep);
    const int rc = sdr_fft_exec(plan, data.data(), 1, reinterpret_cast<float *>(vals.data()));
    sdr_fft_destroy(plan);
    check(rc, "fft::rfft");
    check(sdr_fft_labels(cfg.n, rate, 1, labels.data()), "fft::rfft labels");
    for (size_t i = 0; i < keep; ++i) collated.emplace_back(labels[i], vals[i]);
    return collated;
}

// signal.window(duration).decimate(fps).map(|w| fft::fft(w)) of examples/live.rs:30-39 behind sdr_window_fft_*:
// feed blocks of samples, get the kept windows' spectra (each `window()` values, shifted and 1/sqrt(N)-scaled like
// fft::fft) back to back.  window = (duration * rate).round() (adapters/mod.rs:279), hop = Decimate's wait (mod.rs:22).
class WindowSpectra {
    sdr_window_fft_t *h_ = nullptr;
    size_t n_ = 0;

  public:
    WindowSpectra(float rate, float duration, float fps, bool u8iq = false) {
        sdr_window_fft_config_t cfg{};
        cfg.window = sdr_duration_samples(rate, duration);
        cfg.hop = sdr_decimate_wait(rate, fps);
        cfg.input_format = u8iq ? SDR_FMT_U8IQ : SDR_FMT_C64;
        cfg.flags = SDR_FFT_SHIFT | SDR_FFT_NORM;
        int err = 0;
        h_ = sdr_window_fft_create(&cfg, &err);
        if (!h_) throw Error(err, "WindowSpectra::new");
        n_ = cfg.window;
    }
    WindowSpectra(const WindowSpectra &) = delete;
    WindowSpectra &operator=(const WindowSpectra &) = delete;
    ~WindowSpectra() { sdr_window_fft_destroy(h_); }
    size_t window() const { return n_; }
    // in: n samples (Complex, or 2n bytes when constructed with u8iq); returns the number of spectra appended to out
    size_t process(const void *in, size_t n, std::vector<Complex> &out) {
        const size_t nw = sdr_window_fft_output_count(h_, n);
        const size_t at = out.size();
        out.resize(at + nw * n_);
        size_t got = 0;
        Complex dummy;
        check(sdr_window_fft_process(h_, in, n, reinterpret_cast<float *>(nw ? out.data() + at : &dummy), nw, &got),
              "WindowSpectra::process");
        return got;
    }
};

}  // namespace fft

// =========================================================================================
namespace app {

// The receiver of src/main.rs:32-81 as one device-resident pipeline (sdr_fm_*): rtl_tcp bytes in, 48 kHz
// (left, right) frames out; a batch of stations shares every launch.
class FmStereo {
    sdr_fm_t *h_ = nullptr;
    size_t n_st_;

  public:
    explicit FmStereo(size_t n_stations = 1, float rate = 1800000.0f, unsigned flags = 0) : n_st_(n_stations) {
        sdr_fm_config_t cfg{};
        cfg.n_stations = n_stations;
        cfg.rate = rate;
        cfg.flags = flags;
        int err = 0;
        h_ = sdr_fm_create(&cfg, &err);
        if (!h_) throw Error(err, "FmStereo::new");
    }
    FmStereo(const FmStereo &) = delete;
    FmStereo &operator=(const FmStereo &) = delete;
    ~FmStereo() { sdr_fm_destroy(h_); }
    float rate() const { return sdr_fm_output_rate(h_); }
    // iq: n_stations rows of 2n bytes; out: n_stations rows of the returned number of frames
    size_t process(const uint8_t *iq, size_t n, bool end_of_input, std::vector<std::pair<float, float>> &out) {
        const size_t cap = sdr_fm_max_output(h_, n);
        out.resize(n_st_ * cap);
        size_t got = 0;
        check(sdr_fm_process(h_, iq, n, 2 * n, reinterpret_cast<float *>(out.data()), cap, cap, &got, end_of_input ? 1 : 0),
              "FmStereo::process");
        if (got != cap)
            for (size_t s = 1; s < n_st_; ++s) std::copy(out.begin() + s * cap, out.begin() + s * cap + got, out.begin() + s * got);
        out.resize(n_st_ * got);
        return got;
    }
    void reset() { check(sdr_fm_reset(h_), "FmStereo::reset"); }
};

}  // namespace app
}  // namespace sdr
