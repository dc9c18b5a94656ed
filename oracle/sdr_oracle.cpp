// sdr_oracle.cpp -- CPU ORACLE (test infrastructure, never shipped, never on the product path).
// See sdr_oracle.h for scope and parity status.  Citations are file:line in the read-only
// reference tree (agrif/unnamed-rust-sdr).  Nothing here is copied from it: the reference is
// Rust, this is an independent C++ restatement of its arithmetic, operation by operation.
//
// Build flags matter: -ffp-contract=off (rustc never fuses a*b+c), no -ffast-math.
#include "sdr_oracle.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace {

// Rust `x as usize` for f32: saturating, NaN -> 0.
inline size_t f32_as_usize(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)v;
}
// Rust f32::round = half away from zero = C roundf.
inline float rs_round(float v) { return roundf(v); }
// Rust f32::fract = self - self.trunc()
inline float rs_fract(float v) { return v - truncf(v); }

const float PI_F = 3.14159265358979323846f;  // std::f32::consts::PI

template <class F>
void parallel_for(size_t n_units, int threads, F &&fn) {
    if (threads <= 1 || n_units <= 1) {
        fn(0, n_units, 0);
        return;
    }
    size_t nt = std::min<size_t>((size_t)threads, n_units);
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (size_t t = 0; t < nt; ++t) {
        size_t lo = n_units * t / nt, hi = n_units * (t + 1) / nt;
        pool.emplace_back([&fn, lo, hi, t]() { fn(lo, hi, (int)t); });
    }
    for (auto &th : pool) th.join();
}

}  // namespace

// =====================================================================================
// a1. unpack -- src/rtltcp.rs:158-164:  ((iq.re as f32 - 128.0) / 128.0, (iq.im as f32 - 128.0) / 128.0)
// =====================================================================================
extern "C" void orc_unpack_u8iq(const uint8_t *iq, size_t n, float *out) {
    for (size_t i = 0; i < n; ++i) {
        out[2 * i + 0] = ((float)iq[2 * i + 0] - 128.0f) / 128.0f;
        out[2 * i + 1] = ((float)iq[2 * i + 1] - 128.0f) / 128.0f;
    }
}

// =====================================================================================
// a2. Fir<C,A> -- src/filter/fir.rs:7-33
//   new():   buffer = K zeros (fir.rs:13-18)
//   apply(): pop_back, push_front(value); accum = 0; for (c, v) in coef.zip(buffer): accum += v*c
//   Convolve::accumulate (convolve.rs:13-15):  *self += a.clone() * c.clone()   (mul, then add)
//   num-complex 0.2:  Complex*f32 = (re*c, im*c);  Complex*Complex = (ar*br - ai*bi, ar*bi + ai*br)
// The ring below behaves like the VecDeque: logical index 0 is the newest sample.
// =====================================================================================
struct orc_fir {
    std::vector<float> coef;  // K or 2K floats
    size_t K;
    int taps_complex, kind;
    std::vector<float> ring;  // K * kind floats
    size_t head;              // physical index of logical element 0
};

extern "C" orc_fir_t *orc_fir_new(const float *taps, size_t n_taps, int taps_complex, int kind) {
    if (kind != ORC_KIND_F32 && kind != ORC_KIND_C64) return nullptr;
    if (taps_complex && kind != ORC_KIND_C64) return nullptr;  // f32 * Complex is not a Convolve impl
    orc_fir *f = new orc_fir;
    f->K = n_taps;
    f->taps_complex = taps_complex ? 1 : 0;
    f->kind = kind;
    f->coef.assign(taps, taps + n_taps * (taps_complex ? 2 : 1));
    f->ring.assign(n_taps * kind, 0.0f);
    f->head = 0;
    return f;
}
extern "C" void orc_fir_free(orc_fir_t *f) { delete f; }
extern "C" void orc_fir_reset(orc_fir_t *f) {
    std::fill(f->ring.begin(), f->ring.end(), 0.0f);
    f->head = 0;
}
extern "C" orc_fir_t *orc_fir_clone(const orc_fir_t *f) { return new orc_fir(*f); }

extern "C" void orc_fir_apply(orc_fir_t *f, const float *in, size_t n, float *out) {
    const size_t K = f->K;
    if (K == 0) {  // zip over an empty coef: accum stays zero; pop_back/push_front on empty deque
        // (VecDeque::pop_back on empty is a no-op, push_front grows it; coef.zip stops at 0 terms)
        std::memset(out, 0, n * f->kind * sizeof(float));
        return;
    }
    const float *c = f->coef.data();
    float *ring = f->ring.data();
    if (f->kind == ORC_KIND_F32) {
        for (size_t i = 0; i < n; ++i) {
            f->head = (f->head + K - 1) % K;  // pop_back + push_front
            ring[f->head] = in[i];
            float acc = 0.0f;
            size_t k = 0;
            for (size_t p = f->head; p < K; ++p, ++k) acc += ring[p] * c[k];  // first slice
            for (size_t p = 0; p < f->head; ++p, ++k) acc += ring[p] * c[k];  // wrapped slice
            out[i] = acc;
        }
    } else if (!f->taps_complex) {
        for (size_t i = 0; i < n; ++i) {
            f->head = (f->head + K - 1) % K;
            ring[2 * f->head] = in[2 * i];
            ring[2 * f->head + 1] = in[2 * i + 1];
            float ar = 0.0f, ai = 0.0f;
            size_t k = 0;
            for (size_t p = f->head; p < K; ++p, ++k) {
                ar += ring[2 * p] * c[k];
                ai += ring[2 * p + 1] * c[k];
            }
            for (size_t p = 0; p < f->head; ++p, ++k) {
                ar += ring[2 * p] * c[k];
                ai += ring[2 * p + 1] * c[k];
            }
            out[2 * i] = ar;
            out[2 * i + 1] = ai;
        }
    } else {
        for (size_t i = 0; i < n; ++i) {
            f->head = (f->head + K - 1) % K;
            ring[2 * f->head] = in[2 * i];
            ring[2 * f->head + 1] = in[2 * i + 1];
            float ar = 0.0f, ai = 0.0f;
            size_t k = 0;
            auto mac = [&](size_t p) {
                const float vr = ring[2 * p], vi = ring[2 * p + 1];
                const float cr = c[2 * k], ci = c[2 * k + 1];
                const float pr = vr * cr - vi * ci;  // (v * c).re
                const float pi = vr * ci + vi * cr;  // (v * c).im
                ar += pr;
                ai += pi;
            };
            for (size_t p = f->head; p < K; ++p, ++k) mac(p);
            for (size_t p = 0; p < f->head; ++p, ++k) mac(p);
            out[2 * i] = ar;
            out[2 * i + 1] = ai;
        }
    }
}

extern "C" void orc_fir_f64(const float *taps, size_t K, int taps_complex, int kind, const float *in,
                            size_t n, double *out) {
    for (size_t i = 0; i < n; ++i) {
        double ar = 0.0, ai = 0.0;
        size_t kmax = std::min(K, i + 1);
        for (size_t k = 0; k < kmax; ++k) {
            if (kind == ORC_KIND_F32) {
                ar += (double)in[i - k] * (double)taps[k];
            } else if (!taps_complex) {
                ar += (double)in[2 * (i - k)] * (double)taps[k];
                ai += (double)in[2 * (i - k) + 1] * (double)taps[k];
            } else {
                double vr = in[2 * (i - k)], vi = in[2 * (i - k) + 1];
                double cr = taps[2 * k], ci = taps[2 * k + 1];
                ar += vr * cr - vi * ci;
                ai += vr * ci + vi * cr;
            }
        }
        if (kind == ORC_KIND_F32)
            out[i] = ar;
        else {
            out[2 * i] = ar;
            out[2 * i + 1] = ai;
        }
    }
}

// =====================================================================================
// a4. Decimate -- src/signal/adapters/mod.rs:14-41
//   wait = (signal.rate() / rate).round() as usize           (:22)
//   next(): discard wait-1 upstream samples, return the next one   (:30-37)
//   (wait == 0 makes `self.wait - 1` underflow in the reference; the oracle reports 0 outputs)
// =====================================================================================
extern "C" size_t orc_decimate_wait(float rate_in, float rate_out) {
    return f32_as_usize(rs_round(rate_in / rate_out));
}
extern "C" size_t orc_decimate(const float *in, size_t n, size_t wait, int ef, size_t *phase,
                               float *out) {
    if (wait == 0) return 0;
    size_t ph = phase ? *phase : 0, n_out = 0;
    for (size_t i = 0; i < n; ++i) {
        if (ph == wait - 1) {
            for (int e = 0; e < ef; ++e) out[n_out * ef + e] = in[i * ef + e];
            ++n_out;
            ph = 0;
        } else {
            ++ph;
        }
    }
    if (phase) *phase = ph;
    return n_out;
}

// Take/Skip: (signal.rate() * duration).round() as usize  -- adapters/mod.rs:174,249 ; Window :279
extern "C" size_t orc_round_count(float rate, float duration) {
    return f32_as_usize(rs_round(rate * duration));
}
// Block: (size * signal.rate()).ceil() as usize -- adapters/block.rs:117
extern "C" size_t orc_block_size(float size, float rate) { return f32_as_usize(ceilf(size * rate)); }
// Times: (now as f32) / self.rate -- times.rs:17-21
extern "C" void orc_times(float rate, size_t start, size_t n, float *out) {
    for (size_t i = 0; i < n; ++i) out[i] = (float)(start + i) / rate;
}

// =====================================================================================
// a7. FFT.  fft.rs:10-12 calls rustfft 3.0 (not in the tree; PARITY UNPINNED): a forward,
// unnormalised DFT in f32 whose twiddles are computed in f64 and rounded to f32, re-planned
// on every call.  Restated here as a Stockham mixed-radix (4,2,3,5,generic) transform.
// =====================================================================================
namespace {

template <class T>
struct FftPlan {
    size_t n;
    std::vector<size_t> radices;
    std::vector<std::complex<T>> tw;  // W_n^k, k in [0,n)
    explicit FftPlan(size_t n_) : n(n_) {
        size_t m = n;
        while (m % 4 == 0) { radices.push_back(4); m /= 4; }
        while (m % 2 == 0) { radices.push_back(2); m /= 2; }
        for (size_t p = 3; p * p <= m; p += 2)
            while (m % p == 0) { radices.push_back(p); m /= p; }
        if (m > 1) radices.push_back(m);
        tw.resize(n);
        const double c = -2.0 * M_PI / (double)n;
        for (size_t k = 0; k < n; ++k) tw[k] = std::complex<T>((T)std::cos(c * (double)k), (T)std::sin(c * (double)k));
    }
};

template <class T>
inline std::complex<T> cmul(const std::complex<T> &a, const std::complex<T> &b) {
    return std::complex<T>(a.real() * b.real() - a.imag() * b.imag(),
                           a.real() * b.imag() + a.imag() * b.real());
}

template <class T>
void fft_exec(const FftPlan<T> &pl, std::complex<T> *x, std::complex<T> *y) {
    // Stockham autosort, decimation in time.  x: input/scratch, y: scratch; result pointer returned
    // via swap parity -> caller passes both buffers and we copy at the end if needed.
    const size_t n = pl.n;
    std::complex<T> *src = x, *dst = y;
    size_t p = 1;
    std::vector<std::complex<T>> u;
    for (size_t r : pl.radices) {
        const size_t t_cnt = n / r;    // butterflies in this pass
        const size_t tw_step = n / (p * r);
        if (r == 4) {
            for (size_t i = 0; i < t_cnt; ++i) {
                const size_t k = i % p, j = (i - k) * 4 + k;
                std::complex<T> a0 = src[i];
                std::complex<T> a1 = cmul(src[i + t_cnt], pl.tw[k * tw_step]);
                std::complex<T> a2 = cmul(src[i + 2 * t_cnt], pl.tw[2 * k * tw_step]);
                std::complex<T> a3 = cmul(src[i + 3 * t_cnt], pl.tw[3 * k * tw_step]);
                std::complex<T> s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = a1 - a3;
                std::complex<T> jd13(d13.imag(), -d13.real());  // -i * d13
                dst[j] = s02 + s13;
                dst[j + p] = d02 + jd13;
                dst[j + 2 * p] = s02 - s13;
                dst[j + 3 * p] = d02 - jd13;
            }
        } else if (r == 2) {
            for (size_t i = 0; i < t_cnt; ++i) {
                const size_t k = i % p, j = (i - k) * 2 + k;
                std::complex<T> a0 = src[i];
                std::complex<T> a1 = cmul(src[i + t_cnt], pl.tw[k * tw_step]);
                dst[j] = a0 + a1;
                dst[j + p] = a0 - a1;
            }
        } else {
            u.resize(r);
            const size_t wr_step = n / r;  // W_r = W_n^(n/r)
            for (size_t i = 0; i < t_cnt; ++i) {
                const size_t k = i % p, j = (i - k) * r + k;
                for (size_t s = 0; s < r; ++s) u[s] = cmul(src[i + s * t_cnt], pl.tw[(s * k * tw_step) % n]);
                for (size_t t = 0; t < r; ++t) {
                    std::complex<T> acc = u[0];
                    for (size_t s = 1; s < r; ++s) acc += cmul(u[s], pl.tw[((s * t) % r) * wr_step]);
                    dst[j + t * p] = acc;
                }
            }
        }
        std::swap(src, dst);
        p *= r;
    }
    if (src != y) std::copy(src, src + n, y);
}

}  // namespace

extern "C" void orc_fft_f32(const float *in, size_t n, float *out) {
    if (n == 0) return;
    FftPlan<float> pl(n);  // planned on every call, as fft.rs:10-11 does
    std::vector<std::complex<float>> x(n), y(n);
    for (size_t i = 0; i < n; ++i) x[i] = std::complex<float>(in[2 * i], in[2 * i + 1]);
    fft_exec(pl, x.data(), y.data());
    for (size_t i = 0; i < n; ++i) {
        out[2 * i] = y[i].real();
        out[2 * i + 1] = y[i].imag();
    }
}

extern "C" void orc_dft_f64(const float *in, size_t n, double *out) {
    if (n == 0) return;
    FftPlan<double> pl(n);
    std::vector<std::complex<double>> x(n), y(n);
    for (size_t i = 0; i < n; ++i) x[i] = std::complex<double>(in[2 * i], in[2 * i + 1]);
    fft_exec(pl, x.data(), y.data());
    for (size_t i = 0; i < n; ++i) {
        out[2 * i] = y[i].real();
        out[2 * i + 1] = y[i].imag();
    }
}

// fft.rs:14-26:
//   fstep = rate / (len as f32); start = -(len as isize / 2); norm = 1.0 / (len as f32).sqrt()
//   out[i] = (srci as f32 * fstep, output[srci mod len] * norm)        Complex * f32 = (re*n, im*n)
static void shift_norm(const float *X, size_t n, float rate, float *labels, float *vals) {
    const float fstep = rate / (float)n;
    const long start = -((long)n / 2);
    const float norm = 1.0f / sqrtf((float)n);
    for (size_t i = 0; i < n; ++i) {
        long srci = start + (long)i;
        size_t pos = (size_t)(srci < 0 ? srci + (long)n : srci);
        if (labels) labels[i] = (float)srci * fstep;
        vals[2 * i] = X[2 * pos] * norm;
        vals[2 * i + 1] = X[2 * pos + 1] * norm;
    }
}

extern "C" void orc_fft_shifted(const float *in, size_t n, float rate, float *labels, float *vals) {
    if (n == 0) return;
    std::vector<float> X(2 * n);
    orc_fft_f32(in, n, X.data());
    shift_norm(X.data(), n, rate, labels, vals);
}

// fft.rs:30-37: fft(input.map(|v| Complex::new(v, 0.0))); output.drain(0..len/2)
extern "C" size_t orc_rfft_shifted(const float *in, size_t n, float rate, float *labels, float *vals) {
    if (n == 0) return 0;
    std::vector<float> c(2 * n), l(n), v(2 * n);
    for (size_t i = 0; i < n; ++i) { c[2 * i] = in[i]; c[2 * i + 1] = 0.0f; }
    orc_fft_shifted(c.data(), n, rate, l.data(), v.data());
    const size_t drop = n / 2, keep = n - drop;
    for (size_t i = 0; i < keep; ++i) {
        if (labels) labels[i] = l[drop + i];
        vals[2 * i] = v[2 * (drop + i)];
        vals[2 * i + 1] = v[2 * (drop + i) + 1];
    }
    return keep;
}

extern "C" void orc_fft_batch_u8(const uint8_t *iq, size_t n, size_t batches, int threads, float *vals) {
    parallel_for(batches, threads, [&](size_t lo, size_t hi, int) {
        std::vector<float> x(2 * n);
        for (size_t b = lo; b < hi; ++b) {
            orc_unpack_u8iq(iq + 2 * n * b, n, x.data());
            orc_fft_shifted(x.data(), n, 1.0f, nullptr, vals + 2 * n * b);
        }
    });
}
extern "C" void orc_fft_batch_c64(const float *in, size_t n, size_t batches, int shifted, int threads,
                                  float *vals) {
    parallel_for(batches, threads, [&](size_t lo, size_t hi, int) {
        for (size_t b = lo; b < hi; ++b) {
            if (shifted)
                orc_fft_shifted(in + 2 * n * b, n, 1.0f, nullptr, vals + 2 * n * b);
            else
                orc_fft_f32(in + 2 * n * b, n, vals + 2 * n * b);
        }
    });
}

// =====================================================================================
// a9. Biquad -- src/filter/biquad.rs
//   Biquad::new(a0,a1,a2,b0,b1,b2): b0/a0, b1/a0, b2/a0, -a1/a0, -a2/a0       (:25-38)
//   apply: out = 0; out += v*b0; out += x1*b1; out += x2*b2; out += y1*na1; out += y2*na2   (:43-49)
//          x2<-x1<-v ; y2<-y1<-out                                            (:51-54)
//   BiquadD::design (:83-154): f32 throughout; `2.0 * PI * freq / rate` = ((2.0*PI)*freq)/rate
// =====================================================================================
extern "C" void orc_biquad_design(int kind, float p0, float p1, float rate, float coef[5]) {
    float a0 = 1, a1 = 0, a2 = 0, b0 = 1, b1 = 0, b2 = 0;
    if (kind >= ORC_BQ_LOWPASS && kind <= ORC_BQ_NOTCH) {
        const float freq = p0, q = p1;
        const float omega = 2.0f * PI_F * freq / rate;
        const float cs = cosf(omega);
        const float alpha = sinf(omega) / (2.0f * q);
        a0 = 1.0f + alpha;
        a1 = -2.0f * cs;
        a2 = 1.0f - alpha;
        switch (kind) {
            case ORC_BQ_LOWPASS: b0 = (1.0f - cs) / 2.0f; b1 = 1.0f - cs; b2 = (1.0f - cs) / 2.0f; break;
            case ORC_BQ_HIGHPASS: b0 = (1.0f + cs) / 2.0f; b1 = -1.0f - cs; b2 = (1.0f + cs) / 2.0f; break;
            case ORC_BQ_BANDPASS: b0 = alpha; b1 = 0.0f; b2 = -alpha; break;
            default: b0 = 1.0f; b1 = -2.0f * cs; b2 = 1.0f; break;  // Notch
        }
    } else if (kind == ORC_BQ_LR) {
        const float decayn = p0 / rate;
        a0 = 1.0f; a1 = -expf(-decayn); a2 = 0.0f; b0 = decayn; b1 = 0.0f; b2 = 0.0f;
    }
    coef[0] = b0 / a0;
    coef[1] = b1 / a0;
    coef[2] = b2 / a0;
    coef[3] = -a1 / a0;
    coef[4] = -a2 / a0;
}

namespace {
struct BiquadState {
    int kind = ORC_BQ_IDENTITY;
    float b0 = 1, b1 = 0, b2 = 0, na1 = 0, na2 = 0;
    float x1[2] = {0, 0}, x2[2] = {0, 0}, y1[2] = {0, 0}, y2[2] = {0, 0};
    void design(int k, float p0, float p1, float rate) {
        kind = k;
        if (k != ORC_BQ_IDENTITY) {
            float c[5];
            orc_biquad_design(k, p0, p1, rate, c);
            b0 = c[0]; b1 = c[1]; b2 = c[2]; na1 = c[3]; na2 = c[4];
        }
    }
    inline float apply1(float v, int lane = 0) {
        if (kind == ORC_BQ_IDENTITY) return v;
        float out = 0.0f;
        out += v * b0;
        out += x1[lane] * b1;
        out += x2[lane] * b2;
        out += y1[lane] * na1;
        out += y2[lane] * na2;
        x2[lane] = x1[lane]; x1[lane] = v;
        y2[lane] = y1[lane]; y1[lane] = out;
        return out;
    }
    inline void apply2(float vr, float vi, float &orr, float &oi) {  // Complex<f32> * f32 per component
        orr = apply1(vr, 0);
        oi = apply1(vi, 1);
    }
};
}  // namespace

struct orc_biquad { BiquadState s; int kind; };
extern "C" orc_biquad_t *orc_biquad_new(int kind, float p0, float p1, float rate, int sample_kind) {
    orc_biquad *b = new orc_biquad;
    b->s.design(kind, p0, p1, rate);
    b->kind = sample_kind;
    return b;
}
extern "C" void orc_biquad_free(orc_biquad_t *b) { delete b; }
extern "C" void orc_biquad_apply(orc_biquad_t *b, const float *in, size_t n, float *out) {
    if (b->kind == ORC_KIND_F32)
        for (size_t i = 0; i < n; ++i) out[i] = b->s.apply1(in[i]);
    else
        for (size_t i = 0; i < n; ++i) b->s.apply2(in[2 * i], in[2 * i + 1], out[2 * i], out[2 * i + 1]);
}

// =====================================================================================
// a8. Pll -- src/filter/pll.rs
//   design (:48-60): reference /= rate; nphase = 0; value = 0+0i
//   apply (:70-85):
//     c = value_in * self.value.conj()
//     phasedif = loopfilter.apply(c).arg() * gain              arg = im.atan2(re)
//     nphase += reference + phasedif; nphase = nphase.fract()
//     phase = 2.0 * PI * nphase; self.value = from_polar(1.0, phase) = (1.0*cos, 1.0*sin)
//     locked = lockfilter.apply(c.re); output = outputfilter.apply(phasedif * rate)
//     Some(output) iff locked > 0.01
// =====================================================================================
struct orc_pll {
    float rate, reference, gain;
    BiquadState loopf, outf, lockf;
    float nphase, vre, vim;
};
extern "C" orc_pll_t *orc_pll_new(const orc_pll_design_t *d, float rate) {
    orc_pll *p = new orc_pll;
    p->rate = rate;
    p->reference = d->reference / rate;
    p->gain = d->gain;
    p->loopf.design(d->loop_kind, d->loop_p0, d->loop_p1, rate);
    p->outf.design(d->out_kind, d->out_p0, d->out_p1, rate);
    p->lockf.design(d->lock_kind, d->lock_p0, d->lock_p1, rate);
    p->nphase = 0.0f;
    p->vre = 0.0f;
    p->vim = 0.0f;
    return p;
}
extern "C" void orc_pll_free(orc_pll_t *p) { delete p; }
static inline void pll_step(orc_pll *p, float xr, float xi, float *out, uint8_t *locked) {
    // value * conj(self.value): other = (vre, -vim)
    const float or_ = p->vre, oi_ = -p->vim;
    const float cr = xr * or_ - xi * oi_;
    const float ci = xr * oi_ + xi * or_;
    float lr, li;
    p->loopf.apply2(cr, ci, lr, li);
    const float phasedif = atan2f(li, lr) * p->gain;
    p->nphase += p->reference + phasedif;
    p->nphase = rs_fract(p->nphase);
    const float phase = 2.0f * PI_F * p->nphase;
    p->vre = 1.0f * cosf(phase);
    p->vim = 1.0f * sinf(phase);
    const float lk = p->lockf.apply1(cr);
    const float o = p->outf.apply1(phasedif * p->rate);
    *out = o;
    *locked = lk > 0.01f ? 1 : 0;
}
extern "C" void orc_pll_apply(orc_pll_t *p, const float *in, size_t n, float *out, uint8_t *locked) {
    for (size_t i = 0; i < n; ++i) pll_step(p, in[2 * i], in[2 * i + 1], out + i, locked + i);
}
// FM stereo decode closure of src/main.rs:62-71 around a pilot-tone Pll:
//     let mono = v * 0.5;
//     let diff = if let Some(_) = pllpilot.apply(Complex::new(v, 0.0)) {
//         let diffc = v / pllpilot.value.powi(2);  diffc.re * 0.5 } else { 0.0 };
// num-complex 0.2 (not in the tree, restated from the published crate): powi(2) = z * z by repeated
// multiplication (num_traits::pow), f32 / Complex = (a*c/|z|^2, -a*d/|z|^2) with |z|^2 = c*c + d*d.
extern "C" void orc_fm_stereo_decode(orc_pll_t *p, const float *v, size_t n, float *out_mono_diff) {
    for (size_t i = 0; i < n; ++i) {
        float o;
        uint8_t lk;
        pll_step(p, v[i], 0.0f, &o, &lk);
        const float mono = v[i] * 0.5f;
        float diff = 0.0f;
        if (lk) {
            const float zr = p->vre * p->vre - p->vim * p->vim;
            const float zi = p->vre * p->vim + p->vim * p->vre;
            const float ns = zr * zr + zi * zi;
            diff = (v[i] * zr / ns) * 0.5f;
        }
        out_mono_diff[2 * i] = mono;
        out_mono_diff[2 * i + 1] = diff;
    }
}
extern "C" void orc_pll_state(const orc_pll_t *p, float *nphase, float *vre, float *vim) {
    *nphase = p->nphase;
    *vre = p->vre;
    *vim = p->vim;
}

// =====================================================================================
// FreqSweep -- src/signal/sources.rs:133-194
// =====================================================================================
extern "C" size_t orc_freq_sweep(float rate, float df, int warmup, float rs, float re, float *out_freq,
                                 float *out_c64, size_t cap) {
    float dfdt0 = df * df;  // df.powi(2)
    if (rs > re) dfdt0 = -dfdt0;
    const float endt = (re - rs) / dfdt0;
    const float warmupt = warmup ? 1.0f / df : 0.0f;
    // FreqSweep::new(rate, range.start, dfdt, 0.0, warmupt, warmupt+endt, Some(warmupt+endt))
    const float dt = 1.0f / rate;
    float freq = rs;
    float nphase = 0.0f / (2.0f * PI_F);
    size_t fstart = f32_as_usize(rs_round(warmupt * rate));
    size_t fend = f32_as_usize(rs_round((warmupt + endt) * rate));
    size_t length = f32_as_usize(rs_round((warmupt + endt) * rate));
    const size_t total = length;
    size_t w = 0;
    while (length > 0) {
        --length;
        float dfdt = dfdt0;
        if (fstart > 0) { --fstart; dfdt = 0.0f; }
        if (fend > 0) { --fend; } else { dfdt = 0.0f; }
        freq += dt * dfdt;
        nphase += dt * freq;
        nphase = rs_fract(nphase);
        const float phase = 2.0f * PI_F * nphase;
        if (w < cap) {
            if (out_freq) out_freq[w] = freq;
            if (out_c64) { out_c64[2 * w] = 1.0f * cosf(phase); out_c64[2 * w + 1] = 1.0f * sinf(phase); }
        }
        ++w;
    }
    return total;
}

// =====================================================================================
// a5/a6. Resampler: "sdr-src" specification (libsamplerate-shaped; PARITY UNPINNED).
//
// The reference calls C libsamplerate (src/resample.rs:36,61,...), which is not in the tree and
// whose coefficient tables are unavailable offline.  What IS pinned by reference code is the call
// contract (resample.rs:46-67: output_frames = capacity, end_of_input = input.is_empty(),
// returns input_frames_used, sets len = output_frames_gen) and the adaptor loop
// (adapters/resample.rs:38-82).  The arithmetic below follows libsamplerate's published structure:
//   * ZeroOrderHold / Linear: output m sits at input position  P + m*step  (step = 1/ratio, f64);
//     Linear: (float)(x[i-1] + f*(x[i]-x[i-1])) with the float difference taken in f32 and the
//     rest in f64; the frame "before the first" is the first frame itself (libsamplerate's
//     last_value initialisation).
//   * Sinc*: y = rho * sum_j coef(|j - T| * rho * Q) * x[j], rho = min(ratio, 1), coefficient by
//     linear interpolation in a half-table with Q entries per zero crossing, f64 accumulation,
//     left wing far->near then right wing far->near; no group delay (output m at T = m*step); a call that stops at
//     output_frames consumes only the input its outputs and the next output's window need;
//     outputs stop when T + step > total input length once end_of_input was signalled.
// Deliberate divergences from libsamplerate (documented in DESIGN.md): closed-form positions
// P + m*step instead of a running fmod accumulation (identical whenever step is exactly
// representable, e.g. 5.0, 12.5, 3.0); ratio changes apply at call boundaries (no intra-block
// glide); own Kaiser-windowed-sinc tables with libsamplerate's table geometry.
// =====================================================================================
namespace {

struct SincSpec { int increment; size_t half_len; double fc, beta; };
// table geometry follows libsamplerate's three coefficient sets (entries per zero crossing,
// half length); fc / beta are an own Kaiser design for ~97 / 97 / 145 dB.
const SincSpec kSincSpec[3] = {
    {2381, 340239, 0.9666, 15.0},  // best
    {491, 22438, 0.932, 9.73},     // medium
    {128, 2464, 0.84, 9.73},       // fastest
};

double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 500; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    return sum;
}

std::vector<float> g_tables[3];
std::once_flag g_table_once[3];

const std::vector<float> &sinc_table(int type) {
    std::call_once(g_table_once[type], [type]() {
        const SincSpec &s = kSincSpec[type];
        std::vector<float> &t = g_tables[type];
        t.resize(s.half_len + 2);
        const double i0b = bessel_i0(s.beta);
        for (size_t k = 0; k <= s.half_len + 1; ++k) {
            double v = 0.0;
            if (k <= s.half_len) {
                const double u = (double)k / (double)s.increment;  // in zero crossings of the unit sinc
                const double a = M_PI * s.fc * u;
                const double sinc = (k == 0) ? 1.0 : std::sin(a) / a;
                const double r = (double)k / (double)s.half_len;
                const double w = bessel_i0(s.beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
                v = s.fc * sinc * w;
            }
            t[k] = (float)v;
        }
    });
    return g_tables[type];
}

}  // namespace

extern "C" size_t orc_src_sinc_table(int type, const float **table, int *increment) {
    if (type < 0 || type > 2) return 0;
    const std::vector<float> &t = sinc_table(type);
    if (table) *table = t.data();
    if (increment) *increment = kSincSpec[type].increment;
    return kSincSpec[type].half_len;
}

struct orc_src {
    int type, channels;
    double ratio;  // last ratio, 0 = unset
    bool fresh;
    // ZOH / linear
    double pos;
    std::vector<float> last;
    // sinc
    std::vector<float> buf;  // frames kept (interleaved)
    double spos;             // position of next output relative to buf[0]
    long total_in;           // frames ever appended
    long buf_origin;         // absolute index of buf[0]
    bool ended;
};

enum {
    ORC_ERR_NONE = 0, ORC_ERR_MALLOC = 1, ORC_ERR_BAD_STATE = 2, ORC_ERR_BAD_DATA = 3,
    ORC_ERR_BAD_DATA_PTR = 4, ORC_ERR_BAD_SRC_RATIO = 6, ORC_ERR_BAD_CONVERTER = 10,
    ORC_ERR_BAD_CHANNEL_COUNT = 11, ORC_ERR_DATA_OVERLAP = 16,
};

static bool bad_ratio(double r) { return !(r >= 1.0 / 256.0 && r <= 256.0); }

static void src_reset_state(orc_src *s) {
    s->ratio = 0.0;
    s->fresh = true;
    s->pos = -1.0;
    s->last.assign(s->channels, 0.0f);
    s->buf.clear();
    s->spos = 0.0;
    s->total_in = 0;
    s->buf_origin = 0;
    s->ended = false;
}

extern "C" orc_src_t *orc_src_new(int type, int channels, int *error) {
    if (error) *error = ORC_ERR_NONE;
    if (channels < 1) { if (error) *error = ORC_ERR_BAD_CHANNEL_COUNT; return nullptr; }
    if (type < 0 || type > 4) { if (error) *error = ORC_ERR_BAD_CONVERTER; return nullptr; }
    orc_src *s = new orc_src;
    s->type = type;
    s->channels = channels;
    src_reset_state(s);
    return s;
}
extern "C" orc_src_t *orc_src_delete(orc_src_t *s) { delete s; return nullptr; }
extern "C" int orc_src_reset(orc_src_t *s) { if (!s) return ORC_ERR_BAD_STATE; src_reset_state(s); return 0; }
extern "C" orc_src_t *orc_src_clone(const orc_src_t *s, int *error) {
    if (error) *error = 0;
    if (!s) { if (error) *error = ORC_ERR_BAD_STATE; return nullptr; }
    return new orc_src(*s);
}
extern "C" int orc_src_set_ratio(orc_src_t *s, double r) {
    if (!s) return ORC_ERR_BAD_STATE;
    if (bad_ratio(r)) return ORC_ERR_BAD_SRC_RATIO;
    s->ratio = r;
    return 0;
}
extern "C" int orc_src_get_channels(const orc_src_t *s) { return s ? s->channels : -ORC_ERR_BAD_STATE; }
extern "C" long orc_src_history_frames(const orc_src_t *s) {
    if (!s) return -ORC_ERR_BAD_STATE;
    return (s->type == ORC_SRC_ZERO_ORDER_HOLD || s->type == ORC_SRC_LINEAR) ? (s->fresh ? 0 : 1) : (long)(s->buf.size() / s->channels);
}
extern "C" const char *orc_src_strerror(int e) {
    switch (e) {
        case 0: return "No error.";
        case 1: return "Malloc failed.";
        case 2: return "SRC_STATE pointer is NULL.";
        case 3: return "SRC_DATA pointer is NULL.";
        case 4: return "SRC_DATA->data_out or SRC_DATA->data_in is NULL.";
        case 6: return "SRC ratio outside [1/256, 256] range.";
        case 10: return "Bad converter number.";
        case 11: return "Channel count must be >= 1.";
        case 16: return "Input and output data arrays overlap.";
        default: return nullptr;
    }
}

static int process_zoh_linear(orc_src *s, orc_src_data_t *d) {
    const int ch = s->channels;
    const long n = d->input_frames, cap = d->output_frames;
    if (n <= 0) return 0;
    const float *x = d->data_in;
    if (s->fresh) {
        for (int c = 0; c < ch; ++c) s->last[c] = x[c];
        s->fresh = false;
    }
    const double step = 1.0 / s->ratio;
    const double P = s->pos;
    long m = 0;
    while (m < cap) {
        const double Pm = P + (double)m * step;
        const double fl = std::floor(Pm);
        const long i = (long)fl + 1;  // right neighbour
        if (s->type == ORC_SRC_LINEAR) {
            if (!(Pm < (double)(n - 1))) break;
            const double f = Pm - fl;
            for (int c = 0; c < ch; ++c) {
                const float left = (i == 0) ? s->last[c] : x[(i - 1) * ch + c];
                const float right = x[i * ch + c];
                d->data_out[m * ch + c] = (float)((double)left + f * (double)(right - left));
            }
        } else {
            if (!(Pm <= (double)(n - 1))) break;
            for (int c = 0; c < ch; ++c)
                d->data_out[m * ch + c] = (i == 0) ? s->last[c] : x[(i - 1) * ch + c];
        }
        ++m;
    }
    const double Pn = P + (double)m * step;
    long I = (long)std::floor(Pn) + 1;
    long used = std::min(I, n);
    if (used < 0) used = 0;
    s->pos = Pn - (double)used;
    if (used > 0)
        for (int c = 0; c < ch; ++c) s->last[c] = x[(used - 1) * ch + c];
    d->input_frames_used = used;
    d->output_frames_gen = m;
    return 0;
}

static int process_sinc(orc_src *s, orc_src_data_t *d) {
    const int ch = s->channels;
    const long n = d->input_frames > 0 ? d->input_frames : 0, cap = d->output_frames;
    const SincSpec &sp = kSincSpec[s->type];
    const std::vector<float> &tab = sinc_table(s->type);
    const double ratio = s->ratio, step = 1.0 / ratio;
    const double rho = ratio < 1.0 ? ratio : 1.0;
    const double rq = rho * (double)sp.increment;         // table entries per input frame
    const double wing = (double)sp.half_len / rq;         // half width in input frames
    const long wc = (long)std::ceil(wing);
    // all offered input is visible to this call's outputs; how much of it is CONSUMED is decided below
    const long kept = (long)(s->buf.size() / ch);
    s->buf.insert(s->buf.end(), d->data_in, d->data_in + (size_t)n * ch);
    const bool ending = s->ended || d->end_of_input;
    long have = (long)(s->buf.size() / ch);
    const double end_rel = (double)(s->total_in + n - s->buf_origin);  // one past the last real frame
    long m = 0;
    const double P = s->spos;
    while (m < cap) {
        const double T = P + (double)m * step;
        const long i0 = (long)std::floor(T);
        if (ending) {
            if (T + step > end_rel) break;
        } else {
            if (i0 + wc + 1 > have - 1) break;  // lookahead not yet available
        }
        // left wing: j <= i0, far -> near ; right wing: j > i0, far -> near
        const long jl = i0 - wc - 1;
        for (int c = 0; c < ch; ++c) {
            double left = 0.0, right = 0.0;
            for (long j = jl; j <= i0; ++j) {
                if (j < 0 || j >= have) continue;
                const double fi = (T - (double)j) * rq;
                const long k = (long)fi;
                if (k >= (long)sp.half_len) continue;
                const double fr = fi - (double)k;
                const double co = (double)tab[k] + fr * ((double)tab[k + 1] - (double)tab[k]);
                left += co * (double)s->buf[j * ch + c];
            }
            for (long j = i0 + wc + 1; j > i0; --j) {
                if (j < 0 || j >= have) continue;
                const double fi = ((double)j - T) * rq;
                const long k = (long)fi;
                if (k >= (long)sp.half_len) continue;
                const double fr = fi - (double)k;
                const double co = (double)tab[k] + fr * ((double)tab[k + 1] - (double)tab[k]);
                right += co * (double)s->buf[j * ch + c];
            }
            d->data_out[m * ch + c] = (float)(rho * (left + right));
        }
        ++m;
    }
    d->output_frames_gen = m;
    // Input consumption (libsamplerate consumes only what the outputs it generated needed): when the output
    // capacity ended the call, keep just the frames the NEXT output's window reaches (index floor(Pn)+wc+1) and
    // hand the rest back to the caller -- the carried history stays bounded for any ratio.
    long used = n;
    if (m == cap) {
        const long need = (long)std::floor(P + (double)m * step) + wc + 2 - kept;
        used = need < 0 ? 0 : (need > n ? n : need);
    }
    if (used < n) {
        s->buf.resize((size_t)(kept + used) * ch);
        have = kept + used;
    }
    s->total_in += used;
    d->input_frames_used = used;
    if (d->end_of_input && used == n) s->ended = true;
    // rebase: keep wc+2 frames behind the next output position
    const double Pn = P + (double)m * step;
    long drop = (long)std::floor(Pn) - wc - 2;
    if (drop > have) drop = have;
    if (drop > 0) {
        s->buf.erase(s->buf.begin(), s->buf.begin() + (size_t)drop * ch);
        s->buf_origin += drop;
        s->spos = Pn - (double)drop;
    } else {
        s->spos = Pn;
    }
    return 0;
}

extern "C" int orc_src_process(orc_src_t *s, orc_src_data_t *d) {
    if (!s) return ORC_ERR_BAD_STATE;
    if (!d) return ORC_ERR_BAD_DATA;
    if ((d->data_in == nullptr && d->input_frames > 0) || (d->data_out == nullptr && d->output_frames > 0))
        return ORC_ERR_BAD_DATA_PTR;
    if (bad_ratio(d->src_ratio)) return ORC_ERR_BAD_SRC_RATIO;
    if (d->input_frames < 0) d->input_frames = 0;
    if (d->output_frames < 0) d->output_frames = 0;
    d->input_frames_used = 0;
    d->output_frames_gen = 0;
    s->ratio = d->src_ratio;  // applies from this call on (no intra-block glide; see header note)
    if (s->type == ORC_SRC_ZERO_ORDER_HOLD || s->type == ORC_SRC_LINEAR) return process_zoh_linear(s, d);
    return process_sinc(s, d);
}

// signal::Resample -- src/signal/adapters/resample.rs:38-82, driven to exhaustion over a finite input
extern "C" size_t orc_resample_signal(const float *in, size_t n_frames, int ch, int type, double ratio,
                                      float *out, size_t cap) {
    int err = 0;
    orc_src *sr = orc_src_new(type, ch, &err);
    if (!sr) return 0;
    const size_t buffer_size = 4096;  // resample.rs:21
    std::vector<float> buffer;        // frames * ch
    std::vector<float> resampled(buffer_size * ch);
    size_t src_i = 0, w = 0;
    for (;;) {
        // refill (:46-52)
        while (buffer.size() / ch < buffer_size && src_i < n_frames) {
            for (int c = 0; c < ch; ++c) buffer.push_back(in[src_i * ch + c]);
            ++src_i;
        }
        // SampleRate::process (resample.rs:46-67)
        orc_src_data_t d;
        d.data_in = buffer.data();
        d.data_out = resampled.data();
        d.input_frames = (long)(buffer.size() / ch);
        d.output_frames = (long)buffer_size;  // output.capacity()
        d.input_frames_used = d.output_frames_gen = 0;
        d.end_of_input = buffer.empty() ? 1 : 0;
        d.src_ratio = ratio;
        if (orc_src_process(sr, &d) != 0) break;
        const size_t got = (size_t)d.output_frames_gen;
        if (buffer.empty() && got == 0) break;  // (:62-65)
        buffer.erase(buffer.begin(), buffer.begin() + (size_t)d.input_frames_used * ch);  // (:68)
        for (size_t i = 0; i < got; ++i) {
            if (w < cap)
                for (int c = 0; c < ch; ++c) out[w * ch + c] = resampled[i * ch + c];
            ++w;
        }
    }
    orc_src_delete(sr);
    return w;
}

// =====================================================================================
// CPU-baseline helpers
// =====================================================================================
extern "C" size_t orc_fir_u8_mt(const uint8_t *iq, size_t n, const float *taps, size_t K, int taps_complex,
                                size_t wait, int threads, float *out) {
    if (wait == 0) return 0;
    const size_t n_out = n / wait;
    // ranges are cut on multiples of `wait` so the (k+1)*wait-1 indexing is range-invariant
    const size_t groups = n_out;
    parallel_for(groups, threads, [&](size_t glo, size_t ghi, int) {
        if (ghi <= glo) return;
        const size_t lo = glo * wait, hi = ghi * wait;
        const size_t halo = std::min(lo, K > 0 ? K - 1 : 0);
        orc_fir_t *f = orc_fir_new(taps, K, taps_complex, ORC_KIND_C64);
        float x[2], y[2];
        // prime the history with the halo (outputs discarded)
        for (size_t i = lo - halo; i < lo; ++i) {
            orc_unpack_u8iq(iq + 2 * i, 1, x);
            orc_fir_apply(f, x, 1, y);
        }
        size_t ph = 0, o = glo;
        for (size_t i = lo; i < hi; ++i) {
            orc_unpack_u8iq(iq + 2 * i, 1, x);   // RtlTcpSignal::next
            orc_fir_apply(f, x, 1, y);          // signal::Filter::next -> Fir::apply
            if (ph == wait - 1) {               // Decimate::next
                out[2 * o] = y[0];
                out[2 * o + 1] = y[1];
                ++o;
                ph = 0;
            } else {
                ++ph;
            }
        }
        orc_fir_free(f);
    });
    return n_out;
}

extern "C" void orc_channelizer_mt(const float *in, size_t n_ch, size_t n, const float *taps, size_t K,
                                   const orc_pll_design_t *pd, float rate, int threads, float *out,
                                   uint8_t *locked) {
    parallel_for(n_ch, threads, [&](size_t lo, size_t hi, int) {
        float y[2];
        for (size_t c = lo; c < hi; ++c) {
            orc_fir_t *f = orc_fir_new(taps, K, 0, ORC_KIND_C64);
            orc_pll_t *p = orc_pll_new(pd, rate);
            const float *x = in + 2 * n * c;
            for (size_t i = 0; i < n; ++i) {
                orc_fir_apply(f, x + 2 * i, 1, y);
                pll_step(p, y[0], y[1], out + n * c + i, locked + n * c + i);
            }
            orc_pll_free(p);
            orc_fir_free(f);
        }
    });
}

extern "C" int orc_hardware_threads(void) {
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}
