// host_mirror_test.cpp -- the reference-style usage of the C++ host mirror (host/sdr.hpp), checked against
// the CPU oracle.  Reads like the reference's examples: source.filter(taps).decimate(r).resample(48e3),
// fft::fft(signal.take(t)), PllDesign::new(...).design(rate).apply(v).
// Built and run by tests/test_gpu_host_mirror.py (needs a GPU: there is no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../oracle/sdr_oracle.h"
#include "../../unnamed-rust-sdr_b200/host/sdr.hpp"

using namespace sdr;

static int fails = 0;
#define EXPECT(cond, ...)                          \
    do {                                           \
        if (!(cond)) {                             \
            ++fails;                               \
            printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            printf(__VA_ARGS__);                   \
            printf("\n");                          \
        }                                          \
    } while (0)

static uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static std::vector<float> lowpass(int n, double fc, double fs) {
    std::vector<double> h(n);
    double sum = 0;
    for (int k = 0; k < n; ++k) {
        double t = k - (n - 1) / 2.0, a = 2 * fc / fs;
        double s = (t == 0) ? 1.0 : std::sin(M_PI * a * t) / (M_PI * a * t);
        h[k] = a * s * (0.54 - 0.46 * std::cos(2 * M_PI * k / (n - 1)));
        sum += h[k];
    }
    std::vector<float> o(n);
    for (int k = 0; k < n; ++k) o[k] = (float)(h[k] / sum);
    return o;
}

int main() {
    // ---- rtl_tcp bytes -> filter(255 taps) -> decimate(240e3) -> resample(48e3)  (BASELINE config C3) ----
    const size_t n = 120000;
    std::vector<uint8_t> raw(2 * n);
    for (size_t i = 0; i < 2 * n; ++i) raw[i] = (uint8_t)(splitmix(i ^ 0x5D12B200) >> 56);
    std::vector<float> taps = lowpass(255, 100e3, 2.4e6);

    auto src = signal::from_u8iq(2.4e6f, raw);
    auto dec = signal::decimate(signal::filter(src, taps), 240e3f);
    EXPECT(dec->rate() == 2.4e6f, "Decimate::rate() must report the upstream rate (adapters/mod.rs:38-40)");
    std::vector<Complex> y = dec->collect(50000);
    EXPECT(y.size() == n / 10, "decimated count %zu", y.size());

    std::vector<float> x(2 * n), full(2 * n);
    orc_unpack_u8iq(raw.data(), n, x.data());
    std::vector<double> truth(2 * n);
    orc_fir_f64(taps.data(), taps.size(), 0, ORC_KIND_C64, x.data(), n, truth.data());
    double maxy = 0, maxe = 0;
    for (size_t j = 0; j < y.size(); ++j) {
        const size_t i = (j + 1) * 10 - 1;
        maxy = std::fmax(maxy, std::hypot(truth[2 * i], truth[2 * i + 1]));
        maxe = std::fmax(maxe, std::hypot(y[j].real() - truth[2 * i], y[j].imag() - truth[2 * i + 1]));
    }
    EXPECT(maxe / maxy < 1e-5, "FIR rel err %g", maxe / maxy);

    // relabel the rate as the harness must, then resample
    auto res = signal::resample_with(signal::from_iter(240e3f, y), resample::ConverterType::SincFastest, 48e3f);
    std::vector<Complex> z = res->collect();
    std::vector<float> zr(2 * (y.size() / 5 + 8192));
    const size_t nz = orc_resample_signal(reinterpret_cast<const float *>(y.data()), y.size(), 2, ORC_SRC_SINC_FASTEST,
                                          (double)48e3f / (double)240e3f, zr.data(), zr.size() / 2);
    EXPECT(z.size() == nz, "resampled count %zu vs oracle %zu", z.size(), nz);
    double rmax = 0;
    for (size_t i = 0; i < std::min(z.size(), nz); ++i)
        rmax = std::fmax(rmax, std::hypot(z[i].real() - zr[2 * i], z[i].imag() - zr[2 * i + 1]));
    EXPECT(rmax < 2e-7, "resample max abs diff %g", rmax);

    // ---- SampleRate contract (resample.rs:33-99) ----
    {
        resample::SampleRate<Complex> sr(resample::ConverterType::Linear);
        EXPECT(sr.channels() == 2, "channels");
        std::vector<Complex> in(100, Complex(1, -1)), out;
        out.reserve(4096);
        size_t used = sr.process(0.5, in, out);
        EXPECT(used == 100 && out.size() == 50, "linear used %zu out %zu", used, out.size());
        bool threw = false;
        try { sr.process(1e-9, in, out); } catch (const resample::Error &e) { threw = (e.code == 6); }
        EXPECT(threw, "BadSrcRatio expected");
        auto c = sr.try_clone();
        sr.reset();
    }

    // ---- fft::fft(signal.take(t))  (examples/live.rs:31-38 uses 1000 points) ----
    {
        std::vector<Complex> s(5000);
        for (size_t i = 0; i < s.size(); ++i) s[i] = Complex(std::cos(2 * M_PI * 0.05 * i), std::sin(2 * M_PI * 0.05 * i));
        auto spec = fft::fft(signal::take(signal::from_iter(300000.0f, s), 1000.0f / 300000.0f));
        EXPECT(spec.size() == 1000, "fft len %zu", spec.size());
        std::vector<float> lab(1000), vals(2000);
        orc_fft_shifted(reinterpret_cast<const float *>(s.data()), 1000, 300000.0f, lab.data(), vals.data());
        double e = 0, m = 0;
        for (size_t i = 0; i < 1000; ++i) {
            EXPECT(spec[i].first == lab[i], "label %zu", i);
            e = std::fmax(e, std::hypot(spec[i].second.real() - vals[2 * i], spec[i].second.imag() - vals[2 * i + 1]));
            m = std::fmax(m, std::hypot(vals[2 * i], vals[2 * i + 1]));
        }
        EXPECT(e / m < 1e-4, "fft rel err %g", e / m);
        std::vector<float> r(14400);
        for (size_t i = 0; i < r.size(); ++i) r[i] = (float)std::sin(2 * M_PI * 0.01 * i);
        auto rs = fft::rfft(signal::from_iter(144000.0f, r));  // examples/fft.rs:64,78
        EXPECT(rs.size() == 7200 && rs[0].first == 0.0f, "rfft len %zu", rs.size());
    }

    // ---- PllDesign::new(...).design(rate).apply(v)  (examples/pll.rs:8-18) ----
    {
        std::vector<float> fr(1890), sw(2 * 1890);
        const size_t ns = orc_freq_sweep(1800000.0f, 20000.0f, 1, -200000.0f, 200000.0f, fr.data(), sw.data(), 1890);
        EXPECT(ns == 1890, "sweep len");
        filter::PllDesign d(0.0f, 0.035f, filter::BiquadD::LowPass(80000.0f, 0.7f), filter::BiquadD::LowPass(20000.0f, 0.7f),
                            filter::BiquadD::LowPass(20000.0f, 0.7f));
        filter::Pll pll = d.design(1800000.0f);
        std::vector<std::optional<float>> out;
        pll.process(reinterpret_cast<const Complex *>(sw.data()), ns, out);
        orc_pll_design_t od{0.0f, 0.035f, ORC_BQ_LOWPASS, 80000.0f, 0.7f, ORC_BQ_LOWPASS, 20000.0f, 0.7f, ORC_BQ_LOWPASS, 20000.0f, 0.7f};
        orc_pll_t *op = orc_pll_new(&od, 1800000.0f);
        std::vector<float> ro(ns);
        std::vector<uint8_t> rl(ns);
        orc_pll_apply(op, sw.data(), ns, ro.data(), rl.data());
        orc_pll_free(op);
        size_t mism = 0;
        double e = 0;
        for (size_t i = 0; i < ns; ++i) {
            if ((bool)out[i] != (bool)rl[i]) { ++mism; continue; }
            if (out[i]) e = std::fmax(e, std::fabs(*out[i] - ro[i]));
        }
        EXPECT(mism <= 3 && e < 1e-3 * 1.8e6 * 0.035 * M_PI, "pll mismatches %zu err %g", mism, e);
        EXPECT(std::fabs(std::abs(pll.value()) - 1.0f) < 1e-6, "pll.value on the unit circle");
    }

    // ---- BiquadD::Lr(..).design(rate) as a stream filter (main.rs:75-80 de-emphasis): bit-identical to the reference order ----
    {
        std::vector<float> x(3000), want(3000);
        for (size_t i = 0; i < x.size(); ++i) x[i] = (float)std::sin(0.013 * i) + 0.25f * (float)std::cos(0.31 * i);
        filter::Biquad<float> bq(filter::BiquadD::Lr(1.0f / 75e-6f), 48000.0f);
        std::vector<float> a, b;
        bq.process(x.data(), 1234, a);
        filter::Biquad<float> bq2(bq);  // clone keeps the state
        bq.process(x.data() + 1234, x.size() - 1234, b);
        std::vector<float> b2;
        bq2.process(x.data() + 1234, x.size() - 1234, b2);
        orc_biquad_t *ob = orc_biquad_new(ORC_BQ_LR, 1.0f / 75e-6f, 0.0f, 48000.0f, ORC_KIND_F32);
        orc_biquad_apply(ob, x.data(), x.size(), want.data());
        orc_biquad_free(ob);
        a.insert(a.end(), b.begin(), b.end());
        EXPECT(std::memcmp(a.data(), want.data(), want.size() * sizeof(float)) == 0, "biquad Lr bit-exact");
        EXPECT(b == b2, "biquad clone");
    }

    // ---- Fir::apply per sample == block (filter/mod.rs:23-26) and clone keeps state ----
    {
        std::vector<float> t8 = lowpass(8, 0.2, 1.0);
        filter::Fir<float, Complex> f(t8, 1, false, SDR_FIR_STRICT_ORDER);
        orc_fir_t *of = orc_fir_new(t8.data(), 8, 0, ORC_KIND_C64);
        for (int i = 0; i < 20; ++i) {
            Complex v((float)i, (float)-i), w;
            Complex g = f.apply(v);
            orc_fir_apply(of, reinterpret_cast<const float *>(&v), 1, reinterpret_cast<float *>(&w));
            EXPECT(g == w, "Fir::apply sample %d", i);
        }
        filter::Fir<float, Complex> f2(f);
        Complex v(1, 2);
        EXPECT(f.apply(v) == f2.apply(v), "clone state");
        orc_fir_free(of);
    }

    // ---- src/main.rs:32-81 as one device-resident pipeline == the same chain hopping through host buffers stage by stage ----
    {
        const size_t n = 180000;  // 0.1 s at 1.8 MS/s
        std::vector<uint8_t> iq(2 * n);
        double ph = 0.0;
        for (size_t i = 0; i < n; ++i) {  // stereo multiplex: L = 1 kHz, R = 3 kHz, 19 kHz pilot, 38 kHz subcarrier
            const double t = (double)i / 1.8e6;
            const double l = std::sin(2 * M_PI * 1000 * t), r = std::sin(2 * M_PI * 3000 * t);
            const double mpx = 0.45 * (l + r) + 0.1 * std::sin(2 * M_PI * 19000 * t) + 0.45 * (l - r) * std::sin(2 * M_PI * 38000 * t);
            ph += 2 * M_PI * 75000.0 * mpx / 1.8e6;
            iq[2 * i] = (uint8_t)std::lround(128 + 90 * std::cos(ph));
            iq[2 * i + 1] = (uint8_t)std::lround(128 + 90 * std::sin(ph));
        }
        app::FmStereo fm(1);
        std::vector<std::pair<float, float>> got;
        const size_t ngot = fm.process(iq.data(), n, true, got);
        EXPECT(fm.rate() == 48000.0f && ngot > 4700 && ngot < 4900, "FmStereo output count %zu", ngot);

        // stage by stage with the mirrored reference types, every intermediate in host memory
        std::vector<Complex> x(n);
        for (size_t i = 0; i < n; ++i) x[i] = Complex(((float)iq[2 * i] - 128.0f) / 128.0f, ((float)iq[2 * i + 1] - 128.0f) / 128.0f);
        filter::Pll demod = filter::PllDesign(0.0f, 0.035f, filter::BiquadD::LowPass(80000.0f, 0.7f), filter::BiquadD::Identity(),
                                              filter::BiquadD::LowPass(20000.0f, 0.7f)).design(1800000.0f);
        std::vector<std::optional<float>> d;
        demod.process(x.data(), n, d);
        std::vector<float> v1(n);
        for (size_t i = 0; i < n; ++i) v1[i] = d[i].value_or(0.0f) / 75000.0f;
        auto run_src = [](auto &sr, double ratio, const auto &in, auto &out) {
            using V = std::decay_t<decltype(out)>;
            V chunk;
            chunk.reserve(in.size());
            sr.process(ratio, in, chunk);
            out = chunk;
            for (;;) {  // flush: empty input = end_of_input (resample.rs:56), until nothing comes back
                V none, more;
                more.reserve(4096);
                sr.process(ratio, none, more);
                if (more.empty()) break;
                out.insert(out.end(), more.begin(), more.end());
            }
        };
        resample::SampleRate<float> sr1(resample::ConverterType::SincFastest);
        std::vector<float> v2;
        run_src(sr1, (double)144000.0f / (double)1800000.0f, v1, v2);
        filter::Pll pilot = filter::PllDesign(19000.0f, 0.0002f, filter::BiquadD::LowPass(200.0f, 0.7f), filter::BiquadD::LowPass(20.0f, 0.7f),
                                              filter::BiquadD::LowPass(20.0f, 0.7f)).design(144000.0f);
        std::vector<std::pair<float, float>> md, md2;
        pilot.stereo_decode(v2.data(), v2.size(), md);
        resample::SampleRate<std::pair<float, float>> sr2(resample::ConverterType::SincBestQuality);
        run_src(sr2, (double)48000.0f / (double)144000.0f, md, md2);
        std::vector<float> mono(md2.size()), diff(md2.size()), mo, di;
        for (size_t i = 0; i < md2.size(); ++i) { mono[i] = md2[i].first; diff[i] = md2[i].second; }
        filter::Biquad<float> dm(filter::BiquadD::Lr(1.0f / (75.0f * 0.001f * 0.001f)), 48000.0f), dd(dm);
        dm.process(mono.data(), mono.size(), mo);
        dd.process(diff.data(), diff.size(), di);
        EXPECT(md2.size() == ngot, "stage-by-stage count %zu vs %zu", md2.size(), ngot);
        size_t bad = 0;
        for (size_t i = 0; i < std::min(ngot, md2.size()); ++i) {
            const float L = mo[i] + di[i], R = mo[i] - di[i];
            if (std::memcmp(&L, &got[i].first, 4) || std::memcmp(&R, &got[i].second, 4)) ++bad;
        }
        EXPECT(bad == 0, "FmStereo vs stage-by-stage chain: %zu frames differ", bad);
    }

    // ---- live.rs:30-39: window(d).decimate(fps).map(fft) in one call == fft::fft of each kept window, bit for bit ----
    {
        const float rate = 300000.0f;
        const size_t n = 35000;
        std::vector<Complex> x(n);
        for (size_t i = 0; i < n; ++i) x[i] = Complex((float)std::cos(0.8 * i) + 0.1f * (float)std::sin(0.013 * i), (float)std::sin(0.8 * i));
        fft::WindowSpectra ws(rate, 1000.0f / rate, 30.0f);  // window 1000, hop 10000
        std::vector<Complex> spectra;
        size_t got = ws.process(x.data(), 12345, spectra);
        got += ws.process(x.data() + 12345, n - 12345, spectra);
        EXPECT(ws.window() == 1000 && got == 3 && spectra.size() == 3000, "WindowSpectra count %zu", got);
        size_t bad = 0;
        for (size_t j = 0; j < got; ++j) {
            const size_t e = (j + 1) * 10000 - 1;  // last sample of kept window j
            std::vector<Complex> w(x.begin() + (e + 1 - 1000), x.begin() + e + 1);
            auto f = fft::fft(signal::from_iter<Complex>(rate, w));
            for (size_t i = 0; i < 1000; ++i)
                if (std::memcmp(&f[i].second, &spectra[j * 1000 + i], sizeof(Complex))) ++bad;
        }
        EXPECT(bad == 0, "WindowSpectra vs fft::fft of each window: %zu values differ", bad);
    }

    printf(fails ? "HOST MIRROR: %d failure(s)\n" : "HOST MIRROR OK\n", fails);
    return fails ? 1 : 0;
}
