"""CPU: pin the oracle.  The reference holds no tests or golden vectors, so the pins are
(a) known-answer tests derived from the reference's semantics, (b) independent numpy / scipy / pure-Python
restatements, (c) the committed golden fixtures (regression)."""
import os

import numpy as np
import pytest

import gen
import oracle_lib as O
import pyref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- a1 unpack: src/rtltcp.rs:160-163 -----------------------------------------------------------
def test_unpack_table_exact():
    u = np.arange(256, dtype=np.uint8)
    z = O.unpack_u8iq(np.stack([u, u[::-1]], 1).ravel())
    assert z[0] == complex(-1.0, 0.9921875) and z[128] == complex(0.0, -0.0078125) and z[255] == complex(0.9921875, -1.0)
    want = (u.astype(np.float64) - 128.0) / 128.0
    assert np.array_equal(z.real.astype(np.float64), want)
    assert np.array_equal(z.imag.astype(np.float64), want[::-1])


# ---- a2 FIR: src/filter/fir.rs:23-32 --------------------------------------------------------------
@pytest.mark.parametrize("K", [1, 7, 64, 255])
def test_fir_impulse_returns_taps(K):
    rng = np.random.default_rng(K)
    taps = rng.standard_normal(K).astype(np.float32)
    x = np.zeros(K + 10, np.complex64)
    x[0] = 1
    y = O.Fir(taps).apply(x)
    assert np.array_equal(y[:K].real, taps) and np.all(y[:K].imag == 0) and np.all(y[K:] == 0)
    ct = (rng.standard_normal(K) + 1j * rng.standard_normal(K)).astype(np.complex64)
    y = O.Fir(ct).apply(x)
    assert np.array_equal(y[:K], ct)


def test_fir_constant_gives_f32_running_sums():
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    y = O.Fir(taps, O.KIND_F32).apply(np.ones(100, np.float32))
    acc, want = np.float32(0), []
    for n in range(100):
        acc = np.float32(0)
        for k in range(min(64, n + 1)):
            acc = np.float32(acc + np.float32(np.float32(1.0) * taps[k]))
        want.append(acc)
    assert np.array_equal(y, np.array(want, np.float32))


def test_fir_sequential_f32_order_matches_python_loop():
    rng = np.random.default_rng(1)
    taps = (rng.standard_normal(9) + 1j * rng.standard_normal(9)).astype(np.complex64)
    x = (rng.standard_normal(40) + 1j * rng.standard_normal(40)).astype(np.complex64)
    y = O.Fir(taps).apply(x)
    f = np.float32
    for n in range(40):
        ar, ai = f(0), f(0)
        for k in range(min(9, n + 1)):
            v, c = x[n - k], taps[k]
            pr = f(f(v.real * c.real) - f(v.imag * c.imag))
            pi = f(f(v.real * c.imag) + f(v.imag * c.real))
            ar, ai = f(ar + pr), f(ai + pi)
        assert y[n] == complex(ar, ai)


def test_fir_matches_numpy_convolve_f64():
    rng = np.random.default_rng(2)
    taps = rng.standard_normal(255).astype(np.float32)
    x = gen.complex_noise(4000, 11)
    truth = np.convolve(x.astype(np.complex128), taps.astype(np.float64))[:4000]
    assert np.abs(O.fir_f64(taps, x) - truth).max() < 1e-12
    assert np.abs(O.Fir(taps).apply(x) - truth).max() / np.abs(truth).max() < 2e-6


def test_fir_streaming_and_clone_and_reset():
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    x = gen.complex_noise(3000, 5)
    f = O.Fir(taps)
    whole = O.Fir(taps).apply(x)
    parts = np.concatenate([f.apply(x[:1]), f.apply(x[1:100]), f.apply(x[100:1017])])
    g = f.clone()
    assert np.array_equal(np.concatenate([parts, f.apply(x[1017:])]), whole)
    assert np.array_equal(g.apply(x[1017:]), whole[1017:])
    f.reset()
    assert np.array_equal(f.apply(x[:50]), whole[:50])


# ---- a4 Decimate: src/signal/adapters/mod.rs:19-37 -------------------------------------------------
def test_decimate_indices_and_counts():
    ramp = np.arange(103, dtype=np.float32)
    for D in (1, 2, 10, 50):
        y, ph = O.decimate(ramp, D)
        assert np.array_equal(y, ramp[D - 1::D]) and len(y) == 103 // D and ph == 103 % D
    a, ph = O.decimate(ramp[:37], 10)
    b, ph = O.decimate(ramp[37:], 10, ph)
    assert np.array_equal(np.concatenate([a, b]), ramp[9::10])
    assert O.decimate_wait(2.4e6, 240e3) == 10 and O.decimate_wait(300000.0, 60.0) == 5000
    assert O.decimate_wait(1.0, 3.0) == 0  # the reference underflows here (adapters/mod.rs:31)


def test_take_skip_block_counts():
    assert O.round_count(1.8e6, 0.1) == 180000
    assert O.round_count(44100.0, 1.0 / 100.0) == 441
    assert O.round_count(1000.0, 0.0005) == 1  # half away from zero
    assert O.block_size(0.1, 1.8e6) == 180000
    assert O.block_size(0.1, 144000.0) == 14400
    assert O.block_size(0.1, 44100.0) == 4410
    assert np.array_equal(O.times(4.0, 0, 4), np.array([0, 0.25, 0.5, 0.75], np.float32))


# ---- a7 FFT: src/fft.rs:3-37 -----------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 16, 60, 256, 1000, 1024, 4096, 14400])
def test_fft_matches_numpy(n):
    x = gen.complex_noise(n, 100 + n)
    ref = np.fft.fft(x.astype(np.complex128))
    scale = max(np.abs(ref).max(), 1e-30)
    assert np.abs(O.dft_f64(x) - ref).max() / scale < 1e-13
    assert np.abs(O.fft_f32(x) - ref).max() / scale < 1e-5 * max(1.0, np.log2(n))


def test_fft_tone_lands_in_shifted_bin_with_sqrtN():
    n, b = 1024, 37
    x = np.exp(2j * np.pi * b * np.arange(n) / n).astype(np.complex64)
    labels, vals = O.fft_shifted(x, 2.048e6)
    k = int(np.argmax(np.abs(vals)))
    assert k == b + n // 2
    assert abs(abs(vals[k]) - np.sqrt(n)) < 1e-3
    others = np.delete(np.abs(vals), k)
    assert others.max() < 1e-3
    want = (np.arange(n) - n // 2).astype(np.float32) * (np.float32(2.048e6) / np.float32(n))
    assert np.array_equal(labels, want)


def test_fft_shifted_is_roll_times_norm_exactly():
    for n in (1000, 1024, 7):
        x = gen.complex_noise(n, 7)
        X = O.fft_f32(x)
        _, vals = O.fft_shifted(x, 1.0)
        norm = np.float32(1.0) / np.sqrt(np.float32(n))
        want = np.roll(X, n // 2)
        assert np.array_equal(vals.real, want.real * norm) and np.array_equal(vals.imag, want.imag * norm)


def test_rfft_length_and_content():
    for n in (14400, 1001, 8):
        x = gen.noise(n, 3).astype(np.float32)
        labels, vals = O.rfft_shifted(x, 144000.0)
        assert len(vals) == n - n // 2 and len(labels) == len(vals)
        ref = np.fft.fft(x.astype(np.float64))[:n - n // 2] / np.sqrt(n)
        assert np.abs(vals - ref).max() / np.abs(ref).max() < 1e-5 * np.log2(n)
        assert labels[0] == 0.0


# ---- a9 biquad / a8 PLL: second implementation in pure Python ------------------------------------
def test_biquad_design_matches_python_restatement():
    for (f, q, rate) in [(80000.0, 0.7, 1.8e6), (20000.0, 0.7, 1.8e6), (200.0, 0.7, 144000.0), (20.0, 0.7, 144000.0)]:
        b = pyref.lowpass(f, q, rate)
        c = O.biquad_design(O.BQ_LOWPASS, f, q, rate)
        assert np.array_equal(c, np.array([b.b0, b.b1, b.b2, b.na1, b.na2], np.float32))
    c = O.biquad_design(O.BQ_LR, 13333.0, 0.0, 44100.0)
    d = np.float32(13333.0) / np.float32(44100.0)
    assert c[0] == d and c[1] == 0 and c[2] == 0 and c[4] == 0 and abs(c[3] - np.exp(-d)) < 1e-7


def test_biquad_matches_scipy_lfilter():
    from scipy.signal import lfilter
    c = O.biquad_design(O.BQ_LOWPASS, 20000.0, 0.7, 1.8e6).astype(np.float64)
    x = gen.noise(5000, 9).astype(np.float32)
    y = O.biquad_apply(O.BQ_LOWPASS, 20000.0, 0.7, 1.8e6, x)
    ref = lfilter(c[:3], [1.0, -c[3], -c[4]], x.astype(np.float64))
    assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-4


def test_freq_sweep_and_pll_match_python_restatement():
    fr, v = O.freq_sweep(1800000.0, 20000.0, True, -200000.0, 200000.0)  # examples/pll.rs:5-8
    fr2, v2 = pyref.freq_sweep(1800000.0, 20000.0, True, -200000.0, 200000.0)
    assert len(fr) == 1890 and np.array_equal(fr, fr2) and np.array_equal(v, v2)
    d = O.pll_design(0.0, 0.035, (O.BQ_LOWPASS, 80000.0, 0.7), (O.BQ_LOWPASS, 20000.0, 0.7), (O.BQ_LOWPASS, 20000.0, 0.7))
    out, lk = O.Pll(d, 1.8e6).apply(v)
    p = pyref.Pll(0.0, 0.035, pyref.CBiquad(lambda: pyref.lowpass(80000.0, 0.7, 1.8e6)),
                  pyref.lowpass(20000.0, 0.7, 1.8e6), pyref.lowpass(20000.0, 0.7, 1.8e6), 1.8e6)
    n = 600
    ref = [p.apply((np.float32(z.real), np.float32(z.imag))) for z in v[:n]]
    assert np.array_equal(out[:n], np.array([r[0] for r in ref], np.float32))
    assert np.array_equal(lk[:n].astype(bool), np.array([r[1] for r in ref]))
    # the PLL tracks the sweep once locked (examples/pll.rs plots exactly this)
    sel = lk.astype(bool) & (np.abs(fr) < 150e3)
    assert sel.sum() > 300 and np.abs(out[sel] - fr[sel]).max() < 30e3


def test_pll_identity_output_filter_and_first_sample():
    d = O.pll_design(19000.0, 0.0002, (O.BQ_LOWPASS, 200.0, 0.7), (O.BQ_IDENTITY, 0, 0), (O.BQ_LOWPASS, 20.0, 0.7))
    p = O.Pll(d, 144000.0)
    out, lk = p.apply(np.ones(4, np.complex64))
    assert out[0] == 0.0 and lk[0] == 0  # value starts at 0+0i, so c = 0 (pll.rs:57-58,71)
    nph, val = p.state()
    assert abs(nph) < 1 and abs(abs(val) - 1.0) < 1e-6


# ---- a5/a6 resampler (own spec; parity with libsamplerate UNPINNED) -------------------------------
def test_src_linear_exact_decimation_and_counts():
    x = np.arange(1000, dtype=np.float32)
    y = O.resample_signal(x, O.SRC_LINEAR, 0.2)
    assert y[0] == 0.0 and np.array_equal(y[1:], x[4::5][:len(y) - 1]) and len(y) == 200
    z = O.resample_signal(x, O.SRC_ZOH, 0.2)
    assert z[0] == 0.0 and np.array_equal(z[1:], x[4::5][:len(z) - 1])


def test_src_linear_interpolates_a_ramp():
    x = np.arange(5000, dtype=np.float32)
    y = O.resample_signal(x, O.SRC_LINEAR, 1.5)
    want = np.maximum(-1 + np.arange(len(y)) / 1.5, 0)
    assert abs(len(y) - 7500) <= 2 and np.abs(y - want).max() < 1e-3


@pytest.mark.parametrize("typ,ratio", [(O.SRC_SINC_FASTEST, 0.2), (O.SRC_SINC_FASTEST, 1.5), (O.SRC_SINC_MEDIUM, 0.08),
                                       (O.SRC_SINC_BEST, 1.0 / 3.0)])
def test_src_sinc_resamples_a_tone(typ, ratio):
    n = 20000
    f = 0.02  # cycles / input sample, well inside the passband for every ratio here
    x = np.exp(2j * np.pi * f * np.arange(n)).astype(np.complex64)
    y = O.resample_signal(x, typ, ratio)
    assert abs(len(y) - int(n * ratio)) <= 2
    m = np.arange(len(y))
    want = np.exp(2j * np.pi * f * m / ratio)
    sel = slice(400, len(y) - 400)
    assert np.abs(y[sel] - want[sel]).max() < 2e-4


def test_src_contract_errors_and_state():
    s = O.SampleRate(O.SRC_LINEAR, 2)
    assert s.h and s.err == 0
    assert O.lib().orc_src_new(9, 1, None) is None
    assert O.lib().orc_src_new(O.SRC_LINEAR, 0, None) is None
    with pytest.raises(RuntimeError):
        s.process(1e-4, np.zeros((4, 2), np.float32), 16)
    used, out = s.process(0.5, np.zeros((0, 2), np.float32), 16)  # end_of_input with nothing buffered
    assert used == 0 and len(out) == 0


@pytest.mark.parametrize("typ", [O.SRC_SINC_FASTEST, O.SRC_SINC_MEDIUM])
@pytest.mark.parametrize("ratio", [2.0, 1.5, 0.2])
def test_src_sinc_consumes_only_needed_input_and_history_stays_bounded(typ, ratio):
    """libsamplerate's src_process consumes only the input its generated outputs needed.  The adaptor's pattern
    (adapters/resample.rs:38-82: 4096 frames in, 4096 frames of output capacity) with ratio > 1 must not let the
    carried history grow; and cutting a stream by output capacity must not change a single output."""
    x = gen.complex_noise(60000, 3).view(np.float32).reshape(-1, 2)
    one = O.SampleRate(typ, 2)
    u, whole = one.process(ratio, x, int(len(x) * ratio) + 64)
    tail = one.process(ratio, x[:0], 8192)[1]
    whole = np.concatenate([whole, tail])
    assert u == len(x)
    s = O.SampleRate(typ, 2)
    pos, outs, hist = 0, [], []
    for _ in range(400):
        chunk = x[pos:pos + 4096]
        used, out = s.process(ratio, chunk, 4096)
        assert 0 <= used <= len(chunk)
        pos += used
        outs.append(out)
        hist.append(s.history_frames())
        if len(chunk) == 0 and len(out) == 0:
            break
    got = np.concatenate(outs)
    assert got.shape == whole.shape
    if ratio in (2.0, 0.2):  # step exactly representable: positions P + m*step do not depend on the cut
        assert np.array_equal(got.view(np.uint32), whole.view(np.uint32))
    else:                    # 1/1.5 is rounded: the re-based position differs in its last bits
        assert np.abs(got - whole).max() < 1e-5
    # bounded: two filter wings + one refill, independent of how long the stream is
    bound = 2 * (O.sinc_table(typ)[2] / O.sinc_table(typ)[1] / min(ratio, 1.0) + 3) + 4096
    assert max(hist) <= bound, (max(hist), bound)
    if ratio > 1.0:
        assert pos == len(x) and max(hist[len(hist) // 2:]) <= max(hist[:len(hist) // 2]) + 1


def test_sinc_tables_are_sane():
    for typ, inc, hl in ((O.SRC_SINC_BEST, 2381, 340239), (O.SRC_SINC_MEDIUM, 491, 22438), (O.SRC_SINC_FASTEST, 128, 2464)):
        t, i, n = O.sinc_table(typ)
        assert i == inc and n == hl and t[0] > 0.8 and abs(t[n]) < 1e-4 and t[n + 1] == 0.0
        dc = t[0] + 2 * t[inc:n + 1:inc].sum()  # unit-rate DC gain
        assert abs(dc - 1.0) < 1e-3


# ---- multi-threaded baseline helpers equal the single-thread result --------------------------------
def test_mt_helpers_are_thread_count_invariant():
    iq = gen.tone_noise_u8(20000, 2.048e6, 300e3, 0.5, 0.1, 77)
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    a = O.fir_u8_mt(iq, taps, 1, 1)
    b = O.fir_u8_mt(iq, taps, 1, 5)
    assert np.array_equal(a, b) and np.array_equal(a, O.Fir(taps).apply(O.unpack_u8iq(iq)))
    c = O.fir_u8_mt(iq, taps, 10, 3)
    assert np.array_equal(c, a[9::10])
    f1 = O.fft_batch_u8(iq[:2 * 16 * 1024], 1024, 1)
    f4 = O.fft_batch_u8(iq[:2 * 16 * 1024], 1024, 4)
    assert np.array_equal(f1, f4)


# ---- committed golden fixtures (regression pins; generated by tests/golden/make_golden.py) ----------
def test_golden_fixtures():
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    taps64, taps255 = gen.lowpass_taps(64, 200e3, 2.048e6), gen.lowpass_taps(255, 100e3, 2.4e6)
    iq = gen.tone_noise_u8(4096, 2.048e6, 300e3, 0.5, 0.1, gen.BASE_SEED + 1)
    assert np.array_equal(g["c1_iq"], iq)
    assert np.array_equal(g["c1_fir64"], O.Fir(taps64).apply(O.unpack_u8iq(iq)))
    assert np.array_equal(g["c2_fft1024"], O.fft_batch_u8(iq, 1024, 1))
    fm = gen.fm_u8(8192, 2.4e6, 75e3, 1e3, 0.05, gen.BASE_SEED + 3)
    y = O.Fir(taps255).apply(O.unpack_u8iq(fm))[9::10]
    assert np.array_equal(g["c3_fir255_dec10"], y)
    assert np.array_equal(g["c3_resampled"], O.resample_signal(y, O.SRC_SINC_FASTEST, 0.2))
    fr, v = O.freq_sweep(1800000.0, 20000.0, True, -200000.0, 200000.0)
    d = O.pll_design(0.0, 0.035, (O.BQ_LOWPASS, 80000.0, 0.7), (O.BQ_LOWPASS, 20000.0, 0.7), (O.BQ_LOWPASS, 20000.0, 0.7))
    out, lk = O.Pll(d, 1.8e6).apply(v)
    assert np.array_equal(g["pll_out"], out) and np.array_equal(g["pll_locked"], lk)
    assert np.array_equal(g["bq_lp80k"], O.biquad_design(O.BQ_LOWPASS, 80000.0, 0.7, 1.8e6))
    pd = O.pll_design(19000.0, 0.0002, (O.BQ_LOWPASS, 200.0, 0.7), (O.BQ_LOWPASS, 20.0, 0.7), (O.BQ_LOWPASS, 20.0, 0.7))
    assert np.array_equal(g["fm_mono_diff"], O.Pll(pd, np.float32(144000.0)).stereo_decode(g["fm_mpx"]))


def test_fm_stereo_decode_against_a_python_restatement():
    """orc_fm_stereo_decode == the closure of src/main.rs:62-71 written out in numpy f32 around Pll::apply one sample at a
    time: mono = v * 0.5; Some(_) => (v / value.powi(2)).re * 0.5 with num-complex's z*z and f32 / Complex formulas."""
    n = 4000
    tt = np.arange(n) / 144000.0
    v = (0.3 * np.sin(2 * np.pi * 1500 * tt) + 0.1 * np.sin(2 * np.pi * 19000 * tt) +
         0.25 * np.sin(2 * np.pi * 400 * tt) * np.sin(2 * np.pi * 38000 * tt)).astype(np.float32)
    pd = O.pll_design(19000.0, 0.0002, (O.BQ_LOWPASS, 200.0, 0.7), (O.BQ_LOWPASS, 20.0, 0.7), (O.BQ_LOWPASS, 20.0, 0.7))
    got = O.Pll(pd, np.float32(144000.0)).stereo_decode(v)
    p = O.Pll(pd, np.float32(144000.0))
    want = np.zeros((n, 2), np.float32)
    f = np.float32
    for i in range(n):
        _, lk = p.apply(np.array([complex(v[i], 0.0)], np.complex64))
        _, val = p.state()
        want[i, 0] = v[i] * f(0.5)
        if lk[0]:
            re, im = f(val.real), f(val.imag)
            zr = f(f(re * re) - f(im * im))
            zi = f(f(re * im) + f(im * re))
            ns = f(f(zr * zr) + f(zi * zi))
            want[i, 1] = f(f(f(v[i] * zr) / ns) * f(0.5))
    assert (want[:, 1] != 0).sum() > 500, "the pilot PLL never locked in the test signal"
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
