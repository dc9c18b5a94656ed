"""GPU parity: batched PLL, channelizer, resampler and the adaptor chain through the C ABI vs the oracle."""
import os

import numpy as np
import pytest

import gen
import oracle_lib as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def example_design(sdr):
    B = sdr.BiquadD
    return sdr.PllDesign(0.0, 0.035, B.LowPass(80000.0, 0.7), B.LowPass(20000.0, 0.7), B.LowPass(20000.0, 0.7))


def oracle_design():
    return O.pll_design(0.0, 0.035, (O.BQ_LOWPASS, 80000.0, 0.7), (O.BQ_LOWPASS, 20000.0, 0.7), (O.BQ_LOWPASS, 20000.0, 0.7))


def compare_pll(out, lk, rout, rlk, rate, gain, tol_frac, what):
    """The north star states no PLL tolerance.  The loop feeds back through atan2 / sin / cos, whose last bit
    differs between CUDA and glibc, and atan2 has a branch cut: when the loop-filter output crosses the negative
    real axis a 1-ulp difference flips the phase detector by 2*pi*gain, after which the two trajectories
    re-converge.  Bar used here, relative to the PLL's full-scale output rate*gain*pi:
      - at least 99% of samples within tol_frac, median error < tol_frac / 100,
      - lock flags equal except for isolated samples (the lock filter sitting at the 0.01 threshold)."""
    full = rate * gain * np.pi
    d = np.abs(out.astype(np.float64) - rout.astype(np.float64)) / full
    assert not np.isnan(d).any(), what
    frac_bad = float((d > tol_frac).mean())
    mism = int((lk != rlk).sum())
    assert frac_bad <= 0.01 and np.median(d) < tol_frac / 100, (what, frac_bad, float(np.median(d)), float(d.max()))
    assert mism <= max(2, lk.size // 200), (what, mism)
    return float(d.max()), mism


@pytest.mark.parametrize("fast", [False, True])
def test_pll_example_sweep(sdr, fast):
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    v = g["sweep"]
    p = sdr.PllBatch([example_design(sdr)], 1, 1.8e6, fast_math=fast)
    out, lk = p.process(v)
    # fast = the default f32 atan2 / sincos (~1 ulp of libm), not fast = SDR_PLL_F64_MATH: the same bar for both
    err, mism = compare_pll(out, lk, g["pll_out"], g["pll_locked"], 1.8e6, 0.035, 1e-3, "sweep")
    nph, val = p.state(0)
    assert abs(nph) < 1 and abs(abs(val) - 1) < 1e-6  # f32::fract keeps the sign
    # first sample sees value = 0+0i (pll.rs:57-58)
    assert out[0] == 0.0 and lk[0] == 0


def test_pll_streams_are_independent_and_chunking_is_exact(sdr):
    _, v = O.freq_sweep(1800000.0, 20000.0, True, -200000.0, 200000.0)
    S = 70  # not a multiple of 32
    x = np.stack([np.roll(v, 13 * s) * np.exp(1j * 0.1 * s) for s in range(S)]).astype(np.complex64)
    p = sdr.PllBatch([example_design(sdr)], S, 1.8e6)
    out, lk = p.process(x)
    q = sdr.PllBatch([example_design(sdr)], S, 1.8e6)
    parts = [q.process(np.ascontiguousarray(x[:, a:b])) for a, b in [(0, 1), (1, 33), (33, 1000), (1000, 1890)]]
    assert np.array_equal(np.concatenate([a for a, _ in parts], 1).view(np.uint32), out.view(np.uint32))
    assert np.array_equal(np.concatenate([b for _, b in parts], 1), lk)
    for s in (0, 31, 32, 69):
        ro, rl = O.Pll(oracle_design(), 1.8e6).apply(x[s])
        compare_pll(out[s], lk[s], ro, rl, 1.8e6, 0.035, 1e-3, "stream %d" % s)
    # clone carries state, reset restores the initial one
    c = q.clone()
    a1, _ = q.process(x[:, :100])
    a2, _ = c.process(x[:, :100])
    assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32))
    q.reset()
    a3, _ = q.process(x[:, :100])
    assert np.array_equal(a3.view(np.uint32), out[:, :100].view(np.uint32))


def test_pll_per_stream_designs_and_identity_filters(sdr):
    B = sdr.BiquadD
    d0 = sdr.PllDesign(19000.0, 0.0002, B.LowPass(200.0, 0.7), B.LowPass(20.0, 0.7), B.LowPass(20.0, 0.7))  # main.rs:55-60
    d1 = sdr.PllDesign(0.0, 0.035, B.LowPass(80000.0, 0.7), sdr.Identity(), B.LowPass(20000.0, 0.7))           # main.rs:41-46
    t = np.arange(20000)
    rate = 1.8e6
    x0 = (0.2 * np.cos(2 * np.pi * 19000.0 * t / rate)).astype(np.complex64)
    x1 = np.exp(2j * np.pi * 0.01 * t).astype(np.complex64)
    p = sdr.PllBatch([d0, d1], 2, rate)
    out, lk = p.process(np.stack([x0, x1]))
    o0 = O.pll_design(19000.0, 0.0002, (O.BQ_LOWPASS, 200.0, 0.7), (O.BQ_LOWPASS, 20.0, 0.7), (O.BQ_LOWPASS, 20.0, 0.7))
    o1 = O.pll_design(0.0, 0.035, (O.BQ_LOWPASS, 80000.0, 0.7), (O.BQ_IDENTITY, 0, 0), (O.BQ_LOWPASS, 20000.0, 0.7))
    r0, l0 = O.Pll(o0, rate).apply(x0)
    r1, l1 = O.Pll(o1, rate).apply(x1)
    compare_pll(out[0], lk[0], r0, l0, rate, 0.0002, 2e-2, "pilot")
    compare_pll(out[1], lk[1], r1, l1, rate, 0.035, 1e-3, "demod")


def test_channelizer_small(sdr):
    """C4 in miniature: 40 channels x (255-tap FIR -> PLL)"""
    C_, n = 40, 6000
    taps = gen.lowpass_taps(255, 100e3, 1.8e6)
    t = np.arange(n)
    x = np.stack([np.exp(2j * np.pi * (0.002 * c - 0.03) * t + 0.5j * np.sin(2 * np.pi * 0.001 * t * (1 + c % 3)))
                  for c in range(C_)]).astype(np.complex64)
    x = (x + 0.05 * gen.complex_noise(C_ * n, 4).reshape(C_, n)).astype(np.complex64)
    ch = sdr.Channelizer(taps, example_design(sdr), C_, 1.8e6, strict=True)
    out, lk = ch.process(x)
    rout, rlk = O.channelizer_mt(x, taps, oracle_design(), 1.8e6, threads=4)
    compare_pll(out, lk, rout, rlk, 1.8e6, 0.035, 1e-3, "channelizer")
    # blocks == one call
    ch.reset()
    a = ch.process(np.ascontiguousarray(x[:, :2500]))
    b = ch.process(np.ascontiguousarray(x[:, 2500:]))
    assert np.array_equal(np.concatenate([a[0], b[0]], 1).view(np.uint32), out.view(np.uint32))


@pytest.mark.parametrize("fast", [True, False])
def test_pll_agrees_to_phase_ulps_away_from_the_atan2_branch_cut(sdr, fast):
    """Where the loop-filter output stays away from the negative real axis, a last-bit difference in atan2 / sin /
    cos cannot flip the phase detector, and the loop is contracting: GPU and oracle must then agree to a few ulps of
    the f32 phase accumulator, with identical lock flags.  One ulp of nphase (2^-24 cycles) is 1.2e-7 of full scale
    (rate * gain * pi) at the output, and the loop forgets a perturbation over ~1/gain = 29 samples, so differences in
    the last bit of sin / cos / atan2 random-walk up to ~30 ulps: bar = 1e-5 of full scale for every sample and 1e-6
    for the median (measured on B200: max 2.9e-6 / median 2.2e-7 with the default f32 routines, 3.6e-6 / 1.4e-7 with
    SDR_PLL_F64_MATH: a last-bit difference in arg or in the NCO is far below one ulp of the phase accumulator it is
    added to, so ~1-ulp f32 functions track the oracle as closely as f64 ones).  (a) a locked FM signal (the per-sample restatement in
    tests/pyref.py shows |arg| <= 1.63 rad for all 20 000 samples: no crossing anywhere); (b) the examples/pll.rs
    sweep up to its first sample within 0.25 rad of the cut (sample 11, found with the same restatement)."""
    import pyref
    rate, gain = 1.8e6, 0.035
    full = rate * gain * np.pi
    t = np.arange(20000)
    ph = 2 * np.pi * 30e3 * t / rate + (50e3 / 1e3) * np.sin(2 * np.pi * 1e3 * t / rate)
    x = np.exp(1j * ph).astype(np.complex64)
    ro, rl = O.Pll(oracle_design(), rate).apply(x)
    po, pl_, arg = pyref.pll_trace(0.0, gain, (80000.0, 0.7), (20000.0, 0.7), (20000.0, 0.7), rate, x[:3000])
    assert np.array_equal(po.view(np.uint32), ro[:3000].view(np.uint32))   # the restatement IS the oracle's trajectory
    assert np.abs(arg).max() < np.pi - 0.25
    out, lk = sdr.PllBatch([example_design(sdr)], 1, rate, fast_math=fast).process(x)
    d = np.abs(out.astype(np.float64) - ro) / full
    assert d.max() <= 1e-5 and np.median(d) <= 1e-6, (float(d.max()), float(np.median(d)))
    assert np.array_equal(lk, rl)
    # (b) the sweep's prefix
    _, v = O.freq_sweep(1800000.0, 20000.0, True, -200000.0, 200000.0)
    _, _, arg = pyref.pll_trace(0.0, gain, (80000.0, 0.7), (20000.0, 0.7), (20000.0, 0.7), rate, v[:64])
    first = int(np.argmax(np.abs(arg) > np.pi - 0.25))
    assert first == 11
    so, sl = O.Pll(oracle_design(), rate).apply(v)
    out, lk = sdr.PllBatch([example_design(sdr)], 1, rate, fast_math=fast).process(v)
    d = np.abs(out[:first].astype(np.float64) - so[:first]) / full
    assert d.max() <= 1e-6 and np.array_equal(lk[:first], sl[:first])


@pytest.mark.parametrize("fast", [True, False])
def test_pll_general_and_specialised_kernels_agree_bit_for_bit(sdr, fast):
    """The kernel specialised for designs without Identity sub-filters and with a step per sample below one cycle drops
    the per-sample kind tests and the general trunc of f32::fract; where both apply the two must give the same bits
    (SDR_PLL_GENERAL_KERNEL forces the general one).  A design OUTSIDE the specialised kernel's range (|reference| / rate
    + pi * gain > 1) is checked against the oracle over the prefix before its first approach to the atan2 branch cut."""
    import pyref
    _, v = O.freq_sweep(1800000.0, 20000.0, True, -200000.0, 200000.0)
    x = np.stack([np.roll(v, 7 * s) for s in range(5)]).astype(np.complex64)
    a, la = sdr.PllBatch([example_design(sdr)], 5, 1.8e6, fast_math=fast).process(x)
    b, lb = sdr.PllBatch([example_design(sdr)], 5, 1.8e6, fast_math=fast, general=True).process(x)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(la, lb)
    # wide step: reference 0.494 cycles per sample, gain 0.18 -> |reference| / rate + pi * gain = 1.06: the general kernel,
    # with nphase wrapping through +-1 every other sample; loop filter at 400 kHz keeps this loop well damped (the
    # restatement shows |arg| <= 0.88 rad over the whole signal: nowhere near the cut)
    rate, ref, gain, lbw = 1.8e6, 890000.0, 0.18, 400000.0
    B = sdr.BiquadD
    d_gpu = sdr.PllDesign(ref, gain, B.LowPass(lbw, 0.7), B.LowPass(20000.0, 0.7), B.LowPass(20000.0, 0.7))
    d_cpu = O.pll_design(ref, gain, (O.BQ_LOWPASS, lbw, 0.7), (O.BQ_LOWPASS, 20000.0, 0.7), (O.BQ_LOWPASS, 20000.0, 0.7))
    k = np.arange(1200)
    xs = np.exp(1j * (2 * np.pi * (ref + 300.0) * k / rate + 0.2 * np.sin(2 * np.pi * 1e3 * k / rate))).astype(np.complex64)
    ro, rl = O.Pll(d_cpu, rate).apply(xs)
    _, _, arg = pyref.pll_trace(ref, gain, (lbw, 0.7), (20000.0, 0.7), (20000.0, 0.7), rate, xs)
    assert np.abs(arg).max() < np.pi - 0.25
    out, lk = sdr.PllBatch([d_gpu], 1, rate, fast_math=fast).process(xs)
    d = np.abs(out.astype(np.float64) - ro) / (rate * gain * np.pi)
    assert d.max() <= 1e-5 and np.array_equal(lk, rl), float(d.max())


@pytest.mark.parametrize("n_ch,n", [(128, 8192), (1024, 2048)])
def test_channelizer_default_fir_at_c4_channel_counts(sdr, n_ch, n):
    """C4 as bench.py runs it: the DEFAULT (non-strict) multi-channel FIR feeding the PLLs, at the per-GPU (128) and
    whole-box (1024) channel counts, vs the oracle's per-sample channelizer.  The FIR differs from the oracle's
    sequential f32 sum in the last bits, so the PLL bar applies (see compare_pll)."""
    taps = gen.lowpass_taps(255, 100e3, 1.8e6)
    t = np.arange(n)
    f0 = (0.004 * (np.arange(n_ch) % 16) - 0.03)[:, None]
    x = np.exp(2j * np.pi * f0 * t + 0.5j * np.sin(2 * np.pi * 0.001 * t * (1 + np.arange(n_ch)[:, None] % 3)))
    x = (x + 0.05 * gen.complex_noise(n_ch * n, 4).reshape(n_ch, n)).astype(np.complex64)
    ch = sdr.Channelizer(taps, example_design(sdr), n_ch, 1.8e6)
    out, lk = ch.process(x)
    rout, rlk = O.channelizer_mt(x, taps, oracle_design(), 1.8e6, threads=O.hardware_threads())
    compare_pll(out, lk, rout, rlk, 1.8e6, 0.035, 1e-3, "channelizer %d" % n_ch)
    # the FIR stage alone, same handle type and flags as the channelizer builds: north-star FIR tolerance
    fir = sdr.Fir(taps, "c64", n_channels=n_ch)
    y = fir.process(x)
    sel = [0, n_ch // 2, n_ch - 1]
    for c in sel:
        truth = O.fir_f64(taps, x[c])
        assert np.abs(y[c] - truth).max() / np.abs(truth).max() < 1e-5, c


@pytest.mark.parametrize("typ", ["Linear", "ZeroOrderHold", "SincFastest", "SincMediumQuality", "SincBestQuality"])
@pytest.mark.parametrize("ratio", [0.2, 0.08, 1.0 / 3.0, 1.5, 2.0, 48000.0 / 44100.0])
def test_samplerate_process_matches_oracle(sdr, typ, ratio):
    ct = getattr(sdr.ConverterType, typ)
    x = gen.complex_noise(9000, 12).view(np.float32).reshape(-1, 2)
    a = sdr.SampleRate(ct, 2)
    b = O.SampleRate(int(ct), 2)
    pos = 0
    outs_a, outs_b = [], []
    for blk in (1, 7, 4096, 100, 4096, 0, 0):
        chunk = x[pos:pos + blk]
        ua, oa = a.process(ratio, chunk, 4096)
        ub, ob = b.process(ratio, chunk, 4096)
        assert ua == ub and len(oa) == len(ob), (typ, ratio, blk, ua, ub, len(oa), len(ob))
        outs_a.append(oa)
        outs_b.append(ob)
        pos += ua
    ya, yb = np.concatenate(outs_a), np.concatenate(outs_b)
    assert len(ya) > 0
    if typ in ("Linear", "ZeroOrderHold"):
        assert np.array_equal(ya.view(np.uint32), yb.view(np.uint32))
    else:
        # f64 accumulation in identical order: equal up to the final f32 rounding
        assert np.abs(ya - yb).max() <= 1.2e-7 * max(1.0, np.abs(yb).max())
        assert (ya.view(np.uint32) != yb.view(np.uint32)).mean() < 1e-3


@pytest.mark.parametrize("typ", ["SincBestQuality", "SincMediumQuality", "SincFastest"])
@pytest.mark.parametrize("ratio", [0.2, 1.0 / 3.0, 0.08])
def test_samplerate_long_single_call_matches_oracle(sdr, typ, ratio):
    """One 262 144-frame call -- the size at which the library picks its long-call kernels -- against the ORACLE,
    including SincBestQuality, the reference's default (signal/mod.rs:83).  exact mode (f64 kernels: src_sinc_poly_r_kernel
    for dyadic steps, the per-tap kernel otherwise) equals the specification up to the final f32 rounding; the default
    mode may take the tensor-core polyphase path (integer steps: 0.2 and 1/3 here) and must stay within the north
    star's 1e-5 of max|y| -- bar used: 3e-6."""
    ct = getattr(sdr.ConverterType, typ)
    n = 262144
    x = gen.complex_noise(n, 31).view(np.float32).reshape(-1, 2)
    cap = int(n * ratio) + 64
    b = O.SampleRate(int(ct), 2)
    ub, yb = b.process(ratio, x, cap)
    tb = b.process(ratio, x[:0], 8192)
    yb = np.concatenate([yb, tb[1]])
    for exact in (True, False):
        a = sdr.SampleRate(ct, 2)
        a.set_exact(exact)
        ua, ya = a.process(ratio, x, cap)
        assert ua == ub == n and len(ya) > 0.9 * n * ratio
        ta = a.process(ratio, x[:0], 8192)   # drain (end_of_input)
        assert ta[0] == tb[0] == 0 and len(ta[1]) == len(tb[1])
        ya = np.concatenate([ya, ta[1]])
        assert len(ya) == len(yb)
        if exact:
            assert np.abs(ya - yb).max() <= 1.2e-7 * max(1.0, np.abs(yb).max())
            assert (ya.view(np.uint32) != yb.view(np.uint32)).mean() < 1e-3
        else:
            assert np.abs(ya - yb).max() <= 3e-6 * np.abs(yb).max(), float(np.abs(ya - yb).max() / np.abs(yb).max())


def test_samplerate_tensor_core_path_streaming_and_ratios(sdr):
    """the tensor-core polyphase path across calls (carried history, positions re-based between calls), for the
    reference's two integer-step conversions (240 k -> 48 k: 0.2; 144 k -> 48 k: 1/3) and the three sinc converters;
    launches are counted to prove which path ran (one FIR launch over all S branches + the branch sum)."""
    x = gen.complex_noise(400000, 77).view(np.float32).reshape(-1, 2)
    for typ in (0, 1, 2):
        for ratio, S in ((0.2, 5), (1.0 / 3.0, 3), (0.5, 2), (1.0, 1), (0.1, 10)):
            a, b = sdr.SampleRate(typ, 2), O.SampleRate(typ, 2)
            pos, ya, yb = 0, [], []
            for blk in (100000, 7, 150000, 50000, 99993, 0, 0):
                chunk = x[pos:pos + blk]
                before = sdr.kernel_launch_count()
                ua, oa = a.process(ratio, chunk, 400000)
                launches = sdr.kernel_launch_count() - before
                ub, ob = b.process(ratio, chunk, 400000)
                assert ua == ub and len(oa) == len(ob), (typ, ratio, blk)
                if len(oa) >= 8192:
                    assert launches == (2 if S > 1 else 1), (typ, ratio, blk, launches)
                ya.append(oa)
                yb.append(ob)
                pos += ua
            ya, yb = np.concatenate(ya), np.concatenate(yb)
            assert np.abs(ya - yb).max() <= 3e-6 * np.abs(yb).max(), (typ, ratio, float(np.abs(ya - yb).max() / np.abs(yb).max()))


@pytest.mark.parametrize("typ", ["SincFastest", "SincBestQuality"])
@pytest.mark.parametrize("ratio", [2.0, 1.5])
def test_samplerate_hands_unneeded_input_back_and_history_stays_bounded(sdr, typ, ratio):
    """The adaptor's pattern (adapters/resample.rs:38-82: 4096 frames in, 4096 frames of output capacity) with
    ratio > 1: a call that stops at output_frames consumes only what it needed (libsamplerate's src_process), so the
    carried history stays bounded.  Counts and history equal the oracle's call by call."""
    ct = getattr(sdr.ConverterType, typ)
    x = gen.complex_noise(40000, 3).view(np.float32).reshape(-1, 2)
    a, b = sdr.SampleRate(ct, 2), O.SampleRate(int(ct), 2)
    pos, hist, ya, yb = 0, [], [], []
    for _ in range(200):
        chunk = x[pos:pos + 4096]
        ua, oa = a.process(ratio, chunk, 4096)
        ub, ob = b.process(ratio, chunk, 4096)
        assert ua == ub and len(oa) == len(ob) and a.history_frames() == b.history_frames()
        hist.append(a.history_frames())
        ya.append(oa)
        yb.append(ob)
        pos += ua
        if len(chunk) == 0 and len(oa) == 0:
            break
    assert pos == len(x) and max(hist) <= 4096 + 2 * 160 + 8
    ya, yb = np.concatenate(ya), np.concatenate(yb)
    assert abs(len(ya) - len(x) * ratio) <= 4
    assert np.abs(ya - yb).max() <= 1.2e-7 * max(1.0, np.abs(yb).max())


def test_sinc_polyphase_path_equals_per_tap_kernel_bit_for_bit(sdr):
    """Dyadic steps (ratio 0.2, 0.08, 0.5, 2.0) take the polyphase kernel: per-phase coefficient tables + the same f64
    multiply-add chain.  A second process with SDR_SRC_NO_POLY set runs the per-tap kernel: identical bits, identical
    counts, for every converter including SincBestQuality (the reference's default, signal/mod.rs:83)."""
    import subprocess
    import sys
    import tempfile
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import gen, sdr_b200 as sdr
x = gen.complex_noise(30000, 12).view(np.float32).reshape(-1, 2)
outs = []
for typ in (0, 1, 2):
    for ratio in (0.2, 0.08, 0.5, 2.0, 0.25, 0.0625, 16.0 / 3.0):
        for ch in (1, 2):
            a = sdr.SampleRate(typ, ch)
            a.set_exact(True)
            xx = x if ch == 2 else x[:, :1]
            pos = 0
            for blk in (5, 12000, 4096, 9000, 0, 0):
                u, o = a.process(ratio, xx[pos:pos + blk], 8192)
                pos += u
                outs.append(np.array([u, len(o)], np.float32))
                outs.append(o.ravel())
# one long call: the size at which the library itself picks the R-outputs-per-thread kernel
xl = gen.complex_noise(260000, 13).view(np.float32).reshape(-1, 2)
for typ, ratio in ((0, 0.2), (2, 0.5), (1, 0.25)):
    a = sdr.SampleRate(typ, 2)
    a.set_exact(True)
    u, o = a.process(ratio, xl, 140000)
    outs.append(np.array([u, len(o)], np.float32))
    outs.append(o.ravel())
np.save(sys.argv[1], np.concatenate(outs))
""" % (os.path.join(os.path.dirname(__file__), "..", "unnamed-rust-sdr_b200"), os.path.dirname(__file__))
    res = []
    # SDR_SRC_R forces the R-outputs-per-thread variant that long calls take (the calls here are short)
    for env_extra in ({}, {"SDR_SRC_NO_POLY": "1"}, {"SDR_SRC_R": "5"}, {"SDR_SRC_R": "3"}):
        with tempfile.NamedTemporaryFile(suffix=".npy") as f:
            env = dict(os.environ, **env_extra)
            subprocess.check_call([sys.executable, "-c", code, f.name], env=env)
            res.append(np.load(f.name))
    assert res[0].size > 100000
    for r in res[1:]:
        assert r.shape == res[0].shape
        assert np.array_equal(res[0].view(np.uint32), r.view(np.uint32))


def test_samplerate_contract(sdr):
    s = sdr.SampleRate(sdr.ConverterType.SincBestQuality, 2)
    assert s.get_channels() == 2
    with pytest.raises(sdr.ResampleError) as e:
        s.process(1e-4, np.zeros((4, 2), np.float32), 16)
    assert e.value.code == 6
    used, out = s.process(0.5, np.zeros((0, 2), np.float32), 16)
    assert used == 0 and len(out) == 0
    x = gen.complex_noise(3000, 1).view(np.float32).reshape(-1, 2)
    s.reset()
    u1, o1 = s.process(0.5, x, 4096)
    c = s.try_clone()
    u2, o2 = s.process(0.5, x[:0], 4096)
    u3, o3 = c.process(0.5, x[:0], 4096)
    assert u1 == 3000 and np.array_equal(o2, o3) and len(o1) + len(o2) == 1500
    s.set_ratio(0.25)
    with pytest.raises(sdr.ResampleError):
        s.set_ratio(1e9)
    assert sdr.ConverterType.Linear.name_str() == "Linear Interpolator"


def test_c3_chain_matches_oracle_chain(sdr):
    """config C3: u8 IQ -> 255-tap FIR -> decimate 2.4 MS/s -> 240 kS/s -> (relabel rate) ->
    resample to 48 kHz, through the signal adaptors; vs the oracle's per-sample chain."""
    S = sdr.signal
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    n = 240000
    iq = gen.fm_u8(n, 2.4e6, 75e3, 1e3, 0.05, gen.BASE_SEED + 3)
    assert np.array_equal(iq[:2 * 8192], g["c3_iq"])
    taps = gen.lowpass_taps(255, 100e3, 2.4e6)
    dec = S.from_u8iq(2.4e6, iq).filter(taps).decimate(240e3)
    assert dec.fused and dec.wait == 10
    y = dec.collect(block=50000)
    x = O.unpack_u8iq(iq)
    truth = O.fir_f64(taps, x)[9::10]
    assert len(y) == n // 10
    assert np.abs(y - truth).max() / np.abs(truth).max() < 1e-5
    assert np.abs(y[:len(g["c3_fir255_dec10"])] - g["c3_fir255_dec10"]).max() < 1e-5
    for typ, otyp in ((sdr.ConverterType.SincFastest, O.SRC_SINC_FASTEST), (sdr.ConverterType.Linear, O.SRC_LINEAR)):
        z = S.from_array(240e3, y).resample_with(typ, 48e3).collect()
        ref = O.resample_signal(y, otyp, float(np.float64(np.float32(48e3)) / np.float64(np.float32(240e3))))
        assert len(z) == len(ref) and abs(len(z) - n // 50) <= 2
        assert np.abs(z - ref).max() <= 2e-7 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("strict", [True, False])
def test_c3_device_resident_chain_matches_oracle_chain(sdr, strict):
    """Exactly bench.py's `c3chain` sequence -- Fir.process_dev (u8 IQ -> 255 taps, Decimate 10) into a device buffer,
    SampleRate.process_dev (x0.2, SincBestQuality) out of it, nothing returning to the host in between -- against
    the oracle's Fir::apply + Decimate + SampleRate chain.  Counts are bit-exact; with the reference-order FIR the
    intermediate is bit-identical and the output differs by at most the final f32 rounding; with the default
    (tcgen05) FIR both stay within the north-star 1e-5 of max magnitude."""
    import torch
    dev = torch.device("cuda", 0)
    n = 1 << 20
    iq = gen.fm_u8(n, 2.4e6, 75e3, 1e3, 0.05, gen.BASE_SEED + 3)
    taps = gen.lowpass_taps(255, 100e3, 2.4e6)
    st = torch.cuda.current_stream(dev)
    raw = torch.from_numpy(iq).to(dev)
    n_mid = n // 10 + 1
    mid = torch.zeros(n_mid, dtype=torch.complex64, device=dev)
    fin = torch.zeros(n_mid // 5 + 16, dtype=torch.complex64, device=dev)
    fir = sdr.Fir(taps, "u8iq", decimation=10, strict=strict, stream=st)
    src = sdr.SampleRate(sdr.ConverterType.SincBestQuality, 2, stream=st)
    src.set_exact(strict)   # strict: f64 converter kernels; default: what bench.py's c3chain runs (tensor-core path)
    got_mid = fir.process_dev(raw, n, mid, n_mid)
    used, got_fin = src.process_dev(0.2, mid, got_mid, fin, fin.shape[0])
    torch.cuda.synchronize()
    x = O.unpack_u8iq(iq)
    ref_mid = np.ascontiguousarray(O.Fir(taps).apply(x)[9::10])
    osr = O.SampleRate(O.SRC_SINC_BEST, 2)
    ub, ref_fin = osr.process(0.2, ref_mid.view(np.float32).reshape(-1, 2), fin.shape[0])
    ref_fin = ref_fin.reshape(-1).view(np.complex64)
    assert got_mid == len(ref_mid) == n // 10 and used == ub == got_mid and got_fin == len(ref_fin)
    y_mid = mid[:got_mid].cpu().numpy()
    y_fin = fin[:got_fin].cpu().numpy()
    if strict:
        assert np.array_equal(y_mid.view(np.uint32), ref_mid.view(np.uint32))
        assert np.abs(y_fin - ref_fin).max() <= 1.2e-7 * max(1.0, np.abs(ref_fin).max())
    else:
        truth = O.fir_f64(taps, x)[9::10]
        assert np.abs(y_mid - truth).max() / np.abs(truth).max() < 1e-5
        assert np.abs(y_fin - ref_fin).max() / np.abs(ref_fin).max() < 1e-5


def test_signal_chain_stays_in_hbm_and_equals_the_host_hop_chain(sdr):
    """.filter().decimate().resample() pulled with collect_dev() keeps every intermediate on the
    device (blocks are CUDA tensors between the adaptors) and returns the same samples, bit for bit, as the same
    chain pulled through host numpy blocks -- which test_c3_chain_matches_oracle_chain pins to the oracle."""
    import torch
    S = sdr.signal
    n = 480000
    iq = gen.fm_u8(n, 2.4e6, 75e3, 1e3, 0.05, gen.BASE_SEED + 3)
    taps = gen.lowpass_taps(255, 100e3, 2.4e6)

    def chain(src):
        return src.filter(taps).decimate(240e3)

    host = chain(S.from_u8iq(2.4e6, iq))
    y_host = host.collect(block=50000)
    z_host = S.from_array(240e3, y_host).resample(48e3).collect()
    # device: bytes uploaded once, FIR+decimate -> relabel the rate -> resample, all in HBM
    ctx = S.DeviceCtx(0)
    raw = torch.from_numpy(iq).to(ctx.device)
    before = sdr.kernel_launch_count()
    dec = chain(S.from_device(2.4e6, raw, raw_u8iq=True))
    y_dev = dec.collect_dev(block=50000, ctx=ctx)
    z_dev = S.from_device(240e3, y_dev).resample(48e3).collect_dev(ctx=ctx)
    assert y_dev.is_cuda and z_dev.is_cuda and sdr.kernel_launch_count() > before
    assert np.array_equal(ctx.download(y_dev).view(np.uint32), y_host.view(np.uint32))
    assert np.array_equal(ctx.download(z_dev).view(np.uint32), z_host.view(np.uint32))
    # and against the oracle's adaptor chain (the resampler in 4096-frame chunks, adapters/resample.rs:38-82)
    x = O.unpack_u8iq(iq)
    truth = O.fir_f64(taps, x)[9::10]
    assert np.abs(y_host - truth).max() / np.abs(truth).max() < 1e-5
    ref = O.resample_signal(y_host, O.SRC_SINC_BEST, float(np.float64(np.float32(48e3)) / np.float64(np.float32(240e3))))
    assert len(z_host) == len(ref) and np.abs(z_host - ref).max() <= 2e-7 * max(1.0, np.abs(ref).max())
    # a PLL behind a Block on the device path: same (value, locked) as the host path
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    f_host = S.from_array(1.8e6, g["sweep"]).filter(example_design(sdr))
    o_host = f_host.collect()
    f_dev = S.from_device(1.8e6, torch.from_numpy(g["sweep"]).to(ctx.device)).block(0.0002, dedup=True).filter(example_design(sdr))
    parts, locks = [], []
    while True:
        b = f_dev.next_block_dev(1 << 20, ctx)
        if b.shape[0] == 0:
            break
        parts.append(ctx.download(b))
        locks.append(ctx.download(f_dev.locked))
    assert np.array_equal(np.concatenate(parts).view(np.uint32), o_host.view(np.uint32))
    assert np.array_equal(np.concatenate(locks), f_host.locked)
    # fft::fft of a device-resident signal
    lab, spec = S.fft_dev(S.from_device(240e3, y_dev[:1000]), ctx=ctx)
    lab2, spec2 = sdr.fft(y_host[:1000], 240e3)
    assert np.array_equal(lab, lab2) and np.array_equal(ctx.download(spec).view(np.uint32), spec2.view(np.uint32))


@pytest.mark.parametrize("cplx", [False, True])
def test_biquad_stream_filter_is_bit_identical_to_reference_order(sdr, cplx):
    """Biquad<f32, A> as a stand-alone stream filter (biquad.rs:40-56; the Lr de-emphasis of main.rs:75-80 is one):
    every BiquadD kind, f32 and Complex<f32> samples, ragged blocks, per-stream designs, clone / reset."""
    B = sdr.BiquadD
    rate = 48000.0
    kinds = [(B.LowPass(3000.0, 0.7), (O.BQ_LOWPASS, 3000.0, 0.7)), (B.HighPass(500.0, 0.9), (O.BQ_HIGHPASS, 500.0, 0.9)),
             (B.BandPass(19000.0, 5.0), (O.BQ_BANDPASS, 19000.0, 5.0)), (B.Notch(1000.0, 2.0), (O.BQ_NOTCH, 1000.0, 2.0)),
             (B.Lr(1.0 / 75e-6), (O.BQ_LR, 1.0 / 75e-6, 0.0)), (B.Identity(), (O.BQ_IDENTITY, 0.0, 0.0))]
    n = 5000
    x = gen.complex_noise(n, 33) if cplx else gen.noise(n, 33).astype(np.float32)
    for d, od in kinds:
        want = O.biquad_apply(od[0], od[1], od[2], rate, x)
        f = sdr.Biquad(d, rate, complex_samples=cplx)
        cuts = [0, 1, 2, 33, 1000, 1031, n]
        got = np.concatenate([f.process(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), d.kind
    # one design per stream, 70 streams (more than two warps of sequences), clone mid-stream, reset
    S_ = 70
    designs = [kinds[i % 5][0] for i in range(S_)]
    xs = (gen.complex_noise(S_ * 900, 5).reshape(S_, 900) if cplx else gen.noise(S_ * 900, 5).astype(np.float32).reshape(S_, 900))
    f = sdr.Biquad(designs, rate, n_streams=S_, complex_samples=cplx)
    a = f.process(np.ascontiguousarray(xs[:, :400]))
    g = f.clone()
    b = f.process(np.ascontiguousarray(xs[:, 400:]))
    b2 = g.process(np.ascontiguousarray(xs[:, 400:]))
    got = np.concatenate([a, b], 1)
    for s_ in range(S_):
        od = kinds[s_ % 5][1]
        want = O.biquad_apply(od[0], od[1], od[2], rate, np.ascontiguousarray(xs[s_]))
        assert np.array_equal(got[s_].view(np.uint32), want.view(np.uint32)), s_
    assert np.array_equal(b.view(np.uint32), b2.view(np.uint32))
    f.reset()
    assert np.array_equal(f.process(np.ascontiguousarray(xs[:, :50])).view(np.uint32), got[:, :50].view(np.uint32))


def test_pll_filter_adaptor(sdr):
    """source.filter(PllDesign) (main.rs:49): (value, locked) <-> Option<f32>"""
    S = sdr.signal
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    f = S.from_array(1.8e6, g["sweep"]).filter(example_design(sdr))
    out = f.collect()
    assert len(out) == 1890
    compare_pll(out, f.locked, g["pll_out"], g["pll_locked"], 1.8e6, 0.035, 1e-3, "adaptor")


def test_samplerate_handles_on_two_streams_share_the_constant_table_safely(sdr):
    """The long-call sinc kernel reads its coefficients from one __constant__ table per device.  Two converters of
    different quality on two streams, launched back to back without synchronising in between, must each see their
    own table: results equal the same calls made one at a time."""
    import torch
    dev = torch.device("cuda", 0)
    n_in = 200000
    x = gen.complex_noise(n_in, 21).view(np.float32).reshape(-1, 2)
    d_in = torch.from_numpy(x).to(dev)
    cfgs = [(0, 0.2), (2, 0.5), (1, 0.25), (2, 0.2)]
    # one at a time
    want = []
    for typ, ratio in cfgs:
        s = sdr.SampleRate(typ, 2)
        s.set_exact(True)   # the f64 kernels are the ones that share the __constant__ table
        out = torch.zeros((int(n_in * ratio) + 64, 2), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        used, got = s.process_dev(ratio, d_in, n_in, out, out.shape[0])
        torch.cuda.synchronize()
        want.append((used, got, out[:got].cpu().numpy()))
    # interleaved on separate streams, several rounds, no synchronisation between launches
    streams = [torch.cuda.Stream(dev) for _ in cfgs]
    torch.cuda.synchronize()
    for _ in range(3):
        outs, hs = [], []
        for (typ, ratio), st in zip(cfgs, streams):
            s = sdr.SampleRate(typ, 2, stream=st)
            s.set_exact(True)
            out = torch.zeros((int(n_in * ratio) + 64, 2), dtype=torch.float32, device=dev)
            st.wait_stream(torch.cuda.current_stream(dev))
            used, got = s.process_dev(ratio, d_in, n_in, out, out.shape[0])
            outs.append((used, got, out))
            hs.append(s)
        torch.cuda.synchronize()
        for (used, got, out), (wu, wg, wo) in zip(outs, want):
            assert (used, got) == (wu, wg)
            assert np.array_equal(out[:got].cpu().numpy().view(np.uint32), wo.view(np.uint32))
