"""Sliding-window spectra: signal.window(d).decimate(fps).map(fft::fft) of
examples/live.rs:30-39 against numpy framing + the oracle's fft::fft."""
import numpy as np
import pytest

import gen
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _reference_windows(x, N, D):
    """Window (adapters/mod.rs:271-303) + Decimate (mod.rs:14-41): kept window j ends at sample (j+1)*D - 1"""
    padded = np.concatenate([np.zeros(N, np.complex64), x.astype(np.complex64)])
    out = []
    j = 0
    while (j + 1) * D - 1 < x.size:
        e = (j + 1) * D - 1
        out.append(padded[e + 1:e + 1 + N])
        j += 1
    return out


@pytest.mark.parametrize("N,D,fmt", [(1000, 10000, "u8iq"), (256, 64, "c64"), (1000, 7, "c64"), (300, 300, "u8iq"),
                                     (4096, 1000, "c64"), (1, 3, "c64")])
def test_window_fft_matches_window_decimate_fft(sdr, N, D, fmt):
    n = {10000: 65000, 64: 5000, 7: 900, 300: 2000, 1000: 9000, 3: 50}[D]
    if fmt == "u8iq":
        iq = gen.tone_noise_u8(n, 3.0e5, 4.0e4, 0.5, 0.1, 21)
        x = O.unpack_u8iq(iq)
        raw = iq
    else:
        x = gen.complex_noise(n, 22)
        raw = x
    wins = _reference_windows(x, N, D)
    wf = sdr.WindowFft(N, D, fmt)
    assert wf.output_count(n) == len(wins) == n // D
    got = wf.process(raw)
    assert got.shape == (len(wins), N)
    for j in (0, len(wins) // 2, len(wins) - 1):
        _, want = O.fft_shifted(wins[j], 3.0e5)
        tol = 1e-5 * max(1.0, np.log2(max(N, 2))) * max(1e-6, np.abs(want).max())
        assert np.abs(got[j] - want).max() <= tol, (j, float(np.abs(got[j] - want).max()), tol)


def test_window_fft_streaming_blocks_equal_one_call(sdr):
    N, D = 1000, 333
    n = 20000
    x = gen.complex_noise(n, 23)
    one = sdr.WindowFft(N, D, "c64").process(x)
    wf = sdr.WindowFft(N, D, "c64")
    parts, pos = [], 0
    for blk in (1, 331, 1, 5000, 17, 999, 2000, 11651):
        parts.append(wf.process(x[pos:pos + blk]))
        pos += blk
    assert pos == n
    cat = np.concatenate(parts)
    assert cat.shape == one.shape == (n // D, N)
    assert np.array_equal(cat.view(np.uint32), one.view(np.uint32))
    wf.reset()
    assert np.array_equal(wf.process(x).view(np.uint32), one.view(np.uint32))


def test_window_fft_contract(sdr):
    with pytest.raises(sdr.SdrError):
        sdr.WindowFft(0, 10)
    with pytest.raises(sdr.SdrError):
        sdr.WindowFft(16, 0)
    wf = sdr.WindowFft(16, 4, "c64")
    assert wf.process(np.zeros(3, np.complex64)).shape == (0, 16)  # no window complete yet
    assert wf.process(np.zeros(1, np.complex64)).shape == (1, 16)


def test_window_spectra_over_a_signal(sdr):
    """the live.rs chain through the host mirror: rtl_tcp bytes -> window(1000 / rate).decimate(fps) -> fft"""
    rate, fps = 300000.0, 30.0
    iq = gen.tone_noise_u8(45000, rate, 4.0e4, 0.5, 0.05, 8)
    sig = sdr.signal.from_u8iq(rate, iq)
    got = list(sdr.signal.window_spectra(sig, 1000.0 / rate, fps, block=12345))
    spectra = np.concatenate([s for _, s in got])
    assert spectra.shape == (45000 // 10000, 1000)
    labels = got[0][0]
    peak = labels[np.argmax(np.abs(spectra[-1]))]
    assert abs(peak - 4.0e4) <= rate / 1000  # the tone's bin
    x = O.unpack_u8iq(iq)
    _, want = O.fft_shifted(_reference_windows(x, 1000, 10000)[2], rate)
    assert np.abs(spectra[2] - want).max() <= 1e-4 * np.abs(want).max()
