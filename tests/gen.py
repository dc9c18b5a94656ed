"""Deterministic synthetic inputs: counter-based splitmix64 so that the CPU oracle
and the GPU consume identical bytes regardless of how the work is cut up."""
import numpy as np

BASE_SEED = 0x5D12B200
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(idx, seed):
    z = (np.asarray(idx, np.uint64) ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform(n, seed, offset=0):
    """U[0,1) from the top 53 bits"""
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    return (splitmix64(idx, seed) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def noise(n, seed, offset=0):
    """sum of 4 uniforms - 2 (sigma ~ 0.577)"""
    s = np.zeros(n)
    for k in range(4):
        s += uniform(n, seed + 0x1000 * (k + 1), offset)
    return s - 2.0


def random_u8(n_bytes, seed, offset=0):
    idx = np.arange(offset, offset + n_bytes, dtype=np.uint64)
    return (splitmix64(idx, seed) >> np.uint64(56)).astype(np.uint8)


def quantise_u8iq(z):
    """complex -> rtl_tcp bytes: clamp(round(128 + 128 v), 0, 255), I then Q"""
    out = np.empty(2 * len(z), np.uint8)
    out[0::2] = np.clip(np.round(128.0 + 128.0 * z.real), 0, 255).astype(np.uint8)
    out[1::2] = np.clip(np.round(128.0 + 128.0 * z.imag), 0, 255).astype(np.uint8)
    return out


def tone_noise_u8(n, fs, f_tone, amp, sigma, seed, offset=0):
    t = np.arange(offset, offset + n, dtype=np.float64) / fs
    z = amp * np.exp(2j * np.pi * f_tone * t)
    z = z + sigma / 0.577 * (noise(n, seed, offset) + 1j * noise(n, seed + 7, offset))
    return quantise_u8iq(z)


def complex_noise(n, seed, offset=0, scale=1.0):
    return (scale * (2 * uniform(n, seed, offset) - 1) + 1j * scale * (2 * uniform(n, seed + 3, offset) - 1)).astype(np.complex64)


def lowpass_taps(n_taps, cutoff, fs):
    """Hamming-windowed sinc, unity DC gain, f32 (the reference has no FIR design function: the
    taps are an input of the harness)"""
    k = np.arange(n_taps, dtype=np.float64) - (n_taps - 1) / 2.0
    h = 2 * cutoff / fs * np.sinc(2 * cutoff / fs * k)
    h *= 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(n_taps) / (n_taps - 1))
    h /= h.sum()
    return h.astype(np.float32)


def complex_bandpass_taps(n_taps, cutoff, shift, fs):
    h = lowpass_taps(n_taps, cutoff, fs).astype(np.float64)
    k = np.arange(n_taps, dtype=np.float64)
    return (h * np.exp(2j * np.pi * shift / fs * k)).astype(np.complex64)


def fm_u8(n, fs, dev, f_audio, sigma, seed, offset=0):
    """FM-modulated carrier at baseband + noise, as rtl_tcp bytes (config C3)"""
    t = np.arange(offset, offset + n, dtype=np.float64) / fs
    phase = dev / f_audio * np.sin(2 * np.pi * f_audio * t)
    z = 0.6 * np.exp(1j * phase) + sigma / 0.577 * (noise(n, seed, offset) + 1j * noise(n, seed + 7, offset))
    return quantise_u8iq(z)


def fm_stereo_u8(n, fs, f_left, f_right, seed, sigma=0.01, dev=75e3, offset=0):
    """A broadcast-FM stereo multiplex (L+R, 19 kHz pilot, L-R on the suppressed 38 kHz subcarrier) frequency-
    modulated onto a baseband carrier, as rtl_tcp bytes: what src/main.rs listens to."""
    t = np.arange(offset, offset + n, dtype=np.float64) / fs
    left = np.sin(2 * np.pi * f_left * t)
    right = np.sin(2 * np.pi * f_right * t)
    mpx = 0.45 * (left + right) + 0.1 * np.sin(2 * np.pi * 19000.0 * t) + \
        0.45 * (left - right) * np.sin(2 * np.pi * 38000.0 * t)
    # phase = 2 pi dev * integral(mpx): closed-form per term keeps chunks (offset) consistent
    def integ_sin(f):
        return -np.cos(2 * np.pi * f * t) / (2 * np.pi * f)
    def integ_sin_sin(fa, fb):  # sin(a) sin(b) = (cos(a-b) - cos(a+b)) / 2
        return 0.5 * (np.sin(2 * np.pi * (fa - fb) * t) / (2 * np.pi * (fa - fb)) -
                      np.sin(2 * np.pi * (fa + fb) * t) / (2 * np.pi * (fa + fb)))
    integral = 0.45 * (integ_sin(f_left) + integ_sin(f_right)) + 0.1 * integ_sin(19000.0) + \
        0.45 * (integ_sin_sin(f_left, 38000.0) - integ_sin_sin(f_right, 38000.0))
    del mpx
    z = 0.7 * np.exp(2j * np.pi * dev * integral) + sigma / 0.577 * (noise(n, seed, offset) + 1j * noise(n, seed + 7, offset))
    return quantise_u8iq(z)
