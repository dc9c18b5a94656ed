"""Host-side mirror of the reference's stream abstraction (src/signal/mod.rs:13-123) with the
GPU operators behind the adaptors.  The reference pulls one sample at a time (`next()`); a GPU
needs blocks, so every Signal here exposes `next_block(n)` (at most n samples, empty = end of
stream) and the adaptors batch exactly where the reference's own chunked adaptor does
(src/signal/adapters/resample.rs:38-82).  Sample-for-sample the streams are the same.

    signal::from_iter            -> from_array(rate, samples)
    rtltcp::RtlTcpSignal         -> from_u8iq(rate, bytes)       (unpack fused into the consumer)
    Signal::filter(taps)         -> Filter   (GPU Fir; PllDesign -> (value, locked) pairs)
    Signal::decimate(rate)       -> Decimate (fused into a preceding FIR: only kept outputs computed)
    Signal::resample[_with]      -> Resample (SampleRate on the GPU, 4096-frame chunks)
    Signal::take / skip / block / map / iter
"""
import numpy as np

from . import ops
from ._ffi import FMT_C64, FMT_F32, FMT_U8IQ

DEFAULT_BLOCK = 1 << 20


class Signal:
    def rate(self):
        raise NotImplementedError

    def next_block(self, n):
        raise NotImplementedError

    # ---- combinators (src/signal/mod.rs:18-122) ----
    def block(self, size):
        return Block(self, size)

    def decimate(self, rate):
        return Decimate(self, rate)

    def filter(self, design):
        return Filter(self, design)

    def map(self, f):
        return Map(self, f)

    def resample(self, rate):
        return self.resample_with(ops.ConverterType.SincBestQuality, rate)  # mod.rs:83

    def resample_with(self, typ, rate):
        return Resample(self, typ, rate)

    def skip(self, duration):
        return Skip(self, duration)

    def take(self, duration):
        return Take(self, duration)

    def iter(self):
        while True:
            b = self.next_block(DEFAULT_BLOCK)
            if len(b) == 0:
                return
            yield from b

    def collect(self, block=DEFAULT_BLOCK):
        parts = []
        while True:
            b = self.next_block(block)
            if len(b) == 0:
                break
            parts.append(b)
        if not parts:
            return np.empty(0, getattr(self, "dtype", np.complex64))
        return np.concatenate(parts)


class FromArray(Signal):
    """signal::from_iter (src/signal/sources.rs:31-36) over an in-memory array."""

    def __init__(self, rate, samples):
        self._rate = np.float32(rate)
        self.data = np.asarray(samples)
        self.dtype = self.data.dtype
        self.pos = 0

    def rate(self):
        return self._rate

    def next_block(self, n):
        b = self.data[self.pos:self.pos + n]
        self.pos += len(b)
        return b


class FromU8IQ(Signal):
    """rtl_tcp byte stream (src/rtltcp.rs:151-168).  Consumers that understand u8 IQ (Filter, fft)
    take the raw bytes and unpack on the GPU; anything else gets unpacked complex64."""

    def __init__(self, rate, iq_bytes):
        self._rate = np.float32(rate)
        self.raw = np.ascontiguousarray(iq_bytes, np.uint8)
        self.dtype = np.complex64
        self.pos = 0  # in samples

    def rate(self):
        return self._rate

    def next_raw(self, n):
        b = self.raw[2 * self.pos:2 * (self.pos + n)]
        b = b[:len(b) // 2 * 2]
        self.pos += len(b) // 2
        return b

    def next_block(self, n):
        raw = self.next_raw(n)
        if len(raw) == 0:
            return np.empty(0, np.complex64)
        return ops.unpack_u8iq(raw)


def from_array(rate, samples):
    return FromArray(rate, samples)


def from_u8iq(rate, iq_bytes):
    return FromU8IQ(rate, iq_bytes)


class Filter(Signal):
    """signal::Filter (src/signal/adapters/mod.rs:67-100): design_for(&signal) then apply per sample.
    design: taps array (FilterDesign for Vec<C>, fir.rs:44-50) or a PllDesign."""

    def __init__(self, signal, design, decimation=1):
        self.signal = signal
        self.design = design
        self.decimation = decimation
        self.op = None
        if isinstance(design, ops.PllDesign):
            self.kind = "pll"
            self.dtype = np.float32
        else:
            self.kind = "fir"
            taps = np.asarray(design)
            sample_real = np.dtype(getattr(signal, "dtype", np.complex64)).kind == "f"
            self.dtype = np.float32 if sample_real else np.complex64
            self._fmt = "u8iq" if isinstance(signal, FromU8IQ) else ("f32" if sample_real else "c64")
            self._taps = taps
        self.locked = None

    def rate(self):
        return self.signal.rate()

    def _ensure(self):
        if self.op is None:
            if self.kind == "fir":
                self.op = ops.Fir(self._taps, self._fmt, decimation=self.decimation)
            else:
                self.op = self.design.design(float(self.signal.rate()))

    def next_block(self, n):
        self._ensure()
        want = n * self.decimation
        while True:
            if self.kind == "fir" and self._fmt == "u8iq":
                x = self.signal.next_raw(want)
                cnt = len(x) // 2
            else:
                x = self.signal.next_block(want)
                cnt = len(x)
            if cnt == 0:
                return np.empty(0, self.dtype)
            if self.kind == "fir":
                y = self.op.process(x)
                if len(y) == 0:
                    continue  # a short block that produced no kept output: pull again
                return y
            out, locked = self.op.process(x)
            self.locked = locked
            return out


class Decimate(Signal):
    """signal::Decimate (src/signal/adapters/mod.rs:14-41).  wait = (rate_in/rate).round() as usize;
    keeps the last sample of every group; rate() still reports the upstream rate (:38-40)."""

    def __init__(self, signal, rate):
        self.wait = ops.decimate_wait(float(signal.rate()), float(rate))
        if self.wait == 0:
            raise ValueError("Decimate: wait == 0 (the reference underflows at adapters/mod.rs:31)")
        self.dtype = getattr(signal, "dtype", np.complex64)
        if isinstance(signal, Filter) and signal.kind == "fir" and signal.op is None and signal.decimation == 1:
            # fuse: the FIR computes only the kept outputs
            self.signal = Filter(signal.signal, signal.design, decimation=self.wait)
            self.fused = True
        else:
            self.signal = signal
            self.fused = False
        self.phase = 0

    def rate(self):
        return self.signal.rate()

    def next_block(self, n):
        if self.fused:
            return self.signal.next_block(n)
        while True:
            x = self.signal.next_block(n * self.wait)
            if len(x) == 0:
                return x
            first = self.wait - 1 - self.phase
            y = x[first::self.wait]
            self.phase = (self.phase + len(x)) % self.wait
            if len(y):
                return y


class Resample(Signal):
    """signal::Resample (src/signal/adapters/resample.rs:5-86)."""

    def __init__(self, signal, typ, rate):
        self.signal = signal
        self._rate = np.float32(rate)
        self.dtype = getattr(signal, "dtype", np.complex64)
        self.channels = 2 if np.dtype(self.dtype).kind == "c" else 1
        self.sr = ops.SampleRate(typ, self.channels)
        self.ratio = float(np.float64(np.float32(rate)) / np.float64(np.float32(signal.rate())))  # :25
        self.buffer_size = 4096  # :21
        self.buffer = np.empty((0, self.channels), np.float32)
        self.done = False
        self.pending = np.empty(0, self.dtype)

    def rate(self):
        return self._rate

    def _chunk(self):
        """one turn of the `while buffer_next >= buffer_resampled.len()` loop (:44-77)"""
        while True:
            need = self.buffer_size - len(self.buffer)
            if need > 0:
                x = self.signal.next_block(need)
                if len(x):
                    xf = np.ascontiguousarray(x, self.dtype).view(np.float32).reshape(-1, self.channels)
                    self.buffer = np.concatenate([self.buffer, xf])
            used, out = self.sr.process(self.ratio, self.buffer, self.buffer_size)
            if len(self.buffer) == 0 and len(out) == 0:
                self.done = True
                return np.empty(0, self.dtype)
            self.buffer = self.buffer[used:]
            if len(out) == 0:
                continue
            return out.reshape(-1).view(self.dtype) if self.channels == 2 else out.reshape(-1)

    def next_block(self, n):
        if self.done and len(self.pending) == 0:
            return np.empty(0, self.dtype)
        while len(self.pending) == 0 and not self.done:
            self.pending = self._chunk()
        b = self.pending[:n]
        self.pending = self.pending[n:]
        return b


class Take(Signal):
    """signal::Take (adapters/mod.rs:241-268): (rate * duration).round() samples."""

    def __init__(self, signal, duration):
        self.signal = signal
        self.left = ops.duration_samples(float(signal.rate()), float(duration))
        self.dtype = getattr(signal, "dtype", np.complex64)

    def rate(self):
        return self.signal.rate()

    def next_block(self, n):
        n = min(n, self.left)
        if n == 0:
            return np.empty(0, self.dtype)
        b = self.signal.next_block(n)
        self.left -= len(b)
        return b


class Skip(Signal):
    """signal::Skip (adapters/mod.rs:166-194)."""

    def __init__(self, signal, duration):
        self.signal = signal
        self.left = ops.duration_samples(float(signal.rate()), float(duration))
        self.dtype = getattr(signal, "dtype", np.complex64)

    def rate(self):
        return self.signal.rate()

    def next_block(self, n):
        while self.left > 0:
            b = self.signal.next_block(min(self.left, DEFAULT_BLOCK))
            if len(b) == 0:
                return b
            self.left -= len(b)
        return self.signal.next_block(n)


class Block(Signal):
    """signal::Block (adapters/block.rs:106-207): same samples, delivered in blocks of
    ceil(size * rate) (block.rs:117).  The reference uses it to prefetch on a thread pool; here it
    sets the batch size that flows into the GPU operators downstream."""

    def __init__(self, signal, size):
        self.signal = signal
        self.block_size = ops.block_samples(float(size), float(signal.rate()))
        self.dtype = getattr(signal, "dtype", np.complex64)

    def rate(self):
        return self.signal.rate()

    def next_block(self, n):
        return self.signal.next_block(min(n, self.block_size) if self.block_size else n)

    def next_raw(self, n):
        return self.signal.next_raw(min(n, self.block_size) if self.block_size else n)


class Map(Signal):
    """signal::Map (adapters/mod.rs:139-163); f is applied to whole blocks (vectorised)."""

    def __init__(self, signal, f):
        self.signal, self.f = signal, f
        self.dtype = None

    def rate(self):
        return self.signal.rate()

    def next_block(self, n):
        b = self.signal.next_block(n)
        if len(b) == 0:
            return b
        out = self.f(b)
        self.dtype = out.dtype
        return out


def fft(signal):
    """fft::fft(signal) (src/fft.rs:3-28): drains the signal, one N-point transform."""
    if isinstance(signal, FromU8IQ):
        raw = signal.next_raw(1 << 62)
        n = len(raw) // 2
        if n == 0:
            return np.empty(0, np.float32), np.empty(0, np.complex64)
        plan = ops.FftPlan(n, "u8iq", shift=True, norm=True)
        vals = plan.exec(raw)[0]
        plan.close()
        return ops.fft_labels(n, float(signal.rate())), vals
    x = signal.collect()
    return ops.fft(x, float(signal.rate()))


def rfft(signal):
    """fft::rfft(signal) (src/fft.rs:30-37)."""
    return ops.rfft(signal.collect(), float(signal.rate()))


def window_spectra(signal, duration, fps, block=DEFAULT_BLOCK):
    """signal.window(duration).decimate(fps).map(|w| fft::fft(w)) of examples/live.rs:30-39 as a generator of
    (labels, spectra[n_windows, window]) per input block.  window = round(duration * rate) (adapters/mod.rs:279),
    hop = Decimate's wait (mod.rs:22).  Raw u8 IQ sources are unpacked on the device."""
    rate = float(signal.rate())
    window = ops.duration_samples(rate, duration)
    hop = ops.decimate_wait(rate, fps)
    raw = hasattr(signal, "next_raw")
    wf = ops.WindowFft(window, hop, "u8iq" if raw else "c64")
    labels = ops.fft_labels(window, rate)
    try:
        while True:
            b = signal.next_raw(block) if raw else signal.next_block(block)
            if len(b) == 0:
                return
            s = wf.process(b)
            if len(s):
                yield labels, s
    finally:
        wf.close()
