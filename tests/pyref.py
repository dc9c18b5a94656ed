"""Independent pure-Python (numpy float32 scalar) restatement of the small sequential pieces of the
reference -- used ONLY to cross-check the C++ oracle on small cases (a second implementation written
from the Rust source, not from the C++): Biquad (src/filter/biquad.rs:25-56,83-154),
Pll (src/filter/pll.rs:48-85), FreqSweep (src/signal/sources.rs:133-194)."""
import math

import numpy as np

import ctypes

f32 = np.float32
PI = f32(math.pi)

# numpy's float32 sin/cos/arctan2 are its own SIMD kernels (1 ulp off glibc); the reference calls the
# platform libm through Rust's std, so bind glibc's cosf/sinf/atan2f directly.
_libm = ctypes.CDLL("libm.so.6")
for _n, _na in (("cosf", 1), ("sinf", 1), ("atan2f", 2), ("expf", 1)):
    _f = getattr(_libm, _n)
    _f.restype = ctypes.c_float
    _f.argtypes = [ctypes.c_float] * _na


def cosf(x): return f32(_libm.cosf(float(x)))
def sinf(x): return f32(_libm.sinf(float(x)))
def atan2f(y, x): return f32(_libm.atan2f(float(y), float(x)))


def rs_round(x):
    x = float(x)
    return math.floor(abs(x) + 0.5) * (1 if x >= 0 else -1)


def as_usize(x):
    x = float(x)
    return 0 if not (x > 0) else int(x)


class Biquad:
    def __init__(self, a0, a1, a2, b0, b1, b2, zero=f32(0)):
        a0, a1, a2, b0, b1, b2 = map(f32, (a0, a1, a2, b0, b1, b2))
        self.b0, self.b1, self.b2 = b0 / a0, b1 / a0, b2 / a0
        self.na1, self.na2 = -a1 / a0, -a2 / a0
        self.x1 = self.x2 = self.y1 = self.y2 = zero
        self.zero = zero

    def apply(self, v):
        out = self.zero
        out = out + v * self.b0
        out = out + self.x1 * self.b1
        out = out + self.x2 * self.b2
        out = out + self.y1 * self.na1
        out = out + self.y2 * self.na2
        self.x2, self.x1 = self.x1, v
        self.y2, self.y1 = self.y1, out
        return out


class Identity:
    def apply(self, v):
        return v


def lowpass(freq, q, rate, zero=f32(0)):
    freq, q, rate = f32(freq), f32(q), f32(rate)
    omega = f32(2.0) * PI * freq / rate
    cos = cosf(omega)
    alpha = sinf(omega) / (f32(2.0) * q)
    return Biquad(f32(1) + alpha, f32(-2) * cos, f32(1) - alpha, (f32(1) - cos) / f32(2), f32(1) - cos,
                  (f32(1) - cos) / f32(2), zero)


class CBiquad:
    """Biquad<f32, Complex<f32>>: Complex * f32 = (re*c, im*c), so re and im are two real biquads"""

    def __init__(self, mk):
        self.re, self.im = mk(), mk()

    def apply(self, c):
        return (self.re.apply(c[0]), self.im.apply(c[1]))


class Pll:
    def __init__(self, reference, gain, loopf, outf, lockf, rate):
        self.rate = f32(rate)
        self.reference = f32(reference) / f32(rate)
        self.gain = f32(gain)
        self.loopf, self.outf, self.lockf = loopf, outf, lockf
        self.nphase = f32(0)
        self.value = (f32(0), f32(0))

    def apply(self, v):
        vr, vi = f32(v[0]), f32(v[1])
        o_re, o_im = self.value[0], -self.value[1]
        c = (vr * o_re - vi * o_im, vr * o_im + vi * o_re)
        l = self.loopf.apply(c)
        phasedif = atan2f(l[1], l[0]) * self.gain
        self.nphase = self.nphase + (self.reference + phasedif)
        self.nphase = self.nphase - np.trunc(self.nphase)
        phase = f32(2.0) * PI * self.nphase
        self.value = (f32(1) * cosf(phase), f32(1) * sinf(phase))
        locked = self.lockf.apply(c[0])
        output = self.outf.apply(phasedif * self.rate)
        return output, bool(locked > f32(0.01))


def freq_sweep(rate, df, warmup, start, end):
    rate, df, start, end = f32(rate), f32(df), f32(start), f32(end)
    dfdt0 = df * df
    if start > end:
        dfdt0 = -dfdt0
    endt = (end - start) / dfdt0
    warmupt = f32(1) / df if warmup else f32(0)
    dt = f32(1) / rate
    freq = start
    nphase = f32(0) / (f32(2) * PI)
    fstart = as_usize(rs_round(warmupt * rate))
    fend = as_usize(rs_round((warmupt + endt) * rate))
    length = as_usize(rs_round((warmupt + endt) * rate))
    fr, vals = [], []
    while length > 0:
        length -= 1
        dfdt = dfdt0
        if fstart > 0:
            fstart -= 1
            dfdt = f32(0)
        if fend > 0:
            fend -= 1
        else:
            dfdt = f32(0)
        freq = freq + dt * dfdt
        nphase = nphase + dt * freq
        nphase = nphase - np.trunc(nphase)
        phase = f32(2) * PI * nphase
        fr.append(freq)
        vals.append(complex(f32(1) * cosf(phase), f32(1) * sinf(phase)))
    return np.array(fr, np.float32), np.array(vals, np.complex64)


# ---- signal::Block (src/signal/adapters/block.rs) ---------------------------------------------------
class _Tee:
    """TeeDequeShared (block.rs:7-11): data (newest block at index 0) + one `available` count per reader."""

    def __init__(self):
        self.data = []
        self.available = [0]


class BlockRef:
    """Block::next one sample at a time (block.rs:148-203) with the rayon task run synchronously at the point it is
    spawned (the reference's result does not depend on when the task runs as long as every pop finds its block).
    upstream: a Python iterator of samples.  clone() = Clone for Block (:129-140) + Clone for TeeDeque (:92-103)."""

    def __init__(self, upstream, rate, size, _share=None):
        if _share is not None:
            self.up, self.tee, self.block_size = _share
            self.tee.available.append(len(self.tee.data))  # :95-96
            self.id = len(self.tee.available) - 1
        else:
            self.up = upstream
            self.block_size = as_usize(math.ceil(float(f32(size) * f32(rate))))  # :117
            self.tee = _Tee()
            self.id = 0
        self.current = []
        self.i = 0

    def clone(self):
        return BlockRef(None, 0, 0, _share=(self.up, self.tee, self.block_size))

    def _push(self):  # TeeDequePush::push (:76-89) around the fill closure (:176-187)
        t = self.tee
        if max(t.available) < len(t.data):
            t.data.pop()
        v = []
        for _ in range(self.block_size):
            try:
                v.append(next(self.up))
            except StopIteration:
                pass
        t.data.insert(0, v)
        t.available = [a + 1 for a in t.available]

    def next(self):
        if self.i < len(self.current):      # :149-152
            r = self.current[self.i]
            self.i += 1
            return r
        self.current = []                   # :154-155
        self.i = 0
        needs_extra = True
        t = self.tee
        avail = 0
        if t.available[self.id] > 0:        # try_pop :46-58
            t.available[self.id] -= 1
            avail = t.available[self.id]
            self.current = list(t.data[avail])
            needs_extra = False
        if avail < 1:                       # target = 1 (:165-166)
            jobs = 1 - avail + (1 if needs_extra else 0)
            for _ in range(jobs):
                self._push()
            if needs_extra:                 # pop :60-73
                t.available[self.id] -= 1
                self.current = list(t.data[t.available[self.id]])
        if self.i < len(self.current):      # :197-199 -- note: i is NOT advanced
            return self.current[self.i]
        return None


def pll_trace(reference, gain, loop, out, lock, rate, x):
    """run Pll over x (complex64 array) with LowPass (freq, q) sub-filters; returns (output, locked, arg) where arg is
    the phase detector's atan2 before the gain -- used to find where a trajectory comes near atan2's branch cut"""
    zero2 = f32(0)
    p = Pll(reference, gain, CBiquad(lambda: lowpass(loop[0], loop[1], rate)), lowpass(out[0], out[1], rate),
            lowpass(lock[0], lock[1], rate), rate)
    o = np.empty(len(x), np.float32)
    l = np.empty(len(x), np.uint8)
    a = np.empty(len(x), np.float32)
    for i, v in enumerate(x):
        vr, vi = f32(v.real), f32(v.imag)
        o_re, o_im = p.value[0], -p.value[1]
        c = (vr * o_re - vi * o_im, vr * o_im + vi * o_re)
        lf = p.loopf.apply(c)
        arg = atan2f(lf[1], lf[0])
        a[i] = arg
        phasedif = arg * p.gain
        p.nphase = p.nphase + (p.reference + phasedif)
        p.nphase = p.nphase - np.trunc(p.nphase)
        phase = f32(2.0) * PI * p.nphase
        p.value = (f32(1) * cosf(phase), f32(1) * sinf(phase))
        locked = p.lockf.apply(c[0])
        o[i] = p.outf.apply(phasedif * p.rate)
        l[i] = 1 if locked > f32(0.01) else 0
    return o, l, a
