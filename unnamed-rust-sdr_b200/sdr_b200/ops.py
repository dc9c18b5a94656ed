"""Operator handles over the C ABI.  Names follow the reference crate:

    filter::Fir          (src/filter/fir.rs)           -> Fir
    filter::PllDesign    (src/filter/pll.rs)           -> PllDesign / PllBatch
    filter::BiquadD      (src/filter/biquad.rs)        -> BiquadD
    resample::SampleRate (src/resample.rs)             -> SampleRate, ConverterType
    fft::fft / rfft      (src/fft.rs)                  -> FftPlan, fft(), rfft()

Host entry points take numpy arrays; `*_dev` entry points take anything with .data_ptr()
(torch CUDA tensors) or raw integer device addresses and run asynchronously on the handle's stream.
"""
import ctypes as C
import enum

import numpy as np

from . import _ffi as F
from ._ffi import SdrError, check, lib


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return int(x)


def _stream_ptr(stream):
    if stream is None:
        return None
    h = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
    # NULL means "create your own stream" in the C ABI; the legacy default stream is cudaStreamLegacy (0x1)
    return h if h else 1


def device_count():
    return lib().sdr_device_count()


def device_info(dev=0):
    buf = C.create_string_buffer(256)
    check(lib().sdr_device_info(dev, buf, 256), "sdr_device_info")
    return buf.value.decode()


def kernel_launch_count():
    return int(lib().sdr_kernel_launch_count())


def unpack_u8iq(iq, device=0):
    """RtlTcpSignal::next over a whole buffer (src/rtltcp.rs:158-164)."""
    iq = np.ascontiguousarray(iq, np.uint8)
    n = iq.size // 2
    out = np.empty(n, np.complex64)
    check(lib().sdr_unpack_u8iq(iq.ctypes.data, n, out.ctypes.data, device), "sdr_unpack_u8iq")
    return out


def unpack_u8iq_dev(d_iq, n, d_out, device=0, stream=None):
    """device-resident unpack: d_iq 2n bytes, d_out n complex64, asynchronous on `stream`"""
    check(lib().sdr_unpack_u8iq_dev(_ptr(d_iq), n, _ptr(d_out), device, _stream_ptr(stream)), "sdr_unpack_u8iq_dev")


def decimate_wait(rate_in, rate_out):
    return lib().sdr_decimate_wait(rate_in, rate_out)


def duration_samples(rate, duration):
    return lib().sdr_duration_samples(rate, duration)


def block_samples(size, rate):
    return lib().sdr_block_samples(size, rate)


_FMT_OF = {"u8iq": F.FMT_U8IQ, "c64": F.FMT_C64, "f32": F.FMT_F32}
_IN_DTYPE = {F.FMT_U8IQ: np.uint8, F.FMT_C64: np.complex64, F.FMT_F32: np.float32}


class Fir:
    """filter::Fir<C,A> with the signal::Filter (+ optional signal::Decimate) adaptor behaviour:
    a stateful stream filter.  taps float32 -> Fir<f32,A>; complex64 -> Fir<Complex<f32>,Complex<f32>>."""

    def __init__(self, taps, input_format="c64", decimation=1, n_channels=1, strict=False, device=0,
                 stream=None, flags=0, _handle=None, _meta=None):
        if _handle is not None:
            self.h = _handle
            self.__dict__.update(_meta)
            return
        taps = np.asarray(taps)
        self.taps_complex = int(np.iscomplexobj(taps))
        self._taps = np.ascontiguousarray(taps, np.complex64 if self.taps_complex else np.float32)
        self.fmt = _FMT_OF[input_format] if isinstance(input_format, str) else int(input_format)
        self.decimation = int(decimation)
        self.n_channels = int(n_channels)
        self.device = device
        flags = int(flags) | (F.FIR_STRICT_ORDER if strict else 0)
        cfg = F.FirConfig(self._taps.ctypes.data, self._taps.size, self.taps_complex, self.fmt, self.decimation,
                          self.n_channels, flags, device, _stream_ptr(stream))
        err = C.c_int(0)
        self.h = lib().sdr_fir_create(C.byref(cfg), C.byref(err))
        if not self.h:
            raise SdrError(err.value, "sdr_fir_create")

    @property
    def out_dtype(self):
        return np.float32 if self.fmt == F.FMT_F32 else np.complex64

    def output_count(self, n_in):
        return lib().sdr_fir_output_count(self.h, n_in)

    def process(self, x):
        """x: [n] or [n_channels, n] (u8 input: trailing dim 2n bytes).  Returns outputs for this block."""
        x = np.ascontiguousarray(x, _IN_DTYPE[self.fmt])
        per = 2 if self.fmt == F.FMT_U8IQ else 1
        rows = x.reshape(self.n_channels, -1)
        n_in = rows.shape[1] // per
        n_out = self.output_count(n_in)
        out = np.empty((self.n_channels, n_out), self.out_dtype)
        used, got = C.c_size_t(0), C.c_size_t(0)
        check(lib().sdr_fir_process(self.h, rows.ctypes.data, n_in, n_in, out.ctypes.data, n_out, n_out,
                                    C.byref(used), C.byref(got)), "sdr_fir_process")
        assert got.value == n_out and (used.value == n_in or n_in == 0)
        return out[0] if (self.n_channels == 1 and x.ndim <= 1) else out

    def process_dev(self, d_in, n_in, d_out, out_cap, in_stride=None, out_stride=None):
        used, got = C.c_size_t(0), C.c_size_t(0)
        check(lib().sdr_fir_process_dev(self.h, _ptr(d_in), n_in, in_stride or n_in, _ptr(d_out), out_cap,
                                        out_stride or out_cap, C.byref(used), C.byref(got)), "sdr_fir_process_dev")
        return got.value

    def reset(self):
        check(lib().sdr_fir_reset(self.h), "sdr_fir_reset")

    def clone(self):
        err = C.c_int(0)
        h = lib().sdr_fir_clone(self.h, C.byref(err))
        if not h:
            raise SdrError(err.value, "sdr_fir_clone")
        meta = {k: v for k, v in self.__dict__.items() if k != "h"}
        return Fir(None, _handle=h, _meta=meta)

    @property
    def last_path(self):
        return lib().sdr_fir_last_path(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib().sdr_fir_destroy(self.h)
            self.h = None

    __del__ = close


class FftPlan:
    """Batched forward FFT with the post-processing of fft::fft (src/fft.rs:14-26) selectable by flags."""

    def __init__(self, n, input_format="c64", shift=False, norm=False, rfft=False, device=0, stream=None):
        self.n = int(n)
        self.fmt = _FMT_OF[input_format] if isinstance(input_format, str) else int(input_format)
        self.flags = (F.FFT_SHIFT if shift else 0) | (F.FFT_NORM if norm else 0) | (F.FFT_RFFT if rfft else 0)
        cfg = F.FftConfig(self.n, self.fmt, self.flags, device, _stream_ptr(stream))
        err = C.c_int(0)
        self.h = lib().sdr_fft_create(C.byref(cfg), C.byref(err))
        if not self.h:
            raise SdrError(err.value, "sdr_fft_create")
        self.out_len = lib().sdr_fft_output_len(self.h)

    def exec(self, x):
        x = np.ascontiguousarray(x, _IN_DTYPE[self.fmt])
        per = 2 if self.fmt == F.FMT_U8IQ else 1
        batches = x.size // (per * self.n)
        out = np.empty((batches, self.out_len), np.complex64)
        check(lib().sdr_fft_exec(self.h, x.ctypes.data, batches, out.ctypes.data), "sdr_fft_exec")
        return out

    def exec_dev(self, d_in, batches, d_out):
        check(lib().sdr_fft_exec_dev(self.h, _ptr(d_in), batches, _ptr(d_out)), "sdr_fft_exec_dev")

    def close(self):
        if getattr(self, "h", None):
            lib().sdr_fft_destroy(self.h)
            self.h = None

    __del__ = close


def fft_labels(n, rate, rfft=False):
    out = np.empty(n - n // 2 if rfft else n, np.float32)
    check(lib().sdr_fft_labels(n, rate, int(rfft), out.ctypes.data), "sdr_fft_labels")
    return out


def fft(samples, rate=1.0, device=0):
    """fft::fft (src/fft.rs:3-28): one transform over the whole finite signal.
    Returns (labels f32[N], values complex64[N]) -- the reference's Vec<(f32, Complex<f32>)> unzipped."""
    x = np.ascontiguousarray(samples, np.complex64)
    if x.size == 0:
        return np.empty(0, np.float32), np.empty(0, np.complex64)
    plan = FftPlan(x.size, "c64", shift=True, norm=True, device=device)
    vals = plan.exec(x)[0]
    plan.close()
    return fft_labels(x.size, rate), vals


def rfft(samples, rate=1.0, device=0):
    """fft::rfft (src/fft.rs:30-37)."""
    x = np.ascontiguousarray(samples, np.float32)
    if x.size == 0:
        return np.empty(0, np.float32), np.empty(0, np.complex64)
    plan = FftPlan(x.size, "f32", shift=True, norm=True, rfft=True, device=device)
    vals = plan.exec(x)[0]
    plan.close()
    return fft_labels(x.size, rate, rfft=True), vals


class BiquadD:
    """filter::BiquadD (src/filter/biquad.rs:74-81) and filter::Identity, as (kind, p0, p1)."""

    def __init__(self, kind, p0=0.0, p1=0.0):
        self.kind, self.p0, self.p1 = kind, p0, p1

    @staticmethod
    def LowPass(freq, q): return BiquadD(F.BQ_LOWPASS, freq, q)
    @staticmethod
    def HighPass(freq, q): return BiquadD(F.BQ_HIGHPASS, freq, q)
    @staticmethod
    def BandPass(freq, q): return BiquadD(F.BQ_BANDPASS, freq, q)
    @staticmethod
    def Notch(freq, q): return BiquadD(F.BQ_NOTCH, freq, q)
    @staticmethod
    def Lr(decayrate): return BiquadD(F.BQ_LR, decayrate, 0.0)
    @staticmethod
    def Identity(): return BiquadD(F.BQ_IDENTITY)

    def _c(self):
        return F.BiquadDesign(self.kind, self.p0, self.p1)

    def coefficients(self, rate):
        out = np.empty(5, np.float32)
        d = self._c()
        check(lib().sdr_biquad_design(C.byref(d), rate, out.ctypes.data), "sdr_biquad_design")
        return out


Identity = BiquadD.Identity


class Biquad:
    """filter::Biquad<f32, A> as a stream filter (biquad.rs:40-56): n_streams independent streams of f32 or
    Complex<f32> samples; designs = one BiquadD shared by all streams, or one per stream."""

    def __init__(self, designs, rate, n_streams=1, complex_samples=False, device=0, stream=None, _handle=None, _meta=None):
        if _handle is not None:
            self.h = _handle
            self.__dict__.update(_meta)
            return
        if isinstance(designs, BiquadD):
            designs = [designs]
        self.n_streams, self.complex_samples = int(n_streams), bool(complex_samples)
        arr = (F.BiquadDesign * len(designs))(*[d._c() for d in designs])
        self._arr = arr
        cfg = F.BiquadConfig(C.cast(arr, C.c_void_p), len(designs), self.n_streams, rate, int(self.complex_samples),
                             device, _stream_ptr(stream))
        err = C.c_int(0)
        self.h = lib().sdr_biquad_create(C.byref(cfg), C.byref(err))
        if not self.h:
            raise SdrError(err.value, "sdr_biquad_create")

    def process(self, x):
        dt = np.complex64 if self.complex_samples else np.float32
        x = np.ascontiguousarray(x, dt)
        rows = x.reshape(self.n_streams, -1)
        n = rows.shape[1]
        out = np.empty_like(rows)
        check(lib().sdr_biquad_process(self.h, rows.ctypes.data, n, n, out.ctypes.data, n), "sdr_biquad_process")
        return out[0] if x.ndim <= 1 else out

    def process_dev(self, d_in, n, d_out, in_stride=None, out_stride=None):
        check(lib().sdr_biquad_process_dev(self.h, _ptr(d_in), n, in_stride or n, _ptr(d_out), out_stride or n),
              "sdr_biquad_process_dev")

    def reset(self):
        check(lib().sdr_biquad_reset(self.h), "sdr_biquad_reset")

    def clone(self):
        err = C.c_int(0)
        h = lib().sdr_biquad_clone(self.h, C.byref(err))
        if not h:
            raise SdrError(err.value, "sdr_biquad_clone")
        return Biquad(None, 0, _handle=h, _meta={k: v for k, v in self.__dict__.items() if k not in ("h", "_arr")})

    def close(self):
        if getattr(self, "h", None):
            lib().sdr_biquad_destroy(self.h)
            self.h = None

    __del__ = close


class PllDesign:
    """filter::PllDesign::new(reference, gain, loopfilter, outputfilter, lockfilter) (pll.rs:26-36)."""

    def __init__(self, reference, gain, loopfilter, outputfilter, lockfilter):
        self.reference, self.gain = reference, gain
        self.loopfilter, self.outputfilter, self.lockfilter = loopfilter, outputfilter, lockfilter

    def _c(self):
        return F.PllDesign(self.reference, self.gain, self.loopfilter._c(), self.outputfilter._c(), self.lockfilter._c())

    def design(self, rate, n_streams=1, **kw):
        return PllBatch([self], n_streams, rate, **kw)


class PllBatch:
    """n_streams independent filter::Pll instances (pll.rs:13-85)."""

    def __init__(self, designs, n_streams, rate, fast_math=True, device=0, stream=None, _handle=None, general=False):
        self.n_streams = int(n_streams)
        if _handle is not None:
            self.h = _handle
            return
        arr = (F.PllDesign * len(designs))(*[d._c() for d in designs])
        self._arr = arr
        cfg = F.PllConfig(C.cast(arr, C.c_void_p), len(designs), self.n_streams, rate,
                          (0 if fast_math else F.PLL_F64_MATH) | (F.PLL_GENERAL_KERNEL if general else 0), device,
                          _stream_ptr(stream))
        self._cfg = cfg
        err = C.c_int(0)
        self.h = lib().sdr_pll_create(C.byref(cfg), C.byref(err))
        if not self.h:
            raise SdrError(err.value, "sdr_pll_create")

    def process(self, x):
        """x: complex64 [n] or [n_streams, n] -> (out f32, locked u8) of the same shape."""
        x = np.ascontiguousarray(x, np.complex64)
        rows = x.reshape(self.n_streams, -1)
        n = rows.shape[1]
        out = np.empty((self.n_streams, n), np.float32)
        locked = np.empty((self.n_streams, n), np.uint8)
        check(lib().sdr_pll_process(self.h, rows.ctypes.data, n, n, out.ctypes.data, locked.ctypes.data, n),
              "sdr_pll_process")
        if x.ndim <= 1:
            return out[0], locked[0]
        return out, locked

    def process_dev(self, d_in, n, d_out, d_locked, in_stride=None, out_stride=None):
        check(lib().sdr_pll_process_dev(self.h, _ptr(d_in), n, in_stride or n, _ptr(d_out), _ptr(d_locked),
                                        out_stride or n), "sdr_pll_process_dev")

    def stereo_decode(self, v):
        """the (mono, diff) closure of src/main.rs:62-71 around this (pilot) Pll: v f32 [n] or [n_streams, n] ->
        float32 [..., n, 2]"""
        v = np.ascontiguousarray(v, np.float32)
        rows = v.reshape(self.n_streams, -1)
        n = rows.shape[1]
        out = np.empty((self.n_streams, n, 2), np.float32)
        check(lib().sdr_pll_stereo_decode(self.h, rows.ctypes.data, n, n, out.ctypes.data, n), "sdr_pll_stereo_decode")
        return out[0] if v.ndim <= 1 else out

    def stereo_decode_dev(self, d_v, n, d_out, in_stride=None, out_stride=None):
        check(lib().sdr_pll_stereo_decode_dev(self.h, _ptr(d_v), n, in_stride or n, _ptr(d_out), out_stride or n),
              "sdr_pll_stereo_decode_dev")

    def state(self, idx=0):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        check(lib().sdr_pll_get_state(self.h, idx, C.byref(a), C.byref(b), C.byref(c)), "sdr_pll_get_state")
        return a.value, complex(b.value, c.value)

    def reset(self):
        check(lib().sdr_pll_reset(self.h), "sdr_pll_reset")

    def clone(self):
        err = C.c_int(0)
        h = lib().sdr_pll_clone(self.h, C.byref(err))
        if not h:
            raise SdrError(err.value, "sdr_pll_clone")
        return PllBatch(None, self.n_streams, 0.0, _handle=h)

    def close(self):
        if getattr(self, "h", None):
            lib().sdr_pll_destroy(self.h)
            self.h = None

    __del__ = close


class WindowFft:
    """signal.window(duration).decimate(fps).map(fft::fft) of examples/live.rs:30-39: one spectrum per kept sliding
    window.  window = duration_samples(rate, duration), hop = decimate_wait(rate, fps)."""

    def __init__(self, window, hop, fmt="c64", shift=True, norm=True, device=0, stream=None):
        self.fmt = {"u8iq": F.FMT_U8IQ, "c64": F.FMT_C64}[fmt]
        flags = (F.FFT_SHIFT if shift else 0) | (F.FFT_NORM if norm else 0)
        cfg = F.WindowFftConfig(int(window), int(hop), self.fmt, flags, device, _stream_ptr(stream))
        err = C.c_int(0)
        self.h = lib().sdr_window_fft_create(C.byref(cfg), C.byref(err))
        if not self.h:
            raise SdrError(err.value, "sdr_window_fft_create")
        self.window = int(window)

    def output_count(self, n_in):
        return lib().sdr_window_fft_output_count(self.h, n_in)

    def process(self, x):
        """x: uint8 [2n] (u8iq) or complex64 [n] -> complex64 [n_windows, window]"""
        if self.fmt == F.FMT_U8IQ:
            x = np.ascontiguousarray(x, np.uint8)
            n = x.size // 2
        else:
            x = np.ascontiguousarray(x, np.complex64)
            n = x.size
        nw = self.output_count(n)
        out = np.empty((max(nw, 1), self.window), np.complex64)
        got = C.c_size_t(0)
        check(lib().sdr_window_fft_process(self.h, x.ctypes.data if n else None, n, out.ctypes.data, max(nw, 1),
                                           C.byref(got)), "sdr_window_fft_process")
        return out[:got.value]

    def process_dev(self, d_in, n, d_out, out_cap):
        got = C.c_size_t(0)
        check(lib().sdr_window_fft_process_dev(self.h, _ptr(d_in), n, _ptr(d_out), out_cap, C.byref(got)),
              "sdr_window_fft_process_dev")
        return got.value

    def reset(self):
        check(lib().sdr_window_fft_reset(self.h), "sdr_window_fft_reset")

    def close(self):
        if getattr(self, "h", None):
            lib().sdr_window_fft_destroy(self.h)
            self.h = None

    __del__ = close


class FmStereo:
    """The FM broadcast stereo receiver of src/main.rs:32-81 for a batch of stations, device-resident end to end:
    u8 IQ -> Pll demodulator -> /75000 -> SincFastest to 144 kHz -> pilot Pll + (mono, diff) -> SincBest to 48 kHz ->
    Lr de-emphasis -> (left, right)."""

    def __init__(self, n_stations=1, rate=1.8e6, pilot=0.0, fast_math=True, device=0, stream=None):
        self.n_stations = int(n_stations)
        cfg = F.FmConfig(self.n_stations, rate, pilot, 0 if fast_math else F.PLL_F64_MATH, device, _stream_ptr(stream))
        err = C.c_int(0)
        self.h = lib().sdr_fm_create(C.byref(cfg), C.byref(err))
        if not self.h:
            raise SdrError(err.value, "sdr_fm_create")

    @property
    def output_rate(self):
        return lib().sdr_fm_output_rate(self.h)

    def max_output(self, n):
        return lib().sdr_fm_max_output(self.h, n)

    def process(self, iq, end_of_input=False):
        """iq: uint8 [2n] or [n_stations, 2n] -> float32 [n_out, 2] or [n_stations, n_out, 2] (left, right) at 48 kHz"""
        iq = np.ascontiguousarray(iq, np.uint8)
        rows = iq.reshape(self.n_stations, -1)
        n = rows.shape[1] // 2
        cap = self.max_output(n)
        out = np.empty((self.n_stations, cap, 2), np.float32)
        got = C.c_size_t(0)
        check(lib().sdr_fm_process(self.h, rows.ctypes.data, n, rows.shape[1], out.ctypes.data, cap, cap,
                                   C.byref(got), int(bool(end_of_input))), "sdr_fm_process")
        out = out[:, :got.value]
        return out[0] if iq.ndim <= 1 else out

    def process_dev(self, d_iq, n, in_stride, d_out, out_cap, out_stride=None, end_of_input=False):
        got = C.c_size_t(0)
        check(lib().sdr_fm_process_dev(self.h, _ptr(d_iq), n, in_stride, _ptr(d_out), out_cap, out_stride or out_cap,
                                       C.byref(got), int(bool(end_of_input))), "sdr_fm_process_dev")
        return got.value

    def reset(self):
        check(lib().sdr_fm_reset(self.h), "sdr_fm_reset")

    def close(self):
        if getattr(self, "h", None):
            lib().sdr_fm_destroy(self.h)
            self.h = None

    __del__ = close


class Channelizer:
    """n_channels x (Fir<f32,Complex<f32>> -> Pll): BASELINE config 4."""

    def __init__(self, taps, design, n_channels, rate, input_format="c64", fast_math=True, strict=False,
                 device=0, stream=None):
        taps = np.ascontiguousarray(taps, np.float32)
        self._taps = taps
        self.n_channels = int(n_channels)
        self.fmt = _FMT_OF[input_format]
        fc = F.FirConfig(taps.ctypes.data, taps.size, 0, self.fmt, 1, self.n_channels,
                         F.FIR_STRICT_ORDER if strict else 0, device, _stream_ptr(stream))
        arr = (F.PllDesign * 1)(design._c())
        self._arr = arr
        pc = F.PllConfig(C.cast(arr, C.c_void_p), 1, self.n_channels, rate, 0 if fast_math else F.PLL_F64_MATH,
                         device, None)
        err = C.c_int(0)
        self.h = lib().sdr_channelizer_create(C.byref(fc), C.byref(pc), C.byref(err))
        if not self.h:
            raise SdrError(err.value, "sdr_channelizer_create")

    def process(self, x):
        x = np.ascontiguousarray(x, _IN_DTYPE[self.fmt])
        per = 2 if self.fmt == F.FMT_U8IQ else 1
        rows = x.reshape(self.n_channels, -1)
        n = rows.shape[1] // per
        out = np.empty((self.n_channels, n), np.float32)
        locked = np.empty((self.n_channels, n), np.uint8)
        check(lib().sdr_channelizer_process(self.h, rows.ctypes.data, n, n, out.ctypes.data, locked.ctypes.data, n),
              "sdr_channelizer_process")
        return out, locked

    def process_dev(self, d_in, n, d_out, d_locked, in_stride=None, out_stride=None):
        check(lib().sdr_channelizer_process_dev(self.h, _ptr(d_in), n, in_stride or n, _ptr(d_out), _ptr(d_locked),
                                                out_stride or n), "sdr_channelizer_process_dev")

    def reset(self):
        check(lib().sdr_channelizer_reset(self.h), "sdr_channelizer_reset")

    def close(self):
        if getattr(self, "h", None):
            lib().sdr_channelizer_destroy(self.h)
            self.h = None

    __del__ = close


class ConverterType(enum.IntEnum):
    """resample::ConverterType (src/resample.rs:112-119)."""
    SincBestQuality = F.SRC_SINC_BEST_QUALITY
    SincMediumQuality = F.SRC_SINC_MEDIUM_QUALITY
    SincFastest = F.SRC_SINC_FASTEST
    ZeroOrderHold = F.SRC_ZERO_ORDER_HOLD
    Linear = F.SRC_LINEAR

    def name_str(self):
        return lib().sdr_src_get_name(int(self)).decode()

    def description(self):
        return lib().sdr_src_get_description(int(self)).decode()


class ResampleError(RuntimeError):
    def __init__(self, code):
        self.code = code
        msg = lib().sdr_src_strerror(code)
        super().__init__(msg.decode() if msg else "Unknown(%d)" % code)


class SampleRate:
    """resample::SampleRate<A> (src/resample.rs:11-110).  channels: 1 for f32, 2 for Complex<f32>."""

    def __init__(self, typ, channels, device=0, stream=None, _handle=None):
        self.channels = int(channels)
        if _handle is not None:
            self.h = _handle
            return
        err = C.c_int(0)
        self.h = lib().sdr_src_new_on(int(typ), self.channels, device, _stream_ptr(stream), C.byref(err))
        if not self.h:
            raise ResampleError(err.value)

    def process(self, ratio, inp, out_capacity):
        """SampleRate::process(ratio, &input, &mut output) (resample.rs:46-67): returns
        (input_frames_used, output[frames, channels]); end_of_input = input.is_empty()."""
        inp = np.ascontiguousarray(inp, np.float32).reshape(-1, self.channels)
        out = np.empty((max(out_capacity, 1), self.channels), np.float32)
        d = F.SrcData(inp.ctypes.data if inp.size else None, out.ctypes.data, inp.shape[0], out_capacity, 0, 0,
                      1 if inp.shape[0] == 0 else 0, ratio)
        rc = lib().sdr_src_process(self.h, C.byref(d))
        if rc != 0:
            raise ResampleError(rc)
        return d.input_frames_used, out[:d.output_frames_gen].copy()

    def process_dev(self, ratio, d_in, n_frames, d_out, out_capacity, end_of_input=False):
        """device-resident SampleRate::process: d_in / d_out are device buffers of interleaved f32 frames; the counts
        come back synchronously, the samples land asynchronously on the handle's stream.  Returns (used, generated)."""
        d = F.SrcData(_ptr(d_in) if n_frames else None, _ptr(d_out), int(n_frames), int(out_capacity), 0, 0,
                      1 if (end_of_input or n_frames == 0) else 0, ratio)
        rc = lib().sdr_src_process_dev(self.h, C.byref(d))
        if rc != 0:
            raise ResampleError(rc)
        return d.input_frames_used, d.output_frames_gen

    def reset(self):
        rc = lib().sdr_src_reset(self.h)
        if rc:
            raise ResampleError(rc)

    def try_clone(self):
        err = C.c_int(0)
        h = lib().sdr_src_clone(self.h, C.byref(err))
        if not h:
            raise ResampleError(err.value)
        return SampleRate(0, self.channels, _handle=h)

    def set_ratio(self, ratio):
        rc = lib().sdr_src_set_ratio(self.h, ratio)
        if rc:
            raise ResampleError(rc)

    def get_channels(self):
        return lib().sdr_src_get_channels(self.h)

    def history_frames(self):
        return lib().sdr_src_history_frames(self.h)

    def set_exact(self, exact=True):
        """True: always the f64 kernels (the converter's specification up to the final f32 rounding); False (default):
        long integer-step calls on c64 data may take the tensor-core polyphase path (within 1e-5 of max|y|)"""
        rc = lib().sdr_src_set_exact(self.h, 1 if exact else 0)
        if rc:
            raise ResampleError(rc)

    def close(self):
        if getattr(self, "h", None):
            lib().sdr_src_delete(self.h)
            self.h = None

    __del__ = close


def sinc_table(typ):
    tab = C.c_void_p()
    inc = C.c_int()
    n = lib().sdr_src_sinc_table(int(typ), C.byref(tab), C.byref(inc))
    arr = np.ctypeslib.as_array(C.cast(tab, C.POINTER(C.c_float)), shape=(n + 2,)).copy()
    return arr, inc.value, n


class Timer:
    """CUDA-event timer on a given stream (the stream kernels are launched on)."""

    def __init__(self, device=0, stream=None):
        err = C.c_int(0)
        self.h = lib().sdr_timer_create(device, _stream_ptr(stream), C.byref(err))
        if not self.h:
            raise SdrError(err.value, "sdr_timer_create")

    def begin(self):
        check(lib().sdr_timer_begin(self.h), "sdr_timer_begin")

    def end(self):
        ms = C.c_float(0)
        check(lib().sdr_timer_end(self.h, C.byref(ms)), "sdr_timer_end")
        return ms.value

    def close(self):
        if getattr(self, "h", None):
            lib().sdr_timer_destroy(self.h)
            self.h = None

    __del__ = close
