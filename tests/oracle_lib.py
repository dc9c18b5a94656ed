"""ctypes view of oracle/libsdr_oracle.so -- the CPU oracle (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB_PATH = os.path.join(ORACLE_DIR, "libsdr_oracle.so")

KIND_F32, KIND_C64 = 1, 2
BQ_IDENTITY, BQ_LOWPASS, BQ_HIGHPASS, BQ_BANDPASS, BQ_NOTCH, BQ_LR = range(6)
SRC_SINC_BEST, SRC_SINC_MEDIUM, SRC_SINC_FASTEST, SRC_ZOH, SRC_LINEAR = range(5)


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(ORACLE_DIR, f))
                                          for f in ("sdr_oracle.cpp", "sdr_oracle.h"))):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "-B"])
    return _LIB_PATH


class PllDesign(C.Structure):
    _fields_ = [("reference", C.c_float), ("gain", C.c_float),
                ("loop_kind", C.c_int), ("loop_p0", C.c_float), ("loop_p1", C.c_float),
                ("out_kind", C.c_int), ("out_p0", C.c_float), ("out_p1", C.c_float),
                ("lock_kind", C.c_int), ("lock_p0", C.c_float), ("lock_p1", C.c_float)]


class SrcData(C.Structure):
    _fields_ = [("data_in", C.c_void_p), ("data_out", C.c_void_p),
                ("input_frames", C.c_long), ("output_frames", C.c_long),
                ("input_frames_used", C.c_long), ("output_frames_gen", C.c_long),
                ("end_of_input", C.c_int), ("src_ratio", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, sz, f, i = C.c_void_p, C.c_size_t, C.c_float, C.c_int
        sig = {
            "orc_unpack_u8iq": (None, [vp, sz, vp]),
            "orc_fir_new": (vp, [vp, sz, i, i]),
            "orc_fir_free": (None, [vp]),
            "orc_fir_reset": (None, [vp]),
            "orc_fir_clone": (vp, [vp]),
            "orc_fir_apply": (None, [vp, vp, sz, vp]),
            "orc_fir_f64": (None, [vp, sz, i, i, vp, sz, vp]),
            "orc_decimate_wait": (sz, [f, f]),
            "orc_decimate": (sz, [vp, sz, sz, i, vp, vp]),
            "orc_round_count": (sz, [f, f]),
            "orc_block_size": (sz, [f, f]),
            "orc_times": (None, [f, sz, sz, vp]),
            "orc_fft_f32": (None, [vp, sz, vp]),
            "orc_fft_shifted": (None, [vp, sz, f, vp, vp]),
            "orc_rfft_shifted": (sz, [vp, sz, f, vp, vp]),
            "orc_dft_f64": (None, [vp, sz, vp]),
            "orc_fft_batch_u8": (None, [vp, sz, sz, i, vp]),
            "orc_fft_batch_c64": (None, [vp, sz, sz, i, i, vp]),
            "orc_biquad_design": (None, [i, f, f, f, vp]),
            "orc_biquad_new": (vp, [i, f, f, f, i]),
            "orc_biquad_free": (None, [vp]),
            "orc_biquad_apply": (None, [vp, vp, sz, vp]),
            "orc_pll_new": (vp, [vp, f]),
            "orc_pll_free": (None, [vp]),
            "orc_pll_apply": (None, [vp, vp, sz, vp, vp]),
            "orc_pll_state": (None, [vp, vp, vp, vp]),
            "orc_fm_stereo_decode": (None, [vp, vp, sz, vp]),
            "orc_freq_sweep": (sz, [f, f, i, f, f, vp, vp, sz]),
            "orc_src_new": (vp, [i, i, vp]),
            "orc_src_delete": (vp, [vp]),
            "orc_src_process": (i, [vp, vp]),
            "orc_src_reset": (i, [vp]),
            "orc_src_clone": (vp, [vp, vp]),
            "orc_src_set_ratio": (i, [vp, C.c_double]),
            "orc_src_get_channels": (i, [vp]),
            "orc_src_history_frames": (C.c_long, [vp]),
            "orc_src_strerror": (C.c_char_p, [i]),
            "orc_src_sinc_table": (sz, [i, vp, vp]),
            "orc_resample_signal": (sz, [vp, sz, i, i, C.c_double, vp, sz]),
            "orc_fir_u8_mt": (sz, [vp, sz, vp, sz, i, sz, i, vp]),
            "orc_channelizer_mt": (None, [vp, sz, sz, vp, sz, vp, f, i, vp, vp]),
            "orc_hardware_threads": (i, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c64(a):
    a = np.ascontiguousarray(a, dtype=np.complex64)
    return a, a.view(np.float32)


# ---------------------------------------------------------------------------------------------
def unpack_u8iq(iq):
    iq = np.ascontiguousarray(iq, dtype=np.uint8)
    n = iq.size // 2
    out = np.empty(n, np.complex64)
    lib().orc_unpack_u8iq(_p(iq), n, _p(out))
    return out


class Fir:
    """Fir<C,A> (src/filter/fir.rs).  taps: float32 or complex64; kind: KIND_F32 / KIND_C64."""

    def __init__(self, taps, kind=KIND_C64, _h=None):
        taps = np.asarray(taps)
        self.taps_complex = int(np.iscomplexobj(taps))
        self.taps = np.ascontiguousarray(taps, np.complex64 if self.taps_complex else np.float32)
        self.kind = kind
        self.h = _h or lib().orc_fir_new(_p(self.taps), self.taps.size, self.taps_complex, kind)
        assert self.h

    def apply(self, x):
        if self.kind == KIND_C64:
            x = np.ascontiguousarray(x, np.complex64)
            out = np.empty(x.size, np.complex64)
        else:
            x = np.ascontiguousarray(x, np.float32)
            out = np.empty(x.size, np.float32)
        lib().orc_fir_apply(self.h, _p(x), x.size, _p(out))
        return out

    def reset(self):
        lib().orc_fir_reset(self.h)

    def clone(self):
        return Fir(self.taps, self.kind, _h=lib().orc_fir_clone(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fir_free(self.h)
            self.h = None


def fir_f64(taps, x, kind=KIND_C64):
    taps = np.asarray(taps)
    tc = int(np.iscomplexobj(taps))
    taps = np.ascontiguousarray(taps, np.complex64 if tc else np.float32)
    if kind == KIND_C64:
        x = np.ascontiguousarray(x, np.complex64)
        out = np.empty(x.size, np.complex128)
    else:
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(x.size, np.float64)
    lib().orc_fir_f64(_p(taps), taps.size, tc, kind, _p(x), x.size, _p(out))
    return out


def decimate_wait(rate_in, rate_out):
    return lib().orc_decimate_wait(rate_in, rate_out)


def decimate(x, wait, phase=0):
    """returns (out, new_phase)"""
    x = np.ascontiguousarray(x)
    ef = 2 if np.iscomplexobj(x) else 1
    x = x.astype(np.complex64 if ef == 2 else np.float32, copy=False)
    out = np.empty(x.size // max(wait, 1) + 1, x.dtype)
    ph = C.c_size_t(phase)
    n = lib().orc_decimate(_p(x), x.size, wait, ef, C.byref(ph), _p(out))
    return out[:n].copy(), ph.value


def round_count(rate, duration):
    return lib().orc_round_count(rate, duration)


def block_size(size, rate):
    return lib().orc_block_size(size, rate)


def times(rate, start, n):
    out = np.empty(n, np.float32)
    lib().orc_times(rate, start, n, _p(out))
    return out


def fft_f32(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty_like(x)
    lib().orc_fft_f32(_p(x), x.size, _p(out))
    return out


def fft_shifted(x, rate=1.0):
    x = np.ascontiguousarray(x, np.complex64)
    labels = np.empty(x.size, np.float32)
    vals = np.empty_like(x)
    lib().orc_fft_shifted(_p(x), x.size, rate, _p(labels), _p(vals))
    return labels, vals


def rfft_shifted(x, rate=1.0):
    x = np.ascontiguousarray(x, np.float32)
    labels = np.empty(x.size, np.float32)
    vals = np.empty(x.size, np.complex64)
    k = lib().orc_rfft_shifted(_p(x), x.size, rate, _p(labels), _p(vals))
    return labels[:k].copy(), vals[:k].copy()


def dft_f64(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty(x.size, np.complex128)
    lib().orc_dft_f64(_p(x), x.size, _p(out))
    return out


def fft_batch_u8(iq, n, threads=1):
    iq = np.ascontiguousarray(iq, np.uint8)
    batches = iq.size // (2 * n)
    out = np.empty(batches * n, np.complex64)
    lib().orc_fft_batch_u8(_p(iq), n, batches, threads, _p(out))
    return out.reshape(batches, n)


def fft_batch_c64(x, n, shifted=False, threads=1):
    x = np.ascontiguousarray(x, np.complex64)
    batches = x.size // n
    out = np.empty(batches * n, np.complex64)
    lib().orc_fft_batch_c64(_p(x), n, batches, int(shifted), threads, _p(out))
    return out.reshape(batches, n)


def biquad_design(kind, p0, p1, rate):
    c = np.empty(5, np.float32)
    lib().orc_biquad_design(kind, p0, p1, rate, _p(c))
    return c


def biquad_apply(kind, p0, p1, rate, x):
    x = np.ascontiguousarray(x)
    sk = KIND_C64 if np.iscomplexobj(x) else KIND_F32
    x = x.astype(np.complex64 if sk == KIND_C64 else np.float32, copy=False)
    h = lib().orc_biquad_new(kind, p0, p1, rate, sk)
    out = np.empty_like(x)
    lib().orc_biquad_apply(h, _p(x), x.size, _p(out))
    lib().orc_biquad_free(h)
    return out


def pll_design(reference, gain, loop, out, lock):
    """loop/out/lock are (kind, p0, p1) triples; (BQ_IDENTITY, 0, 0) = filter::Identity"""
    return PllDesign(reference, gain, loop[0], loop[1], loop[2], out[0], out[1], out[2],
                     lock[0], lock[1], lock[2])


class Pll:
    def __init__(self, design, rate):
        self.design = design
        self.h = lib().orc_pll_new(C.byref(design), rate)

    def apply(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        out = np.empty(x.size, np.float32)
        locked = np.empty(x.size, np.uint8)
        lib().orc_pll_apply(self.h, _p(x), x.size, _p(out), _p(locked))
        return out, locked

    def stereo_decode(self, v):
        """the (mono, diff) closure of src/main.rs:62-71 around this (pilot) Pll"""
        v = np.ascontiguousarray(v, np.float32)
        out = np.empty((v.size, 2), np.float32)
        lib().orc_fm_stereo_decode(self.h, _p(v), v.size, _p(out))
        return out

    def state(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        lib().orc_pll_state(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, complex(b.value, c.value)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_pll_free(self.h)
            self.h = None


def freq_sweep(rate, df, warmup, start, end):
    n = lib().orc_freq_sweep(rate, df, int(warmup), start, end, None, None, 0)
    fr = np.empty(n, np.float32)
    v = np.empty(n, np.complex64)
    lib().orc_freq_sweep(rate, df, int(warmup), start, end, _p(fr), _p(v), n)
    return fr, v


class SampleRate:
    """resample::SampleRate<A> (src/resample.rs:11-110) over the oracle's sdr-src."""

    def __init__(self, typ, channels, _h=None):
        err = C.c_int(0)
        self.channels = channels
        self.h = _h or lib().orc_src_new(typ, channels, C.byref(err))
        self.err = err.value

    def process(self, ratio, inp, out_capacity):
        """returns (input_frames_used, output ndarray [frames, channels]) or raises"""
        inp = np.ascontiguousarray(inp, np.float32).reshape(-1, self.channels)
        out = np.empty((out_capacity, self.channels), np.float32)
        d = SrcData(inp.ctypes.data, out.ctypes.data, inp.shape[0], out_capacity, 0, 0,
                    1 if inp.shape[0] == 0 else 0, ratio)
        rc = lib().orc_src_process(self.h, C.byref(d))
        if rc != 0:
            raise RuntimeError(lib().orc_src_strerror(rc))
        return d.input_frames_used, out[:d.output_frames_gen].copy()

    def reset(self):
        return lib().orc_src_reset(self.h)

    def history_frames(self):
        return lib().orc_src_history_frames(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_src_delete(self.h)
            self.h = None


def sinc_table(typ):
    tab = C.c_void_p()
    inc = C.c_int()
    n = lib().orc_src_sinc_table(typ, C.byref(tab), C.byref(inc))
    arr = np.ctypeslib.as_array(C.cast(tab, C.POINTER(C.c_float)), shape=(n + 2,)).copy()
    return arr, inc.value, n


def resample_signal(x, typ, ratio):
    x = np.ascontiguousarray(x)
    ch = 2 if np.iscomplexobj(x) else 1
    x = x.astype(np.complex64 if ch == 2 else np.float32, copy=False)
    cap = int(x.size * ratio) + 8192
    out = np.empty(cap, x.dtype)
    n = lib().orc_resample_signal(_p(x), x.size, ch, typ, ratio, _p(out), cap)
    assert n <= cap
    return out[:n].copy()


def fir_u8_mt(iq, taps, wait=1, threads=1):
    iq = np.ascontiguousarray(iq, np.uint8)
    taps = np.asarray(taps)
    tc = int(np.iscomplexobj(taps))
    taps = np.ascontiguousarray(taps, np.complex64 if tc else np.float32)
    n = iq.size // 2
    out = np.empty(n // wait, np.complex64)
    k = lib().orc_fir_u8_mt(_p(iq), n, _p(taps), taps.size, tc, wait, threads, _p(out))
    return out[:k]


def channelizer_mt(x, taps, design, rate, threads=1):
    x = np.ascontiguousarray(x, np.complex64)
    n_ch, n = x.shape
    taps = np.ascontiguousarray(taps, np.float32)
    out = np.empty((n_ch, n), np.float32)
    locked = np.empty((n_ch, n), np.uint8)
    lib().orc_channelizer_mt(_p(x), n_ch, n, _p(taps), taps.size, C.byref(design), rate, threads,
                             _p(out), _p(locked))
    return out, locked


def hardware_threads():
    return lib().orc_hardware_threads()
