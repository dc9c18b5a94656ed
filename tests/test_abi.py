"""CPU: the C-ABI library loads and exports every symbol include/sdr_b200.h declares; host-pure entry
points agree with the oracle; without a GPU every constructor fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import gen
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sdr_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(sdr):
    syms = declared_symbols()
    assert len(syms) > 50
    L = C.CDLL(sdr.LIB_PATH)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    # and the Python binding prototypes cover exactly the header
    assert sorted(sdr.PROTOTYPES) == syms


def test_exported_symbols_are_only_the_abi(sdr):
    out = subprocess.check_output(["nm", "-D", "--defined-only", sdr.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    extra = [s for s in exported if s.startswith("sdr_") and s not in declared_symbols()]
    assert not extra, extra


def test_product_does_not_link_or_reference_the_oracle(sdr):
    out = subprocess.check_output(["ldd", sdr.LIB_PATH], text=True)
    assert "oracle" not in out
    blob = open(sdr.LIB_PATH, "rb").read()
    assert b"orc_" not in blob and b"libsdr_oracle" not in blob
    pkg = os.path.dirname(os.path.dirname(sdr.LIB_PATH))
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".rs")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "sdr_oracle" not in txt, os.path.join(dp, f)


def test_abi_version_and_strerror(sdr):
    L = sdr.lib()
    assert L.sdr_abi_version() == 1
    assert L.sdr_strerror(0) == b"No error."
    assert b"no CPU fallback" in L.sdr_strerror(102)
    assert L.sdr_src_strerror(6) == O.lib().orc_src_strerror(6)
    assert L.sdr_src_strerror(55) is None
    assert L.sdr_src_get_name(4) == b"Linear Interpolator"
    assert L.sdr_src_get_name(7) is None


def test_host_pure_helpers_match_oracle(sdr):
    for a, b in [(2.4e6, 240e3), (300000.0, 60.0), (1.8e6, 144e3), (44100.0, 8000.0), (1.0, 3.0)]:
        assert sdr.decimate_wait(a, b) == O.decimate_wait(a, b)
    for r, d in [(1.8e6, 0.1), (44100.0, 0.01), (1000.0, 0.0005), (300000.0, 1000.0 / 300000.0)]:
        assert sdr.duration_samples(r, d) == O.round_count(r, d)
        assert sdr.block_samples(d, r) == O.block_size(d, r)
    for n in (1, 7, 1000, 1024, 14400):
        x = np.zeros(n, np.complex64)
        lab, _ = O.fft_shifted(x, 144000.0)
        assert np.array_equal(sdr.fft_labels(n, 144000.0), lab)
        lr, _ = O.rfft_shifted(np.zeros(n, np.float32), 144000.0)
        assert np.array_equal(sdr.fft_labels(n, 144000.0, rfft=True), lr)


def test_biquad_design_matches_oracle(sdr):
    cases = [(sdr.BiquadD.LowPass(80000.0, 0.7), O.BQ_LOWPASS), (sdr.BiquadD.HighPass(1000.0, 0.5), O.BQ_HIGHPASS),
             (sdr.BiquadD.BandPass(19000.0, 5.0), O.BQ_BANDPASS), (sdr.BiquadD.Notch(19000.0, 5.0), O.BQ_NOTCH),
             (sdr.BiquadD.Lr(13333.0), O.BQ_LR)]
    for d, kind in cases:
        for rate in (1.8e6, 144000.0, 44100.0):
            assert np.array_equal(d.coefficients(rate), O.biquad_design(kind, d.p0, d.p1, rate))


def test_sinc_tables_match_oracle(sdr):
    for typ in (0, 1, 2):
        a, ia, na = sdr.sinc_table(typ)
        b, ib, nb = O.sinc_table(typ)
        assert ia == ib and na == nb and np.array_equal(a, b)


def test_no_gpu_means_loud_failure_not_fallback(sdr):
    if sdr.device_count() > 0:
        pytest.skip("a GPU is present")
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    for ctor in (lambda: sdr.Fir(taps, "u8iq"), lambda: sdr.FftPlan(1024, "u8iq"),
                 lambda: sdr.PllDesign(0.0, 0.035, sdr.BiquadD.LowPass(8e4, .7), sdr.Identity(), sdr.Identity()).design(1.8e6)):
        with pytest.raises(sdr.SdrError) as e:
            ctor()
        assert e.value.code == 102
    with pytest.raises(sdr.ResampleError) as e:
        sdr.SampleRate(sdr.ConverterType.Linear, 2)
    assert e.value.code == 102
    with pytest.raises(sdr.SdrError):
        sdr.unpack_u8iq(np.zeros(16, np.uint8))


def test_argument_validation_without_gpu(sdr):
    L = sdr.lib()
    err = C.c_int(0)
    assert L.sdr_fir_create(None, C.byref(err)) is None and err.value == 100
    cfg = sdr._ffi.FirConfig(None, 0, 0, 0, 1, 1, 0, 0, None)
    assert L.sdr_fir_create(C.byref(cfg), C.byref(err)) is None and err.value == 100
    taps = np.ones(4, np.float32)
    cfg = sdr._ffi.FirConfig(taps.ctypes.data, 4, 0, 0, 0, 1, 0, 0, None)  # decimation 0: reference underflows
    assert L.sdr_fir_create(C.byref(cfg), C.byref(err)) is None and err.value == 100
    assert L.sdr_src_new(9, 1, C.byref(err)) is None and err.value == 10
    assert L.sdr_src_new(4, 0, C.byref(err)) is None and err.value == 11
    assert L.sdr_fir_process(None, None, 0, 0, None, 0, 0, None, None) == 103
    assert L.sdr_src_process(None, None) == 2
    fc = sdr._ffi.FftConfig(0, 1, 0, 0, None)
    assert L.sdr_fft_create(C.byref(fc), C.byref(err)) is None and err.value == 100
