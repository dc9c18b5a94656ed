"""Regenerates tests/golden/oracle_golden.npz from the CPU oracle (run in the build container).
The reference itself cannot run here (no Rust toolchain) and ships no golden vectors, so these are
pins of the oracle -- which test_oracle.py independently checks against numpy / scipy / a pure-Python
restatement -- for the GPU box to compare the CUDA path with."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import gen  # noqa: E402
import oracle_lib as O  # noqa: E402

taps64, taps255 = gen.lowpass_taps(64, 200e3, 2.048e6), gen.lowpass_taps(255, 100e3, 2.4e6)
iq = gen.tone_noise_u8(4096, 2.048e6, 300e3, 0.5, 0.1, gen.BASE_SEED + 1)
fm = gen.fm_u8(8192, 2.4e6, 75e3, 1e3, 0.05, gen.BASE_SEED + 3)
y = O.Fir(taps255).apply(O.unpack_u8iq(fm))[9::10]
fr, v = O.freq_sweep(1800000.0, 20000.0, True, -200000.0, 200000.0)
d = O.pll_design(0.0, 0.035, (O.BQ_LOWPASS, 80000.0, 0.7), (O.BQ_LOWPASS, 20000.0, 0.7), (O.BQ_LOWPASS, 20000.0, 0.7))
out, lk = O.Pll(d, 1.8e6).apply(v)
# FM stereo decode (src/main.rs:62-71) around the pilot Pll, on a small synthetic multiplex at 144 kHz
tt = np.arange(6000) / 144000.0
mpx = (0.25 * np.sin(2 * np.pi * 1000 * tt) + 0.1 * np.sin(2 * np.pi * 19000 * tt) +
       0.2 * np.sin(2 * np.pi * 700 * tt) * np.sin(2 * np.pi * 38000 * tt)).astype(np.float32)
pd = O.pll_design(19000.0, 0.0002, (O.BQ_LOWPASS, 200.0, 0.7), (O.BQ_LOWPASS, 20.0, 0.7), (O.BQ_LOWPASS, 20.0, 0.7))
stereo = O.Pll(pd, np.float32(144000.0)).stereo_decode(mpx)
np.savez_compressed(
    os.path.join(HERE, "oracle_golden.npz"),
    c1_iq=iq, c1_fir64=O.Fir(taps64).apply(O.unpack_u8iq(iq)), c2_fft1024=O.fft_batch_u8(iq, 1024, 1),
    c3_iq=fm, c3_fir255_dec10=y, c3_resampled=O.resample_signal(y, O.SRC_SINC_FASTEST, 0.2),
    sweep_freq=fr, sweep=v, pll_out=out, pll_locked=lk,
    bq_lp80k=O.biquad_design(O.BQ_LOWPASS, 80000.0, 0.7, 1.8e6), fm_mpx=mpx, fm_mono_diff=stereo)
print("wrote", os.path.join(HERE, "oracle_golden.npz"))
