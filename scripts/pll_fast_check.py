import os, sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, 'unnamed-rust-sdr_b200')
import oracle_lib as O
import sdr_b200 as sdr
import importlib
T = importlib.import_module('test_gpu_pll_resample')
rate, gain = 1.8e6, 0.035
full = rate * gain * np.pi
t = np.arange(20000)
ph = 2 * np.pi * 30e3 * t / rate + (50e3 / 1e3) * np.sin(2 * np.pi * 1e3 * t / rate)
x = np.exp(1j * ph).astype(np.complex64)
ro, rl = O.Pll(T.oracle_design(), rate).apply(x)
for fast in (False, True):
    out, lk = sdr.PllBatch([T.example_design(sdr)], 1, rate, fast_math=fast).process(x)
    d = np.abs(out.astype(np.float64) - ro) / full
    print('fast', fast, 'max', d.max(), 'median', np.median(d), 'lock equal', np.array_equal(lk, rl))
