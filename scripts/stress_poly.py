"""GPU stress of the polyphase decimating FIR (fir_umma_poly_kernel): random (K, D in the polyphase set, tap kind, channels,
length, 3-way blocking) against the f64 truth, and bit-identity between two different blockings.  Not part of the pytest
suite (minutes); run with `python scripts/stress_poly.py [trials] [seed]`."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "unnamed-rust-sdr_b200"))
import gen, oracle_lib as O
import sdr_b200 as sdr
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
worst = 0.0
for t in range(trials):
    K = int(rng.choice([1, 3, 9, 10, 11, 25, 64, 100, 250, 255, 256, 300, 400, 511]))
    D = int(rng.choice([5, 6, 7, 8, 9, 10, 12]))
    tc = bool(rng.integers(0, 2)); n_ch = int(rng.choice([1, 1, 2, 3]))
    n = int(rng.choice([1, 9, 10, 11, 1023, 10239, 10240, 10241, 20480, 30011, 65536, 102400 + int(rng.integers(0, 50))]))
    taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    if tc: taps = (taps + 1j * rng.standard_normal(K) / np.sqrt(K)).astype(np.complex64)
    raw = gen.random_u8(2 * n_ch * n, 9000 + t).reshape(n_ch, 2 * n)
    outs = []
    for rep in range(2):
        f = sdr.Fir(taps, "u8iq", decimation=D, n_channels=n_ch)
        cuts = sorted(int(c) for c in rng.integers(0, n + 1, 2))
        edges = [0] + cuts + [n]
        parts = [f.process(np.ascontiguousarray(raw[:, 2 * a:2 * b])).reshape(n_ch, -1) for a, b in zip(edges[:-1], edges[1:])]
        outs.append(np.concatenate(parts, axis=1))
        path = f.last_path
    assert outs[0].shape == (n_ch, n // D), (t, K, D, n, outs[0].shape)
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32)) or path != 4, ("blocking", t, K, D, tc, n_ch, n)
    for c in range(n_ch):
        truth = O.fir_f64(taps, O.unpack_u8iq(raw[c]))[D - 1::D]
        if len(truth):
            # scale: the larger of max|truth| and the output's rms for uniform bytes (one- or two-sample calls can land on a
            # cancellation, where an error relative to max|truth| says nothing)
            scale = max(float(np.abs(truth).max()), 0.5 * float(np.linalg.norm(taps)))
            e = float(np.abs(outs[0][c] - truth).max()) / scale; worst = max(worst, e)
            assert e < (1e-6 if path == 4 else 1e-5), ("value", t, K, D, tc, n_ch, n, path, e)
print("stress_poly: %d trials ok, worst relative error %.2e" % (trials, worst))
