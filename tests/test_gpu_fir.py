"""GPU parity: unpack + FIR (+ Decimate) through the C ABI vs the oracle.
Bars (BASELINE.json north_star): bit-exact unpack, decimation indexing and output counts; FIR within
1e-5 of max|y| of the f64 truth (and, in STRICT_ORDER mode, bit-identical to the reference's f32 order)."""
import os

import numpy as np
import pytest

import gen
import oracle_lib as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5


def rel_err(y, truth):
    return float(np.abs(y.astype(np.complex128) - truth).max() / max(np.abs(truth).max(), 1e-30))


def test_unpack_exhaustive_table_and_random(sdr):
    u = np.arange(256, dtype=np.uint8)
    iq = np.stack([u, u[::-1]], 1).ravel()
    assert np.array_equal(sdr.unpack_u8iq(iq).view(np.uint32), O.unpack_u8iq(iq).view(np.uint32))
    for n in (1, 7, 8, 9, 1000, 1 << 20):
        raw = gen.random_u8(2 * n, 42 + n)
        assert np.array_equal(sdr.unpack_u8iq(raw).view(np.uint32), O.unpack_u8iq(raw).view(np.uint32))
    assert len(sdr.unpack_u8iq(np.zeros(0, np.uint8))) == 0


@pytest.mark.parametrize("K", [1, 5, 8, 63, 64, 255])
@pytest.mark.parametrize("tc", [False, True])
def test_fir_strict_is_bit_identical_to_reference_order_u8(sdr, K, tc):
    rng = np.random.default_rng(K)
    taps = rng.standard_normal(K).astype(np.float32)
    if tc:
        taps = (taps + 1j * rng.standard_normal(K)).astype(np.complex64)
    n = 10000 + K
    iq = gen.tone_noise_u8(n, 2.048e6, 300e3, 0.5, 0.1, 7 + K)
    want = O.Fir(taps).apply(O.unpack_u8iq(iq))
    f = sdr.Fir(taps, "u8iq", strict=True)
    got = f.process(iq)
    assert f.last_path == 2
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("fmt", ["c64", "f32"])
def test_fir_strict_c64_and_real_samples(sdr, fmt):
    taps = gen.lowpass_taps(255, 100e3, 2.4e6)
    if fmt == "c64":
        x = gen.complex_noise(20000, 3)
        want = O.Fir(taps).apply(x)
    else:
        x = gen.noise(20000, 3).astype(np.float32)
        want = O.Fir(taps, O.KIND_F32).apply(x)
    got = sdr.Fir(taps, fmt, strict=True).process(x)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("K", [57, 61, 64, 255])
def test_fir_strict_does_not_read_beyond_tap_k_minus_1(sdr, K):
    """Fir::apply multiplies exactly K samples (fir.rs:27-30).  An Inf / NaN sample must poison exactly the K outputs
    whose window holds it -- the kernel's zero padding of the tap table to a multiple of 8 must not extend that by
    Inf * 0 = NaN."""
    taps = gen.lowpass_taps(K, 100e3, 2.4e6)
    x = gen.complex_noise(8192, 9)
    x[3000] = np.inf + 0j
    x[5000] = np.nan
    want = O.Fir(taps).apply(x)
    got = sdr.Fir(taps, "c64", strict=True).process(x)
    assert np.array_equal(np.isnan(got.view(np.float32)), np.isnan(want.view(np.float32)))
    assert np.array_equal(got.view(np.uint32)[~np.isnan(want.view(np.float32))],
                          want.view(np.uint32)[~np.isnan(want.view(np.float32))])
    bad = np.isnan(want.real) | np.isinf(want.real)
    assert bad[3000:3000 + K].all() and not bad[3000 + K:5000].any()


@pytest.mark.parametrize("K,D,tc", [(64, 1, False), (64, 1, True), (255, 10, False), (255, 7, False), (255, 3, False), (33, 1, False)])
def test_tcgen05_path_is_alignment_invariant(sdr, K, D, tc):
    """A stream cut across GPUs on multiples of D hands each shard an input pointer at an arbitrary even address and
    an output pointer at an arbitrary 8-byte one.  The tcgen05 kernels must take such calls themselves (no fallback
    to a kernel with different rounding) and return the same bits as for an aligned copy of the same samples."""
    import torch
    dev = torch.device("cuda", 0)
    n = 200000
    iq = gen.random_u8(2 * n + 64, 41)
    taps = gen.complex_bandpass_taps(K, 200e3, 100e3, 2.048e6) if tc else gen.lowpass_taps(K, 100e3, 2.4e6)
    base = torch.from_numpy(iq).to(dev)
    ref = None
    for off in (0, 2, 6, 8, 14, 30):          # bytes
        for ooff in (0, 1):                   # complex64 elements
            f = sdr.Fir(taps, "u8iq", decimation=D)
            n_out = f.output_count(n)
            out = torch.zeros(n_out + 4, dtype=torch.complex64, device=dev)
            src = base[off:off + 2 * n] if off == 0 else base[0:2 * n].clone()
            if off:
                buf = torch.zeros(2 * n + 64, dtype=torch.uint8, device=dev)
                buf[off:off + 2 * n] = base[0:2 * n]
                src = buf[off:off + 2 * n]
            got = f.process_dev(src, n, out[ooff:], n_out)
            torch.cuda.synchronize()
            assert got == n_out and f.last_path == 4, (off, ooff, f.last_path)
            y = out[ooff:ooff + n_out].cpu().numpy()
            if ref is None:
                ref = y
                truth = O.fir_f64(taps, O.unpack_u8iq(iq[:2 * n]))[D - 1::D]
                assert np.abs(y - truth).max() / np.abs(truth).max() < 1e-5
            else:
                assert np.array_equal(y.view(np.uint32), ref.view(np.uint32)), (off, ooff)


def test_tcgen05_tma_staging_gives_the_same_bits(sdr):
    """fir_umma_kernel can stage its sliding window by TMA tensor copies (SDR_UMMA_TMA=1; off by default because it
    measured 4-5 % slower than cp.async on C1, see fir_umma.cu).  Both loaders must put the same bytes in the same
    places: every output bit equal, for the three row widths, decimation, several channels, odd pointer alignment
    and stream lengths that end inside a window row."""
    import subprocess
    import sys
    import tempfile
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import gen, sdr_b200 as sdr, torch
outs = []
dev = torch.device("cuda", 0)
for K, D, tc, n_ch, n in ((64, 1, False, 1, 70000), (64, 1, True, 1, 2 * 4096 + 4112), (255, 1, False, 3, 30000), (1, 3, False, 1, 9006),
                          (64, 3, True, 1, 41000), (255, 2, False, 2, 50000), (17, 1, False, 1, 12345), (300, 4, False, 1, 66000)):
    rng = np.random.default_rng(K + D)
    taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    if tc:
        taps = (taps + 1j * rng.standard_normal(K) / np.sqrt(K)).astype(np.complex64)
    iq = gen.random_u8(2 * n_ch * n, K * 7 + D).reshape(n_ch, 2 * n)
    f = sdr.Fir(taps, "u8iq", decimation=D, n_channels=n_ch)
    a = f.process(np.ascontiguousarray(iq[:, :2 * 777]))
    b = f.process(np.ascontiguousarray(iq[:, 2 * 777:]))
    assert f.last_path == 4, (K, D, f.last_path)
    outs += [a.ravel().view(np.float32), b.ravel().view(np.float32)]
    if n_ch == 1:   # device pointer at an odd alignment
        buf = torch.zeros(2 * n + 64, dtype=torch.uint8, device=dev)
        buf[6:6 + 2 * n] = torch.from_numpy(iq[0]).to(dev)
        g = sdr.Fir(taps, "u8iq", decimation=D)
        o = torch.zeros(g.output_count(n) + 1, dtype=torch.complex64, device=dev)
        got = g.process_dev(buf[6:6 + 2 * n], n, o, o.numel())
        torch.cuda.synchronize()
        outs.append(o[:got].cpu().numpy().view(np.float32))
np.save(sys.argv[1], np.concatenate(outs))
""" % (os.path.join(os.path.dirname(__file__), "..", "unnamed-rust-sdr_b200"), os.path.dirname(__file__))
    res = []
    for extra in ({}, {"SDR_UMMA_TMA": "1"}, {"SDR_UMMA_TMA": "1", "SDR_UMMA_P": "16"}, {"SDR_UMMA_P": "16"}):
        with tempfile.NamedTemporaryFile(suffix=".npy") as f:
            subprocess.check_call([sys.executable, "-c", code, f.name], env=dict(os.environ, **extra))
            res.append(np.load(f.name))
    assert res[0].size > 300000
    assert np.array_equal(res[0].view(np.uint32), res[1].view(np.uint32))
    assert np.array_equal(res[2].view(np.uint32), res[3].view(np.uint32))


SPLIT2 = 16  # SDR_FIR_SPLIT2


@pytest.mark.parametrize("K", [8, 33, 64, 255, 511])
@pytest.mark.parametrize("n_ch,n", [(1, 10000), (3, 2 * 4096 + 17), (130, 5000), (2, 4096)])
def test_tcgen05_c64_path_within_tolerance_of_f64_truth(sdr, K, n_ch, n):
    """Fir<f32, Complex<f32>> on c64 streams (C4's multi-channel front end) on the tcgen05 Toeplitz kernel: bf16-split
    operands, f32 accumulation in TMEM.  Default = three terms per operand: the reference's own accuracy (bar: 1.5e-6 of
    max|y| and <= 4x the reference-order f32 error); SDR_FIR_SPLIT2 = two terms: inside the north star's 1e-5."""
    taps = gen.lowpass_taps(K, 100e3, 1.8e6)
    x = gen.complex_noise(n_ch * n, 100 + K).reshape(n_ch, n)
    truth = np.stack([O.fir_f64(taps, r) for r in x])
    ref_err = max(rel_err(O.Fir(taps).apply(x[c]), truth[c]) for c in range(min(n_ch, 3)))
    f = sdr.Fir(taps, "c64", n_channels=n_ch)
    y = f.process(x)
    assert f.last_path == 5
    err = rel_err(y, truth)
    if K <= 384:   # three terms fit in shared memory
        assert err < 1.5e-6 and err <= 4 * ref_err + 2e-7, (err, ref_err)
    else:          # two terms
        assert err < 1e-5, err
    g = sdr.Fir(taps, "c64", n_channels=n_ch, flags=SPLIT2)
    y2 = g.process(x)
    assert g.last_path == 5 and rel_err(y2, truth) < 1e-5
    # a tone well inside the passband keeps its amplitude and phase
    t = np.arange(n)
    tone = np.exp(2j * np.pi * 0.004 * t).astype(np.complex64)
    yt = sdr.Fir(taps, "c64").process(tone)
    H = np.sum(taps.astype(np.float64) * np.exp(-2j * np.pi * 0.004 * np.arange(K)))
    assert np.abs(yt[K:] - H * tone[K:]).max() < 3e-6 * max(1.0, abs(H))


def test_tcgen05_c64_path_streaming_impulse_and_device_pointers(sdr):
    import torch
    K = 255
    taps = gen.lowpass_taps(K, 100e3, 1.8e6)
    x = gen.complex_noise(4 * 40000, 5).reshape(4, 40000)
    one = sdr.Fir(taps, "c64", n_channels=4).process(x)
    # blocks that are multiples of the 4096-output tile meet the same tile grid: identical bits
    f = sdr.Fir(taps, "c64", n_channels=4)
    parts = [f.process(np.ascontiguousarray(x[:, a:b])) for a, b in [(0, 8192), (8192, 12288), (12288, 40000)]]
    assert f.last_path == 5
    assert np.array_equal(np.concatenate(parts, 1).view(np.uint32), one.view(np.uint32))
    # ragged blocks: same samples up to the f32 accumulation order inside the tensor core
    g = sdr.Fir(taps, "c64", n_channels=4)
    cuts = [0, 1, 300, 5000, 5001, 20000, 40000]
    rag = np.concatenate([g.process(np.ascontiguousarray(x[:, a:b])) for a, b in zip(cuts[:-1], cuts[1:])], 1)
    assert np.abs(rag - one).max() <= 1e-6 * np.abs(one).max()
    # clone mid-stream, reset
    h = sdr.Fir(taps, "c64", n_channels=4)
    h.process(np.ascontiguousarray(x[:, :8192]))
    c = h.clone()
    a1, a2 = h.process(np.ascontiguousarray(x[:, 8192:16384])), c.process(np.ascontiguousarray(x[:, 8192:16384]))
    assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32))
    h.reset()
    assert np.array_equal(h.process(np.ascontiguousarray(x[:, :8192])).view(np.uint32), one[:, :8192].view(np.uint32))
    # impulse response = the taps (exactly: 1.0 has one bf16 term, the three tap terms add back to the f32 tap)
    imp = np.zeros(1000, np.complex64)
    imp[7] = 1.0 + 0.5j
    yi = sdr.Fir(taps, "c64").process(imp)
    assert np.array_equal(yi[7:7 + K].real, taps) and np.array_equal(yi[7:7 + K].imag, taps * np.float32(0.5))
    assert np.all(yi[:7] == 0) and np.all(yi[7 + K:] == 0)
    # device pointers at odd sample offsets (8-byte but not 16-byte aligned input): still this kernel, same tolerance
    dev = torch.device("cuda", 0)
    dx = torch.from_numpy(x[0]).to(dev)
    truth = O.fir_f64(taps, x[0][3:])
    out = torch.zeros(40000, dtype=torch.complex64, device=dev)
    k = sdr.Fir(taps, "c64")
    got = k.process_dev(dx[3:], 40000 - 3, out, 40000)
    torch.cuda.synchronize()
    assert got == 40000 - 3 and k.last_path == 5
    assert rel_err(out[:got].cpu().numpy(), truth) < 1.5e-6


NO_TENSOR, NO_TCGEN05 = 2, 4  # SDR_FIR_NO_TENSOR, SDR_FIR_NO_TCGEN05


@pytest.mark.parametrize("K,tc", [(64, False), (64, True), (255, False)])
@pytest.mark.parametrize("flags,path", [(0, 4), (NO_TCGEN05, 3)])
def test_fir_fast_path_within_tolerance_of_f64_truth(sdr, K, tc, flags, path):
    taps = gen.complex_bandpass_taps(K, 200e3, 100e3, 2.048e6) if tc else gen.lowpass_taps(K, 200e3, 2.048e6)
    iq = gen.tone_noise_u8(1 << 18, 2.048e6, 300e3, 0.5, 0.1, gen.BASE_SEED + 1)
    x = O.unpack_u8iq(iq)
    truth = O.fir_f64(taps, x)
    f = sdr.Fir(taps, "u8iq", flags=flags)
    got = f.process(iq)
    assert f.last_path == path  # 4: tcgen05 integer Toeplitz path, 3: mma.sync Toeplitz path
    if path == 4:
        assert rel_err(got, truth) < 1e-6  # exact integer accumulation: only tap quantisation (2^-23 of the largest tap) and the final f32 rounding are left
    e_gpu, e_ref = rel_err(got, truth), rel_err(O.Fir(taps).apply(x), truth)
    assert len(got) == len(x)
    cc = sdr.Fir(taps, "u8iq", flags=2)  # SDR_FIR_NO_TENSOR: CUDA-core FMA path
    got_cc = cc.process(iq)
    assert cc.last_path == 1 and rel_err(got_cc, truth) < TOL
    assert e_gpu < TOL, (e_gpu, e_ref)
    assert e_gpu <= 4 * e_ref + 1e-7  # no worse than the reference's own f32 rounding


def test_fir_impulse_response_is_the_taps(sdr):
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    x = np.zeros(300, np.complex64)
    x[0] = 1
    for strict in (False, True):
        y = sdr.Fir(taps, "c64", strict=strict).process(x)
        assert np.array_equal(y[:64].real, taps) and np.all(y[64:] == 0) and np.all(y.imag == 0)


@pytest.mark.parametrize("strict,flags", [(False, 0), (False, NO_TCGEN05), (True, 0)])
def test_fir_streaming_blocks_equal_one_call(sdr, strict, flags):
    taps = gen.lowpass_taps(255, 100e3, 2.4e6)
    iq = gen.fm_u8(50000, 2.4e6, 75e3, 1e3, 0.05, 21)
    whole = sdr.Fir(taps, "u8iq", strict=strict, flags=flags).process(iq)
    f = sdr.Fir(taps, "u8iq", strict=strict, flags=flags)
    cuts = [0, 1, 2, 9, 100, 254, 255, 256, 4096, 4097, 20000, 50000]
    parts = np.concatenate([f.process(iq[2 * a:2 * b]) for a, b in zip(cuts[:-1], cuts[1:])])
    if strict or flags == 0:
        # strict order, and the tcgen05 path (exact integer sums): any blocking gives the same bits
        assert f.last_path == (2 if strict else 4)
        assert np.array_equal(parts.view(np.uint32), whole.view(np.uint32))
    else:
        # tensor path: the grouping of terms into 16-wide MMA steps depends on the block alignment, so different
        # blockings agree to f32 rounding, not bit for bit (STRICT_ORDER is the bit-exact mode)
        assert len(parts) == len(whole) and np.abs(parts - whole).max() <= 2e-6 * np.abs(whole).max()
    assert len(f.process(np.zeros(0, np.uint8))) == 0


@pytest.mark.parametrize("D", [2, 5, 10, 50])
def test_decimating_fir_counts_indices_and_values(sdr, D):
    taps = gen.lowpass_taps(255, 100e3, 2.4e6)
    n = 40000 + 3
    iq = gen.fm_u8(n, 2.4e6, 75e3, 1e3, 0.05, 31)
    x = O.unpack_u8iq(iq)
    full = O.Fir(taps).apply(x)
    want, _ = O.decimate(full, D)
    got = sdr.Fir(taps, "u8iq", decimation=D, strict=True).process(iq)
    assert len(got) == n // D == len(want)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    fast = sdr.Fir(taps, "u8iq", decimation=D).process(iq)
    truth = O.fir_f64(taps, x)[D - 1::D]
    assert len(fast) == n // D and rel_err(fast, truth) < TOL
    # streaming with ragged blocks keeps the phase
    f = sdr.Fir(taps, "u8iq", decimation=D, strict=True)
    cuts = [0, 3, 3 + D - 1, 1000, 1001, 25000, n]
    parts = [f.process(iq[2 * a:2 * b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts).view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("K", [1, 2, 5, 8, 9, 63, 64, 65, 129, 255, 256, 1000])
@pytest.mark.parametrize("D", [1, 3, 10])
@pytest.mark.parametrize("tc", [False, True])
def test_tensor_path_shapes(sdr, K, D, tc):
    rng = np.random.default_rng(K * 31 + D)
    taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    if tc:
        taps = (taps + 1j * rng.standard_normal(K) / np.sqrt(K)).astype(np.complex64)
    n = 9000 + 3 * K + D
    iq = gen.random_u8(2 * n, 1000 + K + D)
    x = O.unpack_u8iq(iq)
    truth = O.fir_f64(taps, x)[D - 1::D]
    for flags in (0, NO_TCGEN05):
        f = sdr.Fir(taps, "u8iq", decimation=D, flags=flags)
        a = f.process(iq[:2 * 777])
        b = f.process(iq[2 * 777:])
        got = np.concatenate([a, b])
        if K <= 160 and flags == 0:
            assert f.last_path == 4
        elif K > 511 or flags:
            assert f.last_path == 3
        assert len(got) == n // D and rel_err(got, truth) < TOL


@pytest.mark.parametrize("D", [2, 3, 4, 5, 6, 7, 9, 10, 12, 20, 50, 8, 16, 96])
@pytest.mark.parametrize("K,tc", [(64, True), (255, False)])
def test_tcgen05_decimating_path(sdr, D, K, tc):
    """Decimate fused behind the filter on the tensor cores: rows stay 32 samples apart, only the offsets a kept
    output can take (multiples of gcd(D, 32)) get columns.  Exact integer sums: any blocking gives the same bits."""
    import math
    rng = np.random.default_rng(K * 3 + D)
    taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    if tc:
        taps = (taps + 1j * rng.standard_normal(K) / np.sqrt(K)).astype(np.complex64)
    n = 5 * 8192 + 37
    iq = gen.random_u8(2 * n, 500 + D)
    truth = O.fir_f64(taps, O.unpack_u8iq(iq))[D - 1::D]
    f = sdr.Fir(taps, "u8iq", decimation=D)
    whole = f.process(iq)
    poly = D in (5, 6, 7, 8, 9, 10, 12)  # phase-plane kernel: every computed output is kept
    if poly or (math.gcd(D, 32) <= 4 and not (D % 2 == 1 and K > 160)):
        assert f.last_path == 4
    elif math.gcd(D, 32) <= 4:
        assert f.last_path in (3, 4)  # odd D needs all 32 candidates: 255 taps' tables leave no room for the stages
    else:
        assert f.last_path in (1, 3)  # D = 16, 96: too few candidates per row for the tcgen05 path
    assert len(whole) == n // D and rel_err(whole, truth) < (1e-6 if f.last_path == 4 else TOL)
    if f.last_path != 4:
        return
    f.reset()
    cuts = [0, 1, D, D + 1, 8192, 8192 + 3, 3 * 8192 - 1, n]
    parts = np.concatenate([f.process(iq[2 * a:2 * b]) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(parts.view(np.uint32), whole.view(np.uint32))


@pytest.mark.parametrize("P", [8, 16, 32])
@pytest.mark.parametrize("K,tc", [(1, False), (17, True), (64, False), (64, True), (255, False), (255, True), (300, False)])
def test_tcgen05_path_every_row_width(sdr, monkeypatch, P, K, tc):
    """The three window-row widths of the tcgen05 kernel (no swizzle / 32-byte / 64-byte swizzled stages) give the
    same bits: the sums are exact integers, so only the (identical) epilogue rounds."""
    rng = np.random.default_rng(K * 7 + P)
    taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    if tc:
        taps = (taps + 1j * rng.standard_normal(K) / np.sqrt(K)).astype(np.complex64)
    n = 3 * 4096 + 5
    iq = gen.random_u8(2 * n, 77 + K)
    truth = O.fir_f64(taps, O.unpack_u8iq(iq))
    monkeypatch.setenv("SDR_UMMA_P", str(P))
    f = sdr.Fir(taps, "u8iq")
    got = f.process(iq)
    monkeypatch.setenv("SDR_UMMA_P", "8")
    ref = sdr.Fir(taps, "u8iq").process(iq)
    if f.last_path != 4:
        assert P == 32 and K > 255  # tables of the widest rows do not fit: falls back
        return
    assert rel_err(got, truth) < 1e-6
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # ragged blocks through the carried history: still the same bits
    f.reset()
    cuts = [0, 5, 4096, 4101, 8191, 8192 + 3, n]
    parts = np.concatenate([f.process(iq[2 * a:2 * b]) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(parts.view(np.uint32), got.view(np.uint32))


def test_planar_and_interleaved_tcgen05_kernels_agree_bit_for_bit(sdr):
    """Real taps can run on de-interleaved byte planes (SDR_FIR_PLANAR: half the tensor work) instead of the interleaved
    bytes.  Both accumulate exact integers and share the epilogue, so the bits must be identical -- D = 1 and
    decimating, ragged blocks, multi-channel, 255 taps included."""
    PLANAR = 8
    for K in (17, 64, 150, 255):
        taps = (np.random.default_rng(K).standard_normal(K) / np.sqrt(K)).astype(np.float32)
        for D in (1, 2, 3, 4, 10, 12):
            if K == 255 and D == 3:
                continue  # the interleaved kernel has no room for 32 candidates x 255 taps
            n = 3 * 8192 + 11
            iq = gen.random_u8(2 * n, K + D)
            cuts = [0, 7, 8192, 2 * 8192 + 5, n]
            got = []
            for flags in (0, PLANAR):
                f = sdr.Fir(taps, "u8iq", decimation=D, flags=flags)
                got.append(np.concatenate([f.process(iq[2 * a:2 * b]) for a, b in zip(cuts[:-1], cuts[1:])]))
                assert f.last_path == 4
            assert len(got[0]) == n // D
            assert np.array_equal(got[0].view(np.uint32), got[1].view(np.uint32)), (K, D)
        raw = gen.random_u8(2 * 3 * 9000, K).reshape(3, -1)
        a = sdr.Fir(taps, "u8iq", n_channels=3).process(raw)
        b = sdr.Fir(taps, "u8iq", n_channels=3, flags=PLANAR).process(raw)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_tensor_path_multichannel_and_impulse(sdr):
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    raw = gen.random_u8(2 * 5 * 4096, 9).reshape(5, -1)
    for flags, path in ((0, 4), (NO_TCGEN05, 3)):
        f = sdr.Fir(taps, "u8iq", n_channels=5, flags=flags)
        got = f.process(raw)
        assert f.last_path == path
        for c in range(5):
            assert rel_err(got[c], O.fir_f64(taps, O.unpack_u8iq(raw[c]))) < TOL
    # impulse: byte 255 at sample 0 over a 128-background is 127/128 * taps, up to f32 rounding of the split sum
    iq = np.full(2 * 300, 128, np.uint8)
    iq[0] = 255
    y = sdr.Fir(taps, "u8iq").process(iq)
    assert np.abs(y[:64].real - taps * np.float32(127.0 / 128.0)).max() < 1e-7 and np.all(y.imag == 0) and np.all(y[64:] == 0)


def test_fir_clone_reset_and_multichannel(sdr):
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    x = gen.complex_noise(6 * 5000, 8).reshape(6, 5000)
    want = np.stack([O.Fir(taps).apply(r) for r in x])
    f = sdr.Fir(taps, "c64", n_channels=6, strict=True)
    a = f.process(x[:, :1234])
    g = f.clone()
    b = f.process(x[:, 1234:])
    b2 = g.process(x[:, 1234:])
    assert np.array_equal(np.concatenate([a, b], 1), want) and np.array_equal(b2, b)
    f.reset()
    assert np.array_equal(f.process(x[:, :100]), want[:, :100])
    fast = sdr.Fir(taps, "c64", n_channels=6).process(x)
    truth = np.stack([O.fir_f64(taps, r) for r in x])
    assert rel_err(fast, truth) < TOL


def test_fir_errors(sdr):
    import ctypes as C
    taps = gen.lowpass_taps(8, 200e3, 2.048e6)
    f = sdr.Fir(taps, "c64")
    x = np.zeros(100, np.complex64)
    out = np.zeros(10, np.complex64)
    used, got = C.c_size_t(0), C.c_size_t(0)
    rc = sdr.lib().sdr_fir_process(f.h, x.ctypes.data, 100, 100, out.ctypes.data, 10, 10, C.byref(used), C.byref(got))
    assert rc == 104 and used.value == 0 and got.value == 0
    with pytest.raises(sdr.SdrError):
        sdr.Fir(taps.astype(np.complex64), "f32")


def test_golden_c1_c3(sdr):
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    taps64, taps255 = gen.lowpass_taps(64, 200e3, 2.048e6), gen.lowpass_taps(255, 100e3, 2.4e6)
    got = sdr.Fir(taps64, "u8iq", strict=True).process(g["c1_iq"])
    assert np.array_equal(got.view(np.uint32), g["c1_fir64"].view(np.uint32))
    got = sdr.Fir(taps255, "u8iq", decimation=10, strict=True).process(g["c3_iq"])
    assert np.array_equal(got.view(np.uint32), g["c3_fir255_dec10"].view(np.uint32))


def test_fir_full_size_linearity_and_device_path(sdr):
    """C1 at 2^26 samples, device-resident: FIR(a + b) == FIR(a) + FIR(b) on tones whose unpacked values add
    exactly, and shard-with-halo outputs equal the single-call outputs bit for bit."""
    import torch
    n = 1 << 26
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    dev = torch.device("cuda:0")
    gen_t = torch.Generator(device=dev).manual_seed(1234)
    iq = torch.randint(0, 256, (2 * n,), dtype=torch.uint8, device=dev, generator=gen_t)
    out = torch.empty(n, dtype=torch.complex64, device=dev)
    f = sdr.Fir(taps, "u8iq")
    assert f.process_dev(iq, n, out, n) == n
    torch.cuda.synchronize()
    # spot-check 3 windows against the oracle
    for s in (0, 12345678, n - 5000):
        seg = iq[2 * max(0, s - 63):2 * (s + 5000)].cpu().numpy()
        want = O.fir_f64(taps, O.unpack_u8iq(seg))[-5000:]
        assert rel_err(out[s:s + 5000].cpu().numpy(), want) < TOL
    # two shards with a K-1 halo reproduce the whole
    half = n // 2
    f2 = sdr.Fir(taps, "u8iq")
    o2 = torch.empty(half + 64, dtype=torch.complex64, device=dev)
    got = f2.process_dev(iq[2 * (half - 64):], half + 64, o2, half + 64)
    torch.cuda.synchronize()
    assert got == half + 64 and torch.equal(o2[64:].view(torch.float32), out[half:].view(torch.float32))


def test_c3_full_size_decimated_equals_every_tenth_output(sdr):
    """BASELINE config 3's FIR stage at 2^26 samples, device resident: the fused Decimate (polyphase tcgen05 kernel) must
    equal every 10th output of the undecimated filter.  Both kernels accumulate exact integers and share the epilogue
    arithmetic, so the comparison is bit for bit; spot windows are checked against the f64 truth."""
    import torch
    n, D = 1 << 26, 10
    taps = gen.lowpass_taps(255, 100e3, 2.4e6)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(4321)
    iq = torch.randint(0, 256, (2 * n,), dtype=torch.uint8, device=dev, generator=g)
    full = torch.empty(n, dtype=torch.complex64, device=dev)
    dec = torch.empty(n // D + 1, dtype=torch.complex64, device=dev)
    f1, fd = sdr.Fir(taps, "u8iq"), sdr.Fir(taps, "u8iq", decimation=D)
    assert f1.process_dev(iq, n, full, n) == n
    got = fd.process_dev(iq, n, dec, n // D + 1)
    torch.cuda.synchronize()
    assert f1.last_path == 4 and fd.last_path == 4 and got == n // D
    assert torch.equal(dec[:got].view(torch.float32), full[D - 1::D][:got].contiguous().view(torch.float32))
    for s in (0, 3333333, n // D - 4000):
        lo = max(0, s * D + D - 1 - 254)
        seg = iq[2 * lo:2 * ((s + 3000) * D + D)].cpu().numpy()
        want = O.fir_f64(taps, O.unpack_u8iq(seg))[(s * D + D - 1 - lo)::D][:3000]
        assert rel_err(dec[s:s + 3000].cpu().numpy(), want) < 1e-6


def test_tensor_paths_random_configurations(sdr):
    """Randomised sweep over (K, D, tap kind, channels, length, blocking): every fast path against the f64 truth, and --
    where the tcgen05 path ran -- bit-identical results for a second, differently cut pass over the same stream."""
    rng = np.random.default_rng(20261018)
    seen = {1: 0, 3: 0, 4: 0}
    for trial in range(48):
        K = int(rng.choice([1, 2, 7, 33, 64, 100, 255, 300, 511, 600]))
        D = int(rng.choice([1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 25]))
        tc = bool(rng.integers(0, 2))
        n_ch = int(rng.choice([1, 1, 2, 3]))
        n = int(rng.choice([1, 5, 100, 1023, 4096, 10240, 30011]))
        taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
        if tc:
            taps = (taps + 1j * rng.standard_normal(K) / np.sqrt(K)).astype(np.complex64)
        raw = gen.random_u8(2 * n_ch * n, 7000 + trial).reshape(n_ch, 2 * n)
        f = sdr.Fir(taps, "u8iq", decimation=D, n_channels=n_ch)
        cut = int(rng.integers(0, n + 1))
        parts = [f.process(np.ascontiguousarray(raw[:, :2 * cut])), f.process(np.ascontiguousarray(raw[:, 2 * cut:]))]
        got = np.concatenate([p_.reshape(n_ch, -1) for p_ in parts], axis=1)
        assert got.shape == (n_ch, n // D), (trial, K, D, n, got.shape)
        seen[f.last_path] = seen.get(f.last_path, 0) + 1
        for c in range(n_ch):
            truth = O.fir_f64(taps, O.unpack_u8iq(raw[c]))[D - 1::D]
            if len(truth):
                assert rel_err(got[c], truth) < (1e-6 if f.last_path == 4 else TOL), (trial, K, D, tc, n_ch, n, f.last_path)
        if f.last_path == 4 and n > 0:
            f.reset()
            cut2 = int(rng.integers(0, n + 1))
            again = np.concatenate([p_.reshape(n_ch, -1) for p_ in (f.process(np.ascontiguousarray(raw[:, :2 * cut2])),
                                                                    f.process(np.ascontiguousarray(raw[:, 2 * cut2:])))], axis=1)
            assert np.array_equal(again.view(np.uint32), got.view(np.uint32)), (trial, K, D, n)
    assert seen[4] >= 20  # the tcgen05 kernels carry most configurations
