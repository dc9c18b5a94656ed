"""CPU: host-side adaptor logic that needs no device (counts, unfused Decimate, sharding maths)."""
import numpy as np

import gen
import oracle_lib as O


def test_unfused_decimate_take_skip_block(sdr):
    S = sdr.signal
    ramp = np.arange(1000, dtype=np.float32)
    s = S.from_array(2.4e6, ramp).decimate(240e3)
    assert s.wait == 10 and not s.fused and s.rate() == np.float32(2.4e6)  # rate() quirk, adapters/mod.rs:38-40
    assert np.array_equal(s.collect(block=37), O.decimate(ramp, 10)[0])
    t = S.from_array(1000.0, ramp).skip(0.1).take(0.25)
    assert np.array_equal(t.collect(block=64), ramp[100:350])
    b = S.from_array(1000.0, ramp).block(0.0305, dedup=True)
    assert b.block_size == O.block_size(0.0305, 1000.0) == 31
    assert len(b.next_block(1 << 20)) == 31
    m = S.from_array(1000.0, ramp).map(lambda v: v * 2)
    assert np.array_equal(m.collect(), ramp * 2)


def test_shard_ranges_cover_and_align(sdr):
    sh = sdr.shard
    for n, w, D, K in [(2 ** 20, 8, 1, 64), (24_000_007, 8, 10, 255), (1000, 3, 7, 5), (10, 4, 1, 3)]:
        prev_hi = 0
        outs = 0
        for r in range(w):
            lo, hi, hlo = sh.sample_range(n, w, r, K - 1, D)
            assert lo == prev_hi and lo % D == 0 and hlo == max(0, lo - (K - 1))
            olo, ohi = sh.output_range(n, w, r, D)
            assert olo == outs
            outs = ohi
            prev_hi = hi
        assert prev_hi == n and outs == n // D
    assert [sh.unit_range(10, 4, r) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]


def test_sharded_fir_equals_single_stream_on_oracle(sdr):
    """the partitioning itself (cut on multiples of D, K-1 halo) preserves every output sample"""
    iq = gen.fm_u8(30011, 2.4e6, 75e3, 1e3, 0.05, 5)
    taps = gen.lowpass_taps(255, 100e3, 2.4e6)
    x = O.unpack_u8iq(iq)
    whole = O.Fir(taps).apply(x)[9::10]
    parts = []
    for r in range(4):
        lo, hi, hlo = sdr.shard.sample_range(len(x), 4, r, 254, 10)
        f = O.Fir(taps)
        f.apply(x[hlo:lo])  # prime with the halo
        parts.append(f.apply(x[lo:hi])[9::10])
    assert np.array_equal(np.concatenate(parts), whole)


def test_rtltcp_wire_protocol_against_a_loopback_server(sdr):
    """src/rtltcp.rs:60-134 on the wire: 12-byte greeting, then SetSampleRate, SetFrequency, gain mode (+ gain),
    SetRtlAgc as 1 + 4 big-endian bytes each, then raw I/Q bytes until the peer closes."""
    import socket
    import threading
    R = sdr.rtltcp
    assert R.command_bytes(R.CMD_SET_FREQUENCY, 100000000) == bytes([0x01, 0x05, 0xF5, 0xE1, 0x00])
    assert R.command_bytes(R.CMD_SET_RTL_AGC, 1) == bytes([0x08, 0, 0, 0, 1])
    assert R.gain_tenths_db(49.6) == 496 and R.gain_tenths_db(-3.0) == 0 and R.gain_tenths_db(0.25) == 3
    for bad in (225000, 300001, 900000, 3200001):
        try:
            R.check_sample_rate(bad)
            raise AssertionError("accepted %d" % bad)
        except ValueError:
            pass
    for ok in (225001, 300000, 900001, 1800000, 2400000, 3200000):
        R.check_sample_rate(ok)

    payload = gen.random_u8(2 * 5000 + 1, 77)  # odd length: the dangling byte is dropped (read error on Q, :138-139)
    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]
    seen = {}

    def serve():
        c, _ = srv.accept()
        c.sendall(b"RTL0" + bytes([0, 0, 0, 5, 0, 0, 0, 29]))
        buf = b""
        want = 5 * 5  # manual gain: rate, freq, gain mode, gain, agc
        while len(buf) < want:
            b = c.recv(want - len(buf))
            if not b:
                break
            buf += b
        seen["cmds"] = buf
        c.sendall(payload.tobytes())
        c.close()

    t = threading.Thread(target=serve, daemon=True)
    t.start()
    sig = R.RtlTcp().address("127.0.0.1:%d" % port).rate(2400000).frequency(99500000).gain(20.7).rtlagc(True).listen(timeout=10)
    assert sig.conn.id == b"RTL0" + bytes([0, 0, 0, 5, 0, 0, 0, 29])
    assert sig.rate() == np.float32(2400000.0)
    got = [sig.next_raw(1234) for _ in range(6)]
    t.join(10)
    srv.close()
    raw = np.concatenate(got)
    assert np.array_equal(raw, payload[:10000]) and len(sig.next_raw(16)) == 0
    import struct
    assert seen["cmds"] == b"".join(struct.pack(">BI", c, a) for c, a in
                                    [(2, 2400000), (1, 99500000), (3, 1), (4, 207), (8, 1)])
    sig.conn.close()


def _drain(sig, n_max, pulls):
    out = []
    for k in pulls:
        b = sig.next_block(k)
        out.extend(b.tolist())
        if len(out) >= n_max:
            break
    return out


def test_block_matches_reference_semantics_sample_for_sample(sdr):
    """signal::Block against a per-sample restatement of adapters/block.rs (tests/pyref.py BlockRef): block size
    ceil(size * rate), the short last block, end of stream -- and the reference's quirk that the first sample of every
    block is delivered twice (block.rs:197-199 returns current[0] without advancing i)."""
    import pyref
    S = sdr.signal
    x = np.arange(1, 1001, dtype=np.float32)
    for size in (0.0305, 0.001, 0.25, 2.0):
        ref = pyref.BlockRef(iter(x.tolist()), 1000.0, size)
        want = []
        while True:
            v = ref.next()
            if v is None:
                break
            want.append(v)
        for pulls in ((1 << 20,), (1,), (7, 1, 64)):
            b = S.from_array(1000.0, x).block(size)
            got = []
            for i in range(100000):
                blk = b.next_block(pulls[i % len(pulls)])
                if len(blk) == 0:
                    break
                got.extend(blk.tolist())
            assert got == want, (size, pulls)
        bs = O.block_size(size, 1000.0)
        assert len(want) == 1000 + -(-1000 // bs)  # one duplicate per block
        clean = S.from_array(1000.0, x).block(size, dedup=True).collect(block=50)
        assert np.array_equal(clean, x)


def test_block_clone_is_a_tee(sdr):
    """Clone for Block (block.rs:129-140): clones share the upstream; every reader sees every block pushed after it
    was created, plus the blocks the deque still held when it was cloned (TeeDeque::clone, :92-103)."""
    import pyref
    S = sdr.signal
    x = np.arange(1, 501, dtype=np.float32)
    # a schedule of (reader, samples to pull); reader 1 is cloned from reader 0 after step 2, reader 2 after step 5
    schedule = [(0, 10), (0, 25), (0, 3), (1, 40), (0, 12), (1, 5), (2, 33), (0, 50), (2, 50), (1, 100), (0, 200),
                (2, 300), (1, 300), (0, 300), (1, 300), (2, 300)]
    clone_at = {3: (0, 1), 6: (1, 2)}
    ref = {0: pyref.BlockRef(iter(x.tolist()), 100.0, 0.2)}
    got_r = {0: S.from_array(100.0, x).block(0.2)}
    want, got = {0: [], 1: [], 2: []}, {0: [], 1: [], 2: []}
    for step, (rid, k) in enumerate(schedule):
        if step in clone_at:
            src, new = clone_at[step]
            ref[new] = ref[src].clone()
            got_r[new] = got_r[src].clone()
        for _ in range(k):
            v = ref[rid].next()
            if v is None:
                break
            want[rid].append(v)
        left = k
        while left > 0:
            blk = got_r[rid].next_block(left)
            if len(blk) == 0:
                break
            got[rid].extend(blk.tolist())
            left -= len(blk)
    for rid in (0, 1, 2):
        assert got[rid] == want[rid] and len(want[rid]) > 100, rid
    # the readers together drained the single shared upstream exactly once
    seen0 = sorted(set(want[0]))
    assert seen0 == x.tolist()


def test_cpp_block_tee_matches_reference_semantics(sdr):
    """the C++ host mirror's signal::Block (host/sdr.hpp) on the same schedule as the per-sample restatement"""
    import os
    import subprocess
    import pyref
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "block_tee_test")
    lib = os.path.dirname(sdr.LIB_PATH)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, exe + ".cpp", "-L" + lib, "-lsdr_b200",
                           "-Wl,-rpath," + lib, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])
    sched = [(0, 10), (0, 25), (0, 3), ("c", 0), (1, 40), (0, 12), (1, 5), ("c", 1), (2, 33), (0, 50), (2, 50), (1, 100),
             (0, 200), (2, 300), (1, 300), (0, 300), (1, 300), (2, 300)]
    args = [str(v) for pair in sched for v in pair]
    out = subprocess.check_output([exe, "500", "100", "0.2", "0"] + args, text=True)
    got = {int(l.split(":")[0]): [float(v) for v in l.split(":")[1].split()] for l in out.splitlines()}
    ref = {0: pyref.BlockRef(iter([float(i) for i in range(1, 501)]), 100.0, 0.2)}
    want = {}
    for a, b in sched:
        if a == "c":
            ref[len(ref)] = ref[b].clone()
            continue
        for _ in range(b):
            v = ref[a].next()
            if v is None:
                break
            want.setdefault(a, []).append(v)
    assert got == want
    out = subprocess.check_output([exe, "100", "1000", "0.0305", "1", "0", "1000"], text=True)
    assert [float(v) for v in out.split(":")[1].split()] == [float(i) for i in range(1, 101)]
