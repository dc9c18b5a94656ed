"""world_size 2 over gloo: the N>1 path of bench.py -- every rank takes its own shard, there is no data-path
collective, timing is reduced with MAX, results concatenate to the single-rank answer.  On the CPU container the
shards are computed by the oracle (this checks the partitioning arithmetic of sdr_b200.shard); on a GPU box the
gpu-marked variant computes them with libsdr_b200 (each rank on cuda:rank % n_devices) and checks the product."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q, product=False):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "unnamed-rust-sdr_b200"))
    import gen
    import oracle_lib as O
    from sdr_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_batches, N = 64, 256
    iq = gen.random_u8(2 * n_batches * N, 99)           # every rank can regenerate any slice
    lo, hi = shard.unit_range(n_batches, world, rank)
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    x = O.unpack_u8iq(iq)
    slo, shi, hlo = shard.sample_range(len(x), world, rank, 63, 1)
    if product:
        import sdr_b200 as sdr
        dev = rank % sdr.device_count()
        plan = sdr.FftPlan(N, "u8iq", shift=True, norm=True, device=dev)
        mine = plan.exec(iq[2 * N * lo:2 * N * hi])
        # halo split: the shard's handle is primed with the K-1 samples to its left, then filters its own range
        f = sdr.Fir(taps, "u8iq", strict=True, device=dev)
        f.process(iq[2 * hlo:2 * slo])
        fir_mine = f.process(iq[2 * slo:2 * shi])
        g = sdr.Fir(taps, "u8iq", device=dev)           # the default (tcgen05) path shards bit-identically too
        g.process(iq[2 * hlo:2 * slo])
        fir_fast = g.process(iq[2 * slo:2 * shi])
        # channel split (C4): each rank runs its own channels of a multi-channel FIR
        xc = gen.complex_noise(6 * 3000, 7).reshape(6, 3000)
        clo, chi = shard.unit_range(6, world, rank)
        chan = sdr.Fir(gen.lowpass_taps(255, 100e3, 1.8e6), "c64", n_channels=chi - clo, device=dev).process(
            np.ascontiguousarray(xc[clo:chi]))
        mine = (mine, fir_fast, chan)
    else:
        mine = O.fft_batch_u8(iq[2 * N * lo:2 * N * hi], N, 1)
        # FIR shard with halo
        f = O.Fir(taps)
        f.apply(x[hlo:slo])
        fir_mine = f.apply(x[slo:shi])
    # the only collectives are off the data path: a barrier and a MAX over the elapsed time
    dist.barrier()
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    units = torch.tensor([float(hi - lo)])
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    q.put((rank, mine, fir_mine, float(t), float(units)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sys.path.insert(0, HERE)
    import gen
    import oracle_lib as O
    iq = gen.random_u8(2 * 64 * 256, 99)
    whole = O.fft_batch_u8(iq, 256, 1)
    assert np.array_equal(np.concatenate([res[0][1], res[1][1]]), whole)
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    assert np.array_equal(np.concatenate([res[0][2], res[1][2]]), O.Fir(taps).apply(O.unpack_u8iq(iq)))
    assert res[0][3] == res[1][3] == 2.0 and res[0][4] == 64.0


import pytest  # noqa: E402


@pytest.mark.gpu
def test_two_rank_sharding_of_the_product_matches_single_rank():
    """the same two-rank run with every shard computed by libsdr_b200: FFT batch split, FIR halo split (reference-order
    kernel bit-identical to the oracle, default tcgen05 kernel bit-identical to its own single-stream run), channel
    split of a multi-channel FIR"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "unnamed-rust-sdr_b200"))
    import gen
    import oracle_lib as O
    import sdr_b200 as sdr
    iq = gen.random_u8(2 * 64 * 256, 99)
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    fft_sh = np.concatenate([res[0][1][0], res[1][1][0]])
    whole = sdr.FftPlan(256, "u8iq", shift=True, norm=True).exec(iq)
    assert np.array_equal(fft_sh.view(np.uint32), whole.view(np.uint32))
    ref = O.fft_batch_u8(iq, 256, 1).reshape(64, 256)
    assert np.abs(fft_sh - ref).max() / np.abs(ref).max() < 8e-5
    assert np.array_equal(np.concatenate([res[0][2], res[1][2]]).view(np.uint32),
                          O.Fir(taps).apply(O.unpack_u8iq(iq)).view(np.uint32))
    fast = np.concatenate([res[0][1][1], res[1][1][1]])
    assert np.array_equal(fast.view(np.uint32), sdr.Fir(taps, "u8iq").process(iq).view(np.uint32))
    xc = gen.complex_noise(6 * 3000, 7).reshape(6, 3000)
    chan = np.concatenate([res[0][1][2], res[1][1][2]])
    assert np.array_equal(chan.view(np.uint32),
                          sdr.Fir(gen.lowpass_taps(255, 100e3, 1.8e6), "c64", n_channels=6).process(xc).view(np.uint32))
    assert res[0][3] == res[1][3] == 2.0 and res[0][4] == 64.0
