"""CPU, world_size 2 over gloo: the N>1 path of bench.py -- every rank takes its own shard, there is
no data-path collective, timing is reduced with MAX, results concatenate to the single-rank answer."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
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
    mine = O.fft_batch_u8(iq[2 * N * lo:2 * N * hi], N, 1)
    # FIR shard with halo
    taps = gen.lowpass_taps(64, 200e3, 2.048e6)
    x = O.unpack_u8iq(iq)
    slo, shi, hlo = shard.sample_range(len(x), world, rank, 63, 1)
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
